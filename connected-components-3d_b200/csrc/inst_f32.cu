// Instantiates the type-dependent kernels (A: face bitmaps, B: unions, P: periodic wrap) for element type float.
#define CC3D_INSTANTIATE
#include <cstring>
#include "cc3d_dispatch.cuh"
template int run_faces_stage<float>(const LabelArgs&);
template int run_union_stage<float>(const LabelArgs&);
template int run_periodic_stage<float>(const LabelArgs&);
template int run_union_global_stage<float>(const LabelArgs&);
