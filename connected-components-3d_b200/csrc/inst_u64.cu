// Instantiates the type-dependent kernels (A: face bitmaps, B: unions, P: periodic wrap) for element type uint64_t.
#define CC3D_INSTANTIATE
#include <cstring>
#include "cc3d_dispatch.cuh"
template int run_faces_stage<uint64_t>(const LabelArgs&);
template int run_union_stage<uint64_t>(const LabelArgs&);
template int run_periodic_stage<uint64_t>(const LabelArgs&);
template int run_union_global_stage<uint64_t>(const LabelArgs&);
