// Instantiates the labelling-stage kernels (A, B1, B2, P) for element type uint16_t.
#define CC3D_INSTANTIATE
#include <cstring>
#include "cc3d_dispatch.cuh"
template int run_label_stage<uint16_t>(const LabelArgs&);
