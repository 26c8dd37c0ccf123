// cc3d_dispatch.cuh — host-side launch of the labelling stage (kernels A, B1, B2, P) for one element
// type. Each inst_<T>.cu instantiates run_label_stage<T> so that the template matrix
// (6 types x 4 predicates x 5 connectivities) compiles in parallel.
#pragma once
#include "cc3d_label.cuh"

#define CC_TX 64

struct LabelArgs {
  const void* in;    // device pointer, element kind T
  u32* L;            // [voxels]
  u32* LR;           // [rows * W]
  u32* XS;           // [rows * (ntx-1)]
  Counters* ctr;     // device
  Geom g;
  int mode;          // MODE_*
  int connectivity;  // 4, 8, 6, 18, 26
  int periodic;
  unsigned char delta[8];  // one element of T
  cudaStream_t stream;
  void (*mark)(const char*, cudaStream_t);  // optional timing hook
  int* launches;     // incremented once per kernel launch
};

template <typename T> int run_label_stage(const LabelArgs& a);

#ifdef CC3D_INSTANTIATE
template <typename T, int MODE, int CONN>
static int launch_label(const LabelArgs& a) {
  Edge<T, MODE> E;
  memcpy(&E.delta, a.delta, sizeof(T));
  const Geom& g = a.g;
  const T* in = static_cast<const T*>(a.in);
  const size_t smem = tile_smem_bytes<T>();
  const i64 ntiles = g.ntx * g.nty * g.ntz;
  // tile shape: 64x8x8 for volumes, 64x64x1 for images (2D connectivities only exist for sz == 1)
  if (g.TZ == 1) {
    auto kA = k_tile_label<T, MODE, CONN, 6, 0>;
    static bool attr_set = false;
    if (!attr_set) { cudaFuncSetAttribute(kA, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr_set = true; }
    kA<<<(unsigned)ntiles, CC_TILE_THREADS, smem, a.stream>>>(in, a.L, a.LR, a.XS, g, E, a.ctr);
  } else {
    if constexpr (CONN == 4 || CONN == 8) { return -1; }
    else {
      auto kA = k_tile_label<T, MODE, CONN, 3, 3>;
      static bool attr_set = false;
      if (!attr_set) { cudaFuncSetAttribute(kA, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr_set = true; }
      kA<<<(unsigned)ntiles, CC_TILE_THREADS, smem, a.stream>>>(in, a.L, a.LR, a.XS, g, E, a.ctr);
    }
  }
  ++*a.launches;
  if (a.mark) a.mark("A_tile_label", a.stream);
  const i64 rows = g.sy * g.sz;
  if (g.nty > 1 || g.ntz > 1) {
    const unsigned nchunks = (unsigned)((g.sx + 255) / 256);
    k_seam_rows<T, MODE, CONN><<<(unsigned)rows * nchunks, 256, 0, a.stream>>>(in, a.L, g, E, nchunks);
    ++*a.launches;
    if (a.mark) a.mark("B1_seam_rows", a.stream);
  }
  if (g.ntx > 1) {
    const i64 n = rows * (g.ntx - 1);
    k_seam_x<CC_TX><<<(unsigned)((n + 255) / 256), 256, 0, a.stream>>>(a.XS, a.L, g);
    ++*a.launches;
    if (a.mark) a.mark("B2_seam_x", a.stream);
  }
  if constexpr ((MODE == MODE_EQ || MODE == MODE_NONZERO) && (CONN == 4 || CONN == 8 || CONN == 6)) {
    if (a.periodic) {
      const i64 n0 = 2 * g.sy * g.sz, n1 = g.sx * g.sz, n2 = g.sx * g.sy;
      k_periodic<T, MODE, CONN><<<(unsigned)((n0 + 255) / 256), 256, 0, a.stream>>>(in, a.L, g, E, 0);
      k_periodic<T, MODE, CONN><<<(unsigned)((n1 + 255) / 256), 256, 0, a.stream>>>(in, a.L, g, E, 1);
      if (CONN == 6) k_periodic<T, MODE, CONN><<<(unsigned)((n2 + 255) / 256), 256, 0, a.stream>>>(in, a.L, g, E, 2);
      *a.launches += (CONN == 6) ? 3 : 2;
      if (a.mark) a.mark("P_periodic", a.stream);
    }
  }
  return 0;
}

template <typename T, int MODE>
static int launch_label_conn(const LabelArgs& a) {
  switch (a.connectivity) {
    case 4: return launch_label<T, MODE, 4>(a);
    case 8: return launch_label<T, MODE, 8>(a);
    case 6: return launch_label<T, MODE, 6>(a);
    case 18: return launch_label<T, MODE, 18>(a);
    case 26: return launch_label<T, MODE, 26>(a);
  }
  return -1;
}

template <typename T> int run_label_stage(const LabelArgs& a) {
  switch (a.mode) {
    case MODE_EQ: return launch_label_conn<T, MODE_EQ>(a);
    case MODE_NONZERO: return launch_label_conn<T, MODE_NONZERO>(a);
    case MODE_DELTA: return launch_label_conn<T, MODE_DELTA>(a);
    case MODE_MASK:
      if constexpr (sizeof(T) <= 2 && !is_float_t<T>::value) return launch_label_conn<T, MODE_MASK>(a);
      else return -1;
  }
  return -1;
}
#endif
