// cc3d_dispatch.cuh — host-side launch of the type-dependent kernels (A: face bitmaps, B: unions,
// P: periodic wrap) for one element type. Each inst_<T>.cu instantiates run_*_stage<T> so that the
// template matrix (6 types x predicates x connectivities) compiles in parallel.
#pragma once
#include <cstdlib>
#include <algorithm>
#include "cc3d_faces.cuh"
#include "cc3d_union.cuh"
#include "cc3d_union_w.cuh"

struct LabelArgs {
  const void* in;    // device pointer, element kind T
  u32* M;            // [PL_COUNT][nwords] bitmaps
  u32* L;            // [max runs] forest over runs
  Counters* ctr;     // device
  Geom g;
  int mode;          // MODE_*
  int connectivity;  // 4, 8, 6, 18, 26
  unsigned char delta[8];  // one element of T
  cudaStream_t stream;
  EdgeQueue GQ;      // edges that leave their union tile
  void (*mark)(const char*, cudaStream_t);  // optional timing hook (per-kernel CUDA events)
  int* launches;     // incremented once per kernel launch
  u32* bigflags;          // one word per union tile (zeroed): set by B1 for tiles with > 16 runs per word
  u32* nbig;              // device counter of such tiles
  bool defer_big;         // the process has met such volumes: flagged tiles are relabelled by the RL = 5 launch
  bool inline_fallback;   // launch the overflow fallback kernel (k_union_global, a no-op unless the edge queue overflowed)
                          // as part of the pipeline; false: the host checks the overflow flag at its next
                          // synchronisation and redoes the unions (run_union_global_stage)
};
#define CC_QUEUE_BLOCKS (148 * 8)

// cudaFuncSetAttribute is per device: remember per (kernel instantiation, device) whether it was done
struct PerDeviceOnce {
  bool done[64] = {false};
  bool first() {
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 64) return true;
    if (done[d]) return false;
    done[d] = true;
    return true;
  }
};

template <typename T> int run_faces_stage(const LabelArgs& a);     // kernel A
template <typename T> int run_union_stage(const LabelArgs& a);     // kernel B
template <typename T> int run_periodic_stage(const LabelArgs& a);  // kernel P (after kernel B)
template <typename T> int run_union_global_stage(const LabelArgs& a);   // every edge on the global forest (overflow redo)

#ifdef CC3D_INSTANTIATE
template <typename T, int MODE, int NW>
static void launch_faces_staged(const LabelArgs& a, const Edge<T, MODE>& E, bool two_d, unsigned nych) {
  const Geom& g = a.g;
  const unsigned nwg = (unsigned)(g.W / NW);
  const i64 ntasks = (i64)nwg * nych * g.sz;
  const unsigned blocks = (unsigned)((ntasks + CC_FACE_WARPS - 1) / CC_FACE_WARPS);
  const T* in = static_cast<const T*>(a.in);
  constexpr size_t smem = faces_async_smem<T, NW>();
  static PerDeviceOnce once;
  if (once.first()) {
    cudaFuncSetAttribute(k_faces_async<T, MODE, false, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k_faces_async<T, MODE, true, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  }
  if (two_d) cc_launch(k_faces_async<T, MODE, false, NW>, dim3(blocks), dim3(CC_FACE_WARPS * 32), (size_t)(smem), a.stream, in, a.M, g, E, a.ctr, nych, nwg, (unsigned)ntasks);
  else cc_launch(k_faces_async<T, MODE, true, NW>, dim3(blocks), dim3(CC_FACE_WARPS * 32), (size_t)(smem), a.stream, in, a.M, g, E, a.ctr, nych, nwg, (unsigned)ntasks);
}


// ---- tensor map of the input volume for k_faces_tma (driver entry point fetched through the runtime: no -lcuda) ----
typedef CUresult (*cc_tmap_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                      const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                      CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static cc_tmap_encode_fn cc_tmap_encoder() {
  static cc_tmap_encode_fn fn = []() -> cc_tmap_encode_fn {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
      cudaGetLastError();
      return nullptr;
    }
    return (cc_tmap_encode_fn)p;
  }();
  return fn;
}

// true when the TMA kernel was launched
#ifndef CC_TMA_V1
#define CC_TMA_KERNEL k_faces_tma2
#define CC_TMA_TRAITS FaceTma2
#define CC_TMA_WARPS_ CC_TMA2_WARPS
#else
#define CC_TMA_KERNEL k_faces_tma
#define CC_TMA_TRAITS FaceTma
#define CC_TMA_WARPS_ CC_FACE_WARPS
#endif
template <typename T, int MODE>
static bool launch_faces_tma(const LabelArgs& a, const Edge<T, MODE>& E, bool two_d) {
  typedef CC_TMA_TRAITS<T> F;
  const Geom& g = a.g;
  const size_t es = sizeof(T);
  if (es > 4 || (reinterpret_cast<uintptr_t>(a.in) & 15) != 0 || ((size_t)g.sx * es) % 16 != 0) return false;
  if (getenv("CC3D_B200_NO_TMA")) return false;
  cc_tmap_encode_fn enc = cc_tmap_encoder();
  if (!enc) return false;
  CUtensorMap map;
  const cuuint64_t dims[3] = {(cuuint64_t)g.sx, (cuuint64_t)g.sy, (cuuint64_t)g.sz};
  const cuuint64_t strides[2] = {(cuuint64_t)g.sx * es, (cuuint64_t)g.sx * (cuuint64_t)g.sy * es};
  const cuuint32_t box[3] = {(cuuint32_t)F::BOXX, (cuuint32_t)(F::TR + 1), 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUtensorMapDataType dt = es == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : (es == 2 ? CU_TENSOR_MAP_DATA_TYPE_UINT16 : CU_TENSOR_MAP_DATA_TYPE_UINT32);
  if (strides[1] >= (1ull << 40)) return false;
  if (enc(&map, dt, 3, const_cast<void*>(a.in), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return false;
  const unsigned nwg = (unsigned)((g.W + F::NW - 1) / F::NW);
  const unsigned nyb = (unsigned)((g.sy + F::TR - 1) / F::TR);
  // z chunks: enough warps to fill the machine several times over, long enough to amortise the z - 1 halo plane
  unsigned zchunk = 16;
  while (zchunk > 2 && (i64)nwg * nyb * ((g.sz + zchunk - 1) / zchunk) < 148 * 24 * 4) zchunk >>= 1;
  { static const int zc_env = []() { const char* e = getenv("CC3D_B200_ZCHUNK"); return e ? atoi(e) : 0; }(); if (zc_env > 0) zchunk = (unsigned)zc_env; }
  const unsigned nzc = (unsigned)((g.sz + zchunk - 1) / zchunk);
  const i64 ntasks = (i64)nwg * nyb * nzc;
  if (ntasks >= (i64(1) << 31)) return false;
  const unsigned blocks = (unsigned)((ntasks + CC_TMA_WARPS_ - 1) / CC_TMA_WARPS_);
  constexpr size_t smem = F::smem();
  static PerDeviceOnce once;
  if (once.first()) {
    cudaFuncSetAttribute(CC_TMA_KERNEL<T, MODE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(CC_TMA_KERNEL<T, MODE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  }
  if (two_d) cc_launch(CC_TMA_KERNEL<T, MODE, false>, dim3(blocks), dim3(CC_TMA_WARPS_ * 32), smem, a.stream, map, a.M, g, E, a.ctr, nyb, nwg, zchunk, (unsigned)ntasks);
  else cc_launch(CC_TMA_KERNEL<T, MODE, true>, dim3(blocks), dim3(CC_TMA_WARPS_ * 32), smem, a.stream, map, a.M, g, E, a.ctr, nyb, nwg, zchunk, (unsigned)ntasks);
  return true;
}

template <typename T, int MODE>
static int launch_faces(const LabelArgs& a) {
  Edge<T, MODE> E;
  memcpy(&E.delta, a.delta, sizeof(T));
  E.zeq = (MODE == MODE_DELTA && a.connectivity == 26) ? 1 : 0;
  const Geom& g = a.g;
  const bool two_d = a.connectivity == 4 || a.connectivity == 8;
  if (two_d && g.sz != 1) return -1;
  const unsigned nych = (unsigned)((g.sy + CC_FACE_YCH - 1) / CC_FACE_YCH);
  const T* in = static_cast<const T*>(a.in);
  // binary images of 1-byte elements: the foreground bitmap comes from byte-parallel arithmetic (one thread per word),
  // the links are word logic on it (k_fg_bitmap_u8 / k_faces_from_fg, cc3d_faces.cuh; the bitmap borrows the forest
  // array, which nothing touches before B1). CC3D_B200_BIN_A=0: generic kernel
  if constexpr (MODE == MODE_NONZERO && sizeof(T) == 1) {
    static const bool bin_a = !(getenv("CC3D_B200_BIN_A") && atoi(getenv("CC3D_B200_BIN_A")) == 0);
    if (bin_a) {
      u32* Fb = a.L;
      const int vec_ok = ((reinterpret_cast<uintptr_t>(in) & 15) == 0 && (g.sx & 15) == 0) ? 1 : 0;
      const unsigned nb = (unsigned)((g.nwords + 255) / 256);
      cc_launch(k_fg_bitmap_u8, dim3(std::min(nb, 148u * 16u)), dim3(256), (size_t)0, a.stream, reinterpret_cast<const uint8_t*>(in), Fb, g, a.ctr, vec_ok);
      if (two_d) cc_launch(k_faces_from_fg<false>, dim3(nb), dim3(256), (size_t)0, a.stream, (const u32*)Fb, a.M, g);
      else cc_launch(k_faces_from_fg<true>, dim3(nb), dim3(256), (size_t)0, a.stream, (const u32*)Fb, a.M, g);
      *a.launches += 2;
      return 0;
    }
  }
  // TMA variant (tensor-map bulk copies): <= 4-byte elements, 16-byte aligned base and rows; any sx / sy / sz
  if constexpr (sizeof(T) <= 4) {
    if (launch_faces_tma<T, MODE>(a, E, two_d)) { ++*a.launches; return 0; }
  }
  // staged (cp.async) variants: every group of NW words lies inside the row and every row is 16-byte aligned
  bool aligned = (reinterpret_cast<uintptr_t>(in) & 15) == 0;
#ifdef CC_FACES_NO_ASYNC
  aligned = false;
#endif
#ifdef CC_FACES_NW8
  if (aligned && g.sx % 256 == 0 && sizeof(T) <= 4) launch_faces_staged<T, MODE, 8>(a, E, two_d, nych);
  else
#endif
  if (aligned && g.sx % 128 == 0) launch_faces_staged<T, MODE, 4>(a, E, two_d, nych);
  else {
    const unsigned nwg = (unsigned)((g.W + CC_FACE_NW - 1) / CC_FACE_NW);
    const i64 ntasks = (i64)nwg * nych * g.sz;
    const unsigned blocks = (unsigned)((ntasks + CC_FACE_WARPS - 1) / CC_FACE_WARPS);
    if (two_d) cc_launch(k_faces<T, MODE, false>, dim3(blocks), dim3(CC_FACE_WARPS * 32), (size_t)(0), a.stream, in, a.M, g, E, a.ctr, nych, nwg, (unsigned)ntasks);
    else cc_launch(k_faces<T, MODE, true>, dim3(blocks), dim3(CC_FACE_WARPS * 32), (size_t)(0), a.stream, in, a.M, g, E, a.ctr, nych, nwg, (unsigned)ntasks);
  }
  ++*a.launches;
  return 0;
}

template <typename T> int run_faces_stage(const LabelArgs& a) {
  switch (a.mode) {
    case MODE_EQ: return launch_faces<T, MODE_EQ>(a);
    case MODE_NONZERO: return launch_faces<T, MODE_NONZERO>(a);
    case MODE_DELTA: return launch_faces<T, MODE_DELTA>(a);
    case MODE_BLOCK:      // occupancy bytes of 2x2x2 blocks (cc3d_blocks.cuh)
      if constexpr (sizeof(T) == 1) return launch_faces<T, MODE_BLOCK>(a);
      else return -1;
  }
  return -1;
}

// CC3D_B200_B1=phased selects the CTA-phased tile kernels of round 1 (k_union_tile_hybrid) for A/B runs
static inline bool b1_phased() {
  static const bool v = []() { const char* e = getenv("CC3D_B200_B1"); return e && e[0] == 'p'; }();
  return v;
}
template <typename T, int MODE, int CONN>
static int launch_union(const LabelArgs& a, bool global_only = false) {
  Edge<T, MODE> E;
  memcpy(&E.delta, a.delta, sizeof(T));
  E.zeq = 0;
  const Geom& g = a.g;
  const T* in = static_cast<const T*>(a.in);
  if (global_only) {
    k_union_global<T, MODE, CONN><<<CC_QUEUE_BLOCKS, 256, 0, a.stream>>>(in, a.M, a.L, g, E, nullptr);
    *a.launches += 1;
    return 0;
  }
  const i64 ntx = (g.W + (1 << g.tw) - 1) >> g.tw, nty = (g.sy + (1 << g.ty) - 1) >> g.ty, ntz = (g.sz + (1 << g.tz) - 1) >> g.tz;
  static PerDeviceOnce once;
  const bool set_attr = once.first();
  // Which tile kernel (measured on B200, 512^3, profiles/r02b_b1_by_mode.md): the warp-owned kernel wins wherever the
  // diagonal candidates are sparse or absent (label volumes 0.28 -> 0.20 ms, binary 6-connected noise 0.76 -> 0.61 ms);
  // where every word carries candidates (binary 18/26-connected noise, continuous values, block grids, multilabel noise)
  // its per-warp item lists overflow and the CTA-phased kernels of round 1 stay ahead (18-conn. noise 2.5 vs 4.8 ms,
  // continuous 0.66 vs 1.13 ms, blocks 0.44 vs 0.76 ms).
  constexpr bool DIAG = CONN == 8 || CONN == 18 || CONN == 26;
  const bool use_w = !b1_phased() && nty < 65536 && ntz < 65536;
  // block grids: the per-word candidate loop with the balanced edge queue (k_union_tile) measured ahead of the item lists
  // (random binary 512^3: 0.380 vs 0.448 ms); CC3D_B200_BLOCK_ITEMS=1 selects the item lists
  static const bool block_items = getenv("CC3D_B200_BLOCK_ITEMS") != nullptr;
  if (MODE == MODE_BLOCK && !block_items) {
    if constexpr (MODE == MODE_BLOCK) {
      const size_t smem = (size_t)TileQueues<MODE>::SMEM_WORDS * 4;
      if (set_attr) cudaFuncSetAttribute(k_union_tile<T, MODE, CONN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      cc_launch(k_union_tile<T, MODE, CONN>, dim3((unsigned)(ntx * nty * ntz)), dim3(CC_TILE_THREADS), (size_t)(smem), a.stream, in, a.M, a.L, g, E, (u32)ntx, (u32)nty, a.GQ);
    }
  } else if constexpr (MODE == MODE_DELTA || MODE == MODE_BLOCK) {
    // continuous predicate / block nodes: edge-parallel item lists (every word has candidates that need a value test)
    const size_t smem = (size_t)CC_TILE_SMEM_WORDS * 4;
    if (set_attr) cudaFuncSetAttribute(k_union_tile_items<T, MODE, CONN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cc_launch(k_union_tile_items<T, MODE, CONN>, dim3((unsigned)(ntx * nty * ntz)), dim3(CC_TILE_THREADS), (size_t)(smem), a.stream, in, a.M, a.L, g, E, (u32)ntx, (u32)nty, a.GQ);
  } else if constexpr (MODE == MODE_NONZERO) {
    if (!DIAG && use_w) {
      // binary images hold at most 16 runs per word: 8 192 per tile
      if constexpr (!DIAG) {
        const size_t smem = (size_t)WarpTile<MODE, 8192>::SMEM_WORDS * 4;
        if (set_attr) cudaFuncSetAttribute(k_union_tile_w<T, MODE, CONN, 8192, false, 8192>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        BigTiles big; big.flags = nullptr; big.count = nullptr; big.defer = 0;
        cc_launch(k_union_tile_w<T, MODE, CONN, 8192, false, 8192>, dim3((unsigned)ntx, (unsigned)nty, (unsigned)ntz), dim3(CC_TILE_THREADS), (size_t)(smem), a.stream, in, a.M, a.L, g, E, a.GQ, big);
      }
    } else {
      const size_t smem = (size_t)TileQueues<MODE>::SMEM_WORDS * 4;
      if (set_attr) cudaFuncSetAttribute(k_union_tile<T, MODE, CONN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      cc_launch(k_union_tile<T, MODE, CONN>, dim3((unsigned)(ntx * nty * ntz)), dim3(CC_TILE_THREADS), (size_t)(smem), a.stream, in, a.M, a.L, g, E, (u32)ntx, (u32)nty, a.GQ);
    }
  } else if (MODE == MODE_EQ && use_w) {
    if constexpr (MODE == MODE_EQ) {
      // warp-owned tiles (cc3d_union_w.cuh): compact 32-bit forest of up to 4 096 runs per tile. Tiles with more runs - with
      // diagonals: more than 3 072, i.e. noise, where the candidates are dense - are flagged and, once the process has met
      // such a volume, relabelled by a second launch of the CTA-phased kernel with a 32-runs-per-word forest (measured faster on
      // noise than the warp-owned kernel with room for 16 384 runs: periodic 6-connected 1024^3 B1 6.4 vs 9.7 ms).
      constexpr u32 DENSE = DIAG ? 3072u : 4096u;
      const size_t smem = (size_t)WarpTile<MODE, 4096>::SMEM_WORDS * 4;
      const size_t smem5 = (size_t)HybridQueues<MODE, 5>::SMEM_WORDS * 4;
      if (set_attr) {
        cudaFuncSetAttribute(k_union_tile_w<T, MODE, CONN, 4096, false, DENSE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(k_union_tile_hybrid<T, MODE, CONN, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem5);
      }
      BigTiles big; big.flags = a.bigflags; big.count = a.nbig; big.defer = a.defer_big ? 1 : 0;
      cc_launch(k_union_tile_w<T, MODE, CONN, 4096, false, DENSE>, dim3((unsigned)ntx, (unsigned)nty, (unsigned)ntz), dim3(CC_TILE_THREADS), (size_t)(smem), a.stream, in, a.M, a.L, g, E, a.GQ, big);
      if (a.defer_big) {
        cc_launch(k_union_tile_hybrid<T, MODE, CONN, 5>, dim3((unsigned)(ntx * nty * ntz)), dim3(CC_TILE_THREADS), (size_t)(smem5), a.stream, in, a.M, a.L, g, E, (u32)ntx, (u32)nty, a.GQ, big);
        *a.launches += 1;
      }
    }
  } else if constexpr (MODE == MODE_EQ || MODE == MODE_MASK) {
    const size_t smem = (size_t)HybridQueues<MODE, 4>::SMEM_WORDS * 4, smem5 = (size_t)HybridQueues<MODE, 5>::SMEM_WORDS * 4;
    if (set_attr) {
      cudaFuncSetAttribute(k_union_tile_hybrid<T, MODE, CONN, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      cudaFuncSetAttribute(k_union_tile_hybrid<T, MODE, CONN, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem5);
    }
    BigTiles big; big.flags = a.bigflags; big.count = a.nbig; big.defer = a.defer_big ? 1 : 0;
    cc_launch(k_union_tile_hybrid<T, MODE, CONN, 4>, dim3((unsigned)(ntx * nty * ntz)), dim3(CC_TILE_THREADS), (size_t)(smem), a.stream, in, a.M, a.L, g, E, (u32)ntx, (u32)nty, a.GQ, big);
    if (a.defer_big) {
      cc_launch(k_union_tile_hybrid<T, MODE, CONN, 5>, dim3((unsigned)(ntx * nty * ntz)), dim3(CC_TILE_THREADS), (size_t)(smem5), a.stream, in, a.M, a.L, g, E, (u32)ntx, (u32)nty, a.GQ, big);
      *a.launches += 1;
    }
  }
  if (a.mark) a.mark("B1_union_tile", a.stream);
  // dense queues (noise) elect one lane per pair of tile roots; CC3D_B200_B2=1: never, =2: always
  static const int b2_mode = getenv("CC3D_B200_B2") ? atoi(getenv("CC3D_B200_B2")) : 0;
  const u32 dense_above = b2_mode == 1 ? 0xFFFFFFFFu : (b2_mode == 2 ? 0u : (u32)(g.nwords >> 1));
  cc_launch(k_union_queue, dim3(CC_QUEUE_BLOCKS * 4), dim3(256), (size_t)(0), a.stream, a.L, a.GQ, dense_above);
  *a.launches += 2;
  if (a.inline_fallback) {
    cc_launch(k_union_global<T, MODE, CONN>, dim3(CC_QUEUE_BLOCKS), dim3(256), (size_t)(0), a.stream, in, a.M, a.L, g, E, (const u32*)a.GQ.ovf);
    *a.launches += 1;
  }
  if (a.mark) a.mark("B2_union_queue", a.stream);
  return 0;
}

// Voxel values are only read for the diagonal candidates of EQ / DELTA with 8/18/26 neighbours; every
// other configuration runs the uint8_t instantiation.
template <typename T, int MODE>
static int launch_union_conn(const LabelArgs& a, bool global_only = false) {
  switch (a.connectivity) {
    case 4: return launch_union<uint8_t, MODE == MODE_DELTA ? MODE_EQ : MODE, 4>(a, global_only);
    case 6: return launch_union<uint8_t, MODE == MODE_DELTA ? MODE_EQ : MODE, 6>(a, global_only);
    case 8: return launch_union<T, MODE, 8>(a, global_only);
    case 18: return launch_union<T, MODE, 18>(a, global_only);
    case 26: return launch_union<T, MODE, 26>(a, global_only);
  }
  return -1;
}

template <typename T> static int union_stage(const LabelArgs& a, bool global_only) {
  switch (a.mode) {
    case MODE_EQ: return launch_union_conn<T, MODE_EQ>(a, global_only);
    case MODE_NONZERO: return launch_union_conn<uint8_t, MODE_NONZERO>(a, global_only);
    case MODE_DELTA: return launch_union_conn<T, MODE_DELTA>(a, global_only);
    case MODE_MASK: return launch_union<uint8_t, MODE_MASK, 8>(a, global_only);
    case MODE_BLOCK:
      if constexpr (sizeof(T) == 1) return launch_union<uint8_t, MODE_BLOCK, 26>(a, global_only);
      else return -1;
  }
  return -1;
}
template <typename T> int run_union_stage(const LabelArgs& a) { return union_stage<T>(a, false); }
template <typename T> int run_union_global_stage(const LabelArgs& a) { return union_stage<T>(a, true); }

template <typename T, int MODE, int CONN>
static int launch_periodic(const LabelArgs& a) {
  Edge<T, MODE> E;
  memcpy(&E.delta, a.delta, sizeof(T));
  E.zeq = 0;
  const Geom& g = a.g;
  const T* in = static_cast<const T*>(a.in);
  const i64 n0 = 2 * g.sy * g.sz, n1 = g.sx * g.sz, n2 = g.sx * g.sy;
  k_periodic<T, MODE, CONN><<<(unsigned)((n0 + 255) / 256), 256, 0, a.stream>>>(in, a.M, a.L, g, E, 0);
  k_periodic<T, MODE, CONN><<<(unsigned)((n1 + 255) / 256), 256, 0, a.stream>>>(in, a.M, a.L, g, E, 1);
  if (CONN == 6) k_periodic<T, MODE, CONN><<<(unsigned)((n2 + 255) / 256), 256, 0, a.stream>>>(in, a.M, a.L, g, E, 2);
  *a.launches += (CONN == 6) ? 3 : 2;
  return 0;
}

template <typename T, int MODE>
static int launch_periodic_conn(const LabelArgs& a) {
  switch (a.connectivity) {
    case 4: return launch_periodic<T, MODE, 4>(a);
    case 8: return launch_periodic<T, MODE, 8>(a);
    case 6: return launch_periodic<T, MODE, 6>(a);
  }
  return -1;
}

template <typename T> int run_periodic_stage(const LabelArgs& a) {
  switch (a.mode) {
    case MODE_EQ: return launch_periodic_conn<T, MODE_EQ>(a);
    case MODE_NONZERO: return launch_periodic_conn<T, MODE_NONZERO>(a);
  }
  return -1;
}
#endif
