// cc3d_runs.cuh — SURVEY 8(f)4: run-length index of a label volume and per-label rendering.
//   R1 k_runs_count(_vec) : run starts per 4096-voxel chunk      (reads the labels once)
//   (scan of the chunk counts: scan_counts, cc3d_b200.cu)
//   R2 k_runs_emit(_vec)  : (value, start, end) of every non-zero run (reads the labels once, writes 24 B per run)
//   The _vec kernels (128-bit loads) need 16-byte aligned labels; the scalar ones serve unaligned views.
//   W1 k_draw_check : validates a run list on the device
//   W2 k_draw_short / k_draw_long : image[start..end) = value
// Replaces extract_runs / set_run_voxels (reference cc3d_graphs.hpp:470-523). A run is a maximal stretch of equal
// non-zero values of the flattened array (runs continue across row ends, exactly like the reference's 1-D walk).
#pragma once
#include "cc3d_common.cuh"

#define CC_RUN_CHUNK 4096      // voxels per block pass: 8 warps x 16 steps x 32 lanes
#define CC_RUN_STEPS 16
#define CC_RUN_LONG 16384ull   // runs longer than this are finished by the whole grid (k_draw_long)

// Start / end flags of the voxel each lane holds. prev / next come from the neighbouring lanes; the two lanes at
// the ends of the 32-voxel group read one extra element (same or adjacent sector: served by L1).
template <typename T>
__device__ __forceinline__ void run_flags(const T* __restrict__ lab, i64 j, i64 n, int lane, T& v, bool& is_start,
                                          bool& is_end) {
  const bool in = j < n;
  v = in ? lab[j] : T(0);
  T prev = __shfl_up_sync(CC_FULL, v, 1);
  T next = __shfl_down_sync(CC_FULL, v, 1);
  if (lane == 0) prev = (in && j > 0) ? lab[j - 1] : T(0);
  if (lane == 31) next = (j + 1 < n) ? lab[j + 1] : T(0);
  const bool first = j == 0, last = j + 1 >= n;
  is_start = in && v != T(0) && (first || prev != v);
  is_end = in && v != T(0) && (last || next != v);
}

// R1: cnt[chunk] = number of run starts in the chunk. Grid-stride over chunks.
template <typename T>
__global__ void __launch_bounds__(256) k_runs_count(const T* __restrict__ lab, i64 n, u32* __restrict__ cnt, i64 nchunks) {
  __shared__ u32 s_w[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (i64 c = blockIdx.x; c < nchunks; c += gridDim.x) {
    const i64 base = c * CC_RUN_CHUNK + (i64)warp * (CC_RUN_STEPS * 32);
    u32 total = 0;
#pragma unroll 4
    for (int k = 0; k < CC_RUN_STEPS; k++) {
      const i64 j = base + k * 32 + lane;
      T v; bool st, en;
      run_flags(lab, j, n, lane, v, st, en);
      total += __popc(__ballot_sync(CC_FULL, st));
    }
    if (lane == 0) s_w[warp] = total;
    __syncthreads();
    if (threadIdx.x == 0) {
      u32 t = 0;
#pragma unroll
      for (int w = 0; w < 8; w++) t += s_w[w];
      cnt[c] = t;
    }
    __syncthreads();
  }
}

// R2: run k (position order) starts at the k-th start flag; its last voxel j is the one with an end flag whose
// inclusive start rank is k + 1, so both sides are placed with the single scan of the start counts.
template <typename T>
__global__ void __launch_bounds__(256)
k_runs_emit(const T* __restrict__ lab, i64 n, const u32* __restrict__ prefix, i64 nchunks, u64* __restrict__ values,
            u64* __restrict__ starts, u64* __restrict__ ends) {
  __shared__ u32 s_w[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const u32 below = (1u << lane) - 1u;
  for (i64 c = blockIdx.x; c < nchunks; c += gridDim.x) {
    const i64 base = c * CC_RUN_CHUNK + (i64)warp * (CC_RUN_STEPS * 32);
    u32 sb[CC_RUN_STEPS];   // start ballots of this warp's 16 steps (warp-uniform)
    u32 total = 0;
#pragma unroll
    for (int k = 0; k < CC_RUN_STEPS; k++) {
      const i64 j = base + k * 32 + lane;
      T v; bool st, en;
      run_flags(lab, j, n, lane, v, st, en);
      sb[k] = __ballot_sync(CC_FULL, st);
      total += __popc(sb[k]);
    }
    if (lane == 0) s_w[warp] = total;
    __syncthreads();
    u32 rank = prefix[c];
    for (int w = 0; w < warp; w++) rank += s_w[w];
#pragma unroll
    for (int k = 0; k < CC_RUN_STEPS; k++) {
      const i64 j = base + k * 32 + lane;
      T v; bool st, en;
      run_flags(lab, j, n, lane, v, st, en);   // second read of the chunk: L1 / L2 resident
      const u32 ex = rank + __popc(sb[k] & below);
      if (st) { values[ex] = (u64)v; starts[ex] = (u64)j; }
      if (en) ends[ex + (st ? 1u : 0u) - 1u] = (u64)j + 1ull;
      rank += __popc(sb[k]);
    }
    __syncthreads();
  }
}

// ---- 128-bit variants (labels 16-byte aligned): every lane holds 16 / sizeof(T) consecutive voxels per step, a warp
// covers its 512-voxel segment in sizeof(T) steps, so the neighbour exchange costs two shuffles per 16 bytes and the
// chunk is read ONCE (R2 keeps its 16 voxels per lane in registers between the count and the emission). ----
template <typename T> __device__ __forceinline__ T shfl_up1(T v) {
  if constexpr (sizeof(T) < 4) return (T)__shfl_up_sync(CC_FULL, (unsigned)v, 1);
  else return __shfl_up_sync(CC_FULL, v, 1);
}
template <typename T> __device__ __forceinline__ T shfl_down1(T v) {
  if constexpr (sizeof(T) < 4) return (T)__shfl_down_sync(CC_FULL, (unsigned)v, 1);
  else return __shfl_down_sync(CC_FULL, v, 1);
}

template <typename T>
__device__ __forceinline__ void run_load_vec(const T* __restrict__ lab, i64 j0, i64 n, T (&v)[16 / sizeof(T)]) {
  constexpr int VEC = 16 / sizeof(T);
  if (j0 + VEC <= n) {
    union { uint4 q; T e[VEC]; } u;
    u.q = *reinterpret_cast<const uint4*>(lab + j0);
#pragma unroll
    for (int i = 0; i < VEC; i++) v[i] = u.e[i];
  } else {
#pragma unroll
    for (int i = 0; i < VEC; i++) v[i] = (j0 + i < n) ? lab[j0 + i] : T(0);
  }
}

// bit i of sm / em: voxel j0 + i starts / ends a run (voxels past the end hold 0 and never flag)
template <typename T>
__device__ __forceinline__ void run_masks_vec(const T* __restrict__ lab, i64 j0, i64 n, int lane,
                                              const T (&v)[16 / sizeof(T)], u32& sm, u32& em) {
  constexpr int VEC = 16 / sizeof(T);
  T prev = shfl_up1(v[VEC - 1]);
  T next = shfl_down1(v[0]);
  if (lane == 0) prev = (j0 > 0 && j0 - 1 < n) ? lab[j0 - 1] : T(0);
  if (lane == 31) next = (j0 + VEC < n) ? lab[j0 + VEC] : T(0);
  sm = 0; em = 0;
#pragma unroll
  for (int i = 0; i < VEC; i++) {
    const T p = i ? v[i ? i - 1 : 0] : prev;
    const T q = (i + 1 < VEC) ? v[(i + 1 < VEC) ? i + 1 : i] : next;
    if (v[i] != T(0) && p != v[i]) sm |= 1u << i;
    if (v[i] != T(0) && q != v[i]) em |= 1u << i;
  }
}

template <typename T>
__global__ void __launch_bounds__(256) k_runs_count_vec(const T* __restrict__ lab, i64 n, u32* __restrict__ cnt, i64 nchunks) {
  constexpr int VEC = 16 / sizeof(T), STEPS = 16 / VEC;
  __shared__ u32 s_w[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (i64 c = blockIdx.x; c < nchunks; c += gridDim.x) {
    const i64 base = c * CC_RUN_CHUNK + (i64)warp * 512;
    T v[STEPS][VEC];
#pragma unroll
    for (int k = 0; k < STEPS; k++) run_load_vec(lab, base + (i64)(k * 32 + lane) * VEC, n, v[k]);
    u32 total = 0;
#pragma unroll
    for (int k = 0; k < STEPS; k++) {
      u32 sm, em;
      run_masks_vec(lab, base + (i64)(k * 32 + lane) * VEC, n, lane, v[k], sm, em);
      total += __popc(sm);
    }
    total = __reduce_add_sync(CC_FULL, total);
    if (lane == 0) s_w[warp] = total;
    __syncthreads();
    if (threadIdx.x == 0) {
      u32 t = 0;
#pragma unroll
      for (int w = 0; w < 8; w++) t += s_w[w];
      cnt[c] = t;
    }
    __syncthreads();
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
k_runs_emit_vec(const T* __restrict__ lab, i64 n, const u32* __restrict__ prefix, i64 nchunks, u64* __restrict__ values,
                u64* __restrict__ starts, u64* __restrict__ ends) {
  constexpr int VEC = 16 / sizeof(T), STEPS = 16 / VEC;
  __shared__ u32 s_w[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (i64 c = blockIdx.x; c < nchunks; c += gridDim.x) {
    const i64 base = c * CC_RUN_CHUNK + (i64)warp * 512;
    T v[STEPS][VEC];
    u32 sm[STEPS], em[STEPS], ex[STEPS];   // ex: start rank of the lane's first voxel within the warp segment
#pragma unroll
    for (int k = 0; k < STEPS; k++) run_load_vec(lab, base + (i64)(k * 32 + lane) * VEC, n, v[k]);
    u32 seg = 0;   // starts of the steps before k (warp-uniform)
#pragma unroll
    for (int k = 0; k < STEPS; k++) {
      run_masks_vec(lab, base + (i64)(k * 32 + lane) * VEC, n, lane, v[k], sm[k], em[k]);
      const u32 cnt_lane = __popc(sm[k]);
      u32 inc = cnt_lane;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const u32 t = __shfl_up_sync(CC_FULL, inc, o);
        if (lane >= o) inc += t;
      }
      ex[k] = seg + inc - cnt_lane;
      seg += __shfl_sync(CC_FULL, inc, 31);
    }
    if (lane == 0) s_w[warp] = seg;
    __syncthreads();
    u32 rank = prefix[c];
    for (int w = 0; w < warp; w++) rank += s_w[w];
#pragma unroll
    for (int k = 0; k < STEPS; k++) {
      const i64 j0 = base + (i64)(k * 32 + lane) * VEC;
      const u32 r0 = rank + ex[k];
      if (sm[k] | em[k]) {
#pragma unroll
        for (int i = 0; i < VEC; i++) {
          const u32 before = r0 + __popc(sm[k] & ((1u << i) - 1u));
          if (sm[k] >> i & 1u) { values[before] = (u64)v[k][i]; starts[before] = (u64)(j0 + i); }
          if (em[k] >> i & 1u) ends[before + (sm[k] >> i & 1u) - 1u] = (u64)(j0 + i) + 1ull;
        }
      }
    }
    __syncthreads();
  }
}

// W1: flags[0] != 0 when any run is invalid (reference set_run_voxels, cc3d_graphs.hpp:511-517:
// start >= end or end > voxels); flags[1] = number of runs longer than CC_RUN_LONG (size of the long list).
__global__ void __launch_bounds__(256)
k_draw_check(const u64* __restrict__ starts, const u64* __restrict__ ends, u64 n_runs, u64 voxels,
             unsigned long long* __restrict__ flags) {
  bool bad = false;
  u32 n_long = 0;
  for (u64 r = (u64)blockIdx.x * blockDim.x + threadIdx.x; r < n_runs; r += (u64)gridDim.x * blockDim.x) {
    const u64 a = starts[r], b = ends[r];
    const bool ok = (a < b) && (b <= voxels);
    bad |= !ok;
    n_long += (ok && b - a > CC_RUN_LONG) ? 1u : 0u;
  }
  const u32 any_bad = __ballot_sync(CC_FULL, bad);
  n_long = __reduce_add_sync(CC_FULL, n_long);
  if ((threadIdx.x & 31) == 0) {
    if (any_bad) atomicOr(&flags[0], 1ull);
    if (n_long) atomicAdd(&flags[1], (unsigned long long)n_long);
  }
}

// W2a: one warp per run; the first CC_RUN_LONG voxels of a run are written here, the rest of a long run goes to
// the long list (long_runs[2 * i] = start, [2 * i + 1] = end; *n_long counts them). `origin` is subtracted from the
// run positions (host images are staged as the window [origin, origin + window) only).
template <typename T>
__global__ void __launch_bounds__(256)
k_draw_short(T* __restrict__ img, T value, const u64* __restrict__ starts, const u64* __restrict__ ends, u64 n_runs,
             u64 origin, u64* __restrict__ long_runs, unsigned long long* __restrict__ n_long) {
  const int lane = threadIdx.x & 31;
  const u64 warps = ((u64)gridDim.x * blockDim.x) >> 5;
  for (u64 r = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n_runs; r += warps) {
    const u64 a = starts[r] - origin, b = ends[r] - origin;
    const u64 stop = (b - a > CC_RUN_LONG) ? a + CC_RUN_LONG : b;
    for (u64 i = a + lane; i < stop; i += 32) img[i] = value;
    if (stop < b && lane == 0) {
      const unsigned long long k = atomicAdd(n_long, 1ull);
      long_runs[2 * k] = stop; long_runs[2 * k + 1] = b;
    }
  }
}

// W2b: every long run is written by the whole grid.
template <typename T>
__global__ void __launch_bounds__(256)
k_draw_long(T* __restrict__ img, T value, const u64* __restrict__ long_runs, const unsigned long long* __restrict__ n_long) {
  const u64 n = *n_long;
  const u64 stride = (u64)gridDim.x * blockDim.x;
  for (u64 r = 0; r < n; r++) {
    const u64 a = long_runs[2 * r], b = long_runs[2 * r + 1];
    for (u64 i = a + (u64)blockIdx.x * blockDim.x + threadIdx.x; i < b; i += stride) img[i] = value;
  }
}
