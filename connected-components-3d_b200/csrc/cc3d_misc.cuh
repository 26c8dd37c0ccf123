// cc3d_misc.cuh — pre-pass (epl / foreground rows / value range), per-label statistics and the
// dust masking kernel.
#pragma once
#include "cc3d_common.cuh"

// ---- row a1: estimate_provisional_label_count (cc3d.hpp:287-315) + value range ----
template <typename T>
__global__ void __launch_bounds__(256)
k_prepass(const T* __restrict__ in, Geom g, Counters* __restrict__ ctr, T* __restrict__ part_min, T* __restrict__ part_max) {
  const int lane = threadIdx.x & 31;
  const i64 nwarps_total = ((i64)gridDim.x * blockDim.x) >> 5;
  const i64 nseg_total = g.sy * g.sz * g.W;
  u32 epl = 0;
  i64 rmin = INT64_MAX, rmax = -1;
  T mn = in[0], mx = in[0];
  for (i64 wid = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5; wid < nseg_total; wid += nwarps_total) {
    const i64 row = wid / g.W, seg = wid - row * g.W;
    const i64 x = seg * 32 + lane;
    const bool inx = x < g.sx;
    const T v = inx ? in[row * g.sx + x] : (T)0;
    T vl = __shfl_up_sync(CC_FULL, v, 1);
    if (lane == 0) vl = (x > 0 && inx) ? in[row * g.sx + x - 1] : (T)0;
    const bool tr = inx && v != (T)0 && (x == 0 || v != vl);
    const u32 tm = __ballot_sync(CC_FULL, tr);
    if (tm) { epl += __popc(tm); rmin = min(rmin, row); rmax = max(rmax, row); }
    if (inx) { if (v < mn) mn = v; if (v > mx) mx = v; }
  }
  // warp reduce min/max
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const T a = __shfl_xor_sync(CC_FULL, mn, o), b = __shfl_xor_sync(CC_FULL, mx, o);
    if (a < mn) mn = a;
    if (b > mx) mx = b;
  }
  __shared__ u32 s_epl;
  __shared__ long long s_rmin, s_rmax;
  __shared__ T s_mn[8], s_mx[8];
  if (threadIdx.x == 0) { s_epl = 0; s_rmin = INT64_MAX; s_rmax = -1; }
  __syncthreads();
  if (lane == 0) {
    s_mn[threadIdx.x >> 5] = mn; s_mx[threadIdx.x >> 5] = mx;
    if (epl) { atomicAdd(&s_epl, epl); atomicMin(&s_rmin, (long long)rmin); atomicMax(&s_rmax, (long long)rmax); }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (s_epl) {
      atomicAdd((unsigned long long*)&ctr->epl, (unsigned long long)s_epl);
      atomicMin((long long*)&ctr->first_row, s_rmin);
      atomicMax((long long*)&ctr->last_row, s_rmax);
    }
    T a = s_mn[0], b = s_mx[0];
    for (int k = 1; k < 8; k++) { if (s_mn[k] < a) a = s_mn[k]; if (s_mx[k] > b) b = s_mx[k]; }
    part_min[blockIdx.x] = a; part_max[blockIdx.x] = b;
  }
}

template <typename T>
__global__ void k_minmax_final(const T* __restrict__ part_min, const T* __restrict__ part_max, int n, T* __restrict__ out2) {
  // single warp
  T mn = part_min[0], mx = part_max[0];
  for (int i = threadIdx.x; i < n; i += 32) { if (part_min[i] < mn) mn = part_min[i]; if (part_max[i] > mx) mx = part_max[i]; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const T a = __shfl_xor_sync(CC_FULL, mn, o), b = __shfl_xor_sync(CC_FULL, mx, o);
    if (a < mn) mn = a;
    if (b > mx) mx = b;
  }
  if (threadIdx.x == 0) { out2[0] = mn; out2[1] = mx; }
}

// ---- row a13: statistics (fastcc3d.pyx:771-938). Memory-axis coordinates. Runs of equal labels inside
// a warp are aggregated, then merged in a per-CTA shared-memory table that is flushed with global
// atomics once per CTA (persistent CTAs), so that a giant component costs O(#CTAs) global atomics. ----
#define CC_STAT_SLOTS 1024
struct StatTable {
  u32 key[CC_STAT_SLOTS];   // label + 1, 0 = empty
  u32 cnt[CC_STAT_SLOTS];
  u32 bb[CC_STAT_SLOTS][6];
  unsigned long long sum[CC_STAT_SLOTS][3];
};

__device__ __forceinline__ void stat_global(u32 l, u32 len, u32 x0, u32 x1, u32 y, u32 z, unsigned long long sx_,
                                            u32* counts, u32* bbox, unsigned long long* sums) {
  atomicAdd(&counts[l], len);
  u32* b = bbox + 6 * (size_t)l;
  atomicMin(&b[0], x0); atomicMax(&b[1], x1);
  atomicMin(&b[2], y); atomicMax(&b[3], y);
  atomicMin(&b[4], z); atomicMax(&b[5], z);
  unsigned long long* s = sums + 3 * (size_t)l;
  atomicAdd(&s[0], sx_);
  atomicAdd(&s[1], (unsigned long long)y * len);
  atomicAdd(&s[2], (unsigned long long)z * len);
}

template <typename LT>
__global__ void __launch_bounds__(256)
k_statistics(const LT* __restrict__ labels, Geom g, u64 N, u32* __restrict__ counts, u32* __restrict__ bbox,
             unsigned long long* __restrict__ sums) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  StatTable& tb = *reinterpret_cast<StatTable*>(smem_raw);
  for (int i = threadIdx.x; i < CC_STAT_SLOTS; i += blockDim.x) {
    tb.key[i] = 0; tb.cnt[i] = 0;
    tb.bb[i][0] = tb.bb[i][2] = tb.bb[i][4] = 0xFFFFFFFFu;
    tb.bb[i][1] = tb.bb[i][3] = tb.bb[i][5] = 0;
    tb.sum[i][0] = tb.sum[i][1] = tb.sum[i][2] = 0;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const i64 nwarps_total = ((i64)gridDim.x * blockDim.x) >> 5;
  const i64 nseg_total = g.sy * g.sz * g.W;
  for (i64 wid = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5; wid < nseg_total; wid += nwarps_total) {
    const i64 row = wid / g.W, seg = wid - row * g.W;
    const i64 x = seg * 32 + lane;
    const bool inx = x < g.sx;
    const u64 l64 = inx ? (u64)labels[row * g.sx + x] : ~0ull;
    const bool ok = inx && l64 <= N;
    const u32 l = ok ? (u32)l64 : 0xFFFFFFFFu;
    const u32 prev = __shfl_up_sync(CC_FULL, l, 1);
    const bool head = lane == 0 || prev != l;
    const u32 heads = __ballot_sync(CC_FULL, head);
    if (head && ok) {
      const u32 after = heads & ~((2u << lane) - 1u);  // heads strictly above this lane
      const int end = after ? (__ffs(after) - 1) : 32;   // one past the run's last lane
      const u32 len = end - lane;
      const u32 x0 = (u32)x, x1 = (u32)(x + len - 1);
      const unsigned long long sumx = (unsigned long long)len * x0 + (unsigned long long)len * (len - 1) / 2;
      const u32 z = (u32)(row / g.sy), y = (u32)(row - (i64)z * g.sy);
      // find / claim a slot
      u32 h = (l * 2654435761u) >> 22;  // 10 bits
      int slot = -1;
#pragma unroll 1
      for (int probe = 0; probe < 16; probe++) {
        const u32 s = (h + probe) & (CC_STAT_SLOTS - 1);
        const u32 k = tb.key[s];
        if (k == l + 1) { slot = s; break; }
        if (k == 0) {
          const u32 old = atomicCAS(&tb.key[s], 0u, l + 1);
          if (old == 0 || old == l + 1) { slot = s; break; }
        }
      }
      if (slot >= 0) {
        atomicAdd(&tb.cnt[slot], len);
        atomicMin(&tb.bb[slot][0], x0); atomicMax(&tb.bb[slot][1], x1);
        atomicMin(&tb.bb[slot][2], y); atomicMax(&tb.bb[slot][3], y);
        atomicMin(&tb.bb[slot][4], z); atomicMax(&tb.bb[slot][5], z);
        atomicAdd(&tb.sum[slot][0], sumx);
        atomicAdd(&tb.sum[slot][1], (unsigned long long)y * len);
        atomicAdd(&tb.sum[slot][2], (unsigned long long)z * len);
      } else {
        stat_global(l, len, x0, x1, y, z, sumx, counts, bbox, sums);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < CC_STAT_SLOTS; i += blockDim.x) {
    if (tb.key[i] == 0) continue;
    const u32 l = tb.key[i] - 1;
    atomicAdd(&counts[l], tb.cnt[i]);
    u32* b = bbox + 6 * (size_t)l;
    atomicMin(&b[0], tb.bb[i][0]); atomicMax(&b[1], tb.bb[i][1]);
    atomicMin(&b[2], tb.bb[i][2]); atomicMax(&b[3], tb.bb[i][3]);
    atomicMin(&b[4], tb.bb[i][4]); atomicMax(&b[5], tb.bb[i][5]);
    unsigned long long* s = sums + 3 * (size_t)l;
    atomicAdd(&s[0], tb.sum[i][0]); atomicAdd(&s[1], tb.sum[i][1]); atomicAdd(&s[2], tb.sum[i][2]);
  }
}

__global__ void __launch_bounds__(256) k_stat_init(u32* __restrict__ counts, u32* __restrict__ bbox, unsigned long long* __restrict__ sums, u64 n) {
  const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  counts[i] = 0;
  bbox[6 * i + 0] = bbox[6 * i + 2] = bbox[6 * i + 4] = 0xFFFFFFFFu;
  bbox[6 * i + 1] = bbox[6 * i + 3] = bbox[6 * i + 5] = 0;
  sums[3 * i] = sums[3 * i + 1] = sums[3 * i + 2] = 0;
}

// ---- row a14: dust masking (cc3d/__init__.py:148-150): img[i] = keep[label[i]] ? img[i] : 0 ----
template <typename IT, typename LT>
__global__ void __launch_bounds__(256)
k_mask_by_label(IT* __restrict__ img, const LT* __restrict__ labels, const unsigned char* __restrict__ keep, u64 N, i64 n) {
  const i64 stride = (i64)gridDim.x * blockDim.x;
  for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const u64 l = (u64)labels[i];
    const bool k = l <= N && keep[l];
    if (!k) img[i] = (IT)0;
  }
}
