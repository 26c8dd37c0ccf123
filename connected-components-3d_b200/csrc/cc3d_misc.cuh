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
      track_rows(ctr, (u64)s_rmin, (u64)s_rmax);
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

// ---- row a13: statistics (fastcc3d.pyx:771-938). Memory-axis coordinates.
// Three levels of accumulation, each fed by the one below only when a label ends:
//   lane  : one lane per voxel column. A warp owns a 32-voxel word column of one z-plane and walks
//           CC_STAT_YCH rows down y; a lane keeps the label and first row of its current VERTICAL run in
//           registers (label volumes have vertical runs of tens to hundreds of voxels).
//   warp  : when lanes finish runs, the lanes that finish the same label are reduced (redux.sync) and added to a
//           per-warp table of 8 labels in shared memory. Its fields are relative to the warp's task (x - xbase,
//           y - y0, at most 2048 voxels), so everything is 32-bit and the update is a plain read-modify-write by one
//           lane: no atomics, no hashing.
//   CTA   : at the end of a task (or when the warp table is full) its entries go to a per-CTA hash table with
//           absolute coordinates, native 32-bit shared atomics only (64-bit sums are lo/hi pairs with an explicit
//           carry; a 64-bit shared atomicAdd compiles to a compare-and-swap spin loop). Persistent CTAs flush the
//           table with global atomics once, so a giant component costs O(#CTAs) global atomics. ----
#define CC_STAT_SLOTS 512
#define CC_STAT_YCH 64
#define CC_STAT_UNR 8
#define CC_STAT_WSLOTS 8
struct StatTable {
  u32 key[CC_STAT_SLOTS];   // label + 1, 0 = empty
  u32 cnt[CC_STAT_SLOTS];
  u32 bb[CC_STAT_SLOTS][6];
  u32 sumlo[CC_STAT_SLOTS][3];
  u32 sumhi[CC_STAT_SLOTS][3];
  // per-warp tables (8 warps): task-relative
  u32 wkey[8][CC_STAT_WSLOTS];   // label + 1, 0 = empty
  u32 wcnt[8][CC_STAT_WSLOTS];
  u32 wsx[8][CC_STAT_WSLOTS];    // sum of (x - xbase)
  u32 wsy[8][CC_STAT_WSLOTS];    // sum of 2 * (y - y0)
  u32 wbb[8][CC_STAT_WSLOTS];    // xmin | xmax << 8 | ymin << 16 | ymax << 24 (relative)
};

// 64-bit add on a {lo, hi} pair of shared 32-bit words
__device__ __forceinline__ void sm_add64(u32* lo, u32* hi, unsigned long long v) {
  const u32 vl = (u32)v, vh = (u32)(v >> 32);
  const u32 old = atomicAdd(lo, vl);
  const u32 carry = (old + vl < old) ? 1u : 0u;
  if (vh + carry) atomicAdd(hi, vh + carry);
}

template <typename LT>
__global__ void __launch_bounds__(256, 4)     // 148 * 4 persistent CTAs: all of them have to be resident (<= 64 registers)
k_statistics(const LT* __restrict__ labels, Geom g, u64 N, u32* __restrict__ counts, u32* __restrict__ bbox,
             unsigned long long* __restrict__ sums, unsigned long long* __restrict__ maxout) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  StatTable& tb = *reinterpret_cast<StatTable*>(smem_raw);
  for (int i = threadIdx.x; i < CC_STAT_SLOTS; i += blockDim.x) {
    tb.key[i] = 0; tb.cnt[i] = 0;
    tb.bb[i][0] = tb.bb[i][2] = tb.bb[i][4] = 0xFFFFFFFFu;
    tb.bb[i][1] = tb.bb[i][3] = tb.bb[i][5] = 0;
    tb.sumlo[i][0] = tb.sumlo[i][1] = tb.sumlo[i][2] = 0;
    tb.sumhi[i][0] = tb.sumhi[i][1] = tb.sumhi[i][2] = 0;
  }
  if (threadIdx.x < 8 * CC_STAT_WSLOTS) (&tb.wkey[0][0])[threadIdx.x] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const u32 sx = (u32)g.sx, sy = (u32)g.sy;
  const i64 W = g.W;
  const i64 nych = (g.sy + CC_STAT_YCH - 1) / CC_STAT_YCH;
  const i64 ntasks = W * nych * g.sz;
  constexpr u32 NONE = 0xFFFFFFFFu;
  const LT nmax = (u64)(LT)~(LT)0 <= N ? (LT)~(LT)0 : (LT)N;   // labels above N are ignored
  u32* wkey = tb.wkey[warp]; u32* wcnt = tb.wcnt[warp]; u32* wsx = tb.wsx[warp]; u32* wsy = tb.wsy[warp]; u32* wbb = tb.wbb[warp];

  // CTA table: cnt voxels of label l with absolute sums and box
  auto cta_add = [&](u32 l, u32 cnt, unsigned long long sumx, unsigned long long sumy, unsigned long long sumz,
                     u32 xmin, u32 xmax, u32 ymin, u32 ymax, u32 z) {
    u32 h = (l * 2654435761u) >> 23;  // 9 bits
    int slot = -1;
#pragma unroll 1
    for (int probe = 0; probe < 16; probe++) {
      const u32 s = (h + probe) & (CC_STAT_SLOTS - 1);
      const u32 k = *(volatile u32*)&tb.key[s];
      if (k == l + 1) { slot = s; break; }
      if (k == 0) {
        const u32 old = atomicCAS(&tb.key[s], 0u, l + 1);
        if (old == 0 || old == l + 1) { slot = s; break; }
      }
    }
    if (slot < 0) {   // table full: straight to global memory
      atomicAdd(&counts[l], cnt);
      u32* b = bbox + 6 * (size_t)l;
      atomicMin(&b[0], xmin); atomicMax(&b[1], xmax);
      atomicMin(&b[2], ymin); atomicMax(&b[3], ymax);
      atomicMin(&b[4], z); atomicMax(&b[5], z);
      unsigned long long* sg = sums + 3 * (size_t)l;
      atomicAdd(&sg[0], sumx); atomicAdd(&sg[1], sumy); atomicAdd(&sg[2], sumz);
      return;
    }
    atomicAdd(&tb.cnt[slot], cnt);
    volatile u32* b = tb.bb[slot];   // most additions do not move the box: read before the atomic
    if (xmin < b[0]) atomicMin(&tb.bb[slot][0], xmin);
    if (xmax > b[1]) atomicMax(&tb.bb[slot][1], xmax);
    if (ymin < b[2]) atomicMin(&tb.bb[slot][2], ymin);
    if (ymax > b[3]) atomicMax(&tb.bb[slot][3], ymax);
    if (z < b[4]) atomicMin(&tb.bb[slot][4], z);
    if (z > b[5]) atomicMax(&tb.bb[slot][5], z);
    sm_add64(&tb.sumlo[slot][0], &tb.sumhi[slot][0], sumx);
    sm_add64(&tb.sumlo[slot][1], &tb.sumhi[slot][1], sumy);
    sm_add64(&tb.sumlo[slot][2], &tb.sumhi[slot][2], sumz);
  };
  // entry e of this warp's table -> CTA table (absolute coordinates of the task: xbase, y0, z)
  auto spill = [&](int e, u32 xbase, u32 y0, u32 z) {
    const u32 k = wkey[e];
    if (!k) return;
    const u32 cnt = wcnt[e], bb = wbb[e];
    cta_add(k - 1, cnt, (unsigned long long)xbase * cnt + wsx[e], (unsigned long long)y0 * cnt + (wsy[e] >> 1),
            (unsigned long long)z * cnt, xbase + (bb & 0xFFu), xbase + ((bb >> 8) & 0xFFu), y0 + ((bb >> 16) & 0xFFu),
            y0 + (bb >> 24), z);
    wkey[e] = 0;
  };
  // warp step (convergent): lanes with need == true finish the run of label cur over rows ystart..yend (relative
  // to y0); one label per iteration of the loop
  auto finish_runs = [&](bool need, u32 cur, u32 ystart, u32 yend, u32 xbase, u32 y0, u32 z) {
    u32 m = __ballot_sync(CC_FULL, need);
    while (m) {
      const int leader = __ffs(m) - 1;
      const u32 lab = __shfl_sync(CC_FULL, cur, leader);
      const bool mine = need && cur == lab;
      const u32 grp = __ballot_sync(CC_FULL, mine);
      m &= ~grp;
      u32 R = 0, SX = 0, SY = 0, YMIN = 0;
      if (mine) {
        const u32 rows = yend - ystart + 1;
        R = __reduce_add_sync(grp, rows);
        SX = __reduce_add_sync(grp, (u32)lane * rows);
        SY = __reduce_add_sync(grp, (ystart + yend) * rows);       // = 2 * sum of (y - y0) over the runs
        YMIN = __reduce_min_sync(grp, ystart);
      }
      // slot of the label in the warp table (lanes 0..7 look at one entry each)
      const u32 k = lane < CC_STAT_WSLOTS ? wkey[lane] : 0xFFFFFFFFu;
      const u32 hit = __ballot_sync(CC_FULL, k == lab + 1);
      int slot;
      if (hit) slot = __ffs(hit) - 1;
      else {
        const u32 empty = __ballot_sync(CC_FULL, k == 0);
        if (empty) slot = __ffs(empty) - 1;
        else {   // full: make room (round robin over the entries by label)
          slot = (int)(lab & (CC_STAT_WSLOTS - 1));
          if (lane == leader) spill(slot, xbase, y0, z);
        }
        __syncwarp();
      }
      if (lane == leader) {
        const u32 xmin = (u32)(__ffs(grp) - 1), xmax = (u32)(31 - __clz(grp));
        if (hit) {
          const u32 bb = wbb[slot];
          wcnt[slot] += R; wsx[slot] += SX; wsy[slot] += SY;
          wbb[slot] = min(bb & 0xFFu, xmin) | (max((bb >> 8) & 0xFFu, xmax) << 8) | (min((bb >> 16) & 0xFFu, YMIN) << 16) |
                      (max(bb >> 24, yend) << 24);
        } else {
          wkey[slot] = lab + 1; wcnt[slot] = R; wsx[slot] = SX; wsy[slot] = SY;
          wbb[slot] = xmin | (xmax << 8) | (YMIN << 16) | (yend << 24);
        }
      }
      __syncwarp();
    }
  };

  // every CTA walks a contiguous range of tasks (a compact region of the volume: few labels in its table)
  const i64 per_cta = (ntasks + gridDim.x - 1) / gridDim.x;
  const i64 task_end = min(ntasks, (i64)(blockIdx.x + 1) * per_cta);
  LT vmax = (LT)0;     // largest label this lane has seen (the caller may not know the maximum: statistics_auto)
  for (i64 task = (i64)blockIdx.x * per_cta + warp; task < task_end; task += blockDim.x >> 5) {
    const i64 w = task % W, t = task / W;
    const u32 ych = (u32)(t % nych), z = (u32)(t / nych);
    const u32 xbase = (u32)(w * 32);
    const u32 x = xbase + lane;
    const bool inx = x < sx;
    const u32 y0 = ych * CC_STAT_YCH;
    const u32 nrow = min(sy, y0 + CC_STAT_YCH) - y0;
    const LT* __restrict__ p = labels + (((size_t)z * sy + y0) * sx + (inx ? x : 0));
    u32 cur = NONE, ystart = 0;
    auto row_step = [&](LT v, u32 r) {
      if (inx && v > vmax) vmax = v;
      const u32 l = (inx && v <= nmax) ? (u32)v : NONE;
      const bool change = l != cur;
      if (__any_sync(CC_FULL, change)) {
        finish_runs(change && cur != NONE, cur, ystart, r - 1, xbase, y0, z);
        if (change) { cur = l; ystart = r; }
      }
    };
    u32 r = 0;
    for (; r + CC_STAT_UNR <= nrow; r += CC_STAT_UNR) {   // full groups of rows: loads issued back to back
      LT v[CC_STAT_UNR];
#pragma unroll
      for (int k = 0; k < CC_STAT_UNR; k++) v[k] = p[(size_t)k * sx];
      p += (size_t)CC_STAT_UNR * sx;
#pragma unroll
      for (int k = 0; k < CC_STAT_UNR; k++) row_step(v[k], r + k);
    }
    for (; r < nrow; r++) { row_step(*p, r); p += sx; }
    finish_runs(cur != NONE, cur, ystart, nrow - 1, xbase, y0, z);
    if (lane < CC_STAT_WSLOTS) spill(lane, xbase, y0, z);
    __syncwarp();
  }
  if (maxout) {
    unsigned long long m = (unsigned long long)vmax;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const unsigned long long t = __shfl_xor_sync(CC_FULL, m, o); if (t > m) m = t; }
    if (lane == 0 && m > *(volatile unsigned long long*)maxout) atomicMax(maxout, m);
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
#pragma unroll
    for (int k = 0; k < 3; k++) atomicAdd(&s[k], ((unsigned long long)tb.sumhi[i][k] << 32) | tb.sumlo[i][k]);
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

