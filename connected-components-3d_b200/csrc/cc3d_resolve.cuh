// cc3d_resolve.cuh — run-table kernels: S (scan of run starts), C1 (compress), C2 (scan of roots),
// C3 (assign), D (expand), the block-order renumbering used by the binary 2D 8-connected path, and the
// face-pair kernels of the sharded path. See cc3d_common.cuh.
#pragma once
#include "cc3d_common.cuh"

#ifndef CC_GRID_BLOCKS
#define CC_GRID_BLOCKS (148 * 8)   // grid of the kernels that loop over a device-side count (148 x 16 / x 32 measured flat)
#endif

// Length of a device-sized array: n_dev == nullptr -> n_host, else ceil(*n_dev / 2^shift)
__device__ __forceinline__ u32 dev_len(i64 n_host, const u64* __restrict__ n_dev, int shift) {
  return n_dev ? (u32)((*n_dev + ((1ull << shift) - 1)) >> shift) : (u32)n_host;
}

// C1: one lane per run. Every run is pointed straight at its root; 32 root flags per word (ballot) and
// the word's popcount are emitted for the scan.
__global__ void __launch_bounds__(256)
k_compress(u32* __restrict__ L, u32* __restrict__ GR, u32* __restrict__ cnt, const u64* __restrict__ n_dev) {
  CC_PDL_WAIT();
  const u32 n = (u32)*n_dev;
  const u32 nwords2 = (n + 31) >> 5;
  const int lane = threadIdx.x & 31;
  const u32 nwarps = (gridDim.x * blockDim.x) >> 5;
  for (u32 wd = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; wd < nwords2; wd += nwarps) {
    const u32 i = (wd << 5) + lane;
    bool isroot = false;
    if (i < n) {
      u32 r = i, p;
      // L1-cached loads: tile roots are shared by many runs; a stale parent is still an ancestor
      while ((p = __ldca(&L[r])) != r) r = p;
      if (r == i) isroot = true;
      else L[i] = r;
    }
    const u32 m = __ballot_sync(CC_FULL, isroot);
    if (lane == 0) { GR[wd] = m; cnt[wd] = __popc(m); }
  }
}

// C1, four runs per lane (round 2d). The sweep is bound by latency x occupancy (one dependent chain per thread: ~1 us per
// iteration with 2 048 resident threads per SM), so a lane takes runs i, i+32, i+64, i+96 of a 128-run group and walks
// the four chains together: four loads in flight per lane, coalesced per k, one ballot per flag word.
__global__ void __launch_bounds__(256)
k_compress4(u32* __restrict__ L, u32* __restrict__ GR, u32* __restrict__ cnt, const u64* __restrict__ n_dev) {
  CC_PDL_WAIT();
  const u32 n = (u32)*n_dev;
  const u32 nwords2 = (n + 31) >> 5;
  const u32 ngroups = (n + 127) >> 7;
  const int lane = threadIdx.x & 31;
  const u32 nwarps = (gridDim.x * blockDim.x) >> 5;
  for (u32 grp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; grp < ngroups; grp += nwarps) {
    const u32 base = (grp << 7) + lane;
    u32 cur[4], nxt[4], first[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const u32 i = base + 32u * k;
      cur[k] = i;
      nxt[k] = i < n ? __ldca(&L[i]) : i;
      first[k] = nxt[k];
    }
    while (true) {
      bool mv[4], any = false;
#pragma unroll
      for (int k = 0; k < 4; k++) { mv[k] = nxt[k] != cur[k]; any |= mv[k]; }
      if (!any) break;
      // L1-cached loads: tile roots are shared by many runs; a stale parent is still an ancestor
#pragma unroll
      for (int k = 0; k < 4; k++) if (mv[k]) { cur[k] = nxt[k]; nxt[k] = __ldca(&L[cur[k]]); }
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const u32 i = base + 32u * k;
      const bool isroot = i < n && cur[k] == i;
      if (i < n && cur[k] != first[k]) L[i] = cur[k];
      const u32 m = __ballot_sync(CC_FULL, isroot);
      const u32 wd = (grp << 2) + k;
      if (lane == 0 && wd < nwords2) { GR[wd] = m; cnt[wd] = __popc(m); }
    }
  }
}

// C2 / S: exclusive scan of u32 counts, three small kernels (block reduce, scan of block sums, apply).
// The length is either a host value or derived from a device-side count (see dev_len); the kernels
// loop over 4096-element chunks, so the grid does not depend on the length.
// track != nullptr (scan S): also reduces the first / last row that has a run start (= any foreground)
// into track->first_row / last_row; W = bitmap words per row.
__global__ void __launch_bounds__(CC_SCAN_THREADS)
k_scan_reduce(const u32* __restrict__ cnt, u64* __restrict__ bsum, i64 n_host, const u64* __restrict__ n_dev, int shift,
              Counters* __restrict__ track, u32 W) {
  const u32 n = dev_len(n_host, n_dev, shift);
  const u32 nb = (n + CC_SCAN_CHUNK - 1) / CC_SCAN_CHUNK;
  u32 imin = 0xFFFFFFFFu, imax = 0;
  bool any = false;
  for (u32 blk = blockIdx.x; blk < nb; blk += gridDim.x) {
    const u32 base = blk * CC_SCAN_CHUNK;
    u32 s = 0;
#pragma unroll
    for (int k = 0; k < CC_SCAN_ITEMS; k++) {
      const u32 i = base + k * CC_SCAN_THREADS + threadIdx.x;
      if (i < n) {
        const u32 v = cnt[i];
        s += v;
        if (track && v) { imin = min(imin, i); imax = i; any = true; }
      }
    }
    u32 tot;
    block_exclusive_scan(s, &tot);
    if (threadIdx.x == 0) bsum[blk] = tot;
  }
  if (track) {   // block-uniform; one pair of global atomics per block
    __shared__ u32 s_min, s_max;
    if (threadIdx.x == 0) { s_min = 0xFFFFFFFFu; s_max = 0; }
    __syncthreads();
    if (__ballot_sync(CC_FULL, any)) {
      imin = __reduce_min_sync(CC_FULL, imin);
      imax = __reduce_max_sync(CC_FULL, imax);
      if ((threadIdx.x & 31) == 0) { atomicMin(&s_min, imin); atomicMax(&s_max, imax); }
    }
    __syncthreads();
    if (threadIdx.x == 0 && s_min != 0xFFFFFFFFu) {
      track_rows(track, (u64)(s_min / W), (u64)(s_max / W));
    }
  }
}

// single block: exclusive scan of the block sums (64-bit), total -> *N (and *N32 when given)
__global__ void __launch_bounds__(1024)
k_scan_blocks(u64* __restrict__ bsum, i64 n_host, const u64* __restrict__ n_dev, int shift, u64* __restrict__ N,
              u32* __restrict__ N32) {
  __shared__ u64 carry;
  __shared__ u64 wsum[32];
  const u32 n = dev_len(n_host, n_dev, shift);
  const u32 nb = (n + CC_SCAN_CHUNK - 1) / CC_SCAN_CHUNK;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (u32 base = 0; base < nb; base += 1024) {
    const u32 i = base + threadIdx.x;
    const u64 v = i < nb ? bsum[i] : 0;
    u64 inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const u64 t = __shfl_up_sync(CC_FULL, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      const u64 s = wsum[lane];
      u64 si = s;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const u64 t = __shfl_up_sync(CC_FULL, si, o);
        if (lane >= o) si += t;
      }
      wsum[lane] = si - s;
    }
    __syncthreads();
    const u64 c = carry;
    if (i < nb) bsum[i] = c + wsum[warp] + inc - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry = c + wsum[warp] + inc;
    __syncthreads();
  }
  if (threadIdx.x == 0) { *N = carry; if (N32) *N32 = (u32)carry; }
}

// prefix[i] = exclusive prefix of cnt (mod 2^32 is fine: ranks are < voxels < 2^32); cnt may alias prefix
__global__ void __launch_bounds__(CC_SCAN_THREADS)
k_scan_apply(const u32* cnt, const u64* __restrict__ bsum, u32* prefix, i64 n_host, const u64* __restrict__ n_dev, int shift) {
  const u32 n = dev_len(n_host, n_dev, shift);
  const u32 nb = (n + CC_SCAN_CHUNK - 1) / CC_SCAN_CHUNK;
  for (u32 blk = blockIdx.x; blk < nb; blk += gridDim.x) {
    const u32 base = blk * CC_SCAN_CHUNK + threadIdx.x * CC_SCAN_ITEMS;
    u32 v[CC_SCAN_ITEMS];
    u32 s = 0;
    if (base + CC_SCAN_ITEMS <= n) {
      const uint4* p = reinterpret_cast<const uint4*>(cnt + base);
#pragma unroll
      for (int k = 0; k < CC_SCAN_ITEMS / 4; k++) {
        const uint4 t = p[k];
        v[4 * k] = t.x; v[4 * k + 1] = t.y; v[4 * k + 2] = t.z; v[4 * k + 3] = t.w;
      }
    } else {
#pragma unroll
      for (int k = 0; k < CC_SCAN_ITEMS; k++) v[k] = (base + k < n) ? cnt[base + k] : 0;
    }
#pragma unroll
    for (int k = 0; k < CC_SCAN_ITEMS; k++) s += v[k];
    u32 ex = block_exclusive_scan(s, nullptr) + (u32)bsum[blk];
    if (base + CC_SCAN_ITEMS <= n) {
      uint4* q = reinterpret_cast<uint4*>(prefix + base);
#pragma unroll
      for (int k = 0; k < CC_SCAN_ITEMS / 4; k++) {
        uint4 t;
        t.x = ex; ex += v[4 * k]; t.y = ex; ex += v[4 * k + 1]; t.z = ex; ex += v[4 * k + 2]; t.w = ex; ex += v[4 * k + 3];
        q[k] = t;
      }
    } else {
#pragma unroll
      for (int k = 0; k < CC_SCAN_ITEMS; k++) { if (base + k < n) prefix[base + k] = ex; ex += v[k]; }
    }
  }
}

// One-pass exclusive scan (decoupled look-back): chunks of 4096 counts are handed out in order by a ticket, a
// chunk publishes its aggregate, then its inclusive prefix as soon as the chunks before it are known. status[c] =
// flag << 32 | value (flag 1 = aggregate, 2 = inclusive prefix; values < 2^32), status[nb_max] = ticket counter;
// both must be zero on entry. Replaces the reduce / scan-of-sums / apply triple (one launch instead of three).
// track != nullptr (scan S): also reduces the first / last row that has a run start into track->first_row/last_row.
#define CC_SCAN_FLAG_A (1ull << 32)
#define CC_SCAN_FLAG_P (2ull << 32)
__global__ void __launch_bounds__(CC_SCAN_THREADS)
k_scan_onepass(const u32* cnt, u32* prefix, unsigned long long* __restrict__ status, u32 nb_max, i64 n_host,
               const u64* __restrict__ n_dev, int shift, u64* __restrict__ total, u32* __restrict__ total32,
               Counters* __restrict__ track, u32 W) {
  CC_PDL_WAIT();
  __shared__ u32 s_chunk, s_prev, s_min, s_max;
  const u32 n = dev_len(n_host, n_dev, shift);
  const u32 nb = (n + CC_SCAN_CHUNK - 1) / CC_SCAN_CHUNK;
  if (threadIdx.x == 0) { s_chunk = (u32)atomicAdd(&status[nb_max], 1ull); s_min = 0xFFFFFFFFu; s_max = 0; }
  __syncthreads();
  const u32 chunk = s_chunk;
  if (chunk >= nb) {
    if (nb == 0 && chunk == 0 && threadIdx.x == 0) { *total = 0; if (total32) *total32 = 0; }
    return;
  }
  const u32 base = chunk * CC_SCAN_CHUNK + threadIdx.x * CC_SCAN_ITEMS;
  u32 v[CC_SCAN_ITEMS];
  if (base + CC_SCAN_ITEMS <= n) {
    const uint4* p = reinterpret_cast<const uint4*>(cnt + base);
#pragma unroll
    for (int k = 0; k < CC_SCAN_ITEMS / 4; k++) {
      const uint4 t = p[k];
      v[4 * k] = t.x; v[4 * k + 1] = t.y; v[4 * k + 2] = t.z; v[4 * k + 3] = t.w;
    }
  } else {
#pragma unroll
    for (int k = 0; k < CC_SCAN_ITEMS; k++) v[k] = (base + k < n) ? cnt[base + k] : 0;
  }
  u32 sum = 0;
#pragma unroll
  for (int k = 0; k < CC_SCAN_ITEMS; k++) sum += v[k];
  if (track && sum) {
    u32 lo = CC_SCAN_ITEMS, hi = 0;
#pragma unroll
    for (int k = 0; k < CC_SCAN_ITEMS; k++) if (v[k]) { lo = min(lo, (u32)k); hi = k; }
    atomicMin(&s_min, base + lo); atomicMax(&s_max, base + hi);
  }
  u32 tot;
  u32 ex = block_exclusive_scan(sum, &tot);
  if (threadIdx.x < 32) {   // warp 0: publish the aggregate, look back, publish the inclusive prefix
    const int lane = threadIdx.x;
    u32 prev = 0;
    if (chunk > 0) {
      if (lane == 0) { __threadfence(); atomicExch(&status[chunk], CC_SCAN_FLAG_A | tot); }
      i64 idx = (i64)chunk - 1;
      while (true) {
        const i64 j = idx - lane;
        unsigned long long st = CC_SCAN_FLAG_P;   // before chunk 0: prefix 0
        if (j >= 0) { do { st = *(volatile unsigned long long*)&status[j]; } while ((st >> 32) == 0); }
        const u32 isp = __ballot_sync(CC_FULL, (st >> 32) == 2);
        const int stop = isp ? (__ffs(isp) - 1) : 32;            // nearest predecessor with a full prefix
        const u32 part = lane <= stop ? (u32)st : 0u;
        prev += __reduce_add_sync(CC_FULL, part);
        if (isp) break;
        idx -= 32;
      }
    }
    if (lane == 0) {
      __threadfence();
      atomicExch(&status[chunk], CC_SCAN_FLAG_P | (unsigned long long)(prev + tot));
      s_prev = prev;
      if (chunk == nb - 1) { *total = (u64)prev + tot; if (total32) *total32 = prev + tot; }
    }
  }
  __syncthreads();
  ex += s_prev;
  if (base + CC_SCAN_ITEMS <= n) {
    uint4* q = reinterpret_cast<uint4*>(prefix + base);
#pragma unroll
    for (int k = 0; k < CC_SCAN_ITEMS / 4; k++) {
      uint4 t;
      t.x = ex; ex += v[4 * k]; t.y = ex; ex += v[4 * k + 1]; t.z = ex; ex += v[4 * k + 2]; t.w = ex; ex += v[4 * k + 3];
      q[k] = t;
    }
  } else {
#pragma unroll
    for (int k = 0; k < CC_SCAN_ITEMS; k++) { if (base + k < n) prefix[base + k] = ex; ex += v[k]; }
  }
  if (track && threadIdx.x == 0 && s_min != 0xFFFFFFFFu) {
    track_rows(track, (u64)(s_min / W), (u64)(s_max / W));
  }
}


// ---------------------------------------------------------------------------------------------
// C (fused): compress + scan of root flags + assign in ONE launch (replaces k_compress / k_scan_onepass / k_assign on
// the label path). Chunks of CC_RANK_RUNS consecutive runs are handed out in order by a ticket (persistent CTAs):
//   1. every run of the chunk is chased to its root (L1-cached loads; a stale parent is still an ancestor); root flags
//      are balloted, 32 per word;
//   2. the words' popcounts are scanned in the block; the chunk publishes its flags (GR), the chunk-relative prefix of
//      every flag word (LP) and its aggregate, then looks back for its exclusive prefix (decoupled look-back,
//      status[c] = flag << 32 | value as in k_scan_onepass) and publishes the inclusive one;
//   3. every run gets the final label of its root, written in place as CC_LABEL_TAG | label. A root in the same chunk is
//      ranked from shared memory; a root in an earlier chunk c' is either already tagged (its label is read directly)
//      or ranked as inclusive(c'-1) + LP[word] + popc(GR[word] & bits below) - chunk c' has published both before its
//      aggregate, and every chunk below ours has published its aggregate before our look-back could finish.
// A chase that meets a tagged entry stops there: the tag carries the final label of the whole component. That is what
// makes the in-place update safe while later chunks are still chasing through this chunk's runs.
// Needs run ids and labels below 2^31 (the host falls back to the three-kernel path otherwise). Readers of L mask the
// tag off (`lmask`). *total = number of components.
// ---------------------------------------------------------------------------------------------
#ifndef CC_RANK_WORDS
#define CC_RANK_WORDS 32          // flag words (of 32 runs) per chunk: small chunks = short serial chains per CTA, many CTAs
#endif
#define CC_RANK_RUNS (CC_RANK_WORDS * 32)
#define CC_LABEL_TAG 0x80000000u

__global__ void __launch_bounds__(256)
k_rank(u32* __restrict__ L, u32* __restrict__ GR, u32* __restrict__ LP, unsigned long long* __restrict__ status, u32 nb_max,
       const u64* __restrict__ n_dev, u64* __restrict__ total) {
  CC_PDL_WAIT();
  __shared__ u32 s_res[CC_RANK_RUNS];
  __shared__ u32 s_mask[CC_RANK_WORDS], s_lp[CC_RANK_WORDS];
  __shared__ u32 s_chunk, s_prev;
  const u32 n = (u32)*n_dev;
  const u32 nb = (n + CC_RANK_RUNS - 1) / CC_RANK_RUNS;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  while (true) {
    __syncthreads();
    if (threadIdx.x == 0) s_chunk = (u32)atomicAdd(&status[nb_max], 1ull);
    __syncthreads();
    const u32 chunk = s_chunk;
    if (chunk >= nb) {
      if (nb == 0 && chunk == 0 && threadIdx.x == 0) *total = 0;
      return;
    }
    const u32 base = chunk * CC_RANK_RUNS;
    // ---- 1. roots ----
#pragma unroll 4
    for (u32 j = warp; j < CC_RANK_WORDS; j += 8) {
      const u32 li = (j << 5) + lane;
      const u32 i = base + li;
      u32 res = CC_LABEL_TAG;       // beyond the last run: label 0, never read
      bool isroot = false;
      if (i < n) {
        u32 r = i;
        while (true) {
          const u32 p = __ldca(&L[r]);
          if (p & CC_LABEL_TAG) { res = p; break; }
          if (p == r) { res = r; isroot = (r == i); break; }
          r = p;
        }
      }
      s_res[li] = res;
      const u32 m = __ballot_sync(CC_FULL, isroot);
      if (lane == 0) s_mask[j] = m;
    }
    __syncthreads();
    // ---- 2. scan ----
    u32 tot;
    const u32 mymask = threadIdx.x < CC_RANK_WORDS ? s_mask[threadIdx.x] : 0u;
    const u32 lp = block_exclusive_scan((u32)__popc(mymask), &tot);
    if (threadIdx.x < CC_RANK_WORDS) {
      s_lp[threadIdx.x] = lp;
      const u32 gw = (base >> 5) + threadIdx.x;
      if ((gw << 5) < n) { GR[gw] = mymask; LP[gw] = lp; }
    }
    __syncthreads();
    if (warp == 0) {
      u32 prev = 0;
      if (chunk > 0) {
        if (lane == 0) { __threadfence(); atomicExch(&status[chunk], CC_SCAN_FLAG_A | tot); }
        i64 idx = (i64)chunk - 1;
        while (true) {
          const i64 j = idx - lane;
          unsigned long long st = CC_SCAN_FLAG_P;   // before chunk 0: prefix 0
          if (j >= 0) { do { st = *(volatile unsigned long long*)&status[j]; } while ((st >> 32) == 0); }
          const u32 isp = __ballot_sync(CC_FULL, (st >> 32) == 2);
          const int stop = isp ? (__ffs(isp) - 1) : 32;
          const u32 part = lane <= stop ? (u32)st : 0u;
          prev += __reduce_add_sync(CC_FULL, part);
          if (isp) break;
          idx -= 32;
        }
      }
      if (lane == 0) {
        __threadfence();      // release our GR / LP, acquire those of the chunks below (their flags were observed above)
        atomicExch(&status[chunk], CC_SCAN_FLAG_P | (unsigned long long)(prev + tot));
        s_prev = prev;
        if (chunk == nb - 1) *total = (u64)prev + tot;
      }
    }
    __syncthreads();
    const u32 prev = s_prev;
    // ---- 3. labels ----
#pragma unroll 4
    for (u32 j = warp; j < CC_RANK_WORDS; j += 8) {
      const u32 li = (j << 5) + lane;
      const u32 i = base + li;
      if (i >= n) continue;
      u32 lab = s_res[li];
      if (!(lab & CC_LABEL_TAG)) {
        const u32 r = lab;
        const u32 below = (1u << (r & 31)) - 1u;
        if (r >= base) {
          const u32 lw = (r - base) >> 5;
          lab = prev + s_lp[lw] + __popc(s_mask[lw] & below) + 1u;
        } else {
          const u32 q = __ldcg(&L[r]);
          if (q & CC_LABEL_TAG) lab = q;
          else {
            const u32 cr = r / CC_RANK_RUNS;
            u32 pb = 0;
            if (cr > 0) {
              unsigned long long st;
              do { st = *(volatile unsigned long long*)&status[cr - 1]; } while ((st >> 32) != 2);
              pb = (u32)st;
            }
            lab = pb + __ldcg(&LP[r >> 5]) + __popc(__ldcg(&GR[r >> 5]) & below) + 1u;
          }
        }
        lab |= CC_LABEL_TAG;
      }
      L[i] = lab;
    }
  }
}

__device__ __forceinline__ u32 rank_in_bitmap(const u32* __restrict__ bm, const u32* __restrict__ prefix, i64 word, int bit) {
  return prefix[word] + __popc(bm[word] & ((1u << bit) - 1u));
}

// C3: every run gets its component's final label (1-based rank of its root).
__global__ void __launch_bounds__(256)
k_assign(u32* __restrict__ L, const u32* __restrict__ GR, const u32* __restrict__ prefix, const u64* __restrict__ n_dev) {
  CC_PDL_WAIT();
  const u32 n = (u32)*n_dev;
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const bool isroot = (GR[i >> 5] >> (i & 31)) & 1u;
    const u32 r = isroot ? i : L[i];
    L[i] = rank_in_bitmap(GR, prefix, r >> 5, (int)(r & 31)) + 1u;
  }
}

// C3, four runs per lane (see k_compress4): four independent gather chains per thread.
__global__ void __launch_bounds__(256)
k_assign4(u32* __restrict__ L, const u32* __restrict__ GR, const u32* __restrict__ prefix, const u64* __restrict__ n_dev) {
  CC_PDL_WAIT();
  const u32 n = (u32)*n_dev;
  const u32 ngroups = (n + 127) >> 7;
  const int lane = threadIdx.x & 31;
  const u32 nwarps = (gridDim.x * blockDim.x) >> 5;
  for (u32 grp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; grp < ngroups; grp += nwarps) {
    const u32 base = (grp << 7) + lane;
    u32 r[4], fl[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const u32 i = base + 32u * k;
      r[k] = i < n ? L[i] : 0u;
      fl[k] = (grp << 2) + k < ((n + 31) >> 5) ? __ldg(&GR[(grp << 2) + k]) : 0u;
    }
    u32 pw[4], gw[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const u32 i = base + 32u * k;
      if ((fl[k] >> lane) & 1u) r[k] = i;
      const bool ok = i < n;
      pw[k] = ok ? __ldg(&prefix[r[k] >> 5]) : 0u;
      gw[k] = ok ? __ldg(&GR[r[k] >> 5]) : 0u;
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const u32 i = base + 32u * k;
      if (i < n) L[i] = pw[k] + __popc(gw[k] & ((1u << (r[k] & 31)) - 1u)) + 1u;
    }
  }
}

// D: the only dense write. One warp per (row, 32-word chunk): lane j first takes word j of the F / X /
// RS bitmaps; then the warp walks the words, one voxel per lane: the voxel's run is the last run that
// started at or before it (popcount), one mostly warp-uniform L2-resident load of the run's label, one
// coalesced store.
// REMAP: 0 = write the label, 1/2 = write remap[label] from a u32/u64 table (sharded volumes:
// per-slab label -> global label). row0 = first row to write; out is indexed relative to row0.
template <typename OUT, int REMAP>
__global__ void __launch_bounds__(256)
k_expand(const u32* __restrict__ L, const u32* __restrict__ M, OUT* __restrict__ out, Geom g, unsigned nchunks,
         u32 row0, u32 nwarps_total, const void* __restrict__ remap, u32 lmask) {
  __shared__ uint4 s_words[8][32];   // per warp: {F, run starts, id of the run that enters the word, -}
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const u32 wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (wid >= nwarps_total) return;
  const u32 rrel = wid / nchunks;
  const u32 chunk = wid - rrel * nchunks;
  const u32 row = row0 + rrel;
  const u32 W = (u32)g.W, sx = (u32)g.sx;
  const u32 wl = (chunk << 5) + lane;
  {
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (wl < W) {
      const u32 j = row * W + wl;
      const uint2 fx = __ldg(reinterpret_cast<const uint2*>(M) + 2 * (size_t)j);
      v.x = fx.x;
      v.y = fx.x & ~fx.y;
      v.z = __ldg(M + g.offRS + j) - 1u;
    }
    s_words[warp][lane] = v;
  }
  __syncwarp();
  const u32 nwd = min(32u, W - (chunk << 5));
  const u32 below = CC_FULL >> (31 - lane);
  const u32 bit = 1u << lane;
  u32 x = (chunk << 10) + lane;
  OUT* __restrict__ o = out + ((size_t)rrel * sx + x);
  auto label_of = [&](const uint4 wv) -> OUT {
    OUT v = 0;
    if (wv.x & bit) {
      const u32 lab = L[wv.z + __popc(wv.y & below)] & lmask;
      if (REMAP == 1) v = (OUT)__ldg(reinterpret_cast<const u32*>(remap) + lab);
      else if (REMAP == 2) v = (OUT)__ldg(reinterpret_cast<const u64*>(remap) + lab);
      else v = (OUT)lab;
    }
    return v;
  };
  const u32 nfull = (x - lane + (nwd << 5) <= sx) ? nwd : nwd - 1;   // words that lie fully inside the row
#pragma unroll 4
  for (u32 j = 0; j < nfull; j++) o[j << 5] = label_of(s_words[warp][j]);
  if (nfull < nwd) {
    const OUT v = label_of(s_words[warp][nfull]);
    if (x + (nfull << 5) < sx) o[nfull << 5] = v;
  }
}

// ---- fused dust (cc3d/__init__.py:71-155): component sizes straight from the run table, then the image is
// masked while it is expanded - the label volume is never written or read. ----
#define CC_CNT_SLOTS 2048
// One thread per bitmap word: every piece of a run inside the word adds its length to its component's count
// (per-CTA shared-memory hash table first, one global atomic per label and CTA).
__global__ void __launch_bounds__(256)
k_run_counts(const u32* __restrict__ L, const u32* __restrict__ M, Geom g, u32* __restrict__ counts, u32 lmask) {
  __shared__ u32 s_key[CC_CNT_SLOTS], s_cnt[CC_CNT_SLOTS];
  for (int i = threadIdx.x; i < CC_CNT_SLOTS; i += blockDim.x) { s_key[i] = 0; s_cnt[i] = 0; }
  __syncthreads();
  const u32 nwords = (u32)g.nwords;
  const u32 per_cta = (nwords + gridDim.x - 1) / gridDim.x;
  const u32 end = min(nwords, (blockIdx.x + 1) * per_cta);
  for (u32 j = blockIdx.x * per_cta + threadIdx.x; j < end; j += blockDim.x) {
    const uint2 fx = __ldg(reinterpret_cast<const uint2*>(M) + 2 * (size_t)j);
    const u32 F = fx.x;
    if (!F) continue;
    const u32 runstarts = F & ~fx.y;
    u32 starts = F & (~fx.y | 1u);            // pieces: a run start, or bit 0 of a run that enters the word
    const u32 stops = ~F | starts;            // a piece ends before the next piece or the next background voxel
    const u32 rid0 = __ldg(M + g.offRS + j) - 1u;
    while (starts) {
      const int b = __ffs(starts) - 1; starts &= starts - 1;
      const u32 rest = b == 31 ? 0u : (stops >> (b + 1));
      const u32 len = rest ? (u32)__ffs(rest) : (u32)(32 - b);
      const u32 lab = L[rid0 + __popc(runstarts & (CC_FULL >> (31 - b)))] & lmask;
      u32 h = (lab * 2654435761u) >> 21;   // 11 bits
      bool done = false;
#pragma unroll 1
      for (int probe = 0; probe < 8 && !done; probe++) {
        const u32 s = (h + probe) & (CC_CNT_SLOTS - 1);
        u32 k = *(volatile u32*)&s_key[s];
        if (k == 0) k = atomicCAS(&s_key[s], 0u, lab + 1);
        if (k == 0 || k == lab + 1) { atomicAdd(&s_cnt[s], len); done = true; }
      }
      if (!done) atomicAdd(&counts[lab], len);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < CC_CNT_SLOTS; i += blockDim.x)
    if (s_key[i]) atomicAdd(&counts[s_key[i] - 1], s_cnt[i]);
}

// keep[l] for l in 0..N: components with lo <= size < hi stay (invert: the others stay); *n_masked = number of
// components outside [lo, hi)
__global__ void __launch_bounds__(256)
k_dust_keep(const u32* __restrict__ counts, unsigned char* __restrict__ keep, u64 N, long long lo, long long hi, int invert,
            unsigned long long* __restrict__ n_masked) {
  const u64 l = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  bool masked = false;
  if (l <= N) {
    if (l == 0) keep[0] = invert ? 0 : 1;
    else {
      const long long c = (long long)counts[l];
      masked = !(lo <= c && c < hi);
      keep[l] = (masked != (invert != 0)) ? 0 : 1;
    }
  }
  const u32 m = __ballot_sync(CC_FULL, masked);
  if (m && (threadIdx.x & 31) == 0) atomicAdd(n_masked, (unsigned long long)__popc(m));
}

// out[v] = keep[label of v] ? img[v] : 0 (same walk as k_expand; out may alias img)
template <typename IT>
__global__ void __launch_bounds__(256)
k_expand_mask(const u32* __restrict__ L, const u32* __restrict__ M, const IT* img, IT* out, Geom g, unsigned nchunks,
              u32 nwarps_total, const unsigned char* __restrict__ keep, u32 lmask) {
  __shared__ uint4 s_words[8][32];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const u32 wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (wid >= nwarps_total) return;
  const u32 row = wid / nchunks;
  const u32 chunk = wid - row * nchunks;
  const u32 W = (u32)g.W, sx = (u32)g.sx;
  const u32 wl = (chunk << 5) + lane;
  {
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (wl < W) {
      const u32 j = row * W + wl;
      const uint2 fx = __ldg(reinterpret_cast<const uint2*>(M) + 2 * (size_t)j);
      v.x = fx.x;
      v.y = fx.x & ~fx.y;
      v.z = __ldg(M + g.offRS + j) - 1u;
    }
    s_words[warp][lane] = v;
  }
  __syncwarp();
  const u32 nwd = min(32u, W - (chunk << 5));
  const u32 below = CC_FULL >> (31 - lane);
  const u32 bit = 1u << lane;
  const u32 x = (chunk << 10) + lane;
  const size_t base = (size_t)row * sx + x;
  for (u32 j = 0; j < nwd; j++) {
    if (x + (j << 5) >= sx) break;
    const uint4 wv = s_words[warp][j];
    const IT v = img[base + (j << 5)];
    bool k = false;
    if (wv.x & bit) k = keep[L[wv.z + __popc(wv.y & below)] & lmask] != 0;
    out[base + (j << 5)] = k ? v : (IT)0;
  }
}

// ---- sharded volumes (z-slabs): equivalences across one slab interface ----
// P = first plane of the upper slab (later in raster order), Q = last plane of the lower slab.
// Emits (label in Q's slab, label in P's slab) for every edge of the chosen predicate/neighbourhood
// between the two planes, skipping an edge when the voxel to the left or above already produced the same pair.
// Replaces the Python face loops of connected_components_stack (cc3d/__init__.py:425-468).
template <typename T, int MODE>
__global__ void __launch_bounds__(256)
k_face_pairs(const T* __restrict__ vP, const u32* __restrict__ lP, const T* __restrict__ vQ, const u32* __restrict__ lQ,
             i64 sx, i64 sy, int connectivity, Edge<T, MODE> E, u64* __restrict__ pairs, unsigned long long cap,
             unsigned long long* __restrict__ count) {
  // per-CTA direct-mapped cache of the pairs already emitted: an interface between two big components reports the same
  // pair from thousands of voxels (a thread that finds its pair here skips it - the thread that stored it appends it)
  __shared__ u64 s_seen[256];
  s_seen[threadIdx.x] = ~0ull;
  __syncthreads();
  const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  u64 mine[9];
  int n = 0;
  if (i < sx * sy) {
    const i64 y = i / sx, x = i - y * sx;
    const T p = vP[i];
    if (E.fg(p)) {
      const u32 lp = lP[i];
      const bool left_same = x > 0 && lP[i - 1] == lp;
      const bool up_same = y > 0 && lP[i - sx] == lp;
      // transitive predicates: when the straight neighbour joins p, every other matching q is its in-plane
      // neighbour and already belongs to the same component of the lower slab
      const bool straight = (MODE != MODE_DELTA) && E(p, vQ[i]);
      const bool zeq_join = (MODE == MODE_DELTA) && connectivity == 26 && E.zedge(p, vQ[i]) && !E(p, vQ[i]);   // +-inf == +-inf below
#pragma unroll
      for (int dy = -1; dy <= 1; dy++) {
#pragma unroll
        for (int dx = -1; dx <= 1; dx++) {
          const int nz = (dx != 0) + (dy != 0);
          if (connectivity == 6 && nz > 0) continue;
          if (connectivity == 18 && nz > 1) continue;
          if (straight && nz > 0) continue;
          const i64 xx = x + dx, yy = y + dy;
          if (xx < 0 || xx >= sx || yy < 0 || yy >= sy) continue;
          const i64 qi = yy * sx + xx;
          const T q = vQ[qi];
          if (!E(p, q) && !(zeq_join && nz == 0)) continue;
          const u32 lq = lQ[qi];
          // the voxel to the left emits the same (lq, lp) through the same direction
          if (left_same && xx > 0 && lQ[qi - 1] == lq && E(vP[i - 1], vQ[qi - 1])) continue;
          // ... or the voxel above
          if (up_same && yy > 0 && lQ[qi - sx] == lq && E(vP[i - sx], vQ[qi - sx])) continue;
          const u64 v = ((u64)lq << 32) | lp;
          const u32 slot = (u32)((v * 0x9E3779B97F4A7C15ull) >> 56);
          if (*(volatile u64*)&s_seen[slot] == v) continue;
          s_seen[slot] = v;
          mine[n++] = v;
        }
      }
    }
  }
  // warp-aggregated append
  u32 total = n;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const u32 t = __shfl_up_sync(CC_FULL, total, o);
    if (lane >= o) total += t;
  }
  const u32 wtotal = __shfl_sync(CC_FULL, total, 31);
  unsigned long long base = 0;
  if (lane == 31 && wtotal) base = atomicAdd(count, (unsigned long long)wtotal);
  base = __shfl_sync(CC_FULL, base, 31);
  const unsigned long long off = base + total - n;
  for (int k = 0; k < n; k++)
    if (off + k < cap) pairs[off + k] = mine[k];
}


// ---------------------------------------------------------------------------------------------
// Sharded volumes: the slab merge ON THE DEVICE (replaces the host union-find of the first round and, with it, the
// host synchronisation in the middle of a sharded step; the reference does this in Python, cc3d/__init__.py:296-321,
// 425-492). Input = the all-gathered buffer: row r = [N_r, epl_r, sz_r, n_pairs_r, pairs...] (int64), pairs packed as
// (label in slab r-1) << 32 | (label in slab r). Global id of (slab r, local label l >= 1) = off[r] + l with
// off[r] = N_0 + ... + N_{r-1}: id order = first-appearance order of the whole volume, so with link-to-smaller unions
// the root of a component is its first label, and
//     final label of a root id = id - #(non-root ids below it),
// i.e. its rank among the roots - no per-slab bookkeeping. Every rank runs the same kernels on the same gathered data
// (redundantly, like the redundant host solve before) and writes only its own slab's remap table.
// result[0] = N of the whole volume, result[1] = label capacity exceeded, result[2] = pair capacity exceeded,
// result[3] = total + 1 (ids incl. 0), result[4] = number of non-root ids, result[8 + 4 r + k] = fact k of slab r.
// ---------------------------------------------------------------------------------------------
struct SlabRows { const long long* rows; int world; long long stride; };
#define CC_MERGE_MAX_WORLD 64

__device__ __forceinline__ void slab_offsets(const SlabRows& f, u64* s_off) {   // s_off[world + 1], one CTA
  if (threadIdx.x == 0) {
    u64 acc = 0;
    for (int r = 0; r < f.world; r++) { s_off[r] = acc; acc += (u64)f.rows[(size_t)r * f.stride]; }
    s_off[f.world] = acc;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(256)
k_merge_init(u32* __restrict__ parent, SlabRows f, u64 label_cap, u64 pair_cap, unsigned long long* __restrict__ result) {
  __shared__ u64 s_off[CC_MERGE_MAX_WORLD + 1];
  slab_offsets(f, s_off);
  const u64 total = s_off[f.world];
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    bool pov = false;
    for (int r = 0; r < f.world; r++) pov |= (u64)f.rows[(size_t)r * f.stride + 3] > pair_cap;
    result[0] = 0; result[1] = (total + 1 > label_cap) ? 1ull : 0ull; result[2] = pov ? 1ull : 0ull;
    result[3] = (total + 1 > label_cap) ? 0ull : total + 1; result[4] = 0; result[5] = 0;
    // the slabs' facts [N, epl, sz, n_pairs] ride along (result[8 + 4 r + k]): the host reads everything with ONE copy
    for (int r = 0; r < f.world; r++)
      for (int k = 0; k < 4; k++) result[8 + 4 * r + k] = (unsigned long long)f.rows[(size_t)r * f.stride + k];
  }
  if (total + 1 > label_cap) return;
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i <= total; i += (u64)gridDim.x * blockDim.x) parent[i] = (u32)i;
}

__global__ void __launch_bounds__(256)
k_merge_union(u32* __restrict__ parent, SlabRows f, u64 label_cap, u64 pair_cap) {
  __shared__ u64 s_off[CC_MERGE_MAX_WORLD + 1];
  slab_offsets(f, s_off);
  if (s_off[f.world] + 1 > label_cap) return;
  const u64 nslots = (u64)f.world * pair_cap;
  for (u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x; t < nslots; t += (u64)gridDim.x * blockDim.x) {
    const int r = (int)(t / pair_cap);
    const u64 k = t - (u64)r * pair_cap;
    if (r == 0) continue;
    const long long* row = f.rows + (size_t)r * f.stride;
    if (k >= (u64)row[3]) continue;
    const u64 v = (u64)row[4 + k];
    const u64 lo = v >> 32, up = v & 0xFFFFFFFFull;
    if (lo < 1 || up < 1 || lo > (u64)f.rows[(size_t)(r - 1) * f.stride] || up > (u64)row[0]) continue;   // malformed pair
    uf_union(parent, (u32)(s_off[r - 1] + lo), (u32)(s_off[r] + up));
  }
}

// one warp per 32 ids: every id is pointed at its root, non-root flags are balloted into NR, popcounts into cnt
__global__ void __launch_bounds__(256)
k_merge_flags(u32* __restrict__ parent, u32* __restrict__ NR, u32* __restrict__ cnt, const unsigned long long* __restrict__ result) {
  const u64 n = result[3];             // total + 1 (0 when the label capacity was exceeded)
  const u64 nwords = (n + 31) >> 5;
  const int lane = threadIdx.x & 31;
  const u64 nwarps = ((u64)gridDim.x * blockDim.x) >> 5;
  for (u64 wd = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5; wd < nwords; wd += nwarps) {
    const u64 i = (wd << 5) + lane;
    bool nonroot = false;
    if (i >= 1 && i < n) {
      u32 r = (u32)i, p;
      while ((p = __ldcg(&parent[r])) != r) r = p;
      if (r != (u32)i) { parent[i] = r; nonroot = true; }
    }
    const u32 m = __ballot_sync(CC_FULL, nonroot);
    if (lane == 0) { NR[wd] = m; cnt[wd] = __popc(m); }
  }
}

// remap[l] = final label of local label l of slab `rank` (remap[0] = 0); result[0] = N of the whole volume
__global__ void __launch_bounds__(256)
k_merge_remap(const u32* __restrict__ parent, const u32* __restrict__ NR, const u32* __restrict__ prefix, SlabRows f, int rank,
              u32* __restrict__ remap, unsigned long long* __restrict__ result) {
  __shared__ u64 s_off[CC_MERGE_MAX_WORLD + 1];
  slab_offsets(f, s_off);
  const u64 n = result[3];
  if (n == 0) return;
  if (blockIdx.x == 0 && threadIdx.x == 0) result[0] = (n - 1) - result[4];
  const u64 off = s_off[rank], nl = (u64)f.rows[(size_t)rank * f.stride];
  for (u64 l = (u64)blockIdx.x * blockDim.x + threadIdx.x; l <= nl; l += (u64)gridDim.x * blockDim.x) {
    u32 out = 0;
    if (l) {
      u32 root = parent[off + l];
      root = parent[root];     // k_merge_flags compresses concurrently: at most one more hop to the root
      const u32 below = __ldg(&prefix[root >> 5]) + __popc(__ldg(&NR[root >> 5]) & ((1u << (root & 31)) - 1u));
      out = root - below;
    }
    remap[l] = out;
  }
}

// The same merge in ONE launch of ONE CTA for small interface graphs (at most CC_MERGE_SMALL slab-label ids: the
// connectomics benchmark has 29 k over eight slabs): the four kernels + scan above are latency, not work (30 us of a
// 0.66 ms sharded step). Phases separated by block barriers; the non-root flags and their prefix live in shared memory.
// result[5] = 1: more ids than this kernel handles - nothing else was written, the caller repeats with the general path.
#define CC_MERGE_SMALL 65536
__global__ void __launch_bounds__(1024)
k_merge_small(u32* __restrict__ parent, SlabRows f, int rank, u64 label_cap, u64 pair_cap, u32* __restrict__ remap,
              unsigned long long* __restrict__ result) {
  __shared__ u64 s_off[CC_MERGE_MAX_WORLD + 1];
  __shared__ u32 s_nr[CC_MERGE_SMALL / 32], s_pre[CC_MERGE_SMALL / 32], s_wsum[32], s_nonroot;
  slab_offsets(f, s_off);
  const u64 total = s_off[f.world];
  const bool fits = total + 1 <= CC_MERGE_SMALL && total + 1 <= label_cap;
  if (threadIdx.x == 0) {
    bool pov = false;
    for (int r = 0; r < f.world; r++) pov |= (u64)f.rows[(size_t)r * f.stride + 3] > pair_cap;
    result[0] = 0; result[1] = (total + 1 > label_cap) ? 1ull : 0ull; result[2] = pov ? 1ull : 0ull;
    result[3] = fits ? total + 1 : 0ull; result[4] = 0; result[5] = (total + 1 > CC_MERGE_SMALL) ? 1ull : 0ull;
    for (int r = 0; r < f.world; r++)
      for (int k = 0; k < 4; k++) result[8 + 4 * r + k] = (unsigned long long)f.rows[(size_t)r * f.stride + k];
  }
  if (!fits) return;
  const u32 n = (u32)total + 1u;
  for (u32 i = threadIdx.x; i < n; i += blockDim.x) parent[i] = i;
  __syncthreads();
  for (int r = 1; r < f.world; r++) {
    const long long* row = f.rows + (size_t)r * f.stride;
    const u64 np = min((u64)row[3], pair_cap);
    const u64 nlo = (u64)f.rows[(size_t)(r - 1) * f.stride], nup = (u64)row[0];
    for (u64 k = threadIdx.x; k < np; k += blockDim.x) {
      const u64 v = (u64)row[4 + k];
      const u64 lo = v >> 32, up = v & 0xFFFFFFFFull;
      if (lo < 1 || up < 1 || lo > nlo || up > nup) continue;   // malformed pair
      uf_union(parent, (u32)(s_off[r - 1] + lo), (u32)(s_off[r] + up));
    }
  }
  __syncthreads();
  // every id -> its root; non-root flags per 32 ids
  const u32 nwords = (n + 31) >> 5;
  const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (u32 wd = warp; wd < nwords; wd += blockDim.x >> 5) {
    const u32 i = (wd << 5) + lane;
    bool nonroot = false;
    if (i >= 1 && i < n) {
      u32 r = i, p;
      while ((p = __ldcg(&parent[r])) != r) r = p;
      if (r != i) { parent[i] = r; nonroot = true; }
    }
    const u32 m = __ballot_sync(CC_FULL, nonroot);
    if (lane == 0) s_nr[wd] = m;
  }
  __syncthreads();
  // exclusive prefix of the popcounts over the words (two words per thread at most)
  {
    const u32 w0 = 2 * threadIdx.x, w1 = w0 + 1;
    const u32 c0 = w0 < nwords ? __popc(s_nr[w0]) : 0u, c1 = w1 < nwords ? __popc(s_nr[w1]) : 0u;
    u32 inc = c0 + c1;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const u32 v = __shfl_up_sync(CC_FULL, inc, o); if ((int)lane >= o) inc += v; }
    if (lane == 31) s_wsum[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      u32 v = s_wsum[lane], si = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const u32 t = __shfl_up_sync(CC_FULL, si, o); if ((int)lane >= o) si += t; }
      s_wsum[lane] = si - v;
      if (lane == 31) s_nonroot = si;
    }
    __syncthreads();
    const u32 excl = s_wsum[warp] + inc - (c0 + c1);
    if (w0 < nwords) s_pre[w0] = excl;
    if (w1 < nwords) s_pre[w1] = excl + c0;
  }
  __syncthreads();
  if (threadIdx.x == 0) { result[4] = s_nonroot; result[0] = (unsigned long long)(n - 1u - s_nonroot); }
  const u32 off = (u32)s_off[rank], nl = (u32)f.rows[(size_t)rank * f.stride];
  for (u32 l = threadIdx.x; l <= nl; l += blockDim.x) {
    u32 out = 0;
    if (l) {
      u32 root = parent[off + l];
      root = parent[root];
      out = root - (s_pre[root >> 5] + __popc(s_nr[root >> 5] & ((1u << (root & 31)) - 1u)));
    }
    remap[l] = out;
  }
}

// union-find over compact node ids (pairs given as two u32 arrays); parent must hold 0..n-1 on entry
__global__ void __launch_bounds__(256) k_iota(u32* __restrict__ p, i64 n) {
  const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = (u32)i;
}
__global__ void __launch_bounds__(256)
k_union_pairs(u32* __restrict__ parent, const u32* __restrict__ a, const u32* __restrict__ b, i64 n) {
  const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) uf_union(parent, a[i], b[i]);
}
__global__ void __launch_bounds__(256) k_flatten(u32* __restrict__ parent, i64 n) {
  const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  u32 r = (u32)i, p;
  while ((p = __ldcg(&parent[r])) != r) r = p;
  parent[i] = r;
}

// ---- binary 2D 8-connected: number components by their first 2x2 block in block-raster order
// (cc3d_binary.hpp:1016-1023, 1215-1231). K[root] = min block key over the component's runs. ----
__global__ void __launch_bounds__(256) k_fill_n(u32* __restrict__ K, u32 v, const u64* __restrict__ n_dev) {
  const u32 n = (u32)*n_dev;
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) K[i] = v;
}

// one atomic per run: along a row the block key grows with x, so a run's minimum is at its start
__global__ void __launch_bounds__(256)
k_blockkey_min(const u32* __restrict__ L, const u32* __restrict__ M, u32* __restrict__ K, Geom g) {
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (u32)g.nwords) return;
  const uint2 fx = __ldg(reinterpret_cast<const uint2*>(M) + 2 * (size_t)i);
  if (!fx.x) return;
  const u32 W = (u32)g.W;
  const u32 row = i / W, w = i - row * W;
  u32 bits = fx.x & ~fx.y;
  u32 id = __ldg(M + g.offRS + i);
  const u32 osx = ((u32)g.sx + 1) >> 1;
  while (bits) {
    const int b = __ffs(bits) - 1;
    bits &= bits - 1;
    const u32 x = (w << 5) + b;
    const u32 root = L[id];                              // after k_compress L[id] is the root (roots: themselves)
    const u32 key = (x >> 1) + osx * (row >> 1);
    if (key < __ldcg(&K[root])) atomicMin(&K[root], key);  // the plain read keeps giant components off one address
    id++;
  }
}

__global__ void __launch_bounds__(256)
k_blockkey_mark(const u32* __restrict__ K, const u32* __restrict__ GR, u32* __restrict__ BK, const u64* __restrict__ n_dev) {
  const u32 n = (u32)*n_dev;
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    if ((GR[i >> 5] >> (i & 31)) & 1u) {
      const u32 k = K[i];
      atomicOr(&BK[k >> 5], 1u << (k & 31));
    }
  }
}

__global__ void __launch_bounds__(256) k_popc(const u32* __restrict__ bm, u32* __restrict__ cnt, i64 n) {
  const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) cnt[i] = __popc(bm[i]);
}

__global__ void __launch_bounds__(256)
k_assign_blockorder(u32* __restrict__ L, const u32* __restrict__ K, const u32* __restrict__ BK,
                    const u32* __restrict__ bprefix, const u64* __restrict__ n_dev) {
  const u32 n = (u32)*n_dev;
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const u32 k = K[L[i]];
    L[i] = rank_in_bitmap(BK, bprefix, k >> 5, (int)(k & 31)) + 1u;
  }
}
