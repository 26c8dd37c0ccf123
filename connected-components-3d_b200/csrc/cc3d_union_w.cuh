// cc3d_union_w.cuh — kernel B1, warp-owned formulation (round 2): the tile-local unions of cc3d_union.cuh without
// block-wide phases. Replaces cc3d.hpp:64-149 (DisjointSet) for the runs of one union tile.
//
// What the CTA-phased kernels (k_union_tile / k_union_tile_hybrid) paid for, measured with ncu on the connectomics
// volume (profiles/r02_ncu_full_summary.md): a third of the stall samples at the six barriers between the
// enumerate / union phases, 4.5 % of the instructions to initialise a 8 192-entry forest of which ~600 entries are
// used, 15 % for a per-word write-back loop with 7 active lanes, 6 % for the compare-and-swap loop that emulates a
// 16-bit atomic minimum. Here:
//   * COMPACT nodes: the runs that start in a tile row segment have contiguous global ids, so an exclusive scan of
//     the segment sizes (every warp computes it redundantly with one shuffle scan: no barrier) gives every run of the
//     tile a dense local id that keeps the raster order (link-to-smaller stays valid). The forest holds as many
//     32-bit entries as the tile has runs: native shared-memory atomicMin instead of a CAS loop, no 16-runs-per-word limit
//     (a tile may hold up to CAPN runs in total), nothing to initialise beyond the runs that exist.
//   * WARP-OWNED words: a warp owns 64 consecutive words of the tile (whole row segments, so every edge of a run is
//     enumerated by the warp that owns the run) and takes them through every phase on its own: straight edges,
//     deferred unions, diagonal candidates. The forest is lock-free, so warps union concurrently; the CTA meets at
//     two barriers only: after the forest / stash initialisation and before the write-back.
//   * FIRST LINK WITHOUT FINDS: an edge (p, q) has q < p (q lies in an earlier row). atomicMin(&parent[p], q) returns
//     p when p was still a root: the link is made and nothing else is needed (70 % of the edges of a label volume).
//     Otherwise the displaced / competing parent `old` is queued as union(old, q) in a warp-private queue that the
//     warp drains with full lanes.
//   * NODE-PARALLEL write-back: one thread per run of the tile (not per word) chases its root and converts both ids
//     back to global run ids with a binary search over the segment table. When the tile spans whole rows of the
//     volume (sx <= 512) the rows of one plane of the tile are contiguous in the global numbering, so a segment is
//     a PLANE of the tile: four segments (one in 2D) instead of 32.
#pragma once
#include "cc3d_union.cuh"

__device__ __forceinline__ u32 sm_find32(volatile u32* A, u32 i) {
  u32 p = A[i];
  while (p != i) {
    const u32 gp = A[p];
    if (gp == p) return p;
    A[i] = gp;
    i = gp;
    p = A[i];
  }
  return i;
}
__device__ __forceinline__ void sm_union32(u32* A, u32 a, u32 b) {
  bool done;
  do {
    a = sm_find32(A, a);
    b = sm_find32(A, b);
    if (a < b) { const u32 old = atomicMin(&A[b], a); done = (old == b); b = old; }
    else if (b < a) { const u32 old = atomicMin(&A[a], b); done = (old == a); a = old; }
    else done = true;
  } while (!done);
}

template <int MODE, u32 CAPN> struct WarpTile {
  static constexpr u32 NWARPS = CC_TILE_THREADS / 32;
  static constexpr u32 WWORDS = CC_TILE_WORDS / NWARPS;     // words a warp owns
  static constexpr u32 WQ = 128;                            // per-warp queue: deferred unions (round 0) / diagonal items (round 1)
  static constexpr u32 GQ = MODE == MODE_EQ ? CC_TILE_GQ_EQ : CC_TILE_GQ;
  static constexpr u32 SEG = CC_TILE_WORDS + 4;             // segment tables (a segment has at least one word)
  static constexpr u32 SMEM_WORDS = CAPN + 2 * SEG + NWARPS * WQ + 2 * GQ + CC_TILE_WORDS / 2;
};

// grid = (tiles in x, tiles in y, tiles in z)
// DENSE: tiles with more runs are flagged for the second launch (<= CAPN)
template <typename T, int MODE, int CONN, u32 CAPN, bool SECOND, u32 DENSE>
__global__ void __launch_bounds__(CC_TILE_THREADS, CAPN <= 4096 ? CC_TILE_MINB(CC_B1W_MINB) : (CAPN <= 8192 ? CC_TILE_MINB(4) : CC_TILE_MINB(2)))
k_union_tile_w(const T* __restrict__ in, const u32* __restrict__ M, u32* __restrict__ L, Geom g, Edge<T, MODE> E,
               EdgeQueue GQ, BigTiles big) {
  CC_PDL_WAIT();
  typedef WarpTile<MODE, CAPN> WT;
  typedef WordEdges<T, MODE, CONN> WE;
  const u32 bx = blockIdx.x, by = blockIdx.y, bz = blockIdx.z;
  const u32 tile = bx + gridDim.x * (by + gridDim.y * bz);
  if constexpr (SECOND) { if (big.flags[tile] == 0) return; }      // second launch: flagged tiles only
  constexpr u32 GQN = WT::GQ, WQ = WT::WQ;
  extern __shared__ __align__(16) u32 smem_u32[];
  u32* lab = smem_u32;                                   // [CAPN] parents (local ids)
  u32* nb = lab + CAPN;                                  // [nsegE + 1] first local id of every segment
  u32* segD = nb + WT::SEG;                              // [nsegE] global id - local id of the runs of the segment
  u32* wqs = segD + WT::SEG;                             // [NWARPS][WQ]
  u64* gq = reinterpret_cast<u64*>(wqs + WT::NWARPS * WQ);   // [GQN] edges that leave the tile
  uint16_t* todos = reinterpret_cast<uint16_t*>(reinterpret_cast<u32*>(gq) + 2 * GQN);   // [NWARPS][WWORDS]
  __shared__ u32 s_wqn[WT::NWARPS], s_gn, s_gbase;
  const u32 W = (u32)g.W, sy = (u32)g.sy, sz = (u32)g.sz, sx = (u32)g.sx;
  const u32 TW = 1u << g.tw, TY = 1u << g.ty;
  const u32 w0 = bx << g.tw, y0 = by << g.ty, z0 = bz << g.tz;
  const u32 wend = min(w0 + TW, W);
  // a segment = runs with contiguous global ids: a row segment of the tile, or - when the tile spans whole rows - all
  // rows of one plane of the tile
  const u32 sshift = (w0 == 0 && wend == W) ? (u32)g.ty : 0u;
  const u32 nsegE = (CC_TILE_WORDS >> g.tw) >> sshift;
  const u32* __restrict__ RS = M + g.offRS;
  const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  // ---- segment table: exclusive scan of the runs per segment (every warp, redundantly: no barrier) ----
  u32 ntot = 0;
#pragma unroll 1
  for (u32 s0 = 0; s0 < nsegE; s0 += 32) {
    const u32 sg = s0 + lane;
    u32 first = 0, cnt = 0;
    if (sg < nsegE) {
      const u32 r = sg << sshift;
      const u32 y = y0 + (r & (TY - 1)), z = z0 + (r >> g.ty);
      if (y < sy && z < sz) {
        const u32 j = (z * sy + y) * W;
        first = __ldg(RS + j + w0);
        const u32 jend = sshift ? (z * sy + min(y0 + TY, sy)) * W : j + wend;
        cnt = __ldg(RS + jend) - first;
      }
    }
    u32 inc = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const u32 v = __shfl_up_sync(CC_FULL, inc, o);
      if (lane >= o) inc += v;
    }
    const u32 excl = ntot + inc - cnt;
    if (sg < nsegE) { nb[sg] = excl; segD[sg] = first - excl; }      // every warp stores the same values
    ntot += __shfl_sync(CC_FULL, inc, 31);
  }
  if (lane == 0) { nb[nsegE] = ntot; s_wqn[warp] = 0; }
  if (nsegE < 4 && lane >= nsegE + 1 && lane < 4) nb[lane] = 0xFFFFFFFFu;      // pads for the compare chain of the write-back
  if (threadIdx.x == 0) s_gn = 0;
  const bool tile_ok = ntot <= (DENSE < CAPN ? DENSE : CAPN);                   // block-uniform
  if constexpr (!SECOND) {
    if (!tile_ok && big.count) {
      if (threadIdx.x == 0) { atomicAdd(big.count, 1u); if (big.defer) big.flags[tile] = 1u; }
      if (big.defer) return;     // the second launch relabels this tile
    }
  }
  if (tile_ok) for (u32 k = threadIdx.x; k < ntot; k += blockDim.x) lab[k] = k;
  __syncthreads();

  auto push_global = [&](u32 gp, u32 gq_) {
    const u32 pos = atomicAdd(GQ.count, 1u);
    if (pos < GQ.cap) GQ.q[pos] = (u64)gp | ((u64)gq_ << 32);
    else *GQ.ovf = 1u;
  };
  auto stage_global = [&](u32 gp, u32 gq_) {
    const u32 pos = atomicAdd(&s_gn, 1u);
    if (pos < GQN) gq[pos] = (u64)gp | ((u64)gq_ << 32);
    else push_global(gp, gq_);
  };

  WE we(in, M, g, E);
  u32* wq = wqs + warp * WQ;
  uint16_t* todo = todos + warp * WT::WWORDS;
  u32 ntodo = 0;                                  // warp-uniform

  // ---- round 0: straight edges of the warp's words ----
#pragma unroll 1
  for (u32 it = 0; it < WT::WWORDS / 32; it++) {
    const u32 q = warp * WT::WWORDS + it * 32 + lane;
    const u32 wx = q & (TW - 1), r = q >> g.tw;
    const u32 ly = r & (TY - 1), lz = r >> g.ty;
    const u32 w = w0 + wx, y = y0 + ly, z = z0 + lz;
    bool td = false;
    if (w < W && y < sy && z < sz) {
      const u32 row = z * sy + y;
      // eager: every word the straight edges can need is requested at once - the kernel is bound by its chains of
      // dependent loads, not by instruction issue (profiles/r02_experiments.md section 8)
      if (we.load_eager(row * W + w, row, w, y, z)) {
        const u32 Sp = we.Sp;
        const u32 gP = we.RSp;                                   // global id of the first run that starts in the word, minus 1
        const u32 needY = we.need_y(), needZ = we.need_z();
        // an edge is tile-local when both of its runs START in the tile: the neighbour row lies in the tile and neither
        // voxel belongs to a run that enters the tile from the left. Such a run exists only when the tile does not begin
        // at x = 0, and it owns the bits below the first run start of a word as long as no run has started in the tile
        // row before that word (local id base == first local id of the row segment - 1).
        const u32 SqY = we.U.F & ~we.U.X, SqZ = we.D.F & ~we.D.X;
        u32 locY = 0, locZ = 0, gQY = 0, gQZ = 0, lQY = 0, lQZ = 0;
        const u32 sgP = r >> sshift;
        const u32 dP = segD[sgP];
        auto from_first_start = [](u32 S) { return S ? ~((S & (0u - S)) - 1u) : 0u; };
        u32 locP = CC_FULL;
        if (w0 != 0 && gP - dP + 1u == nb[sgP]) locP = from_first_start(Sp);
        if (needY) {
          gQY = we.RSu;
          if (tile_ok && ly > 0) {
            const u32 sgQ = (r - 1) >> sshift;
            lQY = gQY - segD[sgQ];
            locY = locP;
            if (w0 != 0 && lQY + 1u == nb[sgQ]) locY &= from_first_start(SqY);
          }
        }
        if (needZ) {
          gQZ = we.RSd;
          if (tile_ok && lz > 0) {
            const u32 sgQ = (r - TY) >> sshift;
            lQZ = gQZ - segD[sgQ];
            locZ = locP;
            if (w0 != 0 && lQZ + 1u == nb[sgQ]) locZ &= from_first_start(SqZ);
          }
        }
        // tile-local edges: first link by one atomicMin
        {
          const u32 lP = gP - dP;
          u32 need = needY & locY, need2 = needZ & locZ, Sq = SqY, lQ = lQY;
          if (!need) { need = need2; need2 = 0; Sq = SqZ; lQ = lQZ; }
          while (need) {
            const int b = __ffs(need) - 1; need &= need - 1;
            const u32 below = CC_FULL >> (31 - b);
            const u32 a = lP + __popc(Sp & below), c = lQ + __popc(Sq & below);
            const u32 old = atomicMin(&lab[a], c);
            if (old != a && old != c) {             // a already had a parent: union(old, c) is still owed
              const u32 pos = atomicAdd(&s_wqn[warp], 1u);
              if (pos < WQ) wq[pos] = old | (c << 16);
              else sm_union32(lab, old, c);
            }
            if (!need) { need = need2; need2 = 0; Sq = SqZ; lQ = lQZ; }
          }
        }
        // edges that leave the tile: one reservation per word in the staging buffer
        {
          u32 need = needY & ~locY, need2 = needZ & ~locZ, Sq = SqY, gQ = gQY;
          const u32 n = __popc(need) + __popc(need2);
          if (n) {
            u32 pos = atomicAdd(&s_gn, n);
            if (!need) { need = need2; need2 = 0; Sq = SqZ; gQ = gQZ; }
            while (need) {
              const int b = __ffs(need) - 1; need &= need - 1;
              const u32 below = CC_FULL >> (31 - b);
              const u32 ga = gP + __popc(Sp & below), gc = gQ + __popc(Sq & below);
              if (pos < GQN) gq[pos] = (u64)ga | ((u64)gc << 32);
              else push_global(ga, gc);
              pos++;
              if (!need) { need = need2; need2 = 0; Sq = SqZ; gQ = gQZ; }
            }
          }
        }
        td = we.may_have_diagonals();
      }
    }
    if constexpr (WE::DIAG0) {
      const u32 tm = __ballot_sync(CC_FULL, td);
      if (td) todo[ntodo + __popc(tm & ((1u << lane) - 1u))] = (uint16_t)q;
      ntodo += __popc(tm);
    }
    __syncwarp();
    // deferred unions, one per lane
    const u32 nq = min(*(volatile u32*)&s_wqn[warp], WQ);
    if (nq) {
      for (u32 e = lane; e < nq; e += 32) {
        const u32 v = wq[e];
        sm_union32(lab, v & 0xFFFFu, v >> 16);
      }
      __syncwarp();
      if (lane == 0) s_wqn[warp] = 0;
      __syncwarp();
    }
  }

  // ---- round 1: diagonal candidates of the to-do words. Word-parallel masks, expanded into work items
  //      (word, direction, bit) in the warp's queue, then one item per lane: value test for EQ / DELTA, run ids of
  //      both ends from the bitmaps, union or staging. ----
  if constexpr (WE::DIAG0) {
    // two bits per diagonal direction of WordEdges::diag_masks: d + 1
    constexpr u32 DXP = (0u << 0) | (2u << 2) | (0u << 4) | (2u << 6) | (1u << 8) | (0u << 10) | (2u << 12) | (1u << 14) | (0u << 16) | (2u << 18);
    constexpr u32 DYP = (0u << 0) | (0u << 2) | (1u << 4) | (1u << 6) | (0u << 8) | (0u << 10) | (0u << 12) | (2u << 14) | (2u << 16) | (2u << 18);
    constexpr u32 DZP = (1u << 0) | (1u << 2) | (0u << 4) | (0u << 6) | (0u << 8) | (0u << 10) | (0u << 12) | (0u << 14) | (0u << 16) | (0u << 18);
    auto resolve = [&](const u32 item) {
      const u32 b = item & 31u, tdir = (item >> 5) & 15u, q = item >> 9;
      const u32 wx = q & (TW - 1), r = q >> g.tw;
      const int ly = (int)(r & (TY - 1)), lz = (int)(r >> g.ty);
      const int dx = (int)((DXP >> (2 * tdir)) & 3u) - 1, dy = (int)((DYP >> (2 * tdir)) & 3u) - 1, dz = (int)((DZP >> (2 * tdir)) & 3u) - 1;
      const int xl = (int)((wx << 5) + b) + dx;            // x of q relative to the tile
      const int lyq = ly + dy, lzq = lz + dz;
      const bool inside = xl >= 0 && xl < (int)(TW << 5) && lyq >= 0 && lyq < (int)TY && lzq >= 0;
      const u32 rowP = (z0 + lz) * sy + y0 + ly;
      const u32 rowQ = (u32)((int)rowP + dy + dz * (int)sy);
      const u32 xp = ((w0 + wx) << 5) + b;
      const u32 xq = (u32)((int)(w0 << 5) + xl);
      if constexpr (MODE == MODE_EQ || MODE == MODE_DELTA || MODE == MODE_BLOCK) {
        if (!E.diag((int)tdir, in[(size_t)rowP * sx + xp], in[(size_t)rowQ * sx + xq])) return;
      }
      const u32 gp = run_id(M, g, rowP * W, xp), gq_ = run_id(M, g, rowQ * W, xq);
      bool local = tile_ok && inside;
      u32 lp = 0, lq_ = 0;
      if (local) {
        const u32 sp = r >> sshift, sq = (((u32)lzq << g.ty) + (u32)lyq) >> sshift;
        lp = gp - segD[sp]; lq_ = gq_ - segD[sq];
        local = (int)(lp - nb[sp]) >= 0 && (int)(lq_ - nb[sq]) >= 0;
      }
      if (local) sm_union32(lab, lp, lq_);
      else stage_global(gp, gq_);
    };
#pragma unroll 1
    for (u32 base = 0; base < ntodo; base += 32) {
      const u32 e = base + lane;
      u32 m[10];
#pragma unroll
      for (int k = 0; k < 10; k++) m[k] = 0;
      u32 q = 0;
      if (e < ntodo) {
        q = (u32)todo[e];
        const u32 wx = q & (TW - 1), r = q >> g.tw;
        const u32 w = w0 + wx, y = y0 + (r & (TY - 1)), z = z0 + (r >> g.ty);
        const u32 row = z * sy + y;
        if (we.load(row * W + w, row, w, y, z)) we.diag_masks(m);
      }
      u32 n = 0;
#pragma unroll
      for (int k = 0; k < 10; k++) n += __popc(m[k]);
      if (!__any_sync(CC_FULL, n != 0)) continue;
      u32 inc = n;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const u32 v = __shfl_up_sync(CC_FULL, inc, o);
        if (lane >= o) inc += v;
      }
      const u32 wtot = __shfl_sync(CC_FULL, inc, 31);
      u32 pos = inc - n;
      if (n) {
        const u32 qb = q << 9;
#pragma unroll
        for (int k = 0; k < 10; k++) {
          u32 mk = m[k];
          while (mk) {
            const u32 bb = __ffs(mk) - 1; mk &= mk - 1;
            const u32 item = qb | ((u32)k << 5) | bb;
            if (pos < WQ) wq[pos] = item;
            else resolve(item);          // list full: resolve in place
            pos++;
          }
        }
      }
      __syncwarp();
      const u32 ni = min(wtot, WQ);
      for (u32 kk = lane; kk < ni; kk += 32) resolve(wq[kk]);
      __syncwarp();
    }
  }

  // ---- runs -> tile roots (one thread per run); staged edges -> global queue ----
  __syncthreads();
  const u32 gn = min(s_gn, GQN);
  if (threadIdx.x == 0 && gn) s_gbase = atomicAdd(GQ.count, gn);
  if (tile_ok) {
    auto seg_of = [&](u32 k) {
      if (nsegE <= 4) return (u32)(k >= nb[1]) + (u32)(k >= nb[2]) + (u32)(k >= nb[3]);     // plane segments: no search
      u32 lo = 0, len = nsegE;
      while (len > 1) {
        const u32 half = len >> 1;
        if (nb[lo + half] <= k) { lo += half; len -= half; } else len = half;
      }
      return lo;
    };
#pragma unroll 1
    for (u32 k = threadIdx.x; k < ntot; k += blockDim.x) {
      u32 l = k, p;
      while ((p = lab[l]) != l) l = p;
      const u32 dk = segD[seg_of(k)];
      L[k + dk] = l == k ? k + dk : l + segD[seg_of(l)];
    }
  } else {
    // every run of the tile stays its own root (all edges went to the global queue)
#pragma unroll 1
    for (u32 sg = 0; sg < nsegE; sg++) {
      const u32 n0 = nb[sg], n1 = nb[sg + 1], d = segD[sg];
      for (u32 k = n0 + threadIdx.x; k < n1; k += blockDim.x) L[k + d] = k + d;
    }
  }
  __syncthreads();
  for (u32 e = threadIdx.x; e < gn; e += blockDim.x) {
    const u32 pos = s_gbase + e;
    if (pos < GQ.cap) GQ.q[pos] = gq[e];
    else *GQ.ovf = 1u;
  }
}
