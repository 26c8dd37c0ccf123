// cc3d_label.cuh — labelling kernels A (tile), B1/B2 (seams), P (periodic). See cc3d_common.cuh.
#pragma once
#include "cc3d_common.cuh"

#define CC_TILE_THREADS 256

// Shared-memory bytes kernel A needs (TX = 64, 64 rows per tile).
template <typename T>
__host__ __device__ constexpr size_t tile_smem_bytes() {
  // slab[4096 + 64] + slink[128] + sB[256] (u32) + sval[64][66] (T)
  return (size_t)(4096 + 64 + 128 + 256) * 4 + (size_t)64 * 66 * sizeof(T);
}

template <typename T> __device__ __forceinline__ T shfl_t(T v, int src) { return (T)__shfl_sync(CC_FULL, v, src); }
template <typename T> __device__ __forceinline__ T shfl_up_t(T v) { return (T)__shfl_up_sync(CC_FULL, v, 1); }

// path-compressing find for the shared-memory forest: the start node is re-pointed at the root
__device__ __forceinline__ u32 uf_find_c(u32* A, u32 i) {
  volatile u32* V = A;
  u32 r = i, p;
  while ((p = V[r]) != r) r = p;
  if (r != i) atomicMin(&A[i], r);
  return r;
}
static __device__ __noinline__ void uf_union_c(u32* A, u32 a, u32 b) {
  bool done;
  do {
    a = uf_find_c(A, a);
    b = uf_find_c(A, b);
    if (a < b) { u32 old = atomicMin(&A[b], a); done = (old == b); b = old; }
    else if (b < a) { u32 old = atomicMin(&A[a], b); done = (old == a); a = old; }
    else done = true;
  } while (!done);
}

// ---------------------------------------------------------------------------------------------
// Kernel A. One CTA (8 warps) labels one 64 x TY x TZ tile (TY*TZ = 64 rows) in shared memory;
// warp w owns rows 8w..8w+7 (one z-plane of an 8x8 tile) and keeps their voxels in registers.
//  phase 1: 16 coalesced loads per warp issued back to back (+ the x0-1 halo column), x-runs by ballot:
//           every voxel starts out pointing at the first voxel of its run inside its 32-wide segment;
//           the epl transition count (cc3d.hpp:300-303) falls out of the same ballots
//  phase 2a: ballot masks of the two "straight" backward edges (dy=-1 and dz=-1) of every segment
//  phase 2b: remaining neighbour rows, evaluated only on the lanes where they can matter (for the
//           transitive predicates EQ/NONZERO a voxel that matches its -y or -z neighbour inherits that
//           neighbour's diagonal connections), then redundancy elimination on the masks (an edge is
//           dropped when a neighbouring lane / row already joins the same two runs) and shared-memory
//           atomicMin unions for the few edges left; the halo column joins through the same forest
//           with indices >= TILE so that it never becomes a root
//  phase 3: flatten, write L (raster index of the local root), the local-root bitmap word and the
//           x-seam slot (own-tile root that the halo voxel to the left belongs to)
// ---------------------------------------------------------------------------------------------
template <typename T, int MODE, int CONN, int TYL, int TZL>
__global__ void __launch_bounds__(CC_TILE_THREADS)
k_tile_label(const T* __restrict__ in, u32* __restrict__ L, u32* __restrict__ LR, u32* __restrict__ XS,
             Geom g, Edge<T, MODE> E, Counters* __restrict__ ctr) {
  constexpr int TX = 64, NSEG = 2, SVX = TX + 2;
  constexpr int TY = 1 << TYL, TZ = 1 << TZL, ROWS = TY * TZ, TILE = ROWS * TX;
  static_assert(ROWS == 64, "8 warps x 8 rows");
  constexpr int RPW = 8;
  constexpr bool TRANS = (MODE == MODE_EQ || MODE == MODE_NONZERO);
  constexpr int NR = hood_rows(CONN);
  constexpr bool HAS_Z = (NR >= 2) && (TZ > 1);

  extern __shared__ __align__(16) unsigned char smem_raw[];
  u32* slab = reinterpret_cast<u32*>(smem_raw);   // [TILE + ROWS]
  u32* slink = slab + TILE + ROWS;                // [ROWS * NSEG]   x-link masks
  u32* sB = slink + ROWS * NSEG;                  // [ROWS * NSEG][2] dy=-1 / dz=-1 straight-edge masks
  T* sval = reinterpret_cast<T*>(sB + ROWS * NSEG * 2);  // [ROWS][SVX]

  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int wrow0 = warp * RPW;

  unsigned t = blockIdx.x;
  const unsigned ntx = (unsigned)g.ntx, nty = (unsigned)g.nty;
  const unsigned bx = t % ntx; t /= ntx;
  const unsigned by = t % nty;
  const unsigned bz = t / nty;
  const i64 x0 = (i64)bx * TX, y0 = (i64)by * TY, z0 = (i64)bz * TZ;

  // ---- phase 1: load + x-runs ----
  T v[RPW][NSEG];
#pragma unroll
  for (int j = 0; j < RPW; j++) {
    const int row = wrow0 + j;
    const i64 gy = y0 + (row & (TY - 1)), gz = z0 + (row >> TYL);
    const bool rowok = gy < g.sy && gz < g.sz;
    const T* src = in + ((gz * g.sy + gy) * g.sx + x0);
#pragma unroll
    for (int k = 0; k < NSEG; k++) {
      const int lx = k * 32 + lane;
      v[j][k] = (rowok && x0 + lx < g.sx) ? src[lx] : (T)0;
    }
  }
  T hv = (T)0;
  if (lane < RPW && x0 > 0) {
    const int row = wrow0 + lane;
    const i64 gy = y0 + (row & (TY - 1)), gz = z0 + (row >> TYL);
    if (gy < g.sy && gz < g.sz) hv = in[(gz * g.sy + gy) * g.sx + x0 - 1];
  }
  if (lane < RPW) {
    sval[(wrow0 + lane) * SVX] = hv;
    sval[(wrow0 + lane) * SVX + TX + 1] = (T)0;
    slab[TILE + wrow0 + lane] = E.fg(hv) ? (u32)(TILE + wrow0 + lane) : CC_BG;
  }
  u32 epl_local = 0;
  int row_min = ROWS, row_max = -1;
#pragma unroll
  for (int j = 0; j < RPW; j++) {
    const int row = wrow0 + j;
    const T hvj = shfl_t(hv, j);
    const T t31 = shfl_t(v[j][0], 31);
#pragma unroll
    for (int k = 0; k < NSEG; k++) {
      const int lx = k * 32 + lane;
      const T cur = v[j][k];
      sval[row * SVX + 1 + lx] = cur;
      T vl = shfl_up_t(cur);
      if (lane == 0) vl = (k == 0) ? hvj : t31;
      const bool f = E.fg(cur);
      const bool link = f && (k > 0 || lane > 0) && E(cur, vl, dir_code(-1, 0, 0));
      const u32 m = __ballot_sync(CC_FULL, link);
      const u32 below = (~m & (CC_FULL >> (31 - lane))) | 1u;
      const int start = 31 - __clz(below);
      slab[row * TX + lx] = f ? (u32)(row * TX + k * 32 + start) : CC_BG;
      if (lane == 0) slink[row * NSEG + k] = m;
      if constexpr (MODE != MODE_MASK) {
        const bool tr = f && ((x0 + lx == 0) || (cur != vl));
        const u32 tm = __ballot_sync(CC_FULL, tr);
        if (tm) { epl_local += __popc(tm); row_min = min(row_min, row); row_max = max(row_max, row); }
      }
    }
  }
  __syncthreads();

  // ---- phase 2a: straight-edge masks B0 (dy=-1) and B1 (dz=-1) ----
  // (rolled loops from here on: the unrolled form is ~300 KB of SASS and thrashes the instruction cache)
#pragma unroll 1
  for (int s = wrow0 * NSEG; s < (wrow0 + RPW) * NSEG; s++) {
    const int row = s >> 1, k = s & 1;
    const int ly = row & (TY - 1), lz = row >> TYL;
    const int lx = k * 32 + lane;
    const T cur = sval[row * SVX + 1 + lx];
    const bool f = E.fg(cur);
    u32 B0 = 0, B1 = 0;
    if (ly > 0) B0 = __ballot_sync(CC_FULL, f && E(cur, sval[(row - 1) * SVX + 1 + lx], dir_code(0, -1, 0)));
    if constexpr (HAS_Z) {
      if (lz > 0) B1 = __ballot_sync(CC_FULL, f && E(cur, sval[(row - TY) * SVX + 1 + lx], dir_code(0, 0, -1)));
    }
    if (lane == 0) { sB[s * 2] = B0; sB[s * 2 + 1] = B1; }
  }
  __syncthreads();

  // ---- phase 2b: remaining edges, redundancy elimination, unions ----
#pragma unroll 1
  for (int s = wrow0 * NSEG; s < (wrow0 + RPW) * NSEG; s++) {
    const int row = s >> 1, k = s & 1;
    const int ly = row & (TY - 1), lz = row >> TYL;
    const int lx = k * 32 + lane;
    const u32 li = row * TX + lx;
    const T cur = sval[row * SVX + 1 + lx];
    const bool f = E.fg(cur);
    const u32 F = __ballot_sync(CC_FULL, f);
    if (F == 0) continue;
    const u32 m_own = slink[s];
    const u32 LRs1 = (m_own >> 1) | ((k == 0 ? slink[s + 1] : 0u) << 31);
    if (k > 0 && lane == 0 && (m_own & 1u)) uf_union_c(slab, li, li - 1);  // run continues from the previous segment
    const u32 B0 = sB[s * 2], B1 = sB[s * 2 + 1];

    // one neighbour row: edges a (dx-1), b (dx0, mask given or computed), c (dx+1); in-row elimination; unions
    auto do_row = [&](const int r, const int row2, const bool have_b, u32 Bm, const u32 Bprev, const u32 Bnext,
                      const u32 cand_b, const u32 cand_ac, const u32 kill_b) {
      const int dy = row_dy(r), dz = row_dz(r);
      const int dxm = hood_dx(CONN, r);
      const T* q = sval + row2 * SVX + 1 + lx;
      if (!have_b) {
        if (cand_b == 0 && cand_ac == 0) return;
        Bm = __ballot_sync(CC_FULL, ((cand_b >> lane) & 1u) && E(cur, q[0], dir_code(0, dy, dz)));
      }
      u32 Am = 0, Cm = 0;
      if ((dxm & 5) && cand_ac) {
        const bool ca = (cand_ac >> lane) & 1u;
        Am = __ballot_sync(CC_FULL, ca && lx > 0 && E(cur, q[-1], dir_code(-1, dy, dz)));
        Cm = __ballot_sync(CC_FULL, ca && lx + 1 < TX && E(cur, q[1], dir_code(1, dy, dz)));
      }
      if ((Am | Bm | Cm) == 0) return;
      const u32 LP = slink[row2 * NSEG + k];
      const u32 LPs1 = (LP >> 1) | ((k == 0 ? slink[row2 * NSEG + 1] : 0u) << 31);
      const u32 Bl = (Bm << 1) | (Bprev >> 31);
      const u32 Br = (Bm >> 1) | (Bnext << 31);
      const u32 needB = Bm & ~(m_own & LP & Bl) & ~kill_b;
      const u32 needA = Am & ~(Bm & LP) & ~(m_own & Bl);
      const u32 needC = Cm & ~(Bm & LPs1) & ~(LRs1 & Br);
      const u32 qi = row2 * TX + lx;
      if ((needB >> lane) & 1u) uf_union_c(slab, li, qi);
      if ((needA >> lane) & 1u) uf_union_c(slab, li, qi - 1);
      if ((needC >> lane) & 1u) uf_union_c(slab, li, qi + 1);
    };

    // R0: (dy=-1, dz=0)
    if (ly > 0) {
      const u32 Bprev = (k == 1) ? sB[(s - 1) * 2] : 0u, Bnext = (k == 0) ? sB[(s + 1) * 2] : 0u;
      u32 cand_ac = F;
      if constexpr (TRANS) cand_ac = (CONN == 26) ? (F & ~(B0 | B1)) : (F & ~B0);
      do_row(0, row - 1, true, B0, Bprev, Bnext, 0u, cand_ac, 0u);
    }
    if constexpr (HAS_Z) {
      if (lz > 0) {
        // R1: (dy=0, dz=-1); square rule: (x,y,z)-(x,y-1,z)-(x,y-1,z-1)-(x,y,z-1) already closes the loop
        {
          const u32 Bprev = (k == 1) ? sB[(s - 1) * 2 + 1] : 0u, Bnext = (k == 0) ? sB[(s + 1) * 2 + 1] : 0u;
          const u32 B1up = (ly > 0) ? sB[(s - NSEG) * 2 + 1] : 0u;
          const u32 B0down = sB[(s - TY * NSEG) * 2];
          u32 cand_ac = F;
          if constexpr (TRANS) cand_ac = (CONN == 26) ? (F & ~(B0 | B1)) : (F & ~B1);
          do_row(1, row - TY, true, B1, Bprev, Bnext, 0u, cand_ac, B0 & B1up & B0down);
        }
        if constexpr (NR >= 4) {
          // R2: (dy=-1, dz=-1)
          if (ly > 0) {
            const u32 B1up = sB[(s - NSEG) * 2 + 1];
            const u32 B0down = sB[(s - TY * NSEG) * 2];
            const u32 cand = TRANS ? (F & ~(B0 | B1)) : F;
            do_row(2, row - TY - 1, false, 0u, 0u, 0u, cand, cand, (B0 & B1up) | (B1 & B0down));
          }
          // R3: (dy=+1, dz=-1)
          if (ly < TY - 1) {
            const u32 B0d1 = sB[(s - TY * NSEG + NSEG) * 2];  // (x,y+1,z-1)-(x,y,z-1)
            const u32 cand = TRANS ? (F & ~B1) : F;
            do_row(3, row - TY + 1, false, 0u, 0u, 0u, cand, cand, B1 & B0d1);
          }
        }
      }
    }
  }
  // halo column: edges between the halo voxel h=(x0-1,y,z) and own voxels (x0, y+ddy, z+ddz),
  // plus halo-halo links used to drop redundant x-seam slots.
  if (x0 > 0 && threadIdx.x < ROWS) {
    const int row = threadIdx.x;
    const T hvv = sval[row * SVX];
    if (E.fg(hvv)) {
      const int ly = row & (TY - 1), lz = row >> TYL;
      const u32 hi = TILE + row;
#pragma unroll 1
      for (int ddz = -1; ddz <= 1; ddz++) {
#pragma unroll 1
        for (int ddy = -1; ddy <= 1; ddy++) {
          const int nz = (ddy != 0) + (ddz != 0);
          bool allowed;
          if (CONN == 4 || CONN == 6) allowed = nz == 0;
          else if (CONN == 8) allowed = ddz == 0;
          else if (CONN == 18) allowed = nz <= 1;
          else allowed = true;
          if (!allowed) continue;
          const int ly2 = ly + ddy, lz2 = lz + ddz;
          if (ly2 < 0 || ly2 >= TY || lz2 < 0 || lz2 >= TZ) continue;
          const int row2 = lz2 * TY + ly2;
          const T ov = sval[row2 * SVX + 1];
          if (!E.fg(ov)) continue;
          bool e;
          if (ddz > 0 || (ddz == 0 && ddy >= 0)) e = E(ov, hvv, dir_code(-1, -ddy, -ddz));  // own voxel is later
          else e = E(hvv, ov, dir_code(1, ddy, ddz));                                        // halo voxel is later
          if (e) uf_union_c(slab, hi, (u32)(row2 * TX));
        }
      }
      if (ly > 0) {
        const T pv = sval[(row - 1) * SVX];
        if (E.fg(pv) && E(hvv, pv, dir_code(0, -1, 0))) uf_union_c(slab, hi, hi - 1);
      }
      if (CONN != 4 && CONN != 8 && lz > 0) {
        const T pv = sval[(row - TY) * SVX];
        if (E.fg(pv) && E(hvv, pv, dir_code(0, 0, -1))) uf_union_c(slab, hi, hi - TY);
      }
    }
  }
  __syncthreads();

  // ---- phase 3: flatten + write ----
#pragma unroll 1
  for (int s = wrow0 * NSEG; s < (wrow0 + RPW) * NSEG; s++) {
    const int row = s >> 1, k = s & 1;
    const i64 gy = y0 + (row & (TY - 1)), gz = z0 + (row >> TYL);
    const bool rowok = gy < g.sy && gz < g.sz;
    const i64 rbase = (gz * g.sy + gy) * g.sx + x0;
    const int lx = k * 32 + lane;
    const u32 li = row * TX + lx;
    u32 l = slab[li];
    u32 out = CC_BG;
    bool is_root = false;
    if (l != CC_BG) {
      u32 p;
      while ((p = slab[l]) != l) l = p;
      is_root = (l == li);
      const int rrow = l >> 6, rlx = l & (TX - 1);
      out = (u32)(((z0 + (rrow >> TYL)) * g.sy + (y0 + (rrow & (TY - 1)))) * g.sx + x0 + rlx);
    }
    const u32 rm = __ballot_sync(CC_FULL, is_root);
    if (rowok && x0 + lx < g.sx) L[rbase + lx] = out;
    if (lane == 0 && rowok && x0 + k * 32 < g.sx) LR[(gz * g.sy + gy) * g.W + ((x0 + k * 32) >> 5)] = rm;
  }
  if (x0 > 0 && threadIdx.x < ROWS) {
    const int row = threadIdx.x;
    const int ly = row & (TY - 1), lz = row >> TYL;
    const i64 gy = y0 + ly, gz = z0 + lz;
    if (gy < g.sy && gz < g.sz) {
      u32 out = CC_BG;
      u32 l = slab[TILE + row];
      if (l != CC_BG) {
        // skip when an earlier halo voxel of the same column is linked to this one (both tiles know that link)
        const T hvv = sval[row * SVX];
        bool covered = false;
        if (ly > 0) { const T pv = sval[(row - 1) * SVX]; covered = E.fg(pv) && E(hvv, pv, dir_code(0, -1, 0)); }
        if (!covered && CONN != 4 && CONN != 8 && lz > 0) {
          const T pv = sval[(row - TY) * SVX]; covered = E.fg(pv) && E(hvv, pv, dir_code(0, 0, -1));
        }
        if (!covered) {
          u32 p;
          while ((p = slab[l]) != l) l = p;
          if (l < (u32)TILE) {
            const int rrow = l >> 6, rlx = l & (TX - 1);
            out = (u32)(((z0 + (rrow >> TYL)) * g.sy + (y0 + (rrow & (TY - 1)))) * g.sx + x0 + rlx);
          }
        }
      }
      XS[(gz * g.sy + gy) * (g.ntx - 1) + (bx - 1)] = out;
    }
  }

  if constexpr (MODE != MODE_MASK) {
    // block-level reduction of epl and the foreground row range
    __shared__ u32 s_epl;
    __shared__ int s_rmin, s_rmax;
    if (threadIdx.x == 0) { s_epl = 0; s_rmin = ROWS; s_rmax = -1; }
    __syncthreads();
    if (lane == 0 && epl_local) {
      atomicAdd(&s_epl, epl_local);
      atomicMin(&s_rmin, row_min);
      atomicMax(&s_rmax, row_max);
    }
    __syncthreads();
    if (threadIdx.x == 0 && s_epl) {
      atomicAdd((unsigned long long*)&ctr->epl, (unsigned long long)s_epl);
      // rows of a tile are ordered like global rows, so the extreme local rows give the extreme global rows
      const i64 gmin = (z0 + (s_rmin >> TYL)) * g.sy + (y0 + (s_rmin & (TY - 1)));
      const i64 gmax = (z0 + (s_rmax >> TYL)) * g.sy + (y0 + (s_rmax & (TY - 1)));
      atomicMin((long long*)&ctr->first_row, (long long)gmin);
      atomicMax((long long*)&ctr->last_row, (long long)gmax);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Kernel B1. Unions across y/z tile seams, straight on the global forest L (atomicMin link-to-smaller).
// One CTA per (row, 256-voxel chunk); CTAs of rows that touch no seam exit at once. A warp loads its
// 32-voxel segment of the row P and of the rows U=(y-1,z), D=(y,z-1), UD=(y-1,z-1), DN=(y+1,z-1),
// ballots the straight edges, and applies the same eliminations as kernel A so that only the first
// voxel of every contact patch between two runs touches global memory:
//   - x: an edge is implied by the same edge one voxel to the left when both runs continue
//   - square: (P-D) is implied by (P-U),(U-UD),(D-UD) on z-seams; (P-U) by (P-D),(D-UD),(U-UD) on pure
//     y-seam rows (kernel A keeps P-D there because it does not see across the y seam)
//   - transitive predicates: diagonals only where the straight edges do not already connect
// ---------------------------------------------------------------------------------------------
template <typename T, int MODE, int CONN>
__global__ void __launch_bounds__(256)
k_seam_rows(const T* __restrict__ in, u32* __restrict__ L, Geom g, Edge<T, MODE> E, unsigned nchunks) {
  constexpr bool TRANS = (MODE == MODE_EQ || MODE == MODE_NONZERO);
  constexpr int NR = hood_rows(CONN);
  const unsigned row = blockIdx.x / nchunks;
  const unsigned chunk = blockIdx.x - row * nchunks;
  const unsigned sy = (unsigned)g.sy;
  const unsigned z = row / sy, y = row - z * sy;
  const int ly = y & (g.TY - 1), lz = z & (g.TZ - 1);
  const bool hasz = NR >= 2 && z > 0;
  const bool xR0 = y > 0 && ly == 0;
  const bool xR1 = hasz && lz == 0;
  const bool xR2 = NR >= 4 && hasz && y > 0 && (ly == 0 || lz == 0);
  const bool xR3 = NR >= 4 && hasz && y + 1 < sy && (ly == g.TY - 1 || lz == 0);
  if (!(xR0 || xR1 || xR2 || xR3)) return;

  const int lane = threadIdx.x & 31;
  const i64 x = (i64)chunk * 256 + threadIdx.x;
  if (x - lane >= g.sx) return;
  const bool inx = x < g.sx;
  const i64 baseP = (i64)row * g.sx;

  // a row segment with its left/right neighbours
  struct Seg { T c, l, r; };
  auto load = [&](i64 base, bool want) -> Seg {
    Seg q; q.c = (T)0; q.l = (T)0; q.r = (T)0;
    if (!want) return q;
    q.c = inx ? in[base + x] : (T)0;
    q.l = (T)__shfl_up_sync(CC_FULL, q.c, 1);
    q.r = (T)__shfl_down_sync(CC_FULL, q.c, 1);
    if (lane == 0) q.l = (x > 0 && inx) ? in[base + x - 1] : (T)0;
    if (lane == 31) q.r = (x + 1 < g.sx) ? in[base + x + 1] : (T)0;
    return q;
  };
  const Seg P = load(baseP, true);
  const bool f = E.fg(P.c);
  const u32 F = __ballot_sync(CC_FULL, f);
  if (F == 0) return;
  const bool haveU = y > 0, haveD = hasz, haveUD = hasz && y > 0, haveDN = NR >= 4 && hasz && y + 1 < sy;
  const i64 sxy = g.sx * g.sy;
  const Seg U = load(baseP - g.sx, haveU);
  const Seg D = load(baseP - sxy, haveD);
  const Seg UD = load(baseP - sxy - g.sx, haveUD && (xR2 || xR1 || xR0));
  const Seg DN = load(baseP - sxy + g.sx, haveDN && xR3);

  const int DXL = dir_code(-1, 0, 0);
  const u32 LxP = __ballot_sync(CC_FULL, f && x > 0 && E(P.c, P.l, DXL));
  const u32 LxPs1 = LxP >> 1;
  const u32 B0 = haveU ? __ballot_sync(CC_FULL, f && E(P.c, U.c, dir_code(0, -1, 0))) : 0u;
  const u32 B1 = haveD ? __ballot_sync(CC_FULL, f && E(P.c, D.c, dir_code(0, 0, -1))) : 0u;
  const u32 B1u = haveUD ? __ballot_sync(CC_FULL, E.fg(U.c) && E(U.c, UD.c, dir_code(0, 0, -1))) : 0u;
  const u32 B0d = haveUD ? __ballot_sync(CC_FULL, E.fg(D.c) && E(D.c, UD.c, dir_code(0, -1, 0))) : 0u;
  const u32 pi = (u32)(baseP + x);

  auto do_row = [&](const int r, const Seg& Q, const i64 baseQ, const bool have_b, u32 Bm, const u32 cand_b,
                    const u32 cand_ac, const u32 kill_b) {
    const int dy = row_dy(r), dz = row_dz(r);
    const int dxm = hood_dx(CONN, r);
    if (!have_b) {
      if (cand_b == 0 && cand_ac == 0) return;
      Bm = __ballot_sync(CC_FULL, ((cand_b >> lane) & 1u) && E(P.c, Q.c, dir_code(0, dy, dz)));
    }
    u32 Am = 0, Cm = 0;
    if ((dxm & 5) && cand_ac) {
      const bool ca = (cand_ac >> lane) & 1u;
      Am = __ballot_sync(CC_FULL, ca && x > 0 && E(P.c, Q.l, dir_code(-1, dy, dz)));
      Cm = __ballot_sync(CC_FULL, ca && x + 1 < g.sx && E(P.c, Q.r, dir_code(1, dy, dz)));
    }
    if ((Am | Bm | Cm) == 0) return;
    const u32 LP = __ballot_sync(CC_FULL, E.fg(Q.c) && x > 0 && E(Q.c, Q.l, DXL));
    const u32 Bl = Bm << 1, Br = Bm >> 1;
    const u32 needB = Bm & ~(LxP & LP & Bl) & ~kill_b;
    const u32 needA = Am & ~(Bm & LP) & ~(LxP & Bl);
    const u32 needC = Cm & ~(Bm & (LP >> 1)) & ~(LxPs1 & Br);
    const u32 qi = (u32)(baseQ + x);
    if ((needB >> lane) & 1u) uf_union(L, pi, qi);
    if ((needA >> lane) & 1u) uf_union(L, pi, qi - 1);
    if ((needC >> lane) & 1u) uf_union(L, pi, qi + 1);
  };

  if (xR0) {
    u32 cand_ac = F;
    if constexpr (TRANS) cand_ac = (CONN == 26) ? (F & ~(B0 | B1)) : (F & ~B0);
    const u32 kill = (lz > 0) ? (B1 & B0d & B1u) : 0u;  // pure y-seam row: the z-1 side closes the square
    do_row(0, U, baseP - g.sx, true, B0, 0u, cand_ac, kill);
  }
  if constexpr (NR >= 2) {
    if (xR1) {
      u32 cand_ac = F;
      if constexpr (TRANS) cand_ac = (CONN == 26) ? (F & ~(B0 | B1)) : (F & ~B1);
      do_row(1, D, baseP - sxy, true, B1, 0u, cand_ac, B0 & B1u & B0d);
    }
  }
  if constexpr (NR >= 4) {
    if (xR2) {
      const u32 cand = TRANS ? (F & ~(B0 | B1)) : F;
      do_row(2, UD, baseP - sxy - g.sx, false, 0u, cand, cand, (B0 & B1u) | (B1 & B0d));
    }
    if (xR3) {
      const u32 B0dn = __ballot_sync(CC_FULL, E.fg(DN.c) && E(DN.c, D.c, dir_code(0, -1, 0)));
      const u32 cand = TRANS ? (F & ~B1) : F;
      do_row(3, DN, baseP - sxy + g.sx, false, 0u, cand, cand, B1 & B0dn);
    }
  }
}

// Kernel B2. x-seam slots written by kernel A: slot (row, k) holds the raster index of the own-tile
// root that the halo voxel (x = (k+1)*TX - 1, row) was found connected to, or CC_BG.
template <int TX>
__global__ void __launch_bounds__(256) k_seam_x(const u32* __restrict__ XS, u32* __restrict__ L, Geom g) {
  const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  const i64 nb = g.ntx - 1;
  if (i >= g.sy * g.sz * nb) return;
  const u32 r = XS[i];
  if (r == CC_BG) return;
  const i64 row = i / nb, k = i - row * nb;
  const u32 h = (u32)(row * g.sx + (k + 1) * TX - 1);
  uf_union(L, h, r);
}

// ---------------------------------------------------------------------------------------------
// Kernel P. Periodic (torus) wrap edges for 4/8/6-connectivity, delta == 0
// (cc3d.hpp:1048-1073, 1265-1277, 1377-1418; cc3d_binary.hpp:733-, 938-, 1163-1210).
// One thread per voxel of the boundary shell; every backward direction that leaves the volume wraps.
// ---------------------------------------------------------------------------------------------
template <typename T, int MODE, int CONN>
__global__ void __launch_bounds__(256)
k_periodic(const T* __restrict__ in, u32* __restrict__ L, Geom g, Edge<T, MODE> E, int face) {
  // face 0: x == 0 and x == sx-1 planes; face 1: y == 0 plane; face 2: z == 0 plane
  const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  i64 x, y, z;
  if (face == 0) {
    if (i >= 2 * g.sy * g.sz) return;
    const i64 j = i >> 1;
    x = (i & 1) ? g.sx - 1 : 0; y = j % g.sy; z = j / g.sy;
    if ((i & 1) && g.sx == 1) return;
  } else if (face == 1) {
    if (i >= g.sx * g.sz) return;
    x = i % g.sx; y = 0; z = i / g.sx;
  } else {
    if (i >= g.sx * g.sy) return;
    x = i % g.sx; y = i / g.sx; z = 0;
  }
  const i64 pi = (z * g.sy + y) * g.sx + x;
  const T v = in[pi];
  if (!E.fg(v)) return;
  constexpr int NDIR = (CONN == 4) ? 2 : (CONN == 8 ? 4 : 3);
  const int D4[2][3] = {{-1, 0, 0}, {0, -1, 0}};
  const int D8[4][3] = {{-1, 0, 0}, {0, -1, 0}, {-1, -1, 0}, {1, -1, 0}};
  const int D6[3][3] = {{-1, 0, 0}, {0, -1, 0}, {0, 0, -1}};
#pragma unroll
  for (int k = 0; k < NDIR; k++) {
    const int dx = CONN == 4 ? D4[k][0] : (CONN == 8 ? D8[k][0] : D6[k][0]);
    const int dy = CONN == 4 ? D4[k][1] : (CONN == 8 ? D8[k][1] : D6[k][1]);
    const int dz = CONN == 4 ? D4[k][2] : (CONN == 8 ? D8[k][2] : D6[k][2]);
    i64 x2 = x + dx, y2 = y + dy, z2 = z + dz;
    if (x2 >= 0 && x2 < g.sx && y2 >= 0 && z2 >= 0) continue;  // interior edge
    x2 = (x2 + g.sx) % g.sx; y2 = (y2 + g.sy) % g.sy; z2 = (z2 + g.sz) % g.sz;
    const i64 qi = (z2 * g.sy + y2) * g.sx + x2;
    if (qi == pi) continue;
    const T q = in[qi];
    if (E(v, q, 0)) uf_union(L, (u32)pi, (u32)qi);
  }
}
