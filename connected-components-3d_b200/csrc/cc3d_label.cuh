// cc3d_label.cuh — labelling kernels A (tile), B1/B2 (seams), P (periodic). See cc3d_common.cuh.
#pragma once
#include "cc3d_common.cuh"

#define CC_TILE_THREADS 256

// Shared-memory bytes kernel A needs for a TXxTYxTZ tile of T.
template <typename T, int TX>
__host__ __device__ constexpr size_t tile_smem_bytes(int TY, int TZ) {
  // slab[TILE + rows] + slink[rows*NSEG] (u32) + sval[rows*(TX+2)] (T), T region 8-byte aligned
  return (((size_t)TY * TZ * TX + (size_t)TY * TZ + (size_t)TY * TZ * (TX / 32)) * 4 + 7) / 8 * 8 +
         (size_t)TY * TZ * (TX + 2) * sizeof(T);
}

// ---------------------------------------------------------------------------------------------
// Kernel A. One CTA labels one tile entirely in shared memory.
//  phase 0: coalesced load of the tile (+ the x0-1 halo column) into sval; voxels outside the volume = 0
//  phase 1: x-runs by warp ballot: every voxel starts out pointing at the first voxel of its run
//           inside its 32-wide segment (no atomics); epl transitions are counted here too
//  phase 2: for each neighbour row of the backward neighbourhood, ballot the three candidate edges,
//           drop the ones already implied by a neighbouring lane (same pair of runs), and union the rest
//           with shared-memory atomicMin; the halo column joins through the same union-find with
//           indices >= TILE so that it can never become a root
//  phase 3: flatten, write L (global raster index of the local root), the local-root bitmap word and
//           the x-seam slot (own-tile root that the halo voxel to the left belongs to)
// ---------------------------------------------------------------------------------------------
template <typename T, int MODE, int CONN, int TX>
__global__ void __launch_bounds__(CC_TILE_THREADS)
k_tile_label(const T* __restrict__ in, u32* __restrict__ L, u32* __restrict__ LR, u32* __restrict__ XS,
             Geom g, Edge<T, MODE> E, Counters* __restrict__ ctr) {
  constexpr int NSEG = TX / 32;
  constexpr int SVX = TX + 2;
  const int TY = g.TY, TZ = g.TZ;
  const int rows = TY * TZ;
  const int TILE = rows * TX;

  extern __shared__ __align__(16) unsigned char smem_raw[];
  u32* slab = reinterpret_cast<u32*>(smem_raw);          // [TILE + rows]
  u32* slink = slab + TILE + rows;                       // [rows * NSEG]
  T* sval = reinterpret_cast<T*>(smem_raw + (((size_t)(TILE + rows + rows * NSEG) * 4 + 7) / 8 * 8));  // [rows][SVX]

  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  constexpr int NWARPS = CC_TILE_THREADS / 32;

  // tile origin
  i64 t = blockIdx.x;
  const i64 bx = t % g.ntx; t /= g.ntx;
  const i64 by = t % g.nty;
  const i64 bz = t / g.nty;
  const i64 x0 = bx * TX, y0 = by * TY, z0 = bz * TZ;

  // ---- phase 0: load ----
  for (int s = warp; s < rows * NSEG; s += NWARPS) {
    const int row = s / NSEG, k = s - row * NSEG;
    const int lz = row / TY, ly = row - lz * TY;
    const i64 gx = x0 + k * 32 + lane, gy = y0 + ly, gz = z0 + lz;
    T v = (T)0;
    if (gx < g.sx && gy < g.sy && gz < g.sz) v = in[(gz * g.sy + gy) * g.sx + gx];
    sval[row * SVX + 1 + k * 32 + lane] = v;
  }
  for (int row = threadIdx.x; row < rows; row += CC_TILE_THREADS) {
    const int lz = row / TY, ly = row - lz * TY;
    const i64 gy = y0 + ly, gz = z0 + lz;
    T v = (T)0;
    if (x0 > 0 && gy < g.sy && gz < g.sz) v = in[(gz * g.sy + gy) * g.sx + x0 - 1];
    sval[row * SVX] = v;
    sval[row * SVX + TX + 1] = (T)0;
  }
  __syncthreads();

  // ---- phase 1: x-runs ----
  u32 epl_local = 0;
  i64 row_min = INT64_MAX, row_max = -1;
  for (int s = warp; s < rows * NSEG; s += NWARPS) {
    const int row = s / NSEG, k = s - row * NSEG;
    const int lx = k * 32 + lane;
    const T v = sval[row * SVX + 1 + lx];
    const T vl = sval[row * SVX + lx];
    const bool f = E.fg(v);
    const bool link = (lx > 0) && f && E(v, vl, dir_code(-1, 0, 0));
    const u32 m = __ballot_sync(CC_FULL, link);
    u32 below = (~m & (CC_FULL >> (31 - lane))) | 1u;
    const int start = 31 - __clz(below);
    slab[row * TX + lx] = f ? (u32)(row * TX + k * 32 + start) : CC_BG;
    if (lane == 0) slink[s] = m;
    if constexpr (MODE != MODE_MASK) {
      // cc3d.hpp:300-303: (row[0] != 0) + sum_x (row[x] != row[x-1] && row[x] != 0)
      const bool tr = f && ((x0 + lx == 0) || (v != vl));
      const u32 tm = __ballot_sync(CC_FULL, tr);
      if (tm) {
        epl_local += __popc(tm);
        const int lz = row / TY, ly = row - lz * TY;
        const i64 grow = (z0 + lz) * g.sy + (y0 + ly);
        row_min = min(row_min, grow);
        row_max = max(row_max, grow);
      }
    }
  }
  // halo column entries of the union-find
  for (int row = threadIdx.x; row < rows; row += CC_TILE_THREADS)
    slab[TILE + row] = E.fg(sval[row * SVX]) ? (u32)(TILE + row) : CC_BG;
  __syncthreads();

  // ---- phase 2: unions ----
  for (int s = warp; s < rows * NSEG; s += NWARPS) {
    const int row = s / NSEG, k = s - row * NSEG;
    const int lz = row / TY, ly = row - lz * TY;
    const int lx = k * 32 + lane;
    const u32 li = row * TX + lx;
    const T v = sval[row * SVX + 1 + lx];
    const u32 m_own = slink[s];
    const u32 m_own_next = (k + 1 < NSEG) ? slink[s + 1] : 0u;
    const u32 LRs1 = (m_own >> 1) | (m_own_next << 31);
    if (k > 0 && lane == 0 && (m_own & 1u)) uf_union(slab, li, li - 1);  // run continues from the previous segment
    if (!__any_sync(CC_FULL, E.fg(v))) continue;
#pragma unroll
    for (int r = 0; r < hood_rows(CONN); r++) {
      constexpr int dummy = 0; (void)dummy;
      const int dy = row_dy(r), dz = row_dz(r);
      const int dxm = hood_dx(CONN, r);
      const int ly2 = ly + dy, lz2 = lz + dz;
      if (ly2 < 0 || ly2 >= TY || lz2 < 0) continue;  // other tile: seam kernel
      const int row2 = lz2 * TY + ly2;
      const T* q = sval + row2 * SVX + 1 + lx;
      const bool f = E.fg(v);
      const bool b = (dxm & 2) && f && E(v, q[0], dir_code(0, dy, dz));
      const bool a = (dxm & 1) && f && lx > 0 && E(v, q[-1], dir_code(-1, dy, dz));
      const bool c = (dxm & 4) && f && lx + 1 < TX && E(v, q[1], dir_code(1, dy, dz));
      const u32 B = __ballot_sync(CC_FULL, b);
      const u32 A = (dxm & 1) ? __ballot_sync(CC_FULL, a) : 0u;
      const u32 C = (dxm & 4) ? __ballot_sync(CC_FULL, c) : 0u;
      if ((A | B | C) == 0) continue;
      const u32 LP = slink[row2 * NSEG + k];
      const u32 LP_next = (k + 1 < NSEG) ? slink[row2 * NSEG + k + 1] : 0u;
      const u32 LPs1 = (LP >> 1) | (LP_next << 31);
      const u32 Bl = B << 1;  // bit j: dx=0 edge exists at x_j - 1 (unknown across the segment start)
      const u32 Br = B >> 1;  // bit j: dx=0 edge exists at x_j + 1
      const u32 needB = B & ~(m_own & LP & Bl);
      const u32 needA = A & ~(B & LP) & ~(m_own & Bl);
      const u32 needC = C & ~(B & LPs1) & ~(LRs1 & Br);
      const u32 qi = row2 * TX + lx;
      if ((needB >> lane) & 1u) uf_union(slab, li, qi);
      if ((needA >> lane) & 1u) uf_union(slab, li, qi - 1);
      if ((needC >> lane) & 1u) uf_union(slab, li, qi + 1);
    }
  }
  // halo column: edges between the halo voxel h=(x0-1,y,z) and own voxels (x0, y+ddy, z+ddz),
  // plus halo-halo links used to drop redundant x-seam slots.
  if (x0 > 0) {
    for (int row = threadIdx.x; row < rows; row += CC_TILE_THREADS) {
      const T hv = sval[row * SVX];
      if (!E.fg(hv)) continue;
      const int lz = row / TY, ly = row - lz * TY;
      const u32 hi = TILE + row;
#pragma unroll
      for (int ddz = -1; ddz <= 1; ddz++) {
#pragma unroll
        for (int ddy = -1; ddy <= 1; ddy++) {
          // is (dx=+-1, ddy, ddz) part of this connectivity?
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
          if (ddz > 0 || (ddz == 0 && ddy >= 0)) e = E(ov, hv, dir_code(-1, -ddy, -ddz));  // own voxel is later
          else e = E(hv, ov, dir_code(1, ddy, ddz));                                        // halo voxel is later
          if (e) uf_union(slab, hi, (u32)(row2 * TX));
        }
      }
      if (ly > 0) {
        const T pv = sval[(row - 1) * SVX];
        if (E.fg(pv) && E(hv, pv, dir_code(0, -1, 0))) uf_union(slab, hi, hi - 1);
      }
      if (CONN != 4 && CONN != 8 && lz > 0) {
        const T pv = sval[(row - TY) * SVX];
        if (E.fg(pv) && E(hv, pv, dir_code(0, 0, -1))) uf_union(slab, hi, hi - TY);
      }
    }
  }
  __syncthreads();

  // ---- phase 3: flatten + write ----
  for (int s = warp; s < rows * NSEG; s += NWARPS) {
    const int row = s / NSEG, k = s - row * NSEG;
    const int lz = row / TY, ly = row - lz * TY;
    const int lx = k * 32 + lane;
    const i64 gx = x0 + lx, gy = y0 + ly, gz = z0 + lz;
    const bool inside = gx < g.sx && gy < g.sy && gz < g.sz;
    const u32 li = row * TX + lx;
    u32 l = slab[li];
    u32 out = CC_BG;
    bool is_root = false;
    if (l != CC_BG) {
      u32 p;
      while ((p = slab[l]) != l) l = p;
      is_root = (l == li);
      const int rrow = l / TX, rlx = l - rrow * TX;
      const int rlz = rrow / TY, rly = rrow - rlz * TY;
      out = (u32)(((z0 + rlz) * g.sy + (y0 + rly)) * g.sx + x0 + rlx);
    }
    const u32 rm = __ballot_sync(CC_FULL, is_root);
    if (inside) L[(gz * g.sy + gy) * g.sx + gx] = out;
    if (lane == 0 && gy < g.sy && gz < g.sz && x0 + k * 32 < g.sx)
      LR[(gz * g.sy + gy) * g.W + ((x0 + k * 32) >> 5)] = rm;
  }
  if (x0 > 0) {
    for (int row = threadIdx.x; row < rows; row += CC_TILE_THREADS) {
      const int lz = row / TY, ly = row - lz * TY;
      const i64 gy = y0 + ly, gz = z0 + lz;
      if (gy >= g.sy || gz >= g.sz) continue;
      u32 out = CC_BG;
      u32 l = slab[TILE + row];
      if (l != CC_BG) {
        // skip when an earlier halo voxel of the same column is linked to this one (both tiles know that link)
        const T hv = sval[row * SVX];
        bool covered = false;
        if (ly > 0) { const T pv = sval[(row - 1) * SVX]; covered = E.fg(pv) && E(hv, pv, dir_code(0, -1, 0)); }
        if (!covered && CONN != 4 && CONN != 8 && lz > 0) {
          const T pv = sval[(row - TY) * SVX]; covered = E.fg(pv) && E(hv, pv, dir_code(0, 0, -1));
        }
        if (!covered) {
          u32 p;
          while ((p = slab[l]) != l) l = p;
          if (l < (u32)TILE) {
            const int rrow = l / TX, rlx = l - rrow * TX;
            const int rlz = rrow / TY, rly = rrow - rlz * TY;
            out = (u32)(((z0 + rlz) * g.sy + (y0 + rly)) * g.sx + x0 + rlx);
          }
        }
      }
      XS[(gz * g.sy + gy) * (g.ntx - 1) + (bx - 1)] = out;
    }
  }

  if constexpr (MODE != MODE_MASK) {
    // block-level reduction of epl and the foreground row range
    __shared__ u32 s_epl;
    __shared__ long long s_rmin, s_rmax;
    if (threadIdx.x == 0) { s_epl = 0; s_rmin = INT64_MAX; s_rmax = -1; }
    __syncthreads();
    if (lane == 0 && epl_local) {
      atomicAdd(&s_epl, epl_local);
      atomicMin(&s_rmin, (long long)row_min);
      atomicMax(&s_rmax, (long long)row_max);
    }
    __syncthreads();
    if (threadIdx.x == 0 && s_epl) {
      atomicAdd((unsigned long long*)&ctr->epl, (unsigned long long)s_epl);
      atomicMin((long long*)&ctr->first_row, s_rmin);
      atomicMax((long long*)&ctr->last_row, s_rmax);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Kernel B1. Unions across y/z tile seams. One warp per (row, 32-voxel x segment); a row only does
// work for those neighbour rows that live in a different (y,z) tile. Same ballot-based redundancy
// elimination as kernel A, unions go to the global forest L with atomicMin.
// ---------------------------------------------------------------------------------------------
template <typename T, int MODE, int CONN>
__global__ void __launch_bounds__(256)
k_seam_rows(const T* __restrict__ in, u32* __restrict__ L, Geom g, Edge<T, MODE> E) {
  const int lane = threadIdx.x & 31;
  const i64 wid = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const i64 nseg = g.W;
  const i64 row = wid / nseg;
  if (row >= g.sy * g.sz) return;
  const i64 seg = wid - row * nseg;
  const i64 z = row / g.sy, y = row - z * g.sy;
  const int ly = (int)(y % g.TY), lz = (int)(z % g.TZ);
  // does any neighbour row fall into another tile?
  bool any = false;
#pragma unroll
  for (int r = 0; r < hood_rows(CONN); r++) {
    const int dy = row_dy(r), dz = row_dz(r);
    const i64 y2 = y + dy, z2 = z + dz;
    if (y2 < 0 || y2 >= g.sy || z2 < 0) continue;
    if (ly + dy < 0 || ly + dy >= g.TY || lz + dz < 0) any = true;
  }
  if (!any) return;

  const i64 x = seg * 32 + lane;
  const bool inx = x < g.sx;
  const i64 base = row * g.sx;
  const T v = inx ? in[base + x] : (T)0;
  T vl = __shfl_up_sync(CC_FULL, v, 1);
  if (lane == 0) vl = (x > 0 && inx) ? in[base + x - 1] : (T)0;
  const bool f = E.fg(v);
  if (!__any_sync(CC_FULL, f)) return;
  const u32 m_own = __ballot_sync(CC_FULL, f && x > 0 && E(v, vl, dir_code(-1, 0, 0)));
  const u32 LRs1 = m_own >> 1;

#pragma unroll
  for (int r = 0; r < hood_rows(CONN); r++) {
    const int dy = row_dy(r), dz = row_dz(r);
    const int dxm = hood_dx(CONN, r);
    const i64 y2 = y + dy, z2 = z + dz;
    if (y2 < 0 || y2 >= g.sy || z2 < 0) continue;
    if (!(ly + dy < 0 || ly + dy >= g.TY || lz + dz < 0)) continue;  // same tile: kernel A did it
    const i64 base2 = (z2 * g.sy + y2) * g.sx;
    const T qb = inx ? in[base2 + x] : (T)0;
    T qa = __shfl_up_sync(CC_FULL, qb, 1);
    T qc = __shfl_down_sync(CC_FULL, qb, 1);
    if (lane == 0) qa = (x > 0 && inx) ? in[base2 + x - 1] : (T)0;
    if (lane == 31) qc = (x + 1 < g.sx) ? in[base2 + x + 1] : (T)0;
    const bool b = (dxm & 2) && f && E(v, qb, dir_code(0, dy, dz));
    const bool a = (dxm & 1) && f && x > 0 && E(v, qa, dir_code(-1, dy, dz));
    const bool c = (dxm & 4) && f && x + 1 < g.sx && E(v, qc, dir_code(1, dy, dz));
    const u32 B = __ballot_sync(CC_FULL, b);
    const u32 A = (dxm & 1) ? __ballot_sync(CC_FULL, a) : 0u;
    const u32 C = (dxm & 4) ? __ballot_sync(CC_FULL, c) : 0u;
    if ((A | B | C) == 0) continue;
    const u32 LP = __ballot_sync(CC_FULL, E.fg(qb) && x > 0 && E(qb, qa, dir_code(-1, 0, 0)));
    const u32 LPs1 = LP >> 1;
    const u32 Bl = B << 1, Br = B >> 1;
    const u32 needB = B & ~(m_own & LP & Bl);
    const u32 needA = A & ~(B & LP) & ~(m_own & Bl);
    const u32 needC = C & ~(B & LPs1) & ~(LRs1 & Br);
    const u32 pi = (u32)(base + x), qi = (u32)(base2 + x);
    if ((needB >> lane) & 1u) uf_union(L, pi, qi);
    if ((needA >> lane) & 1u) uf_union(L, pi, qi - 1);
    if ((needC >> lane) & 1u) uf_union(L, pi, qi + 1);
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
