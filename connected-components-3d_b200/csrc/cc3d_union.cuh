// cc3d_union.cuh — kernels B1/B2 (word-parallel edge elimination + unions) and kernel P (periodic wrap).
// See cc3d_common.cuh for the pipeline.
#pragma once
#include "cc3d_common.cuh"

// A face-bitmap word seen from the 32 voxels of word w: c = bit at x, l = bit at x-1, r = bit at x+1.
struct S3 { u32 c, l, r; };
// The four faces of one row around word w.
struct R3 { S3 F, X, Y, Z; };

// ---------------------------------------------------------------------------------------------
// Edge enumeration for one bitmap word (32 voxels p = (x,y,z) of row P); rows U=(y-1,z) D=(y,z-1)
// UD=(y-1,z-1) DN=(y+1,z-1) V=(y+1,z). Calls emit(b, gid_p, rowQ, xq) for every edge that has to be
// united: p = bit b of the word (run id gid_p), q = voxel xq of row rowQ.
//
// Straight edges (Y: P-U, Z: P-D), dropped when kept edges imply them:
//   x rule:  the edge at x follows from the edge at x-1 when both rows are x-linked from x-1 to x;
//   square:  P-D follows from P-U, U-UD, D-UD at the same x.
//   Every justification refers to straight edges at a smaller (y, x): the kept set spans the same sets.
//
// Diagonal edges p-q (8/18/26). "Between" p and q lie the voxels that are face neighbours of one end
// and neighbours of the other (2 for an in-plane or 18-type diagonal, 6 for a corner diagonal).
//   transitive predicates (EQ, NONZERO): if a voxel between them belongs to the same object it is
//     joined to both ends by edges of a lower class (face < 18-type < corner), so the diagonal is
//     redundant. "Belongs to the same object" shows in the face bitmaps as a link to p or to q, so
//     the candidates are: q foreground and none of those links set.
//   DELTA: the diagonal is redundant when one of the monotone face paths from p to q is fully linked.
// Candidates (a few per thousand voxels on label volumes) load their two voxel values and test
// the predicate itself; NONZERO needs no load. MODE_MASK takes the diagonals from the A0/C0 planes.
// ---------------------------------------------------------------------------------------------
template <typename T, int MODE, int CONN, typename EMIT>
__device__ __forceinline__ void for_each_edge(const T* __restrict__ in, const u32* __restrict__ M, const Geom& g,
                                              const Edge<T, MODE>& E, const u32 i, EMIT&& emit) {
  constexpr int NR = hood_rows(CONN);
  constexpr bool DIAG0 = CONN == 8 || CONN == 18 || CONN == 26;
  constexpr bool DIAGZ = CONN == 18 || CONN == 26;
  constexpr bool CORNER = CONN == 26;
  constexpr bool TRANS = (MODE == MODE_EQ || MODE == MODE_NONZERO);
  const Q4 P = ldq(M, i);
  if (P.F == 0) return;
  const u32 W = (u32)g.W, sy = (u32)g.sy, sx = (u32)g.sx;
  const u32 row = i / W, w = i - row * W;
  const u32 z = row / sy, y = row - z * sy;
  const u32 x0 = w << 5;
  const bool hasL = w > 0, hasR = w + 1 < W;
  const bool hasU = y > 0, hasD = NR >= 2 && z > 0, hasV = y + 1 < sy;
  const u32 WS = W * sy;                  // words per plane
  const Q4 Z4 = {0u, 0u, 0u, 0u};

  // faces of the words left and right of word j (0 outside the row)
  auto sh3 = [](u32 c, u32 lw, u32 rw) -> S3 { S3 s; s.c = c; s.l = (c << 1) | (lw >> 31); s.r = (c >> 1) | (rw << 31); return s; };
  auto row3 = [&](u32 j) -> R3 {
    const Q4 c = ldq(M, j);
    const Q4 l = hasL ? ldq(M, j - 1) : Z4;
    const Q4 r = hasR ? ldq(M, j + 1) : Z4;
    R3 o;
    o.F = sh3(c.F, l.F, r.F); o.X = sh3(c.X, l.X, r.X); o.Y = sh3(c.Y, l.Y, r.Y); o.Z = sh3(c.Z, l.Z, r.Z);
    return o;
  };

  const u32 RSp = __ldg(M + g.offRS + i);
  const u32 Sp = P.F & ~P.X;
  auto pid = [&](int b) -> u32 { return RSp + __popc(Sp & (CC_FULL >> (31 - b))) - 1u; };
  auto joined = [&](int b, u32 rowQ, u32 xq) -> bool {   // value test of a diagonal candidate
    if constexpr (MODE == MODE_NONZERO || MODE == MODE_MASK) return true;
    else return E(in[row * sx + x0 + b], in[rowQ * sx + xq]);
  };
  auto diag = [&](u32 cand, u32 rowQ, int dx) {
    while (cand) {
      const int b = __ffs(cand) - 1; cand &= cand - 1;
      if (joined(b, rowQ, x0 + b + dx)) emit(b, pid(b), rowQ, x0 + b + dx);
    }
  };
  auto straight = [&](u32 need, u32 rowQ) {
    while (need) { const int b = __ffs(need) - 1; need &= need - 1; emit(b, pid(b), rowQ, x0 + b); }
  };

  // ---- straight edges ----
  const Q4 Pl = hasL ? ldq(M, i - 1) : Z4;
  const Q4 U = hasU ? ldq(M, i - W) : Z4;
  const Q4 D = hasD ? ldq(M, i - WS) : Z4;
  const u32 Yl = (P.Y << 1) | (Pl.Y >> 31), Zl = (P.Z << 1) | (Pl.Z >> 31);
  if (hasU) straight(P.Y & ~(P.X & U.X & Yl), row - 1);
  if (hasD) straight(P.Z & ~(P.X & D.X & Zl) & ~(P.Y & U.Z & D.Y), row - sy);

  // ---- diagonal edges ----
  if constexpr (MODE == MODE_MASK) {
    if (hasU) {
      diag(__ldg(M + g.offA0 + i), row - 1, -1);
      diag(__ldg(M + g.offC0 + i), row - 1, +1);
    }
    return;
  }
  if constexpr (!DIAG0) return;
  // p-side precondition of every transitive candidate: no Y link (in-plane) / no Z link (z-1 plane)
  const bool want0 = hasU && (!TRANS || (P.F & ~P.Y));
  const bool wantz = DIAGZ && hasD && (!TRANS || (P.F & ~P.Z));
  if (!want0 && !wantz) return;
  const Q4 Pr = hasR ? ldq(M, i + 1) : Z4;
  const S3 Xp = sh3(P.X, Pl.X, Pr.X), Yp = sh3(P.Y, Pl.Y, Pr.Y), Zp = sh3(P.Z, Pl.Z, Pr.Z);
  const u32 Fp = P.F;
  R3 RU;
  RU.F = RU.X = RU.Y = RU.Z = S3{0u, 0u, 0u};
  if (hasU) RU = row3(i - W);
  if (want0) {
    u32 A0, C0;
    if constexpr (TRANS) {
      A0 = Fp & RU.F.l & ~(Xp.c | Yp.c | Yp.l | RU.X.c);
      C0 = Fp & RU.F.r & ~(Xp.r | Yp.r | Yp.c | RU.X.r);
    } else {
      A0 = Fp & RU.F.l & ~(Xp.c & Yp.l) & ~(Yp.c & RU.X.c);
      C0 = Fp & RU.F.r & ~(Xp.r & Yp.r) & ~(Yp.c & RU.X.r);
    }
    diag(A0, row - 1, -1);
    diag(C0, row - 1, +1);
  }
  if constexpr (DIAGZ) {
    if (wantz) {
      const u32 iD = i - WS;
      const R3 RD = row3(iD);
      const S3 XD = RD.X, YD = RD.Y, ZU = RU.Z, XU = RU.X;
      // (dy=0, dz=-1): A1, C1
      u32 A1, C1;
      if constexpr (TRANS) {
        A1 = Fp & RD.F.l & ~(Xp.c | Zp.l | Zp.c | XD.c);
        C1 = Fp & RD.F.r & ~(Xp.r | Zp.r | Zp.c | XD.r);
      } else {
        A1 = Fp & RD.F.l & ~(Xp.c & Zp.l) & ~(Zp.c & XD.c);
        C1 = Fp & RD.F.r & ~(Xp.r & Zp.r) & ~(Zp.c & XD.r);
      }
      diag(A1, row - sy, -1);
      diag(C1, row - sy, +1);
      // (dy=-1, dz=-1): B2, A2, C2
      if (hasU && (!TRANS || (Fp & ~(Yp.c | Zp.c)))) {
        const R3 RUD = row3(iD - W);
        const S3 FUD = RUD.F, XUD = RUD.X;
        u32 B2;
        if constexpr (TRANS) B2 = Fp & FUD.c & ~(Yp.c | ZU.c | Zp.c | YD.c);
        else B2 = Fp & FUD.c & ~(Yp.c & ZU.c) & ~(Zp.c & YD.c);
        diag(B2, row - sy - 1, 0);
        if constexpr (CORNER) {
          u32 A2, C2;
          if constexpr (TRANS) {
            A2 = Fp & FUD.l & ~(Xp.c | Yp.c | Zp.c | XUD.c | YD.l | ZU.l);
            C2 = Fp & FUD.r & ~(Xp.r | Yp.c | Zp.c | XUD.r | YD.r | ZU.r);
          } else {
            A2 = Fp & FUD.l & ~(Xp.c & Yp.l & ZU.l) & ~(Xp.c & Zp.l & YD.l) & ~(Yp.c & XU.c & ZU.l)
                 & ~(Yp.c & ZU.c & XUD.c) & ~(Zp.c & XD.c & YD.l) & ~(Zp.c & YD.c & XUD.c);
            C2 = Fp & FUD.r & ~(Xp.r & Yp.r & ZU.r) & ~(Xp.r & Zp.r & YD.r) & ~(Yp.c & XU.r & ZU.r)
                 & ~(Yp.c & ZU.c & XUD.r) & ~(Zp.c & XD.r & YD.r) & ~(Zp.c & YD.c & XUD.r);
          }
          diag(A2, row - sy - 1, -1);
          diag(C2, row - sy - 1, +1);
        }
      }
      // (dy=+1, dz=-1): B3, A3, C3
      if (hasV) {
        const R3 RV = row3(i + W);
        const S3 YV = RV.Y, ZV = RV.Z, XV = RV.X;
        if (!TRANS || (Fp & ~(Zp.c | YV.c))) {
          const R3 RDN = row3(iD + W);
          const S3 FDN = RDN.F, XDN = RDN.X, YDN = RDN.Y;
          u32 B3;
          if constexpr (TRANS) B3 = Fp & FDN.c & ~(YV.c | ZV.c | Zp.c | YDN.c);
          else B3 = Fp & FDN.c & ~(YV.c & ZV.c) & ~(Zp.c & YDN.c);
          diag(B3, row - sy + 1, 0);
          if constexpr (CORNER) {
            u32 A3, C3;
            if constexpr (TRANS) {
              A3 = Fp & FDN.l & ~(Xp.c | YV.c | Zp.c | XDN.c | YDN.l | ZV.l);
              C3 = Fp & FDN.r & ~(Xp.r | YV.c | Zp.c | XDN.r | YDN.r | ZV.r);
            } else {
              A3 = Fp & FDN.l & ~(Xp.c & YV.l & ZV.l) & ~(Xp.c & Zp.l & YDN.l) & ~(YV.c & XV.c & ZV.l)
                   & ~(YV.c & ZV.c & XDN.c) & ~(Zp.c & XD.c & YDN.l) & ~(Zp.c & YDN.c & XDN.c);
              C3 = Fp & FDN.r & ~(Xp.r & YV.r & ZV.r) & ~(Xp.r & Zp.r & YDN.r) & ~(YV.c & XV.r & ZV.r)
                   & ~(YV.c & ZV.c & XDN.r) & ~(Zp.c & XD.r & YDN.r) & ~(Zp.c & YDN.c & XDN.r);
            }
            diag(A3, row - sy + 1, -1);
            diag(C3, row - sy + 1, +1);
          }
        }
      }
    }
  }
}

// Union tiles: 2^tw words x 2^ty rows x 2^tz planes = CC_TILE_WORDS words. A run belongs to the tile its
// first voxel lies in; an edge is tile-local when both of its runs belong to the same tile.
struct TilePos { u32 tx, ty, tz; };
__device__ __forceinline__ TilePos tile_of(const Geom& g, u32 w, u32 y, u32 z) {
  TilePos t; t.tx = w >> g.tw; t.ty = y >> g.ty; t.tz = z >> g.tz; return t;
}

// ---------------------------------------------------------------------------------------------
// Kernel B1. One CTA per union tile, union-find in shared memory. A tile row segment (2^tw words of one
// row) owns the contiguous run ids [RS[first word], RS[first word] + cap), so local node = segment *
// cap + (run id - first id of the segment) keeps the raster order of the runs (link-to-smaller stays
// valid). Tile-local edges are united in shared memory; every run of the tile then gets L[run] = run id
// of its tile root. Edges that leave the tile are left to kernel B2.
// ---------------------------------------------------------------------------------------------
template <typename T, int MODE, int CONN>
__global__ void __launch_bounds__(256)
k_union_tile(const T* __restrict__ in, const u32* __restrict__ M, u32* __restrict__ L, Geom g, Edge<T, MODE> E,
             u32 ntx, u32 nty) {
  extern __shared__ u32 smem_u32[];
  u32* lab = smem_u32;                          // [CC_TILE_NODES]
  u32* segRS = smem_u32 + CC_TILE_NODES;        // [CC_TILE_WORDS] first run id of every row segment
  const u32 W = (u32)g.W, sy = (u32)g.sy, sz = (u32)g.sz;
  const u32 TW = 1u << g.tw, TY = 1u << g.ty;
  const u32 nseg = CC_TILE_WORDS >> g.tw;        // TY * TZ
  const u32 capl = g.tw + 5;                     // log2(runs a segment can hold)
  u32 t = blockIdx.x;
  const u32 bx = t % ntx; t /= ntx;
  const u32 by = t % nty;
  const u32 bz = t / nty;
  const u32 w0 = bx << g.tw, y0 = by << g.ty, z0 = bz << g.tz;

  for (u32 r = threadIdx.x; r < nseg; r += blockDim.x) {
    const u32 y = y0 + (r & (TY - 1)), z = z0 + (r >> g.ty);
    segRS[r] = (y < sy && z < sz) ? __ldg(M + g.offRS + (z * sy + y) * W + w0) : 0xFFFFFFFFu;
  }
  for (u32 k = threadIdx.x; k < CC_TILE_NODES; k += blockDim.x) lab[k] = k;
  __syncthreads();

#pragma unroll 1
  for (u32 q = threadIdx.x; q < CC_TILE_WORDS; q += blockDim.x) {
    const u32 wx = q & (TW - 1), r = q >> g.tw;
    const u32 w = w0 + wx, y = y0 + (r & (TY - 1)), z = z0 + (r >> g.ty);
    if (w >= W || y >= sy || z >= sz) continue;
    const u32 i = (z * sy + y) * W + w;
    const u32 base = segRS[r];
    for_each_edge<T, MODE, CONN>(in, M, g, E, i, [&](int b, u32 gp, u32 rowQ, u32 xq) {
      if (gp < base) return;                                    // p's run started left of the tile
      const u32 zq = rowQ / sy, yq = rowQ - zq * sy;
      if ((xq >> 5) >> g.tw != bx || yq >> g.ty != by || zq >> g.tz != bz) return;
      const u32 rq = ((zq - z0) << g.ty) + (yq - y0);
      const u32 gq = run_id(M, g, rowQ * W, xq);
      if (gq < segRS[rq]) return;
      uf_union_h(lab, (r << capl) + (gp - base), (rq << capl) + (gq - segRS[rq]));
    });
  }
  __syncthreads();

  // flatten: every run that starts in the tile -> run id of its tile root
#pragma unroll 1
  for (u32 q = threadIdx.x; q < CC_TILE_WORDS; q += blockDim.x) {
    const u32 wx = q & (TW - 1), r = q >> g.tw;
    const u32 w = w0 + wx, y = y0 + (r & (TY - 1)), z = z0 + (r >> g.ty);
    if (w >= W || y >= sy || z >= sz) continue;
    const u32 i = (z * sy + y) * W + w;
    const uint2 fx = __ldg(reinterpret_cast<const uint2*>(M) + 2 * (size_t)i);
    const int n = __popc(fx.x & ~fx.y);
    if (n == 0) continue;
    const u32 g0 = __ldg(M + g.offRS + i);
    const u32 base = segRS[r];
    for (int k = 0; k < n; k++) {
      u32 l = (r << capl) + (g0 + k - base), p;
      while ((p = lab[l]) != l) l = p;
      L[g0 + k] = segRS[l >> capl] + (l & ((1u << capl) - 1u));
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Kernel B2. One thread per bitmap word: the edges that are not tile-local (same test as B1) are united
// on the global forest L (atomicMin link-to-smaller with path halving).
// ---------------------------------------------------------------------------------------------
template <typename T, int MODE, int CONN>
__global__ void __launch_bounds__(256)
k_union_global(const T* __restrict__ in, const u32* __restrict__ M, u32* __restrict__ L, Geom g, Edge<T, MODE> E) {
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (u32)g.nwords) return;
  const u32 W = (u32)g.W, sy = (u32)g.sy;
  const u32 row = i / W, w = i - row * W;
  const u32 z = row / sy, y = row - z * sy;
  const u32 bx = w >> g.tw, by = y >> g.ty, bz = z >> g.tz;
  const u32 w0 = bx << g.tw;
  u32 base = 0xFFFFFFFFu;   // first run id of p's tile row segment, loaded on demand
  for_each_edge<T, MODE, CONN>(in, M, g, E, i, [&](int b, u32 gp, u32 rowQ, u32 xq) {
    const u32 gq = run_id(M, g, rowQ * W, xq);
    const u32 zq = rowQ / sy, yq = rowQ - zq * sy;
    if ((xq >> 5) >> g.tw == bx && yq >> g.ty == by && zq >> g.tz == bz) {
      if (base == 0xFFFFFFFFu) base = __ldg(M + g.offRS + row * W + w0);
      if (gp >= base && gq >= __ldg(M + g.offRS + rowQ * W + w0)) return;   // tile-local: done by B1
    }
    uf_union_h(L, gp, gq);
  });
}

// ---------------------------------------------------------------------------------------------
// Kernel P. Periodic (torus) wrap edges for 4/8/6-connectivity, delta == 0
// (cc3d.hpp:1048-1073, 1265-1277, 1377-1418; cc3d_binary.hpp:733-, 938-, 1163-1210).
// One thread per voxel of the boundary shell; every backward direction that leaves the volume wraps.
// ---------------------------------------------------------------------------------------------
template <typename T, int MODE, int CONN>
__global__ void __launch_bounds__(256)
k_periodic(const T* __restrict__ in, const u32* __restrict__ M, u32* __restrict__ L, Geom g, Edge<T, MODE> E, int face) {
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
    if (E(v, q))
      uf_union_h(L, run_id(M, g, (u32)((z * g.sy + y) * g.W), (u32)x), run_id(M, g, (u32)((z2 * g.sy + y2) * g.W), (u32)x2));
  }
}
