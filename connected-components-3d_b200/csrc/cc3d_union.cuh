// cc3d_union.cuh — kernel B (word-parallel edge elimination + unions) and kernel P (periodic wrap).
// See cc3d_common.cuh for the pipeline.
#pragma once
#include "cc3d_common.cuh"

// A face-bitmap word seen from the 32 voxels of word w: c = bit at x, l = bit at x-1, r = bit at x+1.
struct S3 { u32 c, l, r; };

// ---------------------------------------------------------------------------------------------
// Kernel B. One thread per bitmap word (32 voxels p = (x,y,z) of row P); rows U=(y-1,z) D=(y,z-1)
// UD=(y-1,z-1) DN=(y+1,z-1) V=(y+1,z).
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
template <typename T, int MODE, int CONN>
__global__ void __launch_bounds__(256)
k_union(const T* __restrict__ in, const u32* __restrict__ M, u32* __restrict__ L, Geom g, Edge<T, MODE> E) {
  constexpr int NR = hood_rows(CONN);
  constexpr bool DIAG0 = CONN == 8 || CONN == 18 || CONN == 26;
  constexpr bool DIAGZ = CONN == 18 || CONN == 26;
  constexpr bool CORNER = CONN == 26;
  constexpr bool TRANS = (MODE == MODE_EQ || MODE == MODE_NONZERO);
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (u32)g.nwords) return;
  const size_t nw = (size_t)g.nwords;
  const u32* PF = M + PL_F * nw;
  const u32* PX = M + PL_X * nw;
  const u32* PY = M + PL_Y * nw;
  const u32* PZ = M + PL_Z * nw;
  const u32 Fp = __ldg(PF + i);
  if (Fp == 0) return;
  const u32 W = (u32)g.W, sy = (u32)g.sy, sx = (u32)g.sx;
  const u32 row = i / W, w = i - row * W;
  const u32 z = row / sy, y = row - z * sy;
  const u32 x0 = w << 5;
  const bool hasL = w > 0, hasR = w + 1 < W;
  const bool hasU = y > 0, hasD = NR >= 2 && z > 0, hasV = y + 1 < sy;
  const u32 jP = row * W;                 // first word of row P
  const u32 WS = W * sy;                  // words per plane

  auto ld3 = [&](const u32* P, u32 j) -> S3 {
    S3 s;
    s.c = __ldg(P + j);
    s.l = (s.c << 1) | (hasL ? (__ldg(P + j - 1) >> 31) : 0u);
    s.r = (s.c >> 1) | (hasR ? (__ldg(P + j + 1) << 31) : 0u);
    return s;
  };
  const S3 Z3 = {0u, 0u, 0u};

  const S3 Xp = ld3(PX, i);
  const u32 RSp = __ldg(M + PL_RS * nw + i);
  const u32 Sp = Fp & ~Xp.c;
  auto pid = [&](int b) -> u32 { return RSp + __popc(Sp & (CC_FULL >> (31 - b))) - 1u; };
  // union of p = bit b of this word with voxel xq of the row whose first word is jQ
  auto unite = [&](int b, u32 jQ, u32 xq) { uf_union_h(L, pid(b), run_id(M, g, jQ, xq)); };
  // value test of a diagonal candidate
  auto joined = [&](int b, u32 rowQ, u32 xq) -> bool {
    if constexpr (MODE == MODE_NONZERO || MODE == MODE_MASK) return true;
    else return E(in[(size_t)row * sx + x0 + b], in[(size_t)rowQ * sx + xq]);
  };
  auto diag = [&](u32 cand, u32 rowQ, int dx) {
    while (cand) {
      const int b = __ffs(cand) - 1; cand &= cand - 1;
      if (joined(b, rowQ, x0 + b + dx)) unite(b, rowQ * W, x0 + b + dx);
    }
  };

  // ---- straight edges ----
  S3 Yp = Z3, Zp = Z3;
  u32 XUc = 0, ZUc = 0, YDc = 0, XDc = 0;
  if (hasU) {
    Yp = ld3(PY, i);
    XUc = __ldg(PX + i - W);
    u32 need = Yp.c & ~(Xp.c & XUc & Yp.l);
    while (need) { const int b = __ffs(need) - 1; need &= need - 1; unite(b, jP - W, x0 + b); }
  }
  if (hasD) {
    Zp = ld3(PZ, i);
    XDc = __ldg(PX + i - WS);
    if (hasU) { ZUc = __ldg(PZ + i - W); YDc = __ldg(PY + i - WS); }
    u32 need = Zp.c & ~(Xp.c & XDc & Zp.l) & ~(Yp.c & ZUc & YDc);
    while (need) { const int b = __ffs(need) - 1; need &= need - 1; unite(b, jP - WS, x0 + b); }
  }

  // ---- diagonal edges ----
  if constexpr (MODE == MODE_MASK) {
    if (hasU) {
      diag(__ldg(M + PL_A0 * nw + i), row - 1, -1);
      diag(__ldg(M + PL_C0 * nw + i), row - 1, +1);
    }
    return;
  }
  if constexpr (DIAG0) {
    if (hasU && (!TRANS || (Fp & ~Yp.c))) {
      const S3 FU = ld3(PF, i - W);
      const S3 XU = ld3(PX, i - W);
      u32 A0, C0;
      if constexpr (TRANS) {
        A0 = Fp & FU.l & ~(Xp.c | Yp.c | Yp.l | XU.c);
        C0 = Fp & FU.r & ~(Xp.r | Yp.r | Yp.c | XU.r);
      } else {
        A0 = Fp & FU.l & ~(Xp.c & Yp.l) & ~(Yp.c & XU.c);
        C0 = Fp & FU.r & ~(Xp.r & Yp.r) & ~(Yp.c & XU.r);
      }
      diag(A0, row - 1, -1);
      diag(C0, row - 1, +1);
    }
  }
  if constexpr (DIAGZ) {
    if (hasD && (!TRANS || (Fp & ~Zp.c))) {
      const u32 iD = i - WS;
      // (dy=0, dz=-1): A1, C1
      {
        const S3 FD = ld3(PF, iD);
        const S3 XD = ld3(PX, iD);
        u32 A1, C1;
        if constexpr (TRANS) {
          A1 = Fp & FD.l & ~(Xp.c | Zp.l | Zp.c | XD.c);
          C1 = Fp & FD.r & ~(Xp.r | Zp.r | Zp.c | XD.r);
        } else {
          A1 = Fp & FD.l & ~(Xp.c & Zp.l) & ~(Zp.c & XD.c);
          C1 = Fp & FD.r & ~(Xp.r & Zp.r) & ~(Zp.c & XD.r);
        }
        diag(A1, row - sy, -1);
        diag(C1, row - sy, +1);
        // (dy=-1, dz=-1): B2, A2, C2
        if (hasU) {
          const S3 FUD = ld3(PF, iD - W);
          const S3 ZU = ld3(PZ, i - W);
          const S3 YD = ld3(PY, iD);
          u32 B2;
          if constexpr (TRANS) B2 = Fp & FUD.c & ~(Yp.c | ZU.c | Zp.c | YD.c);
          else B2 = Fp & FUD.c & ~(Yp.c & ZU.c) & ~(Zp.c & YD.c);
          diag(B2, row - sy - 1, 0);
          if constexpr (CORNER) {
            const S3 XUD = ld3(PX, iD - W);
            u32 A2, C2;
            if constexpr (TRANS) {
              A2 = Fp & FUD.l & ~(Xp.c | Yp.c | Zp.c | XUD.c | YD.l | ZU.l);
              C2 = Fp & FUD.r & ~(Xp.r | Yp.c | Zp.c | XUD.r | YD.r | ZU.r);
            } else {
              const S3 XU = ld3(PX, i - W);
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
          const S3 FDN = ld3(PF, iD + W);
          const S3 YV = ld3(PY, i + W);
          const S3 ZV = ld3(PZ, i + W);
          const S3 YDN = ld3(PY, iD + W);
          u32 B3;
          if constexpr (TRANS) B3 = Fp & FDN.c & ~(YV.c | ZV.c | Zp.c | YDN.c);
          else B3 = Fp & FDN.c & ~(YV.c & ZV.c) & ~(Zp.c & YDN.c);
          diag(B3, row - sy + 1, 0);
          if constexpr (CORNER) {
            const S3 XDN = ld3(PX, iD + W);
            u32 A3, C3;
            if constexpr (TRANS) {
              A3 = Fp & FDN.l & ~(Xp.c | YV.c | Zp.c | XDN.c | YDN.l | ZV.l);
              C3 = Fp & FDN.r & ~(Xp.r | YV.c | Zp.c | XDN.r | YDN.r | ZV.r);
            } else {
              const S3 XV = ld3(PX, i + W);
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
