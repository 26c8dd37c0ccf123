// cc3d_union.cuh — kernels B1/B2 (word-parallel edge elimination, tile-local and global unions) and
// kernel P (periodic wrap).
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
                                              const Edge<T, MODE>& E, const u32 i, const u32 row, const u32 w,
                                              const u32 y, const u32 z, EMIT&& emit) {
  constexpr int NR = hood_rows(CONN);
  constexpr bool DIAG0 = CONN == 8 || CONN == 18 || CONN == 26;
  constexpr bool DIAGZ = CONN == 18 || CONN == 26;
  constexpr bool CORNER = CONN == 26;
  constexpr bool TRANS = (MODE == MODE_EQ || MODE == MODE_NONZERO);
  const Q4 P = ldq(M, i);
  if (P.F == 0) return;
  const u32 W = (u32)g.W, sy = (u32)g.sy, sx = (u32)g.sx;
  const u32 x0 = w << 5;
  const bool hasL = w > 0, hasR = w + 1 < W;
  const bool hasU = y > 0, hasD = NR >= 2 && z > 0, hasV = y + 1 < sy;
  const u32 WS = W * sy;                  // words per plane
  const Q4 Z4 = {0u, 0u, 0u, 0u};
  const u32* __restrict__ RS = M + g.offRS;

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

  const u32 RSp = __ldg(RS + i) - 1u;
  const u32 Sp = P.F & ~P.X;
  auto pid = [&](int b) -> u32 { return RSp + __popc(Sp & (CC_FULL >> (31 - b))); };
  auto joined = [&](int b, u32 rowQ, u32 xq) -> bool {   // value test of a diagonal candidate
    if constexpr (MODE == MODE_NONZERO || MODE == MODE_MASK) return true;
    else return E(in[row * sx + x0 + b], in[rowQ * sx + xq]);
  };
  auto diag = [&](u32 cand, int dy, int dz, int dx) {
    const u32 rowQ = row + dy + dz * (int)sy;
    while (cand) {
      const int b = __ffs(cand) - 1; cand &= cand - 1;
      const u32 xq = x0 + b + dx;
      if (joined(b, rowQ, xq)) emit(pid(b), run_id(M, g, rowQ * W, xq), dy, dz, xq);
    }
  };
  // straight edge to the same x of row Q (faces Qf, run counter RSq = RS[word of Q] - 1)
  auto straight = [&](u32 need, const Q4& Qf, u32 RSq, int dy, int dz) {
    const u32 Sq = Qf.F & ~Qf.X;
    while (need) {
      const int b = __ffs(need) - 1; need &= need - 1;
      const u32 below = CC_FULL >> (31 - b);
      emit(RSp + __popc(Sp & below), RSq + __popc(Sq & below), dy, dz, x0 + b);
    }
  };

  // ---- straight edges ----
  const Q4 Pl = hasL ? ldq(M, i - 1) : Z4;
  const Q4 U = hasU ? ldq(M, i - W) : Z4;
  const Q4 D = hasD ? ldq(M, i - WS) : Z4;
  const u32 Yl = (P.Y << 1) | (Pl.Y >> 31), Zl = (P.Z << 1) | (Pl.Z >> 31);
  if (hasU) {
    const u32 need = P.Y & ~(P.X & U.X & Yl);
    if (need) straight(need, U, __ldg(RS + i - W) - 1u, -1, 0);
  }
  if (hasD) {
    const u32 need = P.Z & ~(P.X & D.X & Zl) & ~(P.Y & U.Z & D.Y);
    if (need) straight(need, D, __ldg(RS + i - WS) - 1u, 0, -1);
  }

  // ---- diagonal edges ----
  if constexpr (MODE == MODE_MASK) {
    if (hasU) {
      diag(__ldg(M + g.offA0 + i), -1, 0, -1);
      diag(__ldg(M + g.offC0 + i), -1, 0, +1);
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
    diag(A0, -1, 0, -1);
    diag(C0, -1, 0, +1);
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
      diag(A1, 0, -1, -1);
      diag(C1, 0, -1, +1);
      // (dy=-1, dz=-1): B2, A2, C2
      if (hasU && (!TRANS || (Fp & ~(Yp.c | Zp.c)))) {
        const R3 RUD = row3(iD - W);
        const S3 FUD = RUD.F, XUD = RUD.X;
        u32 B2;
        if constexpr (TRANS) B2 = Fp & FUD.c & ~(Yp.c | ZU.c | Zp.c | YD.c);
        else B2 = Fp & FUD.c & ~(Yp.c & ZU.c) & ~(Zp.c & YD.c);
        diag(B2, -1, -1, 0);
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
          diag(A2, -1, -1, -1);
          diag(C2, -1, -1, +1);
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
          diag(B3, +1, -1, 0);
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
            diag(A3, +1, -1, -1);
            diag(C3, +1, -1, +1);
          }
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Kernel B1. One CTA per union tile (2^tw words x 2^ty rows x 2^tz planes = CC_TILE_WORDS words).
// A run belongs to the tile its first voxel lies in; an edge is tile-local when both of its runs belong
// to the tile. The runs that start in a tile row segment (2^tw words of one row) have contiguous ids
// [RS[first word], RS[word after the segment]), so an exclusive scan of the segment sizes gives dense
// local ids that keep the raster order of the runs (link-to-smaller stays valid).
//   phase 1: one thread per word enumerates the edges (for_each_edge) and only classifies them:
//            tile-local edges go to a shared-memory queue, the others to a staging buffer
//   phase 2: the queue is worked off by all threads, one edge each (balanced): union-find in shared memory
//   phase 3: every run of the tile gets L[run] = run id of its tile root; the staged edges are appended
//            to the global edge queue GQ for kernel B2
// A tile with more than CC_TILE_LAB runs sends all its edges to B2. If GQ overflows, *ovf is raised
// and kernel B2s redoes every edge on the global forest.
// ---------------------------------------------------------------------------------------------
struct EdgeQueue { u64* q; u32* count; u32* ovf; u32 cap; };

template <typename T, int MODE, int CONN>
__global__ void __launch_bounds__(256)
k_union_tile(const T* __restrict__ in, const u32* __restrict__ M, u32* __restrict__ L, Geom g, Edge<T, MODE> E,
             u32 ntx, u32 nty, EdgeQueue GQ) {
  extern __shared__ __align__(16) u32 smem_u32[];
  u32* lab = smem_u32;                                   // [CC_TILE_LAB]
  u32* lq = lab + CC_TILE_LAB;                           // [CC_TILE_LQ]
  u64* gq = reinterpret_cast<u64*>(lq + CC_TILE_LQ);     // [CC_TILE_GQ]
  u32* segRS = lq + CC_TILE_LQ + 2 * CC_TILE_GQ;         // [CC_TILE_WORDS] first run id of every row segment
  u32* segLB = segRS + CC_TILE_WORDS;                    // [CC_TILE_WORDS] first local id of every row segment
  __shared__ u32 s_ln, s_gn, s_gbase, s_total;
  const u32 W = (u32)g.W, sy = (u32)g.sy, sz = (u32)g.sz;
  const u32 TW = 1u << g.tw, TY = 1u << g.ty;
  const u32 nseg = CC_TILE_WORDS >> g.tw;        // TY * TZ
  u32 t = blockIdx.x;
  const u32 bx = t % ntx; t /= ntx;
  const u32 by = t % nty;
  const u32 bz = t / nty;
  const u32 w0 = bx << g.tw, y0 = by << g.ty, z0 = bz << g.tz;
  const u32 wend = min(w0 + TW, W);
  const u32* __restrict__ RS = M + g.offRS;

  // segment tables (two segments per thread; nseg <= 512)
  {
    u32 c[2];
#pragma unroll
    for (int k = 0; k < 2; k++) {
      const u32 r = 2 * threadIdx.x + k;
      c[k] = 0;
      if (r < nseg) {
        const u32 y = y0 + (r & (TY - 1)), z = z0 + (r >> g.ty);
        u32 first = 0xFFFFFFFFu;
        if (y < sy && z < sz) {
          const u32 j = (z * sy + y) * W;
          first = __ldg(RS + j + w0);
          c[k] = __ldg(RS + j + wend) - first;
        }
        segRS[r] = first;
      }
    }
    if (threadIdx.x == 0) { s_ln = 0; s_gn = 0; }
    u32 total;
    const u32 ex = block_exclusive_scan(c[0] + c[1], &total);
    if (2 * threadIdx.x < nseg) segLB[2 * threadIdx.x] = ex;
    if (2 * threadIdx.x + 1 < nseg) segLB[2 * threadIdx.x + 1] = ex + c[0];
    if (threadIdx.x == 0) s_total = total;
  }
  __syncthreads();
  const u32 total = s_total;
  const bool local_ok = total <= CC_TILE_LAB;
  if (local_ok) for (u32 k = threadIdx.x; k < total; k += blockDim.x) lab[k] = k;
  __syncthreads();

  auto push_global = [&](u32 gp, u32 gq_) {
    const u32 pos = atomicAdd(GQ.count, 1u);
    if (pos < GQ.cap) GQ.q[pos] = (u64)gp | ((u64)gq_ << 32);
    else *GQ.ovf = 1u;
  };

  // ---- phase 1: enumerate + classify ----
#pragma unroll 1
  for (u32 q = threadIdx.x; q < CC_TILE_WORDS; q += blockDim.x) {
    const u32 wx = q & (TW - 1), r = q >> g.tw;
    const int ly = (int)(r & (TY - 1)), lz = (int)(r >> g.ty);
    const u32 w = w0 + wx, y = y0 + ly, z = z0 + lz;
    if (w >= W || y >= sy || z >= sz) continue;
    const u32 row = z * sy + y;
    const u32 baseP = segRS[r], lbP = segLB[r];
    for_each_edge<T, MODE, CONN>(in, M, g, E, row * W + w, row, w, y, z, [&](u32 gp, u32 gq_, int dy, int dz, u32 xq) {
      bool local = local_ok && gp >= baseP;
      u32 rq = 0;
      if (local) {
        const int lyq = ly + dy, lzq = lz + dz;
        local = ((xq >> 5) >> g.tw) == bx && lyq >= 0 && lyq < (int)TY && lzq >= 0;
        if (local) { rq = ((u32)lzq << g.ty) + (u32)lyq; local = gq_ >= segRS[rq]; }
      }
      if (local) {
        const u32 lp = lbP + (gp - baseP), lq_ = segLB[rq] + (gq_ - segRS[rq]);
        const u32 pos = atomicAdd(&s_ln, 1u);
        if (pos < CC_TILE_LQ) lq[pos] = lp | (lq_ << 16);
        else uf_union_h(lab, lp, lq_);
      } else {
        const u32 pos = atomicAdd(&s_gn, 1u);
        if (pos < CC_TILE_GQ) gq[pos] = (u64)gp | ((u64)gq_ << 32);
        else push_global(gp, gq_);
      }
    });
  }
  __syncthreads();

  // ---- phase 2: tile-local unions, one queued edge per thread and step ----
  const u32 ln = min(s_ln, (u32)CC_TILE_LQ), gn = min(s_gn, (u32)CC_TILE_GQ);
  if (threadIdx.x == 0 && gn) s_gbase = atomicAdd(GQ.count, gn);
  for (u32 e = threadIdx.x; e < ln; e += blockDim.x) {
    const u32 v = lq[e];
    uf_union_h(lab, v & 0xFFFFu, v >> 16);
  }
  __syncthreads();

  // ---- phase 3: staged edges -> global queue; runs -> tile roots ----
  for (u32 e = threadIdx.x; e < gn; e += blockDim.x) {
    const u32 pos = s_gbase + e;
    if (pos < GQ.cap) GQ.q[pos] = gq[e];
    else *GQ.ovf = 1u;
  }
#pragma unroll 1
  for (u32 q = threadIdx.x; q < CC_TILE_WORDS; q += blockDim.x) {
    const u32 wx = q & (TW - 1), r = q >> g.tw;
    const u32 w = w0 + wx, y = y0 + (r & (TY - 1)), z = z0 + (r >> g.ty);
    if (w >= W || y >= sy || z >= sz) continue;
    const u32 i = (z * sy + y) * W + w;
    const uint2 fx = __ldg(reinterpret_cast<const uint2*>(M) + 2 * (size_t)i);
    const int n = __popc(fx.x & ~fx.y);
    if (n == 0) continue;
    const u32 g0 = __ldg(RS + i);
    if (!local_ok) { for (int k = 0; k < n; k++) L[g0 + k] = g0 + k; continue; }
    const u32 l0 = segLB[r] + (g0 - segRS[r]);
    for (int k = 0; k < n; k++) {
      u32 l = l0 + k, p;
      while ((p = lab[l]) != l) l = p;
      u32 root = g0 + k;
      if (l != l0 + (u32)k) {
        // segment of the root: last segment whose first local id is <= l
        u32 lo = 0, hi = nseg - 1;
        while (lo < hi) { const u32 mid = (lo + hi + 1) >> 1; if (segLB[mid] <= l) lo = mid; else hi = mid - 1; }
        root = segRS[lo] + (l - segLB[lo]);
      }
      L[g0 + k] = root;
    }
  }
}

// Kernel B2. One thread per queued edge: union on the global forest L (atomicMin link-to-smaller with
// path halving). Tile roots are at most one hop away, so the finds are short.
static __global__ void __launch_bounds__(256) k_union_queue(u32* __restrict__ L, EdgeQueue GQ) {
  const u32 n = min(*GQ.count, GQ.cap);
  for (u32 e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
    const u64 v = GQ.q[e];
    uf_union_h(L, (u32)v, (u32)(v >> 32));
  }
}

// Kernel B2s (fallback, does nothing unless the edge queue overflowed): one thread per bitmap word,
// every edge is united on the global forest.
template <typename T, int MODE, int CONN>
__global__ void __launch_bounds__(256)
k_union_global(const T* __restrict__ in, const u32* __restrict__ M, u32* __restrict__ L, Geom g, Edge<T, MODE> E,
               const u32* __restrict__ ovf) {
  if (*ovf == 0) return;
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (u32)g.nwords) return;
  const u32 W = (u32)g.W, sy = (u32)g.sy;
  const u32 row = i / W, w = i - row * W;
  const u32 z = row / sy, y = row - z * sy;
  for_each_edge<T, MODE, CONN>(in, M, g, E, i, row, w, y, z,
                               [&](u32 gp, u32 gq_, int, int, u32) { uf_union_h(L, gp, gq_); });
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
