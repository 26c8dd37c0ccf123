// cc3d_union.cuh — kernels B1/B2 (word-parallel edge elimination, tile-local and global unions) and
// kernel P (periodic wrap). See cc3d_common.cuh for the pipeline.
#pragma once
#include "cc3d_common.cuh"

// A face-bitmap word seen from the 32 voxels of word w: c = bit at x, l = bit at x-1, r = bit at x+1.
struct S3 { u32 c, l, r; };
// The four faces of one row around word w.
struct R3 { S3 F, X, Y, Z; };

// ---------------------------------------------------------------------------------------------
// Edge enumeration for one bitmap word (32 voxels p = (x,y,z) of row P); rows U=(y-1,z) D=(y,z-1)
// UD=(y-1,z-1) DN=(y+1,z-1) V=(y+1,z). emit(gid_p, gid_q, dy, dz, xq) is called for every edge that has
// to be united: run ids of both ends, q = voxel xq of row (y+dy, z+dz).
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
struct WordEdges {
  static constexpr int NR = hood_rows(CONN);
  static constexpr bool DIAG0 = CONN == 8 || CONN == 18 || CONN == 26;
  static constexpr bool DIAGZ = CONN == 18 || CONN == 26;
  static constexpr bool CORNER = CONN == 26;
  static constexpr bool TRANS = (MODE == MODE_EQ || MODE == MODE_NONZERO);

  const T* __restrict__ in;
  const u32* __restrict__ M;
  const u32* __restrict__ RS;
  const Geom& g;
  const Edge<T, MODE>& E;
  u32 i, row, w, y, z, W, sy, sx, WS, x0;
  bool hasL, hasR, hasU, hasD, hasV;
  Q4 P, Pl, U, D;
  u32 RSp, Sp;

  __device__ __forceinline__ WordEdges(const T* in_, const u32* M_, const Geom& g_, const Edge<T, MODE>& E_)
      : in(in_), M(M_), RS(M_ + g_.offRS), g(g_), E(E_) {}

  // loads the word; false when it has no foreground
  __device__ __forceinline__ bool load(u32 i_, u32 row_, u32 w_, u32 y_, u32 z_) {
    i = i_; row = row_; w = w_; y = y_; z = z_;
    P = ldq(M, i);
    if (P.F == 0) return false;
    W = (u32)g.W; sy = (u32)g.sy; sx = (u32)g.sx; WS = W * sy; x0 = w << 5;
    hasL = w > 0; hasR = w + 1 < W; hasU = y > 0; hasD = NR >= 2 && z > 0; hasV = y + 1 < sy;
    const Q4 Z4 = {0u, 0u, 0u, 0u};
    Pl = hasL ? ldq(M, i - 1) : Z4;
    U = hasU ? ldq(M, i - W) : Z4;
    D = hasD ? ldq(M, i - WS) : Z4;
    RSp = __ldg(RS + i) - 1u;
    Sp = P.F & ~P.X;
    return true;
  }
  __device__ __forceinline__ u32 pid(int b) const { return RSp + __popc(Sp & (CC_FULL >> (31 - b))); }

  // straight edges that have to be united (x rule and square rule applied)
  __device__ __forceinline__ u32 need_y() const {
    const u32 Yl = (P.Y << 1) | (Pl.Y >> 31);
    return hasU ? (P.Y & ~(P.X & U.X & Yl)) : 0u;
  }
  __device__ __forceinline__ u32 need_z() const {
    const u32 Zl = (P.Z << 1) | (Pl.Z >> 31);
    return hasD ? (P.Z & ~(P.X & D.X & Zl) & ~(P.Y & U.Z & D.Y)) : 0u;
  }
  template <typename EMIT>
  __device__ __forceinline__ void straight(EMIT&& emit) const {
    u32 needY = need_y();
    u32 needZ = need_z();
    if (needY) {
      const u32 RSq = __ldg(RS + i - W) - 1u, Sq = U.F & ~U.X;
      while (needY) {
        const int b = __ffs(needY) - 1; needY &= needY - 1;
        const u32 below = CC_FULL >> (31 - b);
        emit(RSp + __popc(Sp & below), RSq + __popc(Sq & below), -1, 0, x0 + b);
      }
    }
    if (needZ) {
      const u32 RSq = __ldg(RS + i - WS) - 1u, Sq = D.F & ~D.X;
      while (needZ) {
        const int b = __ffs(needZ) - 1; needZ &= needZ - 1;
        const u32 below = CC_FULL >> (31 - b);
        emit(RSp + __popc(Sp & below), RSq + __popc(Sq & below), 0, -1, x0 + b);
      }
    }
  }

  // Cheap superset test "this word may have a diagonal candidate", from the words straight() loaded
  // anyway plus row V. Unknown bits of neighbour words count as possible.
  __device__ __forceinline__ bool may_have_diagonals() const {
    if constexpr (!DIAG0) return false;
    if constexpr (MODE == MODE_MASK) return hasU && (__ldg(M + g.offA0 + i) | __ldg(M + g.offC0 + i)) != 0;
    if constexpr (!TRANS) return true;
    const u32 F = P.F;
    const u32 Yl = (P.Y << 1) | (Pl.Y >> 31);
    const u32 nXr = ~(P.X >> 1) , nYr = ~(P.Y >> 1);            // bit 31 unknown -> possible
    u32 any = 0;
    if (hasU) {
      any |= F & ~P.X & ~P.Y & ~Yl & ~U.X & ((U.F << 1) | 1u);                    // A0
      any |= F & ~P.Y & nXr & nYr & ~(U.X >> 1) & ((U.F >> 1) | 0x80000000u);      // C0
    }
    if constexpr (DIAGZ) {
      if (hasD) {
        const u32 nZ = F & ~P.Z;
        const u32 Zl = (P.Z << 1) | (Pl.Z >> 31);
        any |= nZ & ~P.X & ~Zl & ~D.X & ((D.F << 1) | 1u);                         // A1
        any |= nZ & nXr & ~(P.Z >> 1) & ~(D.X >> 1) & ((D.F >> 1) | 0x80000000u);  // C1
        if (hasU) {
          any |= nZ & ~P.Y & ~U.Z & ~D.Y;                                          // B2
          if constexpr (CORNER) any |= nZ & ~P.Y & (~P.X | nXr);                   // A2, C2
        }
        if (hasV && nZ) {
          const Q4 V = ldq(M, i + W);
          any |= nZ & ~V.Y & ~V.Z;                                                 // B3
          if constexpr (CORNER) any |= nZ & ~V.Y & (~P.X | nXr);                   // A3, C3
        }
      }
    }
    return any != 0;
  }

  template <typename EMIT>
  __device__ __forceinline__ void diagonals(EMIT&& emit) const {
    if constexpr (!DIAG0) return;
    const Q4 Z4 = {0u, 0u, 0u, 0u};
    auto sh3 = [](u32 c, u32 lw, u32 rw) -> S3 { S3 s; s.c = c; s.l = (c << 1) | (lw >> 31); s.r = (c >> 1) | (rw << 31); return s; };
    auto row3 = [&](u32 j) -> R3 {
      const Q4 c = ldq(M, j);
      const Q4 l = hasL ? ldq(M, j - 1) : Z4;
      const Q4 r = hasR ? ldq(M, j + 1) : Z4;
      R3 o;
      o.F = sh3(c.F, l.F, r.F); o.X = sh3(c.X, l.X, r.X); o.Y = sh3(c.Y, l.Y, r.Y); o.Z = sh3(c.Z, l.Z, r.Z);
      return o;
    };
    // candidate masks: 0 A0, 1 C0, 2 A1, 3 C1, 4 B2, 5 A2, 6 C2, 7 B3, 8 A3, 9 C3
    u32 m[10];
#pragma unroll
    for (int t = 0; t < 10; t++) m[t] = 0;
    if constexpr (MODE == MODE_MASK) {
      if (hasU) { m[0] = __ldg(M + g.offA0 + i); m[1] = __ldg(M + g.offC0 + i); }
    } else {
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
        if constexpr (TRANS) {
          m[0] = Fp & RU.F.l & ~(Xp.c | Yp.c | Yp.l | RU.X.c);
          m[1] = Fp & RU.F.r & ~(Xp.r | Yp.r | Yp.c | RU.X.r);
        } else {
          m[0] = Fp & RU.F.l & ~(Xp.c & Yp.l) & ~(Yp.c & RU.X.c);
          m[1] = Fp & RU.F.r & ~(Xp.r & Yp.r) & ~(Yp.c & RU.X.r);
        }
      }
      if constexpr (DIAGZ) {
        if (wantz) {
          const u32 iD = i - WS;
          const R3 RD = row3(iD);
          const S3 XD = RD.X, YD = RD.Y, ZU = RU.Z, XU = RU.X;
          if constexpr (TRANS) {
            m[2] = Fp & RD.F.l & ~(Xp.c | Zp.l | Zp.c | XD.c);
            m[3] = Fp & RD.F.r & ~(Xp.r | Zp.r | Zp.c | XD.r);
          } else {
            m[2] = Fp & RD.F.l & ~(Xp.c & Zp.l) & ~(Zp.c & XD.c);
            m[3] = Fp & RD.F.r & ~(Xp.r & Zp.r) & ~(Zp.c & XD.r);
          }
          // (dy=-1, dz=-1): B2, A2, C2
          if (hasU && (!TRANS || (Fp & ~(Yp.c | Zp.c)))) {
            const R3 RUD = row3(iD - W);
            const S3 FUD = RUD.F, XUD = RUD.X;
            if constexpr (TRANS) m[4] = Fp & FUD.c & ~(Yp.c | ZU.c | Zp.c | YD.c);
            else m[4] = Fp & FUD.c & ~(Yp.c & ZU.c) & ~(Zp.c & YD.c);
            if constexpr (CORNER) {
              if constexpr (TRANS) {
                m[5] = Fp & FUD.l & ~(Xp.c | Yp.c | Zp.c | XUD.c | YD.l | ZU.l);
                m[6] = Fp & FUD.r & ~(Xp.r | Yp.c | Zp.c | XUD.r | YD.r | ZU.r);
              } else {
                m[5] = Fp & FUD.l & ~(Xp.c & Yp.l & ZU.l) & ~(Xp.c & Zp.l & YD.l) & ~(Yp.c & XU.c & ZU.l)
                       & ~(Yp.c & ZU.c & XUD.c) & ~(Zp.c & XD.c & YD.l) & ~(Zp.c & YD.c & XUD.c);
                m[6] = Fp & FUD.r & ~(Xp.r & Yp.r & ZU.r) & ~(Xp.r & Zp.r & YD.r) & ~(Yp.c & XU.r & ZU.r)
                       & ~(Yp.c & ZU.c & XUD.r) & ~(Zp.c & XD.r & YD.r) & ~(Zp.c & YD.c & XUD.r);
              }
            }
          }
          // (dy=+1, dz=-1): B3, A3, C3
          if (hasV) {
            const R3 RV = row3(i + W);
            const S3 YV = RV.Y, ZV = RV.Z, XV = RV.X;
            if (!TRANS || (Fp & ~(Zp.c | YV.c))) {
              const R3 RDN = row3(iD + W);
              const S3 FDN = RDN.F, XDN = RDN.X, YDN = RDN.Y;
              if constexpr (TRANS) m[7] = Fp & FDN.c & ~(YV.c | ZV.c | Zp.c | YDN.c);
              else m[7] = Fp & FDN.c & ~(YV.c & ZV.c) & ~(Zp.c & YDN.c);
              if constexpr (CORNER) {
                if constexpr (TRANS) {
                  m[8] = Fp & FDN.l & ~(Xp.c | YV.c | Zp.c | XDN.c | YDN.l | ZV.l);
                  m[9] = Fp & FDN.r & ~(Xp.r | YV.c | Zp.c | XDN.r | YDN.r | ZV.r);
                } else {
                  m[8] = Fp & FDN.l & ~(Xp.c & YV.l & ZV.l) & ~(Xp.c & Zp.l & YDN.l) & ~(YV.c & XV.c & ZV.l)
                         & ~(YV.c & ZV.c & XDN.c) & ~(Zp.c & XD.c & YDN.l) & ~(Zp.c & YDN.c & XDN.c);
                  m[9] = Fp & FDN.r & ~(Xp.r & YV.r & ZV.r) & ~(Xp.r & Zp.r & YDN.r) & ~(YV.c & XV.r & ZV.r)
                         & ~(YV.c & ZV.c & XDN.r) & ~(Zp.c & XD.r & YDN.r) & ~(Zp.c & YDN.c & XDN.r);
                }
              }
            }
          }
        }
      }
    }
    // one loop over the voxels that have any candidate (instead of one divergent loop per direction)
    u32 any = 0;
#pragma unroll
    for (int t = 0; t < 10; t++) any |= m[t];
    while (any) {
      const int b = __ffs(any) - 1; any &= any - 1;
      const u32 gp = pid(b);
#pragma unroll
      for (int t = 0; t < 10; t++) {
        constexpr int DY[10] = {-1, -1, 0, 0, -1, -1, -1, 1, 1, 1};
        constexpr int DZ[10] = {0, 0, -1, -1, -1, -1, -1, -1, -1, -1};
        constexpr int DX[10] = {-1, 1, -1, 1, 0, -1, 1, 0, -1, 1};
        if ((m[t] >> b) & 1u) {
          const u32 rowQ = row + DY[t] + DZ[t] * (int)sy;
          const u32 xq = x0 + b + DX[t];
          bool joined = true;
          if constexpr (MODE == MODE_EQ || MODE == MODE_DELTA) joined = E(in[row * sx + x0 + b], in[rowQ * sx + xq]);
          if (joined) emit(gp, run_id(M, g, rowQ * W, xq), DY[t], DZ[t], xq);
        }
      }
    }
  }
};

// ---------------------------------------------------------------------------------------------
// Kernel B1. One CTA per union tile (2^tw words x 2^ty rows x 2^tz planes = CC_TILE_WORDS words).
// A run belongs to the tile its first voxel lies in; an edge is tile-local when both of its runs belong
// to the tile. The runs that start in a tile row segment (2^tw words of one row) have contiguous ids
// from RS[first word] on, so local node = segment * cap + (run id - first id of the segment) keeps the
// raster order of the runs (link-to-smaller stays valid) and converts back with one table lookup.
//   round 0: one thread per word: straight edges are enumerated and only CLASSIFIED: tile-local edges
//            go to a shared-memory queue, the others to a staging buffer; words that may have diagonal
//            candidates are put on a to-do list (so that the costly candidate code runs with full warps)
//   round 1: one thread per to-do word: diagonal candidates, classified the same way
//   after each round the queue is worked off by all threads, one edge each (balanced): union-find in
//   shared memory. Finally every run of the tile gets L[run] = run id of its tile root, and the staged
//   edges are appended to the global edge queue GQ for kernel B2.
// A tile whose rows hold more than 16 runs per word (possible for multilabel input only) sends all its
// edges to B2. If GQ overflows, *ovf is raised and kernel B2s redoes every edge on the global forest.
// ---------------------------------------------------------------------------------------------
struct EdgeQueue { u64* q; u32* count; u32* ovf; u32 cap; };

template <typename T, int MODE, int CONN>
__global__ void __launch_bounds__(256)
k_union_tile(const T* __restrict__ in, const u32* __restrict__ M, u32* __restrict__ L, Geom g, Edge<T, MODE> E,
             u32 ntx, u32 nty, EdgeQueue GQ) {
  extern __shared__ __align__(16) u32 smem_u32[];
  uint16_t* lab = reinterpret_cast<uint16_t*>(smem_u32);   // [CC_TILE_NODES] 16-bit parents
  u32* lq = smem_u32 + CC_TILE_NODES / 2;                  // [CC_TILE_LQ]
  u64* gq = reinterpret_cast<u64*>(lq + CC_TILE_LQ);       // [CC_TILE_GQ]
  u32* segRS = lq + CC_TILE_LQ + 2 * CC_TILE_GQ;           // [CC_TILE_WORDS] first run id of every row segment
  uint16_t* todo = reinterpret_cast<uint16_t*>(segRS + CC_TILE_WORDS);   // [CC_TILE_WORDS]
  __shared__ u32 s_ln[4], s_gn, s_gbase, s_tn, s_big, s_runs;
  const u32 W = (u32)g.W, sy = (u32)g.sy, sz = (u32)g.sz;
  const u32 TW = 1u << g.tw, TY = 1u << g.ty;
  const u32 nseg = CC_TILE_WORDS >> g.tw;        // TY * TZ
  const u32 capl = g.tw + 4;                     // log2(runs a segment can hold locally)
  u32 t = blockIdx.x;
  const u32 bx = t % ntx; t /= ntx;
  const u32 by = t % nty;
  const u32 bz = t / nty;
  const u32 w0 = bx << g.tw, y0 = by << g.ty, z0 = bz << g.tz;
  const u32 wend = min(w0 + TW, W);
  const u32* __restrict__ RS = M + g.offRS;

  if (threadIdx.x < 4) s_ln[threadIdx.x] = 0;
  if (threadIdx.x == 0) { s_gn = 0; s_tn = 0; s_big = 0; s_runs = 0; }
  __syncthreads();
  for (u32 r = threadIdx.x; r < nseg; r += blockDim.x) {
    const u32 y = y0 + (r & (TY - 1)), z = z0 + (r >> g.ty);
    u32 first = 0xFFFFFFFFu;
    if (y < sy && z < sz) {
      const u32 j = (z * sy + y) * W;
      first = __ldg(RS + j + w0);
      const u32 cnt = __ldg(RS + j + wend) - first;
      if (cnt > (1u << capl)) s_big = 1;   // multilabel rows with > 16 runs per word
      atomicAdd(&s_runs, cnt);
    }
    segRS[r] = first;
  }
  for (u32 k = threadIdx.x; k < CC_TILE_NODES / 2; k += blockDim.x) smem_u32[k] = (2 * k) | ((2 * k + 1) << 16);
  __syncthreads();
  const bool tile_ok = s_big == 0;   // otherwise every edge of this tile goes to kernel B2
  // dense tiles enumerate 256 words at a time so that the edge queue is drained before it overflows
  const u32 step = s_runs > CC_TILE_LQ / 2 ? 256u : (u32)CC_TILE_WORDS;

  auto push_global = [&](u32 gp, u32 gq_) {
    const u32 pos = atomicAdd(GQ.count, 1u);
    if (pos < GQ.cap) GQ.q[pos] = (u64)gp | ((u64)gq_ << 32);
    else *GQ.ovf = 1u;
  };

  WordEdges<T, MODE, CONN> we(in, M, g, E);
  int sub = 0;
#pragma unroll 1
  for (int round = 0; round < 2; round++) {
    // ---- enumerate + classify: round 0 = straight edges of every word, round 1 = diagonals of the to-do words ----
    const u32 nitems = round == 0 ? (u32)CC_TILE_WORDS : s_tn;
#pragma unroll 1
    for (u32 base = 0; base < nitems; base += step, sub++) {
    const u32 iend = min(nitems, base + step);
#pragma unroll 1
    for (u32 e = base + threadIdx.x; e < iend; e += blockDim.x) {
      const u32 q = round == 0 ? e : (u32)todo[e];
      const u32 wx = q & (TW - 1), r = q >> g.tw;
      const int ly = (int)(r & (TY - 1)), lz = (int)(r >> g.ty);
      const u32 w = w0 + wx, y = y0 + ly, z = z0 + lz;
      if (w >= W || y >= sy || z >= sz) continue;
      const u32 row = z * sy + y;
      if (!we.load(row * W + w, row, w, y, z)) continue;
      const u32 baseP = segRS[r];
      auto classify = [&](u32 gp, u32 gq_, int dy, int dz, u32 xq) {
        bool local = tile_ok && gp >= baseP;
        u32 rq = 0;
        if (local) {
          const int lyq = ly + dy, lzq = lz + dz;
          local = ((xq >> 5) >> g.tw) == bx && lyq >= 0 && lyq < (int)TY && lzq >= 0;
          if (local) { rq = ((u32)lzq << g.ty) + (u32)lyq; local = gq_ >= segRS[rq]; }
        }
        if (local) {
          const u32 lp = (r << capl) + (gp - baseP), lq_ = (rq << capl) + (gq_ - segRS[rq]);
          const u32 pos = atomicAdd(&s_ln[sub], 1u);
          if (pos < CC_TILE_LQ) lq[pos] = lp | (lq_ << 16);
          else sm_union16(lab, lp, lq_);
        } else {
          const u32 pos = atomicAdd(&s_gn, 1u);
          if (pos < CC_TILE_GQ) gq[pos] = (u64)gp | ((u64)gq_ << 32);
          else push_global(gp, gq_);
        }
      };
      if (round == 0) {
        // straight edges: q is the same x in the row above / the plane below, so whether the edge stays in
        // the tile is decided per word (row inside the tile) except for runs that entered the tile from the
        // left. One queue reservation per word; slots of edges that turn out to leave the tile get a no-op.
        auto straight_fast = [&](u32 need, const Q4& Qf, const u32 jq, const bool rowlocal, const u32 rq) {
          if (!need) return;
          const u32 Sq = Qf.F & ~Qf.X;
          const u32 RSq = __ldg(RS + jq) - 1u;
          const u32 n = __popc(need);
          if (rowlocal) {
            const u32 baseQ = segRS[rq];
            const u32 pos0 = atomicAdd(&s_ln[sub], n);
            u32 k = 0;
            while (need) {
              const int b = __ffs(need) - 1; need &= need - 1;
              const u32 below = CC_FULL >> (31 - b);
              const u32 gp = we.RSp + __popc(we.Sp & below), gq_ = RSq + __popc(Sq & below);
              const bool loc = gp >= baseP && gq_ >= baseQ;
              const u32 lp = (r << capl) + (gp - baseP), lq_ = (rq << capl) + (gq_ - baseQ);
              const u32 pos = pos0 + k++;
              if (pos < CC_TILE_LQ) lq[pos] = loc ? (lp | (lq_ << 16)) : 0u;
              else if (loc) sm_union16(lab, lp, lq_);
              if (!loc) {
                const u32 gpos = atomicAdd(&s_gn, 1u);
                if (gpos < CC_TILE_GQ) gq[gpos] = (u64)gp | ((u64)gq_ << 32);
                else push_global(gp, gq_);
              }
            }
          } else {
            const u32 pos0 = atomicAdd(&s_gn, n);
            u32 k = 0;
            while (need) {
              const int b = __ffs(need) - 1; need &= need - 1;
              const u32 below = CC_FULL >> (31 - b);
              const u32 gp = we.RSp + __popc(we.Sp & below), gq_ = RSq + __popc(Sq & below);
              const u32 pos = pos0 + k++;
              if (pos < CC_TILE_GQ) gq[pos] = (u64)gp | ((u64)gq_ << 32);
              else push_global(gp, gq_);
            }
          }
        };
        straight_fast(we.need_y(), we.U, we.i - W, tile_ok && ly > 0, r - 1);
        straight_fast(we.need_z(), we.D, we.i - W * sy, tile_ok && lz > 0, r - TY);
        if (we.may_have_diagonals()) todo[atomicAdd(&s_tn, 1u)] = (uint16_t)q;
      } else {
        we.diagonals(classify);
      }
    }
    __syncthreads();
    // ---- tile-local unions, one queued edge per thread and step ----
    const u32 ln = min(s_ln[sub], (u32)CC_TILE_LQ);
    for (u32 e = threadIdx.x; e < ln; e += blockDim.x) {
      const u32 v = lq[e];
      sm_union16(lab, v & 0xFFFFu, v >> 16);
    }
    __syncthreads();
    }
  }

  // ---- staged edges -> global queue; runs -> tile roots ----
  const u32 gn = min(s_gn, (u32)CC_TILE_GQ);
  if (threadIdx.x == 0 && gn) s_gbase = atomicAdd(GQ.count, gn);
  __syncthreads();
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
    if (!tile_ok) { for (int k = 0; k < n; k++) L[g0 + k] = g0 + k; continue; }
    const u32 l0 = (r << capl) + (g0 - segRS[r]);
    for (int k = 0; k < n; k++) {
      u32 l = l0 + k, p;
      while ((p = lab[l]) != l) l = p;
      L[g0 + k] = segRS[l >> capl] + (l & ((1u << capl) - 1u));
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
  const u32 W = (u32)g.W, sy = (u32)g.sy;
  WordEdges<T, MODE, CONN> we(in, M, g, E);
  auto unite = [&](u32 gp, u32 gq_, int, int, u32) { uf_union_h(L, gp, gq_); };
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < (u32)g.nwords; i += gridDim.x * blockDim.x) {
    const u32 row = i / W, w = i - row * W;
    const u32 z = row / sy, y = row - z * sy;
    if (!we.load(i, row, w, y, z)) continue;
    we.straight(unite);
    we.diagonals(unite);
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
