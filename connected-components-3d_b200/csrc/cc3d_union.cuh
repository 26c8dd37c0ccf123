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
  // Same as load(), but every word the straight edges can need - P, its left neighbour, U, D and the run ids of U and D -
  // is requested before the first of them is looked at: one level of dependent loads instead of three (load() waits
  // for P.F before it asks for the neighbours, and the callers ask for the neighbours' run ids after the need masks).
  u32 RSu, RSd;
  __device__ __forceinline__ bool load_eager(u32 i_, u32 row_, u32 w_, u32 y_, u32 z_) {
    i = i_; row = row_; w = w_; y = y_; z = z_;
    W = (u32)g.W; sy = (u32)g.sy; sx = (u32)g.sx; WS = W * sy; x0 = w << 5;
    hasL = w > 0; hasR = w + 1 < W; hasU = y > 0; hasD = NR >= 2 && z > 0; hasV = y + 1 < sy;
    const Q4 Z4 = {0u, 0u, 0u, 0u};
    P = ldq(M, i);
    Pl = hasL ? ldq(M, i - 1) : Z4;
    U = hasU ? ldq(M, i - W) : Z4;
    D = hasD ? ldq(M, i - WS) : Z4;
    RSp = __ldg(RS + i) - 1u;
    RSu = hasU ? __ldg(RS + i - W) - 1u : 0u;
    RSd = hasD ? __ldg(RS + i - WS) - 1u : 0u;
    Sp = P.F & ~P.X;
    return P.F != 0;
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
          if constexpr (CORNER) {
            const u32 c2 = nZ & ~P.Y;       // A2, C2: neither p nor q = (x -+ 1, y - 1, z - 1) has a face link into the 2x2x2 block
            any |= c2 & ~P.X & ~(U.Z << 1) & ~(D.Y << 1);          // (links of q seen from rows U and D; shifted-in bits unknown -> possible)
            any |= c2 & nXr & ~(U.Z >> 1) & ~(D.Y >> 1);
          }
        }
        if (hasV && nZ) {
          const Q4 V = ldq(M, i + W);
          any |= nZ & ~V.Y & ~V.Z;                                                 // B3
          if constexpr (CORNER) {
            const u32 c3 = nZ & ~V.Y;       // A3, C3
            any |= c3 & ~P.X & ~(V.Z << 1);
            any |= c3 & nXr & ~(V.Z >> 1);
          }
        }
      }
    }
    return any != 0;
  }

  // candidate masks of the diagonal directions: 0 A0, 1 C0, 2 A1, 3 C1, 4 B2, 5 A2, 6 C2, 7 B3, 8 A3, 9 C3
  // (bit b: voxel b of this word may be joined to its neighbour in that direction); false = all empty
  __device__ __forceinline__ bool diag_masks(u32* m) const {
#pragma unroll
    for (int t = 0; t < 10; t++) m[t] = 0;
    if constexpr (!DIAG0) return false;
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
    if constexpr (MODE == MODE_MASK) {
      if (hasU) { m[0] = __ldg(M + g.offA0 + i); m[1] = __ldg(M + g.offC0 + i); }
    } else {
      const bool want0 = hasU && (!TRANS || (P.F & ~P.Y));
      const bool wantz = DIAGZ && hasD && (!TRANS || (P.F & ~P.Z));
      if (!want0 && !wantz) return false;
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
    u32 any = 0;
#pragma unroll
    for (int t = 0; t < 10; t++) any |= m[t];
    return any != 0;
  }

  template <typename EMIT>
  __device__ __forceinline__ void diagonals(EMIT&& emit) const {
    u32 m[10];
    if (!diag_masks(m)) return;
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
          if constexpr (MODE == MODE_EQ || MODE == MODE_DELTA || MODE == MODE_BLOCK) joined = E.diag(t, in[row * sx + x0 + b], in[rowQ * sx + xq]);
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

// (binary volumes; multilabel volumes run k_union_tile_hybrid below)
template <int MODE> struct TileQueues {
  static constexpr u32 GQ = MODE == MODE_EQ ? CC_TILE_GQ_EQ : CC_TILE_GQ;
  static constexpr u32 SMEM_WORDS = CC_TILE_NODES / 2 + CC_TILE_LQ + 2 * GQ + CC_TILE_WORDS + CC_TILE_WORDS / 2;
};
template <typename T, int MODE, int CONN>
__global__ void __launch_bounds__(CC_TILE_THREADS, MODE == MODE_EQ ? CC_TILE_MINB(6) : (MODE == MODE_BLOCK ? CC_TILE_MINB(CC_BLOCK_MINB) : 0))
k_union_tile(const T* __restrict__ in, const u32* __restrict__ M, u32* __restrict__ L, Geom g, Edge<T, MODE> E,
             u32 ntx, u32 nty, EdgeQueue GQ) {
  CC_PDL_WAIT();
  constexpr u32 GQN = TileQueues<MODE>::GQ;
  extern __shared__ __align__(16) u32 smem_u32[];
  uint16_t* lab = reinterpret_cast<uint16_t*>(smem_u32);   // [CC_TILE_NODES] 16-bit parents
  u32* lq = smem_u32 + CC_TILE_NODES / 2;                  // [CC_TILE_LQ]
  u64* gq = reinterpret_cast<u64*>(lq + CC_TILE_LQ);       // [GQN]
  u32* segRS = lq + CC_TILE_LQ + 2 * GQN;           // [CC_TILE_WORDS] first run id of every row segment
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
  const u32 step = s_runs > CC_TILE_LQ / 2 ? (u32)(CC_TILE_WORDS / 2) : (u32)CC_TILE_WORDS;

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
          if (pos < GQN) gq[pos] = (u64)gp | ((u64)gq_ << 32);
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
                if (gpos < GQN) gq[gpos] = (u64)gp | ((u64)gq_ << 32);
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
              if (pos < GQN) gq[pos] = (u64)gp | ((u64)gq_ << 32);
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
  const u32 gn = min(s_gn, GQN);
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

// Kernel B1 for multilabel volumes (EQ, and the explicit-diagonal planes of continuous 2D-8): round 0 as in
// k_union_tile; round 1 expands the diagonal candidate masks of the to-do words into (word, direction, bit) work
// items and resolves one item per thread from a shared-memory stash of run starts / first run ids (wS, wR) instead
// of running the candidate loop per word. Tiles with many runs (dense noise) keep the per-word loop: their item
// lists would overflow. Label volumes are latency / barrier bound here: six resident CTAs per SM (40 registers, a
// 512-entry staging buffer) hide more of it.
// RL = log2 of the runs per word the shared-memory forest has room for: 4 (16 runs per word - every label volume and every
// binary image) for the first launch; tiles whose rows hold more (multilabel noise: up to 32 runs per word, the
// maximum) are flagged and - once the process has met such a volume (BigTiles::defer) - relabelled by a second launch
// with RL = 5 instead of sending every one of their edges through the global queue.
struct BigTiles { u32* flags; u32* count; int defer; };
template <int MODE, int RL = 4> struct HybridQueues {
  static constexpr u32 GQ = MODE == MODE_EQ ? CC_TILE_GQ_EQ : CC_TILE_GQ;
  static constexpr bool ITEMS = true;
  static constexpr u32 NODES = CC_TILE_WORDS << RL;
  static constexpr u32 SMEM_WORDS = NODES / 2 + CC_TILE_LQ + 2 * GQ + CC_TILE_WORDS + CC_TILE_WORDS / 2 + (ITEMS ? 2 * CC_TILE_WORDS : 0);
};
template <typename T, int MODE, int CONN, int RL>
__global__ void __launch_bounds__(CC_TILE_THREADS, (MODE == MODE_EQ && RL == 4) ? CC_TILE_MINB(CC_B1_MINB) : 0)
k_union_tile_hybrid(const T* __restrict__ in, const u32* __restrict__ M, u32* __restrict__ L, Geom g, Edge<T, MODE> E,
             u32 ntx, u32 nty, EdgeQueue GQ, BigTiles big) {
  CC_PDL_WAIT();
  if constexpr (RL != 4) { if (big.flags[blockIdx.x] == 0) return; }      // second launch: flagged tiles only
  constexpr u32 GQN = HybridQueues<MODE, RL>::GQ;
  constexpr bool ITEMS = HybridQueues<MODE, RL>::ITEMS;   // round 1 through item lists (needs the wS / wR stash)
  constexpr u32 NODES = HybridQueues<MODE, RL>::NODES;
  extern __shared__ __align__(16) u32 smem_u32[];
  uint16_t* lab = reinterpret_cast<uint16_t*>(smem_u32);   // [NODES] 16-bit parents
  u32* lq = smem_u32 + NODES / 2;                          // [CC_TILE_LQ]
  u64* gq = reinterpret_cast<u64*>(lq + CC_TILE_LQ);       // [GQN]
  u32* segRS = lq + CC_TILE_LQ + 2 * GQN;           // [CC_TILE_WORDS] first run id of every row segment
  uint16_t* todo = reinterpret_cast<uint16_t*>(segRS + CC_TILE_WORDS);   // [CC_TILE_WORDS]
  u32* wS = segRS + CC_TILE_WORDS + CC_TILE_WORDS / 2;     // [CC_TILE_WORDS] run starts of every word (round 0 -> round 1)
  u32* wR = wS + CC_TILE_WORDS;                            // [CC_TILE_WORDS] id of the first run that starts in the word, minus 1
  __shared__ u32 s_ln[4], s_in[2], s_gn, s_gbase, s_tn, s_big, s_runs;
  const u32 W = (u32)g.W, sy = (u32)g.sy, sz = (u32)g.sz;
  const u32 TW = 1u << g.tw, TY = 1u << g.ty;
  const u32 nseg = CC_TILE_WORDS >> g.tw;        // TY * TZ
  const u32 capl = g.tw + RL;                    // log2(runs a segment can hold locally)
  u32 t = blockIdx.x;
  const u32 bx = t % ntx; t /= ntx;
  const u32 by = t % nty;
  const u32 bz = t / nty;
  const u32 w0 = bx << g.tw, y0 = by << g.ty, z0 = bz << g.tz;
  const u32 wend = min(w0 + TW, W);
  const u32* __restrict__ RS = M + g.offRS;

  if (threadIdx.x < 4) s_ln[threadIdx.x] = 0;
  if (threadIdx.x == 0) { s_gn = 0; s_tn = 0; s_big = 0; s_runs = 0; s_in[0] = 0; s_in[1] = 0; }
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
  for (u32 k = threadIdx.x; k < NODES / 2; k += blockDim.x) smem_u32[k] = (2 * k) | ((2 * k + 1) << 16);
  __syncthreads();
  if constexpr (RL == 4) {
    if (s_big) {      // block-uniform
      if (threadIdx.x == 0) { atomicAdd(big.count, 1u); if (big.defer) big.flags[blockIdx.x] = 1u; }
      if (big.defer) return;     // the RL = 5 launch relabels this tile
    }
  }
  const bool tile_ok = s_big == 0;   // otherwise every edge of this tile goes to kernel B2
  // dense tiles enumerate 256 words at a time so that the edge queue is drained before it overflows
  const u32 step = s_runs > CC_TILE_LQ / 2 ? (u32)(CC_TILE_WORDS / 2) : (u32)CC_TILE_WORDS;

  auto push_global = [&](u32 gp, u32 gq_) {
    const u32 pos = atomicAdd(GQ.count, 1u);
    if (pos < GQ.cap) GQ.q[pos] = (u64)gp | ((u64)gq_ << 32);
    else *GQ.ovf = 1u;
  };

  WordEdges<T, MODE, CONN> we(in, M, g, E);
  int sub = 0;
#pragma unroll 1
  for (int round = 0; round < 1; round++) {
    // ---- round 0: straight edges of every word, enumerated and classified per word; words that may have
    //      diagonal candidates are put on the to-do list of round 1 ----
    const u32 nitems = (u32)CC_TILE_WORDS;
#pragma unroll 1
    for (u32 base = 0; base < nitems; base += step, sub++) {
    const u32 iend = min(nitems, base + step);
#pragma unroll 1
    for (u32 e = base + threadIdx.x; e < iend; e += blockDim.x) {
      const u32 q = e;
      const u32 wx = q & (TW - 1), r = q >> g.tw;
      const int ly = (int)(r & (TY - 1)), lz = (int)(r >> g.ty);
      const u32 w = w0 + wx, y = y0 + ly, z = z0 + lz;
      if constexpr (ITEMS) wS[q] = 0;
      if (w >= W || y >= sy || z >= sz) continue;
      const u32 row = z * sy + y;
      if (!we.load(row * W + w, row, w, y, z)) continue;
      if constexpr (ITEMS) { wS[q] = we.Sp; wR[q] = we.RSp; }
      const u32 baseP = segRS[r];
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
                if (gpos < GQN) gq[gpos] = (u64)gp | ((u64)gq_ << 32);
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
              if (pos < GQN) gq[pos] = (u64)gp | ((u64)gq_ << 32);
              else push_global(gp, gq_);
            }
          }
        };
        straight_fast(we.need_y(), we.U, we.i - W, tile_ok && ly > 0, r - 1);
        straight_fast(we.need_z(), we.D, we.i - W * sy, tile_ok && lz > 0, r - TY);
        if (we.may_have_diagonals()) todo[atomicAdd(&s_tn, 1u)] = (uint16_t)q;
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

  // ---- round 1: diagonal candidates of the to-do words. Word-parallel masks, expanded into work items
  //      (word, direction, bit) in the edge-queue buffer, then one item per thread: run ids of both ends from the
  //      shared-memory stash (wS / wR), value test for EQ / DELTA, union or staging. ----
  // Dense tiles (random binary volumes: a dozen candidates per word) keep the per-word candidate loop with the
  // balanced union queue - their item lists would overflow; label volumes take the item lists.
  const bool dense_tile = !ITEMS || s_runs > CC_TILE_LQ / 2;
  if (WordEdges<T, MODE, CONN>::DIAG0 && dense_tile) {
    const u32 ntodo = s_tn;
#pragma unroll 1
    for (u32 base = 0; base < ntodo; base += step, sub++) {
      const u32 iend = min(ntodo, base + step);
#pragma unroll 1
      for (u32 e = base + threadIdx.x; e < iend; e += blockDim.x) {
        const u32 q = (u32)todo[e];
        const u32 wx = q & (TW - 1), r = q >> g.tw;
        const int ly = (int)(r & (TY - 1)), lz = (int)(r >> g.ty);
        const u32 w = w0 + wx, y = y0 + ly, z = z0 + lz;
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
            const u32 pos = atomicAdd(&s_ln[sub & 3], 1u);
            if (pos < CC_TILE_LQ) lq[pos] = lp | (lq_ << 16);
            else sm_union16(lab, lp, lq_);
          } else {
            const u32 pos = atomicAdd(&s_gn, 1u);
            if (pos < GQN) gq[pos] = (u64)gp | ((u64)gq_ << 32);
            else push_global(gp, gq_);
          }
        };
        we.diagonals(classify);
      }
      __syncthreads();
      const u32 ln = min(s_ln[sub & 3], (u32)CC_TILE_LQ);
      for (u32 e = threadIdx.x; e < ln; e += blockDim.x) {
        const u32 v = lq[e];
        sm_union16(lab, v & 0xFFFFu, v >> 16);
      }
      __syncthreads();
    }
  } else if constexpr (WordEdges<T, MODE, CONN>::DIAG0 && ITEMS) {
    // two bits per diagonal direction of WordEdges::diag_masks: d + 1
    constexpr u32 DXP = (0u << 0) | (2u << 2) | (0u << 4) | (2u << 6) | (1u << 8) | (0u << 10) | (2u << 12) | (1u << 14) | (0u << 16) | (2u << 18);
    constexpr u32 DYP = (0u << 0) | (0u << 2) | (1u << 4) | (1u << 6) | (0u << 8) | (0u << 10) | (0u << 12) | (2u << 14) | (2u << 16) | (2u << 18);
    constexpr u32 DZP = (1u << 0) | (1u << 2) | (0u << 4) | (0u << 6) | (0u << 8) | (0u << 10) | (0u << 12) | (0u << 14) | (0u << 16) | (0u << 18);
    const u32 sx = (u32)g.sx;
    const int lane = threadIdx.x & 31;
    auto resolve = [&](const u32 item) {
      const u32 b = item & 31u, tdir = (item >> 5) & 15u, q = item >> 9;
      const u32 wx = q & (TW - 1), r = q >> g.tw;
      const int ly = (int)(r & (TY - 1)), lz = (int)(r >> g.ty);
      const int dx = (int)((DXP >> (2 * tdir)) & 3u) - 1, dy = (int)((DYP >> (2 * tdir)) & 3u) - 1, dz = (int)((DZP >> (2 * tdir)) & 3u) - 1;
      const u32 gp = wR[q] + __popc(wS[q] & (CC_FULL >> (31 - b)));
      const int xl = (int)((wx << 5) + b) + dx;            // x of q relative to the tile
      const int lyq = ly + dy, lzq = lz + dz;
      const bool inside = xl >= 0 && xl < (int)(TW << 5) && lyq >= 0 && lyq < (int)TY && lzq >= 0;
      const u32 rowP = (z0 + lz) * sy + y0 + ly;
      const u32 rowQ = (u32)((int)rowP + dy + dz * (int)sy);
      const u32 xq = (u32)((int)(w0 << 5) + xl);
      if constexpr (MODE == MODE_EQ || MODE == MODE_DELTA) {
        if (!E(in[(size_t)rowP * sx + ((w0 + wx) << 5) + b], in[(size_t)rowQ * sx + xq])) return;
      }
      u32 gq_, rq = 0;
      if (inside) {
        rq = ((u32)lzq << g.ty) + (u32)lyq;
        const u32 qq = (rq << g.tw) + ((u32)xl >> 5);
        gq_ = wR[qq] + __popc(wS[qq] & (CC_FULL >> (31 - (xl & 31))));
      } else {
        gq_ = run_id(M, g, rowQ * W, xq);
      }
      bool local = tile_ok && inside && gp >= segRS[r];
      if (local) local = gq_ >= segRS[rq];
      if (local) {
        sm_union16(lab, (r << capl) + (gp - segRS[r]), (rq << capl) + (gq_ - segRS[rq]));
      } else {
        const u32 pos = atomicAdd(&s_gn, 1u);
        if (pos < GQN) gq[pos] = (u64)gp | ((u64)gq_ << 32);
        else push_global(gp, gq_);
      }
    };
    const u32 ntodo = s_tn;
    u32 par = 0;
#pragma unroll 1
    for (u32 base = 0; base < ntodo; base += blockDim.x, par ^= 1u) {
      const u32 e = base + threadIdx.x;
      u32 m[10];
#pragma unroll
      for (int t = 0; t < 10; t++) m[t] = 0;
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
      for (int t = 0; t < 10; t++) n += __popc(m[t]);
      u32 inc = n;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const u32 v = __shfl_up_sync(CC_FULL, inc, o);
        if (lane >= o) inc += v;
      }
      const u32 wtot = __shfl_sync(CC_FULL, inc, 31);
      u32 wbase = 0;
      if (wtot) {
        if (lane == 31) wbase = atomicAdd(&s_in[par], wtot);
        wbase = __shfl_sync(CC_FULL, wbase, 31);
      }
      u32 pos = wbase + inc - n;
      if (n) {
        const u32 qb = q << 9;
#pragma unroll
        for (int t = 0; t < 10; t++) {
          u32 mk = m[t];
          while (mk) {
            const u32 bb = __ffs(mk) - 1; mk &= mk - 1;
            const u32 item = qb | ((u32)t << 5) | bb;
            if (pos < CC_TILE_LQ) lq[pos] = item;
            else resolve(item);          // list full: resolve in place
            pos++;
          }
        }
      }
      __syncthreads();
      const u32 ni = min(s_in[par], (u32)CC_TILE_LQ);
      if (threadIdx.x == 0) s_in[par ^ 1u] = 0;
      for (u32 kk = threadIdx.x; kk < ni; kk += blockDim.x) resolve(lq[kk]);
      __syncthreads();
    }
  }

  // ---- staged edges -> global queue; runs -> tile roots ----
  const u32 gn = min(s_gn, GQN);
  if (threadIdx.x == 0 && gn) s_gbase = atomicAdd(GQ.count, gn);
  __syncthreads();
  for (u32 e = threadIdx.x; e < gn; e += blockDim.x) {
    const u32 pos = s_gbase + e;
    if (pos < GQ.cap) GQ.q[pos] = gq[e];
    else *GQ.ovf = 1u;
  }
#pragma unroll 1
  for (u32 q = threadIdx.x; q < CC_TILE_WORDS; q += blockDim.x) {
    const u32 r = q >> g.tw;
    const u32 wx = q & (TW - 1);
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

// ---------------------------------------------------------------------------------------------
// Kernel B1, item-list variant (used for the continuous predicate, where every word has diagonal candidates
// whose value test is expensive and diverges inside a per-word loop). One CTA per union tile (2^tw words x 2^ty rows x 2^tz planes = CC_TILE_WORDS words).
// A run belongs to the tile its first voxel lies in; an edge is tile-local when both of its runs belong
// to the tile. The runs that start in a tile row segment (2^tw words of one row) have contiguous ids
// from RS[first word] on, so local node = segment * cap + (run id - first id of the segment) keeps the
// raster order of the runs (link-to-smaller stays valid) and converts back with one table lookup.
//
// Work is split into a word-parallel part and an edge-parallel part so that neither diverges:
//   pre-pass : run starts and first run id of every word of the tile -> shared memory (wS, wR)
//   enumerate: one thread per word computes edge MASKS with whole-word logic (round 0: the straight edges
//              left by the x / square rules; round 1: diagonal candidates of the words flagged in round 0)
//              and expands them into 18-bit work items (word, direction, bit) in a shared-memory list
//              (slots reserved with one warp scan + one atomic per warp)
//   resolve  : one thread per item: run ids of both ends from wS/wR (global bitmaps when the neighbour lies
//              outside the tile), value test for EQ / DELTA candidates, then union in the 16-bit
//              shared-memory forest, or - for edges that leave the tile - a staging buffer that is
//              appended to the global edge queue GQ for kernel B2.
// Finally every run of the tile gets L[run] = run id of its tile root.
// A tile whose rows hold more than 16 runs per word (possible for multilabel input only) sends all its
// edges to B2. If GQ overflows, *ovf is raised and kernel B2s redoes every edge on the global forest.
// ---------------------------------------------------------------------------------------------
#define CC_TILE_ITEMS (CC_TILE_WORDS * 8)    // work items per enumerate step (one word per thread)
#define CC_TILE_SMEM_WORDS (CC_TILE_NODES / 2 + CC_TILE_ITEMS + 2 * CC_TILE_GQ + 3 * CC_TILE_WORDS + CC_TILE_WORDS / 2)

template <typename T, int MODE, int CONN>
__global__ void __launch_bounds__(CC_TILE_THREADS)
k_union_tile_items(const T* __restrict__ in, const u32* __restrict__ M, u32* __restrict__ L, Geom g, Edge<T, MODE> E,
             u32 ntx, u32 nty, EdgeQueue GQ) {
  CC_PDL_WAIT();
  extern __shared__ __align__(16) u32 smem_u32[];
  uint16_t* lab = reinterpret_cast<uint16_t*>(smem_u32);   // [CC_TILE_NODES] 16-bit parents
  u32* items = smem_u32 + CC_TILE_NODES / 2;               // [CC_TILE_ITEMS]
  u64* gq = reinterpret_cast<u64*>(items + CC_TILE_ITEMS); // [CC_TILE_GQ] edges that leave the tile
  u32* segRS = items + CC_TILE_ITEMS + 2 * CC_TILE_GQ;     // [CC_TILE_WORDS] first run id of every row segment
  u32* wS = segRS + CC_TILE_WORDS;                         // [CC_TILE_WORDS] run starts of every word
  u32* wR = wS + CC_TILE_WORDS;                            // [CC_TILE_WORDS] id of the first run that starts in the word, minus 1
  uint16_t* todo = reinterpret_cast<uint16_t*>(wR + CC_TILE_WORDS);   // [CC_TILE_WORDS]
  __shared__ u32 s_in[2], s_gn, s_gbase, s_tn, s_big;   // s_in: item counters of the current / next enumerate step
  const u32 W = (u32)g.W, sy = (u32)g.sy, sz = (u32)g.sz, sx = (u32)g.sx;
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
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) { s_in[0] = 0; s_in[1] = 0; s_gn = 0; s_tn = 0; s_big = 0; }
  __syncthreads();
  for (u32 r = threadIdx.x; r < nseg; r += blockDim.x) {
    const u32 y = y0 + (r & (TY - 1)), z = z0 + (r >> g.ty);
    u32 first = 0xFFFFFFFFu;
    if (y < sy && z < sz) {
      const u32 j = (z * sy + y) * W;
      first = __ldg(RS + j + w0);
      if (__ldg(RS + j + wend) - first > (1u << capl)) s_big = 1;   // multilabel rows with > 16 runs per word
    }
    segRS[r] = first;
  }
  for (u32 q = threadIdx.x; q < CC_TILE_WORDS; q += blockDim.x) {
    const u32 wx = q & (TW - 1), r = q >> g.tw;
    const u32 w = w0 + wx, y = y0 + (r & (TY - 1)), z = z0 + (r >> g.ty);
    u32 st = 0, rid = 0;
    if (w < W && y < sy && z < sz) {
      const u32 i = (z * sy + y) * W + w;
      const uint2 fx = __ldg(reinterpret_cast<const uint2*>(M) + 2 * (size_t)i);
      st = fx.x & ~fx.y;
      rid = __ldg(RS + i) - 1u;
    }
    wS[q] = st; wR[q] = rid;
  }
  for (u32 k = threadIdx.x; k < CC_TILE_NODES / 2; k += blockDim.x) smem_u32[k] = (2 * k) | ((2 * k + 1) << 16);
  __syncthreads();
  const bool tile_ok = s_big == 0;   // otherwise every edge of this tile goes to kernel B2

  auto push_global = [&](u32 gp, u32 gq_) {
    const u32 pos = atomicAdd(GQ.count, 1u);
    if (pos < GQ.cap) GQ.q[pos] = (u64)gp | ((u64)gq_ << 32);
    else *GQ.ovf = 1u;
  };

  // direction table: 0 Y (0,-1,0), 1 Z (0,0,-1), then the ten diagonal directions of WordEdges::diag_masks;
  // two bits per direction: d + 1
  constexpr u32 DXP = (1u << 0) | (1u << 2) | (0u << 4) | (2u << 6) | (0u << 8) | (2u << 10) | (1u << 12) | (0u << 14) | (2u << 16) | (1u << 18) | (0u << 20) | (2u << 22);
  constexpr u32 DYP = (0u << 0) | (1u << 2) | (0u << 4) | (0u << 6) | (1u << 8) | (1u << 10) | (0u << 12) | (0u << 14) | (0u << 16) | (2u << 18) | (2u << 20) | (2u << 22);
  constexpr u32 DZP = (1u << 0) | (0u << 2) | (1u << 4) | (1u << 6) | (0u << 8) | (0u << 10) | (0u << 12) | (0u << 14) | (0u << 16) | (0u << 18) | (0u << 20) | (0u << 22);

  // one edge: word q of the tile, direction t, voxel b of the word
  auto resolve = [&](const u32 item) {
    const u32 b = item & 31u, tdir = (item >> 5) & 15u, q = item >> 9;
    const u32 wx = q & (TW - 1), r = q >> g.tw;
    const int ly = (int)(r & (TY - 1)), lz = (int)(r >> g.ty);
    const int dx = (int)((DXP >> (2 * tdir)) & 3u) - 1, dy = (int)((DYP >> (2 * tdir)) & 3u) - 1, dz = (int)((DZP >> (2 * tdir)) & 3u) - 1;
    const u32 gp = wR[q] + __popc(wS[q] & (CC_FULL >> (31 - b)));
    const int xl = (int)((wx << 5) + b) + dx;            // x of q relative to the tile
    const int lyq = ly + dy, lzq = lz + dz;
    const bool inside = xl >= 0 && xl < (int)(TW << 5) && lyq >= 0 && lyq < (int)TY && lzq >= 0;
    const u32 rowP = (z0 + lz) * sy + y0 + ly;
    const u32 rowQ = (u32)((int)rowP + dy + dz * (int)sy);
    const u32 xq = (u32)((int)((w0 << 5)) + xl);
    if constexpr (MODE == MODE_EQ || MODE == MODE_DELTA || MODE == MODE_BLOCK) {
      if (tdir >= 2) {   // diagonal candidate: test the predicate on the two voxel values
        if (!E.diag((int)tdir - 2, in[(size_t)rowP * sx + ((w0 + wx) << 5) + b], in[(size_t)rowQ * sx + xq])) return;
      }
    }
    u32 gq_, rq = 0;
    if (inside) {
      rq = ((u32)lzq << g.ty) + (u32)lyq;
      const u32 qq = (rq << g.tw) + ((u32)xl >> 5);
      gq_ = wR[qq] + __popc(wS[qq] & (CC_FULL >> (31 - (xl & 31))));
    } else {
      gq_ = run_id(M, g, rowQ * W, xq);
    }
    bool local = tile_ok && inside && gp >= segRS[r];
    if (local) local = gq_ >= segRS[rq];
    if (local) {
      sm_union16(lab, (r << capl) + (gp - segRS[r]), (rq << capl) + (gq_ - segRS[rq]));
    } else {
      const u32 pos = atomicAdd(&s_gn, 1u);
      if (pos < CC_TILE_GQ) gq[pos] = (u64)gp | ((u64)gq_ << 32);
      else push_global(gp, gq_);
    }
  };

  WordEdges<T, MODE, CONN> we(in, M, g, E);
  u32 par = 0;   // which item counter this step uses
#pragma unroll 1
  for (int round = 0; round < 2; round++) {
    // ---- round 0 = straight edges of every word, round 1 = diagonals of the to-do words ----
    const u32 nwords_round = round == 0 ? (u32)CC_TILE_WORDS : s_tn;
#pragma unroll 1
    for (u32 base = 0; base < nwords_round; base += blockDim.x, par ^= 1u) {
      // -- enumerate: masks of one word per thread --
      const u32 e = base + threadIdx.x;
      u32 m[12];
#pragma unroll
      for (int k = 0; k < 12; k++) m[k] = 0;
      u32 q = 0;
      if (e < nwords_round) {
        q = round == 0 ? e : (u32)todo[e];
        const u32 wx = q & (TW - 1), r = q >> g.tw;
        const u32 w = w0 + wx, y = y0 + (r & (TY - 1)), z = z0 + (r >> g.ty);
        if (w < W && y < sy && z < sz) {
          const u32 row = z * sy + y;
          if (we.load(row * W + w, row, w, y, z)) {
            if (round == 0) {
              m[0] = we.need_y();
              m[1] = we.need_z();
              if (we.may_have_diagonals()) todo[atomicAdd(&s_tn, 1u)] = (uint16_t)q;
            } else {
              we.diag_masks(m + 2);
            }
          }
        }
      }
      // -- expand the masks into work items; slots: warp scan + one atomic per warp --
      u32 n = 0;
#pragma unroll
      for (int k = 0; k < 12; k++) n += __popc(m[k]);
      u32 inc = n;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const u32 v = __shfl_up_sync(CC_FULL, inc, o);
        if (lane >= o) inc += v;
      }
      const u32 wtot = __shfl_sync(CC_FULL, inc, 31);
      u32 wbase = 0;
      if (wtot) {
        if (lane == 31) wbase = atomicAdd(&s_in[par], wtot);
        wbase = __shfl_sync(CC_FULL, wbase, 31);
      }
      u32 pos = wbase + inc - n;
      if (n) {
        const u32 qb = q << 9;
#pragma unroll
        for (int k = 0; k < 12; k++) {
          u32 mk = m[k];
          while (mk) {
            const u32 b = __ffs(mk) - 1; mk &= mk - 1;
            const u32 item = qb | ((u32)k << 5) | b;
            if (pos < CC_TILE_ITEMS) items[pos] = item;
            else resolve(item);          // list full: resolve in place
            pos++;
          }
        }
      }
      __syncthreads();
      // -- resolve: one item per thread and step --
      const u32 ni = min(s_in[par], (u32)CC_TILE_ITEMS);
      if (threadIdx.x == 0) s_in[par ^ 1u] = 0;   // last read before the previous step's second barrier
      for (u32 k = threadIdx.x; k < ni; k += blockDim.x) resolve(items[k]);
      __syncthreads();
    }
    __syncthreads();
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
    u32 st = wS[q];
    if (!st) continue;
    const u32 r = q >> g.tw;
    const u32 g0 = wR[q] + 1u;
    const int n = __popc(st);
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
// Dense queues (round 2d; chosen on the device when the queue holds more than one edge per two bitmap words: noise
// volumes) take the second loop. A tile's edges sit next to each other in the queue, so the lanes of a warp often carry the SAME pair
// of tile roots and would all race for one link (the losers retry both finds). Every lane takes the first hop of both
// ends (after B1 that is the tile root), lanes with equal pairs elect one of them, and only that lane unites.
// Measured (profiles/r02e_experiments.md): binary 6-connected noise 512^3 B2 0.180 -> 0.143 ms, periodic 1024^3 noise
// 2.13 -> 1.80 ms; label volumes lose 5 - 10 % (two more loads and a match per edge), so they keep the plain loop.
static __global__ void __launch_bounds__(256) k_union_queue(u32* __restrict__ L, EdgeQueue GQ, u32 dense_above) {
  CC_PDL_WAIT();
  const u32 n = min(*GQ.count, GQ.cap);
  if (n <= dense_above) {
    for (u32 e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
      const u64 v = GQ.q[e];
      uf_union_h(L, (u32)v, (u32)(v >> 32));
    }
    return;
  }
  const u32 lane = threadIdx.x & 31u;
  const u32 stride = gridDim.x * blockDim.x;
  for (u32 e0 = blockIdx.x * blockDim.x + threadIdx.x - lane; e0 < n; e0 += stride) {
    const u32 e = e0 + lane;
    u64 key = ~(u64)lane;          // lanes past the end: unique keys that no edge can have (run ids are < 2^32 - 1)
    if (e < n) {
      const u64 v = GQ.q[e];
      const u32 a = __ldca(L + (u32)v), b = __ldca(L + (u32)(v >> 32));
      key = ((u64)max(a, b) << 32) | min(a, b);
    }
    const u32 grp = __match_any_sync(CC_FULL, key);
    if (e < n && (u32)(__ffs(grp) - 1) == lane && (u32)key != (u32)(key >> 32)) uf_union_h(L, (u32)key, (u32)(key >> 32));
  }
}

// Kernel B2s (fallback, does nothing unless the edge queue overflowed): one thread per bitmap word,
// every edge is united on the global forest.
template <typename T, int MODE, int CONN>
__global__ void __launch_bounds__(256)
k_union_global(const T* __restrict__ in, const u32* __restrict__ M, u32* __restrict__ L, Geom g, Edge<T, MODE> E,
               const u32* __restrict__ ovf) {
  CC_PDL_WAIT();
  if (ovf && *ovf == 0) return;     // ovf == nullptr: unconditional (host-driven redo after an overflow)
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
