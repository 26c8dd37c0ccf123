// cc3d_common.cuh — shared types for the B200 connected-components kernels.
//
// Pipeline (one volume, x fastest, voxel index i = x + sx*(y + sy*z), voxels < 2^32-1).
// The volume is reduced to FACE BITMAPS: one 32-bit word per 32 consecutive voxels of a row for the
// foreground (F) and for the three straight backward edges (X: joined to x-1, Y: to y-1, Z: to z-1).
// The nodes of the union-find are x-RUNS (maximal chains of x-links inside a row). Runs are numbered
// densely in raster order (run table), so the forest, the root flags and the final labels are dense
// arrays of one entry per run that stay L2-resident; the minimum run of a component is the run of
// its first voxel in raster order.
//
//   A  k_faces      : the only pass over the input. One voxel per lane, three compares + four
//                     ballots per 32 voxels -> F/X/Y/Z words, run starts per word, and the epl
//                     transition count (cc3d.hpp:287-315).
//   S  scan         : exclusive scan of run starts per word -> RS (id of the first run that starts in
//                     a word); L[i] = i for every run.
//   B  k_union      : one thread per bitmap word, everything word-parallel. Straight edges: an edge
//                     is dropped when kept edges already join the same two runs (x rule, square rule).
//                     Diagonal edges (8/18/26): a diagonal can only matter when no voxel "between" its
//                     end points joins them, which is decided from the face bitmaps of the neighbour
//                     rows; the handful of candidates left compare their two voxel values.
//                     Kept edges do a lock-free atomicMin link-to-smaller union of two run ids.
//   P  k_periodic   : torus wrap unions (4/8/6-connected, delta == 0).
//   C1 k_compress   : every run -> its root; root flags (+ popcounts).
//   C2 scan         : exclusive scan of root flags -> rank of every root = first-appearance order of
//                     its component.
//   C3 k_assign     : L[run] = final label.
//   D  k_expand     : out[i] = label of i's run, in the out dtype (the only dense write).
// Dense traffic is sizeof(T) + sizeof(OUT) bytes per voxel (the compulsory bytes); everything between
// A and D touches bitmaps (1/8 byte per voxel and plane) and the run table.
// Replaces cc3d.hpp:64-149 (DisjointSet), :245-285 (relabel), :344-1421 (decision-tree kernels),
// cc3d_binary.hpp:31-1263, cc3d_continuous.hpp:90-392.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

typedef uint32_t u32;
typedef uint64_t u64;
typedef int64_t i64;

#define CC_BG 0xFFFFFFFFu
#define CC_FULL 0xFFFFFFFFu

enum { MODE_EQ = 0, MODE_NONZERO = 1, MODE_DELTA = 2, MODE_MASK = 3, MODE_BLOCK = 4 };

// Backward-direction codes. Order follows the reference's compute_neighborhood
// (cc3d_continuous.hpp:38-73).
__host__ __device__ constexpr int dir_code(int dx, int dy, int dz) {
  return (dz == 0) ? ((dy == 0) ? 0 /*(-1,0,0)*/ : (dx == 0 ? 1 : (dx < 0 ? 3 : 4)))
                   : ((dy == 0) ? (dx == 0 ? 2 : (dx < 0 ? 7 : 8))
                                : (dy < 0 ? (dx == 0 ? 5 : (dx < 0 ? 9 : 10))
                                          : (dx == 0 ? 6 : (dx < 0 ? 11 : 12))));
}
// Bitmap storage M (u32 words): [0, 4*nwords) holds one uint4 {F, X, Y, Z} per bitmap word (one 16-byte
// access gives all four faces of a word); RS (nwords + 1 entries: runs that start in the word, turned by
// the scan into the id of the first of them, RS[nwords] = number of runs) follows at g.offRS; A0 / C0
// (explicit in-plane diagonals, continuous 2D-8 path only) at g.offA0 / g.offC0.
struct Q4 { u32 F, X, Y, Z; };
__device__ __forceinline__ Q4 ldq(const u32* __restrict__ M, u32 j) {
  const uint4 v = __ldg(reinterpret_cast<const uint4*>(M) + j);
  Q4 q; q.F = v.x; q.X = v.y; q.Y = v.z; q.Z = v.w;
  return q;
}

// Neighbour rows of a voxel's backward neighbourhood other than its own row:
// R0=(dy-1,dz0) R1=(dy0,dz-1) R2=(dy-1,dz-1) R3=(dy+1,dz-1); dx mask bit0: dx=-1, bit1: dx=0, bit2: dx=+1.
__host__ __device__ constexpr int hood_rows(int conn) { return conn == 4 || conn == 8 ? 1 : (conn == 6 ? 2 : 4); }
__host__ __device__ constexpr int hood_dx(int conn, int r) {
  return conn == 4 ? (r == 0 ? 2 : 0)
       : conn == 8 ? (r == 0 ? 7 : 0)
       : conn == 6 ? (r < 2 ? 2 : 0)
       : conn == 18 ? (r < 2 ? 7 : 2)
       : 7;
}

struct Geom {
  i64 sx, sy, sz;  // volume
  i64 W;           // bitmap words per row = ceil(sx/32)
  i64 rows;        // sy * sz
  i64 nwords;      // rows * W = words per bitmap plane
  i64 offRS, offA0, offC0;   // word offsets into M
  int tw, ty, tz;  // log2 of the union tile extent in words (x), rows (y) and planes (z); tw+ty+tz = 9
};
#ifndef CC_TILE_LOG
#define CC_TILE_LOG 9        // log2 of the words of a union tile (tw + ty + tz)
#endif
#define CC_TILE_WORDS (1 << CC_TILE_LOG)
#ifndef CC_TILE_TZ
#define CC_TILE_TZ 2         // log2 of the planes of a union tile (3D volumes)
#endif
#ifndef CC_TILE_THREADS
#define CC_TILE_THREADS (CC_TILE_WORDS / 2)
#endif
#define CC_TILE_NODES (CC_TILE_WORDS * 16)   // shared-memory forest: 16 runs per word (the binary maximum)
#ifndef CC_TILE_LQ_FACTOR
#define CC_TILE_LQ_FACTOR 4
#endif
#define CC_TILE_LQ (CC_TILE_WORDS * CC_TILE_LQ_FACTOR)   // queue of tile-local edges (packed 16+16 bit local run ids)
#define CC_TILE_GQ (CC_TILE_WORDS * 2)       // staging buffer of edges that leave the tile (64-bit: two run ids)
#define CC_TILE_GQ_EQ CC_TILE_WORDS          // ... for multilabel volumes (fewer edges leave a tile)
#ifndef CC_B1_MINB
#define CC_B1_MINB 6
#endif
#ifndef CC_BLOCK_MINB
#define CC_BLOCK_MINB 0      // resident CTAs per SM asked of k_union_tile on block grids (0: whatever 48 registers allow)
#endif
#ifndef CC_B1W_MINB
#define CC_B1W_MINB 6
#endif
#define CC_TILE_MINB(n) ((n) * 256 / CC_TILE_THREADS)

// Device-side results of a labelling pass. The block is valid after being ZEROED (one memset clears it together with
// the scan status words that follow it in memory), hence the encodings of the foreground row range.
struct Counters {
  u64 epl;
  u64 first_inv;       // max over rows with foreground of ~row (0 = no foreground): first row = ~first_inv
  u64 last_p1;         // max over rows with foreground of row + 1 (0 = none): last row = last_p1 - 1
  u64 N;
  u64 nruns;           // number of x-runs (scan S)
  u32 gq_count;        // entries in the global edge queue (B1 -> B2)
  u32 gq_ovf;          // the queue overflowed: the unions have to be redone on the global forest
  u32 pad[2];
};
__device__ __forceinline__ void track_rows(Counters* c, u64 first_row, u64 last_row) {
  atomicMax((unsigned long long*)&c->first_inv, (unsigned long long)~first_row);
  atomicMax((unsigned long long*)&c->last_p1, (unsigned long long)(last_row + 1));
}

template <typename T> struct is_float_t { static constexpr bool value = false; };
template <> struct is_float_t<float> { static constexpr bool value = true; };
template <> struct is_float_t<double> { static constexpr bool value = true; };

// Edge predicate between a voxel p and a neighbour q (out-of-volume neighbours are passed as 0).
//   EQ      : v[p] == v[q] != 0                      (cc3d.hpp multilabel kernels)
//   NONZERO : v[p] != 0 && v[q] != 0                 (cc3d_binary.hpp)
//   DELTA   : both non-zero and |v[p]-v[q]| <= delta in T arithmetic (cc3d_continuous.hpp:79-88)
template <typename T, int MODE> struct Edge {
  T delta;
  // DELTA, 3D 26-connected only: the reference copies the label of an EQUAL voxel at z-1 before it evaluates match()
  // (cc3d_continuous.hpp:147-150); for finite values equality implies a match, but +-inf == +-inf joins two voxels whose
  // difference is NaN. zeq = 1 reproduces that on the straight -z edge.
  int zeq;
  __device__ __forceinline__ bool fg(T v) const { return v != (T)0; }
  // BLOCK (binary 26-connected volumes, cc3d_blocks.cuh): a "voxel" is the occupancy byte of a 2x2x2 block (bit
  // x + 2y + 4z); two blocks are joined in backward direction (dx, dy, dz) when p has a voxel on the side that faces q
  // and q has one on the side that faces p (all such voxel pairs are 26-adjacent).
  static __device__ __forceinline__ u32 side_mask(int d, u32 lo, u32 hi) { return d < 0 ? lo : (d > 0 ? hi : 0xFFu); }
  static __device__ __forceinline__ bool block_edge(u32 p, u32 q, int dx, int dy, int dz) {
    const u32 mp = side_mask(dx, 0x55u, 0xAAu) & side_mask(dy, 0x33u, 0xCCu) & side_mask(dz, 0x0Fu, 0xF0u);
    const u32 mq = side_mask(-dx, 0x55u, 0xAAu) & side_mask(-dy, 0x33u, 0xCCu) & side_mask(-dz, 0x0Fu, 0xF0u);
    return (p & mp) != 0 && (q & mq) != 0;
  }
  __device__ __forceinline__ bool xedge(T p, T q) const {
    if constexpr (MODE == MODE_BLOCK) return block_edge((u32)p, (u32)q, -1, 0, 0);
    else return (*this)(p, q);
  }
  __device__ __forceinline__ bool yedge(T p, T q) const {
    if constexpr (MODE == MODE_BLOCK) return block_edge((u32)p, (u32)q, 0, -1, 0);
    else return (*this)(p, q);
  }
  __device__ __forceinline__ bool zedge(T p, T q) const {
    if constexpr (MODE == MODE_BLOCK) return block_edge((u32)p, (u32)q, 0, 0, -1);
    if constexpr (MODE == MODE_DELTA && is_float_t<T>::value) { if (zeq && p == q && p != (T)0) return true; }
    return (*this)(p, q);
  }
  // diagonal direction t of WordEdges::diag_masks: 0 A0, 1 C0, 2 A1, 3 C1, 4 B2, 5 A2, 6 C2, 7 B3, 8 A3, 9 C3
  __device__ __forceinline__ bool diag(int t, T p, T q) const {
    if constexpr (MODE == MODE_BLOCK) {
      const int dx = (int)((0x86188u >> (2 * t)) & 3u) - 1;      // -1 +1 -1 +1 0 -1 +1 0 -1 +1
      const int dy = t < 2 ? -1 : (t < 4 ? 0 : (t < 7 ? -1 : 1));
      const int dz = t < 2 ? 0 : -1;
      return block_edge((u32)p, (u32)q, dx, dy, dz);
    } else return (*this)(p, q);
  }
  __device__ __forceinline__ bool operator()(T p, T q) const {
    if constexpr (MODE == MODE_EQ) { return p == q && p != (T)0; }
    else if constexpr (MODE == MODE_NONZERO || MODE == MODE_BLOCK) { return p != (T)0 && q != (T)0; }
    else {
      if (p == (T)0 || q == (T)0) return false;
      if constexpr (is_float_t<T>::value) { return fabs(p - q) <= delta; }
      else { return (p > q ? (T)(p - q) : (T)(q - p)) <= delta; }
    }
  }
};

// Byte-parallel helpers (four 1-byte voxels / blocks per 32-bit word): bit 7 of every byte that is non-zero, and the
// gather of those four flags into a nibble (one multiply: the partial products land on distinct bits).
__device__ __forceinline__ u32 cc_nz_flags4(u32 v) { return (((v & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | v) & 0x80808080u; }
__device__ __forceinline__ u32 cc_nz_nibble4(u32 v) { return (cc_nz_flags4(v) * 0x00204081u) >> 28; }

// Lock-free union with link-to-smaller: the root of every set is its minimum index, which is what
// makes the final scan reproduce first-appearance numbering. `A` is shared or global memory.
__device__ __forceinline__ u32 uf_find(volatile u32* A, u32 i) {
  u32 p;
  while ((p = A[i]) != i) i = p;
  return i;
}
__device__ __forceinline__ void uf_union(u32* A, u32 a, u32 b) {
  bool done;
  do {
    a = uf_find(A, a);
    b = uf_find(A, b);
    if (a < b) { u32 old = atomicMin(&A[b], a); done = (old == b); b = old; }
    else if (b < a) { u32 old = atomicMin(&A[a], b); done = (old == a); a = old; }
    else done = true;
  } while (!done);
}

// Global forest (one u32 per run). Finds use ordinary L1-cached loads: a stale parent is still an
// ancestor (parents only ever move up), so a find may stop early but never leaves the set, and the
// link itself is an atomicMin whose return value is always current. This keeps the millions of finds
// that end at the root of a giant component out of one L2 slice. Path halving: every find re-points
// the nodes it passes at their grandparents.
__device__ __forceinline__ u32 uf_find_h(u32* A, u32 i) {
  u32 p = __ldca(A + i);
  while (p != i) {
    const u32 gp = __ldca(A + p);
    if (gp == p) return p;
    A[i] = gp;
    i = gp;
    p = __ldca(A + i);
  }
  return i;
}
__device__ __forceinline__ void uf_union_h(u32* A, u32 a, u32 b) {
  bool done;
  do {
    a = uf_find_h(A, a);
    b = uf_find_h(A, b);
    if (a < b) { u32 old = atomicMin(&A[b], a); done = (old == b); b = old; }
    else if (b < a) { u32 old = atomicMin(&A[a], b); done = (old == a); a = old; }
    else done = true;
  } while (!done);
}

// Shared-memory forest of a union tile: 16-bit parents (a tile holds at most 2^14 runs), link =
// compare-and-swap minimum on the containing 32-bit word.
__device__ __forceinline__ u32 sm_find16(volatile uint16_t* A, u32 i) {
  u32 p = A[i];
  while (p != i) {
    const u32 gp = A[p];
    if (gp == p) return p;
    A[i] = (uint16_t)gp;
    i = gp;
    p = A[i];
  }
  return i;
}
__device__ __forceinline__ u32 sm_min16(uint16_t* A, u32 idx, u32 val) {   // returns the previous value
  u32* Wd = reinterpret_cast<u32*>(A) + (idx >> 1);
  const int sh = (idx & 1) * 16;
  u32 old = *reinterpret_cast<volatile u32*>(Wd);
  while (true) {
    const u32 cur = (old >> sh) & 0xFFFFu;
    if (val >= cur) return cur;
    const u32 nw = (old & ~(0xFFFFu << sh)) | (val << sh);
    const u32 prev = atomicCAS(Wd, old, nw);
    if (prev == old) return cur;
    old = prev;
  }
}
__device__ __forceinline__ void sm_union16(uint16_t* A, u32 a, u32 b) {
  bool done;
  do {
    a = sm_find16(A, a);
    b = sm_find16(A, b);
    if (a < b) { u32 old = sm_min16(A, b, a); done = (old == b); b = old; }
    else if (b < a) { u32 old = sm_min16(A, a, b); done = (old == a); a = old; }
    else done = true;
  } while (!done);
}

// Programmatic dependent launch (default; -DCC_NO_PDL restores plain launches): the kernels of the label pipeline
// are launched with programmaticStreamSerialization, so a kernel's launch overlaps the tail of its predecessor; every
// such kernel begins with griddepcontrol.wait, i.e. it touches no memory before the predecessor has completed and
// flushed. Verified on B200 (145 parity tests; same-box A/B in profiles/r01_pdl_ab_*.log: 512^3 step 0.6920 ->
// 0.6780 ms, 256^3 0.2227 -> 0.2112 ms).
#if !defined(CC_NO_PDL) && !defined(CC_PDL)
#define CC_PDL 1
#endif
#ifdef CC_PDL
#define CC_PDL_WAIT() asm volatile("griddepcontrol.wait;" ::: "memory")
#else
#define CC_PDL_WAIT()
#endif

template <typename... KArgs, typename... Args>
static inline void cc_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
#ifdef CC_PDL
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
#else
  kernel<<<grid, block, smem, s>>>(static_cast<KArgs>(args)...);
#endif
}

// Block-wide exclusive scan helper (256 threads) shared by the scan kernels and the union tiles.
#define CC_SCAN_THREADS 256
#define CC_SCAN_ITEMS 16
#define CC_SCAN_CHUNK (CC_SCAN_THREADS * CC_SCAN_ITEMS)

__device__ __forceinline__ u32 block_exclusive_scan(u32 v, u32* total) {
  __shared__ u32 wsum[CC_SCAN_THREADS / 32];
  __shared__ u32 wtot;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  u32 inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const u32 n = __shfl_up_sync(CC_FULL, inc, o);
    if (lane >= o) inc += n;
  }
  if (lane == 31) wsum[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    u32 s = lane < CC_SCAN_THREADS / 32 ? wsum[lane] : 0;
    u32 si = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const u32 n = __shfl_up_sync(CC_FULL, si, o);
      if (lane >= o) si += n;
    }
    if (lane < CC_SCAN_THREADS / 32) wsum[lane] = si - s;
    if (lane == 31) wtot = si;
  }
  __syncthreads();
  const u32 r = inc - v + wsum[warp];
  if (total) *total = wtot;
  __syncthreads();
  return r;
}


// Id of the run that contains foreground voxel x of the row whose first word is j0 (= row * W):
// runs are numbered in raster order, so it is the last run that started at or before x.
__device__ __forceinline__ u32 run_id(const u32* __restrict__ M, const Geom& g, u32 j0, u32 x) {
  const u32 j = j0 + (x >> 5);
  const uint2 fx = __ldg(reinterpret_cast<const uint2*>(M) + 2 * (size_t)j);   // {F, X}
  return __ldg(M + g.offRS + j) + __popc(fx.x & ~fx.y & (CC_FULL >> (31 - (x & 31)))) - 1u;
}
