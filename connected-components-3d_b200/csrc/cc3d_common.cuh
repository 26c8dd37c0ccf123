// cc3d_common.cuh — shared types for the B200 connected-components kernels.
//
// Pipeline (one volume, x fastest, index i = x + sx*(y + sy*z), voxels < 2^32-1):
//   A  k_tile_label   : TXxTYxTZ tiles labelled in shared memory; L[i] = raster index of the tile-local
//                       root (the tile component's minimum raster index), LR bitmap = local roots,
//                       XS slots = x-seam equivalences found through the one-column halo.
//   B1 k_seam_rows    : unions across y/z tile seams (atomicMin link-to-smaller on L).
//   B2 k_seam_x       : unions recorded in XS (x tile seams).
//   P  k_periodic     : torus wrap unions (4/8/6-connected, delta == 0).
//   C1 k_compress     : every local root -> its global root; GR bitmap = global roots (+ popcounts).
//   C2 scan           : exclusive scan of GR popcounts -> rank of every global root = first-appearance
//                       order of its component (roots are minimum raster indices).
//   C3 k_assign       : L[local root] = final label.
//   D  k_write        : out[i] = final label in the out dtype.
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

enum { MODE_EQ = 0, MODE_NONZERO = 1, MODE_DELTA = 2, MODE_MASK = 3 };

// Backward-direction codes (bit positions in MODE_MASK inputs). Order follows the reference's
// compute_neighborhood (cc3d_continuous.hpp:38-73).
__host__ __device__ constexpr int dir_code(int dx, int dy, int dz) {
  return (dz == 0) ? ((dy == 0) ? 0 /*(-1,0,0)*/ : (dx == 0 ? 1 : (dx < 0 ? 3 : 4)))
                   : ((dy == 0) ? (dx == 0 ? 2 : (dx < 0 ? 7 : 8))
                                : (dy < 0 ? (dx == 0 ? 5 : (dx < 0 ? 9 : 10))
                                          : (dx == 0 ? 6 : (dx < 0 ? 11 : 12))));
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
__host__ __device__ constexpr int row_dy(int r) { return r == 0 ? -1 : (r == 1 ? 0 : (r == 2 ? -1 : 1)); }
__host__ __device__ constexpr int row_dz(int r) { return r == 0 ? 0 : -1; }

struct Geom {
  i64 sx, sy, sz;     // volume
  int TY, TZ;         // tile extent in y and z (TX is a template parameter)
  i64 ntx, nty, ntz;  // tiles per axis
  i64 W;              // bitmap words per row = ceil(sx/32)
};

struct Counters {      // device-side results of a labelling pass
  u64 epl;
  i64 first_row;       // initialised to INT64_MAX
  i64 last_row;        // initialised to -1
  u64 N;
};

template <typename T> struct is_float_t { static constexpr bool value = false; };
template <> struct is_float_t<float> { static constexpr bool value = true; };
template <> struct is_float_t<double> { static constexpr bool value = true; };

// Edge predicate between a voxel p and an EARLIER (in raster order) neighbour q.
//   EQ      : v[p] == v[q] != 0                      (cc3d.hpp multilabel kernels)
//   NONZERO : v[p] != 0 && v[q] != 0                 (cc3d_binary.hpp)
//   DELTA   : both non-zero and |v[p]-v[q]| <= delta in T arithmetic (cc3d_continuous.hpp:79-88)
//   MASK    : bit `dir` of p's value (precomputed backward-edge bitfield; top bit = foreground)
template <typename T, int MODE> struct Edge {
  T delta;
  __device__ __forceinline__ bool fg(T v) const {
    if constexpr (MODE == MODE_MASK) return (v >> (8 * sizeof(T) - 1)) & 1;
    else return v != (T)0;
  }
  __device__ __forceinline__ bool operator()(T p, T q, int dir) const {
    if constexpr (MODE == MODE_EQ) { return p == q && p != (T)0; }
    else if constexpr (MODE == MODE_NONZERO) { return p != (T)0 && q != (T)0; }
    else if constexpr (MODE == MODE_DELTA) {
      if (p == (T)0 || q == (T)0) return false;
      if constexpr (is_float_t<T>::value) { return fabs(p - q) <= delta; }
      else { return (p > q ? (T)(p - q) : (T)(q - p)) <= delta; }
    } else {
      return (p >> dir) & 1;
    }
  }
};

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
