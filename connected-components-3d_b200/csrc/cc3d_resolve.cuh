// cc3d_resolve.cuh — kernels C1 (compress), C2 (scan), C3 (assign), D (write) and the
// block-order renumbering used by the binary 2D 8-connected path. See cc3d_common.cuh.
#pragma once
#include "cc3d_common.cuh"

// C1: one thread per word of the local-root bitmap. Every local root is pointed straight at its
// global root; the word of global roots and its popcount are emitted for the scan.
__global__ void __launch_bounds__(256)
k_compress(u32* __restrict__ L, const u32* __restrict__ LR, u32* __restrict__ GR, u32* __restrict__ cnt,
           Geom g, i64 nwords) {
  const i64 w = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= nwords) return;
  u32 bits = LR[w];
  u32 gr = 0;
  if (bits) {
    const i64 row = w / g.W;
    const i64 base = row * g.sx + (w - row * g.W) * 32;
    while (bits) {
      const int b = __ffs(bits) - 1;
      bits &= bits - 1;
      const u32 l = (u32)(base + b);
      u32 r = l, p;
      while ((p = __ldcg(&L[r])) != r) r = p;
      if (r == l) gr |= 1u << b;
      else L[l] = r;
    }
  }
  GR[w] = gr;
  cnt[w] = __popc(gr);
}

// C2: exclusive scan of u32 counts, three small kernels (block reduce, scan of block sums, apply).
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

__global__ void __launch_bounds__(CC_SCAN_THREADS)
k_scan_reduce(const u32* __restrict__ cnt, u64* __restrict__ bsum, i64 n) {
  const i64 base = (i64)blockIdx.x * CC_SCAN_CHUNK;
  u32 s = 0;
#pragma unroll
  for (int k = 0; k < CC_SCAN_ITEMS; k++) {
    const i64 i = base + k * CC_SCAN_THREADS + threadIdx.x;
    if (i < n) s += cnt[i];
  }
  u32 tot;
  block_exclusive_scan(s, &tot);
  if (threadIdx.x == 0) bsum[blockIdx.x] = tot;
}

// single block: exclusive scan of the block sums (64-bit), total -> *N
__global__ void __launch_bounds__(1024) k_scan_blocks(u64* __restrict__ bsum, i64 nb, u64* __restrict__ N) {
  __shared__ u64 carry;
  __shared__ u64 wsum[32];
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (i64 base = 0; base < nb; base += 1024) {
    const i64 i = base + threadIdx.x;
    const u64 v = i < nb ? bsum[i] : 0;
    u64 inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const u64 n = __shfl_up_sync(CC_FULL, inc, o);
      if (lane >= o) inc += n;
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      const u64 s = wsum[lane];
      u64 si = s;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const u64 n = __shfl_up_sync(CC_FULL, si, o);
        if (lane >= o) si += n;
      }
      wsum[lane] = si - s;
    }
    __syncthreads();
    const u64 c = carry;
    if (i < nb) bsum[i] = c + wsum[warp] + inc - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry = c + wsum[warp] + inc;
    __syncthreads();
  }
  if (threadIdx.x == 0) *N = carry;
}

// prefix[i] = exclusive prefix of cnt (mod 2^32 is fine: ranks are < voxels < 2^32)
__global__ void __launch_bounds__(CC_SCAN_THREADS)
k_scan_apply(const u32* __restrict__ cnt, const u64* __restrict__ bsum, u32* __restrict__ prefix, i64 n) {
  const i64 base = (i64)blockIdx.x * CC_SCAN_CHUNK + (i64)threadIdx.x * CC_SCAN_ITEMS;
  u32 v[CC_SCAN_ITEMS];
  u32 s = 0;
#pragma unroll
  for (int k = 0; k < CC_SCAN_ITEMS; k++) {
    const i64 i = base + k;
    v[k] = i < n ? cnt[i] : 0;
    s += v[k];
  }
  u32 ex = block_exclusive_scan(s, nullptr) + (u32)bsum[blockIdx.x];
#pragma unroll
  for (int k = 0; k < CC_SCAN_ITEMS; k++) {
    const i64 i = base + k;
    if (i < n) prefix[i] = ex;
    ex += v[k];
  }
}

__device__ __forceinline__ u32 rank_in_bitmap(const u32* __restrict__ bm, const u32* __restrict__ prefix, i64 word, int bit) {
  return prefix[word] + __popc(bm[word] & ((1u << bit) - 1u));
}

// C3: every local root gets its component's final label (1-based rank of its global root).
__global__ void __launch_bounds__(256)
k_assign(u32* __restrict__ L, const u32* __restrict__ LR, const u32* __restrict__ GR,
         const u32* __restrict__ prefix, Geom g, i64 nwords) {
  const i64 w = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= nwords) return;
  u32 bits = LR[w];
  if (!bits) return;
  const i64 row = w / g.W;
  const i64 base = row * g.sx + (w - row * g.W) * 32;
  const u32 gw = GR[w];
  const u32 pw = prefix[w];
  while (bits) {
    const int b = __ffs(bits) - 1;
    bits &= bits - 1;
    const u32 l = (u32)(base + b);
    u32 label;
    if ((gw >> b) & 1u) {
      label = pw + __popc(gw & ((1u << b) - 1u)) + 1u;
    } else {
      const u32 r = L[l];
      const i64 rrow = r / g.sx;
      const i64 rx = r - rrow * g.sx;
      label = rank_in_bitmap(GR, prefix, rrow * g.W + (rx >> 5), (int)(rx & 31)) + 1u;
    }
    L[l] = label;
  }
}

// D: final write. A voxel either is a local root (its L entry already is the label) or points at one.
// Each thread handles 4 consecutive voxels of one row (128-bit load of L, one gather per distinct
// pointer, one vector store); a CTA of 128 threads covers a 512-voxel chunk of a row.
template <typename OUT> struct Out4;
template <> struct Out4<uint16_t> { typedef ushort4 type; };
template <> struct Out4<uint32_t> { typedef uint4 type; };
template <> struct Out4<uint64_t> { typedef ulonglong4 type; };

// REMAP: 0 = write the local label, 1/2 = write remap[label] from a u32/u64 table (sharded volumes:
// per-slab label -> global label). row0 = first row to write; out is indexed relative to row0.
template <typename OUT, bool VEC, int REMAP>
__global__ void __launch_bounds__(128)
k_write(const u32* __restrict__ L, const u32* __restrict__ LR, OUT* __restrict__ out, Geom g, unsigned nchunks,
        unsigned row0, const void* __restrict__ remap) {
  const unsigned rrel = blockIdx.x / nchunks;
  const unsigned chunk = blockIdx.x - rrel * nchunks;
  const unsigned row = row0 + rrel;
  const i64 x = (i64)chunk * 512 + threadIdx.x * 4;
  if (x >= g.sx) return;
  const i64 base = (i64)row * g.sx + x;
  const i64 obase = (i64)rrel * g.sx + x;
  const u32 lrw = LR[(i64)row * g.W + (x >> 5)] >> (x & 31);
  u32 l[4];
  if (VEC) {
    const uint4 t = *reinterpret_cast<const uint4*>(L + base);
    l[0] = t.x; l[1] = t.y; l[2] = t.z; l[3] = t.w;
  } else {
#pragma unroll
    for (int i = 0; i < 4; i++) l[i] = (x + i < g.sx) ? L[base + i] : CC_BG;
  }
  OUT lab[4];
  u32 prev_ptr = CC_BG;
  OUT prev_lab = 0;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    OUT v = 0;
    if (l[i] != CC_BG) {
      const bool isroot = (lrw >> i) & 1u;
      if (!isroot && l[i] == prev_ptr) v = prev_lab;
      else {
        const u32 loc = isroot ? l[i] : __ldg(&L[l[i]]);
        if (REMAP == 1) v = (OUT)__ldg(reinterpret_cast<const u32*>(remap) + loc);
        else if (REMAP == 2) v = (OUT)__ldg(reinterpret_cast<const u64*>(remap) + loc);
        else v = (OUT)loc;
        if (!isroot) { prev_ptr = l[i]; prev_lab = v; }
      }
    }
    lab[i] = v;
  }
  if (VEC) {
    typename Out4<OUT>::type o;
    o.x = lab[0]; o.y = lab[1]; o.z = lab[2]; o.w = lab[3];
    *reinterpret_cast<typename Out4<OUT>::type*>(out + obase) = o;
  } else {
#pragma unroll
    for (int i = 0; i < 4; i++) if (x + i < g.sx) out[obase + i] = lab[i];
  }
}

// ---- sharded volumes (z-slabs): equivalences across one slab interface ----
// P = first plane of the upper slab (later in raster order), Q = last plane of the lower slab.
// Emits (label in Q's slab, label in P's slab) for every edge of the chosen predicate/neighbourhood
// between the two planes, skipping an edge when the voxel to the left already produced the same pair.
// Replaces the Python face loops of connected_components_stack (cc3d/__init__.py:425-468).
template <typename T, int MODE>
__global__ void __launch_bounds__(256)
k_face_pairs(const T* __restrict__ vP, const u32* __restrict__ lP, const T* __restrict__ vQ, const u32* __restrict__ lQ,
             i64 sx, i64 sy, int connectivity, Edge<T, MODE> E, u64* __restrict__ pairs, unsigned long long cap,
             unsigned long long* __restrict__ count) {
  const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  u64 mine[9];
  int n = 0;
  if (i < sx * sy) {
    const i64 y = i / sx, x = i - y * sx;
    const T p = vP[i];
    if (E.fg(p)) {
      const u32 lp = lP[i];
      const bool left_same = x > 0 && lP[i - 1] == lp;
#pragma unroll
      for (int dy = -1; dy <= 1; dy++) {
#pragma unroll
        for (int dx = -1; dx <= 1; dx++) {
          const int nz = (dx != 0) + (dy != 0);
          if (connectivity == 6 && nz > 0) continue;
          if (connectivity == 18 && nz > 1) continue;
          const i64 xx = x + dx, yy = y + dy;
          if (xx < 0 || xx >= sx || yy < 0 || yy >= sy) continue;
          const i64 qi = yy * sx + xx;
          const T q = vQ[qi];
          if (!E(p, q, dir_code(dx, dy, -1))) continue;
          const u32 lq = lQ[qi];
          // the voxel to the left emits the same (lq, lp) through the same direction
          if (left_same && xx > 0 && lQ[qi - 1] == lq && E(vP[i - 1], vQ[qi - 1], dir_code(dx, dy, -1))) continue;
          mine[n++] = ((u64)lq << 32) | lp;
        }
      }
    }
  }
  // warp-aggregated append
  u32 total = n;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const u32 t = __shfl_up_sync(CC_FULL, total, o);
    if (lane >= o) total += t;
  }
  const u32 wtotal = __shfl_sync(CC_FULL, total, 31);
  unsigned long long base = 0;
  if (lane == 31 && wtotal) base = atomicAdd(count, (unsigned long long)wtotal);
  base = __shfl_sync(CC_FULL, base, 31);
  const unsigned long long off = base + total - n;
  for (int k = 0; k < n; k++)
    if (off + k < cap) pairs[off + k] = mine[k];
}

// union-find over compact node ids (pairs given as two u32 arrays); parent must hold 0..n-1 on entry
__global__ void __launch_bounds__(256) k_iota(u32* __restrict__ p, i64 n) {
  const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = (u32)i;
}
__global__ void __launch_bounds__(256)
k_union_pairs(u32* __restrict__ parent, const u32* __restrict__ a, const u32* __restrict__ b, i64 n) {
  const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) uf_union(parent, a[i], b[i]);
}
__global__ void __launch_bounds__(256) k_flatten(u32* __restrict__ parent, i64 n) {
  const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  u32 r = (u32)i, p;
  while ((p = __ldcg(&parent[r])) != r) r = p;
  parent[i] = r;
}

// ---- binary 2D 8-connected: number components by their first 2x2 block in block-raster order
// (cc3d_binary.hpp:1016-1023, 1215-1231). K[root] = min block key over the component. ----
__global__ void __launch_bounds__(256)
k_blockkey_init(u32* __restrict__ K, const u32* __restrict__ GR, Geom g, i64 nwords) {
  const i64 w = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= nwords) return;
  u32 bits = GR[w];
  const i64 row = w / g.W;
  const i64 base = row * g.sx + (w - row * g.W) * 32;
  while (bits) {
    const int b = __ffs(bits) - 1;
    bits &= bits - 1;
    K[base + b] = CC_BG;
  }
}

__global__ void __launch_bounds__(256)
k_blockkey_min(const u32* __restrict__ L, const u32* __restrict__ LR, u32* __restrict__ K, Geom g) {
  const i64 wid = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const i64 row = wid / g.W;
  if (row >= g.sy * g.sz) return;
  const i64 seg = wid - row * g.W;
  const i64 x = seg * 32 + lane;
  u32 root = CC_BG;
  if (x < g.sx) {
    const i64 i = row * g.sx + x;
    const u32 l = L[i];
    if (l != CC_BG) {
      const u32 lr = LR[wid];
      // after k_compress: local roots hold their global root (or themselves)
      root = ((lr >> lane) & 1u) ? l : L[l];
    }
  }
  // one atomic per run of equal roots inside the warp
  const u32 prev = __shfl_up_sync(CC_FULL, root, 1);
  const bool head = root != CC_BG && (lane == 0 || prev != root);
  if (head) {
    const i64 osx = (g.sx + 1) >> 1;
    const u32 key = (u32)((x >> 1) + osx * (row >> 1));
    atomicMin(&K[root], key);
  }
}

__global__ void __launch_bounds__(256)
k_blockkey_mark(const u32* __restrict__ K, const u32* __restrict__ GR, u32* __restrict__ BK, Geom g, i64 nwords) {
  const i64 w = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= nwords) return;
  u32 bits = GR[w];
  const i64 row = w / g.W;
  const i64 base = row * g.sx + (w - row * g.W) * 32;
  while (bits) {
    const int b = __ffs(bits) - 1;
    bits &= bits - 1;
    const u32 k = K[base + b];
    atomicOr(&BK[k >> 5], 1u << (k & 31));
  }
}

__global__ void __launch_bounds__(256) k_popc(const u32* __restrict__ bm, u32* __restrict__ cnt, i64 n) {
  const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) cnt[i] = __popc(bm[i]);
}

__global__ void __launch_bounds__(256)
k_assign_blockorder(u32* __restrict__ L, const u32* __restrict__ LR, const u32* __restrict__ GR,
                    const u32* __restrict__ K, const u32* __restrict__ BK, const u32* __restrict__ bprefix,
                    Geom g, i64 nwords) {
  const i64 w = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= nwords) return;
  u32 bits = LR[w];
  if (!bits) return;
  const i64 row = w / g.W;
  const i64 base = row * g.sx + (w - row * g.W) * 32;
  const u32 gw = GR[w];
  while (bits) {
    const int b = __ffs(bits) - 1;
    bits &= bits - 1;
    const u32 l = (u32)(base + b);
    const u32 r = ((gw >> b) & 1u) ? l : L[l];
    const u32 k = K[r];
    L[l] = rank_in_bitmap(BK, bprefix, k >> 5, (int)(k & 31)) + 1u;
  }
}
