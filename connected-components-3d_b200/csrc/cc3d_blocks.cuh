// cc3d_blocks.cuh — union stage of BINARY 26-connected volumes through 2x2x2 block nodes (replaces kernels B1 / B2 for
// that configuration; cc3d_binary.hpp:31-329 is the reference's kernel for it).
//
// In a 26-connected binary image all voxels of a 2x2x2 block are mutually adjacent, so a block is ONE node, and two
// blocks in (backward) direction d are joined iff the first has a voxel on the side that faces the second and vice
// versa (Edge<., MODE_BLOCK>::block_edge). On random noise - the reference's worst case, BASELINE configs[1] - the
// run-level forest has 33 M runs and an edge for almost every one of them, and every find ends at the root of one giant
// component (B1 3.4 ms at 512^3). The block grid has 1/8 of the voxels, almost every block is occupied and almost
// every face link is set: it behaves like a smooth volume with sparse holes. So the unions are solved by running the
// SAME pipeline one level up:
//   k_block_occ      voxel foreground bitmap F -> occupancy byte per block (bit x + 2y + 4z)
//   A / S / B1 / B2  on the block grid with the BLOCK predicate (non-transitive: the continuous-value code path)
//   k_block_flatten  every block run -> its root block run
//   k_block_minrun   every voxel run -> root of the block run that holds its first voxel; per root the minimum voxel
//                    run id (= the run of the component's first voxel in raster order: first-appearance numbering
//                    is decided on VOXEL runs, block order does not matter)
//   k_block_rootflags / k_popc_n / scan / k_block_assign_rank   root flags of the voxel-run table from the block roots'
//                    minimum runs, then L[run] = first-appearance rank of its component (the C stage of this path)
// The session keeps its voxel-level run table, so every consumer (expansion, dust, slabs, statistics) is unchanged.
#pragma once
#include "cc3d_common.cuh"

// one thread per voxel bitmap word of every EVEN row pair: 16 blocks
static __global__ void __launch_bounds__(256)
k_block_occ(const u32* __restrict__ M, Geom g, uint8_t* __restrict__ occ, u32 BX, u32 BY, u32 BZ) {
  const u32 W = (u32)g.W, sy = (u32)g.sy, sz = (u32)g.sz;
  const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= W * BY * BZ) return;
  const u32 w = t % W, br = t / W;
  const u32 by = br % BY, bz = br / BY;
  const u32 y0 = 2 * by, z0 = 2 * bz;
  auto F = [&](u32 y, u32 z) -> u32 { return (y < sy && z < sz) ? __ldg(M + 4 * ((size_t)(z * sy + y) * W + w)) : 0u; };
  const u32 f00 = F(y0, z0), f10 = F(y0 + 1, z0), f01 = F(y0, z0 + 1), f11 = F(y0 + 1, z0 + 1);
  u32 o[4];
#pragma unroll
  for (int k = 0; k < 4; k++) {
    u32 v = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int sh = 2 * (4 * k + j);
      const u32 b = ((f00 >> sh) & 3u) | (((f10 >> sh) & 3u) << 2) | (((f01 >> sh) & 3u) << 4) | (((f11 >> sh) & 3u) << 6);
      v |= b << (8 * j);
    }
    o[k] = v;
  }
  uint8_t* dst = occ + (size_t)br * BX + 16 * w;
  if ((BX & 15u) == 0) {
    *reinterpret_cast<uint4*>(dst) = make_uint4(o[0], o[1], o[2], o[3]);
  } else {
    const u32 n = min(16u, BX - 16 * w);
    for (u32 j = 0; j < n; j++) dst[j] = (uint8_t)(o[j >> 2] >> (8 * (j & 3)));
  }
}

static __global__ void __launch_bounds__(256) k_block_flatten(u32* __restrict__ L2, const u64* __restrict__ n_dev) {
  const u32 n = (u32)*n_dev;
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    u32 r = i, p;
    while ((p = __ldca(&L2[r])) != r) r = p;      // a stale parent is still an ancestor
    if (r != i) L2[i] = r;
  }
}

// One thread per voxel bitmap word: every run that starts in the word -> L[run] = root block run (flattened forest L2 of
// the block grid, bitmaps M2); minrun[root] = min(run id). A warp first settles the FIRST run of its 32 words with one
// atomic per distinct root (match + min reduction: on noise every lane meets the same giant root), later runs of a word
// only act when their root differs from the previous one; a plain read of the current minimum keeps all but the first
// wave of warps off the atomic.
static __global__ void __launch_bounds__(256)
k_block_minrun(const u32* __restrict__ M, Geom g, const u32* __restrict__ M2, Geom g2, const u32* __restrict__ L2,
               u32* __restrict__ L, u32* __restrict__ minrun) {
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  const u32 W = (u32)g.W, W2 = (u32)g2.W;
  u32 starts = 0, id = 0, row2w = 0, x0 = 0;
  if (i < (u32)g.nwords) {
    const uint2 fx = __ldg(reinterpret_cast<const uint2*>(M) + 2 * (size_t)i);
    starts = fx.x & ~fx.y;
    if (starts) {
      id = __ldg(M + g.offRS + i);
      const u32 row = i / W, w = i - row * W;
      const u32 z = row / (u32)g.sy, y = row - z * (u32)g.sy;
      row2w = ((z >> 1) * (u32)g2.sy + (y >> 1)) * W2;
      x0 = w << 5;
    }
  }
  // the 16 blocks under a voxel word lie in ONE block word: its {F, X} / RS are loaded once, the roots of the word's
  // runs are then gathered four at a time (independent loads in flight instead of one dependent load per run)
  u32 cS = 0, cR = 0;
  if (starts) {
    const u32 j = row2w + (x0 >> 6);
    const uint2 fx2 = __ldg(reinterpret_cast<const uint2*>(M2) + 2 * (size_t)j);
    cS = fx2.x & ~fx2.y; cR = __ldg(M2 + g2.offRS + j) - 1u;
  }
  auto root_of = [&](u32 x) -> u32 {
    const u32 bx = x >> 1;
    const u32 br = cR + __popc(cS & (CC_FULL >> (31 - (bx & 31))));      // 32-bit sum: cR is 'first id - 1' and may be 0xFFFFFFFF
    return __ldg(L2 + br);
  };
  // first run of every word: warp-aggregated
  u32 root = 0xFFFFFFFFu;
  if (starts) {
    const int b = __ffs(starts) - 1; starts &= starts - 1;
    root = root_of(x0 + b);
    L[id] = root;
  }
  {
    const u32 active = __ballot_sync(CC_FULL, root != 0xFFFFFFFFu);
    if (root != 0xFFFFFFFFu) {
      const u32 grp = __match_any_sync(active, root);
      const u32 m = __reduce_min_sync(grp, id);
      if ((u32)(__ffs(grp) - 1) == (threadIdx.x & 31u) && m < __ldcg(&minrun[root])) atomicMin(&minrun[root], m);
    }
  }
  u32 last = root;
  while (starts) {
    u32 r[4]; int nr = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      r[k] = 0;
      if (starts) { const int b = __ffs(starts) - 1; starts &= starts - 1; r[k] = root_of(x0 + b); nr = k + 1; }
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
      if (k < nr) {
        id++;
        L[id] = r[k];
        if (r[k] != last) {
          if (id < __ldcg(&minrun[r[k]])) atomicMin(&minrun[r[k]], id);
          last = r[k];
        }
      }
    }
  }
}

// The C stage of the block path (round 2d). After k_block_minrun the first run of every component is known per ROOT BLOCK
// RUN, so the root flags of the voxel-run table are set by the (few) block roots instead of a sweep over all voxel runs
// (k_compress), and the final write L[run] = rank(minrun[root]) replaces k_block_assign + k_assign: one pass over the
// voxel runs instead of three. GR must be zero on entry.
static __global__ void __launch_bounds__(256)
k_block_rootflags(const u32* __restrict__ L2, const u32* __restrict__ minrun, u32* __restrict__ GR, const u64* __restrict__ n2_dev) {
  const u32 n = (u32)*n2_dev;
  for (u32 r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
    if (__ldg(&L2[r]) != r) continue;
    const u32 m = __ldg(&minrun[r]);
    if (m != CC_BG) atomicOr(&GR[m >> 5], 1u << (m & 31));
  }
}
static __global__ void __launch_bounds__(256)
k_popc_n(const u32* __restrict__ GR, u32* __restrict__ cnt, const u64* __restrict__ n_dev) {
  const u32 nw = (u32)((*n_dev + 31) >> 5);
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < nw; i += gridDim.x * blockDim.x) cnt[i] = __popc(GR[i]);
}
// four runs per lane: three dependent gathers per run, four chains in flight (see k_compress4)
static __global__ void __launch_bounds__(256)
k_block_assign_rank(u32* __restrict__ L, const u32* __restrict__ minrun, const u32* __restrict__ GR,
                    const u32* __restrict__ prefix, const u64* __restrict__ n_dev) {
  const u32 n = (u32)*n_dev;
  const u32 ngroups = (n + 127) >> 7;
  const int lane = threadIdx.x & 31;
  const u32 nwarps = (gridDim.x * blockDim.x) >> 5;
  for (u32 grp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; grp < ngroups; grp += nwarps) {
    const u32 base = (grp << 7) + lane;
    u32 m[4], pw[4], gw[4];
#pragma unroll
    for (int k = 0; k < 4; k++) { const u32 i = base + 32u * k; m[k] = i < n ? L[i] : 0u; }
#pragma unroll
    for (int k = 0; k < 4; k++) { const u32 i = base + 32u * k; m[k] = i < n ? __ldg(&minrun[m[k]]) : 0u; }
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const u32 i = base + 32u * k;
      pw[k] = i < n ? __ldg(&prefix[m[k] >> 5]) : 0u;
      gw[k] = i < n ? __ldg(&GR[m[k] >> 5]) : 0u;
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const u32 i = base + 32u * k;
      if (i < n) L[i] = pw[k] + __popc(gw[k] & ((1u << (m[k] & 31)) - 1u)) + 1u;
    }
  }
}
