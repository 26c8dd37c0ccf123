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
//   k_block_minrun   per root block run the minimum voxel-run id of its component (= the run of the component's first
//                    voxel in raster order: first-appearance numbering is decided on VOXEL runs, block order does not
//                    matter), from the first voxel of every block-run piece (one thread per block word)
//   k_block_rootflags / k_popc_n / scan / k_block_labels   root flags of the voxel-run table from the block roots'
//                    minimum runs, their scan, then every BLOCK run gets the rank of its component (the C stage)
//   k_expand_blocks  D: out[voxel] = label of the block run of its block (REMAP / row ranges as k_expand)
//   k_block_fill_L   only for consumers of the voxel-level run table (fused dust): L[run] = label, on demand
// Nothing sweeps the 33 M voxel runs of a noise volume any more (round 2c: three sweeps; 2d: one).
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

// One thread per BLOCK bitmap word (32 blocks of a block row): every piece of a block run inside the word reports the
// voxel run that holds the piece's first voxel in raster order; minrun[root] = the smallest of them = the run of the
// component's first voxel (a voxel run lies inside ONE block run, and the first voxel of the component is the first voxel
// of its piece). Bit b of an occupancy byte is voxel (x, y, z) = (b & 1, b >> 1 & 1, b >> 2), so voxel row q = b >> 1 of
// the block row comes first in raster order for the smallest q; per q a mask of the blocks that have a voxel in that
// row (and of those whose voxel at x even is set) comes from byte-parallel arithmetic on the 32 occupancy bytes.
// Lanes that may lower a minimum (a plain read first) elect one lane per root: on noise every piece of the first wave
// meets the same giant root.
static __global__ void __launch_bounds__(256)
k_block_minrun(const u32* __restrict__ M, Geom g, const uint8_t* __restrict__ occ, const u32* __restrict__ M2, Geom g2,
               const u32* __restrict__ L2, u32* __restrict__ minrun) {
  const u32 j2 = blockIdx.x * blockDim.x + threadIdx.x;
  const u32 W2 = (u32)g2.W, BX = (u32)g2.sx, BY = (u32)g2.sy;
  u32 F2 = 0, starts2 = 0, id2 = 0, w2 = 0, by = 0, bz = 0;
  u32 mrow[4] = {0, 0, 0, 0}, meven[4] = {0, 0, 0, 0};
  if (j2 < (u32)g2.nwords) {
    const uint2 fx2 = __ldg(reinterpret_cast<const uint2*>(M2) + 2 * (size_t)j2);
    F2 = fx2.x; starts2 = fx2.x & ~fx2.y;
  }
  if (F2) {
    id2 = __ldg(M2 + g2.offRS + j2) - 1u;      // may be 0xFFFFFFFF: 32-bit sums below
    const u32 brow = j2 / W2;
    w2 = j2 - brow * W2; bz = brow / BY; by = brow - bz * BY;
    const uint8_t* __restrict__ p = occ + ((size_t)brow * BX + (w2 << 5));
    const u32 n = min(32u, BX - (w2 << 5));
#pragma unroll
    for (int k = 0; k < 8; k++) {
      u32 v = 0;
      if ((BX & 3u) == 0) { if (4u * k < n) v = __ldg(reinterpret_cast<const u32*>(p) + k); }
      else {
#pragma unroll
        for (int b = 0; b < 4; b++) if (4u * k + b < n) v |= (u32)p[4 * k + b] << (8 * b);
      }
#pragma unroll
      for (int q = 0; q < 4; q++) {
        mrow[q] |= cc_nz_nibble4(v & (0x03030303u << (2 * q))) << (4 * k);
        meven[q] |= cc_nz_nibble4(v & (0x01010101u << (2 * q))) << (4 * k);
      }
    }
  }
  const u32 cont = F2 & ~starts2;      // blocks that continue a run from their left neighbour
  u32 rem = F2;
  while (__any_sync(CC_FULL, rem != 0)) {
    u32 root = 0, id = 0;
    bool need = false;
    if (rem) {
      const int s = __ffs(rem) - 1;
      const u32 t = s == 31 ? 0u : (cont >> (s + 1));
      const int len = 1 + (__ffs(~t) - 1);                       // ~t != 0: t has at most 31 - s bits
      const u32 pm = (len >= 32 ? CC_FULL : ((1u << len) - 1u)) << s;
      rem &= ~pm;
      root = __ldg(L2 + (id2 + __popc(starts2 & (CC_FULL >> (31 - s)))));
      // first voxel row of the block row that the piece touches (it holds at least one voxel); no dynamic indexing
      const u32 c0 = mrow[0] & pm, c1 = mrow[1] & pm, c2 = mrow[2] & pm, c3 = mrow[3] & pm;
      const int q = c0 ? 0 : (c1 ? 1 : (c2 ? 2 : 3));
      const u32 cm = c0 ? c0 : (c1 ? c1 : (c2 ? c2 : c3));
      const u32 ev = c0 ? meven[0] : (c1 ? meven[1] : (c2 ? meven[2] : meven[3]));
      const int b = __ffs(cm) - 1;
      const u32 x = 2u * ((w2 << 5) + b) + (((ev >> b) & 1u) ? 0u : 1u);
      const u32 y = 2u * by + (q & 1), z = 2u * bz + (q >> 1);
      id = run_id(M, g, (z * (u32)g.sy + y) * (u32)g.W, x);
      need = id < __ldcg(&minrun[root]);
    }
    const u32 active = __ballot_sync(CC_FULL, need);
    if (need) {
      const u32 grp = __match_any_sync(active, root);
      const u32 m = __reduce_min_sync(grp, id);
      if ((u32)(__ffs(grp) - 1) == (threadIdx.x & 31u)) atomicMin(&minrun[root], m);
    }
  }
}

// The C stage of the block path (round 2d). After k_block_minrun the first run of every component is known per ROOT BLOCK
// RUN, so the root flags of the voxel-run table are set by the (few) block roots instead of a sweep over all voxel runs
// (k_compress), and every BLOCK run takes the rank of its component (k_block_labels): no sweep over the voxel runs at all.
// GR must be zero on entry.
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
// every block run -> the first-appearance rank of its component, in place (L2: root block run -> label)
static __global__ void __launch_bounds__(256)
k_block_labels(u32* __restrict__ L2, const u32* __restrict__ minrun, const u32* __restrict__ GR,
               const u32* __restrict__ prefix, const u64* __restrict__ n2_dev) {
  const u32 n = (u32)*n2_dev;
  for (u32 r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
    const u32 m = __ldg(&minrun[L2[r]]);      // every thread reads and writes its own entry of L2 only
    L2[r] = __ldg(&prefix[m >> 5]) + __popc(__ldg(&GR[m >> 5]) & ((1u << (m & 31)) - 1u)) + 1u;
  }
}

// D for the block path: out[voxel] = label of the block run of its 2x2x2 block (same walk as k_expand: one warp per
// (row, chunk of 32 voxel words); lane j stages voxel word j and the block word above it, then one voxel per lane).
// The voxel-level run labels are not needed for this: they are only filled in (k_block_fill_L) for the consumers that
// work on the run table (fused dust).
template <typename OUT, int REMAP>
__global__ void __launch_bounds__(256)
k_expand_blocks(const u32* __restrict__ LB, const u32* __restrict__ M, const u32* __restrict__ M2, OUT* __restrict__ out,
                Geom g, Geom g2, unsigned nchunks, u32 row0, u32 nwarps_total, const void* __restrict__ remap) {
  __shared__ uint4 s_words[8][32];   // per warp: {F of the voxel word, run starts of the block word, id of the run that enters it, -}
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const u32 wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (wid >= nwarps_total) return;
  const u32 rrel = wid / nchunks;
  const u32 chunk = wid - rrel * nchunks;
  const u32 row = row0 + rrel;
  const u32 W = (u32)g.W, sx = (u32)g.sx, sy = (u32)g.sy;
  const u32 z = row / sy, y = row - z * sy;
  const u32 row2w = ((z >> 1) * (u32)g2.sy + (y >> 1)) * (u32)g2.W;
  const u32 wl = (chunk << 5) + lane;
  {
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (wl < W) {
      v.x = __ldg(M + 4 * ((size_t)row * W + wl));
      const u32 j2 = row2w + (wl >> 1);
      const uint2 fx2 = __ldg(reinterpret_cast<const uint2*>(M2) + 2 * (size_t)j2);
      v.y = fx2.x & ~fx2.y;
      v.z = __ldg(M2 + g2.offRS + j2) - 1u;
    }
    s_words[warp][lane] = v;
  }
  __syncwarp();
  const u32 nwd = min(32u, W - (chunk << 5));
  const u32 bit = 1u << lane;
  const u32 below0 = CC_FULL >> (31 - (lane >> 1)), below1 = CC_FULL >> (15 - (lane >> 1));   // block bit of the lane in an even / odd voxel word
  u32 x = (chunk << 10) + lane;
  OUT* __restrict__ o = out + ((size_t)rrel * sx + x);
  auto label_of = [&](const uint4 wv, u32 j) -> OUT {
    OUT v = 0;
    if (wv.x & bit) {
      const u32 lab = LB[wv.z + __popc(wv.y & ((j & 1u) ? below1 : below0))];
      if (REMAP == 1) v = (OUT)__ldg(reinterpret_cast<const u32*>(remap) + lab);
      else if (REMAP == 2) v = (OUT)__ldg(reinterpret_cast<const u64*>(remap) + lab);
      else v = (OUT)lab;
    }
    return v;
  };
  const u32 nfull = (x - lane + (nwd << 5) <= sx) ? nwd : nwd - 1;   // words that lie fully inside the row
#pragma unroll 4
  for (u32 j = 0; j < nfull; j++) o[j << 5] = label_of(s_words[warp][j], j);
  if (nfull < nwd) {
    const OUT v = label_of(s_words[warp][nfull], nfull);
    if (x + (nfull << 5) < sx) o[nfull << 5] = v;
  }
}

// The voxel-level run labels of a block-path session, on demand: L[run] = label of the block run that holds the run's
// first voxel. One thread per voxel bitmap word; the 16 blocks under it lie in one block word.
static __global__ void __launch_bounds__(256)
k_block_fill_L(const u32* __restrict__ M, Geom g, const u32* __restrict__ M2, Geom g2, const u32* __restrict__ LB,
               u32* __restrict__ L) {
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (u32)g.nwords) return;
  const uint2 fx = __ldg(reinterpret_cast<const uint2*>(M) + 2 * (size_t)i);
  u32 starts = fx.x & ~fx.y;
  if (!starts) return;
  u32 id = __ldg(M + g.offRS + i);
  const u32 W = (u32)g.W;
  const u32 row = i / W, w = i - row * W;
  const u32 z = row / (u32)g.sy, y = row - z * (u32)g.sy;
  const u32 j2 = ((z >> 1) * (u32)g2.sy + (y >> 1)) * (u32)g2.W + (w >> 1);
  const uint2 fx2 = __ldg(reinterpret_cast<const uint2*>(M2) + 2 * (size_t)j2);
  const u32 cS = fx2.x & ~fx2.y, cR = __ldg(M2 + g2.offRS + j2) - 1u;
  const u32 half = (w & 1u) << 4;
  while (starts) {
    const int b = __ffs(starts) - 1; starts &= starts - 1;
    L[id++] = __ldg(LB + (cR + __popc(cS & (CC_FULL >> (31 - (half + (b >> 1)))))));
  }
}
