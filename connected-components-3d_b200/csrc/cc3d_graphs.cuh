// cc3d_graphs.cuh — the callers either side of the labelling path:
//   k_voxel_graph   voxel connectivity graph of a label image      (cc3d_graphs.hpp:31-247)
//   k_vcg_union     colouring of a voxel connectivity graph         (cc3d_graphs.hpp:583-1106): voxel-level
//                   lock-free union-find on the output array itself; C1/C2/C3 of the labelling path then
//                   number the roots in first-appearance order
//   k_remap_labels  out[i] = table[labels[i]]                        (largest_k, cc3d/__init__.py:199-279)
#pragma once
#include "cc3d_common.cuh"

// ---- voxel connectivity graph -----------------------------------------------------------------
// Bit b of graph[p] stays set unless the neighbour in direction b exists and holds a different value
// (background is a value like any other). Direction tables in the reference's bit order:
//   2D (4/8): +x -x +y -y | +x+y -x+y +x-y -x-y
//   3D (6/18/26): +x -x +y -y +z -z | +x+y -x+y +x-y -x-y +x+z -x+z +y+z -y+z +x-z -x-z +y-z -y-z |
//                 +++ -++ +-+ --+ ++- -+- +-- ---
__constant__ signed char c_vcg_dir3[26][3] = {
  {1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, 1}, {0, 0, -1},
  {1, 1, 0}, {-1, 1, 0}, {1, -1, 0}, {-1, -1, 0}, {1, 0, 1}, {-1, 0, 1}, {0, 1, 1}, {0, -1, 1},
  {1, 0, -1}, {-1, 0, -1}, {0, 1, -1}, {0, -1, -1},
  {1, 1, 1}, {-1, 1, 1}, {1, -1, 1}, {-1, -1, 1}, {1, 1, -1}, {-1, 1, -1}, {1, -1, -1}, {-1, -1, -1}};
__constant__ signed char c_vcg_dir2[8][2] = {{1, 0}, {-1, 0}, {0, 1}, {0, -1}, {1, 1}, {-1, 1}, {1, -1}, {-1, -1}};

// One thread per voxel, x fastest. NDIR = number of directions (bits) of the connectivity; TWO_D selects
// the 2D table. The neighbour reads of a warp fall into the same few cache lines as its own row.
template <typename T, typename OUT, int NDIR, bool TWO_D>
__global__ void __launch_bounds__(256)
k_voxel_graph(const T* __restrict__ in, OUT* __restrict__ graph, i64 sx, i64 sy, i64 sz) {
  const i64 voxels = sx * sy * sz;
  const i64 stride = (i64)gridDim.x * blockDim.x;
  for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < voxels; i += stride) {
    const i64 x = i % sx, t = i / sx;
    const i64 y = t % sy, z = t / sy;
    const T cur = in[i];
    u32 bits = NDIR == 32 ? 0xFFFFFFFFu : ((1u << NDIR) - 1u);
#pragma unroll
    for (int d = 0; d < NDIR; d++) {
      const int dx = TWO_D ? c_vcg_dir2[d][0] : c_vcg_dir3[d][0];
      const int dy = TWO_D ? c_vcg_dir2[d][1] : c_vcg_dir3[d][1];
      const int dz = TWO_D ? 0 : c_vcg_dir3[d][2];
      const i64 xx = x + dx, yy = y + dy, zz = z + dz;
      if (xx < 0 || xx >= sx || yy < 0 || yy >= sy || zz < 0 || zz >= sz) continue;
      if (in[(zz * sy + yy) * sx + xx] != cur) bits &= ~(1u << d);
    }
    graph[i] = (OUT)bits;
  }
}

// ---- colouring of a voxel connectivity graph ---------------------------------------------------
// The reference follows the BACKWARD bits of every voxel (cc3d_graphs.hpp:746-826, 1018-1074): an edge
// p - (p + d) exists iff the neighbour is inside the volume and bit mask[d] of vcg[p] is set.
struct VcgDirs { int n; signed char d[13][3]; u32 mask[13]; };

template <typename V>
__global__ void __launch_bounds__(256)
k_vcg_union(const V* __restrict__ vcg, u32* __restrict__ parent, i64 sx, i64 sy, i64 sz, VcgDirs D) {
  const i64 voxels = sx * sy * sz;
  const i64 stride = (i64)gridDim.x * blockDim.x;
  for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < voxels; i += stride) {
    const u32 v = (u32)vcg[i];
    if (!v) continue;
    const i64 x = i % sx, t = i / sx;
    const i64 y = t % sy, z = t / sy;
    for (int k = 0; k < D.n; k++) {
      if (!(v & D.mask[k])) continue;
      const i64 xx = x + D.d[k][0], yy = y + D.d[k][1], zz = z + D.d[k][2];
      if (xx < 0 || xx >= sx || yy < 0 || yy >= sy || zz < 0) continue;
      uf_union_h(parent, (u32)i, (u32)((zz * sy + yy) * sx + xx));
    }
  }
}

// ---- out[i] = table[labels[i]] (labels above N map to 0) ----------------------------------------
template <typename LT, typename OT>
__global__ void __launch_bounds__(256)
k_remap_labels(const LT* __restrict__ labels, const u32* __restrict__ table, u64 N, OT* __restrict__ out, i64 n) {
  const i64 stride = (i64)gridDim.x * blockDim.x;
  for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const u64 l = (u64)labels[i];
    out[i] = l <= N ? (OT)__ldg(table + l) : (OT)0;
  }
}
