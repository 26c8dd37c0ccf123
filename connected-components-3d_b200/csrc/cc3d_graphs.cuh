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

// ---- contacts / region graph (cc3d_graphs.hpp:300-468) ------------------------------------------------
// For every voxel and every backward direction of the connectivity: if both voxels are non-zero and differ, the
// pair (min, max) gains one contact of the direction's class (0: across x, 1: across y, 2: across z, 3: edge or
// corner). Offsets are computed exactly as the reference's compute_neighborhood does (cc3d_graphs.hpp:259-313),
// including its behaviour at the volume border, where some diagonal offsets degenerate to a face / edge neighbour.
// Lanes of a warp that found the same pair in the same direction add once (match_any) to an open-addressing hash
// table in global memory: keys = min << 32 | max (labels must be < 2^32), four u32 counters per key.
struct ContactTable { unsigned long long* keys; u32* vals; u32 mask; u32* flags; };   // flags[0]: table full, flags[1]: label >= 2^32

__device__ __forceinline__ void contact_add(const ContactTable& tb, unsigned long long key, int cls, u32 n) {
  u32 h = (u32)((key * 0x9E3779B97F4A7C15ull) >> 32) & tb.mask;
#pragma unroll 1
  for (int probe = 0; probe < 128; probe++) {
    const u32 s = (h + probe) & tb.mask;
    unsigned long long k = *(volatile unsigned long long*)&tb.keys[s];
    if (k == 0) k = atomicCAS(&tb.keys[s], 0ull, key);
    if (k == 0 || k == key) { atomicAdd(&tb.vals[4 * (size_t)s + cls], n); return; }
  }
  tb.flags[0] = 1u;
}

template <typename T>
__global__ void __launch_bounds__(256)
k_contacts(const T* __restrict__ in, i64 sx, i64 sy, i64 sz, int connectivity, ContactTable tb) {
  const i64 voxels = sx * sy * sz, sxy = sx * sy;
  const bool two_d = connectivity == 4 || connectivity == 8;
  const int ndir = connectivity / 2;
  const int lane = threadIdx.x & 31;
  const i64 stride = (i64)gridDim.x * blockDim.x;
  const i64 rounds = (voxels + stride - 1) / stride;
  for (i64 it = 0; it < rounds; it++) {            // every lane runs every round: the ballots below stay convergent
    const i64 i = it * stride + (i64)blockIdx.x * blockDim.x + threadIdx.x;
    const bool inside = i < voxels;
    T cur = (T)0;
    i64 off[13];
#pragma unroll
    for (int d = 0; d < 13; d++) off[d] = 0;
    if (inside) {
      cur = in[i];
      const i64 x = i % sx, t = i / sx;
      const i64 y = t % sy, z = t / sy;
      const i64 px = (x < sx - 1), mx = -(i64)(x > 0), py = sx * (y < sy - 1), my = -sx * (y > 0), mz = -sxy * (z > 0);
      off[0] = mx; off[1] = my;
      if (two_d) {
        off[2] = (connectivity > 4) * (mx + my);
        off[3] = (connectivity > 4) * (px + my);
      } else {
        off[2] = mz;
        if (connectivity > 6) {
          off[3] = (mx + my) * (mx && my); off[4] = (px + my) * (px && my);
          off[5] = (mx + mz) * (mx && mz); off[6] = (px + mz) * (px && mz);
          off[7] = (my + mz) * (my && mz); off[8] = (py + mz) * (py && mz);
        }
        if (connectivity > 18) {
          off[9] = (mx + my + mz) * (my && mz); off[10] = (px + my + mz) * (my && mz);
          off[11] = (mx + py + mz) * (py && mz); off[12] = (px + py + mz) * (py && mz);
        }
      }
    }
#pragma unroll
    for (int d = 0; d < 13; d++) {
      if (d < ndir) {
        bool act = false;
        unsigned long long key = 0;
        if (inside && cur != (T)0 && off[d] != 0) {
          const T q = in[i + off[d]];
          if (q != (T)0 && q != cur) {
            const unsigned long long a = (unsigned long long)(cur < q ? cur : q), b = (unsigned long long)(cur < q ? q : cur);
            if (b >> 32) tb.flags[1] = 1u;
            else { act = true; key = (a << 32) | b; }
          }
        }
        const u32 m = __ballot_sync(CC_FULL, act);
        if (act) {
          const u32 grp = __match_any_sync(m, key);
          if (lane == __ffs(grp) - 1) contact_add(tb, key, two_d ? (d < 2 ? d : 3) : (d < 3 ? d : 3), __popc(grp));
        }
      }
    }
  }
}

// dense list of the non-empty table entries
__global__ void __launch_bounds__(256)
k_contacts_compact(ContactTable tb, unsigned long long* __restrict__ out_keys, u32* __restrict__ out_vals,
                   unsigned long long* __restrict__ count, unsigned long long cap) {
  const u32 s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s > tb.mask) return;
  const unsigned long long k = tb.keys[s];
  if (!k) return;
  const unsigned long long pos = atomicAdd(count, 1ull);
  if (pos < cap) {
    out_keys[pos] = k;
    reinterpret_cast<uint4*>(out_vals)[pos] = reinterpret_cast<const uint4*>(tb.vals)[s];
  }
}
