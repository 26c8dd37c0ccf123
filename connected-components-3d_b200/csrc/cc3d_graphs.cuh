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

// ---- crackle v0 crack codes -> per-pixel 4-bit connectivity graph (the colouring step that
// follows is color_connectivity_graph, cc3d_graphs.hpp:583-1106). One THREAD per z slice: the chain parse is a stack
// machine over a difference-coded 2-bit symbol stream whose chains start where the previous one ended, i.e. it is
// sequential within a slice; the 512 slices of the benchmark volume run in parallel.
//   slice blob = u32 index_bytes | u16 n_rows, n_rows x { y_delta, n, n x x_delta } (chain start corners)
//              | symbols, 4 per byte, LSB first; dir[i] = (sym[0] + ... + sym[i]) mod 4, 0 up 1 right 2 down 3 left
//   a reversal ((d - last) mod 4 == 2) is a control symbol: the pending move is dropped; up / left terminate a branch
//   (pop the position), right / down open one (push the position); a chain ends when no branch is open.
//   a move from corner (x, y) cuts one pixel adjacency: clear the facing bit on both pixels (bit0 +x, bit1 -x,
//   bit2 +y, bit3 -y; the graph starts as 0b1111 everywhere).
// Cuts are fire-and-forget atomicAnd on the containing 32-bit word (no other thread touches the slice), so the only
// dependent chain is symbol fetch + control.
__global__ void __launch_bounds__(32)
k_crackle_cuts(const unsigned char* __restrict__ stream, const unsigned long long* __restrict__ slice_off, int sx, int sy, int sz,
               unsigned char* __restrict__ vcg, u32* __restrict__ stack, u32* __restrict__ err) {
  const int z = blockIdx.x * blockDim.x + threadIdx.x;
  if (z >= sz) return;
  const unsigned char* blob = stream + slice_off[z];
  const size_t blob_len = (size_t)(slice_off[z + 1] - slice_off[z]);
  if (blob_len < 6) { if (blob_len) atomicOr(err, 1u); return; }
  const u32 index_bytes = (u32)blob[0] | ((u32)blob[1] << 8) | ((u32)blob[2] << 16) | ((u32)blob[3] << 24);
  if (4 + (size_t)index_bytes > blob_len) { atomicOr(err, 1u); return; }
  const unsigned char* idx = blob + 4;
  auto rd16 = [&](u32 p) -> u32 { return (u32)idx[2 * p] | ((u32)idx[2 * p + 1] << 8); };
  const unsigned char* codes = blob + 4 + index_bytes;
  const size_t nsym = (blob_len - 4 - index_bytes) * 4;
  u32* stk = stack + 2 * (size_t)slice_off[z];          // room for one entry per two symbols of this slice
  const size_t stk_cap = 2 * blob_len;
  u32* plane = reinterpret_cast<u32*>(vcg + (size_t)z * sx * sy);   // sx * sy is a multiple of 4 (checked by the host)
  auto clear_bit = [&](int px, int py, int bit) {
    if (px < 0 || px >= sx || py < 0 || py >= sy) return;
    const size_t i = (size_t)py * sx + px;
    atomicAnd(plane + (i >> 2), ~((1u << bit) << (8 * (i & 3))));
  };
  size_t i = 0;
  u32 acc = 0;
  u32 p = 0;
  const u32 nidx = index_bytes / 2;
  if (nidx == 0) return;
  const u32 n_rows = rd16(p++);
  int y0 = 0;
  for (u32 r = 0; r < n_rows; r++) {
    if (p + 2 > nidx) { atomicOr(err, 2u); return; }
    y0 += (int)rd16(p); const u32 n = rd16(p + 1); p += 2;
    int x0 = 0;
    for (u32 k = 0; k < n; k++) {
      if (p >= nidx) { atomicOr(err, 2u); return; }
      x0 += (int)rd16(p++);
      int x = x0, y = y0;
      int branches = 1, last = -1;
      size_t sp = 0;
      while (branches > 0) {
        if (i >= nsym) { atomicOr(err, 4u); return; }
        acc = (acc + ((codes[i >> 2] >> (2 * (i & 3))) & 3u)) & 3u;
        i++;
        const int d = (int)acc;
        if (last >= 0 && ((d - last) & 3) == 2) {
          if (d == 0 || d == 3) {
            branches--;
            if (sp) { const u32 v = stk[--sp]; x = (int)(v & 0xFFFFu); y = (int)(v >> 16); }
          } else {
            branches++;
            if (sp >= stk_cap) { atomicOr(err, 8u); return; }
            stk[sp++] = (u32)x | ((u32)y << 16);
          }
          last = -1;
          continue;
        }
        if (last >= 0) {
          if (last == 0) { clear_bit(x - 1, y - 1, 0); clear_bit(x, y - 1, 1); y--; }
          else if (last == 2) { clear_bit(x - 1, y, 0); clear_bit(x, y, 1); y++; }
          else if (last == 3) { clear_bit(x - 1, y - 1, 2); clear_bit(x - 1, y, 3); x--; }
          else { clear_bit(x, y - 1, 2); clear_bit(x, y, 3); x++; }
        }
        last = d;
      }
    }
  }
}
