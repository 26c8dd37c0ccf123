// cc3d_faces.cuh — kernel A (the single pass over the input) and the continuous 2D-8 variant.
// See cc3d_common.cuh for the pipeline.
#pragma once
#include "cc3d_common.cuh"

#define CC_FACE_WARPS 8
#define CC_FACE_YCH 32   // rows of one plane a warp walks through
#define CC_FACE_UNR 4    // rows whose loads are issued back to back

template <typename T> __device__ __forceinline__ T shfl_up1(T v) { return __shfl_up_sync(CC_FULL, v, 1); }
template <> __device__ __forceinline__ uint8_t shfl_up1(uint8_t v) { return (uint8_t)__shfl_up_sync(CC_FULL, (unsigned)v, 1); }
template <> __device__ __forceinline__ uint16_t shfl_up1(uint16_t v) { return (uint16_t)__shfl_up_sync(CC_FULL, (unsigned)v, 1); }
template <> __device__ __forceinline__ uint64_t shfl_up1(uint64_t v) { return (uint64_t)__shfl_up_sync(CC_FULL, (unsigned long long)v, 1); }

// ---------------------------------------------------------------------------------------------
// Kernel A. One warp owns one bitmap-word column (32 voxels in x) of one z-plane and walks CC_FACE_YCH
// rows down y, one voxel per lane. The row above stays in registers; the loads of CC_FACE_UNR rows
// (this plane and plane z-1) are issued back to back before any of them is used. Per row: three
// compares and four ballots (F, X, Y, Z); lane 0 stores the four words with one 16-byte store.
// The row body is branch free (rows past the chunk end are predicated off). HASZ = 3D connectivity.
// ---------------------------------------------------------------------------------------------
template <typename T, int MODE, bool HASZ>
__global__ void __launch_bounds__(CC_FACE_WARPS * 32)
k_faces(const T* __restrict__ in, u32* __restrict__ M, Geom g, Edge<T, MODE> E, Counters* __restrict__ ctr,
        unsigned nych, unsigned ntasks) {
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const unsigned task = blockIdx.x * CC_FACE_WARPS + warp;

  u32 epl = 0, rfirst = 0xFFFFFFFFu, rlast = 0, anyfg = 0;
  if (task < ntasks) {
    const u32 W = (u32)g.W, sx = (u32)g.sx, sy = (u32)g.sy;
    const u32 w = task % W;
    const u32 t = task / W;
    const u32 ych = t % nych, z = t / nych;
    const u32 y0 = ych * CC_FACE_YCH;
    const u32 y1 = min(sy, y0 + CC_FACE_YCH);
    const u32 x = (w << 5) + lane;
    const bool inx = x < sx;
    const bool edge = lane == 0 && x > 0;
    const bool hasz = HASZ && z > 0;
    const u32 plane = sy * sx;
    const u32 row0 = z * sy + y0;
    const u32 off0 = row0 * sx + (inx ? x : 0);   // voxel (x, y0, z); voxels < 2^32
    u32 idx = row0 * W + w;
    uint4* __restrict__ MQ = reinterpret_cast<uint4*>(M);
    u32* __restrict__ RS = M + g.offRS;
    const T* __restrict__ p = in + off0;           // walks down the column
    T up = (T)0;
    if (y0 > 0 && inx) up = *(p - sx);
    u32 rowbits = 0;                               // bit k: row y0 + k has foreground

    // one row: faces of the 32 voxels c (left neighbour l, -y neighbour up, -z neighbour d)
    auto step = [&](const T c, const T pe, const T d, const u32 k, const bool store) {
      T l = shfl_up1(c);
      if (lane == 0) l = pe;
      const bool f = E.fg(c);
      const u32 F = __ballot_sync(CC_FULL, f);
      const u32 X = __ballot_sync(CC_FULL, E(c, l));
      const u32 Y = __ballot_sync(CC_FULL, E(c, up));
      const u32 Z = HASZ ? __ballot_sync(CC_FULL, E(c, d)) : 0u;
      const u32 ns = __popc(F & ~X);
      // cc3d.hpp:300-303: a provisional label per x-transition into a non-zero value
      if constexpr (MODE == MODE_EQ) epl += ns;
      else epl += __popc(__ballot_sync(CC_FULL, f && c != l));
      if (F) rowbits |= 1u << k;
      if (store && lane == 0) {
        MQ[idx] = make_uint4(F, X, Y, Z);
        RS[idx] = ns;
      }
      up = c;
      idx += W;
    };

    u32 yb = y0;
    if ((w << 5) + 32 <= sx && yb + CC_FACE_UNR <= y1) {   // warp-uniform: the whole word lies inside the row
      // full groups of CC_FACE_UNR rows, software pipelined: the unpredicated loads of the next group are
      // issued before the current group is evaluated
      T pc[CC_FACE_UNR], pe[CC_FACE_UNR], dc[CC_FACE_UNR];
      auto load_group = [&](T* a, T* b, T* c, const T* q) {
#pragma unroll
        for (int k = 0; k < CC_FACE_UNR; k++) {
          a[k] = q[k * sx];
          b[k] = (T)0;
          if (edge) b[k] = *(q + k * sx - 1);
          c[k] = (T)0;
          if (hasz) c[k] = *(q + k * sx - plane);
        }
      };
      load_group(pc, pe, dc, p);
      for (; yb + 2 * CC_FACE_UNR <= y1; yb += CC_FACE_UNR) {
        T npc[CC_FACE_UNR], npe[CC_FACE_UNR], ndc[CC_FACE_UNR];
        load_group(npc, npe, ndc, p + CC_FACE_UNR * sx);
#pragma unroll
        for (int k = 0; k < CC_FACE_UNR; k++) step(pc[k], pe[k], dc[k], yb - y0 + k, true);
#pragma unroll
        for (int k = 0; k < CC_FACE_UNR; k++) { pc[k] = npc[k]; pe[k] = npe[k]; dc[k] = ndc[k]; }
        p += CC_FACE_UNR * sx;
      }
#pragma unroll
      for (int k = 0; k < CC_FACE_UNR; k++) step(pc[k], pe[k], dc[k], yb - y0 + k, true);
      p += CC_FACE_UNR * sx;
      yb += CC_FACE_UNR;
    }
    // remaining rows, and every row of a partial last word (lanes past sx hold 0)
    for (; yb < y1; yb++) {
      T c = (T)0, e = (T)0, d = (T)0;
      if (inx) { c = *p; if (hasz) d = *(p - plane); }
      if (edge) e = *(p - 1);
      step(c, e, d, yb - y0, true);
      p += sx;
    }
    if (rowbits) {
      anyfg = 1;
      rfirst = row0 + __ffs(rowbits) - 1;
      rlast = row0 + 31 - __clz(rowbits);
    }
  }

  // block-level reduction of epl and the foreground row range
  __shared__ u32 s_epl, s_rmin, s_rmax, s_any;
  if (threadIdx.x == 0) { s_epl = 0; s_rmin = 0xFFFFFFFFu; s_rmax = 0; s_any = 0; }
  __syncthreads();
  if (lane == 0 && anyfg) {
    atomicAdd(&s_epl, epl);
    atomicMin(&s_rmin, rfirst);
    atomicMax(&s_rmax, rlast);
    s_any = 1;
  }
  __syncthreads();
  if (threadIdx.x == 0 && s_any) {
    if (s_epl) atomicAdd((unsigned long long*)&ctr->epl, (unsigned long long)s_epl);
    atomicMin((long long*)&ctr->first_row, (long long)s_rmin);
    atomicMax((long long*)&ctr->last_row, (long long)s_rmax);
  }
}

// ---------------------------------------------------------------------------------------------
// Continuous 2D 8-connected (cc3d_continuous.hpp:270-392): the reference's raster rule picks the
// backward edges of every pixel from its neighbourhood AND the global value range (gmin/gmax
// shortcut, :298-303, 341-349). We evaluate exactly that rule per pixel and ballot the result into
// the X / Y / A0 / C0 planes; kernel B then takes the diagonals from A0/C0 (MODE_MASK).
// One warp per bitmap word. `range` = {min, max} of the image (k_prepass).
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
k_c8_edges(const T* __restrict__ in, u32* __restrict__ M, Geom g, T delta, const T* __restrict__ range) {
  const int lane = threadIdx.x & 31;
  const i64 idx = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (idx >= g.nwords) return;
  const i64 sx = g.sx;
  const i64 y = idx / g.W, w = idx - y * g.W;
  const i64 x = (w << 5) + lane;
  const i64 i = y * sx + x;
  const T cur = x < sx ? in[i] : (T)0;
  bool eX = false, eY = false, eA = false, eC = false;
  if (cur != (T)0) {
    const T gmin = range[0], gmax = range[1];
    auto match = [&](T a, T b) -> bool {
      if constexpr (is_float_t<T>::value) return fabs(a - b) <= delta;
      else return (a > b ? (T)(a - b) : (T)(b - a)) <= delta;
    };
    bool shortcut = false;
    T vB = (T)0;
    if (y > 0) {
      vB = in[i - sx];
      if (cur == vB) shortcut = true;
      else if (vB != (T)0) {
        const T lo = cur < vB ? cur : vB, hi = cur > vB ? cur : vB;
        if ((lo - gmin <= delta) && (gmax - hi <= delta)) shortcut = true;
      }
    }
    if (shortcut) eY = true;
    else {
      if (y > 0 && vB != (T)0 && match(cur, vB)) eY = true;
      if (x > 0 && y > 0) { const T q = in[i - sx - 1]; if (q != (T)0 && match(cur, q)) eA = true; }
      if (x < sx - 1 && y > 0) { const T q = in[i - sx + 1]; if (q != (T)0 && match(cur, q)) eC = true; }
      if (x > 0) { const T q = in[i - 1]; if (q != (T)0 && match(cur, q)) eX = true; }
    }
  }
  const u32 F = __ballot_sync(CC_FULL, cur != (T)0);
  const u32 X = __ballot_sync(CC_FULL, eX);
  const u32 Y = __ballot_sync(CC_FULL, eY);
  const u32 A0 = __ballot_sync(CC_FULL, eA);
  const u32 C0 = __ballot_sync(CC_FULL, eC);
  if (lane == 0) {
    reinterpret_cast<uint4*>(M)[idx] = make_uint4(F, X, Y, 0u);
    M[g.offA0 + idx] = A0; M[g.offC0 + idx] = C0;
    M[g.offRS + idx] = __popc(F & ~X);
  }
}
