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
// compares and four ballots (F, X, Y, Z); lane 0 stores the words. HASZ = 3D connectivity.
// ---------------------------------------------------------------------------------------------
template <typename T, int MODE, bool HASZ>
__global__ void __launch_bounds__(CC_FACE_WARPS * 32)
k_faces(const T* __restrict__ in, u32* __restrict__ M, Geom g, Edge<T, MODE> E, Counters* __restrict__ ctr,
        unsigned nych, unsigned ntasks) {
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const unsigned task = blockIdx.x * CC_FACE_WARPS + warp;

  u32 epl = 0;
  long long rmin = INT64_MAX, rmax = -1;
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
    u32* __restrict__ mF = M + (size_t)PL_F * g.nwords;
    u32* __restrict__ mX = M + (size_t)PL_X * g.nwords;
    u32* __restrict__ mY = M + (size_t)PL_Y * g.nwords;
    u32* __restrict__ mZ = M + (size_t)PL_Z * g.nwords;
    u32* __restrict__ mR = M + (size_t)PL_RS * g.nwords;
    const bool hasz = HASZ && z > 0;
    const T* col = in + ((size_t)z * sy * sx + x);       // (x, 0, z)
    const T* colD = col - (size_t)sy * sx;               // (x, 0, z-1), only dereferenced when hasz
    T up = (T)0;
    if (y0 > 0 && inx) up = col[(size_t)(y0 - 1) * sx];
    u32 idx = (z * sy + y0) * W + w;

    for (u32 yb = y0; yb < y1; yb += CC_FACE_UNR) {
      T pc[CC_FACE_UNR], pe[CC_FACE_UNR], dc[CC_FACE_UNR];
#pragma unroll
      for (int k = 0; k < CC_FACE_UNR; k++) {
        const u32 y = yb + k;
        pc[k] = (T)0; pe[k] = (T)0; dc[k] = (T)0;
        if (y < y1) {
          const size_t o = (size_t)y * sx;
          if (inx) pc[k] = col[o];
          if (edge) pe[k] = col[o - 1];
          if (hasz && inx) dc[k] = colD[o];
        }
      }
#pragma unroll
      for (int k = 0; k < CC_FACE_UNR; k++) {
        const u32 y = yb + k;
        if (y < y1) {
          const T c = pc[k];
          T l = shfl_up1(c);
          if (lane == 0) l = pe[k];
          const bool f = E.fg(c);
          const u32 F = __ballot_sync(CC_FULL, f);
          u32 X = 0, Y = 0, Z = 0, S = 0;
          if (F) {
            X = __ballot_sync(CC_FULL, E(c, l));
            Y = __ballot_sync(CC_FULL, E(c, up));
            if (hasz) Z = __ballot_sync(CC_FULL, E(c, dc[k]));
            S = F & ~X;
            // cc3d.hpp:300-303: a provisional label per x-transition into a non-zero value
            if constexpr (MODE == MODE_EQ) epl += __popc(S);
            else epl += __popc(__ballot_sync(CC_FULL, f && c != l));
            const long long row = (long long)(z * sy + y);
            rmin = min(rmin, row);
            rmax = row;
          }
          if (lane == 0) {
            mF[idx] = F;
            mX[idx] = X;
            mY[idx] = Y;
            if (HASZ) mZ[idx] = Z;
            mR[idx] = __popc(S);
          }
          up = c;
          idx += W;
        }
      }
    }
  }

  // block-level reduction of epl and the foreground row range
  __shared__ u32 s_epl;
  __shared__ long long s_rmin, s_rmax;
  if (threadIdx.x == 0) { s_epl = 0; s_rmin = INT64_MAX; s_rmax = -1; }
  __syncthreads();
  if (lane == 0 && rmax >= 0) {
    atomicAdd(&s_epl, epl);
    atomicMin(&s_rmin, rmin);
    atomicMax(&s_rmax, rmax);
  }
  __syncthreads();
  if (threadIdx.x == 0 && s_rmax >= 0) {
    if (s_epl) atomicAdd((unsigned long long*)&ctr->epl, (unsigned long long)s_epl);
    atomicMin((long long*)&ctr->first_row, s_rmin);
    atomicMax((long long*)&ctr->last_row, s_rmax);
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
    const i64 nw = g.nwords;
    M[PL_F * nw + idx] = F; M[PL_X * nw + idx] = X; M[PL_Y * nw + idx] = Y;
    M[PL_A0 * nw + idx] = A0; M[PL_C0 * nw + idx] = C0;
    M[PL_RS * nw + idx] = __popc(F & ~X);
  }
}
