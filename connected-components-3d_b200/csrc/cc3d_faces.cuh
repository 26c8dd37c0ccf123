// cc3d_faces.cuh — kernel A (the single pass over the input) and the continuous 2D-8 variant.
// See cc3d_common.cuh for the pipeline.
#pragma once
#include "cc3d_common.cuh"

#ifndef CC_FACE_WARPS
#define CC_FACE_WARPS 8
#endif
#ifndef CC_FACE_YCH
#define CC_FACE_YCH 32   // rows of one plane a warp walks through
#endif
#ifndef CC_FACE_LANES
#define CC_FACE_LANES true
#endif
#define CC_FACE_NW 4     // bitmap words (of one row) a warp handles per row step

// Faces of one row of CC_FACE_NW words (c: voxels, l: -x neighbours, d: -z neighbours, up: -y neighbours);
// lane 0 stores the group's four {F,X,Y,Z} and run-start counts.
// LANES (NW == 4): lanes 0..3 each pick the {F,X,Y,Z} of one word (12 selects on the ALU pipe) and the group leaves the warp
// as ONE 64-byte store, one POPC and one 16-byte store of the run-start counts instead of four 16-byte stores, four
// POPCs and a store issued by lane 0: the memory-instruction queue (LDS + STG + POPC) is what throttles the unrolled
// kernel (ncu: mio_throttle 25 % of the stall samples).
template <typename T, int MODE, bool HASZ, int NW, bool LANES = false>
__device__ __forceinline__ void faces_eval_store(const Edge<T, MODE>& E, const T* c, const T* l, const T* d, const T* up,
                                                 int lane, uint4* __restrict__ mq, u32* __restrict__ rs, bool rs_vec, u32& epl) {
  u32 F[NW], X[NW], Y[NW], Z[NW];
#pragma unroll
  for (int k = 0; k < NW; k++) {
    const bool f = E.fg(c[k]);
    F[k] = __ballot_sync(CC_FULL, f);
    X[k] = __ballot_sync(CC_FULL, E.xedge(c[k], l[k]));
    Y[k] = __ballot_sync(CC_FULL, E.yedge(c[k], up[k]));
    Z[k] = HASZ ? __ballot_sync(CC_FULL, E.zedge(c[k], d[k])) : 0u;
    // cc3d.hpp:300-303: a provisional label per x-transition into a non-zero value
    if constexpr (MODE != MODE_EQ) epl += __popc(__ballot_sync(CC_FULL, f && c[k] != l[k]));
  }
  if constexpr (LANES && NW == 4) {
    const bool b0 = lane & 1, b1 = lane & 2;
    auto pick = [&](const u32* a) { const u32 lo = b0 ? a[1] : a[0], hi = b0 ? a[3] : a[2]; return b1 ? hi : lo; };
    const uint4 v = make_uint4(pick(F), pick(X), pick(Y), pick(Z));
    if (lane < 4) {
      mq[lane] = v;
      rs[lane] = __popc(v.x & ~v.y);
    }
    return;
  }
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < NW; k++) mq[k] = make_uint4(F[k], X[k], Y[k], Z[k]);
    if (rs_vec) {
#pragma unroll
      for (int k = 0; k < NW; k += 4)
        reinterpret_cast<uint4*>(rs)[k >> 2] = make_uint4(__popc(F[k] & ~X[k]), __popc(F[k + 1] & ~X[k + 1]), __popc(F[k + 2] & ~X[k + 2]), __popc(F[k + 3] & ~X[k + 3]));
    } else {
#pragma unroll
      for (int k = 0; k < NW; k++) rs[k] = __popc(F[k] & ~X[k]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Kernel A. One warp owns CC_FACE_NW consecutive bitmap words (128 voxels in x) of one z-plane and walks
// CC_FACE_YCH rows down y, one voxel per lane and word. Per word: the voxel c, its -x neighbour l (an
// L1-resident reload of the same lines shifted by one element), its -z neighbour d; the row above stays
// in registers. Three compares and four ballots (F, X, Y, Z) per word; lane 0 stores the four uint4
// {F,X,Y,Z} of the group (64 contiguous bytes) and the four run-start counts. The loads of row r+1 are
// issued before row r is evaluated (software pipeline), all addresses are one pointer + immediates.
// The foreground row range and (MODE_EQ) epl = number of runs come out of scan S; the other predicates
// count the reference's value transitions here (cc3d.hpp:300-303). HASZ = 3D connectivity.
// Groups that do not lie fully inside the row (sx not a multiple of 128) take the bounds-checked path.
// ---------------------------------------------------------------------------------------------
template <typename T, int MODE, bool HASZ>
__global__ void __launch_bounds__(CC_FACE_WARPS * 32)
k_faces(const T* __restrict__ in, u32* __restrict__ M, Geom g, Edge<T, MODE> E, Counters* __restrict__ ctr,
        unsigned nych, unsigned nwg, unsigned ntasks) {
  CC_PDL_WAIT();
  constexpr int NW = CC_FACE_NW;
  const int lane = threadIdx.x & 31;
  const unsigned task = blockIdx.x * CC_FACE_WARPS + (threadIdx.x >> 5);
  if (task >= ntasks) return;
  const u32 W = (u32)g.W, sx = (u32)g.sx, sy = (u32)g.sy;
  const u32 wg = task % nwg;
  const u32 t = task / nwg;
  const u32 ych = t % nych, z = t / nych;
  const u32 y0 = ych * CC_FACE_YCH;
  const u32 nrow = min(sy, y0 + CC_FACE_YCH) - y0;
  const u32 w0 = wg * NW;
  const bool hasz = HASZ && z > 0;
  const size_t plane = (size_t)sy * sx;
  const u32 row0 = z * sy + y0;
  uint4* __restrict__ mq = reinterpret_cast<uint4*>(M) + ((size_t)row0 * W + w0);
  u32* __restrict__ rs = M + g.offRS + ((size_t)row0 * W + w0);
  u32 epl = 0;

  if ((w0 + NW) * 32 <= sx) {
    // ---- fast path: NW full words ----
    const T* __restrict__ p = in + ((size_t)row0 * sx + (w0 << 5) + lane);
    const bool noleft = (w0 == 0 && lane == 0);      // voxel x == 0 has no -x neighbour
    const bool rs_vec = (W & 3) == 0;
    // three register sets rotate through the roles {row above, current row, row being loaded}
    T c0[NW], l0[NW], d0[NW], c1[NW], l1[NW], d1[NW], c2[NW], l2[NW], d2[NW];
    auto load_row = [&](T* cc, T* ll, T* dd, const T* q) {
#pragma unroll
      for (int k = 0; k < NW; k++) {
        cc[k] = q[32 * k];
        if (k == 0) { ll[0] = (T)0; if (!noleft) ll[0] = *(q - 1); }
        else ll[k] = q[32 * k - 1];
        dd[k] = (T)0;
        if (hasz) dd[k] = *(q + 32 * k - plane);
      }
    };
    auto eval_row = [&](const T* c, const T* l, const T* d, const T* up) {
      faces_eval_store<T, MODE, HASZ, NW>(E, c, l, d, up, lane, mq, rs, rs_vec, epl);
      mq += W; rs += W;
    };
#pragma unroll
    for (int k = 0; k < NW; k++) { c2[k] = (T)0; if (y0 > 0) c2[k] = *(p + 32 * k - sx); }
    load_row(c0, l0, d0, p);
    u32 r = 0;
    for (; r + 3 < nrow; r += 3) {
      const T* p1 = p + sx; const T* p2 = p1 + sx; p = p2 + sx;
      load_row(c1, l1, d1, p1); eval_row(c0, l0, d0, c2);
      load_row(c2, l2, d2, p2); eval_row(c1, l1, d1, c0);
      load_row(c0, l0, d0, p);  eval_row(c2, l2, d2, c1);
    }
    const u32 rem = nrow - r;   // 1..3 rows left, the first of them is loaded in set 0
    if (rem >= 2) load_row(c1, l1, d1, p + sx);
    eval_row(c0, l0, d0, c2);
    if (rem >= 2) {
      if (rem == 3) load_row(c2, l2, d2, p + 2 * (size_t)sx);
      eval_row(c1, l1, d1, c0);
      if (rem == 3) eval_row(c2, l2, d2, c1);
    }
  } else {
    // ---- bounds-checked path: words of a group that crosses the end of the row ----
    const u32 nw = min((u32)NW, W - w0);
    for (u32 k = 0; k < nw; k++) {
      const u32 x = ((w0 + k) << 5) + lane;
      const bool inx = x < sx;
      const T* __restrict__ p = in + ((size_t)row0 * sx + (inx ? x : 0));
      T up = (T)0;
      if (y0 > 0 && inx) up = *(p - sx);
      for (u32 r = 0; r < nrow; r++) {
        T c = (T)0, l = (T)0, d = (T)0;
        if (inx) { c = *p; if (x > 0) l = *(p - 1); if (hasz) d = *(p - plane); }
        const bool f = E.fg(c);
        const u32 F = __ballot_sync(CC_FULL, f);
        const u32 X = __ballot_sync(CC_FULL, E.xedge(c, l));
        const u32 Y = __ballot_sync(CC_FULL, E.yedge(c, up));
        const u32 Z = HASZ ? __ballot_sync(CC_FULL, E.zedge(c, d)) : 0u;
        if constexpr (MODE != MODE_EQ) epl += __popc(__ballot_sync(CC_FULL, f && c != l));
        if (lane == 0) {
          mq[(size_t)r * W + k] = make_uint4(F, X, Y, Z);
          rs[(size_t)r * W + k] = __popc(F & ~X);
        }
        up = c;
        p += sx;
      }
    }
  }
  if constexpr (MODE != MODE_EQ) {
    if (lane == 0 && epl) atomicAdd((unsigned long long*)&ctr->epl, (unsigned long long)epl);
  }
}

// ---------------------------------------------------------------------------------------------
// Kernel A, staged variant (the one that runs when rows are 16-byte aligned and sx is a multiple of 128).
// Same task decomposition and arithmetic as k_faces, but the rows travel global -> shared memory with
// cp.async (16 bytes per lane, no registers held while in flight) through a per-warp ring of
// CC_FACE_STAGES row slots, so the bytes in flight per SM no longer depend on the register budget.
// Slot layout: [16 B: the 16 bytes left of the group (-x neighbour of its first voxel)][row z][row z-1].
// ---------------------------------------------------------------------------------------------
#ifndef CC_FACE_STAGES
#define CC_FACE_STAGES 4
#endif
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <typename T, int NW> constexpr size_t faces_async_smem() {
  return (size_t)CC_FACE_WARPS * CC_FACE_STAGES * (16 + 2 * NW * 32 * sizeof(T));
}

#ifndef CC_FACE_MINB
#define CC_FACE_MINB 0
#endif
template <typename T, int MODE, bool HASZ, int NW>
__global__ void __launch_bounds__(CC_FACE_WARPS * 32, CC_FACE_MINB)
k_faces_async(const T* __restrict__ in, u32* __restrict__ M, Geom g, Edge<T, MODE> E, Counters* __restrict__ ctr,
              unsigned nych, unsigned nwg, unsigned ntasks) {
  CC_PDL_WAIT();
  constexpr int RB = NW * 32 * (int)sizeof(T);       // bytes of one row of the group
  constexpr int SLOT = 16 + 2 * RB;
  extern __shared__ __align__(16) unsigned char face_smem[];
  const int lane = threadIdx.x & 31;
  const unsigned task = blockIdx.x * CC_FACE_WARPS + (threadIdx.x >> 5);
  if (task >= ntasks) return;
  unsigned char* ring = face_smem + (size_t)(threadIdx.x >> 5) * CC_FACE_STAGES * SLOT;
  const u32 W = (u32)g.W, sx = (u32)g.sx, sy = (u32)g.sy;
  const u32 wg = task % nwg;
  const u32 t = task / nwg;
  const u32 ych = t % nych, z = t / nych;
  const u32 y0 = ych * CC_FACE_YCH;
  const u32 nrow = min(sy, y0 + CC_FACE_YCH) - y0;
  const u32 w0 = wg * NW;
  const bool hasz = HASZ && z > 0;
  const bool hasleft = w0 > 0;
  const size_t rowbytes = (size_t)sx * sizeof(T);
  const size_t planebytes = (size_t)sy * rowbytes;
  const u32 row0 = z * sy + y0;
  uint4* __restrict__ mq = reinterpret_cast<uint4*>(M) + ((size_t)row0 * W + w0);
  u32* __restrict__ rs = M + g.offRS + ((size_t)row0 * W + w0);
  const bool rs_vec = (W & 3) == 0;
  u32 epl = 0;
  const unsigned char* gsrc = reinterpret_cast<const unsigned char*>(in + ((size_t)row0 * sx + (w0 << 5)));   // next row to issue
  auto issue = [&](unsigned char* slot) {
#pragma unroll
    for (int off = 0; off < RB; off += 512) {
      const int o = off + lane * 16;
      if (RB >= 512 || o < RB) {
        cp_async16(slot + 16 + o, gsrc + o);
        if (hasz) cp_async16(slot + 16 + RB + o, gsrc - planebytes + o);
      }
    }
    if (lane == 0 && hasleft) cp_async16(slot, gsrc - 16);
    gsrc += rowbytes;
  };
  T up[NW];
  {
    const T* p = in + ((size_t)row0 * sx + (w0 << 5) + lane);
#pragma unroll
    for (int k = 0; k < NW; k++) { up[k] = (T)0; if (y0 > 0) up[k] = *(p + 32 * k - sx); }
  }
  // slots keep zeros where nothing is ever copied: the left pad of the first group, row z-1 of plane 0
  if (!hasleft || (HASZ && !hasz)) {
#pragma unroll
    for (int s = 0; s < CC_FACE_STAGES; s++) {
      if (!hasleft && lane == 0) *reinterpret_cast<uint4*>(ring + s * SLOT) = make_uint4(0u, 0u, 0u, 0u);
      if (HASZ && !hasz)
        for (int o = lane * 16; o < RB; o += 512) *reinterpret_cast<uint4*>(ring + s * SLOT + 16 + RB + o) = make_uint4(0u, 0u, 0u, 0u);
    }
    __syncwarp();
  }
  auto eval_slot = [&](const unsigned char* slot) {
    const T* sc = reinterpret_cast<const T*>(slot + 16);
    const T* sd = reinterpret_cast<const T*>(slot + 16 + RB);
    T c[NW], l[NW], d[NW];
#pragma unroll
    for (int k = 0; k < NW; k++) {
      c[k] = sc[32 * k + lane];
      l[k] = sc[32 * k + lane - 1];
      d[k] = HASZ ? sd[32 * k + lane] : (T)0;
    }
    faces_eval_store<T, MODE, HASZ, NW>(E, c, l, d, up, lane, mq, rs, rs_vec, epl);
    mq += W; rs += W;
#pragma unroll
    for (int k = 0; k < NW; k++) up[k] = c[k];
    __syncwarp();   // every lane has read the slot before the copy of a later row is issued into it
  };
  u32 issued = 0;
#pragma unroll
  for (int s = 0; s < CC_FACE_STAGES - 1; s++) {
    if (issued < nrow) { issue(ring + s * SLOT); issued++; }
    cp_async_commit();
  }
  u32 r = 0;
  // full trips: every row of the trip exists and so does the row issued CC_FACE_STAGES-1 ahead of it
  for (; r + 2 * CC_FACE_STAGES - 1 <= nrow; r += CC_FACE_STAGES) {
#pragma unroll
    for (int s = 0; s < CC_FACE_STAGES; s++) {
      issue(ring + ((s + CC_FACE_STAGES - 1) % CC_FACE_STAGES) * SLOT);
      cp_async_commit();
      cp_async_wait<CC_FACE_STAGES - 1>();
      __syncwarp();
      eval_slot(ring + s * SLOT);
    }
  }
  issued = r + CC_FACE_STAGES - 1;
  for (; r < nrow; r += CC_FACE_STAGES) {
#pragma unroll
    for (int s = 0; s < CC_FACE_STAGES; s++) {
      if (r + s < nrow) {
        if (issued < nrow) { issue(ring + ((s + CC_FACE_STAGES - 1) % CC_FACE_STAGES) * SLOT); issued++; }
        cp_async_commit();
        cp_async_wait<CC_FACE_STAGES - 1>();
        __syncwarp();
        eval_slot(ring + s * SLOT);
      }
    }
  }
  if constexpr (MODE != MODE_EQ) {
    if (lane == 0 && epl) atomicAdd((unsigned long long*)&ctr->epl, (unsigned long long)epl);
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

// ---------------------------------------------------------------------------------------------
// Kernel A, TMA variant (sm_100a data movement: cp.async.bulk.tensor + mbarrier; inputs of <= 4-byte elements whose
// rows are 16-byte multiples). A warp owns CC_FACE_NW bitmap words in x and TR rows in y and WALKS DOWN z through a
// chunk of planes. Per plane ONE tensor-map bulk copy (issued by lane 0, completion on a per-buffer mbarrier) brings
// the box [x0 - pad, x0 + 128) x [y0 - 1, y0 + TR) of plane z into a ring of three shared-memory buffers:
//   * the -x neighbour is the same row one element to the left (the box starts `pad` elements early),
//   * the -y neighbour of the first row is the box's halo row, the others are the previous row (kept in registers),
//   * the -z neighbour is the SAME position in the previous buffer of the ring - every voxel crosses L2 -> SM once
//     (plus the one-row halo) instead of twice,
//   * everything outside the volume (x = -1, y = -1, z = -1, the ragged end of a row or of a y block) is zero-filled by
//     the TMA unit, i.e. background: no edge cases in the kernel, any sx / sy / sz.
// No per-lane copy instructions remain on the MIO queue (the cp.async variant issued 2-3 LDGSTS per lane and row).
// ---------------------------------------------------------------------------------------------
#include <cuda.h>   // CUtensorMap (types only; the encoder is fetched through cudaGetDriverEntryPoint)

template <typename T> struct FaceTma {
  static constexpr int NW = CC_FACE_NW;                          // words per row of the box
  static constexpr int PAD = 16 / (int)sizeof(T);                // elements left of the group (16 bytes)
  static constexpr int BOXX = NW * 32 + PAD;                     // box extent in x (elements)
#ifndef CC_TMA_TR4
#define CC_TMA_TR4 4
#endif
#ifndef CC_TMA_NBUF
#define CC_TMA_NBUF 3
#endif
  static constexpr int TR = sizeof(T) == 1 ? 4 * CC_TMA_TR4 : (sizeof(T) == 2 ? 2 * CC_TMA_TR4 : CC_TMA_TR4);   // rows per warp and plane (~2.5 KB boxes)
  static constexpr int PITCH = BOXX * (int)sizeof(T);            // bytes per box row
  static constexpr int BOXB = PITCH * (TR + 1);                  // bytes per box
  static constexpr int BUFB = (BOXB + 127) & ~127;               // 128-byte aligned ring slots
  static constexpr int NBUF = CC_TMA_NBUF;     // ring slots: planes z - 1, z and NBUF - 2 planes in flight
  static constexpr size_t smem() { return (size_t)CC_FACE_WARPS * (NBUF * BUFB) + CC_FACE_WARPS * NBUF * 8 + 128; }
};

__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile(
    "{\n\t.reg .pred p;\n\t"
    "WAIT_%=:\n\t"
    "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
    "@p bra DONE_%=;\n\t"
    "bra WAIT_%=;\n\t"
    "DONE_%=:\n\t}" ::"r"(a), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, unsigned long long* bar, int c0, int c1, int c2) {
  asm volatile(
    "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
    ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(reinterpret_cast<unsigned long long>(map)),
      "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

template <typename T, int MODE, bool HASZ>
__global__ void __launch_bounds__(CC_FACE_WARPS * 32)
k_faces_tma(const __grid_constant__ CUtensorMap tmap, u32* __restrict__ M, Geom g, Edge<T, MODE> E, Counters* __restrict__ ctr,
            unsigned nyb, unsigned nwg, unsigned zchunk, unsigned ntasks) {
  typedef FaceTma<T> F;
  constexpr int NW = F::NW, TR = F::TR, PAD = F::PAD;
  extern __shared__ __align__(128) unsigned char tma_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // 128-byte aligned base (dynamic shared memory is only guaranteed 16-byte alignment)
  unsigned char* base = tma_smem + ((128u - ((unsigned)__cvta_generic_to_shared(tma_smem) & 127u)) & 127u);
  unsigned char* ring = base + (size_t)warp * (F::NBUF * F::BUFB);
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(base + (size_t)CC_FACE_WARPS * (F::NBUF * F::BUFB)) + warp * F::NBUF;
  if (lane == 0) {
#pragma unroll
    for (int b = 0; b < F::NBUF; b++) mbar_init(bars + b, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncwarp();
  CC_PDL_WAIT();
  const unsigned task = blockIdx.x * CC_FACE_WARPS + warp;
  if (task >= ntasks) return;
  const u32 W = (u32)g.W, sy = (u32)g.sy, sz = (u32)g.sz;
  const u32 wg = task % nwg;
  const u32 t = task / nwg;
  const u32 yb = t % nyb, zc = t / nyb;
  const u32 w0 = wg * NW, y0 = yb * TR;
  const u32 z0 = zc * zchunk, z1 = min(sz, z0 + zchunk);
  const int cx = (int)(w0 << 5) - PAD, cy = (int)y0 - 1;
  const u32 nrow = min(sy, y0 + TR) - y0;
  const u32 nw = min((u32)NW, W - w0);
  const bool rs_vec = (W & 3) == 0 && nw == NW;
  u32 epl = 0;

  // it-th plane of the walk (it = 0: plane z0 - 1, only fetched for 3D connectivities) -> ring slot it % NBUF
  auto issue = [&](u32 it) {
    if (lane == 0) {
      const u32 b = it % F::NBUF;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the slot was last READ through the generic proxy
      mbar_expect_tx(bars + b, (unsigned)F::BOXB);
      tma_load_3d(ring + b * F::BUFB, &tmap, bars + b, cx, cy, (int)(z0 + it) - 1);
    }
  };
  constexpr u32 PD = F::NBUF - 2;            // planes in flight ahead of the one being evaluated
  if (HASZ) issue(0);
#pragma unroll
  for (u32 k = 0; k < PD; k++) if (z0 + k < z1) issue(1 + k);
  if (HASZ) mbar_wait(bars + 0, 0);
  for (u32 z = z0; z < z1; z++) {
    const u32 it = z - z0 + 1;
    if (z + PD < z1) issue(it + PD);        // slot of plane z - 2: every lane finished reading it before the __syncwarp below
    mbar_wait(bars + it % F::NBUF, (it / F::NBUF) & 1u);
    const unsigned char* cur = ring + (it % F::NBUF) * F::BUFB;
    const unsigned char* prv = ring + ((it - 1) % F::NBUF) * F::BUFB;
    const u32 row0 = z * sy + y0;
    uint4* __restrict__ mq = reinterpret_cast<uint4*>(M) + ((size_t)row0 * W + w0);
    u32* __restrict__ rs = M + g.offRS + ((size_t)row0 * W + w0);
    T up[NW];
    {
      const T* h = reinterpret_cast<const T*>(cur) + PAD + lane;     // halo row y0 - 1
#pragma unroll
      for (int k = 0; k < NW; k++) up[k] = h[32 * k];
    }
    for (u32 r = 0; r < nrow; r++) {
      const T* sc = reinterpret_cast<const T*>(cur + (size_t)(r + 1) * F::PITCH) + PAD + lane;
      const T* sd = reinterpret_cast<const T*>(prv + (size_t)(r + 1) * F::PITCH) + PAD + lane;
      T c[NW], l[NW], d[NW];
#pragma unroll
      for (int k = 0; k < NW; k++) {
        c[k] = sc[32 * k];
        l[k] = sc[32 * k - 1];
        d[k] = HASZ ? sd[32 * k] : (T)0;
      }
      if (nw == NW) {
        faces_eval_store<T, MODE, HASZ, NW>(E, c, l, d, up, lane, mq, rs, rs_vec, epl);
      } else {
        // ragged last group of a row: the words beyond W are zero-filled background, only the real ones are stored
        u32 Fw[NW], Xw[NW], Yw[NW], Zw[NW];
#pragma unroll
        for (int k = 0; k < NW; k++) {
          const bool f = E.fg(c[k]);
          Fw[k] = __ballot_sync(CC_FULL, f);
          Xw[k] = __ballot_sync(CC_FULL, E.xedge(c[k], l[k]));
          Yw[k] = __ballot_sync(CC_FULL, E.yedge(c[k], up[k]));
          Zw[k] = HASZ ? __ballot_sync(CC_FULL, E.zedge(c[k], d[k])) : 0u;
          if constexpr (MODE != MODE_EQ) epl += __popc(__ballot_sync(CC_FULL, f && c[k] != l[k]));
        }
        if (lane == 0) {
#pragma unroll
          for (int k = 0; k < NW; k++)
            if ((u32)k < nw) { mq[k] = make_uint4(Fw[k], Xw[k], Yw[k], Zw[k]); rs[k] = __popc(Fw[k] & ~Xw[k]); }
        }
      }
      mq += W; rs += W;
#pragma unroll
      for (int k = 0; k < NW; k++) up[k] = c[k];
    }
    __syncwarp();
  }
  if constexpr (MODE != MODE_EQ) {
    if (lane == 0 && epl) atomicAdd((unsigned long long*)&ctr->epl, (unsigned long long)epl);
  }
}


// ---------------------------------------------------------------------------------------------
// Kernel A, TMA variant 2: as k_faces_tma, but the plane that has just been evaluated stays in REGISTERS as the -z
// operand of the next one (TR2 x NW values per lane), so the ring only holds planes that are in flight (NBUF2 - 1 of
// them ahead of the evaluation) and a row costs two shared-memory loads per word (voxel, -x neighbour) instead of
// three. The plane below the first one of a z chunk is fetched like any other and only fills the registers.
// ---------------------------------------------------------------------------------------------
#ifndef CC_TMA2_NBUF
#define CC_TMA2_NBUF 2     // measured on B200 (profiles/r02_faces_ab.md): 2 slots 0.170-0.174 ms, 3 slots 0.174-0.176, 4 slots 0.196 (512^3 u32)
#endif
#ifndef CC_TMA2_TR
#define CC_TMA2_TR 8      // measured on B200 (profiles/r02b_faces_variants.md): 4 rows / 8 warps 0.148 ms, 8 rows / 8 warps 0.137, 8 rows / 4 warps 0.128, 16 rows / 2 warps 0.142
#endif
#ifndef CC_TMA2_MINB
#define CC_TMA2_MINB 1
#endif
#ifndef CC_TMA2_WARPS
#define CC_TMA2_WARPS 4
#endif
template <typename T> struct FaceTma2 {
  static constexpr int WARPS = CC_TMA2_WARPS;
  static constexpr int NW = CC_FACE_NW;
  static constexpr int PAD = 16 / (int)sizeof(T);
  static constexpr int BOXX = NW * 32 + PAD;
  static constexpr int TR = CC_TMA2_TR;
  static constexpr int PITCH = BOXX * (int)sizeof(T);
  static constexpr int BOXB = PITCH * (TR + 1);
  static constexpr int BUFB = (BOXB + 127) & ~127;
  static constexpr int NBUF = CC_TMA2_NBUF;
  static constexpr size_t smem() { return (size_t)WARPS * (NBUF * BUFB) + WARPS * NBUF * 8 + 128; }
};

template <typename T, int MODE, bool HASZ>
__global__ void __launch_bounds__(CC_TMA2_WARPS * 32, CC_TMA2_MINB)
k_faces_tma2(const __grid_constant__ CUtensorMap tmap, u32* __restrict__ M, Geom g, Edge<T, MODE> E, Counters* __restrict__ ctr,
             unsigned nyb, unsigned nwg, unsigned zchunk, unsigned ntasks) {
  typedef FaceTma2<T> F;
  constexpr int NW = F::NW, TR = F::TR, PAD = F::PAD;
  extern __shared__ __align__(128) unsigned char tma_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned char* base = tma_smem + ((128u - ((unsigned)__cvta_generic_to_shared(tma_smem) & 127u)) & 127u);
  unsigned char* ring = base + (size_t)warp * (F::NBUF * F::BUFB);
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(base + (size_t)F::WARPS * (F::NBUF * F::BUFB)) + warp * F::NBUF;
  if (lane == 0) {
#pragma unroll
    for (int b = 0; b < F::NBUF; b++) mbar_init(bars + b, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncwarp();
  CC_PDL_WAIT();
  const unsigned task = blockIdx.x * F::WARPS + warp;
  if (task >= ntasks) return;
  const u32 W = (u32)g.W, sy = (u32)g.sy, sz = (u32)g.sz;
  const u32 wg = task % nwg;
  const u32 t = task / nwg;
  const u32 yb = t % nyb, zc = t / nyb;
  const u32 w0 = wg * NW, y0 = yb * TR;
  const u32 z0 = zc * zchunk, z1 = min(sz, z0 + zchunk);
  const int cx = (int)(w0 << 5) - PAD, cy = (int)y0 - 1;
  const u32 nrow = min(sy, y0 + TR) - y0;
  const u32 nw = min((u32)NW, W - w0);
  const bool rs_vec = (W & 3) == 0 && nw == NW;
  u32 epl = 0;
  // walk: step 0 = plane z0 - 1 (3D connectivities only; it just fills the registers), step s = plane z0 - 1 + s
  const u32 first = HASZ ? 0u : 1u;
  const u32 nsteps = z1 - z0 + 1;
  auto issue = [&](u32 s_) {
    if (lane == 0) {
      const u32 b = (s_ - first) % F::NBUF;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_expect_tx(bars + b, (unsigned)F::BOXB);
      tma_load_3d(ring + b * F::BUFB, &tmap, bars + b, cx, cy, (int)(z0 + s_) - 1);
    }
  };
  constexpr u32 PD = F::NBUF;          // a slot is refilled as soon as its plane has been evaluated: NBUF - 1 planes in flight
#pragma unroll
  for (u32 k = 0; k < PD; k++) if (first + k < nsteps) issue(first + k);
#ifndef CC_TMA2_NO_FAST
  if (nrow == (u32)TR && rs_vec) {
    // ---- full blocks (every block of a volume whose sy is a multiple of 4 and whose rows are multiples of 128 voxels):
    //      rows unrolled without conditions, the planes z - 1 / z ping-pong between two register sets (no register
    //      moves), ONE 32-bit word index per row instead of two 64-bit pointers. ncu of the conditional loop below
    //      (profiles/r02_ncu_full_summary.md): 110 instructions per row of four words, of which 19 were pointer
    //      arithmetic and 16 register moves. ----
    T A[TR][NW], B[TR][NW];
#pragma unroll
    for (int r = 0; r < TR; r++)
#pragma unroll
      for (int k = 0; k < NW; k++) A[r][k] = (T)0;
    uint4* __restrict__ M4 = reinterpret_cast<uint4*>(M);
    u32* __restrict__ RSb = M + g.offRS;
    const u32 wplane = sy * W;
    u32 widx = (z0 * sy + y0) * W + w0;
    auto eval_plane = [&](const T (&prev)[TR][NW], T (&cur)[TR][NW], const unsigned char* buf, u32 wrow) {
      T up[NW];
      {
        const T* h = reinterpret_cast<const T*>(buf) + PAD + lane;     // halo row y0 - 1
#pragma unroll
        for (int k = 0; k < NW; k++) up[k] = h[32 * k];
      }
#pragma unroll
      for (int r = 0; r < TR; r++) {
        const T* sc = reinterpret_cast<const T*>(buf + (size_t)(r + 1) * F::PITCH) + PAD + lane;
        T l[NW];
#pragma unroll
        for (int k = 0; k < NW; k++) { cur[r][k] = sc[32 * k]; l[k] = sc[32 * k - 1]; }
        faces_eval_store<T, MODE, HASZ, NW, CC_FACE_LANES>(E, cur[r], l, prev[r], r == 0 ? up : cur[r > 0 ? r - 1 : 0], lane, M4 + wrow, RSb + wrow, true, epl);
        wrow += W;
      }
    };
    u32 st = first;
    if (HASZ) {
      // plane z0 - 1: only the -z operands of the first plane
      mbar_wait(bars, 0u);
#pragma unroll
      for (int r = 0; r < TR; r++) {
        const T* sc = reinterpret_cast<const T*>(ring + (size_t)(r + 1) * F::PITCH) + PAD + lane;
#pragma unroll
        for (int k = 0; k < NW; k++) A[r][k] = sc[32 * k];
      }
      __syncwarp();
      if (st + PD < nsteps) issue(st + PD);
      st++;
    }
    while (st < nsteps) {
      {
        const u32 q = st - first;
        mbar_wait(bars + q % F::NBUF, (q / F::NBUF) & 1u);
        eval_plane(A, B, ring + (q % F::NBUF) * F::BUFB, widx);
        widx += wplane;
        __syncwarp();
        if (st + PD < nsteps) issue(st + PD);
        st++;
      }
      if (st >= nsteps) break;
      {
        const u32 q = st - first;
        mbar_wait(bars + q % F::NBUF, (q / F::NBUF) & 1u);
        eval_plane(B, A, ring + (q % F::NBUF) * F::BUFB, widx);
        widx += wplane;
        __syncwarp();
        if (st + PD < nsteps) issue(st + PD);
        st++;
      }
    }
    if constexpr (MODE != MODE_EQ) {
      if (lane == 0 && epl) atomicAdd((unsigned long long*)&ctr->epl, (unsigned long long)epl);
    }
    return;
  }
#endif
  T dreg[TR][NW];
#pragma unroll
  for (int r = 0; r < TR; r++)
#pragma unroll
    for (int k = 0; k < NW; k++) dreg[r][k] = (T)0;
  for (u32 st = first; st < nsteps; st++) {
    const u32 q = st - first;
    mbar_wait(bars + q % F::NBUF, (q / F::NBUF) & 1u);
    const unsigned char* cur = ring + (q % F::NBUF) * F::BUFB;
    if (st == 0) {
      // plane z0 - 1: only the -z operands of the first plane
#pragma unroll
      for (int r = 0; r < TR; r++) {
        const T* sc = reinterpret_cast<const T*>(cur + (size_t)(r + 1) * F::PITCH) + PAD + lane;
#pragma unroll
        for (int k = 0; k < NW; k++) dreg[r][k] = sc[32 * k];
      }
    } else {
      const u32 z = z0 + st - 1;
      const u32 row0 = z * sy + y0;
      uint4* __restrict__ mq = reinterpret_cast<uint4*>(M) + ((size_t)row0 * W + w0);
      u32* __restrict__ rs = M + g.offRS + ((size_t)row0 * W + w0);
      T up[NW];
      {
        const T* h = reinterpret_cast<const T*>(cur) + PAD + lane;     // halo row y0 - 1
#pragma unroll
        for (int k = 0; k < NW; k++) up[k] = h[32 * k];
      }
#pragma unroll
      for (int r = 0; r < TR; r++) {
        if ((u32)r < nrow) {
          const T* sc = reinterpret_cast<const T*>(cur + (size_t)(r + 1) * F::PITCH) + PAD + lane;
          T c[NW], l[NW];
#pragma unroll
          for (int k = 0; k < NW; k++) { c[k] = sc[32 * k]; l[k] = sc[32 * k - 1]; }
          if (nw == NW) {
            faces_eval_store<T, MODE, HASZ, NW>(E, c, l, dreg[r], up, lane, mq, rs, rs_vec, epl);
          } else {
            u32 Fw[NW], Xw[NW], Yw[NW], Zw[NW];
#pragma unroll
            for (int k = 0; k < NW; k++) {
              const bool f = E.fg(c[k]);
              Fw[k] = __ballot_sync(CC_FULL, f);
              Xw[k] = __ballot_sync(CC_FULL, E.xedge(c[k], l[k]));
              Yw[k] = __ballot_sync(CC_FULL, E.yedge(c[k], up[k]));
              Zw[k] = HASZ ? __ballot_sync(CC_FULL, E.zedge(c[k], dreg[r][k])) : 0u;
              if constexpr (MODE != MODE_EQ) epl += __popc(__ballot_sync(CC_FULL, f && c[k] != l[k]));
            }
            if (lane == 0) {
#pragma unroll
              for (int k = 0; k < NW; k++)
                if ((u32)k < nw) { mq[k] = make_uint4(Fw[k], Xw[k], Yw[k], Zw[k]); rs[k] = __popc(Fw[k] & ~Xw[k]); }
            }
          }
          mq += W; rs += W;
#pragma unroll
          for (int k = 0; k < NW; k++) { up[k] = c[k]; dreg[r][k] = c[k]; }
        }
      }
    }
    __syncwarp();                                   // every lane has read the slot: it can be refilled
    if (st + PD < nsteps) issue(st + PD);
  }
  if constexpr (MODE != MODE_EQ) {
    if (lane == 0 && epl) atomicAdd((unsigned long long*)&ctr->epl, (unsigned long long)epl);
  }
}

// ---------------------------------------------------------------------------------------------
// Kernel A for BINARY images of 1-byte elements (MODE_NONZERO on uint8 / bool, round 2d). With the non-zero predicate
// every face link is a conjunction of two foreground bits, so the three compares + four votes per 32 voxels of the
// generic kernel (one voxel per lane) are not needed at all:
//   k_fg_bitmap_u8   one THREAD per bitmap word: 32 bytes (two 16-byte loads) -> 32 foreground bits with byte-parallel
//                    arithmetic on 4 voxels at a time (non-zero test: ((v & 0x7f7f7f7f) + 0x7f7f7f7f | v) & 0x80808080,
//                    bit gather: one multiply), and the reference's transition count epl on the RAW values
//                    (cc3d.hpp:300-303: c != 0 && c != left neighbour) with the same trick on v ^ (v shifted by a voxel)
//   k_faces_from_fg  X = F & (F << 1 | carry), Y = F & F(y-1), Z = F & F(z-1) on whole words, run-start counts
// ~4 warp instructions per bitmap word instead of 18. The compact F bitmap (1/8 byte per voxel) lives in the not yet
// used forest array L.
// ---------------------------------------------------------------------------------------------

static __global__ void __launch_bounds__(256)
k_fg_bitmap_u8(const uint8_t* __restrict__ in, u32* __restrict__ Fb, Geom g, Counters* __restrict__ ctr, int vec_ok) {
  CC_PDL_WAIT();
  const u32 W = (u32)g.W, sx = (u32)g.sx, nwords = (u32)g.nwords;
  u32 epl = 0;
  for (u32 j = blockIdx.x * blockDim.x + threadIdx.x; j < nwords; j += gridDim.x * blockDim.x) {
    const u32 row = j / W, w = j - row * W;
    const uint8_t* __restrict__ p = in + ((size_t)row * sx + (w << 5));
    u32 F = 0;
    if (vec_ok && (w << 5) + 32u <= sx) {
      const uint4 a = __ldg(reinterpret_cast<const uint4*>(p)), b = __ldg(reinterpret_cast<const uint4*>(p) + 1);
      u32 pv = w > 0 ? __ldg(reinterpret_cast<const u32*>(p) - 1) : 0u;    // voxel x == 0 has no -x neighbour
      const u32 v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
      for (int k = 0; k < 8; k++) {
        const u32 nz = cc_nz_flags4(v[k]);
        F |= ((nz * 0x00204081u) >> 28) << (4 * k);
        const u32 l4 = __funnelshift_l(pv, v[k], 8);      // the four -x neighbours
        epl += __popc(nz & cc_nz_flags4(v[k] ^ l4));
        pv = v[k];
      }
    } else {
      const u32 n = min(32u, sx - (w << 5));
      uint8_t l = w > 0 ? p[-1] : (uint8_t)0;
      for (u32 i = 0; i < n; i++) {
        const uint8_t c = p[i];
        if (c) { F |= 1u << i; epl += (c != l); }
        l = c;
      }
    }
    Fb[j] = F;
  }
  epl = __reduce_add_sync(CC_FULL, epl);
  if ((threadIdx.x & 31) == 0 && epl) atomicAdd((unsigned long long*)&ctr->epl, (unsigned long long)epl);
}

template <bool HASZ>
__global__ void __launch_bounds__(256)
k_faces_from_fg(const u32* __restrict__ Fb, u32* __restrict__ M, Geom g) {
  CC_PDL_WAIT();
  const u32 j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= (u32)g.nwords) return;
  const u32 W = (u32)g.W, sy = (u32)g.sy;
  const u32 row = j / W, w = j - row * W;
  const u32 z = row / sy, y = row - z * sy;
  const u32 F = __ldg(Fb + j);
  const u32 left = w > 0 ? (__ldg(Fb + j - 1) >> 31) : 0u;
  const u32 X = F & ((F << 1) | left);
  const u32 Y = y > 0 ? (F & __ldg(Fb + j - W)) : 0u;
  const u32 Z = (HASZ && z > 0) ? (F & __ldg(Fb + j - W * sy)) : 0u;
  reinterpret_cast<uint4*>(M)[j] = make_uint4(F, X, Y, Z);
  M[g.offRS + j] = __popc(F & ~X);
}
