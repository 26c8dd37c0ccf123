// cc3d_b200.cu — C-ABI (include/cc3d_b200.h) and host orchestration of the kernel pipeline.
// No CPU fallback exists: every entry point runs CUDA kernels or fails with an error code.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/cc3d_b200.h"
#include "cc3d_dispatch.cuh"
#include "cc3d_misc.cuh"
#include "cc3d_graphs.cuh"
#include "cc3d_resolve.cuh"
#include "cc3d_runs.cuh"
#include "cc3d_blocks.cuh"

namespace {

thread_local std::string g_err;
int fail(int code, const std::string& msg) { g_err = msg; return code; }

#define CUDA_OK(expr)                                                                       \
  do {                                                                                      \
    cudaError_t e__ = (expr);                                                               \
    if (e__ != cudaSuccess)                                                                 \
      return fail(CC3D_B200_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__)); \
  } while (0)

size_t kind_size(int kind) {
  switch (kind) {
    case CC3D_B200_U8: return 1;
    case CC3D_B200_U16: return 2;
    case CC3D_B200_U32: case CC3D_B200_F32: return 4;
    case CC3D_B200_U64: case CC3D_B200_F64: return 8;
  }
  return 0;
}

// ---- device arena: one cached allocation reused across calls ----
struct Arena {
  char* base = nullptr;
  size_t cap = 0;
  size_t off = 0;
  int device = -1;
  cudaEvent_t pending = nullptr;   // set by a stream-ordered release: work that still uses the memory
  void* take(size_t bytes) {
    off = (off + 255) & ~size_t(255);
    void* p = base + off;
    off += bytes;
    return p;
  }
};
std::mutex g_pool_mu;
// Idle arenas. One is enough for back-to-back calls; the multi-slab path on ONE device (every slab's session is
// alive until the merge) parks up to CC_ARENA_POOL of them so that the next volume does not pay cudaMalloc / cudaFree
// (which synchronise the device) per slab.
#define CC_ARENA_POOL 16
std::vector<Arena> g_pool;
int g_live = 0, g_peak_live = 0;   // arenas checked out now / at most at the same time

// the work recorded by a stream-ordered release has to finish before the memory is reused (or freed)
void arena_settle(Arena& a, cudaStream_t s, bool on_stream) {
  if (!a.pending) return;
  if (on_stream) cudaStreamWaitEvent(s, a.pending, 0);
  else cudaEventSynchronize(a.pending);
  cudaEventDestroy(a.pending);
  a.pending = nullptr;
}
void arena_free(Arena& a) {
  arena_settle(a, nullptr, false);
  if (a.base) cudaFree(a.base);
  a = Arena();
}
int arena_acquire(size_t bytes, Arena* out, cudaStream_t s = nullptr, bool on_stream = false) {
  int dev = 0;
  CUDA_OK(cudaGetDevice(&dev));
  {
    std::lock_guard<std::mutex> lk(g_pool_mu);
    // best fit among the idle arenas of this device
    int best = -1;
    for (int i = 0; i < (int)g_pool.size(); i++)
      if (g_pool[i].device == dev && g_pool[i].cap >= bytes && (best < 0 || g_pool[i].cap < g_pool[best].cap)) best = i;
    if (best >= 0) {
      *out = g_pool[best]; out->off = 0;
      g_pool.erase(g_pool.begin() + best);
      arena_settle(*out, s, on_stream);
      g_peak_live = std::max(g_peak_live, ++g_live);
      return 0;
    }
  }
  Arena a;
  a.device = dev;
  a.cap = bytes + (bytes >> 4) + (1 << 20);
  cudaError_t e = cudaMalloc((void**)&a.base, a.cap);
  if (e != cudaSuccess) {
    // out of memory: drop the idle arenas and try once more
    cudaGetLastError();
    {
      std::lock_guard<std::mutex> lk(g_pool_mu);
      for (auto& p : g_pool) arena_free(p);
      g_pool.clear();
    }
    e = cudaMalloc((void**)&a.base, a.cap);
  }
  if (e != cudaSuccess) return fail(CC3D_B200_ERR_CUDA, std::string("cudaMalloc workspace: ") + cudaGetErrorString(e));
  *out = a;
  {
    std::lock_guard<std::mutex> lk(g_pool_mu);
    g_peak_live = std::max(g_peak_live, ++g_live);
  }
  return 0;
}
void arena_release(Arena& a) {
  if (!a.base) return;
  std::lock_guard<std::mutex> lk(g_pool_mu);
  a.off = 0;
  if (g_live > 0) g_live--;
  g_pool.push_back(a);
  a = Arena();
  // keep as many idle arenas as were ever checked out at the same time (1 for plain calls), the largest ones
  while ((int)g_pool.size() > std::min(std::max(g_peak_live, 1), CC_ARENA_POOL)) {
    int smallest = 0;
    for (int i = 1; i < (int)g_pool.size(); i++) if (g_pool[i].cap < g_pool[smallest].cap) smallest = i;
    arena_free(g_pool[smallest]);
    g_pool.erase(g_pool.begin() + smallest);
  }
}
// release while work enqueued on `s` still uses the arena: the next user waits for it (on its own stream)
void arena_release_after(Arena& a, cudaStream_t s) {
  if (!a.base) return;
  if (cudaEventCreateWithFlags(&a.pending, cudaEventDisableTiming) == cudaSuccess) cudaEventRecord(a.pending, s);
  else { a.pending = nullptr; cudaStreamSynchronize(s); }
  arena_release(a);
}

std::atomic<unsigned long long> g_launches{0};
std::atomic<unsigned long long> g_queue_cap_override{0};
std::atomic<bool> g_seen_big_tiles{false};   // a volume with > 16 runs per word was labelled: later calls add the RL = 5 launch   // tests: force the global edge queue to overflow

// ---- optional per-kernel timing ----
thread_local bool g_timing = false;
struct Mark { const char* name; cudaEvent_t ev; };
thread_local std::vector<Mark> g_marks;
thread_local std::vector<std::pair<const char*, float>> g_last_timings;
void mark(const char* name, cudaStream_t s) {
  if (!g_timing) return;
  Mark m; m.name = name;
  cudaEventCreate(&m.ev);
  cudaEventRecord(m.ev, s);
  g_marks.push_back(m);
}
void marks_begin(cudaStream_t s) { if (g_timing) { for (auto& m : g_marks) cudaEventDestroy(m.ev); g_marks.clear(); mark("start", s); } }
void marks_collect(bool append) {
  if (!g_timing) return;
  if (!append) g_last_timings.clear();
  for (size_t i = 1; i < g_marks.size(); i++) {
    float ms = 0;
    cudaEventSynchronize(g_marks[i].ev);
    cudaEventElapsedTime(&ms, g_marks[i - 1].ev, g_marks[i].ev);
    g_last_timings.push_back({g_marks[i].name, ms});
  }
  for (auto& m : g_marks) cudaEventDestroy(m.ev);
  g_marks.clear();
}


Geom make_geom(i64 sx, i64 sy, i64 sz) {
  Geom g;
  g.sx = sx; g.sy = sy; g.sz = sz;
  g.W = (sx + 31) / 32;
  g.rows = sy * sz;
  g.nwords = g.rows * g.W;
  auto pad4 = [](i64 n) { return (n + 3) & ~i64(3); };
  g.offRS = 4 * g.nwords;
  g.offA0 = g.offRS + pad4(g.nwords + 1);
  g.offC0 = g.offA0 + pad4(g.nwords);
  // union tile: CC_TILE_WORDS words = 2^tw words x 2^ty rows x 2^tz planes
  int tw = 0;
  while (tw < 4 && (i64(1) << tw) < g.W) tw++;
  g.tw = tw;
  g.tz = sz > 1 ? CC_TILE_TZ : 0;
  g.ty = CC_TILE_LOG - g.tw - g.tz;
  return g;
}
size_t bitmap_words(const Geom& g, bool with_diagonals) {
  return (size_t)(with_diagonals ? g.offC0 + ((g.nwords + 3) & ~i64(3)) : g.offA0);
}

// exclusive scan of n counts (n on the host, or ceil(*n_dev / 2^shift) when n_dev is given; n_max bounds it)
// cleared: the caller has already zeroed bsum[0 .. nb] earlier in the stream (keeps the kernels of the label pipeline
// back to back, see CC_PDL)
int scan_counts(const u32* cnt, u32* prefix, u64* bsum, i64 n_max, const u64* n_dev, int shift, u64* total_dev,
                u32* total32_dev, cudaStream_t s, Counters* track = nullptr, u32 W = 1, bool cleared = false) {
  const i64 nb = std::max<i64>(1, (n_max + CC_SCAN_CHUNK - 1) / CC_SCAN_CHUNK);
#ifdef CC_SCAN_THREEPASS   // reduce / scan-of-sums / apply triple; the default is the one-pass look-back scan (verified on B200: 145 parity tests, 512^3 step 0.6945 -> 0.6908 ms, 256^3 0.2243 -> 0.2189 ms)
  const unsigned grid = (unsigned)std::min<i64>(nb, CC_GRID_BLOCKS);
  k_scan_reduce<<<grid, CC_SCAN_THREADS, 0, s>>>(cnt, bsum, n_max, n_dev, shift, track, W);
  k_scan_blocks<<<1, 1024, 0, s>>>(bsum, n_max, n_dev, shift, total_dev, total32_dev);
  k_scan_apply<<<grid, CC_SCAN_THREADS, 0, s>>>(cnt, bsum, prefix, n_max, n_dev, shift);
  g_launches += 3;
#else
  // bsum holds nb chunk states + the ticket counter (callers size it (nb + 1) * 8 bytes)
  if (!cleared) {
    cudaMemsetAsync(bsum, 0, (size_t)(nb + 1) * 8, s);
    k_scan_onepass<<<(unsigned)nb, CC_SCAN_THREADS, 0, s>>>(cnt, prefix, (unsigned long long*)bsum, (u32)nb, n_max, n_dev, shift,
                                                             total_dev, total32_dev, track, W);
  } else {
    cc_launch(k_scan_onepass, dim3((unsigned)nb), dim3(CC_SCAN_THREADS), 0, s, cnt, prefix, (unsigned long long*)bsum, (u32)nb,
              n_max, n_dev, shift, total_dev, total32_dev, track, W);
  }
  g_launches += 1;
#endif
  return 0;
}

}  // namespace

struct cc3d_b200_session {
  Arena arena;
  Geom g;
  i64 voxels = 0;
  u32* L = nullptr;   // final label of every run, at the run's first voxel
  u32* M = nullptr;   // edge bitmaps (F and X planes are what the expansion needs)
  Counters* ctr = nullptr;   // device-side counters of the resolve phase
  Counters* hctr = nullptr;  // this session's pinned landing slot for the counters (never shared between sessions)
  bool epl_is_runs = false;  // multilabel: epl (cc3d.hpp:287-315) = number of x-runs, counted by scan S
  const void* din = nullptr; // device copy of the input (the caller's buffer, or the staged copy of a host buffer)
  u64 N = 0;
  u32 lmask = 0xFFFFFFFFu;   // 0x7FFFFFFF when L holds CC_LABEL_TAG | label (fused rank kernel)
  // what a host-driven redo after an edge-queue overflow needs (see redo_unions_global)
  LabelArgs args;
  int in_kind = 0;
  bool periodic = false, block_order = false, inline_fallback = false, redone = false;
  u32 *GR = nullptr, *cnt = nullptr, *prefix = nullptr;
  u64* status2 = nullptr;    // look-back status words of the C stage
  // block path (binary 26-connected, cc3d_blocks.cuh): labels live on the BLOCK runs (LB); the voxel-run labels L are
  // filled in on demand (ensure_run_labels) for the consumers that work on the run table
  const u32* M2 = nullptr;
  u32* LB = nullptr;
  Geom g2 = {};
  bool L_lazy = false;
  i64 status2_words = 0, nwords2 = 0, maxruns = 0, nbwords = 0;
};

// (definitions below inherit C linkage from their declarations in include/cc3d_b200.h)

const char* cc3d_b200_last_error(void) { return g_err.c_str(); }
const char* cc3d_b200_version(void) { return "cc3d_b200 0.1 (sm_100a)"; }
void cc3d_b200_set_timing(int enabled) { g_timing = enabled != 0; }
int cc3d_b200_last_timings(const char** names, float* ms, int cap) {
  int n = 0;
  for (auto& t : g_last_timings) { if (n >= cap) break; names[n] = t.first; ms[n] = t.second; n++; }
  return n;
}
unsigned long long cc3d_b200_launch_count(void) { return g_launches.load(); }
void cc3d_b200_debug_set_queue_capacity(uint64_t entries) { g_queue_cap_override.store(entries); }
void cc3d_b200_debug_set_big_tiles(int seen) { g_seen_big_tiles.store(seen != 0); }
size_t cc3d_b200_workspace_bytes(void) {
  std::lock_guard<std::mutex> lk(g_pool_mu);
  size_t n = 0;
  for (auto& p : g_pool) n += p.cap;
  return n;
}
void cc3d_b200_release_workspace(void) {
  std::lock_guard<std::mutex> lk(g_pool_mu);
  for (auto& p : g_pool) arena_free(p);
  g_pool.clear();
  g_peak_live = g_live;
}

static int check_shape(i64 sx, i64 sy, i64 sz) {
  if (sx < 0 || sy < 0 || sz < 0) return fail(CC3D_B200_ERR_ARGUMENT, "negative dimension");
  if (sx >= (1ll << 31) || sy >= (1ll << 31) || sz >= (1ll << 31)) return fail(CC3D_B200_ERR_ARGUMENT, "dimension >= 2^31");
  return 0;
}

template <typename T>
static int prepass_typed(const T* in, const Geom& g, Counters* ctr, T* range2, Arena& ar, cudaStream_t s) {
  const int blocks = 148 * 8;
  T* pmin = (T*)ar.take(sizeof(T) * blocks);
  T* pmax = (T*)ar.take(sizeof(T) * blocks);
  k_prepass<T><<<blocks, 256, 0, s>>>(in, g, ctr, pmin, pmax);
  k_minmax_final<T><<<1, 32, 0, s>>>(pmin, pmax, blocks, range2);
  g_launches += 2;
  return 0;
}

static int prepass_dispatch(const void* in, int kind, const Geom& g, Counters* ctr, void* range2, Arena& ar, cudaStream_t s) {
  switch (kind) {
    case CC3D_B200_U8: return prepass_typed((const uint8_t*)in, g, ctr, (uint8_t*)range2, ar, s);
    case CC3D_B200_U16: return prepass_typed((const uint16_t*)in, g, ctr, (uint16_t*)range2, ar, s);
    case CC3D_B200_U32: return prepass_typed((const uint32_t*)in, g, ctr, (uint32_t*)range2, ar, s);
    case CC3D_B200_U64: return prepass_typed((const uint64_t*)in, g, ctr, (uint64_t*)range2, ar, s);
    case CC3D_B200_F32: return prepass_typed((const float*)in, g, ctr, (float*)range2, ar, s);
    case CC3D_B200_F64: return prepass_typed((const double*)in, g, ctr, (double*)range2, ar, s);
  }
  return -1;
}

int cc3d_b200_prepass(const void* in, int in_kind, int64_t sx, int64_t sy, int64_t sz, int mem_space,
                      uint64_t* epl, int64_t* first_row, int64_t* last_row, void* vmin, void* vmax, void* stream) {
  if (int rc = check_shape(sx, sy, sz)) return rc;
  const size_t es = kind_size(in_kind);
  if (!es) return fail(CC3D_B200_ERR_KIND, "unsupported input kind");
  const i64 voxels = sx * sy * sz;
  if (epl) *epl = 0;
  if (first_row) *first_row = -1;
  if (last_row) *last_row = -1;
  if (voxels == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  Arena ar;
  const size_t need = (mem_space == CC3D_B200_HOST ? (size_t)voxels * es : 0) + (1 << 16);
  if (int rc = arena_acquire(need, &ar)) return rc;
  const void* din = in;
  if (mem_space == CC3D_B200_HOST) {
    void* d = ar.take((size_t)voxels * es);
    cudaError_t e = cudaMemcpyAsync(d, in, (size_t)voxels * es, cudaMemcpyHostToDevice, s);
    if (e != cudaSuccess) { arena_release(ar); return fail(CC3D_B200_ERR_CUDA, cudaGetErrorString(e)); }
    din = d;
  }
  Counters* ctr = (Counters*)ar.take(sizeof(Counters));
  void* range2 = ar.take(16);
  Geom g = make_geom(sx, sy, sz);
  cudaMemsetAsync(ctr, 0, sizeof(Counters), s);
  prepass_dispatch(din, in_kind, g, ctr, range2, ar, s);
  Counters h;
  unsigned char hr[16];
  cudaError_t e = cudaMemcpyAsync(&h, ctr, sizeof(h), cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaMemcpyAsync(hr, range2, 16, cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  arena_release(ar);
  if (e != cudaSuccess) return fail(CC3D_B200_ERR_CUDA, cudaGetErrorString(e));
  if (epl) *epl = h.epl;
  if (first_row) *first_row = h.epl ? (int64_t)~h.first_inv : -1;
  if (last_row) *last_row = h.epl ? (int64_t)h.last_p1 - 1 : -1;
  if (vmin) memcpy(vmin, hr, es);
  if (vmax) memcpy(vmax, hr + es, es);
  return 0;
}

#define CC_KIND_SWITCH(kind, CALL)                                  \
  switch (kind) {                                                    \
    case CC3D_B200_U8: { typedef uint8_t KT; CALL; break; }          \
    case CC3D_B200_U16: { typedef uint16_t KT; CALL; break; }        \
    case CC3D_B200_U32: { typedef uint32_t KT; CALL; break; }        \
    case CC3D_B200_U64: { typedef uint64_t KT; CALL; break; }        \
    case CC3D_B200_F32: { typedef float KT; CALL; break; }           \
    case CC3D_B200_F64: { typedef double KT; CALL; break; }          \
  }

template <typename T>
static void c8_edges_typed(const T* in, u32* M, const Geom& g, const void* delta, const void* range, cudaStream_t s) {
  T d;
  memcpy(&d, delta, sizeof(T));
  const i64 nthreads = g.nwords * 32;
  k_c8_edges<T><<<(unsigned)((nthreads + 255) / 256), 256, 0, s>>>(in, M, g, d, (const T*)range);
  g_launches += 1;
}

// Pinned landing slots for the counters, so that their copy back is truly asynchronous. One slot per live session
// (a thread-local slot would be overwritten by a second enqueue-only session of the same thread before the first one
// is read); slots are recycled through a free list.
static std::vector<Counters*> g_pin_free;
static Counters* pinned_slot_take() {
  {
    std::lock_guard<std::mutex> lk(g_pool_mu);
    if (!g_pin_free.empty()) { Counters* h = g_pin_free.back(); g_pin_free.pop_back(); return h; }
  }
  Counters* h = nullptr;
  if (cudaMallocHost((void**)&h, sizeof(Counters)) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return h;
}
static void pinned_slot_give(Counters* h) {
  if (!h) return;
  std::lock_guard<std::mutex> lk(g_pool_mu);
  g_pin_free.push_back(h);
}
static int redo_unions_global(cc3d_b200_session* S, cudaStream_t s);
// Reads the counters of an enqueued resolve phase (one stream synchronisation). If the global edge queue overflowed
// (and the fallback kernel was not launched inline) the unions are redone on the global forest here.
static int resolve_finish(cc3d_b200_session* S, cudaStream_t s, cc3d_b200_resolve_info* info) {
  if (S->voxels == 0) return 0;
  Counters hloc;
  for (int pass = 0; pass < 2; pass++) {
    Counters* h = S->hctr;
    cudaError_t e = cudaSuccess;
    if (!h || pass) { h = h ? h : &hloc; e = cudaMemcpyAsync(h, S->ctr, sizeof(Counters), cudaMemcpyDeviceToHost, s); }   // pass 0: enqueued by resolve_enqueue
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) return fail(CC3D_B200_ERR_CUDA, std::string("label_resolve: ") + cudaGetErrorString(e));
    if (h->gq_ovf && !S->inline_fallback && !S->redone) {
      if (int rc = redo_unions_global(S, s)) return rc;
      continue;
    }
    if (h->pad[0]) g_seen_big_tiles.store(true);
    S->N = h->N;
    info->N = h->N;
    info->epl = S->epl_is_runs ? h->nruns : h->epl;
    info->first_foreground_row = h->nruns ? (int64_t)~h->first_inv : -1;
    info->last_foreground_row = h->nruns ? (int64_t)h->last_p1 - 1 : -1;
    return 0;
  }
  return fail(CC3D_B200_ERR_CUDA, "label_resolve: redo did not converge");
}

static int resolve_enqueue(const void* in, int in_kind, int64_t sx, int64_t sy, int64_t sz, int connectivity,
                           const void* delta, int binary_image, int periodic_boundary, int mem_space,
                           void* stream, cc3d_b200_resolve_info* info, cc3d_b200_session** session,
                           bool inline_fallback = false);
static void enqueue_rank_stage(cc3d_b200_session* S, cudaStream_t s, bool cleared);

// Block-path sessions keep their labels on the block runs; consumers of the voxel-level run table call this first.
static void ensure_run_labels(cc3d_b200_session* S, cudaStream_t s) {
  if (!S->L_lazy) return;
  k_block_fill_L<<<(unsigned)((S->g.nwords + 255) / 256), 256, 0, s>>>(S->M, S->g, S->M2, S->g2, S->LB, S->L);
  g_launches += 1;
  S->L_lazy = false;
}

int cc3d_b200_label_resolve(const void* in, int in_kind, int64_t sx, int64_t sy, int64_t sz, int connectivity,
                            const void* delta, int binary_image, int periodic_boundary, int mem_space,
                            void* stream, cc3d_b200_resolve_info* info, cc3d_b200_session** session) {
  int rc = resolve_enqueue(in, in_kind, sx, sy, sz, connectivity, delta, binary_image, periodic_boundary, mem_space,
                           stream, info, session);
  if (rc) return rc;
  rc = resolve_finish(*session, (cudaStream_t)stream, info);
  if (rc) { cc3d_b200_session_release(*session); *session = nullptr; return rc; }
  marks_collect(false);
  return 0;
}

static int resolve_enqueue(const void* in, int in_kind, int64_t sx, int64_t sy, int64_t sz, int connectivity,
                           const void* delta, int binary_image, int periodic_boundary, int mem_space,
                           void* stream, cc3d_b200_resolve_info* info, cc3d_b200_session** session,
                           bool inline_fallback) {
  if (!info || !session) return fail(CC3D_B200_ERR_ARGUMENT, "info/session must not be NULL");
  *session = nullptr;
  info->N = 0; info->epl = 0; info->first_foreground_row = -1; info->last_foreground_row = -1;
  if (int rc = check_shape(sx, sy, sz)) return rc;
  const size_t es = kind_size(in_kind);
  if (!es) return fail(CC3D_B200_ERR_KIND, "unsupported input kind");
  if (connectivity != 4 && connectivity != 8 && connectivity != 6 && connectivity != 18 && connectivity != 26)
    return fail(CC3D_B200_ERR_CONNECTIVITY, "Only 4 and 8 2D and 6, 18, and 26 3D connectivities are supported.");
  if ((connectivity == 4 || connectivity == 8) && sz != 1)
    return fail(CC3D_B200_ERR_2D_NEEDS_SZ1, "sz must be 1 for 2D connectivities.");
  // dispatch precedence of cc3d_continuous.hpp:394-455: binary -> delta == 0 -> continuous
  bool delta_zero = true;
  if (delta) { for (size_t i = 0; i < es; i++) if (((const unsigned char*)delta)[i]) delta_zero = false; }
  if (delta && !delta_zero && (in_kind == CC3D_B200_F32 || in_kind == CC3D_B200_F64)) {
    // -0.0 compares equal to 0
    double d = in_kind == CC3D_B200_F32 ? (double)*(const float*)delta : *(const double*)delta;
    if (d == 0.0) delta_zero = true;
  }
  int mode = binary_image ? MODE_NONZERO : (delta_zero ? MODE_EQ : MODE_DELTA);
  if (mode == MODE_DELTA && periodic_boundary)
    return fail(CC3D_B200_ERR_PERIODIC_CONTINUOUS, "periodic_boundary is not currently supported for continuous data.");
  const bool c8 = (mode == MODE_DELTA && connectivity == 8);
  const bool block_order = (mode == MODE_NONZERO && connectivity == 8);

  const i64 voxels = sx * sy * sz;
  cc3d_b200_session* S = new cc3d_b200_session();
  S->voxels = voxels;
  if (voxels == 0) { *session = S; return 0; }
  if ((u64)voxels >= 0xFFFFFFFFull) {
    delete S;
    return fail(CC3D_B200_ERR_TOO_LARGE, "volume has >= 2^32-1 voxels; use the sharded path");
  }
  cudaStream_t s = (cudaStream_t)stream;
  Geom g = make_geom(sx, sy, sz);
  S->g = g;
  const i64 nwords = g.nwords;
  const i64 nb = (nwords + CC_SCAN_CHUNK - 1) / CC_SCAN_CHUNK;
  // upper bound on the number of x-runs: every voxel (multilabel / continuous), every other voxel (binary)
  const i64 maxruns = (mode == MODE_NONZERO) ? g.rows * ((sx + 1) / 2) : voxels;
  const i64 nwords2 = (maxruns + 31) / 32;            // root-flag words
  const i64 nb2 = (nwords2 + CC_SCAN_CHUNK - 1) / CC_SCAN_CHUNK;
  const i64 nblocks2d = block_order ? ((sx + 1) / 2) * ((sy + 1) / 2) : 0;
  const i64 nbwords = (nblocks2d + 31) / 32;

  // fused rank kernel (C stage in one launch): run ids and labels must stay below 2^31, block-order numbering keeps
  // the three-kernel path
  // Measured on B200 (profiles/r02_rank_ab.md): the fused kernel costs 92 us at 512^3 against 66 us for the three small
  // kernels (its look-back sits between two block-wide phases of every chunk), so the three-kernel path stays the
  // default; CC3D_B200_FUSED_RANK=1 selects the fused kernel (one launch less).
  static const bool want_fused_rank = getenv("CC3D_B200_FUSED_RANK") != nullptr;
  // binary 26-connected volumes: the unions are solved on the grid of 2x2x2 blocks, which also ranks (cc3d_blocks.cuh)
  static const bool no_blocks = getenv("CC3D_B200_NO_BLOCKS") != nullptr;
  const bool use_blocks = mode == MODE_NONZERO && connectivity == 26 && !no_blocks;
  const bool fused_rank = !block_order && !use_blocks && maxruns < (i64(1) << 31) && want_fused_rank;
  const i64 nb_rank = (maxruns + CC_RANK_RUNS - 1) / CC_RANK_RUNS;
  const i64 status2_words = (fused_rank ? nb_rank : nb2) + 2;
  // control block: Counters | scan-S status | C-stage status, zeroed by ONE memset per call
  const i64 ntiles = ((g.W + (i64(1) << g.tw) - 1) >> g.tw) * ((sy + (i64(1) << g.ty) - 1) >> g.ty) * ((sz + (i64(1) << g.tz) - 1) >> g.tz);
  const bool defer_big = g_seen_big_tiles.load() && (mode == MODE_EQ || c8);
  const size_t ctl_bytes = ((sizeof(Counters) + 255) & ~size_t(255)) + (size_t)(nb + 2) * 8 + (size_t)status2_words * 8 +
                           (defer_big ? (size_t)ntiles * 4 : 0);

  size_t need = 4096;
  auto add = [&](size_t b) { need += ((b + 255) & ~size_t(255)) + 256; };
  if (mem_space == CC3D_B200_HOST) { add((size_t)voxels * es); add((size_t)voxels * 4 + 512); }   // staged input + (u16/u32) output
  add((size_t)maxruns * 4);                // L
  add(bitmap_words(g, c8) * 4);            // M
  add((size_t)nwords2 * 4 * 3);            // GR cnt prefix
  add(ctl_bytes);
  size_t gqcap = (size_t)std::min<i64>(8 * nwords + 4096, 0x7FFFFFFF);   // global edge queue entries
  if (g_queue_cap_override.load()) gqcap = (size_t)g_queue_cap_override.load();
  add(gqcap * 8); add(64);
  add(64); add(148 * 8 * 8 * 2 + 512);
  if (block_order) { add((size_t)maxruns * 4); add((size_t)nbwords * 4 * 3 + 64); add(((nbwords + CC_SCAN_CHUNK - 1) / CC_SCAN_CHUNK + 1) * 8); }
  Geom g2 = {};
  i64 maxruns_b = 0, nb_b = 0; size_t gqcap_b = 0, ctl_b = 0, occ_b = 0;
  if (use_blocks) {
    g2 = make_geom((sx + 1) / 2, (sy + 1) / 2, (sz + 1) / 2);
    maxruns_b = g2.rows * g2.sx;          // neighbouring occupied blocks need not be joined: up to one run per block
    nb_b = (g2.nwords + CC_SCAN_CHUNK - 1) / CC_SCAN_CHUNK;
    gqcap_b = (size_t)std::min<i64>(8 * g2.nwords + 4096, 0x7FFFFFFF);
    ctl_b = ((sizeof(Counters) + 255) & ~size_t(255)) + (size_t)(nb_b + 2) * 8;
    occ_b = (size_t)g2.sx * g2.sy * g2.sz + 64;
    add(occ_b); add(bitmap_words(g2, false) * 4); add((size_t)maxruns_b * 4); add((size_t)maxruns_b * 4); add(gqcap_b * 8); add(ctl_b);
  }
  if (int rc = arena_acquire(need, &S->arena, (cudaStream_t)stream, true)) { delete S; return rc; }
  Arena& ar = S->arena;
  // enqueue-only sessions (slab_begin) never read the counters on the host: no landing slot, no copy
  S->hctr = inline_fallback ? nullptr : pinned_slot_take();

  marks_begin(s);
  const void* din = in;
  if (mem_space == CC3D_B200_HOST) {
    void* d = ar.take((size_t)voxels * es);
    cudaError_t e = cudaMemcpyAsync(d, in, (size_t)voxels * es, cudaMemcpyHostToDevice, s);
    if (e != cudaSuccess) { cc3d_b200_session_release(S); return fail(CC3D_B200_ERR_CUDA, cudaGetErrorString(e)); }
    din = d;
    mark("H2D", s);
  }
  u32* M = (u32*)ar.take(bitmap_words(g, c8) * 4);
  u32* L = (u32*)ar.take((size_t)maxruns * 4);
  u32* GR = (u32*)ar.take((size_t)nwords2 * 4);
  u32* cnt = (u32*)ar.take((size_t)nwords2 * 4);
  u32* prefix = (u32*)ar.take((size_t)nwords2 * 4);
  char* ctl = (char*)ar.take(ctl_bytes);
  Counters* ctr = (Counters*)ctl;
  u64* bsum = (u64*)(ctl + ((sizeof(Counters) + 255) & ~size_t(255)));
  u64* bsum2 = bsum + (nb + 2);
  u32* bigflags = defer_big ? (u32*)(bsum2 + status2_words) : nullptr;
  u64* gqbuf = (u64*)ar.take(gqcap * 8);
  S->L = L; S->M = M; S->din = din; S->ctr = ctr;
  S->epl_is_runs = (mode == MODE_EQ);
  S->in_kind = in_kind; S->periodic = periodic_boundary && (connectivity == 4 || connectivity == 8 || connectivity == 6);
  S->block_order = block_order; S->inline_fallback = inline_fallback;
  S->GR = GR; S->cnt = cnt; S->prefix = prefix; S->status2 = bsum2; S->status2_words = status2_words;
  S->nwords2 = nwords2; S->maxruns = maxruns; S->nbwords = nbwords;
  S->lmask = fused_rank ? 0x7FFFFFFFu : 0xFFFFFFFFu;

  // counters, edge-queue control and the look-back status words of both scans: one memset, no init kernel
  cudaMemsetAsync(ctl, 0, ctl_bytes, s);
  const bool scans_cleared = true;

  LabelArgs& a = S->args;
  a.GQ.q = gqbuf; a.GQ.count = &ctr->gq_count; a.GQ.ovf = &ctr->gq_ovf; a.GQ.cap = (u32)gqcap;
  int stage_launches = 0;
  a.launches = &stage_launches;
  a.in = din; a.M = M; a.L = L; a.ctr = ctr; a.g = g; a.mode = mode;
  a.connectivity = connectivity; a.stream = s;
  a.mark = g_timing ? mark : nullptr;
  a.inline_fallback = inline_fallback;
  a.bigflags = bigflags; a.nbig = &ctr->pad[0]; a.defer_big = defer_big;
  memset(a.delta, 0, 8);
  if (delta) memcpy(a.delta, delta, es);
  int rc = 0;
  bool ranked = false;     // the block path numbers its components itself
  // A: face bitmaps
  if (c8) {
    // epl + value range, then the reference's per-pixel edge rule straight into the bitmaps
    void* range2 = ar.take(16);
    prepass_dispatch(din, in_kind, g, ctr, range2, ar, s);
    CC_KIND_SWITCH(in_kind, c8_edges_typed((const KT*)din, M, g, delta, range2, s));
    a.mode = MODE_MASK;
    mark("A_c8_edges", s);
  } else {
    rc = -1;
    CC_KIND_SWITCH(in_kind, rc = run_faces_stage<KT>(a));
    mark("A_faces", s);
  }
  if (rc == 0) {
    // S: number the runs
    u32* RS = M + g.offRS;
    scan_counts(RS, RS, bsum, nwords, nullptr, 0, &ctr->nruns, RS + nwords, s, ctr, (u32)g.W, scans_cleared);
    mark("S_scan_runs", s);
    // B: unions
    if (use_blocks) {
      // the same pipeline one level up, on the occupancy bytes of the 2x2x2 blocks
      uint8_t* occ = (uint8_t*)ar.take(occ_b);
      u32* M2 = (u32*)ar.take(bitmap_words(g2, false) * 4);
      u32* L2 = (u32*)ar.take((size_t)maxruns_b * 4);
      u32* minrun = (u32*)ar.take((size_t)maxruns_b * 4);
      u64* gq2 = (u64*)ar.take(gqcap_b * 8);
      char* ctl2 = (char*)ar.take(ctl_b);
      Counters* ctr2 = (Counters*)ctl2;
      u64* bsum_b = (u64*)(ctl2 + ((sizeof(Counters) + 255) & ~size_t(255)));
      cudaMemsetAsync(ctl2, 0, ctl_b, s);
      const u32 BX = (u32)g2.sx, BY = (u32)g2.sy, BZ = (u32)g2.sz;
      const i64 nthr = g.W * (i64)BY * BZ;
      k_block_occ<<<(unsigned)((nthr + 255) / 256), 256, 0, s>>>(M, g, occ, BX, BY, BZ);
      mark("Bb_occupancy", s);
      LabelArgs b = a;
      int launches_b = 0;
      b.launches = &launches_b;
      b.in = occ; b.M = M2; b.L = L2; b.ctr = ctr2; b.g = g2; b.mode = MODE_BLOCK; b.connectivity = 26;
      b.GQ.q = gq2; b.GQ.count = &ctr2->gq_count; b.GQ.ovf = &ctr2->gq_ovf; b.GQ.cap = (u32)gqcap_b;
      b.mark = nullptr; b.inline_fallback = true; b.bigflags = nullptr; b.nbig = nullptr; b.defer_big = false;
      memset(b.delta, 0, 8);
      rc = run_faces_stage<uint8_t>(b);
      mark("Bb_faces", s);
      if (rc == 0) {
        u32* RS2 = M2 + g2.offRS;
        scan_counts(RS2, RS2, bsum_b, g2.nwords, nullptr, 0, &ctr2->nruns, RS2 + g2.nwords, s, nullptr, 1, true);
        mark("Bb_scan", s);
        rc = run_union_stage<uint8_t>(b);
        mark("Bb_unions", s);
      }
      if (rc == 0) {
        k_block_flatten<<<CC_GRID_BLOCKS, 256, 0, s>>>(L2, &ctr2->nruns);
        k_fill_n<<<CC_GRID_BLOCKS, 256, 0, s>>>(minrun, CC_BG, &ctr2->nruns);
        mark("Bb_flatten", s);
        cudaMemsetAsync(GR, 0, (size_t)nwords2 * 4, s);
        k_block_minrun<<<(unsigned)((g2.nwords + 255) / 256), 256, 0, s>>>(M, g, occ, M2, g2, L2, minrun);
        mark("Bb_minrun", s);
        // C stage of this path: root flags from the block roots, scan, labels on the block runs (cc3d_blocks.cuh)
        k_block_rootflags<<<CC_GRID_BLOCKS, 256, 0, s>>>(L2, minrun, GR, &ctr2->nruns);
        k_popc_n<<<CC_GRID_BLOCKS, 256, 0, s>>>(GR, cnt, &ctr->nruns);
        scan_counts(cnt, prefix, bsum2, nwords2, &ctr->nruns, 5, &ctr->N, nullptr, s, nullptr, 1, scans_cleared);
        mark("C2_scan", s);
        k_block_labels<<<CC_GRID_BLOCKS, 256, 0, s>>>(L2, minrun, GR, prefix, &ctr2->nruns);
        mark("C3_assign", s);
        stage_launches += launches_b + 7;
        ranked = true;
        S->M2 = M2; S->LB = L2; S->g2 = g2; S->L_lazy = true;
        // CC3D_B200_BLOCK_LAZY=0: fill the voxel-run labels right away and expand from them (round 2c behaviour)
        static const bool eager = getenv("CC3D_B200_BLOCK_LAZY") && atoi(getenv("CC3D_B200_BLOCK_LAZY")) == 0;
        if (eager) { ensure_run_labels(S, s); S->LB = nullptr; mark("C3_fill_L", s); }
      }
    } else {
      rc = -1;
      CC_KIND_SWITCH(in_kind, rc = run_union_stage<KT>(a));
    }
  }
  if (rc == 0 && S->periodic) {
    CC_KIND_SWITCH(in_kind, rc = run_periodic_stage<KT>(a));
    mark("P_periodic", s);
  }
  g_launches += stage_launches;
  a.launches = nullptr;
  if (rc != 0) { cc3d_b200_session_release(S); return fail(CC3D_B200_ERR_ARGUMENT, "no kernel for this configuration"); }
  if (!ranked) enqueue_rank_stage(S, s, scans_cleared);
  if (S->hctr) cudaMemcpyAsync(S->hctr, ctr, sizeof(Counters), cudaMemcpyDeviceToHost, s);
  *session = S;
  return 0;
}

// C stage: root flags, their scan and the final label of every run (fused k_rank, or the three-kernel path for
// block-order numbering / >= 2^31 runs). cleared: the status words were zeroed earlier in the stream.
static void enqueue_rank_stage(cc3d_b200_session* S, cudaStream_t s, bool cleared) {
  Counters* ctr = S->ctr;
  u32 *L = S->L, *GR = S->GR, *cnt = S->cnt, *prefix = S->prefix;
  const Geom& g = S->g;
  if (S->lmask != 0xFFFFFFFFu) {
    const i64 nb_rank = std::max<i64>(1, (S->maxruns + CC_RANK_RUNS - 1) / CC_RANK_RUNS);
    if (!cleared) cudaMemsetAsync(S->status2, 0, (size_t)S->status2_words * 8, s);
    const unsigned grid = (unsigned)std::min<i64>(nb_rank, 148 * 8);
    cc_launch(k_rank, dim3(grid), dim3(256), 0, s, L, GR, prefix, (unsigned long long*)S->status2, (u32)nb_rank, &ctr->nruns, &ctr->N);
    g_launches += 1;
    mark("C_rank", s);
    return;
  }
  // CC3D_B200_C_ILP=0: one run per lane (rounds 1-2c); default: four (k_compress4 / k_assign4, round 2d)
  static const bool ilp4 = !(getenv("CC3D_B200_C_ILP") && atoi(getenv("CC3D_B200_C_ILP")) == 0);
  if (ilp4) cc_launch(k_compress4, dim3(CC_GRID_BLOCKS), dim3(256), 0, s, L, GR, cnt, &ctr->nruns);
  else cc_launch(k_compress, dim3(CC_GRID_BLOCKS), dim3(256), 0, s, L, GR, cnt, &ctr->nruns);
  g_launches += 1;
  mark("C1_compress", s);
  scan_counts(cnt, prefix, S->status2, S->nwords2, &ctr->nruns, 5, &ctr->N, nullptr, s, nullptr, 1, cleared);
  mark("C2_scan", s);
  if (S->block_order) {
    Arena& ar = S->arena;
    const i64 nbwords = S->nbwords, nwords = g.nwords;
    u32* K = (u32*)ar.take((size_t)S->maxruns * 4);
    u32* BK = (u32*)ar.take((size_t)nbwords * 4);
    u32* bcnt = (u32*)ar.take((size_t)nbwords * 4);
    u32* bprefix = (u32*)ar.take((size_t)nbwords * 4);
    u64* bsum3 = (u64*)ar.take(((nbwords + CC_SCAN_CHUNK - 1) / CC_SCAN_CHUNK + 2) * 8);
    u64* dummyN = (u64*)ar.take(8);
    cudaMemsetAsync(BK, 0, (size_t)nbwords * 4, s);
    k_fill_n<<<CC_GRID_BLOCKS, 256, 0, s>>>(K, CC_BG, &ctr->nruns);
    k_blockkey_min<<<(unsigned)((nwords + 255) / 256), 256, 0, s>>>(L, S->M, K, g);
    k_blockkey_mark<<<CC_GRID_BLOCKS, 256, 0, s>>>(K, GR, BK, &ctr->nruns);
    k_popc<<<(unsigned)((nbwords + 255) / 256), 256, 0, s>>>(BK, bcnt, nbwords);
    scan_counts(bcnt, bprefix, bsum3, nbwords, nullptr, 0, dummyN, nullptr, s);
    k_assign_blockorder<<<CC_GRID_BLOCKS, 256, 0, s>>>(L, K, BK, bprefix, &ctr->nruns);
    g_launches += 5;
    mark("C3_assign_blockorder", s);
  } else {
    if (ilp4) cc_launch(k_assign4, dim3(CC_GRID_BLOCKS), dim3(256), 0, s, L, GR, prefix, &ctr->nruns);
    else cc_launch(k_assign, dim3(CC_GRID_BLOCKS), dim3(256), 0, s, L, GR, prefix, &ctr->nruns);
    g_launches += 1;
    mark("C3_assign", s);
  }
}

// The global edge queue overflowed (B1 raised gq_ovf) and the fallback kernel was not part of the enqueued pipeline:
// reset the forest and unite EVERY edge on it (k_union_global), then the wrap edges and the C stage again. Rare (the
// queue holds 8 entries per bitmap word); costs one extra synchronisation only when it happens.
__global__ void __launch_bounds__(256) k_iota_n(u32* __restrict__ p, const u64* __restrict__ n_dev) {
  const u32 n = (u32)*n_dev;
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) p[i] = i;
}
static int redo_unions_global(cc3d_b200_session* S, cudaStream_t s) {
  S->redone = true;
  LabelArgs& a = S->args;
  int launches = 0;
  a.launches = &launches;
  a.stream = s;
  a.mark = nullptr;
  k_iota_n<<<CC_GRID_BLOCKS, 256, 0, s>>>(S->L, &S->ctr->nruns);
  g_launches += 1;
  int rc = -1;
  CC_KIND_SWITCH(S->in_kind, rc = run_union_global_stage<KT>(a));
  if (rc == 0 && S->periodic) { CC_KIND_SWITCH(S->in_kind, rc = run_periodic_stage<KT>(a)); }
  g_launches += launches;
  a.launches = nullptr;
  if (rc != 0) return fail(CC3D_B200_ERR_ARGUMENT, "no kernel for this configuration");
  cudaMemsetAsync(&S->ctr->N, 0, sizeof(u64), s);
  enqueue_rank_stage(S, s, false);
  return 0;
}

void cc3d_b200_session_release(cc3d_b200_session* S) {
  if (!S) return;
  arena_release(S->arena);
  pinned_slot_give(S->hctr);
  delete S;
}
// release while work enqueued on `s` may still use the session's workspace (enqueue-only paths, error paths)
static void session_release_after(cc3d_b200_session* S, cudaStream_t s) {
  if (!S) return;
  arena_release_after(S->arena, s);
  // the pinned slot may still receive the counters copy enqueued on s: recycle it only once the stream has passed it
  if (S->hctr) { cudaStreamSynchronize(s); pinned_slot_give(S->hctr); }
  delete S;
}

template <typename OUT>
static void launch_write(const cc3d_b200_session* S, OUT* dout, i64 row0, i64 nrows, const void* remap,
                         int remap_kind, cudaStream_t s) {
  const Geom& g = S->g;
  const unsigned nchunks = (unsigned)((g.W + 31) / 32);
  const i64 nwarps = nrows * nchunks;
  const unsigned blocks = (unsigned)((nwarps + 7) / 8);
  if (S->LB) {
    // block path: straight from the labels of the block runs
    if (!remap) k_expand_blocks<OUT, 0><<<blocks, 256, 0, s>>>(S->LB, S->M, S->M2, dout, g, S->g2, nchunks, (u32)row0, (u32)nwarps, remap);
    else if (remap_kind == CC3D_B200_U32) k_expand_blocks<OUT, 1><<<blocks, 256, 0, s>>>(S->LB, S->M, S->M2, dout, g, S->g2, nchunks, (u32)row0, (u32)nwarps, remap);
    else k_expand_blocks<OUT, 2><<<blocks, 256, 0, s>>>(S->LB, S->M, S->M2, dout, g, S->g2, nchunks, (u32)row0, (u32)nwarps, remap);
    g_launches += 1;
    return;
  }
  if (!remap) k_expand<OUT, 0><<<blocks, 256, 0, s>>>(S->L, S->M, dout, g, nchunks, (u32)row0, (u32)nwarps, remap, S->lmask);
  else if (remap_kind == CC3D_B200_U32) k_expand<OUT, 1><<<blocks, 256, 0, s>>>(S->L, S->M, dout, g, nchunks, (u32)row0, (u32)nwarps, remap, S->lmask);
  else k_expand<OUT, 2><<<blocks, 256, 0, s>>>(S->L, S->M, dout, g, nchunks, (u32)row0, (u32)nwarps, remap, S->lmask);
  g_launches += 1;
}

// shared implementation of label_write / label_write_rows / label_write_remap
static int write_impl(cc3d_b200_session* S, void* out, int out_kind, int mem_space, void* stream, i64 row0, i64 nrows,
                      const void* remap, int remap_kind, u64 max_label, bool release, bool collect_marks = true) {
  if (!S) return fail(CC3D_B200_ERR_ARGUMENT, "NULL session");
  auto done = [&](int rc) { if (release) cc3d_b200_session_release(S); return rc; };
  const size_t os = (out_kind == CC3D_B200_U16) ? 2 : (out_kind == CC3D_B200_U32 ? 4 : (out_kind == CC3D_B200_U64 ? 8 : 0));
  if (!os) return done(fail(CC3D_B200_ERR_KIND, "out kind must be u16, u32 or u64"));
  if (S->voxels == 0 || nrows == 0) return done(0);
  if (row0 < 0 || nrows < 0 || row0 + nrows > S->g.sy * S->g.sz) return done(fail(CC3D_B200_ERR_ARGUMENT, "row range outside the volume"));
  if (remap && remap_kind != CC3D_B200_U32 && remap_kind != CC3D_B200_U64) return done(fail(CC3D_B200_ERR_KIND, "remap kind must be u32 or u64"));
  if ((os == 2 && max_label > 0xFFFFull) || (os == 4 && max_label > 0xFFFFFFFFull))
    return done(fail(CC3D_B200_ERR_OUT_RANGE, "N does not fit the requested output kind"));
  cudaStream_t s = (cudaStream_t)stream;
  const Geom& g = S->g;
  const size_t nvox = (size_t)nrows * (size_t)g.sx;
  void* dout = out;
  const void* dremap = remap;
  void* tmp_out = nullptr;
  void* tmp_remap = nullptr;
  if (mem_space == CC3D_B200_HOST) {
    // the arena was sized for the resolve phase; staging buffers may need their own allocation
    if (S->arena.off + nvox * os + 512 <= S->arena.cap) dout = S->arena.take(nvox * os);
    else {
      cudaError_t e = cudaMalloc(&tmp_out, nvox * os);
      if (e != cudaSuccess) return done(fail(CC3D_B200_ERR_CUDA, cudaGetErrorString(e)));
      dout = tmp_out;
    }
    if (remap) {
      const size_t rb = (size_t)(S->N + 1) * (remap_kind == CC3D_B200_U32 ? 4 : 8);
      cudaError_t e = cudaMalloc(&tmp_remap, rb);
      if (e == cudaSuccess) e = cudaMemcpyAsync(tmp_remap, remap, rb, cudaMemcpyHostToDevice, s);
      if (e != cudaSuccess) { if (tmp_out) cudaFree(tmp_out); return done(fail(CC3D_B200_ERR_CUDA, cudaGetErrorString(e))); }
      dremap = tmp_remap;
    }
  }
  marks_begin(s);
  if (os == 2) launch_write<uint16_t>(S, (uint16_t*)dout, row0, nrows, dremap, remap_kind, s);
  else if (os == 4) launch_write<uint32_t>(S, (uint32_t*)dout, row0, nrows, dremap, remap_kind, s);
  else launch_write<uint64_t>(S, (uint64_t*)dout, row0, nrows, dremap, remap_kind, s);
  mark("D_expand", s);
  cudaError_t e = cudaSuccess;
  if (mem_space == CC3D_B200_HOST) {
    e = cudaMemcpyAsync(out, dout, nvox * os, cudaMemcpyDeviceToHost, s);
    mark("D2H", s);
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (collect_marks) marks_collect(true);
  if (tmp_out) cudaFree(tmp_out);
  if (tmp_remap) cudaFree(tmp_remap);
  if (e != cudaSuccess) return done(fail(CC3D_B200_ERR_CUDA, std::string("label_write: ") + cudaGetErrorString(e)));
  return done(0);
}

int cc3d_b200_label_write(cc3d_b200_session* S, void* out, int out_kind, int mem_space, void* stream) {
  if (!S) return fail(CC3D_B200_ERR_ARGUMENT, "NULL session");
  return write_impl(S, out, out_kind, mem_space, stream, 0, S->g.sy * S->g.sz, nullptr, 0, S->N, true);
}

int cc3d_b200_label_write_rows(cc3d_b200_session* S, int64_t row_begin, int64_t row_end, uint32_t* out, int mem_space,
                               void* stream) {
  return write_impl(S, out, CC3D_B200_U32, mem_space, stream, row_begin, row_end - row_begin, nullptr, 0, 0, false);
}

int cc3d_b200_label_write_remap(cc3d_b200_session* S, const void* remap, int remap_kind, uint64_t max_label, void* out,
                                int out_kind, int mem_space, void* stream) {
  if (!S) return fail(CC3D_B200_ERR_ARGUMENT, "NULL session");
  if (!remap) { cc3d_b200_session_release(S); return fail(CC3D_B200_ERR_ARGUMENT, "NULL remap table"); }
  return write_impl(S, out, out_kind, mem_space, stream, 0, S->g.sy * S->g.sz, remap, remap_kind, max_label, true);
}

template <typename T>
static void face_pairs_typed(const T* vP, const u32* lP, const T* vQ, const u32* lQ, i64 sx, i64 sy, int connectivity,
                             int mode, const void* delta, u64* pairs, unsigned long long cap, unsigned long long* count,
                             cudaStream_t s) {
  const i64 n = sx * sy;
  const unsigned blocks = (unsigned)((n + 255) / 256);
  if (mode == MODE_EQ) { Edge<T, MODE_EQ> E; E.delta = (T)0; E.zeq = 0; k_face_pairs<T, MODE_EQ><<<blocks, 256, 0, s>>>(vP, lP, vQ, lQ, sx, sy, connectivity, E, pairs, cap, count); }
  else if (mode == MODE_NONZERO) { Edge<T, MODE_NONZERO> E; E.delta = (T)0; E.zeq = 0; k_face_pairs<T, MODE_NONZERO><<<blocks, 256, 0, s>>>(vP, lP, vQ, lQ, sx, sy, connectivity, E, pairs, cap, count); }
  else { Edge<T, MODE_DELTA> E; memcpy(&E.delta, delta, sizeof(T)); E.zeq = connectivity == 26 ? 1 : 0; k_face_pairs<T, MODE_DELTA><<<blocks, 256, 0, s>>>(vP, lP, vQ, lQ, sx, sy, connectivity, E, pairs, cap, count); }
  g_launches += 1;
}

int cc3d_b200_face_pairs(const void* values_upper, const uint32_t* labels_upper, const void* values_lower,
                         const uint32_t* labels_lower, int in_kind, int64_t sx, int64_t sy, int connectivity,
                         const void* delta, int binary_image, uint64_t* pairs, uint64_t capacity, uint64_t* count,
                         void* stream) {
  const size_t es = kind_size(in_kind);
  if (!es) return fail(CC3D_B200_ERR_KIND, "unsupported input kind");
  if (connectivity != 6 && connectivity != 18 && connectivity != 26)
    return fail(CC3D_B200_ERR_CONNECTIVITY, "sharded volumes support 6, 18 and 26 connectivity");
  if (!count) return fail(CC3D_B200_ERR_ARGUMENT, "count must not be NULL");
  *count = 0;
  if (sx * sy == 0) return 0;
  bool delta_zero = true;
  if (delta) { for (size_t i = 0; i < es; i++) if (((const unsigned char*)delta)[i]) delta_zero = false; }
  const int mode = binary_image ? MODE_NONZERO : (delta_zero ? MODE_EQ : MODE_DELTA);
  cudaStream_t s = (cudaStream_t)stream;
  // one 8-byte device counter per device, allocated once (cudaMalloc/cudaFree would synchronise every call)
  static unsigned long long* dcounts[64] = {nullptr};
  int devid = 0;
  CUDA_OK(cudaGetDevice(&devid));
  if (devid < 0 || devid >= 64) return fail(CC3D_B200_ERR_ARGUMENT, "device index out of range");
  {
    std::lock_guard<std::mutex> lk(g_pool_mu);
    if (!dcounts[devid]) CUDA_OK(cudaMalloc((void**)&dcounts[devid], 8));
  }
  unsigned long long* dcount = dcounts[devid];
  cudaMemsetAsync(dcount, 0, 8, s);
  const u32 *lP = labels_upper, *lQ = labels_lower;
  switch (in_kind) {
    case CC3D_B200_U8: face_pairs_typed((const uint8_t*)values_upper, lP, (const uint8_t*)values_lower, lQ, sx, sy, connectivity, mode, delta, pairs, capacity, dcount, s); break;
    case CC3D_B200_U16: face_pairs_typed((const uint16_t*)values_upper, lP, (const uint16_t*)values_lower, lQ, sx, sy, connectivity, mode, delta, pairs, capacity, dcount, s); break;
    case CC3D_B200_U32: face_pairs_typed((const uint32_t*)values_upper, lP, (const uint32_t*)values_lower, lQ, sx, sy, connectivity, mode, delta, pairs, capacity, dcount, s); break;
    case CC3D_B200_U64: face_pairs_typed((const uint64_t*)values_upper, lP, (const uint64_t*)values_lower, lQ, sx, sy, connectivity, mode, delta, pairs, capacity, dcount, s); break;
    case CC3D_B200_F32: face_pairs_typed((const float*)values_upper, lP, (const float*)values_lower, lQ, sx, sy, connectivity, mode, delta, pairs, capacity, dcount, s); break;
    default: face_pairs_typed((const double*)values_upper, lP, (const double*)values_lower, lQ, sx, sy, connectivity, mode, delta, pairs, capacity, dcount, s); break;
  }
  unsigned long long h = 0;
  cudaError_t e = cudaMemcpyAsync(&h, dcount, 8, cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) return fail(CC3D_B200_ERR_CUDA, std::string("face_pairs: ") + cudaGetErrorString(e));
  *count = h;
  return 0;
}

// ---- sharded fast path: everything enqueued on the caller's stream, no host synchronisation ----
__global__ void k_slab_facts(const Counters* __restrict__ ctr, int epl_is_runs, long long sz, long long* __restrict__ out) {
  out[0] = (long long)ctr->N;
  out[1] = (long long)(epl_is_runs ? ctr->nruns : ctr->epl);
  out[2] = sz;
}

int cc3d_b200_slab_begin(const void* in, int in_kind, int64_t sx, int64_t sy, int64_t sz, int connectivity,
                         const void* delta, int binary_image, void* stream, cc3d_b200_session** session,
                         uint32_t* labels_first_plane, uint32_t* labels_last_plane, int64_t* facts) {
  if (!session || !facts) return fail(CC3D_B200_ERR_ARGUMENT, "session/facts must not be NULL");
  if (connectivity != 6 && connectivity != 18 && connectivity != 26)
    return fail(CC3D_B200_ERR_CONNECTIVITY, "sharded volumes support 6, 18 and 26 connectivity");
  if (sx <= 0 || sy <= 0 || sz <= 0) return fail(CC3D_B200_ERR_ARGUMENT, "empty slab");
  cc3d_b200_resolve_info info;
  cc3d_b200_session* S = nullptr;
  int rc = resolve_enqueue(in, in_kind, sx, sy, sz, connectivity, delta, binary_image, 0, CC3D_B200_DEVICE, stream, &info, &S,
                           /*inline_fallback=*/true);   // nothing synchronises here: the overflow fallback rides along
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  if (labels_first_plane) launch_write<uint32_t>(S, labels_first_plane, 0, sy, nullptr, 0, s);
  if (labels_last_plane) launch_write<uint32_t>(S, labels_last_plane, (sz - 1) * sy, sy, nullptr, 0, s);
  k_slab_facts<<<1, 1, 0, s>>>(S->ctr, S->epl_is_runs ? 1 : 0, (long long)sz, (long long*)facts);
  g_launches += 1;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { session_release_after(S, s); return fail(CC3D_B200_ERR_CUDA, std::string("slab_begin: ") + cudaGetErrorString(e)); }
  *session = S;
  return 0;
}

int cc3d_b200_slab_finish(cc3d_b200_session* S, const void* remap, int remap_kind, void* out, int out_kind, void* stream) {
  if (!S) return fail(CC3D_B200_ERR_ARGUMENT, "NULL session");
  cudaStream_t s = (cudaStream_t)stream;
  int rc = 0;
  if (!remap || !out) rc = fail(CC3D_B200_ERR_ARGUMENT, "NULL remap table / output");
  else if (remap_kind != CC3D_B200_U32 && remap_kind != CC3D_B200_U64) rc = fail(CC3D_B200_ERR_KIND, "remap kind must be u32 or u64");
  else {
    const i64 nrows = S->g.sy * S->g.sz;
    if (out_kind == CC3D_B200_U16) launch_write<uint16_t>(S, (uint16_t*)out, 0, nrows, remap, remap_kind, s);
    else if (out_kind == CC3D_B200_U32) launch_write<uint32_t>(S, (uint32_t*)out, 0, nrows, remap, remap_kind, s);
    else if (out_kind == CC3D_B200_U64) launch_write<uint64_t>(S, (uint64_t*)out, 0, nrows, remap, remap_kind, s);
    else rc = fail(CC3D_B200_ERR_KIND, "out kind must be u16, u32 or u64");
    if (rc == 0 && cudaGetLastError() != cudaSuccess) rc = fail(CC3D_B200_ERR_CUDA, "slab_finish: launch failed");
  }
  arena_release_after(S->arena, s);
  pinned_slot_give(S->hctr);   // nullptr for slab sessions
  delete S;
  return rc;
}

int cc3d_b200_face_pairs_async(const void* values_upper, const uint32_t* labels_upper, const void* values_lower,
                               const uint32_t* labels_lower, int in_kind, int64_t sx, int64_t sy, int connectivity,
                               const void* delta, int binary_image, uint64_t* pairs, uint64_t capacity,
                               uint64_t* count_dev, void* stream) {
  const size_t es = kind_size(in_kind);
  if (!es) return fail(CC3D_B200_ERR_KIND, "unsupported input kind");
  if (connectivity != 6 && connectivity != 18 && connectivity != 26)
    return fail(CC3D_B200_ERR_CONNECTIVITY, "sharded volumes support 6, 18 and 26 connectivity");
  if (!count_dev || !pairs) return fail(CC3D_B200_ERR_ARGUMENT, "pairs/count must not be NULL");
  if (sx * sy == 0) return 0;
  bool delta_zero = true;
  if (delta) { for (size_t i = 0; i < es; i++) if (((const unsigned char*)delta)[i]) delta_zero = false; }
  const int mode = binary_image ? MODE_NONZERO : (delta_zero ? MODE_EQ : MODE_DELTA);
  cudaStream_t s = (cudaStream_t)stream;
  unsigned long long* dcount = (unsigned long long*)count_dev;   // caller zeroes it (stream ordered)
  const u32 *lP = labels_upper, *lQ = labels_lower;
  CC_KIND_SWITCH(in_kind, face_pairs_typed((const KT*)values_upper, lP, (const KT*)values_lower, lQ, sx, sy, connectivity, mode, delta, pairs, capacity, dcount, s));
  if (cudaGetLastError() != cudaSuccess) return fail(CC3D_B200_ERR_CUDA, "face_pairs_async: launch failed");
  return 0;
}


// ---- sharded volumes: slab merge on the device (enqueue-only; see cc3d_resolve.cuh) ----
static size_t merge_ws_layout(uint64_t label_cap, size_t* o_remap, size_t* o_nr, size_t* o_cnt, size_t* o_prefix, size_t* o_status,
                              size_t* o_result) {
  auto up = [](size_t b) { return (b + 255) & ~size_t(255); };
  const size_t words = (size_t)(label_cap / 32 + 2);
  const size_t nb = (words + CC_SCAN_CHUNK - 1) / CC_SCAN_CHUNK;
  size_t off = up((size_t)label_cap * 4);          // parent
  *o_remap = off;  off += up((size_t)label_cap * 4);
  *o_nr = off;     off += up(words * 4);
  *o_cnt = off;    off += up(words * 4);
  *o_prefix = off; off += up(words * 4);
  *o_status = off; off += up((nb + 2) * 8);
  *o_result = off; off += 8 * (8 + 4 * CC_MERGE_MAX_WORLD) + 256;
  return off;
}
size_t cc3d_b200_merge_workspace_bytes(uint64_t label_cap) {
  size_t a, b, c, d, e, f;
  return merge_ws_layout(label_cap, &a, &b, &c, &d, &e, &f);
}

int cc3d_b200_merge_slabs_device(const int64_t* gathered, int world, int64_t row_stride, int rank, uint64_t pair_cap,
                                 void* workspace, uint64_t label_cap, uint32_t** remap, uint64_t** result, void* stream) {
  if (!gathered || !workspace || !remap || !result || world <= 0 || world > CC_MERGE_MAX_WORLD || rank < 0 || rank >= world ||
      row_stride < 4 + (int64_t)pair_cap || label_cap < 64 || label_cap > 0xFFFFFFFFull)
    return fail(CC3D_B200_ERR_ARGUMENT, "merge_slabs_device: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  size_t o_remap, o_nr, o_cnt, o_prefix, o_status, o_result;
  merge_ws_layout(label_cap, &o_remap, &o_nr, &o_cnt, &o_prefix, &o_status, &o_result);
  char* ws = (char*)workspace;
  u32* parent = (u32*)ws;
  u32* rm = (u32*)(ws + o_remap);
  u32 *NR = (u32*)(ws + o_nr), *cnt = (u32*)(ws + o_cnt), *prefix = (u32*)(ws + o_prefix);
  u64* status = (u64*)(ws + o_status);
  unsigned long long* res = (unsigned long long*)(ws + o_result);
  SlabRows f; f.rows = (const long long*)gathered; f.world = world; f.stride = row_stride;
  const unsigned gb = 148 * 4;
  k_merge_init<<<gb, 256, 0, s>>>(parent, f, label_cap, pair_cap, res);
  k_merge_union<<<gb, 256, 0, s>>>(parent, f, label_cap, pair_cap);
  k_merge_flags<<<gb, 256, 0, s>>>(parent, NR, cnt, res);
  const i64 words = (i64)(label_cap / 32 + 2);
  scan_counts(cnt, prefix, status, words, (const u64*)&res[3], 5, (u64*)&res[4], nullptr, s);
  k_merge_remap<<<gb, 256, 0, s>>>(parent, NR, prefix, f, rank, rm, res);
  g_launches += 4;
  if (cudaGetLastError() != cudaSuccess) return fail(CC3D_B200_ERR_CUDA, "merge_slabs_device: launch failed");
  *remap = rm;
  *result = (uint64_t*)res;
  return 0;
}

int cc3d_b200_merge_slabs_device_small(const int64_t* gathered, int world, int64_t row_stride, int rank, uint64_t pair_cap,
                                       void* workspace, uint64_t label_cap, uint32_t** remap, uint64_t** result, void* stream) {
  if (!gathered || !workspace || !remap || !result || world <= 0 || world > CC_MERGE_MAX_WORLD || rank < 0 || rank >= world ||
      row_stride < 4 + (int64_t)pair_cap || label_cap < 64 || label_cap > 0xFFFFFFFFull)
    return fail(CC3D_B200_ERR_ARGUMENT, "merge_slabs_device_small: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  size_t o_remap, o_nr, o_cnt, o_prefix, o_status, o_result;
  merge_ws_layout(label_cap, &o_remap, &o_nr, &o_cnt, &o_prefix, &o_status, &o_result);
  char* ws = (char*)workspace;
  unsigned long long* res = (unsigned long long*)(ws + o_result);
  SlabRows f; f.rows = (const long long*)gathered; f.world = world; f.stride = row_stride;
  k_merge_small<<<1, 1024, 0, s>>>((u32*)ws, f, rank, label_cap, pair_cap, (u32*)(ws + o_remap), res);
  g_launches += 1;
  if (cudaGetLastError() != cudaSuccess) return fail(CC3D_B200_ERR_CUDA, "merge_slabs_device_small: launch failed");
  *remap = (u32*)(ws + o_remap);
  *result = (uint64_t*)res;
  return 0;
}

int cc3d_b200_solve_pairs(uint32_t* parent, int64_t n_nodes, const uint32_t* a, const uint32_t* b, int64_t n_pairs,
                          void* stream) {
  if (n_nodes <= 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  k_iota<<<(unsigned)((n_nodes + 255) / 256), 256, 0, s>>>(parent, n_nodes);
  if (n_pairs > 0) k_union_pairs<<<(unsigned)((n_pairs + 255) / 256), 256, 0, s>>>(parent, a, b, n_pairs);
  k_flatten<<<(unsigned)((n_nodes + 255) / 256), 256, 0, s>>>(parent, n_nodes);
  g_launches += 3;
  cudaError_t e = cudaStreamSynchronize(s);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) return fail(CC3D_B200_ERR_CUDA, std::string("solve_pairs: ") + cudaGetErrorString(e));
  return 0;
}

int cc3d_b200_label_with_info(const void* in, int in_kind, int64_t sx, int64_t sy, int64_t sz, int connectivity,
                              const void* delta, int binary_image, int periodic_boundary, void* out, int out_kind,
                              int mem_space, cc3d_b200_resolve_info* info, void* stream) {
  if (!info) return fail(CC3D_B200_ERR_ARGUMENT, "info must not be NULL");
  // both phases are enqueued back to back; one synchronisation at the end
  cc3d_b200_session* S = nullptr;
  int rc = resolve_enqueue(in, in_kind, sx, sy, sz, connectivity, delta, binary_image, periodic_boundary,
                           mem_space, stream, info, &S);
  if (rc) return rc;
  if (g_timing) {   // per-kernel timings need the two-phase path (events are collected per phase)
    rc = resolve_finish(S, (cudaStream_t)stream, info);
    if (rc) { cc3d_b200_session_release(S); return rc; }
    marks_collect(false);
    return cc3d_b200_label_write(S, out, out_kind, mem_space, stream);
  }
  cudaStream_t s = (cudaStream_t)stream;
  if (mem_space == CC3D_B200_DEVICE && S->voxels > 0 && S->hctr &&
      (out_kind == CC3D_B200_U16 || out_kind == CC3D_B200_U32 || out_kind == CC3D_B200_U64)) {
    // Device-resident call: the host only needs the counters, which are final BEFORE the expansion kernel runs. An
    // event marks their copy; D is enqueued behind it and the host returns as soon as the event has fired - the output
    // is stream-ordered like any other CUDA result, and the caller's next enqueue overlaps with D instead of waiting
    // for it (the workspace goes back to the cache behind an event of its own).
    thread_local cudaEvent_t ev = nullptr;
    if (!ev && cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) ev = nullptr;
    if (ev) {
      cudaEventRecord(ev, s);
      const i64 nrows = S->g.sy * S->g.sz;
      auto write = [&]() {
        if (out_kind == CC3D_B200_U16) launch_write<uint16_t>(S, (uint16_t*)out, 0, nrows, nullptr, 0, s);
        else if (out_kind == CC3D_B200_U32) launch_write<uint32_t>(S, (uint32_t*)out, 0, nrows, nullptr, 0, s);
        else launch_write<uint64_t>(S, (uint64_t*)out, 0, nrows, nullptr, 0, s);
      };
      write();
      cudaError_t e = cudaEventSynchronize(ev);
      if (e == cudaSuccess) e = cudaGetLastError();
      if (e != cudaSuccess) { session_release_after(S, s); return fail(CC3D_B200_ERR_CUDA, std::string("label: ") + cudaGetErrorString(e)); }
      const Counters* h = S->hctr;
      if (h->gq_ovf) {
        // rare: the edge queue overflowed - redo the unions on the global forest (synchronises), write again
        rc = resolve_finish(S, s, info);
        if (rc == 0) { write(); if (cudaStreamSynchronize(s) != cudaSuccess) rc = fail(CC3D_B200_ERR_CUDA, "label: redo failed"); }
        cc3d_b200_session_release(S);
        if (rc == 0 && out_kind == CC3D_B200_U16 && info->N > 0xFFFFull) rc = fail(CC3D_B200_ERR_OUT_RANGE, "N does not fit the requested output kind");
        return rc;
      }
      if (h->pad[0]) g_seen_big_tiles.store(true);
      S->N = h->N;
      info->N = h->N;
      info->epl = S->epl_is_runs ? h->nruns : h->epl;
      info->first_foreground_row = h->nruns ? (int64_t)~h->first_inv : -1;
      info->last_foreground_row = h->nruns ? (int64_t)h->last_p1 - 1 : -1;
      rc = (out_kind == CC3D_B200_U16 && info->N > 0xFFFFull) ? fail(CC3D_B200_ERR_OUT_RANGE, "N does not fit the requested output kind") : 0;
      // the copy into the pinned slot has completed (the event fired): the slot can be recycled now, the arena after D
      arena_release_after(S->arena, s);
      pinned_slot_give(S->hctr);
      delete S;
      return rc;
    }
  }
  rc = write_impl(S, out, out_kind, mem_space, stream, 0, S->g.sy * S->g.sz, nullptr, 0, 0, false, false);
  if (rc == 0) rc = resolve_finish(S, (cudaStream_t)stream, info);
  if (rc == 0 && S->redone)   // the edge queue overflowed: the labels written above predate the redo
    rc = write_impl(S, out, out_kind, mem_space, stream, 0, S->g.sy * S->g.sz, nullptr, 0, 0, false, false);
  if (rc == 0 && ((out_kind == CC3D_B200_U16 && info->N > 0xFFFFull)))
    rc = fail(CC3D_B200_ERR_OUT_RANGE, "N does not fit the requested output kind");
  cc3d_b200_session_release(S);
  return rc;
}

int cc3d_b200_label(const void* in, int in_kind, int64_t sx, int64_t sy, int64_t sz, int connectivity,
                    const void* delta, int binary_image, int periodic_boundary, void* out, int out_kind,
                    int mem_space, uint64_t* N, void* stream) {
  cc3d_b200_resolve_info info;
  cc3d_b200_session* S = nullptr;
  int rc = cc3d_b200_label_resolve(in, in_kind, sx, sy, sz, connectivity, delta, binary_image, periodic_boundary,
                                   mem_space, stream, &info, &S);
  if (rc) return rc;
  if (N) *N = info.N;
  return cc3d_b200_label_write(S, out, out_kind, mem_space, stream);
}

template <typename LT>
static int statistics_typed(const LT* labels, const Geom& g, u64 N, u32* counts, u32* bbox, u64* sums, cudaStream_t s,
                            unsigned long long* maxout = nullptr) {
  // (A second formulation - x-run records, voxel-parallel emission + record-parallel accumulation - was built and
  // measured this round: 0.58-0.77 ms at 512^3 against 0.9 ms here on the connectomics labelling but 5.3 ms against
  // 3.4 ms on 2048 x 2048 x 512 Voronoi labels, and 110 warp instructions per 32 voxels; it was removed again.
  // profiles/r02_statistics_ab.md has the numbers.)
  static PerDeviceOnce once;
  auto k = k_statistics<LT>;
  if (once.first()) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(StatTable));
  k_stat_init<<<(unsigned)((N + 1 + 255) / 256), 256, 0, s>>>(counts, bbox, (unsigned long long*)sums, N + 1);
  k<<<148 * 4, 256, sizeof(StatTable), s>>>(labels, g, N, counts, bbox, (unsigned long long*)sums, maxout);
  g_launches += 2;
  return 0;
}

int cc3d_b200_statistics(const void* labels, int kind, int64_t sx, int64_t sy, int64_t sz, uint64_t N,
                         uint32_t* counts, uint32_t* bbox, uint64_t* sums, int mem_space, void* stream) {
  if (int rc = check_shape(sx, sy, sz)) return rc;
  const size_t es = kind_size(kind);
  if (!es || kind > CC3D_B200_U64) return fail(CC3D_B200_ERR_KIND, "labels must be u8/u16/u32/u64");
  if (N >= 0xFFFFFFFEull) return fail(CC3D_B200_ERR_TOO_LARGE, "N must be < 2^32-2");
  const i64 voxels = sx * sy * sz;
  if (voxels == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  Geom g = make_geom(sx, sy, sz);
  Arena ar;
  const void* dl = labels;
  u32 *dc = counts, *db = bbox;
  u64* ds = sums;
  const size_t n1 = (size_t)N + 1;
  if (mem_space == CC3D_B200_HOST) {
    if (int rc = arena_acquire((size_t)voxels * es + n1 * (4 + 24 + 24) + 4096, &ar)) return rc;
    void* d = ar.take((size_t)voxels * es);
    dc = (u32*)ar.take(n1 * 4); db = (u32*)ar.take(n1 * 24); ds = (u64*)ar.take(n1 * 24);
    cudaError_t e = cudaMemcpyAsync(d, labels, (size_t)voxels * es, cudaMemcpyHostToDevice, s);
    if (e != cudaSuccess) { arena_release(ar); return fail(CC3D_B200_ERR_CUDA, cudaGetErrorString(e)); }
    dl = d;
  }
  switch (kind) {
    case CC3D_B200_U8: statistics_typed((const uint8_t*)dl, g, N, dc, db, ds, s); break;
    case CC3D_B200_U16: statistics_typed((const uint16_t*)dl, g, N, dc, db, ds, s); break;
    case CC3D_B200_U32: statistics_typed((const uint32_t*)dl, g, N, dc, db, ds, s); break;
    default: statistics_typed((const uint64_t*)dl, g, N, dc, db, ds, s); break;
  }
  cudaError_t e = cudaSuccess;
  if (mem_space == CC3D_B200_HOST) {
    e = cudaMemcpyAsync(counts, dc, n1 * 4, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(bbox, db, n1 * 24, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(sums, ds, n1 * 24, cudaMemcpyDeviceToHost, s);
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  if (e == cudaSuccess) e = cudaGetLastError();
  arena_release(ar);
  if (e != cudaSuccess) return fail(CC3D_B200_ERR_CUDA, std::string("statistics: ") + cudaGetErrorString(e));
  return 0;
}


// statistics without knowing the largest label beforehand (the reference takes np.max first, fastcc3d.pyx:713-720; a
// separate maximum pass would read the volume twice): ONE sweep accumulates the labels below `cap` into tables of
// `cap` entries and tracks the true maximum. *max_label < cap: the first *max_label + 1 entries of the tables are the
// statistics (host tables receive exactly those). Otherwise nothing useful was produced: call again with
// cap > *max_label.
int cc3d_b200_statistics_auto(const void* labels, int kind, int64_t sx, int64_t sy, int64_t sz, uint64_t cap,
                              uint64_t* max_label, uint32_t* counts, uint32_t* bbox, uint64_t* sums, int mem_space,
                              void* stream) {
  if (int rc = check_shape(sx, sy, sz)) return rc;
  const size_t es = kind_size(kind);
  if (!es || kind > CC3D_B200_U64) return fail(CC3D_B200_ERR_KIND, "labels must be u8/u16/u32/u64");
  if (!max_label || cap == 0 || cap >= 0xFFFFFFFEull) return fail(CC3D_B200_ERR_ARGUMENT, "statistics_auto: bad max_label / cap");
  *max_label = 0;
  const i64 voxels = sx * sy * sz;
  if (voxels == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  Geom g = make_geom(sx, sy, sz);
  Arena ar;
  const bool host = mem_space == CC3D_B200_HOST;
  const size_t n1 = (size_t)cap;
  if (int rc = arena_acquire((host ? (size_t)voxels * es + n1 * (4 + 24 + 24) : 0) + 4096, &ar, s, true)) return rc;
  const void* dl = labels;
  u32 *dc = counts, *db = bbox;
  u64* ds = sums;
  if (host) {
    void* d = ar.take((size_t)voxels * es);
    dc = (u32*)ar.take(n1 * 4); db = (u32*)ar.take(n1 * 24); ds = (u64*)ar.take(n1 * 24);
    cudaError_t e = cudaMemcpyAsync(d, labels, (size_t)voxels * es, cudaMemcpyHostToDevice, s);
    if (e != cudaSuccess) { arena_release(ar); return fail(CC3D_B200_ERR_CUDA, cudaGetErrorString(e)); }
    dl = d;
  }
  unsigned long long* dmax = (unsigned long long*)ar.take(8);
  cudaMemsetAsync(dmax, 0, 8, s);
  switch (kind) {
    case CC3D_B200_U8: statistics_typed((const uint8_t*)dl, g, cap - 1, dc, db, ds, s, dmax); break;
    case CC3D_B200_U16: statistics_typed((const uint16_t*)dl, g, cap - 1, dc, db, ds, s, dmax); break;
    case CC3D_B200_U32: statistics_typed((const uint32_t*)dl, g, cap - 1, dc, db, ds, s, dmax); break;
    default: statistics_typed((const uint64_t*)dl, g, cap - 1, dc, db, ds, s, dmax); break;
  }
  unsigned long long hmax = 0;
  cudaError_t e = cudaMemcpyAsync(&hmax, dmax, 8, cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  if (e == cudaSuccess && host && hmax < cap) {
    const size_t m1 = (size_t)hmax + 1;
    e = cudaMemcpyAsync(counts, dc, m1 * 4, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(bbox, db, m1 * 24, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(sums, ds, m1 * 24, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  }
  if (e == cudaSuccess) e = cudaGetLastError();
  arena_release(ar);
  if (e != cudaSuccess) return fail(CC3D_B200_ERR_CUDA, std::string("statistics_auto: ") + cudaGetErrorString(e));
  *max_label = hmax;
  return 0;
}

template <typename IT>
static void mask_typed(IT* img, const void* labels, int label_kind, const unsigned char* keep, u64 N, i64 n, cudaStream_t s) {
  const unsigned blocks = (unsigned)std::min<i64>((n + 255) / 256, 148 * 32);
  switch (label_kind) {
    case CC3D_B200_U8: k_mask_by_label<IT, uint8_t><<<blocks, 256, 0, s>>>(img, (const uint8_t*)labels, keep, N, n); break;
    case CC3D_B200_U16: k_mask_by_label<IT, uint16_t><<<blocks, 256, 0, s>>>(img, (const uint16_t*)labels, keep, N, n); break;
    case CC3D_B200_U32: k_mask_by_label<IT, uint32_t><<<blocks, 256, 0, s>>>(img, (const uint32_t*)labels, keep, N, n); break;
    default: k_mask_by_label<IT, uint64_t><<<blocks, 256, 0, s>>>(img, (const uint64_t*)labels, keep, N, n); break;
  }
}

int cc3d_b200_mask_by_label(void* img, int img_itemsize, const void* labels, int label_kind, int64_t voxels,
                            const uint8_t* keep, uint64_t N, int mem_space, void* stream) {
  const size_t ls = kind_size(label_kind);
  if (!ls || label_kind > CC3D_B200_U64) return fail(CC3D_B200_ERR_KIND, "labels must be u8/u16/u32/u64");
  if (img_itemsize != 1 && img_itemsize != 2 && img_itemsize != 4 && img_itemsize != 8)
    return fail(CC3D_B200_ERR_KIND, "img itemsize must be 1, 2, 4 or 8");
  if (voxels <= 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  Arena ar;
  void* dimg = img; const void* dl = labels; const unsigned char* dk = keep;
  if (mem_space == CC3D_B200_HOST) {
    if (int rc = arena_acquire((size_t)voxels * (img_itemsize + ls) + N + 1 + 4096, &ar)) return rc;
    dimg = ar.take((size_t)voxels * img_itemsize);
    void* l = ar.take((size_t)voxels * ls);
    unsigned char* k = (unsigned char*)ar.take(N + 1);
    cudaMemcpyAsync(dimg, img, (size_t)voxels * img_itemsize, cudaMemcpyHostToDevice, s);
    cudaMemcpyAsync(l, labels, (size_t)voxels * ls, cudaMemcpyHostToDevice, s);
    cudaMemcpyAsync(k, keep, N + 1, cudaMemcpyHostToDevice, s);
    dl = l; dk = k;
  }
  switch (img_itemsize) {
    case 1: mask_typed((uint8_t*)dimg, dl, label_kind, dk, N, voxels, s); break;
    case 2: mask_typed((uint16_t*)dimg, dl, label_kind, dk, N, voxels, s); break;
    case 4: mask_typed((uint32_t*)dimg, dl, label_kind, dk, N, voxels, s); break;
    default: mask_typed((uint64_t*)dimg, dl, label_kind, dk, N, voxels, s); break;
  }
  g_launches += 1;
  cudaError_t e = cudaSuccess;
  if (mem_space == CC3D_B200_HOST) e = cudaMemcpyAsync(img, dimg, (size_t)voxels * img_itemsize, cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  if (e == cudaSuccess) e = cudaGetLastError();
  arena_release(ar);
  if (e != cudaSuccess) return fail(CC3D_B200_ERR_CUDA, std::string("mask_by_label: ") + cudaGetErrorString(e));
  return 0;
}


// ---- sharded volumes: host side of the merge (no GPU work) ----
// Slab r has local labels 1..n_labels[r]; pairs[r] holds n_pairs[r] packed equivalences
// (label in slab r-1) << 32 | (label in slab r) across the interface below slab r (pairs[0] is unused).
// A component is owned by the lowest slab it touches; owned components are numbered slab by slab in local
// label order, which is the first-appearance order of the whole volume (cc3d/__init__.py:296-321, 425-468
// do the same with a Python DisjointSet + renumber). Writes remap[0..n_labels[rank]] for slab `rank`
// (remap[0] = 0) and the global component count.
// ---- callers either side of the labelling path ----
template <typename T>
static int voxel_graph_typed(const T* in, void* graph, i64 sx, i64 sy, i64 sz, int connectivity, cudaStream_t s) {
  const i64 voxels = sx * sy * sz;
  const unsigned blocks = (unsigned)std::min<i64>((voxels + 255) / 256, 148 * 64);
  switch (connectivity) {
    case 4: k_voxel_graph<T, uint8_t, 4, true><<<blocks, 256, 0, s>>>(in, (uint8_t*)graph, sx, sy, sz); break;
    case 8: k_voxel_graph<T, uint8_t, 8, true><<<blocks, 256, 0, s>>>(in, (uint8_t*)graph, sx, sy, sz); break;
    case 6: k_voxel_graph<T, uint8_t, 6, false><<<blocks, 256, 0, s>>>(in, (uint8_t*)graph, sx, sy, sz); break;
    case 18: k_voxel_graph<T, uint32_t, 18, false><<<blocks, 256, 0, s>>>(in, (uint32_t*)graph, sx, sy, sz); break;
    case 26: k_voxel_graph<T, uint32_t, 26, false><<<blocks, 256, 0, s>>>(in, (uint32_t*)graph, sx, sy, sz); break;
    default: return -1;
  }
  g_launches += 1;
  return 0;
}

int cc3d_b200_voxel_connectivity_graph(const void* labels, int kind, int64_t sx, int64_t sy, int64_t sz, int connectivity,
                                       void* graph, int mem_space, void* stream) {
  if (int rc = check_shape(sx, sy, sz)) return rc;
  const size_t es = kind_size(kind);
  if (!es || kind > CC3D_B200_U64) return fail(CC3D_B200_ERR_KIND, "labels must be u8/u16/u32/u64");
  if (connectivity != 4 && connectivity != 8 && connectivity != 6 && connectivity != 18 && connectivity != 26)
    return fail(CC3D_B200_ERR_CONNECTIVITY, "Only 4 and 8 2D and 6, 18, and 26 3D connectivities are supported.");
  if ((connectivity == 4 || connectivity == 8) && sz != 1)
    return fail(CC3D_B200_ERR_2D_NEEDS_SZ1, "sz must be 1 for 2D connectivities.");
  const i64 voxels = sx * sy * sz;
  if (voxels == 0) return 0;
  const size_t os = (connectivity == 18 || connectivity == 26) ? 4 : 1;
  cudaStream_t s = (cudaStream_t)stream;
  Arena ar;
  const void* din = labels; void* dg = graph;
  if (mem_space == CC3D_B200_HOST) {
    if (int rc = arena_acquire((size_t)voxels * (es + os) + 4096, &ar)) return rc;
    void* d = ar.take((size_t)voxels * es);
    dg = ar.take((size_t)voxels * os);
    cudaMemcpyAsync(d, labels, (size_t)voxels * es, cudaMemcpyHostToDevice, s);
    din = d;
  }
  int rc = -1;
  switch (kind) {
    case CC3D_B200_U8: rc = voxel_graph_typed((const uint8_t*)din, dg, sx, sy, sz, connectivity, s); break;
    case CC3D_B200_U16: rc = voxel_graph_typed((const uint16_t*)din, dg, sx, sy, sz, connectivity, s); break;
    case CC3D_B200_U32: rc = voxel_graph_typed((const uint32_t*)din, dg, sx, sy, sz, connectivity, s); break;
    case CC3D_B200_U64: rc = voxel_graph_typed((const uint64_t*)din, dg, sx, sy, sz, connectivity, s); break;
  }
  cudaError_t e = cudaSuccess;
  if (rc == 0 && mem_space == CC3D_B200_HOST) e = cudaMemcpyAsync(graph, dg, (size_t)voxels * os, cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  if (e == cudaSuccess) e = cudaGetLastError();
  arena_release(ar);
  if (rc) return fail(CC3D_B200_ERR_ARGUMENT, "voxel_connectivity_graph: no kernel for this configuration");
  if (e != cudaSuccess) return fail(CC3D_B200_ERR_CUDA, std::string("voxel_connectivity_graph: ") + cudaGetErrorString(e));
  return 0;
}

__global__ void k_set_u64(u64* p, u64 v) { *p = v; }

int cc3d_b200_color_connectivity_graph(const void* vcg, int vcg_kind, int64_t sx, int64_t sy, int64_t sz, int connectivity,
                                       uint32_t* out, uint64_t* N, int mem_space, void* stream) {
  if (int rc = check_shape(sx, sy, sz)) return rc;
  if (vcg_kind != CC3D_B200_U8 && vcg_kind != CC3D_B200_U32) return fail(CC3D_B200_ERR_KIND, "Only uint8 and uint32 are supported.");
  if (connectivity != 4 && connectivity != 8 && connectivity != 6 && connectivity != 26)
    return fail(CC3D_B200_ERR_CONNECTIVITY, "Only 4, 8, 6 and 26 connectivities are supported.");
  if (sz > 1 && connectivity != 6 && connectivity != 26)
    return fail(CC3D_B200_ERR_CONNECTIVITY, "Only 6 and 26 connectivity is supported in 3D.");
  if (sz > 1 && connectivity == 26 && vcg_kind != CC3D_B200_U32)
    return fail(CC3D_B200_ERR_KIND, "26-connectivity requires a 32-bit voxel graph.");
  if (N) *N = 0;
  const i64 voxels = sx * sy * sz;
  if (voxels == 0) return 0;
  if ((u64)voxels >= 0xFFFFFFFFull) return fail(CC3D_B200_ERR_TOO_LARGE, "graph has >= 2^32-1 voxels");
  // backward directions the reference follows (cc3d_graphs.hpp:584-1106); 2D graphs come in two bit layouts
  VcgDirs D;
  D.n = 0;
  auto add = [&](int dx, int dy, int dz, int bit_number) {
    D.d[D.n][0] = (signed char)dx; D.d[D.n][1] = (signed char)dy; D.d[D.n][2] = (signed char)dz;
    D.mask[D.n] = 1u << (bit_number - 1); D.n++;
  };
  add(-1, 0, 0, 2); add(0, -1, 0, 4);
  if (sz == 1) {
    if (connectivity == 8 || connectivity == 26) {
      if (vcg_kind == CC3D_B200_U8) { add(-1, -1, 0, 8); add(1, -1, 0, 7); }
      else { add(-1, -1, 0, 10); add(1, -1, 0, 9); }
    }
  } else {
    add(0, 0, -1, 6);
    if (connectivity == 26) {
      add(-1, -1, 0, 10); add(1, -1, 0, 9);
      add(-1, 0, -1, 16); add(1, 0, -1, 15); add(0, -1, -1, 18); add(0, 1, -1, 17);
      add(-1, -1, -1, 26); add(1, -1, -1, 25); add(-1, 1, -1, 24); add(1, 1, -1, 23);
    }
  }
  cudaStream_t s = (cudaStream_t)stream;
  const size_t vs = kind_size(vcg_kind);
  const i64 nwords2 = (voxels + 31) / 32;
  const i64 nb2 = (nwords2 + CC_SCAN_CHUNK - 1) / CC_SCAN_CHUNK;
  size_t need = 4096 + (size_t)nwords2 * 12 + (size_t)(nb2 + 1) * 8 + 1024 + 4 * 256;
  if (mem_space == CC3D_B200_HOST) need += (size_t)voxels * (vs + 4) + 512;
  Arena ar;
  if (int rc = arena_acquire(need, &ar, s, true)) return rc;
  const void* dv = vcg; u32* dout = out;
  if (mem_space == CC3D_B200_HOST) {
    void* d = ar.take((size_t)voxels * vs);
    dout = (u32*)ar.take((size_t)voxels * 4);
    cudaMemcpyAsync(d, vcg, (size_t)voxels * vs, cudaMemcpyHostToDevice, s);
    dv = d;
  }
  u32* GR = (u32*)ar.take((size_t)nwords2 * 4);
  u32* cnt = (u32*)ar.take((size_t)nwords2 * 4);
  u32* prefix = (u32*)ar.take((size_t)nwords2 * 4);
  u64* bsum = (u64*)ar.take((size_t)(nb2 + 1) * 8);
  u64* nvox = (u64*)ar.take(8);
  u64* ntot = (u64*)ar.take(8);
  const unsigned blocks = (unsigned)std::min<i64>((voxels + 255) / 256, 148 * 64);
  k_set_u64<<<1, 1, 0, s>>>(nvox, (u64)voxels);
  k_iota<<<(unsigned)((voxels + 255) / 256), 256, 0, s>>>(dout, voxels);
  if (vcg_kind == CC3D_B200_U8) k_vcg_union<uint8_t><<<blocks, 256, 0, s>>>((const uint8_t*)dv, dout, sx, sy, sz, D);
  else k_vcg_union<uint32_t><<<blocks, 256, 0, s>>>((const uint32_t*)dv, dout, sx, sy, sz, D);
  k_compress<<<CC_GRID_BLOCKS, 256, 0, s>>>(dout, GR, cnt, nvox);
  scan_counts(cnt, prefix, bsum, nwords2, nullptr, 0, ntot, nullptr, s);
  k_assign<<<CC_GRID_BLOCKS, 256, 0, s>>>(dout, GR, prefix, nvox);
  g_launches += 5;
  u64 hN = 0;
  cudaError_t e = cudaMemcpyAsync(&hN, ntot, 8, cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess && mem_space == CC3D_B200_HOST) e = cudaMemcpyAsync(out, dout, (size_t)voxels * 4, cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  if (e == cudaSuccess) e = cudaGetLastError();
  arena_release(ar);
  if (e != cudaSuccess) return fail(CC3D_B200_ERR_CUDA, std::string("color_connectivity_graph: ") + cudaGetErrorString(e));
  if (N) *N = hN;
  return 0;
}


// ---- crackle v0 decode on the device: crack codes -> 4-bit pixel graph (k_crackle_cuts),
// coloured slice by slice as ONE 6-connected graph without z links (k_vcg_union + the rank kernels: components are
// numbered in raster order, i.e. slice by slice, which is the order of the file's key table), then out[i] =
// lut[component of i]. Replaces the per-slice cc3d.color_connectivity_graph calls of a crackle decoder
// (cc3d_graphs.hpp:1018-1074 is the colouring it would call). ----
int cc3d_b200_crackle_v0_decode(const uint8_t* stream, const uint64_t* slice_off, int64_t sx, int64_t sy, int64_t sz,
                                const uint32_t* lut, uint64_t n_lut, uint32_t* out, uint64_t* n_components, int mem_space,
                                void* cuda_stream) {
  if (int rc = check_shape(sx, sy, sz)) return rc;
  if (!stream || !slice_off || !lut || !out) return fail(CC3D_B200_ERR_ARGUMENT, "crackle_v0_decode: NULL argument");
  const i64 voxels = sx * sy * sz;
  if (n_components) *n_components = 0;
  if (voxels == 0) return 0;
  if ((u64)voxels >= 0xFFFFFFFFull) return fail(CC3D_B200_ERR_TOO_LARGE, "crackle_v0_decode: >= 2^32-1 voxels");
  if (sx >= 65536 || sy >= 65536 || ((sx * sy) & 3)) return fail(CC3D_B200_ERR_ARGUMENT, "crackle_v0_decode: slices must be < 65536 wide / high and sx * sy a multiple of 4");
  if (mem_space != CC3D_B200_HOST) return fail(CC3D_B200_ERR_ARGUMENT, "crackle_v0_decode: the code stream is parsed from host memory (the labels may stay on the device: see out)");
  cudaStream_t s = (cudaStream_t)cuda_stream;
  const size_t total = (size_t)slice_off[sz];
  const i64 nwords2 = (voxels + 31) / 32;
  const i64 nb2 = (nwords2 + CC_SCAN_CHUNK - 1) / CC_SCAN_CHUNK;
  cudaPointerAttributes pa;
  const bool out_on_device = cudaPointerGetAttributes(&pa, out) == cudaSuccess && pa.type == cudaMemoryTypeDevice;
  cudaGetLastError();
  size_t need = 8192 + total + 256 + (size_t)(sz + 1) * 8 + 256 + (size_t)voxels + 256 + 8 * total + 256 + (size_t)n_lut * 4 + 256 +
                (size_t)nwords2 * 12 + 1024 + (size_t)(nb2 + 2) * 8 + 256 + (out_on_device ? 0 : (size_t)voxels * 4 + 256);
  Arena ar;
  if (int rc = arena_acquire(need, &ar, s, true)) return rc;
  unsigned char* dstream = (unsigned char*)ar.take(total + 16);
  unsigned long long* doff = (unsigned long long*)ar.take((size_t)(sz + 1) * 8);
  unsigned char* vcg = (unsigned char*)ar.take((size_t)voxels);
  u32* stack = (u32*)ar.take(8 * total + 16);
  u32* dlut = (u32*)ar.take((size_t)n_lut * 4);
  u32* GR = (u32*)ar.take((size_t)nwords2 * 4);
  u32* cnt = (u32*)ar.take((size_t)nwords2 * 4);
  u32* prefix = (u32*)ar.take((size_t)nwords2 * 4);
  u64* bsum = (u64*)ar.take((size_t)(nb2 + 2) * 8);
  u64* small = (u64*)ar.take(64);          // [0] voxels, [1] N, [2] error flags
  u32* dout = out_on_device ? out : (u32*)ar.take((size_t)voxels * 4);
  cudaMemcpyAsync(dstream, stream, total, cudaMemcpyHostToDevice, s);
  cudaMemcpyAsync(doff, slice_off, (size_t)(sz + 1) * 8, cudaMemcpyHostToDevice, s);
  cudaMemcpyAsync(dlut, lut, (size_t)n_lut * 4, cudaMemcpyHostToDevice, s);
  cudaMemsetAsync(vcg, 0x0F, (size_t)voxels, s);
  cudaMemsetAsync(small, 0, 64, s);
  k_set_u64<<<1, 1, 0, s>>>(small, (u64)voxels);
  k_crackle_cuts<<<(unsigned)((sz + 31) / 32), 32, 0, s>>>(dstream, doff, (int)sx, (int)sy, (int)sz, vcg, stack, (u32*)(small + 2));
  VcgDirs D;
  D.n = 0;
  auto add = [&](int dx, int dy, int dz, int bit_number) {
    D.d[D.n][0] = (signed char)dx; D.d[D.n][1] = (signed char)dy; D.d[D.n][2] = (signed char)dz;
    D.mask[D.n] = 1u << (bit_number - 1); D.n++;
  };
  add(-1, 0, 0, 2); add(0, -1, 0, 4);      // the slices are independent images: no z links
  const unsigned blocks = (unsigned)std::min<i64>((voxels + 255) / 256, 148 * 64);
  k_iota<<<(unsigned)((voxels + 255) / 256), 256, 0, s>>>(dout, voxels);
  k_vcg_union<uint8_t><<<blocks, 256, 0, s>>>(vcg, dout, sx, sy, sz, D);
  k_compress<<<CC_GRID_BLOCKS, 256, 0, s>>>(dout, GR, cnt, small);
  scan_counts(cnt, prefix, bsum, nwords2, nullptr, 0, small + 1, nullptr, s);
  k_assign<<<CC_GRID_BLOCKS, 256, 0, s>>>(dout, GR, prefix, small);
  k_remap_labels<u32, u32><<<(unsigned)std::min<i64>((voxels + 255) / 256, 148 * 32), 256, 0, s>>>(dout, dlut, n_lut ? n_lut - 1 : 0, dout, voxels);
  g_launches += 7;
  u64 h[3] = {0, 0, 0};
  cudaError_t e = cudaMemcpyAsync(h, small, 24, cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess && !out_on_device) e = cudaMemcpyAsync(out, dout, (size_t)voxels * 4, cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  if (e == cudaSuccess) e = cudaGetLastError();
  arena_release(ar);
  if (e != cudaSuccess) return fail(CC3D_B200_ERR_CUDA, std::string("crackle_v0_decode: ") + cudaGetErrorString(e));
  if (h[2]) return fail(CC3D_B200_ERR_ARGUMENT, "crackle_v0_decode: malformed crack code stream (flags " + std::to_string(h[2]) + ")");
  if (h[1] + 1 > n_lut) return fail(CC3D_B200_ERR_ARGUMENT, "crackle_v0_decode: more components than keys");
  if (n_components) *n_components = h[1];
  return 0;
}

template <typename LT>
static void remap_typed(const LT* labels, const u32* table, u64 N, void* out, int out_kind, i64 n, cudaStream_t s) {
  const unsigned blocks = (unsigned)std::min<i64>((n + 255) / 256, 148 * 32);
  switch (out_kind) {
    case CC3D_B200_U8: k_remap_labels<LT, uint8_t><<<blocks, 256, 0, s>>>(labels, table, N, (uint8_t*)out, n); break;
    case CC3D_B200_U16: k_remap_labels<LT, uint16_t><<<blocks, 256, 0, s>>>(labels, table, N, (uint16_t*)out, n); break;
    case CC3D_B200_U32: k_remap_labels<LT, uint32_t><<<blocks, 256, 0, s>>>(labels, table, N, (uint32_t*)out, n); break;
    default: k_remap_labels<LT, uint64_t><<<blocks, 256, 0, s>>>(labels, table, N, (uint64_t*)out, n); break;
  }
}

int cc3d_b200_remap_labels(const void* labels, int label_kind, int64_t voxels, const uint32_t* table, uint64_t N,
                           void* out, int out_kind, int mem_space, void* stream) {
  const size_t ls = kind_size(label_kind), os = kind_size(out_kind);
  if (!ls || label_kind > CC3D_B200_U64) return fail(CC3D_B200_ERR_KIND, "labels must be u8/u16/u32/u64");
  if (!os || out_kind > CC3D_B200_U64) return fail(CC3D_B200_ERR_KIND, "out must be u8/u16/u32/u64");
  if (voxels <= 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  Arena ar;
  const void* dl = labels; const u32* dt = table; void* dout = out;
  if (mem_space == CC3D_B200_HOST) {
    if (int rc = arena_acquire((size_t)voxels * (ls + os) + (N + 1) * 4 + 4096, &ar)) return rc;
    void* l = ar.take((size_t)voxels * ls);
    dout = ar.take((size_t)voxels * os);
    u32* t = (u32*)ar.take((N + 1) * 4);
    cudaMemcpyAsync(l, labels, (size_t)voxels * ls, cudaMemcpyHostToDevice, s);
    cudaMemcpyAsync(t, table, (N + 1) * 4, cudaMemcpyHostToDevice, s);
    dl = l; dt = t;
  }
  switch (label_kind) {
    case CC3D_B200_U8: remap_typed((const uint8_t*)dl, dt, N, dout, out_kind, voxels, s); break;
    case CC3D_B200_U16: remap_typed((const uint16_t*)dl, dt, N, dout, out_kind, voxels, s); break;
    case CC3D_B200_U32: remap_typed((const uint32_t*)dl, dt, N, dout, out_kind, voxels, s); break;
    default: remap_typed((const uint64_t*)dl, dt, N, dout, out_kind, voxels, s); break;
  }
  g_launches += 1;
  cudaError_t e = cudaSuccess;
  if (mem_space == CC3D_B200_HOST) e = cudaMemcpyAsync(out, dout, (size_t)voxels * os, cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  if (e == cudaSuccess) e = cudaGetLastError();
  arena_release(ar);
  if (e != cudaSuccess) return fail(CC3D_B200_ERR_CUDA, std::string("remap_labels: ") + cudaGetErrorString(e));
  return 0;
}

int cc3d_b200_contacts(const void* labels, int kind, int64_t sx, int64_t sy, int64_t sz, int connectivity,
                       uint64_t* pairs, uint32_t* class_counts, uint64_t capacity, uint64_t* count, int mem_space,
                       void* stream) {
  if (int rc = check_shape(sx, sy, sz)) return rc;
  const size_t es = kind_size(kind);
  if (!es || kind > CC3D_B200_U64) return fail(CC3D_B200_ERR_KIND, "labels must be u8/u16/u32/u64");
  if (connectivity != 4 && connectivity != 8 && connectivity != 6 && connectivity != 18 && connectivity != 26)
    return fail(CC3D_B200_ERR_CONNECTIVITY, "Only (2d) 4, 8, or (3d) 6, 18, and 26 connectivities are supported.");
  if ((connectivity == 4 || connectivity == 8) && sz != 1)
    return fail(CC3D_B200_ERR_2D_NEEDS_SZ1, "z thickness must be 1 for 2d region graph extraction.");
  if (!count) return fail(CC3D_B200_ERR_ARGUMENT, "count must not be NULL");
  *count = 0;
  const i64 voxels = sx * sy * sz;
  if (voxels == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  // hash table: grows (x8) until the pairs fit
  u32 slots = 1u << 16;
  while ((u64)slots < 4 * capacity && slots < (1u << 28)) slots <<= 1;
  while (true) {
    Arena ar;
    size_t need = 4096 + (size_t)slots * (8 + 16) + 512 + capacity * (8 + 16) + 1024;
    if (mem_space == CC3D_B200_HOST) need += (size_t)voxels * es + 512;
    if (int rc = arena_acquire(need, &ar, s, true)) return rc;
    const void* din = labels;
    if (mem_space == CC3D_B200_HOST) {
      void* d = ar.take((size_t)voxels * es);
      cudaMemcpyAsync(d, labels, (size_t)voxels * es, cudaMemcpyHostToDevice, s);
      din = d;
    }
    ContactTable tb;
    tb.keys = (unsigned long long*)ar.take((size_t)slots * 8);
    tb.vals = (u32*)ar.take((size_t)slots * 16);
    tb.flags = (u32*)ar.take(64);
    tb.mask = slots - 1;
    unsigned long long* dcount = (unsigned long long*)(tb.flags + 4);
    unsigned long long* dkeys = (unsigned long long*)ar.take(capacity * 8 + 16);
    u32* dvals = (u32*)ar.take(capacity * 16 + 16);
    cudaMemsetAsync(tb.keys, 0, (size_t)slots * 8, s);
    cudaMemsetAsync(tb.vals, 0, (size_t)slots * 16, s);
    cudaMemsetAsync(tb.flags, 0, 64, s);
    const unsigned blocks = (unsigned)std::min<i64>((voxels + 255) / 256, 148 * 16);
    switch (kind) {
      case CC3D_B200_U8: k_contacts<uint8_t><<<blocks, 256, 0, s>>>((const uint8_t*)din, sx, sy, sz, connectivity, tb); break;
      case CC3D_B200_U16: k_contacts<uint16_t><<<blocks, 256, 0, s>>>((const uint16_t*)din, sx, sy, sz, connectivity, tb); break;
      case CC3D_B200_U32: k_contacts<uint32_t><<<blocks, 256, 0, s>>>((const uint32_t*)din, sx, sy, sz, connectivity, tb); break;
      default: k_contacts<uint64_t><<<blocks, 256, 0, s>>>((const uint64_t*)din, sx, sy, sz, connectivity, tb); break;
    }
    k_contacts_compact<<<(slots + 255) / 256, 256, 0, s>>>(tb, dkeys, dvals, dcount, capacity);
    g_launches += 2;
    u32 hflags[4] = {0, 0, 0, 0};
    unsigned long long hcount = 0;
    cudaError_t e = cudaMemcpyAsync(hflags, tb.flags, 16, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(&hcount, dcount, 8, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) { arena_release(ar); return fail(CC3D_B200_ERR_CUDA, std::string("contacts: ") + cudaGetErrorString(e)); }
    if (hflags[1]) { arena_release(ar); return fail(CC3D_B200_ERR_TOO_LARGE, "contacts: label values must be < 2^32"); }
    if (hflags[0] && slots < (1u << 28)) { arena_release(ar); slots <<= 3; continue; }   // table full: larger table
    if (hflags[0]) { arena_release(ar); return fail(CC3D_B200_ERR_TOO_LARGE, "contacts: too many distinct pairs"); }
    *count = hcount;
    if (hcount <= capacity && hcount > 0) {
      const cudaMemcpyKind k = mem_space == CC3D_B200_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
      cudaMemcpyAsync(pairs, dkeys, hcount * 8, k, s);
      cudaMemcpyAsync(class_counts, dvals, hcount * 16, k, s);
      e = cudaStreamSynchronize(s);
    }
    arena_release(ar);
    if (e != cudaSuccess) return fail(CC3D_B200_ERR_CUDA, std::string("contacts: ") + cudaGetErrorString(e));
    return 0;
  }
}

template <typename IT>
static void expand_mask_typed(const cc3d_b200_session* S, const IT* img, IT* out, const unsigned char* keep, cudaStream_t s) {
  const Geom& g = S->g;
  const unsigned nchunks = (unsigned)((g.W + 31) / 32);
  const i64 nwarps = g.sy * g.sz * nchunks;
  k_expand_mask<IT><<<(unsigned)((nwarps + 7) / 8), 256, 0, s>>>(S->L, S->M, img, out, g, nchunks, (u32)nwarps, keep, S->lmask);
}

int cc3d_b200_dust(const void* img, void* out, int kind, int64_t sx, int64_t sy, int64_t sz, int connectivity,
                   int binary_image, int64_t lo, int64_t hi, int invert, int mem_space, uint64_t* N,
                   uint64_t* n_masked, void* stream) {
  if (kind < CC3D_B200_U8 || kind > CC3D_B200_U64) return fail(CC3D_B200_ERR_KIND, "dust: img must be u8/u16/u32/u64");
  if (!out || !N || !n_masked) return fail(CC3D_B200_ERR_ARGUMENT, "dust: out / N / n_masked must not be NULL");
  *N = 0; *n_masked = 0;
  const size_t es = kind_size(kind);
  const i64 voxels = sx * sy * sz;
  if (voxels == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  unsigned char zero_delta[8] = {0};
  cc3d_b200_resolve_info info;
  cc3d_b200_session* S = nullptr;
  int rc = resolve_enqueue(img, kind, sx, sy, sz, connectivity, zero_delta, binary_image, 0, mem_space, stream, &info, &S);
  if (rc) return rc;
  rc = resolve_finish(S, s, &info);          // N decides the size of the count / keep tables
  if (rc) { cc3d_b200_session_release(S); return rc; }
  marks_collect(false);
  const u64 n = info.N;
  Arena ar;
  if (int rc2 = arena_acquire((size_t)(n + 1) * 5 + 4096, &ar, s, true)) { cc3d_b200_session_release(S); return rc2; }
  u32* counts = (u32*)ar.take((size_t)(n + 1) * 4);
  unsigned char* keep = (unsigned char*)ar.take((size_t)(n + 1));
  unsigned long long* dmasked = (unsigned long long*)ar.take(8);
  cudaMemsetAsync(counts, 0, (size_t)(n + 1) * 4, s);
  cudaMemsetAsync(dmasked, 0, 8, s);
  ensure_run_labels(S, s);      // block-path sessions: the run table's labels are filled in on demand
  k_run_counts<<<148 * 4, 256, 0, s>>>(S->L, S->M, S->g, counts, S->lmask);
  k_dust_keep<<<(unsigned)((n + 1 + 255) / 256), 256, 0, s>>>(counts, keep, n, (long long)lo, (long long)hi, invert, dmasked);
  // host images are masked in place in their staged copy; device images go straight to `out` (which may be `img`)
  void* dout = mem_space == CC3D_B200_HOST ? const_cast<void*>(S->din) : out;
  switch (kind) {
    case CC3D_B200_U8: expand_mask_typed(S, (const uint8_t*)S->din, (uint8_t*)dout, keep, s); break;
    case CC3D_B200_U16: expand_mask_typed(S, (const uint16_t*)S->din, (uint16_t*)dout, keep, s); break;
    case CC3D_B200_U32: expand_mask_typed(S, (const uint32_t*)S->din, (uint32_t*)dout, keep, s); break;
    default: expand_mask_typed(S, (const uint64_t*)S->din, (uint64_t*)dout, keep, s); break;
  }
  g_launches += 3;
  unsigned long long hmasked = 0;
  cudaError_t e = cudaMemcpyAsync(&hmasked, dmasked, 8, cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess && mem_space == CC3D_B200_HOST) e = cudaMemcpyAsync(out, dout, (size_t)voxels * es, cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  if (e == cudaSuccess) e = cudaGetLastError();
  arena_release(ar);
  cc3d_b200_session_release(S);
  if (e != cudaSuccess) return fail(CC3D_B200_ERR_CUDA, std::string("dust: ") + cudaGetErrorString(e));
  *N = n;
  *n_masked = hmasked;
  return 0;
}

int cc3d_b200_merge_slabs(int world, const int64_t* n_labels, const uint64_t* const* pairs, const int64_t* n_pairs,
                          int rank, int64_t* remap, int64_t* n_total) {
  if (world <= 0 || rank < 0 || rank >= world || !n_labels || !remap || !n_total)
    return fail(CC3D_B200_ERR_ARGUMENT, "merge_slabs: bad arguments");
  // global id of (slab r, local label l >= 1) = off[r] + l: id order = first-appearance order of the whole volume
  std::vector<i64> off(world + 1, 0);
  for (int r = 0; r < world; r++) off[r + 1] = off[r] + n_labels[r];
  const i64 total = off[world];
  if (total >= 0xFFFFFFFFll) return fail(CC3D_B200_ERR_TOO_LARGE, "merge_slabs: more than 2^32-2 slab labels");
  // union-find over the ids directly (no sorting, duplicates are harmless); root = smallest id of the set.
  // The buffers are thread-local vectors; the loops work on raw pointers (a thread_local access inside a shared
  // library is a __tls_get_addr call, which the per-element form paid several times per pair).
  static thread_local std::vector<u32> parent_tl;
  static thread_local std::vector<i64> final_tl;
  parent_tl.resize((size_t)total + 1);
  u32* const parent = parent_tl.data();
  for (i64 i = 0; i <= total; i++) parent[i] = (u32)i;
  auto find = [parent](u32 i) { while (parent[i] != i) { parent[i] = parent[parent[i]]; i = parent[i]; } return i; };
  for (int r = 1; r < world; r++) {
    const u64* pr = pairs[r];
    const i64 nlo = n_labels[r - 1], nup = n_labels[r];
    const i64 off_lo = off[r - 1], off_up = off[r];
    // the face kernel reports a pair once per touching voxel pair (a few hundred distinct pairs among thousands):
    // a small direct-mapped cache of the pairs already united skips the repeats without touching the forest
    constexpr int CACHE_BITS = 12;
    u64 seen[1 << CACHE_BITS];
    for (int i = 0; i < (1 << CACHE_BITS); i++) seen[i] = ~0ull;   // lo = up = 2^32-1 is never a valid pair
    for (i64 k = 0; k < n_pairs[r]; k++) {
      const u64 v = pr[k];
      u64& slot = seen[(v * 0x9E3779B97F4A7C15ull) >> (64 - CACHE_BITS)];
      if (slot == v) continue;
      slot = v;
      const i64 lo = (i64)(v >> 32), up = (i64)(v & 0xFFFFFFFFull);
      if (lo < 1 || lo > nlo || up < 1 || up > nup)
        return fail(CC3D_B200_ERR_ARGUMENT, "merge_slabs: pair label out of range");
      const u32 a = find((u32)(off_lo + lo)), b = find((u32)(off_up + up));
      if (a < b) parent[b] = a; else if (b < a) parent[a] = b;
    }
  }
  // flatten once (ids ascend, roots are minima: a parent is final when its child is visited), then every later
  // question is a single load. A component is owned by the slab of its root; owned components are numbered slab by
  // slab in label order.
  for (i64 id = 1; id <= total; id++) parent[id] = parent[parent[id]];
  std::vector<i64> base(world + 1, 0);
  for (int r = 0; r < world; r++) {
    i64 owned = 0;
    for (i64 id = off[r] + 1; id <= off[r + 1]; id++) owned += parent[id] == (u32)id;
    base[r + 1] = base[r] + owned;
  }
  *n_total = base[world];
  // final labels of the roots that the labels of slab `rank` point at: roots lie in slabs <= rank
  // (final label of root id in slab r = base[r] + rank of id among the roots of slab r)
  final_tl.resize((size_t)off[rank + 1] + 1);
  i64* const final_of = final_tl.data();
  for (int r = 0; r <= rank; r++) {
    i64 next = base[r];
    for (i64 id = off[r] + 1; id <= off[r + 1]; id++)
      final_of[id] = parent[id] == (u32)id ? ++next : 0;
  }
  remap[0] = 0;
  const i64 off_me = off[rank];
  for (i64 l = 1; l <= n_labels[rank]; l++) remap[l] = final_of[parent[off_me + l]];
  return 0;
}

// ---- runs / draw (reference cc3d_graphs.hpp:470-523, fastcc3d.pyx:1258-1314) ----

template <typename T>
static void runs_count_typed(const T* lab, i64 n, u32* cnt, i64 nchunks, cudaStream_t s) {
  const unsigned blocks = (unsigned)std::min<i64>(nchunks, 148 * 32);
  if (((uintptr_t)lab & 15) == 0) k_runs_count_vec<T><<<blocks, 256, 0, s>>>(lab, n, cnt, nchunks);
  else k_runs_count<T><<<blocks, 256, 0, s>>>(lab, n, cnt, nchunks);
}
template <typename T>
static void runs_emit_typed(const T* lab, i64 n, const u32* prefix, i64 nchunks, u64* values, u64* starts, u64* ends,
                            cudaStream_t s) {
  const unsigned blocks = (unsigned)std::min<i64>(nchunks, 148 * 32);
  if (((uintptr_t)lab & 15) == 0) k_runs_emit_vec<T><<<blocks, 256, 0, s>>>(lab, n, prefix, nchunks, values, starts, ends);
  else k_runs_emit<T><<<blocks, 256, 0, s>>>(lab, n, prefix, nchunks, values, starts, ends);
}

int cc3d_b200_runs(const void* labels, int kind, int64_t voxels, uint64_t* values, uint64_t* starts, uint64_t* ends,
                   uint64_t capacity, uint64_t* count, int mem_space, void* stream) {
  const size_t es = kind_size(kind);
  if (!es || kind > CC3D_B200_U64) return fail(CC3D_B200_ERR_KIND, "Unsupported type: labels must be u8/u16/u32/u64");
  if (!count) return fail(CC3D_B200_ERR_ARGUMENT, "count must not be NULL");
  if (voxels < 0) return fail(CC3D_B200_ERR_ARGUMENT, "negative size");
  if (voxels >= (1ll << 43)) return fail(CC3D_B200_ERR_TOO_LARGE, "runs: more than 2^43 voxels");
  *count = 0;
  if (voxels == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  const bool host = mem_space == CC3D_B200_HOST;
  const i64 nchunks = (voxels + CC_RUN_CHUNK - 1) / CC_RUN_CHUNK;
  const i64 nb = std::max<i64>(1, (nchunks + CC_SCAN_CHUNK - 1) / CC_SCAN_CHUNK);
  size_t need = 4096 + (size_t)(nchunks + 4) * 4 + (size_t)(nb + 1) * 8 + 1024;
  if (host) need += (size_t)voxels * es + 512 + (size_t)capacity * 24 + 1024;
  Arena ar;
  if (int rc = arena_acquire(need, &ar, s, true)) return rc;
  const void* dl = labels;
  if (host) {
    void* d = ar.take((size_t)voxels * es);
    cudaError_t e = cudaMemcpyAsync(d, labels, (size_t)voxels * es, cudaMemcpyHostToDevice, s);
    if (e != cudaSuccess) { arena_release(ar); return fail(CC3D_B200_ERR_CUDA, cudaGetErrorString(e)); }
    dl = d;
  }
  u32* cnt = (u32*)ar.take((size_t)(nchunks + 4) * 4);
  u64* bsum = (u64*)ar.take((size_t)(nb + 1) * 8);
  u64* total_dev = (u64*)ar.take(16);
  switch (kind) {
    case CC3D_B200_U8: runs_count_typed((const uint8_t*)dl, voxels, cnt, nchunks, s); break;
    case CC3D_B200_U16: runs_count_typed((const uint16_t*)dl, voxels, cnt, nchunks, s); break;
    case CC3D_B200_U32: runs_count_typed((const uint32_t*)dl, voxels, cnt, nchunks, s); break;
    default: runs_count_typed((const uint64_t*)dl, voxels, cnt, nchunks, s); break;
  }
  g_launches += 1;
  scan_counts(cnt, cnt, bsum, nchunks, nullptr, 0, total_dev, nullptr, s);
  u64 total = 0;
  cudaError_t e = cudaMemcpyAsync(&total, total_dev, 8, cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) { arena_release(ar); return fail(CC3D_B200_ERR_CUDA, std::string("runs: ") + cudaGetErrorString(e)); }
  *count = total;
  if (total >= (1ull << 32)) { arena_release(ar); return fail(CC3D_B200_ERR_TOO_LARGE, "runs: more than 2^32 runs"); }
  if (total == 0 || total > capacity) { arena_release(ar); return 0; }
  if (!values || !starts || !ends) { arena_release(ar); return fail(CC3D_B200_ERR_ARGUMENT, "runs: output arrays must not be NULL"); }
  u64 *dv = values, *ds = starts, *de = ends;
  if (host) {
    dv = (u64*)ar.take((size_t)total * 8); ds = (u64*)ar.take((size_t)total * 8); de = (u64*)ar.take((size_t)total * 8);
  }
  switch (kind) {
    case CC3D_B200_U8: runs_emit_typed((const uint8_t*)dl, voxels, cnt, nchunks, dv, ds, de, s); break;
    case CC3D_B200_U16: runs_emit_typed((const uint16_t*)dl, voxels, cnt, nchunks, dv, ds, de, s); break;
    case CC3D_B200_U32: runs_emit_typed((const uint32_t*)dl, voxels, cnt, nchunks, dv, ds, de, s); break;
    default: runs_emit_typed((const uint64_t*)dl, voxels, cnt, nchunks, dv, ds, de, s); break;
  }
  g_launches += 1;
  if (host) {
    cudaMemcpyAsync(values, dv, (size_t)total * 8, cudaMemcpyDeviceToHost, s);
    cudaMemcpyAsync(starts, ds, (size_t)total * 8, cudaMemcpyDeviceToHost, s);
    cudaMemcpyAsync(ends, de, (size_t)total * 8, cudaMemcpyDeviceToHost, s);
  }
  e = cudaStreamSynchronize(s);
  if (e == cudaSuccess) e = cudaGetLastError();
  arena_release(ar);
  if (e != cudaSuccess) return fail(CC3D_B200_ERR_CUDA, std::string("runs: ") + cudaGetErrorString(e));
  return 0;
}

template <typename T>
static void draw_typed(T* img, u64 value, const u64* starts, const u64* ends, u64 n_runs, u64 origin, u64* long_runs,
                       unsigned long long* n_long, bool any_long, cudaStream_t s) {
  const u64 warps_needed = n_runs;
  const unsigned blocks = (unsigned)std::min<u64>((warps_needed + 7) / 8, 148 * 32);
  k_draw_short<T><<<std::max(1u, blocks), 256, 0, s>>>(img, (T)value, starts, ends, n_runs, origin, long_runs, n_long);
  g_launches += 1;
  if (any_long) {
    k_draw_long<T><<<148 * 8, 256, 0, s>>>(img, (T)value, long_runs, n_long);
    g_launches += 1;
  }
}
static void draw_kind(void* img, int kind, u64 value, const u64* starts, const u64* ends, u64 n_runs, u64 origin,
                      u64* long_runs, unsigned long long* n_long, bool any_long, cudaStream_t s) {
  switch (kind) {
    case CC3D_B200_U8: draw_typed((uint8_t*)img, value, starts, ends, n_runs, origin, long_runs, n_long, any_long, s); break;
    case CC3D_B200_U16: draw_typed((uint16_t*)img, value, starts, ends, n_runs, origin, long_runs, n_long, any_long, s); break;
    case CC3D_B200_U32: draw_typed((uint32_t*)img, value, starts, ends, n_runs, origin, long_runs, n_long, any_long, s); break;
    default: draw_typed((uint64_t*)img, value, starts, ends, n_runs, origin, long_runs, n_long, any_long, s); break;
  }
}

int cc3d_b200_draw(void* image, int kind, int64_t voxels, uint64_t value, const uint64_t* starts, const uint64_t* ends,
                   uint64_t n_runs, int mem_space, void* stream) {
  const size_t es = kind_size(kind);
  if (!es || kind > CC3D_B200_U64) return fail(CC3D_B200_ERR_KIND, "Unsupported type: image must be u8/u16/u32/u64");
  if (voxels < 0) return fail(CC3D_B200_ERR_ARGUMENT, "negative size");
  if (n_runs == 0) return 0;
  if (!image || !starts || !ends) return fail(CC3D_B200_ERR_ARGUMENT, "draw: NULL argument");
  cudaStream_t s = (cudaStream_t)stream;
  if (mem_space == CC3D_B200_HOST) {
    // the run list is on the host: validate it here, stage only the window of the image the runs span
    u64 lo = ~0ull, hi = 0, n_long = 0;
    for (u64 r = 0; r < n_runs; r++) {
      const u64 a = starts[r], b = ends[r];
      if (a >= b || b > (u64)voxels) return fail(CC3D_B200_ERR_ARGUMENT, "Invalid run.");
      lo = std::min(lo, a); hi = std::max(hi, b);
      n_long += (b - a > CC_RUN_LONG);
    }
    const size_t window = (size_t)(hi - lo);
    Arena ar;
    if (int rc = arena_acquire(window * es + (size_t)n_runs * 16 + (size_t)n_long * 16 + 4096, &ar, s, true)) return rc;
    void* dimg = ar.take(window * es);
    u64* dst = (u64*)ar.take((size_t)n_runs * 8);
    u64* den = (u64*)ar.take((size_t)n_runs * 8);
    u64* dlong = (u64*)ar.take((size_t)n_long * 16 + 16);
    unsigned long long* dn = (unsigned long long*)ar.take(16);
    char* himg = (char*)image + (size_t)lo * es;
    cudaMemcpyAsync(dimg, himg, window * es, cudaMemcpyHostToDevice, s);
    cudaMemcpyAsync(dst, starts, (size_t)n_runs * 8, cudaMemcpyHostToDevice, s);
    cudaMemcpyAsync(den, ends, (size_t)n_runs * 8, cudaMemcpyHostToDevice, s);
    cudaMemsetAsync(dn, 0, 16, s);
    draw_kind(dimg, kind, value, dst, den, n_runs, lo, dlong, dn, n_long > 0, s);
    cudaError_t e = cudaMemcpyAsync(himg, dimg, window * es, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e == cudaSuccess) e = cudaGetLastError();
    arena_release(ar);
    if (e != cudaSuccess) return fail(CC3D_B200_ERR_CUDA, std::string("draw: ") + cudaGetErrorString(e));
    return 0;
  }
  // device run list: validate with a kernel (one small read back), then draw
  Arena ar;
  if (int rc = arena_acquire(4096, &ar, s, true)) return rc;
  unsigned long long* flags = (unsigned long long*)ar.take(32);   // [0] invalid, [1] long runs, [2] long-list cursor
  cudaMemsetAsync(flags, 0, 32, s);
  k_draw_check<<<(unsigned)std::min<u64>((n_runs + 255) / 256, 148 * 8), 256, 0, s>>>(starts, ends, n_runs, (u64)voxels, flags);
  g_launches += 1;
  unsigned long long hflags[2] = {0, 0};
  cudaError_t e = cudaMemcpyAsync(hflags, flags, 16, cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) { arena_release(ar); return fail(CC3D_B200_ERR_CUDA, std::string("draw: ") + cudaGetErrorString(e)); }
  if (hflags[0]) { arena_release(ar); return fail(CC3D_B200_ERR_ARGUMENT, "Invalid run."); }
  Arena ar2;
  u64* dlong = nullptr;
  if (hflags[1]) {
    if (int rc = arena_acquire((size_t)hflags[1] * 16 + 4096, &ar2, s, true)) { arena_release(ar); return rc; }
    dlong = (u64*)ar2.take((size_t)hflags[1] * 16);
  }
  draw_kind(image, kind, value, starts, ends, n_runs, 0, dlong, flags + 2, hflags[1] > 0, s);
  e = cudaStreamSynchronize(s);
  if (e == cudaSuccess) e = cudaGetLastError();
  arena_release(ar2);
  arena_release(ar);
  if (e != cudaSuccess) return fail(CC3D_B200_ERR_CUDA, std::string("draw: ") + cudaGetErrorString(e));
  return 0;
}
