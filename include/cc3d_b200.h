/* cc3d_b200.h — C-ABI of the B200-native connected-components engine.
 *
 * Drop-in boundary for the hot path of seung-lab/connected-components-3d (cc3d 4.x): these entry
 * points replace the C++ template the reference's Cython layer binds,
 *
 *   cc3d::connected_components3d<T,OUT>(T* in_labels, sx, sy, sz, max_labels, connectivity, delta,
 *                                      OUT* out_labels, size_t& N, periodic_boundary, binary_image)
 *   declared  cc3d/fastcc3d.pyx:67-74, defined cc3d/cc3d_continuous.hpp:394-455,
 *   18 instantiations at cc3d/fastcc3d.pyx:471-608,
 *
 * plus the pre-pass cc3d::estimate_provisional_label_count<T> (cc3d/cc3d.hpp:287-315, bound at
 * cc3d/fastcc3d.pyx:60-66) and the Cython statistics loops (cc3d/fastcc3d.pyx:771-938) and the
 * masking step of dust (cc3d/__init__.py:121-151).
 *
 * Conventions
 *   - Plain pointers and sizes only. `x` is the fastest-varying memory axis (the reference reverses
 *     C-order shapes before the call, fastcc3d.pyx:352-359), so index = x + sx*(y + sy*z).
 *   - `mem_space`: CC3D_B200_HOST pointers are staged through device memory inside the call;
 *     CC3D_B200_DEVICE pointers are used in place (zero copy).
 *   - `stream` is a cudaStream_t (NULL = default stream). Calls are synchronous with respect to
 *     the host on return unless stated otherwise.
 *   - All functions return 0 on success or a negative cc3d_b200_status; cc3d_b200_last_error()
 *     gives the message (thread-local).
 *   - Signed integers are passed as their unsigned views, bool as u8, float16 (delta==0) as u16,
 *     exactly as the reference does (fastcc3d.pyx:346-350, 472-563).
 */
#ifndef CC3D_B200_H
#define CC3D_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  CC3D_B200_U8 = 0, CC3D_B200_U16 = 1, CC3D_B200_U32 = 2, CC3D_B200_U64 = 3,
  CC3D_B200_F32 = 4, CC3D_B200_F64 = 5
} cc3d_b200_kind;

typedef enum { CC3D_B200_HOST = 0, CC3D_B200_DEVICE = 1 } cc3d_b200_mem_space;

typedef enum {
  CC3D_B200_OK = 0,
  CC3D_B200_ERR_CONNECTIVITY = -1, /* "Only 4 and 8 2D and 6, 18, and 26 3D connectivities are supported." cc3d.hpp:1469 */
  CC3D_B200_ERR_2D_NEEDS_SZ1 = -2, /* "sz must be 1 for 2D connectivities." cc3d.hpp:1451-1452 */
  CC3D_B200_ERR_PERIODIC_CONTINUOUS = -3, /* cc3d_continuous.hpp:423-425 */
  CC3D_B200_ERR_KIND = -4,
  CC3D_B200_ERR_TOO_LARGE = -5,    /* more voxels than one call supports (use the sharded path) */
  CC3D_B200_ERR_CUDA = -6,
  CC3D_B200_ERR_ARGUMENT = -7,
  CC3D_B200_ERR_OUT_RANGE = -8     /* N does not fit the requested out kind */
} cc3d_b200_status;

/* Results of the resolve phase that the caller needs before it can allocate the output
 * (the reference's out-dtype rule depends on epl, fastcc3d.pyx:373-434). */
typedef struct {
  uint64_t N;                 /* number of connected components */
  uint64_t epl;               /* cc3d.hpp:287-315 transition count (0 for mask mode) */
  int64_t first_foreground_row; /* -1 when there is no foreground */
  int64_t last_foreground_row;
} cc3d_b200_resolve_info;

typedef struct cc3d_b200_session cc3d_b200_session;

const char* cc3d_b200_last_error(void);
const char* cc3d_b200_version(void);

/* Row a1: estimate_provisional_label_count (cc3d.hpp:287-315; fastcc3d.pyx:169-242).
 * Also returns the value range needed by the continuous 2D-8 path (vmin/vmax may be NULL; each
 * points at one element of the input kind). */
int cc3d_b200_prepass(const void* in, int in_kind, int64_t sx, int64_t sy, int64_t sz, int mem_space,
                      uint64_t* epl, int64_t* first_foreground_row, int64_t* last_foreground_row,
                      void* vmin, void* vmax, void* stream);

/* Rows a3-a11, phase 1: label + merge + resolve. Leaves the resolved union-find forest in the
 * session and reports N / epl so that the caller can apply the out-dtype rule and allocate.
 * `delta` points at one element of the input kind (ignored when binary_image != 0).
 * Dispatch precedence is the reference's: binary -> delta == 0 -> continuous. */
int cc3d_b200_label_resolve(const void* in, int in_kind, int64_t sx, int64_t sy, int64_t sz,
                            int connectivity, const void* delta, int binary_image,
                            int periodic_boundary, int mem_space, void* stream,
                            cc3d_b200_resolve_info* info, cc3d_b200_session** session);

/* Rows a10/a12, phase 2: write final labels 1..N (first-appearance order) as out_kind
 * (CC3D_B200_U16/U32/U64). May be called once per session; releases the session. */
int cc3d_b200_label_write(cc3d_b200_session* session, void* out, int out_kind, int mem_space,
                          void* stream);

/* Sharded volumes (z-slabs across GPUs); device pointers only where noted.
 * label_write_rows: local labels (uint32, 1..N of this session) of rows [row_begin,row_end) into a
 * compact buffer; keeps the session. Used to publish a slab's boundary planes. */
int cc3d_b200_label_write_rows(cc3d_b200_session* session, int64_t row_begin, int64_t row_end, uint32_t* out,
                               int mem_space, void* stream);

/* label_write_remap: out[i] = remap[local label of i] (remap has N+1 entries of remap_kind u32/u64,
 * remap[0] = 0; max_label = largest value in remap, checked against out_kind). Releases the session. */
int cc3d_b200_label_write_remap(cc3d_b200_session* session, const void* remap, int remap_kind, uint64_t max_label,
                                void* out, int out_kind, int mem_space, void* stream);

/* Equivalences across one slab interface (DEVICE pointers): `upper` = first plane of the slab that
 * comes later in memory, `lower` = last plane of the slab before it; values of in_kind, labels = the
 * slabs' local labels. Appends (lower_label << 32 | upper_label) for every edge between the planes
 * (6/18/26 neighbourhood, same predicate as label_resolve) to pairs[capacity]; *count receives the
 * number of pairs produced (if > capacity the list was truncated: call again with more room).
 * GPU analogue of the face loops in connected_components_stack (cc3d/__init__.py:425-468). */
int cc3d_b200_face_pairs(const void* values_upper, const uint32_t* labels_upper, const void* values_lower,
                         const uint32_t* labels_lower, int in_kind, int64_t sx, int64_t sy, int connectivity,
                         const void* delta, int binary_image, uint64_t* pairs, uint64_t capacity, uint64_t* count,
                         void* stream);

/* Union-find over n_nodes compact node ids joined by n_pairs (a[i], b[i]) pairs (DEVICE pointers):
 * parent[i] = smallest node id of i's set. Replaces the Python DisjointSet of
 * connected_components_stack (cc3d/__init__.py:296-321). */
int cc3d_b200_solve_pairs(uint32_t* parent, int64_t n_nodes, const uint32_t* a, const uint32_t* b, int64_t n_pairs,
                          void* stream);

/* Host side of the slab merge (no GPU work; replaces the Python DisjointSet + renumber of
 * connected_components_stack, cc3d/__init__.py:296-321, 425-492). Slab r has local labels 1..n_labels[r];
 * pairs[r][0..n_pairs[r]) are the packed face_pairs of the interface below slab r (pairs[0] unused).
 * Writes remap[0..n_labels[rank]] (local label -> global label, remap[0] = 0) for slab `rank` and the
 * global component count. Global numbering = first appearance in the whole volume. */
int cc3d_b200_merge_slabs(int world, const int64_t* n_labels, const uint64_t* const* pairs, const int64_t* n_pairs,
                          int rank, int64_t* remap, int64_t* n_total);

/* Slab merge ON THE DEVICE, enqueue-only (no host synchronisation; replaces the Python DisjointSet + renumber of
 * connected_components_stack, cc3d/__init__.py:296-321, 425-492, and the host merge cc3d_b200_merge_slabs above).
 * `gathered` (device) is the all-gathered buffer of the slab step: `world` rows of `row_stride` int64, row r =
 * [N_r, epl_r, sz_r, n_pairs_r, pairs...], pairs packed as (label in slab r-1) << 32 | (label in slab r), at most
 * `pair_cap` per row. `workspace` (device, >= cc3d_b200_merge_workspace_bytes(label_cap) bytes) holds the union-find
 * over the slab-label ids (the sum of N_r must stay below label_cap). On return *remap points at the uint32 remap
 * table of slab `rank` inside the workspace (remap[local label] = global label, remap[0] = 0; feed it to
 * cc3d_b200_slab_finish) and *result at five device uint64: [0] N of the whole volume, [1] label_cap exceeded,
 * [2] pair_cap exceeded (in either case the tables are invalid: grow the capacity and repeat the step), [3..4] internal,
 * [8 + 4 r + k] = fact k (N, epl, sz, n_pairs) of slab r, so that one small copy gives the host all it needs. */
size_t cc3d_b200_merge_workspace_bytes(uint64_t label_cap);
int cc3d_b200_merge_slabs_device(const int64_t* gathered, int world, int64_t row_stride, int rank, uint64_t pair_cap,
                                 void* workspace, uint64_t label_cap, uint32_t** remap, uint64_t** result, void* stream);
/* The same merge in ONE single-CTA launch for small interface graphs (the sum of N_r below 65 535; same arguments,
 * workspace and results). result[5] = 1: the graph is larger - nothing was computed, call cc3d_b200_merge_slabs_device. */
int cc3d_b200_merge_slabs_device_small(const int64_t* gathered, int world, int64_t row_stride, int rank, uint64_t pair_cap,
                                 void* workspace, uint64_t label_cap, uint32_t** remap, uint64_t** result, void* stream);

/* Sharded fast path (one process per GPU, small slabs): the three calls below only ENQUEUE work on `stream`;
 * none of them synchronises, so a whole slab step needs one host synchronisation (after the all-gather of the
 * facts and face pairs). Device memory only.
 *   slab_begin : resolve phase of a slab + the resolved local labels of its first / last z-plane (uint32,
 *                sy*sx each, either may be NULL) + facts[3] = {N_local, epl, sz} (int64, device), all on stream.
 *   face_pairs_async : as cc3d_b200_face_pairs, count stays on the device (*count_dev must be zero on entry
 *                in stream order; it may exceed `capacity`, pairs past the capacity are dropped).
 *   slab_finish: final write through the local->global remap table (device), then the session's workspace is
 *                returned to the cache behind an event (the next user waits for the write on its own stream).
 * Together they replace the per-slab body of the reference's connected_components_stack loop
 * (cc3d/__init__.py:399-470) for device-resident slabs. */
int cc3d_b200_slab_begin(const void* in, int in_kind, int64_t sx, int64_t sy, int64_t sz, int connectivity,
                         const void* delta, int binary_image, void* stream, cc3d_b200_session** session,
                         uint32_t* labels_first_plane, uint32_t* labels_last_plane, int64_t* facts);
int cc3d_b200_face_pairs_async(const void* values_upper, const uint32_t* labels_upper, const void* values_lower,
                               const uint32_t* labels_lower, int in_kind, int64_t sx, int64_t sy, int connectivity,
                               const void* delta, int binary_image, uint64_t* pairs, uint64_t capacity,
                               uint64_t* count_dev, void* stream);
int cc3d_b200_slab_finish(cc3d_b200_session* session, const void* remap, int remap_kind, void* out, int out_kind,
                          void* stream);

/* Drops a session without writing. */
void cc3d_b200_session_release(cc3d_b200_session* session);

/* One-shot convenience: resolve + write into a caller-chosen out kind. */
int cc3d_b200_label(const void* in, int in_kind, int64_t sx, int64_t sy, int64_t sz, int connectivity,
                    const void* delta, int binary_image, int periodic_boundary, void* out,
                    int out_kind, int mem_space, uint64_t* N, void* stream);

/* Same, and also returns what label_resolve reports (N, epl, foreground rows), so that a caller that guessed
 * the out kind before the call (normally u32) can check the reference's out-dtype rule afterwards and only
 * convert in the rare case the guess was wrong. No host round trip between the two phases. */
int cc3d_b200_label_with_info(const void* in, int in_kind, int64_t sx, int64_t sy, int64_t sz, int connectivity,
                              const void* delta, int binary_image, int periodic_boundary, void* out,
                              int out_kind, int mem_space, cc3d_b200_resolve_info* info, void* stream);

/* Row a13: per-label statistics in MEMORY axes (x fastest). counts[N+1] (uint32, wraps like the
 * reference), bbox[(N+1)*6] uint32 as xmin,xmax,ymin,ymax,zmin,zmax (absent label: min=UINT32_MAX,
 * max=0), sums[(N+1)*3] uint64 coordinate sums (centroid = sum / count, exact below 2^53).
 * Labels > N are ignored. Output arrays live in the same mem_space as `labels`. */
int cc3d_b200_statistics(const void* labels, int kind, int64_t sx, int64_t sy, int64_t sz, uint64_t N,
                         uint32_t* counts, uint32_t* bbox, uint64_t* sums, int mem_space, void* stream);

/* statistics when the largest label is not known beforehand (the reference runs np.max first, fastcc3d.pyx:713-720;
 * here the same sweep tracks the maximum): labels below `cap` are accumulated into tables of `cap` entries
 * (counts[cap], bbox[cap][6], sums[cap][3], layout as cc3d_b200_statistics) and *max_label (host) receives the true
 * maximum. If *max_label < cap the first *max_label + 1 entries are the statistics (host tables receive exactly those);
 * otherwise call again with cap > *max_label. */
int cc3d_b200_statistics_auto(const void* labels, int kind, int64_t sx, int64_t sy, int64_t sz, uint64_t cap,
                              uint64_t* max_label, uint32_t* counts, uint32_t* bbox, uint64_t* sums, int mem_space,
                              void* stream);

/* Row a14: masking step of dust: img[i] = keep[labels[i]] ? img[i] : 0, in place.
 * keep has N+1 bytes (index = label). img_itemsize in {1,2,4,8}. */
int cc3d_b200_mask_by_label(void* img, int img_itemsize, const void* labels, int label_kind,
                            int64_t voxels, const uint8_t* keep, uint64_t N, int mem_space, void* stream);

/* ---- callers either side of the labelling path ---- */

/* cc3d.voxel_connectivity_graph (fastcc3d.pyx:1021-1170 -> extract_voxel_connectivity_graph,
 * cc3d_graphs.hpp:31-247): bit b of graph[p] is cleared iff the neighbour in direction b exists and holds a
 * different value. graph is uint8 for connectivity 4, 8, 6 and uint32 for 18, 26 (the reference's bit order). */
int cc3d_b200_voxel_connectivity_graph(const void* labels, int kind, int64_t sx, int64_t sy, int64_t sz,
                                       int connectivity, void* graph, int mem_space, void* stream);

/* cc3d.color_connectivity_graph (fastcc3d.pyx:941-1018 -> color_connectivity_graph_N, cc3d_graphs.hpp:1076-1106):
 * labels the components of a voxel connectivity graph (uint8 or uint32; connectivity 4/8 for sz == 1, 6/26
 * otherwise; 26 needs uint32) following the backward bits of every voxel, numbered by first appearance.
 * out: uint32[voxels]; *N = number of components. */
int cc3d_b200_color_connectivity_graph(const void* vcg, int vcg_kind, int64_t sx, int64_t sy, int64_t sz,
                                       int connectivity, uint32_t* out, uint64_t* N, int mem_space, void* stream);

/* cc3d.contacts / cc3d.region_graph (fastcc3d.pyx:1180-1252 -> extract_region_graph, cc3d_graphs.hpp:300-468):
 * every pair of different non-zero labels that touch under the connectivity, with the NUMBER of contacts per class
 * (class_counts[4*i + 0..3]: across x, across y, across z, edge/corner; 2D: across x, across y, -, diagonal), so that
 * the caller forms surface areas (count x face area) or voxel counts exactly. pairs[i] = min << 32 | max; label
 * values must be < 2^32. *count = number of pairs; if it exceeds `capacity` nothing is written: call again with
 * more room. Border behaviour of the reference's compute_neighborhood is reproduced. */
int cc3d_b200_contacts(const void* labels, int kind, int64_t sx, int64_t sy, int64_t sz, int connectivity,
                       uint64_t* pairs, uint32_t* class_counts, uint64_t capacity, uint64_t* count, int mem_space,
                       void* stream);

/* crackle v0 decode on the device (the on-disk format either side of the labelling path;
 * its colouring step is color_connectivity_graph, cc3d_graphs.hpp:583-1106, 1018-1074). `stream` (HOST) holds the
 * concatenated per-slice crack-code blobs, slice z = bytes [slice_off[z], slice_off[z+1]); the crack codes are turned
 * into the 4-bit pixel graph by one thread per slice, all slices are coloured at once (4-connected, numbered slice by
 * slice in raster order = the order of the file's key table) and out[i] = lut[component of i] (lut[0] unused, n_lut
 * entries, HOST). `out` (uint32, x fastest, sx*sy*sz) may be a HOST or a DEVICE pointer. *n_components = total number
 * of per-slice components. Flat labels / "impermissible" crack format only (the format of the reference's fixture). */
int cc3d_b200_crackle_v0_decode(const uint8_t* stream, const uint64_t* slice_off, int64_t sx, int64_t sy, int64_t sz,
                                const uint32_t* lut, uint64_t n_lut, uint32_t* out, uint64_t* n_components, int mem_space,
                                void* cuda_stream);

/* out[i] = table[labels[i]] (labels above N give 0): the relabelling step of cc3d.largest_k
 * (cc3d/__init__.py:262-276, fastremap.mask_except + renumber / runs + draw). out kind u8/u16/u32/u64. */
int cc3d_b200_remap_labels(const void* labels, int label_kind, int64_t voxels, const uint32_t* table, uint64_t N,
                           void* out, int out_kind, int mem_space, void* stream);

/* Row a14 fused: cc3d.dust (cc3d/__init__.py:71-155) without materialising the label volume. The image is
 * labelled (connectivity, binary_image as in cc3d_b200_label, delta = 0), component sizes come from the run table,
 * components with lo <= size < hi stay (a scalar threshold t is lo = t, hi = INT64_MAX; invert swaps which side
 * stays and, like np.isin(..., invert=True), also clears the background), and out[v] = stays ? img[v] : 0 is
 * written in the same pass that would expand the labels. out may be img (in place). *N = number of components,
 * *n_masked = number of components outside [lo, hi) (dust_N = N - n_masked, or n_masked when inverted). */
int cc3d_b200_dust(const void* img, void* out, int kind, int64_t sx, int64_t sy, int64_t sz, int connectivity,
                   int binary_image, int64_t lo, int64_t hi, int invert, int mem_space, uint64_t* N,
                   uint64_t* n_masked, void* stream);

/* cc3d.runs (fastcc3d.pyx:1258-1279 -> extract_runs, cc3d_graphs.hpp:470-503): the maximal runs of equal non-zero
 * values of the FLATTENED label array (memory order; runs continue across row ends), in position order:
 * values[k], starts[k], ends[k] (half open). *count = number of runs. When count > capacity nothing is written:
 * call again with more room (capacity 0 = count only; the arrays may then be NULL). Output arrays live in the
 * same mem_space as `labels`. The caller groups the table by value (the reference returns a std::map) and adds the
 * reference's single-voxel quirk (a 1-voxel array reports its run even when it is background, cc3d_graphs.hpp:481-484). */
int cc3d_b200_runs(const void* labels, int kind, int64_t voxels, uint64_t* values, uint64_t* starts, uint64_t* ends,
                   uint64_t capacity, uint64_t* count, int mem_space, void* stream);

/* cc3d.draw / erase (fastcc3d.pyx:1281-1326 -> set_run_voxels, cc3d_graphs.hpp:505-523): image[starts[k]..ends[k]) =
 * value for every run (value is truncated to the image kind; bool images are u8 with value 0/1). The run list is
 * validated first: a run with start >= end or end > voxels gives CC3D_B200_ERR_ARGUMENT "Invalid run." and leaves
 * the image untouched (the reference throws at the first bad run, after drawing the ones before it). image, starts
 * and ends share mem_space; for host images only the window the runs span is staged through the device. */
int cc3d_b200_draw(void* image, int kind, int64_t voxels, uint64_t value, const uint64_t* starts, const uint64_t* ends,
                   uint64_t n_runs, int mem_space, void* stream);

/* Device-memory workspace currently cached by the library on the active device (bytes), and a
 * call that frees it. */
size_t cc3d_b200_workspace_bytes(void);
void cc3d_b200_release_workspace(void);

/* Testing aid: capacity (entries) of the global edge queue between the tile kernel and the global union
 * kernel; 0 restores the default (8 per bitmap word). A tiny value forces the overflow fallback path. */
void cc3d_b200_debug_set_queue_capacity(uint64_t entries);
/* Tests: multilabel tiles with more than 16 runs per word are relabelled by a second launch with a larger shared-memory
 * forest once the process has met such a volume; this sets / clears that memory (1 / 0). */
void cc3d_b200_debug_set_big_tiles(int seen);

/* Number of CUDA kernels this library has launched in this process (all threads). */
unsigned long long cc3d_b200_launch_count(void);

/* Per-kernel device times (ms) of the last label call on this thread, for bench/profiling.
 * names[i] points at static strings; returns the number of entries filled (<= cap). */
int cc3d_b200_last_timings(const char** names, float* ms, int cap);
void cc3d_b200_set_timing(int enabled);

#ifdef __cplusplus
}
#endif
#endif /* CC3D_B200_H */
