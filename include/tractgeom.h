/*
 * tractgeom.h — C ABI of the B200-native streamline-metrics path.
 *
 * Drop-in boundary for ONE reference function:
 *     compute_streamline_metrics(vtk_path, max_streamlines=None) -> (df_sl, df_bundle)
 *     /root/reference/src/geometry/tract_geom_proc.py:153-212
 * The reference has no FFI of its own (pure Python/numpy); these entry points are what a ctypes
 * binding for that function needs (see INTEGRATION.md for the stub a maintainer would add).
 *
 * Data layout ("CSR tractogram"):
 *     xyz      P x 3 coordinates, row-major (x0 y0 z0 x1 y1 z1 ...), float64 or float32
 *     offsets  int64[S+1]; polyline s owns points offsets[s] .. offsets[s+1]-1
 *              (what tract_geom_proc.py:17-20 gathers one line at a time from the VTK `lines` array)
 *     out      float64[17 x S], column-major by metric: out[m*S + s]; metric order = TG_METRIC_* below
 *              = the key order of the per-streamline dict at tract_geom_proc.py:164-187
 *     keep     uint8[S] bit flags: TG_KEEP_LOADER = passes the loader filter (n > 2 and all
 *              coordinates finite, tract_geom_proc.py:21); TG_KEEP_LENGTH = arc length > 1e-8
 *              (tract_geom_proc.py:160).  A row of df_sl exists iff both bits are set.
 *              Rows that fail either filter hold NaN in all 17 columns.
 *
 * All functions return 0 on success, a negative TG_E_* code on failure; tg_last_error() returns a
 * static thread-local message for the last failure.  Nothing here calls exit()/abort().
 * A context is bound to one CUDA device and is not thread-safe (the reference is single-threaded,
 * comprehensive_tract_geometry_analysis.py:169-195 calls it serially).  Its device scratch (work queue,
 * tile partials, float64 copy of float32 points) is shared by all calls: calls on ONE context must be
 * stream-ordered with respect to each other — use one stream per context, or one context per stream.
 */
#ifndef TRACTGEOM_H
#define TRACTGEOM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TG_ABI_VERSION 3   /* 2: + tg_bundle_spread_dev, tg_metrics_csr_host_ex, tg_resample_csr_*, host ingest helpers; 3: + tg_build_id,
                              tg_bundle_partials_dev, tg_batch_*, big-endian point storage (earlier entry points unchanged) */

#define TG_N_METRICS 17
enum tg_metric {                 /* tract_geom_proc.py:164-187 */
    TG_LENGTH = 0, TG_END_TO_END, TG_TORTUOSITY, TG_STRAIGHTNESS, TG_CURV_MEAN, TG_CURV_STD,
    TG_CURV_ENERGY, TG_TORSION_MEAN, TG_BEND_ANGLE_MEAN, TG_BBOX_VOL, TG_ELONGATION_RATIO,
    TG_PLANARITY_RATIO, TG_ANISOTROPY_RATIO, TG_CENTROID_X, TG_CENTROID_Y, TG_CENTROID_Z,
    TG_ANG_DISPERSION
};

#define TG_N_BUNDLE_COLS 13      /* tract_geom_proc.py:197-209: the 13 nan-means, in this order */
/* df_sl column index feeding bundle column j */
static const int TG_BUNDLE_SOURCE[TG_N_BUNDLE_COLS] = {
    TG_LENGTH, TG_TORTUOSITY, TG_CURV_MEAN, TG_CURV_ENERGY, TG_TORSION_MEAN, TG_BEND_ANGLE_MEAN,
    TG_ELONGATION_RATIO, TG_PLANARITY_RATIO, TG_ANISOTROPY_RATIO, TG_ANG_DISPERSION,
    TG_CENTROID_X, TG_CENTROID_Y, TG_CENTROID_Z
};

#define TG_KEEP_LOADER 1u
#define TG_KEEP_LENGTH 2u
#define TG_KEEP_BOTH   3u

enum tg_status {
    TG_OK = 0,
    TG_E_INVALID = -1,    /* bad argument (null pointer, negative size, non-monotone offsets) */
    TG_E_CUDA = -2,       /* a CUDA runtime call failed; message has the CUDA error string */
    TG_E_NOMEM = -3,      /* host or device allocation failed */
    TG_E_NODEVICE = -4    /* no usable CUDA device: there is NO CPU fallback */
};

/* Point storage.  The _BE forms are what a legacy BINARY VTK file holds (`POINTS n float|double`, big-endian): the loader
 * hands the file's bytes over as they are and the byte swap / exact upcast to float64 happens on the device. */
enum tg_dtype { TG_F64 = 0, TG_F32 = 1, TG_F64_BE = 2, TG_F32_BE = 3 };

typedef struct tg_context tg_context;

/* Library / device ------------------------------------------------------------------------- */
int         tg_abi_version(void);
/* Identity of the binary: sha256 prefix over the kernel sources, this header and the nvcc command line
 * (lesion_condition_vae_b200/build.py::source_id).  Tests and bench.py compare it with the sources on disk. */
const char* tg_build_id(void);
const char* tg_last_error(void);
int         tg_device_count(int* count);

/* Create a context on `device` (one CUDA stream, grow-only scratch).  Reused across calls: the
 * reference driver makes 2,368 of them per run (comprehensive_tract_geometry_analysis.py:169-195). */
int tg_create(int device, tg_context** ctx);
int tg_destroy(tg_context* ctx);
/* Block until everything queued on the context's stream has finished. */
int tg_synchronize(tg_context* ctx);
/* The context's cudaStream_t, as an opaque pointer (for callers that record their own events). */
int tg_stream(tg_context* ctx, void** stream);

/* Pinned host memory for the loader to parse straight into. */
int tg_host_alloc(void** ptr, size_t bytes);
int tg_host_free(void* ptr);

/* Per-streamline metrics ------------------------------------------------------------------- */
/* DEVICE pointers.  Replaces the loop at tract_geom_proc.py:158-187 (and the finite / n>2 test of
 * :21 and the length test of :160, reported through `keep`).  Asynchronous on `stream`
 * (a cudaStream_t, or NULL for the context's own stream).  S may be 0. */
int tg_metrics_csr_dev(tg_context* ctx, const void* d_xyz, int xyz_dtype, const int64_t* d_offsets,
                       int64_t S, int64_t P, double* d_out, uint8_t* d_keep, void* stream);

/* Per-bundle partial moments, DEVICE data, HOST bundle table.  Replaces tract_geom_proc.py:191-210.
 * Bundle b owns streamlines bundle_offsets[b] .. bundle_offsets[b+1]-1 (host int64[B+1]).
 * If d_select != NULL it is a uint8[S] mask ANDed with (keep == TG_KEEP_BOTH) (used for the
 * max_streamlines prefix rule).  Outputs (device):
 *     d_sums   float64[B x 13]  sum over kept rows of each bundle column, NaN entries skipped
 *                               (+-inf is NOT skipped, exactly like np.nanmean)
 *     d_counts int64  [B x 14]  [b][0] = kept rows (n_streamlines); [b][1+j] = non-NaN count of column j
 * mean_j = sums[b][j] / counts[b][1+j]  (NaN when the count is 0).  Deterministic summation order. */
int tg_bundle_reduce_dev(tg_context* ctx, const double* d_out, const uint8_t* d_keep,
                         const uint8_t* d_select, int64_t S, const int64_t* h_bundle_offsets,
                         int64_t B, double* d_sums, int64_t* d_counts, void* stream);

/* The same reduction, delivered as the multi-GPU exchange payload (the aggregate being split across
 * GPUs is tract_geom_proc.py:191-210): d_partials float64[B x 27] = {13 sums | kept rows | 13 non-NaN counts} per bundle,
 * counts as doubles (exact below 2^53).  A rank calls this on its CSR shard and all-gathers the B x 27 row block on the
 * same stream; every rank then adds the gathered blocks in rank order (sharding.combine_partials). */
int tg_bundle_partials_dev(tg_context* ctx, const double* d_out, const uint8_t* d_keep,
                           const uint8_t* d_select, int64_t S, const int64_t* h_bundle_offsets,
                           int64_t B, double* d_partials, void* stream);

/* OPT-IN spread of the same 13 bundle columns (not part of the reference's df_bundle:
 * tract_geom_proc.py:193 defines `_safe_std` = np.nanstd and never calls it).  Call after
 * tg_bundle_reduce_dev with the same table, mask and bundle table, passing its d_sums / d_counts.
 *     d_spread float64[B x 13 x 3]  {np.nanstd (ddof 0, two-pass about the bundle mean), np.nanmin, np.nanmax}
 *                                   over the kept rows; NaN where the column has no non-NaN entry;
 *                                   a column holding +-inf has std NaN (inf - inf), like numpy. */
int tg_bundle_spread_dev(tg_context* ctx, const double* d_out, const uint8_t* d_keep,
                         const uint8_t* d_select, int64_t S, const int64_t* h_bundle_offsets,
                         int64_t B, const double* d_sums, const int64_t* d_counts,
                         double* d_spread, void* stream);

/* HOST-buffer convenience calls (the end-to-end path: H2D + kernels + D2H, synchronous).
 * h_xyz / h_offsets / h_out / h_keep are host pointers (pinned is faster, pageable works).
 * h_out may be NULL when only bundle statistics are wanted (what the reference driver consumes,
 * comprehensive_tract_geometry_analysis.py:109); h_keep may be NULL.  h_sums/h_counts as above. */
int tg_metrics_csr_host(tg_context* ctx, const void* h_xyz, int xyz_dtype, const int64_t* h_offsets,
                        int64_t S, int64_t P, const int64_t* h_bundle_offsets, int64_t B,
                        double* h_out, uint8_t* h_keep, double* h_sums, int64_t* h_counts);

/* tg_metrics_csr_host plus the opt-in spread: h_spread float64[B x 13 x 3] as in tg_bundle_spread_dev, or NULL. */
int tg_metrics_csr_host_ex(tg_context* ctx, const void* h_xyz, int xyz_dtype, const int64_t* h_offsets,
                           int64_t S, int64_t P, const int64_t* h_bundle_offsets, int64_t B,
                           double* h_out, uint8_t* h_keep, double* h_sums, int64_t* h_counts,
                           double* h_spread);

/* A BATCH of tract files in one device call (the loop at comprehensive_tract_geometry_analysis.py:169-195 makes 2,368
 * separate calls): begin with capacities, PUSH every file's points as soon as it is parsed — the host-to-device copy
 * is queued on a copy stream and returns at once when h_xyz is pinned (tg_host_alloc), so parsing file i+1 overlaps
 * the transfer of file i; the buffer must stay untouched until tg_batch_run returns — then RUN: metrics of all pushed
 * polylines + the bundle reduction over h_bundle_offsets (indices into the concatenated polylines, usually one bundle
 * per file).  Files may differ in storage type.  h_out (17 x S_total, column-major), h_keep, h_spread may be NULL. */
int tg_batch_begin(tg_context* ctx, int64_t P_capacity, int64_t S_capacity);
int tg_batch_push(tg_context* ctx, const void* h_xyz, int xyz_dtype, int64_t P, const int64_t* h_offsets, int64_t S);
int tg_batch_run(tg_context* ctx, const int64_t* h_bundle_offsets, int64_t B, double* h_out, uint8_t* h_keep,
                 double* h_sums, int64_t* h_counts, double* h_spread);
int tg_batch_size(tg_context* ctx, int64_t* S_total, int64_t* P_total);

/* Arc-length resampling of every polyline to n_nodes points: the ragged-to-fixed step in
 * front of src/vae/data_loader.py:94-100, which expects exactly 100 `point_id`s per streamline and for which
 * the reference ships no producer.  Node k lies at arc length k L/(n_nodes-1), linearly interpolated inside
 * its segment; node 0 / node n_nodes-1 are the first / last point.  nodes: float64[S x n_nodes x 3].
 * A polyline of zero length (or one point) repeats its first point; an empty polyline, or one holding a
 * non-finite coordinate, gives NaN nodes.  n_nodes >= 2. */
int tg_resample_csr_dev(tg_context* ctx, const void* d_xyz, int xyz_dtype, const int64_t* d_offsets,
                        int64_t S, int64_t P, int n_nodes, double* d_nodes, void* stream);
int tg_resample_csr_host(tg_context* ctx, const void* h_xyz, int xyz_dtype, const int64_t* h_offsets,
                         int64_t S, int64_t P, int n_nodes, double* h_nodes);

/* Host-side ingest helpers for legacy VTK tract files (replace the Python `while` walk of
 * tract_geom_proc.py:17-25 and the text parsing behind pv.read).  No device, no context needed.
 *   tg_vtk_lines_to_csr: legacy cell array [n, i0..i(n-1), n, ...] of L ints -> CSR offsets (capacity L + 1)
 *                        and connectivity (capacity L); TG_E_INVALID when a count is negative or overruns.
 *   tg_parse_ascii_f64 / _i64: `want` whitespace-separated numbers from text[0, len) (inf / nan accepted for
 *                        doubles); *consumed = bytes read; TG_E_INVALID when the text ends or is malformed. */
int tg_vtk_lines_to_csr(const int64_t* lines, int64_t L, int64_t* offsets, int64_t* conn,
                        int64_t* n_cells, int64_t* n_conn);
/*   tg_vtk_cells_be32_to_csr: the same walk on the cell array AS THE BINARY FILE HOLDS IT (L big-endian int32), one pass, no
 *                        widened copy; *identity = 1 when the connectivity is 0,1,2,... (conn is then not written). */
int tg_vtk_cells_be32_to_csr(const void* cells_be, int64_t L, int64_t* offsets, int64_t* conn,
                             int64_t* n_cells, int64_t* n_conn, int* identity);
int tg_parse_ascii_f64(const char* text, int64_t len, int64_t want, double* out, int64_t* consumed);
int tg_parse_ascii_i64(const char* text, int64_t len, int64_t want, int64_t* out, int64_t* consumed);

/* Introspection for bench.py / tests: kernels launched by this context since creation. */
int tg_launch_count(tg_context* ctx, int64_t* launches);

#ifdef __cplusplus
}
#endif
#endif /* TRACTGEOM_H */
