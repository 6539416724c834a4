/*
 * pypairs_b200.h -- C ABI of the B200-native PyPairs hot paths.
 *
 * The reference (rfechtner/pypairs) is pure Python + Numba; it has no FFI.  Its
 * de-facto operator seam is the pair of kernel entry points
 *
 *     check_pairs(X[N,G] f64, cats[N] i64, thr[C] i64) -> i64[G,G]
 *         pypairs/tools/sandbag.py:126 (call site), :160-199 (body)
 *     get_phase_scores(X[N,G] f64, iterations, min_iter, min_pairs,
 *                      pairs[P,2] i64, used[G] bool) -> f64[N]
 *         pypairs/tools/cyclone.py:155 (call site), :223-257 (body)
 *
 * Each entry point below names the reference interface it replaces.  Plain C:
 * pointers and sizes only; no torch / C++ types cross the boundary.  One
 * process drives one GPU (one pp_ctx = one device + its streams and scratch);
 * multi-GPU jobs run one process per GPU, give every context a communicator
 * (pp_comm_init) and call the *_dist entry points, which issue the NCCL
 * collectives themselves; (shard, n_shards) of pp_sandbag remains for callers that
 * merge the shards' lists on their own.
 *
 * Conventions
 *   - every function returns 0 on success, a PP_E* code otherwise; the message
 *     is available from pp_last_error(ctx) (ctx may be NULL for create errors);
 *   - the library never exits or aborts the process;
 *   - `x` may be a host pointer (pageable or pinned) or a device pointer on the
 *     ctx's device (x_is_device != 0); the caller owns every input;
 *   - variable-length outputs are allocated by the library (large ones as pinned host
 *     blocks from an internal pool) and released with pp_free(); fixed-size outputs
 *     are caller-allocated host buffers (pp_alloc_result gives pinned ones);
 *   - a pp_ctx is not thread-safe; calls block until results are on the host;
 *   - STREAM CONTRACT for device-resident inputs (x_is_device / pp_csr.is_device): the
 *     library reads them on its own non-blocking streams, which do not synchronise with
 *     the caller's streams (not even the legacy default stream).  Everything that
 *     produces the buffers must have COMPLETED before the call (cudaStreamSynchronize /
 *     cudaEventSynchronize on the producing stream); the Python binding does this for
 *     torch tensors (pypairs_b200/_ffi.py:_order_after_torch).  Device buffers must live
 *     on the ctx's device.  When a call returns, the library no longer touches them.
 */
#ifndef PYPAIRS_B200_H
#define PYPAIRS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PP_ABI_VERSION 2

/* error codes */
#define PP_OK 0
#define PP_EINVAL 1   /* bad argument (message says which) */
#define PP_ECUDA 2    /* CUDA runtime error */
#define PP_ENOMEM 3   /* host or device allocation failed */
#define PP_ENAN 4     /* NaN in the expression matrix (undefined in the reference: fastmath, utils.py:185) */
#define PP_ELIMIT 5   /* size beyond what the kernels support (message says which limit) */
#define PP_ENCCL 6    /* NCCL could not be loaded / a collective failed / no communicator */

/* element type of the expression matrix */
#define PP_F32 0
#define PP_F64 1

/* cyclone permutation source */
#define PP_RNG_MT19937 0 /* bit-exact replay of Numba's np.random.seed(seed)+shuffle stream (cyclone.py:196-199,236) */
#define PP_RNG_PHILOX 1  /* counter-based Philox4x32-10 table generated on the device */
#define PP_RNG_TABLE 2   /* caller supplies the permutation table (validation mode) */

typedef struct pp_ctx pp_ctx;

/*
 * CSR form of the expression matrix (rows = samples / cells, columns = genes), e.g. the arrays of a
 * scipy.sparse.csr_matrix / AnnData.X.  The reference densifies such input on the host
 * (pypairs/utils.py:278-281, 290-292, 299-301); here it is densified chunk by chunk on the device, so
 * only the non-zeros cross PCIe.  Duplicate (row, column) entries are not allowed (scipy: sum_duplicates()),
 * explicit zeros are; column indices need not be sorted.  All three arrays live on the host, or all three on
 * the ctx's device (is_device != 0).
 */
typedef struct pp_csr {
    const void *indptr;     /* [n_rows + 1] int32 or int64 */
    int32_t indptr_is_64;
    const int32_t *indices; /* [nnz] column indices */
    const void *data;       /* [nnz] PP_F32 or PP_F64 */
    int32_t data_dtype;
    int32_t is_device;
} pp_csr;

/* Device-side timing of the last call, milliseconds (CUDA events on the ctx stream). */
typedef struct pp_timings {
    float h2d_ms;     /* host -> device copy of the expression matrix (0 if x_is_device) */
    float prep_ms;    /* rank transform + layout kernels */
    float main_ms;    /* the dominant kernel (sandbag pair kernel / cyclone score kernel) */
    float post_ms;    /* compaction sort / table build + device -> host of results */
    float total_ms;   /* first enqueue to results on host */
    int32_t n_planes; /* sandbag: bit planes used for the rank encoding */
    int32_t n_launches; /* kernels launched by this call */
    int64_t n_units;  /* sandbag: gene pairs evaluated by this shard; cyclone: cells scored by this shard */
    int64_t h2d_bytes; /* bytes of the expression matrix actually copied host -> device by this call */
    float avg_planes;  /* sandbag: mean number of bit planes compared per gene pair of this shard (<= n_planes) */
    int32_t reserved;
} pp_timings;

int pp_version(void);
int pp_device_count(void);
/* device < 0: use the current CUDA device. */
int pp_ctx_create(int device, pp_ctx **out);
void pp_ctx_destroy(pp_ctx *ctx);
const char *pp_last_error(pp_ctx *ctx);
void pp_free(void *p);
/*
 * pp_alloc_result -- host memory for an output array the caller passes to pp_cyclone* (out_scores, out_obs_*): blocks
 * of 64 KB and more come pinned from the library's pool, so the device -> host copy of the result lands in them
 * directly (no staging copy, no page faults of a fresh allocation); any other host pointer works too, through a
 * staging buffer.  Released with pp_free().  NULL if the allocation fails.
 */
void *pp_alloc_result(size_t bytes);
int pp_get_timings(pp_ctx *ctx, pp_timings *out);

/*
 * pp_sandbag -- replaces check_pairs + the np.where scan
 * (pypairs/tools/sandbag.py:126-129, 160-199).
 *
 * x            [n_rows, n_cols] matrix, row-major with leading dimension ld
 *              (elements), rows = samples, cols = genes (as the reference's
 *              `data`, sandbag.py:98-100).
 * cell_rows    [n_cells] row index of every KEPT sample, class-sorted: all
 *              samples of class 0 first, then class 1, ... (the host's
 *              equivalent of data[sample_mask] + cats, sandbag.py:120,126).
 * class_sizes  [n_classes] number of kept samples per class (all > 0).
 * gene_cols    [n_genes] column index of every KEPT gene, ascending
 *              (data[:, gene_mask], sandbag.py:126).
 * thresholds   [n_classes] ceil(n_k * fraction) (sandbag.py:212-217).
 * shard/n_shards  this call evaluates gene-pair tiles t with t % n_shards == shard.
 *
 * out_keys     *out_keys = malloc'ed uint64[n]: row * n_genes + col of every
 *              non -1 entry of the reference's result matrix that falls in this
 *              shard, ascending (== the reference's row-major np.where order,
 *              sandbag.py:129).  row/col index the KEPT genes.
 * out_class    *out_class = malloc'ed int32[n]: result[row, col].
 * drop_unexpressed  != 0: gene_cols are CANDIDATES; the library first drops every
 *              candidate whose column is zero in ALL n_rows rows of x (the
 *              reference's ~np.all(data == 0, axis=0), utils.py:368, evaluated
 *              before sample filtering) and returns the kept columns in
 *              *out_kept_genes (malloc'ed int32[*out_n_kept]); row/col of the
 *              keys then index that kept list.
 * probe_pairs  optional [n_probe,2] int64 (g1 < g2, kept-gene indices): the
 *              library writes num_pos / num_neg (sandbag.py:175-182) of those
 *              pairs to probe_pos / probe_neg [n_probe, n_classes] int32.
 */
int pp_sandbag(pp_ctx *ctx, const void *x, int x_dtype, int x_is_device, int64_t n_rows, int64_t n_cols,
               int64_t ld, const int32_t *cell_rows, int64_t n_cells, const int64_t *class_sizes,
               int32_t n_classes, const int32_t *gene_cols, int64_t n_genes, const int64_t *thresholds,
               int32_t shard, int32_t n_shards, uint64_t **out_keys, int32_t **out_class, int64_t *out_n,
               const int64_t *probe_pairs, int64_t n_probe, int32_t *probe_pos, int32_t *probe_neg,
               int32_t drop_unexpressed, int32_t **out_kept_genes, int64_t *out_n_kept);

/* pp_sandbag with the matrix given in CSR form; every other argument as in pp_sandbag. */
int pp_sandbag_csr(pp_ctx *ctx, const pp_csr *x, int64_t n_rows, int64_t n_cols, const int32_t *cell_rows,
                   int64_t n_cells, const int64_t *class_sizes, int32_t n_classes, const int32_t *gene_cols,
                   int64_t n_genes, const int64_t *thresholds, int32_t shard, int32_t n_shards,
                   uint64_t **out_keys, int32_t **out_class, int64_t *out_n, const int64_t *probe_pairs,
                   int64_t n_probe, int32_t *probe_pos, int32_t *probe_neg, int32_t drop_unexpressed,
                   int32_t **out_kept_genes, int64_t *out_n_kept);

/*
 * pp_cyclone -- replaces the per-class loop over get_phase_scores
 * (pypairs/tools/cyclone.py:151-158, 223-257) for all classes in one call.
 *
 * x              as above; rows = cells.  Rows [row_begin, row_begin+n_cells)
 *                are scored (a cell shard of the job).
 * n_classes      number of marker classes.
 * used_idx       concatenated ascending gene-column indices of each class's
 *                `used` mask (np.where(used[c])[0], cyclone.py:291);
 *                used_off[n_classes+1] delimits the classes.
 * pairs          concatenated [P_c,2] int32 pairs indexing the class's compacted
 *                used-gene vector (cyclone.py:294-306); pair_off[n_classes+1].
 * iterations, min_iter, min_pairs   as cyclone.py:17-26.
 * rng_mode/seed  PP_RNG_*; seed = settings.seed (2803) for parity.
 * perm_tables    PP_RNG_TABLE only: concatenated int32[iterations, U_c] tables,
 *                class after class (row k = cumulative permutation after k+1
 *                shuffles of arange(U_c)).
 * out_scores     host f64[n_classes, n_cells] (class-major), NaN per cyclone.py:243-249.
 * out_obs_hits/out_obs_total  optional host int32[n_classes, n_cells]: the
 *                observed counters of get_proportion (cyclone.py:206-215).
 */
int pp_cyclone(pp_ctx *ctx, const void *x, int x_dtype, int x_is_device, int64_t n_rows, int64_t n_cols,
               int64_t ld, int64_t row_begin, int64_t n_cells, int32_t n_classes, const int32_t *used_idx,
               const int64_t *used_off, const int32_t *pairs, const int64_t *pair_off, int32_t iterations,
               int32_t min_iter, int32_t min_pairs, int32_t rng_mode, uint64_t seed,
               const int32_t *perm_tables, double *out_scores, int32_t *out_obs_hits,
               int32_t *out_obs_total);

/* pp_cyclone with the matrix given in CSR form; every other argument as in pp_cyclone. */
int pp_cyclone_csr(pp_ctx *ctx, const pp_csr *x, int64_t n_rows, int64_t n_cols, int64_t row_begin, int64_t n_cells,
                   int32_t n_classes, const int32_t *used_idx, const int64_t *used_off, const int32_t *pairs,
                   const int64_t *pair_off, int32_t iterations, int32_t min_iter, int32_t min_pairs,
                   int32_t rng_mode, uint64_t seed, const int32_t *perm_tables, double *out_scores,
                   int32_t *out_obs_hits, int32_t *out_obs_total);

/*
 * pp_phase_scores -- single-class form with the reference kernel's exact
 * argument meaning: get_phase_scores(matrix, iterations, min_iter, min_pairs,
 * pairs, used) (pypairs/tools/cyclone.py:253-257).  used_mask is uint8[n_cols].
 */
int pp_phase_scores(pp_ctx *ctx, const void *x, int x_dtype, int x_is_device, int64_t n_rows,
                    int64_t n_cols, int64_t ld, int32_t iterations, int32_t min_iter, int32_t min_pairs,
                    const int64_t *pairs, int64_t n_pairs, const uint8_t *used_mask, double *out_scores);

/*
 * ---- multi-GPU: one process per GPU, one communicator per pp_ctx --------------------------------
 *
 * The reference is single-process (Numba threads only); these entry points are what
 * BASELINE.json's multi-GPU configurations add.  Both paths shard without a data-path exchange -- sandbag over
 * gene-pair tiles (pypairs/tools/sandbag.py:168-169, the g1 x g2 loop nest), cyclone over cells
 * (pypairs/tools/cyclone.py:223-224, the gufunc's loop over rows) -- and the library itself issues, on the ctx
 * stream, the few NCCL collectives around them (see pp_sandbag_dist / pp_cyclone_dist).  NCCL is resolved at run
 * time with dlopen (pp_nccl_load; default: the libnccl.so.2 already mapped into the process).
 *
 * Bootstrap: rank 0 calls pp_comm_unique_id(), the host sends the PP_COMM_ID_BYTES bytes to every rank over any
 * channel it has, every rank calls pp_comm_init() (collective, blocking).  All ranks of a communicator must then
 * make the same sequence of pp_*_dist calls with consistent arguments; a rank that fails validation before the
 * first collective of a call leaves the others waiting, exactly as with NCCL itself.
 */
#define PP_COMM_ID_BYTES 128
int pp_nccl_load(const char *path /* NULL or "": "libnccl.so.2" from the process / loader path */);
int pp_nccl_version(void); /* NCCL_VERSION_CODE of the loaded library, 0 if none */
int pp_comm_unique_id(uint8_t *out_id /* [PP_COMM_ID_BYTES] */);
int pp_comm_init(pp_ctx *ctx, const uint8_t *id /* [PP_COMM_ID_BYTES] */, int32_t rank, int32_t world);
int pp_comm_destroy(pp_ctx *ctx);
int pp_comm_info(pp_ctx *ctx, int32_t *rank, int32_t *world); /* world = 0: no communicator */

/*
 * pp_sandbag_dist / pp_sandbag_csr_dist -- pp_sandbag over all ranks of the ctx's communicator.
 *
 * Every rank passes the SAME arguments (the whole matrix -- host, device or CSR -- and the whole cell / gene /
 * class description, as for pp_sandbag); the work is split inside:
 *   - the rows of x are cut into `world` contiguous blocks holding equal numbers of kept cells; a rank copies
 *     (host input), scans for unexpressed genes and ranks ONLY its own block;
 *   - all-reduce(max) of the expressed-gene flags and of (max rank, NaN flag, per-gene rank width);
 *   - every rank bit-slices its cells into its region of the plane buffer, one all-gather completes it;
 *   - gene-pair tiles t with t % world == rank are evaluated (as pp_sandbag with shard = rank);
 *   - the compacted keys of all ranks are all-gathered and sorted on the device.
 * gather_to < 0: every rank receives the complete, sorted marker list; gather_to = r: only rank r does (the others
 * get *out_n = 0 and skip the device -> host copy).  out_kept_genes / probes are filled on every rank.
 * The result is identical to pp_sandbag on one GPU.
 */
int pp_sandbag_dist(pp_ctx *ctx, const void *x, int x_dtype, int x_is_device, int64_t n_rows, int64_t n_cols,
                    int64_t ld, const int32_t *cell_rows, int64_t n_cells, const int64_t *class_sizes,
                    int32_t n_classes, const int32_t *gene_cols, int64_t n_genes, const int64_t *thresholds,
                    int32_t gather_to, uint64_t **out_keys, int32_t **out_class, int64_t *out_n,
                    const int64_t *probe_pairs, int64_t n_probe, int32_t *probe_pos, int32_t *probe_neg,
                    int32_t drop_unexpressed, int32_t **out_kept_genes, int64_t *out_n_kept);
int pp_sandbag_csr_dist(pp_ctx *ctx, const pp_csr *x, int64_t n_rows, int64_t n_cols, const int32_t *cell_rows,
                        int64_t n_cells, const int64_t *class_sizes, int32_t n_classes, const int32_t *gene_cols,
                        int64_t n_genes, const int64_t *thresholds, int32_t gather_to, uint64_t **out_keys,
                        int32_t **out_class, int64_t *out_n, const int64_t *probe_pairs, int64_t n_probe,
                        int32_t *probe_pos, int32_t *probe_neg, int32_t drop_unexpressed,
                        int32_t **out_kept_genes, int64_t *out_n_kept);

/*
 * pp_cyclone_dist / pp_cyclone_csr_dist -- pp_cyclone on this rank's cell shard, then one all-gather of the score
 * columns (pypairs/tools/cyclone.py:151-170: the per-class score vectors that become the DataFrame columns).
 *
 * x, row_begin, n_cells describe THIS rank's shard exactly as in pp_cyclone (x may be the whole matrix with a row
 * range, or just the rank's own rows with row_begin = 0); marker tables, iterations, rng are the same on all ranks.
 * shard_cells[world] = n_cells of every rank (rank order = cell order of the result).
 * gather_to < 0: every rank receives out_* of shape [n_classes, sum(shard_cells)]; gather_to = r: only rank r does
 * (the out_* pointers of the other ranks may be NULL).  want_observed != 0 (the same on all ranks) also gathers the
 * observed counters.  A NaN on any rank fails the call on every rank (PP_ENAN).
 */
int pp_cyclone_dist(pp_ctx *ctx, const void *x, int x_dtype, int x_is_device, int64_t n_rows, int64_t n_cols,
                    int64_t ld, int64_t row_begin, int64_t n_cells, int32_t n_classes, const int32_t *used_idx,
                    const int64_t *used_off, const int32_t *pairs, const int64_t *pair_off, int32_t iterations,
                    int32_t min_iter, int32_t min_pairs, int32_t rng_mode, uint64_t seed,
                    const int32_t *perm_tables, const int64_t *shard_cells, int32_t gather_to,
                    int32_t want_observed, double *out_scores, int32_t *out_obs_hits, int32_t *out_obs_total);
int pp_cyclone_csr_dist(pp_ctx *ctx, const pp_csr *x, int64_t n_rows, int64_t n_cols, int64_t row_begin,
                        int64_t n_cells, int32_t n_classes, const int32_t *used_idx, const int64_t *used_off,
                        const int32_t *pairs, const int64_t *pair_off, int32_t iterations, int32_t min_iter,
                        int32_t min_pairs, int32_t rng_mode, uint64_t seed, const int32_t *perm_tables,
                        const int64_t *shard_cells, int32_t gather_to, int32_t want_observed, double *out_scores,
                        int32_t *out_obs_hits, int32_t *out_obs_total);

/*
 * pp_perm_table -- the permutation stream on its own (validation / export):
 * int32[iterations, n] for class `class_index` (ignored by MT19937, whose
 * stream depends on (seed, n) only; numba/cpython/randomimpl.py:1929-1946).
 */
int pp_perm_table(pp_ctx *ctx, int32_t rng_mode, uint64_t seed, int32_t class_index, int32_t n,
                  int32_t iterations, int32_t *out_table);

/*
 * pp_dense_rank -- the prep stage on its own (tests, profiling): per kept cell,
 * dense rank (ties share a rank, +0.0 == -0.0) of the kept genes.
 * out_rank: host uint16[n_cells, n_genes]; out_max_rank: max over the matrix.
 */
int pp_dense_rank(pp_ctx *ctx, const void *x, int x_dtype, int x_is_device, int64_t n_rows, int64_t n_cols,
                  int64_t ld, const int32_t *cell_rows, int64_t n_cells, const int32_t *gene_cols,
                  int64_t n_genes, uint16_t *out_rank, int32_t *out_max_rank);

/*
 * pp_score_labels -- the label columns of the result frame (pypairs/tools/cyclone.py:172-181) from the score
 * columns, one host pass (no device work, no ctx):
 *   out_max_class[i] = first class with the largest score of cell i, NaN skipped (DataFrame.idxmax), or n_classes
 *                      ("N/A") unless the NaN-skipping sum of the cell's scores is > 0 (cyclone.py:172-173);
 *   out_cc[i]        (only when idx_g1 >= 0 and idx_g2m >= 0; cyclone.py:176-181) = 2 ("S") if max(G1, G2M) < 0.5,
 *                      else 0 ("G1") / 1 ("G2M") = first maximum of the two; -1 if both are NaN.
 * scores: double[n_classes, n_cells] as returned by pp_cyclone.
 */
int pp_score_labels(const double *scores, int32_t n_classes, int64_t n_cells, int32_t idx_g1, int32_t idx_g2m,
                    int32_t *out_max_class, int32_t *out_cc);

/*
 * pp_alu_peak -- roofline denominator for the bit-sliced / packed-integer kernels: measured
 * throughput of independent 32-bit LOP3 instructions on this device, in lane-operations per
 * second (threads x instructions / time), best of `repeats` launches.  Also returns the SM
 * count.  (MEASURED_PEAKS.json only carries HBM and tensor peaks.)
 */
int pp_alu_peak(pp_ctx *ctx, int32_t repeats, double *out_lop3_lane_ops_per_s, int32_t *out_n_sm);

#ifdef __cplusplus
}
#endif
#endif /* PYPAIRS_B200_H */
