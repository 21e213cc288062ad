/*
 * vgs.h -- C ABI of libvgs.so: B200-native (sm_100a) all-pairs VLMC distance matrices.
 *
 * This is the drop-in boundary for the one hot path of Kalior/clustering-genomic-signatures that
 * this repository accelerates: everything a `d.distance(left_vlmc, right_vlmc)` call does, for all
 * (left, right), plus the matrix hand-off to the clustering code.  Each entry point names the
 * reference interface it replaces (paths relative to the reference repository root).  The reference
 * is Python/Cython, so the reference-side binding is a ctypes stub.
 *
 * Conventions
 *   - plain C types only; every size / index is int64_t unless noted;
 *   - pointers are DEVICE pointers unless the parameter name ends in `_h` (host memory; pinned
 *     host memory makes the copies asynchronous);
 *   - `stream` is a cudaStream_t passed as void* (e.g. torch.cuda.current_stream().cuda_stream);
 *     functions taking a stream only enqueue work on it; `_h` functions synchronise before return;
 *   - every function returns a vgs_status; vgs_last_error() gives the message (thread local);
 *   - a handle owns device copies of the model tables and sequences it was given; the caller owns
 *     every buffer it passes in.  Calls on one handle must be serialised by the caller.
 *   - there is no CPU fallback: without a CUDA device vgs_create fails with VGS_ERR_CUDA.
 *
 * Encoding (same as the Python `flat` module): symbols A,C,G,T = 0..3 (src/vlmc/vlmc.pyx:21); a
 * context of length l is (len, code) with two bits per symbol, NEWEST symbol in the two lowest bits;
 * ctx_id = position of the context in the reference's `tree` dict iteration order; packed sequences
 * hold 16 symbols per uint32 with the FIRST symbol in the two HIGHEST bits, every sequence starting on
 * a 16-byte boundary.
 */
#ifndef VGS_H
#define VGS_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct vgs_context *vgs_handle;

typedef enum {
    VGS_OK = 0,
    VGS_ERR_INVALID = 1,      /* bad argument (ValueError) */
    VGS_ERR_CUDA = 2,         /* CUDA runtime failure / no device (RuntimeError) */
    VGS_ERR_NO_CONTEXT = 3,   /* a history matched no context, not even "": the reference raises
                                 RuntimeError("get_context vlmc.pyx") (src/vlmc/vlmc.pyx:141) */
    VGS_ERR_SYMBOL = 4,       /* non-ACGT symbol: KeyError in the reference (src/vlmc/vlmc.pyx:125) */
    VGS_ERR_UNSUPPORTED = 5,  /* shape outside what this build handles (e.g. order > 15) */
    VGS_ERR_MEMORY = 6        /* device or host allocation failed */
} vgs_status;

/* distance kinds for vgs_pair_* (src/distance/frobenius.pyx, src/distance/projection.pyx) */
#define VGS_PAIR_FROBENIUS_INTERSECTION 0
#define VGS_PAIR_FROBENIUS_UNION 1
#define VGS_PAIR_PROJECTION 2

/* NLL accumulation modes for vgs_nll_* */
#define VGS_NLL_SEQUENTIAL 0 /* one lane per (sequence, model), strictly left-to-right fp64: bit-exact */
#define VGS_NLL_WARP 1       /* one warp per (sequence, model), 128-bit loads, shuffle reduction     */
#define VGS_NLL_KMER 2       /* (order+1)-mer histogram per sequence x dense tables as an exact int8 contraction
                                on the tcgen05 tensor cores (ln p in 2^-43 fixed point, vgs_set_kmer_fraction_bits; fp64 MMA fallback); sets
                                with every order <= 6 and a root context, else falls back to WARP; the tile edge
                                of vgs_nll_tiles must then be a multiple of 128                           */

/* skip modes (src/vlmc/vlmc.pyx:79-85) */
#define VGS_SKIP_ORDER 0 /* log_likelihood_ignore_initial_bias: first `order` symbols not scored */
#define VGS_SKIP_NONE 1  /* log_likelihood: every symbol scored                                   */

/* ---- lifetime ------------------------------------------------------------------------------ */
int vgs_version(void);
const char *vgs_last_error(void);
int vgs_create(int device, vgs_handle *out);
int vgs_destroy(vgs_handle h);
int vgs_device_info(vgs_handle h, int *sm_count, int *cc_major, int *cc_minor, int64_t *free_bytes);
/* number of kernels this handle has launched so far (bench.py's gpu_launches) */
int64_t vgs_launch_count(vgs_handle h);
/* device-pointer entry points only enqueue work: this synchronises `stream` and reports what the kernels
 * flagged (VGS_ERR_NO_CONTEXT, VGS_ERR_SYMBOL) -- the errors the reference raises from get_context */
int vgs_poll_error(vgs_handle h, void *stream);

/* ---- per-kernel timing for bench.py's roofline: when enabled, the dominant kernel of each path is
 *      bracketed by CUDA events on its own stream; vgs_profile_read synchronises, returns the summed device
 *      time and number of launches since the last read, and resets.  kernel ids: VGS_PROF_* ------------ */
#define VGS_PROF_NLL_SCORES 0
#define VGS_PROF_PAIR_PARTIALS 1
#define VGS_PROF_ESTIMATE_PAIRS 2
#define VGS_PROF_ESTIMATE_COUNTS 3
int vgs_profile_enable(vgs_handle h, int on);
int vgs_profile_read(vgs_handle h, int kernel_id, double *total_ms, int64_t *launches);

/* ---- model tables: replaces VLMC.__init__/from_json table construction (src/vlmc/vlmc.pyx:8-60,
 *      :179-180).  prob_h is [total][4] fp64 in A,C,G,T order, bit-identical to the dict values.  The
 *      library derives on the host ln p with libm log() (== CPython math.log, src/vlmc/vlmc.pyx:99;
 *      -1000.0 where p == 0, :94-97) unless logp_h is given, and on the device the per-model hash, the
 *      longest-suffix maps and the fp32 rows. ---------------------------------------------------- */
int vgs_load_models(vgs_handle h, int64_t n_models, const int64_t *ctx_offsets_h, const uint8_t *ctx_len_h,
                    const uint32_t *ctx_code_h, const double *prob_h, const double *logp_h /* nullable */);
/* same, with flags: VGS_LOAD_SKIP_NLL skips ln p and the NLL tables (enough for Frobenius / Projection /
 * Estimate / sampling / lookups; the vgs_nll_* entry points then return VGS_ERR_INVALID) */
#define VGS_LOAD_SKIP_NLL 1
/* VGS_LOAD_NLL_ONLY: only what the log-likelihood path reads goes to the device (keys, hash, host ln p): the transition rows
 * themselves are not uploaded (half of the host->device bytes of an NLL run); Frobenius / Projection / Estimate / the
 * sampler then return VGS_ERR_INVALID.  VGS_LOAD_LAZY_TABLES: do not build the k-mer contraction's digit operand during the
 * load (default for sets with every order <= 6: built block by block behind the ln p upload); it is then built by the
 * first k-mer-mode call */
#define VGS_LOAD_NLL_ONLY 2
#define VGS_LOAD_LAZY_TABLES 4
/* VGS_LOAD_DEVICE_LOG (with the NLL tables, without logp_h): ln p is taken on the device (CUDA's log(), at most 1 ulp from
 * libm's) for the k-mer contraction, whose fixed-point quantisation of ln p (2^-43) is 64 times coarser -- the load is one pass
 * over PCIe with no host arithmetic.  The bit-exact modes (sequential / warp scan, vgs_get_logp) still see libm's ln p: it is
 * computed from the resident transition rows the first time one of them runs, and the contraction keeps its own, so no result
 * depends on the order of calls. */
#define VGS_LOAD_DEVICE_LOG 8
int vgs_load_models_ex(vgs_handle h, int64_t n_models, const int64_t *ctx_offsets_h, const uint8_t *ctx_len_h,
                       const uint32_t *ctx_code_h, const double *prob_h, const double *logp_h /* nullable */, int flags);
int vgs_model_count(vgs_handle h, int64_t *n_models, int64_t *total_contexts);
/* order (src/vlmc/vlmc.pyx:179-180), number of contexts and NLL table tier of one model */
int vgs_model_info(vgs_handle h, int64_t model, int *order, int64_t *n_ctx, int *tier);
/* copy back the host-computed ln p table [total][4] (parity hook against math.log) */
int vgs_get_logp(vgs_handle h, double *logp_out_h);

/* ---- `.tree` ingest: the text of one classifier / PST `.tree` file straight to the flat context table (row f3; replaces
 *      src/parse_trees_to_json.py:18-55 + the JSON round trip of VLMC.from_json, src/vlmc/vlmc.pyx:33-60).  Contexts come
 *      out in file order (= the reference's dict order; a repeated key overwrites its row in place); prob_out [n][4] is
 *      count / total (:45-49) or, with deltas != 0, float(numbers[14..17]) (:40-44); occurrence_out[n] is :29.  Any output
 *      pointer may be NULL (n_ctx_out alone = counting pass).  Errors mirror the reference's exceptions: VGS_ERR_INVALID
 *      for a malformed line (IndexError / ValueError / ZeroDivisionError), VGS_ERR_SYMBOL for a non-ACGT context.
 *      Host code: needs no device and no handle. ----------------------------------------------------------------- */
int vgs_parse_tree_text(const char *text, int64_t n_bytes, int deltas, int64_t capacity, uint8_t *ctx_len_out,
                        uint32_t *ctx_code_out, double *prob_out, double *occurrence_out, int64_t *n_ctx_out);

/* ---- longest-suffix lookup: VLMC.get_context / get_all_contexts (src/vlmc/vlmc.pyx:131-154).
 *      ctx_id_out_h[q] = ctx_id of the longest suffix (length <= order) of the history that is a
 *      context, or -1 (reference: RuntimeError).  all_mask_out_h[q] (nullable): bit l set <=> the
 *      length-l suffix is a context.  Runs on the device. ------------------------------------------ */
int vgs_lookup(vgs_handle h, int64_t n_queries, const int32_t *model_h, const uint32_t *hist_code_h,
               const uint8_t *hist_len_h, int32_t *ctx_id_out_h, uint32_t *all_mask_out_h);

/* ---- sequences: the strings VLMC.generate_sequence returns / log_likelihood takes
 *      (src/vlmc/vlmc.pyx:79-101, :156-163) ------------------------------------------------------- */
int vgs_load_sequences_ascii(vgs_handle h, int64_t n_seq, const uint8_t *ascii_h, const int64_t *offsets_h);
int vgs_load_sequences_packed(vgs_handle h, int64_t n_seq, const uint32_t *words_h, const int64_t *word_offsets_h,
                              const int64_t *lengths_h);
int vgs_sequence_count(vgs_handle h, int64_t *n_seq, int64_t *total_symbols);
/* copy sequence s back as symbol codes 0..3 */
int vgs_get_sequence(vgs_handle h, int64_t s, uint8_t *codes_out_h, int64_t capacity);

/* ---- sampler: VLMC.generate_sequence / generate_sequence_from / _generate_next_letter
 *      (src/vlmc/vlmc.pyx:156-177) for every loaded model at once.  mt_state_h is CPython's
 *      random.getstate()[1] (624 MT19937 words + position), advanced in place exactly as
 *      n_models * (length + pre_sample) calls of random.random() would; model m consumes the
 *      m-th block of that stream (list order, as the reference's N^2 loop does).  The sampled
 *      sequences replace the handle's sequence set. --------------------------------------------- */
int vgs_sample_sequences(vgs_handle h, int64_t length, int64_t pre_sample, uint32_t *mt_state_h);
/* VLMC.generate_sequence_from(sequence_length, context) with a non-empty context (src/vlmc/vlmc.pyx:165-172): every
 * chain starts from the history `context` (symbol codes 0..3, oldest first; a model reads its last `order` symbols,
 * :168) instead of the empty one; the context itself is not part of the output.  context_len = 0 is
 * vgs_sample_sequences. */
int vgs_sample_sequences_from(vgs_handle h, int64_t length, int64_t pre_sample, uint32_t *mt_state_h,
                              const uint8_t *context_codes_h, int64_t context_len);

/* ---- log-likelihood scores: VLMC.log_likelihood / log_likelihood_ignore_initial_bias
 *      (src/vlmc/vlmc.pyx:79-101).  M[s * ld + m] for s in [s_begin, s_end), m in [m_begin, m_end);
 *      NaN marks a failed lookup.  M_d is a device pointer to the full [n_seq][ld] matrix. ----------- */
/* which kernel the most recent vgs_nll_* call used for the scores (bench.py picks the roofline from it) */
#define VGS_NLL_KERNEL_SCAN 0   /* k_nll_scores (sequential / warp)                                         */
#define VGS_NLL_KERNEL_DMMA 1   /* k_kmer_dmma: fp64 mma.sync contraction                                   */
#define VGS_NLL_KERNEL_FMA 2    /* k_kmer_gemm: fp64 FMA contraction (VGS_KMER_GEMM=fma)                    */
#define VGS_NLL_KERNEL_I8MMA 3  /* k_kmer_i8mma: exact int8 contraction on tcgen05 tensor cores (default)   */
int vgs_last_nll_kernel(vgs_handle h);
/* k-mer contraction: ln p enters as the fixed-point integer round(ln p * 2^f) in 6..8 base-256 digit rows per model (the
 * GEMM's work is proportional to the digit count).  f = 43 by default (environment VGS_I8_FRAC_BITS): |d score| <=
 * L * 2^-44, |d distance| <= 2^-43 against the exact sum, below the fp64 scan's own rounding noise; 51 reproduces the
 * scan to 2e-14.  vgs_kmer_digit_info: the digit rows per model of the operand as built (0: not built) and f. */
int vgs_set_kmer_fraction_bits(vgs_handle h, int bits /* 35..51 */);
int vgs_kmer_digit_info(vgs_handle h, int *digits_out /* nullable */, int *fraction_bits_out /* nullable */);

int vgs_nll_scores(vgs_handle h, int64_t s_begin, int64_t s_end, int64_t m_begin, int64_t m_end, int skip_mode,
                   int mode, double *M_d, int64_t ld, void *stream);
int vgs_nll_scores_h(vgs_handle h, int skip_mode, int mode, double *M_out_h /* [n_seq][n_models] */);

/* ---- NegativeLogLikelihood.distance for all pairs (src/distance/negloglikelihood.pyx:17-26) and
 *      the float32 [N,N] matrix of src/clustering/util.pyx:23-33.  Sequence i belongs to model i.
 *      D32_d float32 [N][N]; D64_d nullable fp64 [N][N]. ------------------------------------------ */
int vgs_nll_matrix(vgs_handle h, int64_t length, int mode, float *D32_d, double *D64_d, void *stream);
int vgs_nll_matrix_h(vgs_handle h, int64_t length, int mode, float *D32_out_h, double *D64_out_h);

/* ---- FrobeniusNorm.distance (src/distance/frobenius.pyx:17-53) and Projection.distance
 *      (src/distance/projection.pyx:7-36) for all pairs; `kind` is a VGS_PAIR_* value. ------------- */
int vgs_pair_matrix(vgs_handle h, int kind, float *D32_d, double *D64_d, void *stream);
/* which kernels the most recent vgs_pair_matrix[_h] call used: 0 = CUDA-core pair kernels (sparse contexts, any depth),
 * 1 = exact integer GEMMs on the tcgen05 tensor cores over the dense universe (Frobenius intersection / Projection,
 * depth <= 7, N <= 4096, contexts filling >= 3 % of the universe; VGS_PAIR_TC=0 / 1 forces the choice) */
int vgs_last_pair_kernel(vgs_handle h);
/* the kernels vgs_pair_matrix WOULD use for `kind` on the loaded set (same 0 / 1).  A set the tensor-core path takes is cheap
 * enough (N <= 4096) that a multi-GPU job computes it on every rank instead of sharding its tiles -- the engine asks here,
 * so that one model set gives the same bits on 1, 2, 4 or 8 GPUs */
int vgs_pair_kernel_choice(vgs_handle h, int kind);
/* path of the following vgs_pair_matrix[_h] calls: -1 (default) chosen by universe density, 0 CUDA cores (<= 1e-9 of the fp64
 * oracle), 1 tensor cores wherever they apply (|d - d_oracle| <= 1.3e-7 d + 1e-9: rows quantised to 2^-31, exact integers) */
int vgs_set_pair_path(vgs_handle h, int path);
int vgs_pair_matrix_h(vgs_handle h, int kind, float *D32_out_h, double *D64_out_h);

/* ---- EstimateVLMC.distance for all (left, right) (src/distance/estimate.pyx:19-108):
 *      D[i][j] = statistic of model i's contexts counted along sequence j (asymmetric). ------------- */
int vgs_estimate_matrix(vgs_handle h, float *D32_d, double *D64_d, void *stream);
int vgs_estimate_matrix_h(vgs_handle h, float *D32_out_h, double *D64_out_h);

/* ---- multi-GPU tiling (north_star (c)): the N x N matrix is cut into T x T tiles; the upper
 *      triangle of the tile grid (symmetric kinds) or the whole grid (Estimate) is enumerated row-major
 *      as k = 0..n_tiles-1 and tile k belongs to rank k % world.  A rank computes its tiles into a
 *      payload of slots [slots][T][T] float32 (slot q <-> tile k = q * world + rank, zero padded);
 *      the payloads are all-gathered (NCCL) and vgs_assemble scatters / mirrors them into [N][N]. -- */
int vgs_tile_plan(int64_t n, int64_t tile, int world, int symmetric, int64_t *n_tiles, int64_t *slots_per_rank);
int vgs_nll_tiles(vgs_handle h, int64_t length, int mode, int64_t tile, int rank, int world, float *payload_d,
                  void *stream);
int vgs_pair_tiles(vgs_handle h, int kind, int64_t tile, int rank, int world, float *payload_d, void *stream);
int vgs_estimate_tiles(vgs_handle h, int64_t tile, int rank, int world, float *payload_d, void *stream);
int vgs_assemble(vgs_handle h, int64_t n, int64_t tile, int world, int symmetric, const float *gathered_d,
                 float *D32_d, void *stream);

/* ---- model-sharded NLL over several GPUs (north_star (c) for NegativeLogLikelihood, src/distance/negloglikelihood.pyx:17-26
 *      evaluated for all pairs as src/clustering/util.pyx:12-14 does).  Rank r loads only ITS models with vgs_load_models
 *      (local model m is model m_col0 + m of the job: host ln p, uploads and table builds shrink with the shard) and ALL
 *      sequences.  vgs_nll_scores_band scores every sequence against the shard and writes the band
 *      Mt[(m_col0 + m) * ld + s] of the TRANSPOSED fp64 score matrix into every destination in Mt_dsts_h (device
 *      pointers, this GPU's first): with peer-mapped destinations (vgs_peer_*) the int8 GEMM's epilogue stores the band
 *      straight into the other ranks' matrices over NVLink; with one destination the bands are contiguous and assemble
 *      with a single NCCL all-gather, no scatter kernel.  vgs_nll_combine turns a complete Mt into the float32 (and
 *      optionally fp64) [n][n] distance matrix; entry (i, j) needs Mt[j][i], Mt[i][j] and both diagonal scores. ------- */
int vgs_nll_scores_band(vgs_handle h, int mode, int64_t m_col0, int64_t ld, double *const *Mt_dsts_h, int n_dst, void *stream);
int vgs_nll_combine(vgs_handle h, int64_t n, const double *Mt_d, int64_t ld, int64_t length, float *D32_d, double *D64_d, void *stream);
/* one whole step of the model-sharded job in a single call: vgs_nll_scores_band into Mt_dsts_h (leading dimension n_total),
 * vgs_peer_barrier when there are peer destinations (n_dst == world; flags_h as for vgs_peer_barrier), vgs_nll_combine of
 * Mt_dsts_h[0] into D32_d */
int vgs_nll_sharded_step(vgs_handle h, int mode, int64_t m_col0, int64_t n_total, double *const *Mt_dsts_h, int n_dst,
                         uint32_t *const *flags_h /* nullable when n_dst == 1 */, int rank, int world, uint32_t epoch, int64_t length,
                         float *D32_d, void *stream);

/* ---- peer memory over NVLink (one process per GPU): vgs_peer_alloc returns a zeroed device buffer and its 64-byte CUDA
 *      IPC handle; the other ranks map it with vgs_peer_open (and unmap with vgs_peer_close); the owner frees it with
 *      vgs_peer_free.  vgs_peer_barrier: flags_h[r] is rank r's array of `world` uint32 flags (peer-mapped here, own
 *      array at index `rank`); every rank writes `epoch` to its slot in every rank's array behind the work already on
 *      `stream` and waits for all slots of its own array -- the device-side barrier of the peer-store path. ------------ */
int vgs_peer_alloc(vgs_handle h, int64_t bytes, void **dev_ptr_out, void *ipc_handle_out /* 64 bytes */);
int vgs_peer_open(vgs_handle h, const void *ipc_handle /* 64 bytes */, void **dev_ptr_out);
int vgs_peer_close(vgs_handle h, void *dev_ptr);
int vgs_peer_free(vgs_handle h, void *dev_ptr);
int vgs_peer_barrier(vgs_handle h, uint32_t *const *flags_h, int rank, int world, uint32_t epoch, void *stream);

/* ---- measurement hooks for bench.py's roofline (no reference counterpart).  vgs_tensor_peak_i8: int8 ops/s (in TOPS)
 *      of back-to-back tcgen05.mma kind::i8 M256 N256 K32 (cta_group 2) or M128 N256 K32 (cta_group 1) from resident
 *      shared memory on every SM: the measured tensor-pipe ceiling of the contraction's own instruction.
 *      vgs_g8_plan: the unit list the int8 GEMM would run for the given output rectangles {256-row A block, first B row,
 *      B rows} (host only; rows in multiples of the granule: 32, or vgs_g8_plan_gran's `gran` -- the NLL contraction
 *      with 6 / 7 digit rows per model uses 96 / 224): units_out[i] = {a_blk, b_row0, n, kb0, kbn, partial}. ------------ */
int vgs_tensor_peak_i8(vgs_handle h, int cta_group, double *tops_out, double *ms_out /* nullable */);
/* the same probe at instruction width N (multiple of 16, <= 256): the tensor pipe's rate as a function of N */
int vgs_tensor_rate_i8(vgs_handle h, int cta_group, int n_cols, double *tops_out, double *ms_out /* nullable */);
int vgs_g8_plan(const int32_t *rects_h /* [n_rects][3] */, int n_rects, int kblocks, int n_cta_pairs, int max_split,
                int32_t *units_out_h /* [capacity][6] */, int capacity, int *n_units_out, int *split_out);
int vgs_g8_plan_gran(const int32_t *rects_h /* [n_rects][3] */, int n_rects, int kblocks, int n_cta_pairs, int max_split, int gran,
                     int32_t *units_out_h /* [capacity][6] */, int capacity, int *n_units_out, int *split_out);

/* ---- matrix hand-off layouts of src/clustering/util.pyx:6-21: float32 [N*N][3] rows (i, j, d) ---- */
int vgs_matrix_to_triples(vgs_handle h, int64_t n, const float *D32_d, float *triples_d, void *stream);

/* ---- average-link clustering of the float32 [N][N] matrix on the device, with the reference's exact merge order,
 *      tie-breaking and mixed float32/float64 arithmetic (src/clustering/average_link_clustering.pyx:21-146): merges
 *      until k clusters remain.  labels_out_h[n]: clusters numbered by their smallest member; merge_dist_out_h[n-k]
 *      (nullable): the merge distances in order; n_merges_out (nullable). ------------------------------------------- */
int vgs_average_link(vgs_handle h, int64_t n, const float *D32_d, int64_t k, int32_t *labels_out_h,
                     double *merge_dist_out_h, int64_t *n_merges_out);

/* ---- passes over the finished matrix.
 *      vgs_retrieval_metrics: src/test_distance_function.py:75-79, :96-104 + src/util/distance_metrics.py:1-60.
 *      For every model i its distances to all n models (itself included) are ranked ascending, ties in list order
 *      (Python's stable sorted()); labels_h is [n_tax][n] (e.g. genus, family).  With k = number of models sharing
 *      i's label: in_top[t][i] = (# of the k nearest that share the label) / k (procent_of_taxonomy_in_top),
 *      dist_to[t][i] = mean |d| over the models sharing the label (average_distance_to_taxonomy),
 *      avg[i] = mean of row i (average_distance).  Exactly one of D32_d / D64_d is non-null. ---------------- */
int vgs_retrieval_metrics(vgs_handle h, int64_t n, const float *D32_d, const double *D64_d, int n_tax,
                          const int32_t *labels_h, double *in_top_out_h, double *dist_to_out_h, double *avg_out_h);
/* ClusteringMetrics.silhouette_metric (src/clustering/clustering_metrics.pyx:29-84) on the float32 [N][N] matrix:
 * a[i] = mean distance to the other members of i's cluster (0 for a singleton), b[i] = smallest mean distance to
 * another cluster (both rounded to float32 as the reference's C-float returns), s[i] = (b - a) / max(b, a).
 * cluster_h[n]: any integer labels.  a_out_h / b_out_h nullable. */
int vgs_silhouette(vgs_handle h, int64_t n, const float *D32_d, const int32_t *cluster_h, double *a_out_h, double *b_out_h,
                   double *s_out_h);
/* ClusteringMetrics.average_distance_std (src/clustering/clustering_metrics.pyx:168-185): population standard
 * deviation (numpy.std, fp64) of the distances between distinct members, per cluster in ascending label order;
 * 0 for singletons.  std_out_h has room for `capacity` clusters. */
int vgs_cluster_distance_std(vgs_handle h, int64_t n, const float *D32_d, const double *D64_d, const int32_t *cluster_h,
                             int64_t capacity, double *std_out_h, int64_t *n_clusters_out);

#ifdef __cplusplus
}
#endif
#endif /* VGS_H */
