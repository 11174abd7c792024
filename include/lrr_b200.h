/*
 * lrr_b200 -- C ABI of the B200-native per-variant linear regression (Hail `linear_regression_rows`).
 *
 * This is the drop-in boundary for ONE reference path.  What each entry point replaces:
 *
 *   reference plugin           abstract class MatrixToTableFunction { typ; execute(ctx, MatrixValue): TableValue }
 *                              hail/hail/src/is/hail/expr/ir/functions/RelationalFunctions.scala:24-32
 *   lrr_add_group              the driver prologue's broadcasts (completeColIdx, y, Qt, Qty, yyp)
 *                              hail/hail/src/is/hail/methods/LinearRegression.scala:47-78 (Single), :228-257 (Chained)
 *   lrr_run                    the per-partition hot loop: setMeanImputedDoubles + block algebra + T.cumulative
 *                              LinearRegression.scala:95-193 / :274-402, stats/RegressionUtils.scala:16-58
 *   lrr_pack_bed / _dosage_i8  the upstream entry decode that feeds x (io/plink/LoadPlink.scala:470-530 codes;
 *                              `GT.n_alt_alleles()` as the entry expression, methods/statgen.py:387-392)
 *
 * FFI precedent in the reference (caller-owned buffers, raw addresses, no exceptions across the ABI):
 *   hail/hail/src/is/hail/methods/IBSFFI.scala:14-23  <->  hail/c/ibs.cpp:105-107   (JNA direct mapping)
 *   hail/hail/src/is/hail/linalg/BLAS.scala:116-143                                  (raw Long addresses)
 *   hail/hail/resources/include/hail/NativeStatus.h:12-37                            (errno + message)
 *
 * Conventions: every function returns 0 on success or a non-zero LRR_E* code; the message is available
 * from lrr_last_error().  All `d_*` pointers are DEVICE pointers on the context's device.  A context is
 * bound to one device and is not thread-safe; distinct contexts are independent.  `stream` is a
 * cudaStream_t passed as void* (NULL = default stream).  No allocation happens inside lrr_run once
 * lrr_reserve() has been called for the largest M.
 */
#ifndef LRR_B200_H
#define LRR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LRR_OK 0
#define LRR_EINVAL 1   /* bad argument */
#define LRR_ECUDA 2    /* CUDA runtime error (message has the CUDA string) */
#define LRR_ESTATE 3   /* call order / missing groups */
#define LRR_ENOMEM 4

/* kernels selectable in lrr_run.  AUTO: FP64 for tiny inputs (n_variants * n_samples <= 2.5e8), else TC4 (as many
 * sweeps as the digit columns need) when its exactness bound holds (<= 699k samples guaranteed), else TC, else FP64. */
#define LRR_KERNEL_AUTO 0
#define LRR_KERNEL_FP64 1  /* CUDA-core float64 reference-order kernel */
#define LRR_KERNEL_TC 2    /* tcgen05 int8-sliced exact-integer kernel (INT32 accumulators) */
#define LRR_KERNEL_TC4 3   /* tcgen05 4-bit (E2M1 digits) exact-integer kernel (f32 accumulators within 2^24) */

typedef struct lrr_ctx lrr_ctx;

/* per-group device output pointers; any pointer may be NULL to skip that field.  Arrays of
 * phenotype-indexed fields are [M, P] row-major (row = variant), matching the reference's
 * array<float64> of length P per row (LinearRegression.scala:26-34, 173-186). */
typedef struct {
  int32_t* n;             /* [M]   number of columns used (same for every row)       LR:170 */
  int32_t* n_missing;     /* [M]   missing calls among the group's samples (extension; RU:51) */
  double* sum_x;          /* [M]   LR:136,171 */
  double* y_transpose_x;  /* [M,P] LR:143 */
  double* beta;           /* [M,P] LR:150-155 */
  double* standard_error; /* [M,P] LR:157 */
  double* t_stat;         /* [M,P] LR:159 */
  double* p_value;        /* [M,P] LR:160 */
  double* log10_p;        /* [M,P] optional: log10 of p_value, finite below 1e-308 (extension) */
} lrr_group_out;

const char* lrr_version(void);
int lrr_create(lrr_ctx** out, int device);
void lrr_destroy(lrr_ctx* ctx);
/* last error message of `ctx` (or of the failed lrr_create when ctx == NULL) */
const char* lrr_last_error(const lrr_ctx* ctx);

/* ---- ingest: the device genotype store ---------------------------------------------------------
 * 2 bits per call, variant-major rows of lrr_packed_stride(N) bytes (a multiple of 128).  Codes:
 * 0,1,2 = number of alternate alleles, 3 = missing.  Within each little-endian 32-bit word, which
 * covers samples 16w..16w+15, sample 16w+4s+i sits at bits [8i+2s, 8i+2s+1] (so that
 * (word >> 2s) & 0x03030303 yields four consecutive samples as four bytes).  Padding samples are 0.
 * Every writer of the store can also emit the "missing mask" side array d_row_flags [n_variants] (uint8,
 * may be NULL): 1 iff the row holds at least one missing call.  lrr_run uses it to skip the
 * missing-indicator plane for tiles without missing calls; passing NULL there is always correct
 * (every tile is then treated as possibly missing). */
int64_t lrr_packed_stride(int64_t n_samples);
/* PLINK .bed SNP-major rows (no 3-byte header), a2_reference=True semantics (LoadPlink.scala:475-481) */
int lrr_pack_bed(lrr_ctx* ctx, const uint8_t* d_bed, int64_t n_variants, int64_t bed_stride, int64_t n_samples,
                 uint8_t* d_packed, int64_t packed_stride, uint8_t* d_row_flags, void* stream);
/* int8 dosages [M, N] row-major: 0/1/2, anything else (e.g. -1) = missing */
int lrr_pack_dosage_i8(lrr_ctx* ctx, const int8_t* d_dosage, int64_t n_variants, int64_t n_samples,
                       uint8_t* d_packed, int64_t packed_stride, uint8_t* d_row_flags, void* stream);
/* inverse of lrr_pack_dosage_i8 (missing -> -1); for tests and export */
int lrr_unpack_dosage_i8(lrr_ctx* ctx, const uint8_t* d_packed, int64_t packed_stride, int64_t n_variants,
                         int64_t n_samples, int8_t* d_dosage, void* stream);
/* device store -> PLINK .bed SNP-major rows (inverse of lrr_pack_bed; the bytes ExportPlink would write,
 * hail/hail/src/is/hail/expr/ir/MatrixWriter.scala:2270-2285).  bed_stride >= ceil(n_samples/4). */
int lrr_unpack_bed(lrr_ctx* ctx, const uint8_t* d_packed, int64_t packed_stride, int64_t n_variants, int64_t n_samples,
                   uint8_t* d_bed, int64_t bed_stride, void* stream);
/* seeded Balding-Nichols style synthetic fill (methods/statgen.py:4182-4291 distributionally).  The host
 * draws per-variant per-population allele frequencies and turns them into integer thresholds
 * d_thresholds [M, n_pops, 3] (uint32, 16-bit scale, 0..65536): with u = 16 counter-based random bits,
 * u < t0 -> missing, u < t1 -> 0 alt, u < t2 -> 1 alt, else 2 alt (call ~ Cat(q^2, 2pq, p^2), SG:4290-4291).
 * d_pop [N] (uint8) is each sample's population (SG:4254).  The value of call (first_variant + r, j)
 * depends only on (seed, first_variant + r, j), so any variant range can be regenerated anywhere. */
int lrr_bn_fill(lrr_ctx* ctx, const uint32_t* d_thresholds, int n_pops, const uint8_t* d_pop, int64_t n_variants,
                int64_t first_variant, int64_t n_samples, uint64_t seed, uint8_t* d_packed, int64_t packed_stride,
                uint8_t* d_row_flags, void* stream);

/* ---- basis: one call per group of phenotypes (Single = 1 group, Chained = G groups) -------------
 * The host has selected the group's complete samples (RU:88-128), computed an orthonormal basis Q of
 * the covariates (LR:65-69) and residualised y against it.  Arrays may be host or device memory.
 *   complete_idx [n]     ascending sample indices kept (RU:116-127)
 *   q_cols  [Kd, n]      the Kd dot-product columns of Q restricted to the kept samples, row-major
 *                        by column; Kd = K - has_intercept.  When has_intercept != 0 the caller has
 *                        rotated Q so that its first column is the constant 1/sqrt(n) (dropped here:
 *                        its projection is formed exactly from integer genotype counts) and the other
 *                        K-1 columns are orthogonal to it.
 *   y_res   [P, n]       y - Q Q^T y, row-major by phenotype
 *   qty     [K, P]       Q^T y (full K rows, row 0 = the constant column when has_intercept), LR:71
 *   yyp     [P]          y.y - Qty.Qty, LR:78
 * Fails with LRR_EINVAL when n - K - 1 < 1 (LR:55-58). */
int lrr_clear_groups(lrr_ctx* ctx);
int lrr_add_group(lrr_ctx* ctx, int64_t n_samples_total, int32_t n, int32_t K, int32_t P, int32_t has_intercept,
                  const int32_t* complete_idx, const double* q_cols, const double* y_res, const double* qty,
                  const double* yyp);
/* Weighted least squares group (`weights=` of hl.linear_regression_rows; the reference routes it through
 * _linear_regression_rows_nd, methods/statgen.py:557-581, 636-660): x is mean-imputed first, then every quantity is
 * scaled by sqrt(w).  The host passes the sqrt(w)-SCALED arrays: Q = orthonormal basis of sqrt(w) * covariates,
 * q_cols [K, n] = Q columns TIMES sqrt(w) again (so that a dot product with the unscaled imputed x is Q^T (sqrt(w) x)),
 * y_res [P, n] = (sqrt(w) y - Q Q^T sqrt(w) y) TIMES sqrt(w), qty / yyp from the scaled y, and sqrt_w [n].
 * `sum_x` is then the column sum of the scaled x (statgen.py:646).  No intercept shortcut applies; weighted groups run
 * on the float64 kernel (x.x = sum w x^2 is not linear in the call codes). */
int lrr_add_group_weighted(lrr_ctx* ctx, int64_t n_samples_total, int32_t n, int32_t K, int32_t P,
                           const int32_t* complete_idx, const double* q_cols, const double* y_res, const double* qty,
                           const double* yyp, const double* sqrt_w);
int lrr_num_groups(const lrr_ctx* ctx);

/* ---- the hot call ------------------------------------------------------------------------------ */
/* pre-size internal workspaces for up to max_variants rows per lrr_run call */
int lrr_reserve(lrr_ctx* ctx, int64_t max_variants);
/* regress `n_variants` packed rows against every group; outs[g] receives group g's fields */
int lrr_run(lrr_ctx* ctx, const uint8_t* d_packed, const uint8_t* d_row_flags, int64_t n_variants, int64_t packed_stride,
            int64_t n_samples_total, const lrr_group_out* outs, int32_t n_outs, int32_t kernel, void* stream);
/* The same hot call for a DENSE float64 x: `x` may be any entry-indexed float64 expression in the reference
 * (methods/statgen.py:229, 391; e.g. PL / GP dosages, test_statgen.py:286-364), not only GT.n_alt_alleles().
 * d_x is [n_variants, ldx] row-major on the device, one double per entry over ALL n_samples_total columns, NaN =
 * missing (mean-imputed per group as RU:16-58).  Float64 CUDA-core kernel; weighted groups (lrr_add_group_weighted) too. */
int lrr_run_dense(lrr_ctx* ctx, const double* d_x, int64_t n_variants, int64_t ldx, int64_t n_samples_total,
                  const lrr_group_out* outs, int32_t n_outs, void* stream);
/* Tolerance guard of the quantised (TC / TC4) sweeps.  The basis columns are stored as exact digits of a fixed-point
 * value, so a dot product is off by at most (quantum / 2) * sum_j x_j for ANY genotype row; the per-variant epilogue
 * propagates that bound through xxp, beta, standard_error, t_stat, p_value and y_transpose_x and lists every row whose
 * bound leaves half of the parity tolerance (relative 1e-6, p-values 1e-5; an absolute |dt| <= 5e-10 also passes, 5e-7
 * for groups of more than two phenotypes).  Listed rows are recomputed in float64 inside the same lrr_run (asynchronous,
 * same stream).  lrr_last_recomputed returns how many rows of the last lrr_run were recomputed, summed over the groups
 * (it synchronises on that run).  Precision is adaptive and deterministic: on the first lrr_run after the groups
 * changed, the first 8,192 rows run as a pilot and, when more than 2 % of them were listed (structured or badly scaled
 * covariates), the covariate columns are re-quantised with two more base-13 digits, up to three times.
 * lrr_set_guard(ctx, 0) switches the guard off (kernel tuning, tests of the raw quantised path). */
int lrr_set_guard(lrr_ctx* ctx, int enabled);
int64_t lrr_last_recomputed(lrr_ctx* ctx);
/* The dense call for a COMPACT dosage store (the 8 / 16-bit dosages of imputed data, io/bgen/): one
 * uint16 per entry, entry value = q * scale, q = 0xFFFF = missing.  BGEN's 8-bit genotype probabilities give dosages that
 * are exact multiples of 1/255 (q <= 510, scale = 1/255); any other float dosage is stored to 2/65534 (scale = 2/65534).
 * d_xq is [n_variants, ldx] row-major with ldx a multiple of 8 entries and 16-byte aligned rows.  2 bytes of HBM per entry
 * instead of 8; the results are those of lrr_run_dense on the dequantised values q * scale, bit for bit. */
int lrr_run_dense_u16(lrr_ctx* ctx, const uint16_t* d_xq, int64_t n_variants, int64_t ldx, int64_t n_samples_total, double scale,
                      const lrr_group_out* outs, int32_t n_outs, void* stream);
/* number of kernel launches issued by this context since creation (for bench accounting) */
int64_t lrr_launch_count(const lrr_ctx* ctx);
/* which kernel LRR_KERNEL_AUTO resolved to on the last lrr_run */
int lrr_last_kernel(const lrr_ctx* ctx);

/* measurement hook: when enabled, lrr_run brackets its sweep kernel(s) -- not the per-variant epilogue -- with
 * CUDA events on `stream`; lrr_last_sweep_ms synchronises on them and returns the elapsed milliseconds of the
 * last lrr_run's sweep (negative if none was recorded). */
int lrr_set_timing(lrr_ctx* ctx, int enabled);
float lrr_last_sweep_ms(lrr_ctx* ctx);
/* measurement hook: shape of the last lrr_run's tensor-core sweep(s) -- out4[0] = sweep launches (passes x planes),
 * out4[1] = MMA columns N summed over those launches (padded to 16), out4[2] = the part of out4[1] issued by launches whose
 * tile pairs ALSO run the missing-indicator plane when they hold a missing call (narrow two-plane sweeps; split wide passes
 * count each plane as its own launch instead), out4[3] = digit columns in use, summed like out4[1].  Tensor work of the run =
 * variants x padded samples x (out4[1] + out4[2] x share of two-plane tiles) multiply-adds.  All zero after a float64 run. */
int lrr_last_sweep_shape(const lrr_ctx* ctx, int64_t* out4);

/* ---- the hot call on HOST-resident input: the streaming loop ------------------------------------------
 * The reference's loop consumes each partition's rows as they are decoded from storage (LinearRegression.scala:95,
 * io/plink/LoadPlink.scala:470-530).  lrr_stream_* is that loop for a SNP-major PLINK .bed body (no 3-byte
 * header) in host memory -- page-locked memory gives the full PCIe rate: blocks of `block_variants` rows are
 * copied to the device, packed, swept and their result rows copied back to host arrays, copies and compute
 * overlapped on internal streams over a ring of `depth` device slots (0 = defaults: ~256 MB blocks, up to 16 GB
 * in flight).  lrr_stream_begin needs only the genotype bytes and returns at once, so it can be called BEFORE
 * the groups exist: the first `depth` blocks are copied and packed while the host runs the driver prologue
 * (LR:47-78).  lrr_stream_run takes lrr_group_out structs of HOST pointers ([M] / [M, P] arrays as in lrr_run,
 * NULL = skip) and returns when every result row is in host memory.  A stream runs once; lrr_stream_end frees
 * it (always call it, also after an error). */
typedef struct lrr_stream lrr_stream;
int lrr_stream_begin(lrr_ctx* ctx, lrr_stream** out, const uint8_t* h_bed, int64_t n_variants, int64_t bed_stride,
                     int64_t n_samples, int64_t block_variants, int32_t depth);
int lrr_stream_run(lrr_ctx* ctx, lrr_stream* stream, const lrr_group_out* h_outs, int32_t n_outs, int32_t kernel);
void lrr_stream_end(lrr_ctx* ctx, lrr_stream* stream);
/* measurement hook: device time from the start of the first block's host-to-device copy to the end of the last one, of
 * the last lrr_stream_run (milliseconds; negative if none) -- what the host link gave the call */
float lrr_last_stream_h2d_ms(const lrr_ctx* ctx);
/* with lrr_set_timing(ctx, 1): the block timeline of the last lrr_stream_run, three floats per block (milliseconds since
 * the first copy started: copy done, sweep started, statistics done); copies up to `capacity` floats, returns the total */
int lrr_last_stream_timeline(const lrr_ctx* ctx, float* out, int capacity);
/* The streaming arena (staging buffers + slots, up to 16 GB and never more than half of the free device memory) and the
 * dense path's missing-bit plane stay cached on the context between calls, and so do the buffers of retired groups, the
 * workspaces of the hot call and the host staging buffer (that is what keeps lrr_clear_groups / lrr_add_group / lrr_run
 * free of cudaMalloc and device synchronisation); lrr_trim gives all of it back (synchronises; LRR_ESTATE while a stream
 * is open).  lrr_destroy frees everything. */
int lrr_trim(lrr_ctx* ctx);

/* ---- logistic regression, score test (`hl.logistic_regression_rows(test='score', ...)`) --------
 * Replaces the per-row loop of hail/hail/src/is/hail/methods/LogisticRegression.scala:115-157 for
 * LogisticScoreTest (stats/LogisticRegressionModel.scala:211-264) on the same ingest, complete-sample and
 * mean-imputation logic as the linear path.  The host fits the null model (covariates only, LogisticRegression.scala:67-91)
 * and passes, for the n complete samples (ascending complete_idx): wc [K, n] = w_j c_jk with w = mu (1 - mu),
 * resid [n] = y - mu, w [n], finv [K, K] = inverse of the null Fisher information, score0 [K] = null score.
 * lrr_set_score_model replaces any groups of the context; lrr_run_score sweeps packed rows (float64 kernel: the
 * x'Wx term is not linear in the call codes) and writes chi_sq_stat / p_value [M] (NaN where the reference yields
 * missing: x in the span of the covariates) and optionally n_missing [M]. */
typedef struct {
  double* chi_sq_stat;
  double* p_value;
  int32_t* n_missing;
} lrr_score_out;
int lrr_set_score_model(lrr_ctx* ctx, int64_t n_samples_total, int32_t n, int32_t K, const int32_t* complete_idx,
                        const double* wc, const double* resid, const double* w, const double* finv, const double* score0);
int lrr_run_score(lrr_ctx* ctx, const uint8_t* d_packed, const uint8_t* d_row_flags, int64_t n_variants, int64_t packed_stride,
                  int64_t n_samples_total, const lrr_score_out* out, void* stream);
/* the same test for a dense float64 x ([n_variants, ldx] on the device, NaN = missing; pl_dosage / gp_dosage inputs) */
int lrr_run_score_dense(lrr_ctx* ctx, const double* d_x, int64_t n_variants, int64_t ldx, int64_t n_samples_total,
                        const lrr_score_out* out, void* stream);

/* ---- logistic regression, Wald / likelihood-ratio / Firth tests (`hl.logistic_regression_rows(test='wald'|'lrt'|'firth')`)
 * The per-row loop of hail/hail/src/is/hail/methods/LogisticRegression.scala:115-157 with WaldTest, LikelihoodRatioTest and
 * LogisticFirthTest (stats/LogisticRegressionModel.scala:55-199; the Newton fits :294-408): one CTA per variant iterates
 * in float64 on the mean-imputed genotype column.  The host fits the null model once (LogisticRegression.scala:67-91)
 * and passes, for the n complete samples (ascending complete_idx): cov [K, n], y [n] (0 / 1), the null coefficients
 * b0 [K], and the null fit's last score [K], Fisher matrix [K, K] and log-likelihood (the first Newton step of the full
 * model reuses them, LogisticRegressionModel.scala:311-325).  K <= 63 (up to 19 covariates the per-thread register form,
 * above that the Fisher matrix is accumulated as 4 x 4 register blocks over shared-memory tiles of the design matrix).
 * lrr_run_logit writes, per variant, the fields of the test's schema (`standard_error`, `z_stat`: Wald only;
 * `chi_sq_stat`: LRT / Firth only; NaN where the reference leaves them missing: fit not converged or singular) and the
 * `fit` struct (n_iterations, converged, exploded).  NULL output pointers are skipped. */
#define LRR_LOGIT_WALD 1
#define LRR_LOGIT_LRT 2
#define LRR_LOGIT_FIRTH 3
typedef struct {
  double* beta;
  double* standard_error;
  double* z_stat;
  double* chi_sq_stat;
  double* p_value;
  int32_t* n_iterations;
  uint8_t* converged;
  uint8_t* exploded;
} lrr_logit_out;
int lrr_set_logit_model(lrr_ctx* ctx, int64_t n_samples_total, int32_t n, int32_t K, const int32_t* complete_idx,
                        const double* cov, const double* y, const double* b0, const double* score0, const double* fisher0,
                        double loglik0);
int lrr_run_logit(lrr_ctx* ctx, const uint8_t* d_packed, int64_t n_variants, int64_t packed_stride, int64_t n_samples_total,
                  int32_t test, int32_t max_iterations, double tolerance, const lrr_logit_out* out, void* stream);
/* the same fits for a dense float64 x ([n_variants, ldx] on the device, NaN = missing): pl_dosage / gp_dosage inputs,
 * hail/python/test/hail/methods/test_statgen.py:851-938 */
int lrr_run_logit_dense(lrr_ctx* ctx, const double* d_x, int64_t n_variants, int64_t ldx, int64_t n_samples_total, int32_t test,
                        int32_t max_iterations, double tolerance, const lrr_logit_out* out, void* stream);

/* two-sided Student-t p-value on the device, exposed for unit tests of the epilogue:
 * p[i] = 2 * P[T_df <= -|t[i]|]  (jdistlib T.cumulative call sites LR:160, LR:344) */
int lrr_student_t_two_sided(lrr_ctx* ctx, const double* d_t, int64_t count, double df, double* d_p, double* d_log10_p,
                            void* stream);
/* chi2[i] = qchisqtail(p[i], 1), the 1-d.o.f. chi-squared statistic whose upper tail is p: the per-row half of
 * `hl.lambda_gc` (hail/python/hail/methods/statgen.py:3096-3128), a downstream consumer of `p_value` */
int lrr_qchisqtail1(lrr_ctx* ctx, const double* d_p, int64_t count, double* d_chi2, void* stream);

/* ---- PCA building block (`hl.hwe_normalized_pca`, hail/python/hail/methods/pca.py:15-33, 345-372) ----
 * The power iteration G <- A' (A G) needs, next to the per-variant sweep A G (lrr_run with the columns of G as
 * phenotypes: y_transpose_x), the TRANSPOSED product over the same packed rows:
 *     out[s][j][c] = sum over the variants v of split s of  coef[v][code(v, j)] * t[v][c]
 * coef [n_variants][4] tabulates the entry value of call codes 0, 1, 2 and missing per variant (HWE normalisation:
 * (code - mean_v) / sd_v and 0), t is [n_variants][L] (L <= 24), out is [n_splits][n_samples_total][L]; the variant range
 * is cut into n_splits contiguous parts whose partial sums the caller adds (deterministic). */
int lrr_at_times(lrr_ctx* ctx, const uint8_t* d_packed, int64_t n_variants, int64_t packed_stride, int64_t n_samples_total,
                 const double* d_coef, const double* d_t, int32_t L, int32_t n_splits, double* d_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LRR_B200_H */
