/*
 * CPU oracle (plain C) for Hail's per-variant linear regression.
 *
 * TEST INFRASTRUCTURE ONLY -- not shipped, not on the product path.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * It restates (does not copy) the reference algorithm, one loop per reference statement:
 *   LR = hail/hail/src/is/hail/methods/LinearRegression.scala
 *   RU = hail/hail/src/is/hail/stats/RegressionUtils.scala
 * and, like the reference's Spark path, works on one block of `block_size` variants at a time
 * per thread (LR:95-111: one task per partition, trueGroupedIterator(rowBlockSize)).
 *
 * Input rows are PLINK .bed SNP-major bytes (io/plink/LoadPlink.scala:475-481,525 with the
 * default a2_reference=True: code 0 -> 2 alt alleles, 1 -> missing, 2 -> 1, 3 -> 0); decoding a
 * code to a float64 entry stands in for the reference's upstream `GT.n_alt_alleles()` entry map.
 *
 * Parity pin: checked against oracle/linreg_oracle.py (itself pinned to the reference's R-derived
 * goldens) in tests/test_oracle_c.py.
 *
 * Student-t: jdistlib 0.4.5 T.cumulative (third-party; port of R's pt) -> regularised incomplete
 * beta I_x(d/2, 1/2), evaluated here with the Lentz continued fraction.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static double betacf(double a, double b, double x) {
  const double tiny = 1e-300, eps = 3e-15;  // a tighter test can bounce a few ulp around 1 for hundreds of iterations
  double qab = a + b, qap = a + 1.0, qam = a - 1.0;
  double c = 1.0, d = 1.0 - qab * x / qap;
  if (fabs(d) < tiny) d = tiny;
  d = 1.0 / d;
  double h = d;
  for (int m = 1; m <= 200000; ++m) {
    double m2 = 2.0 * m;
    double aa = m * (b - m) * x / ((qam + m2) * (a + m2));
    d = 1.0 + aa * d; if (fabs(d) < tiny) d = tiny;
    c = 1.0 + aa / c; if (fabs(c) < tiny) c = tiny;
    d = 1.0 / d; h *= d * c;
    aa = -(a + m) * (qab + m) * x / ((a + m2) * (qap + m2));
    d = 1.0 + aa * d; if (fabs(d) < tiny) d = tiny;
    c = 1.0 + aa / c; if (fabs(c) < tiny) c = tiny;
    d = 1.0 / d;
    double del = d * c;
    h *= del;
    if (fabs(del - 1.0) < eps) break;
  }
  return h;
}

/* log B(a, 1/2), stable for large a (Stirling series of the gamma ratio). */
static double log_beta_half(double a) {
  if (a < 50.0) return lgamma(a) + lgamma(0.5) - lgamma(a + 0.5);
  double ia = 1.0 / a, ia2 = ia * ia;
  double series = ia * (1.0 / 8.0 + ia2 * (-1.0 / 192.0 + ia2 * (1.0 / 640.0 + ia2 * (-17.0 / 14336.0))));
  return 0.5 * log(M_PI) - 0.5 * log(a) + series;
}

/* 2 * P[T_d <= -|t|]   (LR:160) */
double lrr_oracle_two_sided_p(double t, double d) {
  if (isnan(t)) return NAN;
  if (isinf(t)) return 0.0;
  double a = 0.5 * d, b = 0.5;
  double t2d = (t / d) * t;
  double x = 1.0 / (1.0 + t2d);
  double lbeta = log_beta_half(a);
  if (x < (a + 1.0) / (a + b + 2.0)) {
    double lf = -a * log1p(t2d) + b * log(t2d / (1.0 + t2d)) - log(a) - lbeta;
    return exp(lf) * betacf(a, b, x);
  }
  double xc = t2d / (1.0 + t2d);
  if (xc == 0.0) return 1.0;
  double lf = b * log(xc) - a * log1p(t2d) - log(b) - lbeta;
  return 1.0 - exp(lf) * betacf(b, a, xc);
}

/*
 * One group (LinearRegressionRowsSingle.execute, LR:46-195) over `M` bed rows.
 *   bed        : M rows of `stride` bytes (no 3-byte header)
 *   idx        : completeColIdx[n] ascending sample indices (RU:116-127)
 *   Qt         : K x n row-major (LR:65-69)      y : n x P row-major (RU:116-118)
 *   Qty        : K x P row-major (LR:71)         yyp : P (LR:78)
 * Outputs: sum_x[M], and ytx/beta/se/t/p as [M, P] row-major.
 */
void lrr_oracle_bed(const uint8_t* bed, int64_t M, int64_t stride, const int32_t* idx, int32_t n,
                    const double* Qt, const double* y, const double* Qty, const double* yyp,
                    int32_t K, int32_t P, int32_t block_size, int32_t n_threads,
                    double* sum_x, double* ytx_out, double* beta, double* se_out, double* t_out, double* p_out) {
  static const double decode[4] = {2.0, NAN, 1.0, 0.0};
  const int32_t d = n - K - 1;
  const double dRec = 1.0 / (double)d; /* LR:51 */
  const int64_t n_blocks = (M + block_size - 1) / block_size;
#ifdef _OPENMP
  if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel
  {
    double* data = (double*)malloc(sizeof(double) * (size_t)n * block_size); /* LR:102 */
    int32_t* missing = (int32_t*)malloc(sizeof(int32_t) * (size_t)n);        /* LR:101 */
    double* qtx = (double*)malloc(sizeof(double) * (size_t)(K > 0 ? K : 1) * block_size);
    double* ytx = (double*)malloc(sizeof(double) * (size_t)P * block_size);
    double* xxpRec = (double*)malloc(sizeof(double) * block_size);
    double* AC = (double*)malloc(sizeof(double) * block_size);
#pragma omp for schedule(static)
    for (int64_t blk = 0; blk < n_blocks; ++blk) {
      const int64_t r0 = blk * block_size;
      const int B = (int)((M - r0 < block_size) ? (M - r0) : block_size);
      /* RU:16-58 setMeanImputedDoubles into column i of data */
      for (int i = 0; i < B; ++i) {
        const uint8_t* row = bed + (r0 + i) * stride;
        double* col = data + (size_t)i * n;
        double sum = 0.0;
        int nMissing = 0;
        for (int32_t j = 0; j < n; ++j) {
          int32_t k = idx[j];
          int code = (row[k >> 2] >> ((k & 3) << 1)) & 3;
          if (code != 1) {
            double e = decode[code];
            sum += e;
            col[j] = e;
          } else {
            missing[nMissing++] = j;
          }
        }
        double mean = sum / (double)(n - nMissing); /* RU:52 */
        for (int m = 0; m < nMissing; ++m) col[missing[m]] = mean;
      }
      /* LR:136 AC = column sums */
      for (int i = 0; i < B; ++i) {
        const double* col = data + (size_t)i * n;
        double s = 0.0;
        for (int32_t j = 0; j < n; ++j) s += col[j];
        AC[i] = s;
      }
      /* LR:139 qtx = Qt * X */
      for (int i = 0; i < B; ++i) {
        const double* col = data + (size_t)i * n;
        for (int c = 0; c < K; ++c) {
          const double* q = Qt + (size_t)c * n;
          double s = 0.0;
          for (int32_t j = 0; j < n; ++j) s += q[j] * col[j];
          qtx[(size_t)c * block_size + i] = s;
        }
      }
      /* LR:141-142 xxpRec = 1 / (x.x - qtx.qtx) */
      for (int i = 0; i < B; ++i) {
        const double* col = data + (size_t)i * n;
        double xx = 0.0, qq = 0.0;
        for (int32_t j = 0; j < n; ++j) xx += col[j] * col[j];
        for (int c = 0; c < K; ++c) qq += qtx[(size_t)c * block_size + i] * qtx[(size_t)c * block_size + i];
        xxpRec[i] = 1.0 / (xx - qq);
      }
      /* LR:143 ytx = y^T * X (raw y, n x P row-major) */
      for (int i = 0; i < B; ++i) {
        const double* col = data + (size_t)i * n;
        for (int p = 0; p < P; ++p) ytx[(size_t)p * block_size + i] = 0.0;
        if (P == 1) {
          double s = 0.0;
          for (int32_t j = 0; j < n; ++j) s += y[j] * col[j];
          ytx[i] = s;
        } else {
          for (int32_t j = 0; j < n; ++j) {
            const double xv = col[j];
            const double* yr = y + (size_t)j * P;
            for (int p = 0; p < P; ++p) ytx[(size_t)p * block_size + i] += yr[p] * xv;
          }
        }
      }
      /* LR:146-160 */
      for (int i = 0; i < B; ++i) {
        const int64_t r = r0 + i;
        sum_x[r] = AC[i];
        for (int p = 0; p < P; ++p) {
          double proj = 0.0;
          for (int c = 0; c < K; ++c) proj += Qty[(size_t)c * P + p] * qtx[(size_t)c * block_size + i];
          double yt = ytx[(size_t)p * block_size + i];
          double xyp = yt - proj;                                    /* LR:146 */
          double b = xyp * xxpRec[i];                                /* LR:150-155 */
          double se = sqrt(dRec * (yyp[p] * xxpRec[i] - b * b));     /* LR:157 */
          double t = b / se;                                         /* LR:159 */
          ytx_out[r * P + p] = yt;
          beta[r * P + p] = b;
          se_out[r * P + p] = se;
          t_out[r * P + p] = t;
          p_out[r * P + p] = lrr_oracle_two_sided_p(t, (double)d);   /* LR:160 */
        }
      }
    }
    free(data); free(missing); free(qtx); free(ytx); free(xxpRec); free(AC);
  }
}

int lrr_oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
