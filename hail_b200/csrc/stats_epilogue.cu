// Per-variant epilogue: from exact genotype counts and the projected dot products to
// (sum_x, y_transpose_x, beta, standard_error, t_stat, p_value).
//
// Restates hail/hail/src/is/hail/methods/LinearRegression.scala:136-160 (same operation order for
// xxpRec, b, se, t) and the Student-t call 2 * T.cumulative(-|t|, d, true, false) (LR:160).  jdistlib's
// T.cumulative is a port of R's pt(): a regularised incomplete beta I_x(d/2, 1/2); it is evaluated here with
// the Lentz continued fraction (direct form in the tail, complement form near the centre).
#include "common.cuh"

namespace lrr {

namespace {

__device__ double betacf_dev(double a, double b, double x) {
  const double tiny = 1e-300, eps = 3e-15;  // a tighter test can bounce a few ulp around 1 for hundreds of iterations
  const double qab = a + b, qap = a + 1.0, qam = a - 1.0;
  double c = 1.0, d = 1.0 - qab * x / qap;
  if (fabs(d) < tiny) d = tiny;
  d = 1.0 / d;
  double h = d;
  for (int m = 1; m <= 1000; ++m) {
    const double m2 = 2.0 * m;
    double aa = m * (b - m) * x / ((qam + m2) * (a + m2));
    d = 1.0 + aa * d;
    if (fabs(d) < tiny) d = tiny;
    c = 1.0 + aa / c;
    if (fabs(c) < tiny) c = tiny;
    d = 1.0 / d;
    h *= d * c;
    aa = -(a + m) * (qab + m) * x / ((a + m2) * (qap + m2));
    d = 1.0 + aa * d;
    if (fabs(d) < tiny) d = tiny;
    c = 1.0 + aa / c;
    if (fabs(c) < tiny) c = tiny;
    d = 1.0 / d;
    const double del = d * c;
    h *= del;
    if (fabs(del - 1.0) < eps) break;
  }
  return h;
}

// p = 2 P[T_df <= -|t|]; lbeta = log B(df/2, 1/2).  Optionally log10(p), finite where p underflows.
__device__ double two_sided_p_dev(double t, double df, double lbeta, double* log10_p) {
  const double kInvLn10 = 0.43429448190325182765;
  if (isnan(t)) {
    if (log10_p) *log10_p = t;
    return t;
  }
  if (isinf(t)) {
    if (log10_p) *log10_p = -INFINITY;
    return 0.0;
  }
  const double a = 0.5 * df, b = 0.5;
  const double t2d = (t / df) * t;
  const double x = 1.0 / (1.0 + t2d);
  if (x < (a + 1.0) / (a + b + 2.0)) {
    const double lf = -a * log1p(t2d) + b * log(t2d / (1.0 + t2d)) - log(a) - lbeta;
    const double cf = betacf_dev(a, b, x);
    if (log10_p) *log10_p = (lf + log(cf)) * kInvLn10;
    return exp(lf) * cf;
  }
  const double xc = t2d / (1.0 + t2d);
  if (xc == 0.0) {
    if (log10_p) *log10_p = 0.0;
    return 1.0;
  }
  const double lf = b * log(xc) - a * log1p(t2d) - log(b) - lbeta;
  const double lower = exp(lf) * betacf_dev(b, a, xc);
  if (log10_p) *log10_p = log1p(-lower) * kInvLn10;
  return 1.0 - lower;
}

struct EpiArgs {
  const int32_t* counts;  // [M][4]
  const double* dots;     // [M][C]
  const double* qty;      // [K][P]
  const double* yyp;      // [P]
  int64_t M;
  int n, K, Kd, P, C, has_intercept, d, weighted;
  double lbeta;
  lrr_group_out out;
};

__global__ void stats_epilogue_kernel(EpiArgs a) {
  const int64_t total = a.M * a.P;
  const double dRec = 1.0 / (double)a.d;  // LR:51
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t v = idx / a.P;
    const int p = (int)(idx - v * a.P);
    const int4 cnt = reinterpret_cast<const int4*>(a.counts)[v];
    const int n1 = cnt.x, n2 = cnt.y, nm = cnt.z;
    const double nv = (double)(a.n - nm);
    const double S = (double)(n1 + 2 * n2);
    const double xx_int = (double)(n1 + 4 * n2);
    const double mean = S / nv;                        // RU:52
    const double* dv = a.dots + v * a.C;
    // weighted groups (statgen.py:636-660): the column sum and x.x of the sqrt(w)-scaled imputed x are dot products
    const double sum_x = a.weighted ? dv[a.Kd + a.P] : S + (double)nm * mean;        // LR:136
    const double xx_imp = a.weighted ? dv[a.Kd + a.P + 1] : xx_int + (double)nm * mean * mean;

    double qq = 0.0;
    for (int c = 0; c < a.Kd; ++c) qq += dv[c] * dv[c];
    double xxp;  // x.x - qtx.qtx  (LR:141-142)
    if (a.has_intercept) {
      // constant column handled exactly: x.x - (sum_x)^2/n == xx_int - S^2/nv for the mean-imputed column
      xxp = (xx_int - S * S / nv) - qq;
    } else {
      xxp = xx_imp - qq;
    }
    const double xyp = dv[a.Kd + p];                   // y_res . x  == ytx - Qty^T qtx (LR:146)
    double proj = 0.0;
    if (a.has_intercept) proj = a.qty[p] * (sum_x / sqrt((double)a.n));
    for (int c = 0; c < a.Kd; ++c) proj += a.qty[(c + a.has_intercept) * a.P + p] * dv[c];
    const double ytx = xyp + proj;                     // LR:143

    double b, se, t, pv, l10 = 0.0;
    // Degenerate (constant / collinear) x: the reference leaves roundoff garbage here (xxp = +-1e-15 and
    // sqrt of a negative -> NaN se; test_statgen.py:277-284).  Rule: no information -> NaN statistics.
    const bool degenerate = !(xxp > 1e-11 * xx_imp);
    if (degenerate && !isnan(xxp)) {
      b = se = t = pv = l10 = __longlong_as_double(0x7ff8000000000000ll);
    } else {
      const double xxpRec = 1.0 / xxp;
      b = xyp * xxpRec;                                          // LR:150-155
      se = sqrt(dRec * (a.yyp[p] * xxpRec - b * b));             // LR:157
      t = b / se;                                                // LR:159
      pv = two_sided_p_dev(t, (double)a.d, a.lbeta, a.out.log10_p ? &l10 : nullptr);  // LR:160
    }
    if (p == 0) {
      if (a.out.n) a.out.n[v] = a.n;
      if (a.out.n_missing) a.out.n_missing[v] = nm;
      if (a.out.sum_x) a.out.sum_x[v] = sum_x;
    }
    if (a.out.y_transpose_x) a.out.y_transpose_x[idx] = ytx;
    if (a.out.beta) a.out.beta[idx] = b;
    if (a.out.standard_error) a.out.standard_error[idx] = se;
    if (a.out.t_stat) a.out.t_stat[idx] = t;
    if (a.out.p_value) a.out.p_value[idx] = pv;
    if (a.out.log10_p) a.out.log10_p[idx] = l10;
  }
}

// Logistic score test per variant (LogisticRegressionModel.scala:211-264 restated through the Schur complement of the
// null block): with f01 = C' W x, f11 = x' W x, s1 = x' (y - mu) from the sweep and F00^-1, u = F00^-1 s0, a0 = s0' u
// from the host,  chi2 = a0 + (s1 - f01' u)^2 / (f11 - f01' F00^-1 f01),  p = pchisqtail(chi2, 1) = erfc(sqrt(chi2 / 2)).
struct ScoreArgs {
  const int32_t* counts;  // [M][4]
  const double* dots;     // [M][K + 3]: f01 (K), s1, sum sqrt(w) x (unused), f11
  const double* finv;     // [K][K]
  const double* aux;      // [K + 1]: u, a0
  int64_t M;
  int K;
  lrr_score_out out;
};

__global__ void score_epilogue_kernel(ScoreArgs a) {
  for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < a.M; v += (int64_t)gridDim.x * blockDim.x) {
    const double* dv = a.dots + v * (a.K + 3);
    const double s1 = dv[a.K], f11 = dv[a.K + 2];
    double quad = 0.0, fu = 0.0;
    for (int i = 0; i < a.K; ++i) {
      double t = 0.0;
      for (int j = 0; j < a.K; ++j) t += a.finv[i * a.K + j] * dv[j];
      quad += dv[i] * t;
      fu += dv[i] * a.aux[i];
    }
    const double denom = f11 - quad;
    double chi2, p;
    // x in the span of the covariates (e.g. a constant call): the full Fisher matrix is singular -> missing (:256-259)
    if (!(denom > 1e-11 * f11)) {
      chi2 = p = __longlong_as_double(0x7ff8000000000000ll);
    } else {
      const double num = s1 - fu;
      chi2 = a.aux[a.K] + num * num / denom;
      p = erfc(sqrt(0.5 * chi2));
    }
    if (a.out.chi_sq_stat) a.out.chi_sq_stat[v] = chi2;
    if (a.out.p_value) a.out.p_value[v] = p;
    if (a.out.n_missing) a.out.n_missing[v] = reinterpret_cast<const int4*>(a.counts)[v].z;
  }
}

__global__ void student_t_kernel(const double* t, int64_t count, double df, double lbeta, double* p, double* l10) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
    double l = 0.0;
    const double pv = two_sided_p_dev(t[i], df, lbeta, l10 ? &l : nullptr);
    if (p) p[i] = pv;
    if (l10) l10[i] = l;
  }
}

}  // namespace

double log_beta_half(double a) {
  if (a < 50.0) return lgamma(a) + lgamma(0.5) - lgamma(a + 0.5);
  const double ia = 1.0 / a, ia2 = ia * ia;
  const double series = ia * (1.0 / 8.0 + ia2 * (-1.0 / 192.0 + ia2 * (1.0 / 640.0 + ia2 * (-17.0 / 14336.0))));
  return 0.5 * log(3.14159265358979323846) - 0.5 * log(a) + series;
}

int launch_stats_epilogue(Ctx* c, int g, int64_t M, const lrr_group_out& out, cudaStream_t st) {
  if (M == 0) return LRR_OK;
  const Group& G = c->groups[g];
  EpiArgs a;
  a.counts = c->d_counts + (int64_t)g * c->reserved_variants * 4;
  a.dots = c->d_dots + c->dots_offset[g];
  a.qty = G.d_qty;
  a.yyp = G.d_yyp;
  a.M = M;
  a.n = G.n;
  a.K = G.K;
  a.Kd = G.Kd;
  a.P = G.P;
  a.C = G.C;
  a.has_intercept = G.has_intercept;
  a.weighted = G.weighted;
  a.d = G.d;
  a.lbeta = G.lbeta;
  a.out = out;
  const int64_t total = M * G.P;
  int64_t grid = (total + 127) / 128;
  if (grid > (int64_t)c->sm_count * 32) grid = (int64_t)c->sm_count * 32;
  stats_epilogue_kernel<<<(int)grid, 128, 0, st>>>(a);
  c->launches++;
  LRR_CUDA(c, cudaGetLastError());
  return LRR_OK;
}

int launch_score_epilogue(Ctx* c, int64_t M, const lrr_score_out& out, cudaStream_t st) {
  if (M == 0) return LRR_OK;
  const Group& G = c->groups[0];
  ScoreArgs a;
  a.counts = c->d_counts;
  a.dots = c->d_dots + c->dots_offset[0];
  a.finv = G.d_qty;
  a.aux = G.d_yyp;
  a.M = M;
  a.K = G.K;
  a.out = out;
  int64_t grid = (M + 127) / 128;
  if (grid > (int64_t)c->sm_count * 32) grid = (int64_t)c->sm_count * 32;
  score_epilogue_kernel<<<(int)grid, 128, 0, st>>>(a);
  c->launches++;
  LRR_CUDA(c, cudaGetLastError());
  return LRR_OK;
}

int launch_student_t(Ctx* c, const double* d_t, int64_t count, double df, double* d_p, double* d_l10,
                     cudaStream_t st) {
  if (count == 0) return LRR_OK;
  int64_t grid = (count + 127) / 128;
  if (grid > 4096) grid = 4096;
  student_t_kernel<<<(int)grid, 128, 0, st>>>(d_t, count, df, log_beta_half(0.5 * df), d_p, d_l10);
  c->launches++;
  LRR_CUDA(c, cudaGetLastError());
  return LRR_OK;
}

}  // namespace lrr
