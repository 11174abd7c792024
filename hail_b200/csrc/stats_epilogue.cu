// Per-variant epilogue kernels: from exact genotype counts and the projected dot products to
// (sum_x, y_transpose_x, beta, standard_error, t_stat, p_value) -- the mathematics lives in stats_device.cuh -- and the
// logistic score-test epilogue.
#include <algorithm>

#include "stats_device.cuh"

namespace lrr {

namespace {

struct EpiArgs {
  const int32_t* counts;  // [M][4]
  const double* dots;     // [M][C]
  int64_t M;
  StatModel model;
};

__global__ void stats_epilogue_kernel(EpiArgs a) {
  const int64_t total = a.M * a.model.P;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t v = idx / a.model.P;
    const int p = (int)(idx - v * a.model.P);
    const int4 cnt = reinterpret_cast<const int4*>(a.counts)[v];
    variant_stats(a.model, v, p, cnt.x, cnt.y, cnt.z, cnt.w, a.dots + v * a.model.stride);
  }
}

// the same statistics for the rows the guard listed, after their float64 recompute (a.model.quantum == NULL here)
__global__ void stats_epilogue_listed_kernel(EpiArgs a, const int32_t* __restrict__ list, const int32_t* __restrict__ count) {
  const int64_t total = (int64_t)(*count) * a.model.P;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = idx / a.model.P;
    const int p = (int)(idx - i * a.model.P);
    const int64_t v = list[i];
    const int4 cnt = reinterpret_cast<const int4*>(a.counts)[v];
    variant_stats(a.model, v, p, cnt.x, cnt.y, cnt.z, cnt.w, a.dots + v * a.model.stride);
  }
}

// the deferred p-values: every lane runs the long continued fraction
__global__ void stats_tail_kernel(const TailEntry* __restrict__ list, const int32_t* __restrict__ count, int capacity, double df,
                                  double lbeta, double* __restrict__ p_value, double* __restrict__ log10_p) {
  const int n = min(*count, capacity);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const TailEntry e = list[i];
    double l = 0.0;
    const double pv = two_sided_p_dev(e.t, df, lbeta, log10_p ? &l : nullptr);
    if (p_value) p_value[e.idx] = pv;
    if (log10_p) log10_p[e.idx] = l;
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
  int stride;             // doubles per dots row: K + 3 (packed sweep) or K + 5 (dense sweep)
  lrr_score_out out;
};

__global__ void score_epilogue_kernel(ScoreArgs a) {
  for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < a.M; v += (int64_t)gridDim.x * blockDim.x) {
    const double* dv = a.dots + v * a.stride;
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

int launch_stats_epilogue(Ctx* c, int g, int64_t M, const lrr_group_out& out, cudaStream_t st, bool dense,
                          const double* quantum, int n_fit, int stride, double qscale, const double* err_sum) {
  if (M == 0) return LRR_OK;
  const Group& G = c->groups[g];
  EpiArgs a;
  a.counts = c->d_counts + (int64_t)g * c->reserved_variants * 4;
  a.dots = c->d_dots + c->dots_offset[g];
  a.M = M;
  a.model = stat_model_of(G, out);
  a.model.dense = dense ? 1 : 0;
  a.model.stride = stride > 0 ? stride : G.C + (dense ? 2 : 0);
  a.model.n_fit = n_fit;
  if (err_sum && !dense && !G.weighted) a.model.err_sum = err_sum;
  if (quantum && c->guard && !G.weighted) {
    a.model.quantum = quantum;
    a.model.qscale = qscale;
    a.model.flag_mark = c->d_flag_mark + (int64_t)g * c->reserved_variants;
    a.model.flag_list = c->d_flag_list + (int64_t)g * c->reserved_variants;
    a.model.flag_count = c->d_flag_count + g;
  }
  const int64_t total = M * G.P;
  const bool defer = c->d_tail && total >= 4096 && (out.p_value || out.log10_p);
  if (defer) {
    a.model.tail_list = static_cast<TailEntry*>(c->d_tail);
    a.model.tail_count = c->d_tail_count;
    a.model.tail_capacity = (int32_t)std::min<int64_t>(c->tail_capacity, 1 << 30);
    LRR_CUDA(c, cudaMemsetAsync(c->d_tail_count, 0, sizeof(int32_t), st));
  }
  int64_t grid = (total + 127) / 128;
  if (grid > (int64_t)c->sm_count * 32) grid = (int64_t)c->sm_count * 32;
  stats_epilogue_kernel<<<(int)grid, 128, 0, st>>>(a);
  c->launches++;
  LRR_CUDA(c, cudaGetLastError());
  if (defer) {
    stats_tail_kernel<<<c->sm_count * 4, 128, 0, st>>>(a.model.tail_list, a.model.tail_count, a.model.tail_capacity, (double)G.d, G.lbeta,
                                                      out.p_value, out.log10_p);
    c->launches++;
    LRR_CUDA(c, cudaGetLastError());
  }
  return LRR_OK;
}

int launch_stats_epilogue_listed(Ctx* c, int g, const lrr_group_out& out, int stride, cudaStream_t st) {
  const Group& G = c->groups[g];
  EpiArgs a;
  a.counts = c->d_counts + (int64_t)g * c->reserved_variants * 4;
  a.dots = c->d_dots + c->dots_offset[g];
  a.M = 0;
  a.model = stat_model_of(G, out);
  a.model.stride = stride;
  stats_epilogue_listed_kernel<<<c->sm_count, 128, 0, st>>>(a, c->d_flag_list + (int64_t)g * c->reserved_variants,
                                                          c->d_flag_count + g);
  c->launches++;
  LRR_CUDA(c, cudaGetLastError());
  return LRR_OK;
}

int launch_score_epilogue(Ctx* c, int64_t M, const lrr_score_out& out, cudaStream_t st, bool dense) {
  if (M == 0) return LRR_OK;
  const Group& G = c->groups[0];
  ScoreArgs a;
  a.counts = c->d_counts;
  a.dots = c->d_dots + c->dots_offset[0];
  a.finv = G.d_qty;
  a.aux = G.d_yyp;
  a.M = M;
  a.K = G.K;
  a.stride = G.C + (dense ? 2 : 0);
  a.out = out;
  int64_t grid = (M + 127) / 128;
  if (grid > (int64_t)c->sm_count * 32) grid = (int64_t)c->sm_count * 32;
  score_epilogue_kernel<<<(int)grid, 128, 0, st>>>(a);
  c->launches++;
  LRR_CUDA(c, cudaGetLastError());
  return LRR_OK;
}

// chi2[i] = qchisqtail(p[i], 1) = 2 erfcinv(p)^2: the statistic whose upper tail is p (lambda_gc, statgen.py:3121-3128)
__global__ void qchisqtail1_kernel(const double* __restrict__ p, int64_t count, double* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
    const double e = erfcinv(p[i]);
    out[i] = 2.0 * e * e;
  }
}

int launch_qchisqtail1(Ctx* c, const double* d_p, int64_t count, double* d_out, cudaStream_t st) {
  if (count == 0) return LRR_OK;
  int64_t grid = (count + 255) / 256;
  if (grid > 4096) grid = 4096;
  qchisqtail1_kernel<<<(int)grid, 256, 0, st>>>(d_p, count, d_out);
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
