// Per-variant statistics shared by the stand-alone epilogue kernel (stats_epilogue.cu) and the fused epilogue of the
// 4-bit sweep (tc4_kernel.cu).  Restates hail/hail/src/is/hail/methods/LinearRegression.scala:136-160 (same operation
// order for xxpRec, b, se, t) and the Student-t call 2 * T.cumulative(-|t|, d, true, false) (LR:160): jdistlib's
// T.cumulative is a port of R's pt(), a regularised incomplete beta I_x(d/2, 1/2), evaluated here with the Lentz continued
// fraction (direct form in the tail, complement form near the centre).
#pragma once
#include "common.cuh"

namespace lrr {

// Continued fraction of the regularised incomplete beta function, 1 / (1 + d_1 / (1 + d_2 / (1 + ...))) with
// d_{2m+1} = -(a + m)(a + b + m) x / ((a + 2m)(a + 2m + 1)) and d_{2m} = m (b - m) x / ((a + 2m - 1)(a + 2m)), evaluated by the
// FORWARD recurrence of its convergents A_n / B_n with every coefficient kept as numerator p_n over denominator q_n
// (equivalence transformation: A_n = q_n A_{n-1} + q_{n-1} p_n A_{n-2}, same for B_n): no division inside the loop -- the
// modified Lentz form it replaces chains four dependent float64 divisions per step, and the statistics kernel was bound by
// exactly that latency.  The convergents are rescaled by a power of two every step (exact); the loop ends when two
// consecutive convergents agree to 3e-15, tested as a cross product.  Same values as the Lentz form to ~1e-10 of the
// fraction where it is of order a (large df), 1e-13 elsewhere; both sit equally far from scipy's pt().
__device__ inline double betacf_dev(double a, double b, double x) {
  const double eps = 3e-15;
  const double qab = a + b, qap = a + 1.0, qam = a - 1.0;
  double Am1 = 0.0, Bm1 = 1.0, A = 1.0, B = 1.0, qprev = 1.0;
  for (int m = 0; m < 1000; ++m) {
    const double dm = (double)m, m2 = 2.0 * dm;
    double p = -(a + dm) * (qab + dm) * x, q = (a + m2) * (qap + m2);      // d_{2m+1}
    double w = qprev * p;
    double An = fma(q, A, w * Am1), Bn = fma(q, B, w * Bm1);
    Am1 = A; Bm1 = B; A = An; B = Bn; qprev = q;
    const double mm = dm + 1.0, mm2 = m2 + 2.0;
    p = mm * (b - mm) * x;                                                    // d_{2m+2}
    q = (qam + mm2) * (a + mm2);
    w = qprev * p;
    An = fma(q, A, w * Am1);
    Bn = fma(q, B, w * Bm1);
    Am1 = A; Bm1 = B; A = An; B = Bn; qprev = q;
    // bring B back to [1, 2): multiply all four by 2^-exponent(B)
    const int e = ((__double2hiint(B) >> 20) & 0x7ff) - 1023;
    if (e > -1000 && e < 1000) {   // (B == 0, Inf or NaN: leave it to the caller's isnan / the final division)
      const double sc = __hiloint2double((1023 - e) << 20, 0);
      A *= sc; B *= sc; Am1 *= sc; Bm1 *= sc;
    }
    const double cross = A * Bm1;
    if (fabs(cross - Am1 * B) <= eps * fabs(cross)) break;
  }
  return A / B;
}

// p = 2 P[T_df <= -|t|]; lbeta = log B(df/2, 1/2).  Optionally log10(p), finite where p underflows.
__device__ inline double two_sided_p_dev(double t, double df, double lbeta, double* log10_p) {
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
  // Direct form I_x(a, 1/2) in the tail, complement 1 - I_{1-x}(1/2, a) near the centre.  The textbook switch is
  // x < (a + 1) / (a + b + 2); for large df that point is |t| ~ 1.73, where the direct continued fraction still needs
  // ~50 iterations while the complement form needs ~8 -- and under the null almost every variant sits there.  The
  // complement form is therefore kept up to |t| = 2.6 (p >= 0.009: its subtraction loses at most 2 of 16 digits).
  if (x < (a + 1.0) / (a + b + 2.0) && t * t >= 6.76) {
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


// the tail region of two_sided_p_dev: the direct continued fraction, ~50 iterations for large df (the centre takes ~8)
__device__ inline bool p_value_is_tail(double t, double df) {
  if (!(fabs(t) < INFINITY)) return false;   // NaN / Inf are answered at once
  const double a = 0.5 * df, b = 0.5;
  const double x = 1.0 / (1.0 + (t / df) * t);
  return x < (a + 1.0) / (a + b + 2.0) && t * t >= 6.76;
}

struct TailEntry {
  int64_t idx;   // v * P + p
  double t;
};

// what the statistics of one group need besides the per-variant sums
struct StatModel {
  const double* qty;      // [K][P]
  const double* yyp;      // [P]
  int n, K, Kd, P, C, has_intercept, d, weighted;
  int dense;              // dense-dosage sweep: dv[C] = sum of the defined entries, dv[C + 1] = their sum of squares
  int stride;             // doubles per dots row
  int n_fit;              // > 0: dv[C + p] = (fitted-value column of phenotype p) . x, the covariate part of y_transpose_x
  // tolerance guard of the quantised sweeps (NULL = the dots are float64 sums, nothing to guard)
  const double* quantum;  // [C + n_fit] quantisation step of every dot column
  const double* err_sum;  // [C + n_fit] sum over the samples of (true - stored) basis value, or NULL (4-bit sweep only)
  double qscale;          // the per-sample error is at most qscale * quantum / 2 (int8 sweep: 4, its shifted fields)
  double tol, tol_p;      // relative bounds that pass (half of BASELINE's 1e-6 / 1e-5: the other half is the reference's own roundoff)
  double t_floor;         // an absolute error bound of t (of beta in units of its standard error) below this passes too
  int32_t* flag_mark;     // [M]
  int32_t* flag_list;     // [M]
  int32_t* flag_count;
  // the p-values of the tail region are deferred to a compacted second kernel (one slow lane would otherwise hold its whole
  // warp for ~50 iterations: under the null a quarter of the warps hold one); NULL = compute every p-value in place
  TailEntry* tail_list;
  int32_t* tail_count;
  int32_t tail_capacity;
  double lbeta;
  lrr_group_out out;
};

// statistics of variant v, phenotype p of one group from the exact counts and the dot products dv[0..stride).
// `aux` = counts[v].w: dense sweep only, bit 0 / 1 = a +Inf / -Inf entry inside the group.
__device__ inline void variant_stats(const StatModel& a, int64_t v, int p, int n1, int n2, int nm, int aux, const double* dv) {
  const double dRec = 1.0 / (double)a.d;  // LR:51
  const int64_t idx = v * a.P + p;
  const double nv = (double)(a.n - nm);
  const double S = a.dense ? dv[a.C] : (double)(n1 + 2 * n2);
  const double xx_int = (double)(n1 + 4 * n2);
  const double mean = S / nv;                        // RU:52
  // weighted groups (statgen.py:636-660): the column sum and x.x of the sqrt(w)-scaled imputed x are dot products
  double sum_x = a.weighted ? dv[a.Kd + a.P] : S + (double)nm * mean;        // LR:136
  // dense: centred sum of squares of the imputed column (the imputed entries sit at the mean and add nothing)
  const double xxc = a.dense ? dv[a.C + 1] - S * S / nv : 0.0;
  const double xx_imp = a.weighted ? dv[a.Kd + a.P + 1]
                                   : a.dense ? xxc + sum_x * sum_x / (double)a.n : xx_int + (double)nm * mean * mean;

  // Centring (quantised 4-bit sweep): with e_j = true - stored basis value, sum_j e_j x_j = ac sum_j e_j + sum_j e_j (x_j - ac)
  // for any constant ac.  The first term is known (err_sum) and is added to the dot product; the second is bounded by
  // (quantum / 2) sum_j |x_j - ac| -- with ac = the call that makes that L1 distance smallest (the row's prevailing
  // genotype) instead of (quantum / 2) sum_j x_j: 2 x tighter at allele frequency 1/2 (ac = 1), ~1 / (1 - p) x near p = 1.
  double ac = 0.0;
  double l1 = S + (double)nm * mean;   // sum_j |x_j - 0| of the imputed column
  if (a.err_sum && !a.dense) {
    const double n0 = nv - (double)(n1 + n2);
    const double l1_1 = n0 + (double)n2 + (double)nm * fabs(mean - 1.0);
    const double l1_2 = 2.0 * n0 + (double)n1 + (double)nm * (2.0 - mean);
    if (l1_1 < l1) { l1 = l1_1; ac = 1.0; }
    if (l1_2 < l1) { l1 = l1_2; ac = 2.0; }
  }
  auto dot = [&](int c) { return ac != 0.0 ? fma(ac, a.err_sum[c], dv[c]) : dv[c]; };

  double qq = 0.0;
  for (int c = 0; c < a.Kd; ++c) qq += dot(c) * dot(c);
  double xxp;  // x.x - qtx.qtx  (LR:141-142)
  if (a.has_intercept) {
    // constant column handled exactly: x.x - (sum_x)^2/n == xx_int - S^2/nv for the mean-imputed column
    xxp = (a.dense ? xxc : xx_int - S * S / nv) - qq;
  } else {
    xxp = xx_imp - qq;
  }
  const double xyp = dot(a.Kd + p);                  // y_res . x  == ytx - Qty^T qtx (LR:146)
  double proj = 0.0;
  if (a.has_intercept) proj = a.qty[p] * (sum_x / sqrt((double)a.n));
  if (a.n_fit) {
    proj += dot(a.C + p);
  } else {
    for (int c = 0; c < a.Kd; ++c) proj += a.qty[(c + a.has_intercept) * a.P + p] * dot(c);
  }
  double ytx = xyp + proj;                           // LR:143

  double b, se, t, pv, l10 = 0.0;
  // Degenerate (constant / collinear) x: the reference leaves roundoff garbage here (xxp = +-1e-15 and
  // sqrt of a negative -> NaN se; test_statgen.py:277-284).  Rule: no information -> NaN statistics.
  const double nan = __longlong_as_double(0x7ff8000000000000ll);
  const double thr = 1e-11 * xx_imp;
  const bool degenerate = !(xxp > thr);
  if (degenerate && !isnan(xxp)) {
    b = se = t = pv = l10 = nan;
  } else {
    const double xxpRec = 1.0 / xxp;
    b = xyp * xxpRec;                                          // LR:150-155
    se = sqrt(dRec * (a.yyp[p] * xxpRec - b * b));             // LR:157
    t = b / se;                                                // LR:159
    bool deferred = false;
    if (a.tail_list && (a.out.p_value || a.out.log10_p) && p_value_is_tail(t, (double)a.d)) {
      const int k = atomicAdd(a.tail_count, 1);
      if (k < a.tail_capacity) {
        a.tail_list[k].idx = idx;
        a.tail_list[k].t = t;
        deferred = true;
      }
    }
    pv = deferred ? 0.0 : two_sided_p_dev(t, (double)a.d, a.lbeta, a.out.log10_p ? &l10 : nullptr);  // LR:160
  }
  if (a.dense && (aux & 3)) {
    // an infinite entry is a defined value in the reference: sum_x = +-Inf (NaN when both signs occur), x.x = Inf and every
    // statistic NaN (Inf - Inf); the sweep zeroed the entry to keep its other sums finite
    const double inf = __longlong_as_double(0x7ff0000000000000ll);
    sum_x = (aux & 3) == 1 ? inf : (aux & 3) == 2 ? -inf : nan;
    ytx = b = se = t = pv = l10 = nan;
  }

  // ---- tolerance guard (quantised sweeps): rigorous bounds on what the digit quantisation can have changed ----
  if (a.quantum && nv > 0.0) {
    // every stored basis value is within quantum / 2 of the true one, so a (centred, see above) dot product is off by at
    // most (quantum / 2) * sum_j |x_j - ac|, whatever the genotypes are (ac = 0: sum_j x_j, the imputed column is
    // non-negative); + the 2^-41 per sample of the fixed-point error total that ac multiplies
    const double h = 0.5 * a.qscale * l1 * (1.0 + 1e-9) + ac * (double)a.n * 9.1e-13;
    double Eq = 0.0, Eproj = 0.0;
    for (int c = 0; c < a.Kd; ++c) {
      const double e = h * a.quantum[c];
      Eq += (2.0 * fabs(dot(c)) + e) * e;
      if (!a.n_fit) Eproj += fabs(a.qty[(c + a.has_intercept) * a.P + p]) * e;
    }
    const double Ey = h * a.quantum[a.Kd + p];
    const double Eytx = Ey + (a.n_fit ? h * a.quantum[a.C + p] : Eproj);
    bool flag;
    if (fabs(xxp - thr) <= Eq) {
      flag = true;                        // the degenerate-or-not decision itself is in doubt
    } else if (degenerate) {
      flag = false;                       // NaN statistics either way
    } else {
      const double inv = 1.0 / xxp;
      const double rx = Eq / (xxp - Eq);                                   // relative bound of 1 / xxp
      const double Eb = inv * (Ey * (1.0 + rx) + fabs(xyp) * rx);
      const double se2 = dRec * (a.yyp[p] * inv - b * b);
      const double Ese2 = dRec * (a.yyp[p] * inv * rx + (2.0 * fabs(b) + Eb) * Eb);
      const double Ese = Ese2 / (se + sqrt(fmax(se2 - Ese2, 0.0)));
      const double Et = (Eb + fabs(t) * Ese) / (se - Ese);
      const bool ok_b = Eb <= a.tol * fabs(b) || Eb <= a.t_floor * se;
      const bool ok_se = se2 > Ese2 && Ese <= a.tol * se;
      const bool ok_t = Et <= a.tol * fabs(t) || Et <= a.t_floor;
      // |d log p / d t| <= |t| + 2.7 for the two-sided Student-t tail (Mills-ratio bound; 2 f(t) / p <= 2.7 for |t| <= 1)
      const bool ok_p = Et * (fabs(t) + 2.7) <= a.tol_p || Et <= a.t_floor;
      // wide profile: the same absolute floor, expressed on the dot-product scale (d beta = d xyp / xxp)
      const bool ok_ytx = Eytx <= a.tol * fabs(ytx) || (a.P > 2 && Eytx <= a.t_floor * se * xxp);
      flag = !(ok_b && ok_se && ok_t && ok_p && ok_ytx);   // a NaN anywhere fails its test
    }
    if (flag && atomicExch(a.flag_mark + v, 1) == 0) a.flag_list[atomicAdd(a.flag_count, 1)] = (int32_t)v;
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

// fill a StatModel from a group (host)
inline StatModel stat_model_of(const Group& G, const lrr_group_out& out) {
  StatModel a;
  a.qty = G.d_qty;
  a.yyp = G.d_yyp;
  a.n = G.n;
  a.K = G.K;
  a.Kd = G.Kd;
  a.P = G.P;
  a.C = G.C;
  a.has_intercept = G.has_intercept;
  a.d = G.d;
  a.weighted = G.weighted;
  a.dense = 0;
  a.stride = G.C;
  a.n_fit = 0;
  a.quantum = nullptr;
  a.err_sum = nullptr;
  a.qscale = 1.0;
  a.tol = 0.5e-6;
  a.tol_p = 0.5e-5;
  // strict profile: |dt| <= 5e-10 passes (the float64 floor the parity tests allow); wide profile (P > 2, the
  // dense-contraction configuration with its stated tolerance): |dt| <= 5e-7
  a.t_floor = G.P > 2 ? 0.5e-6 : 0.5e-9;
  a.flag_mark = a.flag_list = a.flag_count = nullptr;
  a.tail_list = nullptr;
  a.tail_count = nullptr;
  a.tail_capacity = 0;
  a.lbeta = G.lbeta;
  a.out = out;
  return a;
}

}  // namespace lrr
