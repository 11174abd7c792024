// Kernel 5: per-variant logistic regression fits -- the Wald, likelihood-ratio and Firth tests of
// `hl.logistic_regression_rows`.  One CTA per variant; Newton iterations in float64.
//
// Reference, per row (hail/hail/src/is/hail/methods/LogisticRegression.scala:115-157): x is mean-imputed over the complete
// samples into the last column of the design matrix X = [covariates | x] (RegressionUtils.scala:16-58), then
//   WaldTest / LikelihoodRatioTest   stats/LogisticRegressionModel.scala:55-145: model.fit(Some(nullFit)) (:294-370) -- start
//                                    at b = [b_null, 0]; the covariate blocks of the first score / Fisher matrix are the
//                                    null fit's (:311-325); iterate delta = fisher \ score until max|delta| < tol;
//                                    Wald: se = sqrt(diag(inv(fisher))), z = b / se, p = 2 pnorm(-|z|);
//                                    LRT: chi2 = 2 (logLkhd - logLkhd_null), p = pchisqtail(chi2, 1)
//   LogisticFirthTest                :155-199 with fitFirth (:372-408): a null fit (K free coefficients) and a full fit
//                                    (K + 1), both with the hat diagonal of the FULL design; the reference takes it from a
//                                    QR of sqrt(W) X, here it comes from the normal equations: h_i = w_i x_i' F^-1 x_i,
//                                    delta = F_00^-1 X_0' (y - mu + h (1/2 - mu)), sum log|diag R| = 1/2 log det F.
// Every evaluation is one pass over the n complete samples: the CTA's 256 threads stride over them (covariate planes
// [K][n] are L2-resident and read coalesced, genotype codes are gathered from the packed row through the complete-sample
// index), each thread keeps the m + m (m + 1) / 2 partial sums of the score and the Fisher matrix in registers
// (compile-time MM >= m; in 2 / 4 slices with one pass each for m > 12), then a warp-shuffle + shared-memory reduction; thread 0 solves the m x m system (LU with
// partial pivoting, singular = an exactly zero pivot, as LAPACK dgesv under breeze's `\`).
#include "common.cuh"

namespace lrr {

namespace {

constexpr int LT = 256;   // threads per CTA

struct LogitModel {
  int n = 0, K = 0;
  int64_t n_samples_total = 0;
  int32_t* d_idx = nullptr;   // [n] stored-sample index of each complete sample
  double* d_cov = nullptr;    // [K][n]
  double* d_y = nullptr;      // [n]
  double* d_null = nullptr;   // [K] b0 | [K] score0 | [K*K] fisher0 | loglk0
};

struct LogitArgs {
  const uint8_t* packed;   // 2-bit rows, or
  const double* dense;     // [M][ldx] float64 entries over all stored samples, NaN = missing (packed == nullptr)
  int64_t ldx;
  int64_t M, stride;
  int n, K;
  const int32_t* idx;
  const double* cov;
  const double* y;
  const double* null_fit;
  int test;        // 1 wald, 2 lrt, 3 firth
  int max_iter;
  double tol;
  lrr_logit_out out;
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ double nan_value() { return __longlong_as_double(0x7ff8000000000000ll); }

// Solve A z = rhs in place (A [m][MM] row-major in shared memory, destroyed).  Returns false when a pivot is exactly 0.
template <int MM>
__device__ bool lu_solve(double (*A)[MM], double* rhs, int m) {
  for (int k = 0; k < m; ++k) {
    int p = k;
    double best = fabs(A[k][k]);
    for (int i = k + 1; i < m; ++i) {
      const double v = fabs(A[i][k]);
      if (v > best) { best = v; p = i; }
    }
    if (A[p][k] == 0.0) return false;
    if (p != k) {
      for (int j = 0; j < m; ++j) { const double t = A[k][j]; A[k][j] = A[p][j]; A[p][j] = t; }
      const double t = rhs[k]; rhs[k] = rhs[p]; rhs[p] = t;
    }
    const double inv = 1.0 / A[k][k];
    for (int i = k + 1; i < m; ++i) {
      const double f = A[i][k] * inv;
      if (f != 0.0) {
        for (int j = k + 1; j < m; ++j) A[i][j] -= f * A[k][j];
        rhs[i] -= f * rhs[k];
      }
    }
  }
  for (int k = m - 1; k >= 0; --k) {
    double s = rhs[k];
    for (int j = k + 1; j < m; ++j) s -= A[k][j] * rhs[j];
    rhs[k] = s / A[k][k];
  }
  return true;
}

// Inverse of A (destroyed) into Inv by Gauss-Jordan with partial pivoting; also log|det A|.  false = singular.
template <int MM>
__device__ bool invert(double (*A)[MM], double (*Inv)[MM], int m, double* logdet) {
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < m; ++j) Inv[i][j] = (i == j) ? 1.0 : 0.0;
  double ld = 0.0;
  for (int k = 0; k < m; ++k) {
    int p = k;
    double best = fabs(A[k][k]);
    for (int i = k + 1; i < m; ++i) {
      const double v = fabs(A[i][k]);
      if (v > best) { best = v; p = i; }
    }
    if (A[p][k] == 0.0) return false;
    if (p != k)
      for (int j = 0; j < m; ++j) {
        double t = A[k][j]; A[k][j] = A[p][j]; A[p][j] = t;
        t = Inv[k][j]; Inv[k][j] = Inv[p][j]; Inv[p][j] = t;
      }
    ld += log(fabs(A[k][k]));
    const double inv = 1.0 / A[k][k];
    for (int j = 0; j < m; ++j) { A[k][j] *= inv; Inv[k][j] *= inv; }
    for (int i = 0; i < m; ++i) {
      if (i == k) continue;
      const double f = A[i][k];
      if (f != 0.0)
        for (int j = 0; j < m; ++j) { A[i][j] -= f * A[k][j]; Inv[i][j] -= f * Inv[k][j]; }
    }
  }
  *logdet = ld;
  return true;
}

constexpr int TS = 128;  // samples per tile of the tiled form (MM > 20): two threads per sample form eta; the serial exp / log of
                         // a tile then runs on 128 lanes instead of 32 (with 32-sample tiles it was the longest phase)
constexpr int TPS = LT / TS;   // threads per sample in that phase

template <int MM>
struct Shared {
  // register form: per-warp partial sums of the score, the Fisher triangle and the log-likelihood; tiled form: scratch only
  double red[LT / 32][MM <= 20 ? MM + MM * (MM + 1) / 2 + 1 : 4];
  double F[MM][MM];      // Fisher matrix (symmetric, full)
  double W[MM][MM];      // work copy for the solves
  double Inv[MM][MM];
  double score[MM], b[MM], delta[MM];
  double loglik, logdet, mean;
  int counts[3];
  int status;            // 0 continue, 1 converged, 2 exploded
};

// tiled form (MM > 20): one tile of the design matrix [TS samples][MM columns] with the samples' weights and residuals
template <int MM>
struct Tile {
  double X[TS][MM + 2];   // row stride a multiple of 16 bytes (128-bit loads of four columns), +2: rows start on different banks
  double w[TS], r[TS];
};

// One evaluation over the complete samples at coefficients b[0..m0) (columns m0..m-1 of X do not enter eta).
//   MODE 0: score[a] = sum x_a (y - mu), F[a][b] = sum w x_a x_b, loglik = sum log(y mu + (1 - y)(1 - mu))   (all m columns)
//   MODE 1: Firth second pass: score[a] = sum x_a (y - mu + h (1/2 - mu)) with h = w x' Inv x
// The partial sums live in registers.  For MM <= 12 the whole lower triangle of F fits (one pass over the samples); for
// MM = 16 / 20 it is accumulated in 2 / 4 slices, one pass each (eta and mu are recomputed), and MODE 1 reads the inverse
// from shared memory instead of holding it in registers.
template <int MM>
struct Slices {
  static constexpr int value = MM <= 12 ? 1 : MM <= 16 ? 2 : 4;
};

template <int MM, int MODE>
__device__ void eval_pass(const LogitArgs& a, Shared<MM>& sh, const uint32_t* row, const double* drow, int m, int m0,
                          bool want_loglik) {
  constexpr int NF = MM * (MM + 1) / 2;
  constexpr int PARTS = (MODE == 0) ? Slices<MM>::value : 1;
  constexpr int NFP = (NF + PARTS - 1) / PARTS;
  constexpr bool INV_SMEM = (MODE == 1) && (MM > 12);
  constexpr int NFI = (MODE == 0) ? NFP : (INV_SMEM ? 1 : NF);
  const double mean = sh.mean;
  const int K = a.K;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int part = 0; part < PARTS; ++part) {
    double sc[MM], fi[NFI];
    double ll = 0.0;
#pragma unroll
    for (int i = 0; i < MM; ++i) sc[i] = 0.0;
#pragma unroll
    for (int i = 0; i < NFI; ++i) fi[i] = 0.0;
    if (MODE == 1 && !INV_SMEM) {   // the symmetric inverse, off-diagonal entries doubled
#pragma unroll
      for (int p = 0, q = 0; p < MM; ++p)
#pragma unroll
        for (int r = 0; r <= p; ++r, ++q)
          fi[INV_SMEM ? 0 : q] = (p < m && r < m) ? (p == r ? sh.Inv[p][r] : 2.0 * sh.Inv[p][r]) : 0.0;
    }
    double bb[MM];
#pragma unroll
    for (int i = 0; i < MM; ++i) bb[i] = (i < m0) ? sh.b[i] : 0.0;
    for (int i = threadIdx.x; i < a.n; i += LT) {
      double xa[MM];
#pragma unroll
      for (int k = 0; k < MM; ++k) xa[k] = 0.0;
#pragma unroll
      for (int k = 0; k < MM - 1; ++k)
        if (k < K) xa[k] = __ldg(a.cov + (int64_t)k * a.n + i);
      const int s = __ldg(a.idx + i);
      double x;
      if (drow) {
        x = __ldg(drow + s);
        if (x != x) x = mean;
      } else {
        const uint32_t code = (__ldg(row + (s >> 4)) >> sample_shift(s & 15)) & 3u;
        x = (code == 3u) ? mean : (double)code;
      }
#pragma unroll
      for (int k = 0; k < MM; ++k)
        if (k == K) xa[k] = x;
      double e4[4] = {0.0, 0.0, 0.0, 0.0};   // four short chains instead of one of length MM (FP64 latency)
#pragma unroll
      for (int k = 0; k < MM; ++k) e4[k & 3] = fma(bb[k], xa[k], e4[k & 3]);
      const double eta = (e4[0] + e4[1]) + (e4[2] + e4[3]);
      const double mu = 1.0 / (1.0 + exp(-eta));
      const double w = mu * (1.0 - mu);
      const double yi = __ldg(a.y + i);
      double r = yi - mu;
      if (MODE == 0) {
        if (part == 0 && want_loglik) ll += log(yi * mu + (1.0 - yi) * (1.0 - mu));
#pragma unroll
        for (int p = 0, q = 0; p < MM; ++p) {
          const double wx = w * xa[p];
#pragma unroll
          for (int c = 0; c <= p; ++c, ++q)
            if (q >= part * NFP && q < (part + 1) * NFP) fi[(q - part * NFP) % NFI] = fma(wx, xa[c], fi[(q - part * NFP) % NFI]);
        }
      } else {
        double h = 0.0;
#pragma unroll
        for (int p = 0, q = 0; p < MM; ++p) {
          double t = 0.0;
#pragma unroll
          for (int c = 0; c <= p; ++c, ++q)
            t = fma(INV_SMEM ? ((p < m) ? (p == c ? sh.Inv[p][c] : 2.0 * sh.Inv[p][c]) : 0.0) : fi[INV_SMEM ? 0 : q], xa[c], t);
          h = fma(t, xa[p], h);
        }
        r += w * h * (0.5 - mu);
      }
      if (part == 0) {
#pragma unroll
        for (int k = 0; k < MM; ++k) sc[k] = fma(xa[k], r, sc[k]);
      }
    }
    // ---- reduce over the lanes ----
    if (part == 0) {
#pragma unroll
      for (int k = 0; k < MM; ++k) {
        const double t = warp_sum(sc[k]);
        if (lane == 0) sh.red[warp][k] = t;
      }
    }
    if (MODE == 0) {
#pragma unroll
      for (int q = 0; q < NFP; ++q) {
        if (part * NFP + q < NF) {
          const double t = warp_sum(fi[q % NFI]);
          if (lane == 0) sh.red[warp][MM + part * NFP + q] = t;
        }
      }
      if (part == 0) {
        const double t = warp_sum(ll);
        if (lane == 0) sh.red[warp][MM + NF] = t;
      }
    }
  }
  __syncthreads();
  const int total = (MODE == 0) ? MM + NF + 1 : MM;
  for (int q = threadIdx.x; q < total; q += LT) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < LT / 32; ++w) t += sh.red[w][q];
    if (q < MM) {
      sh.score[q] = t;
    } else if (q < MM + NF) {
      int p = 0, rem = q - MM;
      while (rem > p) { rem -= p + 1; ++p; }
      sh.F[p][rem] = t;
      sh.F[rem][p] = t;
    } else {
      sh.loglik = t;
    }
  }
  __syncthreads();
}

// The same evaluation for wide models (20 < m <= MM = 32 / 48 / 64): the (m + 1) m / 2 sums no longer fit the registers of
// one thread, so the CTA walks the samples in tiles of TS: the tile of X = [covariates | x] is staged in shared memory (the
// next tile's values travel in registers meanwhile), two threads per sample form eta -> mu, w, the residual (MODE 1: and
// the hat value x' Inv x), then the weighted Gram update F += X' diag(w) X runs as 4 x 4 register blocks: thread `tid` owns
// block (bi >= bj) of the lower triangle for ALL samples, so F needs no reduction and its sum order is fixed; threads
// nblk .. nblk + MM - 1 own one score entry each.
template <int MM, int MODE>
__device__ void eval_pass_tiled(const LogitArgs& a, Shared<MM>& sh, Tile<MM>& tl, const uint32_t* row, const double* drow, int m,
                                int m0, bool want_loglik) {
  constexpr int NB = MM / 4, NBLK = NB * (NB + 1) / 2;
  // G thread groups share the samples of a tile (group g takes samples g, g + G, ...): 4 x 36 block threads for MM = 32,
  // 2 x 78 for 48, 1 x 136 for 64; the groups' partial blocks are added in group order after the last tile
  constexpr int G = (LT - MM) / NBLK >= 4 ? 4 : (LT - MM) / NBLK >= 2 ? 2 : 1;
  constexpr int PER = TS * MM / LT;   // staged values per thread and tile
  static_assert(TS * MM % LT == 0 && G * NBLK + MM <= LT && TPS * TS == LT && (TPS & (TPS - 1)) == 0 && TS % G == 0,
                "thread mapping of the tiled form");
  static_assert((G - 1) * NBLK * 16 <= 2 * MM * MM, "the partial blocks of groups 1.. fit the W | Inv scratch");
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int K = a.K;
  const double mean = sh.mean;
  const int grp = tid / NBLK, blk = tid % NBLK;   // block threads: tid < G * NBLK
  int bi = 0, bj = 0;
  if (tid < G * NBLK) {
    int p = 0, rem = blk;
    while (rem > p) { rem -= p + 1; ++p; }
    bi = p;
    bj = rem;
  }
  const int sk = tid - G * NBLK;   // score entry of this thread (0 <= sk < MM)
  double acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
  double sc = 0.0, ll = 0.0;
  // value (k, t) of the tile starting at sample i0: covariate k, the imputed x (k == K), or 0 (padding)
  auto fetch = [&](int q, int i0) -> double {
    const int idx = tid + q * LT;
    const int k = idx / TS, t = idx % TS;
    const int i = i0 + t;
    if (i >= a.n || k > K) return 0.0;
    if (k < K) return __ldg(a.cov + (int64_t)k * a.n + i);
    const int smp = __ldg(a.idx + i);
    if (drow) {
      const double x = __ldg(drow + smp);
      return x != x ? mean : x;
    }
    const uint32_t code = (__ldg(row + (smp >> 4)) >> sample_shift(smp & 15)) & 3u;
    return code == 3u ? mean : (double)code;
  };
  double nxt[PER];
#pragma unroll
  for (int q = 0; q < PER; ++q) nxt[q] = fetch(q, 0);
  for (int i0 = 0; i0 < a.n; i0 += TS) {
#pragma unroll
    for (int q = 0; q < PER; ++q) {
      const int idx = tid + q * LT;
      tl.X[idx % TS][idx / TS] = nxt[q];
    }
    __syncthreads();
    if (i0 + TS < a.n) {
#pragma unroll
      for (int q = 0; q < PER; ++q) nxt[q] = fetch(q, i0 + TS);
    }
    {   // eta, mu, w, residual: TPS threads per sample
      const int t = tid / TPS, part = tid % TPS;
      const double* xr = tl.X[t];
      double e = 0.0;
      for (int k = part; k < m0; k += TPS) e = fma(sh.b[k], xr[k], e);
#pragma unroll
      for (int o = 1; o < TPS; o <<= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
      double h = 0.0;
      if (MODE == 1) {   // hat value x' Inv x (Inv symmetric)
        for (int p = part; p < m; p += TPS) {
          double tp = 0.0;
          for (int c = 0; c < m; ++c) tp = fma(sh.Inv[p][c], xr[c], tp);
          h = fma(tp, xr[p], h);
        }
#pragma unroll
        for (int o = 1; o < TPS; o <<= 1) h += __shfl_xor_sync(0xffffffffu, h, o);
      }
      if (part == 0) {
        const int i = i0 + t;
        double w = 0.0, r = 0.0;
        if (i < a.n) {
          const double mu = 1.0 / (1.0 + exp(-e));
          const double yi = __ldg(a.y + i);
          w = mu * (1.0 - mu);
          r = yi - mu;
          if (MODE == 0 && want_loglik) ll += log(yi * mu + (1.0 - yi) * (1.0 - mu));
          if (MODE == 1) r += w * h * (0.5 - mu);
        }
        tl.w[t] = w;
        tl.r[t] = r;
      }
    }
    __syncthreads();
    if (MODE == 0 && tid < G * NBLK) {
#pragma unroll 4
      for (int t = grp; t < TS; t += G) {
        const double2 a01 = *reinterpret_cast<const double2*>(&tl.X[t][4 * bi]), a23 = *reinterpret_cast<const double2*>(&tl.X[t][4 * bi + 2]);
        const double2 b01 = *reinterpret_cast<const double2*>(&tl.X[t][4 * bj]), b23 = *reinterpret_cast<const double2*>(&tl.X[t][4 * bj + 2]);
        const double wt = tl.w[t];
        const double av[4] = {a01.x * wt, a01.y * wt, a23.x * wt, a23.y * wt};
        const double bv[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fma(av[i], bv[j], acc[i][j]);
      }
    } else if (sk >= 0 && sk < MM) {
#pragma unroll 8
      for (int t = 0; t < TS; ++t) sc = fma(tl.X[t][sk], tl.r[t], sc);
    }
    __syncthreads();
  }
  if (MODE == 0 && G > 1) {   // W | Inv are free during a MODE 0 evaluation: scratch for the partial blocks of groups 1..
    double* scratch = &sh.W[0][0];
    if (tid >= NBLK && tid < G * NBLK) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) scratch[((grp - 1) * NBLK + blk) * 16 + i * 4 + j] = acc[i][j];
    }
    __syncthreads();
    if (tid < NBLK) {
      for (int g = 1; g < G; ++g)
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] += scratch[((g - 1) * NBLK + blk) * 16 + i * 4 + j];
    }
  }
  if (MODE == 0 && tid < NBLK) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int p = 4 * bi + i, c = 4 * bj + j;
        if (c <= p) {
          sh.F[p][c] = acc[i][j];
          sh.F[c][p] = acc[i][j];
        }
      }
  }
  if (sk >= 0 && sk < MM) sh.score[sk] = sc;
  if (MODE == 0) {
    ll = warp_sum(ll);
    if (lane == 0) sh.red[warp][0] = ll;
    __syncthreads();
    if (tid == 0) {
      double t = 0.0;
      for (int w = 0; w < LT / 32; ++w) t += sh.red[w][0];
      sh.loglik = t;
    }
  }
  __syncthreads();
}

template <int MM, int MODE>
__device__ __forceinline__ void eval_any(const LogitArgs& a, Shared<MM>& sh, const uint32_t* row, const double* drow, int m, int m0,
                                         bool want_loglik) {
  if constexpr (MM <= 20) {
    eval_pass<MM, MODE>(a, sh, row, drow, m, m0, want_loglik);
  } else {
    extern __shared__ __align__(16) unsigned char s_dyn[];
    Tile<MM>& tl = *reinterpret_cast<Tile<MM>*>(s_dyn + (sizeof(Shared<MM>) + 15) / 16 * 16);
    eval_pass_tiled<MM, MODE>(a, sh, tl, row, drow, m, m0, want_loglik);
  }
}

template <int MM>
constexpr size_t logit_smem_bytes() {
  return (sizeof(Shared<MM>) + 15) / 16 * 16 + (MM > 20 ? sizeof(Tile<MM>) : 0);
}

template <int MM>
__global__ void __launch_bounds__(LT) logit_fit_kernel(LogitArgs a) {
  extern __shared__ __align__(16) unsigned char s_dyn[];
  Shared<MM>& sh = *reinterpret_cast<Shared<MM>*>(s_dyn);
  const int K = a.K, m = K + 1;
  const double* b0 = a.null_fit;
  const double* score0 = a.null_fit + K;
  const double* fisher0 = a.null_fit + 2 * K;
  const double loglk0 = a.null_fit[2 * K + K * K];
  for (int64_t v = blockIdx.x; v < a.M; v += gridDim.x) {
    const uint32_t* row = a.packed ? reinterpret_cast<const uint32_t*>(a.packed + v * a.stride) : nullptr;
    const double* drow = a.packed ? nullptr : a.dense + v * a.ldx;
    // ---- exact genotype counts over the complete samples -> mean of the defined calls (RU:33-52) ----
    if (threadIdx.x < 3) sh.counts[threadIdx.x] = 0;
    if (threadIdx.x == 0) sh.loglik = 0.0;
    __syncthreads();
    if (drow) {   // dense entries: sum and count of the defined ones
      double sum = 0.0;
      int nd = 0;
      for (int i = threadIdx.x; i < a.n; i += LT) {
        const double xv = __ldg(drow + __ldg(a.idx + i));
        if (xv == xv) { sum += xv; ++nd; }
      }
      sum = warp_sum(sum);
      nd = __reduce_add_sync(0xffffffffu, nd);
      if ((threadIdx.x & 31) == 0) {
        atomicAdd(&sh.counts[0], nd);
        sh.red[threadIdx.x >> 5][0] = sum;
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < LT / 32; ++w) t += sh.red[w][0];
        sh.loglik = t;   // scratch: the sum of the defined entries
      }
    } else {
      int n1 = 0, n2 = 0, nm = 0;
      for (int i = threadIdx.x; i < a.n; i += LT) {
        const int s = __ldg(a.idx + i);
        const uint32_t code = (__ldg(row + (s >> 4)) >> sample_shift(s & 15)) & 3u;
        n1 += code == 1u;
        n2 += code == 2u;
        nm += code == 3u;
      }
      n1 = __reduce_add_sync(0xffffffffu, n1);
      n2 = __reduce_add_sync(0xffffffffu, n2);
      nm = __reduce_add_sync(0xffffffffu, nm);
      if ((threadIdx.x & 31) == 0) {
        atomicAdd(&sh.counts[0], n1);
        atomicAdd(&sh.counts[1], n2);
        atomicAdd(&sh.counts[2], nm);
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      sh.mean = drow ? sh.loglik / (double)sh.counts[0]
                     : (double)(sh.counts[0] + 2 * sh.counts[1]) / (double)(a.n - sh.counts[2]);
      sh.status = 0;
    }
    if (threadIdx.x < MM) sh.b[threadIdx.x] = ((int)threadIdx.x < K) ? b0[threadIdx.x] : 0.0;
    __syncthreads();

    int iter = 0;
    double beta = nan_value(), se = nan_value(), z = nan_value(), chi2 = nan_value(), pv = nan_value();
    bool converged = false, exploded = false;

    if (a.test != 3) {
      // =============== Wald / LRT: LogisticRegressionModel.fit(Some(nullFit)) ===============
      eval_any<MM, 0>(a, sh, row, drow, m, m, false);
      if (threadIdx.x == 0) {   // the covariate blocks of the first step are the null fit's (:311-325)
        for (int i = 0; i < K; ++i) {
          sh.score[i] = score0[i];
          for (int j = 0; j < K; ++j) sh.F[i][j] = fisher0[i * K + j];
        }
      }
      __syncthreads();
      while (true) {
        if (iter >= a.max_iter) break;
        ++iter;
        if (threadIdx.x == 0) {
          for (int i = 0; i < m; ++i) {
            sh.delta[i] = sh.score[i];
            for (int j = 0; j < m; ++j) sh.W[i][j] = sh.F[i][j];
          }
          int st = 0;
          if (!lu_solve<MM>(sh.W, sh.delta, m)) {
            st = 2;                                   // MatrixSingularException -> exploded (:360-361)
          } else if (isnan(sh.delta[0])) {
            st = 2;
          } else {
            double mx = 0.0;
            bool any_nan = false;
            for (int i = 0; i < m; ++i) {
              mx = fmax(mx, fabs(sh.delta[i]));   // breeze max(abs(.)): NaN entries other than [0] compare false
              any_nan |= isnan(sh.delta[i]);
            }
            if (mx < a.tol && !any_nan) st = 1;
            else
              for (int i = 0; i < m; ++i) sh.b[i] += sh.delta[i];
          }
          sh.status = st;
        }
        __syncthreads();
        const int st = sh.status;
        if (st == 1) { converged = true; break; }
        if (st == 2) { exploded = true; break; }
        eval_any<MM, 0>(a, sh, row, drow, m, m, false);
      }
      if (converged) {
        if (a.test == 1) {
          if (threadIdx.x == 0) {
            for (int i = 0; i < m; ++i)
              for (int j = 0; j < m; ++j) sh.W[i][j] = sh.F[i][j];
            sh.status = invert<MM>(sh.W, sh.Inv, m, &sh.logdet) ? 1 : 2;
          }
          __syncthreads();
          if (sh.status == 1) {
            beta = sh.b[K];
            se = sqrt(sh.Inv[K][K]);
            z = beta / se;
            pv = erfc(fabs(z) * 0.70710678118654752440);   // 2 pnorm(-|z|)
          }
        } else {
          eval_any<MM, 0>(a, sh, row, drow, m, m, true);         // logLkhd at the final mu (:365)
          beta = sh.b[K];
          chi2 = 2.0 * (sh.loglik - loglk0);
          pv = chi2 > 0.0 ? erfc(sqrt(0.5 * chi2)) : (chi2 == chi2 ? 1.0 : chi2);   // pchisqtail(chi2, 1)
        }
      }
    } else {
      // =============== Firth: fitFirth with K, then K + 1 free coefficients (:155-199, :372-408) ===============
      double ll_null = 0.0;
      for (int stage = 0; stage < 2; ++stage) {
        const int m0 = K + stage;
        iter = 0;
        converged = exploded = false;
        while (!converged && !exploded && iter < a.max_iter) {
          ++iter;
          eval_any<MM, 0>(a, sh, row, drow, m, m0, true);         // F = X' W X over all m columns at mu(b[0..m0))
          if (threadIdx.x == 0) {
            for (int i = 0; i < m; ++i)
              for (int j = 0; j < m; ++j) sh.W[i][j] = sh.F[i][j];
            sh.status = invert<MM>(sh.W, sh.Inv, m, &sh.logdet) ? 0 : 2;
          }
          __syncthreads();
          if (sh.status == 2) { exploded = true; break; }
          const double ll_here = sh.loglik + 0.5 * sh.logdet;   // + sum log|diag R| (:397-399)
          eval_any<MM, 1>(a, sh, row, drow, m, m0, false);
          if (threadIdx.x == 0) {
            for (int i = 0; i < m0; ++i) {
              sh.delta[i] = sh.score[i];
              for (int j = 0; j < m0; ++j) sh.W[i][j] = sh.F[i][j];
            }
            int st = 0;
            if (!lu_solve<MM>(sh.W, sh.delta, m0)) {
              st = 2;
            } else if (isnan(sh.delta[0])) {
              st = 2;
            } else {
              double mx = 0.0;
              bool any_nan = false;
              for (int i = 0; i < m0; ++i) {
                mx = fmax(mx, fabs(sh.delta[i]));
                any_nan |= isnan(sh.delta[i]);
              }
              if (mx < a.tol && !any_nan && iter > 1) st = 1;
              else
                for (int i = 0; i < m0; ++i) sh.b[i] += sh.delta[i];
            }
            sh.status = st;
          }
          __syncthreads();
          if (sh.status == 1) {
            converged = true;
            if (stage == 0) ll_null = ll_here;
            else {
              beta = sh.b[K];
              chi2 = 2.0 * (ll_here - ll_null);
              pv = chi2 > 0.0 ? erfc(sqrt(0.5 * chi2)) : (chi2 == chi2 ? 1.0 : chi2);
            }
          } else if (sh.status == 2) {
            exploded = true;
          }
          __syncthreads();
        }
        if (!converged) break;   // the null Firth fit did not converge: its fit record is reported (:196-197)
        if (stage == 0 && threadIdx.x == 0) sh.b[K] = 0.0;
        __syncthreads();
      }
    }

    if (threadIdx.x == 0) {
      if (a.out.beta) a.out.beta[v] = beta;
      if (a.out.standard_error) a.out.standard_error[v] = se;
      if (a.out.z_stat) a.out.z_stat[v] = z;
      if (a.out.chi_sq_stat) a.out.chi_sq_stat[v] = chi2;
      if (a.out.p_value) a.out.p_value[v] = pv;
      if (a.out.n_iterations) a.out.n_iterations[v] = iter;
      if (a.out.converged) a.out.converged[v] = converged ? 1 : 0;
      if (a.out.exploded) a.out.exploded[v] = exploded ? 1 : 0;
    }
    __syncthreads();
  }
}

template <int MM>
cudaError_t launch_mm(const LogitArgs& a, int grid, cudaStream_t st) {
  constexpr size_t smem = logit_smem_bytes<MM>();
  if (smem > 48 * 1024) {   // per device, so set on every launch (a second GPU in the same process has its own attribute)
    const cudaError_t e = cudaFuncSetAttribute(logit_fit_kernel<MM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  logit_fit_kernel<MM><<<grid, LT, smem, st>>>(a);
  return cudaSuccess;
}

void free_model(LogitModel* m) {
  if (!m) return;
  cudaFree(m->d_idx);
  cudaFree(m->d_cov);
  cudaFree(m->d_y);
  cudaFree(m->d_null);
  delete m;
}

}  // namespace

void logit_release(Ctx* c) {
  free_model(reinterpret_cast<LogitModel*>(c->logit_state));
  c->logit_state = nullptr;
}

int logit_set_model(Ctx* c, int64_t n_samples_total, int32_t n, int32_t K, const int32_t* idx, const double* cov,
                    const double* y, const double* b0, const double* score0, const double* fisher0, double loglk0) {
  if (n_samples_total <= 0 || n <= 0 || n > n_samples_total) return fail(c, LRR_EINVAL, "lrr_set_logit_model: bad sample counts");
  if (K < 1) return fail(c, LRR_EINVAL, "logistic regression requires at least one covariate expression");
  if (K + 1 > 64) return fail(c, LRR_EINVAL, "lrr_set_logit_model: at most 63 covariates");
  if (n - K - 1 < 1) {
    char buf[160];
    snprintf(buf, sizeof buf, "%d samples and %d %s (including x) implies %d degrees of freedom.", n, K + 1,
             K == 1 ? "covariate" : "covariates", n - K - 1);
    return fail(c, LRR_EINVAL, buf);
  }
  if (!idx || !cov || !y || !b0 || !score0 || !fisher0) return fail(c, LRR_EINVAL, "lrr_set_logit_model: NULL input array");
  logit_release(c);
  LogitModel* m = new LogitModel();
  c->logit_state = m;
  m->n = n;
  m->K = K;
  m->n_samples_total = n_samples_total;
  std::vector<double> nf(2 * (size_t)K + (size_t)K * K + 1);
  for (int i = 0; i < K; ++i) nf[i] = b0[i];
  for (int i = 0; i < K; ++i) nf[K + i] = score0[i];
  for (int i = 0; i < K * K; ++i) nf[2 * K + i] = fisher0[i];
  nf[2 * K + K * K] = loglk0;
  LRR_CUDA(c, cudaMalloc(&m->d_idx, sizeof(int32_t) * (size_t)n));
  LRR_CUDA(c, cudaMemcpy(m->d_idx, idx, sizeof(int32_t) * (size_t)n, cudaMemcpyDefault));
  LRR_CUDA(c, cudaMalloc(&m->d_cov, sizeof(double) * (size_t)K * n));
  LRR_CUDA(c, cudaMemcpy(m->d_cov, cov, sizeof(double) * (size_t)K * n, cudaMemcpyDefault));
  LRR_CUDA(c, cudaMalloc(&m->d_y, sizeof(double) * (size_t)n));
  LRR_CUDA(c, cudaMemcpy(m->d_y, y, sizeof(double) * (size_t)n, cudaMemcpyDefault));
  LRR_CUDA(c, cudaMalloc(&m->d_null, sizeof(double) * nf.size()));
  LRR_CUDA(c, cudaMemcpy(m->d_null, nf.data(), sizeof(double) * nf.size(), cudaMemcpyHostToDevice));
  return LRR_OK;
}

int logit_run(Ctx* c, const uint8_t* d_packed, const double* d_dense, int64_t M, int64_t stride, int64_t n_samples_total,
              int test, int max_iter, double tol, const lrr_logit_out& out, cudaStream_t st) {
  LogitModel* m = reinterpret_cast<LogitModel*>(c->logit_state);
  if (!m) return fail(c, LRR_ESTATE, "lrr_run_logit: call lrr_set_logit_model first");
  if (test < 1 || test > 3) return fail(c, LRR_EINVAL, "lrr_run_logit: test must be LRR_LOGIT_WALD, _LRT or _FIRTH");
  if (M < 0 || max_iter < 0 || !(tol > 0.0)) return fail(c, LRR_EINVAL, "lrr_run_logit: bad arguments");
  if (n_samples_total != m->n_samples_total) return fail(c, LRR_EINVAL, "lrr_run_logit: n_samples_total differs from the model's");
  if (d_dense ? stride < n_samples_total : (stride % 4 != 0 || stride * 4 < n_samples_total))
    return fail(c, LRR_EINVAL, "lrr_run_logit: bad row stride");
  if (M == 0) return LRR_OK;
  if (!d_packed && !d_dense) return fail(c, LRR_EINVAL, "lrr_run_logit: the row pointer is NULL");
  LogitArgs a;
  a.packed = d_dense ? nullptr : d_packed;
  a.dense = d_dense;
  a.ldx = stride;
  a.M = M;
  a.stride = stride;
  a.n = m->n;
  a.K = m->K;
  a.idx = m->d_idx;
  a.cov = m->d_cov;
  a.y = m->d_y;
  a.null_fit = m->d_null;
  a.test = test;
  a.max_iter = max_iter;
  a.tol = tol;
  a.out = out;
  const int64_t want = M < (int64_t)c->sm_count * 8 ? M : (int64_t)c->sm_count * 8;
  const int grid = (int)want;
  const int mm = m->K + 1;
  cudaError_t le;
  int tiled_from = 21;   // smallest m that takes the tiled form
  if (const char* e = tuning_env("LRR_LOGIT_TILED_FROM")) tiled_from = atoi(e);
  if (mm >= tiled_from && mm <= 32) le = launch_mm<32>(a, grid, st);
  else if (mm <= 2) le = launch_mm<2>(a, grid, st);
  else if (mm <= 3) le = launch_mm<3>(a, grid, st);
  else if (mm <= 4) le = launch_mm<4>(a, grid, st);
  else if (mm <= 5) le = launch_mm<5>(a, grid, st);
  else if (mm <= 6) le = launch_mm<6>(a, grid, st);
  else if (mm <= 8) le = launch_mm<8>(a, grid, st);
  else if (mm <= 10) le = launch_mm<10>(a, grid, st);
  else if (mm <= 12) le = launch_mm<12>(a, grid, st);
  else if (mm <= 16) le = launch_mm<16>(a, grid, st);
  else if (mm <= 20) le = launch_mm<20>(a, grid, st);
  else if (mm <= 32) le = launch_mm<32>(a, grid, st);   // tiled form: Fisher blocks in registers, X tiles in shared memory
  else if (mm <= 48) le = launch_mm<48>(a, grid, st);
  else le = launch_mm<64>(a, grid, st);
  LRR_CUDA(c, le);
  c->launches++;
  LRR_CUDA(c, cudaGetLastError());
  return LRR_OK;
}

}  // namespace lrr
