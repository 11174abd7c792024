// Kernel 6: the TRANSPOSED product of the packed genotype matrix -- out[j, c] = sum_v coef[v][code(v, j)] * t[v, c] --
// the half of PCA's power iteration (`hl.hwe_normalized_pca`, hail/python/hail/methods/pca.py:15-33, 345-372: G <- A' (A G))
// that the per-variant sweep does not provide.  A[v, j] = coef[v][code]: the caller tabulates the entry value of each
// of the four call codes per variant (HWE normalisation: (c - mean_v) / sd_v for c = 0, 1, 2 and 0 for a missing call,
// pca.py:26-31), so the kernel is independent of the normalisation.
//
// A CTA owns a strip of 512 samples (128 packed bytes of every row) and a contiguous range of variants; each thread owns
// ONE packed byte = 4 samples and keeps their 4 x L partial sums in float64 registers.  Rows are processed in blocks of
// 32: the block's t rows and coefficient tables are staged in shared memory (broadcast reads), the thread's bytes are
// loaded four rows ahead.  Compute-bound on the FP64 pipe (4 L FMAs per packed byte); genotype bytes are
// read once per variant split.  Partial results of the variant splits go to out[split][j][c]; the host adds them.
#include "common.cuh"

namespace lrr {

namespace {

constexpr int GT_THREADS = 128;
constexpr int RB = 32;   // rows per staged block

template <int LP>
__global__ void __launch_bounds__(GT_THREADS) at_times_kernel(const uint8_t* __restrict__ packed, int64_t M, int64_t stride,
                                                              int64_t n_total, const double* __restrict__ coef,
                                                              const double* __restrict__ t, int L, int64_t rows_per_split,
                                                              double* __restrict__ out) {
  static_assert(LP % 2 == 0, "t rows are read as double2");
  __shared__ __align__(16) double s_t[RB][LP];
  __shared__ double s_coef[RB][4];
  const int64_t byte = (int64_t)blockIdx.x * GT_THREADS + threadIdx.x;   // packed byte of every row owned by this thread
  const bool active = byte < stride;
  const int64_t v_lo = (int64_t)blockIdx.y * rows_per_split;
  const int64_t v_hi = min(M, v_lo + rows_per_split);
  const uint8_t* col = packed + (active ? byte : 0);
  double acc[4][LP];
#pragma unroll
  for (int s = 0; s < 4; ++s)
#pragma unroll
    for (int c = 0; c < LP; ++c) acc[s][c] = 0.0;

  for (int64_t v0 = v_lo; v0 < v_hi; v0 += RB) {
    const int nr = (int)min((int64_t)RB, v_hi - v0);
    __syncthreads();
    for (int i = threadIdx.x; i < RB * LP; i += GT_THREADS) {
      const int r = i / LP, c = i - r * LP;
      s_t[r][c] = (r < nr && c < L) ? t[(v0 + r) * L + c] : 0.0;
    }
    if (threadIdx.x < RB * 4) {
      const int r = threadIdx.x >> 2, k = threadIdx.x & 3;
      s_coef[r][k] = (r < nr) ? coef[(v0 + r) * 4 + k] : 0.0;   // rows past the end: every code maps to 0
    }
    __syncthreads();
    // four rows at a time; the next four bytes are loaded while the current ones are consumed
    uint32_t nb[4];
#pragma unroll
    for (int rr = 0; rr < 4; ++rr) nb[rr] = (rr < nr) ? col[(v0 + rr) * stride] : 0u;
#pragma unroll 1
    for (int g = 0; g < RB / 4; ++g) {
      uint32_t cur[4];
#pragma unroll
      for (int rr = 0; rr < 4; ++rr) cur[rr] = nb[rr];
      if (g + 1 < RB / 4) {
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) {
          const int r = 4 * (g + 1) + rr;
          nb[rr] = (r < nr) ? col[(v0 + r) * stride] : 0u;
        }
      }
#pragma unroll
      for (int rr = 0; rr < 4; ++rr) {
        const int r = 4 * g + rr;
        double a[4];
#pragma unroll
        for (int s = 0; s < 4; ++s) a[s] = s_coef[r][(cur[rr] >> (2 * s)) & 3u];
#pragma unroll
        for (int c = 0; c < LP; c += 2) {
          const double2 tv = *reinterpret_cast<const double2*>(&s_t[r][c]);
#pragma unroll
          for (int s = 0; s < 4; ++s) {
            acc[s][c] = fma(a[s], tv.x, acc[s][c]);
            acc[s][c + 1] = fma(a[s], tv.y, acc[s][c + 1]);
          }
        }
      }
    }
  }
  if (!active) return;
  // byte B of a row holds samples 16 (B >> 2) + (B & 3) + 4 s, s = 0..3 (common.cuh sample_shift)
  const int64_t j0 = 16 * (byte >> 2) + (byte & 3);
  double* o = out + (int64_t)blockIdx.y * n_total * L;
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    const int64_t j = j0 + 4 * s;
    if (j < n_total) {
#pragma unroll
      for (int c = 0; c < LP; ++c)
        if (c < L) o[j * L + c] = acc[s][c];
    }
  }
}

}  // namespace

int launch_at_times(Ctx* c, const uint8_t* d_packed, int64_t M, int64_t stride, int64_t n_total, const double* d_coef,
                    const double* d_t, int L, int n_splits, double* d_out, cudaStream_t st) {
  if (L < 1 || L > 24) return fail(c, LRR_EINVAL, "lrr_at_times: 1 <= L <= 24");
  if (n_splits < 1) return fail(c, LRR_EINVAL, "lrr_at_times: n_splits >= 1");
  const int64_t rows_per_split = ((M + n_splits - 1) / n_splits + RB - 1) / RB * RB;
  dim3 grid((unsigned)((stride + GT_THREADS - 1) / GT_THREADS), (unsigned)n_splits);
  const int64_t rps = rows_per_split > 0 ? rows_per_split : RB;
  if (L <= 4) at_times_kernel<4><<<grid, GT_THREADS, 0, st>>>(d_packed, M, stride, n_total, d_coef, d_t, L, rps, d_out);
  else if (L <= 8) at_times_kernel<8><<<grid, GT_THREADS, 0, st>>>(d_packed, M, stride, n_total, d_coef, d_t, L, rps, d_out);
  else if (L <= 12) at_times_kernel<12><<<grid, GT_THREADS, 0, st>>>(d_packed, M, stride, n_total, d_coef, d_t, L, rps, d_out);
  else if (L <= 16) at_times_kernel<16><<<grid, GT_THREADS, 0, st>>>(d_packed, M, stride, n_total, d_coef, d_t, L, rps, d_out);
  else if (L <= 20) at_times_kernel<20><<<grid, GT_THREADS, 0, st>>>(d_packed, M, stride, n_total, d_coef, d_t, L, rps, d_out);
  else at_times_kernel<24><<<grid, GT_THREADS, 0, st>>>(d_packed, M, stride, n_total, d_coef, d_t, L, rps, d_out);
  c->launches++;
  LRR_CUDA(c, cudaGetLastError());
  return LRR_OK;
}

}  // namespace lrr
