// Kernel 1: CUDA-core float64 sweep.  One warp owns VW variants; a CTA of 8 warps shares a staged block of
// the basis [Q' | Y_res] in shared memory.  Restates, per variant,
//   RegressionUtils.setMeanImputedDoubles      hail/hail/src/is/hail/stats/RegressionUtils.scala:16-58
//   qtx = Qt * X, ytx = y^T X, x.x             hail/hail/src/is/hail/methods/LinearRegression.scala:136-146
// with two differences that do not change the mathematics: genotype counts (n_het, n_homvar, n_missing) are
// exact integer popcounts, and the projection uses the covariate-residualised y (xyp = y_res . x) so that
// y_transpose_x is rebuilt as xyp + Qty . qtx in the epilogue.
#include "common.cuh"

namespace lrr {

namespace {

constexpr int VW = 4;        // variants per warp
constexpr int WARPS = 8;     // warps per CTA
constexpr int CB = 12;       // basis columns per pass
constexpr int SB = 512;      // samples per staged block (= 32 packed words)

struct SweepArgs {
  const uint8_t* packed;
  int64_t M;
  int64_t stride;
  int64_t ns_pad;
  const double* basis;   // [C][ns_pad]
  const uint32_t* mask;  // [ns_pad / 16]
  int n;
  int C;
  int32_t* counts;       // [M][4]
  double* dots;          // [M][C]
  int sq_col;            // column accumulated against x^2 instead of x (weighted groups: sum_j w_j x_j^2), or -1
};

__device__ __forceinline__ int warp_sum(int v) { return __reduce_add_sync(0xffffffffu, v); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(WARPS * 32, 1) fp64_sweep_kernel(SweepArgs a) {
  extern __shared__ double s_basis[];  // [CB][SB]
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int64_t v0 = ((int64_t)blockIdx.x * WARPS + warp) * VW;
  const int64_t words_per_row = a.stride / 4;

  const uint32_t* rows[VW];
#pragma unroll
  for (int v = 0; v < VW; ++v) {
    int64_t vv = v0 + v;
    if (vv >= a.M) vv = a.M - 1;  // clamp: loads stay in bounds, stores are skipped
    rows[v] = reinterpret_cast<const uint32_t*>(a.packed + vv * a.stride);
  }

  // ---- phase A: exact counts over the group's samples, mean of the defined calls (RU:33-52) ----
  double mean[VW];
#pragma unroll
  for (int v = 0; v < VW; ++v) {
    int n1 = 0, n2 = 0, nm = 0;
    const uint4* row4 = reinterpret_cast<const uint4*>(rows[v]);
    const uint4* mask4 = reinterpret_cast<const uint4*>(a.mask);
    for (int64_t q = lane; q < words_per_row / 4; q += 32) {
      const uint4 w = __ldg(row4 + q);
      const uint4 m = __ldg(mask4 + q);
      const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
      const uint32_t mm[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t lo = ww[k] & mm[k];         // mask has the low bit of each kept field set
        const uint32_t hi = (ww[k] >> 1) & mm[k];
        n1 += __popc(lo & ~hi);
        n2 += __popc(hi & ~lo);
        nm += __popc(hi & lo);
      }
    }
    n1 = warp_sum(n1);
    n2 = warp_sum(n2);
    nm = warp_sum(nm);
    const double s = (double)(n1 + 2 * n2);
    mean[v] = s / (double)(a.n - nm);  // 0/0 -> NaN for an all-missing variant, as RU:52
    if (lane == 0 && v0 + v < a.M) {
      int4 c = make_int4(n1, n2, nm, 0);
      reinterpret_cast<int4*>(a.counts)[v0 + v] = c;
    }
  }

  // ---- phase B: dot products of the imputed column with every basis column ----
  const int sh = sample_shift(lane & 15);
  const int64_t n_blocks = a.ns_pad / SB;
  for (int c0 = 0; c0 < a.C; c0 += CB) {
    const int cb = min(CB, a.C - c0);
    double acc[VW][CB];
#pragma unroll
    for (int v = 0; v < VW; ++v)
#pragma unroll
      for (int c = 0; c < CB; ++c) acc[v][c] = 0.0;

    for (int64_t blk = 0; blk < n_blocks; ++blk) {
      __syncthreads();
      for (int i = threadIdx.x; i < cb * SB; i += WARPS * 32) {
        const int c = i / SB, j = i - c * SB;
        s_basis[c * SB + j] = __ldg(a.basis + (int64_t)(c0 + c) * a.ns_pad + blk * SB + j);
      }
      __syncthreads();
      uint32_t wv[VW];
#pragma unroll
      for (int v = 0; v < VW; ++v) wv[v] = __ldg(rows[v] + blk * 32 + lane);
#pragma unroll 4
      for (int r = 0; r < 16; ++r) {
        double q[CB];
#pragma unroll
        for (int c = 0; c < CB; ++c) q[c] = (c < cb) ? s_basis[c * SB + 32 * r + lane] : 0.0;
        const int src = 2 * r + (lane >> 4);
#pragma unroll
        for (int v = 0; v < VW; ++v) {
          const uint32_t w = __shfl_sync(0xffffffffu, wv[v], src);
          const uint32_t code = (w >> sh) & 3u;
          const double x = (code == 3u) ? mean[v] : (double)code;
#pragma unroll
          for (int c = 0; c < CB; ++c) acc[v][c] = fma(q[c], (c0 + c == a.sq_col) ? x * x : x, acc[v][c]);
        }
      }
    }
#pragma unroll
    for (int v = 0; v < VW; ++v) {
#pragma unroll
      for (int c = 0; c < CB; ++c) {
        if (c < cb) {
          const double s = warp_sum(acc[v][c]);
          if (lane == 0 && v0 + v < a.M) a.dots[(v0 + v) * a.C + c0 + c] = s;
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Float64 recompute of LISTED rows: the tolerance guard of the quantised sweeps (stats_device.cuh) lists the rows whose
// rigorous error bound leaves the tolerance; they are few and scattered, so one CTA takes RV of them and splits the SAMPLE
// axis over its 8 warps (the sweep above gives a CTA 32 consecutive rows and would leave the device idle behind one CTA).
// Same arithmetic as fp64_sweep_kernel: exact integer counts, mean, float64 FMA dot products with the full-precision basis
// (read through L2, coalesced), fixed reduction order.
// ------------------------------------------------------------------------------------------------
constexpr int RV = 4;

struct RecomputeArgs {
  const uint8_t* packed;
  int64_t stride;
  int64_t ns_pad;
  const double* basis;
  const uint32_t* mask;
  int n;
  int C;
  int sq_col;
  int32_t* counts;
  double* dots;
  int dots_stride;
  const int32_t* list;
  const int32_t* count;
};

__global__ void __launch_bounds__(WARPS * 32, 1) fp64_recompute_kernel(RecomputeArgs a) {
  __shared__ int s_cnt[WARPS][RV][3];
  __shared__ double s_acc[WARPS][RV][CB];
  __shared__ double s_mean[RV];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int total = *a.count;
  const int64_t words_per_row = a.stride / 4;
  const int64_t n_blocks = a.ns_pad / SB;
  const int sh = sample_shift(lane & 15);
  for (int i0 = blockIdx.x * RV; i0 < total; i0 += gridDim.x * RV) {
    int64_t vrow[RV];
    const uint32_t* rows[RV];
#pragma unroll
    for (int v = 0; v < RV; ++v) {
      vrow[v] = a.list[min(i0 + v, total - 1)];
      rows[v] = reinterpret_cast<const uint32_t*>(a.packed + vrow[v] * a.stride);
    }
    // ---- exact counts over the group's samples ----
#pragma unroll
    for (int v = 0; v < RV; ++v) {
      int n1 = 0, n2 = 0, nm = 0;
      const uint4* row4 = reinterpret_cast<const uint4*>(rows[v]);
      const uint4* mask4 = reinterpret_cast<const uint4*>(a.mask);
      for (int64_t q = threadIdx.x; q < words_per_row / 4; q += WARPS * 32) {
        const uint4 w = __ldg(row4 + q);
        const uint4 m = __ldg(mask4 + q);
        const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
        const uint32_t mm[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t lo = ww[k] & mm[k];
          const uint32_t hi = (ww[k] >> 1) & mm[k];
          n1 += __popc(lo & ~hi);
          n2 += __popc(hi & ~lo);
          nm += __popc(hi & lo);
        }
      }
      n1 = warp_sum(n1);
      n2 = warp_sum(n2);
      nm = warp_sum(nm);
      if (lane == 0) { s_cnt[warp][v][0] = n1; s_cnt[warp][v][1] = n2; s_cnt[warp][v][2] = nm; }
    }
    __syncthreads();
    if (threadIdx.x < RV) {
      const int v = threadIdx.x;
      int n1 = 0, n2 = 0, nm = 0;
      for (int w = 0; w < WARPS; ++w) { n1 += s_cnt[w][v][0]; n2 += s_cnt[w][v][1]; nm += s_cnt[w][v][2]; }
      s_mean[v] = (double)(n1 + 2 * n2) / (double)(a.n - nm);
      if (i0 + v < total) reinterpret_cast<int4*>(a.counts)[vrow[v]] = make_int4(n1, n2, nm, 0);
    }
    __syncthreads();
    double mean[RV];
#pragma unroll
    for (int v = 0; v < RV; ++v) mean[v] = s_mean[v];
    // ---- dot products: warp w takes the 512-sample blocks w, w + 8, ... ----
    for (int c0 = 0; c0 < a.C; c0 += CB) {
      const int cb = min(CB, a.C - c0);
      double acc[RV][CB];
#pragma unroll
      for (int v = 0; v < RV; ++v)
#pragma unroll
        for (int c = 0; c < CB; ++c) acc[v][c] = 0.0;
      for (int64_t blk = warp; blk < n_blocks; blk += WARPS) {
        uint32_t wv[RV];
#pragma unroll
        for (int v = 0; v < RV; ++v) wv[v] = __ldg(rows[v] + blk * 32 + lane);
        const double* bp = a.basis + (int64_t)c0 * a.ns_pad + blk * SB + lane;
#pragma unroll 4
        for (int r = 0; r < 16; ++r) {
          double q[CB];
#pragma unroll
          for (int c = 0; c < CB; ++c) q[c] = (c < cb) ? __ldg(bp + (int64_t)c * a.ns_pad + 32 * r) : 0.0;
          const int src = 2 * r + (lane >> 4);
#pragma unroll
          for (int v = 0; v < RV; ++v) {
            const uint32_t w = __shfl_sync(0xffffffffu, wv[v], src);
            const uint32_t code = (w >> sh) & 3u;
            const double x = (code == 3u) ? mean[v] : (double)code;
#pragma unroll
            for (int c = 0; c < CB; ++c) acc[v][c] = fma(q[c], (c0 + c == a.sq_col) ? x * x : x, acc[v][c]);
          }
        }
      }
#pragma unroll
      for (int v = 0; v < RV; ++v)
#pragma unroll
        for (int c = 0; c < CB; ++c) {
          const double t = warp_sum(acc[v][c]);
          if (lane == 0) s_acc[warp][v][c] = t;
        }
      __syncthreads();
      if (threadIdx.x < RV * CB) {
        const int v = threadIdx.x / CB, c = threadIdx.x % CB;
        double t = 0.0;
        for (int w = 0; w < WARPS; ++w) t += s_acc[w][v][c];
        if (c < cb && i0 + v < total) a.dots[vrow[v] * a.dots_stride + c0 + c] = t;
      }
      __syncthreads();
    }
  }
}

}  // namespace

int launch_fp64_recompute(Ctx* c, int g, const uint8_t* d_packed, int64_t stride, const int32_t* d_list, const int32_t* d_count,
                          int dots_stride, cudaStream_t st) {
  const Group& G = c->groups[g];
  RecomputeArgs a;
  a.packed = d_packed;
  a.stride = stride;
  a.ns_pad = G.ns_pad;
  a.basis = G.d_basis;
  a.mask = G.d_mask;
  a.n = G.n;
  a.C = G.C;
  a.sq_col = G.weighted ? G.C - 1 : -1;
  a.counts = c->d_counts + (int64_t)g * c->reserved_variants * 4;
  a.dots = c->d_dots + c->dots_offset[g];
  a.dots_stride = dots_stride;
  a.list = d_list;
  a.count = d_count;
  fp64_recompute_kernel<<<c->sm_count * 2, WARPS * 32, 0, st>>>(a);
  c->launches++;
  LRR_CUDA(c, cudaGetLastError());
  return LRR_OK;
}

int launch_fp64_sweep(Ctx* c, const uint8_t* d_packed, int64_t M, int64_t stride, cudaStream_t st) {
  if (M == 0) return LRR_OK;
  const int smem = CB * SB * (int)sizeof(double);
  // per device, so not cached in a process-wide flag
  LRR_CUDA(c, cudaFuncSetAttribute(fp64_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  for (size_t g = 0; g < c->groups.size(); ++g) {
    const Group& G = c->groups[g];
    SweepArgs a;
    a.packed = d_packed;
    a.M = M;
    a.stride = stride;
    a.ns_pad = G.ns_pad;
    a.basis = G.d_basis;
    a.mask = G.d_mask;
    a.n = G.n;
    a.C = G.C;
    a.sq_col = G.weighted ? G.C - 1 : -1;
    a.counts = c->d_counts + (int64_t)g * c->reserved_variants * 4;
    a.dots = c->d_dots + c->dots_offset[g];
    const int64_t per_cta = (int64_t)WARPS * VW;
    const int grid = (int)((M + per_cta - 1) / per_cta);
    fp64_sweep_kernel<<<grid, WARPS * 32, smem, st>>>(a);
    c->launches++;
    LRR_CUDA(c, cudaGetLastError());
  }
  return LRR_OK;
}

}  // namespace lrr
