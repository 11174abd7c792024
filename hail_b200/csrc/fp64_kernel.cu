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
// rigorous error bound leaves the tolerance.  They are few and scattered, and their number is known only on the device,
// so a fixed grid walks a list of work items:
//   * more than RC_SPLIT_ROWS listed rows: one item = one row; a CTA streams the whole basis once for it;
//   * fewer: one item = 1 / RC_SPLITS of a row's samples, so that even a single listed row is spread over 32 SMs
//     (a lone CTA needs ~3 ms to pull 35 MB of basis through one SM); the partial sums of a row are combined, in split
//     order, by whichever CTA finishes the row's last split (deterministic).
// Arithmetic: exact integer counts; for every column the float64 sums D_c = sum over the defined calls of code * b_j and
// D_m = sum over the missing calls of b_j; the mean-imputed dot product is D_c + mean * D_m (RU:16-58, LR:139-146).
// A thread takes one packed word = 16 consecutive samples and reads each column's 16 values as one 128-byte run, so a
// warp has 4 KB per column in flight.
// ------------------------------------------------------------------------------------------------
constexpr int RC_THREADS = 256;
constexpr int RC_SPLITS = 32;
constexpr int RC_SPLIT_ROWS = 256;
constexpr int RC_CB = 8;                 // columns per pass: 2 * 8 accumulators + 16 staged values per thread
constexpr int RC_VALS = 2 * RC_CB + 4;   // doubles per (row, pass, split) partial: D_c, D_m, then n1, n2, nm as doubles

struct RecomputeArgs {
  const uint8_t* packed;
  int64_t stride;
  int64_t ns_pad;
  const double* basis;
  const uint32_t* mask;
  int n;
  int C;
  int n_pass;
  int32_t* counts;
  double* dots;
  int dots_stride;
  const int32_t* list;
  const int32_t* count;
  double* partial;     // [RC_SPLIT_ROWS][n_pass][RC_SPLITS][RC_VALS]
  int32_t* arrive;     // [RC_SPLIT_ROWS], zero between launches
};

__global__ void __launch_bounds__(RC_THREADS) fp64_recompute_kernel(RecomputeArgs a) {
  __shared__ double s_red[RC_THREADS / 32][RC_VALS];
  __shared__ double s_tot[RC_VALS];
  __shared__ int s_ticket;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int total = *a.count;
  if (total <= 0) return;
  const int splits = total > RC_SPLIT_ROWS ? 1 : RC_SPLITS;
  const int64_t n_words = a.ns_pad / 16;
  const int64_t words_per_split = (n_words + splits - 1) / splits;
  const int64_t n_items = (int64_t)total * splits;
  for (int64_t item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int li = (int)(item / splits), sp = (int)(item - (int64_t)li * splits);
    const int64_t v = a.list[li];
    const uint32_t* row = reinterpret_cast<const uint32_t*>(a.packed + v * a.stride);
    const int64_t w0 = sp * words_per_split, w1 = min(n_words, w0 + words_per_split);
    double mean = 0.0;
    int cnt1 = 0, cnt2 = 0, cntm = 0;
    for (int pass = 0; pass < a.n_pass; ++pass) {
      const int c0 = pass * RC_CB, cb = min(RC_CB, a.C - c0);
      double dc[RC_CB], dm[RC_CB];
#pragma unroll
      for (int c = 0; c < RC_CB; ++c) dc[c] = dm[c] = 0.0;
      int n1 = 0, n2 = 0, nm = 0;
      for (int64_t w = w0 + threadIdx.x; w < w1; w += RC_THREADS) {
        const uint32_t word = __ldg(row + w), mk = __ldg(a.mask + w);   // mask: low bit of each kept field
        const uint32_t lo = word & mk, hi = (word >> 1) & mk;
        if (pass == 0) {
          n1 += __popc(lo & ~hi);
          n2 += __popc(hi & ~lo);
          nm += __popc(hi & lo);
        }
        const double* bp = a.basis + (int64_t)c0 * a.ns_pad + 16 * w;
#pragma unroll
        for (int c = 0; c < RC_CB; ++c) {
          if (c < cb) {
            double b[16];
            const double2* p2 = reinterpret_cast<const double2*>(bp + (int64_t)c * a.ns_pad);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const double2 t = __ldg(p2 + q);
              b[2 * q] = t.x;
              b[2 * q + 1] = t.y;
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const uint32_t code = (word >> sample_shift(j)) & 3u;   // samples outside the group have b == 0
              if (code == 3u) dm[c] += b[j]; else dc[c] = fma((double)code, b[j], dc[c]);
            }
          }
        }
      }
      // ---- block reduction in a fixed order ----
#pragma unroll
      for (int c = 0; c < RC_CB; ++c) {
        const double t0 = warp_sum(dc[c]), t1 = warp_sum(dm[c]);
        if (lane == 0) { s_red[warp][c] = t0; s_red[warp][RC_CB + c] = t1; }
      }
      if (pass == 0) {
        n1 = warp_sum(n1); n2 = warp_sum(n2); nm = warp_sum(nm);
        if (lane == 0) { s_red[warp][2 * RC_CB] = (double)n1; s_red[warp][2 * RC_CB + 1] = (double)n2; s_red[warp][2 * RC_CB + 2] = (double)nm; }
      }
      __syncthreads();
      if (threadIdx.x < RC_VALS) {
        double t = 0.0;
        if (threadIdx.x < 2 * RC_CB || (pass == 0 && threadIdx.x < 2 * RC_CB + 3))
          for (int w = 0; w < RC_THREADS / 32; ++w) t += s_red[w][threadIdx.x];
        s_tot[threadIdx.x] = t;
      }
      __syncthreads();
      if (splits == 1) {
        // the whole row is here: finish it
        if (pass == 0) {
          cnt1 = (int)s_tot[2 * RC_CB]; cnt2 = (int)s_tot[2 * RC_CB + 1]; cntm = (int)s_tot[2 * RC_CB + 2];
          mean = (double)(cnt1 + 2 * cnt2) / (double)(a.n - cntm);   // 0 / 0 -> NaN for an all-missing row, as RU:52
          if (threadIdx.x == 0) reinterpret_cast<int4*>(a.counts)[v] = make_int4(cnt1, cnt2, cntm, 0);
        }
        if ((int)threadIdx.x < cb) {
          const double d_c = s_tot[threadIdx.x], d_m = s_tot[RC_CB + threadIdx.x];
          a.dots[v * a.dots_stride + c0 + threadIdx.x] = cntm > 0 ? d_c + mean * d_m : d_c;
        }
      } else if (threadIdx.x < RC_VALS) {
        a.partial[(((int64_t)li * a.n_pass + pass) * RC_SPLITS + sp) * RC_VALS + threadIdx.x] = s_tot[threadIdx.x];
      }
      __syncthreads();
    }
    if (splits > 1) {
      // the CTA that completes the row's last split combines the partial sums (split order) and writes the row
      __threadfence();
      if (threadIdx.x == 0) s_ticket = atomicAdd(a.arrive + li, 1);
      __syncthreads();
      if (s_ticket == RC_SPLITS - 1) {
        __threadfence();
        const double* pr = a.partial + (int64_t)li * a.n_pass * RC_SPLITS * RC_VALS;
        double f1 = 0.0, f2 = 0.0, fm = 0.0;
        for (int s2 = 0; s2 < RC_SPLITS; ++s2) {
          f1 += __ldcg(pr + s2 * RC_VALS + 2 * RC_CB);
          f2 += __ldcg(pr + s2 * RC_VALS + 2 * RC_CB + 1);
          fm += __ldcg(pr + s2 * RC_VALS + 2 * RC_CB + 2);
        }
        const int t1 = (int)f1, t2 = (int)f2, tm = (int)fm;
        const double mn = (double)(t1 + 2 * t2) / (double)(a.n - tm);
        if (threadIdx.x == 0) {
          reinterpret_cast<int4*>(a.counts)[v] = make_int4(t1, t2, tm, 0);
          a.arrive[li] = 0;   // ready for the next launch
        }
        for (int c = threadIdx.x; c < a.C; c += RC_THREADS) {
          const double* pc = pr + (int64_t)(c / RC_CB) * RC_SPLITS * RC_VALS + (c % RC_CB);
          double d_c = 0.0, d_m = 0.0;
          for (int s2 = 0; s2 < RC_SPLITS; ++s2) {
            d_c += __ldcg(pc + s2 * RC_VALS);
            d_m += __ldcg(pc + s2 * RC_VALS + RC_CB);
          }
          a.dots[v * a.dots_stride + c] = tm > 0 ? d_c + mn * d_m : d_c;
        }
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
  a.n_pass = (G.C + RC_CB - 1) / RC_CB;
  a.counts = c->d_counts + (int64_t)g * c->reserved_variants * 4;
  a.dots = c->d_dots + c->dots_offset[g];
  a.dots_stride = dots_stride;
  a.list = d_list;
  a.count = d_count;
  const size_t need = sizeof(double) * (size_t)RC_SPLIT_ROWS * (size_t)a.n_pass * RC_SPLITS * RC_VALS + sizeof(int32_t) * RC_SPLIT_ROWS;
  if (need > c->recompute_bytes) {
    LRR_CUDA(c, cudaStreamSynchronize(st));
    cudaFree(c->d_recompute);
    c->d_recompute = nullptr;
    c->recompute_bytes = 0;
    LRR_CUDA(c, cudaMalloc(&c->d_recompute, need));
    LRR_CUDA(c, cudaMemsetAsync(c->d_recompute, 0, need, st));
    c->recompute_bytes = need;
  }
  a.arrive = static_cast<int32_t*>(c->d_recompute);
  a.partial = reinterpret_cast<double*>(static_cast<char*>(c->d_recompute) + ((sizeof(int32_t) * RC_SPLIT_ROWS + 255) / 256 * 256));
  fp64_recompute_kernel<<<c->sm_count * 4, RC_THREADS, 0, st>>>(a);
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
