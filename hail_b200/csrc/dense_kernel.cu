// Kernel 4: float64 sweep over DENSE dosages -- x is any float64 entry expression (statgen.py:229, 391: e.g. the
// PL / GP dosages of test_statgen.py:286-364), one double per entry, NaN = missing.  Reference steps per variant:
//   RegressionUtils.setMeanImputedDoubles      hail/hail/src/is/hail/stats/RegressionUtils.scala:16-58
//   qtx = Qt * X, ytx = y^T X, x.x, sum(x)     hail/hail/src/is/hail/methods/LinearRegression.scala:136-146
//
// HBM-bound at 8 bytes per entry, so x is read exactly ONCE per group:
//   dense_sweep_kernel   a CTA owns VT = 8 variants, its 8 warps stride over 64-sample chunks (one 128-bit load per lane
//                        per variant, the basis columns [Q' | Y_res] read as 128-bit loads from L2 where they stay
//                        resident: 32 MB at 400k samples x 10 columns).  Per variant it accumulates, over the DEFINED
//                        entries of the group's samples: the count, the pivot-shifted sum and sum of squares, and the
//                        dot product with every basis column (missing entries contribute 0 here), and it writes one bit
//                        per (variant, sample) saying "missing and in the group" (1/64 of the traffic of x).
//   dense_impute_kernel  the mean-imputed column is x + mean * [missing]: a warp per variant walks the set bits and adds
//                        mean * sum_{missing} basis[c] to every dot product, reading a sample-major copy of the basis
//                        (one sample's C values are contiguous).  Cost is proportional to the number of missing entries.
// The statistics epilogue (stats_device.cuh, `dense`) then uses dots[C] = sum of the defined entries and
// dots[C + 1] = their centred sum of squares instead of the exact genotype counts of the packed kernels.
#include <type_traits>

#include "common.cuh"

namespace lrr {

namespace {

constexpr int VW = 4;        // variants per warp
constexpr int WARPS = 8;     // warps per CTA
constexpr int VT = VW * WARPS;  // variants per CTA
constexpr int CHUNK = 64;    // samples per step (two per lane)
constexpr int STAGES = 6;    // cp.async ring depth: (STAGES - 1) x 16 KB of x in flight per SM

struct DenseArgs {
  const double* x;       // [M][ldx]
  int64_t M, ldx, n_total;
  int64_t ns_pad;
  const double* basis;   // [C][ns_pad]
  const uint32_t* mask;  // [ns_pad / 16], bit sample_shift(j & 15) of word j >> 4 set iff sample j is in the group
  int64_t first_sample;  // lowest sample index of the group: its entry is the pivot of the shifted sums
  int n;
  int C;                 // all dot columns of the group
  int c0;                // first column of this pass
  int32_t* counts;       // [M][4]: (0, 0, n_missing, 0)
  double* dots;          // [M][C + 2]: C dot products, sum of the defined entries, their centred sum of squares
  uint2* nanmask;        // [M][n_chunks]: .x bit l = sample 64 k + 2 l missing (and in the group), .y = sample 64 k + 2 l + 1
  int64_t n_chunks;
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// 16- / 8-byte asynchronous global -> shared copies; `bytes` < size zero-fills the rest (samples past the last one)
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem),
               "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem, int bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem),
               "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// CB = basis columns of this pass (compile-time: the accumulators live in registers), FIRST = this pass also produces
// the counts / sums / missing bits, VEC = rows are 16-byte aligned (128-bit copies of x)
template <int CB, bool FIRST, bool VEC>
__global__ void __launch_bounds__(WARPS * 32, 1) dense_sweep_kernel(DenseArgs a) {
  extern __shared__ __align__(16) double s_ring[];   // STAGES x { x [VT][CHUNK], basis [CB][CHUNK] }
  constexpr int STAGE_DOUBLES = (VT + CB) * CHUNK;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int64_t v0 = (int64_t)blockIdx.x * VT;
  const int cb = min(CB, a.C - a.c0);

  auto issue = [&](int64_t chunk) {   // one stage: 64 samples of the CTA's 32 variants and of this pass's columns
    double* st = s_ring + (chunk % STAGES) * STAGE_DOUBLES;
    const int64_t j0 = chunk * CHUNK;
    if (VEC) {
#pragma unroll
      for (int i = 0; i < VT * (CHUNK / 2) / (WARPS * 32); ++i) {
        const int idx = threadIdx.x + i * WARPS * 32;
        const int r = idx >> 5, l2 = (idx & 31) * 2;
        int64_t vv = v0 + r;
        if (vv >= a.M) vv = a.M - 1;   // clamp: copies stay in bounds, stores are skipped
        const int64_t left = a.n_total - (j0 + l2);
        const int bytes = left >= 2 ? 16 : (left == 1 ? 8 : 0);
        cp_async16(st + r * CHUNK + l2, a.x + vv * a.ldx + (bytes ? j0 + l2 : 0), bytes);
      }
    } else {
#pragma unroll
      for (int i = 0; i < VT * CHUNK / (WARPS * 32); ++i) {
        const int idx = threadIdx.x + i * WARPS * 32;
        const int r = idx >> 6, l = idx & 63;
        int64_t vv = v0 + r;
        if (vv >= a.M) vv = a.M - 1;
        const int bytes = (j0 + l < a.n_total) ? 8 : 0;
        cp_async8(st + r * CHUNK + l, a.x + vv * a.ldx + (bytes ? j0 + l : 0), bytes);
      }
    }
    for (int idx = threadIdx.x; idx < cb * (CHUNK / 2); idx += WARPS * 32) {   // ns_pad >= 64 n_chunks: in bounds
      const int c = idx >> 5, l2 = (idx & 31) * 2;
      cp_async16(st + (VT + c) * CHUNK + l2, a.basis + (int64_t)(a.c0 + c) * a.ns_pad + j0 + l2, 16);
    }
  };

  double piv[VW];
#pragma unroll
  for (int r = 0; r < VW; ++r) {
    piv[r] = 0.0;
    if (FIRST) {
      int64_t vv = v0 + warp * VW + r;
      if (vv >= a.M) vv = a.M - 1;
      const double p = a.x[vv * a.ldx + a.first_sample];
      piv[r] = (p == p && fabs(p) <= 1.7e308) ? p : 0.0;
    }
  }
  double acc[VW][CB];
  double s[VW], ss[VW];
  int n_miss = 0;   // lane r counts variant r's missing entries
#pragma unroll
  for (int r = 0; r < VW; ++r) {
    s[r] = ss[r] = 0.0;
#pragma unroll
    for (int c = 0; c < CB; ++c) acc[r][c] = 0.0;
  }
  if (cb < CB) {   // unused columns of the last pass: zero once, never copied into
    for (int st = 0; st < STAGES; ++st)
      for (int i = threadIdx.x; i < (CB - cb) * CHUNK; i += WARPS * 32) s_ring[st * STAGE_DOUBLES + (VT + cb) * CHUNK + i] = 0.0;
  }

  const int sh0 = sample_shift((lane & 7) * 2), sh1 = sample_shift((lane & 7) * 2 + 1);
#pragma unroll
  for (int k = 0; k < STAGES - 1; ++k) {
    if (k < a.n_chunks) issue(k);
    cp_async_commit();
  }
  for (int64_t chunk = 0; chunk < a.n_chunks; ++chunk) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();   // stage `chunk` has landed for every thread; stage chunk - 1 is free again
    if (chunk + STAGES - 1 < a.n_chunks) issue(chunk + STAGES - 1);
    cp_async_commit();
    const double* st = s_ring + (chunk % STAGES) * STAGE_DOUBLES;
    const uint32_t mw = __ldg(a.mask + chunk * 4 + (lane >> 3));
    const bool g0 = (mw >> sh0) & 1u, g1 = (mw >> sh1) & 1u;
    double2 xv[VW];
#pragma unroll
    for (int r = 0; r < VW; ++r) xv[r] = *reinterpret_cast<const double2*>(st + (warp * VW + r) * CHUNK + lane * 2);
    double x0[VW], x1[VW];
#pragma unroll
    for (int r = 0; r < VW; ++r) {
      const bool m0 = xv[r].x != xv[r].x, m1 = xv[r].y != xv[r].y;
      const bool u0 = g0 && !m0, u1 = g1 && !m1;      // entries that enter the sums
      x0[r] = u0 ? xv[r].x : 0.0;
      x1[r] = u1 ? xv[r].y : 0.0;
      if (FIRST) {
        const double d0 = u0 ? x0[r] - piv[r] : 0.0, d1 = u1 ? x1[r] - piv[r] : 0.0;
        s[r] += d0 + d1;
        ss[r] = fma(d0, d0, fma(d1, d1, ss[r]));
        const uint32_t b0 = __ballot_sync(0xffffffffu, g0 && m0), b1 = __ballot_sync(0xffffffffu, g1 && m1);
        if (lane == r) {
          n_miss += __popc(b0) + __popc(b1);
          const int64_t vv = v0 + warp * VW + r;
          if (vv < a.M) a.nanmask[vv * a.n_chunks + chunk] = make_uint2(b0, b1);
        }
      }
    }
#pragma unroll
    for (int c = 0; c < CB; ++c) {
      const double2 q = *reinterpret_cast<const double2*>(st + (VT + c) * CHUNK + lane * 2);
#pragma unroll
      for (int r = 0; r < VW; ++r) acc[r][c] = fma(q.x, x0[r], fma(q.y, x1[r], acc[r][c]));
    }
  }

  // ---- reduce over the lanes: every warp owns its variants ----
#pragma unroll
  for (int r = 0; r < VW; ++r) {
    const int64_t vv = v0 + warp * VW + r;
    double* d = a.dots + vv * (a.C + 2);
#pragma unroll
    for (int c = 0; c < CB; ++c) {
      const double t = warp_sum(acc[r][c]);
      if (lane == 0 && c < cb && vv < a.M) d[a.c0 + c] = t;
    }
    if (FIRST) {
      const double rs = warp_sum(s[r]), rss = warp_sum(ss[r]);
      const int nm = __shfl_sync(0xffffffffu, n_miss, r);
      if (lane == 0 && vv < a.M) {
        const int ndv = a.n - nm;
        d[a.C] = rs + (double)ndv * piv[r];                    // sum of the defined entries
        d[a.C + 1] = rss - rs * rs / (double)ndv;              // centred squares (the mean-imputed entries add 0)
        reinterpret_cast<int4*>(a.counts)[vv] = make_int4(0, 0, nm, 0);
      }
    }
  }
}

struct ImputeArgs {
  int64_t M;
  int C;
  int n;
  const double* basis_t;  // [ns_pad][C] sample-major copy
  const int32_t* counts;
  double* dots;
  const uint2* nanmask;
  int64_t n_chunks;
};

// dots[v][c] += mean_v * sum over the missing in-group samples j of basis[c][j]   (RU:52-57: those slots hold the mean)
__global__ void __launch_bounds__(256) dense_impute_kernel(ImputeArgs a) {
  constexpr int CP = 12;
  const int lane = threadIdx.x & 31;
  const int64_t v = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (v >= a.M) return;
  const int nm = a.counts[v * 4 + 2];
  if (nm == 0) return;
  double* d = a.dots + v * (a.C + 2);
  const double mean = d[a.C] / (double)(a.n - nm);   // 0 / 0 -> NaN for an all-missing variant, as RU:52
  const uint2* mk = a.nanmask + v * a.n_chunks;
  for (int c0 = 0; c0 < a.C; c0 += CP) {
    double corr[CP];
#pragma unroll
    for (int c = 0; c < CP; ++c) corr[c] = 0.0;
    const int cb = min(CP, a.C - c0);
    for (int64_t k = lane; k < a.n_chunks; k += 32) {
      const uint2 w = mk[k];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t bits = h ? w.y : w.x;
        while (bits) {
          const int b = __ffs(bits) - 1;
          bits &= bits - 1;
          const double* row = a.basis_t + (k * CHUNK + 2 * b + h) * a.C + c0;
#pragma unroll
          for (int c = 0; c < CP; ++c)
            if (c < cb) corr[c] += __ldg(row + c);
        }
      }
    }
#pragma unroll
    for (int c = 0; c < CP; ++c) {
      const double r = warp_sum(corr[c]);
      if (lane == 0 && c < cb) d[c0 + c] += mean * r;
    }
  }
}

__global__ void transpose_basis_kernel(const double* __restrict__ basis, int C, int64_t ns_pad, double* __restrict__ out) {
  const int64_t total = (int64_t)C * ns_pad;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t j = i / C;
    const int c = (int)(i - j * C);
    out[i] = basis[(int64_t)c * ns_pad + j];
  }
}

__global__ void first_sample_kernel(const uint32_t* __restrict__ mask, int64_t n_words, unsigned long long* out) {
  for (int64_t w = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; w < n_words; w += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t m = mask[w];
    if (!m) continue;
    for (int j = 0; j < 16; ++j)
      if ((m >> sample_shift(j)) & 1u) {
        atomicMin(out, (unsigned long long)(w * 16 + j));
        break;
      }
  }
}

template <int CB, bool FIRST, bool VEC>
void launch_pass_v(const DenseArgs& a, int grid, cudaStream_t st) {
  constexpr int smem = STAGES * (VT + CB) * CHUNK * (int)sizeof(double);
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(dense_sweep_kernel<CB, FIRST, VEC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    attr_set = true;
  }
  dense_sweep_kernel<CB, FIRST, VEC><<<grid, WARPS * 32, smem, st>>>(a);
}

template <int CB, bool FIRST>
void launch_pass(const DenseArgs& a, bool vec, int grid, cudaStream_t st) {
  if (vec) launch_pass_v<CB, FIRST, true>(a, grid, st);
  else launch_pass_v<CB, FIRST, false>(a, grid, st);
}

template <bool FIRST>
void launch_pass_cb(const DenseArgs& a, int cb, bool vec, int grid, cudaStream_t st) {
  if (cb <= 4) launch_pass<4, FIRST>(a, vec, grid, st);
  else if (cb <= 8) launch_pass<8, FIRST>(a, vec, grid, st);
  else launch_pass<12, FIRST>(a, vec, grid, st);
}

}  // namespace

int launch_dense_sweep(Ctx* c, const double* d_x, int64_t M, int64_t ldx, cudaStream_t st) {
  if (M == 0) return LRR_OK;
  const int64_t n_total = c->n_samples_total;
  for (size_t g = 0; g < c->groups.size(); ++g) {
    Group& G = c->groups[g];
    const int64_t n_chunks = (n_total + CHUNK - 1) / CHUNK;
    const size_t need = sizeof(uint2) * (size_t)M * (size_t)n_chunks;
    if (need > c->nanmask_bytes) {
      LRR_CUDA(c, cudaStreamSynchronize(st));
      cudaFree(c->d_nanmask);
      c->d_nanmask = nullptr;
      c->nanmask_bytes = 0;
      LRR_CUDA(c, cudaMalloc(&c->d_nanmask, need));
      c->nanmask_bytes = need;
    }
    if (!G.d_basis_t) {   // sample-major copy of the basis for the imputation pass + the pivot sample, once per group
      LRR_CUDA(c, cudaMalloc(&G.d_basis_t, sizeof(double) * (size_t)G.C * (size_t)G.ns_pad + sizeof(unsigned long long)));
      unsigned long long* d_first = reinterpret_cast<unsigned long long*>(G.d_basis_t + (size_t)G.C * (size_t)G.ns_pad);
      LRR_CUDA(c, cudaMemsetAsync(d_first, 0xff, sizeof(unsigned long long), st));
      transpose_basis_kernel<<<c->sm_count * 8, 256, 0, st>>>(G.d_basis, G.C, G.ns_pad, G.d_basis_t);
      first_sample_kernel<<<c->sm_count, 256, 0, st>>>(G.d_mask, G.ns_pad / 16, d_first);
      c->launches += 2;
      unsigned long long h_first = 0;
      LRR_CUDA(c, cudaMemcpyAsync(&h_first, d_first, sizeof(h_first), cudaMemcpyDeviceToHost, st));
      LRR_CUDA(c, cudaStreamSynchronize(st));
      G.first_sample = (h_first < (unsigned long long)n_total) ? (int64_t)h_first : 0;
    }
    DenseArgs a;
    a.x = d_x;
    a.M = M;
    a.ldx = ldx;
    a.n_total = n_total;
    a.ns_pad = G.ns_pad;
    a.basis = G.d_basis;
    a.mask = G.d_mask;
    a.first_sample = G.first_sample;
    a.n = G.n;
    a.C = G.C;
    a.counts = c->d_counts + (int64_t)g * c->reserved_variants * 4;
    a.dots = c->d_dots + c->dots_offset[g];
    a.nanmask = reinterpret_cast<uint2*>(c->d_nanmask);
    a.n_chunks = n_chunks;
    const bool vec = (ldx % 2 == 0) && (reinterpret_cast<uintptr_t>(d_x) % 16 == 0);
    const int grid = (int)((M + VT - 1) / VT);
    // column passes: one when C <= 12 (x is read once); otherwise passes of 12 re-read x
    int c0 = 0;
    do {
      a.c0 = c0;
      const int cb = G.C - c0;
      if (c0 == 0) launch_pass_cb<true>(a, cb, vec, grid, st);
      else launch_pass_cb<false>(a, cb, vec, grid, st);
      c->launches++;
      LRR_CUDA(c, cudaGetLastError());
      c0 += 12;
    } while (c0 < G.C);
    ImputeArgs ia;
    ia.M = M;
    ia.C = G.C;
    ia.n = G.n;
    ia.basis_t = G.d_basis_t;
    ia.counts = a.counts;
    ia.dots = a.dots;
    ia.nanmask = a.nanmask;
    ia.n_chunks = n_chunks;
    dense_impute_kernel<<<(int)((M + 7) / 8), 256, 0, st>>>(ia);
    c->launches++;
    LRR_CUDA(c, cudaGetLastError());
  }
  return LRR_OK;
}

}  // namespace lrr
