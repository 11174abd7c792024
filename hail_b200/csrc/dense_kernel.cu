// Kernel 4: float64 sweep over DENSE dosages -- x is any float64 entry expression (statgen.py:229, 391: e.g. the
// PL / GP dosages of test_statgen.py:286-364), one double per entry, NaN = missing.  Reference steps per variant:
//   RegressionUtils.setMeanImputedDoubles      hail/hail/src/is/hail/stats/RegressionUtils.scala:16-58
//   qtx = Qt * X, ytx = y^T X, x.x, sum(x)     hail/hail/src/is/hail/methods/LinearRegression.scala:136-146
//
// HBM-bound at 8 bytes per entry, so x is read exactly ONCE per group:
//   dense_sweep_kernel   a CTA owns VT = 8 variants, its 8 warps stride over 64-sample chunks (one 128-bit load per lane
//                        per variant, the basis columns [Q' | Y_res] read as 128-bit loads from L2 where they stay
//                        resident: 32 MB at 400k samples x 10 columns).  Per variant it accumulates, over the DEFINED
//                        entries of the group's samples: the count, the sum and the sum of squares, and the
//                        dot product with every basis column (missing entries contribute 0 here), and it writes one bit
//                        per (variant, sample) saying "missing and in the group" (1/64 of the traffic of x).
//   dense_impute_kernel  the mean-imputed column is x + mean * [missing]: a warp per variant walks the set bits and adds
//                        mean * sum_{missing} basis[c] to every dot product, reading a sample-major copy of the basis
//                        (one sample's C values are contiguous).  Cost is proportional to the number of missing entries.
// The statistics epilogue (stats_device.cuh, `dense`) then uses dots[C] = sum of the defined entries and
// dots[C + 1] = their sum of squares instead of the exact genotype counts of the packed kernels.
#include <type_traits>

#include "common.cuh"

namespace lrr {

namespace {

constexpr int VW = 4;            // variants per dot warp
constexpr int DOT_WARPS = 8;     // warps that accumulate the dot products
constexpr int STAT_WARPS = 8;    // warps that count / sum the defined entries, write the missing bits, zero the holes
constexpr int SVW = 4;           // variants per stat warp
constexpr int VT = VW * DOT_WARPS;  // variants per CTA (= SVW * STAT_WARPS)
constexpr int THREADS = (DOT_WARPS + STAT_WARPS) * 32;
constexpr int CHUNK = 64;        // samples per step (two per lane)
constexpr int STAGES = 8;        // cp.async ring depth (16 KB of x + 5 KB of basis per stage)

struct DenseArgs {
  const double* x;       // [M][ldx]  (float64 input)
  const uint16_t* xq;    // [M][ldx]  (compact input: entry = xq * xscale, 0xFFFF = missing), exclusive with x
  double xscale;
  int64_t M, ldx, n_total;
  int64_t ns_pad;
  const double* basis;   // [C][ns_pad]
  const uint32_t* mask;  // [ns_pad / 16], bit sample_shift(j & 15) of word j >> 4 set iff sample j is in the group
  const double* indicator;  // [ns_pad] 1.0 for the group's samples, else 0.0
  int n;
  int C;                 // all dot columns of the group
  int c0;                // first column of this pass
  int sq_col;            // weighted groups: the column accumulated against x^2 (sum_j w_j x_j^2, statgen.py:646), else -1
  int32_t* counts;       // [M][4]: (0, 0, n_missing, 0)
  double* dots;          // [M][C + 2]: C dot products, sum of the defined entries, their sum of squares
  uint2* nanmask;        // [M][n_chunks]: .x bit l = sample 64 k + 2 l missing (and in the group), .y = sample 64 k + 2 l + 1
  int64_t n_chunks;
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// 16- / 8-byte asynchronous global -> shared copies; `bytes` < size zero-fills the rest (samples past the last one)
__device__ __forceinline__ void cp_async16(uint32_t smem, const void* gmem, int bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem), "l"(gmem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async16_full(uint32_t smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async8(uint32_t smem, const void* gmem, int bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(smem), "l"(gmem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async8_full(uint32_t smem, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// exponent all ones: NaN (missing) or +-Inf.  Both are zeroed in shared memory so that the other rows of the CTA stay
// clean, but only NaN is a missing entry (RU:16-58): an infinite entry is a DEFINED value in the reference and turns
// every statistic of its row into NaN (sum_x into +-Inf); the row is marked in counts[v].w (bit 0: +Inf seen,
// bit 1: -Inf seen) and the statistics epilogue emits exactly that.
__device__ __forceinline__ bool not_finite(double v) { return (__double2hiint(v) & 0x7ff00000) == 0x7ff00000; }

// CB = basis columns of this pass (compile-time: the accumulators live in registers), FIRST = this pass also produces
// the counts / sums / missing bits, VEC = rows are 16-byte aligned (128-bit copies of x).
// Warp-specialised, one barrier per 64-sample step:
//   stat warps     (8 x 4 variants) issue every cp.async copy (stage k + 7) and run one stage AHEAD of the dot warps: on
//                  stage k + 1 they ballot the "missing and in the group" bits, count them, and overwrite every
//                  non-finite entry with 0 in shared memory
//   dot warps      (8 x 4 variants) then read clean data on stage k: acc[r][c] += basis[c][j] * x[r][j] and, with the
//                  group's 0 / 1 indicator column staged next to the basis, sum += ind[j] * x[r][j] and
//                  squares += (ind[j] * x[r][j]) * x[r][j]  (samples outside the group have zero basis rows)
// XK: how x arrives -- 0: float64, 8-byte copies; 1: float64, 16-byte copies (aligned rows); 2: compact uint16 entries
// (value = q * xscale, 0xFFFF = missing; 16-byte copies of 8 entries into a raw region of the stage, converted to float64
// by the stat warps one stage ahead of the dot warps, which never see the difference).
template <int CB, bool FIRST, int XK, bool SQ>
__global__ void __launch_bounds__(THREADS, 1) dense_sweep_kernel(DenseArgs a) {
  extern __shared__ __align__(16) double s_ring[];   // STAGES x { x [VT][CHUNK], indicator [CHUNK], basis [CB][CHUNK], raw u16 }
  constexpr bool VEC = XK == 1;
  constexpr bool U16 = XK == 2;
  constexpr int RAW_DOUBLES = U16 ? VT * CHUNK / 4 : 0;
  constexpr int STAGE_DOUBLES = (VT + 1 + CB) * CHUNK + RAW_DOUBLES;
  constexpr uint32_t STAGE_BYTES = STAGE_DOUBLES * 8;
  constexpr uint32_t RAW_OFF = (uint32_t)((VT + 1 + CB) * CHUNK) * 8u;   // byte offset of the raw region inside a stage
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int64_t v0 = (int64_t)blockIdx.x * VT;
  const int cb = min(CB, a.C - a.c0);
  const int sw = warp - DOT_WARPS;

  if (cb < CB) {   // unused columns of the last pass: zero once, never copied into
    for (int st = 0; st < STAGES; ++st)
      for (int i = threadIdx.x; i < (CB - cb) * CHUNK; i += THREADS)
        s_ring[st * STAGE_DOUBLES + (VT + 1 + cb) * CHUNK + i] = 0.0;
  }

  if (sw < 0) {
    // =================== dot warps ===================
    double acc[VW][CB];
    double s[VW], ss[VW];
#pragma unroll
    for (int r = 0; r < VW; ++r) {
      s[r] = ss[r] = 0.0;
#pragma unroll
      for (int c = 0; c < CB; ++c) acc[r][c] = 0.0;
    }
    const int sqc = SQ ? a.sq_col - a.c0 : -1;
    __syncthreads();   // (the stat warps clean stage 0 now)
    int slot = 0;      // ring slot of stage `chunk`
    for (int64_t chunk = 0; chunk < a.n_chunks; ++chunk) {
      __syncthreads();   // stage chunk is complete and clean
      const double* st = s_ring + slot * STAGE_DOUBLES + lane * 2;
      double2 xv[VW];
#pragma unroll
      for (int r = 0; r < VW; ++r) xv[r] = *reinterpret_cast<const double2*>(st + (warp * VW + r) * CHUNK);
      if (FIRST) {
        const double2 g = *reinterpret_cast<const double2*>(st + VT * CHUNK);
#pragma unroll
        for (int r = 0; r < VW; ++r) {
          const double t0 = g.x * xv[r].x, t1 = g.y * xv[r].y;
          s[r] += t0 + t1;
          ss[r] = fma(t0, xv[r].x, fma(t1, xv[r].y, ss[r]));
        }
      }
#pragma unroll
      for (int c = 0; c < CB; ++c) {
        const double2 q = *reinterpret_cast<const double2*>(st + (VT + 1 + c) * CHUNK);
        if (SQ && c == sqc) {   // the w column of a weighted group meets x^2
#pragma unroll
          for (int r = 0; r < VW; ++r) acc[r][c] = fma(q.x, xv[r].x * xv[r].x, fma(q.y, xv[r].y * xv[r].y, acc[r][c]));
        } else {
#pragma unroll
          for (int r = 0; r < VW; ++r) acc[r][c] = fma(q.x, xv[r].x, fma(q.y, xv[r].y, acc[r][c]));
        }
      }
      slot = slot == STAGES - 1 ? 0 : slot + 1;
    }
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
        if (lane == 0 && vv < a.M) {
          d[a.C] = rs;        // sum of the defined entries of the group's samples
          d[a.C + 1] = rss;   // their sum of squares
        }
      }
    }
  } else {
    // =================== stat warps: loader + missing bits + hole filling ===================
    const int t = threadIdx.x - DOT_WARPS * 32;
    constexpr int LT = STAT_WARPS * 32;   // loader threads
    const uint32_t ring_base = (uint32_t)__cvta_generic_to_shared(s_ring);
    constexpr int X_ELEMS = U16 ? VT * CHUNK / 8 : VEC ? VT * CHUNK / 2 : VT * CHUNK;         // x copies per stage
    static_assert(X_ELEMS % LT == 0, "every loader thread issues the same number of x copies");
    constexpr int XI = X_ELEMS / LT;
    constexpr int XBYTES = U16 ? 2 : 8;   // bytes per entry in global memory
    const char* xsrc[XI];
    uint32_t xdst[XI];   // byte offset inside a stage
    int xl[XI];          // first entry of the copy within its chunk
#pragma unroll
    for (int i = 0; i < XI; ++i) {
      const int idx = t + i * LT;
      const int r = U16 ? idx >> 3 : VEC ? idx >> 5 : idx >> 6;
      const int l = U16 ? (idx & 7) * 8 : VEC ? (idx & 31) * 2 : idx & 63;
      int64_t vv = v0 + r;
      if (vv >= a.M) vv = a.M - 1;   // clamp: copies stay in bounds, stores are skipped
      xsrc[i] = (U16 ? reinterpret_cast<const char*>(a.xq) : reinterpret_cast<const char*>(a.x)) + (vv * a.ldx + l) * XBYTES;
      xdst[i] = U16 ? RAW_OFF + (uint32_t)(r * CHUNK + l) * 2u : (uint32_t)(r * CHUNK + l) * 8u;
      xl[i] = l;
    }
    const char* const xbase = U16 ? reinterpret_cast<const char*>(a.xq) : reinterpret_cast<const char*>(a.x);
    // column copies: slot 0 = the group indicator, slot 1 + c = basis column c0 + c (all 128-bit, ns_pad is padded)
    constexpr int BI = ((1 + CB) * (CHUNK / 2) + LT - 1) / LT;
    const double* bsrc[BI];
    uint32_t bdst[BI];
    bool bok[BI];
#pragma unroll
    for (int i = 0; i < BI; ++i) {
      const int idx = t + i * LT;
      const int col = idx >> 5, l2 = (idx & 31) * 2;
      bok[i] = col < 1 + cb;
      bsrc[i] = (col == 0 || !bok[i] ? a.indicator : a.basis + (int64_t)(a.c0 + col - 1) * a.ns_pad) + l2;
      bdst[i] = (uint32_t)((VT + col) * CHUNK + l2) * 8u;
    }
    const int64_t n_full = a.n_total / CHUNK;
    // copies of stage `chunk` into ring slot `slot`; the source pointers advance with every call (calls are in order)
    auto issue = [&](int64_t chunk, int slot) {
      const uint32_t st = ring_base + (uint32_t)slot * STAGE_BYTES;
      if (chunk < n_full) {
#pragma unroll
        for (int i = 0; i < XI; ++i) {
          if (VEC || U16) cp_async16_full(st + xdst[i], xsrc[i]);
          else cp_async8_full(st + xdst[i], xsrc[i]);
        }
      } else {   // the last, partial chunk: zero-fill past the last sample
#pragma unroll
        for (int i = 0; i < XI; ++i) {
          const int64_t left = a.n_total - (chunk * CHUNK + (int64_t)xl[i]);
          if (U16) {
            const int bytes = left >= 8 ? 16 : (left > 0 ? (int)left * 2 : 0);
            cp_async16(st + xdst[i], bytes ? xsrc[i] : xbase, bytes);
          } else if (VEC) {
            const int bytes = left >= 2 ? 16 : (left == 1 ? 8 : 0);
            cp_async16(st + xdst[i], bytes ? xsrc[i] : xbase, bytes);
          } else {
            const int bytes = left >= 1 ? 8 : 0;
            cp_async8(st + xdst[i], bytes ? xsrc[i] : xbase, bytes);
          }
        }
      }
#pragma unroll
      for (int i = 0; i < BI; ++i)
        if (bok[i]) cp_async16_full(st + bdst[i], bsrc[i]);
#pragma unroll
      for (int i = 0; i < XI; ++i) xsrc[i] += CHUNK * XBYTES;
#pragma unroll
      for (int i = 0; i < BI; ++i) bsrc[i] += CHUNK;
    };
#pragma unroll
    for (int k = 0; k < STAGES - 1; ++k) {
      if (k < a.n_chunks) issue(k, k);
      cp_async_commit();
    }

    int nm[SVW], inf_seen[SVW];
#pragma unroll
    for (int r = 0; r < SVW; ++r) nm[r] = inf_seen[r] = 0;
    // one step on the stage in `slot`: missing bits of chunk `chunk`, then zero the holes in place
    auto stat_step = [&](int64_t chunk, int slot) {
      double* st = s_ring + slot * STAGE_DOUBLES + (sw * SVW) * CHUNK + lane * 2;
      bool g0 = false, g1 = false;
      if (FIRST) {   // the staged indicator column (no global load on the per-step critical path)
        const double2 g = *reinterpret_cast<const double2*>(s_ring + slot * STAGE_DOUBLES + VT * CHUNK + lane * 2);
        g0 = g.x != 0.0;
        g1 = g.y != 0.0;
      }
#pragma unroll
      for (int r = 0; r < SVW; ++r) {
        if (U16) {
          // compact entries: convert this lane's two entries to float64 (missing -> 0 + its bit), always written
          const uint32_t q2 = *reinterpret_cast<const uint32_t*>(
              reinterpret_cast<const char*>(s_ring + slot * STAGE_DOUBLES) + RAW_OFF + ((sw * SVW + r) * CHUNK + lane * 2) * 2);
          const uint32_t q0 = q2 & 0xFFFFu, q1 = q2 >> 16;
          const bool m0 = q0 == 0xFFFFu, m1 = q1 == 0xFFFFu;
          *reinterpret_cast<double2*>(st + r * CHUNK) = make_double2(m0 ? 0.0 : (double)q0 * a.xscale, m1 ? 0.0 : (double)q1 * a.xscale);
          if (FIRST) {
            const uint32_t b0 = __ballot_sync(0xffffffffu, g0 && m0), b1 = __ballot_sync(0xffffffffu, g1 && m1);
            nm[r] += __popc(b0) + __popc(b1);
            const int64_t vv = v0 + sw * SVW + r;
            if (lane == 0 && vv < a.M) a.nanmask[vv * a.n_chunks + chunk] = make_uint2(b0, b1);
          }
          continue;
        }
        double2 xv = *reinterpret_cast<const double2*>(st + r * CHUNK);
        bool m0 = not_finite(xv.x), m1 = not_finite(xv.y);
        if (m0 || m1) {
          if (FIRST) {   // +-Inf inside the group: a defined value, not a missing one
            if (m0 && isinf(xv.x)) { if (g0) inf_seen[r] |= xv.x > 0.0 ? 1 : 2; m0 = false; }
            if (m1 && isinf(xv.y)) { if (g1) inf_seen[r] |= xv.y > 0.0 ? 1 : 2; m1 = false; }
          }
          if (not_finite(xv.x)) xv.x = 0.0;
          if (not_finite(xv.y)) xv.y = 0.0;
          *reinterpret_cast<double2*>(st + r * CHUNK) = xv;
        }
        if (FIRST) {
          const uint32_t b0 = __ballot_sync(0xffffffffu, g0 && m0), b1 = __ballot_sync(0xffffffffu, g1 && m1);
          nm[r] += __popc(b0) + __popc(b1);
          const int64_t vv = v0 + sw * SVW + r;
          if (lane == 0 && vv < a.M) a.nanmask[vv * a.n_chunks + chunk] = make_uint2(b0, b1);
        }
      }
    };
    cp_async_wait<STAGES - 2>();
    __syncthreads();   // stage 0 has landed for every loader thread
    if (a.n_chunks > 0) stat_step(0, 0);
    int slot = 0;
    for (int64_t chunk = 0; chunk < a.n_chunks; ++chunk) {
      cp_async_wait<STAGES - 3>();
      __syncthreads();   // stage chunk + 1 has landed, stage chunk is clean, stage chunk - 1 is free again
      if (chunk + STAGES - 1 < a.n_chunks) issue(chunk + STAGES - 1, slot == 0 ? STAGES - 1 : slot - 1);
      cp_async_commit();
      slot = slot == STAGES - 1 ? 0 : slot + 1;
      if (chunk + 1 < a.n_chunks) stat_step(chunk + 1, slot);
    }
    if (FIRST) {
#pragma unroll
      for (int r = 0; r < SVW; ++r) {
        const int64_t vv = v0 + sw * SVW + r;
        const int inf_bits = (int)__reduce_or_sync(0xffffffffu, (unsigned)inf_seen[r]);
        if (lane == 0 && vv < a.M) reinterpret_cast<int4*>(a.counts)[vv] = make_int4(0, 0, nm[r], inf_bits);
      }
    }
  }
}

struct ImputeArgs {
  int64_t M;
  int C;
  int n;
  int sq_col;             // column whose imputed entries contribute mean^2 (weighted groups), else -1
  const double* basis_t;  // [ns_pad][C] sample-major copy
  const int32_t* counts;
  double* dots;
  const uint2* nanmask;
  int64_t n_chunks;
};

// dots[v][c] += mean_v * sum over the missing in-group samples j of basis[c][j]   (RU:52-57: those slots hold the mean)
// One CTA per variant with missing entries; threads stride over the 64-sample mask words.
constexpr int IMPUTE_THREADS = 128;
__global__ void __launch_bounds__(IMPUTE_THREADS) dense_impute_kernel(ImputeArgs a) {
  constexpr int CP = 12;
  __shared__ double s_corr[IMPUTE_THREADS / 32][CP];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t v = blockIdx.x;
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
    for (int64_t k = threadIdx.x; k < a.n_chunks; k += IMPUTE_THREADS) {
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
      if (lane == 0) s_corr[warp][c] = r;
    }
    __syncthreads();
    if ((int)threadIdx.x < cb) {
      double r = 0.0;
#pragma unroll
      for (int w = 0; w < IMPUTE_THREADS / 32; ++w) r += s_corr[w][threadIdx.x];
      d[c0 + threadIdx.x] += ((c0 + (int)threadIdx.x == a.sq_col) ? mean * mean : mean) * r;
    }
    __syncthreads();
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

// 0 / 1 indicator of the group's samples as float64 (a staged "basis column" for the sums of the defined entries)
__global__ void indicator_kernel(const uint32_t* __restrict__ mask, int64_t ns_pad, double* __restrict__ out) {
  for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < ns_pad; j += (int64_t)gridDim.x * blockDim.x)
    out[j] = ((mask[j >> 4] >> sample_shift((int)(j & 15))) & 1u) ? 1.0 : 0.0;
}

template <int CB, bool FIRST, int XK, bool SQ>
cudaError_t launch_pass_v(const DenseArgs& a, int grid, cudaStream_t st) {
  constexpr int smem = STAGES * ((VT + 1 + CB) * CHUNK + (XK == 2 ? VT * CHUNK / 4 : 0)) * (int)sizeof(double);
  // the attribute is per device: set it on every launch (a host-side table write) instead of caching one flag per process
  const cudaError_t e = cudaFuncSetAttribute(dense_sweep_kernel<CB, FIRST, XK, SQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  dense_sweep_kernel<CB, FIRST, XK, SQ><<<grid, THREADS, smem, st>>>(a);
  return cudaSuccess;
}

template <int CB, bool FIRST, bool SQ = false>
cudaError_t launch_pass(const DenseArgs& a, int xk, int grid, cudaStream_t st) {
  if (xk == 2) {
    if constexpr (SQ) return cudaErrorNotSupported;   // (weighted groups take float64 x)
    else return launch_pass_v<CB, FIRST, 2, false>(a, grid, st);
  }
  return xk == 1 ? launch_pass_v<CB, FIRST, 1, SQ>(a, grid, st) : launch_pass_v<CB, FIRST, 0, SQ>(a, grid, st);
}

template <bool FIRST>
cudaError_t launch_pass_cb(const DenseArgs& a, int cb, int xk, int grid, cudaStream_t st) {
  if (a.sq_col >= 0) {   // weighted groups: fewer column-tile sizes (each is one more kernel to build)
    if (cb <= 4) return launch_pass<4, FIRST, true>(a, xk, grid, st);
    if (cb <= 8) return launch_pass<8, FIRST, true>(a, xk, grid, st);
    return launch_pass<12, FIRST, true>(a, xk, grid, st);
  }
  if (cb <= 2) return launch_pass<2, FIRST>(a, xk, grid, st);
  if (cb <= 4) return launch_pass<4, FIRST>(a, xk, grid, st);
  if (cb <= 6) return launch_pass<6, FIRST>(a, xk, grid, st);
  if (cb <= 8) return launch_pass<8, FIRST>(a, xk, grid, st);
  if (cb <= 10) return launch_pass<10, FIRST>(a, xk, grid, st);
  return launch_pass<12, FIRST>(a, xk, grid, st);
}

}  // namespace

int launch_dense_sweep(Ctx* c, const double* d_x, int64_t M, int64_t ldx, cudaStream_t st, const uint16_t* d_xq, double xscale) {
  if (M == 0) return LRR_OK;
  if (d_xq) {
    if (ldx % 8 != 0 || reinterpret_cast<uintptr_t>(d_xq) % 16 != 0)
      return fail(c, LRR_EINVAL, "compact dosage rows must be 16-byte aligned (ldx a multiple of 8 entries)");
    for (const Group& G : c->groups)
      if (G.weighted) return fail(c, LRR_EINVAL, "weighted groups take float64 x (lrr_run_dense)");
  }
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
    if (!G.d_basis_t) {   // once per group: sample-major copy of the basis (imputation pass) + the indicator column
      LRR_CUDA(c, cudaMalloc(&G.d_basis_t, sizeof(double) * ((size_t)G.C + 1) * (size_t)G.ns_pad));
      transpose_basis_kernel<<<c->sm_count * 8, 256, 0, st>>>(G.d_basis, G.C, G.ns_pad, G.d_basis_t);
      indicator_kernel<<<c->sm_count * 2, 256, 0, st>>>(G.d_mask, G.ns_pad, G.d_basis_t + (size_t)G.C * (size_t)G.ns_pad);
      c->launches += 2;
      LRR_CUDA(c, cudaGetLastError());
    }
    DenseArgs a;
    a.x = d_x;
    a.xq = d_xq;
    a.xscale = xscale;
    a.M = M;
    a.ldx = ldx;
    a.n_total = n_total;
    a.ns_pad = G.ns_pad;
    a.basis = G.d_basis;
    a.mask = G.d_mask;
    a.indicator = G.d_basis_t + (size_t)G.C * (size_t)G.ns_pad;
    a.n = G.n;
    a.C = G.C;
    a.sq_col = G.weighted ? G.C - 1 : -1;
    a.counts = c->d_counts + (int64_t)g * c->reserved_variants * 4;
    a.dots = c->d_dots + c->dots_offset[g];
    a.nanmask = reinterpret_cast<uint2*>(c->d_nanmask);
    a.n_chunks = n_chunks;
    const int vec = d_xq ? 2 : ((ldx % 2 == 0) && (reinterpret_cast<uintptr_t>(d_x) % 16 == 0)) ? 1 : 0;
    const int grid = (int)((M + VT - 1) / VT);
    // column passes: one when C <= 12 (x is read once); otherwise passes of 12 re-read x
    int c0 = 0;
    do {
      a.c0 = c0;
      const int cb = G.C - c0;
      LRR_CUDA(c, c0 == 0 ? launch_pass_cb<true>(a, cb, vec, grid, st) : launch_pass_cb<false>(a, cb, vec, grid, st));
      c->launches++;
      LRR_CUDA(c, cudaGetLastError());
      c0 += 12;
    } while (c0 < G.C);
    ImputeArgs ia;
    ia.M = M;
    ia.C = G.C;
    ia.n = G.n;
    ia.sq_col = G.weighted ? G.C - 1 : -1;
    ia.basis_t = G.d_basis_t;
    ia.counts = a.counts;
    ia.dots = a.dots;
    ia.nanmask = a.nanmask;
    ia.n_chunks = n_chunks;
    dense_impute_kernel<<<(unsigned)M, IMPUTE_THREADS, 0, st>>>(ia);
    c->launches++;
    LRR_CUDA(c, cudaGetLastError());
  }
  return LRR_OK;
}

}  // namespace lrr
