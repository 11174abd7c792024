// Probe: tcgen05.mma.kind::mxf4.block_scale (E2M1 x E2M1 -> f32, K = 64) with the A operand in tensor memory.
// Questions: (1) element / nibble layout of A in TMEM and B in 128B-swizzled shared memory, (2) is the f32
// accumulation EXACT for small-integer data over ~400k-sample sums, (3) cycles per MMA for N = 96..128.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scratch/fp4_probe scratch/fp4_probe.cu
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// block-scaled instruction descriptor (cute/arch/mma_sm100_desc.hpp InstrDescriptorBlockScaled)
__host__ __device__ inline uint32_t make_idesc_mxf4(int n, int m) {
  uint32_t d = 0;
  d |= 1u << 7;                   // a_format = E2M1 (MXF4Format)
  d |= 1u << 10;                  // b_format = E2M1
  d |= (uint32_t)(n >> 3) << 17;  // N
  d |= 1u << 23;                  // scale format UE8M0
  d |= (uint32_t)(m >> 4) << 24;  // M
  return d;                       // a_sf_id = b_sf_id = 0, K = 64 dense
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mma_mxf4_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t sfa, uint32_t sfb,
                                            uint32_t acc) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::mxf4.block_scale.scale_vec::2X [%0], [%1], %2, %3, [%5], [%6], p;\n}\n" ::"r"(d),
      "r"(a), "l"(b), "r"(idesc), "r"(acc), "r"(sfa), "r"(sfb)
      : "memory");
}
__device__ __forceinline__ void commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
      "[%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

__host__ __device__ inline uint32_t hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}
__host__ __device__ inline uint32_t a_code(int r, int k) { return hash32(0x1000u + r * 977u + k * 131u) & 3u; }         // 0, .5, 1, 1.5
__host__ __device__ inline uint32_t b_code(int n, int k) { return hash32(0x9000u + n * 613u + k * 29u) & 15u; }        // any E2M1
static double e2m1(uint32_t c) {
  static const double v[8] = {0, 0.5, 1, 1.5, 2, 3, 4, 6};
  return (c & 8) ? -v[c & 7] : v[c & 7];
}

constexpr int KSTEPS = 4;   // 4 x 64 = 256 samples per smem row / per 32 TMEM columns

// 128 threads; warp w owns TMEM lanes 32w..32w+31
__global__ void __launch_bounds__(128, 1) probe(int N, int reps, int timing_groups, float* out /*[128][N]*/,
                                                unsigned long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t s0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sB = smem_raw + (s0 - smem_u32(smem_raw));   // N rows x 128 B, 128B swizzle
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = threadIdx.x;
  // B: element k of row n is nibble (k & 1) of byte k / 2; 16-byte chunk c of row n is stored at chunk c ^ (n & 7)
  for (int i = threadIdx.x; i < N * 128; i += blockDim.x) {
    const int n = i / 128, byte = i % 128;
    const uint32_t lo = b_code(n, 2 * byte), hi = b_code(n, 2 * byte + 1);
    const int chunk = byte >> 4, within = byte & 15;
    sB[n * 128 + (((chunk ^ (n & 7)) << 4) | within)] = (uint8_t)(lo | (hi << 4));
  }
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base;
  const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
  const uint32_t A_COL = 256, SFA_COL = 320, SFB_COL = 352;
  {
    // A: lane = row; 32 registers = 128 bytes = 256 nibbles; element k in nibble (k & 1) of byte k / 2
    uint32_t r[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      uint32_t w = 0;
#pragma unroll
      for (int e = 0; e < 8; ++e) w |= a_code(row, 8 * i + e) << (4 * e);
      r[i] = w;
    }
    tmem_st32(tmem + lane_addr + A_COL, r);
    // scale factors: every byte = UE8M0 1.0, so that the result does not depend on the SF layout
#pragma unroll
    for (int i = 0; i < 32; ++i) r[i] = 0x7F7F7F7Fu;
    tmem_st32(tmem + lane_addr + SFA_COL, r);   // covers SFA_COL .. SFB_COL + 31 partially: 32 columns from 320
    tmem_st32(tmem + lane_addr + SFB_COL, r);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  const uint32_t idesc = make_idesc_mxf4(N, 128);
  const uint64_t bd = make_desc(s0);
  uint32_t phase = 0;
  if (warp == 0) {
    if (elect_one()) {
      for (int rep = 0; rep < reps; ++rep)
#pragma unroll
        for (int j = 0; j < KSTEPS; ++j)
          mma_mxf4_ts(tmem, tmem + A_COL + j * 8, bd + (uint64_t)(j * 2), idesc, tmem + SFA_COL, tmem + SFB_COL, (rep | j) ? 1u : 0u);
      commit(smem_u32(&bar));
    }
    __syncwarp();
  }
  mbar_wait(smem_u32(&bar), phase);
  phase ^= 1;
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (out) {
    for (int c0 = 0; c0 < N; c0 += 16) {
      uint32_t r[16];
      tmem_ld16(tmem + lane_addr + c0, r);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if (c0 + i < N) out[row * N + c0 + i] = __uint_as_float(r[i]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // timing: groups of 16 MMAs, two groups in flight
  if (timing_groups > 0 && warp == 0) {
    long long t0 = clock64();
    for (int g = 0; g < timing_groups; ++g) {
      if (elect_one()) {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          mma_mxf4_ts(tmem, tmem + A_COL + (j & 3) * 8, bd + (uint64_t)((j & 3) * 2), idesc, tmem + SFA_COL, tmem + SFB_COL, 1u);
        commit(smem_u32(&bar));
      }
      __syncwarp();
      mbar_wait(smem_u32(&bar), phase);   // serialised groups: an upper bound of the per-MMA cost
      phase ^= 1;
    }
    long long t1 = clock64();
    if (lane == 0) cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

int main() {
  const size_t smem = 256 * 128 + 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int N : {16, 96, 112, 128}) {
    for (int reps : {1, 1563}) {
      float* d_out;
      cudaMalloc(&d_out, sizeof(float) * 128 * N);
      cudaMemset(d_out, 0xFF, sizeof(float) * 128 * N);
      probe<<<1, 128, smem>>>(N, reps, 0, d_out, nullptr);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("N=%d reps=%d: CUDA error %s\n", N, reps, cudaGetErrorString(e)); return 1; }
      std::vector<float> h(128 * N);
      cudaMemcpy(h.data(), d_out, sizeof(float) * h.size(), cudaMemcpyDeviceToHost);
      cudaFree(d_out);
      long bad = 0; double maxerr = 0; int shown = 0;
      for (int r = 0; r < 128; ++r)
        for (int n = 0; n < N; ++n) {
          double base = 0;
          for (int k = 0; k < 64 * KSTEPS; ++k) base += e2m1(a_code(r, k)) * e2m1(b_code(n, k));
          const double want = base * reps;
          const double got = h[r * N + n];
          if (got != (double)(float)want) {
            ++bad;
            maxerr = fmax(maxerr, fabs(got - want));
            if (shown < 4) { printf("   mismatch r=%d n=%d got %.4f want %.4f\n", r, n, got, want); ++shown; }
          }
        }
      printf("N=%3d reps=%4d (%6d samples): %ld / %d mismatches, max abs err %.3f\n", N, reps, reps * 64 * KSTEPS, bad, 128 * N, maxerr);
    }
  }
  // timing on all SMs
  for (int N : {48, 96, 112, 128}) {
    unsigned long long* d_c;
    cudaMalloc(&d_c, sizeof(unsigned long long) * 148);
    const int groups = 2000;
    probe<<<148, 128, smem>>>(N, 1, groups, nullptr, d_c);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("timing N=%d: CUDA error %s\n", N, cudaGetErrorString(e)); return 1; }
    unsigned long long h[148];
    cudaMemcpy(h, d_c, sizeof h, cudaMemcpyDeviceToHost);
    unsigned long long mx = 0;
    for (auto v : h) mx = v > mx ? v : mx;
    printf("mxf4 TS N=%3d: %.1f cycles per MMA (K=64), serialised groups of 16\n", N, (double)mx / (groups * 16.0));
    cudaFree(d_c);
  }
  return 0;
}
