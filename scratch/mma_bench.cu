// Microbenchmark: back-to-back tcgen05.mma.kind::i8 (M=128 per CTA, K=32) for small N.
//   mode 0: A from TMEM (TS), cta_group::1        mode 1: A from shared memory (SS), cta_group::1
//   mode 2: A from TMEM (TS), cta_group::2 (M=256 over a CTA pair, B split across the pair)
// Reports cycles per MMA (per issuing CTA) so that the sweep kernel's per-MMA cost can be attributed.
// Optional: `sttm` = 1 runs 16 extra warps that stream tcgen05.st into a separate TMEM region (contention probe).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scratch/mma_bench scratch/mma_bench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

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
__device__ __forceinline__ uint32_t make_idesc(int n, int m) {
  uint32_t d = 0;
  d |= 2u << 4;
  d |= 0u << 7;
  d |= 1u << 10;
  d |= (uint32_t)(n >> 3) << 17;
  d |= (uint32_t)(m >> 4) << 24;
  return d;
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
  return pred != 0;
}

template <int CG>
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (CG == 1)
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
  else
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::i8 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
template <int CG>
__device__ __forceinline__ void commit(uint32_t bar) {
  if (CG == 1)
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
  else
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
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

// warps: 0 = MMA issuer, 1..16 = optional STTM streamers
template <int CG, int PG, int W, int mode>
__global__ void __launch_bounds__(17 * 32, 1) bench(int N, int groups, int sttm, int sttm_gap, int Mcta,
                                                      unsigned long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t s0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_smem = s0;                   // 128 rows x 128 B
  const uint32_t b_smem = s0 + 16384;           // 256 rows x 128 B
  __shared__ uint64_t bars[10];
  __shared__ uint32_t tmem_base;
  __shared__ volatile int stop;
  const int warp = threadIdx.x >> 5;
  // deterministic smem fill (values irrelevant for timing)
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(smem_raw + (s0 - smem_u32(smem_raw)))[i] = 0x01010101u * (i & 3);
  if (threadIdx.x == 0) {
    for (int i = 0; i < 10; ++i) mbar_init(smem_u32(&bars[i]), 1);
    stop = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    if (CG == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base)) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base)) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base;
  const bool leader = CG == 1 || cluster_ctarank() == 0;

  if (warp == 0) {
    const uint32_t idesc = make_idesc(N, CG == 2 ? 2 * Mcta : Mcta);
    const uint64_t bd = make_desc(b_smem), ad = make_desc(a_smem);
    const uint32_t a_t = tmem + 256;
    long long t0 = 0, t1 = 0;
    if (leader) {
      // warm-up group
      if (elect_one()) {
        for (int j = 0; j < 16; ++j) {
          if (mode == 1) mma_ss(tmem, ad + (uint64_t)((j & 3) * 2), bd + (uint64_t)((j & 3) * 2), idesc, j ? 1u : 0u);
          else mma_ts<CG>(tmem, a_t + (j & 3) * 8, bd + (uint64_t)((j & 3) * 2), idesc, j ? 1u : 0u);
        }
        commit<CG>(smem_u32(&bars[9]));
      }
      __syncwarp();
    }
    mbar_wait(smem_u32(&bars[9]), 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    t0 = clock64();
    uint32_t ph = 1;
    // groups is a multiple of W: the loop body is unrolled over the W barriers so every address is a constant
    for (int g0 = 0; g0 < groups; g0 += W) {
      const uint32_t par = (uint32_t)((g0 / W) & 1);
#pragma unroll
      for (int u = 0; u < W; ++u) {
        if (leader) {
          if (elect_one()) {
#pragma unroll
            for (int j = 0; j < PG; ++j) {
              if (mode == 1) mma_ss(tmem, ad + (uint64_t)((j & 3) * 2), bd + (uint64_t)((j & 3) * 2), idesc, 1u);
              else mma_ts<CG>(tmem, a_t + (j & 3) * 8 + ((j >> 2) & 3) * 32, bd + (uint64_t)((j & 3) * 2), idesc, 1u);
            }
            commit<CG>(smem_u32(&bars[u]));
          }
          __syncwarp();
        }
        // wait for the group issued W-1 groups ago
        if (u == W - 1) mbar_wait(smem_u32(&bars[0]), par);
        else if (g0 > 0) mbar_wait(smem_u32(&bars[u + 1]), par ^ 1u);
      }
    }
#pragma unroll
    for (int u = 1; u < W; ++u) mbar_wait(smem_u32(&bars[u]), (uint32_t)(((groups / W) - 1) & 1));
    (void)ph;
    t1 = clock64();
    if (threadIdx.x == 0 && leader) out[blockIdx.x] = (unsigned long long)(t1 - t0);
    stop = 1;
  } else if (sttm) {
    // STTM streamers: warp w writes lanes (w&3)*32.., columns 384 + ((w-1)>>2)*32
    const int w = warp - 1;
    const uint32_t addr = tmem + (((uint32_t)(w & 3) * 32) << 16) + 384 + (w >> 2) * 32;
    uint32_t r[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) r[i] = threadIdx.x + i;
    while (!stop) {
      tmem_st32(addr, r);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      for (int k = 0; k < sttm_gap; ++k) asm volatile("nanosleep.u32 20;");
#pragma unroll
      for (int i = 0; i < 32; ++i) r[i] += 1;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  if (warp == 0) {
    if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

template <int CG, int PG, int W, int MODE>
static double run_t(int N, int groups, int sttm, int gap, int ctas, int Mcta) {
  unsigned long long* d_out;
  cudaMalloc(&d_out, sizeof(unsigned long long) * ctas);
  cudaMemset(d_out, 0, sizeof(unsigned long long) * ctas);
  const size_t smem = 16384 + 32768 + 1024;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(ctas);
  cfg.blockDim = dim3(17 * 32);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  groups = groups / W * W;
  cudaFuncSetAttribute(bench<CG, PG, W, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaError_t e = cudaLaunchKernelEx(&cfg, bench<CG, PG, W, MODE>, N, groups, sttm, gap, Mcta, d_out);
  if (e != cudaSuccess) { printf("launch error %s\n", cudaGetErrorString(e)); exit(1); }
  e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("run error %s (cg=%d N=%d mode=%d)\n", cudaGetErrorString(e), CG, N, MODE); exit(1); }
  unsigned long long* h = (unsigned long long*)malloc(sizeof(unsigned long long) * ctas);
  cudaMemcpy(h, d_out, sizeof(unsigned long long) * ctas, cudaMemcpyDeviceToHost);
  double mx = 0;
  for (int i = 0; i < ctas; ++i) if ((double)h[i] > mx) mx = (double)h[i];
  free(h);
  cudaFree(d_out);
  return mx / ((double)groups * PG);
}

int main() {
  const int groups = 2400;
  const int Ns[] = {16, 32, 48, 64, 96, 128, 192, 256};
  printf("cycles per MMA (max over 148 CTAs), K=32 i8, fully unrolled issue\n");
  printf("\nN=48 TS cg1: per-group MMAs (PG) x in-flight window (W groups)\n");
  printf("%6s %8s %8s %8s %8s\n", "PG", "W=2", "W=3", "W=4", "W=8");
  printf("%6d %8.1f %8.1f %8.1f %8.1f\n", 4, run_t<1, 4, 2, 0>(48, groups, 0, 0, 148, 128), run_t<1, 4, 3, 0>(48, groups, 0, 0, 148, 128), run_t<1, 4, 4, 0>(48, groups, 0, 0, 148, 128), run_t<1, 4, 8, 0>(48, groups, 0, 0, 148, 128));
  printf("%6d %8.1f %8.1f %8.1f %8.1f\n", 8, run_t<1, 8, 2, 0>(48, groups, 0, 0, 148, 128), run_t<1, 8, 3, 0>(48, groups, 0, 0, 148, 128), run_t<1, 8, 4, 0>(48, groups, 0, 0, 148, 128), run_t<1, 8, 8, 0>(48, groups, 0, 0, 148, 128));
  printf("%6d %8.1f %8.1f %8.1f %8.1f\n", 16, run_t<1, 16, 2, 0>(48, groups, 0, 0, 148, 128), run_t<1, 16, 3, 0>(48, groups, 0, 0, 148, 128), run_t<1, 16, 4, 0>(48, groups, 0, 0, 148, 128), run_t<1, 16, 8, 0>(48, groups, 0, 0, 148, 128));
  printf("%6d %8.1f %8.1f %8.1f %8.1f\n", 32, run_t<1, 32, 2, 0>(48, groups, 0, 0, 148, 128), run_t<1, 32, 3, 0>(48, groups, 0, 0, 148, 128), run_t<1, 32, 4, 0>(48, groups, 0, 0, 148, 128), run_t<1, 32, 8, 0>(48, groups, 0, 0, 148, 128));
  printf("\nPG=16 W=4: cycles per MMA vs N\n");
  printf("%6s %10s %10s %10s %10s %10s\n", "N", "TS M128", "TS M64", "SS M128", "SS M64", "TS2 M256");
  for (int N : Ns) {
    printf("%6d %10.1f %10.1f %10.1f %10.1f %10.1f\n", N, run_t<1, 16, 4, 0>(N, groups, 0, 0, 148, 128), run_t<1, 16, 4, 0>(N, groups, 0, 0, 148, 64),
           run_t<1, 16, 4, 1>(N, groups, 0, 0, 148, 128), run_t<1, 16, 4, 1>(N, groups, 0, 0, 148, 64),
           (N % 16 == 0) ? run_t<2, 16, 4, 2>(N, groups, 0, 0, 148, 128) : 0.0);
    fflush(stdout);
  }
  printf("\nTS cg1 N=48 PG=16 W=4 with 16 STTM streamer warps (gap = nanosleep(20) count between stores)\n");
  for (int gap : {0, 1, 2, 4, 8, 16}) printf("gap %d: %.1f\n", gap, run_t<1, 16, 4, 0>(48, groups, 1, gap, 148, 128));
  return 0;
}
