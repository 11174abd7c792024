// Genotype ingest: PLINK .bed / int8 dosages -> the device 2-bit store, its inverse, and the seeded
// Balding-Nichols style synthetic fill used by the benchmark.
//
// Reference behaviour restated here (not copied):
//   .bed decode   hail/hail/src/is/hail/io/plink/LoadPlink.scala:475-481 (a2_reference=True: code 0 -> hom-alt,
//                 1 -> missing, 2 -> het, 3 -> hom-ref) and :525 (sample i at byte i>>2, bits (i&3)<<1)
//   x expression  GT.n_alt_alleles()  (hail/hail/src/is/hail/variant/Call.scala:430-437)
//   BN generator  hail/python/hail/methods/statgen.py:4254-4291 (genotype ~ Cat(q^2, 2pq, p^2) given the
//                 sample's population allele frequency)
#include "common.cuh"

namespace lrr {

// (byte b, position p) -> (byte p, position b) for the sixteen 2-bit fields of a word.
__device__ __forceinline__ uint32_t transpose_fields(uint32_t x) {
  uint32_t t = ((x >> 6) ^ x) & 0x00CC00CCu;
  x ^= t ^ (t << 6);
  t = ((x >> 12) ^ x) & 0x0000F0F0u;
  x ^= t ^ (t << 12);
  return x;
}

// mask with the 2-bit fields of the first `valid` samples (store order) set
__device__ __forceinline__ uint32_t valid_mask_store_order(int valid) {
  if (valid >= 16) return 0xFFFFFFFFu;
  uint32_t m = 0;
  for (int j = 0; j < valid; ++j) m |= 3u << sample_shift(j);
  return m;
}

// 1 where a store word holds at least one missing call (code 3)
__device__ __forceinline__ bool word_has_missing(uint32_t w) { return (w & (w >> 1) & 0x55555555u) != 0u; }

__global__ void pack_bed_kernel(const uint8_t* __restrict__ bed, int64_t M, int64_t bed_stride, int64_t N,
                                uint8_t* __restrict__ packed, int64_t packed_stride, uint8_t* __restrict__ flags) {
  const int64_t words_per_row = packed_stride / 4;
  const int64_t total = M * words_per_row;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t v = idx / words_per_row;
    const int64_t w = idx - v * words_per_row;
    const int64_t s0 = w * 16;
    uint32_t out = 0;
    if (s0 < N) {
      const uint8_t* row = bed + v * bed_stride;
      uint32_t x = 0;
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int64_t byte = w * 4 + b;
        if (byte < bed_stride && byte * 4 < N) x |= (uint32_t)row[byte] << (8 * b);
      }
      // recode: bed (hi,lo) 00->10(2 alt) 01->11(missing) 10->01(het) 11->00(hom-ref)
      const uint32_t hi = (x >> 1) & 0x55555555u, lo = x & 0x55555555u;
      const uint32_t nhi = (~hi) & 0x55555555u;
      x = (nhi << 1) | (hi ^ lo);
      x = transpose_fields(x);
      const int64_t left = N - s0;
      out = x & valid_mask_store_order(left >= 16 ? 16 : (int)left);
    }
    if (flags && word_has_missing(out)) flags[v] = 1;  // benign race: every writer stores 1
    reinterpret_cast<uint32_t*>(packed + v * packed_stride)[w] = out;
  }
}

// inverse of pack_bed_kernel: device store -> PLINK SNP-major bytes (what ExportPlink writes,
// hail/hail/src/is/hail/expr/ir/MatrixWriter.scala:2270-2285: hom-ref -> 3, het -> 2, hom-alt -> 0, missing -> 1)
__global__ void unpack_bed_kernel(const uint8_t* __restrict__ packed, int64_t packed_stride, int64_t M, int64_t N,
                                  uint8_t* __restrict__ bed, int64_t bed_stride) {
  const int64_t words = (bed_stride + 3) / 4;
  const int64_t total = M * words;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t v = idx / words;
    const int64_t w = idx - v * words;
    uint32_t x = (w * 4 < packed_stride) ? reinterpret_cast<const uint32_t*>(packed + v * packed_stride)[w] : 0u;
    x = transpose_fields(x);  // the field transpose is an involution
    const uint32_t hi = (x >> 1) & 0x55555555u, lo = x & 0x55555555u;
    x = (((~hi) & 0x55555555u) << 1) | ((~(hi ^ lo)) & 0x55555555u);
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int64_t byte = w * 4 + b;
      if (byte < bed_stride) {
        uint32_t val = (x >> (8 * b)) & 0xFFu;
        const int64_t left = N - byte * 4;  // samples covered by this byte
        if (left <= 0) val = 0;
        else if (left < 4) val &= (1u << (2 * left)) - 1u;  // PLINK pads the last byte with zero bits
        bed[v * bed_stride + byte] = (uint8_t)val;
      }
    }
  }
}

__global__ void pack_i8_kernel(const int8_t* __restrict__ dos, int64_t M, int64_t N, uint8_t* __restrict__ packed,
                               int64_t packed_stride, uint8_t* __restrict__ flags) {
  const int64_t words_per_row = packed_stride / 4;
  const int64_t total = M * words_per_row;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t v = idx / words_per_row;
    const int64_t w = idx - v * words_per_row;
    const int64_t s0 = w * 16;
    uint32_t out = 0;
    const int8_t* row = dos + v * N;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      if (s0 + j < N) {
        const int d = row[s0 + j];
        const uint32_t code = (d >= 0 && d <= 2) ? (uint32_t)d : 3u;
        out |= code << sample_shift(j);
      }
    }
    if (flags && word_has_missing(out)) flags[v] = 1;
    reinterpret_cast<uint32_t*>(packed + v * packed_stride)[w] = out;
  }
}

__global__ void unpack_i8_kernel(const uint8_t* __restrict__ packed, int64_t packed_stride, int64_t M, int64_t N,
                                 int8_t* __restrict__ dos) {
  const int64_t total = M * N;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t v = idx / N;
    const int64_t j = idx - v * N;
    const uint32_t word = reinterpret_cast<const uint32_t*>(packed + v * packed_stride)[j >> 4];
    const uint32_t code = (word >> sample_shift((int)(j & 15))) & 3u;
    dos[idx] = code == 3u ? (int8_t)-1 : (int8_t)code;
  }
}

// splitmix64 finaliser: a counter-based generator that numpy can mirror bit-for-bit (tests/bn_model.py)
__device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

// thresholds [M][n_pops][3] as 16-bit-scaled integers in [0, 65536]:
//   u < t0 -> missing; u < t1 -> 0 alt; u < t2 -> 1 alt; else 2 alt      (u = 16 random bits)
__global__ void bn_fill_kernel(const uint32_t* __restrict__ thresh, int n_pops, const uint8_t* __restrict__ pop,
                               int64_t M, int64_t first_variant, int64_t N, uint64_t seed,
                               uint8_t* __restrict__ packed, int64_t packed_stride, uint8_t* __restrict__ flags) {
  const int64_t words_per_row = packed_stride / 4;
  const int64_t total = M * words_per_row;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = idx / words_per_row;
    const int64_t w = idx - r * words_per_row;
    const int64_t s0 = w * 16;
    uint32_t out = 0;
    if (s0 < N) {
      const uint64_t v = (uint64_t)(first_variant + r);
      const uint32_t* th = thresh + r * n_pops * 3;
      const uint64_t base = seed ^ (v * 0xD1B54A32D192ED03ull) ^ ((uint64_t)w * 0x8CB92BA72F3D8DD7ull);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint64_t bits = mix64(base + (uint64_t)q);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int j = q * 4 + e;
          if (s0 + j < N) {
            const uint32_t u = (uint32_t)(bits >> (16 * e)) & 0xFFFFu;
            const uint32_t* t = th + 3 * pop[s0 + j];
            const uint32_t code = u < t[0] ? 3u : (u < t[1] ? 0u : (u < t[2] ? 1u : 2u));
            out |= code << sample_shift(j);
          }
        }
      }
    }
    if (flags && word_has_missing(out)) flags[r] = 1;
    reinterpret_cast<uint32_t*>(packed + r * packed_stride)[w] = out;
  }
}

static int grid_for(const Ctx* c, int64_t total, int block) {
  int64_t g = (total + block - 1) / block;
  const int64_t cap = (int64_t)c->sm_count * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

int launch_pack_bed(Ctx* c, const uint8_t* bed, int64_t M, int64_t bed_stride, int64_t N, uint8_t* packed,
                    int64_t packed_stride, uint8_t* flags, cudaStream_t st) {
  if (M == 0) return LRR_OK;
  if (flags) LRR_CUDA(c, cudaMemsetAsync(flags, 0, (size_t)M, st));
  pack_bed_kernel<<<grid_for(c, M * (packed_stride / 4), 256), 256, 0, st>>>(bed, M, bed_stride, N, packed,
                                                                               packed_stride, flags);
  c->launches++;
  LRR_CUDA(c, cudaGetLastError());
  return LRR_OK;
}

int launch_pack_i8(Ctx* c, const int8_t* dos, int64_t M, int64_t N, uint8_t* packed, int64_t packed_stride,
                   uint8_t* flags, cudaStream_t st) {
  if (M == 0) return LRR_OK;
  if (flags) LRR_CUDA(c, cudaMemsetAsync(flags, 0, (size_t)M, st));
  pack_i8_kernel<<<grid_for(c, M * (packed_stride / 4), 256), 256, 0, st>>>(dos, M, N, packed, packed_stride, flags);
  c->launches++;
  LRR_CUDA(c, cudaGetLastError());
  return LRR_OK;
}

int launch_unpack_i8(Ctx* c, const uint8_t* packed, int64_t packed_stride, int64_t M, int64_t N, int8_t* dos,
                     cudaStream_t st) {
  if (M == 0 || N == 0) return LRR_OK;
  unpack_i8_kernel<<<grid_for(c, M * N, 256), 256, 0, st>>>(packed, packed_stride, M, N, dos);
  c->launches++;
  LRR_CUDA(c, cudaGetLastError());
  return LRR_OK;
}

int launch_unpack_bed(Ctx* c, const uint8_t* packed, int64_t packed_stride, int64_t M, int64_t N, uint8_t* bed,
                      int64_t bed_stride, cudaStream_t st) {
  if (M == 0) return LRR_OK;
  unpack_bed_kernel<<<grid_for(c, M * ((bed_stride + 3) / 4), 256), 256, 0, st>>>(packed, packed_stride, M, N, bed,
                                                                                   bed_stride);
  c->launches++;
  LRR_CUDA(c, cudaGetLastError());
  return LRR_OK;
}

int launch_bn_fill(Ctx* c, const uint32_t* thresh, int n_pops, const uint8_t* pop, int64_t M, int64_t first_variant,
                   int64_t N, uint64_t seed, uint8_t* packed, int64_t packed_stride, uint8_t* flags, cudaStream_t st) {
  if (M == 0) return LRR_OK;
  if (flags) LRR_CUDA(c, cudaMemsetAsync(flags, 0, (size_t)M, st));
  bn_fill_kernel<<<grid_for(c, M * (packed_stride / 4), 256), 256, 0, st>>>(thresh, n_pops, pop, M, first_variant, N,
                                                                              seed, packed, packed_stride, flags);
  c->launches++;
  LRR_CUDA(c, cudaGetLastError());
  return LRR_OK;
}

}  // namespace lrr
