// What the two tensor-core sweeps (tc_kernel.cu: int8 digits, tc4_kernel.cu: E2M1 digits) share besides the PTX layer
// (tc_ptx.cuh): compile-time tags, the per-tile "any missing call" test on the row flags, the tensor-map encoder and the
// persistent cluster launch.
#pragma once
#include <cuda.h>
#include <string.h>

#include <string>

#include "common.cuh"

namespace lrr {
namespace tcc {

template <int V> struct IntTag { static constexpr int value = V; };
struct TrueTag { static constexpr bool value = true; };
struct FalseTag { static constexpr bool value = false; };

// does any of the `tile_rows` rows of tile `tile` hold a missing call?  (row_flags: uint8 per row, NULL = unknown)
__device__ __forceinline__ bool tile_flags_any(const uint8_t* __restrict__ row_flags, int64_t M, int tile, int tile_rows) {
  const int64_t r0 = (int64_t)tile * tile_rows;
  if (r0 >= M) return false;   // padding tile of a cluster
  if (!row_flags) return true;
  uint32_t any = 0;
  if (r0 + tile_rows <= M && ((reinterpret_cast<uintptr_t>(row_flags + r0) & 15) == 0)) {
    const uint4* f = reinterpret_cast<const uint4*>(row_flags + r0);
    for (int i = 0; i < tile_rows / 16; ++i) {
      const uint4 v = __ldg(f + i);
      any |= v.x | v.y | v.z | v.w;
    }
  } else {
    for (int64_t r = r0; r < M && r < r0 + tile_rows; ++r) any |= row_flags[r];
  }
  return any != 0;
}

__global__ void mask_hi_kernel(const uint32_t* __restrict__ mask_lo, int64_t words, uint32_t* __restrict__ mask_hi);

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda); NULL + `why`
// when the driver does not provide it
inline EncodeTiledFn get_encode_fn(std::string* why) {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
    cudaGetLastError();
    if (why) *why = "cuTensorMapEncodeTiled is not available from the driver";
    return nullptr;
  }
  return reinterpret_cast<EncodeTiledFn>(fn);
}

// 2-D uint8 tensor map, 128-byte swizzle: [outer rows][inner bytes], row stride `row_stride`, box [box_outer][box_inner]
inline int encode_2d_u8(EncodeTiledFn encode, CUtensorMap* map, const void* ptr, uint64_t inner, uint64_t outer, uint64_t row_stride,
                        uint32_t box_inner, uint32_t box_outer) {
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {row_stride};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                      tuning_env("LRR_ABL_L2P128") ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B
                      : tuning_env("LRR_ABL_L2PNONE") ? CU_TENSOR_MAP_L2_PROMOTION_NONE : CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : (int)r;
}

// Launch a persistent sweep: clusters of `cs` CTAs, as many as are co-resident (one CTA per SM), never more than the
// tiles need.  `args` = the kernel's three parameters (genotype map, basis map, Params).
inline int launch_persistent_clusters(Ctx* c, void* kfn, int cs, int threads, int smem_bytes, int n_tiles, void** args,
                                      cudaStream_t st) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.blockDim = dim3((unsigned)threads);
  cfg.dynamicSmemBytes = (size_t)smem_bytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)cs;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int max_clusters = c->sm_count / cs;
  if (cs > 1) {
    cfg.gridDim = dim3((unsigned)(c->sm_count / cs * cs));
    int nc = 0;
    if (cudaOccupancyMaxActiveClusters(&nc, kfn, &cfg) == cudaSuccess && nc > 0) max_clusters = nc;
    else cudaGetLastError();
  }
  int n_cta = max_clusters * cs;
  const int need = (n_tiles + cs - 1) / cs * cs;
  if (n_cta > need) n_cta = need;
  cfg.gridDim = dim3((unsigned)n_cta);
  LRR_CUDA(c, cudaLaunchKernelExC(&cfg, kfn, args));
  c->launches++;
  LRR_CUDA(c, cudaGetLastError());
  return LRR_OK;
}

}  // namespace tcc
}  // namespace lrr
