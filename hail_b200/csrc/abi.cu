// extern "C" boundary of lrr_b200 (see include/lrr_b200.h for what each entry point replaces in the reference).
#include <math.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include <chrono>
#include <cstdio>

#include "common.cuh"

namespace lrr {

static thread_local std::string g_create_error;

int fail(Ctx* c, int code, const std::string& msg) {
  if (c) c->err = msg; else g_create_error = msg;
  return code;
}

int cuda_fail(Ctx* c, cudaError_t e, const char* what) {
  std::string m = std::string("CUDA error: ") + cudaGetErrorString(e) + " in " + what;
  cudaGetLastError();  // clear sticky-less errors so the next call reports its own
  return fail(c, LRR_ECUDA, m);
}

int abi_caught(void* ctx, int code, const char* what) noexcept {
  try {
    std::string m = std::string("internal error: ") + (what ? what : "?");
    if (ctx) reinterpret_cast<Ctx*>(ctx)->err = m; else g_create_error = m;
  } catch (...) {
  }
  return code;
}

namespace {

void free_group(Group& g) {
  cudaFree(g.d_basis);
  cudaFree(g.d_qty);
  cudaFree(g.d_yyp);
  cudaFree(g.d_mask);
  cudaFree(g.d_bq);
  cudaFree(g.d_colscale);
  cudaFree(g.d_basis_t);
  g = Group();
}

// a retired group keeps its four input buffers for the next lrr_add_group (everything derived from them goes)
void retire_group(Ctx* c, Group& g) {
  if (g.d_bq) cudaFree(g.d_bq);
  if (g.d_colscale) cudaFree(g.d_colscale);
  if (g.d_basis_t) cudaFree(g.d_basis_t);
  g.d_bq = nullptr;
  g.d_colscale = nullptr;
  g.d_basis_t = nullptr;
  if (c->spare.size() < 8) c->spare.push_back(g); else free_group(g);
}

void free_workspace(Ctx* c) {
  cudaFree(c->d_counts);
  cudaFree(c->d_dots);
  cudaFree(c->d_flag_mark);
  cudaFree(c->d_flag_list);
  cudaFree(c->d_flag_count);
  c->d_counts = nullptr;
  c->d_dots = nullptr;
  c->d_flag_mark = c->d_flag_list = c->d_flag_count = nullptr;
  c->flag_pending = false;
  c->reserved_variants = 0;
  c->ws_groups = c->ws_cols = 0;
  c->dots_offset.clear();
}

// scatter compact per-kept-sample columns into zero-initialised full-width planes and build the sample mask.  The
// columns come from up to four sources (covariate block, phenotype block, and the two extra columns of weighted /
// score groups); a source may be device memory or page-locked host memory mapped into the device's address space.
struct ScatterSrc {
  const double* ptr[4];
  int first[5];   // source k holds columns [first[k], first[k + 1])
};

__global__ void scatter_basis_kernel(ScatterSrc src, int C, int n, const int32_t* __restrict__ idx, int64_t ns_pad,
                                     double* __restrict__ basis, uint32_t* __restrict__ mask) {
  const int64_t total = (int64_t)C * n;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i / n);
    const int j = (int)(i - (int64_t)c * n);
    const int64_t s = idx[j];
    const int k = c < src.first[1] ? 0 : c < src.first[2] ? 1 : c < src.first[3] ? 2 : 3;
    basis[(int64_t)c * ns_pad + s] = src.ptr[k][(int64_t)(c - src.first[k]) * n + j];
    if (c == 0) atomicOr(mask + (s >> 4), 1u << sample_shift((int)(s & 15)));
  }
}

__global__ void copy_doubles_kernel(double* __restrict__ dst, const double* __restrict__ src, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) dst[i] = src[i];
}

// Where lrr_add_group's input arrays are read from.  HOST arrays are copied (by the CPU) into a page-locked buffer that is
// mapped into the device's address space, and the kernels above read it over PCIe directly: a cudaMemcpy would queue behind
// every host-to-device copy the streaming loop has in flight on the same copy engine (lrr_stream_begin runs first) and the
// first sweep could not start before the last genotype block had arrived.  DEVICE arrays are read in place.
struct Staging {
  Ctx* c;
  size_t off = 0;
  explicit Staging(Ctx* ctx) : c(ctx) {}
  static bool on_device(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
      cudaGetLastError();
      return false;
    }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
  }
  // device-readable pointer to `bytes` at `p` (NULL on allocation failure; *err is set then)
  const void* in(const void* p, size_t bytes, cudaError_t* err) {
    if (!p || bytes == 0 || on_device(p)) return p;
    const size_t at = (off + 255) / 256 * 256;
    memcpy(static_cast<char*>(c->h_stage) + at, p, bytes);
    off = at + bytes;
    (void)err;
    return static_cast<const char*>(c->d_stage_view) + at;
  }
};

int ensure_staging(Ctx* c, size_t bytes) {
  // the previous lrr_add_group's kernels may still be reading the buffer
  if (c->stage_ev_valid) LRR_CUDA(c, cudaEventSynchronize(c->stage_ev));
  if (bytes <= c->h_stage_bytes) return LRR_OK;
  if (c->h_stage) cudaFreeHost(c->h_stage);
  c->h_stage = nullptr;
  c->h_stage_bytes = 0;
  LRR_CUDA(c, cudaHostAlloc(&c->h_stage, bytes, cudaHostAllocMapped | cudaHostAllocPortable));
  LRR_CUDA(c, cudaHostGetDevicePointer(&c->d_stage_view, c->h_stage, 0));
  c->h_stage_bytes = bytes;
  return LRR_OK;
}

// grow-only: [G][rows] counts / flag arrays and [sum_g (C_g + 2)][rows] dot products
int ensure_workspace(Ctx* c, int64_t M) {
  const int64_t G = (int64_t)c->groups.size();
  int64_t total_c = 0;
  for (const Group& g : c->groups) total_c += g.C + 2;   // + 2: fitted-value dots (tc4) / column sum and squares (dense)
  if (M > c->reserved_variants || G > c->ws_groups || total_c > c->ws_cols) {
    const int64_t rows = std::max<int64_t>(M, c->reserved_variants);
    const int64_t groups = std::max<int64_t>(std::max<int64_t>(G, c->ws_groups), 1);
    const int64_t cols = std::max<int64_t>(std::max<int64_t>(total_c, c->ws_cols), 1);
    LRR_CUDA(c, cudaDeviceSynchronize());   // (growing only: a run may still be using the old buffers)
    free_workspace(c);
    LRR_CUDA(c, cudaMalloc(&c->d_counts, sizeof(int32_t) * 4 * (size_t)rows * (size_t)groups));
    LRR_CUDA(c, cudaMalloc(&c->d_dots, sizeof(double) * (size_t)rows * (size_t)cols));
    LRR_CUDA(c, cudaMalloc(&c->d_flag_mark, sizeof(int32_t) * (size_t)rows * (size_t)groups));
    LRR_CUDA(c, cudaMalloc(&c->d_flag_list, sizeof(int32_t) * (size_t)rows * (size_t)groups));
    LRR_CUDA(c, cudaMalloc(&c->d_flag_count, sizeof(int32_t) * (size_t)groups));
    if ((int)groups > c->h_flag_groups) {
      if (c->h_flag_count) cudaFreeHost(c->h_flag_count);
      c->h_flag_count = nullptr;
      c->h_flag_groups = 0;
      LRR_CUDA(c, cudaMallocHost(&c->h_flag_count, sizeof(int32_t) * (size_t)groups));
      c->h_flag_groups = (int)groups;
    }
    if (!c->flag_ev) LRR_CUDA(c, cudaEventCreateWithFlags(&c->flag_ev, cudaEventDisableTiming));
    c->reserved_variants = rows;
    c->ws_groups = groups;
    c->ws_cols = cols;
  }
  {
    // the deferred-tail list of the statistics epilogue: 1/16 of the largest group's (variant, phenotype) entries
    int64_t maxP = 1;
    for (const Group& g : c->groups) maxP = std::max<int64_t>(maxP, g.P);
    const int64_t want = M * maxP / 16 + 4096;
    if (want > c->tail_capacity) {
      LRR_CUDA(c, cudaDeviceSynchronize());
      cudaFree(c->d_tail);
      c->d_tail = nullptr;
      c->tail_capacity = 0;
      LRR_CUDA(c, cudaMalloc(&c->d_tail, sizeof(int64_t) * 2 * (size_t)want));
      if (!c->d_tail_count) LRR_CUDA(c, cudaMalloc(&c->d_tail_count, sizeof(int32_t)));
      c->tail_capacity = want;
    }
  }
  if (c->dots_offset.size() != (size_t)G) {
    c->dots_offset.resize((size_t)G);
    int64_t off = 0;
    for (int64_t g = 0; g < G; ++g) {
      c->dots_offset[(size_t)g] = off * c->reserved_variants;
      off += c->groups[(size_t)g].C + 2;
    }
  }
  return LRR_OK;
}

}  // namespace

void release_caches(Ctx* c) {
  for (auto& g : c->spare) free_group(g);
  c->spare.clear();
  cudaFree(c->d_recompute);
  c->d_recompute = nullptr;
  c->recompute_bytes = 0;
  cudaFree(c->d_tail);
  c->d_tail = nullptr;
  c->tail_capacity = 0;
  if (c->h_stage) cudaFreeHost(c->h_stage);
  c->h_stage = nullptr;
  c->d_stage_view = nullptr;
  c->h_stage_bytes = 0;
  c->stage_ev_valid = false;
  if (c->groups.empty()) {
    free_workspace(c);
    tc4_trim(c);
  }
}

int run_begin(Ctx* c, cudaStream_t st) {
  if (c->ready_valid) LRR_CUDA(c, cudaStreamWaitEvent(st, c->ready_ev, 0));
  return LRR_OK;
}

int run_end(Ctx* c, cudaStream_t st) {
  if (!c->busy_ev) LRR_CUDA(c, cudaEventCreateWithFlags(&c->busy_ev, cudaEventDisableTiming));
  LRR_CUDA(c, cudaEventRecord(c->busy_ev, st));
  c->busy_valid = true;
  return LRR_OK;
}

namespace {
constexpr int64_t kPilotRows = 8192;   // rows of the precision pilot (64 tiles)
int run_rows_once(Ctx* c, const uint8_t* d_packed, const uint8_t* d_row_flags, int64_t n_variants, int64_t packed_stride,
                  const lrr_group_out* outs, int k, cudaStream_t st);
}  // namespace

// the hot call behind lrr_run and the streaming loop (stream.cu): sweep + per-variant statistics of one row block
int run_rows(Ctx* c, const uint8_t* d_packed, const uint8_t* d_row_flags, int64_t n_variants, int64_t packed_stride,
             int64_t n_samples_total, const lrr_group_out* outs, int32_t n_outs, int32_t kernel, cudaStream_t st) {
  if (c->groups.empty()) return fail(c, LRR_ESTATE, "lrr_run: no groups (call lrr_add_group)");
  if (c->groups[0].score) return fail(c, LRR_ESTATE, "lrr_run: the context holds a logistic score model (use lrr_run_score)");
  if (n_outs != (int32_t)c->groups.size() || !outs) return fail(c, LRR_EINVAL, "lrr_run: need one lrr_group_out per group");
  if (n_variants < 0) return fail(c, LRR_EINVAL, "lrr_run: negative n_variants");
  if (n_samples_total != c->n_samples_total) return fail(c, LRR_EINVAL, "lrr_run: n_samples_total differs from the groups'");
  if (packed_stride % kRowAlignBytes != 0 || packed_stride * 4 < n_samples_total)
    return fail(c, LRR_EINVAL, "packed_stride must be a multiple of 128 bytes covering n_samples (use lrr_packed_stride)");
  if (packed_stride * 4 != c->groups[0].ns_pad) return fail(c, LRR_EINVAL, "lrr_run: packed_stride must equal lrr_packed_stride(n_samples_total)");
  if (n_variants == 0) return LRR_OK;
  if (!d_packed) return fail(c, LRR_EINVAL, "lrr_run: d_packed is NULL");
  if (int r = ensure_workspace(c, n_variants)) return r;
  if (int r = run_begin(c, st)) return r;

  int k = kernel;
  const bool may_miss = true;  // the column budget is checked for the general (two-plane) mode
  // AUTO: small problems go straight to the float64 CUDA-core kernel (quantising the basis would cost more than the
  // sweep: ~4e12 genotype-columns / s against ~0.5 ms of preparation); otherwise the 4-bit tensor-core sweep (as many
  // passes as the columns need) when its exactness bound holds, else the int8 tensor-core sweep, else the float64 kernel
  if (k == LRR_KERNEL_AUTO) {
    double cols = 0.0;
    for (const Group& g : c->groups) cols += g.C;
    const double work = (double)n_variants * (double)c->groups[0].ns_pad * cols;
    if (work <= 2e9) k = LRR_KERNEL_FP64;
    else k = tc4_supported(c, false, d_row_flags, n_variants, st) ? LRR_KERNEL_TC4 : tc_supported(c, may_miss) ? LRR_KERNEL_TC : LRR_KERNEL_FP64;
  }
  // Adaptive precision (4-bit sweep), once per group set and deterministic: the first kPilotRows rows (or the whole of
  // a shorter run, e.g. one block of the streaming loop) run as a pilot;
  // when more than 2 % of them leave the tolerance guard (structured or badly scaled covariates: |Q'x| is large for
  // every row), the covariate / fitted-value columns are re-quantised with two more digits, up to three times.  This
  // is the only place the hot call waits for the device, and only on the first call after the groups changed.
  if (k == LRR_KERNEL_TC4 && c->guard && !c->pilot_done && n_variants >= 1024) {
    c->pilot_done = true;
    const int64_t pilot_rows = std::min<int64_t>(kPilotRows, n_variants);   // (a short block of the streaming loop: all of it)
    const int saved_timing = c->timing;
    if (pilot_rows != n_variants) c->timing = 0;   // (the sweep timer brackets the run proper)
    for (;;) {
      int r = LRR_OK;
      if (!tc4_supported(c, false, d_row_flags, n_variants, st))
        r = fail(c, LRR_EINVAL, "lrr_run: 4-bit tensor-core kernel does not support this configuration: " + c->err);
      if (!r) r = run_rows_once(c, d_packed, d_row_flags, pilot_rows, packed_stride, outs, k, st);
      if (!r) {
        const cudaError_t e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) r = cuda_fail(c, e, "cudaStreamSynchronize (precision pilot)");
      }
      if (r) {
        c->timing = saved_timing;
        return r;
      }
      int64_t worst = 0;
      for (size_t g = 0; g < c->groups.size(); ++g) worst = std::max<int64_t>(worst, c->h_flag_count[g]);
      if (worst * 50 <= pilot_rows || c->digit_boost >= 6) break;
      c->digit_boost += 2;
      tc4_invalidate(c);
    }
    c->timing = saved_timing;
    if (pilot_rows == n_variants) return LRR_OK;   // the pilot was the run
  }
  return run_rows_once(c, d_packed, d_row_flags, n_variants, packed_stride, outs, k, st);
}

namespace {
int run_rows_once(Ctx* c, const uint8_t* d_packed, const uint8_t* d_row_flags, int64_t n_variants, int64_t packed_stride,
                  const lrr_group_out* outs, int k, cudaStream_t st) {
  const bool may_miss = true;
  if (c->timing) {
    if (!c->ev0) {
      LRR_CUDA(c, cudaEventCreate(&c->ev0));
      LRR_CUDA(c, cudaEventCreate(&c->ev1));
    }
    LRR_CUDA(c, cudaEventRecord(c->ev0, st));
  }
  c->sweep_shape[0] = c->sweep_shape[1] = c->sweep_shape[2] = c->sweep_shape[3] = 0;
  if (k == LRR_KERNEL_TC4) {
    if (!tc4_supported(c, false, d_row_flags, n_variants, st))
      return fail(c, LRR_EINVAL, "lrr_run: 4-bit tensor-core kernel does not support this configuration: " + c->err);
    if (int r = launch_tc4_sweep(c, d_packed, d_row_flags, n_variants, packed_stride, st)) return r;
  } else if (k == LRR_KERNEL_TC) {
    if (!tc_supported(c, may_miss))
      return fail(c, LRR_EINVAL, "lrr_run: tensor-core kernel does not support this configuration: " + c->err);
    if (int r = launch_tc_sweep(c, d_packed, d_row_flags, n_variants, packed_stride, st)) return r;
  } else if (k == LRR_KERNEL_FP64) {
    if (int r = launch_fp64_sweep(c, d_packed, n_variants, packed_stride, st)) return r;
  } else {
    return fail(c, LRR_EINVAL, "lrr_run: unknown kernel id");
  }
  c->last_kernel = k;
  if (c->timing) {
    LRR_CUDA(c, cudaEventRecord(c->ev1, st));
    c->ev_valid = true;
  }
  const size_t G = c->groups.size();
  const bool quantised = (k == LRR_KERNEL_TC4 || k == LRR_KERNEL_TC);
  const bool guarded = quantised && c->guard;
  if (guarded) {
    LRR_CUDA(c, cudaMemsetAsync(c->d_flag_count, 0, sizeof(int32_t) * G, st));
    for (size_t g = 0; g < G; ++g)
      LRR_CUDA(c, cudaMemsetAsync(c->d_flag_mark + (int64_t)g * c->reserved_variants, 0, sizeof(int32_t) * (size_t)n_variants, st));
  }
  for (size_t g = 0; g < G; ++g) {
    const double* quantum = nullptr;
    const double* err_sum = nullptr;
    int n_fit = 0, stride = c->groups[g].C;
    double qscale = 1.0;
    if (k == LRR_KERNEL_TC4) {
      quantum = tc4_quantum(c, (int)g, &n_fit);
      err_sum = tc4_errsum(c, (int)g);
      stride = c->groups[g].C + 2;
    } else if (k == LRR_KERNEL_TC) {
      quantum = tc_quantum(c, (int)g);
      qscale = 4.0;
    }
    if (int r = launch_stats_epilogue(c, (int)g, n_variants, outs[g], st, false, quantum, n_fit, stride, qscale, err_sum)) return r;
    if (guarded && !c->groups[g].weighted) {
      // rows the guard listed: float64 recompute of their counts and dot products, then their statistics once more
      const int32_t* list = c->d_flag_list + (int64_t)g * c->reserved_variants;
      if (int r = launch_fp64_recompute(c, (int)g, d_packed, packed_stride, list, c->d_flag_count + g, stride, st)) return r;
      if (int r = launch_stats_epilogue_listed(c, (int)g, outs[g], stride, st)) return r;
    }
  }
  if (guarded) {
    LRR_CUDA(c, cudaMemcpyAsync(c->h_flag_count, c->d_flag_count, sizeof(int32_t) * G, cudaMemcpyDeviceToHost, st));
    LRR_CUDA(c, cudaEventRecord(c->flag_ev, st));
    c->flag_pending = true;
    c->flag_rows = n_variants;
  } else {
    c->flag_pending = false;
    c->flag_rows = 0;
    for (int g = 0; g < c->h_flag_groups; ++g) c->h_flag_count[g] = 0;
  }
  return run_end(c, st);
}
}  // namespace

}  // namespace lrr

using namespace lrr;

extern "C" {

const char* lrr_version(void) { return "lrr_b200 0.1 (sm_100a)"; }

int lrr_create(lrr_ctx** out, int device) try {
  if (!out) return fail(nullptr, LRR_EINVAL, "lrr_create: out is NULL");
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return fail(nullptr, LRR_ECUDA,
                std::string("lrr_create: no CUDA device available (") + cudaGetErrorString(e) +
                    "); this library has no CPU fallback");
  if (device < 0 || device >= count) return fail(nullptr, LRR_EINVAL, "lrr_create: device index out of range");
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) return fail(nullptr, LRR_ECUDA, std::string("lrr_create: ") + cudaGetErrorString(e));
  if (prop.major < 10)
    return fail(nullptr, LRR_ECUDA, "lrr_create: device is not Blackwell (sm_100a) -- kernels are built for sm_100a only");
  Ctx* c = new Ctx();
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  *out = reinterpret_cast<lrr_ctx*>(c);
  return LRR_OK;
}
LRR_ABI_CATCH(nullptr)

void lrr_destroy(lrr_ctx* ctx) try {
  if (!ctx) return;
  Ctx* c = reinterpret_cast<Ctx*>(ctx);
  DeviceGuard guard(c->device);
  for (auto& g : c->groups) free_group(g);
  for (auto& g : c->spare) free_group(g);
  if (c->h_stage) cudaFreeHost(c->h_stage);
  if (c->stage_ev) cudaEventDestroy(c->stage_ev);
  cudaFree(c->d_tail);
  cudaFree(c->d_tail_count);
  cudaFree(c->d_recompute);
  if (c->busy_ev) cudaEventDestroy(c->busy_ev);
  if (c->ready_ev) cudaEventDestroy(c->ready_ev);
  free_workspace(c);
  tc_release(c);
  tc4_release(c);
  logit_release(c);
  cudaFree(c->arena);
  cudaFree(c->d_nanmask);
  if (c->h_flag_count) cudaFreeHost(c->h_flag_count);
  if (c->flag_ev) cudaEventDestroy(c->flag_ev);
  if (c->ev0) cudaEventDestroy(c->ev0);
  if (c->ev1) cudaEventDestroy(c->ev1);
  delete c;
} catch (...) {
}

const char* lrr_last_error(const lrr_ctx* ctx) {
  if (!ctx) return g_create_error.c_str();
  return reinterpret_cast<const Ctx*>(ctx)->err.c_str();
}

int64_t lrr_packed_stride(int64_t n_samples) {
  if (n_samples <= 0) return kRowAlignBytes;
  const int64_t bytes = (n_samples + 3) / 4;
  return (bytes + kRowAlignBytes - 1) / kRowAlignBytes * kRowAlignBytes;
}

#define CTX_PROLOGUE                                      \
  if (!ctx) return LRR_EINVAL;                            \
  Ctx* c = reinterpret_cast<Ctx*>(ctx);                   \
  DeviceGuard guard(c->device);                           \
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream)

static int check_packed(Ctx* c, int64_t n_samples, int64_t packed_stride) {
  if (packed_stride % kRowAlignBytes != 0 || packed_stride * 4 < n_samples)
    return fail(c, LRR_EINVAL, "packed_stride must be a multiple of 128 bytes covering n_samples (use lrr_packed_stride)");
  return LRR_OK;
}

int lrr_pack_bed(lrr_ctx* ctx, const uint8_t* d_bed, int64_t n_variants, int64_t bed_stride, int64_t n_samples,
                 uint8_t* d_packed, int64_t packed_stride, uint8_t* d_row_flags, void* stream) try {
  CTX_PROLOGUE;
  if (n_variants < 0 || n_samples <= 0 || bed_stride < (n_samples + 3) / 4)
    return fail(c, LRR_EINVAL, "lrr_pack_bed: bed_stride must be >= ceil(n_samples/4) (LoadPlink.scala:240-251)");
  if (int r = check_packed(c, n_samples, packed_stride)) return r;
  return launch_pack_bed(c, d_bed, n_variants, bed_stride, n_samples, d_packed, packed_stride, d_row_flags, st);
}
LRR_ABI_CATCH(ctx)

int lrr_pack_dosage_i8(lrr_ctx* ctx, const int8_t* d_dosage, int64_t n_variants, int64_t n_samples, uint8_t* d_packed,
                       int64_t packed_stride, uint8_t* d_row_flags, void* stream) try {
  CTX_PROLOGUE;
  if (n_variants < 0 || n_samples <= 0) return fail(c, LRR_EINVAL, "lrr_pack_dosage_i8: bad shape");
  if (int r = check_packed(c, n_samples, packed_stride)) return r;
  return launch_pack_i8(c, d_dosage, n_variants, n_samples, d_packed, packed_stride, d_row_flags, st);
}
LRR_ABI_CATCH(ctx)

int lrr_unpack_dosage_i8(lrr_ctx* ctx, const uint8_t* d_packed, int64_t packed_stride, int64_t n_variants,
                         int64_t n_samples, int8_t* d_dosage, void* stream) try {
  CTX_PROLOGUE;
  if (n_variants < 0 || n_samples <= 0) return fail(c, LRR_EINVAL, "lrr_unpack_dosage_i8: bad shape");
  if (int r = check_packed(c, n_samples, packed_stride)) return r;
  return launch_unpack_i8(c, d_packed, packed_stride, n_variants, n_samples, d_dosage, st);
}
LRR_ABI_CATCH(ctx)

int lrr_unpack_bed(lrr_ctx* ctx, const uint8_t* d_packed, int64_t packed_stride, int64_t n_variants, int64_t n_samples,
                   uint8_t* d_bed, int64_t bed_stride, void* stream) try {
  CTX_PROLOGUE;
  if (n_variants < 0 || n_samples <= 0 || bed_stride < (n_samples + 3) / 4)
    return fail(c, LRR_EINVAL, "lrr_unpack_bed: bed_stride must be >= ceil(n_samples/4)");
  if (int r = check_packed(c, n_samples, packed_stride)) return r;
  return launch_unpack_bed(c, d_packed, packed_stride, n_variants, n_samples, d_bed, bed_stride, st);
}
LRR_ABI_CATCH(ctx)

int lrr_bn_fill(lrr_ctx* ctx, const uint32_t* d_thresholds, int n_pops, const uint8_t* d_pop, int64_t n_variants,
                int64_t first_variant, int64_t n_samples, uint64_t seed, uint8_t* d_packed, int64_t packed_stride,
                uint8_t* d_row_flags, void* stream) try {
  CTX_PROLOGUE;
  if (n_variants < 0 || n_samples <= 0 || n_pops <= 0 || n_pops > 255) return fail(c, LRR_EINVAL, "lrr_bn_fill: bad shape");
  if (int r = check_packed(c, n_samples, packed_stride)) return r;
  return launch_bn_fill(c, d_thresholds, n_pops, d_pop, n_variants, first_variant, n_samples, seed, d_packed,
                        packed_stride, d_row_flags, st);
}
LRR_ABI_CATCH(ctx)

int lrr_clear_groups(lrr_ctx* ctx) try {
  if (!ctx) return LRR_EINVAL;
  Ctx* c = reinterpret_cast<Ctx*>(ctx);
  DeviceGuard guard(c->device);
  // no device synchronisation: the buffers are kept (retire_group) and their next writer, lrr_add_group, is ordered
  // behind the runs still in flight through busy_ev
  const bool trace = tuning_env("LRR_TRACE") != nullptr;
  auto now_us = [] { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  const double tr0 = trace ? now_us() : 0.0;
  for (auto& g : c->groups) retire_group(c, g);
  c->groups.clear();
  tc_invalidate(c);
  tc4_invalidate(c);
  if (trace) fprintf(stderr, "[lrr trace] clear_groups: %.0f us\n", now_us() - tr0);
  c->dots_offset.clear();
  c->flag_pending = false;
  c->digit_boost = 0;
  c->pilot_done = false;
  c->n_samples_total = 0;
  return LRR_OK;
}
LRR_ABI_CATCH(ctx)

int lrr_num_groups(const lrr_ctx* ctx) {
  return ctx ? (int)reinterpret_cast<const Ctx*>(ctx)->groups.size() : 0;
}

// shared body of lrr_add_group / lrr_add_group_weighted (`sqrt_w` != NULL: weighted group, has_intercept must be 0)
static int add_group_impl(lrr_ctx* ctx, int64_t n_samples_total, int32_t n, int32_t K, int32_t P, int32_t has_intercept,
                          const int32_t* complete_idx, const double* q_cols, const double* y_res, const double* qty,
                          const double* yyp, const double* sqrt_w, const double* w_col = nullptr, int qty_len = -1,
                          int yyp_len = -1) {
  if (!ctx) return LRR_EINVAL;
  Ctx* c = reinterpret_cast<Ctx*>(ctx);
  DeviceGuard guard(c->device);
  if (n_samples_total <= 0 || n <= 0 || n > n_samples_total) return fail(c, LRR_EINVAL, "lrr_add_group: bad sample counts");
  if (P <= 0) return fail(c, LRR_EINVAL, "No phenotypes present.");  // RU:97-98
  if (K < 0 || K + P > kMaxGroupCols) return fail(c, LRR_EINVAL, "lrr_add_group: K + P out of range");
  has_intercept = has_intercept ? 1 : 0;
  if (has_intercept && K < 1) return fail(c, LRR_EINVAL, "lrr_add_group: has_intercept needs K >= 1");
  const int d = n - K - 1;
  if (d < 1) {  // LR:55-58
    char buf[160];
    snprintf(buf, sizeof buf, "%d samples and %d %s (including x) implies %d degrees of freedom.", n, K + 1,
             K == 1 ? "covariate" : "covariates", d);
    return fail(c, LRR_EINVAL, buf);
  }
  if (!c->groups.empty() && c->n_samples_total != n_samples_total)
    return fail(c, LRR_EINVAL, "lrr_add_group: all groups must share n_samples_total");
  if (!complete_idx || !y_res || !yyp || (K > 0 && !qty) || (K - has_intercept > 0 && !q_cols))
    return fail(c, LRR_EINVAL, "lrr_add_group: NULL input array");

  Group g;
  g.n = n;
  g.K = K;
  g.P = P;
  g.has_intercept = has_intercept;
  g.Kd = K - has_intercept;
  g.weighted = sqrt_w ? 1 : 0;
  g.C = g.Kd + P + (sqrt_w ? 2 : 0);
  g.d = d;
  g.lbeta = log_beta_half(0.5 * (double)d);
  g.ns_pad = lrr_packed_stride(n_samples_total) * 4;

  // steady state: no cudaMalloc, no cudaMemcpy and no device synchronisation here.  The four group buffers come from a
  // retired group when one is large enough, host inputs travel through the mapped staging buffer (Staging above), and the
  // kernels below run on the default stream behind busy_ev (the last run that may still read those buffers elsewhere).
  const size_t basis_bytes = sizeof(double) * (size_t)g.C * (size_t)g.ns_pad;
  const size_t mask_bytes = sizeof(uint32_t) * (size_t)(g.ns_pad / 16);
  const size_t n_qty = qty_len >= 0 ? (size_t)qty_len : (size_t)K * P;
  const size_t n_yyp = yyp_len >= 0 ? (size_t)yyp_len : (size_t)P;
  const size_t qty_bytes = sizeof(double) * (n_qty ? n_qty : 1), yyp_bytes = sizeof(double) * (n_yyp ? n_yyp : 1);
  bool pooled = false;
  auto bail = [&](cudaError_t e, const char* what) {
    free_group(g);
    return cuda_fail(c, e, what);
  };
  cudaError_t e;
#define TRY(call) if ((e = (call)) != cudaSuccess) return bail(e, #call)
  // weighted / score groups: column C-2 = sqrt(w), column C-1 = w (host arithmetic on n values)
  std::vector<double> sw, w;
  if (sqrt_w) {
    sw.resize((size_t)n);
    w.resize((size_t)n);
    TRY(cudaMemcpy(sw.data(), sqrt_w, sizeof(double) * (size_t)n, cudaMemcpyDefault));
    for (int i = 0; i < n; ++i) w[(size_t)i] = sw[(size_t)i] * sw[(size_t)i];
    if (w_col) TRY(cudaMemcpy(w.data(), w_col, sizeof(double) * (size_t)n, cudaMemcpyDefault));
  }
  const size_t stage_need = sizeof(int32_t) * (size_t)n + sizeof(double) * ((size_t)g.C * (size_t)n + n_qty + n_yyp) + 8 * 256;
  if (int r = ensure_staging(c, stage_need)) {
    free_group(g);
    return r;
  }
  Staging stage(c);
  const int32_t* v_idx = static_cast<const int32_t*>(stage.in(complete_idx, sizeof(int32_t) * (size_t)n, &e));
  ScatterSrc src;
  src.ptr[0] = static_cast<const double*>(stage.in(q_cols, sizeof(double) * (size_t)g.Kd * (size_t)n, &e));
  src.ptr[1] = static_cast<const double*>(stage.in(y_res, sizeof(double) * (size_t)P * (size_t)n, &e));
  src.ptr[2] = sqrt_w ? static_cast<const double*>(stage.in(sw.data(), sizeof(double) * (size_t)n, &e)) : nullptr;
  src.ptr[3] = sqrt_w ? static_cast<const double*>(stage.in(w.data(), sizeof(double) * (size_t)n, &e)) : nullptr;
  src.first[0] = 0;
  src.first[1] = g.Kd;
  src.first[2] = g.Kd + P;
  src.first[3] = g.Kd + P + (sqrt_w ? 1 : 0);
  src.first[4] = g.C;
  const double* v_qty = static_cast<const double*>(stage.in(qty, sizeof(double) * n_qty, &e));
  const double* v_yyp = static_cast<const double*>(stage.in(yyp, sizeof(double) * n_yyp, &e));
  if (c->busy_valid) TRY(cudaStreamWaitEvent(0, c->busy_ev, 0));
  for (size_t i = 0; i < c->spare.size(); ++i) {
    const Group& sp = c->spare[i];
    if (sp.cap_basis >= basis_bytes && sp.cap_mask >= mask_bytes && sp.cap_qty >= qty_bytes && sp.cap_yyp >= yyp_bytes &&
        sp.cap_basis <= 2 * basis_bytes + (1u << 20)) {   // (do not pin a huge buffer under a tiny group)
      g.d_basis = sp.d_basis; g.d_mask = sp.d_mask; g.d_qty = sp.d_qty; g.d_yyp = sp.d_yyp;
      g.cap_basis = sp.cap_basis; g.cap_mask = sp.cap_mask; g.cap_qty = sp.cap_qty; g.cap_yyp = sp.cap_yyp;
      c->spare.erase(c->spare.begin() + (long)i);
      pooled = true;
      break;
    }
  }
  if (!pooled) {
    TRY(cudaMalloc(&g.d_basis, basis_bytes));
    TRY(cudaMalloc(&g.d_mask, mask_bytes));
    TRY(cudaMalloc(&g.d_qty, qty_bytes));
    TRY(cudaMalloc(&g.d_yyp, yyp_bytes));
    g.cap_basis = basis_bytes; g.cap_mask = mask_bytes; g.cap_qty = qty_bytes; g.cap_yyp = yyp_bytes;
  }
  TRY(cudaMemsetAsync(g.d_basis, 0, basis_bytes, 0));
  TRY(cudaMemsetAsync(g.d_mask, 0, mask_bytes, 0));
  if (n_qty) copy_doubles_kernel<<<(unsigned)std::min<size_t>((n_qty + 255) / 256, 1024), 256>>>(g.d_qty, v_qty, (int64_t)n_qty);
  copy_doubles_kernel<<<(unsigned)std::min<size_t>((n_yyp + 255) / 256, 1024), 256>>>(g.d_yyp, v_yyp, (int64_t)n_yyp);
  {
    const int64_t total = (int64_t)g.C * n;
    int grid = (int)((total + 255) / 256);
    if (grid > 65535) grid = 65535;
    scatter_basis_kernel<<<grid, 256>>>(src, g.C, n, v_idx, g.ns_pad, g.d_basis, g.d_mask);
    c->launches += 3;
    TRY(cudaGetLastError());
  }
  if (!c->stage_ev) TRY(cudaEventCreateWithFlags(&c->stage_ev, cudaEventDisableTiming));
  TRY(cudaEventRecord(c->stage_ev, 0));
  c->stage_ev_valid = true;
  if (!c->ready_ev) TRY(cudaEventCreateWithFlags(&c->ready_ev, cudaEventDisableTiming));
  TRY(cudaEventRecord(c->ready_ev, 0));
  c->ready_valid = true;
#undef TRY
  c->groups.push_back(g);
  tc_invalidate(c);
  tc4_invalidate(c);
  c->n_samples_total = n_samples_total;
  c->dots_offset.clear();  // workspace layout depends on the group list
  c->pilot_done = false;
  return LRR_OK;
}

int lrr_add_group(lrr_ctx* ctx, int64_t n_samples_total, int32_t n, int32_t K, int32_t P, int32_t has_intercept,
                  const int32_t* complete_idx, const double* q_cols, const double* y_res, const double* qty,
                  const double* yyp) try {
  return add_group_impl(ctx, n_samples_total, n, K, P, has_intercept, complete_idx, q_cols, y_res, qty, yyp, nullptr);
}
LRR_ABI_CATCH(ctx)

int lrr_add_group_weighted(lrr_ctx* ctx, int64_t n_samples_total, int32_t n, int32_t K, int32_t P,
                           const int32_t* complete_idx, const double* q_cols, const double* y_res, const double* qty,
                           const double* yyp, const double* sqrt_w) try {
  if (ctx && !sqrt_w) return fail(reinterpret_cast<Ctx*>(ctx), LRR_EINVAL, "lrr_add_group_weighted: sqrt_w is NULL");
  return add_group_impl(ctx, n_samples_total, n, K, P, 0, complete_idx, q_cols, y_res, qty, yyp, sqrt_w);
}
LRR_ABI_CATCH(ctx)

int lrr_set_score_model(lrr_ctx* ctx, int64_t n_samples_total, int32_t n, int32_t K, const int32_t* complete_idx,
                        const double* wc, const double* resid, const double* w, const double* finv, const double* score0) try {
  if (!ctx) return LRR_EINVAL;
  Ctx* c = reinterpret_cast<Ctx*>(ctx);
  if (K < 1) return fail(c, LRR_EINVAL, "logistic regression requires at least one covariate expression");
  if (!wc || !resid || !w || !finv || !score0) return fail(c, LRR_EINVAL, "lrr_set_score_model: NULL input array");
  if (int r = lrr_clear_groups(ctx)) return r;
  // u = F00^-1 s0 and s0' u: the part of chi2 = s' F^-1 s that does not depend on the variant
  std::vector<double> aux((size_t)K + 1, 0.0), sw((size_t)n);
  for (int i = 0; i < K; ++i)
    for (int j = 0; j < K; ++j) aux[i] += finv[(size_t)i * K + j] * score0[j];
  for (int i = 0; i < K; ++i) aux[K] += score0[i] * aux[i];
  for (int i = 0; i < n; ++i) sw[i] = sqrt(w[i]);
  if (int r = add_group_impl(ctx, n_samples_total, n, K, 1, 0, complete_idx, wc, resid, finv, aux.data(), sw.data(), w,
                             K * K, K + 1))
    return r;
  c->groups.back().score = 1;
  return LRR_OK;
}
LRR_ABI_CATCH(ctx)

int lrr_run_dense(lrr_ctx* ctx, const double* d_x, int64_t n_variants, int64_t ldx, int64_t n_samples_total,
                  const lrr_group_out* outs, int32_t n_outs, void* stream) try {
  CTX_PROLOGUE;
  if (c->groups.empty()) return fail(c, LRR_ESTATE, "lrr_run_dense: no groups (call lrr_add_group)");
  if (c->groups[0].score) return fail(c, LRR_ESTATE, "lrr_run_dense: the context holds a logistic score model");
  if (n_outs != (int32_t)c->groups.size() || !outs) return fail(c, LRR_EINVAL, "lrr_run_dense: need one lrr_group_out per group");
  if (n_variants < 0 || ldx < n_samples_total) return fail(c, LRR_EINVAL, "lrr_run_dense: bad shape (ldx >= n_samples_total)");
  if (n_samples_total != c->n_samples_total) return fail(c, LRR_EINVAL, "lrr_run_dense: n_samples_total differs from the groups'");
  if (n_variants == 0) return LRR_OK;
  if (!d_x) return fail(c, LRR_EINVAL, "lrr_run_dense: d_x is NULL");
  if (int r = ensure_workspace(c, n_variants)) return r;
  if (int r = run_begin(c, st)) return r;
  if (int r = launch_dense_sweep(c, d_x, n_variants, ldx, st)) return r;
  c->last_kernel = LRR_KERNEL_FP64;
  for (size_t g = 0; g < c->groups.size(); ++g)
    if (int r = launch_stats_epilogue(c, (int)g, n_variants, outs[g], st, true)) return r;
  return run_end(c, st);
}
LRR_ABI_CATCH(ctx)

int lrr_run_dense_u16(lrr_ctx* ctx, const uint16_t* d_xq, int64_t n_variants, int64_t ldx, int64_t n_samples_total, double scale,
                      const lrr_group_out* outs, int32_t n_outs, void* stream) try {
  CTX_PROLOGUE;
  if (c->groups.empty()) return fail(c, LRR_ESTATE, "lrr_run_dense_u16: no groups (call lrr_add_group)");
  if (c->groups[0].score) return fail(c, LRR_ESTATE, "lrr_run_dense_u16: the context holds a logistic score model");
  if (n_outs != (int32_t)c->groups.size() || !outs) return fail(c, LRR_EINVAL, "lrr_run_dense_u16: need one lrr_group_out per group");
  if (n_variants < 0 || ldx < n_samples_total) return fail(c, LRR_EINVAL, "lrr_run_dense_u16: bad shape (ldx >= n_samples_total)");
  if (!(scale > 0.0)) return fail(c, LRR_EINVAL, "lrr_run_dense_u16: scale must be positive");
  if (n_samples_total != c->n_samples_total) return fail(c, LRR_EINVAL, "lrr_run_dense_u16: n_samples_total differs from the groups'");
  if (n_variants == 0) return LRR_OK;
  if (!d_xq) return fail(c, LRR_EINVAL, "lrr_run_dense_u16: d_xq is NULL");
  if (int r = ensure_workspace(c, n_variants)) return r;
  if (int r = run_begin(c, st)) return r;
  if (int r = launch_dense_sweep(c, nullptr, n_variants, ldx, st, d_xq, scale)) return r;
  c->last_kernel = LRR_KERNEL_FP64;
  for (size_t g = 0; g < c->groups.size(); ++g)
    if (int r = launch_stats_epilogue(c, (int)g, n_variants, outs[g], st, true)) return r;
  return run_end(c, st);
}
LRR_ABI_CATCH(ctx)

int lrr_run_score(lrr_ctx* ctx, const uint8_t* d_packed, const uint8_t* d_row_flags, int64_t n_variants, int64_t packed_stride,
                  int64_t n_samples_total, const lrr_score_out* out, void* stream) try {
  CTX_PROLOGUE;
  (void)d_row_flags;
  if (c->groups.size() != 1 || !c->groups[0].score) return fail(c, LRR_ESTATE, "lrr_run_score: call lrr_set_score_model first");
  if (!out) return fail(c, LRR_EINVAL, "lrr_run_score: out is NULL");
  if (n_variants < 0) return fail(c, LRR_EINVAL, "lrr_run_score: negative n_variants");
  if (n_samples_total != c->n_samples_total) return fail(c, LRR_EINVAL, "lrr_run_score: n_samples_total differs from the model's");
  if (int r = check_packed(c, n_samples_total, packed_stride)) return r;
  if (packed_stride * 4 != c->groups[0].ns_pad) return fail(c, LRR_EINVAL, "lrr_run_score: packed_stride must equal lrr_packed_stride(n_samples_total)");
  if (n_variants == 0) return LRR_OK;
  if (!d_packed) return fail(c, LRR_EINVAL, "lrr_run_score: d_packed is NULL");
  if (int r = ensure_workspace(c, n_variants)) return r;
  if (int r = run_begin(c, st)) return r;
  if (int r = launch_fp64_sweep(c, d_packed, n_variants, packed_stride, st)) return r;
  c->last_kernel = LRR_KERNEL_FP64;
  if (int r = launch_score_epilogue(c, n_variants, *out, st)) return r;
  return run_end(c, st);
}
LRR_ABI_CATCH(ctx)

int lrr_run_score_dense(lrr_ctx* ctx, const double* d_x, int64_t n_variants, int64_t ldx, int64_t n_samples_total,
                        const lrr_score_out* out, void* stream) try {
  CTX_PROLOGUE;
  if (c->groups.size() != 1 || !c->groups[0].score) return fail(c, LRR_ESTATE, "lrr_run_score_dense: call lrr_set_score_model first");
  if (!out) return fail(c, LRR_EINVAL, "lrr_run_score_dense: out is NULL");
  if (n_variants < 0 || ldx < n_samples_total) return fail(c, LRR_EINVAL, "lrr_run_score_dense: bad shape (ldx >= n_samples_total)");
  if (n_samples_total != c->n_samples_total) return fail(c, LRR_EINVAL, "lrr_run_score_dense: n_samples_total differs from the model's");
  if (n_variants == 0) return LRR_OK;
  if (!d_x) return fail(c, LRR_EINVAL, "lrr_run_score_dense: d_x is NULL");
  if (int r = ensure_workspace(c, n_variants)) return r;
  if (int r = run_begin(c, st)) return r;
  if (int r = launch_dense_sweep(c, d_x, n_variants, ldx, st)) return r;
  c->last_kernel = LRR_KERNEL_FP64;
  if (int r = launch_score_epilogue(c, n_variants, *out, st, true)) return r;
  return run_end(c, st);
}
LRR_ABI_CATCH(ctx)

int lrr_set_logit_model(lrr_ctx* ctx, int64_t n_samples_total, int32_t n, int32_t K, const int32_t* complete_idx,
                        const double* cov, const double* y, const double* b0, const double* score0, const double* fisher0,
                        double loglik0) try {
  if (!ctx) return LRR_EINVAL;
  Ctx* c = reinterpret_cast<Ctx*>(ctx);
  DeviceGuard guard(c->device);
  return logit_set_model(c, n_samples_total, n, K, complete_idx, cov, y, b0, score0, fisher0, loglik0);
}
LRR_ABI_CATCH(ctx)

int lrr_run_logit(lrr_ctx* ctx, const uint8_t* d_packed, int64_t n_variants, int64_t packed_stride, int64_t n_samples_total,
                  int32_t test, int32_t max_iterations, double tolerance, const lrr_logit_out* out, void* stream) try {
  CTX_PROLOGUE;
  if (!out) return fail(c, LRR_EINVAL, "lrr_run_logit: out is NULL");
  if (int r = check_packed(c, n_samples_total, packed_stride)) return r;
  return logit_run(c, d_packed, nullptr, n_variants, packed_stride, n_samples_total, test, max_iterations, tolerance, *out, st);
}
LRR_ABI_CATCH(ctx)

int lrr_run_logit_dense(lrr_ctx* ctx, const double* d_x, int64_t n_variants, int64_t ldx, int64_t n_samples_total, int32_t test,
                        int32_t max_iterations, double tolerance, const lrr_logit_out* out, void* stream) try {
  CTX_PROLOGUE;
  if (!out) return fail(c, LRR_EINVAL, "lrr_run_logit_dense: out is NULL");
  return logit_run(c, nullptr, d_x, n_variants, ldx, n_samples_total, test, max_iterations, tolerance, *out, st);
}
LRR_ABI_CATCH(ctx)

int lrr_reserve(lrr_ctx* ctx, int64_t max_variants) try {
  if (!ctx) return LRR_EINVAL;
  Ctx* c = reinterpret_cast<Ctx*>(ctx);
  DeviceGuard guard(c->device);
  if (c->groups.empty()) return fail(c, LRR_ESTATE, "lrr_reserve: add groups first");
  if (max_variants < 0) return fail(c, LRR_EINVAL, "lrr_reserve: negative size");
  return ensure_workspace(c, max_variants);
}
LRR_ABI_CATCH(ctx)

int lrr_run(lrr_ctx* ctx, const uint8_t* d_packed, const uint8_t* d_row_flags, int64_t n_variants, int64_t packed_stride,
            int64_t n_samples_total, const lrr_group_out* outs, int32_t n_outs, int32_t kernel, void* stream) try {
  CTX_PROLOGUE;
  return run_rows(c, d_packed, d_row_flags, n_variants, packed_stride, n_samples_total, outs, n_outs, kernel, st);
}
LRR_ABI_CATCH(ctx)

int lrr_set_guard(lrr_ctx* ctx, int enabled) {
  if (!ctx) return LRR_EINVAL;
  reinterpret_cast<Ctx*>(ctx)->guard = enabled ? 1 : 0;
  return LRR_OK;
}

int64_t lrr_last_recomputed(lrr_ctx* ctx) {
  if (!ctx) return -1;
  Ctx* c = reinterpret_cast<Ctx*>(ctx);
  DeviceGuard guard(c->device);
  if (c->flag_pending && cudaEventSynchronize(c->flag_ev) != cudaSuccess) return -1;
  int64_t total = 0;
  for (size_t g = 0; g < c->groups.size() && (int)g < c->h_flag_groups; ++g) total += c->h_flag_count[g];
  return total;
}

int64_t lrr_launch_count(const lrr_ctx* ctx) { return ctx ? reinterpret_cast<const Ctx*>(ctx)->launches : 0; }
int lrr_last_kernel(const lrr_ctx* ctx) { return ctx ? reinterpret_cast<const Ctx*>(ctx)->last_kernel : 0; }

int lrr_last_sweep_shape(const lrr_ctx* ctx, int64_t* out4) {
  if (!ctx || !out4) return LRR_EINVAL;
  const Ctx* c = reinterpret_cast<const Ctx*>(ctx);
  for (int i = 0; i < 4; ++i) out4[i] = c->sweep_shape[i];
  return LRR_OK;
}

int lrr_set_timing(lrr_ctx* ctx, int enabled) try {
  if (!ctx) return LRR_EINVAL;
  Ctx* c = reinterpret_cast<Ctx*>(ctx);
  c->timing = enabled ? 1 : 0;
  if (!enabled) c->ev_valid = false;
  return LRR_OK;
}
LRR_ABI_CATCH(ctx)

float lrr_last_sweep_ms(lrr_ctx* ctx) {
  if (!ctx) return -1.f;
  Ctx* c = reinterpret_cast<Ctx*>(ctx);
  if (!c->ev_valid) return -1.f;
  DeviceGuard guard(c->device);
  if (cudaEventSynchronize(c->ev1) != cudaSuccess) return -1.f;
  float ms = -1.f;
  if (cudaEventElapsedTime(&ms, c->ev0, c->ev1) != cudaSuccess) return -1.f;
  return ms;
}

int lrr_student_t_two_sided(lrr_ctx* ctx, const double* d_t, int64_t count, double df, double* d_p, double* d_log10_p,
                            void* stream) try {
  CTX_PROLOGUE;
  if (count < 0 || !(df > 0)) return fail(c, LRR_EINVAL, "lrr_student_t_two_sided: bad arguments");
  return launch_student_t(c, d_t, count, df, d_p, d_log10_p, st);
}
LRR_ABI_CATCH(ctx)

int lrr_qchisqtail1(lrr_ctx* ctx, const double* d_p, int64_t count, double* d_chi2, void* stream) try {
  CTX_PROLOGUE;
  if (count < 0 || (count > 0 && (!d_p || !d_chi2))) return fail(c, LRR_EINVAL, "lrr_qchisqtail1: bad arguments");
  return launch_qchisqtail1(c, d_p, count, d_chi2, st);
}
LRR_ABI_CATCH(ctx)

int lrr_at_times(lrr_ctx* ctx, const uint8_t* d_packed, int64_t n_variants, int64_t packed_stride, int64_t n_samples_total,
                 const double* d_coef, const double* d_t, int32_t L, int32_t n_splits, double* d_out, void* stream) try {
  CTX_PROLOGUE;
  if (n_variants < 0 || n_samples_total <= 0) return fail(c, LRR_EINVAL, "lrr_at_times: bad shape");
  if (int r = check_packed(c, n_samples_total, packed_stride)) return r;
  if (!d_out || (n_variants > 0 && (!d_packed || !d_coef || !d_t))) return fail(c, LRR_EINVAL, "lrr_at_times: NULL array");
  return launch_at_times(c, d_packed, n_variants, packed_stride, n_samples_total, d_coef, d_t, L, n_splits, d_out, st);
}
LRR_ABI_CATCH(ctx)

}  // extern "C"
