// Shared definitions for the lrr_b200 library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include <new>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/lrr_b200.h"

// Tuning switches read from the environment exist only in tuning builds (scratch/build_abl.sh passes -DLRR_TUNING=1):
// the shipped library never consults the environment, so no variable can change its results.
#ifndef LRR_TUNING
#define LRR_TUNING 0
#endif

namespace lrr {

inline const char* tuning_env(const char* name) {
#if LRR_TUNING
  return getenv(name);
#else
  (void)name;
  return nullptr;
#endif
}

constexpr int kSamplesPerWord = 16;    // 2 bits per call, 32-bit words
constexpr int kRowAlignBytes = 128;    // packed row stride granularity (TMA boxes, 128-bit loads)
constexpr int kMaxGroupCols = 4096;    // sanity bound on K + P per group

// One group of phenotypes sharing a complete-sample set (LinearRegression.scala:228-255 ChainedLinregInput).
struct Group {
  int n = 0;              // complete samples
  int K = 0;              // covariates
  int P = 0;              // phenotypes
  int has_intercept = 0;  // constant column handled exactly from integer counts
  int Kd = 0;             // dot-product covariate columns = K - has_intercept
  int C = 0;              // dot-product columns = Kd + P (weighted groups: + 2, see `weighted`)
  int score = 0;          // logistic score-test model (lrr_set_score_model): columns [w c_k (K) | y - mu | sqrt(w) | w^sq],
                          // d_qty = F00^-1 [K, K], d_yyp = [u = F00^-1 s0 (K) | s0' u]
  int weighted = 0;       // WLS group (statgen.py:557-581): columns are sqrt(w)-scaled, column C-2 = sqrt(w) (its dot is
                          // sum_x of the scaled x), column C-1 = w accumulated against x^2 (the scaled x.x)
  int d = 0;              // degrees of freedom n - K - 1 (LR:50)
  double lbeta = 0.0;     // log B(d/2, 1/2) for the Student-t epilogue
  int64_t ns_pad = 0;     // padded sample count (= 4 * packed stride)
  double* d_basis = nullptr;  // [C][ns_pad] column planes; zero rows for excluded / padding samples
  double* d_qty = nullptr;    // [K][P]
  double* d_yyp = nullptr;    // [P]
  uint32_t* d_mask = nullptr; // [ns_pad/16] bit (8i+2s) set iff sample 16w+4s+i is in the group
  // tensor-core path (filled lazily by tc_prepare_group)
  int8_t* d_bq = nullptr;     // [ncols_pad][ns_pad] int8 digit planes, K-major rows
  double* d_colscale = nullptr;  // [C] value of one unit of the lowest digit
  int ncols_pad = 0;
  int n_slices = 0;
  // dense-dosage path (filled lazily by launch_dense_sweep)
  double* d_basis_t = nullptr;   // [ns_pad][C] sample-major copy of d_basis, then [ns_pad] the 0 / 1 group indicator
  // capacities (bytes) of d_basis / d_mask / d_qty / d_yyp: a retired group's buffers are reused by the next lrr_add_group
  size_t cap_basis = 0, cap_mask = 0, cap_qty = 0, cap_yyp = 0;
};

struct Ctx {
  int device = 0;
  std::string err;
  std::vector<Group> groups;
  int64_t n_samples_total = 0;
  // workspaces
  int64_t reserved_variants = 0;
  int32_t* d_counts = nullptr;  // [G][M][4] n1, n2, nmiss, pad
  double* d_dots = nullptr;     // [sum_g C_g][M]... laid out per group: [M][C_g]
  std::vector<int64_t> dots_offset;  // per group offset (in doubles) into d_dots
  int64_t launches = 0;
  int last_kernel = 0;
  int sm_count = 148;
  void* tc_state = nullptr;  // opaque, owned by tc_kernel.cu
  void* tc4_state = nullptr; // opaque, owned by tc4_kernel.cu
  void* logit_state = nullptr; // opaque, owned by logit_kernel.cu (Wald / LRT / Firth model)
  // optional device timing of the sweep kernel(s) of the last lrr_run (bench roofline)
  int timing = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  bool ev_valid = false;
  // shape of the last lrr_run's tensor-core sweep(s), for the bench's tensor roofline: launches, MMA columns summed over
  // the launches (N of every pass, padded), the part of them issued by launches whose tiles may also run the
  // missing-indicator plane (narrow two-plane sweeps), and the digit columns actually used (unpadded)
  int64_t sweep_shape[4] = {0, 0, 0, 0};
  // device arena of the streaming loop (staging + slots), kept between streams (stream.cu)
  void* arena = nullptr;
  size_t arena_bytes = 0;
  int streams_alive = 0;
  int64_t stream_budget = 0;   // device bytes the default stream depth may take, asked from the driver when the arena is (re)built
  std::vector<float> stream_timeline;   // lrr_set_timing(1): per block of the last lrr_stream_run, ms since the first copy
                                        // started: [copy done, sweep started, statistics done]
  float last_stream_h2d_ms = -1.f;   // first block copy issued -> last block copy done, of the last lrr_stream_run
  // dense-dosage path: one bit per (variant, sample) "missing and in the group" (dense_kernel.cu), grow-only
  void* d_nanmask = nullptr;
  size_t nanmask_bytes = 0;
  // tolerance guard of the quantised (tensor-core) sweeps: rows whose rigorous error bound leaves the tolerance are
  // listed by the statistics epilogue and recomputed in float64 (stats_device.cuh, fp64_kernel.cu)
  int32_t* d_flag_mark = nullptr;   // [G][reserved] 1 = row already listed
  int32_t* d_flag_list = nullptr;   // [G][reserved] listed rows
  int32_t* d_flag_count = nullptr;  // [G]
  int32_t* h_flag_count = nullptr;  // page-locked mirror of d_flag_count after the last run (see lrr_last_recomputed)
  int h_flag_groups = 0;
  cudaEvent_t flag_ev = nullptr;    // the mirror is valid once this has completed
  bool flag_pending = false;
  int64_t flag_rows = 0;            // rows of the run the mirror belongs to
  // per-call latency: nothing on the lrr_clear_groups / lrr_add_group / lrr_run path synchronises the device or calls
  // cudaMalloc in the steady state.  Retired groups keep their buffers for the next lrr_add_group, the workspaces are
  // grow-only, and ordering against kernels still in flight on the caller's streams is done with two events.
  std::vector<Group> spare;         // retired groups (buffers only)
  void* h_stage = nullptr;          // add_group staging: page-locked host buffer mapped into the device (see abi.cu Staging)
  void* d_stage_view = nullptr;     // its device address
  size_t h_stage_bytes = 0;
  cudaEvent_t stage_ev = nullptr;   // the kernels reading the staging buffer have finished
  bool stage_ev_valid = false;
  cudaEvent_t busy_ev = nullptr;    // recorded on the caller's stream after the last run that read the group buffers
  cudaEvent_t ready_ev = nullptr;   // recorded on the default stream after the last lrr_add_group's device work
  bool busy_valid = false, ready_valid = false;
  int64_t ws_groups = 0, ws_cols = 0;   // workspace capacities next to reserved_variants
  void* d_tail = nullptr;           // deferred tail p-values of the statistics epilogue: (entry, t) pairs (grow-only)
  int32_t* d_tail_count = nullptr;
  int64_t tail_capacity = 0;
  void* d_recompute = nullptr;      // split-row partial sums + arrival counters of fp64_recompute_kernel (grow-only)
  size_t recompute_bytes = 0;
  int guard = 1;                    // 0: no tolerance guard (kernel tuning / tests of the raw quantised path)
  int digit_boost = 0;              // extra base-13 digits for covariate / fitted columns (raised by the pilot of run_rows)
  bool pilot_done = false;          // the precision pilot ran for the current group set
};

// make `dev` current for the lifetime of the guard
struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev);
  }
  ~DeviceGuard() {
    int cur = -1;
    cudaGetDevice(&cur);
    if (prev >= 0 && cur != prev) cudaSetDevice(prev);
  }
};

// order the caller's stream behind the last lrr_add_group (call at the start of every run that reads group buffers) and
// mark the group buffers as in use by that stream (call at its end)
void release_caches(Ctx* c);   // lrr_trim: everything the context keeps between calls that no live group needs
int run_begin(Ctx* c, cudaStream_t st);
int run_end(Ctx* c, cudaStream_t st);
int fail(Ctx* c, int code, const std::string& msg);
int cuda_fail(Ctx* c, cudaError_t e, const char* what);
// No C++ exception crosses the C ABI: every extern "C" entry point is a function-try-block ending in
// LRR_ABI_CATCH, which turns std::bad_alloc into LRR_ENOMEM and anything else into LRR_ESTATE with the message kept.
int abi_caught(void* ctx, int code, const char* what) noexcept;
#define LRR_ABI_CATCH(ctx)                                                                              \
  catch (const std::bad_alloc&) { return lrr::abi_caught((void*)(ctx), LRR_ENOMEM, "out of host memory"); } \
  catch (const std::exception& e) { return lrr::abi_caught((void*)(ctx), LRR_ESTATE, e.what()); }          \
  catch (...) { return lrr::abi_caught((void*)(ctx), LRR_ESTATE, "unknown C++ exception"); }

#define LRR_CUDA(ctx, call)                                  \
  do {                                                       \
    cudaError_t _e = (call);                                 \
    if (_e != cudaSuccess) return cuda_fail(ctx, _e, #call); \
  } while (0)

// kernels / stages implemented in the other translation units
int launch_pack_bed(Ctx*, const uint8_t*, int64_t, int64_t, int64_t, uint8_t*, int64_t, uint8_t*, cudaStream_t);
int launch_pack_i8(Ctx*, const int8_t*, int64_t, int64_t, uint8_t*, int64_t, uint8_t*, cudaStream_t);
int launch_unpack_i8(Ctx*, const uint8_t*, int64_t, int64_t, int64_t, int8_t*, cudaStream_t);
int launch_unpack_bed(Ctx*, const uint8_t*, int64_t, int64_t, int64_t, uint8_t*, int64_t, cudaStream_t);
int launch_bn_fill(Ctx*, const uint32_t*, int, const uint8_t*, int64_t, int64_t, int64_t, uint64_t, uint8_t*, int64_t,
                   uint8_t*, cudaStream_t);
int launch_fp64_sweep(Ctx*, const uint8_t* d_packed, int64_t M, int64_t stride, cudaStream_t);
// d_xq != NULL: compact uint16 entries (value = q * xscale, 0xFFFF = missing) instead of the float64 d_x
int launch_dense_sweep(Ctx*, const double* d_x, int64_t M, int64_t ldx, cudaStream_t, const uint16_t* d_xq = nullptr,
                       double xscale = 0.0);
int launch_tc_sweep(Ctx*, const uint8_t* d_packed, const uint8_t* d_row_flags, int64_t M, int64_t stride, cudaStream_t);
bool tc_supported(Ctx*, bool may_have_missing);
const double* tc_quantum(Ctx*, int g);
const double* tc4_quantum(Ctx*, int g, int* n_fit);
const double* tc4_errsum(Ctx*, int g);
int launch_fp64_recompute(Ctx*, int g, const uint8_t* d_packed, int64_t stride, const int32_t* d_list, const int32_t* d_count,
                          int dots_stride, cudaStream_t);
void tc_invalidate(Ctx*);
void tc_release(Ctx*);
int launch_tc4_sweep(Ctx*, const uint8_t* d_packed, const uint8_t* d_row_flags, int64_t M, int64_t stride, cudaStream_t);
// (d_row_flags, M, st: lets the first quantisation of a group set pick the wide one-plane plan when no row has a missing call)
bool tc4_supported(Ctx*, bool single_pass_only, const uint8_t* d_row_flags = nullptr, int64_t M = 0, cudaStream_t st = nullptr);
void tc4_invalidate(Ctx*);
void tc4_release(Ctx*);
void tc4_trim(Ctx*);
// `quantum` != NULL: per-column quantisation step of the sweep that produced the dots (tolerance guard on); `n_fit`:
// fitted-value dot products behind the C dot columns; `stride`: doubles per dots row; `err_sum`: per-column totals of the
// basis rounding errors (4-bit sweep), added to the dot products times the row's centring constant
int launch_stats_epilogue(Ctx*, int g, int64_t M, const lrr_group_out& out, cudaStream_t, bool dense = false,
                          const double* quantum = nullptr, int n_fit = 0, int stride = 0, double qscale = 1.0,
                          const double* err_sum = nullptr);
// the same statistics for the rows listed in d_flag_list (after launch_fp64_recompute)
int launch_stats_epilogue_listed(Ctx*, int g, const lrr_group_out& out, int stride, cudaStream_t);
int run_rows(Ctx*, const uint8_t* d_packed, const uint8_t* d_row_flags, int64_t M, int64_t stride, int64_t n_samples_total,
             const lrr_group_out* outs, int32_t n_outs, int32_t kernel, cudaStream_t);
int launch_student_t(Ctx*, const double*, int64_t, double, double*, double*, cudaStream_t);
int launch_qchisqtail1(Ctx*, const double* d_p, int64_t count, double* d_out, cudaStream_t);
int launch_score_epilogue(Ctx*, int64_t M, const lrr_score_out& out, cudaStream_t, bool dense = false);
int logit_set_model(Ctx*, int64_t n_samples_total, int32_t n, int32_t K, const int32_t* idx, const double* cov, const double* y,
                    const double* b0, const double* score0, const double* fisher0, double loglk0);
int logit_run(Ctx*, const uint8_t* d_packed, const double* d_dense, int64_t M, int64_t stride, int64_t n_samples_total,
              int test, int max_iter, double tol, const lrr_logit_out& out, cudaStream_t);
void logit_release(Ctx*);
int launch_at_times(Ctx*, const uint8_t* d_packed, int64_t M, int64_t stride, int64_t n_total, const double* d_coef,
                    const double* d_t, int L, int n_splits, double* d_out, cudaStream_t);

// position of sample `j` (0..15 within its word) in the packed word: bits [8i+2s, 8i+2s+1], j = 4s+i
__host__ __device__ inline int sample_shift(int j) { return 8 * (j & 3) + 2 * (j >> 2); }

// log B(a, 1/2) stable for large a (host)
double log_beta_half(double a);

}  // namespace lrr
