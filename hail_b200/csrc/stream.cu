// Host-resident input: the streaming form of the hot loop (include/lrr_b200.h "lrr_stream_*").
//
// The reference consumes a partition's rows as they are decoded from storage (LinearRegression.scala:95 one task per
// partition, io/plink/LoadPlink.scala:470-530 decodes the .bed bytes of each row).  Here the rows are a SNP-major
// .bed body in host memory; blocks of rows flow through four streams
//
//   h2d   cudaMemcpyAsync of a block's .bed bytes into one of N_STAGE staging buffers      (PCIe, the pacer)
//   pack  .bed -> 2-bit store + row flags into one of `depth` device slots                  (HBM speed)
//   comp  sweep + per-variant statistics of the slot into one of N_RES result buffers       (the hot path)
//   d2h   result rows back to the caller's host arrays
//
// chained by events, so that copies in both directions overlap the compute and `depth` blocks can be in flight
// before the first sweep: lrr_stream_begin needs only the genotype bytes and returns at once, the host runs the
// driver prologue (covariate QR) meanwhile, and lrr_stream_run then drains the ring block by block.
#include <string.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <vector>

#include "common.cuh"

namespace lrr {
namespace {

constexpr int N_STAGE = 3;   // .bed staging buffers (h2d -> pack)
constexpr int N_RES = 2;     // device result buffers (comp -> d2h)

struct ResultBuf {
  // per group: int32 n, n_missing [B]; double sum_x [B]; double ytx, beta, se, t, p, log10_p [B, P]
  std::vector<lrr_group_out> outs;
  void* base = nullptr;
  cudaEvent_t ready = nullptr, copied = nullptr;
  bool used = false;
};

}  // namespace

struct Stream {
  Ctx* ctx = nullptr;
  const uint8_t* h_bed = nullptr;
  int64_t M = 0, bed_stride = 0, N = 0, block = 0, n_blocks = 0, stride = 0;
  int depth = 0;
  cudaStream_t s_h2d = nullptr, s_pack = nullptr, s_comp = nullptr, s_d2h = nullptr;
  uint8_t* d_stage[N_STAGE] = {nullptr, nullptr, nullptr};
  cudaEvent_t stage_loaded[N_STAGE] = {}, stage_free[N_STAGE] = {};
  bool stage_used[N_STAGE] = {};
  std::vector<uint8_t*> d_packed, d_flags;
  std::vector<cudaEvent_t> packed, swept;
  std::vector<char> slot_swept_valid;
  ResultBuf res[N_RES];
  int64_t next_load = 0;   // next block to copy + pack
  bool ran = false;
  cudaEvent_t h2d_first = nullptr, h2d_last = nullptr;   // timing events around the first / last block copy (measurement)
  std::vector<cudaEvent_t> t_loaded, t_comp0, t_comp1;     // per block: copy done, sweep started, statistics done (c->timing only)
};

namespace {

int64_t rows_of(const Stream* s, int64_t b) { return std::min(s->block, s->M - b * s->block); }

// copy block b to the device and pack it into its slot (asynchronous)
int issue_load(Stream* s, int64_t b) {
  Ctx* c = s->ctx;
  const int k = (int)(b % std::min<int64_t>(N_STAGE, s->n_blocks)), slot = (int)(b % s->depth);
  const int64_t rows = rows_of(s, b);
  if (s->stage_used[k]) LRR_CUDA(c, cudaStreamWaitEvent(s->s_h2d, s->stage_free[k], 0));
  if (b == 0 && s->h2d_first) LRR_CUDA(c, cudaEventRecord(s->h2d_first, s->s_h2d));
  LRR_CUDA(c, cudaMemcpyAsync(s->d_stage[k], s->h_bed + b * s->block * s->bed_stride, (size_t)(rows * s->bed_stride),
                              cudaMemcpyHostToDevice, s->s_h2d));
  if (b == s->n_blocks - 1 && s->h2d_last) LRR_CUDA(c, cudaEventRecord(s->h2d_last, s->s_h2d));
  if (c->timing && b < 4096) {
    if ((int64_t)s->t_loaded.size() <= b) s->t_loaded.resize((size_t)b + 1, nullptr);
    LRR_CUDA(c, cudaEventCreate(&s->t_loaded[(size_t)b]));
    LRR_CUDA(c, cudaEventRecord(s->t_loaded[(size_t)b], s->s_h2d));
  }
  LRR_CUDA(c, cudaEventRecord(s->stage_loaded[k], s->s_h2d));
  LRR_CUDA(c, cudaStreamWaitEvent(s->s_pack, s->stage_loaded[k], 0));
  if (s->slot_swept_valid[slot]) LRR_CUDA(c, cudaStreamWaitEvent(s->s_pack, s->swept[slot], 0));   // slot still being read
  if (int r = launch_pack_bed(c, s->d_stage[k], rows, s->bed_stride, s->N, s->d_packed[slot], s->stride, s->d_flags[slot],
                              s->s_pack))
    return r;
  LRR_CUDA(c, cudaEventRecord(s->stage_free[k], s->s_pack));
  LRR_CUDA(c, cudaEventRecord(s->packed[slot], s->s_pack));
  s->stage_used[k] = true;
  return LRR_OK;
}

void destroy(Stream* s) {
  if (!s) return;
  DeviceGuard g(s->ctx->device);
  if (s->s_h2d) cudaStreamSynchronize(s->s_h2d);
  if (s->s_pack) cudaStreamSynchronize(s->s_pack);
  if (s->s_comp) cudaStreamSynchronize(s->s_comp);
  if (s->s_d2h) cudaStreamSynchronize(s->s_d2h);
  for (int k = 0; k < N_STAGE; ++k) {
    if (s->stage_loaded[k]) cudaEventDestroy(s->stage_loaded[k]);
    if (s->stage_free[k]) cudaEventDestroy(s->stage_free[k]);
  }
  // the staging / slot memory is the context's arena: it stays allocated for the next stream (lrr_destroy frees it)
  for (auto* v : {&s->t_loaded, &s->t_comp0, &s->t_comp1})
    for (auto e : *v) if (e) cudaEventDestroy(e);
  if (s->h2d_first) cudaEventDestroy(s->h2d_first);
  if (s->h2d_last) cudaEventDestroy(s->h2d_last);
  for (auto e : s->packed) if (e) cudaEventDestroy(e);
  for (auto e : s->swept) if (e) cudaEventDestroy(e);
  for (auto& r : s->res) {
    cudaFree(r.base);
    if (r.ready) cudaEventDestroy(r.ready);
    if (r.copied) cudaEventDestroy(r.copied);
  }
  if (s->s_h2d) cudaStreamDestroy(s->s_h2d);
  if (s->s_pack) cudaStreamDestroy(s->s_pack);
  if (s->s_comp) cudaStreamDestroy(s->s_comp);
  if (s->s_d2h) cudaStreamDestroy(s->s_d2h);
  s->ctx->streams_alive--;
  delete s;
}

// carve one device allocation into the fields of every group for `B` rows
int alloc_results(Stream* s, ResultBuf& r, int64_t B) {
  Ctx* c = s->ctx;
  size_t bytes = 0;
  auto take = [&](size_t n) { size_t off = bytes; bytes += (n + 255) / 256 * 256; return off; };
  std::vector<std::vector<size_t>> offs;
  for (const Group& g : c->groups) {
    std::vector<size_t> o;
    o.push_back(take(sizeof(int32_t) * B));
    o.push_back(take(sizeof(int32_t) * B));
    o.push_back(take(sizeof(double) * B));
    for (int f = 0; f < 6; ++f) o.push_back(take(sizeof(double) * B * g.P));
    offs.push_back(o);
  }
  LRR_CUDA(c, cudaMalloc(&r.base, bytes ? bytes : 256));
  char* b = static_cast<char*>(r.base);
  r.outs.resize(c->groups.size());
  for (size_t g = 0; g < c->groups.size(); ++g) {
    lrr_group_out& o = r.outs[g];
    o.n = reinterpret_cast<int32_t*>(b + offs[g][0]);
    o.n_missing = reinterpret_cast<int32_t*>(b + offs[g][1]);
    o.sum_x = reinterpret_cast<double*>(b + offs[g][2]);
    o.y_transpose_x = reinterpret_cast<double*>(b + offs[g][3]);
    o.beta = reinterpret_cast<double*>(b + offs[g][4]);
    o.standard_error = reinterpret_cast<double*>(b + offs[g][5]);
    o.t_stat = reinterpret_cast<double*>(b + offs[g][6]);
    o.p_value = reinterpret_cast<double*>(b + offs[g][7]);
    o.log10_p = reinterpret_cast<double*>(b + offs[g][8]);
  }
  LRR_CUDA(c, cudaEventCreateWithFlags(&r.ready, cudaEventDisableTiming));
  LRR_CUDA(c, cudaEventCreateWithFlags(&r.copied, cudaEventDisableTiming));
  return LRR_OK;
}

}  // namespace
}  // namespace lrr

using namespace lrr;

extern "C" {

int lrr_stream_begin(lrr_ctx* ctx, lrr_stream** out, const uint8_t* h_bed, int64_t n_variants, int64_t bed_stride,
                     int64_t n_samples, int64_t block_variants, int32_t depth) try {
  if (!ctx) return LRR_EINVAL;
  Ctx* c = reinterpret_cast<Ctx*>(ctx);
  if (!out) return fail(c, LRR_EINVAL, "lrr_stream_begin: out is NULL");
  *out = nullptr;
  if (n_variants < 0 || n_samples <= 0 || bed_stride < (n_samples + 3) / 4)
    return fail(c, LRR_EINVAL, "lrr_stream_begin: bed_stride must be >= ceil(n_samples/4) (LoadPlink.scala:240-251)");
  if (n_variants > 0 && !h_bed) return fail(c, LRR_EINVAL, "lrr_stream_begin: h_bed is NULL");
  if (c->streams_alive) return fail(c, LRR_ESTATE, "lrr_stream_begin: another stream of this context is still open (lrr_stream_end it first)");
  DeviceGuard guard(c->device);
  const bool trace = tuning_env("LRR_TRACE") != nullptr;   // tuning builds only: where lrr_stream_begin spends host time
  auto now_us = [] { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  const double tr0 = trace ? now_us() : 0.0;
  double tr1 = 0, tr2 = 0, tr3 = 0, tr4 = 0;
  Stream* s = new Stream();
  c->streams_alive++;
  s->ctx = c;
  s->h_bed = h_bed;
  s->M = n_variants;
  s->bed_stride = bed_stride;
  s->N = n_samples;
  s->stride = lrr_packed_stride(n_samples);
  // default block: about 256 MB of .bed bytes, a multiple of the 128-variant tile
  if (block_variants <= 0) block_variants = std::max<int64_t>(128, (256ll << 20) / bed_stride / 128 * 128);
  s->block = std::max<int64_t>(1, std::min<int64_t>(block_variants, std::max<int64_t>(n_variants, 1)));
  s->n_blocks = (n_variants + s->block - 1) / s->block;
  // default depth: up to 16 GB of packed rows in flight (covers the host prologue at PCIe rate) but never more than
  // half of what the device has free (the cached arena counts as free: it is reused), at least 3 slots
  const int n_stage = (int)std::min<int64_t>(N_STAGE, s->n_blocks);
  const size_t stage_bytes = ((size_t)(s->block * bed_stride) + 255) / 256 * 256;
  const size_t slot_bytes = (size_t)(s->block * s->stride);
  const size_t flag_bytes = ((size_t)s->block + 255) / 256 * 256;
  auto arena_need = [&](int64_t dep) {
    const int64_t d = std::max<int64_t>(1, std::min<int64_t>(dep, std::max<int64_t>(s->n_blocks, 1)));
    return n_stage * stage_bytes + (size_t)d * (slot_bytes + flag_bytes);
  };
  if (depth <= 0) {
    // The budget is asked from the driver when the arena has to be (re)built and remembered with it: cudaMemGetInfo goes
    // through the kernel driver and took up to 80 ms in one call out of five on a shared box.
    int64_t budget = c->arena ? c->stream_budget : 0;
    if (budget > 0 && arena_need(std::max<int64_t>(3, budget / (s->block * s->stride))) > c->arena_bytes) budget = 0;
    if (budget <= 0) {
      size_t free_b = 0, total_b = 0;
      budget = 16ll << 30;
      if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess)
        budget = std::min<int64_t>(budget, (int64_t)((free_b + c->arena_bytes) / 2));
      else
        cudaGetLastError();
      c->stream_budget = budget;
    }
    depth = (int)std::max<int64_t>(3, budget / (s->block * s->stride));
  }
  s->depth = (int)std::max<int64_t>(1, std::min<int64_t>(depth, std::max<int64_t>(s->n_blocks, 1)));
  if (trace) tr1 = now_us();
  auto bail = [&](int code) {
    destroy(s);
    return code;
  };
#define TRY(call)                                                      \
  do {                                                                 \
    cudaError_t _e = (call);                                           \
    if (_e != cudaSuccess) return bail(cuda_fail(c, _e, #call));       \
  } while (0)
  TRY(cudaStreamCreateWithFlags(&s->s_h2d, cudaStreamNonBlocking));
  TRY(cudaStreamCreateWithFlags(&s->s_pack, cudaStreamNonBlocking));
  TRY(cudaStreamCreateWithFlags(&s->s_comp, cudaStreamNonBlocking));
  TRY(cudaStreamCreateWithFlags(&s->s_d2h, cudaStreamNonBlocking));
  TRY(cudaEventCreate(&s->h2d_first));
  TRY(cudaEventCreate(&s->h2d_last));
  if (trace) tr2 = now_us();
  // one arena for the staging buffers and the slots, cached on the context across streams (cudaMalloc / cudaFree of
  // tens of GB would otherwise sit in front of the first copy of every call)
  const size_t need = arena_need(s->depth);
  if (need > c->arena_bytes) {
    cudaFree(c->arena);
    c->arena = nullptr;
    c->arena_bytes = 0;
    TRY(cudaMalloc(&c->arena, need));
    c->arena_bytes = need;
  }
  uint8_t* a = static_cast<uint8_t*>(c->arena);
  for (int k = 0; k < n_stage; ++k) {
    s->d_stage[k] = a;
    a += stage_bytes;
    TRY(cudaEventCreateWithFlags(&s->stage_loaded[k], cudaEventDisableTiming));
    TRY(cudaEventCreateWithFlags(&s->stage_free[k], cudaEventDisableTiming));
  }
  s->d_packed.assign(s->depth, nullptr);
  s->d_flags.assign(s->depth, nullptr);
  s->packed.assign(s->depth, nullptr);
  s->swept.assign(s->depth, nullptr);
  s->slot_swept_valid.assign(s->depth, 0);
  for (int i = 0; i < s->depth && s->n_blocks > 0; ++i) {
    s->d_packed[i] = a;
    a += slot_bytes;
    TRY(cudaEventCreateWithFlags(&s->packed[i], cudaEventDisableTiming));
    TRY(cudaEventCreateWithFlags(&s->swept[i], cudaEventDisableTiming));
  }
  for (int i = 0; i < s->depth && s->n_blocks > 0; ++i) {
    s->d_flags[i] = a;
    a += flag_bytes;
  }
#undef TRY
  if (trace) tr3 = now_us();
  for (int64_t b = 0; b < s->n_blocks && b < s->depth; ++b) {
    if (int r = issue_load(s, b)) return bail(r);
    s->next_load = b + 1;
  }
  if (trace) {
    tr4 = now_us();
    fprintf(stderr, "[lrr trace] stream_begin: mem info %.0f us, streams %.0f us, arena + events %.0f us, first loads %.0f us\n",
            tr1 - tr0, tr2 - tr1, tr3 - tr2, tr4 - tr3);
  }
  *out = reinterpret_cast<lrr_stream*>(s);
  return LRR_OK;
}
LRR_ABI_CATCH(ctx)

int lrr_stream_run(lrr_ctx* ctx, lrr_stream* stream, const lrr_group_out* h_outs, int32_t n_outs, int32_t kernel) try {
  if (!ctx || !stream) return LRR_EINVAL;
  Ctx* c = reinterpret_cast<Ctx*>(ctx);
  Stream* s = reinterpret_cast<Stream*>(stream);
  if (s->ctx != c) return fail(c, LRR_EINVAL, "lrr_stream_run: stream belongs to another context");
  if (s->ran) return fail(c, LRR_ESTATE, "lrr_stream_run: a stream runs once");
  if (c->groups.empty()) return fail(c, LRR_ESTATE, "lrr_stream_run: no groups (call lrr_add_group)");
  if (n_outs != (int32_t)c->groups.size() || !h_outs) return fail(c, LRR_EINVAL, "lrr_stream_run: need one lrr_group_out per group");
  if (s->N != c->n_samples_total) return fail(c, LRR_EINVAL, "lrr_stream_run: n_samples differs from the groups'");
  DeviceGuard guard(c->device);
  s->ran = true;
  if (s->n_blocks == 0) return LRR_OK;
  if (int r = lrr_reserve(ctx, s->block)) return r;
  for (auto& rb : s->res)
    if (int r = alloc_results(s, rb, s->block)) return r;
  const int timing = c->timing;
  c->timing = 0;   // the sweep timer brackets one lrr_run; not meaningful across overlapped blocks
  int rc = LRR_OK;
  for (int64_t b = 0; b < s->n_blocks && rc == LRR_OK; ++b) {
    const int slot = (int)(b % s->depth);
    ResultBuf& rb = s->res[b % N_RES];
    const int64_t rows = rows_of(s, b), row0 = b * s->block;
    cudaError_t e;
    if ((e = cudaStreamWaitEvent(s->s_comp, s->packed[slot], 0)) != cudaSuccess) { rc = cuda_fail(c, e, "wait packed"); break; }
    if (rb.used && (e = cudaStreamWaitEvent(s->s_comp, rb.copied, 0)) != cudaSuccess) { rc = cuda_fail(c, e, "wait copied"); break; }
    if (timing && b < 4096) {
      cudaEvent_t e0 = nullptr, e1 = nullptr;
      if (cudaEventCreate(&e0) == cudaSuccess && cudaEventCreate(&e1) == cudaSuccess) {
        s->t_comp0.push_back(e0);
        s->t_comp1.push_back(e1);
        cudaEventRecord(e0, s->s_comp);
      }
    }
    rc = run_rows(c, s->d_packed[slot], s->d_flags[slot], rows, s->stride, s->N, rb.outs.data(), n_outs, kernel, s->s_comp);
    if (rc != LRR_OK) break;
    if (timing && b < 4096 && (int64_t)s->t_comp1.size() == b + 1) cudaEventRecord(s->t_comp1[(size_t)b], s->s_comp);
#define STEP(call) if ((e = (call)) != cudaSuccess) { rc = cuda_fail(c, e, #call); break; }
    STEP(cudaEventRecord(s->swept[slot], s->s_comp));
    s->slot_swept_valid[slot] = 1;
    STEP(cudaEventRecord(rb.ready, s->s_comp));
    STEP(cudaStreamWaitEvent(s->s_d2h, rb.ready, 0));
    for (size_t g = 0; g < c->groups.size(); ++g) {
      const int P = c->groups[g].P;
      const lrr_group_out& d = rb.outs[g];
      const lrr_group_out& h = h_outs[g];
      auto copy = [&](void* dst, const void* src, size_t elem, int64_t per_row) {
        if (!dst || e != cudaSuccess) return;   // the first failed enqueue is kept in `e` and reported below
        e = cudaMemcpyAsync(static_cast<char*>(dst) + (size_t)row0 * per_row * elem, src, (size_t)rows * per_row * elem,
                            cudaMemcpyDeviceToHost, s->s_d2h);
      };
      copy(h.n, d.n, sizeof(int32_t), 1);
      copy(h.n_missing, d.n_missing, sizeof(int32_t), 1);
      copy(h.sum_x, d.sum_x, sizeof(double), 1);
      copy(h.y_transpose_x, d.y_transpose_x, sizeof(double), P);
      copy(h.beta, d.beta, sizeof(double), P);
      copy(h.standard_error, d.standard_error, sizeof(double), P);
      copy(h.t_stat, d.t_stat, sizeof(double), P);
      copy(h.p_value, d.p_value, sizeof(double), P);
      copy(h.log10_p, d.log10_p, sizeof(double), P);
    }
    if (e != cudaSuccess) { rc = cuda_fail(c, e, "cudaMemcpyAsync(result rows, device -> host)"); break; }
    STEP(cudaEventRecord(rb.copied, s->s_d2h));
#undef STEP
    rb.used = true;
    if (s->next_load < s->n_blocks) {
      rc = issue_load(s, s->next_load);
      s->next_load++;
    }
  }
  c->timing = timing;
  cudaError_t e = cudaStreamSynchronize(s->s_d2h);
  cudaError_t e2 = cudaStreamSynchronize(s->s_comp);
  c->stream_timeline.clear();
  if (timing && rc == LRR_OK && s->h2d_first) {   // per block, milliseconds since the first copy started: copy done, sweep started, statistics done
    cudaStreamSynchronize(s->s_comp);
    for (size_t b = 0; b < s->t_comp1.size() && b < s->t_loaded.size(); ++b) {
      float t[3] = {-1.f, -1.f, -1.f};
      cudaEventElapsedTime(&t[0], s->h2d_first, s->t_loaded[b]);
      cudaEventElapsedTime(&t[1], s->h2d_first, s->t_comp0[b]);
      cudaEventElapsedTime(&t[2], s->h2d_first, s->t_comp1[b]);
      c->stream_timeline.insert(c->stream_timeline.end(), t, t + 3);
    }
    cudaGetLastError();
  }
  c->last_stream_h2d_ms = -1.f;
  if (rc == LRR_OK && s->h2d_first && s->h2d_last && cudaEventSynchronize(s->h2d_last) == cudaSuccess) {
    float ms = -1.f;
    if (cudaEventElapsedTime(&ms, s->h2d_first, s->h2d_last) == cudaSuccess) c->last_stream_h2d_ms = ms;
    else cudaGetLastError();
  }
  if (rc == LRR_OK && e != cudaSuccess) rc = cuda_fail(c, e, "cudaStreamSynchronize(d2h)");
  if (rc == LRR_OK && e2 != cudaSuccess) rc = cuda_fail(c, e2, "cudaStreamSynchronize(comp)");
  return rc;
}
LRR_ABI_CATCH(ctx)

int lrr_last_stream_timeline(const lrr_ctx* ctx, float* out, int capacity) {
  if (!ctx) return 0;
  const Ctx* c = reinterpret_cast<const Ctx*>(ctx);
  const int n = (int)std::min<size_t>(c->stream_timeline.size(), capacity > 0 ? (size_t)capacity : 0);
  if (out) memcpy(out, c->stream_timeline.data(), sizeof(float) * (size_t)n);
  return (int)c->stream_timeline.size();
}

float lrr_last_stream_h2d_ms(const lrr_ctx* ctx) { return ctx ? reinterpret_cast<const Ctx*>(ctx)->last_stream_h2d_ms : -1.f; }

int lrr_trim(lrr_ctx* ctx) try {
  if (!ctx) return LRR_EINVAL;
  Ctx* c = reinterpret_cast<Ctx*>(ctx);
  if (c->streams_alive) return fail(c, LRR_ESTATE, "lrr_trim: a stream of this context is still open");
  DeviceGuard guard(c->device);
  LRR_CUDA(c, cudaDeviceSynchronize());
  cudaFree(c->arena);
  c->arena = nullptr;
  c->arena_bytes = 0;
  c->stream_budget = 0;
  cudaFree(c->d_nanmask);
  c->d_nanmask = nullptr;
  c->nanmask_bytes = 0;
  release_caches(c);   // retired groups' buffers, workspaces of the hot call, staging (abi.cu)
  return LRR_OK;
}
LRR_ABI_CATCH(ctx)

void lrr_stream_end(lrr_ctx* ctx, lrr_stream* stream) try {
  (void)ctx;
  destroy(reinterpret_cast<Stream*>(stream));
} catch (...) {
}

}  // extern "C"
