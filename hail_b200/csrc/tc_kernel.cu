// Kernel 2: the throughput sweep on the 5th-generation tensor cores (tcgen05, sm_100a).
//
// Formulation.  For a tile of 128 variants the projections  D[v, c] = sum_j x[v, j] * B[j, c]  against the
// basis [Q' | Y_res] (LinearRegression.scala:139-146 restated: qtx = Qt * X, xyp = y_res^T X) are one skinny
// GEMM whose A operand is the genotype tile and whose K dimension is the sample axis.  A is EXACT in 8 bits
// (call codes 0..3), so only B needs precision: every basis column is stored as S balanced base-256 digits
// (int8 planes, fixed point relative to the column's max), the MMA accumulates int8 x uint8 products in INT32
// -- exactly, no rounding and no dependence on summation order -- and the per-variant epilogue recombines the
// S exact integers in float64.  With S = 6 the only error is the 2^-47 quantisation of B.
//
// Dataflow per CTA (persistent over variant tiles; CTA pairs with tcgen05.mma.cta_group::2, see the kernel's comment):
//   TMA warps     cp.async.bulk.tensor: packed genotype tile [128 variants x 512 samples] (16 KB, 128B swizzle) into
//                 a deep ring; the matching basis panels [this CTA's half of ncols x 512 samples] int8 into a 6-stage ring
//   unpack warps  (16) LDS.128 of the thread's own variant row, 2-bit -> uint8 with shift/mask, exact popcount,
//                 tcgen05.st of the uint8 row into a TMEM ring (the A operand lives in TMEM)
//   MMA warp      one elected lane issues tcgen05.mma.kind::i8 (M=128 or 256, N=ncols, K=32) with A from TMEM, B from
//                 the swizzled shared-memory panels, D (int32) in TMEM; tcgen05.commit releases ring slots / stages
//   epilogue      (unpack warps 0-3) tcgen05.ld the int32 accumulators, recombine digits, mean-impute correction
//                 from the missing-indicator plane, write counts + float64 dot products for the stats epilogue
// This kernel is the multi-pass path (chained groups, many phenotypes); tc4_kernel.cu is the single-pass 4-bit variant.
//
// Missing calls (RegressionUtils.scala:16-58 mean imputation): code 3.  Plane "c" carries the raw code
// (0..3), plane "m" the indicator [code == 3]; sum_j B (x0 + mean * m) = (Dc - 3 Dm) + mean * Dm.  Tiles whose
// rows carry no missing call (row_flags from ingest) skip plane m entirely.
#include <cuda.h>
#include <algorithm>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "tc_common.cuh"
#include "tc_ptx.cuh"

namespace lrr {
namespace tc {

constexpr int TILE_M = 128;            // variants per tile == TMEM lanes
constexpr int CHUNK = 512;             // samples per shared-memory stage (128 packed bytes per row)
constexpr int SLOT = 128;              // samples per TMEM A slot (32 columns of 4 x uint8)
constexpr int SLOTS = CHUNK / SLOT;    // 4
constexpr int UNPACK_WARPS = 16;       // warp w: TMEM lane quarter w & 3, slot (w >> 2) of every chunk
constexpr int WARP_TMA_G = 16, WARP_TMA_B = 17, WARP_MMA = 18;
constexpr int THREADS = 19 * 32;
// timing-ablation switches (kernel tuning only; results are WRONG when any is non-zero)
#ifndef LRR_ABL_NO_MMA
#define LRR_ABL_NO_MMA 0
#endif
#ifndef LRR_ABL_NO_GENO
#define LRR_ABL_NO_GENO 0    // skip the genotype TMA (barriers still cycle)
#endif
#ifndef LRR_ABL_MMA_J
#define LRR_ABL_MMA_J 4      // MMAs issued per slot (4 = all)
#endif
#ifndef LRR_ABL_MMA_N
#define LRR_ABL_MMA_N 0      // override the MMA N dimension (0 = ncols)
#endif
#ifndef LRR_ABL_MMA_NOWAIT
#define LRR_ABL_MMA_NOWAIT 0
#endif
#ifndef LRR_ABL_NO_STTM
#define LRR_ABL_NO_STTM 0
#endif
#ifndef LRR_ABL_NO_UNPACK
#define LRR_ABL_NO_UNPACK 0
#endif
#ifndef LRR_ABL_HALF_B
#define LRR_ABL_HALF_B 0     // 1: fetch only half of every basis-panel stage (pair mode)
#endif
#ifndef LRR_ABL_NO_BPANEL
#define LRR_ABL_NO_BPANEL 0
#endif
#ifndef LRR_TC_SHIFT_ON_FMA
#define LRR_TC_SHIFT_ON_FMA 0           // 1: right shifts as mul.hi (IMAD.HI, FMA pipe) instead of SHF -- measured slower
#endif
constexpr int GENO_BYTES = TILE_M * 128;   // 16 KB
constexpr int MAX_GROUPS = 4;
constexpr int MAX_GSTAGES = 10;        // genotype ring (16 KB per stage)
#ifndef LRR_TC_NB
#define LRR_TC_NB 6                     // preferred depth of the basis-panel ring (6 or 3: the depths the MMA fast path is unrolled for)
#endif
constexpr int MAX_BSTAGES = 6;         // basis-panel ring (4 * ncols * 128 B per stage)
constexpr int MAX_RING = 4;            // A ring groups (GROUP_COLS TMEM columns each)
constexpr int GROUP_COLS = 128;        // one-plane: 4 slots of plane c; two-plane: 2 slots of plane c + 2 of plane m
// Balanced base-256 digits per basis column: 48 bits for every column (error 2^-47 of the column max, 2^-45 for the
// samples whose field reaches the MMA as 4c -- the quantum the tolerance guard of stats_device.cuh is given is
// therefore 4 colscale).  The covariate columns carry y_transpose_x = xyp + Qty . qtx, whose RIGOROUS bound needs the
// same class of precision as the phenotype columns (32-bit covariate digits left 1 % of the rows outside it).
constexpr int N_SLICES_Y = 6;
constexpr int N_SLICES_Q = 6;
// TMEM column map (512 columns x 128 lanes x 32 bit), ncols = digit columns padded to 16:
//   [0, ncols)            accumulators of plane c (raw call code)
//   [ncols, 2 ncols)      accumulators of plane m (missing indicator), two-plane tiles only
//   [ring_base, 512)      A operand ring of `ring_groups` groups x 128 columns; ring_base = 2 ncols rounded up
//                         to 32.  A group is the unit handed to the MMA warp: the 4 slots (512 samples) of a
//                         chunk for one-plane tiles, or 2 slots of plane c + 2 slots of plane m (256 samples)
//                         for two-plane tiles

struct GroupMeta {
  int col_off;        // first digit column of this group in B
  int C;              // dot-product columns (Kd + P)
  int Kd;             // of which covariate columns (N_SLICES_Q digits each; the rest have N_SLICES_Y)
  int n;              // complete samples
  int mask_all;       // every stored sample is in the group: no masking needed for the hom-alt count
  int32_t* counts;    // [M][4]
  double* dots;       // [M][dots_stride], already offset to this segment's first dot column
  int dots_stride;    // the group's total number of dot columns
  const double* colscale;   // [C]
  const uint32_t* mask_hi;  // [ns_pad/16], high bit of each kept field
};

struct Params {
  int64_t M;
  int n_tiles;
  int n_chunks;
  int ncols;          // padded to 16
  int n_gstages, n_bstages;
  int n_groups;
  int ring_base;      // first TMEM column of the A ring
  int ring_groups;    // groups of GROUP_COLS columns in the A ring (2..MAX_RING)
  int gstage_bytes, bstage_bytes;
  int mask_bytes;     // n_groups * 128 when any group needs masking, else 0
  const uint8_t* row_flags;  // nullable
  GroupMeta g[MAX_GROUPS];
};

// PTX wrappers (mbarrier, TMA, tcgen05, cluster): tc_ptx.cuh, shared with tc4_kernel.cu
using namespace ptx;
using tcc::IntTag;
using tcc::TrueTag;
using tcc::FalseTag;
using tcc::EncodeTiledFn;

// w >> k for k in {2, 4, 6}
template <int K>
__device__ __forceinline__ uint32_t shr(uint32_t w) {
#if LRR_TC_SHIFT_ON_FMA
  uint32_t r;
  asm("mul.hi.u32 %0, %1, %2;" : "=r"(r) : "r"(w), "r"(1u << (32 - K)));
  return r;
#else
  return w >> K;
#endif
}

// instruction descriptor: D = s32, A = u8 (K-major, from TMEM), B = s8 (K-major), M = 128, N = ncols
__device__ __forceinline__ uint32_t make_idesc(int n, int m) {
  uint32_t d = 0;
  d |= 2u << 4;                   // c_format = S32
  d |= 0u << 7;                   // a_format = unsigned 8-bit
  d |= 1u << 10;                  // b_format = signed 8-bit
  d |= (uint32_t)(n >> 3) << 17;  // N
  d |= (uint32_t)(m >> 4) << 24;       // M (256 for a CTA pair)
  return d;
}


struct Barriers {
  uint64_t gfull[MAX_GSTAGES];   // genotype stage filled by TMA
  uint64_t gempty[MAX_GSTAGES];  // genotype stage read out by the 16 unpack warps
  uint64_t bfull[MAX_BSTAGES];   // basis-panel stage filled by TMA (both CTAs of a pair complete on the leader's)
  uint64_t bempty[MAX_BSTAGES];  // basis-panel stage consumed (MMA commit of every CTA in the cluster)
  uint64_t a_full[MAX_RING];     // A ring slot written (4 quarter-warps)
  uint64_t a_empty[MAX_RING];    // A ring slot consumed (MMA commit)
  uint64_t d_full;               // accumulators complete (MMA commit)
  uint64_t d_empty;              // accumulators read out (4 epilogue warps)
  uint32_t tmem_base;
  uint32_t pad;
  // hom-alt popcounts of slots 1..3, handed to the epilogue warps; double-buffered by tile parity because the
  // producing warps may finish the whole next tile (when it has <= ring-depth chunks) during this tile's epilogue
  int32_t n2_xchg[2][SLOTS - 1][MAX_GROUPS][TILE_M];
};

__device__ __forceinline__ bool tile_has_missing(const Params& p, int tile) {
  return tcc::tile_flags_any(p.row_flags, p.M, tile, TILE_M);
}

// NG = number of groups known at compile time (1, 2) or 0 = run-time p.n_groups.
// CS = 1: one CTA per 128-variant tile, tcgen05.mma.cta_group::1.
// CS = 2: a CTA pair (cluster of 2, two SMs of one TPC) sweeps two consecutive tiles as ONE M = 256 MMA
//   (tcgen05.mma.cta_group::2 issued by the leader, rank 0): each CTA unpacks its own 128 variants into its own
//   tensor memory and holds HALF of the rows of every basis panel in its shared memory, so the L2 -> SM traffic of
//   the basis halves against two single CTAs and one instruction feeds both tensor cores.  Cross-CTA hand-offs:
//   "A ring slot written" and "accumulators read out" are counted on the leader's barriers (remote arrives of the
//   peer's warps); the leader's commits are multicast to both CTAs' "slot / stage / accumulator" barriers; the basis
//   TMA of either CTA completes its bytes on the leader's "stage filled" barrier.
//
// Two independent shared-memory rings: the genotype tiles come from HBM (long latency, 16 KB per chunk, released
// as soon as the unpack warps have read them -> deep prefetch), the basis panels come from L2 (short latency,
// 4 * ncols * 128 B per chunk, held until the tensor core has consumed them).
template <int NG, int CS>
__global__ void __launch_bounds__(THREADS, 1)
tc_sweep_kernel(const __grid_constant__ CUtensorMap geno_map, const __grid_constant__ CUtensorMap b_map, const Params p) {
  static_assert(CS == 1 || CS == 2, "one CTA or a CTA pair");
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;   // genotype ring base (shared window address)
  uint8_t* smem_gen = smem_raw + (smem0 - smem_u32(smem_raw));
  const uint32_t bring0 = smem0 + p.n_gstages * p.gstage_bytes;   // basis-panel ring base
  Barriers* bars = reinterpret_cast<Barriers*>(smem_gen + (size_t)p.n_gstages * p.gstage_bytes +
                                               (size_t)p.n_bstages * p.bstage_bytes);
  const uint32_t bar0 = smem_u32(bars);
  auto GFULL = [&](int s) { return bar0 + 8u * s; };
  auto GEMPTY = [&](int s) { return bar0 + 8u * (MAX_GSTAGES + s); };
  auto BFULL = [&](int s) { return bar0 + 8u * (2 * MAX_GSTAGES + s); };
  auto BEMPTY = [&](int s) { return bar0 + 8u * (2 * MAX_GSTAGES + MAX_BSTAGES + s); };
  auto AFULL = [&](int s) { return bar0 + 8u * (2 * MAX_GSTAGES + 2 * MAX_BSTAGES + s); };
  auto AEMPTY = [&](int s) { return bar0 + 8u * (2 * MAX_GSTAGES + 2 * MAX_BSTAGES + MAX_RING + s); };
  const uint32_t DFULL = bar0 + 8u * (2 * MAX_GSTAGES + 2 * MAX_BSTAGES + 2 * MAX_RING);
  const uint32_t DEMPTY = DFULL + 8u;
  const int n_groups = NG ? NG : p.n_groups;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.n_gstages; ++s) {
      mbar_init(GFULL(s), 1);
      mbar_init(GEMPTY(s), UNPACK_WARPS);
    }
    for (int s = 0; s < p.n_bstages; ++s) {
      mbar_init(BFULL(s), 1);
      mbar_init(BEMPTY(s), 1);    // the (pair-multicast) MMA commit
    }
    for (int s = 0; s < p.ring_groups; ++s) {
      mbar_init(AFULL(s), UNPACK_WARPS * CS);   // per CTA, one-plane: 16 warps x 1 arrival; two-plane: 8 warps x 2
      mbar_init(AEMPTY(s), 1);
    }
    mbar_init(DFULL, 1);
    mbar_init(DEMPTY, 4 * CS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == WARP_MMA) {
    if (CS == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&bars->tmem_base))
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&bars->tmem_base))
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  if (warp == WARP_TMA_G && lane == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(&geno_map) : "memory");
  if (warp == WARP_TMA_B && lane == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(&b_map) : "memory");
  tc_fence_before();
  __syncthreads();
  if (CS > 1) cluster_sync_all();   // peers must see initialised barriers before any multicast / remote arrive
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;

  // tile schedule: cluster k sweeps tiles CS * (k + i * n_clusters) + rank; the trip count is cluster-uniform
  // (tiles past the end are all-zero boxes whose results are never stored)
  const int cta_rank = CS > 1 ? (int)cluster_ctarank() : 0;
  const int first_tile = (CS > 1 ? (int)cluster_id_x() * CS : (int)blockIdx.x) + cta_rank;
  const int tile_step = CS > 1 ? (int)n_clusters_x() * CS : (int)gridDim.x;
  const int tile_end = CS > 1 ? p.n_tiles + cta_rank : p.n_tiles;   // (tile - rank) < n_tiles for every CTA alike
  const int panel_bytes = p.ncols / CS * 128;   // this CTA's rows of one basis panel
  // the mode of a tile pair is the pair's: both CTAs, and the one MMA stream, must agree on the ring layout
  auto tile_mode = [&](int tile) {
    if (CS == 1) return tile_has_missing(p, tile);
    const int t0 = tile - cta_rank;
    return tile_has_missing(p, t0) || tile_has_missing(p, t0 + 1);
  };

  if (warp == WARP_TMA_G) {
    // ============================== genotype producer ==============================
    int gs = 0;
    uint32_t g_phase = 0;   // parity of the number of completed passes over the ring
    for (int tile = first_tile; tile < tile_end; tile += tile_step) {
      for (int ch = 0; ch < p.n_chunks; ++ch) {
        mbar_wait(GEMPTY(gs), g_phase ^ 1);
        const uint32_t sbase = smem0 + gs * p.gstage_bytes;
        if (elect_one()) {
#if LRR_ABL_NO_GENO
          mbar_arrive_expect_tx(GFULL(gs), (uint32_t)(p.mask_bytes));
#else
          mbar_arrive_expect_tx(GFULL(gs), (uint32_t)(GENO_BYTES + p.mask_bytes));
          tma_load_2d(&geno_map, GFULL(gs), sbase, ch * 128, tile * TILE_M);
#endif
          if (p.mask_bytes) {
            for (int g = 0; g < n_groups; ++g)
              bulk_load_1d(sbase + GENO_BYTES + g * 128, p.g[g].mask_hi + ch * (CHUNK / 16), 128, GFULL(gs));
          }
        }
        __syncwarp();
        if (++gs == p.n_gstages) { gs = 0; g_phase ^= 1; }
      }
    }
  } else if (warp == WARP_TMA_B) {
    // ============================== basis-panel producer ==============================
    int bs = 0;
    uint32_t b_phase = 0;
    for (int tile = first_tile; tile < tile_end; tile += tile_step) {
      for (int ch = 0; ch < p.n_chunks; ++ch) {
        mbar_wait(BEMPTY(bs), b_phase ^ 1);
        const uint32_t sbase = bring0 + bs * p.bstage_bytes;
        if (elect_one()) {
#if LRR_ABL_NO_BPANEL
          mbar_arrive(BFULL(bs));
#else
          if (CS == 1) {
            mbar_arrive_expect_tx(BFULL(bs), (uint32_t)(SLOTS * panel_bytes));
#pragma unroll
            for (int s = 0; s < SLOTS; ++s)
              tma_load_2d(&b_map, BFULL(bs), sbase + s * panel_bytes, ch * CHUNK + s * SLOT, 0);
          } else {
            // both halves complete on the leader's barrier; only the leader arms it (for the bytes of both)
            if (cta_rank == 0) mbar_arrive_expect_tx(BFULL(bs), (uint32_t)(2 * (SLOTS >> LRR_ABL_HALF_B) * panel_bytes));
            const uint32_t full_leader = mapa_leader(BFULL(bs));
#pragma unroll
            for (int s = 0; s < (SLOTS >> LRR_ABL_HALF_B); ++s)
              tma_load_2d_pair(&b_map, full_leader, sbase + s * panel_bytes, ch * CHUNK + s * SLOT,
                               cta_rank * (p.ncols / 2));
          }
#endif
        }
        __syncwarp();
        if (++bs == p.n_bstages) { bs = 0; b_phase ^= 1; }
      }
    }
  } else if (warp == WARP_MMA) {
    // ============================== MMA issuer ==============================
    // (CTA pair: only the leader issues; the peer's MMA warp just owns its half of the tensor-memory allocation)
    if (CS == 1 || cta_rank == 0) {
    // The whole warp runs the loop in uniform control flow; one elected lane issues tcgen05.mma / commit.
    // This warp's serial instruction stream is what paces the kernel, so the steady state of one-plane tiles is
    // unrolled over the (equal) depths of the basis-panel ring and the A ring: every shared-memory, tensor-memory
    // and barrier address is then a base plus a compile-time constant and stays on the uniform datapath.
    const uint32_t idesc = make_idesc(LRR_ABL_MMA_N ? LRR_ABL_MMA_N : p.ncols, TILE_M * CS);
    auto mma = [&](uint32_t d, uint32_t a, uint64_t b, uint32_t acc) {
      if (CS == 1) mma_i8_ts(d, a, b, idesc, acc); else mma_i8_ts_pair(d, a, b, idesc, acc);
    };
    auto commit = [&](uint32_t bar) {
      if (CS == 1) tc_commit(bar); else tc_commit_pair(bar);
    };
    auto wait_peer = [&](uint32_t bar, uint32_t parity) {   // barriers the peer CTA's warps arrive on, too
      mbar_wait(bar, parity);
    };
    int bs = 0;
    uint32_t b_phase = 0;
    int rg = 0;              // ring group of the next group instance (instances are numbered across tiles)
    uint32_t rg_par = 0;     // parity of the number of completed passes over the ring groups
    uint32_t tile_i = 0;
    const uint64_t desc0 = make_kmajor_sw128_desc(bring0);
    const uint32_t stage_d = (uint32_t)p.bstage_bytes >> 4;   // descriptor-address units (16 B)
    const uint32_t panel_d = (uint32_t)panel_bytes >> 4;
    const uint32_t a_ring = tmem + p.ring_base;

    // one group instance, any mode (alignment, tails, two-plane tiles)
    auto generic_group = [&](int ch, int s0, bool two_plane, int gsl) {
      wait_peer(AFULL(rg), rg_par);
      tc_fence_after();
      const uint32_t a_g = a_ring + rg * GROUP_COLS;
      const uint64_t bd = desc0 + (uint64_t)(bs * stage_d);
      if (elect_one()) {
#if !LRR_ABL_NO_MMA
        if (!two_plane) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int j = 0; j < LRR_ABL_MMA_J; ++j)
              mma(tmem, a_g + k * 32 + j * 8, bd + (uint64_t)(k * panel_d + j * 2), (ch | k | j) ? 1u : 0u);
        } else {
#pragma unroll
          for (int k = 0; k < 2; ++k)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint32_t acc = (ch | s0 | k | j) ? 1u : 0u;
              const uint64_t d = bd + (uint64_t)((s0 + k) * panel_d + j * 2);
              mma(tmem, a_g + k * 32 + j * 8, d, acc);
              mma(tmem + p.ncols, a_g + 64 + k * 32 + j * 8, d, acc);
            }
        }
#endif
        commit(AEMPTY(rg));
        if (s0 + gsl == SLOTS) commit(BEMPTY(bs));
      }
      __syncwarp();
      if (++rg == p.ring_groups) { rg = 0; rg_par ^= 1; }
    };
    auto generic_chunk = [&](int ch, bool two_plane) {
      const int gsl = two_plane ? 2 : 4;   // slots per group
      mbar_wait(BFULL(bs), b_phase);
      for (int s0 = 0; s0 < SLOTS; s0 += gsl) generic_group(ch, s0, two_plane, gsl);
      if (++bs == p.n_bstages) { bs = 0; b_phase ^= 1; }
    };

    // the fast path is unrolled over NB chunks == the depth of the basis-panel ring (6 when shared memory allows:
    // the panels come from L2 behind the HBM stream and 3 stages do not cover that latency; measured +10 %)
    // Fast path of one tile, compile-time mode (TP) and A-ring depth (RG).  Requires b-stage 0 and ring group 0 at
    // entry and (NB * groups-per-chunk) % RG == 0, so that ring positions repeat every NB chunks.
    auto fast_chunks = [&](auto tp_tag, auto rg_tag, auto nb_tag, int& ch) {
      constexpr bool TP = decltype(tp_tag)::value;
      constexpr int RG = decltype(rg_tag)::value;
      constexpr int NB = decltype(nb_tag)::value;
      constexpr int GPC = TP ? 2 : 1;              // group instances per chunk
      constexpr int WRAPS = NB * GPC / RG;         // ring passes per unrolled block
      static_assert((NB * GPC) % RG == 0, "ring positions must repeat every NB chunks");
      for (; ch + NB <= p.n_chunks; ch += NB) {
#pragma unroll
        for (int u = 0; u < NB; ++u) {
          mbar_wait(BFULL(u), b_phase);
          const uint64_t bd = desc0 + (uint64_t)(u * stage_d);
#pragma unroll
          for (int h = 0; h < GPC; ++h) {
            constexpr int dummy = 0; (void)dummy;
            const int inst = u * GPC + h;
            const int g = inst % RG;
            const uint32_t par = rg_par ^ (uint32_t)((inst / RG) & 1);
            wait_peer(AFULL(g), par);
            tc_fence_after();
            if (elect_one()) {
#if !LRR_ABL_NO_MMA
              const uint32_t a_g = a_ring + g * GROUP_COLS;
              if (!TP) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
#pragma unroll
                  for (int j = 0; j < LRR_ABL_MMA_J; ++j) {
                    const uint32_t acc = (u | k | j) ? 1u : (ch ? 1u : 0u);
                    mma(tmem, a_g + k * 32 + j * 8, bd + (uint64_t)(k * panel_d + j * 2), acc);
                  }
              } else {
#pragma unroll
                for (int k = 0; k < 2; ++k)
#pragma unroll
                  for (int j = 0; j < 4; ++j) {
                    const uint32_t acc = (u | h | k | j) ? 1u : (ch ? 1u : 0u);
                    const uint64_t d = bd + (uint64_t)((2 * h + k) * panel_d + j * 2);
                    mma(tmem, a_g + k * 32 + j * 8, d, acc);
                    mma(tmem + p.ncols, a_g + 64 + k * 32 + j * 8, d, acc);
                  }
              }
#endif
              commit(AEMPTY(g));
              if (h == GPC - 1) commit(BEMPTY(u));
            }
            __syncwarp();
          }
        }
        b_phase ^= 1;
        if (WRAPS & 1) rg_par ^= 1;
      }
    };

    for (int tile = first_tile; tile < tile_end; tile += tile_step, ++tile_i) {
      const bool two_plane = tile_mode(tile);
      wait_peer(DEMPTY, (tile_i & 1) ^ 1);   // the previous tile's accumulators have been read out (by both CTAs)
      tc_fence_after();
      int ch = 0;
      if (p.n_bstages == 6 || p.n_bstages == 3) {
        while (ch < p.n_chunks && bs != 0) generic_chunk(ch++, two_plane);   // align to b-stage 0
        if (rg == 0 && ch < p.n_chunks) {
          if (p.n_bstages == 6) {
            if (!two_plane && p.ring_groups == 3) fast_chunks(FalseTag{}, IntTag<3>{}, IntTag<6>{}, ch);
            else if (two_plane && p.ring_groups == 3) fast_chunks(TrueTag{}, IntTag<3>{}, IntTag<6>{}, ch);
            else if (two_plane && p.ring_groups == 2) fast_chunks(TrueTag{}, IntTag<2>{}, IntTag<6>{}, ch);
          } else {
            if (!two_plane && p.ring_groups == 3) fast_chunks(FalseTag{}, IntTag<3>{}, IntTag<3>{}, ch);
            else if (two_plane && p.ring_groups == 3) fast_chunks(TrueTag{}, IntTag<3>{}, IntTag<3>{}, ch);
            else if (two_plane && p.ring_groups == 2) fast_chunks(TrueTag{}, IntTag<2>{}, IntTag<3>{}, ch);
          }
        }
      }
      for (; ch < p.n_chunks; ++ch) generic_chunk(ch, two_plane);
      if (elect_one()) commit(DFULL);
      __syncwarp();
    }
    }
  } else {
    // ============================== unpack + epilogue warps ==============================
    const int q = warp & 3;          // TMEM lane quarter this warp may access
    const int s = warp >> 2;         // slot of every chunk this warp produces
    const int row = q * 32 + lane;   // variant row within the tile == TMEM lane
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const uint32_t row_off = (uint32_t)row * 128u;
    const uint32_t swz = (uint32_t)(row & 7);
    const uint32_t ld0 = row_off + (((uint32_t)(2 * s) ^ swz) << 4);       // 16-byte chunks 2s, 2s+1 (swizzled)
    const uint32_t ld1 = row_off + (((uint32_t)(2 * s + 1) ^ swz) << 4);
    uint32_t gaddr = smem0;          // shared address of the current genotype stage
    uint32_t gbar = GFULL(0);        // its "full" barrier ("empty" is MAX_GSTAGES * 8 bytes further)
    const uint32_t gaddr_end = smem0 + p.n_gstages * p.gstage_bytes;
    uint32_t g_phase = 0;
    int rg0 = 0;                     // ring group of the next chunk's first group instance (see the MMA warp)
    uint32_t rg0_par = 0;
    uint32_t tile_i = 0;
    bool prev_two_plane = false;
    const int ring_groups = p.ring_groups;
    const uint32_t a_ring = tmem + lane_addr + p.ring_base;
    int n2[NG ? NG : MAX_GROUPS];
    // "slot written" / "accumulators read out" are counted on the pair leader's barriers
    const uint32_t afull_leader = (CS == 2) ? mapa_leader(AFULL(0)) : 0u;
    const uint32_t dempty_leader = (CS == 2) ? mapa_leader(DEMPTY) : 0u;
    auto arrive_afull = [&](int g) {
      if (CS == 1) mbar_arrive(AFULL(g)); else mbar_arrive_cluster(afull_leader + 8u * (uint32_t)g);
    };

    // The chunk loop of one tile, specialised at compile time on the tile's mode (TP: two planes, the tile has
    // missing calls) and on whether any group needs its sample mask for the hom-alt count (MA: none does).
    auto run_chunks = [&](auto tp_tag, auto ma_tag) {
      constexpr bool TP = decltype(tp_tag)::value;
      constexpr bool MA = decltype(ma_tag)::value;
      // one-plane: the 4 slots of a chunk form one group; two-plane: slots {0,1} and {2,3} form two groups
      const int hh = TP ? (s >> 1) : 0;                   // which group of the chunk this warp feeds
      const uint32_t in_group = TP ? (uint32_t)(s & 1) * 32u : (uint32_t)s * 32u;
      int pend_rg = -1;    // TMEM store issued but not yet published to the MMA warp
      // one chunk: `rg` / `rg_par` = ring group and pass parity of this warp's group instance, `pend` = the ring
      // group of the previous chunk's store (-1: none).  With compile-time arguments everything folds.
      auto chunk_body = [&](const int rg, const uint32_t rg_par, const int pend) {
        mbar_wait(gbar, g_phase);
        // packed bytes of samples [128 s, 128 s + 128) of this row
        const uint4 w0 = lds128(gaddr + ld0);
        const uint4 w1 = lds128(gaddr + ld1);
        const uint32_t w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
        uint32_t mm[NG ? NG : MAX_GROUPS][8];
        if (!MA) {
#pragma unroll
          for (int g = 0; g < (NG ? NG : MAX_GROUPS); ++g) {
            if (NG || g < n_groups) {
              const uint4 m0 = lds128(gaddr + GENO_BYTES + g * 128 + s * 32);
              const uint4 m1 = lds128(gaddr + GENO_BYTES + g * 128 + s * 32 + 16);
              mm[g][0] = m0.x; mm[g][1] = m0.y; mm[g][2] = m0.z; mm[g][3] = m0.w;
              mm[g][4] = m1.x; mm[g][5] = m1.y; mm[g][6] = m1.z; mm[g][7] = m1.w;
            }
          }
        }
        // 2-bit fields -> bytes.  Fields at bit offsets 0 and 4 of every byte keep their value c; fields at
        // offsets 2 and 6 are left in place (value 4c) -- their basis rows were quantised as v/4 -- so a
        // 16-call word costs one shift and four masks.  (Leaving all four fields in place -- c, 4c, 16c, 64c, no
        // shift -- was measured: no faster, and it costs 4 more bits of the covariate digits.)
        uint32_t rc[32];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
#if !LRR_ABL_NO_UNPACK
          const uint32_t t = shr<4>(w[i]);
          rc[4 * i + 0] = w[i] & 0x03030303u;
          rc[4 * i + 1] = w[i] & 0x0C0C0C0Cu;
          rc[4 * i + 2] = t & 0x03030303u;
          rc[4 * i + 3] = t & 0x0C0C0C0Cu;
#else
          rc[4 * i + 0] = w[i]; rc[4 * i + 1] = w[i]; rc[4 * i + 2] = w[i]; rc[4 * i + 3] = w[i];
#endif
        }
        // exact counts per group for x.x = n1 + 4 n2.  Tiles without missing calls whose groups keep every sample
        // count all set bits (n1 + n2; the epilogue solves with the ones column n1 + 2 n2); otherwise the hom-alt
        // calls (code 2: high bit set, low bit clear) of the group's samples are counted directly.
#pragma unroll
        for (int g = 0; g < (NG ? NG : MAX_GROUPS); ++g) {
          if (NG || g < n_groups) {
            int acc = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              if (MA && !TP) {
                acc += __popc(w[i]);
              } else {
                const uint32_t m = MA ? 0xAAAAAAAAu : mm[g][i];
                acc += TP ? __popc(w[i] & ~(w[i] << 1) & m) : __popc(w[i] & m);
              }
            }
            n2[g] += acc;
          }
        }
        // the genotype stage is in registers now: hand it back to the TMA producer
        __syncwarp();
        if (lane == 0) mbar_arrive(gbar + 8u * MAX_GSTAGES);
        // retire the previous chunk's TMEM store only now: its latency hid behind this chunk's unpack arithmetic
        if (pend >= 0) {
          tmem_wait_st();
          tc_fence_before();
          __syncwarp();
          if (lane < (TP ? 2 : 1)) arrive_afull(pend);
        }
        mbar_wait(AEMPTY(rg), rg_par ^ 1u);
        tc_fence_after();
        const uint32_t a_c = a_ring + rg * GROUP_COLS + in_group;
#if !LRR_ABL_NO_STTM
        tmem_st32(a_c, rc);
        if (TP) {
          // missing-indicator plane, same byte scaling as plane c (1 or 4)
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const uint32_t mw = w[i] & (w[i] >> 1);   // bit 2f set iff field f is code 3
            const uint32_t t = shr<4>(mw);
            rc[4 * i + 0] = mw & 0x01010101u;
            rc[4 * i + 1] = mw & 0x04040404u;
            rc[4 * i + 2] = t & 0x01010101u;
            rc[4 * i + 3] = t & 0x04040404u;
          }
          tmem_st32(a_c + 64, rc);
        }
#else
        { uint32_t x = 0;
#pragma unroll
          for (int i = 0; i < 32; ++i) x ^= rc[i];
          if (x == 0x12345678u) n2[0] += 1; }
#endif
        gaddr += p.gstage_bytes;
        gbar += 8u;
        if (gaddr == gaddr_end) { gaddr = smem0; gbar = GFULL(0); g_phase ^= 1; }
      };
      // generic step: ring position from the running counters
      auto generic_chunk = [&]() {
        int rg = rg0 + hh;   // this warp's group instance: rg0 (+1 for the second group of a two-plane chunk)
        uint32_t rg_par = rg0_par;
        if (rg >= ring_groups) { rg -= ring_groups; rg_par ^= 1; }
        chunk_body(rg, rg_par, pend_rg);
        pend_rg = rg;
        rg0 += TP ? 2 : 1;
        if (rg0 >= ring_groups) { rg0 -= ring_groups; rg0_par ^= 1; }
      };
      int ch = 0;
      if (!TP && ring_groups == 3) {
        // steady state of one-plane tiles: unrolled over the 3 ring groups (positions become constants)
        do { generic_chunk(); ++ch; } while (ch < p.n_chunks && rg0 != 0);
        for (; ch + 3 <= p.n_chunks; ch += 3) {   // entered with rg0 == 0 and pend_rg == 2
          chunk_body(0, rg0_par, pend_rg);
          chunk_body(1, rg0_par, 0);
          chunk_body(2, rg0_par, 1);
          rg0_par ^= 1;
          pend_rg = 2;
        }
      }
      for (; ch < p.n_chunks; ++ch) generic_chunk();
      if (pend_rg >= 0) {   // flush the last chunk of the tile
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane < (TP ? 2 : 1)) arrive_afull(pend_rg);
      }
    };

    for (int tile = first_tile; tile < tile_end; tile += tile_step, ++tile_i) {
      const bool two_plane = tile_mode(tile);
      if (tile_i > 0 && two_plane != prev_two_plane) {
        // the ring is laid out differently: wait until every MMA of the previous tile has retired
        mbar_wait(DFULL, (tile_i - 1) & 1);
        tc_fence_after();
      }
      prev_two_plane = two_plane;
#pragma unroll
      for (int g = 0; g < (NG ? NG : MAX_GROUPS); ++g) n2[g] = 0;
      if (two_plane) {
        if (p.mask_bytes == 0) run_chunks(TrueTag{}, TrueTag{}); else run_chunks(TrueTag{}, FalseTag{});
      } else {
        if (p.mask_bytes == 0) run_chunks(FalseTag{}, TrueTag{}); else run_chunks(FalseTag{}, FalseTag{});
      }

      // ------------------------------ per-tile epilogue ------------------------------
      if (s > 0) {
#pragma unroll
        for (int g = 0; g < (NG ? NG : MAX_GROUPS); ++g)
          if (NG || g < n_groups) bars->n2_xchg[tile_i & 1][s - 1][g][row] = n2[g];
      }
      named_bar_sync(1, UNPACK_WARPS * 32);
      if (s == 0) {
        mbar_wait(DFULL, tile_i & 1);
        tc_fence_after();
        const int64_t v = (int64_t)tile * TILE_M + row;
        const uint32_t d_c = tmem + lane_addr, d_m = tmem + lane_addr + p.ncols;
#pragma unroll
        for (int g = 0; g < (NG ? NG : MAX_GROUPS); ++g) {
          if (!(NG || g < n_groups)) continue;
          const GroupMeta& G = p.g[g];
          const int cnt = n2[g] + bars->n2_xchg[tile_i & 1][0][g][row] + bars->n2_xchg[tile_i & 1][1][g][row] +
                          bars->n2_xchg[tile_i & 1][2][g][row];
          // the group's columns: Kd x N_SLICES_Q then P x N_SLICES_Y digit columns, then one "ones" column
          const int n_digit_cols = G.Kd * N_SLICES_Q + (G.C - G.Kd) * N_SLICES_Y;
          const int ones_col = G.col_off + n_digit_cols;
          uint32_t r16[16];
          tmem_ld16(d_c + (ones_col & ~15), r16);
          tmem_wait_ld();
          int sc = 0;
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (i == (ones_col & 15)) sc = (int)r16[i];
          int nm = 0;
          if (two_plane) {
            tmem_ld16(d_m + (ones_col & ~15), r16);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (i == (ones_col & 15)) nm = (int)r16[i];
          }
          sc >>= 2;                           // the ones column carries 4 per unit (see the quantiser)
          nm >>= 2;
          const int S = sc - 3 * nm;          // n1 + 2 n2
          // one-plane tiles of unmasked passes counted every set bit: n1 + n2 (see the unpack warps)
          const int n2g = (!two_plane && p.mask_bytes == 0) ? S - cnt : cnt;
          const int n1 = S - 2 * n2g;
          const double mean = (double)S / (double)(G.n - nm);
          if (v < p.M) reinterpret_cast<int4*>(G.counts)[v] = make_int4(n1, n2g, nm, 0);
          // digit columns, 16 TMEM columns at a time; (c, sl) walk the columns and their digits
          const int c_lo = G.col_off, c_hi = ones_col;
          long long hi = 0, lo = 0, mhi = 0, mlo = 0;
          int c = 0, sl = 0, nd = G.Kd > 0 ? N_SLICES_Q : N_SLICES_Y;
          for (int base = c_lo & ~15; base < c_hi; base += 16) {
            uint32_t dc[16], dm[16];
            tmem_ld16(d_c + base, dc);
            if (two_plane) tmem_ld16(d_m + base, dm);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int col = base + i;
              if (col >= c_lo && col < c_hi) {
                const long long dv = (long long)(int)dc[i] - (two_plane ? 3ll * (long long)(int)dm[i] : 0ll);
                const long long mv = two_plane ? (long long)(int)dm[i] : 0ll;
                if (sl < 3) {
                  lo += dv * (1ll << (8 * sl));
                  mlo += mv * (1ll << (8 * sl));
                } else {
                  hi += dv * (1ll << (8 * (sl - 3)));
                  mhi += mv * (1ll << (8 * (sl - 3)));
                }
                if (++sl == nd) {
                  const double scale = G.colscale[c];
                  double dot = fma((double)hi, 16777216.0, (double)lo) * scale;
                  if (two_plane && nm > 0) dot += mean * (fma((double)mhi, 16777216.0, (double)mlo) * scale);
                  if (v < p.M) G.dots[v * G.dots_stride + c] = dot;
                  hi = lo = mhi = mlo = 0;
                  sl = 0;
                  ++c;
                  nd = c < G.Kd ? N_SLICES_Q : N_SLICES_Y;
                }
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (CS == 1) mbar_arrive(DEMPTY); else mbar_arrive_cluster(dempty_leader);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (CS > 1) cluster_sync_all();   // no CTA may exit while its peer can still signal / run MMAs into it
  if (warp == WARP_MMA) {
    if (CS == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// basis quantisation: float64 column -> N_SLICES balanced base-256 digits (int8), plus the mask "ones" column
// ------------------------------------------------------------------------------------------------
__global__ void colmax_kernel(const double* __restrict__ basis, int C, int64_t ns_pad, unsigned long long* colmax_bits) {
  const int c = blockIdx.y;
  double m = 0.0;
  for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < ns_pad; j += (int64_t)gridDim.x * blockDim.x)
    m = fmax(m, fabs(basis[(int64_t)c * ns_pad + j]));
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(colmax_bits + c, (unsigned long long)__double_as_longlong(m));
}

__global__ void quantize_kernel(const double* __restrict__ basis, const uint32_t* __restrict__ mask, int C, int Kd,
                                int64_t ns_pad, const unsigned long long* __restrict__ colmax_bits, int col_off,
                                int8_t* __restrict__ bq, double* __restrict__ colscale) {
  const int c = blockIdx.y;  // 0..C-1 data columns, C = ones column
  const int nd = c < Kd ? N_SLICES_Q : N_SLICES_Y;
  const int first = col_off + (c < Kd ? c * N_SLICES_Q : Kd * N_SLICES_Q + (c - Kd) * N_SLICES_Y);
  // Imax = 127 * (256^S - 1) / 255 : the largest integer with S balanced digits in [-128, 127]
  const double imax = 127.0 * ((double)((1ull << (8 * nd)) - 1ull) / 255.0);
  for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < ns_pad; j += (int64_t)gridDim.x * blockDim.x) {
    const int f = (int)((j >> 2) & 1);   // 1: the sample's field sits at bit offset 2 or 6 of its byte and reaches the MMA as 4c
    if (c == C) {
      const uint32_t mw = mask[j >> 4];
      // fields at bit offsets 2 / 6 of a byte reach the tensor core as 4c, the others as c (see the unpack warps)
      const int in_mask = (int)((mw >> sample_shift((int)(j & 15))) & 1u);
      bq[(int64_t)first * ns_pad + j] = (int8_t)(in_mask * (f ? 1 : 4));
      continue;
    }
    const double cm = __longlong_as_double((long long)colmax_bits[c]);
    long long I = 0;
    if (cm > 0.0) I = __double2ll_rn(basis[(int64_t)c * ns_pad + j] / cm * (f ? 0.25 * imax : imax));
    for (int s = 0; s < nd; ++s) {
      long long d = ((I + 128) & 255) - 128;  // balanced digit in [-128, 127]
      bq[(int64_t)(first + s) * ns_pad + j] = (int8_t)d;
      I = (I - d) >> 8;
    }
    if (j == 0) colscale[c] = cm > 0.0 ? cm / imax : 0.0;
  }
}

// Worst-case magnitude of any partial sum of one accumulator column: the A bytes are 3 (or 12, fields left in place)
// at most, so for any genotypes |sum_j A_j d_j| <= max(sum_{d>0} a_j d_j, sum_{d<0} a_j |d_j|).  Checked against INT32
// on the host: the tensor-core path is only used when no accumulator can overflow.
__global__ void acc_bound_kernel(const int8_t* __restrict__ bq, int64_t ns_pad, unsigned long long* __restrict__ bound) {
  const int64_t r = blockIdx.y;
  unsigned long long pos = 0, neg = 0;
  for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < ns_pad; j += (int64_t)gridDim.x * blockDim.x) {
    const int d = bq[r * ns_pad + j];
    const unsigned long long a = ((j >> 2) & 1) ? 12ull : 3ull;
    if (d > 0) pos += a * (unsigned long long)d; else neg += a * (unsigned long long)(-d);
  }
  for (int o = 16; o > 0; o >>= 1) {
    pos += __shfl_xor_sync(0xffffffffu, pos, o);
    neg += __shfl_xor_sync(0xffffffffu, neg, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(bound + 2 * r, pos);
    atomicAdd(bound + 2 * r + 1, neg);
  }
}

// One sweep over the genotypes covers at most PASS_COLS digit columns (so that both accumulator planes and the A ring
// fit the 512 TMEM columns).  Wider problems (e.g. 128 phenotypes) run several passes, each over a contiguous range
// of every group's dot columns ("segment") plus that group's ones column.
constexpr int PASS_COLS = 128;

struct Segment {
  int group;     // index into Ctx::groups
  int c_first;   // first dot column of the group in this segment
  int n_cols;    // dot columns in this segment
  int kd_in;     // of which (leading) covariate columns
  int row0;      // first row of this segment in the pass's panel matrix (digit rows, then the ones row)
};

// launch shape of one pass for a given cluster size (the basis-panel box and ring depend on it: a CTA of a pair
// holds half of the panel rows)
struct PassShape {
  int n_gstages = 0, n_bstages = 0, bstage_bytes = 0, smem_bytes = 0;
  CUtensorMap b_map;
};

struct Pass {
  int ncols = 0;                 // panel rows, padded to 16
  int64_t bq_row0 = 0;           // first row of this pass in State::d_bq
  std::vector<Segment> segs;
  int gstage_bytes = 0;
  int ring_base = 0, ring_groups = 0, mask_bytes = 0;
  PassShape shape[2];            // [cluster size - 1]
};

struct State {
  bool prepared = false;
  bool usable = false;
  std::string why;
  EncodeTiledFn encode = nullptr;
  std::vector<Pass> passes;
  int8_t* d_bq = nullptr;
  double* d_colscale = nullptr;   // concatenated per group
  unsigned long long* d_colmax = nullptr;
  uint32_t* d_mask_hi = nullptr;  // [G][ns_pad/16] group masks shifted to the high bit of each field
  unsigned long long* d_bound = nullptr;   // scratch of acc_bound_kernel (grow-only)
  size_t bound_bytes = 0;
  std::vector<int> scale_off;
  int cluster = 2;
  bool attr_set = false;
};

static State* state(Ctx* c) {
  if (!c->tc_state) c->tc_state = new State();
  return static_cast<State*>(c->tc_state);
}

static void free_prepared(State* s) {
  cudaFree(s->d_bq);
  cudaFree(s->d_colscale);
  cudaFree(s->d_colmax);
  cudaFree(s->d_mask_hi);
  s->d_mask_hi = nullptr;
  s->d_bq = nullptr;
  s->d_colscale = nullptr;
  s->d_colmax = nullptr;
  s->passes.clear();
  s->prepared = false;
  s->usable = false;
}

static int encode_2d(State* s, CUtensorMap* map, const void* ptr, uint64_t inner, uint64_t outer, uint64_t row_stride,
                     uint32_t box_inner, uint32_t box_outer) {
  return tcc::encode_2d_u8(s->encode, map, ptr, inner, outer, row_stride, box_inner, box_outer);
}

// split every group's dot columns into passes of at most PASS_COLS panel rows
static void plan_passes(const Ctx* c, std::vector<Pass>& passes) {
  passes.clear();
  Pass cur;
  int used = 0;
  auto close = [&]() {
    if (!cur.segs.empty()) {
      cur.ncols = (used + 15) / 16 * 16;
      passes.push_back(cur);
    }
    cur = Pass();
    used = 0;
  };
  for (size_t g = 0; g < c->groups.size(); ++g) {
    const Group& gr = c->groups[g];
    int col = 0;
    while (col < gr.C) {
      // open a segment of group g in the current pass (needs room for one column + the ones row)
      const int nd0 = col < gr.Kd ? N_SLICES_Q : N_SLICES_Y;
      if (used + nd0 + 1 > PASS_COLS || (int)cur.segs.size() == MAX_GROUPS) close();
      Segment sg;
      sg.group = (int)g;
      sg.c_first = col;
      sg.n_cols = 0;
      sg.kd_in = 0;
      sg.row0 = used;
      int rows = 0;
      while (col < gr.C) {
        const int nd = col < gr.Kd ? N_SLICES_Q : N_SLICES_Y;
        if (used + rows + nd + 1 > PASS_COLS) break;
        rows += nd;
        if (col < gr.Kd) sg.kd_in++;
        sg.n_cols++;
        col++;
      }
      used += rows + 1;   // + ones row
      cur.segs.push_back(sg);
      if (col < gr.C) close();
    }
  }
  close();
}

// build the quantised basis for the current group list
static int prepare(Ctx* c) {
  State* s = state(c);
  if (s->prepared) return LRR_OK;
  free_prepared(s);
  s->prepared = true;
  s->usable = false;
  if (!s->encode && !(s->encode = tcc::get_encode_fn(&s->why))) return LRR_OK;
  const size_t G = c->groups.size();
  if (G == 0) {
    s->why = "no groups";
    return LRR_OK;
  }
  for (const Group& gr : c->groups)
    if (gr.weighted) {
      s->why = "weighted groups (x.x is not linear in the call codes) run on the float64 kernel";
      return LRR_OK;
    }
  int nscale = 0;
  s->scale_off.assign(G, 0);
  for (size_t g = 0; g < G; ++g) {
    s->scale_off[g] = nscale;
    nscale += c->groups[g].C;
  }
  plan_passes(c, s->passes);
  int64_t total_rows = 0;
  for (auto& ps : s->passes) {
    ps.bq_row0 = total_rows;
    total_rows += ps.ncols;
  }
  const int64_t ns_pad = c->groups[0].ns_pad;
  LRR_CUDA(c, cudaMalloc(&s->d_bq, (size_t)total_rows * ns_pad));
  LRR_CUDA(c, cudaMemset(s->d_bq, 0, (size_t)total_rows * ns_pad));
  LRR_CUDA(c, cudaMalloc(&s->d_colscale, sizeof(double) * (size_t)nscale));
  LRR_CUDA(c, cudaMalloc(&s->d_colmax, sizeof(unsigned long long) * (size_t)nscale));
  LRR_CUDA(c, cudaMemset(s->d_colmax, 0, sizeof(unsigned long long) * (size_t)nscale));
  const int64_t mask_words = ns_pad / 16;
  LRR_CUDA(c, cudaMalloc(&s->d_mask_hi, sizeof(uint32_t) * (size_t)mask_words * G));
  bool any_masked = false;
  const unsigned gx = (unsigned)std::min<int64_t>((ns_pad + 255) / 256, 1024);
  for (size_t g = 0; g < G; ++g) {
    const Group& gr = c->groups[g];
    if ((int64_t)gr.n != c->n_samples_total) any_masked = true;
    tcc::mask_hi_kernel<<<(unsigned)std::min<int64_t>((mask_words + 255) / 256, 1024), 256>>>(gr.d_mask, mask_words,
                                                                                        s->d_mask_hi + g * mask_words);
    for (int c0 = 0; c0 < gr.C; c0 += 65535)   // grid.y limit
      colmax_kernel<<<dim3(gx, (unsigned)std::min(gr.C - c0, 65535)), 256>>>(gr.d_basis + (int64_t)c0 * ns_pad,
                                                                            std::min(gr.C - c0, 65535), ns_pad,
                                                                            s->d_colmax + s->scale_off[g] + c0);
    c->launches += 2;
  }
  for (auto& ps : s->passes) {
    for (const Segment& sg : ps.segs) {
      const Group& gr = c->groups[sg.group];
      quantize_kernel<<<dim3(gx, (unsigned)(sg.n_cols + 1)), 256>>>(
          gr.d_basis + (int64_t)sg.c_first * ns_pad, gr.d_mask, sg.n_cols, sg.kd_in, ns_pad,
          s->d_colmax + s->scale_off[sg.group] + sg.c_first, (int)(ps.bq_row0 + sg.row0), s->d_bq,
          s->d_colscale + s->scale_off[sg.group] + sg.c_first);
      c->launches++;
    }
  }
  LRR_CUDA(c, cudaGetLastError());
  {
    // exactness guard: no INT32 accumulator can overflow, whatever the genotypes are
    // (grow-only scratch kept on the state: a cudaFree here would wait for every copy the streaming loop has in flight)
    const size_t bound_bytes = sizeof(unsigned long long) * 2 * (size_t)total_rows;
    if (bound_bytes > s->bound_bytes) {
      cudaFree(s->d_bound);
      s->d_bound = nullptr;
      s->bound_bytes = 0;
      LRR_CUDA(c, cudaMalloc(&s->d_bound, bound_bytes));
      s->bound_bytes = bound_bytes;
    }
    unsigned long long* d_bound = s->d_bound;
    LRR_CUDA(c, cudaMemsetAsync(d_bound, 0, bound_bytes, 0));
    for (int64_t r0 = 0; r0 < total_rows; r0 += 65535)
      acc_bound_kernel<<<dim3(gx, (unsigned)std::min<int64_t>(total_rows - r0, 65535)), 256>>>(s->d_bq + r0 * ns_pad, ns_pad,
                                                                                            d_bound + 2 * r0);
    c->launches++;
    std::vector<unsigned long long> h_bound(2 * (size_t)total_rows);
    cudaError_t e = cudaMemcpy(h_bound.data(), d_bound, sizeof(unsigned long long) * h_bound.size(), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return cuda_fail(c, e, "acc_bound_kernel");
    unsigned long long worst = 0;
    for (unsigned long long v : h_bound) worst = std::max(worst, v);
    if (worst > 2147483647ull) {
      s->why = "INT32 accumulators could overflow for this many samples (worst-case column bound " +
               std::to_string(worst) + " > 2^31 - 1)";
      return LRR_OK;
    }
  }
  LRR_CUDA(c, cudaDeviceSynchronize());
  const int budget = 227 * 1024 - (int)sizeof(Barriers) - 1024;
  for (auto& ps : s->passes) {
    // shared memory: genotype ring (16 KB [+ group masks] per stage) + basis-panel ring + barriers + 1 KB slack
    ps.mask_bytes = any_masked ? (int)ps.segs.size() * 128 : 0;
    ps.gstage_bytes = (GENO_BYTES + ps.mask_bytes + 1023) / 1024 * 1024;
    ps.ring_base = (2 * ps.ncols + 31) / 32 * 32;
    ps.ring_groups = (512 - ps.ring_base) / GROUP_COLS;
    if (ps.ring_groups > MAX_RING) ps.ring_groups = MAX_RING;
    if (ps.ring_groups < 2) {
      s->why = "not enough tensor memory for the A ring";
      return LRR_OK;
    }
    for (int cs = 1; cs <= 2; ++cs) {
      PassShape& sh = ps.shape[cs - 1];
      const int rows = ps.ncols / cs;   // panel rows held by one CTA
      if (encode_2d(s, &sh.b_map, s->d_bq + ps.bq_row0 * ns_pad, (uint64_t)ns_pad, (uint64_t)ps.ncols, (uint64_t)ns_pad,
                    SLOT, (uint32_t)rows)) {
        s->why = "cuTensorMapEncodeTiled failed for the basis panels";
        return LRR_OK;
      }
      sh.bstage_bytes = SLOTS * rows * 128;
      // basis-panel ring: 6 stages if at least 6 genotype stages still fit, else 3, else whatever fits (generic path)
      int bst = LRR_TC_NB;
      if (const char* e = tuning_env("LRR_TC_BSTAGES")) bst = atoi(e);
      if (bst > MAX_BSTAGES) bst = MAX_BSTAGES;
      if (bst > 3 && budget - bst * sh.bstage_bytes < 6 * ps.gstage_bytes) bst = 3;
      while (bst > 2 && budget - bst * sh.bstage_bytes < 3 * ps.gstage_bytes) --bst;
      int gst = (budget - bst * sh.bstage_bytes) / ps.gstage_bytes;
      if (gst > MAX_GSTAGES) gst = MAX_GSTAGES;
      if (const char* e = tuning_env("LRR_TC_GSTAGES")) { const int v = atoi(e); if (v >= 2 && v < gst) gst = v; }
      if (bst < 2 || gst < 2) {
        s->why = "not enough shared memory for the genotype / basis-panel rings";
        return LRR_OK;
      }
      sh.n_gstages = gst;
      sh.n_bstages = bst;
      sh.smem_bytes = gst * ps.gstage_bytes + bst * sh.bstage_bytes + (int)sizeof(Barriers) + 1024;
    }
  }
  if (!s->attr_set) {
#define LRR_SET_SMEM(NG_, CS_) \
  LRR_CUDA(c, cudaFuncSetAttribute(tc_sweep_kernel<NG_, CS_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024))
    LRR_SET_SMEM(0, 1); LRR_SET_SMEM(1, 1); LRR_SET_SMEM(2, 1);
    LRR_SET_SMEM(0, 2); LRR_SET_SMEM(1, 2); LRR_SET_SMEM(2, 2);
#undef LRR_SET_SMEM
    s->attr_set = true;
  }
  if (const char* e = tuning_env("LRR_TC_CLUSTER")) {
    const int v = atoi(e);
    if (v == 1 || v == 2) s->cluster = v;
  }
  s->usable = true;
  s->why.clear();
  return LRR_OK;
}

}  // namespace tc

bool tc_supported(Ctx* c, bool /*may_have_missing*/) {
  if (tc::prepare(c) != LRR_OK) return false;
  tc::State* s = tc::state(c);
  if (!s->usable) {
    c->err = s->why;
    return false;
  }
  return true;
}

// per-column quantum (value of one unit of the lowest digit) of group g's dot columns, after tc_supported
const double* tc_quantum(Ctx* c, int g) {
  tc::State* s = tc::state(c);
  return s->d_colscale ? s->d_colscale + s->scale_off[g] : nullptr;
}

void tc_invalidate(Ctx* c) {
  if (!c->tc_state) return;
  tc::State* s = static_cast<tc::State*>(c->tc_state);
  tc::free_prepared(s);
}

void tc_release(Ctx* c) {
  if (!c->tc_state) return;
  tc::State* s = static_cast<tc::State*>(c->tc_state);
  tc::free_prepared(s);
  cudaFree(s->d_bound);
  delete s;
  c->tc_state = nullptr;
}

int launch_tc_sweep(Ctx* c, const uint8_t* d_packed, const uint8_t* d_row_flags, int64_t M, int64_t stride,
                    cudaStream_t st) {
  using namespace tc;
  if (M == 0) return LRR_OK;
  if (int r = prepare(c)) return r;
  State* s = state(c);
  if (!s->usable) return fail(c, LRR_EINVAL, "tensor-core kernel unavailable: " + s->why);
  CUtensorMap geno_map;
  if (encode_2d(s, &geno_map, d_packed, (uint64_t)stride, (uint64_t)M, (uint64_t)stride, 128, TILE_M))
    return fail(c, LRR_ECUDA, "cuTensorMapEncodeTiled failed for the genotype store (pointer must be 16-byte aligned)");
  for (const Pass& ps : s->passes) {
    Params p;
    memset(&p, 0, sizeof p);
    p.M = M;
    p.n_tiles = (int)((M + TILE_M - 1) / TILE_M);
    p.n_chunks = (int)(stride / 128);
    p.ncols = ps.ncols;
    // cluster size: a CTA pair by default (LRR_TC_CLUSTER=1 forces single CTAs); tiny inputs run single CTAs
    int cs = s->cluster;
    if (p.n_tiles < 2 * cs) cs = 1;
    const PassShape& sh = ps.shape[cs - 1];
    p.n_gstages = sh.n_gstages;
    p.n_bstages = sh.n_bstages;
    p.n_groups = (int)ps.segs.size();
    p.ring_base = ps.ring_base;
    p.ring_groups = ps.ring_groups;
    p.gstage_bytes = ps.gstage_bytes;
    p.bstage_bytes = sh.bstage_bytes;
    p.mask_bytes = ps.mask_bytes;
    p.row_flags = d_row_flags;
    for (int i = 0; i < p.n_groups; ++i) {
      const Segment& sg = ps.segs[i];
      const Group& gr = c->groups[sg.group];
      p.g[i].col_off = sg.row0;
      p.g[i].C = sg.n_cols;
      p.g[i].Kd = sg.kd_in;
      p.g[i].n = gr.n;
      p.g[i].counts = c->d_counts + (int64_t)sg.group * c->reserved_variants * 4;
      p.g[i].dots = c->d_dots + c->dots_offset[sg.group] + sg.c_first;
      p.g[i].dots_stride = gr.C;
      p.g[i].colscale = s->d_colscale + s->scale_off[sg.group] + sg.c_first;
      p.g[i].mask_hi = s->d_mask_hi + (int64_t)sg.group * (gr.ns_pad / 16);
      p.g[i].mask_all = ((int64_t)gr.n == c->n_samples_total) ? 1 : 0;
    }
    void* kfn = nullptr;
#define LRR_PICK(NG_)                                                                       \
  (cs == 2 ? (void*)tc_sweep_kernel<NG_, 2> : (void*)tc_sweep_kernel<NG_, 1>)
    kfn = p.n_groups == 1 ? LRR_PICK(1) : p.n_groups == 2 ? LRR_PICK(2) : LRR_PICK(0);
#undef LRR_PICK
    void* args[3] = {(void*)&geno_map, (void*)&sh.b_map, (void*)&p};
    if (int r = tcc::launch_persistent_clusters(c, kfn, cs, THREADS, sh.smem_bytes, p.n_tiles, args, st)) return r;
    c->sweep_shape[0] += 1;
    c->sweep_shape[1] += ps.ncols;
    c->sweep_shape[2] += ps.ncols;   // every int8 sweep is two-plane capable
    c->sweep_shape[3] += ps.ncols;
  }
  return LRR_OK;
}

namespace tcc {
__global__ void mask_hi_kernel(const uint32_t* __restrict__ mask_lo, int64_t words, uint32_t* __restrict__ mask_hi) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < words; i += (int64_t)gridDim.x * blockDim.x)
    mask_hi[i] = mask_lo[i] << 1;
}
}  // namespace tcc

}  // namespace lrr
