// Kernel 3: the sweep on the tensor cores' 4-bit path (tcgen05.mma.kind::mxf4, sm_100a) -- half the tensor-memory
// traffic and half the unpack arithmetic of the int8 kernel (tc_kernel.cu), same exact-integer results.
//
// Same formulation as tc_kernel.cu (D[v, c] = sum_j x[v, j] B[j, c] per 128-variant tile, LinearRegression.scala:139-146
// restated), but both operands are E2M1 (4-bit float) elements:
//   A  call code c in {0..3} as the nibble 00cc = c / 2 (exactly representable: 0, .5, 1, 1.5), 64 samples per MMA;
//   B  every basis column as balanced base-13 digits u in {0, +-1, +-2, +-3, +-4, +-6, +-8} stored as u / 2 (the E2M1
//      values 0, .5, 1, 1.5, 2, 3, 4): a complete residue system mod 13.  The number of digits is chosen PER COLUMN by
//      the host (digit_policy below: 13 digits = 46.5 bits for the residualised phenotype columns, 6 for well-scaled
//      covariate columns plus one 11-digit "fitted value" column per phenotype that carries y_transpose_x, more for
//      heavy-tailed columns); what the chosen quantum costs is bounded per variant in the statistics epilogue
//      (stats_device.cuh) and rows outside the tolerance are recomputed in float64.  Block scales are all 1 (UE8M0 0x7F).
// Every product is a multiple of 1/4 and the f32 accumulator holds sum c u / 4 EXACTLY as long as |sum c u| <= 2^24;
// the host checks the worst case over all possible genotypes for every column (acc_bound_kernel) and refuses the
// kernel otherwise, so the result is independent of tiling and summation order, like the int8 kernel's INT32 sums.
// (Exactness of the hardware accumulation for such sums: scratch/fp4_probe.cu, and the bit-exact n / sum_x tests.)
//
// On-SM cost per 512-sample chunk against the int8 kernel (measured components, profiles/README.md): tcgen05.st
// 32 KB instead of 64 KB (it blocks the issuing SM sub-partition at 256 B/clk), 3 instead of 5 ALU operations per
// 16 calls, 8 instead of 16 MMAs of the same duration (N/2 cycles, K = 64 instead of 32).
//
// Structure (warps, rings, CTA pair with tcgen05.mma.cta_group::2) is the int8 kernel's; differences:
//   * A ring in units of 64 TMEM columns (one chunk of one plane); a one-plane chunk takes one unit, a two-plane
//     chunk (tile with missing calls) two consecutive units [plane c | plane m]; every unit has its own
//     full / empty barrier pair and BOTH sides track the phase parity per barrier, so units may be skipped;
//   * one-plane tiles use the columns of the unused plane-m accumulators as ring space (6 units instead of 4): the
//     MMAs of a chunk take ~200 ns plus commit latency and sit inside the unit's reuse loop;
//   * the scale-factor operands point at 16 TMEM columns filled with 0x7F bytes -- all scales are 1, which makes
//     the result independent of the scale-factor layout.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "common.cuh"
#include "tc_common.cuh"
#include "tc_ptx.cuh"

// Timing ablations (LRR_ABL_BITS / LRR_ABL_STREAM / LRR_ABL_CONTIG, see Params) are compiled in only with
// -DLRR_TC4_ABLATIONS=1 -DLRR_TUNING=1 (scratch/build_abl.sh); the shipped kernel carries none of their tests and the
// shipped library never reads the environment (common.cuh tuning_env).
#ifndef LRR_TC4_ABLATIONS
#define LRR_TC4_ABLATIONS 0
#endif
#if LRR_TC4_ABLATIONS
#define ABL(bit) (p.abl & (bit))
#define ABL_STREAM() (p.abl_stream)
#define ABL_CONTIG() (p.abl_contig)
#else
#define ABL(bit) (0)
#define ABL_STREAM() (0)
#define ABL_CONTIG() (0)
#endif

namespace lrr {
namespace tc4 {

using namespace ptx;
using tcc::IntTag;
using tcc::TrueTag;
using tcc::FalseTag;
using tcc::EncodeTiledFn;

constexpr int TILE_M = 128;            // variants per tile == TMEM lanes
constexpr int CHUNK = 512;             // samples per genotype stage (128 packed bytes per row)
constexpr int SLOT = 128;              // samples per unpack warp and chunk (16 TMEM columns of 8 nibbles)
constexpr int SLOTS = CHUNK / SLOT;    // 4
constexpr int SLOT_COLS = 16;
constexpr int UNIT_COLS = SLOTS * SLOT_COLS;   // 64: one chunk of one plane
constexpr int MAX_UNITS = 6;           // ring units: 6 (or 4) for one-plane tiles, 4 (two pairs) for two-plane tiles
constexpr int NU2 = 4;
constexpr int SF_BASE = 512 - 16;      // scale-factor columns sit at the top of tensor memory in both modes
constexpr int PANEL = 256;             // samples per basis-panel row (128 bytes of nibbles)
constexpr int PANELS = CHUNK / PANEL;  // 2
constexpr int UNPACK_WARPS = 16;       // warp w: TMEM lane quarter w & 3, slot w >> 2 of every chunk
constexpr int WARP_TMA_G = 16, WARP_TMA_B = 17, WARP_MMA = 18;
constexpr int THREADS = 19 * 32;
constexpr int GENO_BYTES = TILE_M * 128;   // 16 KB
constexpr int MAX_GROUPS = 4;
constexpr int MAX_GSTAGES = 10;
constexpr int NB = 6;                  // basis-panel ring depth the MMA fast path is unrolled for
constexpr int MAX_BSTAGES = NB;
constexpr int SF_COLS = 16;            // scale-factor columns (A: first 8, B: last 8), all bytes 0x7F
#ifndef LRR_TC4_TP_EAGER
#define LRR_TC4_TP_EAGER 1
#endif
constexpr int MAX_DIGITS = 13;         // base-13 digits per column: 13^13 / 3 < 2^53 / 3, recombined in two int64 halves
constexpr int PASS_COLS = 112;         // digit columns per sweep: 2 * 112 accumulators + 16 + ring of 4 * 64 <= 512
constexpr int PASS_COLS_WIDE = 240;    // one-plane sweeps of multi-pass configurations: 240 accumulators | ring of 4 * 64 | 16 = 512
constexpr int UNROLL = 12;             // chunks per unrolled block of the MMA fast path (lcm of NB and NU)

struct GroupMeta {
  int col_off;        // first digit column of this group in B
  int C;              // dot-product columns of this segment
  int n_digit_cols;   // sum of nd[0..C): the "ones" column follows them
  int n;              // complete samples
  int32_t* counts;    // [M][4]
  double* dots;       // [M][dots_stride], already offset to this segment's first dot column
  int dots_stride;
  const double* colscale;   // [C]
  const uint32_t* mask_hi;  // [ns_pad/16], both bits of each kept sample's field
  uint8_t nd[PASS_COLS_WIDE];   // base-13 digits of each dot column of the segment
};

struct Params {
  int64_t M;
  int n_tiles;
  int n_chunks;
  int ncols;          // padded to 16
  int n_gstages, n_bstages;
  int n_groups;
  int ring_base1, nu1;   // one-plane tiles: first TMEM column of the A ring (after D_c) and its units (6 or 4)
  int ring_base2;        // two-plane tiles: ring after D_c and D_m, NU2 units
  int gstage_bytes, bstage_bytes;
  int mask_bytes;     // n_groups * 128 when any group needs masking, else 0
  int one_plane_only; // wide passes: the host has checked that no row holds a missing call
  int plane_mode;     // 0: one sweep, both planes where a tile pair holds a missing call;  split sweeps (wide passes over
                      // data WITH missing calls): 1 = plane c of every tile (raw sums for flagged pairs), 2 = plane m of the
                      // flagged pairs only, which also finishes their rows from the sums of the plane-c sweep
  const uint8_t* row_flags;  // nullable
  int abl_contig;     // timing ablation: read every genotype box as one contiguous 16 KB block (results are WRONG)
  int abl;            // timing ablation bits: 1 no tcgen05.st, 2 no MMA, 4 no basis-panel loads, 8 no unpack ALU, 16 no popcount (results are WRONG)
  int abl_stream;     // timing ablation: genotype stream only -- no unpack, no basis panels, no MMA (results are WRONG)
  GroupMeta g[MAX_GROUPS];
};

// D[tmem] (+)= A[tmem] * B[smem desc], E2M1 x E2M1 -> f32, K = 64, block scales from tensor memory
template <int CS>
__device__ __forceinline__ void mma_mxf4_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t sfa,
                                            uint32_t sfb, uint32_t accumulate) {
  if (CS == 1)
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::mxf4.block_scale.scale_vec::2X [%0], [%1], %2, %3, [%5], [%6], p;\n}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(sfa), "r"(sfb)
        : "memory");
  else
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::mxf4.block_scale.scale_vec::2X [%0], [%1], %2, %3, [%5], [%6], p;\n}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(sfa), "r"(sfb)
        : "memory");
}

// block-scaled instruction descriptor: A = B = E2M1, scales UE8M0, K = 64 dense, both K-major
__device__ __forceinline__ uint32_t make_idesc(int n, int m) {
  uint32_t d = 0;
  d |= 1u << 7;                   // a_format = E2M1
  d |= 1u << 10;                  // b_format = E2M1
  d |= (uint32_t)(n >> 3) << 17;  // N
  d |= 1u << 23;                  // scale format UE8M0
  d |= (uint32_t)(m >> 4) << 24;  // M (256 for a CTA pair)
  return d;                       // scale-factor ids 0
}


struct Barriers {
  uint64_t gfull[MAX_GSTAGES];   // genotype stage filled by TMA
  uint64_t gempty[MAX_GSTAGES];  // genotype stage read out by the 16 unpack warps
  uint64_t bfull[MAX_BSTAGES];   // basis-panel stage filled by TMA (both CTAs of a pair complete on the leader's)
  uint64_t bempty[MAX_BSTAGES];  // basis-panel stage consumed (MMA commit)
  uint64_t a_full[MAX_UNITS];    // ring unit written (16 warps of each CTA)
  uint64_t a_empty[MAX_UNITS];   // ring unit consumed (MMA commit)
  uint64_t d_full;               // accumulators complete (MMA commit)
  uint64_t d_empty;              // accumulators read out (4 epilogue warps of each CTA)
  uint32_t tmem_base;
  uint32_t pad;
  int32_t n2_xchg[2][SLOTS - 1][MAX_GROUPS][TILE_M];   // see tc_kernel.cu
};

__device__ __forceinline__ bool tile_has_missing(const Params& p, int tile) {
  if (p.one_plane_only) return false;   // wide passes: the host has checked the row flags
  return tcc::tile_flags_any(p.row_flags, p.M, tile, TILE_M);
}

// NG = number of groups known at compile time (1, 2) or 0 = run-time p.n_groups; CS = 1 (one CTA per tile) or 2 (CTA
// pair, tcgen05.mma.cta_group::2, see tc_kernel.cu for the cross-CTA protocol).
template <int NG, int CS>
__global__ void __launch_bounds__(THREADS, 1)
tc4_sweep_kernel(const __grid_constant__ CUtensorMap geno_map, const __grid_constant__ CUtensorMap b_map, const Params p) {
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
  auto AEMPTY = [&](int s) { return bar0 + 8u * (2 * MAX_GSTAGES + 2 * MAX_BSTAGES + MAX_UNITS + s); };
  const uint32_t DFULL = bar0 + 8u * (2 * MAX_GSTAGES + 2 * MAX_BSTAGES + 2 * MAX_UNITS);
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
      mbar_init(BEMPTY(s), 1);
    }
    for (int s = 0; s < MAX_UNITS; ++s) {
      mbar_init(AFULL(s), UNPACK_WARPS * CS);
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
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;
  // unit scale factors: every byte of the SF columns is UE8M0 1.0 in every lane of both CTAs
  if (warp < 4) {
    uint32_t sf[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) sf[i] = 0x7F7F7F7Fu;
    tmem_st16(tmem + ((uint32_t)(warp * 32) << 16) + SF_BASE, sf);
    tmem_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  if (CS > 1) cluster_sync_all();   // peers must see initialised barriers / scale factors before any remote arrive / MMA
  tc_fence_after();

  const int cta_rank = CS > 1 ? (int)cluster_ctarank() : 0;
  const int first_tile = (CS > 1 ? (int)cluster_id_x() * CS : (int)blockIdx.x) + cta_rank;
  const int tile_step = CS > 1 ? (int)n_clusters_x() * CS : (int)gridDim.x;
  const int tile_end = CS > 1 ? p.n_tiles + cta_rank : p.n_tiles;   // (tile - rank) < n_tiles for every CTA alike
  const int panel_bytes = p.ncols / CS * 128;   // this CTA's rows of one basis panel (256 samples)
  auto pair_flag = [&](int tile) {   // does the tile (pair) hold a missing call, as far as the row flags know?
    if (CS == 1) return tile_has_missing(p, tile);
    const int t0 = tile - cta_rank;
    return tile_has_missing(p, t0) || tile_has_missing(p, t0 + 1);
  };
  // two planes in one sweep only in mode 0; the split sweeps run every tile as a one-plane tile
  auto tile_mode = [&](int tile) { return p.plane_mode == 0 && pair_flag(tile); };
  // the plane-m sweep visits the flagged pairs only (every role skips the others alike)
  auto tile_skipped = [&](int tile) { return p.plane_mode == 2 && !pair_flag(tile); };

  if (warp == WARP_TMA_G) {
    // ============================== genotype producer ==============================
    int gs = 0;
    uint32_t g_phase = 0;
    for (int tile = first_tile; tile < tile_end; tile += tile_step) {
      if (tile_skipped(tile)) continue;
      for (int ch = 0; ch < p.n_chunks; ++ch) {
        mbar_wait(GEMPTY(gs), g_phase ^ 1);
        const uint32_t sbase = smem0 + gs * p.gstage_bytes;
        if (elect_one()) {
          mbar_arrive_expect_tx(GFULL(gs), (uint32_t)(GENO_BYTES + p.mask_bytes));
          if (ABL_CONTIG()) tma_load_2d(&geno_map, GFULL(gs), sbase, 0, (tile * p.n_chunks + ch) * TILE_M);
          else tma_load_2d(&geno_map, GFULL(gs), sbase, ch * 128, tile * TILE_M);
          if (p.mask_bytes) {
            for (int g = 0; g < n_groups; ++g)
              bulk_load_1d(sbase + GENO_BYTES + g * 128, p.g[g].mask_hi + ch * (CHUNK / 16), 128, GFULL(gs));
          }
        }
        __syncwarp();
        if (++gs == p.n_gstages) { gs = 0; g_phase ^= 1; }
      }
    }
  } else if (warp == WARP_TMA_B && ABL_STREAM()) {
    // (ablation: no basis panels)
  } else if (warp == WARP_MMA && ABL_STREAM()) {
    // (ablation: no MMAs)
  } else if (warp == WARP_TMA_B) {
    // ============================== basis-panel producer ==============================
    int bs = 0;
    uint32_t b_phase = 0;
    for (int tile = first_tile; tile < tile_end; tile += tile_step) {
      if (tile_skipped(tile)) continue;
      for (int ch = 0; ch < p.n_chunks; ++ch) {
        mbar_wait(BEMPTY(bs), b_phase ^ 1);
        const uint32_t sbase = bring0 + bs * p.bstage_bytes;
        if (elect_one()) {
          if ABL(4) {
            if (cta_rank == 0) mbar_arrive(BFULL(bs));
          } else if (CS == 1) {
            mbar_arrive_expect_tx(BFULL(bs), (uint32_t)(PANELS * panel_bytes));
#pragma unroll
            for (int s = 0; s < PANELS; ++s)
              tma_load_2d(&b_map, BFULL(bs), sbase + s * panel_bytes, ch * (CHUNK / 2) + s * (PANEL / 2), 0);
          } else {
            if (cta_rank == 0) mbar_arrive_expect_tx(BFULL(bs), (uint32_t)(2 * PANELS * panel_bytes));
            const uint32_t full_leader = mapa_leader(BFULL(bs));
#pragma unroll
            for (int s = 0; s < PANELS; ++s)
              tma_load_2d_pair(&b_map, full_leader, sbase + s * panel_bytes, ch * (CHUNK / 2) + s * (PANEL / 2),
                               cta_rank * (p.ncols / 2));
          }
        }
        __syncwarp();
        if (++bs == p.n_bstages) { bs = 0; b_phase ^= 1; }
      }
    }
  } else if (warp == WARP_MMA) {
    // ============================== MMA issuer (pair: the leader only) ==============================
    if (CS == 1 || cta_rank == 0) {
      const uint32_t idesc = make_idesc(p.ncols, TILE_M * CS);
      const uint32_t sfa = tmem + SF_BASE, sfb = tmem + SF_BASE + 8;
      auto commit = [&](uint32_t bar) {
        if (CS == 1) tc_commit(bar); else tc_commit_pair(bar);
      };
      int bs = 0;
      uint32_t b_phase = 0;
      int ru = 0;               // next ring unit
      uint32_t full_par = 0;    // bit u: parity of the next completion of AFULL(u) this warp waits for
      uint32_t tile_i = 0;
      bool prev_two_plane = false;
      const uint64_t desc0 = make_kmajor_sw128_desc(bring0);
      const uint32_t stage_d = (uint32_t)p.bstage_bytes >> 4;   // descriptor-address units (16 B)
      const uint32_t panel_d = (uint32_t)panel_bytes >> 4;
      uint32_t a_ring = tmem + p.ring_base1;   // set per tile from its mode

      // the MMAs of one chunk: unit `u` (+ `u + 1` = plane m), basis stage descriptor `bd`
      auto issue = [&](auto tp_tag, const int u, const uint64_t bd, const uint32_t first_acc) {
        constexpr bool TP = decltype(tp_tag)::value;
        const uint32_t a_c = a_ring + u * UNIT_COLS;
#pragma unroll
        for (int s = 0; s < SLOTS; ++s)
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const uint32_t acc = (s | j) ? 1u : first_acc;
            const uint64_t d = bd + (uint64_t)((s >> 1) * panel_d + ((s & 1) * 2 + j) * 2);
            mma_mxf4_ts<CS>(tmem, a_c + s * SLOT_COLS + j * 8, d, idesc, sfa, sfb, acc);
            if (TP) mma_mxf4_ts<CS>(tmem + p.ncols, a_c + UNIT_COLS + s * SLOT_COLS + j * 8, d, idesc, sfa, sfb, acc);
          }
      };
      auto generic_chunk = [&](int ch, bool two_plane) {
        mbar_wait(BFULL(bs), b_phase);
        const int nu = two_plane ? NU2 : p.nu1;
        if (two_plane && (ru & 1)) { if (++ru >= nu) ru -= nu; }
        mbar_wait(AFULL(ru), (full_par >> ru) & 1u);
        full_par ^= 1u << ru;
        tc_fence_after();
        const uint64_t bd = desc0 + (uint64_t)(bs * stage_d);
        if (elect_one()) {
          if ABL(2) {
          } else if (two_plane) issue(TrueTag{}, ru, bd, ch ? 1u : 0u); else issue(FalseTag{}, ru, bd, ch ? 1u : 0u);
          commit(AEMPTY(ru));
          commit(BEMPTY(bs));
        }
        __syncwarp();
        ru += two_plane ? 2 : 1;
        if (ru >= nu) ru -= nu;
        if (++bs == p.n_bstages) { bs = 0; b_phase ^= 1; }
      };
      // fast path: UNROLL chunks with every ring position a compile-time constant (entered with bs == 0, ru == 0)
      auto fast_chunks = [&](auto tp_tag, auto nu_tag, auto nb_tag, int& ch) {
        constexpr bool TP = decltype(tp_tag)::value;
        constexpr int NU = decltype(nu_tag)::value;
        constexpr int NB = decltype(nb_tag)::value;   // basis-panel ring depth: 6, or 3 for the wide passes
        static_assert(UNROLL % NU == 0 && UNROLL % NB == 0 && (UNROLL / NB) % 2 == 0,
                      "ring positions and the basis ring's parity must repeat every UNROLL chunks");
        for (; ch + UNROLL <= p.n_chunks; ch += UNROLL) {
#pragma unroll
          for (int k = 0; k < UNROLL; ++k) {
            constexpr int dummy = 0; (void)dummy;
            const int b = k % NB;
            const int u = TP ? (2 * k) % NU : k % NU;
            mbar_wait(BFULL(b), b_phase ^ (uint32_t)((k / NB) & 1));
            mbar_wait(AFULL(u), (full_par >> u) & 1u);
            full_par ^= 1u << u;
            tc_fence_after();
            const uint64_t bd = desc0 + (uint64_t)(b * stage_d);
            if (elect_one()) {
              if (!ABL(2)) issue(tp_tag, u, bd, k ? 1u : (ch ? 1u : 0u));
              commit(AEMPTY(u));
              commit(BEMPTY(b));
            }
            __syncwarp();
          }
          // UNROLL / NB is even: the parity of the basis ring is unchanged
        }
      };

      for (int tile = first_tile; tile < tile_end; tile += tile_step) {
        if (tile_skipped(tile)) continue;
        const bool two_plane = tile_mode(tile);
        if (tile_i > 0 && two_plane != prev_two_plane) ru = 0;   // the ring is laid out per mode (the unpack warps drain first)
        prev_two_plane = two_plane;
        a_ring = tmem + (two_plane ? p.ring_base2 : p.ring_base1);
        mbar_wait(DEMPTY, (tile_i & 1) ^ 1);   // the previous tile's accumulators have been read out (by both CTAs)
        tc_fence_after();
        int ch = 0;
        if (p.n_bstages == NB || p.n_bstages == 3) {
          while (ch < p.n_chunks && (bs != 0 || ru != 0)) generic_chunk(ch++, two_plane);
          if (ch < p.n_chunks) {
            if (p.n_bstages == NB) {
              if (two_plane) fast_chunks(TrueTag{}, IntTag<NU2>{}, IntTag<NB>{}, ch);
              else if (p.nu1 == 6) fast_chunks(FalseTag{}, IntTag<6>{}, IntTag<NB>{}, ch);
              else fast_chunks(FalseTag{}, IntTag<4>{}, IntTag<NB>{}, ch);
            } else {
              if (two_plane) fast_chunks(TrueTag{}, IntTag<NU2>{}, IntTag<3>{}, ch);
              else if (p.nu1 == 6) fast_chunks(FalseTag{}, IntTag<6>{}, IntTag<3>{}, ch);
              else fast_chunks(FalseTag{}, IntTag<4>{}, IntTag<3>{}, ch);
            }
          }
        }
        for (; ch < p.n_chunks; ++ch) generic_chunk(ch, two_plane);
        if (elect_one()) commit(DFULL);
        __syncwarp();
        ++tile_i;
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
    int ru = 0;                      // next ring unit (same sequence as the MMA warp's)
    uint32_t empty_par = 0;          // bit u: parity of the last completion of AEMPTY(u) this warp relies on
    uint32_t tile_i = 0;
    bool prev_two_plane = false;
    uint32_t a_slot = tmem + lane_addr + p.ring_base1 + s * SLOT_COLS;   // set per tile from its mode
    int n2[NG ? NG : MAX_GROUPS];
    const uint32_t afull_leader = (CS == 2) ? mapa_leader(AFULL(0)) : 0u;
    const uint32_t dempty_leader = (CS == 2) ? mapa_leader(DEMPTY) : 0u;
    auto arrive_afull = [&](int u) {
      if (CS == 1) mbar_arrive(AFULL(u)); else mbar_arrive_cluster(afull_leader + 8u * (uint32_t)u);
    };

    // The chunk loop of one tile, specialised on the tile's mode (TP: two planes) and on whether any group needs its
    // sample mask for the hom-alt count (MA: none does).
    // MO: the plane-m sweep of a split pass -- the A operand is the missing indicator alone, nothing is counted.
    auto run_chunks = [&](auto tp_tag, auto ma_tag, auto mo_tag) {
      constexpr bool TP = decltype(tp_tag)::value;
      constexpr bool MO = decltype(mo_tag)::value;
      constexpr bool MA = decltype(ma_tag)::value || MO;
      constexpr bool EAGER = TP && (LRR_TC4_TP_EAGER != 0);
      static_assert(!(TP && MO), "the plane-m sweep runs one-plane tiles");
      int pend = -1;    // unit whose TMEM store is issued but not yet published to the MMA warp
      auto chunk_body = [&](const int u, const int pend_u) {
        mbar_wait(gbar, g_phase);
        const uint4 w0 = lds128(gaddr + ld0);   // packed bytes of samples [128 s, 128 s + 128) of this row
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
        // 2-bit fields -> nibbles 00cc: word i gives column 2i (fields 0 / 2 of every byte in the low / high nibble)
        // and column 2i + 1 (fields 1 / 3); the basis panels are stored in the same element order (quantize_kernel)
        uint32_t rc[16];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if ABL(8) {
            rc[2 * i + 0] = w[i];
            rc[2 * i + 1] = w[i];
          } else if (MO) {
            const uint32_t mw = w[i] & (w[i] >> 1);   // bit 2f set iff field f is code 3: nibble 0001 (= 0.5)
            rc[2 * i + 0] = mw & 0x11111111u;
            rc[2 * i + 1] = (mw >> 2) & 0x11111111u;
          } else {
            rc[2 * i + 0] = w[i] & 0x33333333u;
            rc[2 * i + 1] = (w[i] >> 2) & 0x33333333u;
          }
        }
        // exact counts for x.x = n1 + 4 n2: the number of set bits among the group's fields, T = n1 + n2 + 2 n_miss (one
        // POPC per word, one AND more under a sample mask); with S = n1 + 2 n2 and n_miss from the "ones" column of the
        // two planes the epilogue solves n2 = S - (T - 2 n_miss), n1 = S - 2 n2
        auto count_bits = [&]() {
#pragma unroll
          for (int g = 0; g < (NG ? NG : MAX_GROUPS); ++g) {
            if ((NG || g < n_groups) && !ABL(16) && !MO) {
              int acc = 0;
#pragma unroll
              for (int i = 0; i < 8; ++i) acc += MA ? __popc(w[i]) : __popc(w[i] & mm[g][i]);
              n2[g] += acc;
            }
          }
        };
        if (!EAGER) count_bits();   // (two-plane tiles count behind their TMEM stores, see below)
        __syncwarp();
        if (lane == 0) mbar_arrive(gbar + 8u * MAX_GSTAGES);   // genotype stage back to the TMA producer
        if (pend_u >= 0 && !EAGER) {   // retire the previous chunk's TMEM store behind this chunk's arithmetic
          tmem_wait_st();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) arrive_afull(pend_u);
        }
        mbar_wait(AEMPTY(u), ((empty_par >> u) & 1u) ^ 1u);
        empty_par ^= 1u << u;
        tc_fence_after();
        if (!ABL(1)) tmem_st16(a_slot + u * UNIT_COLS, rc);
        else if ((rc[0] ^ rc[5] ^ rc[10] ^ rc[15]) == 0x12345678u) n2[0] += 1;   // keep the arithmetic alive
        if (TP) {
          // missing-indicator plane: nibble 0001 (= 0.5) where the call is code 3
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const uint32_t mw = w[i] & (w[i] >> 1);   // bit 2f set iff field f is code 3
            rc[2 * i + 0] = mw & 0x11111111u;
            rc[2 * i + 1] = (mw >> 2) & 0x11111111u;
          }
          tmem_st16(a_slot + (u + 1) * UNIT_COLS, rc);
        }
        if (EAGER) {
          // two-plane tiles: the ring holds only two chunks (4 units), so a store published one chunk late would leave the
          // MMA warp nothing to overlap with -- publish at once (35.4 -> 32.4 ms per C3 pass), with the population counts
          // in the shadow of the stores
          count_bits();
          tmem_wait_st();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) arrive_afull(u);
        }
        gaddr += p.gstage_bytes;
        gbar += 8u;
        if (gaddr == gaddr_end) { gaddr = smem0; gbar = GFULL(0); g_phase ^= 1; }
      };
      const int nu = TP ? NU2 : p.nu1;
      auto generic_chunk = [&]() {
        if (TP && (ru & 1)) { if (++ru >= nu) ru -= nu; }
        chunk_body(ru, pend);
        pend = ru;
        ru += TP ? 2 : 1;
        if (ru >= nu) ru -= nu;
      };
      int ch = 0;
      // steady state unrolled over the ring (positions become constants): entered with ru == 0
      do { generic_chunk(); ++ch; } while (ch < p.n_chunks && ru != 0);
      if (!TP && nu == 6) {
        for (; ch + 6 <= p.n_chunks; ch += 6) {   // pend == 5 here
          chunk_body(0, pend);
          chunk_body(1, 0);
          chunk_body(2, 1);
          chunk_body(3, 2);
          chunk_body(4, 3);
          chunk_body(5, 4);
          pend = 5;
        }
      } else if (!TP) {
        for (; ch + 4 <= p.n_chunks; ch += 4) {   // pend == 3 here
          chunk_body(0, pend);
          chunk_body(1, 0);
          chunk_body(2, 1);
          chunk_body(3, 2);
          pend = 3;
        }
      } else {
        for (; ch + 2 <= p.n_chunks; ch += 2) {   // pend == 2 here
          chunk_body(0, pend);
          chunk_body(2, 0);
          pend = 2;
        }
      }
      for (; ch < p.n_chunks; ++ch) generic_chunk();
      if (pend >= 0 && !EAGER) {   // flush the last chunk of the tile
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) arrive_afull(pend);
      }
    };

    if (ABL_STREAM()) {
      // ablation: consume the genotype stages and nothing else (abl_stream == 2: also LDS the thread's 32 bytes)
      uint32_t sink = 0;
      for (int tile = first_tile; tile < tile_end; tile += tile_step)
        for (int ch = 0; ch < p.n_chunks; ++ch) {
          mbar_wait(gbar, g_phase);
          if (ABL_STREAM() == 2) { const uint4 w0 = lds128(gaddr + ld0); const uint4 w1 = lds128(gaddr + ld1); sink ^= w0.x ^ w1.w; }
          __syncwarp();
          if (lane == 0) mbar_arrive(gbar + 8u * MAX_GSTAGES);
          gaddr += p.gstage_bytes;
          gbar += 8u;
          if (gaddr == gaddr_end) { gaddr = smem0; gbar = GFULL(0); g_phase ^= 1; }
        }
      if (sink == 0x12345u) bars->pad = sink;
    } else
    for (int tile = first_tile; tile < tile_end; tile += tile_step) {
      if (tile_skipped(tile)) continue;
      const bool two_plane = tile_mode(tile);
      if (tile_i > 0 && two_plane != prev_two_plane) {
        // the ring is laid out differently: wait until every MMA of the previous tile has retired
        mbar_wait(DFULL, (tile_i - 1) & 1);
        tc_fence_after();
        // ... and until the epilogue warps have read the previous tile's accumulators: the one-plane ring reuses the
        // plane-m accumulator columns (the epilogue warps arrive here after their tcgen05.ld, in program order)
        named_bar_sync(2, UNPACK_WARPS * 32);
        ru = 0;
      }
      prev_two_plane = two_plane;
      a_slot = tmem + lane_addr + (two_plane ? p.ring_base2 : p.ring_base1) + s * SLOT_COLS;
#pragma unroll
      for (int g = 0; g < (NG ? NG : MAX_GROUPS); ++g) n2[g] = 0;
      if (p.plane_mode == 2) {
        run_chunks(FalseTag{}, TrueTag{}, TrueTag{});
      } else if (two_plane) {
        if (p.mask_bytes == 0) run_chunks(TrueTag{}, TrueTag{}, FalseTag{}); else run_chunks(TrueTag{}, FalseTag{}, FalseTag{});
      } else {
        if (p.mask_bytes == 0) run_chunks(FalseTag{}, TrueTag{}, FalseTag{}); else run_chunks(FalseTag{}, FalseTag{}, FalseTag{});
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
        // split sweeps: the plane-c sweep leaves RAW sums for the flagged pairs (sum of codes, set bits, D_c); the
        // plane-m sweep, which visits exactly those pairs, finishes them: counts, mean, (D_c - 3 D_m) + mean D_m
        const bool raw_c = p.plane_mode == 1 && pair_flag(tile);
        const bool m_only = p.plane_mode == 2;
#pragma unroll
        for (int g = 0; g < (NG ? NG : MAX_GROUPS); ++g) {
          if (!(NG || g < n_groups)) continue;
          const GroupMeta& G = p.g[g];
          const int cnt = n2[g] + bars->n2_xchg[tile_i & 1][0][g][row] + bars->n2_xchg[tile_i & 1][1][g][row] +
                          bars->n2_xchg[tile_i & 1][2][g][row];
          // the group's columns: nd[c] digit columns per dot column c, then one "ones" column (digit value 1.0)
          const int ones_col = G.col_off + G.n_digit_cols;
          uint32_t r16[16];
          tmem_ld16(d_c + (ones_col & ~15), r16);
          tmem_wait_ld();
          float sc_f = 0.f;
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (i == (ones_col & 15)) sc_f = __uint_as_float(r16[i]);
          float nm_f = 0.f;
          if (two_plane) {
            tmem_ld16(d_m + (ones_col & ~15), r16);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (i == (ones_col & 15)) nm_f = __uint_as_float(r16[i]);
          }
          int sc = __float2int_rn(2.f * sc_f);         // sum of codes: A = c / 2, ones digit = 1.0
          int nm = __float2int_rn(2.f * nm_f);         // missing calls: A = 1 / 2
          int bits = cnt;                              // set bits among the group's fields = n1 + n2 + 2 n_miss
          if (m_only) {
            nm = sc;                                   // this sweep's accumulators hold the indicator plane
            const int4 raw = v < p.M ? reinterpret_cast<const int4*>(G.counts)[v] : make_int4(0, 0, 0, 0);
            sc = raw.x;
            bits = raw.y;
          }
          const int S = sc - 3 * nm;                   // n1 + 2 n2
          const int n2g = S - (bits - 2 * nm);
          const int n1 = S - 2 * n2g;
          const double mean = (double)S / (double)(G.n - nm);
          if (v < p.M) reinterpret_cast<int4*>(G.counts)[v] = raw_c ? make_int4(sc, cnt, -1, 0) : make_int4(n1, n2g, nm, 0);
          // digit columns, 16 TMEM columns at a time: accumulator = sum c u / 4 with u the base-13 digit
          const int c_lo = G.col_off, c_hi = ones_col;
          long long hi = 0, lo = 0, mhi = 0, mlo = 0;
          long long pw = 1;   // 13^sl (sl < 6) or 13^(sl - 6)
          int c = 0, sl = 0, nd = G.nd[0];
          for (int base = c_lo & ~15; base < c_hi; base += 16) {
            uint32_t dc[16], dm[16];
            tmem_ld16(d_c + base, dc);
            if (two_plane) tmem_ld16(d_m + base, dm);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int col = base + i;
              if (col >= c_lo && col < c_hi) {
                const long long cv = (long long)__float2int_rn(4.f * __uint_as_float(dc[i]));
                const long long mv = two_plane ? (long long)__float2int_rn(4.f * __uint_as_float(dm[i])) : 0ll;
                const long long dv = cv - 3ll * mv;
                if (sl < 6) {
                  lo += dv * pw;
                  mlo += mv * pw;
                } else {
                  hi += dv * pw;
                  mhi += mv * pw;
                }
                pw *= 13;
                if (++sl == 6) pw = 1;
                if (sl == nd) {
                  const double scale = G.colscale[c];
                  double dot = fma((double)hi, 4826809.0, (double)lo) * scale;
                  if (two_plane && nm > 0) dot += mean * (fma((double)mhi, 4826809.0, (double)mlo) * scale);
                  if (m_only && v < p.M) {   // `dot` is D_m here; the plane-c sweep left D_c
                    const double d_raw = G.dots[v * G.dots_stride + c];
                    dot = nm > 0 ? (d_raw - 3.0 * dot) + mean * dot : d_raw;
                  }
                  if (v < p.M) G.dots[v * G.dots_stride + c] = dot;
                  hi = lo = mhi = mlo = 0;
                  sl = 0;
                  pw = 1;
                  ++c;
                  nd = c < G.C ? G.nd[c] : MAX_DIGITS;
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
      ++tile_i;
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
// basis quantisation: float64 column -> balanced base-13 digits as E2M1 nibbles, in the A operand's element order
// ------------------------------------------------------------------------------------------------
// max |v| and sum v^2 of one column per CTA (fixed reduction order: the digit policy below must not depend on atomics)
__global__ void __launch_bounds__(1024) colstat_kernel(const double* __restrict__ cols, int64_t ns_pad, double* __restrict__ colmax_out,
                                                       double* __restrict__ sumsq_out) {
  __shared__ double s_m[32], s_q[32];
  const double* col = cols + (int64_t)blockIdx.x * ns_pad;   // one CTA per column
  double* colmax = colmax_out + blockIdx.x;
  double* sumsq = sumsq_out + blockIdx.x;
  double m = 0.0, q = 0.0;
  for (int64_t j = threadIdx.x; j < ns_pad; j += blockDim.x) {
    const double v = col[j];
    m = fmax(m, fabs(v));
    q = fma(v, v, q);
  }
  for (int o = 16; o > 0; o >>= 1) {
    m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    q += __shfl_xor_sync(0xffffffffu, q, o);
  }
  if ((threadIdx.x & 31) == 0) { s_m[threadIdx.x >> 5] = m; s_q[threadIdx.x >> 5] = q; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) { m = fmax(m, s_m[w]); q += s_q[w]; }
    *colmax = m;
    *sumsq = q;
  }
}

// the "fitted value" column of phenotype p: fit[j] = sum_c q_c[j] * Qty[c + has_intercept][p] over the Kd dot-product
// covariate columns, so that y_transpose_x = xyp + Qty[0][p] sum_x / sqrt(n) + fit . x (LR:143-146) needs no
// high-precision covariate projections
__global__ void fitted_kernel(const double* __restrict__ basis, const double* __restrict__ qty, int Kd, int P, int has_intercept,
                              int p, int64_t ns_pad, double* __restrict__ fit) {
  for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < ns_pad; j += (int64_t)gridDim.x * blockDim.x) {
    double f = 0.0;
    for (int c = 0; c < Kd; ++c) f = fma(basis[(int64_t)c * ns_pad + j], qty[(c + has_intercept) * P + p], f);
    fit[j] = f;
  }
}

// digit of residue r = I mod 13, chosen from {0, +-1, +-2, +-3, +-4, +-6, +-8} (complete residue system mod 13)
__device__ __forceinline__ int digit13(long long& I) {
  const int r = (int)(((I % 13) + 13) % 13);
  const int d = (r <= 4) ? r : (r == 5 ? -8 : (r == 6 ? 6 : (r == 7 ? -6 : (r == 8 ? 8 : r - 13))));
  I = (I - d) / 13;
  return d;
}
// E2M1 code of the value u / 2 for a digit u
__device__ __forceinline__ uint32_t e2m1_code(int u) {
  const int a = u < 0 ? -u : u;
  const uint32_t m = a <= 4 ? (uint32_t)a : (a == 6 ? 5u : 6u);   // 0, .5, 1, 1.5, 2 | 3 | 4
  return m | (u < 0 ? 8u : 0u);
}
__host__ __device__ inline double imax13(int nd) {   // floor(13^nd / 3): every |I| up to it has nd digits
  double p13 = 1.0;
  for (int i = 0; i < nd; ++i) p13 *= 13.0;
  return floor(p13 / 3.0);
}

// One column -> its nd digit rows.  One thread per byte of a panel row.  Byte b of a row covers word wd = b / 8 of the
// packed genotype row (16 samples): with r = (b % 8) / 4 and i = b % 4 its low nibble is sample 16 wd + 4 r + i, its high
// nibble sample 16 wd + 4 (r + 2) + i -- the order in which the unpack warps emit the calls of a word.
// The stored value of sample j is I_j * colscale with I_j = rn(v_j / colscale), |I_j| <= imax: |v_j - I_j colscale| <=
// colscale / 2 (plus 2^-52 |v_j| of the division), the per-sample quantum the statistics epilogue bounds.
// It also accumulates errfx += sum_j (v_j - I_j colscale) / colscale in 2^-40 fixed point (integer atomics: the sum does
// not depend on the order), the column's TOTAL rounding error: sum_j e_j x_j = a sum_j e_j + sum_j e_j (x_j - a) for any
// constant a, so the statistics epilogue adds a * sum_j e_j to the dot product and bounds the rest by
// (colscale / 2) sum_j |x_j - a| with a = the row's most frequent call (stats_device.cuh).
__global__ void quantize_kernel(const double* __restrict__ col, int nd, int64_t ns_pad, const double* __restrict__ colmax,
                                int first_row, uint8_t* __restrict__ bq, double* __restrict__ colscale,
                                unsigned long long* __restrict__ errfx) {
  const double imax = imax13(nd);
  const int64_t row_bytes = ns_pad / 2;
  const double cm = *colmax;
  const double cs = cm > 0.0 ? cm / imax : 0.0;
  long long err = 0;
  for (int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; b < row_bytes; b += (int64_t)gridDim.x * blockDim.x) {
    const int64_t wd = b >> 3;
    const int r = (int)((b >> 2) & 1), i = (int)(b & 3);
    const int64_t j_lo = 16 * wd + 4 * r + i, j_hi = j_lo + 8;
    long long I_lo = 0, I_hi = 0;
    if (cm > 0.0) {
      const double v_lo = col[j_lo], v_hi = col[j_hi];
      I_lo = __double2ll_rn(v_lo / cm * imax);
      I_hi = __double2ll_rn(v_hi / cm * imax);
      // v - I cs with one rounding (fma), in units of cs: within [-1/2, 1/2] up to the roundoff of the division above
      err += __double2ll_rn(fma(-(double)I_lo, cs, v_lo) / cs * 1099511627776.0);
      err += __double2ll_rn(fma(-(double)I_hi, cs, v_hi) / cs * 1099511627776.0);
    }
    for (int s = 0; s < nd; ++s) {
      const uint32_t lo = e2m1_code(digit13(I_lo)), hi = e2m1_code(digit13(I_hi));
      bq[(int64_t)(first_row + s) * row_bytes + b] = (uint8_t)(lo | (hi << 4));
    }
    if (b == 0) *colscale = cs;
  }
  for (int o = 16; o > 0; o >>= 1) err += __shfl_xor_sync(0xffffffffu, err, o);
  if ((threadIdx.x & 31) == 0 && err != 0) atomicAdd(errfx, (unsigned long long)err);
}

// errsum[k] = (fixed-point total of column k) * 2^-40 * colscale[k]
__global__ void errsum_kernel(const unsigned long long* __restrict__ errfx, const double* __restrict__ colscale, int n,
                              double* __restrict__ errsum) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) errsum[k] = (double)(long long)errfx[k] * (1.0 / 1099511627776.0) * colscale[k];
}

// the "ones" row of a segment: digit value 1.0 for the samples of the group
__global__ void ones_row_kernel(const uint32_t* __restrict__ mask, int64_t ns_pad, int row, uint8_t* __restrict__ bq) {
  const int64_t row_bytes = ns_pad / 2;
  for (int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; b < row_bytes; b += (int64_t)gridDim.x * blockDim.x) {
    const int64_t wd = b >> 3;
    const int r = (int)((b >> 2) & 1), i = (int)(b & 3);
    const int64_t j_lo = 16 * wd + 4 * r + i, j_hi = j_lo + 8;
    const uint32_t mw = mask[wd];
    const uint32_t in_lo = (mw >> sample_shift((int)(j_lo & 15))) & 1u, in_hi = (mw >> sample_shift((int)(j_hi & 15))) & 1u;
    bq[(int64_t)row * row_bytes + b] = (uint8_t)((in_lo ? 2u : 0u) | ((in_hi ? 2u : 0u) << 4));
  }
}

// Worst case of |sum_j c_j u_j| over all genotypes (c <= 3) for one panel row; the f32 accumulator holds sum c u / 4
// exactly while this stays <= 2^24.
__global__ void acc_bound_kernel(const uint8_t* __restrict__ bq, int64_t row_bytes, unsigned long long* __restrict__ bound) {
  const int64_t r = blockIdx.y;
  unsigned long long pos = 0, neg = 0;
  const uint4* row = reinterpret_cast<const uint4*>(bq + r * row_bytes);   // row_bytes is a multiple of 256
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < row_bytes / 16; q += (int64_t)gridDim.x * blockDim.x) {
    const uint4 v = __ldg(row + q);
    const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
    uint32_t p32 = 0, n32 = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int h = 0; h < 8; ++h) {
        const uint32_t code = (w4[k] >> (4 * h)) & 15u;
        const uint32_t m = code & 7u;
        const uint32_t u = m <= 4 ? m : (m == 5 ? 6u : (m == 6 ? 8u : 12u));
        if (code & 8u) n32 += 3u * u; else p32 += 3u * u;
      }
    pos += p32;
    neg += n32;
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

struct Segment {
  int group;     // index into Ctx::groups
  int c_first;   // first (extended) dot column of the group in this segment
  int n_cols;    // dot columns in this segment
  int n_digit_cols;   // their digit rows
  int row0;      // first row of this segment in the pass's panel matrix (digit rows, then the ones row)
};

struct PassShape {
  int n_gstages = 0, n_bstages = 0, bstage_bytes = 0, smem_bytes = 0;
  CUtensorMap b_map;
};

struct Pass {
  int ncols = 0;                 // panel rows, padded to 16
  int64_t bq_row0 = 0;           // first row of this pass in State::d_bq
  std::vector<Segment> segs;
  int gstage_bytes = 0;
  int ring_base1 = 0, nu1 = 0, ring_base2 = 0, mask_bytes = 0;
  PassShape shape[2];            // [cluster size - 1]
};

// Per group: the dot columns the sweep produces.  Columns [0, C) are the group's basis planes ([Kd covariate | P
// residualised phenotype] columns); columns [C, C + n_fit) are fitted-value columns (one per phenotype, P <= 2 only).
struct GroupCols {
  int n_fit = 0;
  double* d_fit = nullptr;       // [n_fit][ns_pad]
  std::vector<uint8_t> nd;       // [C + n_fit] base-13 digits per column
};

struct State {
  bool prepared = false;
  bool usable = false;
  bool wide = false;             // the prepared plan uses wide one-plane-only passes
  unsigned long long* d_bound = nullptr;   // scratch of acc_bound_kernel (grow-only)
  size_t bound_bytes = 0;
  int32_t* d_any = nullptr;      // scratch of any_flag_kernel
  int32_t* h_any = nullptr;      // page-locked
  std::string why;
  EncodeTiledFn encode = nullptr;
  std::vector<Pass> passes;
  std::vector<GroupCols> cols;
  uint8_t* d_bq = nullptr;
  double* d_colscale = nullptr;   // [2 * nscale]: quantum of every dot column, then the sum of its rounding errors
  int nscale = 0;
  unsigned long long* d_errfx = nullptr;   // [nscale] fixed-point accumulators of quantize_kernel
  double* d_colstat = nullptr;   // [2][nscale]: column maxima, sums of squares
  uint32_t* d_mask_hi = nullptr;   // per group: both bits of every kept field
  std::vector<int> scale_off;
  int cluster = 2;
  bool attr_set = false;
  // The buffers above live in grow-only pools that survive free_prepared: lrr_clear_groups runs at the start of every public
  // call, and a cudaFree there is a device synchronisation plus a trip through the kernel driver that took 20 - 570 ms in one
  // call out of four on a shared box (profiles/r02b_e2e_clear_groups_trace.txt).  Reuse is ordered like the group buffers:
  // prepare() writes them on the default stream behind lrr_add_group, which waits for the runs still in flight (busy_ev).
  struct Pool {
    void* p = nullptr;
    size_t cap = 0;
  };
  Pool pool_bq, pool_colscale, pool_errfx, pool_colstat, pool_mask, pool_fit;
};

static int pool_take(Ctx* c, State::Pool& pl, size_t need, void** out) {
  if (need > pl.cap) {
    cudaFree(pl.p);
    pl.p = nullptr;
    pl.cap = 0;
    LRR_CUDA(c, cudaMalloc(&pl.p, need));
    pl.cap = need;
  }
  *out = pl.p;
  return LRR_OK;
}

static State* state(Ctx* c) {
  if (!c->tc4_state) c->tc4_state = new State();
  return static_cast<State*>(c->tc4_state);
}

static void free_prepared(State* s) {   // forgets the plan; the device buffers stay in their pools
  s->d_errfx = nullptr;
  s->cols.clear();
  s->d_mask_hi = nullptr;
  s->d_bq = nullptr;
  s->d_colscale = nullptr;
  s->d_colstat = nullptr;
  s->passes.clear();
  s->prepared = false;
  s->usable = false;
}

static void free_pools(State* s) {
  for (State::Pool* pl : {&s->pool_bq, &s->pool_colscale, &s->pool_errfx, &s->pool_colstat, &s->pool_mask, &s->pool_fit}) {
    cudaFree(pl->p);
    pl->p = nullptr;
    pl->cap = 0;
  }
}

static int encode_2d(State* s, CUtensorMap* map, const void* ptr, uint64_t inner, uint64_t outer, uint64_t row_stride,
                     uint32_t box_inner, uint32_t box_outer) {
  return tcc::encode_2d_u8(s->encode, map, ptr, inner, outer, row_stride, box_inner, box_outer);
}

// ------------------------------------------------------------------------------------------------
// Digit policy.  The quantisation error of a dot product is at most (colscale / 2) * sum_j |x_j| for ANY genotype row
// (stats_device.cuh checks exactly that per variant), colscale = colmax / floor(13^nd / 3).  The digit counts below
// keep that bound inside the tolerance for well-scaled columns at any allele frequency (DESIGN.md 5.1 "precision"):
//   strict profile (P <= 2)   residualised phenotype 13 digits (|dt| <= 1e-10), covariate columns 6 digits (they only
//                             enter x.x - |Q'x|^2, relative error ~1e-7), and y_transpose_x comes from one 11-digit
//                             fitted-value column per phenotype instead of the covariate projections;
//   wide profile (P > 2)      phenotype columns 10 digits (|dt| <= 1.5e-7 worst case, ~1e-10 typical: the "stated
//                             tolerance" of the dense-contraction configuration), covariate columns 12 digits (shared by
//                             all phenotypes; they carry y_transpose_x: error bound ~1e-8 in dot-product units, far inside
//                             that profile's 5e-7 floor.  12 rather than 13 is what lets BASELINE's 128 phenotypes + 9
//                             covariates fill exactly six 240-column passes: 1 + 9 * 12 + 13 * 10 and 5 x (1 + 23 * 10)).
// A column whose maximum is far above 4.5 rms (heavy tails, outliers) gets one more digit per factor 13, up to 13;
// `boost` (Ctx::digit_boost, raised when too many rows had to be recomputed) adds digits to the covariate / fitted columns.
// ------------------------------------------------------------------------------------------------
static int tail_digits(double colmax, double sumsq, int64_t n, double slack) {
  if (!(colmax > 0.0) || !(sumsq > 0.0) || n <= 0) return 0;
  const double ratio = colmax / (slack * 4.5 * sqrt(sumsq / (double)n));
  if (!(ratio > 1.0)) return 0;
  return (int)ceil(log(ratio) / log(13.0) - 1e-9);
}

// `slack`: how far above 4.5 rms a column's maximum may lie before it costs a digit.  The strict profile's digit counts were
// sized for exactly that scale with the bound (quantum / 2) sum_j x_j; the centred bound of the statistics epilogue
// ((quantum / 2) sum_j |x_j - ac|, stats_device.cuh) is 1.57 x tighter at its own worst allele frequency (0.3) than the
// old one at its design point (0.5) -- relative to x.x - |Q'x|^2 ~ 2 p (1 - p) n -- and no longer grows towards p = 1, so a
// factor 1.5 is free (a Gaussian column of 400k samples peaks at ~5 rms = 1.1 x).  The wide profile's bounds leave room under its 5e-7 floor: a 10-digit phenotype
// column is bounded by |dt| <= 1.5e-7 at 4.5 rms, so up to 3 x that scale stays inside; the 12-digit covariate columns
// (~1e-8 in dot-product units) have a factor 13 = one whole digit.  A Gaussian column of 400k samples peaks at ~5 rms --
// without the slack every such column paid an eleventh digit (1,526 instead of 1,394 digit columns for BASELINE's
// 128-phenotype configuration: 7 passes instead of 6).  The per-variant guard enforces the tolerance either way.
static void digit_policy(const Group& gr, int n_fit, const double* colmax, const double* sumsq, int boost, std::vector<uint8_t>& nd) {
  const bool wide = gr.P > 2;
  int wide_cov_digits = 12;
  double y_slack = wide ? 3.0 : 1.5, cov_slack = wide ? 13.0 : 1.5;
  if (const char* e = tuning_env("LRR_TC4_WIDE_COVD")) { const int v = atoi(e); if (v >= 6 && v <= MAX_DIGITS) wide_cov_digits = v; }
  if (const char* e = tuning_env("LRR_TC4_SLACK")) { if (atof(e) >= 1.0) { y_slack = atof(e); cov_slack = wide ? 13.0 * atof(e) / 3.0 : atof(e); } }
  const int Cx = gr.C + n_fit;
  nd.resize(Cx);
  for (int c = 0; c < Cx; ++c) {
    int base;
    double slack;
    if (c < gr.Kd) { base = wide ? wide_cov_digits : (n_fit == 0 ? 13 : 6 + boost); slack = cov_slack; }
    else if (c < gr.C) { base = wide ? 10 + boost : 13; slack = y_slack; }
    else { base = 11 + boost; slack = y_slack; }
    const int d = base + tail_digits(colmax[c], sumsq[c], gr.n, slack);
    nd[c] = (uint8_t)std::min(MAX_DIGITS, std::max(1, d));
  }
}

// split every group's dot columns into passes of at most PASS_COLS panel rows
static void plan_passes(const Ctx* c, const std::vector<GroupCols>& cols, std::vector<Pass>& passes, int pass_cols) {
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
    const std::vector<uint8_t>& nd = cols[g].nd;
    const int Cx = (int)nd.size();
    int col = 0;
    while (col < Cx) {
      if (used + nd[col] + 1 > pass_cols || (int)cur.segs.size() == MAX_GROUPS) close();
      Segment sg;
      sg.group = (int)g;
      sg.c_first = col;
      sg.n_cols = 0;
      sg.row0 = used;
      int rows = 0;
      while (col < Cx) {
        if (used + rows + nd[col] + 1 > pass_cols) break;
        rows += nd[col];
        sg.n_cols++;
        col++;
      }
      sg.n_digit_cols = rows;
      used += rows + 1;   // + ones row
      cur.segs.push_back(sg);
      if (col < Cx) close();
    }
  }
  close();
}

// both bits of every kept sample's field (Group::d_mask has the low bit): what the unpack warps AND the packed words with
__global__ void mask_full_kernel(const uint32_t* __restrict__ mask_lo, int64_t words, uint32_t* __restrict__ mask_full) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < words; i += (int64_t)gridDim.x * blockDim.x)
    mask_full[i] = mask_lo[i] | (mask_lo[i] << 1);
}

__global__ void any_flag_kernel(const uint8_t* __restrict__ flags, int64_t M, int32_t* __restrict__ any) {
  int f = 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < M; i += (int64_t)gridDim.x * blockDim.x) f |= flags[i];
  if (__any_sync(0xffffffffu, f != 0) && (threadIdx.x & 31) == 0) atomicOr(any, 1);
}

static int prepare(Ctx* c, bool wide) {
  State* s = state(c);
  if (s->prepared && s->wide == wide) return LRR_OK;
  free_prepared(s);
  s->prepared = true;
  s->wide = wide;
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
  const int64_t ns_pad = c->groups[0].ns_pad;
  const int64_t row_bytes = ns_pad / 2;
  const unsigned gx = (unsigned)std::min<int64_t>((ns_pad + 255) / 256, 1024);
  const unsigned gxb = (unsigned)std::min<int64_t>((row_bytes + 255) / 256, 1024);
  // ---- fitted-value columns, column statistics, digit policy ----
  int nscale = 0;
  s->scale_off.assign(G, 0);
  s->cols.assign(G, GroupCols());
  for (size_t g = 0; g < G; ++g) {
    const Group& gr = c->groups[g];
    GroupCols& gc = s->cols[g];
    gc.n_fit = (gr.P <= 2 && gr.Kd > 0) ? gr.P : 0;   // the spare two slots of a dots row hold their dot products
    s->scale_off[g] = nscale;
    nscale += gr.C + gc.n_fit;
  }
  {
    size_t fit_cols = 0;
    for (size_t g = 0; g < G; ++g) fit_cols += (size_t)s->cols[g].n_fit;
    void* fit_base = nullptr;
    if (fit_cols)
      if (int r = pool_take(c, s->pool_fit, sizeof(double) * fit_cols * (size_t)ns_pad, &fit_base)) return r;
    size_t at = 0;
    for (size_t g = 0; g < G; ++g) {
      const Group& gr = c->groups[g];
      GroupCols& gc = s->cols[g];
      if (!gc.n_fit) continue;
      gc.d_fit = static_cast<double*>(fit_base) + at * (size_t)ns_pad;
      at += (size_t)gc.n_fit;
      for (int p = 0; p < gc.n_fit; ++p)
        fitted_kernel<<<gx, 256>>>(gr.d_basis, gr.d_qty, gr.Kd, gr.P, gr.has_intercept, p, ns_pad, gc.d_fit + (int64_t)p * ns_pad);
      c->launches += gc.n_fit;
    }
  }
  auto column_ptr = [&](size_t g, int col) -> const double* {
    const Group& gr = c->groups[g];
    return col < gr.C ? gr.d_basis + (int64_t)col * ns_pad : s->cols[g].d_fit + (int64_t)(col - gr.C) * ns_pad;
  };
  if (int r = pool_take(c, s->pool_colstat, sizeof(double) * 2 * (size_t)nscale, reinterpret_cast<void**>(&s->d_colstat))) return r;
  if (int r = pool_take(c, s->pool_colscale, sizeof(double) * 2 * (size_t)nscale, reinterpret_cast<void**>(&s->d_colscale))) return r;
  if (int r = pool_take(c, s->pool_errfx, sizeof(unsigned long long) * (size_t)nscale, reinterpret_cast<void**>(&s->d_errfx))) return r;
  LRR_CUDA(c, cudaMemset(s->d_errfx, 0, sizeof(unsigned long long) * (size_t)nscale));
  s->nscale = (int)nscale;
  for (size_t g = 0; g < G; ++g) {
    const Group& gr = c->groups[g];
    const int k = s->scale_off[g];
    colstat_kernel<<<(unsigned)gr.C, 1024>>>(gr.d_basis, ns_pad, s->d_colstat + k, s->d_colstat + nscale + k);
    c->launches++;
    if (s->cols[g].n_fit) {
      colstat_kernel<<<(unsigned)s->cols[g].n_fit, 1024>>>(s->cols[g].d_fit, ns_pad, s->d_colstat + k + gr.C, s->d_colstat + nscale + k + gr.C);
      c->launches++;
    }
  }
  LRR_CUDA(c, cudaGetLastError());
  std::vector<double> h_stat(2 * (size_t)nscale);
  LRR_CUDA(c, cudaMemcpy(h_stat.data(), s->d_colstat, sizeof(double) * h_stat.size(), cudaMemcpyDeviceToHost));
  for (size_t g = 0; g < G; ++g)
    digit_policy(c->groups[g], s->cols[g].n_fit, h_stat.data() + s->scale_off[g], h_stat.data() + nscale + s->scale_off[g],
                 c->digit_boost, s->cols[g].nd);
  int wide_cols = PASS_COLS_WIDE;
  if (const char* e = tuning_env("LRR_TC4_WIDE")) { const int v = atoi(e); if (v >= PASS_COLS && v <= PASS_COLS_WIDE) wide_cols = v / 16 * 16; }
  plan_passes(c, s->cols, s->passes, wide ? wide_cols : PASS_COLS);
  int64_t total_rows = 0;
  for (auto& ps : s->passes) {
    ps.bq_row0 = total_rows;
    total_rows += ps.ncols;
  }
  if (int r = pool_take(c, s->pool_bq, (size_t)total_rows * row_bytes, reinterpret_cast<void**>(&s->d_bq))) return r;
  LRR_CUDA(c, cudaMemset(s->d_bq, 0, (size_t)total_rows * row_bytes));
  const int64_t mask_words = ns_pad / 16;
  if (int r = pool_take(c, s->pool_mask, sizeof(uint32_t) * (size_t)mask_words * G, reinterpret_cast<void**>(&s->d_mask_hi))) return r;
  bool any_masked = false;
  for (size_t g = 0; g < G; ++g) {
    const Group& gr = c->groups[g];
    if ((int64_t)gr.n != c->n_samples_total) any_masked = true;
    mask_full_kernel<<<(unsigned)std::min<int64_t>((mask_words + 255) / 256, 1024), 256>>>(gr.d_mask, mask_words,
                                                                                          s->d_mask_hi + g * mask_words);
    c->launches++;
  }
  for (auto& ps : s->passes) {
    for (const Segment& sg : ps.segs) {
      const Group& gr = c->groups[sg.group];
      int row = (int)ps.bq_row0 + sg.row0;
      for (int i = 0; i < sg.n_cols; ++i) {
        const int col = sg.c_first + i, k = s->scale_off[sg.group] + col;
        const int nd = s->cols[sg.group].nd[col];
        quantize_kernel<<<gxb, 256>>>(column_ptr(sg.group, col), nd, ns_pad, s->d_colstat + k, row, s->d_bq, s->d_colscale + k,
                                      s->d_errfx + k);
        row += nd;
        c->launches++;
      }
      ones_row_kernel<<<gxb, 256>>>(gr.d_mask, ns_pad, row, s->d_bq);
      c->launches++;
    }
  }
  errsum_kernel<<<(unsigned)((nscale + 255) / 256), 256>>>(s->d_errfx, s->d_colscale, (int)nscale, s->d_colscale + nscale);
  c->launches++;
  LRR_CUDA(c, cudaGetLastError());
  {
    // exactness guard: no f32 accumulator can leave the exactly-representable range, whatever the genotypes are
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
    const unsigned gxq = (unsigned)std::max<int64_t>(1, std::min<int64_t>((row_bytes / 16 + 255) / 256, 64));
    for (int64_t r0 = 0; r0 < total_rows; r0 += 65535)
      acc_bound_kernel<<<dim3(gxq, (unsigned)std::min<int64_t>(total_rows - r0, 65535)), 256>>>(s->d_bq + r0 * row_bytes,
                                                                                             row_bytes, d_bound + 2 * r0);
    c->launches++;
    std::vector<unsigned long long> h_bound(2 * (size_t)total_rows);
    cudaError_t e = cudaMemcpy(h_bound.data(), d_bound, sizeof(unsigned long long) * h_bound.size(), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return cuda_fail(c, e, "tc4 acc_bound_kernel");
    unsigned long long worst = 0;
    for (unsigned long long v : h_bound) worst = std::max(worst, v);
    if (worst > (1ull << 24)) {
      s->why = "f32 accumulators could leave the exact range for this many samples (worst-case column bound " +
               std::to_string(worst) + " > 2^24)";
      return LRR_OK;
    }
  }
  const int budget = 227 * 1024 - (int)sizeof(Barriers) - 1024;
  for (auto& ps : s->passes) {
    ps.mask_bytes = any_masked ? (int)ps.segs.size() * 128 : 0;
    ps.gstage_bytes = (GENO_BYTES + ps.mask_bytes + 1023) / 1024 * 1024;
    // TMEM: accumulators from column 0, scale factors in the top 16 columns, the A ring in between
    ps.ring_base1 = ps.ncols > 224 ? ps.ncols : (ps.ncols + 31) / 32 * 32;   // 240 columns: the ring starts right behind them
    ps.ring_base2 = (2 * ps.ncols + 31) / 32 * 32;
    ps.nu1 = (SF_BASE - ps.ring_base1) / UNIT_COLS >= 6 ? 6 : 4;
    if (const char* e = tuning_env("LRR_TC4_NU1")) { if (atoi(e) == 4) ps.nu1 = 4; }
    if ((!wide && ps.ring_base2 + NU2 * UNIT_COLS > SF_BASE) || ps.ring_base1 + ps.nu1 * UNIT_COLS > SF_BASE) {
      s->why = "not enough tensor memory for the A ring";
      return LRR_OK;
    }
    for (int cs = 1; cs <= 2; ++cs) {
      PassShape& sh = ps.shape[cs - 1];
      const int rows = ps.ncols / cs;   // panel rows held by one CTA
      if (encode_2d(s, &sh.b_map, s->d_bq + ps.bq_row0 * row_bytes, (uint64_t)row_bytes, (uint64_t)ps.ncols,
                    (uint64_t)row_bytes, PANEL / 2, (uint32_t)rows)) {
        s->why = "cuTensorMapEncodeTiled failed for the basis panels";
        return LRR_OK;
      }
      sh.bstage_bytes = PANELS * rows * 128;
      int bst = NB;
      if (bst > 3 && budget - bst * sh.bstage_bytes < 6 * ps.gstage_bytes) bst = 3;
      while (bst > 2 && budget - bst * sh.bstage_bytes < 3 * ps.gstage_bytes) --bst;
      if (const char* e = tuning_env("LRR_TC4_BST")) { const int v = atoi(e); if (v >= 2 && v <= NB && budget - v * sh.bstage_bytes >= 2 * ps.gstage_bytes) bst = v; }
      int gst = (budget - bst * sh.bstage_bytes) / ps.gstage_bytes;
      if (gst > MAX_GSTAGES) gst = MAX_GSTAGES;
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
  LRR_CUDA(c, cudaFuncSetAttribute(tc4_sweep_kernel<NG_, CS_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024))
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

// Before the first quantisation of a group set: when the column count makes a multi-pass plan certain (every column has
// at least 6 digits) the basis is quantised once, for the wide plan.
static int first_plan_is_wide(Ctx* c, const uint8_t* d_row_flags, int64_t M, cudaStream_t st, bool* wide) {
  (void)d_row_flags; (void)M; (void)st;
  State* s = state(c);
  *wide = false;
  if (s->prepared) return LRR_OK;
  int64_t dot_cols = 0;
  for (const Group& gr : c->groups) dot_cols += gr.C;
  *wide = 6 * dot_cols > PASS_COLS;   // more than one narrow pass for certain: wide passes (clean input: one plane; else split)
  return LRR_OK;
}

}  // namespace tc4

// usable at all; `single_pass_only`: only when one sweep covers every column (what LRR_KERNEL_AUTO asks)
bool tc4_supported(Ctx* c, bool single_pass_only, const uint8_t* d_row_flags, int64_t M, cudaStream_t st) {
  tc4::State* s0 = tc4::state(c);
  bool first_wide = false;
  if (tc4::first_plan_is_wide(c, d_row_flags, M, st, &first_wide) != LRR_OK) return false;
  if (tc4::prepare(c, s0->prepared ? s0->wide : first_wide) != LRR_OK) return false;
  tc4::State* s = tc4::state(c);
  if (!s->usable) {
    c->err = s->why;
    return false;
  }
  if (single_pass_only && s->passes.size() > 1) {
    c->err = "more digit columns than one 4-bit sweep holds";
    return false;
  }
  return true;
}

// per-column quantum (value of one unit of the lowest digit) of group g's dot columns [C + n_fit], after tc4_supported
const double* tc4_quantum(Ctx* c, int g, int* n_fit) {
  tc4::State* s = tc4::state(c);
  if (!s->usable) return nullptr;
  if (n_fit) *n_fit = s->cols[g].n_fit;
  return s->d_colscale + s->scale_off[g];
}

// per-column sum of the rounding errors of the stored basis values (same layout as tc4_quantum)
const double* tc4_errsum(Ctx* c, int g) {
  tc4::State* s = tc4::state(c);
  if (!s->usable) return nullptr;
  return s->d_colscale + s->nscale + s->scale_off[g];
}

void tc4_invalidate(Ctx* c) {
  if (!c->tc4_state) return;
  tc4::free_prepared(static_cast<tc4::State*>(c->tc4_state));
}

// lrr_trim: give the pools back once no plan uses them
void tc4_trim(Ctx* c) {
  if (!c->tc4_state) return;
  tc4::State* s = static_cast<tc4::State*>(c->tc4_state);
  if (!s->prepared) tc4::free_pools(s);
}

void tc4_release(Ctx* c) {
  if (!c->tc4_state) return;
  tc4::State* s = static_cast<tc4::State*>(c->tc4_state);
  tc4::free_prepared(s);
  tc4::free_pools(s);
  cudaFree(s->d_any);
  cudaFree(s->d_bound);
  if (s->h_any) cudaFreeHost(s->h_any);
  delete s;
  c->tc4_state = nullptr;
}

int launch_tc4_sweep(Ctx* c, const uint8_t* d_packed, const uint8_t* d_row_flags, int64_t M, int64_t stride,
                     cudaStream_t st) {
  using namespace tc4;
  if (M == 0) return LRR_OK;
  State* s = state(c);
  bool first_wide = false;
  if (int r = first_plan_is_wide(c, d_row_flags, M, st, &first_wide)) return r;
  if (int r = prepare(c, s->prepared ? s->wide : first_wide)) return r;
  if (!s->usable) return fail(c, LRR_EINVAL, "4-bit tensor-core kernel unavailable: " + s->why);
  // More digit columns than one sweep holds: every pass re-reads the genotypes, so the passes are WIDE (up to 240 columns,
  // one accumulator set).  When the row flags say that no row holds a missing call (one small reduction + a 4-byte read per
  // call) a wide pass is one one-plane sweep; otherwise it is SPLIT into two one-plane sweeps -- plane c of every tile, then
  // plane m of the flagged tile pairs, which finishes their rows -- instead of twice as many narrow two-plane sweeps (a
  // two-plane tile has tensor memory for two chunks of A operand only and runs 1.6x slower than two one-plane tiles).
  bool split = false;
  if (s->passes.size() > 1 || s->wide) {
    split = true;   // unknown flags: every tile may hold a missing call
    if (d_row_flags) {
      if (!s->d_any) {
        LRR_CUDA(c, cudaMalloc(&s->d_any, sizeof(int32_t)));
        LRR_CUDA(c, cudaMallocHost(&s->h_any, sizeof(int32_t)));
      }
      LRR_CUDA(c, cudaMemsetAsync(s->d_any, 0, sizeof(int32_t), st));
      any_flag_kernel<<<(unsigned)std::min<int64_t>((M + 255) / 256, 1184), 256, 0, st>>>(d_row_flags, M, s->d_any);
      c->launches++;
      LRR_CUDA(c, cudaMemcpyAsync(s->h_any, s->d_any, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
      LRR_CUDA(c, cudaStreamSynchronize(st));
      split = *s->h_any != 0;
    }
    if (!s->wide) {
      if (int r = prepare(c, true)) return r;
      if (!s->usable) return fail(c, LRR_EINVAL, "4-bit tensor-core kernel unavailable: " + s->why);
    }
  }
  CUtensorMap geno_map;
  const bool abl_contig = tuning_env("LRR_ABL_CONTIG") != nullptr;   // timing ablation only
  if (abl_contig) {
    if (encode_2d(s, &geno_map, d_packed, 128, (uint64_t)(M * stride / 128), 128, 128, TILE_M))
      return fail(c, LRR_ECUDA, "cuTensorMapEncodeTiled failed (contiguous ablation)");
  } else if (encode_2d(s, &geno_map, d_packed, (uint64_t)stride, (uint64_t)M, (uint64_t)stride, 128, TILE_M))
    return fail(c, LRR_ECUDA, "cuTensorMapEncodeTiled failed for the genotype store (pointer must be 16-byte aligned)");
  for (const Pass& ps : s->passes)
  for (int mode = split ? 1 : 0; mode <= (split ? 2 : 0); ++mode) {
    Params p;
    memset(&p, 0, sizeof p);
    p.plane_mode = mode;
    p.M = M;
    p.n_tiles = (int)((M + TILE_M - 1) / TILE_M);
    p.n_chunks = (int)(stride / 128);
    p.ncols = ps.ncols;
    int cs = s->cluster;
    if (p.n_tiles < 2 * cs) cs = 1;
    const PassShape& sh = ps.shape[cs - 1];
    p.n_gstages = sh.n_gstages;
    p.n_bstages = sh.n_bstages;
    p.n_groups = (int)ps.segs.size();
    p.ring_base1 = ps.ring_base1;
    p.nu1 = ps.nu1;
    p.ring_base2 = ps.ring_base2;
    p.gstage_bytes = ps.gstage_bytes;
    p.bstage_bytes = sh.bstage_bytes;
    p.mask_bytes = ps.mask_bytes;
    p.row_flags = d_row_flags;
    p.one_plane_only = (s->wide && !split) ? 1 : 0;
    p.abl_contig = abl_contig ? 1 : 0;
    p.abl_stream = tuning_env("LRR_ABL_STREAM") ? atoi(tuning_env("LRR_ABL_STREAM")) : 0;
    p.abl = tuning_env("LRR_ABL_BITS") ? atoi(tuning_env("LRR_ABL_BITS")) : 0;
    for (int i = 0; i < p.n_groups; ++i) {
      const Segment& sg = ps.segs[i];
      const Group& gr = c->groups[sg.group];
      p.g[i].col_off = sg.row0;
      p.g[i].C = sg.n_cols;
      p.g[i].n_digit_cols = sg.n_digit_cols;
      memcpy(p.g[i].nd, s->cols[sg.group].nd.data() + sg.c_first, (size_t)sg.n_cols);
      p.g[i].n = gr.n;
      p.g[i].counts = c->d_counts + (int64_t)sg.group * c->reserved_variants * 4;
      p.g[i].dots = c->d_dots + c->dots_offset[sg.group] + sg.c_first;
      p.g[i].dots_stride = gr.C + 2;   // [C dot products | fitted-value dot products (<= 2)]
      p.g[i].colscale = s->d_colscale + s->scale_off[sg.group] + sg.c_first;
      p.g[i].mask_hi = s->d_mask_hi + (int64_t)sg.group * (gr.ns_pad / 16);
    }
    void* kfn = nullptr;
#define LRR_PICK(NG_) (cs == 2 ? (void*)tc4_sweep_kernel<NG_, 2> : (void*)tc4_sweep_kernel<NG_, 1>)
    kfn = p.n_groups == 1 ? LRR_PICK(1) : p.n_groups == 2 ? LRR_PICK(2) : LRR_PICK(0);
#undef LRR_PICK
    void* args[3] = {(void*)&geno_map, (void*)&sh.b_map, (void*)&p};
    if (int r = tcc::launch_persistent_clusters(c, kfn, cs, THREADS, sh.smem_bytes, p.n_tiles, args, st)) return r;
    int used = 0;
    for (const Segment& sg : ps.segs) used += sg.n_digit_cols + 1;
    c->sweep_shape[0] += 1;
    c->sweep_shape[1] += ps.ncols;
    if (mode == 0 && !p.one_plane_only) c->sweep_shape[2] += ps.ncols;   // narrow sweep: flagged tile pairs run plane m too
    c->sweep_shape[3] += used;
  }
  return LRR_OK;
}

}  // namespace lrr
