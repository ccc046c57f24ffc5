// tcgen05 / TMEM / TMA implicit-GEMM convolution for sm_100a.
//
// GEMM view: M = 128 "virtual pixels" of one image (a BH x BW patch), N = BN output channels,
// K = (taps of the phase) x Cin.  For every tap and KC-channel chunk one pipeline stage is filled by
// two TMA tiled loads:
//   A: box (KC ch, BW px, BH px, 1 img) of the haloed NHWC activation tensor at the tap's offset
//      (element stride S along w/h for strided convs; out-of-bounds = zero fill = zero padding),
//      landing in shared memory as 128 rows x KC in the canonical K-major swizzled UMMA layout;
//   B: box (KC, BN) of the packed weight matrix [tap*Cout + n][Cin].
// One elected thread issues tcgen05.mma (M=128, N=BN, K=16) into a double-buffered TMEM accumulator;
// EG groups of four epilogue warps drain TMEM in alternating passes of 32 output channels (tcgen05.ld 32x32b), convert
// to 16-bit, stage the pass in swizzled shared memory, reduce per-channel sum / sum-of-squares for InstanceNorm
// (deterministic per-tile partials) and write the rows out with 16-byte coalesced stores.
//
// Warp roles (64 + 128*EG threads): warp 0 = TMA producer, warp 1 = TMEM alloc + MMA issuer, then EG groups of four
// epilogue warps (warp_id % 4 selects the TMEM lane quarter).  EG = 2 for the K-deep ResnetBlock convolutions (the
// drain hides under 36 K iterations and the shared memory goes to a fourth operand stage), EG = 4 for the short-K
// layers (down / up convolutions), whose tiles are bound by the drain.  Persistent: grid = min(tiles, #SM).
#include "tc_common.cuh"
#include <mutex>
#include <unordered_map>
#include <string>
#include <stdlib.h>

namespace ng {

// ------------------------------------------------------------------------------------------------
// kernel
// ------------------------------------------------------------------------------------------------
struct TcParams {
  ConvGeom g;
  int BH, BW;                 // patch of virtual pixels (BH*BW <= 128)
  int patches_y, patches_x, co_tiles;
  int total_tiles;
  int group_items;            // G = B * patches_y * patches_x : tiles sharing one (phase, cout-tile) weight slab
  int groups_per;             // ceil(G / cluster size)
  int total_groups;           // co_tiles * nphase * groups_per
  int epilogue, act, crop;
  float slope;
  int bf16;
  int stat_slots;
  const float* bias;
  void* y;
  float* stat_partials;
  float* mean_rstd;           // fused finalisation (optional)
  int* tile_counters;
  unsigned long long* stat_acc;   // fixed-point per-(image, channel) statistics accumulators (optional)
  int sparse_merged;          // merged phases: skip the identically-zero (shift, phase) weight slabs (off by default)
  int image_minor;            // tile order: consecutive work items walk the images first (spreads the accumulator atomics)
  unsigned long long mg_B;    // ceil(2^42 / B)
  // DS ("direct stem"): the A operand is built in shared memory from the caller's NCHW fp32 planes
  const float* src;           // [B][src_c][src_H][src_W]
  int src_c, src_H, src_W, src_wrap;
  unsigned long long mg_groups_per, mg_patches_x, mg_patches_y;   // ceil(2^42 / d) reciprocals for decode()
  int arrivals_per_image;     // epilogue-group arrivals that complete an image
  float inv_count;            // 1 / (Hout * Wout)
};

// RT ("row taps", generator stem in row-merged form): the taps of a K x 1 stride-1 convolution only shift the patch by
// whole rows, so the haloed patch (BH + K - 1 rows of RT_BW pixels) is fetched ONCE per tile and every tap's A operand
// is a descriptor into it (row shifts of RT_BW = 16 pixels are whole 1024-byte swizzle atoms); the K weight slabs stay
// resident in shared memory for the life of the CTA.  L2 -> SM traffic per tile drops from K x (A + B) to one patch.
constexpr int RT_BW = 16, RT_BH = 8, RT_KH = 7;
constexpr int RT_PATCH_ROWS = (RT_BH + RT_KH - 1) * RT_BW;          // 224 pixels

// DS ("direct stem", implies RT): the generator stem straight from the NCHW fp32 tiles.  Four producer warps stage the
// tile's 14 x 22 source window (wrapper reflect pad + stem reflect halo resolved, fp32 -> 16 bit) and write the haloed
// patch in the row-merged form -- element kw*4 + c of pixel (y, x) = channel c at (y, x + kw), 32 elements = 64 bytes per
// pixel, 64-byte swizzle -- directly in the canonical K-major UMMA layout.  No intermediate tensor in HBM (ng_prep_stem
// wrote and the conv re-read 217 + 275 MB per 32 tiles), K = 7 x 32 instead of 7 x 64 (half the MMAs: 3 real channels
// are padded to 4, not 8).
constexpr int DS_WIN_W = RT_BW + 8;                                  // 22 source columns + 2 (8-byte row alignment)
constexpr int DS_SCRATCH_BYTES = (RT_BH + RT_KH - 1) * DS_WIN_W * 8;

// Merged phases: virtual channel block ("slot") s of the 4 * Cout_real outputs holds output parity
// (pa, pb) = (s >> 1, (s & 1) ^ (s >> 1)) -- Gray order (0,0), (0,1), (1,1), (1,0) -- so that the slots an input shift
// (sy, sx) contributes to (pa >= sy and pb >= sx) are always CONTIGUOUS: shift (0,0) -> slots 0..3, (0,1) -> 1..2,
// (1,0) -> 2..3, (1,1) -> 2.  The banded form multiplies each shift by exactly that band: one N = band x Cout_real MMA
// chain per K chunk instead of N = 4 x Cout_real with 7 of 16 slabs structurally zero.
__host__ __device__ __forceinline__ bool merged_slot_uses_shift(int slot, int tp) {
  const int pa = slot >> 1, pb = (slot & 1) ^ (slot >> 1);
  return pa >= (tp >> 1) && pb >= (tp & 1);
}
// band of used slots inside [ph_lo, ph_lo + nph) for shift tp: l in [lo, hi) relative to ph_lo (lo == hi: none)
__device__ __forceinline__ void merged_band(int tp, int ph_lo, int nph, int& lo, int& hi) {
  lo = nph; hi = 0;
  for (int l = 0; l < nph; ++l)
    if (merged_slot_uses_shift(ph_lo + l, tp)) { if (l < lo) lo = l; hi = l + 1; }
  if (hi == 0) lo = 0;
}

template <int BN, int KC, bool RT = false, int EGW = 2, bool DS = false>
struct TcCfg {
  static constexpr int A_BYTES = RT ? RT_PATCH_ROWS * KC * 2 : 128 * KC * 2;
  static constexpr int B_BYTES = RT ? 0 : (BN * KC * 2 + 1023) / 1024 * 1024;
  static constexpr int W_BYTES = RT ? RT_KH * BN * KC * 2 : 0;      // resident weights (RT only)
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  // epilogue: EG groups of four warps drain the accumulator in alternating passes of EB = 32 output channels, each
  // group with its own 8 KB staging tile, stats scratch and row-offset table (a single warp per scheduler is latency
  // bound, so two groups nearly halve the drain time of a tile)
  static constexpr int EG = BN >= 64 ? EGW : 1;
  static constexpr int EB = 32;
  // more groups than a tile has passes (BN = 64: two passes, four groups): the groups split into TG sets that take
  // alternate tiles, i.e. one accumulator buffer each, so two tiles are drained at the same time
  static constexpr int PASSES = BN >= 64 ? BN / EB : 1;
  static constexpr int TG = EG > PASSES ? EG / PASSES : 1;
  static constexpr int EGT = EG / TG;                         // groups that share one tile
  static_assert(TG == 1 || TG == 2, "tile-alternating epilogue sets map onto the two accumulator buffers");
  static constexpr int THREADS = 64 + 128 * EG + (DS ? 128 : 0);
  static constexpr int DS_WARP0 = 2 + 4 * EG;                   // first of the four DS producer warps
  static constexpr int STAGING_BYTES = (BN >= 64 ? EG * 128 * EB * 2 : 0) + (DS ? 2 * DS_SCRATCH_BYTES : 0);
  static constexpr int RED_BYTES = BN >= 64 ? EG * 2048 : 0;         // cross-row-group stats combine
  static constexpr int ROWOFF_BYTES = EG * 1024 + 64;          // + per-group "this group finalises" flags
  static constexpr int BUDGET = 222 * 1024;
  static constexpr int STAGES_RAW = (BUDGET - STAGING_BYTES - RED_BYTES - W_BYTES) / STAGE_BYTES;
  static constexpr int STAGES = DS ? 4 : (STAGES_RAW > 8 ? 8 : STAGES_RAW);
  static constexpr int ACC_STRIDE = BN < 32 ? 32 : BN;   // TMEM columns between the two accumulators
  static constexpr int TMEM_COLS = 2 * ACC_STRIDE <= 64 ? 64 : (2 * ACC_STRIDE <= 128 ? 128 : (2 * ACC_STRIDE <= 256 ? 256 : 512));
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + W_BYTES + STAGING_BYTES + RED_BYTES + ROWOFF_BYTES +
                                    256 /*barriers*/ + 1024 /*alignment slack*/;
  static_assert(STAGES >= 2, "pipeline needs at least two stages");
};

// CS = thread-block-cluster size.  The CS CTAs of a cluster work on CS different pixel patches that share the same
// weight slab; each loads 1/CS of every B (weight) stage and multicasts it to all of them, which divides the
// L2 -> SM weight traffic (the measured limiter of the 128x256 tile) by CS.
template <int BN, int KC, int CS, bool RT = false, int EGW = 2, bool DS = false>
__global__ void __launch_bounds__(TcCfg<BN, KC, RT, EGW, DS>::THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ TcParams p) {
  using Cfg = TcCfg<BN, KC, RT, EGW, DS>;
  static_assert(!RT || (CS == 1 && (KC == 64 || (DS && KC == 32))), "row-tap variant: no clusters, 64-channel rows");
  static_assert(!DS || RT, "the direct stem is a row-tap kernel");
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* stage_base = smem;
  uint8_t* wres = smem + STAGES * Cfg::STAGE_BYTES;            // RT: resident weights [tap][BN][KC]
  uint8_t* staging = wres + Cfg::W_BYTES;
  uint8_t* ds_scratch = staging + (Cfg::STAGING_BYTES - (DS ? 2 * DS_SCRATCH_BYTES : 0));   // DS: two source windows
  float* red = reinterpret_cast<float*>(staging + Cfg::STAGING_BYTES);
  long long* rowoff = reinterpret_cast<long long*>(reinterpret_cast<uint8_t*>(red) + Cfg::RED_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(rowoff) + Cfg::ROWOFF_BYTES);
  uint64_t* full_bar = bars;                 // [STAGES]
  uint64_t* empty_bar = bars + STAGES;       // [STAGES]
  uint64_t* tfull_bar = bars + 2 * STAGES;   // [2]
  uint64_t* tempty_bar = bars + 2 * STAGES + 2;
  uint64_t* wfull_bar = bars + 2 * STAGES + 4;                  // RT: weights landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 5);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const ConvGeom& g = p.g;
  const uint32_t crank = CS > 1 ? cluster_ctarank() : 0u;
  const int cluster_id = blockIdx.x / CS, num_clusters = gridDim.x / CS;
  constexpr uint16_t MC_MASK = (uint16_t)((1u << CS) - 1u);

  if (threadIdx.x == 0) {
    // DS: a stage is filled by the four producer warps (one arrival each), not by a TMA transaction
    for (int s = 0; s < STAGES; ++s) { mbar_init(smem_u32(&full_bar[s]), DS ? 4 : 1); mbar_init(smem_u32(&empty_bar[s]), CS); }
    for (int s = 0; s < 2; ++s) { mbar_init(smem_u32(&tfull_bar[s]), 1); mbar_init(smem_u32(&tempty_bar[s]), 128 * Cfg::EGT); }
    mbar_init(smem_u32(wfull_bar), 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)Cfg::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CS > 1) cluster_sync_all();     // peers' barriers are initialised before any remote arrive / multicast
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int chunks = g.Cin / KC;
  // work item q (one per cluster and round) -> (cout tile, phase, group index); CTA `crank` takes item gi*CS + crank
  // divisions by launch constants use precomputed reciprocals: n / d == (n * ceil(2^42 / d)) >> 42, exact while
  // n * d < 2^42 (checked on the host); the epilogue runs this once per tile in every one of its threads
  auto fdiv = [](int n, unsigned long long magic) { return (int)(((unsigned long long)(unsigned)n * magic) >> 42); };
  auto decode = [&](int q, int& cot, int& ph, int& n, int& py, int& px, bool& dummy) {
    const int t = fdiv(q, p.mg_groups_per);
    const int gi = q - t * p.groups_per;
    cot = g.nphase == 1 ? t : (t >> 2);       // nphase is 1 or 4
    ph = g.nphase == 1 ? 0 : (t & 3);
    int idx = gi * CS + (int)crank;
    dummy = idx >= p.group_items;             // odd tail: recompute the last patch, write nothing
    if (dummy) idx = p.group_items - 1;
    if (p.image_minor) {                      // idx = patch * B + n
      const int pt = fdiv(idx, p.mg_B);
      n = idx - pt * g.B;
      py = fdiv(pt, p.mg_patches_x);
      px = pt - py * p.patches_x;
    } else {                                  // idx = (n * patches_y + py) * patches_x + px
      const int rowi = fdiv(idx, p.mg_patches_x);
      px = idx - rowi * p.patches_x;
      n = fdiv(rowi, p.mg_patches_y);
      py = rowi - n * p.patches_y;
    }
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      if constexpr (RT) {
        const uint32_t wb = smem_u32(wfull_bar);
        mbar_expect_tx(wb, (uint32_t)Cfg::W_BYTES);
        for (int tp = 0; tp < RT_KH; ++tp)
          tma_load_2d(&tmB, wb, smem_u32(wres) + tp * (BN * KC * 2), 0, g.taps[tp].wrow);
      }
      for (int q = cluster_id; q < p.total_groups; q += num_clusters) {
        int cot, ph, n, py, px; bool dummy;
        decode(q, cot, ph, n, py, px, dummy);
        const int i0 = py * p.BH, j0 = px * p.BW;
        if constexpr (DS) break;                                         // the patches come from the DS producer warps
        if constexpr (RT) {
          // one haloed patch per tile: rows i0 + dy0 .. i0 + dy0 + BH + K - 2, columns j0 + dx0 .. + BW - 1
          mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
          const uint32_t fb = smem_u32(&full_bar[stage]);
          mbar_expect_tx(fb, (uint32_t)Cfg::A_BYTES);
          tma_load_4d(&tmA, fb, smem_u32(stage_base + stage * Cfg::STAGE_BYTES), 0, j0 + g.taps[0].dx, i0 + g.taps[0].dy, n);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
          continue;
        }
        for (int tp = g.phase_tap0[ph]; tp < g.phase_tap0[ph + 1]; ++tp) {
          const int by = g.S * i0 + g.taps[tp].dy, bx = g.S * j0 + g.taps[tp].dx;
          const int brow = g.taps[tp].wrow + cot * BN;
          // merged phases: tap tp is the input shift (sy, sx) = (tp >> 1, tp & 1); output phase (pa, pb) uses it iff
          // pa >= sy and pb >= sx -- 9 of the 16 (shift, phase) weight slabs are non-zero.  Only those are fetched (and
          // multiplied, see the MMA issuer); a shift that none of this tile's phases uses is skipped altogether.
          uint32_t used = 0;
          const bool sparse = g.merged && p.sparse_merged;
          const int Cr = g.Cout_real, nph = sparse ? BN / Cr : 0, ph_lo = sparse ? (cot * BN) / Cr : 0;
          for (int l = 0; l < nph; ++l)
            if (merged_slot_uses_shift(ph_lo + l, tp)) used |= 1u << l;
          if (sparse && used == 0) continue;
          for (int c = 0; c < chunks; ++c) {
            mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);       // free in every CTA of the cluster
            const uint32_t fb = smem_u32(&full_bar[stage]);
            const uint32_t sa = smem_u32(stage_base + stage * Cfg::STAGE_BYTES);
            if (sparse) {
              mbar_expect_tx(fb, (uint32_t)(p.BH * p.BW * KC * 2 + __popc(used) * Cr * KC * 2));
              tma_load_4d(&tmA, fb, sa, c * KC, bx, by, n);
              for (int l = 0; l < nph; ++l)          // tmB's box is one phase slab (Cout_real rows) for merged launches
                if (used & (1u << l)) tma_load_2d(&tmB, fb, sa + Cfg::A_BYTES + l * (Cr * KC * 2), c * KC, brow + l * Cr);
              if (++stage == STAGES) { stage = 0; phase ^= 1; }
              continue;
            }
            mbar_expect_tx(fb, (uint32_t)(p.BH * p.BW * KC * 2 + BN * KC * 2));   // the A box holds BH*BW (<=128) rows
            tma_load_4d(&tmA, fb, sa, c * KC, bx, by, n);
            if constexpr (CS == 1) {
              tma_load_2d(&tmB, fb, sa + Cfg::A_BYTES, c * KC, brow);
            } else {
              constexpr int ROWS = BN / CS;                          // this CTA's slice of the weight rows
              tma_load_2d_mc(&tmB, fb, sa + Cfg::A_BYTES + crank * (ROWS * KC * 2), c * KC, brow + (int)crank * ROWS,
                             MC_MASK);
            }
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | ((uint32_t)(p.bf16 ? 1 : 0) << 7) | ((uint32_t)(p.bf16 ? 1 : 0) << 10) |
                             ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      constexpr uint32_t SBO = KC == 64 ? 1024 : (KC == 32 ? 512 : 256);      // 8 rows of KC 16-bit elements
      constexpr uint32_t LAYOUT = KC == 64 ? 2 : (KC == 32 ? 4 : 6);          // SWIZZLE_128B / 64B / 32B
      uint32_t stage = 0, phase = 0, as = 0, as_phase = 0;
      for (int q = cluster_id; q < p.total_groups; q += num_clusters) {
        const int ph = (q / p.groups_per) % g.nphase;
        const int kiters = (g.phase_tap0[ph + 1] - g.phase_tap0[ph]) * chunks;
        mbar_wait(smem_u32(&tempty_bar[as]), as_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_c = tmem_base + as * Cfg::ACC_STRIDE;
        if constexpr (RT) {
          if (q == cluster_id) { mbar_wait(smem_u32(wfull_bar), 0); }      // resident weights (first tile only)
          mbar_wait(smem_u32(&full_bar[stage]), phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(stage_base + stage * Cfg::STAGE_BYTES);
#pragma unroll 1
          for (int tp = 0; tp < RT_KH; ++tp) {
            // tap tp = the patch shifted down by tp rows of RT_BW pixels (tp * 2048 bytes: whole swizzle atoms)
            const uint64_t adesc = make_kmajor_desc(sa + tp * (RT_BW * KC * 2), SBO, LAYOUT);
            const uint64_t bdesc = make_kmajor_desc(smem_u32(wres) + tp * (BN * KC * 2), SBO, LAYOUT);
#pragma unroll
            for (int k = 0; k < KC / 16; ++k)
              umma_f16(tmem_c, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (uint32_t)((tp | k) != 0));
          }
          umma_commit(smem_u32(&empty_bar[stage]));
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
          umma_commit(smem_u32(&tfull_bar[as]));
          if (++as == 2) { as = 0; as_phase ^= 1; }
          continue;
        }
        if (g.merged && p.sparse_merged == 2) {
          // banded merged phases: per shift ONE MMA chain over the contiguous band of slots that use it
          const int Cr = g.Cout_real, nph = BN / Cr;
          const int cot = (q / p.groups_per) / g.nphase, ph_lo = (cot * BN) / Cr;
          int kit = 0;
          for (int tp = 0; tp < g.ntaps; ++tp) {
            int lo, hi;
            merged_band(tp, ph_lo, nph, lo, hi);
            if (hi == lo) continue;
            const uint32_t idesc_b = (idesc & ~(0x3Fu << 17)) | ((uint32_t)(((hi - lo) * Cr) >> 3) << 17);
            for (int c = 0; c < chunks; ++c, ++kit) {
              mbar_wait(smem_u32(&full_bar[stage]), phase);
              tc_fence_after();
              const uint32_t sa = smem_u32(stage_base + stage * Cfg::STAGE_BYTES);
              const uint64_t adesc = make_kmajor_desc(sa, SBO, LAYOUT);
              const uint64_t bdesc = make_kmajor_desc(sa + Cfg::A_BYTES + lo * (Cr * KC * 2), SBO, LAYOUT);
#pragma unroll
              for (int k = 0; k < KC / 16; ++k)       // shift 0 covers every slot of the tile: first MMA at kit == 0
                umma_f16(tmem_c + lo * Cr, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc_b,
                         (uint32_t)((kit | k) != 0));
              umma_commit(smem_u32(&empty_bar[stage]));
              if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
          }
          umma_commit(smem_u32(&tfull_bar[as]));
          if (++as == 2) { as = 0; as_phase ^= 1; }
          continue;
        }
        if (g.merged && p.sparse_merged) {
          // sparse merged phases (see the producer): one N = Cout_real MMA chain per (shift, phase) slab that is not
          // identically zero, each phase accumulating in its own TMEM column block
          const int Cr = g.Cout_real, nph = BN / Cr;
          const int cot = (q / p.groups_per) / g.nphase, ph_lo = (cot * BN) / Cr;
          const uint32_t idesc_s = (idesc & ~(0x3Fu << 17)) | ((uint32_t)(Cr >> 3) << 17);
          int kit = 0;
          for (int tp = 0; tp < g.ntaps; ++tp) {
            uint32_t used = 0;
            for (int l = 0; l < nph; ++l)
              if (merged_slot_uses_shift(ph_lo + l, tp)) used |= 1u << l;
            if (used == 0) continue;
            for (int c = 0; c < chunks; ++c, ++kit) {
              mbar_wait(smem_u32(&full_bar[stage]), phase);
              tc_fence_after();
              const uint32_t sa = smem_u32(stage_base + stage * Cfg::STAGE_BYTES);
              const uint64_t adesc = make_kmajor_desc(sa, SBO, LAYOUT);
              for (int l = 0; l < nph; ++l) {
                if (!(used & (1u << l))) continue;
                const uint64_t bdesc = make_kmajor_desc(sa + Cfg::A_BYTES + l * (Cr * KC * 2), SBO, LAYOUT);
#pragma unroll
                for (int k = 0; k < KC / 16; ++k)       // every phase uses shift 0: its first MMA is at kit == 0
                  umma_f16(tmem_c + l * Cr, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc_s,
                           (uint32_t)((kit | k) != 0));
              }
              umma_commit(smem_u32(&empty_bar[stage]));
              if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
          }
          umma_commit(smem_u32(&tfull_bar[as]));
          if (++as == 2) { as = 0; as_phase ^= 1; }
          continue;
        }
        for (int kit = 0; kit < kiters; ++kit) {
          mbar_wait(smem_u32(&full_bar[stage]), phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(stage_base + stage * Cfg::STAGE_BYTES);
          const uint64_t adesc = make_kmajor_desc(sa, SBO, LAYOUT);
          const uint64_t bdesc = make_kmajor_desc(sa + Cfg::A_BYTES, SBO, LAYOUT);
#pragma unroll
          for (int k = 0; k < KC / 16; ++k)
            umma_f16(tmem_c, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (uint32_t)((kit | k) != 0));
          if constexpr (CS == 1) umma_commit(smem_u32(&empty_bar[stage]));
          else umma_commit_mc(smem_u32(&empty_bar[stage]), MC_MASK);    // release the slot in every CTA of the cluster
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(smem_u32(&tfull_bar[as]));
        if (++as == 2) { as = 0; as_phase ^= 1; }
      }
    }
    __syncwarp();
  } else if (DS && warp >= Cfg::DS_WARP0) {
    // ===================== direct-stem patch producers (4 warps) =====================
    if constexpr (DS) {
      const int pt = threadIdx.x - Cfg::DS_WARP0 * 32;                   // 0..127
      constexpr int PR = RT_BH + RT_KH - 1;                              // 14 patch rows
      const int H1 = g.Hout, W1 = g.Wout, halo = RT_KH / 2;
      const size_t plane = (size_t)p.src_H * p.src_W;
      // Source window of a tile: 14 rows x 24 columns (padded-image columns j0 - 4 .. j0 + 19: the 22 the taps need plus one
      // on either side so that the window starts on an even source column) x up to 4 channels, kept in shared memory as
      // [row][column][4 channels] 16-bit (8 bytes per pixel).  Work item = (row, column PAIR): per channel one 8-byte load
      // of two neighbouring source pixels (a scalar pair with the reflections resolved where the pair touches the image
      // border, or when the geometry rules out aligned pairs), then ONE 16-byte shared store of the two converted pixels.
      // 168 items over 128 threads; the loads of tile i + 1 are issued into registers before tile i's patch is built, so
      // their latency hides behind the build and the wait for a free stage.
      constexpr int CPAIRS = DS_WIN_W / 2;                               // 12 column pairs per row
      constexpr int N_ITEMS = PR * CPAIRS;                               // 168
      constexpr int IPT = (N_ITEMS + 127) / 128;                         // items per thread: 2
      const bool vec_ok = ((p.src_W | p.src_wrap) & 1) == 0 && (reinterpret_cast<uintptr_t>(p.src) & 7) == 0 &&
                          (plane & 1) == 0;
      float2 v[IPT][4];
#pragma unroll
      for (int k = 0; k < IPT; ++k)
#pragma unroll
        for (int c = 0; c < 4; ++c) v[k][c] = make_float2(0.f, 0.f);
      auto issue_loads = [&](int q) {
        int cot, ph, n, py, px; bool dummy;
        decode(q, cot, ph, n, py, px, dummy);
        const int i0 = py * p.BH, j0 = px * p.BW;
        const float* sn = p.src + (size_t)n * p.src_c * plane;
#pragma unroll
        for (int k = 0; k < IPT; ++k) {
          const int item = pt + k * 128;
          if (item >= N_ITEMS) continue;
          const int r = item / CPAIRS, cp = item - r * CPAIRS;
          // row: padded-image row i0 + r - 3 -> both reflections resolved (rows are independent of one another)
          const int y1 = min(i0 + r - halo, H1 - 1 + halo);
          const int y0 = reflect_idx(reflect_idx(y1, H1) - p.src_wrap, p.src_H);
          const float* row = sn + (size_t)y0 * p.src_W;
          const int xp = j0 - 4 + 2 * cp;                                // padded-image column of the pair's first pixel
          const int xs = xp - p.src_wrap;                                // source column, if no reflection is involved
          if (vec_ok && xp >= 0 && xp + 1 < W1 && xs >= 0 && xs + 1 < p.src_W) {
#pragma unroll
            for (int c = 0; c < 4; ++c)
              if (c < p.src_c) v[k][c] = __ldg(reinterpret_cast<const float2*>(row + c * plane + xs));
          } else {
            const int xa = min(max(xp, -halo), W1 - 1 + halo), xb = min(max(xp + 1, -halo), W1 - 1 + halo);
            const int x0 = reflect_idx(reflect_idx(xa, W1) - p.src_wrap, p.src_W);
            const int x1 = reflect_idx(reflect_idx(xb, W1) - p.src_wrap, p.src_W);
#pragma unroll
            for (int c = 0; c < 4; ++c)
              if (c < p.src_c) v[k][c] = make_float2(__ldg(row + c * plane + x0), __ldg(row + c * plane + x1));
          }
        }
      };
      uint32_t stage = 0, phase = 0;
      int it = 0;
      if (cluster_id < p.total_groups) issue_loads(cluster_id);
      for (int q = cluster_id; q < p.total_groups; q += num_clusters, ++it) {
        const uint32_t wsrc = smem_u32(ds_scratch + (it & 1) * DS_SCRATCH_BYTES);
        // ---- this tile's source window: registers -> shared, two pixels x 4 channels (16 bit) per store
#pragma unroll
        for (int k = 0; k < IPT; ++k) {
          const int item = pt + k * 128;
          if (item >= N_ITEMS) continue;
          uint32_t w0, w1, w2, w3;
          if (p.bf16) {
            w0 = pack2<__nv_bfloat16>(v[k][0].x, v[k][1].x); w1 = pack2<__nv_bfloat16>(v[k][2].x, v[k][3].x);
            w2 = pack2<__nv_bfloat16>(v[k][0].y, v[k][1].y); w3 = pack2<__nv_bfloat16>(v[k][2].y, v[k][3].y);
          } else {
            w0 = pack2<__half>(v[k][0].x, v[k][1].x); w1 = pack2<__half>(v[k][2].x, v[k][3].x);
            w2 = pack2<__half>(v[k][0].y, v[k][1].y); w3 = pack2<__half>(v[k][2].y, v[k][3].y);
          }
          sts128(wsrc + item * 16, w0, w1, w2, w3);                      // item = row * 12 + pair -> (row * 24 + 2 * pair) * 8
        }
        asm volatile("bar.sync 6, 128;" ::: "memory");
        if (q + num_clusters < p.total_groups) issue_loads(q + num_clusters);      // next tile's loads fly during the build
        // ---- patch rows: pixel (r, x) -> 32 elements (kw*4 + c) = window columns x + 1 .. x + 8 (kw = 7 meets zero weights)
        if (pt < 32) mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
        asm volatile("bar.sync 6, 128;" ::: "memory");
        const uint32_t sa = smem_u32(stage_base + stage * Cfg::STAGE_BYTES);
#pragma unroll
        for (int k = 0; k < (RT_PATCH_ROWS * 4) / 128; ++k) {
          const int id = pt + k * 128;
          const int pix = id >> 2, qc = id & 3;
          const int r = pix >> 4, x = pix & 15;
          const uint32_t a0 = wsrc + ((r * DS_WIN_W + x + 2 * qc + 1) << 3);
          const long long lo = lds64(a0), hi = lds64(a0 + 8);
          sts128(sa + pix * 64 + ((qc ^ ((pix >> 1) & 3)) << 4), (uint32_t)lo, (uint32_t)(lo >> 32), (uint32_t)hi,
                 (uint32_t)(hi >> 32));
        }
        fence_proxy_async();                 // generic-proxy stores -> visible to the tensor core's async-proxy reads
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&full_bar[stage]));
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue (warps 2..5 = group 0, warps 6..9 = group 1) =====================
    const int qtr = warp & 3;               // TMEM lane quarter this warp may access
    const int row = qtr * 32 + lane;        // accumulator row == TMEM lane
    const int grp = (warp - 2) >> 2;        // epilogue group
    const int we = (warp - 2) & 3;          // warp within the group
    const int et = we * 32 + lane;          // 0..127 thread index within the group
    const int row_i = row / p.BW, row_j = row - row_i * p.BW;     // this row's pixel inside the patch (tile-invariant)
    uint32_t as = 0, as_phase = 0;
    const int tsel = grp / Cfg::EGT, gt = grp % Cfg::EGT;       // tile set of this group, group index within the tile
    int it = 0;
    for (int q = cluster_id; q < p.total_groups; q += num_clusters, ++it) {
      if (Cfg::TG > 1 && (it % Cfg::TG) != tsel) {              // the other set drains this tile (other accumulator)
        if (++as == 2) { as = 0; as_phase ^= 1; }
        continue;
      }
      int cot, ph, n, py, px; bool dummy;
      decode(q, cot, ph, n, py, px, dummy);
      const int vi = py * p.BH + row_i, vj = px * p.BW + row_j;
      const int oy = g.OS * vi + g.phase_oy[ph], ox = g.OS * vj + g.phase_ox[ph];
      const bool valid = !dummy && row < p.BH * p.BW && oy < g.Hout && ox < g.Wout;
      const int n0 = cot * BN;

      mbar_wait(smem_u32(&tfull_bar[as]), as_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + as * Cfg::ACC_STRIDE + ((uint32_t)(qtr * 32) << 16);

      if constexpr (BN < 64) {
        // ---- single real output channel (generator head, PatchGAN last layer) ----
        uint32_t r[16];
        tmem_ld16(taddr, r);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(smem_u32(&tempty_bar[as]));
        if (valid && p.epilogue == NG_EPI_HEAD) {
          const int hy = oy - p.crop, hx = ox - p.crop, HH = g.Hout - 2 * p.crop, WW = g.Wout - 2 * p.crop;
          if (hy >= 0 && hx >= 0 && hy < HH && hx < WW) {
            const float v = apply_act(__uint_as_float(r[0]) + (p.bias ? p.bias[0] : 0.f), p.act, p.slope);
            reinterpret_cast<float*>(p.y)[((size_t)n * HH + hy) * WW + hx] = v;
          }
        } else if (valid) {
          // thin (16 stored channels) NHWC output: data gradient of the PatchGAN input layer
          uint32_t w[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            float a = __uint_as_float(r[2 * k]), b = __uint_as_float(r[2 * k + 1]);
            if (p.epilogue == NG_EPI_BIAS_ACT) {
              a = apply_act(a + (p.bias ? p.bias[2 * k] : 0.f), p.act, p.slope);
              b = apply_act(b + (p.bias ? p.bias[2 * k + 1] : 0.f), p.act, p.slope);
            }
            w[k] = p.bf16 ? pack2<__nv_bfloat16>(a, b) : pack2<__half>(a, b);
          }
          uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(p.y) +
                                                ((((size_t)n * g.Hout + oy) * g.Wout + ox) * 16) * 2);
          dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
          dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
        }
      } else {
        constexpr int EB = Cfg::EB, EG = Cfg::EG, PASSES = BN / EB;
        const uint32_t stg = smem_u32(staging) + grp * (128 * EB * 2);     // shared-space byte addresses
        const uint32_t redg = smem_u32(red) + grp * 2048;
        const uint32_t rowoffg = smem_u32(rowoff) + grp * 1024;
        const uint32_t barid = 1 + grp;
        // element offset of this row's output pixel (merged phases: of its 2x2 output block's top-left pixel)
        sts64(rowoffg + row * 8, valid ? (((long long)n * g.Hout + oy) * g.Wout + ox) * (long long)g.Cout_real : -1ll);
#pragma unroll 1
        for (int ps = gt; ps < PASSES; ps += Cfg::EGT) {
          // where this pass's 32 channels go: plain = channel n0 + ps*EB of the row's pixel; merged phases = channel co0
          // of the pixel (2i + pa, 2j + pb) with phase = virtual channel / Cout_real
          int c_first = n0 + ps * EB, phase_id = ph;
          long long chan_off = c_first;
          if (g.merged) {
            phase_id = c_first / g.Cout_real;
            c_first -= phase_id * g.Cout_real;
            chan_off = ((long long)(phase_id >> 1) * g.Wout + ((phase_id & 1) ^ (phase_id >> 1))) * g.Cout_real + c_first;
          }
          {
            uint32_t r[32];
            tmem_ld32(taddr + ps * EB, r);
            tmem_ld_wait();
            if (ps + Cfg::EGT >= PASSES) {
              tc_fence_before();
              mbar_arrive(smem_u32(&tempty_bar[as]));     // this thread's last read of the accumulator
            }
            uint32_t w[16];
            if (p.epilogue == NG_EPI_BIAS_ACT) {
#pragma unroll
              for (int k = 0; k < 32; ++k) {
                const float b = p.bias ? p.bias[n0 + ps * EB + k] : 0.f;
                r[k] = __float_as_uint(apply_act(__uint_as_float(r[k]) + b, p.act, p.slope));
              }
            }
#pragma unroll
            for (int k = 0; k < 16; ++k) {
              const float a = valid ? __uint_as_float(r[2 * k]) : 0.f, b = valid ? __uint_as_float(r[2 * k + 1]) : 0.f;
              w[k] = p.bf16 ? pack2<__nv_bfloat16>(a, b) : pack2<__half>(a, b);
            }
            // 64-byte staging rows; 16-byte chunk index XOR (row / 2) mod 4 -> conflict-free writes and reads
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
              const int chunk = c4 ^ ((row >> 1) & 3);
              sts128(stg + row * (EB * 2) + chunk * 16, w[4 * c4], w[4 * c4 + 1], w[4 * c4 + 2], w[4 * c4 + 3]);
            }
          }
          bar_sync_id(barid);

          // ---- per-channel partial statistics (deterministic: fixed row order, fixed combine order) ----
          if (p.epilogue == NG_EPI_RAW && (p.stat_partials != nullptr || p.stat_acc != nullptr)) {
            constexpr int PAIRS = EB / 2, G = 128 / PAIRS;        // 16 channel pairs x 8 row groups
            const int cp = et % PAIRS, rs = et / PAIRS;
            float s0 = 0.f, q0 = 0.f, s1 = 0.f, q1 = 0.f;
            // rows rs, rs + 8, ...: the swizzle term ((row >> 1) & 3) is the same for all of them -> one base address
            const uint32_t a0 = stg + rs * (EB * 2) + (((cp >> 2) ^ ((rs >> 1) & 3)) * 16) + (cp & 3) * 4;
#pragma unroll
            for (int k = 0; k < 128 / G; ++k) {
              const uint32_t word = lds32(a0 + k * (G * EB * 2));
              const float2 v = p.bf16 ? unpack2<__nv_bfloat16>(word) : unpack2<__half>(word);
              s0 += v.x; q0 = fmaf(v.x, v.x, q0); s1 += v.y; q1 = fmaf(v.y, v.y, q1);
            }
            if (rs > 0) sts_f4(redg + ((rs - 1) * PAIRS + cp) * 16, make_float4(s0, q0, s1, q1));
            bar_sync_id(barid);
            if (rs == 0) {
#pragma unroll
              for (int k = 1; k < G; ++k) {
                const float4 o = lds_f4(redg + ((k - 1) * PAIRS + cp) * 16);
                s0 += o.x; q0 += o.y; s1 += o.z; q1 += o.w;
              }
              if (!dummy && p.stat_acc != nullptr) {
                // fixed point: integer adds commute, so the per-image totals are independent of tile completion order
                unsigned long long* acc = p.stat_acc + ((size_t)n * g.Cout_real + c_first + 2 * cp) * 2;
                constexpr float SS = (float)(1 << NG_STAT_SUM_SHIFT), SQ = (float)(1 << NG_STAT_SQ_SHIFT);
                atomicAdd(acc + 0, (unsigned long long)__float2ll_rn(s0 * SS));
                atomicAdd(acc + 1, (unsigned long long)__float2ll_rn(q0 * SQ));
                atomicAdd(acc + 2, (unsigned long long)__float2ll_rn(s1 * SS));
                atomicAdd(acc + 3, (unsigned long long)__float2ll_rn(q1 * SQ));
              }
              if (!dummy && p.stat_partials != nullptr) {
                const int slot = (phase_id * p.patches_y + py) * p.patches_x + px;
                float* dst = p.stat_partials + (((size_t)n * p.stat_slots + slot) * g.Cout_real + c_first + 2 * cp) * 2;
                *reinterpret_cast<float4*>(dst) = make_float4(s0, q0, s1, q1);
              }
            }
          }

          // ---- coalesced row stores: 4 lanes x 16 B per 64-byte row slice, 8 rows per warp instruction ----
          {
            constexpr int LPR = EB / 8, RPI = 32 / LPR;
            const int chunk = lane % LPR;
            uint8_t* ybase = reinterpret_cast<uint8_t*>(p.y);
            // rows r0, r0 + 32, ...: constant swizzle term -> one staging base address per thread
            const int r0 = we * RPI + lane / LPR;
            const uint32_t sb = stg + r0 * (EB * 2) + ((chunk ^ ((r0 >> 1) & 3)) * 16);
            const uint32_t rb = rowoffg + r0 * 8;
#pragma unroll
            for (int k = 0; k < 128 / (4 * RPI); ++k) {
              const long long off = lds64(rb + k * (4 * RPI * 8));
              if (off >= 0) {
                const uint4 v = lds128(sb + k * (4 * RPI * EB * 2));
                *reinterpret_cast<uint4*>(ybase + (off + chan_off) * 2 + chunk * 16) = v;
              }
            }
          }
          bar_sync_id(barid);   // staging (and, after the last pass, rowoff) are reused
        }
        // ---- fused InstanceNorm finalisation: the group that completes the image's last tile turns the partials
        // into (mean, rstd).  Release: every partial of this group is written (barrier above) and fenced before the
        // arrival; acquire: fence after observing the final count.  Fixed summation order -> deterministic.
        if (p.mean_rstd != nullptr && p.epilogue == NG_EPI_RAW && p.stat_partials != nullptr && !dummy) {
          int* flagw = reinterpret_cast<int*>(rowoff) + Cfg::EG * 256 + grp;   // flags live behind the EG row-offset tables
          __threadfence();
          bar_sync_id(barid);
          if (et == 0) {
            const int old = atomicAdd(&p.tile_counters[n], 1);
            const int last = old == p.arrivals_per_image - 1;
            if (last) p.tile_counters[n] = 0;                            // leave the counters zero for the next launch
            *flagw = last;
          }
          bar_sync_id(barid);
          if (*flagw) {
            __threadfence();
            const int Cr = g.Cout_real;
            const float2* part = reinterpret_cast<const float2*>(p.stat_partials) + (size_t)n * p.stat_slots * Cr;
            for (int c = et; c < Cr; c += 128) {
              float s0 = 0.f, q0 = 0.f;
              for (int k = 0; k < p.stat_slots; ++k) {
                const float2 v = __ldcg(part + (size_t)k * Cr + c);
                s0 += v.x; q0 += v.y;
              }
              const float m = s0 * p.inv_count;
              const float var = fmaxf(q0 * p.inv_count - m * m, 0.f);
              reinterpret_cast<float2*>(p.mean_rstd)[(size_t)n * Cr + c] = make_float2(m, rsqrtf(var + 1e-5f));
            }
          }
          bar_sync_id(barid);                                            // flag word is reused by the next tile
        }
      }
      if (++as == 2) { as = 0; as_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (CS > 1) cluster_sync_all();     // no CTA leaves while a peer may still multicast into it
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cfg::TMEM_COLS)
                 : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;
static std::mutex g_mu;

int get_tensor_map_encoder(PFN_cuTensorMapEncodeTiled_v12000* out) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (!g_encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
      set_error("cuTensorMapEncodeTiled unavailable (%s)", cudaGetErrorString(e));
      return NG_E_DRIVER;
    }
    g_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  }
  if (out) *out = g_encode;
  return NG_OK;
}

static int get_encode() { return get_tensor_map_encoder(nullptr); }

int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return dev;
}

struct TmapKey {
  uint64_t base;
  uint64_t dims[4], strides[3];
  uint32_t box[4], estr[4];
  uint32_t dt, rank, sw, promo;
  bool operator==(const TmapKey& o) const { return memcmp(this, &o, sizeof(TmapKey)) == 0; }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    const uint64_t* w = reinterpret_cast<const uint64_t*>(&k);
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < sizeof(TmapKey) / 8; ++i) { h ^= w[i]; h *= 1099511628211ull; }
    return (size_t)h;
  }
};
static_assert(sizeof(TmapKey) % 8 == 0, "TmapKey is hashed as 64-bit words");
static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> g_tmaps;
static std::mutex g_tmap_mu;

int cached_tensor_map(CUtensorMap* out, CUtensorMapDataType dt, int rank, const void* base, const cuuint64_t* dims,
                      const cuuint64_t* strides, const cuuint32_t* box, const cuuint32_t* estr, CUtensorMapSwizzle sw,
                      CUtensorMapL2promotion promo, int* cres) {
  *cres = 0;
  int r = get_encode();
  if (r) return r;
  TmapKey k;
  memset(&k, 0, sizeof(k));
  k.base = (uint64_t)(uintptr_t)base;
  for (int i = 0; i < rank; ++i) { k.dims[i] = dims[i]; k.box[i] = box[i]; k.estr[i] = estr[i]; }
  for (int i = 0; i + 1 < rank; ++i) k.strides[i] = strides[i];
  k.dt = (uint32_t)dt; k.rank = (uint32_t)rank; k.sw = (uint32_t)sw; k.promo = (uint32_t)promo;
  {
    std::lock_guard<std::mutex> lk(g_tmap_mu);
    auto it = g_tmaps.find(k);
    if (it != g_tmaps.end()) { *out = it->second; return NG_OK; }
  }
  CUtensorMap tm;
  CUresult cr = g_encode(&tm, dt, (cuuint32_t)rank, const_cast<void*>(base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, sw, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) { *cres = (int)cr; return NG_E_DRIVER; }
  {
    std::lock_guard<std::mutex> lk(g_tmap_mu);
    if (g_tmaps.size() > 16384) g_tmaps.clear();      // bounded: descriptors are cheap to rebuild
    g_tmaps.emplace(k, tm);
  }
  *out = tm;
  return NG_OK;
}

static void choose_patch(int VH, int VW, int S, int& BH, int& BW) {
  long long best = -1;
  BH = 1; BW = 1;
  for (int bw = 1; bw <= 128 && bw * S <= 256 && bw <= VW; ++bw) {
    int bh = 128 / bw;
    if (bh * S > 256) bh = 256 / S;
    if (bh > VH) bh = VH;            // taller than the image only wastes rows
    const long long tiles = (long long)((VH + bh - 1) / bh) * ((VW + bw - 1) / bw);
    // fewer tiles wins; ties prefer wider rows (longer contiguous stores)
    const long long score = tiles * 1024 - bw;
    if (best < 0 || score < best) { best = score; BH = bh; BW = bw; }
  }
}

// Row-tap eligibility: K x 1 stride-1 correlation over 64 stored channels into 64 output channels whose taps are
// consecutive rows at one column offset (the row-merged generator stem), image at least one patch wide.
static bool row_tap_ok(const ng_conv_args& a, const ConvGeom& g) {
  static const int env = [] { const char* v = getenv("NIRGAN_B200_ROWTAP"); return v ? atoi(v) : 1; }();   // C++11 magic static: thread-safe
  if (!env) return false;
  if (a.form != NG_FORM_GATHER || a.sgn != 1 || a.stride != 1 || a.KW != 1 || a.KH != RT_KH) return false;
  if (a.Cin != 64 || a.Cout != 64 || a.epilogue == NG_EPI_HEAD || g.ntaps != RT_KH) return false;
  if (g.VW < RT_BW || g.VH < RT_BH) return false;
  for (int t = 0; t < RT_KH; ++t)
    if (g.taps[t].dx != g.taps[0].dx || g.taps[t].dy != g.taps[0].dy + t) return false;
  return true;
}

static void pick_patch(const ng_conv_args& a, const ConvGeom& g, int& BH, int& BW) {
  if (row_tap_ok(a, g)) { BH = RT_BH; BW = RT_BW; }
  else choose_patch(g.VH, g.VW, g.S, BH, BW);
}

template <int BN, int KC, int CS, bool RT = false, int EGW = 2>
static int launch_tc(const ng_conv_args& a, const ConvGeom& g, cudaStream_t st) {
  using Cfg = TcCfg<BN, KC, RT, EGW>;
  int r = get_encode();
  if (r) return r;
  TcParams p;
  memset(&p, 0, sizeof(p));
  p.g = g;
  pick_patch(a, g, p.BH, p.BW);
  p.patches_y = (g.VH + p.BH - 1) / p.BH;
  p.patches_x = (g.VW + p.BW - 1) / p.BW;
  p.co_tiles = g.Cout / BN;
  const long long tiles = (long long)g.B * p.patches_y * p.patches_x * g.nphase * p.co_tiles;
  NG_REQUIRE(tiles > 0 && tiles < (1ll << 30), NG_E_SHAPE, "conv_tc: tile count out of range");
  p.total_tiles = (int)tiles;
  p.group_items = g.B * p.patches_y * p.patches_x;
  p.groups_per = (p.group_items + CS - 1) / CS;
  p.total_groups = p.co_tiles * g.nphase * p.groups_per;
  NG_REQUIRE(g.nphase == 1 || g.nphase == 4, NG_E_UNSUPPORTED, "conv_tc: 1 or 4 phases");
  NG_REQUIRE((long long)p.total_groups * p.groups_per < (1ll << 42) && (long long)p.group_items * p.patches_x < (1ll << 42),
             NG_E_SHAPE, "conv_tc: problem too large for the reciprocal tile decode");
  p.mg_groups_per = ((1ull << 42) + (unsigned)p.groups_per - 1) / (unsigned)p.groups_per;
  p.mg_patches_x = ((1ull << 42) + (unsigned)p.patches_x - 1) / (unsigned)p.patches_x;
  p.mg_patches_y = ((1ull << 42) + (unsigned)p.patches_y - 1) / (unsigned)p.patches_y;
  p.epilogue = a.epilogue; p.act = a.act; p.crop = a.crop; p.slope = a.slope;
  p.bf16 = a.dtype == NG_BF16;
  p.stat_slots = (g.merged ? 4 : g.nphase) * p.patches_y * p.patches_x;
  p.bias = a.bias; p.y = a.y; p.stat_partials = a.stat_partials;
  p.mean_rstd = a.mean_rstd; p.tile_counters = a.tile_counters;
  p.stat_acc = reinterpret_cast<unsigned long long*>(a.stat_acc);
  p.image_minor = (a.stat_acc != nullptr && g.B > 1) ? 1 : 0;
  p.mg_B = ((1ull << 42) + (unsigned)g.B - 1) / (unsigned)g.B;
  // Measured on B200 (profiles/r2c): fetching / multiplying only the 9 non-zero (shift, phase) slabs of a merged-phase
  // ConvTranspose is SLOWER than the dense N = 256 form (u2 0.142 -> 0.203 ms, u1 0.110 -> 0.121 ms per 32 tiles): an
  // N = 64 tcgen05.mma re-reads the whole 128-row A tile from shared memory for a quarter of the columns, so the 72
  // narrow MMAs cost more than the 32 wide ones they replace.  Kept as a tested switch.
  // NIRGAN_B200_SPARSE_MERGED: 0 (default) = dense N = 4 * Cout_real, 1 = one N = Cout_real chain per non-zero slab (the
  // slower form above), 2 = banded: with the slots in Gray order every shift's slots are contiguous, so each shift is ONE
  // chain of N = 256 / 128 / 128 / 64 (Cout_real = 64) instead of four times N = 256 -- a third fewer tensor-pipe cycles,
  // but no faster (r3u: u2 0.163-0.185 vs 0.153 ms, u1 0.107-0.116 vs 0.114 ms): these short-K tiles are bound by the
  // accumulator drain and the 268 MB of output, not by the MMAs.
  static const int sparse_env = [] { const char* v = getenv("NIRGAN_B200_SPARSE_MERGED"); return v ? atoi(v) : 0; }();
  p.sparse_merged = (g.merged && sparse_env && CS == 1) ? sparse_env : 0;
  p.arrivals_per_image = p.patches_y * p.patches_x * g.nphase * p.co_tiles * Cfg::EGT;
  p.inv_count = 1.0f / ((float)g.Hout * (float)g.Wout);

  CUtensorMap tmA, tmB;
  const CUtensorMapDataType dt = a.dtype == NG_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  const CUtensorMapSwizzle sw = KC == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_32B;
  {
    cuuint64_t dims[4] = {(cuuint64_t)g.Cin, (cuuint64_t)g.Wb, (cuuint64_t)g.Hb, (cuuint64_t)g.B};
    cuuint64_t strides[3] = {(cuuint64_t)g.Cin * 2, (cuuint64_t)g.Wb * g.Cin * 2, (cuuint64_t)g.Hb * g.Wb * g.Cin * 2};
    cuuint32_t box[4] = {(cuuint32_t)KC, (cuuint32_t)(p.BW * g.S), (cuuint32_t)(p.BH * g.S), 1};
    if (RT) box[2] = (cuuint32_t)(RT_BH + RT_KH - 1);       // the haloed patch: every tap's rows in one box
    cuuint32_t estr[4] = {1, (cuuint32_t)g.S, (cuuint32_t)g.S, 1};
    int cr = 0;
    const int er = cached_tensor_map(&tmA, dt, 4, a.x, dims, strides, box, estr, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, &cr);
    NG_REQUIRE(er == NG_OK, NG_E_DRIVER, "cuTensorMapEncodeTiled(A) failed: %d (dims %d %d %d %d box %d %d %d)",
               cr, g.Cin, g.Wb, g.Hb, g.B, KC, p.BW * g.S, p.BH * g.S);
  }
  {
    const int taps_total = g.merged ? g.ntaps : a.KH * a.KW;
    cuuint64_t dims[2] = {(cuuint64_t)g.Cin, (cuuint64_t)taps_total * g.Cout};
    cuuint64_t strides[1] = {(cuuint64_t)g.Cin * 2};
    // merged phases: one box = one (shift, phase) slab of Cout_real rows (only the non-zero slabs are fetched)
    cuuint32_t box[2] = {(cuuint32_t)KC, (cuuint32_t)(p.sparse_merged ? g.Cout_real : BN / CS)};
    cuuint32_t estr[2] = {1, 1};
    int cr = 0;
    const int er = cached_tensor_map(&tmB, dt, 2, a.w, dims, strides, box, estr, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, &cr);
    NG_REQUIRE(er == NG_OK, NG_E_DRIVER, "cuTensorMapEncodeTiled(B) failed: %d", cr);
  }
  // per instantiation AND per device (kernel attributes and the SM count belong to a device); thread-safe
  static PerDeviceOnce once;
  static int max_ctas_dev[64];
  const int dev = current_device();
  if (once.needed(dev)) {
    int e = check_cuda(cudaFuncSetAttribute(conv_tc_kernel<BN, KC, CS, RT, EGW>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            Cfg::SMEM_BYTES), "cudaFuncSetAttribute(conv_tc)");
    if (e) return e;
    int max_ctas = num_sms() / CS * CS;
    if (CS > 1) {
      cudaLaunchConfig_t qc = {};
      qc.gridDim = dim3(max_ctas); qc.blockDim = dim3(Cfg::THREADS); qc.dynamicSmemBytes = Cfg::SMEM_BYTES;
      cudaLaunchAttribute qa[1];
      qa[0].id = cudaLaunchAttributeClusterDimension;
      qa[0].val.clusterDim.x = CS; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
      qc.attrs = qa; qc.numAttrs = 1;
      int ncl = 0;
      if (cudaOccupancyMaxActiveClusters(&ncl, conv_tc_kernel<BN, KC, CS, RT, EGW>, &qc) == cudaSuccess && ncl > 0 &&
          ncl * CS < max_ctas)
        max_ctas = ncl * CS;      // persistent kernel: every cluster must be co-resident
    }
    __atomic_store_n(&max_ctas_dev[dev & 63], max_ctas, __ATOMIC_RELEASE);
    once.done(dev);
  }
  const int max_ctas = __atomic_load_n(&max_ctas_dev[dev & 63], __ATOMIC_ACQUIRE);
  long long want = (long long)p.total_groups * CS;
  const int grid = (int)(want < max_ctas ? want : max_ctas);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(Cfg::THREADS); cfg.dynamicSmemBytes = Cfg::SMEM_BYTES; cfg.stream = st;
  cudaLaunchAttribute attrs[1];
  attrs[0].id = cudaLaunchAttributeClusterDimension;
  attrs[0].val.clusterDim.x = CS; attrs[0].val.clusterDim.y = 1; attrs[0].val.clusterDim.z = 1;
  cfg.attrs = attrs; cfg.numAttrs = 1;
  int e = check_cuda(cudaLaunchKernelEx(&cfg, conv_tc_kernel<BN, KC, CS, RT, EGW>, tmA, tmB, p), "conv_tc_kernel launch");
  if (e) return e;
  NG_LAUNCH_CHECK("conv_tc_kernel");
  return NG_OK;
}

// Generator stem straight from the NCHW fp32 tiles (see the DS notes above).  y: [B][H1][W1][64] pre-norm, 16-bit;
// w_packed: [7][64][32] (ng_pack_weight_rowmerged with 4 channel slots).
int conv_tc_stem_direct(const float* src, int cin, int B, int H, int W, int wrap, const void* w_packed, int dtype, void* y,
                        float* stat_partials, long long* stat_acc, cudaStream_t st) {
  constexpr int BN = 64, KC = 32;
  using Cfg = TcCfg<BN, KC, true, 4, true>;
  NG_REQUIRE(dtype == NG_F16 || dtype == NG_BF16, NG_E_UNSUPPORTED, "stem_direct: operands must be f16 or bf16");
  NG_REQUIRE(src && w_packed && y && cin >= 1 && cin <= 4 && B > 0, NG_E_ARG, "stem_direct: bad arguments (cin <= 4)");
  NG_REQUIRE(((uintptr_t)w_packed & 15) == 0 && ((uintptr_t)y & 15) == 0 && ((uintptr_t)stat_acc & 15) == 0, NG_E_ALIGN,
             "stem_direct: tensors must be 16-byte aligned");
  const int H1 = H + 2 * wrap, W1 = W + 2 * wrap;
  NG_REQUIRE(wrap >= 0 && wrap < H && wrap < W && H1 > 3 && W1 > 3, NG_E_SHAPE, "stem_direct: reflect padding needs a larger tile");
  NG_REQUIRE(H1 >= RT_BH && W1 >= RT_BW, NG_E_SHAPE, "stem_direct: tile smaller than one 8 x 16 patch");
  TcParams p;
  memset(&p, 0, sizeof(p));
  ConvGeom& g = p.g;
  g.B = B; g.Hb = H1 + 6; g.Wb = W1; g.Cin = KC; g.Cout = BN; g.Cout_real = BN;
  g.VH = H1; g.VW = W1; g.S = 1; g.OS = 1; g.Hout = H1; g.Wout = W1; g.nphase = 1; g.ntaps = RT_KH;
  g.phase_tap0[0] = 0; g.phase_tap0[1] = RT_KH;
  for (int t = 0; t < RT_KH; ++t) { g.taps[t].dy = (int16_t)t; g.taps[t].dx = 0; g.taps[t].wrow = t * BN; }
  p.BH = RT_BH; p.BW = RT_BW;
  p.patches_y = (H1 + RT_BH - 1) / RT_BH;
  p.patches_x = (W1 + RT_BW - 1) / RT_BW;
  p.co_tiles = 1;
  const long long tiles = (long long)B * p.patches_y * p.patches_x;
  NG_REQUIRE(tiles > 0 && tiles < (1ll << 30), NG_E_SHAPE, "stem_direct: tile count out of range");
  p.total_tiles = (int)tiles; p.group_items = (int)tiles; p.groups_per = (int)tiles; p.total_groups = (int)tiles;
  p.mg_groups_per = ((1ull << 42) + (unsigned)p.groups_per - 1) / (unsigned)p.groups_per;
  p.mg_patches_x = ((1ull << 42) + (unsigned)p.patches_x - 1) / (unsigned)p.patches_x;
  p.mg_patches_y = ((1ull << 42) + (unsigned)p.patches_y - 1) / (unsigned)p.patches_y;
  p.mg_B = ((1ull << 42) + (unsigned)B - 1) / (unsigned)B;
  p.epilogue = NG_EPI_RAW; p.act = NG_ACT_NONE;
  p.bf16 = dtype == NG_BF16;
  p.stat_slots = p.patches_y * p.patches_x;
  p.y = y; p.stat_partials = stat_partials;
  p.stat_acc = reinterpret_cast<unsigned long long*>(stat_acc);
  p.image_minor = (stat_acc != nullptr && B > 1) ? 1 : 0;
  p.inv_count = 1.0f / ((float)H1 * (float)W1);
  p.src = src; p.src_c = cin; p.src_H = H; p.src_W = W; p.src_wrap = wrap;

  CUtensorMap tmB;
  {
    const CUtensorMapDataType dt = dtype == NG_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    cuuint64_t dims[2] = {(cuuint64_t)KC, (cuuint64_t)RT_KH * BN};
    cuuint64_t strides[1] = {(cuuint64_t)KC * 2};
    cuuint32_t box[2] = {(cuuint32_t)KC, (cuuint32_t)BN};
    cuuint32_t estr[2] = {1, 1};
    int cr = 0;
    const int er = cached_tensor_map(&tmB, dt, 2, w_packed, dims, strides, box, estr, CU_TENSOR_MAP_SWIZZLE_64B,
                                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, &cr);
    NG_REQUIRE(er == NG_OK, NG_E_DRIVER, "stem_direct: cuTensorMapEncodeTiled(W) failed: %d", cr);
  }
  auto kern = conv_tc_kernel<BN, KC, 1, true, 4, true>;
  static PerDeviceOnce once;
  const int dev = current_device();
  if (once.needed(dev)) {
    int e = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES),
                       "cudaFuncSetAttribute(stem_direct)");
    if (e) return e;
    once.done(dev);
  }
  const int sms = num_sms();
  const int grid = (int)(tiles < sms ? tiles : sms);
  kern<<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, st>>>(tmB, tmB, p);
  NG_LAUNCH_CHECK("conv_tc_kernel<direct stem>");
  return NG_OK;
}

static int cluster_size_for(int bn, int kc) {
  // measured on B200: multicast does not shorten the kernel (not L2-bound); kept as a switch
  static const int env = [] {
    const char* v = getenv("NIRGAN_B200_CLUSTER");
    const int e = v ? atoi(v) : 1;
    return (e == 2 || e == 4) ? e : 1;
  }();
  return (bn == 256 && kc == 64) ? env : 1;     // multicast pays on the weight-heavy 256-wide tiles
}

static int tc_block_n(const ng_conv_args& a, const ConvGeom& g) {
  if (a.epilogue == NG_EPI_HEAD || a.Cout == 16) return 16;
  if (g.Cout % 256 == 0) return 256;
  if (g.Cout % 128 == 0) return 128;
  if (g.Cout % 64 == 0) return 64;
  return 0;
}

int conv_tc_stat_slots(const ng_conv_args& a, const ConvGeom& g) {
  int BH, BW;
  pick_patch(a, g, BH, BW);
  return (g.merged ? 4 : g.nphase) * ((g.VH + BH - 1) / BH) * ((g.VW + BW - 1) / BW);
}

int conv_tc(const ng_conv_args& a, const ConvGeom& g, cudaStream_t st) {
  NG_REQUIRE(a.dtype == NG_F16 || a.dtype == NG_BF16, NG_E_UNSUPPORTED, "conv_tc: operands must be f16 or bf16");
  NG_REQUIRE(((uintptr_t)a.x & 15) == 0 && ((uintptr_t)a.w & 15) == 0 && ((uintptr_t)a.y & 15) == 0, NG_E_ALIGN,
             "conv_tc: tensors must be 16-byte aligned");
  const int bn = tc_block_n(a, g);
  const int kc = a.Cin % 64 == 0 ? 64 : (a.Cin == 16 ? 16 : 0);
  NG_REQUIRE(bn != 0 && kc != 0, NG_E_UNSUPPORTED, "conv_tc: Cin %d / Cout %d not tileable", a.Cin, a.Cout);
  NG_REQUIRE(a.epilogue != NG_EPI_HEAD || a.Cout == 16, NG_E_SHAPE, "conv_tc: head epilogue expects Cout stored as 16");
  NG_REQUIRE(bn != 16 || (a.stat_partials == nullptr && a.stat_acc == nullptr), NG_E_UNSUPPORTED,
             "conv_tc: no InstanceNorm statistics for 16-channel outputs");
  NG_REQUIRE(a.stat_acc == nullptr || (a.epilogue == NG_EPI_RAW && ((uintptr_t)a.stat_acc & 15) == 0), NG_E_ARG,
             "conv_tc: stat_acc needs the RAW epilogue and 16-byte alignment");
  NG_REQUIRE(a.stat_acc == nullptr || a.mean_rstd == nullptr, NG_E_ARG,
             "conv_tc: fused finalisation works on stat_partials, not on stat_acc");
  NG_REQUIRE((a.mean_rstd == nullptr) == (a.tile_counters == nullptr), NG_E_ARG,
             "conv_tc: fused finalisation needs both mean_rstd and tile_counters");
  NG_REQUIRE(a.mean_rstd == nullptr || (a.stat_partials != nullptr && a.epilogue == NG_EPI_RAW), NG_E_ARG,
             "conv_tc: fused finalisation needs the RAW epilogue with stat_partials");
  if (bn == 16 && kc == 64) return launch_tc<16, 64, 1>(a, g, st);
  if (bn == 64 && kc == 16) return launch_tc<64, 16, 1>(a, g, st);
  if (bn == 128 && kc == 16) return launch_tc<128, 16, 1>(a, g, st);
  if (bn == 256 && kc == 16) return launch_tc<256, 16, 1>(a, g, st);
  static const int wide_env = [] { const char* v = getenv("NIRGAN_B200_EPI4"); return v ? atoi(v) : 2; }();
  // 64-wide tiles have two drain passes; with four epilogue groups two tiles (both accumulator buffers) drain at once
  if (bn == 64 && kc == 64 && row_tap_ok(a, g))
    return wide_env >= 2 ? launch_tc<64, 64, 1, true, 4>(a, g, st) : launch_tc<64, 64, 1, true>(a, g, st);
  if (bn == 64 && kc == 64) return wide_env >= 2 ? launch_tc<64, 64, 1, false, 4>(a, g, st) : launch_tc<64, 64, 1>(a, g, st);
  // very short K loops on 256-wide tiles (the merged-phase up-convolution to full resolution: 8 operand stages per
  // tile against 8 drain passes) are bound by the epilogue drain: four epilogue groups.  Measured on B200: 0.222 ->
  // 0.179 ms for that layer; for 9..18 K iterations (down convs, first up-conv) the fourth operand stage that EG = 4
  // costs matters more (+7..13 %), so those stay on two groups.
  int max_kiters = 0;
  for (int ph = 0; ph < g.nphase; ++ph) {
    const int it = (g.phase_tap0[ph + 1] - g.phase_tap0[ph]) * (a.Cin / 64);
    if (it > max_kiters) max_kiters = it;
  }
  const bool wide = wide_env && kc == 64 && max_kiters <= 8;
  if (bn == 128 && kc == 64) return wide_env >= 3 ? launch_tc<128, 64, 1, false, 4>(a, g, st) : launch_tc<128, 64, 1>(a, g, st);
  if (bn == 256 && kc == 64 && wide) return launch_tc<256, 64, 1, false, 4>(a, g, st);
  if (bn == 256 && kc == 64) {
    const int cs = g.merged ? 1 : cluster_size_for(bn, kc);
    if (cs == 4) return launch_tc<256, 64, 4>(a, g, st);
    if (cs == 2) return launch_tc<256, 64, 2>(a, g, st);
    return launch_tc<256, 64, 1>(a, g, st);
  }
  set_error("conv_tc: no kernel for BN %d KC %d", bn, kc);
  return NG_E_UNSUPPORTED;
}

}  // namespace ng
