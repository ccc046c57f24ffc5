// tcgen05 / TMEM / TMA weight-gradient kernel for sm_100a.
//
//   dW[tap][n][k] = sum over (image, virtual pixel) of dY[img, OS*v + phase][n] * X[img, S*v + tap offset][k]
//
// GEMM view per tap: M = 128 output channels (n), N = BN input channels (k), reduction K = pixels.  Both operands are
// stored pixel-major with channels contiguous (NHWC), i.e. *MN-major* for the tensor core: a pipeline stage holds 64
// pixels as [64 rows][64 channels = 128 B] slabs (TMA 128-byte swizzle) -- 2 slabs of dY and BN/64 slabs of X fetched by
// 4-D TMA boxes at the tap's offset (element stride S / OS for strided and transposed convolutions, out-of-bounds =
// zero fill = zero padding).  One elected thread issues tcgen05.mma (M=128, N=BN, K=16 pixels, both operands MN-major)
// into a double-buffered fp32 TMEM accumulator.
//
// The pixel range is split across CTAs ("split-K"): work unit = (split, tap, 128-channel n tile); every unit writes its
// fp32 128 x BN partial to the caller's workspace [split][tap][Cout][Cin] and a second kernel sums the splits in a fixed
// order, so the result is deterministic (no atomics).
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM alloc + MMA issuer, warps 2..5 = epilogue.
#include "tc_common.cuh"
#include <stdlib.h>

namespace ng {

constexpr int WG_PIXELS = 64;                       // pixels (GEMM-K) per pipeline stage
constexpr int WG_SLAB = WG_PIXELS * 128;            // one 64-channel slab: 8 KB

struct WgParams {
  ConvGeom g;
  int BH, BW;                 // virtual-pixel patch, BH*BW == 64
  int patches_y, patches_x;
  int P;                      // B * patches_y * patches_x
  int n_tiles;                // ceil(Cout / 128)
  int splits, pps;            // pixel-range splits, patches per split
  int total_units;            // splits * ngroups * n_tiles
  int bf16;
  float* partials;            // [splits][ntaps][Cout][Cin]
  // tap groups: a work unit walks grp_cnt[k] (<= NT) consecutive taps of ONE output phase starting at grp_first[k]; the
  // dY tile of a pipeline stage is loaded once per group instead of once per tap
  int ngroups;
  unsigned char grp_first[64], grp_cnt[64];
  int tap_stride;             // tap index distance between the taps of a group (1; 3 = the kh taps of one kw of a 3x3)
  int k_tiles;                // BN-wide input-channel tiles per tap (1 unless BN < Cin)
};

// NT = taps handled per pipeline stage ("tap group", NT*BN <= 512 TMEM columns): work unit = (split, tap group, n tile); a
// stage holds the dY tile once plus the X tile of every tap of the group, and each tap accumulates into its own BN-column
// TMEM slice.  For the thin layers (few input channels, many taps) the dY operand dominates the L2 -> shared-memory
// traffic and is fetched once per GROUP instead of once per tap: PatchGAN input layer 16 taps x 16 channels and the
// generator stem 7 x 64 in one group; 64-channel layers in groups of 3-4, 128-channel layers in pairs.  Two accumulator
// sets (the epilogue of one unit overlaps the MMAs of the next) whenever 2*NT*BN columns fit.
// RP ("row patch", the row-merged generator stem: NT taps that are pure row shifts of an 8 x 8 pixel patch over 64 stored
// channels): the X operand of a stage is ONE haloed patch of (8 + NT - 1) rows x 8 pixels and tap t's operand is the
// 64-pixel window that starts t rows (t x 1024 bytes = whole swizzle atoms) into it -- 14 KB per stage instead of
// NT x 8 KB, which is what bounded this layer (L2 -> shared-memory traffic).
// RP with BN = 128, NT = 3 serves the stride-1 3x3 convolutions with 128 / 256 input channels (the 18 ResnetBlock
// convolutions): a group is the three kh taps of one kw (tap stride 3), the X operand two 64-channel slabs of the haloed
// 10 x 8 pixel patch (20 KB for three taps instead of 3 x 16 KB), the input channels go in k tiles of 128.  Per stage
// 36 KB feed 3 x 128x128x64 MACs = 87 MAC/B, against 44 MAC/B for the one-tap 128 x 256 unit that was L2 -> SM bound.
constexpr int RP_BW = 8, RP_BH = 8;

template <int BN, int NT, bool RP = false>
struct WgCfg {
  static constexpr int A_BYTES = 2 * WG_SLAB;
  // X operand of one tap: BN/64 slabs of [64 px][128 B] (128-byte swizzle), or for BN == 16 one slab of [64 px][32 B]
  static constexpr int B_TAP_BYTES = BN >= 64 ? (BN / 64) * WG_SLAB : WG_PIXELS * 32;
  static constexpr int RP_SLAB = (RP_BH + NT - 1) * RP_BW * 128;   // one 64-channel slab of the haloed row patch
  static constexpr int B_BYTES = RP ? (BN / 64) * RP_SLAB : NT * B_TAP_BYTES;
  static constexpr int B_LOADS = BN >= 64 ? BN / 64 : 1;
  static constexpr int B_KSTEP = BN >= 64 ? 2048 : 512;            // bytes per 16 pixel rows
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES_RAW = (216 * 1024) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int NACC = 2 * NT * BN <= 512 ? 2 : 1;
  static constexpr int ACC_COLS = NACC * NT * BN;
  static constexpr int TMEM_COLS = ACC_COLS <= 32 ? 32 : (ACC_COLS <= 64 ? 64 : (ACC_COLS <= 128 ? 128 : (ACC_COLS <= 256 ? 256 : 512)));
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 256 /*barriers*/ + 1024 /*alignment slack*/;
  static_assert(ACC_COLS <= 512, "accumulators exceed tensor memory");
  static_assert(STAGES >= 2, "pipeline needs at least two stages");
};

template <int BN, int NT, bool RP = false>
__global__ void __launch_bounds__(192, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmX,
                const __grid_constant__ WgParams p) {
  using Cfg = WgCfg<BN, NT, RP>;
  static_assert(!RP || BN == 64 || BN == 128, "row-patch variant: one or two 64-channel slabs");
  constexpr int NACC = Cfg::NACC;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tfull_bar = bars + 2 * STAGES;
  uint64_t* tempty_bar = bars + 2 * STAGES + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const ConvGeom& g = p.g;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(smem_u32(&full_bar[s]), 1); mbar_init(smem_u32(&empty_bar[s]), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(smem_u32(&tfull_bar[s]), 1); mbar_init(smem_u32(&tempty_bar[s]), 128); }
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
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int per_split = p.ngroups * p.n_tiles * p.k_tiles;
  auto decode = [&](int u, int& split, int& tap, int& cnt, int& nt, int& kt, int& pt0, int& pt1) {
    split = u / per_split;
    int r = u - split * per_split;
    kt = r % p.k_tiles;                     // input-channel tile (BN channels)
    r /= p.k_tiles;
    const int grp = r / p.n_tiles;
    tap = p.grp_first[grp];                 // first tap of the group
    cnt = p.grp_cnt[grp];                   // taps walked inside the unit (<= NT)
    nt = r - grp * p.n_tiles;
    pt0 = split * p.pps;
    pt1 = min(p.P, pt0 + p.pps);
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int u = blockIdx.x; u < p.total_units; u += gridDim.x) {
        int split, tap, cnt, nt, kt, pt0, pt1;
        decode(u, split, tap, cnt, nt, kt, pt0, pt1);
        int ph = 0;
        while (tap >= g.phase_tap0[ph + 1]) ++ph;
        const int oy0 = g.phase_oy[ph], ox0 = g.phase_ox[ph];
        const uint32_t stage_tx = RP ? (uint32_t)Cfg::STAGE_BYTES : (uint32_t)(Cfg::A_BYTES + cnt * Cfg::B_TAP_BYTES);
        const int per_img = p.patches_y * p.patches_x;
        for (int pt = pt0; pt < pt1; ++pt) {
          const int n = pt / per_img;
          const int r = pt - n * per_img;
          const int py = r / p.patches_x, px = r - py * p.patches_x;
          const int vi0 = py * p.BH, vj0 = px * p.BW;
          mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
          const uint32_t fb = smem_u32(&full_bar[stage]);
          const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          mbar_expect_tx(fb, stage_tx);
#pragma unroll
          for (int j = 0; j < 2; ++j)
            tma_load_4d(&tmY, fb, sa + j * WG_SLAB, nt * 128 + 64 * j, g.OS * vj0 + ox0, g.OS * vi0 + oy0, n);
          if constexpr (RP) {
            // one haloed patch: rows vi0 + dy0 .. vi0 + dy0 + 8 + NT - 2 (tmX's box is that tall), 8 pixels wide
#pragma unroll
            for (int j = 0; j < Cfg::B_LOADS; ++j)
              tma_load_4d(&tmX, fb, sa + Cfg::A_BYTES + j * Cfg::RP_SLAB, kt * BN + 64 * j, vj0 + g.taps[tap].dx,
                          vi0 + g.taps[tap].dy, n);
          } else {
#pragma unroll 1
            for (int t = 0; t < cnt; ++t)
#pragma unroll
              for (int j = 0; j < Cfg::B_LOADS; ++j)
                tma_load_4d(&tmX, fb, sa + Cfg::A_BYTES + t * Cfg::B_TAP_BYTES + j * WG_SLAB, kt * BN + 64 * j,
                            g.S * vj0 + g.taps[tap + t].dx, g.S * vi0 + g.taps[tap + t].dy, n);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      // c = f32, a/b format, both operands MN-major (bits 15, 16), N = BN, M = 128
      const uint32_t idesc = (1u << 4) | ((uint32_t)(p.bf16 ? 1 : 0) << 7) | ((uint32_t)(p.bf16 ? 1 : 0) << 10) |
                             (1u << 15) | (1u << 16) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      uint32_t stage = 0, phase = 0, as = 0, as_phase = 0;
      for (int u = blockIdx.x; u < p.total_units; u += gridDim.x) {
        int split, tap, cnt, nt, kt, pt0, pt1;
        decode(u, split, tap, cnt, nt, kt, pt0, pt1);
        const int kiters = pt1 - pt0;
        mbar_wait(smem_u32(&tempty_bar[as]), as_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_c = tmem_base + as * (NT * BN);
        for (int kit = 0; kit < kiters; ++kit) {
          mbar_wait(smem_u32(&full_bar[stage]), phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint64_t adesc = make_mnmajor_desc(sa, WG_SLAB, 1024);
#pragma unroll 1
          for (int t = 0; t < cnt; ++t) {
            // RP: tap t = the patch t rows (t * 8 pixels * 128 B = t swizzle atoms) further down
            const uint32_t sb = sa + Cfg::A_BYTES + t * (RP ? RP_BW * 128 : Cfg::B_TAP_BYTES);
            const uint64_t bdesc = BN >= 64 ? make_mnmajor_desc(sb, RP ? Cfg::RP_SLAB : WG_SLAB, 1024)
                                            : make_mnmajor_desc_sw32(sb, 256);
#pragma unroll
            for (int k = 0; k < WG_PIXELS / 16; ++k)     // 16 pixel rows per K step
              umma_f16(tmem_c + t * BN, adesc + (uint64_t)(k * (2048 >> 4)), bdesc + (uint64_t)(k * (Cfg::B_KSTEP >> 4)),
                       idesc, (uint32_t)((kit | k) != 0));
          }
          umma_commit(smem_u32(&empty_bar[stage]));
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(smem_u32(&tfull_bar[as]));
        if (++as == NACC) { as = 0; as_phase ^= 1; }
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue (warps 2..5): TMEM -> fp32 partial =====================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    uint32_t as = 0, as_phase = 0;
    for (int u = blockIdx.x; u < p.total_units; u += gridDim.x) {
      int split, tap, cnt, nt, kt, pt0, pt1;
      decode(u, split, tap, cnt, nt, kt, pt0, pt1);
      const int n = nt * 128 + row;
      mbar_wait(smem_u32(&tfull_bar[as]), as_phase);
      tc_fence_after();
#pragma unroll 1
      for (int t = 0; t < cnt; ++t) {
        const uint32_t taddr = tmem_base + as * (NT * BN) + t * BN + ((uint32_t)(q * 32) << 16);
        // packed weight-gradient row of this tap: wrow = (kh*KW + kw) * Cout (phased geometries enumerate taps by phase)
        const int tp = tap + t * p.tap_stride;
        float* dst = p.partials + ((size_t)split * g.ntaps * g.Cout + g.taps[tp].wrow + (n < g.Cout ? n : 0)) * (size_t)g.Cin +
                     kt * BN;
        if constexpr (BN >= 32) {
#pragma unroll 1
          for (int c = 0; c < BN / 32; ++c) {
            uint32_t r[32];
            tmem_ld32(taddr + c * 32, r);
            tmem_ld_wait();
            if (n < g.Cout) {
#pragma unroll
              for (int k = 0; k < 8; ++k)
                *reinterpret_cast<float4*>(dst + c * 32 + 4 * k) =
                    make_float4(__uint_as_float(r[4 * k]), __uint_as_float(r[4 * k + 1]), __uint_as_float(r[4 * k + 2]),
                                __uint_as_float(r[4 * k + 3]));
            }
          }
        } else {
          uint32_t r[16];
          tmem_ld16(taddr, r);
          tmem_ld_wait();
          if (n < g.Cout) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              *reinterpret_cast<float4*>(dst + 4 * k) =
                  make_float4(__uint_as_float(r[4 * k]), __uint_as_float(r[4 * k + 1]), __uint_as_float(r[4 * k + 2]),
                              __uint_as_float(r[4 * k + 3]));
          }
        }
      }
      tc_fence_before();
      mbar_arrive(smem_u32(&tempty_bar[as]));
      if (++as == NACC) { as = 0; as_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cfg::TMEM_COLS)
                 : "memory");
  }
}

// dw[i] = sum_s partials[s][i]  (fixed order -> deterministic)
__global__ void __launch_bounds__(256)
wgrad_reduce_kernel(const float4* __restrict__ part, int splits, long long n4, float4* __restrict__ dw) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 s = part[i];
    int k = 1;
    for (; k + 3 < splits; k += 4) {          // four independent loads in flight, added in split order
      const float4 a = part[(size_t)k * n4 + i], b = part[(size_t)(k + 1) * n4 + i], c = part[(size_t)(k + 2) * n4 + i],
                   d = part[(size_t)(k + 3) * n4 + i];
      s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w;
      s.x += b.x; s.y += b.y; s.z += b.z; s.w += b.w;
      s.x += c.x; s.y += c.y; s.z += c.z; s.w += c.w;
      s.x += d.x; s.y += d.y; s.z += d.z; s.w += d.w;
    }
    for (; k < splits; ++k) {
      const float4 v = part[(size_t)k * n4 + i];
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    dw[i] = s;
  }
}

// Same sum for small outputs with many splits (thin layers: 1 k float4 outputs x 134 splits): a thread per output would
// walk the splits as one chain of dependent-latency loads, so eight threads share an output -- thread (o, ks) adds splits
// ks, ks + 8, ... four loads at a time, and the eight slices are then added in slice order (fixed order -> deterministic).
__global__ void __launch_bounds__(256)
wgrad_reduce_sliced_kernel(const float4* __restrict__ part, int splits, long long n4, float4* __restrict__ dw) {
  __shared__ float4 red[8][32];
  const int o = threadIdx.x & 31, ks = threadIdx.x >> 5;
  const long long i = (long long)blockIdx.x * 32 + o;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (i < n4) {
    int k = ks;
    for (; k + 24 < splits; k += 32) {
      const float4 a = part[(size_t)k * n4 + i], b = part[(size_t)(k + 8) * n4 + i], c = part[(size_t)(k + 16) * n4 + i],
                   d = part[(size_t)(k + 24) * n4 + i];
      s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w;
      s.x += b.x; s.y += b.y; s.z += b.z; s.w += b.w;
      s.x += c.x; s.y += c.y; s.z += c.z; s.w += c.w;
      s.x += d.x; s.y += d.y; s.z += d.z; s.w += d.w;
    }
    for (; k < splits; k += 8) {
      const float4 a = part[(size_t)k * n4 + i];
      s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w;
    }
  }
  red[ks][o] = s;
  __syncthreads();
  if (ks == 0 && i < n4) {
    float4 t = red[0][o];
#pragma unroll
    for (int j = 1; j < 8; ++j) { const float4 v = red[j][o]; t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w; }
    dw[i] = t;
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static bool row_patch_ok(const ConvGeom& g) {
  static const bool on = [] { const char* e = getenv("NIRGAN_B200_WGRAD_ROWPATCH"); return !(e && e[0] == '0'); }();
  if (!on || g.nphase != 1 || g.Cin != 64 || g.ntaps != 7 || g.S != 1 || g.OS != 1) return false;
  for (int t = 0; t < g.ntaps; ++t)
    if (g.taps[t].dx != g.taps[0].dx || g.taps[t].dy != g.taps[0].dy + t) return false;
  return g.VW >= RP_BW && g.VH >= RP_BH;
}

// stride-1 3x3 with 128 / 256 input channels, taps enumerated kh-major with unit offsets: the row-patch form over
// (kw group) x (three kh taps)
static bool row_patch3_ok(const ConvGeom& g) {
  static const bool on = [] { const char* e = getenv("NIRGAN_B200_WGRAD_ROWPATCH3"); return !(e && e[0] == '0'); }();
  if (!on || g.nphase != 1 || g.ntaps != 9 || g.S != 1 || g.OS != 1 || (g.Cin != 128 && g.Cin != 256)) return false;
  for (int kh = 0; kh < 3; ++kh)
    for (int kw = 0; kw < 3; ++kw) {
      const int t = kh * 3 + kw;
      if (g.taps[t].dy != g.taps[0].dy + kh || g.taps[t].dx != g.taps[0].dx + kw) return false;
    }
  return g.VW >= RP_BW && g.VH >= RP_BH;
}

struct WgPlan {
  int tap_stride, k_tiles;
  int BH, BW, patches_y, patches_x, P, n_tiles, splits, pps, bn;
  int nt;       // taps per stage (compile-time group capacity)
  int ngroups;
  unsigned char grp_first[64], grp_cnt[64];
};

// taps per group by input-channel count (NT * Cin <= 512 TMEM columns; instantiated pairs only):
// (16, 16) PatchGAN input layer; (64, 7) row-merged generator stem; (64, 3) / (64, 4) 3x3 / 4x4 layers with 64 input
// channels; (128, 2) layers with 128 input channels.  NIRGAN_B200_WGRAD_GROUPS=0: one tap per unit except the two
// round-1 tap-inner cases.
static int taps_per_group(const ConvGeom& g) {
  static const bool on = [] { const char* e = getenv("NIRGAN_B200_WGRAD_GROUPS"); return !(e && e[0] == '0'); }();
  if (g.Cin == 16 && g.ntaps == 16 && g.nphase == 1) return 16;
  if (g.Cin == 64 && g.ntaps == 7 && g.nphase == 1) return 7;
  if (!on) return 1;
  if (g.Cin == 64 && g.ntaps > 1) return g.ntaps % 4 == 0 && g.nphase == 1 ? 4 : 3;
  if (g.Cin == 128 && g.ntaps > 1) return 2;
  return 1;
}

static bool wgrad_tc_supported(const ng_conv_args& a) {
  if (a.dtype != NG_F16 && a.dtype != NG_BF16) return false;
  if (a.Cin != 16 && a.Cin != 64 && a.Cin != 128 && a.Cin != 256) return false;
  if (a.Cout % 64 != 0) return false;
  return true;
}

static void wgrad_plan(const ng_conv_args& a, const ConvGeom& g, WgPlan& w) {
  // patch of exactly 64 virtual pixels (the TMA box may overhang the image: zero fill)
  long long best = -1;
  w.BH = 8; w.BW = 8;
  for (int bw = 64; bw >= 1; bw >>= 1) {
    const int bh = 64 / bw;
    if (bw * g.S > 256 || bh * g.S > 256 || bw * g.OS > 256 || bh * g.OS > 256) continue;
    const long long tiles = (long long)((g.VH + bh - 1) / bh) * ((g.VW + bw - 1) / bw);
    if (best < 0 || tiles < best) { best = tiles; w.BH = bh; w.BW = bw; }
  }
  const bool rp3 = row_patch3_ok(g);
  if (row_patch_ok(g) || rp3) { w.BH = RP_BH; w.BW = RP_BW; }
  w.patches_y = (g.VH + w.BH - 1) / w.BH;
  w.patches_x = (g.VW + w.BW - 1) / w.BW;
  w.P = g.B * w.patches_y * w.patches_x;
  w.n_tiles = (g.Cout + 127) / 128;
  w.bn = g.Cin;
  w.nt = taps_per_group(g);
  w.tap_stride = 1; w.k_tiles = 1;
  // groups of up to nt consecutive taps that share an output phase (the dY box position depends on the phase)
  w.ngroups = 0;
  if (rp3) {
    // group kw = taps kw, kw + 3, kw + 6 (the three row shifts of one haloed patch); input channels in tiles of 128
    w.nt = 3; w.tap_stride = 3; w.bn = 128; w.k_tiles = g.Cin / 128;
    for (int kw = 0; kw < 3; ++kw) { w.grp_first[kw] = (unsigned char)kw; w.grp_cnt[kw] = 3; }
    w.ngroups = 3;
  } else {
    for (int ph = 0; ph < g.nphase; ++ph)
      for (int t = g.phase_tap0[ph]; t < g.phase_tap0[ph + 1]; t += w.nt) {
        const int left = g.phase_tap0[ph + 1] - t;
        w.grp_first[w.ngroups] = (unsigned char)t;
        w.grp_cnt[w.ngroups] = (unsigned char)(left < w.nt ? left : w.nt);
        ++w.ngroups;
      }
  }
  const int base = w.ngroups * w.n_tiles * w.k_tiles, sms = num_sms();
  int best_s = 1; double best_eff = -1.0;
  int max_s = 4 * sms / base;                   // enough splits to fill the GPU even for a single-tap, single-tile problem
  if (max_s < 64) max_s = 64;
  if (max_s > 256) max_s = 256;
  if (max_s > w.P) max_s = w.P;
  for (int s = 1; s <= max_s; ++s) {
    const long long units = (long long)base * s;
    const long long waves = (units + sms - 1) / sms;
    const double eff = (double)units / (double)(waves * sms);
    if (eff > best_eff + 1e-9) { best_eff = eff; best_s = s; }
    if (eff >= 0.9) { best_s = s; break; }
  }
  w.pps = (w.P + best_s - 1) / best_s;
  w.splits = (w.P + w.pps - 1) / w.pps;
}

long long wgrad_tc_workspace_bytes(const ng_conv_args& a, const ConvGeom& g) {
  if (!wgrad_tc_supported(a)) return 0;
  WgPlan w;
  wgrad_plan(a, g, w);
  return (long long)w.splits * g.ntaps * g.Cout * g.Cin * (long long)sizeof(float);
}

template <int BN, int NT, bool RP = false>
static int launch_wgrad_tc(const ng_conv_args& a, const ConvGeom& g, const WgPlan& w, float* dw, void* workspace,
                           cudaStream_t st) {
  using Cfg = WgCfg<BN, NT, RP>;
  WgParams p;
  memset(&p, 0, sizeof(p));
  p.g = g;
  p.BH = w.BH; p.BW = w.BW; p.patches_y = w.patches_y; p.patches_x = w.patches_x; p.P = w.P;
  p.n_tiles = w.n_tiles; p.splits = w.splits; p.pps = w.pps;
  p.ngroups = w.ngroups;
  memcpy(p.grp_first, w.grp_first, sizeof(p.grp_first));
  memcpy(p.grp_cnt, w.grp_cnt, sizeof(p.grp_cnt));
  p.tap_stride = w.tap_stride; p.k_tiles = w.k_tiles;
  p.total_units = w.splits * w.ngroups * w.n_tiles * w.k_tiles;
  p.bf16 = a.dtype == NG_BF16;
  p.partials = w.splits == 1 ? dw : reinterpret_cast<float*>(workspace);

  CUtensorMap tmY, tmX;
  const CUtensorMapDataType dt = a.dtype == NG_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  {
    cuuint64_t dims[4] = {(cuuint64_t)g.Cout, (cuuint64_t)g.Wout, (cuuint64_t)g.Hout, (cuuint64_t)g.B};
    cuuint64_t strides[3] = {(cuuint64_t)g.Cout * 2, (cuuint64_t)g.Wout * g.Cout * 2,
                             (cuuint64_t)g.Hout * g.Wout * g.Cout * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)(w.BW * g.OS), (cuuint32_t)(w.BH * g.OS), 1};
    cuuint32_t estr[4] = {1, (cuuint32_t)g.OS, (cuuint32_t)g.OS, 1};
    int cr = 0;
    const int er = cached_tensor_map(&tmY, dt, 4, a.y, dims, strides, box, estr, CU_TENSOR_MAP_SWIZZLE_128B,
                                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, &cr);
    NG_REQUIRE(er == NG_OK, NG_E_DRIVER, "wgrad_tc: cuTensorMapEncodeTiled(dY) failed: %d", cr);
  }
  {
    cuuint64_t dims[4] = {(cuuint64_t)g.Cin, (cuuint64_t)g.Wb, (cuuint64_t)g.Hb, (cuuint64_t)g.B};
    cuuint64_t strides[3] = {(cuuint64_t)g.Cin * 2, (cuuint64_t)g.Wb * g.Cin * 2, (cuuint64_t)g.Hb * g.Wb * g.Cin * 2};
    cuuint32_t box[4] = {(cuuint32_t)(BN >= 64 ? 64 : BN), (cuuint32_t)(w.BW * g.S), (cuuint32_t)(w.BH * g.S), 1};
    if (RP) box[2] = (cuuint32_t)(RP_BH + NT - 1);          // the haloed patch: every row tap in one box
    cuuint32_t estr[4] = {1, (cuuint32_t)g.S, (cuuint32_t)g.S, 1};
    int cr = 0;
    const int er = cached_tensor_map(&tmX, dt, 4, a.x, dims, strides, box, estr,
                                     BN >= 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_32B,
                                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, &cr);
    NG_REQUIRE(er == NG_OK, NG_E_DRIVER, "wgrad_tc: cuTensorMapEncodeTiled(X) failed: %d", cr);
  }
  static PerDeviceOnce once;      // per instantiation and per device; thread-safe
  const int dev = current_device();
  if (once.needed(dev)) {
    int e = check_cuda(cudaFuncSetAttribute(wgrad_tc_kernel<BN, NT, RP>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            Cfg::SMEM_BYTES), "cudaFuncSetAttribute(wgrad_tc)");
    if (e) return e;
    once.done(dev);
  }
  const int sms = num_sms();
  const int grid = p.total_units < sms ? p.total_units : sms;
  wgrad_tc_kernel<BN, NT, RP><<<grid, 192, Cfg::SMEM_BYTES, st>>>(tmY, tmX, p);
  NG_LAUNCH_CHECK("wgrad_tc_kernel");
  if (w.splits > 1) {
    const long long n4 = (long long)g.ntaps * g.Cout * g.Cin / 4;
    long long blocks = (n4 + 255) / 256;
    if (blocks > (long long)sms * 8) blocks = (long long)sms * 8;
    if (w.splits >= 16 && n4 < (long long)sms * 256 * 2) {
      wgrad_reduce_sliced_kernel<<<(unsigned)((n4 + 31) / 32), 256, 0, st>>>(reinterpret_cast<const float4*>(workspace),
                                                                           w.splits, n4, reinterpret_cast<float4*>(dw));
      NG_LAUNCH_CHECK("wgrad_reduce_sliced_kernel");
      return NG_OK;
    }
    wgrad_reduce_kernel<<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<const float4*>(workspace), w.splits, n4,
                                                         reinterpret_cast<float4*>(dw));
    NG_LAUNCH_CHECK("wgrad_reduce_kernel");
  }
  return NG_OK;
}

// returns NG_E_UNSUPPORTED (without setting an error) when the geometry is not covered: the caller falls back to SIMT
int wgrad_tc(const ng_conv_args& a, const ConvGeom& g, float* dw, void* workspace, long long workspace_bytes,
             cudaStream_t st, bool* handled) {
  *handled = false;
  if (!wgrad_tc_supported(a) || g.ntaps != a.KH * a.KW) return NG_OK;
  WgPlan w;
  wgrad_plan(a, g, w);
  const long long need = w.splits > 1 ? (long long)w.splits * g.ntaps * g.Cout * g.Cin * (long long)sizeof(float) : 0;
  if (need > 0 && (workspace == nullptr || workspace_bytes < need)) return NG_OK;     // no workspace: SIMT path
  NG_REQUIRE(((uintptr_t)a.x & 15) == 0 && ((uintptr_t)a.y & 15) == 0 && ((uintptr_t)dw & 15) == 0 &&
                 ((uintptr_t)workspace & 15) == 0,
             NG_E_ALIGN, "wgrad_tc: tensors must be 16-byte aligned");
  *handled = true;
  if (w.tap_stride == 3) return launch_wgrad_tc<128, 3, true>(a, g, w, dw, workspace, st);
  if (w.nt == 16 && g.Cin == 16) return launch_wgrad_tc<16, 16>(a, g, w, dw, workspace, st);
  if (w.nt == 7 && g.Cin == 64 && row_patch_ok(g) && w.BW == RP_BW && w.BH == RP_BH)
    return launch_wgrad_tc<64, 7, true>(a, g, w, dw, workspace, st);
  if (w.nt == 7 && g.Cin == 64) return launch_wgrad_tc<64, 7>(a, g, w, dw, workspace, st);
  if (w.nt == 4 && g.Cin == 64) return launch_wgrad_tc<64, 4>(a, g, w, dw, workspace, st);
  if (w.nt == 3 && g.Cin == 64) return launch_wgrad_tc<64, 3>(a, g, w, dw, workspace, st);
  if (w.nt == 2 && g.Cin == 128) return launch_wgrad_tc<128, 2>(a, g, w, dw, workspace, st);
  NG_REQUIRE(w.nt == 1, NG_E_UNSUPPORTED, "wgrad_tc: no kernel for %d taps per group at Cin %d", w.nt, g.Cin);
  switch (g.Cin) {
    case 16:  return launch_wgrad_tc<16, 1>(a, g, w, dw, workspace, st);
    case 64:  return launch_wgrad_tc<64, 1>(a, g, w, dw, workspace, st);
    case 128: return launch_wgrad_tc<128, 1>(a, g, w, dw, workspace, st);
    case 256: return launch_wgrad_tc<256, 1>(a, g, w, dw, workspace, st);
  }
  *handled = false;
  return NG_OK;
}

}  // namespace ng
