// Generator head Conv2d(64 -> 1, k7) + Tanh (model/networks.py:366-368) as ONE tcgen05 kernel.
//
// The two-launch form (1x1 "tap GEMM" z[pixel][tap] = <x[pixel,:], w[tap,:]> into HBM, then ng_tap_gather summing the 49
// shifted taps) moves the activation once and z twice: ~840 MB per 32 tiles of 256 px.  Here z never leaves the SM:
//   * one TMA box per tile fetches the haloed 14 x 22 pixel patch (64 channels, K-major, 128-byte swizzle) of an 8 x 16
//     output patch -- every input pixel is read once from HBM (2.4x from L2, which has the bandwidth);
//   * the tap GEMM of the WHOLE patch (308 pixels -> three M = 128 row chunks, N = 64 taps, K = 64) runs as twelve
//     tcgen05.mma into TMEM (3 x 64 columns, double buffered); the [tap][channel] weights stay resident in shared memory;
//   * an epilogue group of four warps drains the accumulators into a shared z tile [patch pixel][tap] (16 bit, 132-byte
//     rows: conflict-free for both the row-per-lane writes and the pixel-per-lane reads), then each thread owns one
//     output pixel: out = tanh(bias + sum_{kh,kw} z[(oy + kh, ox + kw)][kh * 7 + kw]) in a fixed order (deterministic),
//     written as fp32 with the wrapper's crop applied.  Two groups take alternate tiles.
// HBM traffic: the activation once + the fp32 output (~290 MB per 32 tiles).
#include "tc_common.cuh"
#include <stdlib.h>

namespace ng {

// One template serves both single-output-channel layers of the reference:
//   <NCH = 1, KWIN = 7, NTAP = 64>  generator head    Conv2d(64 -> 1, k7) + Tanh   (model/networks.py:366-368)
//   <NCH = 8, KWIN = 4, NTAP = 16>  PatchGAN last     Conv2d(512 -> 1, k4, p1)      (model/networks.py:574-576)
// NCH = 64-channel K chunks (one pipeline stage each, accumulated in TMEM), KWIN = window, NTAP = stored taps = GEMM-N.
// The im2col form of the PatchGAN layer re-read its 512-channel input once per tap (16x: 900 MB from L2 per 64 images,
// 0.18 ms for 0.9 GFLOP); here every input pixel is fetched 1.6x (patch halo) and the 16 taps come out of one GEMM.
constexpr int HD_BH = 8, HD_BW = 16;
constexpr int HD_C = 64;                                                 // channels per K chunk
constexpr int HD_GROUPS = 2;
constexpr int HD_THREADS = 64 + 128 * HD_GROUPS;

template <int NCH, int KWIN, int NTAP>
struct HeadCfg {
  static constexpr int PH = HD_BH + KWIN - 1, PW = HD_BW + KWIN - 1;     // haloed patch (14 x 22 / 11 x 19)
  static constexpr int ROWS = PH * PW;                                   // patch pixels (308 / 209)
  static constexpr int CHUNKS = (ROWS + 127) / 128;                      // M = 128 row chunks (3 / 2)
  // A stage holds the patch rows of one K chunk (whole 1024-byte swizzle atoms).  The last row chunk's MMA reads up to
  // 128 * CHUNKS - ROWS rows past them -- into the next stage / the weights, readable shared memory whose products land in
  // TMEM lanes nobody drains -- which is what lets one more stage fit (with two, the head was bound by the TMA round trip).
  static constexpr int A_BYTES = (ROWS * HD_C * 2 + 1023) / 1024 * 1024;
  static constexpr int BOX_BYTES = ROWS * HD_C * 2;
  static constexpr int W_CHUNK_BYTES = NTAP * HD_C * 2;                  // [NTAP rows][128 B] per K chunk
  static constexpr int W_BYTES = NCH * W_CHUNK_BYTES;
  static constexpr int TAPS = KWIN * KWIN;                               // real taps (49 / 16)
  // z tile: 16-bit for the 49-tap head (the rounding of the stored z of the two-kernel form), fp32 for the 16-tap layer
  // (it is small, and the PatchGAN logit keeps the fp32 accumulation of the im2col form it replaces)
  static constexpr bool Z32 = NTAP == 16;
  static constexpr int ZWORDS = Z32 ? TAPS : (TAPS + 1) / 2;             // 32-bit words stored per z row (25 / 16)
  static constexpr int ZPITCH = (Z32 ? NTAP + 1 : NTAP / 2 + 1) * 4;     // odd word count: conflict-free rows (132 / 68)
  static constexpr int Z_BYTES = (ROWS * ZPITCH + 127) / 128 * 128;
  static constexpr int FIXED = W_BYTES + HD_GROUPS * Z_BYTES + 256 + 1024;
  static constexpr int STAGES_RAW = (232448 - 1024 - FIXED) / A_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int ACC_COLS = CHUNKS * NTAP;                         // one accumulator set (192 / 32)
  static constexpr int TMEM_COLS = 2 * ACC_COLS <= 32 ? 32 : (2 * ACC_COLS <= 64 ? 64 : (2 * ACC_COLS <= 128 ? 128 :
                                   (2 * ACC_COLS <= 256 ? 256 : 512)));
  static constexpr int SMEM = STAGES * A_BYTES + FIXED;
  static_assert(STAGES >= 2 && 2 * ACC_COLS <= 512 && (NTAP == 16 || NTAP == 64), "unsupported head shape");
  // the over-read of the last stage must stay inside the allocation
  static_assert((STAGES - 1) * A_BYTES + CHUNKS * 128 * HD_C * 2 <= SMEM - 1024, "over-read leaves the allocation");
};

struct HeadParams {
  int B, Hc, Wc, crop;            // cropped output geometry
  int org;                        // buffer coordinate of the patch origin of output (0, 0): crop - (pad - halo), may be < 0
  int tiles_y, tiles_x, total;
  int act, bf16;
  const float* bias;
  float* out;
};

__device__ __forceinline__ uint32_t lds16(uint32_t a) {
  uint16_t v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }

template <int NCH, int KWIN, int NTAP>
__global__ void __launch_bounds__(HD_THREADS, 1)
head_fused_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ HeadParams p) {
  using Cfg = HeadCfg<NCH, KWIN, NTAP>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* wres = smem + STAGES * Cfg::A_BYTES;
  uint8_t* zt = wres + Cfg::W_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(zt + HD_GROUPS * Cfg::Z_BYTES);
  uint64_t* full_bar = bars;                   // [STAGES]
  uint64_t* empty_bar = bars + STAGES;         // [STAGES]
  uint64_t* tfull_bar = bars + 2 * STAGES;     // [2]
  uint64_t* tempty_bar = tfull_bar + 2;        // [2]
  uint64_t* wfull_bar = tempty_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wfull_bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(smem_u32(&full_bar[s]), 1); mbar_init(smem_u32(&empty_bar[s]), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(smem_u32(&tfull_bar[s]), 1); mbar_init(smem_u32(&tempty_bar[s]), 128); }
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
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int per_image = p.tiles_y * p.tiles_x;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_expect_tx(smem_u32(wfull_bar), (uint32_t)Cfg::W_BYTES);
      for (int kc = 0; kc < NCH; ++kc)
        tma_load_2d(&tmB, smem_u32(wfull_bar), smem_u32(wres) + kc * Cfg::W_CHUNK_BYTES, kc * HD_C, 0);
      uint32_t stage = 0, phase = 0;
      for (int q = blockIdx.x; q < p.total; q += gridDim.x) {
        const int n = q / per_image, t = q - n * per_image;
        const int ty = t / p.tiles_x, tx = t - ty * p.tiles_x;
        for (int kc = 0; kc < NCH; ++kc) {
          mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
          const uint32_t fb = smem_u32(&full_bar[stage]);
          mbar_expect_tx(fb, (uint32_t)Cfg::BOX_BYTES);
          // buffer coordinates of the patch origin (negative / beyond the buffer = the zero padding: TMA zero fill)
          tma_load_4d(&tmA, fb, smem_u32(smem + stage * Cfg::A_BYTES), kc * HD_C, tx * HD_BW + p.org, ty * HD_BH + p.org, n);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | ((uint32_t)(p.bf16 ? 1 : 0) << 7) | ((uint32_t)(p.bf16 ? 1 : 0) << 10) |
                             ((uint32_t)(NTAP >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      uint32_t stage = 0, phase = 0, as = 0, as_phase = 0;
      mbar_wait(smem_u32(wfull_bar), 0);
      for (int q = blockIdx.x; q < p.total; q += gridDim.x) {
        mbar_wait(smem_u32(&tempty_bar[as]), as_phase ^ 1);
        for (int kc = 0; kc < NCH; ++kc) {
          mbar_wait(smem_u32(&full_bar[stage]), phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * Cfg::A_BYTES);
          const uint64_t bdesc = make_kmajor_desc(smem_u32(wres) + kc * Cfg::W_CHUNK_BYTES, 1024, 2);
#pragma unroll
          for (int ch = 0; ch < Cfg::CHUNKS; ++ch) {
            const uint64_t adesc = make_kmajor_desc(sa + ch * (128 * HD_C * 2), 1024, 2);
#pragma unroll
            for (int k = 0; k < HD_C / 16; ++k)
              umma_f16(tmem_base + as * Cfg::ACC_COLS + ch * NTAP, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k),
                       idesc, (uint32_t)((kc | k) != 0));
          }
          umma_commit(smem_u32(&empty_bar[stage]));
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(smem_u32(&tfull_bar[as]));
        if (++as == 2) { as = 0; as_phase ^= 1; }
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue groups: z tile, then the tap gather =====================
    const int grp = (warp - 2) >> 2, qtr = warp & 3;            // TMEM lane quarter = warp id mod 4
    const int et = ((warp - 2) & 3) * 32 + lane;                // 0..127 within the group: the output pixel it owns
    const uint32_t zg = smem_u32(zt) + grp * Cfg::Z_BYTES;
    const uint32_t barid = 1 + grp;
    const int oy = et / HD_BW, ox = et - oy * HD_BW;
    const float bias = p.bias ? p.bias[0] : 0.f;
    uint32_t as = 0, as_phase = 0;
    int it = 0;
    for (int q = blockIdx.x; q < p.total; q += gridDim.x, ++it) {
      if ((it & 1) != grp) {                                    // the other group drains this tile (other accumulator)
        if (++as == 2) { as = 0; as_phase ^= 1; }
        continue;
      }
      const int n = q / per_image, t = q - n * per_image;
      const int ty = t / p.tiles_x, tx = t - ty * p.tiles_x;
      mbar_wait(smem_u32(&tfull_bar[as]), as_phase);
      tc_fence_after();
#pragma unroll
      for (int ch = 0; ch < Cfg::CHUNKS; ++ch) {
        const int r = ch * 128 + qtr * 32 + lane;               // patch pixel of this thread's TMEM lane
        if (ch * 128 + qtr * 32 >= Cfg::ROWS) continue;         // warp-uniform: no patch pixel in this lane quarter
        const uint32_t taddr = tmem_base + as * Cfg::ACC_COLS + ch * NTAP + ((uint32_t)(qtr * 32) << 16);
        const uint32_t zr = zg + r * Cfg::ZPITCH;
        if constexpr (NTAP == 64) {
          uint32_t v0[32], v1[32];
          tmem_ld32(taddr, v0);
          tmem_ld32(taddr + 32, v1);
          tmem_ld_wait();
          if (r < Cfg::ROWS) {
#pragma unroll
            for (int k = 0; k < 16; ++k) {
              const float a = __uint_as_float(v0[2 * k]), b = __uint_as_float(v0[2 * k + 1]);
              sts32(zr + 4 * k, p.bf16 ? pack2<__nv_bfloat16>(a, b) : pack2<__half>(a, b));
            }
#pragma unroll
            for (int k = 0; k < Cfg::ZWORDS - 16; ++k) {          // taps 32..49 (48 is the last real one)
              const float a = __uint_as_float(v1[2 * k]), b = __uint_as_float(v1[2 * k + 1]);
              sts32(zr + 64 + 4 * k, p.bf16 ? pack2<__nv_bfloat16>(a, b) : pack2<__half>(a, b));
            }
          }
        } else {
          uint32_t v0[16];
          tmem_ld16(taddr, v0);
          tmem_ld_wait();
          if (r < Cfg::ROWS) {
#pragma unroll
            for (int k = 0; k < Cfg::ZWORDS; ++k) sts32(zr + 4 * k, v0[k]);      // fp32 z
          }
        }
      }
      tc_fence_before();
      mbar_arrive(smem_u32(&tempty_bar[as]));                   // this thread's last read of the accumulator
      bar_sync_id(barid);
      // ---- gather: this thread's output pixel (oy, ox) of the 8 x 16 patch ----
      float acc = 0.f;
      const uint32_t z0 = zg + (oy * Cfg::PW + ox) * Cfg::ZPITCH;
#pragma unroll
      for (int kh = 0; kh < KWIN; ++kh) {
#pragma unroll
        for (int kw = 0; kw < KWIN; ++kw) {
          if constexpr (Cfg::Z32) {
            acc += __uint_as_float(lds32(z0 + (kh * Cfg::PW + kw) * Cfg::ZPITCH + (kh * KWIN + kw) * 4));
          } else {
            const uint32_t h = lds16(z0 + (kh * Cfg::PW + kw) * Cfg::ZPITCH + (kh * KWIN + kw) * 2);
            acc += p.bf16 ? __uint_as_float(h << 16) : __half2float(__ushort_as_half((unsigned short)h));
          }
        }
      }
      const int y = ty * HD_BH + oy, x = tx * HD_BW + ox;
      if (y < p.Hc && x < p.Wc) p.out[((size_t)n * p.Hc + y) * p.Wc + x] = apply_act(acc + bias, p.act, 0.f);
      bar_sync_id(barid);                                       // z tile free for this group's next tile
      if (++as == 2) { as = 0; as_phase ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cfg::TMEM_COLS)
                 : "memory");
  }
}

template <int NCH, int KWIN, int NTAP>
static int launch_head(const void* x_haloed, int dtype, int B, int H, int W, int pad, int halo, const void* w_taps,
                       const float* bias, int act, int crop, float* out, cudaStream_t st) {
  using Cfg = HeadCfg<NCH, KWIN, NTAP>;
  HeadParams p;
  p.B = B; p.Hc = H - 2 * crop; p.Wc = W - 2 * crop; p.crop = crop; p.org = crop - (pad - halo);
  p.tiles_y = (p.Hc + HD_BH - 1) / HD_BH; p.tiles_x = (p.Wc + HD_BW - 1) / HD_BW;
  const long long total = (long long)B * p.tiles_y * p.tiles_x;
  NG_REQUIRE(total < (1ll << 31), NG_E_SHAPE, "head_conv: too many tiles");
  p.total = (int)total; p.act = act; p.bf16 = dtype == NG_BF16 ? 1 : 0; p.bias = bias; p.out = out;
  const CUtensorMapDataType dt = dtype == NG_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  const int C = NCH * HD_C, Hz = H + KWIN - 1 - 2 * (pad - halo), Wz = W + KWIN - 1 - 2 * (pad - halo);
  CUtensorMap tmA, tmB;
  {
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)Wz, (cuuint64_t)Hz, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)Wz * C * 2, (cuuint64_t)Hz * Wz * C * 2};
    cuuint32_t box[4] = {(cuuint32_t)HD_C, (cuuint32_t)Cfg::PW, (cuuint32_t)Cfg::PH, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    int cr = 0;
    const int er = cached_tensor_map(&tmA, dt, 4, x_haloed, dims, strides, box, estr, CU_TENSOR_MAP_SWIZZLE_128B,
                                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, &cr);
    NG_REQUIRE(er == NG_OK, NG_E_DRIVER, "head_conv: cuTensorMapEncodeTiled(x) failed: %d", cr);
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)NTAP};
    cuuint64_t strides[1] = {(cuuint64_t)C * 2};
    cuuint32_t box[2] = {(cuuint32_t)HD_C, (cuuint32_t)NTAP};
    cuuint32_t estr[2] = {1, 1};
    int cr = 0;
    const int er = cached_tensor_map(&tmB, dt, 2, w_taps, dims, strides, box, estr, CU_TENSOR_MAP_SWIZZLE_128B,
                                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, &cr);
    NG_REQUIRE(er == NG_OK, NG_E_DRIVER, "head_conv: cuTensorMapEncodeTiled(w) failed: %d", cr);
  }
  auto kern = head_fused_kernel<NCH, KWIN, NTAP>;
  static PerDeviceOnce once;      // per instantiation and per device
  const int dev = current_device();
  if (once.needed(dev)) {
    int e = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM),
                       "cudaFuncSetAttribute(head_fused)");
    if (e) return e;
    once.done(dev);
  }
  const int sms = num_sms();
  const int grid = p.total < sms ? p.total : sms;
  kern<<<grid, HD_THREADS, Cfg::SMEM, st>>>(tmA, tmB, p);
  NG_LAUNCH_CHECK("head_fused_kernel");
  return NG_OK;
}

}  // namespace ng

using namespace ng;

extern "C" int ng_head_conv(const void* x_haloed, int32_t dtype, int32_t B, int32_t H, int32_t W, int32_t C, int32_t K,
                            int32_t pad, int32_t halo, const void* w_taps, const float* bias, int32_t act, int32_t crop,
                            float* out, void* stream) {
  int r = require_sm100(); if (r) return r;
  NG_REQUIRE(x_haloed && w_taps && out, NG_E_ARG, "head_conv: null tensor");
  NG_REQUIRE(dtype == NG_F16 || dtype == NG_BF16, NG_E_UNSUPPORTED, "head_conv: 16-bit storage only (the fp32 verification "
             "mode runs ng_conv2d)");
  NG_REQUIRE((C == 64 && K == 7) || (C == 512 && K == 4), NG_E_UNSUPPORTED,
             "head_conv: built for 64 -> 1 k7 (generator head) and 512 -> 1 k4 (PatchGAN last layer), got %d -> 1 k%d", C, K);
  NG_REQUIRE(pad >= 0 && halo >= 0 && halo <= pad && 2 * pad <= K - 1, NG_E_ARG, "head_conv: pad %d / halo %d for k%d", pad, halo, K);
  NG_REQUIRE(B > 0 && H > 0 && W > 0 && crop >= 0 && H - 2 * crop > 0 && W - 2 * crop > 0, NG_E_SHAPE, "head_conv: empty output");
  NG_REQUIRE(((uintptr_t)x_haloed & 15) == 0 && ((uintptr_t)w_taps & 127) == 0, NG_E_ALIGN, "head_conv: unaligned tensor");
  NG_REQUIRE(act == NG_ACT_TANH || act == NG_ACT_NONE, NG_E_UNSUPPORTED, "head_conv: activation %d", act);
  cudaStream_t st = (cudaStream_t)stream;
  if (K == 7) return launch_head<1, 7, 64>(x_haloed, dtype, B, H, W, pad, halo, w_taps, bias, act, crop, out, st);
  return launch_head<8, 4, 16>(x_haloed, dtype, B, H, W, pad, halo, w_taps, bias, act, crop, out, st);
}
