// Backward of the fused "InstanceNorm -> inject -> activation (+ residual) -> halo" unit, plus the small
// gradient plumbing kernels of the training step (tanh' / channel padding, NHWC->NCHW gradient export,
// SatCLIP injection gradients).  All memory-bound; 16-byte vector accesses over 8 channels.
//
// Forward unit (ng_in_apply):   xh = (y - mean) * rstd ; u = inject(xh) ; o = act(u) [+ residual] ; buffer = halo(o)
// Backward, given g = dL/d(buffer) (haloed) and/or g_skip = dL/do from a skip connection:
//   do  = fold_halo(g) + g_skip
//   du  = do * act'(u)
//   dxh = du * d inject / d xh
//   dy  = rstd * (dxh - mean_hw(dxh) - xh * mean_hw(dxh * xh))          (InstanceNorm backward)
// Pass 1 (ng_in_bwd_reduce) accumulates the two per-(n,c) means (+ injection gradients), pass 2
// (ng_in_bwd_apply) recomputes dxh and writes dy (and, for residual blocks, do for the skip path).
#include "tc_common.cuh"
#include <stdlib.h>

namespace ng {

template <typename T>
__device__ __forceinline__ void ld8(const T* p, float (&f)[8]) {
  if constexpr (sizeof(T) == 2) {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    float2 a = unpack2<T>(u.x), b = unpack2<T>(u.y), c = unpack2<T>(u.z), d = unpack2<T>(u.w);
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
  } else {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  }
}
template <typename T>
__device__ __forceinline__ void st8(T* p, const float (&f)[8]) {
  if constexpr (sizeof(T) == 2) {
    uint4 u;
    u.x = pack2<T>(f[0], f[1]); u.y = pack2<T>(f[2], f[3]); u.z = pack2<T>(f[4], f[5]); u.w = pack2<T>(f[6], f[7]);
    *reinterpret_cast<uint4*>(p) = u;
  } else {
    *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(f[4], f[5], f[6], f[7]);
  }
}

__device__ __forceinline__ float bilerp128_b(const float* __restrict__ e, int oy, int ox, int H, int W) {
  const float sy = fmaxf((oy + 0.5f) * (128.f / H) - 0.5f, 0.f), sx = fmaxf((ox + 0.5f) * (128.f / W) - 0.5f, 0.f);
  const int y0 = min((int)sy, 127), x0 = min((int)sx, 127);
  const int y1 = min(y0 + 1, 127), x1 = min(x0 + 1, 127);
  const float ly = sy - y0, lx = sx - x0;
  return (1.f - ly) * ((1.f - lx) * e[y0 * 128 + x0] + lx * e[y0 * 128 + x1]) +
         ly * ((1.f - lx) * e[y1 * 128 + x0] + lx * e[y1 * 128 + x1]);
}

struct BwdArgs {
  int B, H, W, C, c8_shift;
  int act; float slope;
  int gp, halo_mode;           // halo of the incoming gradient buffer
  int inj_mode;
  float inv_hw;
  int ppb;                     // interior pixels per block
  int nblk1;                   // pass-1 blocks per image = partial-sum slots per image
  int n0;                      // first image of this launch (image-chunked launches: grid.y = images in the chunk)
  int pf_ahead;                // lean kernels: loop iterations of operands kept prefetched into L2 ahead of the loads (0 = off)
  int reverse2;                // lean pass 2 walks images / pixel blocks in the reverse of pass 1's order (L2 reuse)
  unsigned long long w_magic;  // ceil(2^40 / W): p / W == (p * w_magic) >> 40 for p < 2^20
};

__device__ __forceinline__ uint4 ldg16_stream(const void* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
// DRAM -> L2 bulk prefetch (no destination in the SM): the later ld.global of the same bytes hits the L2
__device__ __forceinline__ void l2_prefetch(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
template <typename T>
__device__ __forceinline__ void up8(const uint4& u, float (&f)[8]) {
  float2 a = unpack2<T>(u.x), b = unpack2<T>(u.y), c = unpack2<T>(u.z), d = unpack2<T>(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
// raw 8-channel vector: 16 B on the 16-bit paths, 32 B in fp32 verification mode
template <typename T> struct Raw8 { uint4 lo; };
template <> struct Raw8<float> { uint4 lo, hi; };
template <typename T>
__device__ __forceinline__ Raw8<T> ld_raw8(const T* p) {
  Raw8<T> r;
  r.lo = ldg16_stream(p);
  if constexpr (sizeof(T) == 4) r.hi = ldg16_stream(p + 4);
  return r;
}
template <typename T>
__device__ __forceinline__ void raw_to_f(const Raw8<T>& r, float (&f)[8]) {
  if constexpr (sizeof(T) == 2) up8<T>(r.lo, f);
  else {
    f[0] = __uint_as_float(r.lo.x); f[1] = __uint_as_float(r.lo.y); f[2] = __uint_as_float(r.lo.z); f[3] = __uint_as_float(r.lo.w);
    f[4] = __uint_as_float(r.hi.x); f[5] = __uint_as_float(r.hi.y); f[6] = __uint_as_float(r.hi.z); f[7] = __uint_as_float(r.hi.w);
  }
}

// The reflect halo folds mirrored gradient rows / columns back onto interior pixels next to the border: adds the
// (rare) extra contributions to d[] for interior pixel (y, x).  ptr = gradient buffer of image n at channel group c8.
template <typename T>
__device__ __forceinline__ void fold_halo_extras(const BwdArgs& a, const T* __restrict__ gn, int y, int x, float (&d)[8]) {
  const int p = a.gp, Wb = a.W + 2 * p;
  int ys[3], xs[3];
  ys[0] = y + p; xs[0] = x + p;
  ys[1] = (y >= 1 && y <= p) ? p - y : -1;
  ys[2] = (y >= a.H - 1 - p && y <= a.H - 2) ? p + 2 * (a.H - 1) - y : -1;
  xs[1] = (x >= 1 && x <= p) ? p - x : -1;
  xs[2] = (x >= a.W - 1 - p && x <= a.W - 2) ? p + 2 * (a.W - 1) - x : -1;
  if ((ys[1] & ys[2] & xs[1] & xs[2]) == -1) return;       // all four absent
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    if (ys[i] < 0) continue;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      if (xs[j] < 0 || (i == 0 && j == 0)) continue;
      float f[8];
      ld8<T>(gn + ((size_t)ys[i] * Wb + xs[j]) * a.C, f);
#pragma unroll
      for (int k = 0; k < 8; ++k) d[k] += f[k];
    }
  }
}

// Shared by all passes.  PHASE 1: per-(n,c) sums of dxh and dxh*xh (+ injection gradients).  PHASE 2: dy = rstd * (dxh -
// mean(dxh) - xh * mean(dxh*xh)) and the optional export of do for the skip path.  Block (chunk, n) walks `ppb`
// consecutive interior pixels of image n; thread t owns the channel group c8 = t mod C/8 (statistics in registers) and
// batches UNROLL pixels of independent 16-byte streaming loads (gradient, skip gradient, forward pre-norm tensor).
template <typename T, int PASS, int UNROLL, bool HAS_G, bool HAS_SKIP, bool HAS_INJ>
__device__ __forceinline__ void in_bwd_loop(const BwdArgs& a, const int n, const int c8, const int pstep, const int npix,
                                            const int p_begin, const int p_end, const float s, const bool mr,
                                            const float (&mean)[8], const float (&rstd)[8], const float (&m1)[8],
                                            const float (&m2)[8], const T* __restrict__ gn, const T* __restrict__ sn,
                                            const T* __restrict__ yn, const float* __restrict__ injn, const bool fold,
                                            float (&acc1)[8], float (&acc2)[8], float& ds_acc,
                                            float* __restrict__ de_map, T* __restrict__ dy, T* __restrict__ do_out) {
  const int gp = a.gp, Wb = a.W + 2 * gp;
  for (int p0 = p_begin + (threadIdx.x >> a.c8_shift); p0 < p_end; p0 += pstep * UNROLL) {
    Raw8<T> rg[HAS_G ? UNROLL : 1], rs[HAS_SKIP ? UNROLL : 1], ry[UNROLL];
    int py[UNROLL], px[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const int pp = p0 + u * pstep;
      const int yy = (int)(((unsigned long long)(unsigned)pp * a.w_magic) >> 40);
      py[u] = yy; px[u] = pp - yy * a.W;
      if (pp < p_end) {
        ry[u] = ld_raw8<T>(yn + (size_t)pp * a.C);
        if constexpr (HAS_G) rg[u] = ld_raw8<T>(gn + ((size_t)(yy + gp) * Wb + px[u] + gp) * a.C);
        if constexpr (HAS_SKIP) rs[u] = ld_raw8<T>(sn + (size_t)pp * a.C);
      }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const int pp = p0 + u * pstep;
      if (pp >= p_end) continue;
      float d_o[8], xh[8];
      if constexpr (HAS_G) raw_to_f<T>(rg[u], d_o);
      else {
#pragma unroll
        for (int k = 0; k < 8; ++k) d_o[k] = 0.f;
      }
      if (fold) fold_halo_extras<T>(a, gn, py[u], px[u], d_o);
      if constexpr (HAS_SKIP) {
        float t[8];
        raw_to_f<T>(rs[u], t);
#pragma unroll
        for (int k = 0; k < 8; ++k) d_o[k] += t[k];
      }
      raw_to_f<T>(ry[u], xh);
      if (mr) {
#pragma unroll
        for (int k = 0; k < 8; ++k) xh[k] = (xh[k] - mean[k]) * rstd[k];
      }
      float ev = 0.f, fac = 1.f, add = 0.f;
      if constexpr (HAS_INJ) {
        ev = bilerp128_b(injn, py[u], px[u], a.H, a.W);
        if (a.inj_mode == NG_INJECT_ADD) add = s * ev;
        else if (a.inj_mode == NG_INJECT_MUL_SCALED) fac = 1.f + s * ev;
        else fac = ev;
      }
      float du[8], dxh[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float uu = xh[k] * fac + add;
        float dact = 1.f;
        if (a.act == NG_ACT_RELU) dact = uu > 0.f ? 1.f : 0.f;
        else if (a.act == NG_ACT_LRELU) dact = uu > 0.f ? 1.f : a.slope;
        du[k] = d_o[k] * dact;
        dxh[k] = du[k] * fac;
      }
      if constexpr (PASS == 1) {
#pragma unroll
        for (int k = 0; k < 8; ++k) { acc1[k] += dxh[k]; acc2[k] = fmaf(dxh[k], xh[k], acc2[k]); }
        if constexpr (HAS_INJ) {
          float t = 0.f;     // sum over these 8 channels of du * d u / d(s*e);  u = xh*(1+s*e) | xh + s*e | xh*e
#pragma unroll
          for (int k = 0; k < 8; ++k) t += (a.inj_mode == NG_INJECT_ADD) ? du[k] : du[k] * xh[k];
          if (a.inj_mode == NG_INJECT_MUL) { if (de_map) atomicAdd(&de_map[(size_t)n * npix + pp], t); }
          else {
            ds_acc += t * ev;
            if (de_map) atomicAdd(&de_map[(size_t)n * npix + pp], t * s);
          }
        }
      } else {
        float o[8];
        if (mr) {
#pragma unroll
          for (int k = 0; k < 8; ++k) o[k] = rstd[k] * (dxh[k] - m1[k] - xh[k] * m2[k]);
        } else {
#pragma unroll
          for (int k = 0; k < 8; ++k) o[k] = dxh[k];
        }
        st8<T>(dy + ((size_t)n * npix + pp) * a.C + c8 * 8, o);
        if (do_out) st8<T>(do_out + ((size_t)n * npix + pp) * a.C + c8 * 8, d_o);
      }
    }
  }
}

// PASS 1 / PASS 2: the two-launch form (pass 2 alone serves the units without a norm).  PASS 3: both phases in one launch
// for normalised units -- after phase 1 the block publishes its partial sums and takes a ticket on the image's counter;
// the last block of the image combines the partials (fixed order) and raises the image's flag, the others wait for it,
// then every block runs phase 2 over the SAME pixels, which it has just read: those reads are served by the L2 instead
// of HBM.  Blocks are dispatched image-major, so a waiting block only ever waits for blocks dispatched before or right
// after it; the launcher keeps the blocks of one image (nblk1) well below the number of co-resident blocks.
template <typename T, int PASS, int UNROLL, bool HAS_G, bool HAS_SKIP, bool HAS_INJ>
__global__ void __launch_bounds__(256, UNROLL >= 4 ? 2 : 3)
in_bwd_kernel(const __grid_constant__ BwdArgs a, const T* __restrict__ g, const T* __restrict__ gskip,
              const T* __restrict__ yv, const float* __restrict__ mr, const float* __restrict__ inj,
              const float* __restrict__ inj_scale, float* __restrict__ sums, float* __restrict__ dscale,
              float* __restrict__ de_map, T* __restrict__ dy, T* __restrict__ do_out) {
  extern __shared__ float sm[];            // PASS 1 / 3: [4 accumulators][256 threads]
  const int n = blockIdx.y + a.n0, C8 = a.C >> 3;
  const int c8 = threadIdx.x & (C8 - 1);
  const int pstep = 256 >> a.c8_shift;
  const int npix = a.H * a.W;
  const int p_begin = blockIdx.x * a.ppb, p_end = min(npix, p_begin + a.ppb);
  const float s = (HAS_INJ && inj_scale) ? *inj_scale : 1.f;
  float mean[8], rstd[8], m1[8], m2[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { mean[k] = 0.f; rstd[k] = 1.f; m1[k] = 0.f; m2[k] = 0.f; }
  if (mr) {
    const float4* m4 = reinterpret_cast<const float4*>(mr + ((size_t)n * a.C + c8 * 8) * 2);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float4 m = m4[k];
      mean[2 * k] = m.x; rstd[2 * k] = m.y; mean[2 * k + 1] = m.z; rstd[2 * k + 1] = m.w;
    }
    if (PASS == 2) {
      const float4* s4 = reinterpret_cast<const float4*>(sums + ((size_t)n * a.C + c8 * 8) * 2);   // combined sums
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float4 sv = s4[k];
        m1[2 * k] = sv.x * a.inv_hw; m2[2 * k] = sv.y * a.inv_hw;
        m1[2 * k + 1] = sv.z * a.inv_hw; m2[2 * k + 1] = sv.w * a.inv_hw;
      }
    }
  }
  const int gp = a.gp, Wb = a.W + 2 * gp;
  const T* gn = HAS_G ? g + (size_t)n * (a.H + 2 * gp) * Wb * a.C + c8 * 8 : nullptr;
  const T* sn = HAS_SKIP ? gskip + (size_t)n * npix * a.C + c8 * 8 : nullptr;
  const T* yn = yv + (size_t)n * npix * a.C + c8 * 8;
  const float* injn = HAS_INJ ? inj + (size_t)n * 128 * 128 : nullptr;
  const bool fold = HAS_G && a.halo_mode == NG_HALO_REFLECT && gp > 0;
  float acc1[8], acc2[8], ds_acc = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) acc1[k] = acc2[k] = 0.f;

  in_bwd_loop<T, (PASS == 2 ? 2 : 1), UNROLL, HAS_G, HAS_SKIP, HAS_INJ>(a, n, c8, pstep, npix, p_begin, p_end, s,
                                                                        mr != nullptr, mean, rstd, m1, m2, gn, sn, yn,
                                                                        injn, fold, acc1, acc2, ds_acc, de_map, dy, do_out);
  if constexpr (PASS != 2) {
    // block reduction without atomics, four accumulators at a time (4 KB of shared memory, so that these blocks fit
    // beside a resident tcgen05 CTA of the weight-gradient stream): every thread parks accumulators 4r..4r+3, then each
    // output (channel, stat) is the sum over the 256/C8 threads that own that channel group, in a fixed order.
    if (HAS_INJ && dscale) {
      ds_acc = warp_sum(ds_acc);
      if ((threadIdx.x & 31) == 0 && ds_acc != 0.f) atomicAdd(dscale, ds_acc);
    }
    // partial slots live behind the [B][C][2] combined sums
    float* dst = sums ? sums + (size_t)a.B * a.C * 2 + ((size_t)n * a.nblk1 + blockIdx.x) * a.C * 2 : nullptr;
    const int lanes = 256 >> a.c8_shift;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      sm[0 * 256 + threadIdx.x] = acc1[2 * r];
      sm[1 * 256 + threadIdx.x] = acc2[2 * r];
      sm[2 * 256 + threadIdx.x] = acc1[2 * r + 1];
      sm[3 * 256 + threadIdx.x] = acc2[2 * r + 1];
      __syncthreads();
      if (dst && (int)threadIdx.x < (a.C >> 1)) {
        const int cg = threadIdx.x >> 2, j = threadIdx.x & 3;
        float t = 0.f;
        for (int L = 0; L < lanes; ++L) t += sm[j * 256 + (L << a.c8_shift) + cg];
        dst[cg * 16 + 4 * r + j] = t;
      }
      __syncthreads();
    }
  }
  if constexpr (PASS == 3) {
    __shared__ int last_flag;
    int* counter = reinterpret_cast<int*>(sums + (size_t)a.B * a.C * 2 * (1 + a.nblk1));      // [B] tickets, [B] flags
    int* flag = counter + a.B;
    float* combined = sums + (size_t)n * a.C * 2;
    __threadfence();                        // the block's partial slot is visible device-wide
    __syncthreads();
    if (threadIdx.x == 0) last_flag = (atomicAdd(&counter[n], 1) == a.nblk1 - 1);
    __syncthreads();
    if (last_flag) {
      __threadfence();
      const float* part = sums + (size_t)a.B * a.C * 2 + (size_t)n * a.nblk1 * a.C * 2;
      const int C2 = a.C * 2;
      for (int o = threadIdx.x; o < C2; o += 256) {
        float t = 0.f;
        int k = 0;
        for (; k + 8 <= a.nblk1; k += 8) {          // eight independent L2 loads in flight, summed in slot order
          float v[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = __ldcg(part + (size_t)(k + j) * C2 + o);
#pragma unroll
          for (int j = 0; j < 8; ++j) t += v[j];
        }
        for (; k < a.nblk1; ++k) t += __ldcg(part + (size_t)k * C2 + o);
        combined[o] = t;
      }
      __threadfence();
      __syncthreads();
      if (threadIdx.x == 0) atomicExch(&flag[n], 1);
    } else {
      if (threadIdx.x == 0) {
        while (*reinterpret_cast<volatile int*>(&flag[n]) == 0) __nanosleep(100);
        __threadfence();
      }
      __syncthreads();
    }
    {
      const float4* s4 = reinterpret_cast<const float4*>(combined + c8 * 16);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float4 sv = __ldcg(s4 + k);
        m1[2 * k] = sv.x * a.inv_hw; m2[2 * k] = sv.y * a.inv_hw;
        m1[2 * k + 1] = sv.z * a.inv_hw; m2[2 * k + 1] = sv.w * a.inv_hw;
      }
    }
    in_bwd_loop<T, 2, UNROLL, HAS_G, HAS_SKIP, HAS_INJ>(a, n, c8, pstep, npix, p_begin, p_end, s, true, mean, rstd, m1, m2,
                                                        gn, sn, yn, injn, fold, acc1, acc2, ds_acc, de_map, dy, do_out);
  }
}

// Lean form of the two norm-backward passes for the common case -- 16-bit storage, normalised unit, no injection, the
// activation a compile-time constant -- written for instruction count (ncu r2h: the generic kernel executes ~85 warp
// instructions per 16-byte load, a third of them address arithmetic, at IPC 1.6 and 34-36 % of the HBM peak): xh is one
// FFMA per element, the mask a select, dy two FFMAs with per-channel constants, the (row, column) of a pixel is advanced
// incrementally, offsets are 32-bit.  Same block -> pixel mapping, partial-sum slots and combine kernel as the generic
// form, so either can serve either pass.
template <typename T, int PASS, int UNROLL, bool HAS_G, bool HAS_SKIP, int ACT>
__global__ void __launch_bounds__(256, 2)
in_bwd_fast_kernel(const __grid_constant__ BwdArgs a, const T* __restrict__ g, const T* __restrict__ gskip,
                   const T* __restrict__ yv, const float* __restrict__ mr, float* __restrict__ sums, T* __restrict__ dy,
                   T* __restrict__ do_out) {
  static_assert(sizeof(T) == 2, "16-bit storage only");
  extern __shared__ float sm[];            // PASS 1: [4 accumulators][256 threads]
  // pass 2 in the reverse of pass 1's dispatch order: what pass 1 read LAST (the final images of the batch) is what the L2
  // still holds when pass 2 starts, so its first blocks re-read from the L2 instead of HBM
  const bool rev = PASS == 2 && a.reverse2;
  const int bx = rev ? (int)(gridDim.x - 1 - blockIdx.x) : (int)blockIdx.x;
  const int n = (rev ? (int)(gridDim.y - 1 - blockIdx.y) : (int)blockIdx.y) + a.n0, C8 = a.C >> 3, C = a.C;
  const int c8 = threadIdx.x & (C8 - 1);
  const int pstep = 256 >> a.c8_shift;
  const int npix = a.H * a.W;
  const int p_begin = bx * a.ppb, p_end = min(npix, p_begin + a.ppb);
  float sa[8], sb[8], k1[8], k2[8];        // xh = y * sa + sb ; dy = dxh * sa + xh * k2 + k1
  {
    const float4* m4 = reinterpret_cast<const float4*>(mr + ((size_t)n * C + c8 * 8) * 2);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float4 m = m4[k];
      sa[2 * k] = m.y; sb[2 * k] = -m.x * m.y; sa[2 * k + 1] = m.w; sb[2 * k + 1] = -m.z * m.w;
    }
    if (PASS == 2) {
      const float4* s4 = reinterpret_cast<const float4*>(sums + ((size_t)n * C + c8 * 8) * 2);   // combined sums
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float4 sv = s4[k];
        k1[2 * k] = -sa[2 * k] * (sv.x * a.inv_hw); k2[2 * k] = -sa[2 * k] * (sv.y * a.inv_hw);
        k1[2 * k + 1] = -sa[2 * k + 1] * (sv.z * a.inv_hw); k2[2 * k + 1] = -sa[2 * k + 1] * (sv.w * a.inv_hw);
      }
    }
  }
  const int gp = a.gp, Wb = a.W + 2 * gp;
  const T* gn = HAS_G ? g + (size_t)n * (a.H + 2 * gp) * Wb * C + c8 * 8 : nullptr;
  const T* sn = HAS_SKIP ? gskip + (size_t)n * npix * C + c8 * 8 : nullptr;
  const T* yn = yv + (size_t)n * npix * C + c8 * 8;
  T* dyn = PASS == 2 ? dy + (size_t)n * npix * C + c8 * 8 : nullptr;
  T* don = (PASS == 2 && do_out) ? do_out + (size_t)n * npix * C + c8 * 8 : nullptr;
  const bool fold = HAS_G && a.halo_mode == NG_HALO_REFLECT && gp > 0;
  float acc1[8], acc2[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc1[k] = acc2[k] = 0.f;
  int pp = p_begin + (threadIdx.x >> a.c8_shift);
  int py = pp / a.W, px = pp - py * a.W;
  const int adv_y = pstep / a.W, adv_x = pstep - adv_y * a.W;
  // The loads of an iteration cover pstep * UNROLL consecutive pixels of every operand.  Optionally (pf_ahead > 0) one
  // thread keeps the next iterations' bytes on their way from DRAM into the L2 (bulk prefetch: no registers, no shared
  // memory).  Measured: no gain here -- these kernels are bound by DRAM efficiency over ~600 concurrent streams, not by
  // the latency of the individual loads (see the launcher).
  const int chunk_px = pstep * UNROLL;
  auto prefetch_chunk = [&](int p0) {
    if (p0 >= p_end) return;
    const int p1 = min(p0 + chunk_px, p_end) - 1;
    const uint32_t bytes = (uint32_t)(p1 - p0 + 1) * C * 2;
    l2_prefetch(yv + ((size_t)n * npix + p0) * C, bytes);
    if constexpr (HAS_SKIP) l2_prefetch(gskip + ((size_t)n * npix + p0) * C, bytes);
    if constexpr (HAS_G) {
      const int y0 = p0 / a.W, y1 = p1 / a.W;
      const size_t a0 = (size_t)(y0 + gp) * Wb + (p0 - y0 * a.W) + gp, a1 = (size_t)(y1 + gp) * Wb + (p1 - y1 * a.W) + gp;
      l2_prefetch(g + ((size_t)n * (a.H + 2 * gp) * Wb + a0) * C, (uint32_t)(a1 - a0 + 1) * C * 2);
    }
  };
  int pf_next = p_begin;
  if (a.pf_ahead > 0 && threadIdx.x == 0) {
    for (int i = 0; i < a.pf_ahead; ++i, pf_next += chunk_px) prefetch_chunk(pf_next);
  }
  while (pp < p_end) {
    if (a.pf_ahead > 0 && threadIdx.x == 0) { prefetch_chunk(pf_next); pf_next += chunk_px; }
    Raw8<T> rg[HAS_G ? UNROLL : 1], rs[HAS_SKIP ? UNROLL : 1], ry[UNROLL];
    int ypos[UNROLL], xpos[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      ypos[u] = py; xpos[u] = pp < p_end ? px : -1;
      if (pp < p_end) {
        ry[u] = ld_raw8<T>(yn + pp * C);
        if constexpr (HAS_G) rg[u] = ld_raw8<T>(gn + ((py + gp) * Wb + px + gp) * C);
        if constexpr (HAS_SKIP) rs[u] = ld_raw8<T>(sn + pp * C);
      }
      pp += pstep; py += adv_y; px += adv_x;
      if (px >= a.W) { px -= a.W; ++py; }
    }
    int pq = pp - UNROLL * pstep;
#pragma unroll
    for (int u = 0; u < UNROLL; ++u, pq += pstep) {
      if (xpos[u] < 0) continue;
      float d_o[8], xh[8];
      if constexpr (HAS_G) raw_to_f<T>(rg[u], d_o);
      else {
#pragma unroll
        for (int k = 0; k < 8; ++k) d_o[k] = 0.f;
      }
      // mirrored halo rows / columns fold back onto interior pixels within gp of the border only
      if (fold && ((unsigned)(ypos[u] - 1) >= (unsigned)(a.H - 2) || (unsigned)(xpos[u] - 1) >= (unsigned)(a.W - 2) ||
                   ypos[u] <= gp || xpos[u] <= gp || ypos[u] >= a.H - 1 - gp || xpos[u] >= a.W - 1 - gp))
        fold_halo_extras<T>(a, gn, ypos[u], xpos[u], d_o);
      if constexpr (HAS_SKIP) {
        float t[8];
        raw_to_f<T>(rs[u], t);
#pragma unroll
        for (int k = 0; k < 8; ++k) d_o[k] += t[k];
      }
      raw_to_f<T>(ry[u], xh);
      float dxh[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        xh[k] = fmaf(xh[k], sa[k], sb[k]);
        if (ACT == NG_ACT_RELU) dxh[k] = xh[k] > 0.f ? d_o[k] : 0.f;
        else if (ACT == NG_ACT_LRELU) dxh[k] = xh[k] > 0.f ? d_o[k] : d_o[k] * a.slope;
        else dxh[k] = d_o[k];
      }
      if constexpr (PASS == 1) {
#pragma unroll
        for (int k = 0; k < 8; ++k) { acc1[k] += dxh[k]; acc2[k] = fmaf(dxh[k], xh[k], acc2[k]); }
      } else {
        float o[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] = fmaf(dxh[k], sa[k], fmaf(xh[k], k2[k], k1[k]));
        st8<T>(dyn + pq * C, o);
        if (don) st8<T>(don + pq * C, d_o);
      }
    }
  }
  if constexpr (PASS == 1) {
    // identical block reduction and slot layout as the generic kernel
    float* dst = sums + (size_t)a.B * C * 2 + ((size_t)n * a.nblk1 + blockIdx.x) * C * 2;
    const int lanes = 256 >> a.c8_shift;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      sm[0 * 256 + threadIdx.x] = acc1[2 * r];
      sm[1 * 256 + threadIdx.x] = acc2[2 * r];
      sm[2 * 256 + threadIdx.x] = acc1[2 * r + 1];
      sm[3 * 256 + threadIdx.x] = acc2[2 * r + 1];
      __syncthreads();
      if ((int)threadIdx.x < (C >> 1)) {
        const int cg = threadIdx.x >> 2, j = threadIdx.x & 3;
        float t = 0.f;
        for (int L2 = 0; L2 < lanes; ++L2) t += sm[j * 256 + (L2 << a.c8_shift) + cg];
        dst[cg * 16 + 4 * r + j] = t;
      }
      __syncthreads();
    }
  }
}

// ---- TMA-staged form of the two norm-backward passes ----------------------------------------------------
// The register-staged kernels above are bound by the bytes they keep in flight (ncu r2h: 34-36 % of the HBM peak with
// the load batches of a warp draining while it computes).  Here one producer thread streams the operands with bulk
// asynchronous copies (cp.async.bulk, completion on an mbarrier) into a ring of shared-memory stages -- ~190 KB in
// flight per SM regardless of what the 16 consumer warps are doing -- and the consumers read 16-byte vectors from
// shared memory.  One persistent CTA per SM owns a contiguous range of `stage items` (a row segment, or a few whole
// rows, of one image) of the whole batch; pass 1 leaves one partial-sum slot per (CTA, image it touched), pass 2 sums
// the slots of its image in CTA order (deterministic) while its first stages are already in flight: no combine launch.
// A reflect halo of g folds back onto the interior pixels within gp of the border.  The staged kernels read plain row
// segments, so this pre-pass adds the halo contributions into those interior pixels IN PLACE, once (the register-staged
// kernels redo the fold in both passes).  Reads touch halo positions only, writes interior positions only: no ordering
// between threads is needed.  One thread = one border pixel x 8 channels.
template <typename T>
__global__ void __launch_bounds__(256)
fold_halo_kernel(const __grid_constant__ BwdArgs a, T* __restrict__ g) {
  const int p = a.gp, H = a.H, W = a.W, C8 = a.C >> 3, Wb = W + 2 * p, Hb = H + 2 * p;
  const int nrow_px = 2 * p * W, per_image = nrow_px + (H - 2 * p) * 2 * p;
  const long long total = (long long)a.B * per_image * C8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % C8);
    const long long t = i / C8;
    const int n = (int)(t / per_image);
    int j = (int)(t - (long long)n * per_image), y, x;
    if (j < nrow_px) {                      // rows 1..p and H-1-p..H-2, every column
      const int ri = j / W;
      x = j - ri * W;
      y = ri < p ? 1 + ri : H - 1 - p + (ri - p);
    } else {                                // the other rows (0, p+1..H-2-p, H-1): columns 1..p and W-1-p..W-2
      j -= nrow_px;
      const int yi = j / (2 * p), ci = j - yi * 2 * p;
      y = yi == 0 ? 0 : (yi == H - 2 * p - 1 ? H - 1 : p + yi);
      x = ci < p ? 1 + ci : W - 1 - p + (ci - p);
    }
    T* gn = g + (size_t)n * Hb * Wb * a.C + c8 * 8;
    float d[8], own[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) d[k] = 0.f;
    fold_halo_extras<T>(a, gn, y, x, d);
    T* o = gn + ((size_t)(y + p) * Wb + x + p) * a.C;
    ld8<T>(o, own);
#pragma unroll
    for (int k = 0; k < 8; ++k) own[k] += d[k];
    st8<T>(o, own);
  }
}

struct BwdStream {
  int segw, rows, split, ipi;     // stage item = rows x segw pixels of one image; items per image
  int total, G, kmax;             // items in the batch, CTAs, partial slots per CTA
  int nstages, tensor_bytes;      // shared-memory ring: nstages x tensors x tensor_bytes
};
constexpr int kBwdConsumers = 512;

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void consumers_sync() { asm volatile("bar.sync 1, 512;" ::: "memory"); }

template <typename T, int PASS, bool HAS_G, bool HAS_SKIP, int ACT>
__global__ void __launch_bounds__(kBwdConsumers + 32, 1)
in_bwd_stream_kernel(const __grid_constant__ BwdArgs a, const __grid_constant__ BwdStream q, const T* __restrict__ g,
                     const T* __restrict__ gskip, const T* __restrict__ yv, const float* __restrict__ mr,
                     float* __restrict__ sums, T* __restrict__ dy, T* __restrict__ do_out) {
  static_assert(sizeof(T) == 2, "16-bit storage only");
  constexpr int NTEN = (HAS_G ? 1 : 0) + (HAS_SKIP ? 1 : 0) + 1;
  extern __shared__ __align__(128) uint8_t ring_raw[];
  __shared__ __align__(8) uint64_t bars[16];          // full[8], empty[8]
  const uint32_t ring = smem_u32(ring_raw);
  float* red = reinterpret_cast<float*>(ring_raw + (size_t)q.nstages * NTEN * q.tensor_bytes);    // [4][512]
  const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[8]);
  const int tid = threadIdx.x, C = a.C, npix = a.H * a.W;
  if (tid == 0) {
    for (int s = 0; s < q.nstages; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, kBwdConsumers / 32); }
    fence_barrier_init();
  }
  __syncthreads();
  const long long it0 = (long long)blockIdx.x * q.total / q.G, it1 = (long long)(blockIdx.x + 1) * q.total / q.G;
  const int gp = a.gp, Wb = a.W + 2 * gp, Hb = a.H + 2 * gp;
  const uint32_t seg_bytes = (uint32_t)q.segw * C * 2;

  if (tid >= kBwdConsumers) {
    // ---- producer: one thread keeps the ring full ----
    if (tid != kBwdConsumers) return;
    int k = 0;
    for (long long it = it0; it < it1; ++it, ++k) {
      const int s = k % q.nstages;
      mbar_wait(empty0 + 8 * s, ((k / q.nstages) & 1) ^ 1);
      const int n = (int)(it / q.ipi), r = (int)(it - (long long)n * q.ipi);
      const int yb = r / q.split, xs = r - yb * q.split;
      const int y0 = yb * q.rows, x0 = xs * q.segw, nr = min(q.rows, a.H - y0);
      const uint32_t bytes = seg_bytes * nr;
      const uint32_t bar = full0 + 8 * s;
      mbar_expect_tx(bar, bytes * NTEN);
      uint32_t dst = ring + (uint32_t)s * NTEN * q.tensor_bytes;
      if constexpr (HAS_G) {
        for (int rr = 0; rr < nr; ++rr)
          bulk_g2s(dst + rr * seg_bytes, g + (((size_t)n * Hb + y0 + rr + gp) * Wb + x0 + gp) * C, seg_bytes, bar);
        dst += q.tensor_bytes;
      }
      // the un-haloed tensors are contiguous over the rows of a stage (split == 1 whenever rows > 1)
      bulk_g2s(dst, yv + ((size_t)n * npix + (size_t)y0 * a.W + x0) * C, bytes, bar);
      if constexpr (HAS_SKIP) bulk_g2s(dst + q.tensor_bytes, gskip + ((size_t)n * npix + (size_t)y0 * a.W + x0) * C, bytes, bar);
    }
    return;
  }

  // ---- consumers ----
  const int C8 = C >> 3, c8 = tid & (C8 - 1), lane_px = tid >> a.c8_shift, lanes = kBwdConsumers >> a.c8_shift;
  float sa[8], sb[8], k1[8], k2[8], acc1[8], acc2[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { acc1[k] = acc2[k] = 0.f; k1[k] = k2[k] = 0.f; sa[k] = 1.f; sb[k] = 0.f; }
  float* part = sums + (size_t)a.B * C * 2;
  const int n_first = (int)(it0 / q.ipi);
  int cur_n = -1;

  auto flush = [&](int n) {          // PASS 1: block reduction in a fixed order into the slot of (this CTA, image n)
    float* dst = part + ((size_t)blockIdx.x * q.kmax + (n - n_first)) * C * 2;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      red[0 * kBwdConsumers + tid] = acc1[2 * r];
      red[1 * kBwdConsumers + tid] = acc2[2 * r];
      red[2 * kBwdConsumers + tid] = acc1[2 * r + 1];
      red[3 * kBwdConsumers + tid] = acc2[2 * r + 1];
      consumers_sync();
      if (tid < (C >> 1)) {
        const int cg = tid >> 2, j = tid & 3;
        float t = 0.f;
        for (int L = 0; L < lanes; ++L) t += red[j * kBwdConsumers + (L << a.c8_shift) + cg];
        dst[cg * 16 + 4 * r + j] = t;
      }
      consumers_sync();
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) acc1[k] = acc2[k] = 0.f;
  };

  int k = 0;
  for (long long it = it0; it < it1; ++it, ++k) {
    const int s = k % q.nstages;
    const int n = (int)(it / q.ipi), r = (int)(it - (long long)n * q.ipi);
    const int yb = r / q.split, xs = r - yb * q.split;
    const int y0 = yb * q.rows, x0 = xs * q.segw, nr = min(q.rows, a.H - y0);
    if (n != cur_n) {
      if (PASS == 1 && cur_n >= 0) flush(cur_n);
      cur_n = n;
      const float4* m4 = reinterpret_cast<const float4*>(mr + ((size_t)n * C + c8 * 8) * 2);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 m = m4[j];
        sa[2 * j] = m.y; sb[2 * j] = -m.x * m.y; sa[2 * j + 1] = m.w; sb[2 * j + 1] = -m.z * m.w;
      }
      if constexpr (PASS == 2) {
        // the image's sums: slots of the CTAs whose ranges meet the image, added in CTA order
        const long long i_lo = (long long)n * q.ipi, i_hi = i_lo + q.ipi - 1;
        const int c_lo = (int)(((i_lo + 1) * q.G - 1) / q.total), c_hi = (int)(((i_hi + 1) * q.G - 1) / q.total);
        float s1[8], s2[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) s1[j] = s2[j] = 0.f;
        for (int cc = c_lo; cc <= c_hi; ++cc) {
          const int nf = (int)(((long long)cc * q.total / q.G) / q.ipi);
          const float4* p4 = reinterpret_cast<const float4*>(part + (((size_t)cc * q.kmax + (n - nf)) * C + c8 * 8) * 2);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 v = __ldcg(p4 + j);
            s1[2 * j] += v.x; s2[2 * j] += v.y; s1[2 * j + 1] += v.z; s2[2 * j + 1] += v.w;
          }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) { k1[j] = -sa[j] * (s1[j] * a.inv_hw); k2[j] = -sa[j] * (s2[j] * a.inv_hw); }
      }
    }
    const int npx = nr * q.segw;
    const size_t pix0 = (size_t)n * npix + (size_t)y0 * a.W + x0;       // stage pixels are consecutive interior pixels
    T* dyn = PASS == 2 ? dy + pix0 * C + c8 * 8 : nullptr;
    T* don = (PASS == 2 && do_out) ? do_out + pix0 * C + c8 * 8 : nullptr;
    const uint32_t sbase = ring + (uint32_t)s * NTEN * q.tensor_bytes + c8 * 16;
    const uint32_t off_y = HAS_G ? q.tensor_bytes : 0, off_s = off_y + q.tensor_bytes;
    mbar_wait(full0 + 8 * s, (k / q.nstages) & 1);
#pragma unroll 2
    for (int qx = lane_px; qx < npx; qx += lanes) {
      const uint32_t so = sbase + (uint32_t)qx * C * 2;
      float d_o[8], xh[8];
      if constexpr (HAS_G) up8<T>(lds128(so), d_o);
      else {
#pragma unroll
        for (int j = 0; j < 8; ++j) d_o[j] = 0.f;
      }
      if constexpr (HAS_SKIP) {
        float t[8];
        up8<T>(lds128(so + off_s), t);
#pragma unroll
        for (int j = 0; j < 8; ++j) d_o[j] += t[j];
      }
      up8<T>(lds128(so + off_y), xh);
      float dxh[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        xh[j] = fmaf(xh[j], sa[j], sb[j]);
        if (ACT == NG_ACT_RELU) dxh[j] = xh[j] > 0.f ? d_o[j] : 0.f;
        else if (ACT == NG_ACT_LRELU) dxh[j] = xh[j] > 0.f ? d_o[j] : d_o[j] * a.slope;
        else dxh[j] = d_o[j];
      }
      if constexpr (PASS == 1) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { acc1[j] += dxh[j]; acc2[j] = fmaf(dxh[j], xh[j], acc2[j]); }
      } else {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = fmaf(dxh[j], sa[j], fmaf(xh[j], k2[j], k1[j]));
        st8<T>(dyn + (size_t)qx * C, o);
        if (don) st8<T>(don + (size_t)qx * C, d_o);
      }
    }
    __syncwarp();
    if ((tid & 31) == 0) mbar_arrive(empty0 + 8 * s);
  }
  if (PASS == 1 && cur_n >= 0) flush(cur_n);
}

// Geometry of the staged form; false when the shape does not fit it (the register-staged kernels serve those).
static bool in_bwd_stream_geometry(int B, int H, int W, int C, int ntensors, BwdStream* out, bool sizing_only = false) {
  // opt-in (NIRGAN_B200_BWD_STREAM=1, read per call so that tests can switch it): in isolation the staged kernels reach
  // 52-77 % of the HBM copy peak (profiles/r2q_in_bwd_stream.md; the lean ones 47-79 % on the same units, r3g), but their
  // ~200 KB shared-memory ring cannot sit beside a tcgen05 CTA of the weight-gradient stream, and the training step
  // loses more from that lost overlap than the kernels gain (20.6 vs 19.4 ms per step, r2q).
  const char* env = getenv("NIRGAN_B200_BWD_STREAM");
  const bool on = sizing_only || (env && env[0] == '1');
  if (!on || C > 512 || C < 8) return false;
  const long long rowbytes = (long long)W * C * 2;
  BwdStream q;
  q.rows = 1; q.split = 0;
  for (int budget = 16384; budget <= 32768 && !q.split; budget *= 2) {
    if (rowbytes <= budget) {
      q.split = 1;
      q.rows = (int)(16384 / rowbytes);
      if (q.rows < 1) q.rows = 1;
      if (q.rows > H) q.rows = H;
      if (q.rows > 16) q.rows = 16;
    } else {
      for (int sp = 2; sp <= W; ++sp)        // whole divisors of the row only, and never segments below half the budget
        if (W % sp == 0 && rowbytes / sp <= budget) {
          if (rowbytes / sp >= budget / 2) q.split = sp;
          break;
        }
    }
  }
  if (!q.split) return false;
  q.segw = W / q.split;
  q.ipi = ((H + q.rows - 1) / q.rows) * q.split;
  const long long total = (long long)B * q.ipi;
  if (total >= (1ll << 31) / 1024) return false;
  q.total = (int)total;
  q.G = num_sms() < q.total ? num_sms() : q.total;
  const int per = (q.total + q.G - 1) / q.G;
  q.kmax = (per + q.ipi - 1) / q.ipi + 1;
  q.tensor_bytes = (q.rows * q.segw * C * 2 + 127) / 128 * 128;
  const int avail = 232448 - 4 * kBwdConsumers * (int)sizeof(float) - 1024;
  q.nstages = avail / (ntensors * q.tensor_bytes);
  if (q.nstages > 8) q.nstages = 8;
  if (q.nstages < 2) return false;
  *out = q;
  return true;
}

template <typename T, int PASS>
static int launch_in_bwd_stream(const BwdArgs& a, const BwdStream& q, cudaStream_t st, const void* g, const void* gskip,
                                const void* y, const float* mr, float* sums, void* dy, void* do_out) {
  const bool hg = g != nullptr, hs = gskip != nullptr;
  const int nten = (hg ? 1 : 0) + (hs ? 1 : 0) + 1;
  const size_t smem = (size_t)q.nstages * nten * q.tensor_bytes + 4 * kBwdConsumers * sizeof(float);
#define NG_BWDS(HG, HS, ACT)                                                                                            \
  do {                                                                                                                  \
    auto kern = in_bwd_stream_kernel<T, PASS, HG, HS, ACT>;                                                             \
    int e = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - 1024),          \
                       "in_bwd_stream smem attribute");                                                                 \
    if (e) return e;                                                                                                    \
    kern<<<q.G, kBwdConsumers + 32, smem, st>>>(a, q, (const T*)g, (const T*)gskip, (const T*)y, mr, sums, (T*)dy,      \
                                                (T*)do_out);                                                            \
  } while (0)
#define NG_BWDS_ACT(ACT)                                                                                                \
  do {                                                                                                                  \
    if (hg && hs) NG_BWDS(true, true, ACT); else if (hg) NG_BWDS(true, false, ACT); else NG_BWDS(false, true, ACT);    \
  } while (0)
  if (a.act == NG_ACT_RELU) NG_BWDS_ACT(NG_ACT_RELU);
  else if (a.act == NG_ACT_LRELU) NG_BWDS_ACT(NG_ACT_LRELU);
  else NG_BWDS_ACT(NG_ACT_NONE);
#undef NG_BWDS_ACT
#undef NG_BWDS
  return NG_OK;
}

// combined[n][o] = sum over the image's pass-1 blocks of partial[n][blk][o], in block order (deterministic, no atomics)
__global__ void __launch_bounds__(256)
in_bwd_combine_kernel(const float* __restrict__ partial, int first, int total, int per_image, int nblk,
                      float* __restrict__ out) {
  const int i = first + blockIdx.x * blockDim.x + threadIdx.x;      // [first, total): the images of this chunk
  if (i >= total) return;
  const int n = i / per_image, o = i - n * per_image;
  const float* p = partial + (size_t)n * nblk * per_image + o;
  float t = 0.f;
  int k = 0;
  for (; k + 8 <= nblk; k += 8) {              // eight independent loads in flight, added in block order
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = p[(size_t)(k + j) * per_image];
#pragma unroll
    for (int j = 0; j < 8; ++j) t += v[j];
  }
  for (; k < nblk; ++k) t += p[(size_t)k * per_image];
  out[i] = t;
}

// ---- gradient plumbing ------------------------------------------------------------------------------
// dst[n][y][x][0] = scale * dout[n][y-crop][x-crop] * act'(out)  (zero outside the crop, channels 1.. zero)
template <typename T>
__global__ void __launch_bounds__(256)
head_bwd_prep_kernel(const float* __restrict__ dout, const float* __restrict__ out, int B, int H, int W, int crop,
                     int act, float scale, const float* __restrict__ dev_scale, int cpad, T* __restrict__ dst) {
  const long long total = (long long)B * H * W;
  if (dev_scale) scale *= dev_scale[0];
  const int Hc = H - 2 * crop, Wc = W - 2 * crop;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % W), y = (int)((i / W) % H), n = (int)(i / ((long long)W * H));
    float v = 0.f;
    const int yc = y - crop, xc = x - crop;
    if (yc >= 0 && yc < Hc && xc >= 0 && xc < Wc) {
      const size_t o = ((size_t)n * Hc + yc) * Wc + xc;
      v = dout[o] * scale;
      if (act == NG_ACT_TANH) { const float t = out[o]; v *= (1.f - t * t); }
      // 16-bit gradient storage: saturate instead of overflowing to inf (the NDVI/NDWI/EVI terms are singular where
      // pred + band crosses zero; an inf here would turn the whole backward pass into NaN)
      if constexpr (sizeof(T) == 2) v = fminf(fmaxf(v, -3.0e4f), 3.0e4f);
    }
    T* d = dst + i * cpad;
    d[0] = from_f32<T>(v);
    for (int c = 1; c < cpad; ++c) d[c] = from_f32<T>(0.f);
  }
}

// Backward of ng_tap_gather: dz[n][yy][xx][t] = scale * dout[n][yy-kh-crop][xx-kw-crop] * act'(out) for t = kh*KW + kw
// (zero outside the cropped output window and for t >= KH*KW).  A block owns a TH x TW tile of z pixels: it first builds
// the (TH + KH - 1) x (TW + KW - 1) window of gm = scale * dout * act'(out) in shared memory (each output pixel's product
// computed once instead of once per tap), then one thread per z pixel writes that pixel's ZC taps -- 128 contiguous bytes
// on the 16-bit paths, the tap -> (kh, kw) mapping resolved at compile time.  (The first form, a thread per pixel and 8
// taps with the index arithmetic per tap, ran at 0.31 ms for the 326 MB it writes; bound by instruction issue.)
template <typename T, int KH, int KW, int ZC>
__global__ void __launch_bounds__(256)
tap_scatter_kernel(const float* __restrict__ dout, const float* __restrict__ out, int B, int Hz, int Wz, int act, int crop,
                   float scale, const float* __restrict__ dev_scale, T* __restrict__ dz) {
  constexpr int TH = 8, TW = 32, GH = TH + KH - 1, GW = TW + KW - 1;
  __shared__ float gm[GH][GW + 1];
  const int Hc = Hz - KH + 1 - 2 * crop, Wc = Wz - KW + 1 - 2 * crop;
  if (dev_scale) scale *= dev_scale[0];
  const int n = blockIdx.z, ty0 = blockIdx.y * TH, tx0 = blockIdx.x * TW;
  const float* dn = dout + (size_t)n * Hc * Wc;
  const float* on = out + (size_t)n * Hc * Wc;
  for (int i = threadIdx.x; i < GH * GW; i += 256) {
    const int ly = i / GW, lx = i - ly * GW;
    const int yc = ty0 - (KH - 1) - crop + ly, xc = tx0 - (KW - 1) - crop + lx;
    float v = 0.f;
    if ((unsigned)yc < (unsigned)Hc && (unsigned)xc < (unsigned)Wc) {
      const int o = yc * Wc + xc;
      v = __ldg(dn + o) * scale;
      if (act == NG_ACT_TANH) { const float tv = __ldg(on + o); v *= (1.f - tv * tv); }
      // 16-bit gradient storage: saturate instead of overflowing to inf
      if constexpr (sizeof(T) == 2) v = fminf(fmaxf(v, -3.0e4f), 3.0e4f);
    }
    gm[ly][lx] = v;
  }
  __syncthreads();
  const int py = threadIdx.x / TW, px = threadIdx.x - py * TW;
  const int yy = ty0 + py, xx = tx0 + px;
  if (yy >= Hz || xx >= Wz) return;
  T* d = dz + (((size_t)n * Hz + yy) * Wz + xx) * ZC;
#pragma unroll
  for (int g8 = 0; g8 < ZC / 8; ++g8) {
    float f[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      constexpr int dummy = 0; (void)dummy;
      const int t = g8 * 8 + k;                       // compile-time after unrolling
      f[k] = t < KH * KW ? gm[py + (KH - 1) - t / KW][px + (KW - 1) - t % KW] : 0.f;
    }
    st8<T>(d + g8 * 8, f);
  }
}

// NHWC (cpad channels, T, scaled) -> NCHW fp32 (c channels), dst = src * scale
template <typename T>
__global__ void __launch_bounds__(256)
grad_to_nchw_kernel(const T* __restrict__ src, int B, int H, int W, int cpad, int c0, int c, int s2d, float scale,
                    const float* __restrict__ dev_scale, float* __restrict__ dst) {
  const long long total = (long long)B * c * H * W;
  if (dev_scale) scale *= dev_scale[0];
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % W), y = (int)((i / W) % H);
    const int ch = (int)((i / ((long long)W * H)) % c), n = (int)(i / ((long long)W * H * c));
    if (s2d) {
      // source in the space-to-depth layout of the zero-padded image (ng_prep_input_s2d): [(H+2)/2][(W+2)/2][parity*cpad + ch]
      const int yp = y + 1, xp = x + 1, Ws = (W + 2) >> 1, Hs = (H + 2) >> 1;
      const size_t pix = ((size_t)n * Hs + (yp >> 1)) * Ws + (xp >> 1);
      dst[i] = to_f32<T>(src[pix * (4 * cpad) + ((yp & 1) * 2 + (xp & 1)) * cpad + c0 + ch]) * scale;
    } else {
      dst[i] = to_f32<T>(src[(((size_t)n * H + y) * W + x) * cpad + c0 + ch]) * scale;
    }
  }
}

// adjoint of F.interpolate(bilinear, align_corners=False) 128x128 -> HxW:  de128 += A^T de_map
__global__ void __launch_bounds__(256)
bilerp128_bwd_kernel(const float* __restrict__ de_map, int B, int H, int W, float scale,
                     const float* __restrict__ dev_scale, float* __restrict__ de128) {
  const long long total = (long long)B * H * W;
  if (dev_scale) scale *= dev_scale[0];
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(i % W), oy = (int)((i / W) % H), n = (int)(i / ((long long)W * H));
    const float v = de_map[i] * scale;
    const float sy = fmaxf((oy + 0.5f) * (128.f / H) - 0.5f, 0.f), sx = fmaxf((ox + 0.5f) * (128.f / W) - 0.5f, 0.f);
    const int y0 = min((int)sy, 127), x0 = min((int)sx, 127);
    const int y1 = min(y0 + 1, 127), x1 = min(x0 + 1, 127);
    const float ly = sy - y0, lx = sx - x0;
    float* e = de128 + (size_t)n * 128 * 128;
    atomicAdd(&e[y0 * 128 + x0], v * (1.f - ly) * (1.f - lx));
    atomicAdd(&e[y0 * 128 + x1], v * (1.f - ly) * lx);
    atomicAdd(&e[y1 * 128 + x0], v * ly * (1.f - lx));
    atomicAdd(&e[y1 * 128 + x1], v * ly * lx);
  }
}

// fc backward: dW[n][k] = sum_b dy[b][n] * x[b][k];  db[n] = sum_b dy[b][n]   (fp32)
__global__ void __launch_bounds__(256)
linear_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, int B, int K, int N, float* __restrict__ dw,
                  float* __restrict__ db) {
  const int n = blockIdx.x;
  for (int k = threadIdx.x; k < K; k += 256) {
    float s = 0.f;
    for (int b = 0; b < B; ++b) s = fmaf(dy[(size_t)b * N + n], x[(size_t)b * K + k], s);
    dw[(size_t)n * K + k] = s;
  }
  if (threadIdx.x == 0 && db) {
    float s = 0.f;
    for (int b = 0; b < B; ++b) s += dy[(size_t)b * N + n];
    db[n] = s;
  }
}

// ---- adaptive power-of-two gradient scale ---------------------------------------------------------------
// out[2] (as uint) <- max |g| bit pattern (non-negative floats order like unsigned integers; inf / NaN sort last)
__global__ void __launch_bounds__(256)
grad_amax_kernel(const float* __restrict__ g, long long n, unsigned* __restrict__ amax_bits) {
  unsigned m = 0u;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    m = max(m, __float_as_uint(g[i]) & 0x7fffffffu);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m != 0u) atomicMax(amax_bits, m);
}
// out[0] = f = 2^floor(log2(target / amax)) (1 when amax is 0 or not finite), out[1] = 1/f
__global__ void grad_scale_finalize_kernel(float target, float* __restrict__ out) {
  const unsigned bits = reinterpret_cast<const unsigned*>(out)[2];
  float f = 1.f;
  if (bits != 0u && bits < 0x7f800000u) {
    int e = (int)floorf(log2f(target / __uint_as_float(bits)));
    e = max(-60, min(60, e));
    f = exp2f((float)e);
  }
  out[0] = f;
  out[1] = 1.f / f;
}

static inline unsigned grid_cap(long long items) {
  long long blocks = (items + 255) / 256;
  const long long cap = (long long)num_sms() * 8;
  return (unsigned)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

}  // namespace ng

using namespace ng;

#define DISPATCH_T(dtype, CALL)                                             \
  switch (dtype) {                                                          \
    case NG_F32: { using T = float; CALL; break; }                          \
    case NG_F16: { using T = __half; CALL; break; }                         \
    case NG_BF16: { using T = __nv_bfloat16; CALL; break; }                 \
    default: ng::set_error("bad dtype %d", (int)(dtype)); return NG_E_ARG;  \
  }

#define DISPATCH_T16(dtype, CALL)                                           \
  switch (dtype) {                                                          \
    case NG_F16: { using T = __half; CALL; break; }                         \
    case NG_BF16: { using T = __nv_bfloat16; CALL; break; }                 \
    default: ng::set_error("bad dtype %d", (int)(dtype)); return NG_E_ARG;  \
  }

template <typename T, int PASS>
static void launch_in_bwd(const BwdArgs& a, dim3 grid, size_t smem, cudaStream_t st, const void* g, const void* gskip,
                          const void* y, const float* mr, const float* inj, const float* inj_scale, float* sums,
                          float* dscale, float* de_map, void* dy, void* do_out) {
  const bool hg = g != nullptr, hs = gskip != nullptr, hi = a.inj_mode != NG_INJECT_NONE;
  // lean kernels for the common case (16-bit, normalised, no injection); NIRGAN_B200_BWD_FAST=0 keeps the generic ones
  static const bool fast_on = [] { const char* e = getenv("NIRGAN_B200_BWD_FAST"); return !(e && e[0] == '0'); }();
  if constexpr (sizeof(T) == 2 && (PASS == 1 || PASS == 2)) {
    const long long big = (long long)(a.H + 2 * a.gp) * (a.W + 2 * a.gp) * a.C;
    if (fast_on && !hi && mr != nullptr && dscale == nullptr && de_map == nullptr && big < (1ll << 31) && (256 >> a.c8_shift) < a.W &&
        (a.act == NG_ACT_RELU || a.act == NG_ACT_NONE || a.act == NG_ACT_LRELU)) {
#define NG_BWDF(G, S, ACT)                                                                                          \
      in_bwd_fast_kernel<T, PASS, 4, G, S, ACT><<<grid, 256, smem, st>>>(a, (const T*)g, (const T*)gskip, (const T*)y, \
                                                                        mr, sums, (T*)dy, (T*)do_out)
#define NG_BWDF_ACT(ACT)                                                                                            \
      do {                                                                                                           \
        if (hg && hs) NG_BWDF(true, true, ACT); else if (hg) NG_BWDF(true, false, ACT); else NG_BWDF(false, true, ACT); \
      } while (0)
      if (a.act == NG_ACT_RELU) NG_BWDF_ACT(NG_ACT_RELU);
      else if (a.act == NG_ACT_LRELU) NG_BWDF_ACT(NG_ACT_LRELU);
      else NG_BWDF_ACT(NG_ACT_NONE);
#undef NG_BWDF_ACT
#undef NG_BWDF
      return;
    }
  }
  // four pixels of loads in flight per thread at 2 blocks / SM (<= 128 registers) beat two at 3 blocks / SM: the kernel is
  // bound by the bytes in flight, not by occupancy (4.5 -> 4.15 ms per training step); NIRGAN_B200_BWD_UNROLL4=0 restores
  static const bool deep = [] { const char* e = getenv("NIRGAN_B200_BWD_UNROLL4"); return !(e && e[0] == '0'); }();
#define NG_BWD(G, S, I)                                                                                              \
  do {                                                                                                               \
    if (deep && !(I))                                                                                                \
      in_bwd_kernel<T, PASS, 4, G, S, I><<<grid, 256, smem, st>>>(a, (const T*)g, (const T*)gskip, (const T*)y, mr,  \
                                                                  inj, inj_scale, sums, dscale, de_map, (T*)dy,      \
                                                                  (T*)do_out);                                       \
    else                                                                                                             \
      in_bwd_kernel<T, PASS, 2, G, S, I><<<grid, 256, smem, st>>>(a, (const T*)g, (const T*)gskip, (const T*)y, mr,  \
                                                                  inj, inj_scale, sums, dscale, de_map, (T*)dy,      \
                                                                  (T*)do_out);                                       \
  } while (0)
  if (hi) {                       // the injected unit (d1) receives its gradient from one haloed buffer
    if (hg && hs) NG_BWD(true, true, true); else if (hg) NG_BWD(true, false, true); else NG_BWD(false, true, true);
  } else {
    if (hg && hs) NG_BWD(true, true, false); else if (hg) NG_BWD(true, false, false); else NG_BWD(false, true, false);
  }
#undef NG_BWD
}

// Pixels-per-block multiplier (block = pstep * mult consecutive pixels of one image) for a grid of B * ceil(npix / ppb)
// equal blocks at `resident` blocks per wave: the candidate with the best wave efficiency waves / ceil(waves) (the last,
// partly filled wave of an HBM-bound kernel runs at a fraction of the bandwidth), longer blocks on ties.
static int in_bwd_pick_mult(int B, int npix, int pstep, int lo, int hi) {
  // NIRGAN_B200_BWD_TUNE: 0 = fixed block length, 1 (default) = whole waves only, 2 = a single, nearly full wave counts as
  // filled too (blocks 4-5x longer, the per-block prologue / reduction paid once per SM slot): measured no difference
  // (r3h: 4.17 vs 4.15 ms of norm backward per step)
  static const int tune = [] { const char* e = getenv("NIRGAN_B200_BWD_TUNE"); return e ? atoi(e) : 1; }();
  const double resident = 2.0 * num_sms();             // 2 blocks / SM of the four-pixel kernels
  int best = lo;
  double best_eff = -1.0;
  for (int mult = hi; mult >= lo; --mult) {
    const long long blocks = (long long)B * ((npix + pstep * mult - 1) / (pstep * mult));
    const double waves = blocks / resident;
    if (waves < (tune >= 2 ? 0.9 : 1.0) && mult > lo) continue;   // does not fill the GPU once: use shorter blocks
    const double eff = waves <= 1.0 ? waves : waves / ceil(waves);
    if (eff > best_eff + 0.02) { best_eff = eff; best = mult; }
  }
  if (tune <= 0) return -1;
  return best;
}

static bool in_bwd_fused() {
  static const bool on = [] { const char* e = getenv("NIRGAN_B200_BWD_FUSED"); return e && e[0] == '1'; }();   // measured slower (blocks idle at the hand-over): off
  return on;
}

// pass-1 blocks per image.  Two-launch form: long blocks (fewer partial slots to combine) as long as the grid fills
// whole waves.  One-launch form (PASS 3): ~45 KB per tensor per block so that what the co-resident blocks read in phase 1
// (3 blocks x 148 SMs x 2-3 tensors) is still in the L2 for phase 2, but never more blocks per image than half the
// co-resident capacity (the blocks of an image wait for each other).
static int in_bwd_pass1_blocks(int B, int H, int W, int C) {
  const int pstep = 256 / (C / 8);
  if (in_bwd_fused()) {
    int mult = in_bwd_pick_mult(B, H * W, pstep, 8, 16);
    if (mult < 0) mult = 11;
    const int cap = 3 * num_sms() / 2;
    const int least = (H * W + pstep * cap - 1) / (pstep * cap);
    if (mult < least) mult = least;
    return (H * W + pstep * mult - 1) / (pstep * mult);
  }
  int mult = in_bwd_pick_mult(B, H * W, pstep, 4, 64);
  if (mult < 0) {
    mult = 64;
    while (mult > 4 && (long long)B * ((H * W + pstep * mult - 1) / (pstep * mult)) < 4ll * num_sms()) mult >>= 1;
  }
  return (H * W + pstep * mult - 1) / (pstep * mult);
}

extern "C" int64_t ng_in_bwd_scratch_floats(int32_t B, int32_t H, int32_t W, int32_t C) {
  if (B <= 0 || H <= 0 || W <= 0 || C < 8 || C % 8) return NG_E_ARG;
  // combined sums + per-block partials + per-image ticket counter and flag (one-launch form); the staged form keeps one
  // partial slot per (CTA, image it touches) behind the combined sums
  int64_t need = (int64_t)B * (1 + in_bwd_pass1_blocks(B, H, W, C)) * C * 2 + 2 * (int64_t)B;
  BwdStream q;
  if (in_bwd_stream_geometry(B, H, W, C, 3, &q, true)) {
    const int64_t staged = (int64_t)B * C * 2 + (int64_t)q.G * q.kmax * C * 2;
    if (staged > need) need = staged;
  }
  return need;
}

extern "C" int ng_in_bwd(void* g_halo, int32_t g_pad, int32_t halo_mode, const void* g_skip, const void* y,
                         int32_t dtype, int32_t B, int32_t H, int32_t W, int32_t C, const float* mean_rstd, int32_t act,
                         float slope, const float* inject_e, int32_t inject_mode, const float* inject_scale,
                         float* sums_scratch, void* dy, void* do_out, float* dscale, float* de_map, void* stream) {
  int r = require_sm100(); if (r) return r;
  NG_REQUIRE((g_halo || g_skip) && y && dy, NG_E_ARG, "in_bwd: null tensor");
  NG_REQUIRE(C % 8 == 0 && ((C / 8) & (C / 8 - 1)) == 0 && C <= 512, NG_E_SHAPE, "in_bwd: C = %d unsupported", C);
  NG_REQUIRE(mean_rstd == nullptr || sums_scratch != nullptr, NG_E_ARG, "in_bwd: normalised unit needs a [B][C][2] scratch");
  NG_REQUIRE(inject_mode == NG_INJECT_NONE || inject_e, NG_E_ARG, "in_bwd: injection without embedding map");
  NG_REQUIRE(halo_mode != NG_HALO_REFLECT || (2 * g_pad + 1 < H && 2 * g_pad + 1 < W), NG_E_SHAPE, "in_bwd: halo too large");
  NG_REQUIRE((long long)H * W < (1ll << 20) && W < (1 << 12), NG_E_SHAPE, "in_bwd: image %dx%d too large", H, W);
  cudaStream_t st = (cudaStream_t)stream;
  BwdArgs a;
  a.B = B; a.H = H; a.W = W; a.C = C; a.act = act; a.slope = slope; a.gp = g_halo ? g_pad : 0; a.halo_mode = halo_mode;
  a.inj_mode = inject_mode; a.inv_hw = 1.0f / ((float)H * (float)W);
  a.c8_shift = 0;
  while ((1 << a.c8_shift) < C / 8) ++a.c8_shift;
  a.w_magic = ((1ull << 40) + (unsigned)W - 1) / (unsigned)W;
  a.n0 = 0;
  // off by default: measured no effect on these kernels (r3f: 4.11 ms of norm backward per step without, 3.92-4.18 ms
  // with 1-12 iterations ahead) while the DRAM reads grow by ~20 % (r3g); the apply kernels do gain from theirs
  static const int pf_ahead = [] { const char* e = getenv("NIRGAN_B200_BWD_PREFETCH"); return e ? atoi(e) : 0; }();
  a.pf_ahead = pf_ahead;
  static const int reverse2 = [] { const char* e = getenv("NIRGAN_B200_BWD_REVERSE"); return e ? atoi(e) : 1; }();
  a.reverse2 = reverse2;
  const int pstep = 256 / (C / 8);
  const bool need_pass1 = mean_rstd != nullptr || (inject_mode != NG_INJECT_NONE && (dscale || de_map));
  a.nblk1 = in_bwd_pass1_blocks(B, H, W, C);
  if (mean_rstd && in_bwd_fused()) {
    // one launch: phase 1, per-image hand-over, phase 2 over the same (L2-resident) pixels
    if (de_map) {
      int e = check_cuda(cudaMemsetAsync(de_map, 0, (size_t)B * H * W * sizeof(float), st), "in_bwd memset de");
      if (e) return e;
    }
    float* tickets = sums_scratch + (size_t)B * C * 2 * (1 + a.nblk1);
    int e = check_cuda(cudaMemsetAsync(tickets, 0, 2 * (size_t)B * sizeof(int), st), "in_bwd memset tickets");
    if (e) return e;
    a.ppb = (H * W + a.nblk1 - 1) / a.nblk1;
    a.ppb = (a.ppb + pstep - 1) / pstep * pstep;
    dim3 grid((unsigned)a.nblk1, (unsigned)B);
    DISPATCH_T(dtype, (launch_in_bwd<T, 3>(a, grid, (size_t)4 * 256 * sizeof(float), st, g_halo, g_skip, y, mean_rstd,
                                           inject_e, inject_scale, sums_scratch, dscale, de_map, dy, do_out)));
    NG_LAUNCH_CHECK("in_bwd_kernel<fused>");
    return NG_OK;
  }
  // Staged form (16-bit storage, normalised unit, no injection): two launches, no combine.  The partial-slot geometry
  // depends on the shape only (not on which tensors are present), so both passes and the scratch query agree.
  if (dtype != NG_F32 && mean_rstd && inject_mode == NG_INJECT_NONE && dscale == nullptr && de_map == nullptr &&
      (act == NG_ACT_RELU || act == NG_ACT_NONE || act == NG_ACT_LRELU)) {
    BwdStream q, q3;
    const int nten = (g_halo ? 1 : 0) + (g_skip ? 1 : 0) + 1;
    if (in_bwd_stream_geometry(B, H, W, C, nten, &q) && in_bwd_stream_geometry(B, H, W, C, 3, &q3)) {
      if (g_halo && halo_mode == NG_HALO_REFLECT && a.gp > 0) {
        const long long items = (long long)B * (2 * a.gp * W + (H - 2 * a.gp) * 2 * a.gp) * (C / 8);
        DISPATCH_T16(dtype, (fold_halo_kernel<T><<<grid_cap(items), 256, 0, st>>>(a, (T*)g_halo)));
        NG_LAUNCH_CHECK("fold_halo_kernel");
        a.halo_mode = NG_HALO_ZERO;          // folded: the passes read the interior only
      }
      DISPATCH_T16(dtype, (r = launch_in_bwd_stream<T, 1>(a, q, st, g_halo, g_skip, y, mean_rstd, sums_scratch, nullptr, nullptr)));
      if (r) return r;
      NG_LAUNCH_CHECK("in_bwd_stream_kernel<pass 1>");
      DISPATCH_T16(dtype, (r = launch_in_bwd_stream<T, 2>(a, q, st, g_halo, g_skip, y, mean_rstd, sums_scratch, dy, do_out)));
      if (r) return r;
      NG_LAUNCH_CHECK("in_bwd_stream_kernel<pass 2>");
      return NG_OK;
    }
  }
  // Lean form: optionally fold the reflect halo once, in place, before the two passes (the pre-pass of the staged form) so
  // that neither pass carries the per-pixel border test and the extra loads.  NIRGAN_B200_BWD_PREFOLD (default 1).
  {
    static const bool prefold = [] { const char* e = getenv("NIRGAN_B200_BWD_PREFOLD"); return !(e && e[0] == '0'); }();
    const long long big = (long long)(H + 2 * a.gp) * (W + 2 * a.gp) * C;
    if (prefold && dtype != NG_F32 && mean_rstd && inject_mode == NG_INJECT_NONE && dscale == nullptr && de_map == nullptr &&
        g_halo && halo_mode == NG_HALO_REFLECT && a.gp > 0 && big < (1ll << 31) && (256 >> a.c8_shift) < W &&
        (act == NG_ACT_RELU || act == NG_ACT_NONE || act == NG_ACT_LRELU)) {
      const long long items = (long long)B * (2 * a.gp * W + (H - 2 * a.gp) * 2 * a.gp) * (C / 8);
      DISPATCH_T16(dtype, (fold_halo_kernel<T><<<grid_cap(items), 256, 0, st>>>(a, (T*)g_halo)));
      NG_LAUNCH_CHECK("fold_halo_kernel");
      a.halo_mode = NG_HALO_ZERO;
    }
  }
  // Image-chunked schedule for normalised units: pass 1 and pass 2 of a chunk of images run back to back, with the chunk
  // sized so that what pass 1 streamed (g + y [+ skip]) is still in the 126 MB L2 when pass 2 re-reads it -- the second
  // read of every unit then comes from L2 instead of HBM.  Measured SLOWER at every chunk size (r2o: 19.30 ms per training
  // step unchunked, 19.90 / 20.81 / 21.59 ms with 80 / 40 / 20 MB chunks -- the extra launches and their partly filled
  // waves cost more than the L2 hits return), so the default NIRGAN_B200_BWD_CHUNK_MB=0 keeps the whole batch per launch.
  static const long long chunk_bytes = [] {
    const char* e = getenv("NIRGAN_B200_BWD_CHUNK_MB");
    return (long long)(e ? atoi(e) : 0) << 20;
  }();
  const int esz = dtype == NG_F32 ? 4 : 2;
  const long long per_image = ((long long)(H + 2 * a.gp) * (W + 2 * a.gp) * (g_halo ? 1 : 0) + (long long)H * W * (g_skip ? 2 : 1)) *
                              C * esz;
  int chunk = B;
  if (need_pass1 && mean_rstd && chunk_bytes > 0 && de_map == nullptr) {
    chunk = (int)(chunk_bytes / (per_image > 0 ? per_image : 1));
    if (chunk < (B + 7) / 8) chunk = (B + 7) / 8;             // at most 8 chunks per unit (launch overhead)
    if (chunk < 1) chunk = 1;
    if (chunk > B) chunk = B;
  }
  const int mult2 = in_bwd_pick_mult(chunk, H * W, pstep, 4, 64);
  const int ppb2 = pstep * (mult2 < 0 ? 16 : mult2);
  int ppb1 = (H * W + a.nblk1 - 1) / a.nblk1;
  ppb1 = (ppb1 + pstep - 1) / pstep * pstep;
  if (need_pass1 && de_map) {
    int e = check_cuda(cudaMemsetAsync(de_map, 0, (size_t)B * H * W * sizeof(float), st), "in_bwd memset de");
    if (e) return e;
  }
  for (int n0 = 0; n0 < B; n0 += chunk) {
    const int nb = B - n0 < chunk ? B - n0 : chunk;
    a.n0 = n0;
    if (need_pass1) {
      a.ppb = ppb1;
      dim3 grid((unsigned)a.nblk1, (unsigned)nb);
      DISPATCH_T(dtype, (launch_in_bwd<T, 1>(a, grid, (size_t)4 * 256 * sizeof(float), st, g_halo, g_skip, y, mean_rstd,
                                             inject_e, inject_scale, mean_rstd ? sums_scratch : nullptr, dscale, de_map,
                                             nullptr, nullptr)));
      NG_LAUNCH_CHECK("in_bwd_kernel<pass 1>");
      if (mean_rstd) {
        const int total = B * C * 2, first = n0 * C * 2, count = nb * C * 2;
        in_bwd_combine_kernel<<<(count + 255) / 256, 256, 0, st>>>(sums_scratch + (size_t)total, first, first + count, C * 2,
                                                                  a.nblk1, sums_scratch);
        NG_LAUNCH_CHECK("in_bwd_combine_kernel");
      }
    }
    a.ppb = ppb2;
    dim3 grid((unsigned)((H * W + a.ppb - 1) / a.ppb), (unsigned)nb);
    DISPATCH_T(dtype, (launch_in_bwd<T, 2>(a, grid, 0, st, g_halo, g_skip, y, mean_rstd, inject_e, inject_scale,
                                           sums_scratch, nullptr, nullptr, dy, do_out)));
    NG_LAUNCH_CHECK("in_bwd_kernel<pass 2>");
  }
  return NG_OK;
}

extern "C" int ng_grad_scale_pow2(const float* g, int64_t n, float target, float* out4, void* stream) {
  int r = require_sm100(); if (r) return r;
  NG_REQUIRE(g && out4 && n > 0 && target > 0.f, NG_E_ARG, "grad_scale_pow2: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  int e = check_cuda(cudaMemsetAsync(out4 + 2, 0, sizeof(float), st), "grad_scale_pow2 memset");
  if (e) return e;
  grad_amax_kernel<<<grid_cap((long long)n), 256, 0, st>>>(g, (long long)n, reinterpret_cast<unsigned*>(out4) + 2);
  NG_LAUNCH_CHECK("grad_amax_kernel");
  grad_scale_finalize_kernel<<<1, 1, 0, st>>>(target, out4);
  NG_LAUNCH_CHECK("grad_scale_finalize_kernel");
  return NG_OK;
}

extern "C" int ng_head_bwd_prep(const float* dout, const float* out, int32_t B, int32_t H, int32_t W, int32_t crop,
                                int32_t act, float scale, const float* dev_scale, int32_t c_pad, int32_t dtype,
                                void* dst, void* stream) {
  int r = require_sm100(); if (r) return r;
  NG_REQUIRE(dout && dst && (act != NG_ACT_TANH || out), NG_E_ARG, "head_bwd_prep: null tensor");
  DISPATCH_T(dtype, (head_bwd_prep_kernel<T><<<grid_cap((long long)B * H * W), 256, 0, (cudaStream_t)stream>>>(
                        dout, out, B, H, W, crop, act, scale, dev_scale, c_pad, (T*)dst)));
  NG_LAUNCH_CHECK("head_bwd_prep_kernel");
  return NG_OK;
}

extern "C" int ng_tap_scatter(const float* dout, const float* out, int32_t B, int32_t Hz, int32_t Wz, int32_t zc,
                              int32_t KH, int32_t KW, int32_t act, int32_t crop, float scale, const float* dev_scale,
                              int32_t dtype, void* dz, void* stream) {
  int r = require_sm100(); if (r) return r;
  NG_REQUIRE(dout && dz && (act != NG_ACT_TANH || out), NG_E_ARG, "tap_scatter: null tensor");
  NG_REQUIRE(KH == 7 && KW == 7 && zc == 64, NG_E_UNSUPPORTED, "tap_scatter: built for 7x7 taps over 64 stored channels");
  NG_REQUIRE(Hz - KH + 1 - 2 * crop > 0 && Wz - KW + 1 - 2 * crop > 0, NG_E_SHAPE, "tap_scatter: empty output");
  NG_REQUIRE((long long)B * Hz * Wz * 8 < (1ll << 31), NG_E_SHAPE, "tap_scatter: batch too large for 32-bit indexing");
  NG_REQUIRE(B <= 65535, NG_E_SHAPE, "tap_scatter: batch %d exceeds the grid's z extent", B);
  const dim3 grid((unsigned)((Wz + 31) / 32), (unsigned)((Hz + 7) / 8), (unsigned)B);
  DISPATCH_T(dtype, (tap_scatter_kernel<T, 7, 7, 64><<<grid, 256, 0, (cudaStream_t)stream>>>(
                        dout, out, B, Hz, Wz, act, crop, scale, dev_scale, (T*)dz)));
  NG_LAUNCH_CHECK("tap_scatter_kernel");
  return NG_OK;
}

extern "C" int ng_grad_to_nchw(const void* src, int32_t dtype, int32_t B, int32_t H, int32_t W, int32_t c_pad, int32_t c0,
                               int32_t c, int32_t s2d, float scale, const float* dev_scale, float* dst, void* stream) {
  int r = require_sm100(); if (r) return r;
  NG_REQUIRE(src && dst && c0 >= 0 && c > 0 && c0 + c <= c_pad, NG_E_ARG, "grad_to_nchw: bad arguments");
  DISPATCH_T(dtype, (grad_to_nchw_kernel<T><<<grid_cap((long long)B * c * H * W), 256, 0, (cudaStream_t)stream>>>(
                        (const T*)src, B, H, W, c_pad, c0, c, s2d, scale, dev_scale, dst)));
  NG_LAUNCH_CHECK("grad_to_nchw_kernel");
  return NG_OK;
}

extern "C" int ng_inject_bwd(const float* de_map, int32_t B, int32_t H, int32_t W, float scale, const float* dev_scale,
                             const float* embeds, float* de128_scratch, float* dfc_w, float* dfc_b, void* stream) {
  int r = require_sm100(); if (r) return r;
  NG_REQUIRE(de_map && embeds && de128_scratch && dfc_w, NG_E_ARG, "inject_bwd: null tensor");
  cudaStream_t st = (cudaStream_t)stream;
  int e = check_cuda(cudaMemsetAsync(de128_scratch, 0, (size_t)B * 128 * 128 * sizeof(float), st), "inject_bwd memset");
  if (e) return e;
  bilerp128_bwd_kernel<<<grid_cap((long long)B * H * W), 256, 0, st>>>(de_map, B, H, W, scale, dev_scale, de128_scratch);
  NG_LAUNCH_CHECK("bilerp128_bwd_kernel");
  linear_bwd_kernel<<<128 * 128, 256, 0, st>>>(de128_scratch, embeds, B, 256, 128 * 128, dfc_w, dfc_b);
  NG_LAUNCH_CHECK("linear_bwd_kernel");
  return NG_OK;
}
