// Shared helpers for the nirgan_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "../../include/nirgan_b200.h"

namespace ng {

void set_error(const char* fmt, ...);
int  check_cuda(cudaError_t e, const char* what);   // returns 0 or positive cuda error (and sets message)
int  require_sm100();                                // NG_OK or NG_E_ARCH for the current device
int  num_sms();

#define NG_REQUIRE(cond, code, ...)                  \
  do {                                               \
    if (!(cond)) {                                   \
      ng::set_error(__VA_ARGS__);                    \
      return (code);                                 \
    }                                                \
  } while (0)

#define NG_LAUNCH_CHECK(name)                                        \
  do {                                                               \
    int _e = ng::check_cuda(cudaGetLastError(), name);               \
    if (_e) return _e;                                               \
  } while (0)

// ---- geometry shared by the SIMT and tcgen05 convolution kernels ------------------------------
// A convolution is a sum over "taps" of shifted GEMMs.  A work tile covers a patch of "virtual
// pixels" (i, j); tap t reads input buffer position (S*i + dy_t, S*j + dx_t) (zero outside the
// buffer) and the virtual pixel maps to output pixel (OS*i + oy_phase, OS*j + ox_phase).
struct ConvTap {
  int16_t dy, dx;   // offset in *buffer* coordinates (halo included)
  int32_t wrow;     // first row of this tap's [Cout][Cin] weight slab in the packed weight matrix
};

struct ConvGeom {
  int B, Hb, Wb, Cin;        // input buffer dims (halo included)
  int Cout;
  int VH, VW;                // virtual pixel grid (per phase)
  int S, OS;                 // input / output step per virtual pixel
  int Hout, Wout;
  int nphase;
  int phase_tap0[5];
  int phase_oy[4], phase_ox[4];
  int ntaps;
  int merged;                // NG_FORM_PHASED_MERGED: Cout is the virtual 4*Cout_real, phases live in the channel index
  int Cout_real;
  ConvTap taps[64];
};

int build_geometry(const ng_conv_args& a, ConvGeom& g);   // NG_OK or error

// ---- element conversion -------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// pack two floats into one 32-bit word of 16-bit elements (lo = a)
template <typename T> __device__ __forceinline__ uint32_t pack2(float a, float b);
template <> __device__ __forceinline__ uint32_t pack2<__half>(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
template <> __device__ __forceinline__ uint32_t pack2<__nv_bfloat16>(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
template <typename T> __device__ __forceinline__ float2 unpack2(uint32_t w);
template <> __device__ __forceinline__ float2 unpack2<__half>(uint32_t w) {
  return __half22float2(*reinterpret_cast<__half2*>(&w));
}
template <> __device__ __forceinline__ float2 unpack2<__nv_bfloat16>(uint32_t w) {
  return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&w));
}

__device__ __forceinline__ int reflect_idx(int i, int n) {   // ReflectionPad semantics, |overhang| < n
  i = i < 0 ? -i : i;
  return i >= n ? 2 * (n - 1) - i : i;
}

__device__ __forceinline__ float apply_act(float v, int act, float slope) {
  switch (act) {
    case NG_ACT_RELU:  return fmaxf(v, 0.f);
    case NG_ACT_LRELU: return v > 0.f ? v : v * slope;
    case NG_ACT_TANH:  return tanhf(v);
    default:           return v;
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace ng
