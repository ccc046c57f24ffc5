// Host-side helpers: error reporting, device checks, convolution geometry (tap tables).
#include "common.cuh"
#include <stdarg.h>
#include <mutex>

namespace ng {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return 0;
  set_error("%s: CUDA error %d (%s)", what, (int)e, cudaGetErrorString(e));
  return (int)e;
}

static int g_sm_major[64];
static int g_sm_count[64];
static bool g_dev_known[64];
static std::mutex g_dev_mu;

static int query_device(int dev) {
  if (dev < 0 || dev >= 64) return NG_E_ARG;
  std::lock_guard<std::mutex> lk(g_dev_mu);
  if (!g_dev_known[dev]) {
    cudaDeviceProp p;
    cudaError_t e = cudaGetDeviceProperties(&p, dev);
    if (e != cudaSuccess) {
      set_error("cudaGetDeviceProperties(%d): %s", dev, cudaGetErrorString(e));
      return NG_E_ARCH;
    }
    g_sm_major[dev] = p.major;
    g_sm_count[dev] = p.multiProcessorCount;
    g_dev_known[dev] = true;
  }
  return NG_OK;
}

int require_sm100() {
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    set_error("no CUDA device: nirgan_b200 has no CPU fallback");
    return NG_E_ARCH;
  }
  int r = query_device(dev);
  if (r) return r;
  if (g_sm_major[dev] != 10) {
    set_error("device %d is sm_%d0; nirgan_b200 kernels are built for sm_100a only (no fallback)", dev,
              g_sm_major[dev]);
    return NG_E_ARCH;
  }
  return NG_OK;
}

int num_sms() {
  int dev = 0;
  cudaGetDevice(&dev);
  if (query_device(dev)) return 148;
  return g_sm_count[dev];
}

int build_geometry(const ng_conv_args& a, ConvGeom& g) {
  NG_REQUIRE(a.B > 0 && a.Hin > 0 && a.Win > 0 && a.Cin > 0 && a.Cout > 0, NG_E_SHAPE, "conv: empty shape");
  NG_REQUIRE(a.KH > 0 && a.KW > 0 && a.KH * a.KW <= 64, NG_E_SHAPE, "conv: kernel %dx%d unsupported", a.KH, a.KW);
  NG_REQUIRE(a.stride == 1 || a.stride == 2, NG_E_UNSUPPORTED, "conv: stride %d unsupported", a.stride);
  NG_REQUIRE(a.in_pad >= 0 && a.pad >= 0 && a.in_pad_w >= 0 && a.pad_w >= 0, NG_E_ARG, "conv: negative padding");
  memset(&g, 0, sizeof(g));
  g.B = a.B;
  g.Hb = a.Hin + 2 * a.in_pad;
  g.Wb = a.Win + 2 * a.in_pad_w;
  g.Cin = a.Cin;
  g.Cout = a.Cout;
  g.Hout = a.Hout;
  g.Wout = a.Wout;
  if (a.form == NG_FORM_GATHER) {
    NG_REQUIRE(a.sgn == 1 || a.sgn == -1, NG_E_ARG, "conv: sgn must be +-1");
    NG_REQUIRE(a.sgn == 1 || a.stride == 1, NG_E_UNSUPPORTED, "conv: flipped taps need stride 1");
    int eh = a.sgn == 1 ? (a.Hin + 2 * a.pad - a.KH) / a.stride + 1 : a.Hin + a.KH - 1 - 2 * a.pad;
    int ew = a.sgn == 1 ? (a.Win + 2 * a.pad_w - a.KW) / a.stride + 1 : a.Win + a.KW - 1 - 2 * a.pad_w;
    NG_REQUIRE(a.Hout == eh && a.Wout == ew, NG_E_SHAPE, "conv: Hout/Wout %dx%d, expected %dx%d", a.Hout, a.Wout,
               eh, ew);
    g.VH = a.Hout; g.VW = a.Wout; g.S = a.stride; g.OS = 1; g.nphase = 1;
    g.phase_tap0[0] = 0;
    int t = 0;
    for (int kh = 0; kh < a.KH; ++kh)
      for (int kw = 0; kw < a.KW; ++kw, ++t) {
        g.taps[t].dy = (int16_t)(a.sgn * (kh - a.pad) + a.in_pad);
        g.taps[t].dx = (int16_t)(a.sgn * (kw - a.pad_w) + a.in_pad_w);
        g.taps[t].wrow = (kh * a.KW + kw) * a.Cout;
      }
    g.ntaps = t;
    g.phase_tap0[1] = t;
  } else if (a.form == NG_FORM_PHASED) {
    NG_REQUIRE(a.stride == 2, NG_E_UNSUPPORTED, "phased conv: stride must be 2");
    NG_REQUIRE(a.pad == a.pad_w && a.in_pad == a.in_pad_w && a.KH == a.KW, NG_E_UNSUPPORTED,
               "phased conv: symmetric kernels / padding only");
    int s = a.stride;
    g.VH = (a.Hout + s - 1) / s; g.VW = (a.Wout + s - 1) / s; g.S = 1; g.OS = s; g.nphase = s * s;
    int t = 0;
    for (int pa = 0; pa < s; ++pa)
      for (int pb = 0; pb < s; ++pb) {
        int ph = pa * s + pb;
        g.phase_tap0[ph] = t;
        g.phase_oy[ph] = pa; g.phase_ox[ph] = pb;
        for (int kh = 0; kh < a.KH; ++kh) {
          if (((pa + a.pad - kh) % s) != 0) continue;
          for (int kw = 0; kw < a.KW; ++kw) {
            if (((pb + a.pad - kw) % s) != 0) continue;
            g.taps[t].dy = (int16_t)((pa + a.pad - kh) / s + a.in_pad);
            g.taps[t].dx = (int16_t)((pb + a.pad - kw) / s + a.in_pad);
            g.taps[t].wrow = (kh * a.KW + kw) * a.Cout;
            ++t;
          }
        }
      }
    g.ntaps = t;
    g.phase_tap0[g.nphase] = t;
    // every input pixel contributes to out[s*i - pad + k]: the largest index must fit
    NG_REQUIRE((a.Hin - 1) * s - a.pad + a.KH - 1 <= a.Hout + s - 1 && a.Hout <= (a.Hin - 1) * s - 2 * a.pad + a.KH + s - 1,
               NG_E_SHAPE, "phased conv: Hout %d inconsistent with Hin %d", a.Hout, a.Hin);
  } else if (a.form == NG_FORM_PHASED_MERGED) {
    NG_REQUIRE(a.stride == 2 && a.KH == 3 && a.KW == 3 && a.pad == 1 && a.pad_w == 1, NG_E_UNSUPPORTED,
               "merged-phase conv: built for ConvTranspose2d(k3, s2, p1, op1)");
    NG_REQUIRE(a.in_pad == a.in_pad_w && a.Hout == 2 * a.Hin && a.Wout == 2 * a.Win, NG_E_SHAPE,
               "merged-phase conv: output must be 2x the input");
    g.merged = 1;
    g.Cout_real = a.Cout;
    g.Cout = 4 * a.Cout;
    g.VH = a.Hin; g.VW = a.Win; g.S = 1; g.OS = 2; g.nphase = 1;
    g.phase_tap0[0] = 0;
    int t = 0;
    for (int sy = 0; sy < 2; ++sy)
      for (int sx = 0; sx < 2; ++sx, ++t) {
        g.taps[t].dy = (int16_t)(sy + a.in_pad);
        g.taps[t].dx = (int16_t)(sx + a.in_pad);
        g.taps[t].wrow = t * g.Cout;
      }
    g.ntaps = t;
    g.phase_tap0[1] = t;
  } else {
    set_error("conv: unknown form %d", a.form);
    return NG_E_ARG;
  }
  if (!g.merged) g.Cout_real = g.Cout;
  return NG_OK;
}

}  // namespace ng

extern "C" int ng_version(void) { return NG_VERSION; }
extern "C" const char* ng_last_error(void) { return ng::g_err; }
extern "C" int ng_device_check(int device) {
  int r = ng::query_device(device);
  if (r) return r;
  if (ng::g_sm_major[device] != 10) {
    ng::set_error("device %d is not sm_100", device);
    return NG_E_ARCH;
  }
  return NG_OK;
}
