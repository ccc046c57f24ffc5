// extern "C" convolution entry points: validate, build the tap-table geometry, dispatch.
#include "common.cuh"

namespace ng {
int conv_simt(const ng_conv_args& a, const ConvGeom& g, cudaStream_t st);
int wgrad_simt(const ng_conv_args& a, const ConvGeom& g, float* dw, void* workspace, long long workspace_bytes,
               cudaStream_t st);
long long wgrad_simt_workspace_bytes(const ng_conv_args& a, const ConvGeom& g);
int bias_grad(const ng_conv_args& a, const ConvGeom& g, float* dbias, cudaStream_t st);
int wgrad_tc(const ng_conv_args& a, const ConvGeom& g, float* dw, void* workspace, long long workspace_bytes,
             cudaStream_t st, bool* handled);
long long wgrad_tc_workspace_bytes(const ng_conv_args& a, const ConvGeom& g);
int conv_tc(const ng_conv_args& a, const ConvGeom& g, cudaStream_t st);
int conv_tc_stat_slots(const ng_conv_args& a, const ConvGeom& g);
int conv_tc_stem_direct(const float* src, int cin, int B, int H, int W, int wrap, const void* w_packed, int dtype, void* y,
                        float* stat_partials, long long* stat_acc, cudaStream_t st);
}  // namespace ng

using namespace ng;

static int validate(const ng_conv_args* a) {
  NG_REQUIRE(a != nullptr, NG_E_ARG, "conv: null argument block");
  NG_REQUIRE(a->x && a->w && a->y, NG_E_ARG, "conv: null tensor pointer");
  NG_REQUIRE(a->dtype == NG_F32 || a->dtype == NG_F16 || a->dtype == NG_BF16, NG_E_ARG, "conv: bad dtype %d", a->dtype);
  NG_REQUIRE(a->epilogue >= NG_EPI_RAW && a->epilogue <= NG_EPI_HEAD, NG_E_ARG, "conv: bad epilogue %d", a->epilogue);
  NG_REQUIRE(a->Cin % 16 == 0, NG_E_SHAPE, "conv: stored Cin %d must be a multiple of 16", a->Cin);
  NG_REQUIRE(a->Cout % 8 == 0, NG_E_SHAPE, "conv: stored Cout %d must be a multiple of 8", a->Cout);
  NG_REQUIRE(a->crop >= 0 && (a->epilogue == NG_EPI_HEAD || a->crop == 0), NG_E_ARG, "conv: crop only with the head epilogue");
  return NG_OK;
}

extern "C" int ng_conv_stat_slots(const ng_conv_args* a) {
  if (!a) return NG_E_ARG;
  ConvGeom g;
  int r = build_geometry(*a, g);
  if (r) return r;
  if (a->impl != NG_IMPL_TC) return 0;
  return conv_tc_stat_slots(*a, g);
}

extern "C" int ng_conv2d(const ng_conv_args* a, void* stream) {
  int r = require_sm100();
  if (r) return r;
  r = validate(a);
  if (r) return r;
  ConvGeom g;
  r = build_geometry(*a, g);
  if (r) return r;
  NG_REQUIRE(!g.merged || (a->impl == NG_IMPL_TC && a->epilogue == NG_EPI_RAW), NG_E_UNSUPPORTED,
             "conv: the merged-phase form runs on the tcgen05 kernel with the RAW epilogue only");
  if (a->impl == NG_IMPL_SIMT) return conv_simt(*a, g, (cudaStream_t)stream);
  if (a->impl == NG_IMPL_TC) return conv_tc(*a, g, (cudaStream_t)stream);
  set_error("conv: unknown impl %d", a->impl);
  return NG_E_ARG;
}

extern "C" int64_t ng_conv2d_wgrad_workspace_bytes(const ng_conv_args* a) {
  if (!a) return NG_E_ARG;
  ConvGeom g;
  int r = build_geometry(*a, g);
  if (r) return r;
  const long long simt = wgrad_simt_workspace_bytes(*a, g);      // single-output-channel layers (either impl)
  if (a->impl != NG_IMPL_TC) return simt;
  const long long tc = wgrad_tc_workspace_bytes(*a, g);
  return tc > simt ? tc : simt;
}

extern "C" int ng_conv2d_wgrad(const ng_conv_args* a, float* dw_packed, float* dbias, void* workspace,
                               int64_t workspace_bytes, void* stream) {
  int r = require_sm100();
  if (r) return r;
  r = validate(a);
  if (r) return r;
  NG_REQUIRE(dw_packed != nullptr, NG_E_ARG, "wgrad: null output");
  NG_REQUIRE(a->form != NG_FORM_PHASED_MERGED, NG_E_UNSUPPORTED, "wgrad: describe the ConvTranspose as NG_FORM_PHASED");
  ConvGeom g;
  r = build_geometry(*a, g);
  if (r) return r;
  bool handled = false;
  if (a->impl == NG_IMPL_TC) {
    r = wgrad_tc(*a, g, dw_packed, workspace, workspace_bytes, (cudaStream_t)stream, &handled);
    if (r) return r;
  }
  if (!handled) {
    r = wgrad_simt(*a, g, dw_packed, workspace, workspace_bytes, (cudaStream_t)stream);
    if (r) return r;
  }
  if (dbias) return bias_grad(*a, g, dbias, (cudaStream_t)stream);
  return NG_OK;
}

extern "C" int ng_stem_conv(const float* src, int32_t cin, int32_t B, int32_t H, int32_t W, int32_t wrap_pad,
                            const void* w_packed, int32_t dtype, void* y, float* stat_partials, int64_t* stat_acc,
                            void* stream) {
  int r = require_sm100();
  if (r) return r;
  return conv_tc_stem_direct(src, cin, B, H, W, wrap_pad, w_packed, dtype, y, stat_partials,
                             reinterpret_cast<long long*>(stat_acc), (cudaStream_t)stream);
}

extern "C" int ng_stem_conv_stat_slots(int32_t H, int32_t W, int32_t wrap_pad) {
  const int H1 = H + 2 * wrap_pad, W1 = W + 2 * wrap_pad;
  return ((H1 + 7) / 8) * ((W1 + 15) / 16);
}
