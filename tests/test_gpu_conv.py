"""GPU parity of the convolution kernels (CUDA-core and tcgen05) through the C ABI.

Checker: the same convolution restated with torch fp32 ops on operands rounded to the kernel's
operand precision (TF32 disabled), so only accumulation order and the output rounding differ.
Tolerances (written here, per the parity contract):
  fp32 CUDA-core path    : max-abs <= 2e-5 * max|y|  (+1e-6)
  16-bit operand paths   : max-abs <= 3e-3 * max|y|  (one 16-bit output rounding, fp32 accumulate)
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def _impls():
    from nirgan_b200 import _lib as L
    return [pytest.param(L.IMPL_SIMT, L.F32, id="simt-f32"), pytest.param(L.IMPL_SIMT, L.F16, id="simt-f16"),
            pytest.param(L.IMPL_TC, L.F16, id="tc-f16"), pytest.param(L.IMPL_TC, L.BF16, id="tc-bf16")]


def _tol(dtype, ref):
    from nirgan_b200 import _lib as L
    scale = float(ref.abs().max())
    return (2e-5 if dtype == L.F32 else (3e-3 if dtype == L.F16 else 1.2e-2)) * scale + 1e-6


def _gen(*shape, seed=0, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return torch.randn(*shape, generator=g, device="cuda") * scale


# name, Cin, Cout, K, stride, pad, halo_mode, H, W, B
CONV_CASES = [
    ("stem7x7", 3, 64, 7, 1, 3, "reflect", 32, 32, 2),
    ("stem7x7_odd", 3, 64, 7, 1, 3, "reflect", 36, 44, 1),
    ("down1_s2", 64, 128, 3, 2, 1, "zero", 32, 32, 2),
    ("down2_s2_odd", 128, 256, 3, 2, 1, "zero", 18, 18, 2),
    ("down2_s2_69", 128, 256, 3, 2, 1, "zero", 138, 138, 1),
    ("res3x3", 256, 256, 3, 1, 1, "reflect", 16, 16, 2),
    ("res3x3_64", 256, 256, 3, 1, 1, "reflect", 64, 64, 2),
    ("res3x3_odd21", 256, 256, 3, 1, 1, "reflect", 21, 21, 3),
    ("res3x3_69", 256, 256, 3, 1, 1, "reflect", 69, 69, 1),
    ("d_l1_k4s2", 64, 128, 4, 2, 1, "zero", 32, 32, 2),
    ("d_l2_k4s2", 128, 256, 4, 2, 1, "zero", 16, 16, 2),
    ("d_l3_k4s1", 256, 512, 4, 1, 1, "zero", 9, 9, 2),
]


@pytest.mark.parametrize("impl,dtype", _impls())
@pytest.mark.parametrize("case", CONV_CASES, ids=[c[0] for c in CONV_CASES])
def test_conv_raw_and_stats(case, impl, dtype):
    from nirgan_b200 import _lib as L
    import helpers as Hh
    name, Cin, Cout, K, s, p, mode, H, W, B = case
    x = Hh.rnd(_gen(B, Cin, H, W, seed=1), dtype)
    w = Hh.rnd(_gen(Cout, Cin, K, K, seed=2, scale=0.05), dtype)
    halo = p if mode == "reflect" else 0          # zero padding comes from out-of-bounds reads
    xb = Hh.to_actbuf(x, halo, mode, dtype, c_pad=Hh.rup(Cin, 16))
    wp = Hh.pack_weight(w, 0, Cout, Hh.rup(Cin, 16), dtype)
    Ho, Wo = (H + 2 * p - K) // s + 1, (W + 2 * p - K) // s + 1
    y, mr, _ = Hh.conv_call(xb, wp, Cout, K, s, p, Ho, Wo, dtype, impl, want_stats=True)
    xr = F.pad(x, (p,) * 4, mode="reflect") if mode == "reflect" else F.pad(x, (p,) * 4)
    ref = F.conv2d(xr, w, stride=s)
    got = Hh.from_compact(y, B, Ho, Wo, Cout)
    assert torch.isfinite(got).all(), f"{name}: non-finite output (unwritten rows?)"
    err = float((got - ref).abs().max())
    assert err <= _tol(dtype, ref), f"{name}: max-abs {err:.3e} > {_tol(dtype, ref):.3e}"
    # InstanceNorm statistics describe the *stored* tensor
    mu, rstd = Hh.stats_ref(got)
    assert float((mr[..., 0] - mu).abs().max()) <= 1e-4 * max(1.0, float(ref.abs().max()))
    assert float((mr[..., 1] / rstd - 1).abs().max()) <= 2e-4


CONVT_CASES = [
    ("up1", 256, 128, 8, 8, 2),
    ("up1_odd", 256, 128, 9, 9, 1),
    ("up2", 128, 64, 16, 16, 2),
    ("up2_138", 128, 64, 69, 69, 1),
]


@pytest.mark.parametrize("impl,dtype", _impls())
@pytest.mark.parametrize("case", CONVT_CASES, ids=[c[0] for c in CONVT_CASES])
def test_conv_transpose_phased(case, impl, dtype):
    from nirgan_b200 import _lib as L
    import helpers as Hh
    name, Cin, Cout, H, W, B = case
    x = Hh.rnd(_gen(B, Cin, H, W, seed=3), dtype)
    w = Hh.rnd(_gen(Cin, Cout, 3, 3, seed=4, scale=0.05), dtype)      # ConvTranspose2d layout (Cin, Cout, kh, kw)
    xb = Hh.to_actbuf(x, 0, "zero", dtype)
    wp = Hh.pack_weight(w, 1, Cout, Cin, dtype)
    y, mr, _ = Hh.conv_call(xb, wp, Cout, 3, 2, 1, 2 * H, 2 * W, dtype, impl, form=L.FORM_PHASED, want_stats=True)
    ref = F.conv_transpose2d(x, w, stride=2, padding=1, output_padding=1)
    got = Hh.from_compact(y, B, 2 * H, 2 * W, Cout)
    assert torch.isfinite(got).all()
    err = float((got - ref).abs().max())
    assert err <= _tol(dtype, ref), f"{name}: max-abs {err:.3e}"
    mu, rstd = Hh.stats_ref(got)
    assert float((mr[..., 0] - mu).abs().max()) <= 1e-4 * max(1.0, float(ref.abs().max()))
    assert float((mr[..., 1] / rstd - 1).abs().max()) <= 2e-4


@pytest.mark.parametrize("dt", ["f16", "bf16"])
@pytest.mark.parametrize("case", CONVT_CASES, ids=[c[0] for c in CONVT_CASES])
def test_conv_transpose_merged_phases(case, dt):
    """NG_FORM_PHASED_MERGED (four output phases in GEMM-N, four input shifts in GEMM-K) == ConvTranspose2d, and agrees
    with the four-phase NG_FORM_PHASED kernel (same products, fp32 accumulation in a different order)."""
    import ctypes as C
    from nirgan_b200 import _lib as L
    import helpers as Hh
    dtype = L.F16 if dt == "f16" else L.BF16
    name, Cin, Cout, H, W, B = case
    x = Hh.rnd(_gen(B, Cin, H, W, seed=3), dtype)
    w = Hh.rnd(_gen(Cin, Cout, 3, 3, seed=4, scale=0.05), dtype)
    xb = Hh.to_actbuf(x, 0, "zero", dtype)
    wm = torch.empty(16 * Cout * Cin, dtype=Hh.TORCH_DT[dtype], device="cuda")
    L.call("ng_pack_weight_phasemerged", w.contiguous().data_ptr(), Cin, Cout, dtype, wm.data_ptr(), Hh.stream())
    y, mr, _ = Hh.conv_call(xb, wm, Cout, 3, 2, 1, 2 * H, 2 * W, dtype, L.IMPL_TC, form=L.FORM_PHASED_MERGED,
                            want_stats=True)
    ref = F.conv_transpose2d(x, w, stride=2, padding=1, output_padding=1)
    got = Hh.from_compact(y, B, 2 * H, 2 * W, Cout)
    assert torch.isfinite(got).all()
    err = float((got - ref).abs().max())
    assert err <= _tol(dtype, ref), f"{name}: max-abs {err:.3e}"
    mu, rstd = Hh.stats_ref(got)
    assert float((mr[..., 0] - mu).abs().max()) <= 1e-4 * max(1.0, float(ref.abs().max()))
    assert float((mr[..., 1] / rstd - 1).abs().max()) <= 2e-4
    wp = Hh.pack_weight(w, 1, Cout, Cin, dtype)
    y4, _, _ = Hh.conv_call(xb, wp, Cout, 3, 2, 1, 2 * H, 2 * W, dtype, L.IMPL_TC, form=L.FORM_PHASED)
    assert float((Hh.from_compact(y4, B, 2 * H, 2 * W, Cout) - got).abs().max()) <= _tol(dtype, ref)


@pytest.mark.parametrize("impl,dtype", _impls())
@pytest.mark.parametrize("crop,H", [(0, 32), (10, 44)])
def test_head_conv_tanh(impl, dtype, crop, H):
    from nirgan_b200 import _lib as L
    import helpers as Hh
    B, Cin = 2, 64
    x = Hh.rnd(_gen(B, Cin, H, H, seed=5), dtype)
    w = Hh.rnd(_gen(1, Cin, 7, 7, seed=6, scale=0.02), dtype)
    bias = _gen(1, seed=7, scale=0.1)
    xb = Hh.to_actbuf(x, 3, "reflect", dtype)
    wp = Hh.pack_weight(w, 0, 16, Cin, dtype)
    y, _, _ = Hh.conv_call(xb, wp, 16, 7, 1, 3, H, H, dtype, impl, epilogue=L.EPI_HEAD, act=L.ACT_TANH, crop=crop,
                           bias=bias)
    ref = torch.tanh(F.conv2d(F.pad(x, (3,) * 4, mode="reflect"), w, bias))
    if crop:
        ref = ref[..., crop:-crop, crop:-crop]
    got = y.view(B, 1, H - 2 * crop, H - 2 * crop)
    assert torch.isfinite(got).all()
    assert float((got - ref).abs().max()) <= (2e-5 if dtype == L.F32 else 2e-4)


@pytest.mark.parametrize("impl,dtype", _impls())
def test_patchgan_first_and_last_layer(impl, dtype):
    from nirgan_b200 import _lib as L
    import helpers as Hh
    B, H = 2, 32
    x = Hh.rnd(_gen(B, 4, H, H, seed=8), dtype)
    w = Hh.rnd(_gen(64, 4, 4, 4, seed=9, scale=0.05), dtype)
    bias = _gen(64, seed=10, scale=0.1)
    xb = Hh.to_actbuf(x, 0, "zero", dtype, c_pad=16)
    wp = Hh.pack_weight(w, 0, 64, 16, dtype)
    y, _, _ = Hh.conv_call(xb, wp, 64, 4, 2, 1, H // 2, H // 2, dtype, impl, epilogue=L.EPI_BIAS_ACT, act=L.ACT_LRELU,
                           slope=0.2, bias=bias)
    ref = F.leaky_relu(F.conv2d(x, w, bias, stride=2, padding=1), 0.2)
    got = Hh.from_compact(y, B, H // 2, H // 2, 64)
    assert float((got - ref).abs().max()) <= _tol(dtype, ref)
    # last layer: 512 -> 1, k4 s1 p1, bias, no activation, fp32 output
    x = Hh.rnd(_gen(B, 512, 7, 7, seed=11), dtype)
    w = Hh.rnd(_gen(1, 512, 4, 4, seed=12, scale=0.02), dtype)
    bias = _gen(1, seed=13, scale=0.1)
    xb = Hh.to_actbuf(x, 0, "zero", dtype)
    wp = Hh.pack_weight(w, 0, 16, 512, dtype)
    y, _, _ = Hh.conv_call(xb, wp, 16, 4, 1, 1, 6, 6, dtype, impl, epilogue=L.EPI_HEAD, act=L.ACT_NONE, bias=bias)
    ref = F.conv2d(x, w, bias, stride=1, padding=1)
    assert float((y.view(B, 1, 6, 6) - ref).abs().max()) <= (5e-5 if dtype == L.F32 else 2e-3)


@pytest.mark.parametrize("impl,dtype", _impls())
def test_dgrad_forms(impl, dtype):
    """The three data-gradient shapes of the training step, all served by the forward kernels."""
    from nirgan_b200 import _lib as L
    import helpers as Hh
    B = 2
    # (1) dgrad of a stride-1 3x3 conv on a reflect-haloed input = full correlation with flipped taps
    dy = Hh.rnd(_gen(B, 256, 12, 12, seed=14), dtype)
    w = Hh.rnd(_gen(256, 256, 3, 3, seed=15, scale=0.05), dtype)             # (Cout, Cin, kh, kw)
    xb = Hh.to_actbuf(dy, 0, "zero", dtype)
    wp = Hh.pack_weight(w, 1, 256, 256, dtype)                                 # n = Cin of the forward conv
    y, _, _ = Hh.conv_call(xb, wp, 256, 3, 1, 0, 14, 14, dtype, impl, sgn=-1)
    ref = F.conv_transpose2d(dy, w, stride=1, padding=0)
    assert float((Hh.from_compact(y, B, 14, 14, 256) - ref).abs().max()) <= _tol(dtype, ref)
    # (2) dgrad of a 4x4 stride-2 pad-1 conv (PatchGAN) = phased transposed conv
    dy = Hh.rnd(_gen(B, 128, 8, 8, seed=16), dtype)
    w = Hh.rnd(_gen(128, 64, 4, 4, seed=17, scale=0.05), dtype)
    xb = Hh.to_actbuf(dy, 0, "zero", dtype)
    wp = Hh.pack_weight(w, 1, 64, 128, dtype)
    y, _, _ = Hh.conv_call(xb, wp, 64, 4, 2, 1, 16, 16, dtype, impl, form=L.FORM_PHASED)
    ref = F.conv_transpose2d(dy, w, stride=2, padding=1)
    assert float((Hh.from_compact(y, B, 16, 16, 64) - ref).abs().max()) <= _tol(dtype, ref)
    # (3) dgrad of ConvTranspose2d(k3,s2,p1,op1) = stride-2 3x3 conv of dy with pad 1
    dy = Hh.rnd(_gen(B, 128, 16, 16, seed=18), dtype)
    w = Hh.rnd(_gen(256, 128, 3, 3, seed=19, scale=0.05), dtype)             # ConvT weight (Cin=256, Cout=128)
    xb = Hh.to_actbuf(dy, 0, "zero", dtype)
    wp = Hh.pack_weight(w, 0, 256, 128, dtype)
    y, _, _ = Hh.conv_call(xb, wp, 256, 3, 2, 1, 8, 8, dtype, impl)
    ref = F.conv2d(dy, w, stride=2, padding=1)
    assert float((Hh.from_compact(y, B, 8, 8, 256) - ref).abs().max()) <= _tol(dtype, ref)


def test_argument_errors_are_reported():
    """Error behaviour of the C ABI: negative status + message, never a crash."""
    import ctypes as C
    from nirgan_b200 import _lib as L
    a = L.ConvArgs()
    assert L.load().ng_conv2d(C.byref(a), None) < 0
    assert "conv" in L.last_error()
    assert L.load().ng_conv2d(None, None) < 0
    with pytest.raises(RuntimeError):
        L.call("ng_in_apply", None, L.F16, 1, 4, 4, 8, None, None, None, 0, 0.0, None, 0, None, 0, None, None, 0, 0, None)


@pytest.mark.parametrize("impl,dtype", _impls())
@pytest.mark.parametrize("wrap,H,W", [(0, 32, 32), (10, 24, 36)])
def test_stem_rowmerged(impl, dtype, wrap, H, W):
    """ng_prep_stem + 7x1 conv over 64 merged (kw, c) channels == ReflectionPad2d(3) + Conv2d(3->64, k7)
    on the (optionally wrapper-padded) tile."""
    import ctypes as C
    from nirgan_b200 import _lib as L
    from nirgan_b200.engine import ActBuf
    import helpers as Hh
    B = 2
    x = _gen(B, 3, H, W, seed=21)
    w = Hh.rnd(_gen(64, 3, 7, 7, seed=22, scale=0.05), dtype)
    H1, W1 = H + 2 * wrap, W + 2 * wrap
    x0 = torch.empty(B * (H1 + 6) * W1 * 64, dtype=Hh.TORCH_DT[dtype], device="cuda")
    L.call("ng_prep_stem", x.data_ptr(), 3, B, H, W, wrap, 3, 7, dtype, x0.data_ptr(), Hh.stream())
    wp = torch.empty(7 * 64 * 64, dtype=Hh.TORCH_DT[dtype], device="cuda")
    L.call("ng_pack_weight_rowmerged", w.data_ptr(), 64, 3, 7, 7, 8, dtype, wp.data_ptr(), Hh.stream())
    y = torch.full((B * H1 * W1 * 64,), float("nan"), device="cuda").to(Hh.TORCH_DT[dtype])
    a = L.ConvArgs()
    a.dtype, a.impl, a.form, a.sgn = dtype, impl, L.FORM_GATHER, 1
    a.B, a.Hin, a.Win, a.Cin, a.in_pad, a.in_pad_w = B, H1, W1, 64, 3, 0
    a.Cout, a.KH, a.KW, a.stride, a.pad, a.pad_w, a.Hout, a.Wout = 64, 7, 1, 1, 3, 0, H1, W1
    a.x, a.w, a.y = x0.data_ptr(), wp.data_ptr(), y.data_ptr()
    L.call("ng_conv2d", C.byref(a), Hh.stream())
    xr = Hh.rnd(x, dtype)
    if wrap:
        xr = F.pad(xr, (wrap,) * 4, mode="reflect")
    ref = F.conv2d(F.pad(xr, (3,) * 4, mode="reflect"), w)
    got = Hh.from_compact(y, B, H1, W1, 64)
    assert torch.isfinite(got).all()
    assert float((got - ref).abs().max()) <= _tol(dtype, ref)
    # weight-gradient unpack is the exact inverse of the row-merged pack
    back = torch.empty_like(w)
    L.call("ng_unpack_weight_grad_rowmerged", wp.float().contiguous().data_ptr(), 64, 3, 7, 7, 1.0, None, 0.0,
           back.data_ptr(), Hh.stream())
    assert torch.equal(back, w)
    # beta = 1 accumulates (autograd's AccumulateGrad for a parameter reached twice in one pass)
    L.call("ng_unpack_weight_grad_rowmerged", wp.float().contiguous().data_ptr(), 64, 3, 7, 7, 0.5, None, 1.0,
           back.data_ptr(), Hh.stream())
    assert torch.equal(back, w + 0.5 * w)


@pytest.mark.parametrize("impl,dtype", _impls())
@pytest.mark.parametrize("crop,H", [(0, 32), (10, 44)])
def test_head_tap_gemm_and_gather(impl, dtype, crop, H):
    """1x1 'tap GEMM' (Cout = 49 taps -> 64) over the haloed buffer + ng_tap_gather == Conv2d(64->1, k7) + Tanh."""
    from nirgan_b200 import _lib as L
    from nirgan_b200.engine import ActBuf
    import helpers as Hh
    B, Cin = 2, 64
    x = Hh.rnd(_gen(B, Cin, H, H, seed=5), dtype)
    w = Hh.rnd(_gen(1, Cin, 7, 7, seed=6, scale=0.02), dtype)
    bias = _gen(1, seed=7, scale=0.1)
    xb = Hh.to_actbuf(x, 3, "reflect", dtype)
    xz = ActBuf(xb.t, B, H + 6, H + 6, Cin, 0)
    wt = torch.zeros(64 * 64, dtype=Hh.TORCH_DT[dtype], device="cuda")
    L.call("ng_pack_weight", w.data_ptr(), 1, Cin, 7, 7, 0, 1, 64, dtype, wt.data_ptr(), Hh.stream())
    z, _, _ = Hh.conv_call(xz, wt, 64, 1, 1, 0, H + 6, H + 6, dtype, impl)
    out = torch.full((B * (H - 2 * crop) ** 2,), float("nan"), device="cuda")
    L.call("ng_tap_gather", z.data_ptr(), dtype, B, H + 6, H + 6, 64, 7, 7, bias.data_ptr(), L.ACT_TANH, crop,
           out.data_ptr(), Hh.stream())
    ref = torch.tanh(F.conv2d(F.pad(x, (3,) * 4, mode="reflect"), w, bias))
    if crop:
        ref = ref[..., crop:-crop, crop:-crop]
    got = out.view(B, 1, H - 2 * crop, H - 2 * crop)
    assert torch.isfinite(got).all()
    assert float((got - ref).abs().max()) <= (2e-5 if dtype == L.F32 else (1e-3 if dtype == L.F16 else 1e-2))   # 49 16-bit-rounded partial sums



@pytest.mark.parametrize("dtype_name", ["f16", "bf16"])
@pytest.mark.parametrize("crop,H,W,B", [(0, 32, 32, 2), (10, 44, 44, 2), (10, 37, 61, 3), (0, 8, 16, 1), (10, 276, 276, 2),
                                        (3, 150, 40, 5)])
def test_head_fused_kernel(dtype_name, crop, H, W, B):
    """ng_head_conv (haloed patch by TMA, tap GEMM of the whole patch on tcgen05, 49-tap gather from the shared z tile) ==
    ReflectionPad2d(3) + Conv2d(64 -> 1, k7) + Tanh with the wrapper's crop, incl. partially covered 8 x 16 patches, more
    tiles than CTAs (both epilogue groups, both accumulators, stage reuse), and launch-to-launch bit-exactness."""
    from nirgan_b200 import _lib as L
    import helpers as Hh
    dtype = L.F16 if dtype_name == "f16" else L.BF16
    Cin = 64
    x = Hh.rnd(_gen(B, Cin, H, W, seed=5), dtype)
    w = Hh.rnd(_gen(1, Cin, 7, 7, seed=6, scale=0.02), dtype)
    bias = _gen(1, seed=7, scale=0.1)
    xb = Hh.to_actbuf(x, 3, "reflect", dtype)
    wt = torch.zeros(64 * 64, dtype=Hh.TORCH_DT[dtype], device="cuda")
    L.call("ng_pack_weight", w.data_ptr(), 1, Cin, 7, 7, 0, 1, 64, dtype, wt.data_ptr(), Hh.stream())
    Hc, Wc = H - 2 * crop, W - 2 * crop
    outs = []
    for _ in range(2):
        out = torch.full((B * Hc * Wc,), float("nan"), device="cuda")
        L.call("ng_head_conv", xb.t.data_ptr(), dtype, B, H, W, 64, 7, 3, 3, wt.data_ptr(), bias.data_ptr(), L.ACT_TANH, crop,
               out.data_ptr(), Hh.stream())
        torch.cuda.synchronize()
        outs.append(out)
    assert torch.equal(outs[0], outs[1])
    ref = torch.tanh(F.conv2d(F.pad(x, (3,) * 4, mode="reflect"), w, bias))
    if crop:
        ref = ref[..., crop:-crop, crop:-crop]
    got = outs[0].view(B, 1, Hc, Wc)
    assert torch.isfinite(got).all()
    # 49 partial sums of magnitude ~0.2 rounded to 16 bits before the fp32 sum (as in the tap GEMM + gather pair)
    assert float((got - ref).abs().max()) <= (1.5e-3 if dtype == L.F16 else 1.2e-2)


@pytest.mark.parametrize("dtype_name", ["f32", "f16", "bf16"])
@pytest.mark.parametrize("crop,H,W,B", [(0, 32, 32, 2), (10, 44, 44, 2), (3, 37, 61, 3), (0, 8, 8, 1)])
def test_tap_scatter_matches_definition(dtype_name, crop, H, W, B):
    """ng_tap_scatter (adjoint of the head's tap gather, tiled through shared memory):
    dz[n][yy][xx][kh*7+kw] = scale * dout[n][yy-kh-crop][xx-kw-crop] * (1 - out^2), zero outside the cropped window and
    for the 15 padding taps; H x W = the head conv's output before the crop."""
    from nirgan_b200 import _lib as L
    import helpers as Hh
    dtype = {"f32": L.F32, "f16": L.F16, "bf16": L.BF16}[dtype_name]
    Hz, Wz, Hc, Wc = H + 6, W + 6, H - 2 * crop, W - 2 * crop
    dout = _gen(B, Hc, Wc, seed=3)
    out = torch.tanh(_gen(B, Hc, Wc, seed=4))
    dev_scale = torch.tensor([0.5], device="cuda")
    dz = torch.full((B * Hz * Wz * 64,), float("nan"), device="cuda").to(Hh.TORCH_DT[dtype])
    L.call("ng_tap_scatter", dout.data_ptr(), out.data_ptr(), B, Hz, Wz, 64, 7, 7, L.ACT_TANH, crop, 3.0, dev_scale.data_ptr(),
           dtype, dz.data_ptr(), Hh.stream())
    torch.cuda.synchronize()
    gm = 1.5 * dout * (1 - out * out)
    ref = torch.zeros(B, Hz, Wz, 64, device="cuda")
    for kh in range(7):
        for kw in range(7):
            ref[:, kh + crop:kh + crop + Hc, kw + crop:kw + crop + Wc, kh * 7 + kw] = gm
    got = dz.view(B, Hz, Wz, 64).float()
    assert bool(torch.isfinite(got).all())
    tol = 1e-6 if dtype == L.F32 else (2.0 ** -10 if dtype == L.F16 else 2.0 ** -7)
    assert float((got - ref).abs().max()) <= tol * float(ref.abs().max())


@pytest.mark.parametrize("dtype_name", ["f16", "bf16"])
@pytest.mark.parametrize("halo", [0, 1])
@pytest.mark.parametrize("H,W,B", [(30, 30, 3), (31, 31, 64), (8, 16, 1), (13, 45, 2), (62, 62, 20)])
def test_patchgan_last_layer_fused_kernel(dtype_name, H, W, B, halo):
    """ng_head_conv in its 512 -> 1, k4, p1 form (eight 64-channel K chunks accumulated in TMEM, 16 taps, zero halo) ==
    Conv2d(512, 1, 4, 1, 1) over the same 16-bit-rounded operands; ragged patches, more tiles than CTAs, bit-exact
    from launch to launch."""
    from nirgan_b200 import _lib as L
    import helpers as Hh
    dtype = L.F16 if dtype_name == "f16" else L.BF16
    Cin, K = 512, 4
    Hi, Wi = H + 1, W + 1                                       # input size whose k4 p1 s1 output is H x W
    x = Hh.rnd(_gen(B, Cin, Hi, Wi, seed=5), dtype)
    w = Hh.rnd(_gen(1, Cin, K, K, seed=6, scale=0.02), dtype)
    bias = _gen(1, seed=7, scale=0.1)
    xb = Hh.to_actbuf(x, halo, "zero", dtype)                   # halo 1: padding materialised; 0: all of it by TMA zero fill
    wt = torch.zeros(16 * Cin, dtype=Hh.TORCH_DT[dtype], device="cuda")
    L.call("ng_pack_weight", w.data_ptr(), 1, Cin, K, K, 0, 1, Cin, dtype, wt.data_ptr(), Hh.stream())
    outs = []
    for _ in range(2):
        out = torch.full((B * H * W,), float("nan"), device="cuda")
        L.call("ng_head_conv", xb.t.data_ptr(), dtype, B, H, W, Cin, K, 1, halo, wt.data_ptr(), bias.data_ptr(), L.ACT_NONE, 0,
               out.data_ptr(), Hh.stream())
        torch.cuda.synchronize()
        outs.append(out)
    assert torch.equal(outs[0], outs[1])
    ref = F.conv2d(x, w, bias, padding=1)
    assert ref.shape[-2:] == (H, W)
    got = outs[0].view(B, 1, H, W)
    assert torch.isfinite(got).all()
    assert float((got - ref).abs().max()) <= 2e-4               # identical 16-bit operands, fp32 all the way (fp32 z tile)


@pytest.mark.parametrize("dtype_name", ["f16", "bf16"])
@pytest.mark.parametrize("cin,wrap,H,W,B", [(3, 0, 32, 32, 2), (3, 10, 24, 36, 2), (3, 0, 50, 70, 3), (4, 10, 44, 44, 1),
                                            (3, 10, 256, 256, 2)])
def test_stem_direct_from_fp32_tiles(dtype_name, cin, wrap, H, W, B):
    """ng_stem_conv (im2col tile assembled in shared memory from the NCHW fp32 tiles, K = 7 x 32, 64-byte swizzle) ==
    F.pad(reflect, wrap) + ReflectionPad2d(3) + Conv2d(cin -> 64, k7), incl. partially covered 8 x 16 patches at the right /
    bottom border, both statistics forms, and launch-to-launch bit-exactness."""
    from nirgan_b200 import _lib as L
    import helpers as Hh
    dtype = L.F16 if dtype_name == "f16" else L.BF16
    x = _gen(B, cin, H, W, seed=31)
    w = Hh.rnd(_gen(64, cin, 7, 7, seed=32, scale=0.05), dtype)
    H1, W1 = H + 2 * wrap, W + 2 * wrap
    wp = torch.empty(7 * 64 * 32, dtype=Hh.TORCH_DT[dtype], device="cuda")
    L.call("ng_pack_weight_rowmerged", w.data_ptr(), 64, cin, 7, 7, 4, dtype, wp.data_ptr(), Hh.stream())
    slots = int(L.load().ng_stem_conv_stat_slots(H, W, wrap))
    outs = []
    for rep in range(2):
        y = torch.full((B * H1 * W1 * 64,), float("nan"), device="cuda").to(Hh.TORCH_DT[dtype])
        part = torch.full((B * slots * 64 * 2,), float("nan"), device="cuda")
        acc = torch.zeros(B * 64 * 2, dtype=torch.int64, device="cuda")
        L.call("ng_stem_conv", x.data_ptr(), cin, B, H, W, wrap, wp.data_ptr(), dtype, y.data_ptr(), part.data_ptr(),
               acc.data_ptr(), Hh.stream())
        torch.cuda.synchronize()
        outs.append((y, part, acc))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][2], outs[1][2])
    y, part, acc = outs[0]
    xr = Hh.rnd(x, dtype)
    if wrap:
        xr = F.pad(xr, (wrap,) * 4, mode="reflect")
    ref = F.conv2d(F.pad(xr, (3,) * 4, mode="reflect"), w)
    got = Hh.from_compact(y, B, H1, W1, 64)
    assert torch.isfinite(got).all()
    assert float((got - ref).abs().max()) <= _tol(dtype, ref)
    # statistics of the STORED (rounded) output: per-tile partials -> finalize, and the fixed-point accumulators
    mr = torch.empty(B * 64 * 2, device="cuda")
    L.call("ng_in_stats_finalize", part.data_ptr(), B, slots, 64, H1 * W1, mr.data_ptr(), Hh.stream())
    mu, rstd = Hh.stats_ref(got)
    m = mr.view(B, 64, 2)
    assert float((m[..., 0] - mu).abs().max()) <= 1e-4 * max(1.0, float(mu.abs().max()))
    assert float((m[..., 1] / rstd - 1).abs().max()) <= 1e-3
    a = acc.view(B, 64, 2).double()
    mean_acc = a[..., 0] / 2 ** 24 / (H1 * W1)
    var_acc = a[..., 1] / 2 ** 20 / (H1 * W1) - mean_acc ** 2
    assert float((mean_acc - mu.double()).abs().max()) <= 1e-4 * max(1.0, float(mu.abs().max()))
    assert float((torch.rsqrt(var_acc + 1e-5) / rstd.double() - 1).abs().max()) <= 1e-3
