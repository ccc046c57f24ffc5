"""GPU parity of the memory-bound kernels through the C ABI, against the oracle's arithmetic."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def _gen(*shape, seed=0, scale=1.0, uniform=False):
    g = torch.Generator(device="cuda").manual_seed(seed)
    if uniform:
        return torch.rand(*shape, generator=g, device="cuda") * scale
    return torch.randn(*shape, generator=g, device="cuda") * scale


def _dtypes():
    from nirgan_b200 import _lib as L
    return [pytest.param(L.F32, id="f32"), pytest.param(L.F16, id="f16"), pytest.param(L.BF16, id="bf16")]


@pytest.mark.parametrize("dtype", _dtypes())
@pytest.mark.parametrize("wrap,halo,mode", [(0, 3, "reflect"), (10, 3, "reflect"), (0, 0, "zero"), (0, 2, "zero")])
def test_prep_input(dtype, wrap, halo, mode):
    from nirgan_b200 import _lib as L
    import helpers as Hh
    B, H, W = 2, 24, 28
    rgb, nir = _gen(B, 3, H, W, seed=1), _gen(B, 1, H, W, seed=2)
    Hb, Wb = H + 2 * wrap + 2 * halo, W + 2 * wrap + 2 * halo
    out = torch.empty(B * Hb * Wb * 16, dtype=Hh.TORCH_DT[dtype], device="cuda")
    L.call("ng_prep_input", rgb.data_ptr(), 3, nir.data_ptr(), 1, B, H, W, wrap, halo,
           L.HALO_REFLECT if mode == "reflect" else L.HALO_ZERO, 16, dtype, out.data_ptr(), Hh.stream())
    x = torch.cat((rgb, nir), 1)
    if wrap:
        x = F.pad(x, (wrap,) * 4, mode="reflect")          # Px2Px_PL.forward, pix2pix.py:91-93
    if halo:
        x = F.pad(x, (halo,) * 4, mode="reflect") if mode == "reflect" else F.pad(x, (halo,) * 4)
    got = out.view(B, Hb, Wb, 16).permute(0, 3, 1, 2).float()
    assert torch.equal(got[:, :4], Hh.rnd(x, dtype))         # bit-exact: pure index mapping + one rounding
    assert float(got[:, 4:].abs().max()) == 0.0


@pytest.mark.parametrize("dtype", _dtypes())
def test_in_stats_and_apply(dtype):
    from nirgan_b200 import _lib as L
    import helpers as Hh
    B, H, W, Cn = 2, 13, 17, 64
    y = Hh.rnd(_gen(B, Cn, H, W, seed=3, scale=2.0) + 0.5, dtype)
    yb = Hh.to_actbuf(y, 0, "zero", dtype)
    mr = torch.empty(B * Cn * 2, dtype=torch.float32, device="cuda")
    L.call("ng_in_stats", yb.t.data_ptr(), dtype, B, H * W, Cn, mr.data_ptr(), Hh.stream())
    mu, rstd = Hh.stats_ref(y)
    m = mr.view(B, Cn, 2)
    assert float((m[..., 0] - mu).abs().max()) <= 2e-6 * 4
    assert float((m[..., 1] / rstd - 1).abs().max()) <= 1e-5
    res = Hh.rnd(_gen(B, Cn, H, W, seed=4), dtype)
    tol = {L.F32: 2e-5, L.F16: 4e-3, L.BF16: 3e-2}[dtype]
    for act, slope, op, mode, use_res in [(L.ACT_RELU, 0.0, 1, "reflect", False), (L.ACT_NONE, 0.0, 1, "reflect", True),
                                          (L.ACT_RELU, 0.0, 3, "reflect", False), (L.ACT_LRELU, 0.2, 0, "zero", False),
                                          (L.ACT_NONE, 0.0, 0, "zero", True), (L.ACT_LRELU, 0.2, 2, "zero", False)]:
        rb = Hh.to_actbuf(res, 1, "reflect", dtype) if use_res else None
        out = torch.empty(B * (H + 2 * op) * (W + 2 * op) * Cn, dtype=Hh.TORCH_DT[dtype], device="cuda")
        L.call("ng_in_apply", yb.t.data_ptr(), dtype, B, H, W, Cn, mr.data_ptr(), None, None, act, slope,
               rb.t.data_ptr() if rb else None, 1, None, L.INJECT_NONE, None, out.data_ptr(), op,
               L.HALO_REFLECT if mode == "reflect" else L.HALO_ZERO, Hh.stream())
        ref = (y - mu[..., None, None]) * rstd[..., None, None]
        ref = F.relu(ref) if act == L.ACT_RELU else (F.leaky_relu(ref, slope) if act == L.ACT_LRELU else ref)
        if use_res:
            ref = ref + res
        if op:
            ref = F.pad(ref, (op,) * 4, mode="reflect") if mode == "reflect" else F.pad(ref, (op,) * 4)
        got = out.view(B, H + 2 * op, W + 2 * op, Cn).permute(0, 3, 1, 2).float()
        assert float((got - ref).abs().max()) <= tol, (act, op, mode, use_res)


@pytest.mark.parametrize("Hm", [128, 69, 21, 32])
@pytest.mark.parametrize("style", ["multiply", "add", "multiply_raw"])
def test_apply_with_satclip_injection(Hm, style):
    """x_hat * (1 + s*bilinear(e)) between IN and ReLU (generator_inject.py:113-127), fp32."""
    from nirgan_b200 import _lib as L
    import helpers as Hh
    B, Cn = 2, 128
    dtype = L.F32
    y = _gen(B, Cn, Hm, Hm, seed=5)
    e = _gen(B, 128 * 128, seed=6)
    s = torch.tensor(0.7, device="cuda")
    yb = Hh.to_actbuf(y, 0, "zero", dtype)
    mr = torch.empty(B * Cn * 2, dtype=torch.float32, device="cuda")
    L.call("ng_in_stats", yb.t.data_ptr(), dtype, B, Hm * Hm, Cn, mr.data_ptr(), Hh.stream())
    out = torch.empty(B * Hm * Hm * Cn, dtype=torch.float32, device="cuda")
    mode = {"multiply": L.INJECT_MUL_SCALED, "add": L.INJECT_ADD, "multiply_raw": L.INJECT_MUL}[style]
    L.call("ng_in_apply", yb.t.data_ptr(), dtype, B, Hm, Hm, Cn, mr.data_ptr(), None, None, L.ACT_RELU, 0.0, None, 0,
           e.data_ptr(), mode, s.data_ptr(), out.data_ptr(), 0, L.HALO_ZERO, Hh.stream())
    mu, rstd = Hh.stats_ref(y)
    xh = (y - mu[..., None, None]) * rstd[..., None, None]
    em = F.interpolate(e.view(B, 1, 128, 128), size=(Hm, Hm), mode="bilinear", align_corners=False)
    ref = {"multiply": xh * (1 + s * em), "add": xh + s * em, "multiply_raw": xh * em}[style]
    ref = F.relu(ref)
    got = out.view(B, Hm, Hm, Cn).permute(0, 3, 1, 2)
    assert float((got - ref).abs().max()) <= 5e-5


def test_linear_fc():
    from nirgan_b200 import _lib as L
    import helpers as Hh
    B, K, N = 5, 256, 128 * 128
    x, w, b = _gen(B, K, seed=7), _gen(N, K, seed=8, scale=0.02), _gen(N, seed=9, scale=0.1)
    y = torch.empty(B, N, device="cuda")
    L.call("ng_linear", x.data_ptr(), w.data_ptr(), b.data_ptr(), B, K, N, y.data_ptr(), Hh.stream())
    ref = (x.double() @ w.double().t() + b.double()).float()
    assert float((y - ref).abs().max()) <= 2e-6


def test_lsgan_and_pixel_losses_match_oracle(golden_dir):
    import numpy as np
    import nirgan_oracle as O
    from nirgan_b200.losses import lsgan, pixel_losses
    g = np.load(f"{golden_dir}/losses.npz")
    d_out = torch.from_numpy(g["d_out"]).cuda()
    assert abs(float(lsgan(d_out, 1.0)) - float(g["lsgan_real"])) <= 1e-6
    assert abs(float(lsgan(d_out, 0.0)) - float(g["lsgan_fake"])) <= 1e-6
    rgb, nir, pred = (torch.from_numpy(g[k]).cuda() for k in ("rgb", "nir", "pred"))
    out = pixel_losses(rgb, nir, pred, (1.0, 1.0, 1.0, 1.0))
    for i, k in enumerate(("l1", "ndvi", "ndwi", "evi")):
        ref = float(g[k])
        assert abs(float(out[i]) - ref) <= 2e-5 * max(1.0, abs(ref)), k   # golden values come from the reference
    # gradients against torch autograd on a well-conditioned input (denominators away from 0)
    rgb2, nir2 = rgb + 1.0, nir
    pc = pred.clone().requires_grad_(True)
    w = (100.0, 0.33, 0.33, 0.33)
    parts = pixel_losses(rgb2, nir2, pc, w)
    sum(wi * pi for wi, pi in zip(w, parts)).backward()
    po = pred.clone().requires_grad_(True)
    ref_loss = 100.0 * (po - nir2).abs().mean() + O.rs_weighted_loss(
        rgb2, nir2, po, {"lambda_ndvi": 0.33, "lambda_ndwi": 0.33, "lambda_evi": 0.33})
    ref_loss.backward()
    rel = float((pc.grad - po.grad).norm() / po.grad.norm())
    assert rel <= 1e-5, rel
    pl = lsgan(d_out.clone().requires_grad_(True), 1.0)
    dd = d_out.clone().requires_grad_(True)
    pl2 = lsgan(dd, 1.0)
    pl2.backward()
    assert float((dd.grad - 2 * (d_out - 1.0) / d_out.numel()).abs().max()) <= 1e-7


@pytest.mark.parametrize("criterion", ["l1", "l2"])
def test_all_six_indices_and_modes_match_oracle(criterion):
    """RemoteSensingIndices: every index, both criteria, 'loss' / 'logging_dict' / 'index' modes, and d/dpred of the
    weighted sum against torch autograd of the oracle (utils/remote_sensing_indices.py:23-319, pinned in
    tests/golden/PIN_REPORT.txt as rs_all6_rel)."""
    import nirgan_oracle as O
    from nirgan_b200.utils.remote_sensing_indices import RemoteSensingIndices
    g = torch.Generator().manual_seed(21)
    rgb = (0.05 + 0.6 * torch.rand(3, 3, 40, 56, generator=g)).cuda()      # reflectance-like, denominators away from 0
    nir = (0.1 + 0.7 * torch.rand(3, 1, 40, 56, generator=g)).cuda()
    pred = (0.1 + 0.7 * torch.rand(3, 1, 40, 56, generator=g)).cuda()
    cfg = {"lambda_ndvi": 0.3, "lambda_ndwi": 0.2, "lambda_gndvi": 0.15, "lambda_savi": 0.1, "lambda_msavi": 0.25,
           "lambda_evi": 0.4}
    rs = RemoteSensingIndices(mode="loss", criterion=criterion)
    pc = pred.clone().requires_grad_(True)
    total = rs.get_and_weight_losses(rgb, nir, pc, cfg)
    total.backward()
    po = pred.clone().requires_grad_(True)
    ref = O.rs_weighted_loss(rgb, nir, po, cfg, criterion)
    ref.backward()
    assert abs(float(total) - float(ref)) <= 2e-5 * max(1.0, abs(float(ref)))
    assert float((pc.grad - po.grad).norm() / po.grad.norm()) <= 2e-5
    # a zero / negative weight switches the index off, as `if weight > 0.0` does in the reference
    cfg2 = dict(cfg, lambda_savi=0.0, lambda_msavi=-1.0)
    assert abs(float(rs.get_and_weight_losses(rgb, nir, pred, cfg2)) -
               float(O.rs_weighted_loss(rgb, nir, pred, cfg2, criterion))) <= 2e-5
    # logging_dict and the single-index methods
    log = rs.get_and_weight_losses(rgb, nir, pred, mode="logging_dict")
    pairs = {"ndvi": O.ndvi_pair, "ndwi": O.ndwi_pair, "gndvi": O.gndvi_pair, "savi": O.savi_pair, "msavi": O.msavi_pair,
             "evi": O.evi_pair}
    for name, fn in pairs.items():
        a, b = fn(rgb, nir, pred)
        want = float(O._crit(a, b, criterion))
        assert abs(float(log[f"indices_loss/{name}_error"]) - want) <= 2e-5 * max(1.0, abs(want)), name
        assert abs(float(getattr(rs, name + "_calculation")(rgb, nir, pred)) - want) <= 2e-5 * max(1.0, abs(want)), name
    # index mode: the maps themselves, without the loss-mode epsilons; 3-D inputs are promoted like the reference does
    ri = RemoteSensingIndices(mode="index", criterion=criterion)
    for name, fn in pairs.items():
        kw = {"eps": 0.0} if name in ("ndvi", "ndwi") else ({"loss_mode": False} if name == "evi" else {})
        a, b = fn(rgb, nir, pred, **kw)
        ga, gb = getattr(ri, name + "_calculation")(rgb, nir, pred)
        assert float((ga - a).abs().max()) <= 1e-5 * max(1.0, float(a.abs().max())), name
        assert float((gb - b).abs().max()) <= 1e-5 * max(1.0, float(b.abs().max())), name
    ga, _ = ri.ndvi_calculation(rgb[0], nir[0], pred[0])
    assert ga.shape == (1, 1, 40, 56)
    with pytest.raises(NotImplementedError, match="logging_dict"):
        rs.get_and_weight_losses(rgb, nir, pred, mode="nope")


def test_adam_matches_torch():
    from nirgan_b200 import _lib as L
    import helpers as Hh
    n = 100_003
    p = _gen(n, seed=10)
    ref_p = p.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref_p], lr=2e-4, betas=(0.5, 0.999))
    m, v = torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    for step in range(1, 4):
        g = _gen(n, seed=20 + step, scale=1e-3)
        ref_p.grad = g.clone()
        opt.step()
        L.call("ng_adam_step", p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), n, 2e-4, 0.5, 0.999, 1e-8,
               step, 1.0, Hh.stream())
        assert float((p - ref_p.detach()).abs().max()) <= 5e-7      # 2 ulp at |p| ~ 2


@pytest.mark.parametrize("kind", ["res", "down", "convT", "dk4s2", "head", "dout"])
def test_wgrad_matches_autograd(kind):
    import ctypes as C
    from nirgan_b200 import _lib as L
    import helpers as Hh
    dtype = L.F32
    B = 2
    if kind == "res":
        Cin, Cout, K, s, p, H, mode, form = 256, 256, 3, 1, 1, 12, "reflect", L.FORM_GATHER
    elif kind == "down":
        Cin, Cout, K, s, p, H, mode, form = 64, 128, 3, 2, 1, 18, "zero", L.FORM_GATHER
    elif kind == "dk4s2":
        Cin, Cout, K, s, p, H, mode, form = 64, 128, 4, 2, 1, 16, "zero", L.FORM_GATHER
    elif kind == "head":
        Cin, Cout, K, s, p, H, mode, form = 64, 1, 7, 1, 3, 20, "reflect", L.FORM_GATHER
    elif kind == "dout":
        Cin, Cout, K, s, p, H, mode, form = 512, 1, 4, 1, 1, 13, "zero", L.FORM_GATHER
    else:
        Cin, Cout, K, s, p, H, mode, form = 128, 64, 3, 2, 1, 9, "zero", L.FORM_PHASED
    x = _gen(B, Cin, H, H, seed=11)
    if form == L.FORM_PHASED:
        w = _gen(Cin, Cout, K, K, seed=12, scale=0.05).requires_grad_(True)
        out = F.conv_transpose2d(x, w, stride=2, padding=1, output_padding=1)
        Ho = 2 * H
    else:
        w = _gen(Cout, Cin, K, K, seed=12, scale=0.05).requires_grad_(True)
        xr = F.pad(x, (p,) * 4, mode="reflect") if mode == "reflect" else F.pad(x, (p,) * 4)
        out = F.conv2d(xr, w, stride=s)
        Ho = out.shape[-1]
    dy = _gen(*out.shape, seed=13)
    out.backward(dy)
    co_pad = 16 if kind in ("head", "dout") else Cout
    xb = Hh.to_actbuf(x, p if mode == "reflect" else 0, mode, dtype)
    dyb = Hh.to_actbuf(dy, 0, "zero", dtype, c_pad=co_pad)
    a = L.ConvArgs()
    a.dtype, a.impl, a.form, a.sgn = dtype, L.IMPL_SIMT, form, 1
    a.B, a.Hin, a.Win, a.Cin, a.in_pad, a.in_pad_w = B, H, H, Cin, xb.pad, xb.pad
    a.Cout, a.KH, a.KW, a.stride, a.pad, a.pad_w, a.Hout, a.Wout = co_pad, K, K, s, p, p, Ho, Ho
    a.x, a.w, a.y = xb.t.data_ptr(), xb.t.data_ptr(), dyb.t.data_ptr()
    if kind in ("head", "dout"):
        a.epilogue = L.EPI_HEAD        # single real output channel -> dedicated kernel
    dwp = torch.empty(K * K * co_pad * Cin, device="cuda")
    db = torch.empty(co_pad, device="cuda")
    L.call("ng_conv2d_wgrad", C.byref(a), dwp.data_ptr(), db.data_ptr(), None, 0, Hh.stream())
    dw = torch.empty_like(w)
    n_axis = 1 if form == L.FORM_PHASED else 0
    L.call("ng_unpack_weight_grad", dwp.data_ptr(), w.shape[0], w.shape[1], K, K, n_axis, co_pad, Cin, 1.0, None, 0.0,
           dw.data_ptr(), Hh.stream())
    rel = float((dw - w.grad).norm() / w.grad.norm())
    assert rel <= 2e-5, rel
    assert float((db[:Cout] - dy.sum(dim=(0, 2, 3))).abs().max()) <= 1e-3


@pytest.mark.parametrize("dt", ["f16", "bf16"])
@pytest.mark.parametrize("Cin,K,H,B", [(512, 4, 13, 2), (512, 4, 31, 5), (64, 3, 20, 3), (256, 4, 9, 1)])
def test_wgrad_single_output_channel_reads_input_once(Cin, K, H, B, dt):
    """Weight gradient of a stride-1 convolution with ONE real output channel (PatchGAN last layer: 512 -> 1, k4, p1) in
    the single-read form -- every position of the haloed input buffer visited once, all taps accumulated in registers,
    per-block partials reduced in block order -- against autograd on the same 16-bit-rounded operands; bit-identical
    from launch to launch (no atomics)."""
    import ctypes as C
    from nirgan_b200 import _lib as L
    import helpers as Hh
    dtype = L.F16 if dt == "f16" else L.BF16
    tdt = torch.float16 if dt == "f16" else torch.bfloat16
    p = 1
    x = _gen(B, Cin, H, H, seed=11).to(tdt).float()
    w = _gen(1, Cin, K, K, seed=12, scale=0.05).requires_grad_(True)
    out = F.conv2d(F.pad(x, (p,) * 4), w)
    Ho = out.shape[-1]
    dy = _gen(*out.shape, seed=13).to(tdt).float()
    out.backward(dy)
    xb = Hh.to_actbuf(x, p, "zero", dtype)                       # halo materialised (zeros), as the PatchGAN buffers are
    dyb = Hh.to_actbuf(dy, 0, "zero", dtype, c_pad=16)
    a = L.ConvArgs()
    a.dtype, a.impl, a.form, a.sgn = dtype, L.IMPL_TC, L.FORM_GATHER, 1
    a.B, a.Hin, a.Win, a.Cin, a.in_pad, a.in_pad_w = B, H, H, Cin, xb.pad, xb.pad
    a.Cout, a.KH, a.KW, a.stride, a.pad, a.pad_w, a.Hout, a.Wout = 16, K, K, 1, p, p, Ho, Ho
    a.x, a.w, a.y = xb.t.data_ptr(), xb.t.data_ptr(), dyb.t.data_ptr()
    a.epilogue = L.EPI_HEAD
    need = int(L.load().ng_conv2d_wgrad_workspace_bytes(C.byref(a)))
    assert need > 0                                              # the single-read form asks for its partial slots
    ws = torch.full((need // 4,), float("nan"), device="cuda")
    outs = []
    for _ in range(2):
        dwp = torch.full((K * K * 16 * Cin,), float("nan"), device="cuda")
        L.call("ng_conv2d_wgrad", C.byref(a), dwp.data_ptr(), None, ws.data_ptr(), need, Hh.stream())
        torch.cuda.synchronize()
        outs.append(dwp)
    assert torch.equal(outs[0], outs[1])
    dw = torch.empty_like(w)
    L.call("ng_unpack_weight_grad", outs[0].data_ptr(), 1, Cin, K, K, 0, 16, Cin, 1.0, None, 0.0, dw.data_ptr(), Hh.stream())
    rel = float((dw - w.grad).norm() / w.grad.norm())
    assert rel <= 1e-5, rel
    assert float(outs[0].view(K * K, 16, Cin)[:, 1:].abs().max()) == 0.0       # padding rows of the packed gradient


# tcgen05 split-K weight gradient (MN-major operands) against torch autograd on the same 16-bit-rounded operands.
# Geometries: ResnetBlock 3x3 (reflect halo), strided down conv, ConvTranspose (phased), PatchGAN k4 s2 and k4 s1
# (Cout 512 -> four n tiles), Cout 64 (upper half of the 128-row tile is out-of-bounds zero fill), an image smaller
# than one 64-pixel patch, and a batch large enough for several pixel splits.
@pytest.mark.parametrize("dt", ["f16", "bf16"])
@pytest.mark.parametrize("kind", ["res", "res69", "res128", "down", "down128", "convT", "convT256", "dk4s2", "dk4s1", "tiny",
                                  "l0"])
def test_wgrad_tc_matches_autograd(kind, dt):
    import ctypes as C
    from nirgan_b200 import _lib as L
    import helpers as Hh
    dtype = L.F16 if dt == "f16" else L.BF16
    tdt = torch.float16 if dt == "f16" else torch.bfloat16
    B = 2
    cfgs = {
        "res": (256, 256, 3, 1, 1, 16, "reflect", L.FORM_GATHER),
        "res69": (256, 256, 3, 1, 1, 69, "reflect", L.FORM_GATHER),      # row-patch form, ragged 8 x 8 patches, 3 images
        "res128": (128, 128, 3, 1, 1, 20, "zero", L.FORM_GATHER),        # row-patch form, one k tile, zero padding
        "down": (64, 128, 3, 2, 1, 36, "zero", L.FORM_GATHER),
        "down128": (128, 256, 3, 2, 1, 34, "zero", L.FORM_GATHER),
        "convT": (128, 64, 3, 2, 1, 17, "zero", L.FORM_PHASED),
        "convT256": (256, 128, 3, 2, 1, 9, "zero", L.FORM_PHASED),
        "dk4s2": (64, 128, 4, 2, 1, 32, "zero", L.FORM_GATHER),
        "dk4s1": (256, 512, 4, 1, 1, 8, "zero", L.FORM_GATHER),
        "tiny": (64, 64, 3, 1, 1, 5, "reflect", L.FORM_GATHER),
        "l0": (16, 64, 4, 2, 1, 40, "zero", L.FORM_GATHER),          # PatchGAN input layer (16 stored channels)
    }
    Cin, Cout, K, s, p, H, mode, form = cfgs[kind]
    if kind == "res69":
        B = 3
    x = _gen(B, Cin, H, H, seed=11).to(tdt).float()
    if form == L.FORM_PHASED:
        w = _gen(Cin, Cout, K, K, seed=12, scale=0.05).requires_grad_(True)
        out = F.conv_transpose2d(x, w, stride=2, padding=1, output_padding=1)
        Ho = 2 * H
    else:
        w = _gen(Cout, Cin, K, K, seed=12, scale=0.05).requires_grad_(True)
        xr = F.pad(x, (p,) * 4, mode="reflect") if mode == "reflect" else F.pad(x, (p,) * 4)
        out = F.conv2d(xr, w, stride=s)
        Ho = out.shape[-1]
    dy = _gen(*out.shape, seed=13).to(tdt).float()
    out.backward(dy)
    xb = Hh.to_actbuf(x, p if mode == "reflect" else 0, mode, dtype)
    dyb = Hh.to_actbuf(dy, 0, "zero", dtype)
    a = L.ConvArgs()
    a.dtype, a.impl, a.form, a.sgn = dtype, L.IMPL_TC, form, 1
    a.B, a.Hin, a.Win, a.Cin, a.in_pad, a.in_pad_w = B, H, H, Cin, xb.pad, xb.pad
    a.Cout, a.KH, a.KW, a.stride, a.pad, a.pad_w, a.Hout, a.Wout = Cout, K, K, s, p, p, Ho, Ho
    a.x, a.w, a.y = xb.t.data_ptr(), xb.t.data_ptr(), dyb.t.data_ptr()
    need = L.load().ng_conv2d_wgrad_workspace_bytes(C.byref(a))
    assert need > 0
    ws = torch.full((need // 4,), float("nan"), device="cuda")
    dwp = torch.full((K * K * Cout * Cin,), float("nan"), device="cuda")
    db = torch.empty(Cout, device="cuda")
    L.call("ng_conv2d_wgrad", C.byref(a), dwp.data_ptr(), db.data_ptr(), ws.data_ptr(), need, Hh.stream())
    dw = torch.empty_like(w)
    n_axis = 1 if form == L.FORM_PHASED else 0
    dw.fill_(float("nan"))           # beta = 0 must not read the destination
    L.call("ng_unpack_weight_grad", dwp.data_ptr(), w.shape[0], w.shape[1], K, K, n_axis, Cout, Cin, 1.0, None, 0.0,
           dw.data_ptr(), Hh.stream())
    assert bool(torch.isfinite(dw).all())
    dw2 = dw.clone()
    L.call("ng_unpack_weight_grad", dwp.data_ptr(), w.shape[0], w.shape[1], K, K, n_axis, Cout, Cin, 1.0, None, 1.0,
           dw2.data_ptr(), Hh.stream())
    assert torch.equal(dw2, dw + dw)
    rel = float((dw - w.grad).norm() / w.grad.norm())
    assert rel <= 1e-5, rel          # identical 16-bit operands, fp32 accumulation on both sides
    assert float((db - dy.sum(dim=(0, 2, 3))).abs().max()) <= 1e-3
    # deterministic: a second run gives bit-identical output
    dwp2 = torch.empty_like(dwp)
    L.call("ng_conv2d_wgrad", C.byref(a), dwp2.data_ptr(), None, ws.data_ptr(), need, Hh.stream())
    assert torch.equal(dwp, dwp2)


@pytest.mark.parametrize("H,W,B", [(24, 40, 2), (69, 69, 3), (8, 16, 1)])
def test_wgrad_tc_row_patch_stem(H, W, B):
    """Weight gradient of the row-merged generator stem (7 x 1 taps over 64 stored channels, Cout 64): the tcgen05 kernel's
    row-patch form (one haloed 14 x 8 pixel X patch per stage, every tap a window into it) against torch autograd of the
    same 7 x 1 convolution, incl. partially covered patches at the right / bottom border."""
    import ctypes as C
    from nirgan_b200 import _lib as L
    import helpers as Hh
    dtype, tdt = L.F16, torch.float16
    x = _gen(B, 64, H, W, seed=41).to(tdt).float()
    w = _gen(64, 64, 7, 1, seed=42, scale=0.05).requires_grad_(True)
    xr = F.pad(x, (0, 0, 3, 3))                                   # rows only: the column taps are merged into channels
    out = F.conv2d(xr, w)
    dy = _gen(*out.shape, seed=43).to(tdt).float()
    out.backward(dy)
    xb_t = xr.permute(0, 2, 3, 1).contiguous().to(tdt).reshape(-1)   # [B][H+6][W][64]: halo rows materialised
    dyb = Hh.to_actbuf(dy, 0, "zero", dtype)
    a = L.ConvArgs()
    a.dtype, a.impl, a.form, a.sgn = dtype, L.IMPL_TC, L.FORM_GATHER, 1
    a.B, a.Hin, a.Win, a.Cin, a.in_pad, a.in_pad_w = B, H, W, 64, 3, 0
    a.Cout, a.KH, a.KW, a.stride, a.pad, a.pad_w, a.Hout, a.Wout = 64, 7, 1, 1, 3, 0, H, W
    a.x, a.w, a.y = xb_t.data_ptr(), xb_t.data_ptr(), dyb.t.data_ptr()
    need = L.load().ng_conv2d_wgrad_workspace_bytes(C.byref(a))
    assert need > 0
    ws = torch.full((need // 4,), float("nan"), device="cuda")
    dwp = torch.full((7 * 64 * 64,), float("nan"), device="cuda")
    L.call("ng_conv2d_wgrad", C.byref(a), dwp.data_ptr(), None, ws.data_ptr(), need, Hh.stream())
    dw = torch.empty_like(w)
    L.call("ng_unpack_weight_grad", dwp.data_ptr(), 64, 64, 7, 1, 0, 64, 64, 1.0, None, 0.0, dw.data_ptr(), Hh.stream())
    torch.cuda.synchronize()
    assert bool(torch.isfinite(dw).all())
    rel = float((dw - w.grad).norm() / w.grad.norm())
    assert rel <= 1e-5, rel
