#!/usr/bin/env python
"""Small launches of every hand-written mbarrier / TMEM pipeline for `compute-sanitizer --tool racecheck|synccheck`
(one tool per gpurun call): the tcgen05 convolution in its variants (K-deep 256-wide, 64-wide with four epilogue groups,
row-tap stem, direct stem with the software producer warps, merged-phase ConvTranspose, strided, 16-channel head), the
tcgen05 weight gradient (plain, tap groups, row patch, tap-inner) and the fused apply / norm-backward kernels.  Prints the
max-abs deviation of each result from torch so that a sanitizer-clean run is also a correct one."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402
import nirgan_b200  # noqa: F401,E402
from nirgan_b200 import _lib as L  # noqa: E402
import helpers as Hh  # noqa: E402


def gen(*shape, seed=0, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return torch.randn(*shape, generator=g, device="cuda") * scale


def conv_case(name, Cin, Cout, K, s, p, H, mode, form=L.FORM_GATHER, B=1, stats=True):
    dtype = L.F16
    x = Hh.rnd(gen(B, Cin, H, H, seed=1), dtype)
    if form == L.FORM_GATHER:
        w = Hh.rnd(gen(Cout, Cin, K, K, seed=2, scale=0.05), dtype)
        xr = F.pad(x, (p,) * 4, mode="reflect") if mode == "reflect" else F.pad(x, (p,) * 4)
        ref = F.conv2d(xr, w, stride=s)
        xb = Hh.to_actbuf(x, p if mode == "reflect" else 0, mode, dtype)
        wp = Hh.pack_weight(w, 0, Cout, Cin, dtype)
        y, mr, _ = Hh.conv_call(xb, wp, Cout, K, s, p, ref.shape[-2], ref.shape[-1], dtype, L.IMPL_TC, want_stats=stats)
    else:
        w = Hh.rnd(gen(Cin, Cout, K, K, seed=2, scale=0.05), dtype)
        ref = F.conv_transpose2d(x, w, stride=2, padding=1, output_padding=1)
        xb = Hh.to_actbuf(x, 0, "zero", dtype)
        wp = torch.empty(16 * Cout * Cin, dtype=torch.float16, device="cuda")
        L.call("ng_pack_weight_phasemerged", w.contiguous().data_ptr(), Cin, Cout, dtype, wp.data_ptr(), Hh.stream())
        y, mr, _ = Hh.conv_call(xb, wp, Cout, K, 2, 1, 2 * H, 2 * H, dtype, L.IMPL_TC, form=form, want_stats=stats)
    got = Hh.from_compact(y, B, ref.shape[-2], ref.shape[-1], Cout)
    print(f"{name:28s} max-abs dev {float((got - ref).abs().max()):.3e} (|ref| max {float(ref.abs().max()):.2f})", flush=True)


def main():
    torch.cuda.set_device(0)
    conv_case("conv_tc<256,64> res 3x3", 256, 256, 3, 1, 1, 16, "reflect")
    conv_case("conv_tc<128,64> down s2", 64, 128, 3, 2, 1, 32, "zero")
    conv_case("conv_tc<64,64,EG4> 1x1", 64, 64, 1, 1, 0, 24, "zero", stats=False)
    conv_case("conv_tc merged-phase convT", 128, 64, 3, 2, 1, 16, "zero", form=L.FORM_PHASED_MERGED)
    # direct stem (software producer warps + SW64 operands)
    dtype = L.F16
    x = gen(2, 3, 24, 36, seed=5)
    w = Hh.rnd(gen(64, 3, 7, 7, seed=6, scale=0.05), dtype)
    wp = torch.empty(7 * 64 * 32, dtype=torch.float16, device="cuda")
    L.call("ng_pack_weight_rowmerged", w.data_ptr(), 64, 3, 7, 7, 4, dtype, wp.data_ptr(), Hh.stream())
    H1, W1 = 44, 56
    y = torch.empty(2 * H1 * W1 * 64, dtype=torch.float16, device="cuda")
    acc = torch.zeros(2 * 64 * 2, dtype=torch.int64, device="cuda")
    L.call("ng_stem_conv", x.data_ptr(), 3, 2, 24, 36, 10, wp.data_ptr(), dtype, y.data_ptr(), None, acc.data_ptr(), Hh.stream())
    ref = F.conv2d(F.pad(F.pad(Hh.rnd(x, dtype), (10,) * 4, mode="reflect"), (3,) * 4, mode="reflect"), w)
    print(f"{'ng_stem_conv (direct stem)':28s} max-abs dev {float((Hh.from_compact(y, 2, H1, W1, 64) - ref).abs().max()):.3e}", flush=True)
    # weight gradients: tap groups (64 ch, 3x3), pairs (128 ch), plain (256 ch)
    import ctypes as C
    for name, Cin, Cout, K, s, H in (("wgrad_tc<64,3> groups", 64, 128, 3, 2, 16), ("wgrad_tc<128,2> pairs", 128, 256, 3, 2, 16),
                                      ("wgrad_tc<256,1>", 256, 256, 3, 1, 12)):
        xx = gen(2, Cin, H, H, seed=7).half().float()
        ww = gen(Cout, Cin, K, K, seed=8, scale=0.05).requires_grad_(True)
        mode = "reflect" if s == 1 else "zero"
        xr = F.pad(xx, (1,) * 4, mode="reflect") if s == 1 else F.pad(xx, (1,) * 4)
        out = F.conv2d(xr, ww, stride=s)
        dy = gen(*out.shape, seed=9).half().float()
        out.backward(dy)
        xb = Hh.to_actbuf(xx, 1 if s == 1 else 0, mode, dtype)
        dyb = Hh.to_actbuf(dy, 0, "zero", dtype)
        a = L.ConvArgs()
        a.dtype, a.impl, a.form, a.sgn = dtype, L.IMPL_TC, L.FORM_GATHER, 1
        a.B, a.Hin, a.Win, a.Cin, a.in_pad, a.in_pad_w = 2, H, H, Cin, xb.pad, xb.pad
        a.Cout, a.KH, a.KW, a.stride, a.pad, a.pad_w, a.Hout, a.Wout = Cout, K, K, s, 1, 1, out.shape[-1], out.shape[-1]
        a.x, a.w, a.y = xb.t.data_ptr(), xb.t.data_ptr(), dyb.t.data_ptr()
        need = L.load().ng_conv2d_wgrad_workspace_bytes(C.byref(a))
        ws = torch.empty(max(need // 4, 4), device="cuda")
        dwp = torch.empty(K * K * Cout * Cin, device="cuda")
        L.call("ng_conv2d_wgrad", C.byref(a), dwp.data_ptr(), None, ws.data_ptr(), need, Hh.stream())
        dw = torch.empty_like(ww)
        L.call("ng_unpack_weight_grad", dwp.data_ptr(), Cout, Cin, K, K, 0, Cout, Cin, 1.0, None, 0.0, dw.data_ptr(), Hh.stream())
        torch.cuda.synchronize()
        print(f"{name:28s} rel-L2 {float((dw - ww.grad).norm() / ww.grad.norm()):.3e}", flush=True)
    torch.cuda.synchronize()
    print("sanitize targets done")


if __name__ == "__main__":
    main()
