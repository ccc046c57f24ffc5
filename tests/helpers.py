"""Test helpers: drive single C-ABI kernels from torch tensors and restate them in torch fp32."""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn.functional as F

import nirgan_b200  # noqa: F401  (alias of the nir-gan_b200 package)
from nirgan_b200 import _lib as L
from nirgan_b200.engine import ActBuf

TORCH_DT = {L.F32: torch.float32, L.F16: torch.float16, L.BF16: torch.bfloat16}


def stream():
    return torch.cuda.current_stream().cuda_stream


def rup(x, m):
    return (x + m - 1) // m * m


def to_actbuf(x_nchw: torch.Tensor, halo: int, mode: str, dtype: int, c_pad: int = 0) -> ActBuf:
    """NCHW fp32 -> haloed NHWC ActBuf (torch ops; test-side only)."""
    B, Cn, H, W = x_nchw.shape
    x = x_nchw
    if halo:
        x = F.pad(x, (halo,) * 4, mode="reflect") if mode == "reflect" else F.pad(x, (halo,) * 4)
    cp = c_pad or Cn
    if cp > Cn:
        x = torch.cat([x, x.new_zeros(B, cp - Cn, x.shape[2], x.shape[3])], 1)
    t = x.permute(0, 2, 3, 1).contiguous().to(TORCH_DT[dtype]).reshape(-1)
    return ActBuf(t, B, H, W, cp, halo)


def from_compact(t: torch.Tensor, B, H, W, Cn) -> torch.Tensor:
    return t.view(B, H, W, Cn).permute(0, 3, 1, 2).float()


def pack_weight(w: torch.Tensor, n_axis: int, n_pad: int, k_pad: int, dtype: int) -> torch.Tensor:
    d0, d1, kh, kw = w.shape
    dst = torch.empty(kh * kw * n_pad * k_pad, dtype=TORCH_DT[dtype], device=w.device)
    L.call("ng_pack_weight", w.contiguous().data_ptr(), d0, d1, kh, kw, n_axis, n_pad, k_pad, dtype, dst.data_ptr(),
           stream())
    return dst


def conv_call(x: ActBuf, wp: torch.Tensor, Cout: int, K: int, stride: int, pad: int, Hout: int, Wout: int, dtype: int,
              impl: int, form=L.FORM_GATHER, sgn=1, epilogue=L.EPI_RAW, act=L.ACT_NONE, slope=0.0, crop=0, bias=None,
              want_stats=False):
    """Returns (y tensor, mean_rstd or None)."""
    dev = x.t.device
    if epilogue == L.EPI_HEAD:
        y = torch.full((x.B * (Hout - 2 * crop) * (Wout - 2 * crop),), float("nan"), dtype=torch.float32, device=dev)
    else:
        y = torch.full((x.B * Hout * Wout * Cout,), float("nan"), dtype=torch.float32, device=dev).to(TORCH_DT[dtype])
    a = L.ConvArgs()
    a.dtype, a.impl, a.form, a.sgn = dtype, impl, form, sgn
    a.B, a.Hin, a.Win, a.Cin, a.in_pad, a.in_pad_w = x.B, x.H, x.W, x.C, x.pad, x.pad
    a.Cout, a.KH, a.KW, a.stride, a.pad, a.pad_w = Cout, K, K, stride, pad, pad
    a.Hout, a.Wout = Hout, Wout
    a.epilogue, a.act, a.slope, a.crop = epilogue, act, slope, crop
    a.x, a.w, a.y = x.t.data_ptr(), wp.data_ptr(), y.data_ptr()
    a.bias = bias.data_ptr() if bias is not None else None
    mr = None
    part = None
    if want_stats and impl == L.IMPL_TC:
        slots = L.load().ng_conv_stat_slots(C.byref(a))
        assert slots > 0, L.last_error()
        part = torch.full((x.B * slots * Cout * 2,), float("nan"), dtype=torch.float32, device=dev)
        a.stat_partials = part.data_ptr()
    fused_mr = cnt = None
    if want_stats and impl == L.IMPL_TC:
        # also run the fused finalisation (last CTA per image) and check it against the stand-alone kernel below
        fused_mr = torch.full((x.B * Cout * 2,), float("nan"), dtype=torch.float32, device=dev)
        cnt = torch.zeros(x.B, dtype=torch.int32, device=dev)
        a.mean_rstd, a.tile_counters = fused_mr.data_ptr(), cnt.data_ptr()
    L.call("ng_conv2d", C.byref(a), stream())
    if want_stats:
        mr = torch.empty(x.B * Cout * 2, dtype=torch.float32, device=dev)
        if impl == L.IMPL_TC:
            L.call("ng_in_stats_finalize", part.data_ptr(), x.B, slots, Cout, Hout * Wout, mr.data_ptr(), stream())
            # the product path: fixed-point accumulators filled by integer atomics, turned into (mean, rstd) by the apply
            # kernel itself (mean_rstd_out); must agree with the per-tile partials + finalize form, and be bit-identical
            # from launch to launch whatever order the tiles finish in
            accs = []
            for _ in range(2):
                b = L.ConvArgs()
                C.memmove(C.byref(b), C.byref(a), C.sizeof(L.ConvArgs))
                b.stat_partials, b.mean_rstd, b.tile_counters = None, None, None
                acc = torch.full((x.B * Cout * 2,), 7, dtype=torch.int64, device=dev)
                L.call("ng_memset_zero", acc.data_ptr(), acc.numel() * 8, stream())
                b.stat_acc = acc.data_ptr()
                y2 = torch.empty_like(y)
                b.y = y2.data_ptr()
                L.call("ng_conv2d", C.byref(b), stream())
                accs.append(acc)
            torch.cuda.synchronize()
            assert torch.equal(accs[0], accs[1]), "fixed-point statistics must not depend on tile completion order"
            assert torch.equal(y2, y), "image-minor tile order must not change the convolution output"
            mr2 = torch.full((x.B * Cout * 2,), float("nan"), dtype=torch.float32, device=dev)
            o2 = torch.empty(x.B * Hout * Wout * Cout, dtype=TORCH_DT[dtype], device=dev)
            L.call("ng_in_apply", y.data_ptr(), dtype, x.B, Hout, Wout, Cout, None, accs[0].data_ptr(), mr2.data_ptr(),
                   L.ACT_NONE, 0.0, None, 0, None, L.INJECT_NONE, None, o2.data_ptr(), 0, L.HALO_ZERO, stream())
            torch.cuda.synchronize()
            m1, m2 = mr.view(-1, 2), mr2.view(-1, 2)
            assert float((m1[:, 0] - m2[:, 0]).abs().max()) <= 1e-5 * max(1.0, float(m1[:, 0].abs().max())), "acc mean"
            assert float((m1[:, 1] / m2[:, 1] - 1).abs().max()) <= 1e-4, "acc rstd"
            torch.cuda.synchronize()
            assert int(cnt.abs().max()) == 0, "tile counters must be left at zero"
            assert torch.isfinite(fused_mr).all()
            assert float((fused_mr - mr).abs().max()) <= 1e-5 * max(1.0, float(mr.abs().max())), "fused finalisation"
            L.call("ng_conv2d", C.byref(a), stream())          # second launch on the same counters
            torch.cuda.synchronize()
            assert float((fused_mr - mr).abs().max()) <= 1e-5 * max(1.0, float(mr.abs().max()))
        else:
            L.call("ng_in_stats", y.data_ptr(), dtype, x.B, Hout * Wout, Cout, mr.data_ptr(), stream())
        mr = mr.view(x.B, Cout, 2)
    torch.cuda.synchronize()
    return y, mr, a


def rnd(x: torch.Tensor, dtype: int) -> torch.Tensor:
    """Round fp32 values to the kernel's operand precision (so the fp32 torch restatement sees the same operands)."""
    return x.to(TORCH_DT[dtype]).float()


def stats_ref(y_nchw: torch.Tensor):
    mu = y_nchw.mean(dim=(2, 3))
    var = y_nchw.var(dim=(2, 3), unbiased=False)
    return mu, torch.rsqrt(var + 1e-5)
