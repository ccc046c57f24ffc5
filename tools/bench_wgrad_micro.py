#!/usr/bin/env python
"""Times ng_conv2d_wgrad alone on the thin layers of the B=32 training step (CUDA events, 20 launches each), with the
bytes both operands occupy and the arithmetic, to see which floor each one is near.  Diagnostic tool."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import nirgan_b200  # noqa: F401,E402
from nirgan_b200 import _lib as L  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    B = int(os.environ.get("B", "32"))
    only = os.environ.get("ONLY", "")
    # name: Cin, Cout, K, stride, pad, Hin, halo (buffer pad), form
    cfgs = {
        "head_tapgemm": (64, 64, 1, 1, 0, 282, 0, L.FORM_GATHER),
        "d1": (64, 128, 3, 2, 1, 276, 1, L.FORM_GATHER),
        "d2": (128, 256, 3, 2, 1, 138, 1, L.FORM_GATHER),
        "u1": (256, 128, 3, 2, 1, 69, 0, L.FORM_PHASED),
        "u2": (128, 64, 3, 2, 1, 138, 0, L.FORM_PHASED),
        "res": (256, 256, 3, 1, 1, 69, 1, L.FORM_GATHER),
        "D_l0": (16, 64, 4, 2, 1, 256, 1, L.FORM_GATHER),
        "D_l1": (64, 128, 4, 2, 1, 128, 1, L.FORM_GATHER),
        "D_l2": (128, 256, 4, 2, 1, 64, 1, L.FORM_GATHER),
        "D_l3": (256, 512, 4, 1, 1, 32, 1, L.FORM_GATHER),
    }
    stream = torch.cuda.current_stream(dev).cuda_stream
    out = {}
    for name, (Cin, Cout, K, s, p, H, halo, form) in cfgs.items():
        if only and name not in only.split(","):
            continue
        Ho = 2 * H if form == L.FORM_PHASED else (H + 2 * p - K) // s + 1
        x = (torch.randn(B, H + 2 * halo, H + 2 * halo, Cin, device=dev) * 0.5).half()
        dy = (torch.randn(B, Ho, Ho, Cout, device=dev) * 0.5).half()
        a = L.ConvArgs()
        a.dtype, a.impl, a.form, a.sgn = L.F16, L.IMPL_TC, form, 1
        a.B, a.Hin, a.Win, a.Cin, a.in_pad, a.in_pad_w = B, H, H, Cin, halo, halo
        a.Cout, a.KH, a.KW, a.stride, a.pad, a.pad_w, a.Hout, a.Wout = Cout, K, K, s, p, p, Ho, Ho
        a.x, a.w, a.y = x.data_ptr(), x.data_ptr(), dy.data_ptr()
        need = int(L.load().ng_conv2d_wgrad_workspace_bytes(C.byref(a)))
        ws = torch.empty(max(need // 4, 4), device=dev)
        dwp = torch.empty(K * K * Cout * Cin, device=dev)
        for _ in range(3):
            L.call("ng_conv2d_wgrad", C.byref(a), dwp.data_ptr(), None, ws.data_ptr(), need, stream)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(20):
            L.call("ng_conv2d_wgrad", C.byref(a), dwp.data_ptr(), None, ws.data_ptr(), need, stream)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 20 * 1e3
        flop = 2.0 * B * Ho * Ho * Cout * Cin * K * K / (s * s if form == L.FORM_PHASED else 1)
        if form == L.FORM_PHASED:
            flop = 2.0 * B * H * H * Cout * Cin * K * K
        mb = (x.numel() + dy.numel()) * 2 / 1e6
        out[name] = {"us": round(us, 1), "operand_MB": round(mb, 1), "hbm_floor_us": round(mb / 6443.0 * 1e3, 1),
                     "tflops": round(flop / us / 1e6, 1), "ws_MB": round(need / 1e6, 1)}
    print(json.dumps({"wgrad_micro": out}))


if __name__ == "__main__":
    main()
