#!/usr/bin/env python
"""Pins oracle/postprocess_oracle.py::emd_loss (and records ssim_loss) -- run in the build container only.

``utils/losses.py`` of the reference imports kornia and scipy.stats at module level; kornia is not installed here, so the
module is imported with a stub ``kornia`` package: ``emd_loss`` (losses.py:64-78) is pure torch and runs unmodified, which
pins the oracle's restatement bit-exactly, values AND autograd gradients.  ``ssim_loss`` needs kornia.metrics.ssim itself
-> stays PARITY UNPINNED (oracle = restatement of kornia 0.7.3's published formula); its fixture values come from the
oracle and are marked as such.

Writes tests/golden/losses_small.npz and appends to tests/golden/PIN_REPORT_losses.txt.
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import postprocess_oracle as P  # noqa: E402


def reference_losses():
    sys.modules.setdefault("kornia", types.ModuleType("kornia"))
    sys.path.insert(0, REF)
    import importlib.util
    spec = importlib.util.spec_from_file_location("_ref_losses", os.path.join(REF, "utils", "losses.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def main():
    ref = reference_losses()
    lines = []
    g = torch.Generator().manual_seed(0)
    cases = {"a": (2, 1, 32, 32), "b": (3, 1, 40, 24), "c": (1, 2, 16, 16), "d": (2, 1, 64, 64)}
    out = {}
    for name, shape in cases.items():
        target = torch.rand(shape, generator=g)
        pred = (target + 0.3 * torch.randn(shape, generator=g)).clamp(-1, 1)
        for fn in ("emd_loss",):
            p1 = pred.clone().requires_grad_(True)
            r = getattr(ref, fn)(p1, target)
            r.backward()
            p2 = pred.clone().requires_grad_(True)
            o = getattr(P, fn)(p2, target)
            o.backward()
            dv, dg = float((r - o).abs()), float((p1.grad - p2.grad).abs().max())
            lines.append(f"{fn}[{name} {shape}]  value ref {float(r):.9e} oracle {float(o):.9e} |d| {dv:.1e}  grad max|d| {dg:.1e}  "
                         f"{'ok' if dv == 0.0 and dg == 0.0 else 'MISMATCH'}")
            out[f"{name}.emd"] = np.float32(r.item())
            out[f"{name}.emd_grad"] = p1.grad.numpy()
        for win in (5, 11):
            p3 = pred.clone().requires_grad_(True)
            s = P.ssim_loss(p3, target, win)
            s.backward()
            out[f"{name}.ssim{win}"] = np.float32(s.item())
            out[f"{name}.ssim{win}_grad"] = p3.grad.numpy()
            lines.append(f"ssim_loss[{name} {shape} w{win}]  oracle {float(s):.9e}  (kornia absent: unpinned)")
        out[f"{name}.pred"], out[f"{name}.target"] = pred.numpy(), target.numpy()
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "losses_small.npz"), **out)
    rep = "\n".join(lines)
    print(rep)
    with open(os.path.join(ROOT, "tests", "golden", "PIN_REPORT_losses.txt"), "w") as f:
        f.write("oracle/pin_losses.py: oracle vs the reference's utils/losses.py (kornia stubbed)\n" + rep + "\n")
    if "MISMATCH" in rep:
        raise SystemExit(1)


if __name__ == "__main__":
    main()
