#!/usr/bin/env python
"""Parity of the full generator / discriminator forward per numeric mode against the reference-produced golden fixtures
(tests/golden): max-abs and mean-abs error on the tanh [-1,1] output.  Prints a markdown table (profiles/)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import nirgan_oracle as O  # noqa: E402
from test_gpu_models import make_G  # noqa: E402


def main():
    gd = os.path.join(ROOT, "tests", "golden")
    rows = []
    for prec, impl in (("fp32", "simt"), ("fp16", "tc"), ("bf16", "tc")):
        for name, inj in (("g_plain_64", False), ("g_plain_64_pad10", False), ("g_plain_256", False),
                          ("g_inject_64_s1.0", True), ("g_inject_64_pad10_s1.0", True)):
            g = np.load(f"{gd}/{name}.npz")
            kw = {"bias_std": float(g["bias_std"])} if "bias_std" in g else {}
            if inj:
                kw["scale_param"] = float(g["scale"]) if "scale" in g else 1.0
            sd = O.random_state_dict(O.generator_param_shapes(inject=inj), seed=int(g["sd_seed"]), **kw)
            net = make_G(sd, prec, impl, inject=inj)
            if "x" in g:
                x = torch.from_numpy(g["x"])
            else:
                x = torch.rand(1, 3, 256, 256, generator=torch.Generator().manual_seed(int(g["x_seed"])))
            pad = 10 if "pad10" in name else 0
            with torch.no_grad():
                y = net(x.cuda(), torch.from_numpy(g["embeds"]).cuda(), wrap_pad=pad) if inj else net(x.cuda(), wrap_pad=pad)
            d = (y.float().cpu() - torch.from_numpy(g["y"])).abs()
            rows.append((prec + "/" + impl, name, float(d.max()), float(d.mean())))
    print("| mode | fixture | max-abs | mean-abs |\n|---|---|---|---|")
    for r in rows:
        print(f"| {r[0]} | {r[1]} | {r[2]:.3e} | {r[3]:.3e} |")


if __name__ == "__main__":
    main()
