#!/usr/bin/env python
"""The "existing Blackwell kernel" bar of SURVEY.md 8(d): the same network evaluated by stock torch modules (cuDNN /
cuBLAS kernels as shipped with torch) on the B200 -- NOT the CPU reference and NOT part of the product.  The generator
below is an independent torch.nn statement of the architecture (model/networks.py:322-374,391-434 +
model/generator_inject.py:105-135) with random weights; only its speed is of interest.

Modes: fp32 with TF32 convolutions, and bf16 autocast + channels_last.  Workloads: configs[1] inference (B=64, 256^2,
injected) and configs[3] training step (B=32, 256^2 behind the pad-10 wrapper, G+D, LSGAN + 100*L1 + index losses, Adam).
Prints one JSON line."""
import argparse
import json
import time

import torch
import torch.nn as nn
import torch.nn.functional as F


class Block(nn.Module):
    def __init__(s, c):
        super().__init__()
        s.body = nn.Sequential(nn.ReflectionPad2d(1), nn.Conv2d(c, c, 3), nn.InstanceNorm2d(c), nn.ReLU(True),
                               nn.ReflectionPad2d(1), nn.Conv2d(c, c, 3), nn.InstanceNorm2d(c))

    def forward(s, x):
        return x + s.body(x)


class G(nn.Module):
    def __init__(s, inject=True, ngf=64):
        super().__init__()
        s.inject = inject
        s.stem = nn.Sequential(nn.ReflectionPad2d(3), nn.Conv2d(3, ngf, 7), nn.InstanceNorm2d(ngf), nn.ReLU(True))
        s.d1 = nn.Sequential(nn.Conv2d(ngf, 2 * ngf, 3, 2, 1), nn.InstanceNorm2d(2 * ngf))
        s.d2 = nn.Sequential(nn.Conv2d(2 * ngf, 4 * ngf, 3, 2, 1), nn.InstanceNorm2d(4 * ngf), nn.ReLU(True))
        s.blocks = nn.Sequential(*[Block(4 * ngf) for _ in range(9)])
        s.up = nn.Sequential(nn.ConvTranspose2d(4 * ngf, 2 * ngf, 3, 2, 1, 1), nn.InstanceNorm2d(2 * ngf), nn.ReLU(True),
                             nn.ConvTranspose2d(2 * ngf, ngf, 3, 2, 1, 1), nn.InstanceNorm2d(ngf), nn.ReLU(True),
                             nn.ReflectionPad2d(3), nn.Conv2d(ngf, 1, 7), nn.Tanh())
        if inject:
            s.fc = nn.Linear(256, 128 * 128)
            s.scale = nn.Parameter(torch.tensor(0.01))

    def forward(s, x, e=None):
        x = s.d1(s.stem(x))
        if s.inject:
            m = s.fc(e).view(-1, 1, 128, 128)
            m = F.interpolate(m, size=x.shape[-2:], mode="bilinear")
            x = x * (1 + s.scale * m)
        return s.up(s.blocks(s.d2(F.relu(x))))


def D():
    L = [nn.Conv2d(4, 64, 4, 2, 1), nn.LeakyReLU(0.2, True)]
    for ci, co, st in ((64, 128, 2), (128, 256, 2), (256, 512, 1)):
        L += [nn.Conv2d(ci, co, 4, st, 1), nn.InstanceNorm2d(co), nn.LeakyReLU(0.2, True)]
    return nn.Sequential(*L, nn.Conv2d(512, 1, 4, 1, 1))


def timed(fn, steps, warmup):
    for _ in range(warmup):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def rs(rgb, n):
    r, g, b = rgb[:, 0:1], rgb[:, 1:2], rgb[:, 2:3]
    return (n - r) / (n + r + 1e-6), (n - g) / (n + g + 1e-6), 2.5 * (n - r) / ((n + 6) * (r - 7.5) * (b + 1) + 1e-6)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    torch.backends.cudnn.benchmark = True
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.manual_seed(0)
    out = {}
    for mode in ("tf32", "bf16_channels_last"):
        cl = mode != "tf32"
        g = G(True).to(dev).eval()
        if cl:
            g = g.to(memory_format=torch.channels_last)
        x = torch.rand(64, 3, 256, 256, device=dev)
        e = torch.randn(64, 256, device=dev)
        if cl:
            x = x.contiguous(memory_format=torch.channels_last)

        def infer():
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=cl):
                return g(x, e)

        ms = timed(infer, a.steps, a.warmup)
        out[f"infer_{mode}"] = {"ms_per_step": round(ms, 3), "tiles_per_s": round(64 / ms * 1e3, 1)}

        # training step, configs[3]: B=32, pad-10 wrapper, D then G, Adam(2e-4, (0.5, 0.999))
        gt, d = G(False).to(dev).train(), D().to(dev).train()
        if cl:
            gt, d = gt.to(memory_format=torch.channels_last), d.to(memory_format=torch.channels_last)
        og = torch.optim.Adam(gt.parameters(), 2e-4, betas=(0.5, 0.999))
        od = torch.optim.Adam(d.parameters(), 2e-4, betas=(0.5, 0.999))
        rgb = torch.rand(32, 3, 256, 256, device=dev)
        nir = torch.rand(32, 1, 256, 256, device=dev)

        def fwd():
            return gt(F.pad(rgb, (10, 10, 10, 10), mode="reflect"))[..., 10:-10, 10:-10]

        def train():
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=cl):
                pred = fwd()                                   # reference-faithful: G evaluated in both passes
                pf = d(torch.cat((rgb, pred.detach()), 1))
                pr = d(torch.cat((rgb, nir), 1))
                ld = F.mse_loss(pf.float(), torch.zeros_like(pf, dtype=torch.float32)) + \
                    F.mse_loss(pr.float(), torch.ones_like(pr, dtype=torch.float32))
            od.zero_grad(set_to_none=True)
            ld.backward()
            od.step()
            for p in d.parameters():
                p.requires_grad_(False)
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=cl):
                pred = fwd()
                pf = d(torch.cat((rgb, pred), 1))
                lg = F.mse_loss(pf.float(), torch.ones_like(pf, dtype=torch.float32)) + 100 * F.l1_loss(pred.float(), nir)
                for t, p in zip(rs(rgb, nir), rs(rgb, pred.float())):
                    lg = lg + 0.33 * F.l1_loss(p, t)
            og.zero_grad(set_to_none=True)
            lg.backward()
            og.step()
            for p in d.parameters():
                p.requires_grad_(True)

        ms = timed(train, max(3, a.steps // 2), a.warmup)
        out[f"train_{mode}"] = {"ms_per_step": round(ms, 2), "samples_per_s": round(32 / ms * 1e3, 1)}
        del g, gt, d, og, od
        torch.cuda.empty_cache()
    out["torch"] = torch.__version__
    out["cudnn"] = torch.backends.cudnn.version()
    print(json.dumps({"torch_gpu_bar": out}))


if __name__ == "__main__":
    t0 = time.time()
    main()
