#!/usr/bin/env python
"""Post-processing throughput: nearest x4 + histogram matching + float16 for a batch of
512x512 predicted tiles, device kernels vs the CPU oracle (numpy restatement of skimage.exposure.match_histograms)."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--tile", type=int, default=512)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    a = ap.parse_args()
    import nirgan_b200  # noqa: F401
    from nirgan_b200 import postprocess as PP
    import postprocess_oracle as P
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(0)
    pred = torch.tanh(torch.randn(a.batch, 1, a.tile, a.tile, generator=g, device=dev))
    s2 = torch.rand(a.batch, 1, a.tile // 4, a.tile // 4, generator=g, device=dev) * 0.35
    for _ in range(a.warmup):
        PP.postprocess(pred, s2)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        out = PP.postprocess(pred, s2)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    n = min(a.batch, 8)
    pc, sc = pred[:n].cpu(), s2[:n].cpu()
    t0 = time.perf_counter()
    want = P.postprocess(pc, sc)
    cpu_s = time.perf_counter() - t0
    err = float((out[:n].float().cpu() - want.float()).abs().max())
    px = a.batch * a.tile * a.tile
    print(json.dumps({"workload": f"post-processing of {a.batch} predicted {a.tile}x{a.tile} tiles (nearest x4, histogram "
                                  "matching, float16)", "ms_per_batch": ms, "tiles_per_s": a.batch / ms * 1e3,
                      "sort_keys_per_s": 2 * px / ms * 1e3, "cpu_oracle_tiles_per_s": n / cpu_s, "cpu_sample_tiles": n,
                      "max_abs_vs_oracle_fp16": err}))


if __name__ == "__main__":
    main()
