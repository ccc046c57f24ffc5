#!/usr/bin/env python
"""Secondary measurement (BASELINE.json configs[3]): one full Pix2Pix training step (D pass then G pass, LSGAN +
100*L1 + NDVI/NDWI/EVI, Adam) at batch 32, 256x256 tiles, through nirgan_b200.model.pix2pix.Px2Px.
Prints one JSON line: samples/s, ms per step, algorithmic TFLOP/s (391.6 GFLOP per sample)."""
import argparse
import contextlib
import json
import time
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
# more hardware queues than streams (compute slices + copy streams): no false dependencies between streams
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--tile", type=int, default=256)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--precision", default="fp16")
    ap.add_argument("--conv", default="tc")
    ap.add_argument("--profile", action="store_true", help="per-op device times of one extra step -> stderr + gpurun_out")
    args = ap.parse_args()
    import nirgan_b200  # noqa: F401
    from nirgan_b200.model.pix2pix import Px2Px
    from nirgan_b200.config import px2px_config as _cfg
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    with contextlib.redirect_stdout(sys.stderr):
        model = Px2Px(_cfg()).to(dev).train()
    model.netG.configure_b200(precision=args.precision, impl=args.conv)
    model.netD.configure_b200(precision=args.precision, impl=args.conv)
    from nirgan_b200.trainer import Trainer
    trainer = Trainer(model)
    opt_d, opt_g = trainer.opt_d, trainer.opt_g
    g = torch.Generator().manual_seed(1)
    batch = {"rgb": torch.rand(args.batch, 3, args.tile, args.tile, generator=g).to(dev),
             "nir": torch.rand(args.batch, 1, args.tile, args.tile, generator=g).to(dev)}

    # NIRGAN_B200_HP_STREAM=1: run the step on a high-priority stream (the weight-gradient side stream keeps the default
    # priority), so that the main-stream kernels are dispatched first when both streams have work
    import os
    hp = torch.cuda.Stream(dev, priority=-1) if os.environ.get("NIRGAN_B200_HP_STREAM", "0") == "1" else None
    if hp is not None:
        hp.wait_stream(torch.cuda.current_stream(dev))
        torch.cuda.set_stream(hp)

    def step():
        return trainer.step(batch)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    t_host = time.perf_counter()
    for _ in range(args.steps):
        ld, lg = step()
    host_ms = (time.perf_counter() - t_host) * 1e3 / args.steps      # time to ENQUEUE a step (no device sync inside)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    if args.profile:
        import re
        from nirgan_b200 import engine
        engine.PROFILE[0] = []
        e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e2.record()
        step()
        e3.record()
        torch.cuda.synchronize()
        recs = engine.PROFILE[0]
        engine.PROFILE[0] = None
        agg, tot = {}, 0.0
        for label, name, a, b in recs:
            t = a.elapsed_time(b)
            tot += t
            key = (label.split(".")[0], name, re.sub(r"^[a-z0-9]+\.", "", label))
            agg[key] = agg.get(key, 0.0) + t
        rows = sorted(agg.items(), key=lambda kv: -kv[1])
        print(f"profiled step {e2.elapsed_time(e3):.2f} ms wall, {tot:.2f} ms inside plans", file=sys.stderr)
        byname = {}
        for (tag, name, lab), t in rows:
            byname[name] = byname.get(name, 0.0) + t
        print("by entry point:", {k: round(v, 3) for k, v in sorted(byname.items(), key=lambda kv: -kv[1])}, file=sys.stderr)
        for (tag, name, lab), t in rows[:60]:
            print(f"  {t:8.3f} ms  {tag:5s} {name:24s} {lab}", file=sys.stderr)
        out_dir = os.path.join(ROOT, "gpurun_out")
        if os.path.isdir(out_dir):
            with open(os.path.join(out_dir, "train_op_times.json"), "w") as f:
                json.dump([{"plan": tag, "op": name, "label": lab, "ms": round(t, 4)} for (tag, name, lab), t in rows], f, indent=0)
    # op count per sample: 505.8 GFLOP reference-faithful (2 G forwards per batch, pix2pix.py:177-180) or
    # 391.6 when the G pass reuses the D pass's forward (the default here; NIRGAN_B200_REUSE_G=0 disables it)
    gflop = (391.6 if model.reuse_g_forward else 505.8) * (args.tile / 256.0) ** 2
    print(json.dumps({"workload": f"configs[3]: Pix2Pix training step, batch {args.batch}, {args.tile}px, {args.precision}/{args.conv}",
                      "ms_per_step": ms, "samples_per_s": args.batch / ms * 1e3,
                      "algorithmic_tflops": args.batch * gflop / ms, "gflop_per_sample": gflop,
                      "g_forward_reused": model.reuse_g_forward, "host_enqueue_ms_per_step": round(host_ms, 2),
                      "loss_D": float(ld), "loss_G": float(lg),
                      "mem_gb": torch.cuda.max_memory_allocated() / 1e9,
                      "skipped_steps": [opt_d.skipped_steps, opt_g.skipped_steps],
                      "adam_arena_steps": [opt_d.fast_steps, opt_g.fast_steps]}))


if __name__ == "__main__":
    main()
