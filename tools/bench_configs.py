#!/usr/bin/env python
"""Secondary measurements for BASELINE.json configs[2] and configs[4] (one process per GPU; launch with
`python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P
tools/bench_configs.py --config 3|5 ...`, or directly for N=1).

--config 3  create_synthetic_dataset-style tile-sharded inference: synthetic 3x512x512 tiles (ids tile_%06d, sorted),
            plain generator behind the pad-10 wrapper (532/266/133 pyramid), tiles staged on the device, sharded over
            ranks with NO collective.  value = 512x512 tiles/s over all ranks (also as 256x256-tile equivalents).
--config 5  data-parallel Pix2Pix training step: batch 32 per GPU, per-step resolution drawn from a seeded sequence over
            {128,192,256,384,512} (same on all ranks), bucketed NCCL all-reduce of the D and of the G gradients overlapped with
            the backward pass (nirgan_b200.trainer.Trainer).  value = samples/s over all ranks.
Timing: CUDA events on the launching stream, barrier + synchronize on both sides, max over ranks.
"""
import argparse
import contextlib
import json
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
    ap.add_argument("--config", type=int, required=True, choices=[3, 5])
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--precision", default="fp16")
    ap.add_argument("--postprocess", action="store_true",
                    help="config 3: also histogram-match every prediction to a synthetic Sentinel-2 NIR band on the device")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        # stdout carries exactly one JSON line: keep NCCL's banner ("NCCL version ...") off it
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)

    import nirgan_b200  # noqa: F401
    from nirgan_b200.model.pix2pix import Px2Px
    from nirgan_b200 import synth
    from nirgan_b200.config import px2px_config as _cfg

    def barrier():
        torch.cuda.synchronize(dev)
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(ms):
        if dist is None:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    torch.manual_seed(0)                       # same weights on every rank
    with contextlib.redirect_stdout(sys.stderr):
        model = Px2Px(_cfg()).to(dev)
    model.netG.configure_b200(precision=args.precision, impl="tc")
    model.netD.configure_b200(precision=args.precision, impl="tc")

    if args.config == 3:
        model.eval()
        B = args.batch or 16
        gen = torch.Generator(device=dev).manual_seed(100 + rank)
        n_tiles = B * world
        names = [f"tile_{i:06d}.tif" for i in range(n_tiles)]
        mine = synth.shard(n_tiles, rank, world)
        x = torch.rand(len(mine), 3, 512, 512, generator=gen, device=dev)       # this rank's shard, resident in HBM

        s2 = torch.rand(len(mine), 1, 128, 128, generator=gen, device=dev) * 0.35

        def step():
            with torch.no_grad():
                y = model.predict_step(x)
                if args.postprocess:
                    from nirgan_b200.postprocess import postprocess
                    y = postprocess(y, s2)
                return y

        for _ in range(args.warmup):
            step()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            y = step()
        e1.record()
        barrier()
        ms = max_over_ranks(e0.elapsed_time(e1)) / args.steps
        if rank == 0:
            v = n_tiles / ms * 1e3
            print(json.dumps({"workload": "configs[2]: tile-sharded inference, 3x512x512 tiles, plain G behind the pad-10 "
                                          f"wrapper (532^2), {B} tiles/GPU/step, tiles resident on device, no collective",
                              "metric": "rgb2nir_512px_tiles_per_sec", "value": v, "unit": "tiles/s", "n_gpus": world,
                              "ms_per_step": ms, "equiv_256px_tiles_per_sec": 4 * v,
                              "model_tflops": v * 424.436 / 1e3, "precision": args.precision,
                              "postprocess": bool(args.postprocess),
                              "ids": [synth.tile_id(names[0]), synth.tile_id(names[-1])],
                              "out_shape": list(y.shape)}), flush=True)
    else:
        model.train()
        B = args.batch or 32
        from nirgan_b200.trainer import Trainer
        trainer = Trainer(model)
        opt_d, opt_g = trainer.opt_d, trainer.opt_g
        sizes = [128, 192, 256, 384, 512]
        seq_gen = torch.Generator().manual_seed(1234)
        total = args.warmup + args.steps
        seq = [sizes[int(torch.randint(0, len(sizes), (1,), generator=seq_gen))] for _ in range(total)]
        # every resolution is seen once before timing (plan compilation + buffer allocation are one-off costs)
        gen = torch.Generator(device=dev).manual_seed(200 + rank)
        data = {s: {"rgb": torch.rand(B, 3, s, s, generator=gen, device=dev),
                    "nir": torch.rand(B, 1, s, s, generator=gen, device=dev)} for s in sizes}

        def step(s):
            return trainer.step(data[s])

        for s in sorted(sizes, reverse=True):      # largest first: the shared buffer pool is sized once
            step(s)
        for s in seq[:args.warmup]:
            step(s)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in seq[args.warmup:]:
            ld, lg = step(s)
        e1.record()
        barrier()
        ms = max_over_ranks(e0.elapsed_time(e1))
        if rank == 0:
            timed = seq[args.warmup:]
            samples = B * world * len(timed)
            gflop = sum((391.6 if model.reuse_g_forward else 505.8) * (s / 256.0) ** 2 for s in timed) * B * world
            print(json.dumps({"workload": f"configs[4]: data-parallel Pix2Pix training, batch {B}/GPU, resolutions {timed}, "
                                          "NCCL all-reduce of D and G gradients each step",
                              "metric": "train_samples_per_sec", "value": samples / ms * 1e3, "unit": "samples/s",
                              "n_gpus": world, "ms_per_step": ms / len(timed), "algorithmic_tflops": gflop / ms,
                              "loss_D": float(ld.detach()), "loss_G": float(lg.detach()),
                              "mem_gb": torch.cuda.max_memory_allocated() / 1e9, "precision": args.precision,
                              "skipped_steps": [opt_d.skipped_steps, opt_g.skipped_steps]}), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
