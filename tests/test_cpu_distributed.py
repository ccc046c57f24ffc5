"""CPU suite for the N>1 host logic: tile sharding / gather and the DDP gradient all-reduce over gloo,
world_size 2 (no GPU compute: the model is a stand-in callable)."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

import nirgan_b200  # noqa: F401
from nirgan_b200 import synth
from nirgan_b200.optim import allreduce_gradients


def test_shards_partition_the_sorted_list():
    for n in (0, 1, 5, 64, 101):
        for world in (1, 2, 3, 8):
            for mode in ("contiguous", "strided"):
                parts = [synth.shard(n, r, world, mode) for r in range(world)]
                flat = sorted(i for p in parts for i in p)
                assert flat == list(range(n))
                assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    assert synth.tile_id("tile_000123.tif") == "tile_000123"
    assert synth.sorted_tiles(["b.tif", "a.tif", "c.tif"]) == ["a.tif", "b.tif", "c.tif"]


def _fake_model(hr):
    return hr.mean(1, keepdim=True) * 2.0 - 1.0


def _tiles():
    names = [f"tile_{i:06d}.tif" for i in (7, 3, 0, 5, 1, 6, 2, 4, 8)]
    t = {}
    for n in names:
        k = int(n[5:11])
        size = 16 if k % 3 else 24        # ragged: two tile sizes
        t[n] = torch.rand(3, size, size, generator=torch.Generator().manual_seed(k))
    return t


def test_single_rank_equals_sequential_loop():
    tiles = _tiles()
    seq = {synth.tile_id(n): _fake_model(tiles[n][None])[0] for n in sorted(tiles)}
    out = synth.run_shard(_fake_model, tiles, 0, 1, batch_size=2)
    assert list(out.keys()) == list(seq.keys())
    for k in seq:
        assert torch.equal(out[k], seq[k])


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    tiles = _tiles()
    local = synth.run_shard(_fake_model, tiles, rank, world, batch_size=2)
    merged = synth.gather_shards(local)
    # DDP gradient averaging: every rank holds different grads, all end with the mean
    torch.manual_seed(0)
    ps = [torch.nn.Parameter(torch.zeros(5, 3)), torch.nn.Parameter(torch.zeros(7)), torch.nn.Parameter(torch.zeros(2))]
    ps[0].grad = torch.full((5, 3), float(rank + 1))
    ps[1].grad = torch.arange(7.0) * (rank + 1)
    allreduce_gradients(ps)          # third parameter has no grad: skipped
    # bucketed exchange of a flat arena (the training path): buckets go out in reverse order as the backward pass reports
    # "everything at or above this offset is final"; finish() sends the rest; the result is the SUM over ranks
    from nirgan_b200.optim import BucketedAllReduce
    flat = torch.arange(1000.0) * (rank + 1)
    red = BucketedAllReduce(flat, min_bucket_elems=100)
    red.ready(950)                   # 50 elements: below the bucket threshold, held back
    held = red.buckets
    red.ready(600)                   # [600, 1000) goes out
    red.ready(580)                   # held back again
    nb = red.finish()                # [0, 600)
    bucket_ok = bool(torch.equal(flat, torch.arange(1000.0) * 3)) and held == 0 and nb == 2 and red.last_buckets == 2
    # plain numpy payloads: torch tensors travel through shared-memory handles that die with the worker
    q.put((rank, list(merged.keys()), {k: v.numpy().copy() for k, v in merged.items()}, ps[0].grad.numpy().copy(),
           ps[1].grad.numpy().copy(), len(local), bucket_ok))
    dist.destroy_process_group()


def test_two_ranks_gloo_sharded_inference_and_grad_allreduce():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    tiles = _tiles()
    seq = {synth.tile_id(n): _fake_model(tiles[n][None])[0] for n in sorted(tiles)}
    assert sum(r[5] for r in res) == len(tiles)
    assert all(r[6] for r in res), "bucketed all-reduce over gloo"
    for rank, keys, merged, g0, g1, _, _ in res:
        assert keys == list(seq.keys())                       # identical {id -> array} mapping on every rank
        for k in seq:
            assert torch.equal(torch.from_numpy(merged[k]), seq[k])
        assert torch.equal(torch.from_numpy(g0), torch.full((5, 3), 1.5))       # mean of 1 and 2
        assert torch.equal(torch.from_numpy(g1), torch.arange(7.0) * 1.5)
