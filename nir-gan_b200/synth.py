"""Tile-sharded mirror of the reference's inference loop (create_synthetic_dataset.py:93,100-118).

Reference: ``dl = DataLoader(SR_dataset(root), batch_size=2, shuffle=False)``; for every batch
``pred = model(hr)`` and each sample is stored under ``id = fname.split('.')[0]`` where the file list is
``sorted(os.listdir(LR))`` (data/SR_dataset_RGB.py:16-19,55).  Here the sorted tile list is partitioned over
ranks (one process per GPU, no collective on the data path); because InstanceNorm statistics are per sample
and the kernels' work decomposition is per image, a tile's result is bit-identical whichever rank / batch slot
computes it, so the union of the shards equals the sequential loop's ``{id: array}`` mapping exactly.
Histogram matching and the .npz writer that follow the loop live in ``postprocess.py``.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Mapping, Optional, Sequence

import torch


def tile_id(filename: str) -> str:
    """data/SR_dataset_RGB.py:55."""
    return filename.split(".")[0]


def sorted_tiles(filenames: Sequence[str]) -> List[str]:
    """data/SR_dataset_RGB.py:16-19 ordering contract."""
    return sorted(filenames)


def shard(n_items: int, rank: int, world: int, mode: str = "contiguous") -> List[int]:
    """Indices of the sorted tile list owned by `rank`.  'contiguous' keeps neighbouring tiles together
    (balanced to within one tile); 'strided' is i == rank (mod world)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    if mode == "strided":
        return list(range(rank, n_items, world))
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return list(range(start, start + base + (1 if rank < extra else 0)))


def _batches(mine: List[str], tiles: Mapping[str, torch.Tensor], batch_size: int) -> List[List[str]]:
    """Consecutive runs of same-shaped tiles, at most ``batch_size`` long, in list order."""
    out, i = [], 0
    while i < len(mine):
        shape = tiles[mine[i]].shape
        j = i
        while j < len(mine) and j - i < batch_size and tiles[mine[j]].shape == shape:
            j += 1
        out.append(mine[i:j])
        i = j
    return out


def run_shard(model: Callable[[torch.Tensor], torch.Tensor], tiles: Mapping[str, torch.Tensor], rank: int = 0,
              world: int = 1, batch_size: int = 64, device: Optional[torch.device] = None,
              mode: str = "contiguous", s2_nir: Optional[Mapping[str, torch.Tensor]] = None,
              model_async: Optional[Callable] = None, in_flight: int = 3) -> Dict[str, torch.Tensor]:
    """Run ``model(hr)`` (e.g. ``Px2Px.forward`` in eval mode) over this rank's tiles; returns {id: (1,H,W) fp32 CPU}.
    Tiles of different sizes are batched by size in list order.  With ``s2_nir`` ({filename: (1,h,w) Sentinel-2 NIR at a
    quarter of the tile's resolution}) the prediction is histogram-matched to it on the device and returned as float16,
    i.e. the loop body of create_synthetic_dataset.py:107-116 (device-side ``postprocess.postprocess``).

    ``model_async(hr, ready_event) -> (pred, done_events)`` (e.g. ``lambda hr, ev: netG.forward_async(hr, None, pad, ev)``)
    selects the pipelined loop: batches are staged in pinned memory and copied on an H2D stream, the generator's slices
    queue behind those of the previous batch, and post-processing + read-back run on a third stream, with at most
    ``in_flight`` batches between stack and read-back.  Results are identical to the plain loop."""
    names = sorted_tiles(list(tiles.keys()))
    mine = [names[i] for i in shard(len(names), rank, world, mode)]
    out: Dict[str, torch.Tensor] = {}
    if model_async is not None:
        if device is None or torch.device(device).type != "cuda":
            raise RuntimeError("nirgan_b200: the pipelined tile loop needs a CUDA device")
        return _run_shard_pipelined(model_async, tiles, mine, batch_size, torch.device(device), s2_nir, in_flight)
    with torch.no_grad():
        for group in _batches(mine, tiles, batch_size):
            hr = torch.stack([tiles[n] for n in group])
            if device is not None:
                hr = hr.to(device, non_blocking=True)
            pred = model(hr).float()
            if s2_nir is not None:
                from .postprocess import postprocess
                ref = torch.stack([s2_nir[n] for n in group])
                pred = postprocess(pred, ref.to(pred.device, non_blocking=True))
            pred = pred.cpu()
            for n, p in zip(group, pred):
                out[tile_id(n)] = p
    return out


@torch.no_grad()
def _run_shard_pipelined(model_async, tiles, mine, batch_size, device, s2_nir, in_flight) -> Dict[str, torch.Tensor]:
    out: Dict[str, torch.Tensor] = {}
    h2d, post = torch.cuda.Stream(device), torch.cuda.Stream(device)
    pending = []                      # (names, pinned result, event)

    def retire():
        group, host, ev = pending.pop(0)
        ev.synchronize()
        for n, p in zip(group, host):
            out[tile_id(n)] = p.clone()

    for group in _batches(mine, tiles, batch_size):
        hr_host = torch.stack([tiles[n] for n in group]).pin_memory()
        ref_host = torch.stack([s2_nir[n] for n in group]).pin_memory() if s2_nir is not None else None
        with torch.cuda.stream(h2d):
            hr = hr_host.to(device, non_blocking=True)
            ref = ref_host.to(device, non_blocking=True) if ref_host is not None else None
            ready = torch.cuda.Event()
            ready.record(h2d)
        cur = torch.cuda.current_stream(device)
        hr.record_stream(cur)
        pred, done = model_async(hr, ready)
        with torch.cuda.stream(post):
            for ev in done:
                post.wait_event(ev)
            post.wait_event(ready)
            pred.record_stream(post)
            res = pred.float()
            if ref is not None:
                from .postprocess import postprocess
                ref.record_stream(post)
                res = postprocess(res, ref)
            host = torch.empty(res.shape, dtype=res.dtype).pin_memory()
            host.copy_(res, non_blocking=True)
            fin = torch.cuda.Event()
            fin.record(post)
        pending.append((group, host, fin))
        while len(pending) > max(1, in_flight):
            retire()
    while pending:
        retire()
    return out


def gather_shards(local: Dict[str, torch.Tensor], group=None) -> Dict[str, torch.Tensor]:
    """Host-side merge of the per-rank dictionaries (the only cross-rank step; not on the GPU data path)."""
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return dict(sorted(local.items()))
    parts = [None] * dist.get_world_size(group)
    dist.all_gather_object(parts, local, group=group)
    merged: Dict[str, torch.Tensor] = {}
    for p in parts:
        for k, v in p.items():
            if k in merged:
                raise RuntimeError(f"tile id {k} produced by two ranks")
            merged[k] = v
    return dict(sorted(merged.items()))
