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


def run_shard(model: Callable[[torch.Tensor], torch.Tensor], tiles: Mapping[str, torch.Tensor], rank: int = 0,
              world: int = 1, batch_size: int = 64, device: Optional[torch.device] = None,
              mode: str = "contiguous", s2_nir: Optional[Mapping[str, torch.Tensor]] = None) -> Dict[str, torch.Tensor]:
    """Run ``model(hr)`` (e.g. ``Px2Px.forward`` in eval mode) over this rank's tiles; returns {id: (1,H,W) fp32 CPU}.
    Tiles of different sizes are batched by size in list order.  With ``s2_nir`` ({filename: (1,h,w) Sentinel-2 NIR at a
    quarter of the tile's resolution}) the prediction is histogram-matched to it on the device and returned as float16,
    i.e. the loop body of create_synthetic_dataset.py:107-116 (device-side ``postprocess.postprocess``)."""
    names = sorted_tiles(list(tiles.keys()))
    mine = [names[i] for i in shard(len(names), rank, world, mode)]
    out: Dict[str, torch.Tensor] = {}
    i = 0
    with torch.no_grad():
        while i < len(mine):
            shape = tiles[mine[i]].shape
            j = i
            while j < len(mine) and j - i < batch_size and tiles[mine[j]].shape == shape:
                j += 1
            hr = torch.stack([tiles[n] for n in mine[i:j]])
            if device is not None:
                hr = hr.to(device, non_blocking=True)
            pred = model(hr).float()
            if s2_nir is not None:
                from .postprocess import postprocess
                ref = torch.stack([s2_nir[n] for n in mine[i:j]])
                pred = postprocess(pred, ref.to(pred.device, non_blocking=True))
            pred = pred.cpu()
            for n, p in zip(mine[i:j], pred):
                out[tile_id(n)] = p
            i = j
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
