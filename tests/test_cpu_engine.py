"""Host logic of the engine that needs no GPU: buffer cache and its batch-slice view."""
import pytest
import torch

import nirgan_b200  # noqa: F401
from nirgan_b200.engine import Buffers, SliceBuffers


def test_buffers_are_cached_by_name_size_and_dtype():
    b = Buffers(torch.device("cpu"))
    a = b.get("x", 16, torch.float32)
    assert b.get("x", 16, torch.float32) is a
    assert b.get("x", 32, torch.float32) is not a
    assert b.get("x", 16, torch.float16) is not a
    assert float(b.get("z", 8, torch.float32, zero=True).abs().sum()) == 0.0
    assert b.bytes() == 16 * 4 + 32 * 4 + 16 * 2 + 8 * 4


def test_slice_buffers_return_in_place_views_of_per_image_buffers():
    full = Buffers(torch.device("cpu"))
    t = full.get("act", 4 * 6, torch.float32)
    t.copy_(torch.arange(24, dtype=torch.float32))
    halves = [SliceBuffers(full, p, 2) for p in range(2)]
    h0, h1 = halves[0].get("act", 12, torch.float32), halves[1].get("act", 12, torch.float32)
    assert torch.equal(h0, t[:12]) and torch.equal(h1, t[12:])
    h1.fill_(-1.0)
    assert float(t[12:].sum()) == -12.0 and float(t[:12].sum()) == float(sum(range(12)))      # a view, not a copy
    with pytest.raises(KeyError):
        halves[0].get("act", 10, torch.float32)          # not half of a full-batch buffer
    with pytest.raises(KeyError):
        halves[0].get("scalar", 1, torch.float32)        # buffers that do not scale with the batch are refused


def test_plan_records_side_ops_and_launch_counts():
    """Plan bookkeeping that needs no GPU: ops bound to exported symbols, side flags, launch counts."""
    from nirgan_b200.engine import Plan
    p = Plan()
    p.add("ng_in_stats_finalize", 0, 1, 1, 2, 1, 0, label="a.fin")
    p.add("ng_conv2d_wgrad", None, 0, 0, 0, 0, launches=3, label="a.wgrad", side=True)
    p.add("ng_in_apply", label="")
    assert p.side == [False, True, False] and p.launches == 5
    assert p.labels == ["a.fin", "a.wgrad", "ng_in_apply"] and [o[2] for o in p.ops][1] == "ng_conv2d_wgrad"
    with pytest.raises(AttributeError):
        p.add("ng_no_such_entry_point")


def test_tile_batches_group_by_shape_in_list_order():
    from nirgan_b200.synth import _batches
    tiles = {f"t{i}": torch.zeros(3, 8 if i not in (3, 4) else 16, 8) for i in range(7)}
    names = [f"t{i}" for i in range(7)]
    assert _batches(names, tiles, 2) == [["t0", "t1"], ["t2"], ["t3", "t4"], ["t5", "t6"]]
    assert _batches(names, tiles, 64) == [["t0", "t1", "t2"], ["t3", "t4"], ["t5", "t6"]]
    assert _batches([], tiles, 4) == []


def test_plan_memory_never_overlaps_live_buffers():
    """engine.plan_memory: buffers whose live ranges (first..last op that is bound to a pointer inside them) intersect
    never share bytes; pinned buffers stay alive throughout; the total is well below the sum for a chain of layers."""
    import torch
    from nirgan_b200.engine import CountingBuffers, Plan, plan_memory
    cb = CountingBuffers("cpu")
    src = cb.get("in", 1000, torch.float32)
    bufs = [cb.get(f"l{i}", 5000 + 100 * (i % 3), torch.float16) for i in range(12)]
    out = cb.get("out", 300, torch.float32)
    p = Plan()
    prev = src
    for i, b in enumerate(bufs):                      # layer i reads layer i-1 (and i-3 as a residual), writes i
        res = bufs[i - 3].data_ptr() + 64 if i >= 3 else None
        p.add("ng_in_apply", prev.data_ptr(), res, b.data_ptr(), label=f"l{i}")
        prev = b
    p.add("ng_tap_gather", prev.data_ptr(), out.data_ptr(), label="head")
    offsets, total = plan_memory([p], cb, pinned=[src, out])
    live = {}
    for i, (fn, args, _) in enumerate(p.ops):
        for a in args:
            if a is None:
                continue
            for off, nbytes, key, _z in cb.ranges:
                if off <= a - 0x7F0000000000 < off + nbytes:
                    lo, hi = live.get(key, (i, i))
                    live[key] = (min(lo, i), max(hi, i))
    live[("in", 1000, torch.float32)] = (-1, len(p.ops))
    live[("out", 300, torch.float32)] = (-1, len(p.ops))
    size = {key: nbytes for _, nbytes, key, _z in cb.ranges}
    keys = list(size)
    for i, a in enumerate(keys):
        for b in keys[i + 1:]:
            la, lb = live[a], live[b]
            if not (la[1] < lb[0] or lb[1] < la[0]):          # alive at the same time -> disjoint memory
                assert offsets[a] + size[a] <= offsets[b] or offsets[b] + size[b] <= offsets[a], (a, b)
    assert total <= 0.6 * sum(size.values())
    assert total >= max(size.values())
