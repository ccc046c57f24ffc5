"""Host-side execution engine: lays the hot path out as a static plan of C-ABI calls.

The engine owns nothing numerical: every FLOP runs in ``csrc/libnirgan_b200.so``.  It decides
buffer layouts (NHWC, haloed), keeps packed low-precision weight shadows in sync with the fp32
master parameters of the reference-compatible ``nn.Module``s, and compiles, per (batch, height,
width), a flat list of pre-bound ctypes calls -- so a forward is ~70 kernel launches with no Python
arithmetic in between and can be captured into a CUDA graph.

Data layout in HBM:
  activations   NHWC, 16-bit (fp16 default, bf16 optional) or fp32 in verification mode,
                with the halo (reflect or zero) the *next* convolution needs already materialised
  pre-norm      NHWC compact conv outputs + per-(n,c) (mean, rstd) fp32
  weights       [tap][Cout][Cin], K(=Cin)-contiguous, same element type as the activations
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Tuple

import torch

from . import _lib as L

# bumped by B200Adam.step(): parameter storage is updated in place by ng_adam_step (no torch version bump)
WEIGHT_EPOCH = [0]

_PRECISIONS = {
    "fp16": (L.F16, torch.float16),
    "bf16": (L.BF16, torch.bfloat16),
    "fp32": (L.F32, torch.float32),
}


@dataclass
class EngineConfig:
    """precision: 'fp16' (default fast mode; meets the 2e-2 / 2e-3 tolerance),
    'bf16' (same tensor rate, wider range, larger error) or 'fp32' (verification mode, CUDA-core
    path, <= 1e-4).  impl: 'tc' = tcgen05 kernels, 'simt' = CUDA-core kernels (any precision)."""
    precision: str = "fp16"
    impl: str = "tc"
    chunk: int = 0          # images per pass through the network (0 = whole batch)
    streams: int = 0        # inference: run this many batch slices concurrently on separate CUDA streams, so the
                            # HBM-bound apply kernels of one slice overlap the tensor-bound convolutions of another
                            # (0 = auto: 2 slices for batches of >= 32 tiles, measured best on B200; else 1)

    def __post_init__(self):
        if self.precision not in _PRECISIONS:
            raise ValueError(f"precision must be one of {list(_PRECISIONS)}")
        if self.impl not in ("tc", "simt"):
            raise ValueError("impl must be 'tc' or 'simt'")
        if self.precision == "fp32":
            self.impl = "simt"      # tensor cores have no fp32-operand mode that meets 1e-4

    @staticmethod
    def from_env() -> "EngineConfig":
        return EngineConfig(os.environ.get("NIRGAN_B200_PRECISION", "fp16"),
                            os.environ.get("NIRGAN_B200_IMPL", "tc"),
                            int(os.environ.get("NIRGAN_B200_CHUNK", "0")),
                            int(os.environ.get("NIRGAN_B200_STREAMS", "0")))


def require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"nirgan_b200: {what} must be a CUDA tensor on an sm_100 (B200) device; "
                           f"the hot path has no CPU fallback")


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


class Buffers:
    """Named, shape-keyed device buffer cache (the caller -- PyTorch -- owns all memory)."""

    def __init__(self, device):
        self.device = device
        self._b: Dict[Tuple, torch.Tensor] = {}

    def get(self, name: str, numel: int, dtype: torch.dtype, zero: bool = False) -> torch.Tensor:
        key = (name, numel, dtype)
        t = self._b.get(key)
        if t is None:
            t = (torch.zeros if zero else torch.empty)(numel, dtype=dtype, device=self.device)
            self._b[key] = t
        return t

    def bytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in self._b.values())

    def drop(self, tags) -> None:
        """Forget every buffer whose name starts with one of `tags` + '.' (the buffers of an evicted plan)."""
        tags = tuple(t + "." for t in tags)
        if not tags:
            return
        for k in [k for k in self._b if k[0].startswith(tags)]:
            del self._b[k]


class Pool:
    """One device allocation that the training contexts of a runner carve their buffers from.  Contexts of different
    input shapes (mixed resolutions, BASELINE config 5) are never in flight at the same time -- a step's forward and
    backward complete before the next step's forward starts -- so each context bump-allocates from offset 0 and they all
    alias the same memory: the footprint is that of the LARGEST shape, not the sum over shapes."""

    def __init__(self, nbytes: int, device):
        self.capacity = int(nbytes)
        self.t = torch.empty(self.capacity, dtype=torch.uint8, device=device)
        self.owner = None          # the context whose activations currently live in the pool
        self.live = None           # the context whose forward is waiting for its backward


class PoolTooSmall(RuntimeError):
    pass


class PoolBuffers:
    """``Buffers`` interface over a ``Pool``: named buffers of ONE context, 256-byte aligned, allocated in request order.
    Buffers that must keep their contents between steps of other shapes (zero-initialised counters) live outside."""

    ALIGN = 256

    def __init__(self, pool: Pool, device):
        self.pool, self.device = pool, device
        self.off = 0
        self._b: Dict[Tuple, torch.Tensor] = {}
        self._own: Dict[Tuple, torch.Tensor] = {}

    def get(self, name: str, numel: int, dtype: torch.dtype, zero: bool = False) -> torch.Tensor:
        key = (name, numel, dtype)
        t = self._b.get(key)
        if t is not None:
            return t
        if zero:
            t = self._own[key] = torch.zeros(numel, dtype=dtype, device=self.device)
        else:
            esz = torch.empty(0, dtype=dtype).element_size()
            nbytes = (numel * esz + self.ALIGN - 1) // self.ALIGN * self.ALIGN
            if self.off + nbytes > self.pool.capacity:
                raise PoolTooSmall(f"pool of {self.pool.capacity} bytes exhausted at {name!r}")
            t = self.pool.t[self.off:self.off + numel * esz].view(dtype)
            self.off += nbytes
        self._b[key] = t
        return t

    def bytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in self._own.values())

    def drop(self, tags) -> None:
        pass


class _SizedStub:
    """Stands in for a device buffer while a context is built only to learn how much memory it needs: plan compilation
    binds pointers and sizes but never touches buffer contents."""

    FAKE_BASE = 0x7F0000000000         # non-null, 256-byte aligned, never dereferenced

    def __init__(self, numel: int, dtype: torch.dtype, fake_off: int = 0):
        self._n, self._esz = int(numel), torch.empty(0, dtype=dtype).element_size()
        self.dtype = dtype
        self._addr = self.FAKE_BASE + fake_off

    def data_ptr(self) -> int:
        return self._addr

    def numel(self) -> int:
        return self._n

    def element_size(self) -> int:
        return self._esz


class CountingBuffers:
    """``Buffers`` interface that allocates nothing and adds up what a ``PoolBuffers`` would need."""

    def __init__(self, device):
        self.device = device
        self._b: Dict[Tuple, _SizedStub] = {}
        self.total = 0
        self.ranges: list = []             # (fake offset, bytes, key, zero-initialised)

    def get(self, name: str, numel: int, dtype: torch.dtype, zero: bool = False):
        key = (name, numel, dtype)
        t = self._b.get(key)
        if t is None:
            # every stub gets its own range of a fake address space, so that plan_memory() can tell from the pointers
            # bound into a plan which buffers each op touches
            t = self._b[key] = _SizedStub(numel, dtype, self.total)
            a = PoolBuffers.ALIGN
            nbytes = (numel * t.element_size() + a - 1) // a * a
            self.ranges.append((self.total, nbytes, key, bool(zero)))
            self.total += nbytes
        return t

    def bytes(self) -> int:
        return 0

    def drop(self, tags) -> None:
        pass


def plan_memory(plans, counting: "CountingBuffers", pinned=()):
    """Liveness-based layout of the buffers of forward-only plan(s) built over `counting`: a buffer is live from the first
    to the last op that is bound to a pointer inside it (`pinned` stubs -- inputs written before the plan runs, results
    read after it -- are live throughout), and buffers whose live ranges do not overlap share memory.  Greedy by size,
    lowest fitting offset.  Returns ({buffer key: byte offset}, total bytes)."""
    import bisect
    starts = [r[0] for r in counting.ranges]
    base = _SizedStub.FAKE_BASE
    first, last = {}, {}

    def touch(ptr, i):
        if not isinstance(ptr, int) or ptr < base or ptr >= base + counting.total:
            return
        k = bisect.bisect_right(starts, ptr - base) - 1
        first.setdefault(k, i)
        last[k] = i

    i = 0
    for plan in plans:
        for fn, args, _ in plan.ops:
            for a in (args if fn is not None else ()):
                obj = getattr(a, "_obj", None)
                if isinstance(obj, L.ConvArgs):
                    for f in ("x", "w", "bias", "y", "stat_partials", "mean_rstd", "tile_counters", "stat_acc"):
                        touch(getattr(obj, f), i)
                else:
                    touch(a, i)
            i += 1
    n_ops = i
    pinned_ids = set()
    for t in pinned:
        if isinstance(t, _SizedStub):
            pinned_ids.add(bisect.bisect_right(starts, t.data_ptr() - base) - 1)
    items = []
    for k, (off, nbytes, key, zero) in enumerate(counting.ranges):
        if k in pinned_ids or zero or k not in first:
            lo, hi = -1, n_ops            # external, persistent or never seen in an op: keep it alive throughout
        else:
            lo, hi = first[k], last[k]
        items.append((nbytes, lo, hi, key))
    placed, offsets, total = [], {}, 0
    for nbytes, lo, hi, key in sorted(items, key=lambda t: -t[0]):
        busy = sorted((o, o + n) for (o, n, l2, h2) in placed if not (h2 < lo or hi < l2))
        off = 0
        for o, e in busy:
            if off + nbytes <= o:
                break
            off = max(off, e)
        placed.append((off, nbytes, lo, hi))
        offsets[key] = off
        total = max(total, off + nbytes)
    return offsets, total


class PlannedBuffers:
    """``Buffers`` interface over one allocation with the byte offsets computed by ``plan_memory``."""

    def __init__(self, offsets: dict, total: int, device):
        self.device, self.offsets, self.total = device, offsets, total
        self.t = torch.empty(max(total, 256), dtype=torch.uint8, device=device)
        self._b: Dict[Tuple, torch.Tensor] = {}

    def get(self, name: str, numel: int, dtype: torch.dtype, zero: bool = False) -> torch.Tensor:
        key = (name, numel, dtype)
        t = self._b.get(key)
        if t is None:
            off = self.offsets[key]
            esz = torch.empty(0, dtype=dtype).element_size()
            t = self._b[key] = self.t[off:off + numel * esz].view(dtype)
            if zero:
                t.zero_()
        return t

    def bytes(self) -> int:
        return self.total

    def drop(self, tags) -> None:
        pass


class SliceBuffers:
    """View of a parent ``Buffers`` for one of ``parts`` equal batch slices: every request must name a buffer the parent
    already holds at ``parts`` times the size (a [B][...] buffer of the full-batch graph) and gets its slice.  Used to
    compile half-batch forward plans that write the full-batch graph's activations in place."""

    def __init__(self, parent: Buffers, part: int, parts: int):
        self.parent, self.part, self.parts, self.device = parent, part, parts, parent.device

    def get(self, name: str, numel: int, dtype: torch.dtype, zero: bool = False) -> torch.Tensor:
        full = self.parent._b.get((name, numel * self.parts, dtype))
        if full is None:
            raise KeyError(f"buffer {name!r} ({numel} x {self.parts}) is not a per-image buffer of the full-batch graph")
        return full[self.part * numel:(self.part + 1) * numel]

    def bytes(self) -> int:
        return 0


@dataclass
class ActBuf:
    """Haloed NHWC activation buffer [B][H+2p][W+2p][C]."""
    t: torch.Tensor
    B: int
    H: int
    W: int
    C: int
    pad: int

    def view(self) -> torch.Tensor:
        return self.t.view(self.B, self.H + 2 * self.pad, self.W + 2 * self.pad, self.C)

    def interior_nchw(self) -> torch.Tensor:
        v = self.view()
        p = self.pad
        if p:
            v = v[:, p:-p, p:-p, :]
        return v.permute(0, 3, 1, 2).float()


# when a list, Plan.run records (label, start_event, end_event) per op into it (tools/bench_train.py --profile)
PROFILE = [None]
# inference plans are replayed as CUDA graphs (NIRGAN_B200_GRAPH=0 launches every kernel from the host instead)
GRAPHS = [os.environ.get("NIRGAN_B200_GRAPH", "1") != "0"]
# ops added with side=True (the weight gradients: off the critical path of a backward pass) run on a second stream so
# that they overlap the HBM-bound norm-backward kernels of the next layer (NIRGAN_B200_SIDE_STREAM=0: single stream)
SIDE_STREAM = [os.environ.get("NIRGAN_B200_SIDE_STREAM", "1") != "0"]
# training plans (forward and backward of both networks) are replayed as CUDA graphs too: the step is ~250 launches and
# the host needs longer to enqueue them one by one than the GPU needs to run them (NIRGAN_B200_TRAIN_GRAPH=0: eager)
TRAIN_GRAPHS = [os.environ.get("NIRGAN_B200_TRAIN_GRAPH", "1") != "0"]
# training contexts of all input shapes share one memory pool per runner (NIRGAN_B200_POOL=0: one set of buffers per shape)
POOL = [os.environ.get("NIRGAN_B200_POOL", "1") != "0"]
_SIDE: Dict[int, "torch.cuda.Stream"] = {}


class Plan:
    """A compiled list of bound C-ABI calls (plus optional host hooks between them)."""

    def __init__(self):
        self.ops: List[Tuple[Callable, tuple, str]] = []
        self.keepalive: list = []
        self.launches = 0          # kernel launches per run (for bench.py's gpu_launches)
        self.records: dict = {}
        self.labels: List[str] = []
        self.side: List[bool] = []
        self._events: list = []
        self.hook_fn: Optional[Callable] = None     # called as hook_fn(payload, main_stream, side_stream_or_None)
        self._graphs: Optional[list] = None
        self._graph_sig = None

    def add(self, fn_name: str, *args, launches: int = 1, label: str = "", side: bool = False):
        fn = getattr(L.load(), fn_name)
        self.ops.append((fn, args, fn_name))
        self.labels.append(label or fn_name)
        self.side.append(bool(side))
        self.launches += launches

    def add_hook(self, payload, label: str = "hook"):
        """A host-side call-out at this point of the plan (used by the data-parallel gradient exchange: "everything the
        plan has exported so far is final").  Side-stream work issued before the hook is flushed first and the hook is
        handed both streams, so it can order a collective behind either.  Hooks split a graph-replayed plan into one CUDA
        graph per segment."""
        self.ops.append((None, payload, "hook"))
        self.labels.append(label)
        self.side.append(False)

    def _segments(self):
        segs, cur = [], []
        for i, (fn, _, _) in enumerate(self.ops):
            if fn is None:
                segs.append(("ops", cur)); segs.append(("hook", i)); cur = []
            else:
                cur.append(i)
        segs.append(("ops", cur))
        return [s for s in segs if s[0] == "hook" or s[1]]

    def _arg_signature(self):
        """Values of the mutable ctypes scalars among the bound arguments (accumulate flags of the gradient exports):
        a captured graph bakes them in, so a change forces eager launches for that run."""
        sig = []
        for fn, args, _ in self.ops:
            if fn is None:
                continue
            for a in args:
                if isinstance(a, C.c_float):
                    sig.append(a.value)
        return tuple(sig)

    def run_graphed(self, stream: "torch.cuda.Stream"):
        """Replay the plan as CUDA graphs on `stream` (captured on the second call; the first runs eagerly so that
        one-off host work -- kernel attributes, TMA descriptor encoding -- happens outside the capture).  All buffers are
        static and the packed weights are refreshed in place, so the captured graphs stay valid.  Host hooks run between
        the graphs of the segments they separate."""
        if PROFILE[0] is not None or not GRAPHS[0]:
            return self.run(stream.cuda_stream)
        if self._graphs is None:
            if not getattr(self, "_warm", False):
                self._warm = True
                return self.run(stream.cuda_stream)
            try:
                graphs = []
                for kind, body in self._segments():
                    if kind == "hook":
                        graphs.append(("hook", body))
                        continue
                    g = torch.cuda.CUDAGraph()
                    # captured on torch's private capture stream (the caller's may be the legacy default stream, where
                    # capture is illegal); replay() enqueues the graph on whatever stream is current
                    with torch.cuda.graph(g, capture_error_mode="thread_local"):
                        self._run_ops(body, torch.cuda.current_stream().cuda_stream)
                    graphs.append(("graph", g))
                self._graphs, self._graph_sig = graphs, self._arg_signature()
            except Exception as e:          # capture is an optimisation: fall back to eager launches, loudly, once
                import warnings
                warnings.warn(f"nirgan_b200: CUDA graph capture failed ({e}); launching kernels eagerly")
                GRAPHS[0] = False
                return self.run(stream.cuda_stream)
        if self._graph_sig != self._arg_signature():
            return self.run(stream.cuda_stream)
        for kind, g in self._graphs:
            if kind == "hook":
                self._call_hook(g, stream.cuda_stream, None)
            else:
                g.replay()

    def run_training(self, device):
        """Training plans: CUDA-graph replay (NIRGAN_B200_TRAIN_GRAPH=0: eager launches)."""
        st = torch.cuda.current_stream(device)
        if TRAIN_GRAPHS[0]:
            return self.run_graphed(st)
        return self.run(st.cuda_stream)

    def run(self, stream_ptr: int):
        if PROFILE[0] is not None:
            return self._run_profiled(stream_ptr)
        self._run_ops(range(len(self.ops)), stream_ptr)

    def _call_hook(self, i: int, stream_ptr: int, side):
        if self.hook_fn is not None:
            dev = torch.cuda.current_device()
            main = torch.cuda.ExternalStream(stream_ptr, device=dev) if stream_ptr else torch.cuda.default_stream(dev)
            self.hook_fn(self.ops[i][1], main, side)

    def _run_ops(self, idx, stream_ptr: int):
        idx = list(idx)
        if SIDE_STREAM[0] and any(self.side[i] for i in idx):
            return self._run_two_streams(idx, stream_ptr)
        for i in idx:
            fn, args, name = self.ops[i]
            if fn is None:
                self._call_hook(i, stream_ptr, None)
                continue
            st = fn(*args, stream_ptr)
            if st != 0:
                L.check(st, name)

    def _run_two_streams(self, idx, stream_ptr: int):
        """Side ops depend on everything issued before them on the main stream and on earlier side ops; nothing on the
        main stream depends on them until the plan (segment) ends (they only write their own weight-gradient buffers,
        the gradient arena and a workspace that side ops share, serialised by the side stream).  Each side op is
        submitted AFTER the next main op, so the main op takes the SMs first and the side op then overlaps what follows."""
        dev = torch.cuda.current_device()
        side = _SIDE.get(dev)
        if side is None:
            # NIRGAN_B200_SIDE_PRIORITY=-1: high-priority side stream (experiment; default: same priority as the caller's)
            side = _SIDE[dev] = torch.cuda.Stream(device=dev, priority=int(os.environ.get("NIRGAN_B200_SIDE_PRIORITY", "0")))
        main = torch.cuda.ExternalStream(stream_ptr, device=dev) if stream_ptr else torch.cuda.default_stream(dev)
        n_side = sum(1 for i in idx if self.side[i])
        while len(self._events) < n_side + 1:
            self._events.append(torch.cuda.Event())
        side_ptr = side.cuda_stream
        pending = []
        k = 0
        last_side_main = False
        for i in idx:
            fn, args, name = self.ops[i]
            if fn is None:
                for op in pending:
                    self._launch(op, side_ptr)
                pending = []
                self._call_hook(i, stream_ptr, side)
                continue
            if self.side[i]:
                if not last_side_main:
                    # a run of consecutive side ops (weight gradient + its exports) shares one fork point
                    for op in pending:
                        self._launch(op, side_ptr)
                    pending = []
                    ev = self._events[k]
                    k += 1
                    ev.record(main)
                    side.wait_event(ev)
                pending.append((fn, args, name))
                last_side_main = True
                continue
            last_side_main = False
            st = fn(*args, stream_ptr)
            if st != 0:
                L.check(st, name)
            for op in pending:
                self._launch(op, side_ptr)
            pending = []
        for op in pending:
            self._launch(op, side_ptr)
        ev = self._events[n_side]
        ev.record(side)
        main.wait_event(ev)

    @staticmethod
    def _launch(op, stream_ptr: int):
        fn, args, name = op
        st = fn(*args, stream_ptr)
        if st != 0:
            L.check(st, name)


def _run_profiled(self, stream_ptr: int):
    ev = torch.cuda.Event(enable_timing=True)
    ev.record()
    for i, ((fn, args, name), label) in enumerate(zip(self.ops, self.labels)):
        if fn is None:
            self._call_hook(i, stream_ptr, None)
            continue
        st = fn(*args, stream_ptr)
        if st != 0:
            L.check(st, name)
        ev2 = torch.cuda.Event(enable_timing=True)
        ev2.record()
        PROFILE[0].append((label, name, ev, ev2))
        ev = ev2


Plan._run_profiled = _run_profiled


class Engine:
    """Shared machinery for the generator / discriminator runners."""

    def __init__(self, cfg: EngineConfig, device: torch.device):
        L.load()
        self.cfg = cfg
        self.device = device
        self.dt_enum, self.dt_torch = _PRECISIONS[cfg.precision]
        self.impl = L.IMPL_TC if cfg.impl == "tc" else L.IMPL_SIMT
        self.buffers = Buffers(device)
        self._packed: Dict[Tuple, Tuple[int, torch.Tensor]] = {}

    # ---- weights -----------------------------------------------------------------------------
    def packed_weight(self, w: torch.Tensor, n_axis: int, n_pad: int, k_pad: int, stream: int) -> torch.Tensor:
        """fp32 master (4-D) -> packed shadow, refreshed when the master changes.
        n_axis 0/1: [tap][n_pad][k_pad];  n_axis 'rowmerged': [kh][O][kw*8+c] (stem);
        n_axis 'taps': [n = tap (padded to n_pad)][k_pad] for the tap-GEMM form of a 1-output-channel conv,
        'taps_T': its transpose [n = channel][k = tap] for the data gradient."""
        key = (w.data_ptr(), n_axis, n_pad, k_pad, self.dt_enum)
        ver = (w._version, WEIGHT_EPOCH[0])
        hit = self._packed.get(key)
        if hit is not None and hit[0] == ver:
            return hit[1]
        d0, d1, kh, kw = w.shape
        src = w.detach()
        if src.dtype != torch.float32 or not src.is_contiguous():
            src = src.float().contiguous()
        if n_axis in ("rowmerged", "rowmerged4"):
            # [kh][O][kw*cs + c]: cs = 8 (64-wide rows, the ng_prep_stem layout) or 4 (32-wide, ng_stem_conv)
            cs = 8 if n_axis == "rowmerged" else 4
            dst = hit[1] if hit is not None else torch.empty(kh * d0 * 8 * cs, dtype=self.dt_torch, device=self.device)
            L.call("ng_pack_weight_rowmerged", src.data_ptr(), d0, d1, kh, kw, cs, self.dt_enum, dst.data_ptr(), stream)
        elif n_axis in ("s2d", "s2d_T"):
            # PatchGAN input layer in space-to-depth form: (O, I, 4, 4) -> [tap (4)][O][64] / transposed [tap][64][O]
            assert kh == 4 and kw == 4 and d1 <= 16
            dst = hit[1] if hit is not None else torch.empty(4 * d0 * 64, dtype=self.dt_torch, device=self.device)
            L.call("ng_pack_weight_s2d", src.data_ptr(), d0, d1, 1 if n_axis == "s2d_T" else 0, self.dt_enum,
                   dst.data_ptr(), stream)
        elif n_axis == "phasemerged":
            # ConvTranspose2d (Cin, Cout, 3, 3) -> [shift (4)][phase*Cout + co][Cin]  (NG_FORM_PHASED_MERGED)
            assert kh == 3 and kw == 3 and n_pad == d1 and k_pad == d0
            dst = hit[1] if hit is not None else torch.empty(16 * d1 * d0, dtype=self.dt_torch, device=self.device)
            L.call("ng_pack_weight_phasemerged", src.data_ptr(), d0, d1, self.dt_enum, dst.data_ptr(), stream)
        elif n_axis == "taps":
            assert d0 == 1 and kh * kw <= n_pad
            dst = hit[1] if hit is not None else torch.zeros(n_pad * k_pad, dtype=self.dt_torch, device=self.device)
            L.call("ng_pack_weight", src.data_ptr(), d0, d1, kh, kw, 0, 1, k_pad, self.dt_enum, dst.data_ptr(), stream)
        elif n_axis == "taps_T":
            # [n = input channel (n_pad)][k = tap (k_pad)]: (1, C, kh, kw) is already a [C][kh*kw] matrix
            assert d0 == 1 and kh * kw <= k_pad and d1 <= n_pad
            dst = hit[1] if hit is not None else torch.empty(n_pad * k_pad, dtype=self.dt_torch, device=self.device)
            L.call("ng_pack_weight", src.data_ptr(), d1, kh * kw, 1, 1, 0, n_pad, k_pad, self.dt_enum, dst.data_ptr(),
                   stream)
        else:
            dst = hit[1] if hit is not None else torch.empty(kh * kw * n_pad * k_pad, dtype=self.dt_torch,
                                                             device=self.device)
            L.call("ng_pack_weight", src.data_ptr(), d0, d1, kh, kw, n_axis, n_pad, k_pad, self.dt_enum,
                   dst.data_ptr(), stream)
        self._packed[key] = (ver, dst)
        return dst

    # ---- plan building blocks -------------------------------------------------------------------
    def act(self, name: str, B: int, H: int, W: int, Cn: int, pad: int, dtype=None) -> ActBuf:
        dtype = dtype or self.dt_torch
        n = B * (H + 2 * pad) * (W + 2 * pad) * Cn
        return ActBuf(self.buffers.get(name, n, dtype), B, H, W, Cn, pad)

    def conv_args(self, x: ActBuf, w: torch.Tensor, y: torch.Tensor, Cout: int, K: int, stride: int, pad: int,
                  Hout: int, Wout: int, form=L.FORM_GATHER, sgn=1, epilogue=L.EPI_RAW, act=L.ACT_NONE, slope=0.0,
                  crop=0, bias: Optional[torch.Tensor] = None, partials: Optional[torch.Tensor] = None,
                  impl: Optional[int] = None, KW: Optional[int] = None, pad_w: Optional[int] = None,
                  in_pad_w: Optional[int] = None) -> L.ConvArgs:
        a = L.ConvArgs()
        a.dtype, a.impl, a.form, a.sgn = self.dt_enum, self.impl if impl is None else impl, form, sgn
        a.B, a.Hin, a.Win, a.Cin, a.in_pad = x.B, x.H, x.W, x.C, x.pad
        a.in_pad_w = x.pad if in_pad_w is None else in_pad_w
        a.Cout, a.KH, a.KW, a.stride, a.pad = Cout, K, (K if KW is None else KW), stride, pad
        a.pad_w = pad if pad_w is None else pad_w
        a.Hout, a.Wout = Hout, Wout
        a.epilogue, a.act, a.slope, a.crop = epilogue, act, slope, crop
        a.x, a.w, a.bias, a.y = x.t.data_ptr(), w.data_ptr(), _ptr(bias), y.data_ptr()
        a.stat_partials = _ptr(partials)
        return a


def conv_out(h: int, k: int, s: int, p: int) -> int:
    return (h + 2 * p - k) // s + 1
