"""Adam on the ng_adam_* kernels (torch.optim.Adam semantics: eps 1e-8, no weight decay, bias-corrected;
model/pix2pix.py:486-487) and the data-parallel gradient exchange of the DDP path (train.py:118-120).

Gradients of all parameters of a network live in ONE flat fp32 arena (``GradArena``): the backward plans of ``graph.py``
export every weight gradient straight into its slot and autograd adopts views of the slots as ``.grad``.  One optimizer
step is then three launches -- non-finite scan, step-counter advance, multi-tensor update -- with no host
synchronisation (a non-finite gradient turns the update into a no-op on the device, like torch.cuda.amp.GradScaler).
Data parallel: ``BucketedAllReduce`` sums contiguous ranges of the arena over the ranks on a communication stream as soon
as the backward pass has produced them (reverse parameter order, like torch DDP's buckets), so the exchange overlaps the
rest of the backward pass; the 1/world factor of the mean rides on the optimizer's ``grad_scale`` instead of a separate
division kernel.
"""
from __future__ import annotations

from typing import Iterable, List, Optional

import torch

from . import _lib as L


class GradArena:
    """Flat gradient storage of one network (16-byte aligned slots, parameter order)."""

    def __init__(self, params: List[torch.nn.Parameter]):
        dev = params[0].device
        offs, total = [], 0
        for p in params:
            offs.append(total)
            total += (p.numel() + 3) // 4 * 4
        self.params, self.offsets, self.total = params, offs, total
        self.numels = [p.numel() for p in params]
        self.grad = torch.zeros(total, dtype=torch.float32, device=dev)
        self.index = {id(p): i for i, p in enumerate(params)}
        for p, o in zip(params, offs):
            p._b200_grad_slot = self.grad[o:o + p.numel()].view_as(p)
            p._b200_grad_arena = self

    @staticmethod
    def of(params: Iterable[torch.nn.Parameter]) -> Optional["GradArena"]:
        """The arena shared by exactly these parameters (created on first use), or None when they cannot share one
        (CPU / mixed devices / not fp32-contiguous)."""
        ps = list(params)
        if not ps or not all(p.is_cuda and p.dtype == torch.float32 and p.is_contiguous() for p in ps) or \
                len({p.device for p in ps}) != 1:
            return None
        ar = getattr(ps[0], "_b200_grad_arena", None)
        if ar is not None and len(ar.params) == len(ps) and all(a is b for a, b in zip(ar.params, ps)) and \
                ar.grad.device == ps[0].device:
            return ar
        return GradArena(ps)

    def offset_of(self, p: torch.nn.Parameter) -> int:
        return self.offsets[self.index[id(p)]]

    def grads_in_place(self) -> bool:
        """True when every parameter's .grad is (a view of) its arena slot."""
        for p in self.params:
            g = p.grad
            if g is None or g.data_ptr() != p._b200_grad_slot.data_ptr() or not g.is_contiguous():
                return False
        return True


class _AdamState:
    """Moments, device-side step counter and pointer tables of one optimizer over a GradArena."""

    def __init__(self, ar: GradArena):
        dev = ar.grad.device
        self.m = torch.zeros(ar.total, dtype=torch.float32, device=dev)
        self.v = torch.zeros(ar.total, dtype=torch.float32, device=dev)
        self.offs_dev = torch.tensor(ar.offsets, dtype=torch.int64, device=dev)
        self.numel_dev = torch.tensor(ar.numels, dtype=torch.int64, device=dev)
        self.ptrs_dev = None
        self.ptrs_host: List[int] = []
        self.flag = torch.zeros(1, dtype=torch.int32, device=dev)
        self.step_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        self.skipped_dev = torch.zeros(1, dtype=torch.int32, device=dev)

    def refresh_pointers(self, ar: GradArena):
        ptrs = [p.data_ptr() for p in ar.params]
        if ptrs != self.ptrs_host:
            self.ptrs_host = ptrs
            self.ptrs_dev = torch.tensor(ptrs, dtype=torch.int64, device=ar.grad.device)


class B200Adam(torch.optim.Optimizer):
    def __init__(self, params, lr=2e-4, betas=(0.5, 0.999), eps=1e-8):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))
        self.skip_nonfinite = True
        self._skipped_host = 0
        self._arena: Optional[GradArena] = None
        self._st: Optional[_AdamState] = None
        self._host_steps = 0
        self.fast_steps = 0            # steps taken by the three-launch arena path (diagnostic)
        self.arena()                   # parameters already on the GPU: gradients land in the arena from the first backward

    # ---- flat arena -----------------------------------------------------------------------------------
    def arena(self) -> Optional[GradArena]:
        """Built lazily once every parameter lives on one CUDA device (single param group, fp32, contiguous)."""
        if self._arena is None and len(self.param_groups) == 1:
            ar = GradArena.of(self.param_groups[0]["params"])
            if ar is not None:
                self._arena, self._st = ar, _AdamState(ar)
                self._bind_state()
        return self._arena

    def _bind_state(self):
        """torch-style per-parameter state: views of the flat moment arenas (what state_dict() serialises)."""
        ar, st = self._arena, self._st
        for p, o in zip(ar.params, ar.offsets):
            self.state[p] = {"step": self._host_steps, "exp_avg": st.m[o:o + p.numel()].view_as(p),
                             "exp_avg_sq": st.v[o:o + p.numel()].view_as(p)}

    def load_state_dict(self, state_dict):
        """Checkpoint resume (Lightning ``resume_from_checkpoint``, reference train.py:67-69,126): the inherited loader
        replaces ``self.state`` with fresh tensors; the fast path reads the flat arenas, so the loaded moments and step
        count are copied into them and the per-parameter views are re-bound."""
        super().load_state_dict(state_dict)
        ar = self.arena()
        if ar is None:
            return
        steps = 0
        for p, o in zip(ar.params, ar.offsets):
            s = self.state.get(p)
            if not s:
                continue
            if "exp_avg" in s:
                self._st.m[o:o + p.numel()].copy_(s["exp_avg"].reshape(-1).to(self._st.m))
                self._st.v[o:o + p.numel()].copy_(s["exp_avg_sq"].reshape(-1).to(self._st.v))
            steps = max(steps, int(s.get("step", 0)))
        self._host_steps = steps
        self._st.step_dev.fill_(steps)
        self._bind_state()

    def __setstate__(self, state):
        super().__setstate__(state)
        # un-pickled optimizers come back without device arenas: rebuild them from the per-parameter state
        self._arena, self._st = None, None
        saved = {p: dict(s) for p, s in self.state.items()}
        if self.arena() is not None:
            steps = 0
            for p, o in zip(self._arena.params, self._arena.offsets):
                s = saved.get(p)
                if s and "exp_avg" in s:
                    self._st.m[o:o + p.numel()].copy_(s["exp_avg"].reshape(-1))
                    self._st.v[o:o + p.numel()].copy_(s["exp_avg_sq"].reshape(-1))
                    steps = max(steps, int(s.get("step", 0)))
            self._host_steps = steps
            self._st.step_dev.fill_(steps)
            self._bind_state()

    @property
    def skipped_steps(self) -> int:
        """Number of updates skipped because a gradient was not finite (reads the device counter: synchronises)."""
        n = self._skipped_host
        if self._st is not None:
            n += int(self._st.skipped_dev.item())
        return n

    @torch.no_grad()
    def step(self, closure=None, grad_scale: float = 1.0):
        """``grad_scale`` multiplies every gradient inside the update kernel (1/world_size turns summed DDP gradients
        into their mean without a separate division pass)."""
        loss = closure() if closure is not None else None
        from . import engine
        ar = self.arena()
        if ar is not None and ar.grads_in_place():
            st_ = self._st
            group = self.param_groups[0]
            b1, b2 = group["betas"]
            st = torch.cuda.current_stream(ar.grad.device).cuda_stream
            st_.refresh_pointers(ar)
            flag = None
            if self.skip_nonfinite:
                L.call("ng_nonfinite_flag", ar.grad.data_ptr(), ar.total, st_.flag.data_ptr(), st)
                st_.skipped_dev.add_(st_.flag)
                flag = st_.flag.data_ptr()
            L.call("ng_adam_multi", st_.ptrs_dev.data_ptr(), st_.offs_dev.data_ptr(), st_.numel_dev.data_ptr(),
                   len(ar.params), ar.grad.data_ptr(), st_.m.data_ptr(), st_.v.data_ptr(), ar.total, float(group["lr"]),
                   float(b1), float(b2), float(group["eps"]), st_.step_dev.data_ptr(), float(grad_scale), flag, st)
            self._host_steps += 1      # counts calls; updates the device skipped are in skipped_steps
            self.fast_steps += 1
            for p in ar.params:
                p._b200_epoch = engine.WEIGHT_EPOCH[0] + 1
                self.state[p]["step"] = self._host_steps
            engine.WEIGHT_EPOCH[0] += 1
            return loss
        return self._step_per_parameter(loss, grad_scale)

    def _step_per_parameter(self, loss, grad_scale):
        """Gradients that did not come through the arena (user-assigned .grad, CPU tests of the host logic ...)."""
        from . import engine
        grads = [p.grad for g in self.param_groups for p in g["params"] if p.grad is not None]
        if grads and self.skip_nonfinite:
            bad = torch.stack([(~torch.isfinite(g)).any() for g in grads]).any()
            if bool(bad):
                self._skipped_host += 1
                return loss
        for group in self.param_groups:
            b1, b2 = group["betas"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not p.is_cuda:
                    raise RuntimeError("nirgan_b200 B200Adam: parameters must live on a B200 (no CPU fallback)")
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, dtype=torch.float32)
                    st["exp_avg_sq"] = torch.zeros_like(p, dtype=torch.float32)
                st["step"] = int(st["step"]) + 1
                g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                p._b200_epoch = engine.WEIGHT_EPOCH[0] + 1      # this parameter's storage changes in this step
                L.call("ng_adam_step", p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(),
                       st["exp_avg_sq"].data_ptr(), p.numel(), float(group["lr"]), float(b1), float(b2),
                       float(group["eps"]), int(st["step"]), float(grad_scale),
                       torch.cuda.current_stream(p.device).cuda_stream)
        if self._st is not None:         # keep the device-side step counter of the arena path in sync
            self._host_steps += 1
            self._st.step_dev.add_(1)
        engine.WEIGHT_EPOCH[0] += 1     # masters changed in place: packed low-precision shadows must be refreshed
        return loss


# =====================================================================================================================
# data parallel
# =====================================================================================================================
class BucketedAllReduce:
    """Sum-all-reduce of a GradArena in contiguous ranges ("buckets"), each launched on a communication stream the
    moment the backward pass has finished writing it -- ``ready(lo)`` is called by the backward plan's hooks with the
    arena offset below which gradients are still outstanding, so the buckets go out in reverse parameter order while the
    remaining layers are still being differentiated.  ``finish()`` sends what is left, makes the caller's stream wait
    for every bucket and returns; the mean's 1/world factor is NOT applied here (``B200Adam.step(grad_scale=1/world)``).

    DDP semantics of the reference (Lightning ``strategy='ddp'``, train.py:118-120): after ``finish`` every rank holds
    the same summed gradients.  Backend NCCL on GPUs (gloo in the CPU tests, where streams do not exist)."""

    def __init__(self, arena_grad: torch.Tensor, group=None, min_bucket_elems: int = 1 << 18):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.flat = arena_grad
        self.min_elems = int(min_bucket_elems)
        self.cuda = arena_grad.is_cuda
        self.comm = torch.cuda.Stream(arena_grad.device) if self.cuda else None
        self.last_buckets = 0             # collectives launched by the last finish()ed pass
        self.reset()
        # device-side measurement of the exposed (non-overlapped) wait in finish()
        self.timing = False
        self._t0 = self._t1 = None

    def active(self) -> bool:
        return self.dist.is_available() and self.dist.is_initialized() and self.dist.get_world_size(self.group) > 1

    def reset(self):
        self.hi = self.flat.numel()       # everything in [hi, end) has been sent
        self.works = []
        self.buckets = 0

    def _send(self, lo: int, producer_stream: Optional["torch.cuda.Stream"]):
        if lo >= self.hi:
            return
        chunk = self.flat[lo:self.hi]
        self.hi = lo
        self.buckets += 1
        if not self.cuda:
            self.works.append(self.dist.all_reduce(chunk, op=self.dist.ReduceOp.SUM, group=self.group, async_op=True))
            return
        ev = torch.cuda.Event()
        ev.record(producer_stream if producer_stream is not None else torch.cuda.current_stream(self.flat.device))
        self.comm.wait_event(ev)
        with torch.cuda.stream(self.comm):
            self.works.append(self.dist.all_reduce(chunk, op=self.dist.ReduceOp.SUM, group=self.group, async_op=True))

    def ready(self, lo: int, producer_stream=None):
        """Gradients at arena offsets >= lo are final (written on `producer_stream`, default: the current stream)."""
        if not self.active():
            return
        if self.hi - lo >= self.min_elems:
            self._send(lo, producer_stream)

    def finish(self, producer_stream=None):
        if not self.active():
            self.reset()
            return
        self._send(0, producer_stream)
        if self.cuda and self.timing:
            main = torch.cuda.current_stream(self.flat.device)
            self._t0 = torch.cuda.Event(enable_timing=True)
            self._t1 = torch.cuda.Event(enable_timing=True)
            self._t0.record(main)
        for w in self.works:
            w.wait()                      # the caller's stream waits for the collective (no host block on CUDA)
        if self.cuda and self.timing:
            self._t1.record(torch.cuda.current_stream(self.flat.device))
        n = self.last_buckets = self.buckets
        self.reset()
        return n

    def exposed_ms(self) -> float:
        """Time the caller's stream spent waiting in the last finish() (after a device synchronise)."""
        if self._t0 is None:
            return 0.0
        return self._t0.elapsed_time(self._t1)


def allreduce_gradients(params: Iterable[torch.nn.Parameter], world_size: Optional[int] = None, group=None,
                        average: bool = True) -> None:
    """Blocking form: all-reduce the gradients of `params` in ONE collective (on the arena itself when the gradients
    live in a GradArena) and, when ``average``, divide by the world size.  ``trainer.Trainer`` uses the bucketed,
    overlapped ``BucketedAllReduce`` with ``average`` folded into the optimizer step instead; this entry point serves
    callers that drive the optimizers themselves.  Backend: NCCL on GPUs, gloo in the CPU tests."""
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized():
        return
    ws = world_size or dist.get_world_size(group)
    if ws == 1:
        return
    ps = [p for p in params if p.grad is not None]
    if not ps:
        return
    ar = getattr(ps[0], "_b200_grad_arena", None)
    if ar is not None and len(ps) == len(ar.params) and all(getattr(p, "_b200_grad_arena", None) is ar for p in ps) \
            and ar.grads_in_place():
        dist.all_reduce(ar.grad, op=dist.ReduceOp.SUM, group=group)
        if average:
            ar.grad.mul_(1.0 / ws)
        return
    flat = torch.cat([p.grad.reshape(-1) for p in ps])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat.mul_(1.0 / ws)
    off = 0
    for p in ps:
        n = p.numel()
        p.grad.copy_(flat[off:off + n].view_as(p.grad))
        off += n
