"""Adam on the ng_adam_* kernels (torch.optim.Adam semantics: eps 1e-8, no weight decay, bias-corrected;
model/pix2pix.py:486-487) and the data-parallel gradient all-reduce of the DDP path (train.py:118-120).

Gradients of all parameters of an optimizer live in ONE flat fp32 arena (the autograd bridges of ``train.py`` export
weight gradients straight into it and hand autograd views of it), with flat moment arenas beside it.  One optimizer
step is then three launches -- non-finite scan, step-counter advance, multi-tensor update -- instead of several per
parameter, with no host synchronisation (a non-finite gradient turns the update into a no-op on the device, like
torch.cuda.amp.GradScaler), and the DDP all-reduce is a single NCCL call on the arena (no concatenation or copy-back).
"""
from __future__ import annotations

from typing import Iterable, List, Optional

import torch

from . import _lib as L


class _Arena:
    """Flat gradient / moment storage of one optimizer."""

    def __init__(self, params: List[torch.nn.Parameter]):
        dev = params[0].device
        offs, total = [], 0
        for p in params:
            offs.append(total)
            total += (p.numel() + 3) // 4 * 4            # 16-byte aligned slots
        self.params, self.offsets, self.total = params, offs, total
        self.grad = torch.zeros(total, dtype=torch.float32, device=dev)
        self.m = torch.zeros(total, dtype=torch.float32, device=dev)
        self.v = torch.zeros(total, dtype=torch.float32, device=dev)
        self.offs_dev = torch.tensor(offs, dtype=torch.int64, device=dev)
        self.ptrs_dev = None
        self.ptrs_host: List[int] = []
        self.flag = torch.zeros(1, dtype=torch.int32, device=dev)
        self.step_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        self.skipped_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        for p, o in zip(params, offs):
            p._b200_grad_slot = self.grad[o:o + p.numel()].view_as(p)
            p._b200_grad_arena = self
        self.refresh_pointers()

    def refresh_pointers(self):
        ptrs = [p.data_ptr() for p in self.params]
        if ptrs != self.ptrs_host:
            self.ptrs_host = ptrs
            self.ptrs_dev = torch.tensor(ptrs, dtype=torch.int64, device=self.grad.device)

    def slot(self, p, o):
        return self.grad[o:o + p.numel()]

    def grads_in_place(self) -> bool:
        """True when every parameter's .grad is (a view of) its arena slot."""
        for p in self.params:
            g = p.grad
            if g is None or g.data_ptr() != p._b200_grad_slot.data_ptr() or not g.is_contiguous():
                return False
        return True


class B200Adam(torch.optim.Optimizer):
    def __init__(self, params, lr=2e-4, betas=(0.5, 0.999), eps=1e-8):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))
        self.skip_nonfinite = True
        self._skipped_host = 0
        self._arena: Optional[_Arena] = None
        self._host_steps = 0
        self.fast_steps = 0            # steps taken by the three-launch arena path (diagnostic)
        self.arena()                   # parameters already on the GPU: gradients land in the arena from the first backward

    # ---- flat arena -----------------------------------------------------------------------------------
    def arena(self) -> Optional[_Arena]:
        """Built lazily once every parameter lives on one CUDA device (single param group, fp32, contiguous)."""
        if self._arena is None and len(self.param_groups) == 1:
            ps = list(self.param_groups[0]["params"])
            if ps and all(p.is_cuda and p.dtype == torch.float32 and p.is_contiguous() for p in ps) and \
                    len({p.device for p in ps}) == 1:
                self._arena = _Arena(ps)
                for p, o in zip(ps, self._arena.offsets):         # torch-style per-parameter state views
                    self.state[p] = {"step": 0, "exp_avg": self._arena.m[o:o + p.numel()].view_as(p),
                                     "exp_avg_sq": self._arena.v[o:o + p.numel()].view_as(p)}
        return self._arena

    @property
    def skipped_steps(self) -> int:
        """Number of updates skipped because a gradient was not finite (reads the device counter: synchronises)."""
        n = self._skipped_host
        if self._arena is not None:
            n += int(self._arena.skipped_dev.item())
        return n

    @torch.no_grad()
    def step(self, closure=None, grad_scale: float = 1.0):
        loss = closure() if closure is not None else None
        from . import engine
        ar = self.arena()
        if ar is not None and ar.grads_in_place():
            group = self.param_groups[0]
            b1, b2 = group["betas"]
            st = torch.cuda.current_stream(ar.grad.device).cuda_stream
            ar.refresh_pointers()
            flag = None
            if self.skip_nonfinite:
                L.call("ng_nonfinite_flag", ar.grad.data_ptr(), ar.total, ar.flag.data_ptr(), st)
                ar.skipped_dev.add_(ar.flag)
                flag = ar.flag.data_ptr()
            L.call("ng_adam_multi", ar.ptrs_dev.data_ptr(), ar.offs_dev.data_ptr(), len(ar.params), ar.grad.data_ptr(),
                   ar.m.data_ptr(), ar.v.data_ptr(), ar.total, float(group["lr"]), float(b1), float(b2),
                   float(group["eps"]), ar.step_dev.data_ptr(), float(grad_scale), flag, st)
            self._host_steps += 1
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
                st["step"] += 1
                g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                p._b200_epoch = engine.WEIGHT_EPOCH[0] + 1      # this parameter's storage changes in this step
                L.call("ng_adam_step", p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(),
                       st["exp_avg_sq"].data_ptr(), p.numel(), float(group["lr"]), float(b1), float(b2),
                       float(group["eps"]), int(st["step"]), float(grad_scale),
                       torch.cuda.current_stream(p.device).cuda_stream)
        if self._arena is not None:      # keep the device-side step counter of the arena path in sync
            self._host_steps += 1
            self._arena.step_dev.add_(1)
        engine.WEIGHT_EPOCH[0] += 1     # masters changed in place: packed low-precision shadows must be refreshed
        return loss


def allreduce_gradients(params: Iterable[torch.nn.Parameter], world_size: Optional[int] = None, group=None) -> None:
    """DDP semantics: average gradients over ranks with ONE all-reduce per optimizer (the payload is <= 62 MB fp32;
    NVLink 5 / NVSwitch makes it latency-bound, so a single bucket minimises launches).  When the gradients already
    live in a B200Adam arena the collective runs on the arena itself.  Backend: NCCL on GPUs, gloo in the CPU tests."""
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
        ar.grad.div_(ws)
        return
    flat = torch.cat([p.grad.reshape(-1) for p in ps])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat.div_(ws)
    off = 0
    for p in ps:
        n = p.numel()
        p.grad.copy_(flat[off:off + n].view_as(p.grad))
        off += n
