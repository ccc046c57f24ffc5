"""Adam on the ng_adam_step kernel (torch.optim.Adam semantics: eps 1e-8, no weight decay, bias-corrected;
model/pix2pix.py:486-487) and the data-parallel gradient all-reduce of the DDP path (train.py:118-120)."""
from __future__ import annotations

from typing import Iterable, Optional

import torch

from . import _lib as L


class B200Adam(torch.optim.Optimizer):
    def __init__(self, params, lr=2e-4, betas=(0.5, 0.999), eps=1e-8):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))
        self.skip_nonfinite = True
        self.skipped_steps = 0

    @torch.no_grad()
    def step(self, closure=None, grad_scale: float = 1.0):
        loss = closure() if closure is not None else None
        # low-precision backward: a non-finite gradient (fp16 overflow despite the loss scale) skips the update,
        # like torch.cuda.amp.GradScaler does; the flag is read once per step.
        grads = [p.grad for g in self.param_groups for p in g["params"] if p.grad is not None]
        if grads and self.skip_nonfinite:
            bad = torch.stack([(~torch.isfinite(g)).any() for g in grads]).any()
            if bool(bad):
                self.skipped_steps += 1
                return loss
        from . import engine
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
        engine.WEIGHT_EPOCH[0] += 1     # masters changed in place: packed low-precision shadows must be refreshed
        return loss


def allreduce_gradients(params: Iterable[torch.nn.Parameter], world_size: Optional[int] = None, group=None) -> None:
    """DDP semantics: average gradients over ranks with ONE all-reduce per optimizer (the payload is <= 62 MB fp32;
    NVLink 5 / NVSwitch makes it latency-bound, so a single bucket minimises launches).  Backend: NCCL on GPUs,
    gloo in the CPU tests."""
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized():
        return
    ws = world_size or dist.get_world_size(group)
    if ws == 1:
        return
    ps = [p for p in params if p.grad is not None]
    if not ps:
        return
    flat = torch.cat([p.grad.reshape(-1) for p in ps])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat.div_(ws)
    off = 0
    for p in ps:
        n = p.numel()
        p.grad.copy_(flat[off:off + n].view_as(p.grad))
        off += n
