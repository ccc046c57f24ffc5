"""The optimisation loop around ``Px2Px.training_step``: what pytorch_lightning 1.9's automatic optimisation does for
the reference's two-optimizer LightningModule (reference train.py:118-126; model/pix2pix.py:165-257,485-492), one process
per GPU.

Per batch: optimizer 0 (discriminator) -- zero_grad, ``training_step(batch, i, 0)``, backward, gradient exchange,
``step`` -- then the same for optimizer 1 (generator), D frozen (PL ``toggle_optimizer``).  With a process group of
more than one rank the gradients are summed over the ranks by ``optim.BucketedAllReduce``: the backward plans call back
("every gradient at or above this arena offset is final") as soon as a bucket of the flat gradient arena is complete, the
bucket's NCCL all-reduce is launched on a communication stream and overlaps the rest of the backward pass, and the mean's
1/world factor is applied inside the Adam kernel (``grad_scale``).  DDP semantics: every rank ends each step with the
same parameters.
"""
from __future__ import annotations

from typing import Optional

import torch

from .optim import BucketedAllReduce, GradArena


class Trainer:
    def __init__(self, model, group=None, time_exchange: bool = False):
        """model: ``model.pix2pix.Px2Px`` on a CUDA device in train() mode with ``configure_b200`` applied.  When
        torch.distributed is initialised (NCCL, one rank per GPU) gradients are averaged over the ranks."""
        import torch.distributed as dist
        self.model = model
        self.opt_d, self.opt_g = model.configure_optimizers()
        self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        self.group = group
        self.reducers = {}
        if self.world > 1:
            for name, net in (("d", model.netD), ("g", model.netG)):
                ar = GradArena.of(list(net.parameters()))
                r = BucketedAllReduce(ar.grad, group=group)
                r.timing = bool(time_exchange)
                self.reducers[name] = r
        self.exchange = True          # False: skip the collectives (measurement of their cost; ranks then diverge)
        self.step_index = 0

    def _hook(self, reducer: BucketedAllReduce):
        def fn(payload, main, side):
            if payload[0] == "grads_ready" and self.exchange:
                reducer.ready(int(payload[1]), side if side is not None else main)
        return fn

    def _install(self, net, reducer: Optional[BucketedAllReduce]):
        runner = getattr(net, "_runner", None)
        if runner is None:
            return
        fn = self._hook(reducer) if reducer is not None else None
        for c in runner._train.values():
            c["bwd"].hook_fn = fn

    def _pass(self, batch, idx: int, net, opt, key: str):
        opt.zero_grad(set_to_none=True)
        loss = self.model.training_step(batch, self.step_index, idx)
        red = self.reducers.get(key)
        self._install(net, red)           # contexts are created by the forward above; hooks fire during backward
        if red is not None:
            red.reset()
        loss.backward()
        if red is not None and self.exchange:
            red.finish()
        opt.step(grad_scale=1.0 / self.world if (red is not None and self.exchange) else 1.0)
        return loss

    def step(self, batch):
        """One optimisation step on `batch` ({'rgb', 'nir'[, 'embeds' | 'coords']}); returns (loss_D, loss_G) as
        device scalars (no host synchronisation)."""
        ld = self._pass(batch, 0, self.model.netD, self.opt_d, "d")
        lg = self._pass(batch, 1, self.model.netG, self.opt_g, "g")
        self.step_index += 1
        return ld.detach(), lg.detach()

    def exposed_exchange_ms(self) -> float:
        """Device time the training stream spent waiting for the gradient collectives in the last step (needs
        ``time_exchange=True`` and a device synchronise before the call)."""
        return sum(r.exposed_ms() for r in self.reducers.values())
