"""Bridges the reference-style ``nn.Module.forward`` calls to the engine.

Inference (``torch.no_grad()`` / ``module.eval()`` with no parameter requiring grad) goes straight to
the compiled forward plan.  Training goes through ``torch.autograd.Function``s whose backward runs
the sm_100a dgrad / wgrad / InstanceNorm-backward kernels (see ``train.py``).
"""
from __future__ import annotations

import torch


def _needs_grad(module, *tensors) -> bool:
    if not torch.is_grad_enabled():
        return False
    if any(t is not None and t.requires_grad for t in tensors):
        return True
    return any(p.requires_grad for p in module.parameters())


def generator_apply(module, runner, x, embeds, wrap_pad, reuse_token=None):
    if _needs_grad(module, x, embeds):
        from .train import GeneratorFunction
        out = GeneratorFunction.run(module, runner, x, embeds, wrap_pad, reuse_token)
        if getattr(module, "post_correction", False):
            # model/generator_inject.py:133-134: a learnable scalar after the tanh head; torch autograd differentiates
            # this one multiply (d/d(param) = sum(dout * tanh), d/d(tanh) = dout * param) around the compiled plans
            out = out * module.post_correction_param
        return out
    return runner.forward(x, embeds, wrap_pad)


def discriminator_apply(module, runner, parts):
    tensors = [t for part in parts for t in part if t is not None]
    if _needs_grad(module, *tensors):
        from .train import DiscriminatorFunction
        return DiscriminatorFunction.run(module, runner, parts)
    return runner.forward(parts)


def lsgan_apply(pred, target: float):
    from .losses import lsgan
    return lsgan(pred, target)
