"""Autograd bridges for the training step (model/pix2pix.py:165-257): the generator and the
discriminator appear to torch autograd as single differentiable functions of (input, parameters);
their forward and backward are the compiled C-ABI plans of ``graph.UnitGraph``.

Gradient precision: in the fp16 fast mode activations' gradients are carried in fp16 with an adaptive
power-of-two scale f = 2^floor(log2(target / max|dout|)) (target 8, NIRGAN_B200_GRAD_AMAX) computed on
the device by ng_grad_scale_pow2: it is applied when the fp32 output gradient enters the plan and divided
out exactly (the plan is linear in dout) when weight / input gradients are exported as fp32.
"""
from __future__ import annotations

from typing import List

import torch

from . import _lib as L


def _stream(t):
    return torch.cuda.current_stream(t.device).cuda_stream


# parameters whose arena slot has been handed to autograd during the CURRENT backward pass (cleared by an engine
# callback when the pass ends): a parameter reached twice in one pass (the PatchGAN sees the fake and the real batch)
# must not be given the same memory twice -- autograd would sum two aliases of it
_LENT: set = set()


def _slot_of(p: torch.nn.Parameter):
    slot = getattr(p, "_b200_grad_slot", None)
    return slot if (slot is not None and slot.device == p.device) else None


def _grad_dst(p: torch.nn.Parameter) -> torch.Tensor:
    """Where the fp32 gradient of `p` is written: a fresh view of its slot in the optimizer's flat arena
    (optim.B200Adam) the first time the parameter is reached in a backward pass and when no .grad is live -- autograd
    then adopts the view as p.grad without a copy -- else a new tensor (finished by _grad_done)."""
    slot = _slot_of(p)
    if slot is not None and p.grad is None and id(p) not in _LENT:
        if not _LENT:
            torch.autograd.Variable._execution_engine.queue_callback(_LENT.clear)
        _LENT.add(id(p))
        return slot.view_as(slot)
    return torch.empty_like(p, dtype=torch.float32)


def _grad_done(p: torch.nn.Parameter, g: torch.Tensor):
    """What to return to autograd for `p`: the tensor itself, or None after adding it into the slot that an earlier
    function of this pass already handed over (None = zero contribution; the sum lives in the arena)."""
    slot = _slot_of(p)
    if slot is not None and g.data_ptr() != slot.data_ptr() and p.grad is None and id(p) in _LENT:
        slot.add_(g)
        return None
    return g


def _grad_fill(p: torch.nn.Parameter, value) -> torch.Tensor:
    dst = _grad_dst(p)
    if value is None:
        dst.zero_()
    else:
        dst.copy_(value.reshape(dst.shape))
    return _grad_done(p, dst)


def _export_weight_grads(graph, bwd_plan, params: List[torch.nn.Parameter]) -> dict:
    """Packed fp32 weight gradients -> reference-layout fp32 tensors keyed by parameter id (times the inverse of the
    adaptive gradient scale, read on the device)."""
    st = graph.stream
    out = {}
    inv = 1.0
    gs = bwd_plan.records.get("gscale")
    dev_inv = gs.data_ptr() + 4 if gs is not None else None
    for i, dwp in bwd_plan.records["dw"].items():
        u = graph.units[i]
        w = u.conv.weight
        gw = _grad_dst(w)
        d0, d1, kh, kw = w.shape
        if u.pack == "rowmerged":
            L.call("ng_unpack_weight_grad_rowmerged", dwp.data_ptr(), d0, d1, kh, kw, inv, dev_inv, gw.data_ptr(), st)
        else:
            L.call("ng_unpack_weight_grad", dwp.data_ptr(), d0, d1, kh, kw, u.pack, u.cout, u.x.C, inv, dev_inv,
                   gw.data_ptr(), st)
        out[id(w)] = _grad_done(w, gw)
        if u.conv.bias is not None:
            dbb = bwd_plan.records["db"].get(i)
            if dbb is not None:
                db = dbb[:u.conv.bias.numel()]
                out[id(u.conv.bias)] = _grad_fill(u.conv.bias, db * gs[1] if gs is not None else db)
            else:
                # bias feeding InstanceNorm: its gradient is identically zero (the reference returns rounding noise)
                out[id(u.conv.bias)] = _grad_fill(u.conv.bias, None)
    th = bwd_plan.records.get("tap_head")
    if th is not None:
        # head as tap GEMM: dwp is [tap (64 stored)][channel]; the reference layout (1, C, kh, kw) is [channel][tap]
        w = th["conv"].weight
        gw = _grad_dst(w)
        _, cin, kh, kw = w.shape
        L.call("ng_unpack_weight_grad", th["dwp"].data_ptr(), cin, kh * kw, 1, 1, 1, 64, cin, inv, dev_inv,
               gw.data_ptr(), st)
        out[id(w)] = _grad_done(w, gw)
        # every tap column of dz sums to sum(dy): the centre tap's column sum is the bias gradient
        db = th["db"][th["center"]:th["center"] + 1]
        out[id(th["conv"].bias)] = _grad_fill(th["conv"].bias, db * gs[1] if gs is not None else db)
    return out


class GeneratorFunction(torch.autograd.Function):
    @staticmethod
    def run(module, runner, x, embeds, wrap_pad):
        params = [p for p in module.parameters()]
        return GeneratorFunction.apply(module, runner, wrap_pad, x, embeds, *params)

    @staticmethod
    def forward(ctx, module, runner, wrap_pad, x, embeds, *params):
        _LENT.clear()                   # no backward pass is in flight during a forward (belt and braces)
        c = runner.train_forward(x, embeds, wrap_pad)
        runner._live += 1               # buffers of this context must survive until its backward (see _RunnerBase.trim)
        B, Cin, H, W = c["geom"]
        fwd = c["fwd"]
        ctx.c, ctx.module, ctx.runner = c, module, runner
        ctx.params = params
        ctx.has_embeds = embeds is not None
        out = fwd.records["out"].view(B, 1, H, W)
        if getattr(module, "post_correction", False):
            raise NotImplementedError("nirgan_b200: training with post_correction is outside the hot path")
        return out.clone()

    @staticmethod
    def backward(ctx, dout):
        c, module, runner = ctx.c, ctx.module, ctx.runner
        g, bwd = c["graph"], c["bwd"]
        runner._live = max(0, runner._live - 1)
        c["fresh"] = None               # a backward may follow only the forward that produced these activations
        st = _stream(dout)
        c["dout"].view_as(dout).copy_(dout.float())
        inj = bwd.records.get("inject")
        if inj is not None:
            inj["dscale"].zero_()
        bwd.run_training(dout.device)
        grads = _export_weight_grads(g, bwd, ctx.params)
        gs = bwd.records.get("gscale")
        dev_inv = gs.data_ptr() + 4 if gs is not None else None
        if inj is not None:
            eng = runner._engine
            B = c["geom"][0]
            fc = module.fc
            dW, db = _grad_dst(fc.weight), _grad_dst(fc.bias)
            scratch = eng.buffers.get("gt.de128", B * 128 * 128, torch.float32)
            L.call("ng_inject_bwd", inj["de_map"].data_ptr(), B, inj["H"], inj["W"], 1.0, dev_inv,
                   c["fwd"].records["emb"].data_ptr(), scratch.data_ptr(), dW.data_ptr(), db.data_ptr(), st)
            grads[id(fc.weight)], grads[id(fc.bias)] = _grad_done(fc.weight, dW), _grad_done(fc.bias, db)
            if hasattr(module, "scale_param"):
                ds = inj["dscale"] * gs[1] if gs is not None else inj["dscale"]
                grads[id(module.scale_param)] = _grad_fill(module.scale_param, ds)
        out = []
        for p in ctx.params:
            gp = grads.get(id(p))
            out.append(gp if (gp is not None and p.requires_grad) else None)
        return (None, None, None, None, None, *out)


class DiscriminatorFunction(torch.autograd.Function):
    @staticmethod
    def run(module, runner, x):
        params = [p for p in module.parameters()]
        return DiscriminatorFunction.apply(module, runner, x, *params)

    @staticmethod
    def forward(ctx, module, runner, x, *params):
        _LENT.clear()
        need_dw = any(p.requires_grad for p in params)
        need_dx = x.requires_grad
        if runner._live == 0 and runner._engine is not None:
            runner.trim(runner._engine)
        slot = runner._live
        if slot >= 8:
            raise RuntimeError("nirgan_b200: 8 discriminator forwards are waiting for their backward; call "
                               "netD.reset_training_slots() if those graphs were dropped")
        runner._live += 1
        c = runner.train_context(x, slot, need_dw, need_dx)
        B, Cin, H, W = c["geom"]
        st = _stream(x)
        c["fwd"].records["src"].view(B, Cin, H, W).copy_(x.detach().float())
        c["fwd"].run_training(x.device)
        ctx.c, ctx.module, ctx.runner, ctx.params = c, module, runner, params
        ctx.need_dw, ctx.need_dx = need_dw, need_dx
        Ho, Wo = c["graph"].records["out_hw"]
        return c["graph"].units[-1].out_f32.view(B, 1, Ho, Wo).clone()

    @staticmethod
    def backward(ctx, dout):
        c, runner = ctx.c, ctx.runner
        g, bwd = c["graph"], c["bwd"]
        st = _stream(dout)
        c["dout"].view_as(dout).copy_(dout.float())
        bwd.run_training(dout.device)
        runner._live = max(0, runner._live - 1)
        grads = _export_weight_grads(g, bwd, ctx.params) if ctx.need_dw else {}
        gs = bwd.records.get("gscale")
        dev_inv = gs.data_ptr() + 4 if gs is not None else None
        dx = None
        if ctx.need_dx:
            B, Cin, H, W = c["geom"]
            gx = bwd.records["dx"]
            dx = torch.empty(B, Cin, H, W, dtype=torch.float32, device=dout.device)
            L.call("ng_grad_to_nchw", gx.t.data_ptr(), runner._engine.dt_enum, B, H, W, gx.C, Cin, 1.0, dev_inv,
                   dx.data_ptr(), st)
        out = []
        for p in ctx.params:
            gp = grads.get(id(p))
            out.append(gp if (gp is not None and p.requires_grad) else None)
        return (None, None, dx, *out)
