"""Autograd bridges for the training step (model/pix2pix.py:165-257): the generator and the
discriminator appear to torch autograd as single differentiable functions of (input, parameters);
their forward and backward are the compiled C-ABI plans of ``graph.UnitGraph``.

Weight gradients never pass through torch arithmetic: the backward plan itself unpacks every packed fp32 weight
gradient into the parameter's slot of the network's flat gradient arena (``optim.GradArena``; reference layout, adaptive
gradient scale divided out) and autograd is handed views of those slots, which it adopts as ``.grad`` without a copy.
When a parameter is reached a second time in one backward pass, or its ``.grad`` still lives in the slot (no
``zero_grad(set_to_none=True)``), the export kernels accumulate instead (``beta`` = 1) and autograd is told "nothing to
add" -- exactly AccumulateGrad's semantics.

Gradient precision: in the fp16 fast mode activations' gradients are carried in fp16 with an adaptive
power-of-two scale f = 2^floor(log2(target / max|dout|)) (target 8, NIRGAN_B200_GRAD_AMAX) computed on
the device by ng_grad_scale_pow2: it is applied when the fp32 output gradient enters the plan and divided
out exactly (the plan is linear in dout) when weight / input gradients are exported as fp32.
"""
from __future__ import annotations

import torch

from . import _lib as L


def _stream(t):
    return torch.cuda.current_stream(t.device).cuda_stream


# parameters whose arena slot has been handed to autograd during the CURRENT backward pass (cleared by an engine
# callback when the pass ends): a parameter reached twice in one pass must not be given the same memory twice --
# autograd would sum two aliases of it
_LENT: set = set()


def _lend(p: torch.nn.Parameter) -> bool:
    """Decide how this backward pass delivers the gradient of `p`: True = the export overwrites the slot and autograd
    adopts a view of it; False = the export accumulates into the slot (already lent in this pass, or .grad lives there)
    and autograd receives None."""
    slot = p._b200_grad_slot
    if id(p) in _LENT or (p.grad is not None and p.grad.data_ptr() == slot.data_ptr()):
        return False
    if not _LENT:
        torch.autograd.Variable._execution_engine.queue_callback(_LENT.clear)
    _LENT.add(id(p))
    return True


def _deliver(bwd_plan, params) -> dict:
    """Set the accumulate cells of the plan's gradient exports for this pass; returns {id(p): tensor | None} for the
    parameters the plan exports plus the identically-zero ones."""
    out = {}
    need = {id(p) for p in params if p.requires_grad}
    for p, beta in bwd_plan.records["exports"]:
        if id(p) not in need:
            beta.value = 0.0          # frozen parameter: its slot is scratch
            continue
        fresh = _lend(p)
        beta.value = 0.0 if fresh else 1.0
        out[id(p)] = p._b200_grad_slot.view_as(p) if fresh else None
    for p in bwd_plan.records["zero_grads"]:
        if id(p) in need:
            out[id(p)] = p._b200_grad_slot.view_as(p) if _lend(p) else None      # the slot is never written: zeros
    return out


class GeneratorFunction(torch.autograd.Function):
    @staticmethod
    def run(module, runner, x, embeds, wrap_pad, reuse_token=None):
        params = [p for p in module.parameters() if p is not getattr(module, "post_correction_param", None)]
        return GeneratorFunction.apply(module, runner, wrap_pad, reuse_token, x, embeds, *params)

    @staticmethod
    def forward(ctx, module, runner, wrap_pad, reuse_token, x, embeds, *params):
        if x.requires_grad or (embeds is not None and embeds.requires_grad):
            raise NotImplementedError("nirgan_b200: gradients w.r.t. the generator's inputs (tiles / embeddings) are not "
                                      "part of the hot path (the reference trains the networks' parameters only)")
        _LENT.clear()                   # no backward pass is in flight during a forward (belt and braces)
        c = runner.train_forward(x, embeds, wrap_pad, reuse_token=reuse_token)
        c["live"] = True                # buffers of this context must survive until its backward
        if c.get("pool") is not None:
            c["pool"].live = c
        runner._live += 1
        B, Cin, H, W = c["geom"]
        ctx.c, ctx.module, ctx.runner = c, module, runner
        ctx.params = params
        # a view of the plan's static output buffer: valid until the next training forward of this shape
        return c["fwd"].records["out"].view(B, 1, H, W)

    @staticmethod
    def backward(ctx, dout):
        c, module, runner = ctx.c, ctx.module, ctx.runner
        bwd = c["bwd"]
        runner._live = max(0, runner._live - 1)
        c["live"] = False
        c["share_token"] = None         # a backward may follow only the forward that produced these activations
        st = _stream(dout)
        c["dout"].view_as(dout).copy_(dout)
        inj = bwd.records.get("inject")
        if inj is not None:
            inj["dscale"].zero_()
        grads = _deliver(bwd, ctx.params)
        bwd.run_training(dout.device)
        gs = bwd.records.get("gscale")
        dev_inv = gs.data_ptr() + 4 if gs is not None else None
        if inj is not None:
            B = c["geom"][0]
            fc = module.fc
            need = {id(p) for p in ctx.params if p.requires_grad}
            scratch = c["de128"]
            if id(fc.weight) in need or id(fc.bias) in need:
                fw, fb = _lend(fc.weight), _lend(fc.bias)
                dW = fc.weight._b200_grad_slot if fw else torch.empty_like(fc.weight)
                db = fc.bias._b200_grad_slot if fb else torch.empty_like(fc.bias)
                L.call("ng_inject_bwd", inj["de_map"].data_ptr(), B, inj["H"], inj["W"], 1.0, dev_inv,
                       c["fwd"].records["emb"].data_ptr(), scratch.data_ptr(), dW.data_ptr(), db.data_ptr(), st)
                if not fw:
                    fc.weight._b200_grad_slot.add_(dW)       # rare: fc reached twice / .grad not cleared
                if not fb:
                    fc.bias._b200_grad_slot.add_(db)
                grads[id(fc.weight)] = fc.weight._b200_grad_slot.view_as(fc.weight) if fw else None
                grads[id(fc.bias)] = fc.bias._b200_grad_slot.view_as(fc.bias) if fb else None
            sp = getattr(module, "scale_param", None)
            if sp is not None and id(sp) in need:
                fresh = _lend(sp)
                L.call("ng_unpack_weight_grad", inj["dscale"].data_ptr(), 1, 1, 1, 1, 0, 1, 1, 1.0, dev_inv,
                       0.0 if fresh else 1.0, sp._b200_grad_slot.data_ptr(), st)
                grads[id(sp)] = sp._b200_grad_slot.view_as(sp) if fresh else None
        out = []
        for p in ctx.params:
            gp = grads.get(id(p))
            out.append(gp if (gp is not None and p.requires_grad) else None)
        return (None, None, None, None, None, None, *out)


class DiscriminatorFunction(torch.autograd.Function):
    """forward(module, runner, struct, *tensors, *params): `tensors` are the unique input tensors of the parts
    (PatchGANRunner.describe), `struct` says how they combine."""

    @staticmethod
    def run(module, runner, parts):
        uniq, struct, Bp, ca, cb, H, W = runner.describe(parts)
        params = [p for p in module.parameters()]
        return DiscriminatorFunction.apply(module, runner, (struct, Bp, ca, cb, H, W), len(uniq), *uniq, *params)

    @staticmethod
    def forward(ctx, module, runner, desc, n_in, *rest):
        _LENT.clear()
        uniq, params = rest[:n_in], rest[n_in:]
        struct, Bp, ca, cb, H, W = desc
        need_dw = any(p.requires_grad for p in params)
        need_dx = any(t.requires_grad for t in uniq)
        if runner._live == 0 and runner._engine is not None:
            runner.trim(runner._engine)
        slot = runner._live
        if slot >= 8:
            raise RuntimeError("nirgan_b200: 8 discriminator forwards are waiting for their backward; call "
                               "netD.reset_training_slots() if those graphs were dropped")
        runner._live += 1
        chans = tuple(t.shape[1] for t in uniq)
        c = runner.train_context((struct, chans, Bp, ca, cb, H, W), slot, need_dw, need_dx, uniq[0].device)
        runner.load_inputs(c["graph"], uniq)
        c["fwd"].run_training(uniq[0].device)
        ctx.c, ctx.module, ctx.runner, ctx.params = c, module, runner, params
        ctx.need_dw, ctx.need_dx = need_dw, need_dx
        ctx.desc, ctx.n_in = desc, n_in
        ctx.in_needs = tuple(bool(t.requires_grad) for t in uniq)
        B = len(struct) * Bp
        Ho, Wo = c["graph"].records["out_hw"]
        return c["graph"].units[-1].out_f32.view(B, 1, Ho, Wo)

    @staticmethod
    def backward(ctx, dout):
        c, runner = ctx.c, ctx.runner
        bwd = c["bwd"]
        st = _stream(dout)
        c["dout"].view_as(dout).copy_(dout)
        grads = _deliver(bwd, ctx.params) if ctx.need_dw else {}
        bwd.run_training(dout.device)
        runner._live = max(0, runner._live - 1)
        gs = bwd.records.get("gscale")
        dev_inv = gs.data_ptr() + 4 if gs is not None else None
        dxs = [None] * ctx.n_in
        if ctx.need_dx:
            struct, Bp, ca, cb, H, W = ctx.desc
            gx = bwd.records["dx"]                     # gradient of the prepared input: [nparts*Bp][H][W][16], or its
            esz = gx.t.element_size()                  # space-to-depth form [nparts*Bp][(H+2)/2][(W+2)/2][4 x 16]
            s2d = bool(c["graph"].records.get("s2d"))
            img = ((H + 2) // 2) * ((W + 2) // 2) * 64 if s2d else H * W * gx.C
            for i, (ia, ib) in enumerate(struct):
                for j, c0, cn in ((ia, 0, ca), (ib, ca, cb)):
                    if j < 0 or not ctx.in_needs[j]:
                        continue
                    if dxs[j] is not None:
                        raise NotImplementedError("nirgan_b200: one tensor feeding two discriminator parts with "
                                                  "requires_grad is not supported")
                    dx = torch.empty(Bp, cn, H, W, dtype=torch.float32, device=dout.device)
                    L.call("ng_grad_to_nchw", gx.t.data_ptr() + i * Bp * img * esz, runner._engine.dt_enum, Bp,
                           H, W, 16 if s2d else gx.C, c0, cn, 1 if s2d else 0, 1.0, dev_inv, dx.data_ptr(), st)
                    dxs[j] = dx
        out = []
        for p in ctx.params:
            gp = grads.get(id(p))
            out.append(gp if (gp is not None and p.requires_grad) else None)
        return (None, None, None, None, *dxs, *out)
