"""Runners: build the unit graphs of the generator (plain / SatCLIP-injected) and the PatchGAN
discriminator from reference-layout ``nn.Module``s and execute them (inference and training)."""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional, Tuple

import torch

from . import _lib as L
from . import engine as _engine
from .engine import ActBuf, CountingBuffers, Engine, EngineConfig, Plan, Pool, PoolBuffers, conv_out, require_cuda, _ptr
from .graph import Unit, UnitGraph


class _RunnerBase:
    def __init__(self, module: torch.nn.Module, cfg: Optional[EngineConfig] = None):
        self.module = module
        self.cfg = cfg or EngineConfig.from_env()
        self._engine: Optional[Engine] = None
        self._fwd: Dict[Tuple, Tuple[UnitGraph, Plan]] = {}
        self._train: Dict[Tuple, dict] = {}
        self._live = 0                    # training contexts holding un-backwarded activations
        self.last_plan: Optional[Plan] = None

    def engine(self, device) -> Engine:
        if self._engine is None or self._engine.device != device:
            self._engine = Engine(self.cfg, device)
            self._fwd.clear()
            self._train.clear()
        return self._engine

    def trim(self, eng: Engine) -> None:
        """Plans own their memory for life (static plans, no allocator).  With many distinct input shapes the
        least-recently-used plans are dropped once the total exceeds the budget (NIRGAN_B200_BUFFER_GB, default 90) --
        only when no forward is waiting for its backward.  Training contexts share one pool per runner (its size is that
        of the largest shape), so evicting them frees nothing unless pooling is off."""
        budget = float(os.environ.get("NIRGAN_B200_BUFFER_GB", "90")) * 1e9
        if self._live != 0:
            return

        def held():
            planned = sum(getattr(v[1], "planned_bytes", 0) for v in self._fwd.values())
            return eng.buffers.bytes() + planned + self.pool_bytes()

        def tags_of(v):
            return set(v.get("tags", ()) if isinstance(v, dict) else getattr(v[1], "tags", ()))

        while held() > budget:
            cands = [(v.get("used", 0) if isinstance(v, dict) else getattr(v[1], "used", 0), d, k)
                     for d in (self._fwd, self._train) for k, v in d.items()
                     if not (isinstance(v, dict) and v.get("pool") is not None)]
            if len(cands) <= 1:
                break
            _, d, k = min(cands, key=lambda c: c[0])
            tags = tags_of(d[k])
            if not tags:
                break
            # contexts that differ only in what they differentiate share their buffers (same tag): they go together
            for dd in (self._fwd, self._train):
                for kk in [kk for kk, vv in dd.items() if tags & tags_of(vv)]:
                    del dd[kk]
            eng.buffers.drop(tags)

    def pooled_build(self, eng: Engine, key, pool_key, build_fn) -> dict:
        """Build a training context whose buffers come from the runner's shared pool `pool_key` (engine.Pool): a first,
        allocation-free pass over the same code measures the context; the pool is (re)allocated when it is too small --
        contexts bound to the old pool are dropped and rebuild on their next use, so feeding the largest shape first
        avoids the churn."""
        if not _engine.POOL[0]:
            return build_fn()
        if not hasattr(self, "_pools"):
            self._pools, self._ctx_bytes = {}, {}
        saved = eng.buffers
        need = self._ctx_bytes.get(key)
        if need is None:
            counter = CountingBuffers(eng.device)
            eng.buffers = counter
            try:
                build_fn()
            finally:
                eng.buffers = saved
            need = self._ctx_bytes[key] = counter.total
        pool = self._pools.get(pool_key)
        if pool is None or pool.capacity < need:
            if pool is not None:
                for k in [k for k, v in self._train.items() if v.get("pool") is pool]:
                    del self._train[k]
                self._pools.pop(pool_key)
                del pool
                torch.cuda.empty_cache()
            pool = self._pools[pool_key] = Pool(need, eng.device)
        eng.buffers = PoolBuffers(pool, eng.device)
        try:
            ctx = build_fn()
        finally:
            eng.buffers = saved
        ctx["pool"] = pool
        return ctx

    def pool_bytes(self) -> int:
        return sum(p.capacity for p in getattr(self, "_pools", {}).values())

    def touch(self, entry) -> None:
        self._clock = getattr(self, "_clock", 0) + 1
        if isinstance(entry, dict):
            entry["used"] = self._clock
        else:
            entry[1].used = self._clock          # (graph, plan) pair of an inference plan

    def grad_arena(self):
        """The flat gradient arena shared with the optimizer of this network (optim.GradArena)."""
        from .optim import GradArena
        ar = GradArena.of(list(self.module.parameters()))
        if ar is None:
            raise RuntimeError("nirgan_b200: training needs every parameter of the network as a contiguous fp32 CUDA tensor")
        return ar

    def ddp_hook_units(self, graph: UnitGraph, buckets: int = 3):
        """Units after whose gradient export the data-parallel exchange may send a bucket: `buckets` - 1 cut points at
        roughly equal parameter mass, walking the units in backward order.  Empty when no process group with more than
        one rank is initialised (single-GPU plans stay hook-free)."""
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1) or buckets <= 1:
            return ()
        sizes = [u.conv.weight.numel() for u in graph.units]
        total, acc, cuts, target = sum(sizes), 0, [], 1
        for i in range(len(sizes) - 1, 0, -1):
            acc += sizes[i]
            if acc >= total * target / buckets and len(cuts) < buckets - 1:
                cuts.append(i)
                target += 1
        return tuple(cuts)

    def loss_scale(self) -> float:
        """Target max |dL/d(output)| of the adaptive power-of-two gradient scaling (ng_grad_scale_pow2) used when
        gradients are stored as fp16; 0 = no scaling (bf16 / fp32 have the range)."""
        return float(os.environ.get("NIRGAN_B200_GRAD_AMAX", "8")) if self.cfg.precision == "fp16" else 0.0


# =================================================================================================
class GeneratorRunner(_RunnerBase):
    """ResnetGenerator / ResnetGenerator_inject (model/networks.py:341-374, model/generator_inject.py:105-135)."""

    def __init__(self, module, cfg=None):
        super().__init__(module, cfg)
        self.head_mode = os.environ.get("NIRGAN_B200_HEAD", "tapgemm")     # inference head: 'tapgemm' | 'direct'
        self._side_streams: list = []

    def _convs(self):
        m = self.module.model
        nb = self.module.n_blocks
        return m[1], m[4], m[7], [m[10 + b] for b in range(nb)], m[10 + nb], m[13 + nb], m[17 + nb]

    def build_graph(self, eng: Engine, B: int, H: int, W: int, wrap: int, inject: bool, stream: int, tag: str,
                    direct_head: bool, direct_stem: bool = False, keep_rowmerged: bool = False) -> UnitGraph:
        mod = self.module
        stem, d1, d2, blocks, u1, u2, head = self._convs()
        ngf, cin = stem.weight.shape[0], stem.weight.shape[1]
        if ngf % 64:
            raise NotImplementedError("nirgan_b200 kernels are tiled for ngf multiples of 64")
        if cin > 8:
            raise NotImplementedError("nirgan_b200 stem kernel: input_nc <= 8")
        H1, W1 = H + 2 * wrap, W + 2 * wrap
        if H1 % 4 or W1 % 4:
            raise RuntimeError(f"nirgan_b200: padded tile {H1}x{W1} must be divisible by 4 (two stride-2 stages)")
        g = UnitGraph(eng, tag)
        # input: NCHW fp32 -> row-merged NHWC [B][H1+6][W1][kw*8+c] (wrapper reflect pad + stem reflect halo fused):
        # the 7x7x3 stem becomes a 7x1 conv over 64 "channels" = 7 K-steps of 128-byte rows instead of 49 thin taps
        src = eng.buffers.get(tag + ".in", B * cin * H * W, torch.float32)
        g.records["src"] = src
        C2, C4 = 2 * ngf, 4 * ngf
        H2, W2 = conv_out(H1, 3, 2, 1), conv_out(W1, 3, 2, 1)
        H3, W3 = conv_out(H2, 3, 2, 1), conv_out(W2, 3, 2, 1)
        direct_stem = direct_stem and eng.impl == L.IMPL_TC and cin <= 4 and ngf == 64 and H1 >= 8 and W1 >= 16
        if direct_stem and keep_rowmerged:
            # training plans: the forward runs ng_stem_conv from the fp32 tiles as well (0.27 -> 0.18 ms per 32 tiles); the
            # row-merged tensor is still written by ng_prep_stem because the stem's weight gradient contracts over it
            x0 = ActBuf(eng.buffers.get(tag + ".x0", B * (H1 + 6) * W1 * 64, eng.dt_torch), B, H1, W1, 64, 3)
            g.pre_ops.append(("ng_prep_stem", (src.data_ptr(), cin, B, H, W, wrap, 3, 7, eng.dt_enum, x0.t.data_ptr()),
                              tag + ".prep"))
            u = g.add(Unit("stem", stem, x0, ngf, 7, 1, 3, H1, W1, pack="rowmerged", KW=1, pad_w=0, in_pad_w=0,
                           act=L.ACT_RELU, out_pad=0,
                           direct={"src": src, "cin": cin, "H": H, "W": W, "wrap": wrap}))
        elif direct_stem:
            # forward-only plans: ng_stem_conv builds the im2col tile in shared memory from the fp32 tiles (no row-merged
            # tensor in HBM; training plans keep it because the stem's weight gradient reads it)
            xd = ActBuf(src, B, H1, W1, 32, 3)            # shape-only description of the virtual row-merged input
            u = g.add(Unit("stem", stem, xd, ngf, 7, 1, 3, H1, W1, pack="rowmerged4", KW=1, pad_w=0, in_pad_w=0,
                           act=L.ACT_RELU, out_pad=0,
                           direct={"src": src, "cin": cin, "H": H, "W": W, "wrap": wrap}))
        else:
            x0 = ActBuf(eng.buffers.get(tag + ".x0", B * (H1 + 6) * W1 * 64, eng.dt_torch), B, H1, W1, 64, 3)
            g.pre_ops.append(("ng_prep_stem", (src.data_ptr(), cin, B, H, W, wrap, 3, 7, eng.dt_enum, x0.t.data_ptr()),
                              tag + ".prep"))
            u = g.add(Unit("stem", stem, x0, ngf, 7, 1, 3, H1, W1, pack="rowmerged", KW=1, pad_w=0, in_pad_w=0,
                           act=L.ACT_RELU, out_pad=0))
        inj = None
        if inject:
            if H2 != W2:
                raise RuntimeError("nirgan_b200: SatCLIP injection is defined for square tiles only "
                                   "(generator_inject.py:116 passes size=(W,H))")
            emb = eng.buffers.get(tag + ".emb", B * 256, torch.float32)
            e = eng.buffers.get(tag + ".e", B * 128 * 128, torch.float32)
            g.records["emb"] = emb
            g.pre_ops.append(("ng_linear", (emb.data_ptr(), mod.fc.weight.data_ptr(), mod.fc.bias.data_ptr(), B, 256,
                                            128 * 128, e.data_ptr()), tag + ".fc"))
            style = mod.inject_style
            if style == "add":
                mode = L.INJECT_ADD
            elif style == "multiply":
                # `and self.scale_param` truthiness quirk (generator_inject.py:124): one host read at plan time
                mode = L.INJECT_MUL_SCALED if bool(mod.scale_param) else L.INJECT_MUL
            else:
                raise NotImplementedError(f"inject style [{style}] is not recognized")
            inj = {"e": e, "mode": mode, "scale": mod.scale_param.data}
        u = g.add(Unit("d1", d1, u.out, C2, 3, 2, 1, H2, W2, act=L.ACT_RELU, out_pad=0, inject=inj))
        nb = len(blocks)
        u = g.add(Unit("d2", d2, u.out, C4, 3, 2, 1, H3, W3, act=L.ACT_RELU, out_pad=1 if nb else 0))
        for b, blk in enumerate(blocks):
            src_idx = len(g.units) - 1
            a = g.add(Unit(f"r{b}a", blk.conv_block[1], u.out, C4, 3, 1, 1, H3, W3, act=L.ACT_RELU, out_pad=1))
            # out = x + IN(conv2(...)); no ReLU after the add (networks.py:433)
            u = g.add(Unit(f"r{b}b", blk.conv_block[5], a.out, C4, 3, 1, 1, H3, W3, act=L.ACT_NONE,
                           out_pad=0 if b == nb - 1 else 1, residual=src_idx))
        # ConvTranspose k3 s2 p1 op1 as 4 output-parity phases
        u = g.add(Unit("u1", u1, u.out, C2, 3, 2, 1, 2 * H3, 2 * W3, form=L.FORM_PHASED, pack=1, act=L.ACT_RELU))
        u = g.add(Unit("u2", u2, u.out, ngf, 3, 2, 1, 4 * H3, 4 * W3, form=L.FORM_PHASED, pack=1, act=L.ACT_RELU,
                       out_pad=3))
        g.records["head_in"] = u.out
        if direct_head:
            g.add(Unit("head", head, u.out, 16, 7, 1, 3, H1, W1, kind="head", act=L.ACT_TANH, crop=wrap))
        else:
            g.tap_head = {"conv": head, "x": u.out, "crop": wrap, "H": H, "W": W}
        g.records["geom"] = (B, H, W, H1, W1, wrap, ngf)
        return g

    def _use_tap_head(self) -> bool:
        return self.head_mode == "tapgemm" and self._convs()[0].weight.shape[0] == 64

    def _inference_plan(self, eng, B, Cin, H, W, wrap, inject, stream, slot: int = 0) -> Plan:
        key = (B, Cin, H, W, wrap, inject, slot)
        hit = self._fwd.get(key)
        if hit is not None:
            g, plan = hit
            g.refresh_weights()
            self.touch(hit)
            return plan
        self.trim(eng)
        tag = f"g{slot}_{B}x{Cin}x{H}x{W}p{wrap}{'i' if inject else ''}"      # shape-unique: plans are evicted by tag

        def build():
            g = self.build_graph(eng, B, H, W, wrap, inject, stream, tag, direct_head=not self._use_tap_head(),
                                 direct_stem=os.environ.get("NIRGAN_B200_STEM_DIRECT", "1") != "0")
            plan = g.compile_forward()
            plan.tags = (tag,)
            if g.tap_head is None:
                plan.records["out"] = g.units[-1].out_f32
            return g, plan

        if os.environ.get("NIRGAN_B200_MEMPLAN", "1") != "0":
            # forward-only plan: a buffer is dead after its last reader, so the ~70 per-layer buffers are laid out by
            # liveness in ONE allocation (engine.plan_memory): peak = the two or three tensors alive at the widest layer
            # instead of the sum over layers (B = 64 at 256x256: ~1.2 GB instead of 9.8 GB).  First pass: sizes and
            # pointer uses only (no allocation); second pass: the real plan over the planned layout.
            saved = eng.buffers
            counter = CountingBuffers(eng.device)
            eng.buffers = counter
            try:
                g0, p0 = build()
                pinned = [p0.records[k] for k in ("src", "emb", "out") if k in p0.records]
                offsets, total = _engine.plan_memory([p0], counter, pinned)
                eng.buffers = _engine.PlannedBuffers(offsets, total, eng.device)
                g, plan = build()
                plan.keepalive.append(eng.buffers)
                plan.planned_bytes = total
            finally:
                eng.buffers = saved
        else:
            g, plan = build()
        self._fwd[key] = (g, plan)
        self.touch(self._fwd[key])
        return plan

    def inference_bytes(self) -> int:
        """Device memory held by the cached inference plans (planned layouts + per-layer buffers)."""
        eng = self._engine
        return sum(getattr(p, "planned_bytes", 0) for _, p in self._fwd.values()) + (eng.buffers.bytes() if eng else 0)

    @torch.no_grad()
    def forward(self, x: torch.Tensor, embeds: Optional[torch.Tensor] = None, wrap_pad: int = 0) -> torch.Tensor:
        """Synchronous-in-stream-order call: the result is valid on the caller's current stream."""
        out, done = self.forward_async(x, embeds, wrap_pad)
        main = torch.cuda.current_stream(x.device)
        for ev in done:
            main.wait_event(ev)
        return out

    @torch.no_grad()
    def forward_async(self, x: torch.Tensor, embeds: Optional[torch.Tensor] = None, wrap_pad: int = 0,
                      ready: Optional["torch.cuda.Event"] = None):
        """Streaming form of ``forward``: returns ``(out, done_events)`` without making the caller's stream wait.  The
        batch is cut into slices that run on the runner's own streams; slice s of this call queues behind slice s of the
        previous call only, so consecutive calls overlap (the tail of one step runs beside the head of the next) instead
        of meeting at a barrier on the caller's stream.  ``ready`` = event after which ``x`` / ``embeds`` may be read
        (default: the caller's stream at call time).  A consumer waits on every event of ``done_events`` (a stream:
        ``stream.wait_event(e)``) before touching ``out``."""
        require_cuda(x, "generator input")
        if x.dim() != 4:
            raise RuntimeError("generator input must be (B, C, H, W)")
        eng = self.engine(x.device)
        Btot, Cin, H, W = x.shape
        inject = embeds is not None
        chunk = eng.cfg.chunk if eng.cfg.chunk > 0 else Btot
        main = torch.cuda.current_stream(x.device)
        if inject:
            require_cuda(embeds, "embeds")
        convert = x.dtype != torch.float32 or not x.is_contiguous() or \
            (inject and (embeds.dtype != torch.float32 or not embeds.is_contiguous()))
        if convert:
            if ready is not None:
                main.wait_event(ready)      # the layout / dtype conversion below reads the inputs on the caller's stream
            x = x.contiguous().float()
            if inject:
                embeds = embeds.contiguous().float()
        out = torch.empty(Btot, 1, H, W, dtype=torch.float32, device=x.device)
        want = eng.cfg.streams if eng.cfg.streams > 0 else (2 if Btot >= 32 else 1)
        nstreams = max(1, min(want, (Btot + chunk - 1) // chunk if eng.cfg.chunk > 0 else want, Btot))
        if eng.cfg.chunk <= 0 and nstreams > 1:
            chunk = (Btot + nstreams - 1) // nstreams
        if len(self._side_streams) < nstreams:
            self._side_streams += [torch.cuda.Stream(x.device) for _ in range(nstreams - len(self._side_streams))]
        streams = self._side_streams[:nstreams]
        # the packed weight shadows are refreshed IN PLACE on the caller's stream: when the masters changed since the
        # previous call, slices of that call may still be reading the old shadows on the runner's streams -- wait for them
        sig = self.weight_signature()
        if sig != getattr(self, "_async_sig", None):
            for ev in getattr(self, "_async_done", ()):
                main.wait_event(ev)
            self._async_sig = sig
        # build / refresh every plan on the caller's stream first (weight packing), then fork
        jobs = []
        for i, b0 in enumerate(range(0, Btot, chunk)):
            B = min(chunk, Btot - b0)
            slot = i % nstreams
            jobs.append((b0, B, slot, self._inference_plan(eng, B, Cin, H, W, wrap_pad, inject, main.cuda_stream, slot)))
        if ready is None:
            ready = torch.cuda.Event()
            ready.record(main)            # inputs (and refreshed weights) are ready once the caller's stream gets here
        else:
            wev = torch.cuda.Event()      # weight packing, if any, was issued on the caller's stream
            wev.record(main)
            for st in streams:
                st.wait_event(wev)
        for st in streams:
            st.wait_event(ready)
            x.record_stream(st)
            out.record_stream(st)
            if inject:
                embeds.record_stream(st)
        for b0, B, slot, plan in jobs:
            st = streams[slot]
            with torch.cuda.stream(st):
                plan.records["src"].view(B, Cin, H, W).copy_(x[b0:b0 + B])
                if inject:
                    plan.records["emb"].view(B, 256).copy_(embeds[b0:b0 + B])
                plan.run_graphed(st)
                o = plan.records["out"].view(B, 1, H, W)
                if getattr(self.module, "post_correction", False):
                    o = o * self.module.post_correction_param
                out[b0:b0 + B].copy_(o)
        done = []
        for st in streams:
            ev = torch.cuda.Event()
            ev.record(st)
            done.append(ev)
        self.last_plan = plan
        self._async_done = done
        return out, done

    # ---- training -------------------------------------------------------------------------------------
    def weight_signature(self):
        ps = list(self.module.parameters())
        return (max(getattr(p, "_b200_epoch", 0) for p in ps), sum(p._version for p in ps))

    def train_forward(self, x: torch.Tensor, embeds, wrap_pad: int, share: bool = False, reuse_token=None) -> dict:
        """Run the training forward plan (activations kept for the backward).

        Px2Px_PL.training_step evaluates G on the same batch for both optimizers (model/pix2pix.py:177-180) with
        identical results, so the second evaluation may be skipped -- but only through an explicit hand-over: the D pass
        calls with ``share=True`` and receives a token in ``ctx['share_token']``; the G pass passes that token back as
        ``reuse_token``.  The activations are reused only if the token is the context's current one (no other forward
        of this shape ran in between) and the generator's weights are unchanged; every other call recomputes."""
        c = self.train_context(x, embeds, wrap_pad)
        if c.get("live"):
            raise RuntimeError(
                "nirgan_b200: a grad-enabled generator forward of this shape is still waiting for its backward; a second "
                "one would overwrite its activations.  Run backward first, or call netG.reset_training_slots() if that "
                "graph was dropped.")
        pool = c.get("pool")
        if pool is not None and pool.live is not None and pool.live is not c and pool.live.get("live"):
            raise RuntimeError("nirgan_b200: a generator forward of another shape is still waiting for its backward; the "
                               "training contexts share one memory pool (NIRGAN_B200_POOL=0 gives every shape its own)")
        sig = self.weight_signature()
        tok = c.get("share_token")
        reuse = reuse_token is not None and tok is not None and reuse_token is tok[0] and tok[1] == sig and \
            (pool is None or pool.owner is c)
        c["share_token"] = None
        if pool is not None:
            pool.owner = c
        if not reuse:
            B, Cin, H, W = c["geom"]
            fwd = c["fwd"]
            fwd.records["src"].view(B, Cin, H, W).copy_(x.detach())
            if embeds is not None:
                fwd.records["emb"].view(B, 256).copy_(embeds.detach())
            main = torch.cuda.current_stream(x.device)
            halves = c.get("fwd_halves")
            if halves:
                # two half-batch plans over the same buffers on two streams: the HBM-bound apply kernels of one half
                # run beside the tensor-bound convolutions of the other (InstanceNorm is per image: same bits)
                if not self._side_streams:
                    self._side_streams.append(torch.cuda.Stream(x.device))
                side = self._side_streams[0]
                side.wait_stream(main)
                halves[0].run(main.cuda_stream)
                halves[1].run(side.cuda_stream)
                main.wait_stream(side)
            else:
                fwd.run_training(x.device)
        if share:
            c["share_token"] = (object(), sig)
        return c

    def train_context(self, x: torch.Tensor, embeds, wrap_pad: int) -> dict:
        """Graph + forward/backward plans with dedicated buffers (activations must survive until backward)."""
        eng = self.engine(x.device)
        B, Cin, H, W = x.shape
        inject = embeds is not None
        stream = torch.cuda.current_stream(x.device).cuda_stream
        key = (B, Cin, H, W, wrap_pad, inject)
        ctx = self._train.get(key)
        slots = tuple(p._b200_grad_slot.data_ptr() if hasattr(p, "_b200_grad_slot") else 0
                      for p in self.module.parameters())
        if ctx is not None and ctx["slots"] != slots:
            ctx = None                 # the gradient arena moved (a new optimizer adopted the parameters): re-bind
        if ctx is None:
            self.trim(eng)
            self.grad_arena()
            slots = tuple(p._b200_grad_slot.data_ptr() for p in self.module.parameters())
            tag = f"gt_{B}x{Cin}x{H}x{W}p{wrap_pad}{'i' if inject else ''}"

            def build():
                g = self.build_graph(eng, B, H, W, wrap_pad, inject, stream, tag, direct_head=not self._use_tap_head(),
                                     direct_stem=os.environ.get("NIRGAN_B200_STEM_DIRECT", "1") != "0", keep_rowmerged=True)
                fwd = g.compile_forward()
                if g.tap_head is None:
                    fwd.records["out"] = g.units[-1].out_f32
                dout = eng.buffers.get(tag + ".dout", B * H * W, torch.float32)
                bwd = g.compile_backward(dout, self.loss_scale(), need_dw=True, need_dx=False, want_inject_grads=inject,
                                         hook_units=self.ddp_hook_units(g))
                c = {"graph": g, "fwd": fwd, "bwd": bwd, "dout": dout, "geom": (B, Cin, H, W), "tags": (tag,),
                     "tag": tag, "slots": slots, "live": False}
                if inject:
                    c["de128"] = eng.buffers.get(tag + ".de128", B * 128 * 128, torch.float32)
                c["fwd_halves"] = self._half_batch_plans(eng, B, H, W, wrap_pad, inject, stream, tag)
                return c

            ctx = self._train[key] = self.pooled_build(eng, key, "g", build)
        else:
            ctx["graph"].refresh_weights(backward=True, need_dx=False)
        self.touch(ctx)
        return ctx


    def _half_batch_plans(self, eng, B, H, W, wrap_pad, inject, stream, tag):
        """Forward plans for the two halves of the batch that write the full-batch graph's buffers in place (None when
        the batch is odd / small, disabled by NIRGAN_B200_TRAIN_SLICES=0, or a buffer is not per-image)."""
        import os
        from .engine import SliceBuffers
        # measured on B200: slower than the single full-batch plan at batch 32 (21.6 vs 20.9 ms per step) and at batch 64
        # (41.3 vs 39.4 ms) -- every conv launch pays its ~25 us pipeline fill twice and the eager launches of the two
        # plans do not interleave the way the graph-replayed inference slices do.  Off unless NIRGAN_B200_TRAIN_SLICES=1.
        mode = os.environ.get("NIRGAN_B200_TRAIN_SLICES", "0")
        if B % 2 or B < 16 or mode != "1" or isinstance(eng.buffers, CountingBuffers):
            return None
        saved, plans = eng.buffers, []
        try:
            for part in range(2):
                eng.buffers = SliceBuffers(saved, part, 2)
                gh = self.build_graph(eng, B // 2, H, W, wrap_pad, inject, stream, tag,
                                      direct_head=not self._use_tap_head(),
                                      direct_stem=os.environ.get("NIRGAN_B200_STEM_DIRECT", "1") != "0", keep_rowmerged=True)
                plans.append(gh.compile_forward())
                plans[-1].keepalive.append(gh)
        except KeyError:
            return None
        finally:
            eng.buffers = saved
        return plans


# =================================================================================================
class PatchGANRunner(_RunnerBase):
    """NLayerDiscriminator (model/networks.py:539-584).

    The input is described as a list of *parts*, each the channel concatenation of one or two NCHW tensors
    (``torch.cat((rgb, pred), 1)``, model/pix2pix.py:197,202,216), stacked along the batch: the concatenation, the
    NCHW->NHWC transform, the channel padding and the conversion to the operand type are ONE ``ng_prep_input`` launch per
    part, and the fake and the real batch of the D pass run as one 2B batch through the network (InstanceNorm is per
    sample, so every sample's result is the one the reference computes with two separate calls)."""

    def conv_modules(self) -> List[torch.nn.Module]:
        return [m for m in self.module.model if isinstance(m, torch.nn.Conv2d)]

    @staticmethod
    def describe(parts):
        """parts [(a, b | None)] -> (unique tensors, ((index of a, index of b | -1), ...), Bp, ca, cb, H, W)."""
        uniq, struct = [], []
        Bp, ca, H, W = parts[0][0].shape
        cb = 0 if parts[0][1] is None else parts[0][1].shape[1]

        def idx(t):
            for j, u in enumerate(uniq):
                if u is t:
                    return j
            uniq.append(t)
            return len(uniq) - 1

        for a, b in parts:
            require_cuda(a, "discriminator input")
            if tuple(a.shape) != (Bp, ca, H, W) or (b is None) != (cb == 0) or \
                    (b is not None and tuple(b.shape) != (Bp, cb, H, W)):
                raise RuntimeError("nirgan_b200: every part of a discriminator batch must have the same shapes")
            struct.append((idx(a), -1 if b is None else idx(b)))
        return uniq, tuple(struct), Bp, ca, cb, H, W

    def build_graph(self, eng: Engine, Bp: int, struct, chans, ca: int, cb: int, H: int, W: int, tag: str) -> UnitGraph:
        convs = self.conv_modules()
        g = UnitGraph(eng, tag)
        cin = convs[0].weight.shape[1]
        if ca + cb != cin:
            raise RuntimeError(f"nirgan_b200: discriminator expects {cin} input channels, got {ca} + {cb}")
        nparts = len(struct)
        B = nparts * Bp
        srcs = [eng.buffers.get(f"{tag}.in{j}", Bp * c * H * W, torch.float32) for j, c in enumerate(chans)]
        g.records["srcs"] = srcs
        c0 = convs[0]
        Hc, Wc = conv_out(H, 4, 2, 1), conv_out(W, 4, 2, 1)
        s2d = (eng.impl == L.IMPL_TC and eng.dt_enum != L.F32 and H % 2 == 0 and W % 2 == 0 and cin <= 16 and
               tuple(c0.kernel_size) == (4, 4) and tuple(c0.stride) == (2, 2) and tuple(c0.padding) == (1, 1) and
               c0.weight.shape[0] % 64 == 0 and os.environ.get("NIRGAN_B200_D_S2D", "1") != "0")
        g.records["s2d"] = s2d
        if s2d:
            # input layer in space-to-depth form: [B][(H+2)/2][(W+2)/2][4 parities x 16 slots] of the zero-padded image,
            # on which Conv2d(k4, s2, p1) is a 2x2 stride-1 convolution over 64 stored channels (ng_prep_input_s2d)
            Hs, Ws = (H + 2) // 2, (W + 2) // 2
            x = eng.act(tag + ".x0", B, Hs, Ws, 64, 0)
            esz = x.t.element_size()
            for i, (ia, ib) in enumerate(struct):
                g.pre_ops.append(("ng_prep_input_s2d", (srcs[ia].data_ptr(), ca, srcs[ib].data_ptr() if ib >= 0 else None, cb,
                                                        Bp, H, W, eng.dt_enum, x.t.data_ptr() + i * Bp * Hs * Ws * 64 * esz),
                                  f"{tag}.prep{i}"))
            u = g.add(Unit("l0", c0, x, c0.weight.shape[0], 2, 1, 0, Hc, Wc, kind="biasact", act=L.ACT_LRELU, slope=0.2,
                           halo_mode=L.HALO_ZERO, pack="s2d"))
        else:
            x = eng.act(tag + ".x0", B, H, W, 16, 0)
            esz = x.t.element_size()
            for i, (ia, ib) in enumerate(struct):
                g.pre_ops.append(("ng_prep_input", (srcs[ia].data_ptr(), ca, srcs[ib].data_ptr() if ib >= 0 else None, cb, Bp,
                                                    H, W, 0, 0, L.HALO_ZERO, 16, eng.dt_enum,
                                                    x.t.data_ptr() + i * Bp * H * W * 16 * esz), f"{tag}.prep{i}"))
            u = g.add(Unit("l0", c0, x, c0.weight.shape[0], 4, 2, 1, Hc, Wc, kind="biasact", act=L.ACT_LRELU, slope=0.2,
                           halo_mode=L.HALO_ZERO))
        for i, conv in enumerate(convs[1:-1], start=1):
            s = conv.stride[0]
            Hn, Wn = conv_out(u.Hout, 4, s, 1), conv_out(u.Wout, 4, s, 1)
            u = g.add(Unit(f"l{i}", conv, u.out, conv.weight.shape[0], 4, s, 1, Hn, Wn, act=L.ACT_LRELU, slope=0.2,
                           out_pad=0, halo_mode=L.HALO_ZERO))
        cl = convs[-1]
        Ho, Wo = conv_out(u.Hout, 4, 1, 1), conv_out(u.Wout, 4, 1, 1)
        g.add(Unit("out", cl, u.out, 16, 4, 1, 1, Ho, Wo, kind="head", act=L.ACT_NONE))
        g.records["out_hw"] = (Ho, Wo)
        g.records["cin"] = cin
        g.records["parts"] = (nparts, Bp, ca, cb)
        return g

    @staticmethod
    def load_inputs(graph: UnitGraph, uniq) -> None:
        for dst, t in zip(graph.records["srcs"], uniq):
            dst.view(t.shape).copy_(t.detach())

    @torch.no_grad()
    def forward(self, parts) -> torch.Tensor:
        uniq, struct, Bp, ca, cb, H, W = self.describe(parts)
        eng = self.engine(uniq[0].device)
        chans = tuple(t.shape[1] for t in uniq)
        key = (Bp, struct, chans, H, W)
        hit = self._fwd.get(key)
        if hit is None:
            self.trim(eng)
            tag = f"d_{len(struct)}x{Bp}x{ca}+{cb}x{H}x{W}s{hash(struct) & 0xffff:x}"
            g = self.build_graph(eng, Bp, struct, chans, ca, cb, H, W, tag)
            plan = g.compile_forward()
            plan.tags = (tag,)
            hit = self._fwd[key] = (g, plan)
        else:
            hit[0].refresh_weights()
        self.touch(hit)
        g, plan = hit
        self.load_inputs(g, uniq)
        plan.run(torch.cuda.current_stream(uniq[0].device).cuda_stream)
        Ho, Wo = g.records["out_hw"]
        self.last_plan = plan
        return g.units[-1].out_f32.view(len(struct) * Bp, 1, Ho, Wo).clone()

    def train_context(self, parts_desc, slot: int, need_dw: bool, need_dx: bool, device) -> dict:
        struct, chans, Bp, ca, cb, H, W = parts_desc
        eng = self.engine(device)
        key = (Bp, struct, chans, H, W, slot, need_dw, need_dx)
        ctx = self._train.get(key)
        slots = tuple(p._b200_grad_slot.data_ptr() if hasattr(p, "_b200_grad_slot") else 0
                      for p in self.module.parameters()) if need_dw else ()
        if ctx is not None and ctx["slots"] != slots:
            ctx = None
        if ctx is None:
            if need_dw:
                self.grad_arena()
                slots = tuple(p._b200_grad_slot.data_ptr() for p in self.module.parameters())
            tag = f"dt{slot}_{len(struct)}x{Bp}x{ca}+{cb}x{H}x{W}s{hash(struct) & 0xffff:x}"

            def build():
                g = self.build_graph(eng, Bp, struct, chans, ca, cb, H, W, tag)
                fwd = g.compile_forward()
                Ho, Wo = g.records["out_hw"]
                B = len(struct) * Bp
                dout = eng.buffers.get(tag + ".dout", B * Ho * Wo, torch.float32)
                bwd = g.compile_backward(dout, self.loss_scale(), need_dw=need_dw, need_dx=need_dx,
                                         hook_units=self.ddp_hook_units(g, buckets=2) if need_dw else ())
                return {"graph": g, "fwd": fwd, "bwd": bwd, "dout": dout, "geom": (B, ca + cb, H, W), "tags": (tag,),
                        "tag": tag, "slots": slots}

            # forwards that wait for their backward at the same time sit in different slots: one pool per slot
            ctx = self._train[key] = self.pooled_build(eng, key, ("d", slot), build)
        else:
            ctx["graph"].refresh_weights(backward=True, need_dx=need_dx)
        self.touch(ctx)
        return ctx
