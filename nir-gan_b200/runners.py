"""Runners: build the unit graphs of the generator (plain / SatCLIP-injected) and the PatchGAN
discriminator from reference-layout ``nn.Module``s and execute them (inference and training)."""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional, Tuple

import torch

from . import _lib as L
from .engine import ActBuf, Engine, EngineConfig, Plan, conv_out, require_cuda, _ptr
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
        """Plans own their buffers for life (static plan, no allocator).  With many distinct input shapes (mixed
        resolutions, BASELINE config 5) drop every cached plan and buffer once they exceed the budget
        (NIRGAN_B200_BUFFER_GB, default 90) -- only when no forward is waiting for its backward."""
        budget = float(os.environ.get("NIRGAN_B200_BUFFER_GB", "90")) * 1e9
        if eng.buffers.bytes() > budget and self._live == 0:
            self._fwd.clear()
            self._train.clear()
            eng.buffers._b.clear()

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
                    direct_head: bool) -> UnitGraph:
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
        g = UnitGraph(eng, tag, stream)
        # input: NCHW fp32 -> row-merged NHWC [B][H1+6][W1][kw*8+c] (wrapper reflect pad + stem reflect halo fused):
        # the 7x7x3 stem becomes a 7x1 conv over 64 "channels" = 7 K-steps of 128-byte rows instead of 49 thin taps
        src = eng.buffers.get(tag + ".in", B * cin * H * W, torch.float32)
        g.records["src"] = src
        x0 = ActBuf(eng.buffers.get(tag + ".x0", B * (H1 + 6) * W1 * 64, eng.dt_torch), B, H1, W1, 64, 3)
        g.pre_ops.append(("ng_prep_stem", (src.data_ptr(), cin, B, H, W, wrap, 3, 7, eng.dt_enum, x0.t.data_ptr()),
                          tag + ".prep"))
        C2, C4 = 2 * ngf, 4 * ngf
        H2, W2 = conv_out(H1, 3, 2, 1), conv_out(W1, 3, 2, 1)
        H3, W3 = conv_out(H2, 3, 2, 1), conv_out(W2, 3, 2, 1)
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
            return plan
        self.trim(eng)
        g = self.build_graph(eng, B, H, W, wrap, inject, stream, "g" if slot == 0 else f"g{slot}",
                             direct_head=not self._use_tap_head())
        plan = g.compile_forward()
        if g.tap_head is None:
            plan.records["out"] = g.units[-1].out_f32
        self._fwd[key] = (g, plan)
        return plan

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
        return out, done

    # ---- training -------------------------------------------------------------------------------------
    def weight_signature(self):
        ps = list(self.module.parameters())
        return (max(getattr(p, "_b200_epoch", 0) for p in ps), sum(p._version for p in ps))

    def train_forward(self, x: torch.Tensor, embeds, wrap_pad: int) -> dict:
        """Run the training forward plan (activations kept for the backward) unless the very same input tensors and
        weights were the last thing this context computed: Px2Px_PL.training_step evaluates G on the same batch for
        both optimizers (model/pix2pix.py:177-180) with identical results, so the second evaluation is reused."""
        c = self.train_context(x, embeds, wrap_pad)
        key = (x.data_ptr(), x._version, tuple(x.shape),
               None if embeds is None else (embeds.data_ptr(), embeds._version), self.weight_signature())
        if c.get("fresh") != key:
            B, Cin, H, W = c["geom"]
            fwd = c["fwd"]
            fwd.records["src"].view(B, Cin, H, W).copy_(x.detach().float())
            if embeds is not None:
                fwd.records["emb"].view(B, 256).copy_(embeds.detach().float())
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
            c["fresh"] = key
        return c

    def train_context(self, x: torch.Tensor, embeds, wrap_pad: int) -> dict:
        """Graph + forward/backward plans with dedicated buffers (activations must survive until backward)."""
        eng = self.engine(x.device)
        B, Cin, H, W = x.shape
        inject = embeds is not None
        stream = torch.cuda.current_stream(x.device).cuda_stream
        key = (B, Cin, H, W, wrap_pad, inject)
        ctx = self._train.get(key)
        if ctx is None:
            self.trim(eng)
            g = self.build_graph(eng, B, H, W, wrap_pad, inject, stream, "gt", direct_head=not self._use_tap_head())
            fwd = g.compile_forward()
            if g.tap_head is None:
                fwd.records["out"] = g.units[-1].out_f32
            dout = eng.buffers.get("gt.dout", B * H * W, torch.float32)
            bwd = g.compile_backward(dout, self.loss_scale(), need_dw=True, need_dx=False, want_inject_grads=inject)
            ctx = self._train[key] = {"graph": g, "fwd": fwd, "bwd": bwd, "dout": dout, "geom": (B, Cin, H, W)}
            ctx["fwd_halves"] = self._half_batch_plans(eng, B, H, W, wrap_pad, inject, stream)
        else:
            ctx["graph"].refresh_weights(backward=True, need_dx=False)
        return ctx


    def _half_batch_plans(self, eng, B, H, W, wrap_pad, inject, stream):
        """Forward plans for the two halves of the batch that write the full-batch graph's buffers in place (None when
        the batch is odd / small, disabled by NIRGAN_B200_TRAIN_SLICES=0, or a buffer is not per-image)."""
        import os
        from .engine import SliceBuffers
        # measured on B200: slower than the single full-batch plan at batch 32 (21.6 vs 20.9 ms per step) and at batch 64
        # (41.3 vs 39.4 ms) -- every conv launch pays its ~25 us pipeline fill twice and the eager launches of the two
        # plans do not interleave the way the graph-replayed inference slices do.  Off unless NIRGAN_B200_TRAIN_SLICES=1.
        mode = os.environ.get("NIRGAN_B200_TRAIN_SLICES", "0")
        if B % 2 or B < 16 or mode != "1":
            return None
        saved, plans = eng.buffers, []
        try:
            for part in range(2):
                eng.buffers = SliceBuffers(saved, part, 2)
                gh = self.build_graph(eng, B // 2, H, W, wrap_pad, inject, stream, "gt",
                                      direct_head=not self._use_tap_head())
                plans.append(gh.compile_forward())
                plans[-1].keepalive.append(gh)
        except KeyError:
            return None
        finally:
            eng.buffers = saved
        return plans


# =================================================================================================
class PatchGANRunner(_RunnerBase):
    """NLayerDiscriminator (model/networks.py:539-584)."""

    def conv_modules(self) -> List[torch.nn.Module]:
        return [m for m in self.module.model if isinstance(m, torch.nn.Conv2d)]

    def build_graph(self, eng: Engine, B: int, H: int, W: int, stream: int, tag: str) -> UnitGraph:
        convs = self.conv_modules()
        g = UnitGraph(eng, tag, stream)
        cin = convs[0].weight.shape[1]
        src = eng.buffers.get(tag + ".in", B * cin * H * W, torch.float32)
        g.records["src"] = src
        x = eng.act(tag + ".x0", B, H, W, 16, 0)
        g.pre_ops.append(("ng_prep_input", (src.data_ptr(), cin, None, 0, B, H, W, 0, 0, L.HALO_ZERO, 16, eng.dt_enum,
                                            x.t.data_ptr()), tag + ".prep"))
        c0 = convs[0]
        Hc, Wc = conv_out(H, 4, 2, 1), conv_out(W, 4, 2, 1)
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
        return g

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        require_cuda(x, "discriminator input")
        eng = self.engine(x.device)
        B, Cin, H, W = x.shape
        stream = torch.cuda.current_stream(x.device).cuda_stream
        key = (B, Cin, H, W)
        hit = self._fwd.get(key)
        if hit is None:
            g = self.build_graph(eng, B, H, W, stream, "d")
            hit = self._fwd[key] = (g, g.compile_forward())
        else:
            hit[0].refresh_weights()
        g, plan = hit
        plan.records["src"].view(B, Cin, H, W).copy_(x.float())
        plan.run(stream)
        Ho, Wo = g.records["out_hw"]
        self.last_plan = plan
        return g.units[-1].out_f32.view(B, 1, Ho, Wo).clone()

    def train_context(self, x: torch.Tensor, slot: int, need_dw: bool, need_dx: bool) -> dict:
        eng = self.engine(x.device)
        B, Cin, H, W = x.shape
        stream = torch.cuda.current_stream(x.device).cuda_stream
        key = (B, Cin, H, W, slot, need_dw, need_dx)
        ctx = self._train.get(key)
        if ctx is None:
            tag = f"dt{slot}"
            g = self.build_graph(eng, B, H, W, stream, tag)
            fwd = g.compile_forward()
            Ho, Wo = g.records["out_hw"]
            dout = eng.buffers.get(tag + ".dout", B * Ho * Wo, torch.float32)
            bwd = g.compile_backward(dout, self.loss_scale(), need_dw=need_dw, need_dx=need_dx)
            ctx = self._train[key] = {"graph": g, "fwd": fwd, "bwd": bwd, "dout": dout, "geom": (B, Cin, H, W)}
        else:
            ctx["graph"].refresh_weights(backward=True, need_dx=need_dx)
        return ctx
