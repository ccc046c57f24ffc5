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
                            int(os.environ.get("NIRGAN_B200_CHUNK", "0")))


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


class Plan:
    """A compiled list of bound C-ABI calls."""

    def __init__(self):
        self.ops: List[Tuple[Callable, tuple]] = []
        self.keepalive: list = []
        self.launches = 0          # kernel launches per run (for bench.py's gpu_launches)
        self.records: dict = {}
        self.labels: List[str] = []

    def add(self, fn_name: str, *args, launches: int = 1, label: str = ""):
        fn = getattr(L.load(), fn_name)
        self.ops.append((fn, args, fn_name))
        self.labels.append(label or fn_name)
        self.launches += launches

    def run(self, stream_ptr: int):
        for fn, args, name in self.ops:
            st = fn(*args, stream_ptr)
            if st != 0:
                L.check(st, name)


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
        n_axis 'taps': [n = tap (padded to n_pad)][k_pad] for the tap-GEMM form of a 1-output-channel conv."""
        key = (w.data_ptr(), n_axis, n_pad, k_pad, self.dt_enum)
        ver = w._version
        hit = self._packed.get(key)
        if hit is not None and hit[0] == ver:
            return hit[1]
        d0, d1, kh, kw = w.shape
        src = w.detach()
        if src.dtype != torch.float32 or not src.is_contiguous():
            src = src.float().contiguous()
        if n_axis == "rowmerged":
            dst = hit[1] if hit is not None else torch.empty(kh * d0 * 64, dtype=self.dt_torch, device=self.device)
            L.call("ng_pack_weight_rowmerged", src.data_ptr(), d0, d1, kh, kw, self.dt_enum, dst.data_ptr(), stream)
        elif n_axis == "taps":
            assert d0 == 1 and kh * kw <= n_pad
            dst = hit[1] if hit is not None else torch.zeros(n_pad * k_pad, dtype=self.dt_torch, device=self.device)
            L.call("ng_pack_weight", src.data_ptr(), d0, d1, kh, kw, 0, 1, k_pad, self.dt_enum, dst.data_ptr(), stream)
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

    def add_conv_norm(self, plan: Plan, name: str, x: ActBuf, w_packed: torch.Tensor, Cout: int, K: int, stride: int,
                      pad: int, Hout: int, Wout: int, form=L.FORM_GATHER, sgn=1, **geom):
        """conv -> compact pre-norm Y + (mean, rstd).  Returns (Y ActBuf, mean_rstd tensor)."""
        y = self.act(name + ".y", x.B, Hout, Wout, Cout, 0)
        mr = self.buffers.get(name + ".mr", x.B * Cout * 2, torch.float32)
        a = self.conv_args(x, w_packed, y.t, Cout, K, stride, pad, Hout, Wout, form, sgn, **geom)
        if self.impl == L.IMPL_TC:
            slots = L.load().ng_conv_stat_slots(C.byref(a))
            if slots <= 0:
                L.check(slots if slots < 0 else -1, "ng_conv_stat_slots")
            part = self.buffers.get(name + ".part", x.B * slots * Cout * 2, torch.float32)
            a.stat_partials = part.data_ptr()
            plan.keepalive.append(a)
            plan.add("ng_conv2d", C.byref(a), label=name)
            plan.add("ng_in_stats_finalize", part.data_ptr(), x.B, slots, Cout, Hout * Wout, mr.data_ptr(),
                     label=name + ".fin")
        else:
            plan.keepalive.append(a)
            plan.add("ng_conv2d", C.byref(a), label=name)
            plan.add("ng_in_stats", y.t.data_ptr(), self.dt_enum, x.B, Hout * Wout, Cout, mr.data_ptr(),
                     label=name + ".stats")
        return y, mr

    def add_apply(self, plan: Plan, name: str, y: ActBuf, mr: Optional[torch.Tensor], act: int, out_pad: int,
                  halo_mode=L.HALO_REFLECT, slope=0.0, residual: Optional[ActBuf] = None,
                  inject_e: Optional[torch.Tensor] = None, inject_mode=L.INJECT_NONE,
                  inject_scale: Optional[torch.Tensor] = None) -> ActBuf:
        out = self.act(name, y.B, y.H, y.W, y.C, out_pad)
        plan.add("ng_in_apply", y.t.data_ptr(), self.dt_enum, y.B, y.H, y.W, y.C, _ptr(mr), act, slope,
                 _ptr(residual.t) if residual else None, residual.pad if residual else 0, _ptr(inject_e),
                 inject_mode, _ptr(inject_scale), out.t.data_ptr(), out_pad, halo_mode, label=name)
        return out


def conv_out(h: int, k: int, s: int, p: int) -> int:
    return (h + 2 * p - k) // s + 1


class GeneratorRunner:
    """Executes ResnetGenerator / ResnetGenerator_inject forward (model/networks.py:341-374,
    model/generator_inject.py:105-135) for a module that keeps the reference state_dict layout."""

    def __init__(self, module: torch.nn.Module, cfg: Optional[EngineConfig] = None):
        self.module = module
        self.cfg = cfg or EngineConfig.from_env()
        self._engine: Optional[Engine] = None
        self._plans: Dict[Tuple, Plan] = {}
        self.head_mode = os.environ.get("NIRGAN_B200_HEAD", "tapgemm")     # 'tapgemm' | 'direct'

    # lazily bound to the device of the first input
    def engine(self, device) -> Engine:
        if self._engine is None or self._engine.device != device:
            self._engine = Engine(self.cfg, device)
            self._plans.clear()
        return self._engine

    def _convs(self):
        m = self.module.model
        nb = self.module.n_blocks
        blocks = [m[10 + b] for b in range(nb)]
        return m[1], m[4], m[7], blocks, m[10 + nb], m[13 + nb], m[17 + nb]

    def _build(self, eng: Engine, B: int, H: int, W: int, wrap: int, inject: bool, stream: int) -> Plan:
        mod = self.module
        stem, d1, d2, blocks, u1, u2, head = self._convs()
        ngf = stem.weight.shape[0]
        cin = stem.weight.shape[1]
        assert ngf % 64 == 0, "nirgan_b200 kernels are tiled for ngf multiples of 64"
        plan = Plan()
        H1, W1 = H + 2 * wrap, W + 2 * wrap
        if H1 % 4 or W1 % 4:
            raise RuntimeError(f"nirgan_b200: padded tile {H1}x{W1} must be divisible by 4 "
                               f"(two stride-2 stages, as in the reference)")
        pw = lambda conv, n_axis, n_pad, k_pad: eng.packed_weight(conv.weight, n_axis, n_pad, k_pad, stream)
        plan.records["weights"] = []   # (conv module, n_axis, n_pad, k_pad) to re-pack when masters change
        wrec = plan.records["weights"]

        def W_(conv, n_axis, n_pad, k_pad):
            wrec.append((conv, n_axis, n_pad, k_pad))
            return pw(conv, n_axis, n_pad, k_pad)

        # input: NCHW fp32 -> row-merged NHWC [B][H1+6][W1][kw*8+c] (wrapper reflect pad + stem reflect halo fused):
        # the 7x7x3 stem becomes a 7x1 conv over 64 "channels" = 7 K-steps of 128-byte rows instead of 49 thin taps
        src = eng.buffers.get("g.in", B * cin * H * W, torch.float32)
        plan.records["src"] = src
        if cin > 8:
            raise NotImplementedError("nirgan_b200 stem kernel: input_nc <= 8")
        x0 = ActBuf(eng.buffers.get("g.x0", B * (H1 + 6) * W1 * 64, eng.dt_torch), B, H1, W1, 64, 3)
        plan.add("ng_prep_stem", src.data_ptr(), cin, B, H, W, wrap, 3, 7, eng.dt_enum, x0.t.data_ptr(),
                 label="g.prep")
        # stem (bias cancelled by InstanceNorm -> skipped)
        y, mr = eng.add_conv_norm(plan, "g.stem", x0, W_(stem, "rowmerged", ngf, 64), ngf, 7, 1, 3, H1, W1,
                                  KW=1, pad_w=0, in_pad_w=0)
        x = eng.add_apply(plan, "g.x1", y, mr, L.ACT_RELU, 0)
        # down 1 (+ SatCLIP injection between IN and ReLU)
        H2, W2 = conv_out(H1, 3, 2, 1), conv_out(W1, 3, 2, 1)
        y, mr = eng.add_conv_norm(plan, "g.d1", x, W_(d1, 0, 2 * ngf, ngf), 2 * ngf, 3, 2, 1, H2, W2)
        if inject:
            if H2 != W2:
                raise RuntimeError("nirgan_b200: SatCLIP injection is defined for square tiles only "
                                   "(generator_inject.py:116 passes size=(W,H))")
            emb = eng.buffers.get("g.emb", B * 256, torch.float32)
            e = eng.buffers.get("g.e", B * 128 * 128, torch.float32)
            plan.records["emb"] = emb
            plan.add("ng_linear", emb.data_ptr(), mod.fc.weight.data_ptr(), mod.fc.bias.data_ptr(), B, 256,
                     128 * 128, e.data_ptr())
            style = mod.inject_style
            # `elif inject_style == "multiply" and self.scale_param` truthiness quirk (generator_inject.py:124)
            # is resolved on the host at plan time.
            if style == "add":
                mode = L.INJECT_ADD
            elif style == "multiply":
                mode = L.INJECT_MUL_SCALED if bool(mod.scale_param) else L.INJECT_MUL   # one-time host read
            else:
                raise NotImplementedError(f"inject style [{style}] is not recognized")
            x = eng.add_apply(plan, "g.x2", y, mr, L.ACT_RELU, 0, inject_e=e, inject_mode=mode,
                              inject_scale=mod.scale_param.data)
            plan.records["inject_mode"] = mode
        else:
            x = eng.add_apply(plan, "g.x2", y, mr, L.ACT_RELU, 0)
        # down 2 -> first haloed (reflect 1) residual-trunk buffer
        H3, W3 = conv_out(H2, 3, 2, 1), conv_out(W2, 3, 2, 1)
        C4 = 4 * ngf
        y, mr = eng.add_conv_norm(plan, "g.d2", x, W_(d2, 0, C4, 2 * ngf), C4, 3, 2, 1, H3, W3)
        nb = len(blocks)
        x = eng.add_apply(plan, "g.t0", y, mr, L.ACT_RELU, 1 if nb else 0)
        for b, blk in enumerate(blocks):
            c1, c2 = blk.conv_block[1], blk.conv_block[5]
            y, mr = eng.add_conv_norm(plan, "g.ra", x, W_(c1, 0, C4, C4), C4, 3, 1, 1, H3, W3)
            a = eng.add_apply(plan, "g.ta", y, mr, L.ACT_RELU, 1)
            y, mr = eng.add_conv_norm(plan, "g.rb", a, W_(c2, 0, C4, C4), C4, 3, 1, 1, H3, W3)
            last = b == nb - 1
            # out = x + IN(conv2(...)); no ReLU after the add (networks.py:433)
            nxt = eng.add_apply(plan, f"g.t{(b + 1) % 2}" if not last else "g.tl", y, mr, L.ACT_NONE,
                                0 if last else 1, residual=x)
            x = nxt
        # up 1, up 2 (ConvTranspose k3 s2 p1 op1 as 4 output-parity phases)
        y, mr = eng.add_conv_norm(plan, "g.u1", x, W_(u1, 1, 2 * ngf, C4), 2 * ngf, 3, 2, 1, 2 * H3, 2 * W3,
                                  form=L.FORM_PHASED)
        x = eng.add_apply(plan, "g.x5", y, mr, L.ACT_RELU, 0)
        y, mr = eng.add_conv_norm(plan, "g.u2", x, W_(u2, 1, ngf, 2 * ngf), ngf, 3, 2, 1, 4 * H3, 4 * W3,
                                  form=L.FORM_PHASED)
        x = eng.add_apply(plan, "g.x6", y, mr, L.ACT_RELU, 3)
        # head 7x7 -> 1 channel (+bias, tanh), fp32 NCHW, wrapper crop fused.
        out = eng.buffers.get("g.out", B * H * W, torch.float32)
        if self.head_mode == "tapgemm" and ngf == 64:
            # tap GEMM + gather: z[pixel][tap] = <x[pixel,:], w[tap,:]> over the haloed buffer (each input pixel
            # read once, no 49x im2col re-read), then out = tanh(b + sum_t z[(y+kh, x+kw), t])
            xz = ActBuf(x.t, B, H1 + 6, W1 + 6, ngf, 0)
            z = eng.act("g.z", B, H1 + 6, W1 + 6, 64, 0)
            a = eng.conv_args(xz, W_(head, "taps", 64, ngf), z.t, 64, 1, 1, 0, H1 + 6, W1 + 6)
            plan.keepalive.append(a)
            plan.add("ng_conv2d", C.byref(a), label="g.head.gemm")
            plan.add("ng_tap_gather", z.t.data_ptr(), eng.dt_enum, B, H1 + 6, W1 + 6, 64, 7, 7, head.bias.data_ptr(),
                     L.ACT_TANH, wrap, out.data_ptr(), label="g.head.gather")
        else:
            a = eng.conv_args(x, W_(head, 0, 16, ngf), out, 16, 7, 1, 3, H1, W1, epilogue=L.EPI_HEAD,
                              act=L.ACT_TANH, crop=wrap, bias=head.bias.data)
            plan.keepalive.append(a)
            plan.add("ng_conv2d", C.byref(a), label="g.head")
        plan.records["out"] = out
        plan.records["post"] = getattr(mod, "post_correction", False)
        return plan

    def _refresh_weights(self, eng: Engine, plan: Plan, stream: int):
        for conv, n_axis, n_pad, k_pad in plan.records["weights"]:
            eng.packed_weight(conv.weight, n_axis, n_pad, k_pad, stream)

    @torch.no_grad()
    def forward(self, x: torch.Tensor, embeds: Optional[torch.Tensor] = None, wrap_pad: int = 0) -> torch.Tensor:
        require_cuda(x, "generator input")
        if x.dim() != 4:
            raise RuntimeError("generator input must be (B, C, H, W)")
        eng = self.engine(x.device)
        Btot, Cin, H, W = x.shape
        inject = embeds is not None
        chunk = eng.cfg.chunk if eng.cfg.chunk > 0 else Btot
        out = torch.empty(Btot, 1, H, W, dtype=torch.float32, device=x.device)
        stream = torch.cuda.current_stream(x.device).cuda_stream
        x = x.contiguous().float()
        if inject:
            require_cuda(embeds, "embeds")
            embeds = embeds.contiguous().float()
        for b0 in range(0, Btot, chunk):
            B = min(chunk, Btot - b0)
            key = (B, Cin, H, W, wrap_pad, inject)
            plan = self._plans.get(key)
            if plan is None:
                plan = self._build(eng, B, H, W, wrap_pad, inject, stream)
                self._plans[key] = plan
            else:
                self._refresh_weights(eng, plan, stream)
            plan.records["src"].view(B, Cin, H, W).copy_(x[b0:b0 + B])
            if inject:
                plan.records["emb"].view(B, 256).copy_(embeds[b0:b0 + B])
            plan.run(stream)
            o = plan.records["out"].view(B, 1, H, W)
            if plan.records["post"]:
                o = o * self.module.post_correction_param
            out[b0:b0 + B].copy_(o)
        self.last_plan = plan
        return out
