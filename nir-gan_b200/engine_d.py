"""PatchGAN (NLayerDiscriminator, model/networks.py:539-584) forward plan."""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Tuple

import torch

from . import _lib as L
from .engine import Engine, EngineConfig, Plan, conv_out, require_cuda


class PatchGANRunner:
    def __init__(self, module: torch.nn.Module, cfg: Optional[EngineConfig] = None):
        self.module = module
        self.cfg = cfg or EngineConfig.from_env()
        self._engine: Optional[Engine] = None
        self._plans: Dict[Tuple, Plan] = {}

    def engine(self, device) -> Engine:
        if self._engine is None or self._engine.device != device:
            self._engine = Engine(self.cfg, device)
            self._plans.clear()
        return self._engine

    def conv_modules(self):
        return [m for m in self.module.model if isinstance(m, torch.nn.Conv2d)]

    def _build(self, eng: Engine, B: int, H: int, W: int, stream: int, tag: str = "d") -> Plan:
        convs = self.conv_modules()
        plan = Plan()
        wrec = plan.records["weights"] = []

        def W_(conv, n_pad, k_pad):
            wrec.append((conv, 0, n_pad, k_pad))
            return eng.packed_weight(conv.weight, 0, n_pad, k_pad, stream)

        cin = convs[0].weight.shape[1]
        src = eng.buffers.get(tag + ".in", B * cin * H * W, torch.float32)
        plan.records["src"] = src
        x = eng.act(tag + ".x0", B, H, W, 16, 0)
        plan.add("ng_prep_input", src.data_ptr(), cin, None, 0, B, H, W, 0, 0, L.HALO_ZERO, 16, eng.dt_enum,
                 x.t.data_ptr())
        # layer 0: conv + bias + LeakyReLU (no norm)
        c0 = convs[0]
        ndf = c0.weight.shape[0]
        Hc, Wc = conv_out(H, 4, 2, 1), conv_out(W, 4, 2, 1)
        y0 = eng.act(tag + ".a0", B, Hc, Wc, ndf, 0)
        a = eng.conv_args(x, W_(c0, ndf, 16), y0.t, ndf, 4, 2, 1, Hc, Wc, epilogue=L.EPI_BIAS_ACT, act=L.ACT_LRELU,
                          slope=0.2, bias=c0.bias.data)
        plan.keepalive.append(a)
        plan.add("ng_conv2d", C.byref(a))
        x = y0
        # middle layers: conv (+bias, cancelled by IN) + InstanceNorm + LeakyReLU
        for i, conv in enumerate(convs[1:-1], start=1):
            stride = conv.stride[0]
            co, ci = conv.weight.shape[0], conv.weight.shape[1]
            Hn, Wn = conv_out(x.H, 4, stride, 1), conv_out(x.W, 4, stride, 1)
            y, mr = eng.add_conv_norm(plan, f"{tag}.l{i}", x, W_(conv, co, ci), co, 4, stride, 1, Hn, Wn)
            x = eng.add_apply(plan, f"{tag}.a{i}", y, mr, L.ACT_LRELU, 0, halo_mode=L.HALO_ZERO, slope=0.2)
        # last layer: conv + bias -> 1 channel, fp32
        cl = convs[-1]
        Ho, Wo = conv_out(x.H, 4, 1, 1), conv_out(x.W, 4, 1, 1)
        out = eng.buffers.get(tag + ".out", B * Ho * Wo, torch.float32)
        a = eng.conv_args(x, W_(cl, 16, cl.weight.shape[1]), out, 16, 4, 1, 1, Ho, Wo, epilogue=L.EPI_HEAD,
                          act=L.ACT_NONE, bias=cl.bias.data)
        plan.keepalive.append(a)
        plan.add("ng_conv2d", C.byref(a))
        plan.records["out"] = out
        plan.records["out_hw"] = (Ho, Wo)
        return plan

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        require_cuda(x, "discriminator input")
        eng = self.engine(x.device)
        B, Cin, H, W = x.shape
        stream = torch.cuda.current_stream(x.device).cuda_stream
        key = (B, Cin, H, W)
        plan = self._plans.get(key)
        if plan is None:
            plan = self._plans[key] = self._build(eng, B, H, W, stream)
        else:
            for conv, n_axis, n_pad, k_pad in plan.records["weights"]:
                eng.packed_weight(conv.weight, n_axis, n_pad, k_pad, stream)
        plan.records["src"].view(B, Cin, H, W).copy_(x.float())
        plan.run(stream)
        Ho, Wo = plan.records["out_hw"]
        self.last_plan = plan
        return plan.records["out"].view(B, 1, Ho, Wo).clone()
