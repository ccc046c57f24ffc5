"""Unit graph: the hot path as a chain of fused units, compiled to forward and backward plans.

A *unit* is   conv -> [InstanceNorm -> inject -> activation (+ residual)] -> haloed output buffer
(or conv + bias + activation fused in the conv epilogue for the layers without a norm).  The same
description drives
  * the forward plan      (ng_conv2d, ng_in_stats[_finalize], ng_in_apply),
  * the backward plan     (ng_in_bwd, ng_conv2d_wgrad, ng_unpack_weight_grad, ng_conv2d as dgrad)
so the training step of model/pix2pix.py:165-257 runs entirely on the C-ABI kernels.  Data gradients
reuse the forward convolution kernels with re-packed weights:
  stride-1 conv            -> flipped-tap full correlation          (GATHER, sgn=-1)
  stride-2 conv            -> phase-decomposed transposed conv      (PHASED)
  ConvTranspose2d(s2)      -> stride-2 conv                         (GATHER, stride 2)
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional

import torch

from . import _lib as L
from .engine import ActBuf, Engine, Plan, _ptr


# NIRGAN_B200_MERGE_PHASES=0 runs ConvTranspose2d as four separate output-phase GEMMs (NG_FORM_PHASED) instead
import os as _os
MERGE_PHASES = [_os.environ.get("NIRGAN_B200_MERGE_PHASES", "1") != "0"]
# NIRGAN_B200_FUSED_FINALIZE=1 lets the conv kernel's last CTA per image finalise the InstanceNorm statistics instead of
# launching ng_in_stats_finalize.  Measured on B200: the per-tile __threadfence + counter arrival it needs costs far
# more than the ~10 us launches it saves (conv time +40 %), so it is OFF by default; kept as a tested switch.
FUSED_FINALIZE = [_os.environ.get("NIRGAN_B200_FUSED_FINALIZE", "0") == "1"]
# InstanceNorm statistics of the tcgen05 path: the conv epilogue adds every tile's per-channel (sum, sum of squares) to
# 64-bit fixed-point accumulators with integer atomics and the apply kernel derives (mean, rstd) from them itself, so a
# normalised unit is two launches (conv, apply) and the plan starts with ONE memset of all its accumulators.
# NIRGAN_B200_STAT_ACC=0: per-tile fp32 partials + an ng_in_stats_finalize launch per unit (round-1 form).
# Conv2d(64 -> 1, k7) + Tanh head as one kernel (ng_head_conv); NIRGAN_B200_HEAD_FUSED=0 keeps the tap GEMM + gather pair
# data gradient of the 64-input-channel stride-2 convolution in merged-phase form; NIRGAN_B200_DGRAD_MERGED=0: phased form
DGRAD_MERGED = [_os.environ.get("NIRGAN_B200_DGRAD_MERGED", "1") != "0"]
HEAD_FUSED = [_os.environ.get("NIRGAN_B200_HEAD_FUSED", "1") != "0"]
STAT_ACC = [_os.environ.get("NIRGAN_B200_STAT_ACC", "1") != "0"]


@dataclass
class Unit:
    name: str
    conv: torch.nn.Module                 # parameter holder (Conv2d / ConvTranspose2d)
    x: ActBuf                             # input buffer (haloed)
    cout: int                             # stored output channels
    K: int
    stride: int
    pad: int
    Hout: int
    Wout: int
    form: int = L.FORM_GATHER
    pack: object = 0                      # n_axis (0 / 1) or 'rowmerged'
    KW: Optional[int] = None
    pad_w: Optional[int] = None
    in_pad_w: Optional[int] = None
    kind: str = "norm"                    # 'norm' | 'biasact' | 'head'
    act: int = L.ACT_NONE
    slope: float = 0.0
    out_pad: int = 0
    halo_mode: int = L.HALO_REFLECT
    residual: Optional[int] = None        # index of the unit whose output buffer is added
    inject: Optional[dict] = None         # {'e': tensor, 'mode': int, 'scale': tensor}
    crop: int = 0
    # generator stem straight from the NCHW fp32 tiles (ng_stem_conv): {'src': tensor, 'cin', 'H', 'W', 'wrap'}; x is then a
    # shape-only description (no row-merged buffer exists)
    direct: Optional[dict] = None
    # filled by the builder
    y: Optional[ActBuf] = None
    mr: Optional[torch.Tensor] = None
    out: Optional[ActBuf] = None
    out_f32: Optional[torch.Tensor] = None
    wkey: tuple = ()


class UnitGraph:
    def __init__(self, eng: Engine, tag: str, stream: int = 0):
        self.eng, self.tag = eng, tag
        self.units: List[Unit] = []
        self.pre_ops = []                   # (fn_name, args, label) executed before the units (input prep, fc)
        self.records: dict = {}
        # single-output-channel 7x7 head as "tap GEMM + gather" (see ng_tap_gather / ng_tap_scatter):
        # {'conv': module, 'x': ActBuf (haloed head input), 'crop': int, 'B','H','W': output geometry}
        self.tap_head: Optional[dict] = None

    @property
    def stream(self) -> int:
        """Weight (re-)packing runs on the stream that is current at CALL time: the plan's kernels are launched on the
        caller's current stream too, so packing is ordered before them whichever stream the module is used from."""
        return torch.cuda.current_stream(self.eng.device).cuda_stream

    # ---- construction -----------------------------------------------------------------------------
    def merged_phases(self, u: Unit) -> bool:
        """ConvTranspose2d(k3, s2, p1, op1) forward on the tcgen05 kernel: merge the four output phases into GEMM-N."""
        return (u.form == L.FORM_PHASED and self.eng.impl == L.IMPL_TC and u.kind == "norm" and u.K == 3 and u.stride == 2
                and u.pad == 1 and u.KW is None and u.x.pad == 0 and u.x.C % 64 == 0 and (4 * u.cout) % 64 == 0
                and u.Hout == 2 * u.x.H and u.Wout == 2 * u.x.W and MERGE_PHASES[0])

    def tap_conv(self, u: Unit) -> bool:
        """PatchGAN last layer (512 -> 1, k4, p1; zero padding by TMA fill) through the one-kernel tap form (ng_head_conv)."""
        return (HEAD_FUSED[0] and u.kind == "head" and self.eng.impl == L.IMPL_TC and self.eng.dt_enum != L.F32
                and u.K == 4 and u.KW in (0, 4, None) and u.stride == 1 and u.pad == 1 and u.x.C == 512 and u.x.pad in (0, 1)
                and u.crop == 0 and u.act == L.ACT_NONE)

    def weight(self, u: Unit) -> torch.Tensor:
        cin = u.x.C
        if self.tap_conv(u):
            return self.eng.packed_weight(u.conv.weight, "taps", 16, cin, self.stream)
        if self.merged_phases(u):
            return self.eng.packed_weight(u.conv.weight, "phasemerged", u.cout, cin, self.stream)
        if u.pack == "rowmerged" and u.direct is None:
            return self.eng.packed_weight(u.conv.weight, "rowmerged", u.cout, 64, self.stream)
        if u.pack == "rowmerged4" or u.direct is not None:        # ng_stem_conv reads the 32-wide row-merged weights
            return self.eng.packed_weight(u.conv.weight, "rowmerged4", u.cout, 32, self.stream)
        if u.pack == "s2d":
            return self.eng.packed_weight(u.conv.weight, "s2d", u.cout, 64, self.stream)
        return self.eng.packed_weight(u.conv.weight, u.pack, u.cout, cin, self.stream)

    def add(self, u: Unit) -> Unit:
        eng, B = self.eng, u.x.B
        pre = f"{self.tag}.{u.name}"
        if u.kind == "head":
            u.out_f32 = eng.buffers.get(pre + ".out", B * (u.Hout - 2 * u.crop) * (u.Wout - 2 * u.crop), torch.float32)
        else:
            u.y = eng.act(pre + ".y", B, u.Hout, u.Wout, u.cout, 0)
            if u.kind == "norm":
                u.mr = eng.buffers.get(pre + ".mr", B * u.cout * 2, torch.float32)
                u.out = eng.act(pre + ".o", B, u.Hout, u.Wout, u.cout, u.out_pad)
            else:
                u.out = u.y                  # conv epilogue already produced the activation (pad 0)
        self.units.append(u)
        return u

    def _args(self, u: Unit, x: ActBuf, w: torch.Tensor, y: torch.Tensor, **over) -> L.ConvArgs:
        kw = dict(form=u.form, sgn=1, KW=u.KW, pad_w=u.pad_w, in_pad_w=u.in_pad_w)
        kw.update(over)
        return self.eng.conv_args(x, w, y, u.cout, u.K, u.stride, u.pad, u.Hout, u.Wout, **kw)

    # ---- forward ------------------------------------------------------------------------------------
    def compile_forward(self) -> Plan:
        eng = self.eng
        plan = Plan()
        plan.records.update(self.records)
        plan.records["weights"] = [u for u in self.units]
        use_acc = eng.impl == L.IMPL_TC and STAT_ACC[0] and not FUSED_FINALIZE[0]
        acc_off, acc_all = {}, None
        if use_acc:
            total = 0
            for i, u in enumerate(self.units):
                if u.kind == "norm":
                    acc_off[i] = total
                    total += u.x.B * u.cout * 2
            if total:
                acc_all = eng.buffers.get(self.tag + ".statacc", total, torch.int64)
                plan.add("ng_memset_zero", acc_all.data_ptr(), total * 8, label=self.tag + ".statacc.zero")
        for name, args, label in self.pre_ops:
            plan.add(name, *args, label=label)
        for ui, u in enumerate(self.units):
            w = self.weight(u)
            pre = f"{self.tag}.{u.name}"
            if u.kind == "head" and self.tap_conv(u):
                plan.add("ng_head_conv", u.x.t.data_ptr(), eng.dt_enum, u.x.B, u.Hout, u.Wout, u.x.C, u.K, u.pad, u.x.pad,
                         w.data_ptr(),
                         u.conv.bias.data_ptr(), u.act, 0, u.out_f32.data_ptr(), label=pre)
                continue
            if u.kind == "head":
                a = self._args(u, u.x, w, u.out_f32, epilogue=L.EPI_HEAD, act=u.act, crop=u.crop, bias=u.conv.bias.data)
                plan.keepalive.append(a)
                plan.add("ng_conv2d", C.byref(a), label=pre)
                continue
            if u.kind == "biasact":
                a = self._args(u, u.x, w, u.y.t, epilogue=L.EPI_BIAS_ACT, act=u.act, slope=u.slope,
                               bias=u.conv.bias.data)
                plan.keepalive.append(a)
                plan.add("ng_conv2d", C.byref(a), label=pre)
                continue
            B = u.x.B
            acc_ptr = None
            if u.direct is not None:
                # the stem straight from the fp32 tiles: im2col tile assembled in shared memory inside the conv
                d = u.direct
                part = None
                if use_acc:
                    acc_ptr = acc_all.data_ptr() + 8 * acc_off[ui]
                else:
                    slots = int(L.load().ng_stem_conv_stat_slots(d["H"], d["W"], d["wrap"]))
                    part = eng.buffers.get(pre + ".part", B * slots * u.cout * 2, torch.float32)
                plan.add("ng_stem_conv", d["src"].data_ptr(), d["cin"], B, d["H"], d["W"], d["wrap"], w.data_ptr(),
                         eng.dt_enum, u.y.t.data_ptr(), _ptr(part), acc_ptr, label=pre)
                if not use_acc:
                    plan.add("ng_in_stats_finalize", part.data_ptr(), B, slots, u.cout, u.Hout * u.Wout,
                             u.mr.data_ptr(), label=pre + ".fin")
            else:
                a = self._args(u, u.x, w, u.y.t, **({"form": L.FORM_PHASED_MERGED} if self.merged_phases(u) else {}))
            if u.direct is not None:
                pass
            elif use_acc:
                acc_ptr = acc_all.data_ptr() + 8 * acc_off[ui]
                a.stat_acc = acc_ptr
                plan.keepalive.append(a)
                plan.add("ng_conv2d", C.byref(a), label=pre)
            elif eng.impl == L.IMPL_TC:
                slots = L.load().ng_conv_stat_slots(C.byref(a))
                if slots <= 0:
                    L.check(slots if slots < 0 else -1, "ng_conv_stat_slots")
                part = eng.buffers.get(pre + ".part", B * slots * u.cout * 2, torch.float32)
                a.stat_partials = part.data_ptr()
                plan.keepalive.append(a)
                if FUSED_FINALIZE[0]:
                    # the conv kernel's last CTA per image reduces the partials to (mean, rstd) itself
                    cnt = eng.buffers.get(pre + ".cnt", B, torch.int32, zero=True)
                    a.mean_rstd, a.tile_counters = u.mr.data_ptr(), cnt.data_ptr()
                    plan.add("ng_conv2d", C.byref(a), label=pre)
                else:
                    plan.add("ng_conv2d", C.byref(a), label=pre)
                    plan.add("ng_in_stats_finalize", part.data_ptr(), B, slots, u.cout, u.Hout * u.Wout,
                             u.mr.data_ptr(), label=pre + ".fin")
            else:
                plan.keepalive.append(a)
                plan.add("ng_conv2d", C.byref(a), label=pre)
                plan.add("ng_in_stats", u.y.t.data_ptr(), eng.dt_enum, B, u.Hout * u.Wout, u.cout, u.mr.data_ptr(),
                         label=pre + ".stats")
            res = self.units[u.residual].out if u.residual is not None else None
            inj = u.inject or {}
            plan.add("ng_in_apply", u.y.t.data_ptr(), eng.dt_enum, B, u.Hout, u.Wout, u.cout,
                     None if use_acc else u.mr.data_ptr(), acc_ptr, u.mr.data_ptr() if use_acc else None, u.act,
                     u.slope, _ptr(res.t) if res else None, res.pad if res else 0, _ptr(inj.get("e")),
                     inj.get("mode", L.INJECT_NONE), _ptr(inj.get("scale")), u.out.t.data_ptr(), u.out_pad, u.halo_mode,
                     label=pre + ".apply")
        if self.tap_head is not None:
            # z[pixel][tap] = <x[pixel,:], w[tap,:]> over the haloed buffer (each input pixel read once, no 49x im2col
            # re-read), then out = tanh(b + sum_t z[(y+kh, x+kw), t])
            th = self.tap_head
            x, head = th["x"], th["conv"]
            K = head.weight.shape[-1]
            Hz, Wz = x.H + 2 * x.pad, x.W + 2 * x.pad
            out = eng.buffers.get(self.tag + ".head.out", x.B * th["H"] * th["W"], torch.float32)
            wt = eng.packed_weight(head.weight, "taps", 64, x.C, self.stream)
            if HEAD_FUSED[0] and eng.impl == L.IMPL_TC and eng.dt_enum != L.F32 and K == 7 and x.C == 64 and x.pad == 3:
                # one kernel: the z tile of an 8 x 16 output patch stays in TMEM / shared memory (ng_head_conv)
                plan.add("ng_head_conv", x.t.data_ptr(), eng.dt_enum, x.B, x.H, x.W, 64, 7, 3, 3, wt.data_ptr(),
                         head.bias.data_ptr(), L.ACT_TANH, th["crop"], out.data_ptr(), label=self.tag + ".head")
            else:
                xz = ActBuf(x.t, x.B, Hz, Wz, x.C, 0)
                z = eng.act(self.tag + ".z", x.B, Hz, Wz, 64, 0)
                a = eng.conv_args(xz, wt, z.t, 64, 1, 1, 0, Hz, Wz)
                plan.keepalive.append(a)
                plan.add("ng_conv2d", C.byref(a), label=self.tag + ".head.gemm")
                plan.add("ng_tap_gather", z.t.data_ptr(), eng.dt_enum, x.B, Hz, Wz, 64, K, K, head.bias.data_ptr(),
                         L.ACT_TANH, th["crop"], out.data_ptr(), label=self.tag + ".head.gather")
            th["out"] = out
            plan.records["out"] = out
        return plan

    def refresh_weights(self, backward: bool = False, need_dx: bool = False):
        """Re-pack the low-precision weight shadows whose fp32 masters changed (no-op otherwise)."""
        for i, u in enumerate(self.units):
            self.weight(u)
            if backward and (i > 0 or need_dx):
                self._dgrad_weight(u)
        if self.tap_head is not None:
            th = self.tap_head
            self.eng.packed_weight(th["conv"].weight, "taps", 64, th["x"].C, self.stream)
            if backward:
                self.eng.packed_weight(th["conv"].weight, "taps_T", th["x"].C, 64, self.stream)

    # ---- backward -------------------------------------------------------------------------------------
    def dgrad_merged(self, u: Unit) -> bool:
        """Data gradient of Conv2d(k3, s2, p1) = ConvTranspose2d(k3, s2, p1, op1) of dY with the SAME weight tensor read as
        (in = Cout, out = Cin, 3, 3): the merged-phase form of the forward up-convolutions applies (the four output
        parities as GEMM-N = 4 * Cin instead of four launches-in-one of N = Cin MMAs).  Taken where it is wider than the
        plain phased form: Cin = 64 (d1: N = 64 -> 256, 0.30 -> 0.2 ms); at Cin >= 128 the phased form is already N >= 128."""
        return (DGRAD_MERGED[0] and MERGE_PHASES[0] and u.form != L.FORM_PHASED and self.eng.impl == L.IMPL_TC
                and self.eng.dt_enum != L.F32 and u.kind == "norm" and u.K == 3 and u.stride == 2 and u.pad == 1
                and u.KW is None and u.x.pad == 0 and u.x.C == 64 and u.cout % 64 == 0
                and u.x.H == 2 * u.Hout and u.x.W == 2 * u.Wout)

    def _dgrad_weight(self, u: Unit) -> torch.Tensor:
        """Weights re-packed for the data gradient: roles of Cin / Cout swapped."""
        w = u.conv.weight
        cin = u.x.C
        if self.dgrad_merged(u):
            return self.eng.packed_weight(w, "phasemerged", cin, u.cout, self.stream)
        if u.pack == "s2d":                         # space-to-depth input layer: n = the 64 input slots, k = Cout
            return self.eng.packed_weight(w, "s2d_T", cin, u.cout, self.stream)
        if u.form == L.FORM_PHASED:                 # ConvTranspose2d (Cin, Cout, k, k): n = Cin, k = Cout
            return self.eng.packed_weight(w, 0, cin, u.cout, self.stream)
        return self.eng.packed_weight(w, 1, cin, u.cout, self.stream)       # Conv2d (Cout, Cin, k, k): n = Cin

    def _export(self, plan: Plan, exports: list, p: torch.nn.Parameter, fn_name: str, packed_ptr: int, dims: tuple,
                dev_inv: Optional[int], label: str):
        """Gradient export as part of the plan (side stream, right behind the weight gradient it unpacks): packed fp32 ->
        the parameter's slot of the flat gradient arena in the reference layout, times the inverse of the adaptive
        gradient scale.  The accumulate factor is a mutable ctypes cell set per backward pass (train.py): 0 = overwrite,
        1 = add (parameter reached twice in one pass, or .grad not cleared)."""
        slot = p._b200_grad_slot
        beta = C.c_float(0.0)
        plan.add(fn_name, packed_ptr, *dims, 1.0, dev_inv, beta, slot.data_ptr(), label=label, side=True)
        exports.append((p, beta))

    def compile_backward(self, dout_f32: torch.Tensor, loss_scale: float, need_dw: bool, need_dx: bool,
                         want_inject_grads: bool = False, hook_units=()) -> Plan:
        """dout_f32: gradient of the fp32 single-channel output of the last (head) unit; loss_scale: 0 = none, else
        the target max |gradient| of the adaptive power-of-two scaling (fp16 mode).  Returns a plan whose
        records hold 'dw' {unit index: packed fp32 grad}, 'db' {unit index: bias grad}, 'dx' (ActBuf-shaped grad of the
        first unit's input buffer, when need_dx), 'exports' [(parameter, accumulate cell)] for the gradients the plan
        writes into the flat arena itself and 'zero_grads' [parameters whose gradient is identically zero].
        hook_units: unit indices after whose export a ("grads_ready", arena offset) hook is placed (data-parallel
        buckets: every gradient at or above that arena offset is final)."""
        eng, tag = self.eng, self.tag
        plan = Plan()
        exports, zero_grads = [], []
        plan.records["exports"], plan.records["zero_grads"] = exports, zero_grads
        S = 1.0
        gsc = None
        if loss_scale:
            # 16-bit (fp16) gradients: adaptive power-of-two scale f (device scalar) so that max |dout * f| is
            # `loss_scale`; gradients are exported times 1/f.  No host round trip.
            gsc = eng.buffers.get(tag + ".gscale", 4, torch.float32)
            plan.add("ng_grad_scale_pow2", dout_f32.data_ptr(), dout_f32.numel(), float(loss_scale), gsc.data_ptr(),
                     launches=2, label=tag + ".gscale")
        plan.records["gscale"] = gsc
        dev_inv = gsc.data_ptr() + 4 if gsc is not None else None
        units = self.units
        n = len(units)
        g_halo = [None] * n         # gradient w.r.t. each unit's haloed output buffer (from the next unit's dgrad)
        g_skip = [None] * n         # gradient w.r.t. each unit's output interior from a residual consumer
        dw, db = {}, {}
        dY = None
        # one split-K workspace shared by every tcgen05 weight-gradient launch of this plan (they run back to back on
        # one stream), sized for the largest unit
        ws_bytes = 0
        if need_dw:
            for u in units:
                q = self._args(u, u.x, u.x.t, u.x.t)
                need = L.load().ng_conv2d_wgrad_workspace_bytes(C.byref(q))
                if need < 0:
                    L.check(int(need), "ng_conv2d_wgrad_workspace_bytes")
                ws_bytes = max(ws_bytes, int(need))
        th = self.tap_head
        if th is not None:
            x = th["x"]
            Hz, Wz = x.H + 2 * x.pad, x.W + 2 * x.pad
            th_B = x.B
            xz = ActBuf(x.t, th_B, Hz, Wz, x.C, 0)
            dz = eng.act(tag + ".head.dz", th_B, Hz, Wz, 64, 0)
            if need_dw:
                q = eng.conv_args(xz, x.t, dz.t, 64, 1, 1, 0, Hz, Wz)
                ws_bytes = max(ws_bytes, int(L.load().ng_conv2d_wgrad_workspace_bytes(C.byref(q))))
        ws = eng.buffers.get(tag + ".wgrad_ws", max(ws_bytes // 4, 4), torch.float32)
        if th is not None:
            head = th["conv"]
            K = head.weight.shape[-1]
            plan.add("ng_tap_scatter", dout_f32.data_ptr(), th["out"].data_ptr(), th_B, Hz, Wz, 64, K, K, L.ACT_TANH,
                     th["crop"], S, _ptr(gsc), eng.dt_enum, dz.t.data_ptr(), label=tag + ".head.dscatter")
            if need_dw:
                dwp = eng.buffers.get(tag + ".head.dwp", 64 * x.C, torch.float32)          # [tap (64)][k = channel]
                dbb = eng.buffers.get(tag + ".head.db", 64, torch.float32)
                a = eng.conv_args(xz, dwp, dz.t, 64, 1, 1, 0, Hz, Wz)
                plan.keepalive.append(a)
                plan.add("ng_conv2d_wgrad", C.byref(a), dwp.data_ptr(), dbb.data_ptr(), ws.data_ptr(), ws.numel() * 4,
                         launches=3, label=tag + ".head.wgrad", side=True)
                plan.records["tap_head"] = {"conv": head, "dwp": dwp, "db": dbb, "center": (K // 2) * K + K // 2}
                # head as tap GEMM: dwp is [tap (64 stored)][channel]; the reference layout (1, C, kh, kw) is
                # [channel][tap].  Every tap column of dz sums to sum(dy): the centre tap's column sum is the bias gradient
                cin_h = head.weight.shape[1]
                self._export(plan, exports, head.weight, "ng_unpack_weight_grad", dwp.data_ptr(),
                             (cin_h, K * K, 1, 1, 1, 64, cin_h), dev_inv, tag + ".head.export")
                center = (K // 2) * K + K // 2
                self._export(plan, exports, head.bias, "ng_unpack_weight_grad", dbb.data_ptr() + 4 * center,
                             (1, 1, 1, 1, 0, 1, 1), dev_inv, tag + ".head.export_b")
            # data gradient: g[pixel][c] = sum_t dz[pixel][t] * w[t][c]  (1x1 conv over the 64 stored taps)
            gbuf = eng.act(tag + ".head.dx", x.B, x.H, x.W, x.C, x.pad)
            wT = eng.packed_weight(head.weight, "taps_T", x.C, 64, self.stream)
            a = eng.conv_args(ActBuf(dz.t, th_B, Hz, Wz, 64, 0), wT, gbuf.t, x.C, 1, 1, 0, Hz, Wz)
            plan.keepalive.append(a)
            plan.add("ng_conv2d", C.byref(a), label=tag + ".head.dgrad")
            g_halo[n - 1] = gbuf
        for i in range(n - 1, -1, -1):
            u = units[i]
            pre = f"{tag}.{u.name}"
            B = u.x.B
            # ---- gradient of the conv output ----
            if u.kind == "head":
                dY = eng.act(pre + ".dy", B, u.Hout, u.Wout, u.cout, 0)
                plan.add("ng_head_bwd_prep", dout_f32.data_ptr(), _ptr(u.out_f32), B, u.Hout, u.Wout, u.crop, u.act, S,
                         _ptr(gsc), u.cout, eng.dt_enum, dY.t.data_ptr(), label=pre + ".dprep")
            else:
                dY = eng.act(pre + ".dy", B, u.Hout, u.Wout, u.cout, 0)
                gh, gs = g_halo[i], g_skip[i]
                assert gh is not None or gs is not None, f"unit {u.name} receives no gradient"
                do_out = None
                if u.residual is not None:
                    do_out = eng.act(pre + ".do", B, u.Hout, u.Wout, u.cout, 0)
                    g_skip[u.residual] = do_out
                inj = u.inject or {}
                sums = None
                if u.kind == "norm":
                    nf = int(L.load().ng_in_bwd_scratch_floats(B, u.Hout, u.Wout, u.cout))
                    if nf <= 0:
                        L.check(nf if nf < 0 else -1, "ng_in_bwd_scratch_floats")
                    sums = eng.buffers.get(pre + ".bsum", nf, torch.float32)
                dscale = de_map = None
                if inj and want_inject_grads:
                    dscale = eng.buffers.get(pre + ".dscale", 1, torch.float32)
                    de_map = eng.buffers.get(pre + ".demap", B * u.Hout * u.Wout, torch.float32)
                    plan.records["inject"] = {"dscale": dscale, "de_map": de_map, "H": u.Hout, "W": u.Wout, "unit": i}
                # for 'biasact' units y holds the activated output: its sign is the activation mask
                plan.add("ng_in_bwd", _ptr(gh.t) if gh else None, u.out_pad if gh else 0, u.halo_mode,
                         _ptr(gs.t) if gs else None, u.y.t.data_ptr(), eng.dt_enum, B, u.Hout, u.Wout, u.cout,
                         _ptr(u.mr) if u.kind == "norm" else None, u.act, u.slope, _ptr(inj.get("e")),
                         inj.get("mode", L.INJECT_NONE), _ptr(inj.get("scale")), _ptr(sums), dY.t.data_ptr(),
                         _ptr(do_out.t) if do_out else None, _ptr(dscale), _ptr(de_map), launches=2, label=pre + ".bwd")
            # ---- weight gradient ----
            if need_dw:
                taps = u.K * (u.KW if u.KW is not None else u.K)
                kdim = 64 if u.pack == "rowmerged" else u.x.C
                dwp = eng.buffers.get(pre + ".dwp", taps * u.cout * kdim, torch.float32)
                has_bias_grad = u.kind in ("head", "biasact")
                dbb = eng.buffers.get(pre + ".db", u.cout, torch.float32) if has_bias_grad else None
                # a.w is unused by wgrad; a.y = dY.  EPI_HEAD marks a single real output channel (dedicated kernel)
                a = self._args(u, u.x, dwp, dY.t, epilogue=L.EPI_HEAD if u.kind == "head" else L.EPI_RAW)
                plan.keepalive.append(a)
                plan.add("ng_conv2d_wgrad", C.byref(a), dwp.data_ptr(), _ptr(dbb), ws.data_ptr(), ws.numel() * 4, launches=2,
                         label=pre + ".wgrad", side=True)
                dw[i], db[i] = dwp, dbb
                w = u.conv.weight
                d0, d1, kh, kw = w.shape
                if u.pack == "rowmerged":
                    self._export(plan, exports, w, "ng_unpack_weight_grad_rowmerged", dwp.data_ptr(), (d0, d1, kh, kw),
                                 dev_inv, pre + ".export")
                elif u.pack == "s2d":
                    self._export(plan, exports, w, "ng_unpack_weight_grad_s2d", dwp.data_ptr(), (d0, d1), dev_inv,
                                 pre + ".export")
                else:
                    self._export(plan, exports, w, "ng_unpack_weight_grad", dwp.data_ptr(),
                                 (d0, d1, kh, kw, u.pack, u.cout, u.x.C), dev_inv, pre + ".export")
                if u.conv.bias is not None:
                    if dbb is not None:
                        nb = u.conv.bias.numel()
                        self._export(plan, exports, u.conv.bias, "ng_unpack_weight_grad", dbb.data_ptr(),
                                     (nb, 1, 1, 1, 0, u.cout, 1), dev_inv, pre + ".export_b")
                    else:
                        # bias feeding InstanceNorm: its gradient is identically zero (the reference returns rounding
                        # noise); its arena slot is never written and stays zero
                        zero_grads.append(u.conv.bias)
                if i in hook_units:
                    plan.add_hook(("grads_ready", w._b200_grad_arena.offset_of(w)), label=pre + ".ddp_hook")
            # ---- data gradient ----
            if i == 0 and not need_dx:
                continue
            xin = u.x
            g = eng.act(pre + ".dx", xin.B, xin.H, xin.W, xin.C, xin.pad)
            wd = self._dgrad_weight(u)
            dyb = ActBuf(dY.t, B, u.Hout, u.Wout, u.cout, 0)
            Kw = u.KW if u.KW is not None else u.K
            assert Kw == u.K, "data gradient of asymmetric kernels is not needed on this path"
            Hg, Wg = xin.H + 2 * xin.pad, xin.W + 2 * xin.pad
            if u.form == L.FORM_PHASED:            # ConvTranspose -> stride-2 conv of dY
                a = eng.conv_args(dyb, wd, g.t, xin.C, u.K, 2, u.pad, Hg, Wg)
            elif u.stride == 2:                     # strided conv -> phased transposed conv
                a = eng.conv_args(dyb, wd, g.t, xin.C, u.K, 2, u.pad, Hg, Wg,
                                  form=L.FORM_PHASED_MERGED if self.dgrad_merged(u) else L.FORM_PHASED)
            else:                                   # stride-1 conv -> flipped full correlation
                # haloed input (in_pad == pad): gradient of the haloed buffer, pad 0;  zero-padded input: pad = pad
                eff_pad = 0 if xin.pad == u.pad else u.pad
                a = eng.conv_args(dyb, wd, g.t, xin.C, u.K, 1, eff_pad, Hg, Wg, sgn=-1)
            if xin.C < 64 and xin.C != 16:
                a.impl = L.IMPL_SIMT          # untileable thin data gradients: CUDA-core kernel
            plan.keepalive.append(a)
            plan.add("ng_conv2d", C.byref(a), label=pre + ".dgrad")
            if i > 0:
                g_halo[i - 1] = g
            else:
                plan.records["dx"] = g
        plan.records["dw"], plan.records["db"] = dw, db
        return plan
