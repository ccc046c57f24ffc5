"""Fused loss kernels behind autograd (LSGAN, L1 + the six remote-sensing indices, SSIM loss, EMD / histogram loss)."""
from __future__ import annotations

import torch

from . import _lib as L
from .engine import require_cuda


def _stream(t):
    return torch.cuda.current_stream(t.device).cuda_stream


class _LsganFn(torch.autograd.Function):
    """mean((p - target)^2); backward = 2 (p - target) / n (computed in the same kernel)."""

    @staticmethod
    def forward(ctx, pred, target):
        p = pred.contiguous().float()
        out = torch.empty(2, dtype=torch.float32, device=p.device)     # [0] loss, [1] scratch
        need_grad = pred.requires_grad
        grad = torch.empty_like(p) if need_grad else None
        L.call("ng_lsgan_loss", p.data_ptr(), p.numel(), float(target), out.data_ptr(), 0,
               grad.data_ptr() if need_grad else None, 1.0, _stream(p))
        if need_grad:
            ctx.save_for_backward(grad)
        ctx.shape = pred.shape
        return out[0]

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return (grad * g).view(ctx.shape), None


def lsgan(pred: torch.Tensor, target: float) -> torch.Tensor:
    require_cuda(pred, "lsgan prediction")
    return _LsganFn.apply(pred, target)


class _PixelLossFn(torch.autograd.Function):
    """One pass over (rgb, nir, pred): returns [L1, NDVI, NDWI, EVI] means and keeps
    d(sum_i w_i * loss_i)/dpred for the backward (weights fixed at call time)."""

    @staticmethod
    def forward(ctx, rgb, nir, pred, weights):
        B, _, H, W = pred.shape
        rgb, nir, p = rgb.contiguous().float(), nir.contiguous().float(), pred.contiguous().float()
        out = torch.empty(4, dtype=torch.float32, device=p.device)
        scratch = torch.empty(4 * 1024, dtype=torch.float32, device=p.device)
        need_grad = pred.requires_grad
        dpred = torch.empty_like(p) if need_grad else None
        w = (L.c_f32 * 4)(*[float(v) for v in weights])
        L.call("ng_g_pixel_losses", rgb.data_ptr(), nir.data_ptr(), p.data_ptr(), B, H * W, w, out.data_ptr(),
               dpred.data_ptr() if need_grad else None, scratch.data_ptr(), _stream(p))
        if need_grad:
            ctx.save_for_backward(dpred)
        ctx.weights = [float(v) for v in weights]
        return out

    @staticmethod
    def backward(ctx, gout):
        # out_i enters the total loss as w_i * out_i, so gout == weights (up to a common factor c):
        # dpred was computed for exactly that combination; recover c from the first non-zero weight.
        (dpred,) = ctx.saved_tensors
        c = None
        for i, w in enumerate(ctx.weights):
            if w != 0.0:
                c = gout[i] / w
                break
        if c is None:
            return None, None, torch.zeros_like(dpred), None
        return None, None, dpred * c, None


def pixel_losses(rgb, nir, pred, weights4):
    """weights4 = (w_L1, w_NDVI, w_NDWI, w_EVI) that the caller will apply to the four returned means."""
    for t, n in ((rgb, "rgb"), (nir, "nir"), (pred, "pred")):
        require_cuda(t, n)
    return _PixelLossFn.apply(rgb, nir, pred, tuple(weights4))


# term order of ng_rs_pixel_losses = the reference's iteration order (utils/remote_sensing_indices.py:45-52)
RS_TERMS = ("l1", "ndvi", "ndwi", "gndvi", "savi", "msavi", "evi")


class _RsPixelLossFn(torch.autograd.Function):
    """One pass over (rgb, nir, pred): [L1, NDVI, NDWI, GNDVI, SAVI, MSAVI, EVI] means (criterion l1 / l2 for the six
    indices) for the terms in `mask`, and d(sum_k w_k * term_k)/dpred for the backward (weights fixed at call time)."""

    @staticmethod
    def forward(ctx, rgb, nir, pred, weights, criterion, mask):
        B, _, H, W = pred.shape
        rgb, nir, p = rgb.contiguous().float(), nir.contiguous().float(), pred.contiguous().float()
        out = torch.empty(8, dtype=torch.float32, device=p.device)
        scratch = torch.empty(8 * 1024, dtype=torch.float32, device=p.device)
        need_grad = pred.requires_grad
        dpred = torch.empty_like(p) if need_grad else None
        w = (L.c_f32 * 7)(*[float(v) for v in weights])
        L.call("ng_rs_pixel_losses", rgb.data_ptr(), nir.data_ptr(), p.data_ptr(), B, H * W, w, int(criterion), int(mask),
               out.data_ptr(), dpred.data_ptr() if need_grad else None, scratch.data_ptr(), _stream(p))
        if need_grad:
            ctx.save_for_backward(dpred)
        ctx.weights = [float(v) for v in weights]
        return out[:7]

    @staticmethod
    def backward(ctx, gout):
        # term_k enters the total loss as w_k * term_k, so gout == weights up to a common factor c
        (dpred,) = ctx.saved_tensors
        c = None
        for i, w in enumerate(ctx.weights):
            if w != 0.0:
                c = gout[i] / w
                break
        if c is None:
            return None, None, torch.zeros_like(dpred), None, None, None
        return None, None, dpred * c, None, None, None


def rs_pixel_losses(rgb, nir, pred, weights7, criterion: str = "l1", mask: int = None):
    """weights7 = weights the caller will apply to (L1, NDVI, NDWI, GNDVI, SAVI, MSAVI, EVI); terms with a non-zero
    weight are evaluated (or exactly those in `mask`).  Returns the 7 means (0 for unselected terms)."""
    for t, n in ((rgb, "rgb"), (nir, "nir"), (pred, "pred")):
        require_cuda(t, n)
    if mask is None:
        mask = sum(1 << k for k, w in enumerate(weights7) if float(w) != 0.0)
    if mask == 0:
        return torch.zeros(7, dtype=torch.float32, device=pred.device)
    return _RsPixelLossFn.apply(rgb, nir, pred, tuple(weights7), 0 if criterion == "l1" else 1, int(mask))


def rs_index(rgb, nir, pred, which: str, loss_eps: bool):
    """Index maps (target, prediction) of one remote-sensing index, shape of `nir`."""
    for t, n in ((rgb, "rgb"), (nir, "nir"), (pred, "pred")):
        require_cuda(t, n)
    B, _, H, W = pred.shape
    rgb, nir, p = rgb.contiguous().float(), nir.contiguous().float(), pred.contiguous().float()
    a, b = torch.empty_like(p), torch.empty_like(p)
    L.call("ng_rs_index", rgb.data_ptr(), nir.data_ptr(), p.data_ptr(), B, H * W, RS_TERMS.index(which), int(loss_eps),
           a.data_ptr(), b.data_ptr(), _stream(p))
    return a, b


class _SsimLossFn(torch.autograd.Function):
    """1 - mean(kornia.metrics.ssim(img1, img2, window)); d/dimg1 from the adjoint-filter kernel (img2 is the target)."""

    @staticmethod
    def forward(ctx, img1, img2, window):
        B, Cn, H, W = img1.shape
        a, b = img1.contiguous().float(), img2.contiguous().float()
        n = L.load().ng_ssim_loss_scratch_floats(B * Cn, H, W)
        scratch = torch.empty(int(n), dtype=torch.float32, device=a.device)
        out = torch.empty(1, dtype=torch.float32, device=a.device)
        need_grad = img1.requires_grad
        grad = torch.empty_like(a) if need_grad else None
        L.call("ng_ssim_loss", a.data_ptr(), b.data_ptr(), B * Cn, H, W, int(window), 1.0, out.data_ptr(),
               grad.data_ptr() if need_grad else None, scratch.data_ptr(), _stream(a))
        if need_grad:
            ctx.save_for_backward(grad)
        return out[0]

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return grad * g, None, None


def ssim_loss(img1, img2, window_size: int = 11):
    """utils/losses.py:10-29.  Gradient flows to img1 (the prediction) only, as it is used in pix2pix.py:234."""
    require_cuda(img1, "img1")
    require_cuda(img2, "img2")
    if img2.requires_grad:
        raise NotImplementedError("ssim_loss: gradient w.r.t. the second image is not implemented")
    return _SsimLossFn.apply(img1, img2, int(window_size))


class _EmdLossFn(torch.autograd.Function):
    """mean |cumsum(softmax(pred_b)) - cumsum(softmax(target_b))|, one block per sample; d/dpred in the same launch."""

    @staticmethod
    def forward(ctx, pred, target):
        B = pred.shape[0]
        p, t = pred.contiguous().float().view(B, -1), target.contiguous().float().view(B, -1)
        scratch = torch.empty(B, dtype=torch.float64, device=p.device)
        out = torch.empty(1, dtype=torch.float32, device=p.device)
        need_grad = pred.requires_grad
        grad = torch.empty_like(p) if need_grad else None
        L.call("ng_emd_loss", p.data_ptr(), t.data_ptr(), B, p.shape[1], out.data_ptr(),
               grad.data_ptr() if need_grad else None, scratch.data_ptr(), _stream(p))
        if need_grad:
            ctx.save_for_backward(grad)
        ctx.shape = pred.shape
        return out[0]

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return (grad * g).view(ctx.shape), None


def emd_loss(pred, target):
    """utils/losses.py:64-78 (imported as hist_loss by pix2pix.py:13).  Inputs must be finite (the reference asserts it on
    the host; here non-finite values propagate to the loss instead of forcing a device sync)."""
    require_cuda(pred, "pred")
    require_cuda(target, "target")
    if target.requires_grad:
        raise NotImplementedError("emd_loss: gradient w.r.t. the target is not implemented")
    if pred.shape != target.shape:
        raise ValueError(f"emd_loss: shapes differ: {tuple(pred.shape)} vs {tuple(target.shape)}")
    return _EmdLossFn.apply(pred, target)
