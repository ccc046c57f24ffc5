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


def pixel_losses(rgb, nir, pred, weights4):
    """(L1, NDVI, NDWI, EVI) means of the shipped configuration; weights4 selects the terms to evaluate (zero weight =
    term skipped, returned as 0).  The gradient follows whatever the caller does with the four means."""
    w7 = (weights4[0], weights4[1], weights4[2], 0.0, 0.0, 0.0, weights4[3])
    t = rs_pixel_losses(rgb, nir, pred, w7)
    return torch.stack((t[0], t[1], t[2], t[6]))


# term order of ng_rs_pixel_losses = the reference's iteration order (utils/remote_sensing_indices.py:45-52)
RS_TERMS = ("l1", "ndvi", "ndwi", "gndvi", "savi", "msavi", "evi")


class _RsPixelLossFn(torch.autograd.Function):
    """One pass over (rgb, nir, pred): [L1, NDVI, NDWI, GNDVI, SAVI, MSAVI, EVI] means (criterion l1 / l2 for the six
    indices) for the terms in `mask`.  Backward: a second launch of the same kernel with the per-term upstream gradients
    as its (device-side) weights writes d(sum_k gout_k * term_k)/dpred -- whatever combination of the returned means the
    caller builds (model/pix2pix.py:222-247 weights them; 'logging_dict' mode uses them one by one), the gradient is
    that of the objective actually formed."""

    @staticmethod
    def forward(ctx, rgb, nir, pred, criterion, mask):
        B, _, H, W = pred.shape
        rgb, nir, p = rgb.contiguous().float(), nir.contiguous().float(), pred.contiguous().float()
        out = torch.empty(8, dtype=torch.float32, device=p.device)
        scratch = torch.empty(8 * 1024, dtype=torch.float32, device=p.device)
        w = (L.c_f32 * 7)(*([0.0] * 7))
        L.call("ng_rs_pixel_losses", rgb.data_ptr(), nir.data_ptr(), p.data_ptr(), B, H * W, w, None, int(criterion),
               int(mask), out.data_ptr(), None, scratch.data_ptr(), _stream(p))
        if pred.requires_grad:
            ctx.save_for_backward(rgb, nir, p)
            ctx.scratch = scratch
        ctx.criterion, ctx.mask = int(criterion), int(mask)
        return out[:7]

    @staticmethod
    def backward(ctx, gout):
        rgb, nir, p = ctx.saved_tensors
        B, _, H, W = p.shape
        g = gout.contiguous().float()
        dpred = torch.empty_like(p)
        out = torch.empty(8, dtype=torch.float32, device=p.device)
        L.call("ng_rs_pixel_losses", rgb.data_ptr(), nir.data_ptr(), p.data_ptr(), B, H * W, None, g.data_ptr(),
               ctx.criterion, ctx.mask, out.data_ptr(), dpred.data_ptr(), ctx.scratch.data_ptr(), _stream(p))
        return None, None, dpred, None, None


def rs_pixel_losses(rgb, nir, pred, weights7, criterion: str = "l1", mask: int = None):
    """weights7 = weights the caller will apply to (L1, NDVI, NDWI, GNDVI, SAVI, MSAVI, EVI); terms with a non-zero
    weight are evaluated (or exactly those in `mask`).  Returns the 7 means (0 for unselected terms)."""
    for t, n in ((rgb, "rgb"), (nir, "nir"), (pred, "pred")):
        require_cuda(t, n)
    if mask is None:
        mask = sum(1 << k for k, w in enumerate(weights7) if float(w) != 0.0)
    if mask == 0:
        return torch.zeros(7, dtype=torch.float32, device=pred.device)
    return _RsPixelLossFn.apply(rgb, nir, pred, 0 if criterion == "l1" else 1, int(mask))


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
