"""B200 drop-in for the reference's ``utils/calculate_metrics.py``.

``calculate_metrics(pred, target, phase)`` keeps its signature and keys (``<phase>/L1``, ``/L2``, ``/PSNR``, ``/SSIM``)
but runs as one fused kernel on the device (``ng_image_metrics``: L1, MSE, PSNR and the mean kornia-style SSIM map with
a 5x5 Gaussian window), so the training loop no longer has to copy ``pred`` / ``nir`` to the host every logging step
(model/pix2pix.py:184 does ``.cpu()`` first); only four floats come back.
"""
from __future__ import annotations

import torch

from .. import _lib as L
from ..engine import require_cuda


def image_metrics(pred: torch.Tensor, target: torch.Tensor, window_size: int = 5, max_val: float = 1.0) -> torch.Tensor:
    """(L1, L2, PSNR, mean SSIM) as a 4-element CUDA tensor (no host synchronisation)."""
    require_cuda(pred, "metrics prediction")
    require_cuda(target, "metrics target")
    if pred.shape != target.shape or pred.dim() != 4:
        raise RuntimeError("calculate_metrics expects two [B, C, H, W] tensors of the same shape")
    B, Cn, H, W = pred.shape
    p, t = pred.detach().contiguous().float(), target.detach().contiguous().float()
    out = torch.empty(4, dtype=torch.float32, device=p.device)
    scratch = torch.empty(int(L.load().ng_image_metrics_scratch_floats(B * Cn, H, W)), dtype=torch.float32, device=p.device)
    L.call("ng_image_metrics", p.data_ptr(), t.data_ptr(), B * Cn, H, W, int(window_size), float(max_val), out.data_ptr(),
           scratch.data_ptr(), torch.cuda.current_stream(p.device).cuda_stream)
    return out


def calculate_metrics(pred, target, phase="train"):
    """utils/calculate_metrics.py:6-37."""
    m = image_metrics(pred, target, window_size=5, max_val=1.0).tolist()      # the one device -> host read
    return {phase + "/L1": m[0], phase + "/L2": m[1], phase + "/PSNR": m[2], phase + "/SSIM": m[3]}
