"""Device-side mirror of the post-processing in the reference's inference loop
(create_synthetic_dataset.py:34-52,111-118): nearest x4 upsampling of the Sentinel-2 NIR band, per-tile histogram
matching of the predicted NIR against it (``skimage.exposure.match_histograms``), float16 output and the ``.npz``
writer.  Everything up to the fp16 tensor runs in ``csrc/postprocess.cu``; only zlib + file IO stay on the host.
"""
from __future__ import annotations

import os
from concurrent.futures import ThreadPoolExecutor
from typing import Iterable, Optional, Sequence

import numpy as np
import torch

from . import _lib as L
from .engine import require_cuda


def _stream(t):
    return torch.cuda.current_stream(t.device).cuda_stream


def resize(x: torch.Tensor, size, mode: str = "nearest") -> torch.Tensor:
    """F.interpolate(x, size=size, mode=mode[, align_corners=False]) for fp32 (B,C,h,w) on the device."""
    require_cuda(x, "resize input")
    if mode not in ("nearest", "bilinear"):
        raise NotImplementedError(f"resize mode {mode}")
    B, Cn, h, w = x.shape
    H, W = size
    x = x.contiguous().float()
    out = torch.empty(B, Cn, H, W, dtype=torch.float32, device=x.device)
    L.call("ng_resize_plane", x.data_ptr(), B * Cn, h, w, H, W, 0 if mode == "nearest" else 1, out.data_ptr(), _stream(x))
    return out


def histogram_match(image: torch.Tensor, reference: torch.Tensor, out_dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """create_synthetic_dataset.py:34-47 on the device.  image (B,1,H,W), reference (B,1,h,w) -> (B,1,H,W).
    The reference is first resized to the image size with bilinear interpolation (align_corners=False), as in the
    reference code (an identity when the sizes already agree)."""
    require_cuda(image, "histogram_match image")
    require_cuda(reference, "histogram_match reference")
    if image.dim() != 4 or image.shape[1] != 1 or reference.dim() != 4 or reference.shape[1] != 1:
        raise RuntimeError("histogram_match expects (B,1,H,W) image and (B,1,h,w) reference")
    if out_dtype not in (torch.float32, torch.float16):
        raise NotImplementedError("histogram_match output must be float32 or float16")
    B, _, H, W = image.shape
    if reference.shape[0] != B:
        raise RuntimeError("histogram_match: batch sizes differ")
    if tuple(reference.shape[-2:]) != (H, W):
        reference = resize(reference, (H, W), "bilinear")
    image = image.contiguous().float()
    reference = reference.contiguous().float()
    N = H * W
    need = int(L.load().ng_hist_match_workspace_bytes(B, N, N))
    ws = torch.empty(need // 4 + 1, dtype=torch.int32, device=image.device)
    out = torch.empty(B, 1, H, W, dtype=out_dtype, device=image.device)
    L.call("ng_hist_match", image.data_ptr(), reference.data_ptr(), B, N, N, L.F16 if out_dtype == torch.float16 else L.F32,
           out.data_ptr(), ws.data_ptr(), need, _stream(image))
    return out


def postprocess(pred: torch.Tensor, s2_nir: torch.Tensor) -> torch.Tensor:
    """create_synthetic_dataset.py:111-116: ``F.interpolate(s2_nir, scale_factor=4)`` (nearest), histogram matching of
    the prediction against it, float16.  Returns a (B,1,H,W) float16 CUDA tensor."""
    require_cuda(pred, "postprocess prediction")
    s2 = s2_nir.to(pred.device, non_blocking=True)
    up = resize(s2, (s2.shape[-2] * 4, s2.shape[-1] * 4), "nearest")
    return histogram_match(pred, up, out_dtype=torch.float16)


def save_image(pred_nir: torch.Tensor, out_path: str, name: str) -> str:
    """create_synthetic_dataset.py:49-52: ``np.savez_compressed(<out_path>/<name>, nir=pred_nir.numpy())``."""
    out_filename = os.path.join(out_path, f"{name}")
    np.savez_compressed(out_filename, nir=pred_nir.detach().cpu().numpy())
    return out_filename + ".npz"


def save_images(pred_nir: torch.Tensor, names: Sequence[str], out_path: str, workers: int = 8) -> list:
    """One ``.npz`` per tile (float16 (1,H,W) arrays keyed 'nir'), written by a thread pool: zlib releases the GIL, so
    the writer keeps up with the device instead of serialising the loop as the reference does."""
    os.makedirs(out_path, exist_ok=True)
    host = pred_nir.detach().to("cpu")
    with ThreadPoolExecutor(max_workers=max(1, workers)) as ex:
        return list(ex.map(lambda a: save_image(a[0], out_path, a[1]), zip(host, names)))
