"""CPU ORACLE for the post-processing that follows the generator -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Restates, in numpy / torch-CPU, what ``create_synthetic_dataset.py`` does between ``pred = model(hr)`` and the
``.npz`` file:

  * ``s2_nir_int = F.interpolate(s2_nir, scale_factor=4)``                      create_synthetic_dataset.py:111
  * ``histogram_match(image=pred, reference=s2_nir_int)``                       create_synthetic_dataset.py:34-47,112
      - bilinear resize of the reference to the image size (align_corners=False) :37
      - per sample ``skimage.exposure.match_histograms(img, ref, channel_axis=None)`` :44
  * ``im.to(torch.float16)`` + ``np.savez_compressed(name, nir=...)``            :49-52,116-117

Third-party dependency: ``match_histograms`` lives in **scikit-image**, which the reference imports
(create_synthetic_dataset.py:5) but neither pins in requirements.txt nor ships, and which is not installed in this
image.  ``match_cumulative_cdf`` below restates its published algorithm
(``skimage/exposure/histogram_matching.py::_match_cumulative_cdf``, unchanged from 0.18 to 0.25, float branch):

    src_values, src_unique_indices, src_counts = np.unique(source.ravel(), return_inverse=True, return_counts=True)
    tmpl_values, tmpl_counts = np.unique(template.ravel(), return_counts=True)
    src_quantiles = np.cumsum(src_counts) / source.size
    tmpl_quantiles = np.cumsum(tmpl_counts) / template.size
    interp_a_values = np.interp(src_quantiles, tmpl_quantiles, tmpl_values)
    return interp_a_values[src_unique_indices].reshape(source.shape)

followed by the cast back to the image's float type (float32).  PARITY UNPINNED for this row: no scikit-image here to
generate golden vectors from, and the reference has no test for it; the restatement is checked against its defining
properties instead (tests/test_cpu_postprocess.py).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


def match_cumulative_cdf(source: np.ndarray, template: np.ndarray) -> np.ndarray:
    """skimage.exposure.histogram_matching._match_cumulative_cdf (float images), result cast to float32."""
    src_values, src_unique_indices, src_counts = np.unique(source.ravel(), return_inverse=True, return_counts=True)
    tmpl_values, tmpl_counts = np.unique(template.ravel(), return_counts=True)
    src_quantiles = np.cumsum(src_counts) / source.size
    tmpl_quantiles = np.cumsum(tmpl_counts) / template.size
    interp_a_values = np.interp(src_quantiles, tmpl_quantiles, tmpl_values)
    return interp_a_values[src_unique_indices.ravel()].reshape(source.shape).astype(np.float32)


def histogram_match(image: torch.Tensor, reference: torch.Tensor) -> torch.Tensor:
    """create_synthetic_dataset.py:34-47: (B,1,H,W), (B,1,h,w) -> (B,1,H,W) float32."""
    reference = F.interpolate(reference, size=image.shape[-2:], mode="bilinear", align_corners=False)
    matched = []
    for img, ref in zip(image, reference):
        m = match_cumulative_cdf(img.squeeze().cpu().numpy(), ref.squeeze().cpu().numpy())
        matched.append(torch.from_numpy(m).unsqueeze(0))
    return torch.stack(matched, dim=0)


def postprocess(pred: torch.Tensor, s2_nir: torch.Tensor) -> torch.Tensor:
    """create_synthetic_dataset.py:111-116: nearest x4 upsampling of the Sentinel-2 NIR, histogram matching, float16."""
    s2_nir_int = F.interpolate(s2_nir, scale_factor=4)
    return histogram_match(pred, s2_nir_int).to(torch.float16)


# --------------------------------------------------------------------------------------
# validation metrics (utils/calculate_metrics.py:6-37) -- kornia==0.7.3 (requirements.txt:9) is not installed here;
# its published algorithms are restated: kornia.metrics.psnr = 10 log10(max_val^2 / mse), kornia.metrics.ssim =
# Gaussian window (sigma 1.5, normalised), filter2d_separable with border_type 'reflect' and 'same' padding,
# C1 = (0.01 max_val)^2, C2 = (0.03 max_val)^2, eps = 1e-12, map = num / (den + eps).  PARITY UNPINNED (no kornia).
# --------------------------------------------------------------------------------------
def gaussian_kernel1d(window_size: int, sigma: float = 1.5) -> torch.Tensor:
    x = torch.arange(window_size, dtype=torch.float32) - window_size // 2
    g = torch.exp(-(x ** 2) / (2.0 * sigma ** 2))
    return g / g.sum()


def _filter(x: torch.Tensor, k1: torch.Tensor) -> torch.Tensor:
    p = k1.numel() // 2
    C = x.shape[1]
    x = F.pad(x, (p, p, p, p), mode="reflect")
    kx = k1.view(1, 1, 1, -1).repeat(C, 1, 1, 1)
    ky = k1.view(1, 1, -1, 1).repeat(C, 1, 1, 1)
    return F.conv2d(F.conv2d(x, kx, groups=C), ky, groups=C)


def ssim_map(img1: torch.Tensor, img2: torch.Tensor, window_size: int, max_val: float = 1.0, eps: float = 1e-12):
    k = gaussian_kernel1d(window_size)
    C1, C2 = (0.01 * max_val) ** 2, (0.03 * max_val) ** 2
    mu1, mu2 = _filter(img1, k), _filter(img2, k)
    s1 = _filter(img1 ** 2, k) - mu1 ** 2
    s2 = _filter(img2 ** 2, k) - mu2 ** 2
    s12 = _filter(img1 * img2, k) - mu1 * mu2
    num = (2.0 * mu1 * mu2 + C1) * (2.0 * s12 + C2)
    den = (mu1 ** 2 + mu2 ** 2 + C1) * (s1 + s2 + C2)
    return num / (den + eps)


def calculate_metrics(pred: torch.Tensor, target: torch.Tensor, phase: str = "train") -> dict:
    """utils/calculate_metrics.py:6-37."""
    mse = F.mse_loss(pred, target)
    return {phase + "/L1": F.l1_loss(pred, target).item(), phase + "/L2": mse.item(),
            phase + "/PSNR": (10.0 * torch.log10(1.0 / mse)).item(),
            phase + "/SSIM": ssim_map(pred, target, 5, 1.0).mean().item()}


def ssim_loss(img1: torch.Tensor, img2: torch.Tensor, window_size: int = 11) -> torch.Tensor:
    """utils/losses.py:10-29: 1 - kornia.metrics.ssim(img1, img2, window_size).mean() (differentiable, torch)."""
    return 1.0 - ssim_map(img1.float(), img2.float(), window_size).mean()


def emd_loss(pred: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """utils/losses.py:64-78 (pix2pix.py's hist_loss): mean |cumsum(softmax) difference| over (B, C*H*W); pinned
    bit-exact against the reference function by oracle/pin_losses.py."""
    pred = pred.reshape(pred.shape[0], -1)
    target = target.reshape(target.shape[0], -1)
    pred_cdf = torch.cumsum(F.softmax(pred, dim=1), dim=1)
    target_cdf = torch.cumsum(F.softmax(target, dim=1), dim=1)
    return torch.mean(torch.abs(pred_cdf - target_cdf))
