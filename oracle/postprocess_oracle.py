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
