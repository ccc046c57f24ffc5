"""CPU ORACLE for the SatCLIP location encoder -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Restates in torch float64 the step right before the injected generator: ``SatClIP_wrapper.predict(coords)``
(model/satclip/satclip_wrapper.py:29-34) = ``LocationEncoder(posenc, nnet)`` (model/satclip/location_encoder.py:267-275)
with

  * posenc = ``SphericalHarmonics(legendre_polys=L)`` (positional_encoding/spherical_harmonics.py:27-42): lon/lat in
    degrees -> phi = deg2rad(lon + 180), theta = deg2rad(lat + 90) -> the L*L real spherical harmonics Y_l^m, l < L,
    m = -l..l, from the closed-form path (positional_encoding/spherical_harmonics_closed_form.py:8-40; the analytic
    ``spherical_harmonics_ylm.py`` the shipped checkpoint was trained with is not in the repository --
    .MISSING_LARGE_BLOBS:1 -- and equals the closed form mathematically);
  * nnet = ``SirenNet(dim_in=L*L, dim_hidden, dim_out=256, num_layers)`` (location_encoder.py:73-151): hidden layers
    ``sin(w0 * (W x + b))`` with w0 = 30 for the first and 1 for the others (dropout is inactive in eval mode), last layer
    linear with identity activation; everything in float64, result cast to float32.

Pinned by ``oracle/pin_satclip.py`` against the reference's own classes (imported with the missing analytic module
stubbed by the closed form) -> ``tests/golden/satclip_small.npz`` and ``tests/golden/PIN_REPORT_satclip.txt``.
"""
from __future__ import annotations

import math
from typing import Dict

import torch

Tensor = torch.Tensor


def _assoc_legendre(l: int, m: int, x: Tensor) -> Tensor:
    """spherical_harmonics_closed_form.py:8-27 (m >= 0)."""
    pmm = torch.ones_like(x)
    if m > 0:
        somx2 = torch.sqrt((1 - x) * (1 + x))
        fact = 1.0
        for _ in range(1, m + 1):
            pmm = pmm * (-fact) * somx2
            fact += 2.0
    if l == m:
        return pmm
    pmmp1 = x * (2.0 * m + 1.0) * pmm
    if l == m + 1:
        return pmmp1
    pll = torch.zeros_like(x)
    for ll in range(m + 2, l + 1):
        pll = ((2.0 * ll - 1.0) * x * pmmp1 - (ll + m - 1.0) * pmm) / (ll - m)
        pmm, pmmp1 = pmmp1, pll
    return pll


def _renorm(l: int, m: int) -> float:
    """spherical_harmonics_closed_form.py:29-31."""
    return math.sqrt((2.0 * l + 1.0) * math.factorial(l - m) / (4 * math.pi * math.factorial(l + m)))


def spherical_harmonics(lonlat: Tensor, L: int) -> Tensor:
    """SphericalHarmonics.forward (spherical_harmonics.py:27-42): (B,2) degrees -> (B, L*L) float64."""
    lonlat = lonlat.double()
    phi = torch.deg2rad(lonlat[:, 0] + 180)
    theta = torch.deg2rad(lonlat[:, 1] + 90)
    ct = torch.cos(theta)
    out = []
    for l in range(L):
        for m in range(-l, l + 1):
            if m == 0:
                y = _renorm(l, 0) * _assoc_legendre(l, 0, ct)
            elif m > 0:
                y = math.sqrt(2.0) * _renorm(l, m) * torch.cos(m * phi) * _assoc_legendre(l, m, ct)
            else:
                y = math.sqrt(2.0) * _renorm(l, -m) * torch.sin(-m * phi) * _assoc_legendre(l, -m, ct)
            out.append(y)
    return torch.stack(out, dim=-1)


def siren_forward(sd: Dict[str, Tensor], x: Tensor, w0: float = 1.0, w0_initial: float = 30.0, prefix: str = "nnet.") -> Tensor:
    """SirenNet.forward in eval mode (location_encoder.py:104-151): state-dict keys ``nnet.layers.{i}.weight/bias``,
    ``nnet.last_layer.weight/bias``."""
    i = 0
    while f"{prefix}layers.{i}.weight" in sd:
        x = torch.sin((w0_initial if i == 0 else w0) * torch.nn.functional.linear(
            x, sd[f"{prefix}layers.{i}.weight"].double(), sd[f"{prefix}layers.{i}.bias"].double()))
        i += 1
    return torch.nn.functional.linear(x, sd[f"{prefix}last_layer.weight"].double(), sd[f"{prefix}last_layer.bias"].double())


def location_encoder(sd: Dict[str, Tensor], lonlat: Tensor, L: int) -> Tensor:
    """SatClIP_wrapper.predict (satclip_wrapper.py:29-34): float64 encoder, float32 embeddings (B, dim_out)."""
    return siren_forward(sd, spherical_harmonics(lonlat, L)).float()


def random_siren_state_dict(dim_in: int, dim_hidden: int = 256, dim_out: int = 256, num_layers: int = 2, seed: int = 0,
                            w0: float = 1.0, c: float = 6.0) -> Dict[str, Tensor]:
    """Siren.init_ (location_encoder.py:133-141) from a seeded generator: uniform(-1/dim, 1/dim) for the first layer,
    uniform(+-sqrt(c/dim)/w0) for the others (weights and biases), float64."""
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def uni(shape, std):
        return (torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) * std

    for i in range(num_layers):
        d = dim_in if i == 0 else dim_hidden
        std = (1 / d) if i == 0 else (math.sqrt(c / d) / w0)
        sd[f"nnet.layers.{i}.weight"] = uni((dim_hidden, d), std)
        sd[f"nnet.layers.{i}.bias"] = uni((dim_hidden,), std)
    std = math.sqrt(c / dim_hidden) / w0
    sd["nnet.last_layer.weight"] = uni((dim_out, dim_hidden), std)
    sd["nnet.last_layer.bias"] = uni((dim_out,), std)
    return sd
