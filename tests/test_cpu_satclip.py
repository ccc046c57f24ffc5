"""CPU suite for the SatCLIP location-encoder oracle (oracle/satclip_oracle.py): the committed fixtures were generated
by oracle/pin_satclip.py from the REFERENCE's own LocationEncoder classes (tests/golden/PIN_REPORT_satclip.txt), so
this is the oracle-vs-reference pin; plus textbook values of the real spherical harmonics."""
import math
import os

import numpy as np
import torch

import satclip_oracle as S

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _load(name):
    z = np.load(os.path.join(GOLD, name))
    return {k: z[k] for k in z.files}


def test_oracle_reproduces_the_reference_fixture_small():
    z = _load("satclip_small.npz")
    sd = {k[3:]: torch.from_numpy(v) for k, v in z.items() if k.startswith("sd.")}
    lonlat = torch.from_numpy(z["lonlat"])
    L = int(round(math.sqrt(z["pe"].shape[1])))
    pe = S.spherical_harmonics(lonlat, L)
    assert torch.equal(pe, torch.from_numpy(z["pe"]))                  # same float64 operations -> bit-exact
    y = S.location_encoder(sd, lonlat, L)
    assert y.dtype == torch.float32 and torch.equal(y, torch.from_numpy(z["y"]))


def test_oracle_reproduces_the_reference_fixture_l10():
    z = _load("satclip_l10.npz")
    sd = S.random_siren_state_dict(100, 256, 256, 2, seed=int(z["seed"]))
    y = S.location_encoder(sd, torch.from_numpy(z["lonlat"]), 10)
    assert torch.equal(y, torch.from_numpy(z["y"]))


def test_low_order_harmonics_have_their_textbook_values():
    lonlat = torch.tensor([[-180.0, 0.0], [-90.0, -60.0], [12.5, 33.0], [179.0, 89.0]], dtype=torch.float64)
    pe = S.spherical_harmonics(lonlat, 3)
    phi = torch.deg2rad(lonlat[:, 0] + 180)
    th = torch.deg2rad(lonlat[:, 1] + 90)
    x, y, zc = torch.sin(th) * torch.cos(phi), torch.sin(th) * torch.sin(phi), torch.cos(th)
    c0 = 0.5 * math.sqrt(1 / math.pi)
    c1 = math.sqrt(3 / (4 * math.pi))
    assert torch.allclose(pe[:, 0], torch.full((4,), c0, dtype=torch.float64), atol=1e-15)
    # Condon-Shortley phase is kept by the reference's recursion: Y_1^{-1} = -c1 y, Y_1^0 = c1 z, Y_1^1 = -c1 x
    assert torch.allclose(pe[:, 1], -c1 * y, atol=1e-14)
    assert torch.allclose(pe[:, 2], c1 * zc, atol=1e-14)
    assert torch.allclose(pe[:, 3], -c1 * x, atol=1e-14)
    assert torch.allclose(pe[:, 6], 0.25 * math.sqrt(5 / math.pi) * (3 * zc * zc - 1), atol=1e-14)


def test_harmonics_are_orthonormal_on_the_sphere():
    """Gauss-Legendre x uniform-phi quadrature of Y_i Y_j over the sphere is the identity for l < 6."""
    L = 6
    xs, ws = np.polynomial.legendre.leggauss(16)
    nphi = 32
    lat = np.degrees(np.arccos(xs)) - 90.0
    lon = np.arange(nphi) * 360.0 / nphi - 180.0
    grid = np.stack(np.meshgrid(lon, lat, indexing="ij"), -1).reshape(-1, 2)
    w = np.tile(ws, nphi) * (2 * math.pi / nphi)
    pe = S.spherical_harmonics(torch.from_numpy(grid), L).numpy()
    gram = (pe * w[:, None]).T @ pe
    assert np.allclose(gram, np.eye(L * L), atol=1e-12)


def test_siren_forward_matches_a_hand_rolled_loop():
    sd = S.random_siren_state_dict(9, 5, 4, 2, seed=3)
    x = torch.linspace(-1, 1, 18, dtype=torch.float64).reshape(2, 9)
    h = torch.sin(30.0 * (x @ sd["nnet.layers.0.weight"].T + sd["nnet.layers.0.bias"]))
    h = torch.sin(1.0 * (h @ sd["nnet.layers.1.weight"].T + sd["nnet.layers.1.bias"]))
    y = h @ sd["nnet.last_layer.weight"].T + sd["nnet.last_layer.bias"]
    assert torch.allclose(S.siren_forward(sd, x), y, atol=1e-15)
