"""GPU parity: ng_satclip_encode (through the SatClIP_wrapper mirror) against the oracle and the reference-generated
fixtures.  Tolerance: the kernel and the reference both compute in float64 but sum the dot products in a different
order, and sin(30 * .) amplifies a 1e-16 relative difference by up to 30x per layer; 1e-9 absolute on the float64 values
is generous, and the float32 result must agree to 1 float32 ulp-ish (2e-6 relative to max|y|)."""
import math
import os

import numpy as np
import pytest
import torch

import satclip_oracle as S

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _wrapper(sd, L):
    from nirgan_b200.model.satclip.satclip_wrapper import SatClIP_wrapper as W
    return W(state_dict=sd, legendre_polys=L, device="cuda")


def _check(y, ref):
    ref = ref.to(y.device)
    scale = float(ref.abs().max())
    assert y.dtype == torch.float32 and y.shape == ref.shape
    assert float((y - ref).abs().max()) <= 2e-6 * max(scale, 1.0)


def test_fixture_small_from_the_reference_classes():
    z = np.load(os.path.join(GOLD, "satclip_small.npz"))
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd.")}
    L = int(round(math.sqrt(z["pe"].shape[1])))
    m = _wrapper(sd, L)
    _check(m.predict(torch.from_numpy(z["lonlat"])), torch.from_numpy(z["y"]))


def test_fixture_l10_checkpoint_shape():
    z = np.load(os.path.join(GOLD, "satclip_l10.npz"))
    sd = S.random_siren_state_dict(100, 256, 256, 2, seed=int(z["seed"]))
    m = _wrapper(sd, 10)
    _check(m.predict(torch.from_numpy(z["lonlat"]).cuda()), torch.from_numpy(z["y"]))


@pytest.mark.parametrize("L,hidden,layers,dout,B", [(1, 8, 1, 3, 5), (4, 32, 1, 16, 7), (16, 128, 3, 64, 33),
                                                    (32, 512, 2, 256, 9), (10, 1024, 2, 300, 4), (40, 512, 2, 256, 64)])
def test_against_the_oracle_over_shapes(L, hidden, layers, dout, B):
    if L * L > 1024:
        sd = S.random_siren_state_dict(L * L, hidden, dout, layers, seed=1)
        with pytest.raises(RuntimeError):
            _wrapper(sd, L).predict(torch.zeros(B, 2))
        return
    g = torch.Generator().manual_seed(L * 100 + B)
    lon = torch.rand(B, generator=g, dtype=torch.float64) * 360 - 180
    lat = torch.rand(B, generator=g, dtype=torch.float64) * 180 - 90
    lonlat = torch.stack([lon, lat], -1)
    sd = S.random_siren_state_dict(L * L, hidden, dout, layers, seed=L + B)
    _check(_wrapper(sd, L).predict(lonlat), S.location_encoder(sd, lonlat, L))


def test_edge_coordinates_and_float32_input():
    """Poles, the date line, the reference's own test ranges (lon in [-90, 90], lat in [-180, 180], satclip_wrapper.py:50-52)
    and float32 coordinates (predict() casts to double first, like the reference)."""
    pts = torch.tensor([[0.0, 90.0], [0.0, -90.0], [180.0, 0.0], [-180.0, 0.0], [45.0, 170.0], [-60.0, -135.0],
                        [1e-9, 1e-9]], dtype=torch.float32)
    sd = S.random_siren_state_dict(100, 256, 256, 2, seed=11)
    m = _wrapper(sd, 10)
    _check(m.predict(pts.cuda()), S.location_encoder(sd, pts, 10))
    assert m.predict(torch.zeros(0, 2)).shape == (0, 256)
    with pytest.raises(ValueError):
        m.predict(torch.zeros(3, 3))


def test_feeds_the_injected_generator_like_predict_step():
    """Px2Px_PL.predict_step (model/pix2pix.py:134-163): coords -> embeds -> netG(rgb, embeds)."""
    sd = S.random_siren_state_dict(100, 256, 256, 2, seed=5)
    m = _wrapper(sd, 10)
    coords = torch.tensor([[11.5, 48.1], [-3.7, 40.4]])
    emb = m.predict(coords)
    assert emb.shape == (2, 256) and emb.is_cuda and torch.isfinite(emb).all()
    # state_dict round trip in the checkpoint's key layout
    from nirgan_b200.model.satclip.satclip_wrapper import SatClIP_wrapper as W
    ck = {"model.location.nnet." + k[5:]: v for k, v in sd.items()}
    ck["model.visual.conv1.weight"] = torch.zeros(1)
    m2 = W(state_dict=ck, legendre_polys=10)
    assert torch.equal(m2.predict(coords), emb)


def test_px2px_predict_step_takes_coordinates():
    """predict_step(rgb, coords) (model/pix2pix.py:134-163, extract_batch :448-459): the attached encoder produces the
    embeddings that are injected; identical to passing the embeddings precomputed; an inject model with neither raises."""
    from nirgan_b200.model.pix2pix import Px2Px
    from nirgan_b200.config import satclip_inject_config as inject_config
    torch.manual_seed(0)
    model = Px2Px(inject_config()).cuda().eval()
    sd = S.random_siren_state_dict(100, 256, 256, 2, seed=2)
    enc = _wrapper(sd, 10).eval()
    rgb = torch.rand(2, 3, 64, 64, device="cuda")
    coords = torch.tensor([[11.5, 48.1], [-70.6, -33.4]], device="cuda")
    with pytest.raises(RuntimeError):
        model.predict_step(rgb, coords)
    model.attach_satclip(enc)
    y = model.predict_step(rgb, coords)
    y2 = model.predict_step(rgb, embeds=S.location_encoder(sd, coords.cpu(), 10).cuda())
    assert y.shape == (2, 1, 64, 64) and float((y - y2).abs().max()) <= 1e-3
    rgb_b, _, emb = model.extract_batch({"rgb": rgb, "nir": None, "coords": coords})
    assert emb.shape == (2, 256) and rgb_b is rgb
