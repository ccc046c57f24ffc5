"""Pin oracle/satclip_oracle.py against the REAL reference classes and write the golden fixture.

TEST INFRASTRUCTURE; runs only in the authoring container (needs /root/reference).  The reference's
``model/satclip/__init__.py`` imports pytorch_lightning (absent) and ``positional_encoding/spherical_harmonics.py``
imports the un-shipped analytic module, so the package init is bypassed and the analytic module is stubbed with the
reference's own closed-form implementation; ``get_positional_encoding(..., harmonics_calculation='closed-form')``,
``get_neural_network('siren', ...)`` and ``LocationEncoder`` are then the reference's unmodified code.
Usage: python oracle/pin_satclip.py
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
REF = os.environ.get("NIRGAN_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
import satclip_oracle as S  # noqa: E402


def reference_classes():
    import model  # noqa: F401  (namespace package of the reference)
    pkg = types.ModuleType("model.satclip")
    pkg.__path__ = [os.path.join(REF, "model", "satclip")]
    sys.modules["model.satclip"] = pkg
    spec = importlib.util.spec_from_file_location(
        "cf", os.path.join(REF, "model", "satclip", "positional_encoding", "spherical_harmonics_closed_form.py"))
    cf = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(cf)
    stub = types.ModuleType("model.satclip.positional_encoding.spherical_harmonics_ylm")
    stub.SH = cf.SH
    sys.modules["model.satclip.positional_encoding.spherical_harmonics_ylm"] = stub
    from model.satclip.location_encoder import LocationEncoder, get_neural_network, get_positional_encoding
    return LocationEncoder, get_neural_network, get_positional_encoding


def main():
    LocationEncoder, get_nn, get_pe = reference_classes()
    lines = ["satclip oracle vs /root/reference classes (float64, closed-form harmonics), max abs diff"]
    g = torch.Generator().manual_seed(3)
    lonlat = torch.stack([torch.rand(64, generator=g) * 360 - 180, torch.rand(64, generator=g) * 180 - 90], -1).double()
    lonlat[0] = torch.tensor([0.0, 0.0])
    lonlat[1] = torch.tensor([-180.0, -90.0])
    lonlat[2] = torch.tensor([180.0, 90.0])
    ok = True
    for L, hidden, layers in ((10, 256, 2), (6, 64, 2), (4, 32, 1), (16, 128, 3)):
        torch.manual_seed(L)
        pe = get_pe("sphericalharmonics", legendre_polys=L, harmonics_calculation="closed-form").double()
        net = get_nn("siren", input_dim=pe.embedding_dim, num_classes=256, dim_hidden=hidden, num_layers=layers).double()
        enc = LocationEncoder(pe, net).eval()
        with torch.no_grad():
            ref_pe = pe(lonlat)
            ref = enc(lonlat).float()
        sd = {k: v.detach().clone() for k, v in enc.state_dict().items()}
        d_pe = float((S.spherical_harmonics(lonlat, L) - ref_pe).abs().max())
        d = float((S.location_encoder(sd, lonlat, L) - ref).abs().max())
        lines.append(f"L={L:2d} hidden={hidden:3d} layers={layers}  harmonics {d_pe:.3e}  embeddings {d:.3e}  "
                     f"{'ok' if d_pe <= 1e-12 and d <= 1e-6 else 'FAIL'}")
        ok &= d_pe <= 1e-12 and d <= 1e-6
        if (L, hidden) == (6, 64):
            np.savez_compressed(os.path.join(HERE, "..", "tests", "golden", "satclip_small.npz"), L=L, lonlat=lonlat.numpy(),
                                y=ref.numpy(), pe=ref_pe.numpy(), **{"sd." + k: v.numpy() for k, v in sd.items()})
        if (L, hidden) == (10, 256):
            # full-size configuration (satclip-resnet50-l10: L = 10, capacity 256, 2 hidden layers): the weights are too
            # large for a fixture, so only a fingerprint of the reference's output for the oracle's seeded weights is kept
            sd2 = S.random_siren_state_dict(100, 256, 256, 2, seed=7)
            enc.load_state_dict(sd2)
            with torch.no_grad():
                ref2 = enc(lonlat).float()
            d2 = float((S.location_encoder(sd2, lonlat, 10) - ref2).abs().max())
            lines.append(f"L=10 full size, oracle-seeded weights (seed 7): embeddings {d2:.3e}  {'ok' if d2 <= 1e-6 else 'FAIL'}")
            ok &= d2 <= 1e-6
            np.savez_compressed(os.path.join(HERE, "..", "tests", "golden", "satclip_l10.npz"), lonlat=lonlat.numpy(),
                                y=ref2.numpy(), seed=7)
    rep = "\n".join(lines)
    print(rep)
    open(os.path.join(HERE, "..", "tests", "golden", "PIN_REPORT_satclip.txt"), "w").write(rep + "\n")
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
