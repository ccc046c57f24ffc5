"""Times ng_satclip_encode (L=10, hidden 256, 2 layers, 256 out -- the shipped satclip-resnet50-l10 shape) against the
oracle's torch-float64 CPU evaluation.  Prints one JSON line."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import nirgan_b200  # noqa: F401,E402
import satclip_oracle as S  # noqa: E402
from nirgan_b200.model.satclip.satclip_wrapper import SatClIP_wrapper  # noqa: E402


def main():
    sd = S.random_siren_state_dict(100, 256, 256, 2, seed=0)
    m = SatClIP_wrapper(state_dict=sd, legendre_polys=10)
    out = {}
    for B in (24, 64, 1024, 16384):
        g = torch.Generator().manual_seed(B)
        c = torch.stack([torch.rand(B, generator=g) * 360 - 180, torch.rand(B, generator=g) * 180 - 90], -1).cuda()
        for _ in range(5):
            m.predict(c)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(20):
            m.predict(c)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        cc = c.cpu()
        t0 = time.perf_counter()
        S.location_encoder(sd, cc, 10)
        cpu_ms = (time.perf_counter() - t0) * 1e3
        out[str(B)] = {"gpu_ms": round(ms, 4), "coords_per_s": round(B / ms * 1e3), "oracle_cpu_ms": round(cpu_ms, 2)}
    print(json.dumps({"satclip_encode": out}))


if __name__ == "__main__":
    main()
