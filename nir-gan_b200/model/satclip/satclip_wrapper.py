"""B200 drop-in for ``model/satclip/satclip_wrapper.py::SatClIP_wrapper`` (the location encoder in front of the injected
generator).

``SatClIP_wrapper(satclip_path, device).predict(coords)`` keeps the reference's contract (satclip_wrapper.py:29-34):
``coords`` (B, 2) lon/lat in degrees -> float32 (B, embed_dim) embeddings, computed in float64.  The whole encoder --
spherical-harmonics positional encoding (positional_encoding/spherical_harmonics.py:27-42, closed form
spherical_harmonics_closed_form.py:8-40) and the SIREN MLP (location_encoder.py:73-151) -- is ONE kernel,
``ng_satclip_encode``.  Only the location tower of the checkpoint is read (``model.location.nnet.*``, the keys
``get_satclip`` keeps, load.py:3-18); the image tower is never used by NIR-GAN.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn as nn

from ... import _lib as L
from ...engine import require_cuda

_PREFIXES = ("model.location.nnet.", "location.nnet.", "nnet.", "")


def _location_tower(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    for p in _PREFIXES:
        if p + "last_layer.weight" in sd:
            return {k[len(p):]: v for k, v in sd.items() if k.startswith(p) and (k[len(p):].startswith("layers.")
                                                                               or k[len(p):].startswith("last_layer."))}
    raise KeyError("no SirenNet weights ('...nnet.last_layer.weight') in the state dict")


class SatClIP_wrapper(nn.Module):
    def __init__(self, satclip_path: Optional[str] = None, device="cuda", state_dict=None, legendre_polys: int = 10,
                 w0: float = 1.0, w0_initial: float = 30.0):
        super().__init__()
        if state_dict is None:
            if satclip_path is None:
                satclip_path = "model/satclip/satclip-resnet50-l10.ckpt"     # the reference's default
            ckpt = torch.load(satclip_path, map_location="cpu", weights_only=False)
            legendre_polys = int(ckpt.get("hyper_parameters", {}).get("legendre_polys", legendre_polys))
            state_dict = ckpt["state_dict"]
        sd = _location_tower(state_dict)
        n = 0
        while f"layers.{n}.weight" in sd:
            n += 1
        if n < 1:
            raise ValueError("SirenNet needs at least one hidden layer")
        self.legendre_polys = int(legendre_polys)
        self.num_layers = n
        self.dim_hidden = int(sd["layers.0.weight"].shape[0])
        self.dim_out = int(sd["last_layer.weight"].shape[0])
        if int(sd["layers.0.weight"].shape[1]) != self.legendre_polys ** 2:
            raise ValueError(f"first layer expects {int(sd['layers.0.weight'].shape[1])} inputs, "
                             f"legendre_polys={self.legendre_polys} gives {self.legendre_polys ** 2}")
        self.w0, self.w0_initial = float(w0), float(w0_initial)
        # reference layout (float64, [out][in]) kept as buffers so state_dict() round-trips; the kernel reads a transposed copy
        for k, v in sd.items():
            self.register_buffer(k.replace(".", "_"), v.detach().to(torch.float64).clone())
        self._keys = list(sd.keys())
        self._packed = None
        self.to(device)

    def _apply(self, fn, *a, **kw):
        self._packed = None
        return super()._apply(fn, *a, **kw)

    def _params_t(self) -> torch.Tensor:
        if self._packed is None:
            parts = []
            for i in range(self.num_layers):
                parts += [getattr(self, f"layers_{i}_weight").t().contiguous().reshape(-1), getattr(self, f"layers_{i}_bias")]
            parts += [self.last_layer_weight.t().contiguous().reshape(-1), self.last_layer_bias]
            self._packed = torch.cat(parts).contiguous()
        return self._packed

    @torch.no_grad()
    def predict(self, x: torch.Tensor) -> torch.Tensor:
        if x.dim() != 2 or x.shape[1] != 2:
            raise ValueError(f"coords must be (B, 2) lon/lat, got {tuple(x.shape)}")
        p = self._params_t()
        require_cuda(p, "the SatCLIP encoder (move it with .to(\"cuda\"))")
        x = x.to(device=p.device, dtype=torch.float64).contiguous()
        out = torch.empty(x.shape[0], self.dim_out, dtype=torch.float32, device=p.device)
        if x.shape[0] == 0:
            return out
        L.call("ng_satclip_encode", x.data_ptr(), x.shape[0], self.legendre_polys, p.data_ptr(), self.dim_hidden,
               self.num_layers, self.dim_out, self.w0_initial, self.w0, out.data_ptr(),
               torch.cuda.current_stream(p.device).cuda_stream)
        return out

    def forward(self, x):
        print("Don't use fwd, use 'predict' step instead")
        return self.predict(x).double()
