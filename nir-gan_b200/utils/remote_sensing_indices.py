"""B200 drop-in for the reference's ``utils/remote_sensing_indices.py``.

``RemoteSensingIndices(mode, criterion).get_and_weight_losses(rgb, nir, nir_pred, loss_config, mode)``
keeps its signature (remote_sensing_indices.py:6,23).  The hot configuration -- mode 'loss',
criterion 'l1', NDVI / NDWI / EVI weights > 0 and GNDVI / SAVI / MSAVI weights 0 (every shipped
config, configs/config_px2px_SatCLIP.yaml:32-38) -- runs as ONE fused CUDA kernel that reads the
five input planes once and also emits d/dpred.  Anything else (criterion 'l2', 'index' mode,
'logging_dict', the three zero-weight indices) is outside the hot path and raises.
"""
from __future__ import annotations

from ..losses import pixel_losses

_HOT = ("lambda_ndvi", "lambda_ndwi", "lambda_evi")
_COLD = ("lambda_gndvi", "lambda_savi", "lambda_msavi")


class RemoteSensingIndices():
    def __init__(self, mode="loss", criterion="l1"):
        assert mode in ["loss", "index"], f"Mode '{mode}' not implemented. 'loss', 'index' are supported."
        self.mode = mode
        if criterion not in ("l1", "l2"):
            raise NotImplementedError(f"Criterion '{criterion}' not implemented. 'l1' or 'l2' are supported.")
        if criterion != "l1" or mode != "loss":
            raise NotImplementedError("nirgan_b200 RemoteSensingIndices: only mode='loss', criterion='l1' is on the "
                                      "accelerated hot path")
        self.criterion_name = criterion

    @staticmethod
    def _prep(t):
        return t.unsqueeze(0) if t.dim() == 3 else t

    def _fused(self, rgb, nir, nir_pred, w_ndvi, w_ndwi, w_evi):
        rgb, nir, nir_pred = self._prep(rgb), self._prep(nir), self._prep(nir_pred)
        return pixel_losses(rgb, nir, nir_pred, (0.0, w_ndvi, w_ndwi, w_evi))

    def get_and_weight_losses(self, rgb, nir, nir_pred, loss_config=None, mode="loss"):
        if loss_config is None:
            loss_config = {"lambda_ndvi": 0.333, "lambda_ndwi": 0.333, "lambda_evi": 0.333,
                           "lambda_savi": 0.0, "lambda_msavi": 0.0, "lambda_gndvi": 0.0}
        if mode != "loss":
            raise NotImplementedError(f"Mode '{mode}' is outside the nirgan_b200 hot path ('loss' only).")
        for k in _COLD:
            if loss_config.get(k, 0.0) > 0.0:
                raise NotImplementedError(f"nirgan_b200: index weight {k} > 0 is outside the accelerated hot path "
                                          f"(weight 0.0 in every shipped config)")
        w = [max(float(loss_config.get(k, 0.0)), 0.0) for k in _HOT]
        if not any(w):
            return 0.0
        out = self._fused(rgb, nir, nir_pred, *w)
        # same accumulation order as the reference: ndvi, ndwi, ..., evi (remote_sensing_indices.py:45-62)
        total = 0.0
        for wi, idx in zip(w, (1, 2, 3)):
            if wi > 0.0:
                total = total + wi * out[idx]
        return total

    def ndvi_calculation(self, rgb, nir, nir_pred):
        return self._fused(rgb, nir, nir_pred, 1.0, 0.0, 0.0)[1]

    def ndwi_calculation(self, rgb, nir, nir_pred):
        return self._fused(rgb, nir, nir_pred, 0.0, 1.0, 0.0)[2]

    def evi_calculation(self, rgb, nir, nir_pred):
        return self._fused(rgb, nir, nir_pred, 0.0, 0.0, 1.0)[3]
