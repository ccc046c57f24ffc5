"""B200 drop-in for the reference's ``utils/remote_sensing_indices.py``.

``RemoteSensingIndices(mode, criterion)`` keeps the reference's surface (remote_sensing_indices.py:4-319):
``get_and_weight_losses(rgb, nir, nir_pred, loss_config=None, mode="loss" | "logging_dict")`` and the six
``*_calculation`` methods in both object modes ('loss' -> scalar, 'index' -> (index, index_pred) maps), criterion
'l1' or 'l2'.  Every evaluation is ONE fused CUDA kernel (``ng_rs_pixel_losses`` / ``ng_rs_index``) that reads the five
input planes once, evaluates only the requested indices and, in loss mode, also emits d/dpred.  Formulas follow the
reference verbatim (loss-mode epsilons, the product-form EVI denominator, GNDVI without epsilon).
"""
from __future__ import annotations

from ..losses import RS_TERMS, rs_index, rs_pixel_losses

_KEYS = ("lambda_ndvi", "lambda_ndwi", "lambda_gndvi", "lambda_savi", "lambda_msavi", "lambda_evi")   # iteration order
_LOG_NAMES = {"lambda_ndvi": "indices_loss/ndvi_error", "lambda_ndwi": "indices_loss/ndwi_error",
              "lambda_gndvi": "indices_loss/gndvi_error", "lambda_savi": "indices_loss/savi_error",
              "lambda_msavi": "indices_loss/msavi_error", "lambda_evi": "indices_loss/evi_error"}


class RemoteSensingIndices():
    def __init__(self, mode="loss", criterion="l1"):
        assert mode in ["loss", "index"], f"Mode '{mode}' not implemented. 'loss', 'index' are supported."
        self.mode = mode
        if criterion not in ("l1", "l2"):
            raise NotImplementedError(f"Criterion '{criterion}' not implemented. 'l1' or 'l2' are supported.")
        self.criterion_name = criterion

    @staticmethod
    def _prep(t):
        return t.unsqueeze(0) if t.dim() == 3 else t

    def prepare_tensor_for_loss(self, rgb, nir, nir_pred):
        return self._prep(rgb), self._prep(nir), self._prep(nir_pred)

    def _terms(self, rgb, nir, nir_pred, weights6):
        """weights6 in _KEYS order -> the 7-term vector of the fused kernel (term 0, the pix2pix L1, is off)."""
        rgb, nir, nir_pred = self.prepare_tensor_for_loss(rgb, nir, nir_pred)
        return rs_pixel_losses(rgb, nir, nir_pred, (0.0,) + tuple(weights6), self.criterion_name)

    def get_and_weight_losses(self, rgb, nir, nir_pred, loss_config=None, mode="loss"):
        if loss_config is None:
            loss_config = {"lambda_ndvi": 0.333, "lambda_ndwi": 0.333, "lambda_evi": 0.333,
                           "lambda_savi": 0.0, "lambda_msavi": 0.0, "lambda_gndvi": 0.0}
        if mode == "loss":
            w = [float(loss_config.get(k, 0.0)) if loss_config.get(k, 0.0) > 0.0 else 0.0 for k in _KEYS]
            if not any(w):
                return 0.0
            out = self._terms(rgb, nir, nir_pred, w)
            total = 0.0                 # same accumulation order as the reference (remote_sensing_indices.py:55-60)
            for i, wi in enumerate(w):
                if wi > 0.0:
                    total = total + wi * out[i + 1]
            return total
        elif mode == "logging_dict":
            out = self._terms(rgb, nir, nir_pred, [1.0] * 6)
            return {_LOG_NAMES[k]: out[i + 1] for i, k in enumerate(_KEYS)}
        raise NotImplementedError(f"Mode '{mode}' not implemented. 'loss' or 'logging_dict' are supported.")

    def _single(self, which, rgb, nir, nir_pred):
        rgb, nir, nir_pred = self.prepare_tensor_for_loss(rgb, nir, nir_pred)
        if self.mode == "loss":
            w = [0.0] * 7
            w[RS_TERMS.index(which)] = 1.0
            return rs_pixel_losses(rgb, nir, nir_pred, w, self.criterion_name)[RS_TERMS.index(which)]
        elif self.mode == "index":
            return rs_index(rgb, nir, nir_pred, which, loss_eps=False)
        raise NotImplementedError(f"Mode '{self.mode}' not implemented. 'loss' or 'index' are supported.")

    def ndvi_calculation(self, rgb, nir, nir_pred):
        return self._single("ndvi", rgb, nir, nir_pred)

    def ndwi_calculation(self, rgb, nir, nir_pred):
        return self._single("ndwi", rgb, nir, nir_pred)

    def gndvi_calculation(self, rgb, nir, nir_pred):
        return self._single("gndvi", rgb, nir, nir_pred)

    def savi_calculation(self, rgb, nir, nir_pred):
        return self._single("savi", rgb, nir, nir_pred)

    def msavi_calculation(self, rgb, nir, nir_pred):
        return self._single("msavi", rgb, nir, nir_pred)

    def evi_calculation(self, rgb, nir, nir_pred):
        return self._single("evi", rgb, nir, nir_pred)
