"""Attribute-style configurations equivalent to the reference's YAML files.

The reference reads ``configs/config_px2px.yaml`` / ``configs/config_px2px_SatCLIP.yaml`` through OmegaConf and only ever
uses attribute access (``config.base_configs.ngf``, ``config.satclip.use_satclip``, ``config.Data.padding_amount`` ...;
model/pix2pix.py:18-86, model/generator_inject.py:30-52,145-200).  OmegaConf is not needed for that: these helpers
return nested ``SimpleNamespace`` objects with the same keys and the shipped values, for ``bench.py``, ``smoke()`` and
the tests (any object with the same attributes works, a real OmegaConf included).
"""
from __future__ import annotations

import types


def namespace(d: dict):
    """dict -> nested SimpleNamespace (dict leaves that must stay dicts are wrapped explicitly by the callers)."""
    return types.SimpleNamespace(**{k: namespace(v) if isinstance(v, dict) else v for k, v in d.items()})


def satclip_inject_config(scale_init: float = 0.01, post_correction: bool = False, post_correction_init: float = 1.0):
    """configs/config_px2px_SatCLIP.yaml: SatCLIP 'inject' / 'multiply' generator, scaling_param_init 0.01, pad 10."""
    return namespace({
        "base_configs": {"input_nc": 3, "output_nc": 1, "ngf": 64, "ndf": 64, "netD": "basic",
                         "netG": "resnet_9blocks", "norm": "instance", "no_dropout": True,
                         "init_type": "normal", "init_gain": 0.02, "n_layers_D": 3, "gan_mode": "lsgan",
                         "lr": 2e-4, "beta1": 0.5, "lambda_GAN": 1.0, "lambda_L1": 100.0,
                         "lambda_ssim": 0.0, "lambda_hist": 0.0, "lambda_rs_losses": 1.0,
                         "rs_losses_criterium": "l1", "isTrain": True,
                         "internal_rs_loss_weights": {"lambda_ndvi": 0.33, "lambda_ndwi": 0.33,
                                                      "lambda_evi": 0.33, "lambda_savi": 0.0,
                                                      "lambda_msavi": 0.0, "lambda_gndvi": 0.0}},
        "satclip": {"use_satclip": True, "satclip_style": "inject", "satclip_inject_style": "multiply",
                    "post_correction": post_correction, "post_correction_init": post_correction_init,
                    "scaling_param": True, "scaling_param_init": scale_init},
        "Data": {"padding": True, "padding_amount": 10}})


def px2px_config(lambda_rs: float = 1.0, inject: bool = False, lambda_l1: float = 100.0, **satclip_kw):
    """configs/config_px2px.yaml (plain generator, ``use_satclip: False``) or, with ``inject=True``, the SatCLIP variant;
    loss weights of :24-39 (lambda_GAN 1, lambda_L1 100, lambda_rs_losses 1, internal weights .33/.33/.33)."""
    c = satclip_inject_config(**satclip_kw)
    c.base_configs.lambda_rs_losses = lambda_rs
    c.base_configs.lambda_L1 = lambda_l1
    if not inject:
        c.satclip.use_satclip = False
    return c
