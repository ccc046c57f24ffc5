"""CPU ORACLE for the NIR-GAN hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain torch-fp32 *functional* restatement of the reference's arithmetic for the
hot path.  It is keyed on reference state_dict names, so a
reference checkpoint (or the state_dict of the B200 drop-in modules) can be fed to
it unchanged.  Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline``
/ ``--impl reference`` legs of ``bench.py`` may import this module; the product
path (``nir-gan_b200/``) never does.

Pinning: ``oracle/pin_against_reference.py`` imports the real reference modules
from /root/reference (authoring container only), checks every function here
against them on seeded inputs and writes the golden fixtures under
``tests/golden/``.  The reference itself ships no golden vectors.

Each function cites the reference file:line (relative to /root/reference) it follows.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
IN_EPS = 1e-5  # nn.InstanceNorm2d default, model/networks.py:29-30

# Optional emulation of the B200 path's STORAGE precision (test infrastructure for the gradient-parity tests): when set
# to torch.float16 / torch.bfloat16, inputs, weights, every convolution output and every unit output are rounded to that
# type where the kernels store them -- all arithmetic stays fp32 and the rounding is straight-through for autograd.  The
# fp32 autograd of THIS function has the same ReLU / LeakyReLU masks as the low-precision forward, so comparing gradients
# against it isolates the backward kernels' own arithmetic from the (large, unavoidable) effect of mask flips.
# None (default) = the reference's exact fp32 arithmetic; the pinning scripts and golden fixtures never set it.
STORAGE_DTYPE = [None]


class storage_rounding:
    def __init__(self, dtype):
        self.dtype = dtype

    def __enter__(self):
        self.prev = STORAGE_DTYPE[0]
        STORAGE_DTYPE[0] = self.dtype
        return self

    def __exit__(self, *exc):
        STORAGE_DTYPE[0] = self.prev
        return False


def _st(x: Tensor) -> Tensor:
    """Round to the emulated storage type (straight-through gradient); identity by default."""
    d = STORAGE_DTYPE[0]
    if d is None:
        return x
    return x + (x.detach().to(d).float() - x.detach())


# --------------------------------------------------------------------------------------
# building blocks
# --------------------------------------------------------------------------------------
def _inorm(x: Tensor) -> Tensor:
    """InstanceNorm2d(affine=False, track_running_stats=False): biased variance, eps 1e-5.
    model/networks.py:29-30."""
    mu = x.mean(dim=(2, 3), keepdim=True)
    var = x.var(dim=(2, 3), unbiased=False, keepdim=True)
    return (x - mu) / torch.sqrt(var + IN_EPS)


def _rpad(x: Tensor, p: int) -> Tensor:
    """nn.ReflectionPad2d(p) (edge pixel not repeated)."""
    return F.pad(x, (p, p, p, p), mode="reflect")


def generator_layout(n_blocks: int = 9) -> Dict[str, int]:
    """Sequential indices of the parameterised modules of ResnetGenerator
    (model/networks.py:341-368): stem conv at 1, down convs at 4 and 7, ResnetBlocks at
    10..10+n-1, ConvTranspose at 10+n and 13+n, head conv at 17+n."""
    return {
        "stem": 1, "down1": 4, "down2": 7, "block0": 10,
        "up1": 10 + n_blocks, "up2": 13 + n_blocks, "head": 17 + n_blocks,
    }


def resnet_generator_forward(sd: Dict[str, Tensor], x: Tensor, n_blocks: int = 9,
                             embeds: Optional[Tensor] = None,
                             inject_style: str = "multiply",
                             post_correction: bool = False,
                             prefix: str = "") -> Tensor:
    """ResnetGenerator.forward (model/networks.py:372-374, layers :341-368, block :431-434)
    and, when ``embeds`` is given, ResnetGenerator_inject.forward
    (model/generator_inject.py:105-135): the injection is applied to the output of
    model[:6] (i.e. after InstanceNorm of down-1, before its ReLU)."""
    L = generator_layout(n_blocks)
    g = lambda k: sd[prefix + k]
    w = lambda i: (_st(g(f"model.{i}.weight")), g(f"model.{i}.bias"))

    h = _st(F.conv2d(_rpad(_st(x), 3), *w(L["stem"])))             # :341-342
    h = _st(F.relu(_inorm(h)))                                     # :343-344
    h = _inorm(_st(F.conv2d(h, *w(L["down1"]), stride=2, padding=1)))   # :349-350
    if embeds is not None:                                         # generator_inject.py:110-127
        e = F.linear(embeds, g("fc.weight"), g("fc.bias")).view(-1, 1, 128, 128)
        # NOTE the reference passes size=(W, H) (generator_inject.py:116): square tiles only.
        e = F.interpolate(e, size=(h.shape[-1], h.shape[-2]), mode="bilinear", align_corners=False)
        e = e.repeat(1, h.shape[-3], 1, 1)
        s = g("scale_param")
        if inject_style == "add":
            h = h + s * e
        elif inject_style == "multiply" and bool(s):               # truthiness quirk :124
            h = h * (1 + s * e)
        elif inject_style == "multiply":
            h = h * e
    h = _st(F.relu(h))                                             # :351
    h = _st(F.relu(_inorm(_st(F.conv2d(h, *w(L["down2"]), stride=2, padding=1)))))
    for b in range(n_blocks):                                      # :354-356, :405-434
        i = L["block0"] + b
        r = _st(F.conv2d(_rpad(h, 1), _st(g(f"model.{i}.conv_block.1.weight")), g(f"model.{i}.conv_block.1.bias")))
        r = _st(F.relu(_inorm(r)))
        r = _st(F.conv2d(_rpad(r, 1), _st(g(f"model.{i}.conv_block.5.weight")), g(f"model.{i}.conv_block.5.bias")))
        h = _st(h + _inorm(r))                                     # no ReLU after the add, :433
    for key in ("up1", "up2"):                                     # :358-365
        h = _st(F.conv_transpose2d(h, *w(L[key]), stride=2, padding=1, output_padding=1))
        h = _st(F.relu(_inorm(h)))
    h = torch.tanh(F.conv2d(_rpad(h, 3), *w(L["head"])))           # :366-368
    if post_correction:                                            # generator_inject.py:133-134
        h = h * g("post_correction_param")
    return h


def patchgan_forward(sd: Dict[str, Tensor], x: Tensor, n_layers: int = 3, prefix: str = "") -> Tensor:
    """NLayerDiscriminator.forward (model/networks.py:539-584): Conv4x4 s2 + LReLU(0.2);
    (n_layers-1) x [Conv4x4 s2 + IN + LReLU]; Conv4x4 s1 + IN + LReLU; Conv4x4 s1 -> 1 channel.
    All zero-pad 1, bias everywhere (norm = instance)."""
    g = lambda k: sd[prefix + k]
    h = _st(F.leaky_relu(F.conv2d(_st(x), _st(g("model.0.weight")), g("model.0.bias"), stride=2, padding=1), 0.2))
    idx = 2
    for _ in range(1, n_layers):
        h = _st(F.conv2d(h, _st(g(f"model.{idx}.weight")), g(f"model.{idx}.bias"), stride=2, padding=1))
        h = _st(F.leaky_relu(_inorm(h), 0.2))
        idx += 3
    h = _st(F.conv2d(h, _st(g(f"model.{idx}.weight")), g(f"model.{idx}.bias"), stride=1, padding=1))
    h = _st(F.leaky_relu(_inorm(h), 0.2))
    idx += 3
    return F.conv2d(h, _st(g(f"model.{idx}.weight")), g(f"model.{idx}.bias"), stride=1, padding=1)


def lsgan_loss(pred: Tensor, target_is_real: bool) -> Tensor:
    """GANLoss('lsgan').__call__ (model/networks.py:232-233,268-270): MSE against 1.0 / 0.0."""
    t = 1.0 if target_is_real else 0.0
    return ((pred - t) ** 2).mean()


# --------------------------------------------------------------------------------------
# remote-sensing index losses (utils/remote_sensing_indices.py)
# --------------------------------------------------------------------------------------
def _crit(a: Tensor, b: Tensor, criterion: str) -> Tensor:
    return (a - b).abs().mean() if criterion == "l1" else ((a - b) ** 2).mean()


def ndvi_pair(rgb, nir, nir_pred, eps=1e-6):
    """utils/remote_sensing_indices.py:104-110 (eps=1e-6 in loss mode, 0 in index mode)."""
    red = rgb[:, 0:1]
    return (nir - red) / (nir + red + eps), (nir_pred - red) / (nir_pred + red + eps)


def ndwi_pair(rgb, nir, nir_pred, eps=1e-6):
    """utils/remote_sensing_indices.py:140-148."""
    green = rgb[:, 1:2]
    return (nir - green) / (nir + green + eps), (nir_pred - green) / (nir_pred + green + eps)


def evi_pair(rgb, nir, nir_pred, loss_mode=True):
    """utils/remote_sensing_indices.py:296-316.  Product-form denominator, verbatim:
    2.5*(n-R)/((n+6)*(R-7.5)*(B+1) [+1e-6])."""
    red, blue = rgb[:, 0:1], rgb[:, 2:3]
    e = 1e-6 if loss_mode else 0.0
    d = (nir + 6) * (red - 7.5) * (blue + 1) + e
    dp = (nir_pred + 6) * (red - 7.5) * (blue + 1) + e
    return 2.5 * ((nir - red) / d), 2.5 * ((nir_pred - red) / dp)


def gndvi_pair(rgb, nir, nir_pred):
    """utils/remote_sensing_indices.py:169-176 (no eps; divides by ndvi + green, verbatim)."""
    green, red = rgb[:, 1:2], rgb[:, 0:1]
    ndvi, ndvi_p = (nir - red) / (nir + red), (nir_pred - red) / (nir_pred + red)
    return (nir - green) / (ndvi + green), (nir_pred - green) / (ndvi_p + green)


def savi_pair(rgb, nir, nir_pred):
    """utils/remote_sensing_indices.py:202-205."""
    red = rgb[:, 0:1]
    return 1.5 * (nir - red) / (nir + red + 0.5), 1.5 * (nir_pred - red) / (nir_pred + red + 0.5)


def msavi_pair(rgb, nir, nir_pred):
    """utils/remote_sensing_indices.py:232-235."""
    red = rgb[:, 0:1]
    f = lambda n: (2 * n + 1 - torch.sqrt((2 * n + 1) ** 2 - 8 * (n - red))) / 2
    return f(nir), f(nir_pred)


_RS_ORDER = ("lambda_ndvi", "lambda_ndwi", "lambda_gndvi", "lambda_savi", "lambda_msavi", "lambda_evi")
_RS_FN = {"lambda_ndvi": ndvi_pair, "lambda_ndwi": ndwi_pair, "lambda_gndvi": gndvi_pair,
          "lambda_savi": savi_pair, "lambda_msavi": msavi_pair, "lambda_evi": evi_pair}


def rs_weighted_loss(rgb, nir, nir_pred, loss_config=None, criterion="l1"):
    """RemoteSensingIndices.get_and_weight_losses(mode='loss')
    (utils/remote_sensing_indices.py:23-62): iterate ndvi, ndwi, gndvi, savi, msavi, evi;
    add weight*loss for weights > 0; default weights .333/.333/.333."""
    if loss_config is None:
        loss_config = {"lambda_ndvi": 0.333, "lambda_ndwi": 0.333, "lambda_evi": 0.333,
                       "lambda_savi": 0.0, "lambda_msavi": 0.0, "lambda_gndvi": 0.0}
    total = 0.0
    for k in _RS_ORDER:
        wgt = loss_config.get(k, 0.0)
        if wgt > 0.0:
            a, b = _RS_FN[k](rgb, nir, nir_pred)
            total = total + wgt * _crit(a, b, criterion)
    return total


# --------------------------------------------------------------------------------------
# Px2Px_PL restatement (model/pix2pix.py) -- Lightning-free
# --------------------------------------------------------------------------------------
def px2px_forward(sd_g, rgb, pad_amount: int = 10, padding: bool = True, embeds=None,
                  inject_style="multiply", n_blocks=9, prefix=""):
    """Px2Px_PL.forward (model/pix2pix.py:88-110): reflect-pad by padding_amount, run netG,
    crop [pad:-pad]."""
    x = F.pad(rgb, (pad_amount,) * 4, mode="reflect") if padding else rgb
    y = resnet_generator_forward(sd_g, x, n_blocks=n_blocks, embeds=embeds,
                                 inject_style=inject_style, prefix=prefix)
    if padding:
        y = y[..., pad_amount:-pad_amount, pad_amount:-pad_amount]
    return y


DEFAULT_LOSS_CFG = dict(lambda_GAN=1.0, lambda_L1=100.0, lambda_rs_losses=1.0,
                        rs_criterion="l1",
                        rs_weights={"lambda_ndvi": 0.33, "lambda_ndwi": 0.33, "lambda_evi": 0.33,
                                    "lambda_savi": 0.0, "lambda_msavi": 0.0, "lambda_gndvi": 0.0})


def d_loss(sd_g, sd_d, rgb, nir, embeds=None, pad_amount=10, padding=True):
    """Discriminator pass, model/pix2pix.py:195-212: MSE(D(cat(rgb,G(rgb)).detach()),0) +
    MSE(D(cat(rgb,nir)),1); NO 0.5 factor (:206)."""
    with torch.no_grad():
        pred = px2px_forward(sd_g, rgb, pad_amount, padding, embeds)
    pf = patchgan_forward(sd_d, torch.cat((rgb, pred), 1))
    pr = patchgan_forward(sd_d, torch.cat((rgb, nir), 1))
    return lsgan_loss(pf, False) + lsgan_loss(pr, True), pred


def g_loss(sd_g, sd_d, rgb, nir, embeds=None, cfg=None, pad_amount=10, padding=True):
    """Generator pass, model/pix2pix.py:214-257: lambda_GAN*MSE(D(cat(rgb,pred)),1) +
    lambda_L1*L1(pred,nir) + lambda_rs*RS(rgb,nir,pred)."""
    cfg = cfg or DEFAULT_LOSS_CFG
    pred = px2px_forward(sd_g, rgb, pad_amount, padding, embeds)
    pf = patchgan_forward(sd_d, torch.cat((rgb, pred), 1))
    loss = cfg["lambda_GAN"] * lsgan_loss(pf, True) + cfg["lambda_L1"] * (pred - nir).abs().mean()
    if cfg["lambda_rs_losses"] > 0.0:
        loss = loss + cfg["lambda_rs_losses"] * rs_weighted_loss(rgb, nir, pred, cfg["rs_weights"],
                                                                 cfg["rs_criterion"])
    return loss, pred


class OracleTrainer:
    """PL-1.9 automatic optimisation with two optimizers restated (model/pix2pix.py:165-257,
    485-492): per batch, optimizer 0 = D (G frozen), then optimizer 1 = G (D frozen);
    Adam(lr=2e-4, betas=(0.5, 0.999)) each."""

    def __init__(self, sd_g, sd_d, lr=2e-4, beta1=0.5, cfg=None, pad_amount=10, padding=True):
        self.g = {k: v.clone().requires_grad_(True) for k, v in sd_g.items()}
        self.d = {k: v.clone().requires_grad_(True) for k, v in sd_d.items()}
        self.opt_d = torch.optim.Adam(list(self.d.values()), lr=lr, betas=(beta1, 0.999))
        self.opt_g = torch.optim.Adam(list(self.g.values()), lr=lr, betas=(beta1, 0.999))
        self.cfg, self.pad, self.padding = cfg or DEFAULT_LOSS_CFG, pad_amount, padding

    def step(self, rgb, nir, embeds=None, apply_update=True):
        self.opt_d.zero_grad(set_to_none=True)
        ld, _ = d_loss({k: v.detach() for k, v in self.g.items()}, self.d, rgb, nir, embeds,
                       self.pad, self.padding)
        ld.backward()
        grads_d = {k: v.grad.clone() for k, v in self.d.items()}
        if apply_update:
            self.opt_d.step()
        self.opt_g.zero_grad(set_to_none=True)
        # D is frozen during the G pass (PL toggle_optimizer): detach its params.
        lg, pred = g_loss(self.g, {k: v.detach() for k, v in self.d.items()}, rgb, nir, embeds,
                          self.cfg, self.pad, self.padding)
        lg.backward()
        grads_g = {k: (v.grad.clone() if v.grad is not None else torch.zeros_like(v))
                   for k, v in self.g.items()}
        if apply_update:
            self.opt_g.step()
        return {"loss_D": ld.detach(), "loss_G": lg.detach(), "pred": pred.detach(),
                "grads_d": grads_d, "grads_g": grads_g}


# --------------------------------------------------------------------------------------
# inference loop (create_synthetic_dataset.py:93,100-118; data/SR_dataset_RGB.py:16-19,55)
# --------------------------------------------------------------------------------------
def tile_ids(filenames):
    """Ordering contract: sorted() file list, id = fname.split('.')[0]."""
    return [f.split(".")[0] for f in sorted(filenames)]


def synth_loop(sd_g, tiles: Dict[str, Tensor], batch_size: int = 2, pad_amount: int = 10):
    """Sequential reference loop: for each batch of `batch_size` sorted tiles run the padded
    forward; return {id: (1,H,W) float32}."""
    names = sorted(tiles.keys())
    out = {}
    with torch.no_grad():
        for i in range(0, len(names), batch_size):
            chunk = names[i:i + batch_size]
            hr = torch.stack([tiles[n] for n in chunk])
            pred = px2px_forward(sd_g, hr, pad_amount)
            for n, p in zip(chunk, pred):
                out[n.split(".")[0]] = p
    return out


# --------------------------------------------------------------------------------------
# deterministic parameter / input generators shared by tests and bench
# --------------------------------------------------------------------------------------
def generator_param_shapes(input_nc=3, output_nc=1, ngf=64, n_blocks=9, inject=False):
    """Shapes in reference state_dict order."""
    L = generator_layout(n_blocks)
    shp = {}
    if inject:
        shp["scale_param"] = ()
        shp["fc.weight"] = (128 * 128, 256)
        shp["fc.bias"] = (128 * 128,)
    shp[f"model.{L['stem']}.weight"] = (ngf, input_nc, 7, 7); shp[f"model.{L['stem']}.bias"] = (ngf,)
    shp[f"model.{L['down1']}.weight"] = (2 * ngf, ngf, 3, 3); shp[f"model.{L['down1']}.bias"] = (2 * ngf,)
    shp[f"model.{L['down2']}.weight"] = (4 * ngf, 2 * ngf, 3, 3); shp[f"model.{L['down2']}.bias"] = (4 * ngf,)
    for b in range(n_blocks):
        for j in (1, 5):
            shp[f"model.{L['block0'] + b}.conv_block.{j}.weight"] = (4 * ngf, 4 * ngf, 3, 3)
            shp[f"model.{L['block0'] + b}.conv_block.{j}.bias"] = (4 * ngf,)
    shp[f"model.{L['up1']}.weight"] = (4 * ngf, 2 * ngf, 3, 3); shp[f"model.{L['up1']}.bias"] = (2 * ngf,)
    shp[f"model.{L['up2']}.weight"] = (2 * ngf, ngf, 3, 3); shp[f"model.{L['up2']}.bias"] = (ngf,)
    shp[f"model.{L['head']}.weight"] = (output_nc, ngf, 7, 7); shp[f"model.{L['head']}.bias"] = (output_nc,)
    return shp


def discriminator_param_shapes(input_nc=4, ndf=64, n_layers=3):
    shp = {"model.0.weight": (ndf, input_nc, 4, 4), "model.0.bias": (ndf,)}
    idx, prev = 2, 1
    for n in range(1, n_layers):
        m = min(2 ** n, 8)
        shp[f"model.{idx}.weight"] = (ndf * m, ndf * prev, 4, 4); shp[f"model.{idx}.bias"] = (ndf * m,)
        idx += 3; prev = m
    m = min(2 ** n_layers, 8)
    shp[f"model.{idx}.weight"] = (ndf * m, ndf * prev, 4, 4); shp[f"model.{idx}.bias"] = (ndf * m,)
    idx += 3
    shp[f"model.{idx}.weight"] = (1, ndf * m, 4, 4); shp[f"model.{idx}.bias"] = (1,)
    return shp


def random_state_dict(shapes, seed=0, gain=0.02, bias_std=0.0, scale_param=0.01):
    """N(0, gain) weights, zero (or N(0,bias_std)) biases -- init_weights('normal')
    (model/networks.py:79-93) without depending on module registration order."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, s in shapes.items():
        if k in ("scale_param", "post_correction_param"):
            sd[k] = torch.tensor(float(scale_param))
        elif k.endswith("bias"):
            sd[k] = torch.randn(s, generator=g) * bias_std if bias_std > 0 else torch.zeros(s)
        else:
            sd[k] = torch.randn(s, generator=g) * gain
    return sd


def g_forward_gflop(h: int, w: int, ngf=64, n_blocks=9, input_nc=3) -> float:
    """Algorithmic 2*MAC FLOPs of ResnetGenerator forward on an (h, w) input."""
    f = 2 * h * w * ngf * input_nc * 49
    f += 2 * (h // 2) * (w // 2) * (2 * ngf) * ngf * 9
    f += 2 * (h // 4) * (w // 4) * (4 * ngf) * (2 * ngf) * 9
    f += n_blocks * 2 * 2 * (h // 4) * (w // 4) * (4 * ngf) * (4 * ngf) * 9
    f += 2 * (h // 4) * (w // 4) * (4 * ngf) * (2 * ngf) * 9     # convT: one MAC per (in px, tap)
    f += 2 * (h // 2) * (w // 2) * (2 * ngf) * ngf * 9
    f += 2 * h * w * ngf * 1 * 49
    return f / 1e9
