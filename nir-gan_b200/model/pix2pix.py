"""Lightning-free B200 mirror of the hot-path parts of ``model/pix2pix.py::Px2Px_PL``.

Kept: the constructor's network selection (pix2pix.py:18-86), ``forward(input, embeds=None)`` with the
reflect-pad / crop wrapper (:88-110), ``predict_step(rgb, coords)`` (:134-163) and ``extract_batch`` (:426-482) with the SatCLIP location
encoder attached through ``attach_satclip`` (one ``ng_satclip_encode`` launch per batch; batches may also carry
precomputed ``embeds``),
``training_step(batch, batch_idx, optimizer_idx)`` loss composition (:165-257) and
``configure_optimizers`` (:485-492).  Dropped: Lightning hooks, wandb logging, validation plots.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import networks
from .generator_inject import define_G_inject
from ..losses import emd_loss as hist_loss
from ..losses import rs_pixel_losses, ssim_loss


def _get(cfg, name, default=None):
    return getattr(cfg, name, default) if not isinstance(cfg, dict) else cfg.get(name, default)


class Px2Px(nn.Module):
    def __init__(self, opt):
        super().__init__()
        self.opt = opt.base_configs
        self.config = opt
        o, sat = self.opt, opt.satclip
        self.satclip = bool(_get(sat, "use_satclip", False))
        style = _get(sat, "satclip_style", None)
        if self.satclip and style == "concat":
            self.netG = networks.define_G(o.input_nc + 1, o.output_nc, o.ngf, o.netG, o.norm, not o.no_dropout,
                                          o.init_type, o.init_gain)
        elif self.satclip and style == "inject":
            self.netG = define_G_inject(self.config)
        else:
            self.netG = networks.define_G(o.input_nc, o.output_nc, o.ngf, o.netG, o.norm, not o.no_dropout,
                                          o.init_type, o.init_gain)
        self.netD = networks.define_D(o.input_nc + o.output_nc, o.ndf, o.netD, o.n_layers_D, o.norm, o.init_type,
                                      o.init_gain)
        self.criterionGAN = networks.GANLoss(o.gan_mode)
        self.inject = self.satclip and style == "inject"
        self.concat = self.satclip and style == "concat"
        self.satclip_model = None      # pix2pix.py:68-74 builds it from the shipped checkpoint; here it is attached
        # Px2Px_PL.training_step runs G twice per batch (once per optimizer) with identical results;
        # NIRGAN_B200_REUSE_G=0 restores the reference-faithful double evaluation
        import os
        self.reuse_g_forward = os.environ.get("NIRGAN_B200_REUSE_G", "1") != "0"

    # pix2pix.py:88-110 -- the pad / crop is fused into the first / last kernel (wrap_pad)
    def forward(self, input, embeds=None, use_padding=True, reuse_token=None):
        pad = self.config.Data.padding_amount if self.config.Data.padding else 0
        if self.inject:
            return self.netG(input, embeds, wrap_pad=pad, reuse_token=reuse_token)
        return self.netG(input, wrap_pad=pad, reuse_token=reuse_token)

    def attach_satclip(self, satclip_model):
        """The reference constructs ``SatClIP_wrapper(device=self.device).eval()`` itself (pix2pix.py:68-74); the
        checkpoint is not part of the repository, so the encoder (``nirgan_b200.model.satclip.satclip_wrapper``) is
        handed in."""
        self.satclip_model = satclip_model
        return self

    def satclip_get_inject(self, coords):               # pix2pix.py:478-482
        if self.satclip_model is None:
            raise RuntimeError("use_satclip is set but no SatCLIP encoder is attached (attach_satclip) and the batch "
                               "carries no precomputed 'embeds'")
        return self.satclip_model.predict(coords)

    def satclip_get_concat(self, coords, rgb):          # pix2pix.py:466-476
        e = self.satclip_get_inject(coords)
        e = e.view(rgb.shape[0], 1, 1, 256).expand(rgb.shape[0], 1, 256, 256)
        e = torch.nn.functional.interpolate(e, size=(rgb.shape[-1], rgb.shape[-2]), mode="bicubic")
        return torch.cat((rgb, e * self.config.satclip.scaling_factor), dim=1)

    @torch.no_grad()
    def forward_async(self, input, embeds=None, ready=None):
        """Streaming form of ``forward`` for inference loops (``synth.run_shard(model_async=...)``): returns
        ``(pred, done_events)`` without joining the caller's stream; the pad / crop wrapper is applied as in ``forward``."""
        pad = self.config.Data.padding_amount if self.config.Data.padding else 0
        return self.netG.forward_async(input, embeds if self.inject else None, pad, ready)

    @torch.no_grad()
    def predict_step(self, rgb, coords=None, embeds=None):
        assert self.training is False, "Model is in training mode, set to eval mode before predicting"
        if not self.satclip:
            return self.forward(rgb)
        batch = {"rgb": rgb, "nir": None, "coords": coords, "embeds": embeds}
        if self.concat:
            rgb, _ = self.extract_batch(batch)
            return self.forward(rgb)
        if self.inject:
            rgb, _, embeds = self.extract_batch(batch)
            return self.forward(rgb, embeds)
        raise NotImplementedError("SatClip Style not recognized, choose 'concat' or 'inject'")

    def extract_batch(self, batch):                     # pix2pix.py:426-463
        rgb, nir = batch["rgb"], batch["nir"]
        if not self.satclip:
            return rgb, nir
        if self.concat:
            return self.satclip_get_concat(batch["coords"], rgb), nir
        if self.inject:
            embeds = batch.get("embeds")
            if embeds is None:
                embeds = self.satclip_get_inject(batch["coords"])
            return rgb, nir, embeds
        raise NotImplementedError("SatClip Style not recognized, choose 'concat' or 'inject'")

    def _toggle(self, active: nn.Module, frozen: nn.Module):
        """PL-1.9 ``toggle_optimizer``: parameters of the optimizer that is not stepping do not require grad."""
        saved = [(p, p.requires_grad) for p in frozen.parameters()]
        for p, _ in saved:
            p.requires_grad_(False)
        return saved

    def training_step(self, batch, batch_idx, optimizer_idx):
        assert self.training is True, "Model is in eval mode, set to training mode before training"
        frozen = self._toggle(self.netD, self.netG) if optimizer_idx == 0 else self._toggle(self.netG, self.netD)
        try:
            return self._training_step(batch, batch_idx, optimizer_idx)
        finally:
            for p, rg in frozen:
                p.requires_grad_(rg)

    def _training_step(self, batch, batch_idx, optimizer_idx):
        if self.inject:
            rgb, nir, embeds = self.extract_batch(batch)
        else:
            (rgb, nir), embeds = self.extract_batch(batch), None
        o = self.config.base_configs
        self.netD.reset_training_slots()
        self.netG.reset_training_slots()
        if optimizer_idx == 0:      # discriminator, pix2pix.py:195-212
            self._shared = None
            with torch.no_grad():    # fake_AB.detach(): G needs no graph in the D pass
                if self.reuse_g_forward:
                    # the G pass of this step evaluates G on the same batch with the same weights: keep the activations
                    # and hand them over through an explicit token (held together with the batch tensors themselves, so
                    # neither their memory nor their identity can be recycled in between)
                    pad = self.config.Data.padding_amount if self.config.Data.padding else 0
                    pred, token = self.netG.forward_shared(rgb, embeds if self.inject else None, wrap_pad=pad)
                    self._shared = (token, rgb, rgb._version, embeds, None if embeds is None else embeds._version)
                else:
                    pred = self.forward(rgb, embeds)
            # torch.cat((rgb, pred), 1) / torch.cat((rgb, nir), 1) are fused into the PatchGAN's input kernel and the fake
            # and the real batch go through D as one 2B batch (per-sample InstanceNorm: same values as two calls)
            B = rgb.shape[0]
            patches = self.netD.forward_parts([(rgb, pred), (rgb, nir)])
            loss_D_fake = self.criterionGAN(patches[:B], False)
            loss_D_real = self.criterionGAN(patches[B:], True)
            return loss_D_fake + loss_D_real          # no 0.5 factor (pix2pix.py:206)
        # generator, pix2pix.py:214-257
        token = None
        sh, self._shared = getattr(self, "_shared", None), None
        if sh is not None and sh[1] is rgb and sh[2] == rgb._version and sh[3] is embeds and \
                (embeds is None or sh[4] == embeds._version):
            token = sh[0]
        pred = self.forward(rgb, embeds, reuse_token=token)
        pred_fake = self.netD(rgb, pred)
        loss_G = self.criterionGAN(pred_fake, True) * o.lambda_GAN
        extra = []                                             # pix2pix.py:231-243, added after GAN + L1
        if _get(o, "lambda_ssim", 0.0) > 0.0:
            extra.append(ssim_loss(pred, nir) * o.lambda_ssim)
        if _get(o, "lambda_hist", 0.0) > 0.0:
            extra.append(hist_loss(pred, nir) * o.lambda_hist)
        lam_rs = float(_get(o, "lambda_rs_losses", 0.0))
        w = _get(o, "internal_rs_loss_weights", None)
        wd = dict(w) if isinstance(w, dict) else (vars(w) if w is not None else {})
        crit = _get(o, "rs_losses_criterium", "l1")
        if crit not in ("l1", "l2"):
            raise NotImplementedError(f"Criterion '{crit}' not implemented. 'l1' or 'l2' are supported.")
        # term order of the fused kernel = the reference's iteration order (remote_sensing_indices.py:45-52)
        keys = ("lambda_ndvi", "lambda_ndwi", "lambda_gndvi", "lambda_savi", "lambda_msavi", "lambda_evi")
        ws = [float(o.lambda_L1)] + [lam_rs * float(wd.get(k, 0.0)) if (lam_rs > 0.0 and float(wd.get(k, 0.0)) > 0.0)
                                     else 0.0 for k in keys]
        parts = rs_pixel_losses(rgb, nir, pred, ws, crit)    # one fused pass: L1 + the weighted indices (+ d/dpred)
        if ws[0] != 0.0:
            loss_G = loss_G + ws[0] * parts[0]
        for e in extra:
            loss_G = loss_G + e
        for wi, pi in zip(ws[1:], parts[1:]):
            if wi != 0.0:
                loss_G = loss_G + wi * pi
        return loss_G

    def configure_optimizers(self):
        """pix2pix.py:485-492: [optim_d, optim_g], Adam(lr, betas=(beta1, 0.999)); the update runs in ng_adam_step."""
        from ..optim import B200Adam
        o = self.opt
        optim_g = B200Adam(self.netG.parameters(), lr=o.lr, betas=(o.beta1, 0.999))
        optim_d = B200Adam(self.netD.parameters(), lr=o.lr, betas=(o.beta1, 0.999))
        return [optim_d, optim_g]
