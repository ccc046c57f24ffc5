"""B200 drop-in for the hot-path subset of the reference's ``model/networks.py``.

Same factory signatures (``define_G`` networks.py:120, ``define_D`` :163), same module tree and
therefore the same ``state_dict`` keys / shapes / init RNG order as the reference
(``model.{1,4,7,...}.weight``, ``model.N.conv_block.{1,5}.weight`` ...), so reference checkpoints
load unchanged and ``torch.manual_seed(s); define_G(...)`` yields bit-identical initial weights.
The ``nn`` layers inside ``self.model`` are parameter containers only: ``forward`` hands the fp32
masters to the sm_100a kernels through ``engine.GeneratorRunner`` / ``PatchGANRunner``.
No CPU fallback: a CPU tensor raises.

Accelerated selections: netG in {resnet_9blocks, resnet_6blocks}, netD in {basic, n_layers},
norm='instance', no dropout, gan_mode='lsgan' (the only ones the shipped configs use,
configs/config_px2px.yaml:12-19).  Other names raise NotImplementedError like the reference does
for unknown names (networks.py:159,203), with a message that says what is out of scope.
"""
from __future__ import annotations

import functools

import torch
import torch.nn as nn
from torch.nn import init

from ..engine import EngineConfig, require_cuda
from ..runners import GeneratorRunner, PatchGANRunner


# ---------------------------------------------------------------------------------------------
# helpers with the reference's names / semantics
# ---------------------------------------------------------------------------------------------
def get_norm_layer(norm_type="instance"):
    """networks.py:18-36.  Only 'instance' (affine=False, no running stats) is on the hot path."""
    if norm_type == "instance":
        return functools.partial(nn.InstanceNorm2d, affine=False, track_running_stats=False)
    if norm_type in ("batch", "none"):
        raise NotImplementedError(
            "normalization layer [%s] is outside the nirgan_b200 hot path (only 'instance' is accelerated)" % norm_type)
    raise NotImplementedError("normalization layer [%s] is not found" % norm_type)


def init_weights(net, init_type="normal", init_gain=0.02):
    """networks.py:68-99: every module whose class name contains 'Conv' or 'Linear' gets its weight
    drawn per ``init_type`` and a zero bias, visited in ``net.apply`` order."""
    fillers = {
        "normal": lambda w: init.normal_(w, 0.0, init_gain),
        "xavier": lambda w: init.xavier_normal_(w, gain=init_gain),
        "kaiming": lambda w: init.kaiming_normal_(w, a=0, mode="fan_in"),
        "orthogonal": lambda w: init.orthogonal_(w, gain=init_gain),
    }

    def visit(m):
        cls = type(m).__name__
        if hasattr(m, "weight") and ("Conv" in cls or "Linear" in cls):
            if init_type not in fillers:
                raise NotImplementedError("initialization method [%s] is not implemented" % init_type)
            fillers[init_type](m.weight.data)
            if getattr(m, "bias", None) is not None:
                init.constant_(m.bias.data, 0.0)

    net.apply(visit)


def init_net(net, init_type="normal", init_gain=0.02, gpu_ids=[]):
    """networks.py:102-117.  ``gpu_ids`` non-empty moves the net to that device (one process per GPU
    is the B200 design; nn.DataParallel is not used)."""
    if len(gpu_ids) > 0:
        assert torch.cuda.is_available()
        net.to(gpu_ids[0])
    init_weights(net, init_type, init_gain=init_gain)
    return net


def _instance_bias(norm_layer) -> bool:
    f = norm_layer.func if isinstance(norm_layer, functools.partial) else norm_layer
    return f == nn.InstanceNorm2d


# ---------------------------------------------------------------------------------------------
# generator
# ---------------------------------------------------------------------------------------------
class ResnetBlock(nn.Module):
    """networks.py:377-434: x + [ReflPad1, Conv3x3, IN, ReLU, ReflPad1, Conv3x3, IN](x).
    ``conv_block`` indices 1 and 5 hold the parameters."""

    def __init__(self, dim, padding_type, norm_layer, use_dropout, use_bias):
        super().__init__()
        if padding_type != "reflect" or use_dropout:
            raise NotImplementedError("nirgan_b200 ResnetBlock: only reflect padding without dropout is accelerated")
        half = lambda: [nn.ReflectionPad2d(1), nn.Conv2d(dim, dim, kernel_size=3, padding=0, bias=use_bias),
                        norm_layer(dim)]
        self.conv_block = nn.Sequential(*half(), nn.ReLU(True), *half())


def _generator_trunk(input_nc, output_nc, ngf, norm_layer, n_blocks):
    """Layer list of networks.py:341-368 from a compact spec (same module order => same keys)."""
    bias = _instance_bias(norm_layer)
    seq = [nn.ReflectionPad2d(3), nn.Conv2d(input_nc, ngf, kernel_size=7, padding=0, bias=bias), norm_layer(ngf),
           nn.ReLU(True)]
    ch = ngf
    for _ in range(2):
        seq += [nn.Conv2d(ch, 2 * ch, kernel_size=3, stride=2, padding=1, bias=bias), norm_layer(2 * ch), nn.ReLU(True)]
        ch *= 2
    seq += [ResnetBlock(ch, "reflect", norm_layer, False, bias) for _ in range(n_blocks)]
    for _ in range(2):
        seq += [nn.ConvTranspose2d(ch, ch // 2, kernel_size=3, stride=2, padding=1, output_padding=1, bias=bias),
                norm_layer(ch // 2), nn.ReLU(True)]
        ch //= 2
    seq += [nn.ReflectionPad2d(3), nn.Conv2d(ngf, output_nc, kernel_size=7, padding=0), nn.Tanh()]
    return seq


class _B200Module(nn.Module):
    """Common plumbing: engine configuration and lazy runner."""

    def configure_b200(self, precision=None, impl=None, chunk=None, streams=None):
        cfg = self.b200_config
        self.b200_config = EngineConfig(precision or cfg.precision, impl or cfg.impl,
                                        cfg.chunk if chunk is None else chunk,
                                        cfg.streams if streams is None else streams)
        self._runner = None
        return self

    def reset_training_slots(self):
        """Forget activations of training forwards whose backward will never run."""
        r = getattr(self, "_runner", None)
        if r is not None:
            r._live = 0
            for c in r._train.values():
                if isinstance(c, dict) and "live" in c:
                    c["live"] = False

    @torch.no_grad()
    def forward_shared(self, input, embeds=None, wrap_pad: int = 0):
        """Generator forward WITHOUT an autograd graph that nevertheless keeps its activations.  Returns
        ``(pred, token)``: handing ``token`` to the next grad-enabled ``forward(..., reuse_token=token)`` on the SAME
        batch lets it adopt these activations instead of recomputing them (the D pass and the G pass of one training
        step evaluate G on the same batch with the same weights, model/pix2pix.py:177-180).  The token dies with any
        other training forward of this shape and with any change of the generator's weights; the caller vouches for the
        batch being the same.  ``pred`` is a view of the plan's output buffer (valid until that next forward)."""
        require_cuda(input, "generator input")
        c = self._get_runner(GeneratorRunner).train_forward(input, embeds, wrap_pad, share=True)
        B, _, H, W = c["geom"]
        out = c["fwd"].records["out"].view(B, 1, H, W)
        if getattr(self, "post_correction", False):
            out = out * self.post_correction_param
        return out, c["share_token"][0]

    @torch.no_grad()
    def forward_async(self, input, embeds=None, wrap_pad: int = 0, ready=None):
        """Streaming inference (no autograd): ``(out, done_events)`` without making the caller's stream wait, so that
        consecutive calls overlap; see ``GeneratorRunner.forward_async``.  Wait on every event before reading ``out``."""
        require_cuda(input, "generator input")
        return self._get_runner(GeneratorRunner).forward_async(input, embeds, wrap_pad, ready)

    def _get_runner(self, factory):
        if getattr(self, "_runner", None) is None:
            object.__setattr__(self, "_runner", factory(self, self.b200_config))
        return self._runner


class ResnetGenerator(_B200Module):
    """networks.py:316-374 with the forward executed by sm_100a kernels."""

    def __init__(self, input_nc, output_nc, ngf=64, norm_layer=nn.BatchNorm2d, use_dropout=False, n_blocks=6,
                 padding_type="reflect"):
        assert n_blocks >= 0
        super().__init__()
        if not _instance_bias(norm_layer) or use_dropout or padding_type != "reflect":
            raise NotImplementedError("nirgan_b200 ResnetGenerator: instance norm, reflect padding, no dropout only")
        if output_nc != 1:
            raise NotImplementedError("nirgan_b200 ResnetGenerator: the fused head kernel emits one band (output_nc=1)")
        self.n_blocks = n_blocks
        self.b200_config = EngineConfig.from_env()
        self._runner = None
        self.model = nn.Sequential(*_generator_trunk(input_nc, output_nc, ngf, norm_layer, n_blocks))

    def forward(self, input, wrap_pad: int = 0, reuse_token=None):
        """(B, input_nc, H, W) fp32 CUDA -> (B, 1, H, W) fp32.  ``wrap_pad`` fuses Px2Px_PL.forward's
        reflect-pad / crop (pix2pix.py:91-93,107-108) into the first and last kernels; ``reuse_token``: see
        ``forward_shared``."""
        require_cuda(input, "ResnetGenerator input")
        from ..autograd import generator_apply
        return generator_apply(self, self._get_runner(GeneratorRunner), input, None, wrap_pad, reuse_token)


def define_G(input_nc, output_nc, ngf, netG, norm="batch", use_dropout=False, init_type="normal", init_gain=0.02,
             gpu_ids=[]):
    """networks.py:120-160."""
    norm_layer = get_norm_layer(norm_type=norm)
    if netG == "resnet_9blocks":
        net = ResnetGenerator(input_nc, output_nc, ngf, norm_layer=norm_layer, use_dropout=use_dropout, n_blocks=9)
    elif netG == "resnet_6blocks":
        net = ResnetGenerator(input_nc, output_nc, ngf, norm_layer=norm_layer, use_dropout=use_dropout, n_blocks=6)
    elif netG in ("unet_128", "unet_256"):
        raise NotImplementedError("Generator model name [%s] is outside the nirgan_b200 hot path "
                                  "(no shipped config selects it)" % netG)
    else:
        raise NotImplementedError("Generator model name [%s] is not recognized" % netG)
    return init_net(net, init_type, init_gain, gpu_ids)


# ---------------------------------------------------------------------------------------------
# discriminator
# ---------------------------------------------------------------------------------------------
class NLayerDiscriminator(_B200Module):
    """networks.py:539-584 (70x70 PatchGAN for n_layers=3)."""

    def __init__(self, input_nc, ndf=64, n_layers=3, norm_layer=nn.BatchNorm2d):
        super().__init__()
        if not _instance_bias(norm_layer):
            raise NotImplementedError("nirgan_b200 NLayerDiscriminator: instance norm only")
        bias = True
        self.n_layers = n_layers
        self.b200_config = EngineConfig.from_env()
        self._runner = None
        seq = [nn.Conv2d(input_nc, ndf, kernel_size=4, stride=2, padding=1), nn.LeakyReLU(0.2, True)]
        mult = 1
        for n in range(1, n_layers + 1):
            prev, mult = mult, min(2 ** n, 8)
            stride = 2 if n < n_layers else 1
            seq += [nn.Conv2d(ndf * prev, ndf * mult, kernel_size=4, stride=stride, padding=1, bias=bias),
                    norm_layer(ndf * mult), nn.LeakyReLU(0.2, True)]
        seq += [nn.Conv2d(ndf * mult, 1, kernel_size=4, stride=1, padding=1)]
        self.model = nn.Sequential(*seq)

    def forward(self, input, second=None):
        """``netD(x)`` as in the reference; ``netD(rgb, pred)`` is ``netD(torch.cat((rgb, pred), 1))`` with the
        concatenation fused into the input kernel (model/pix2pix.py:197,202,216)."""
        require_cuda(input, "NLayerDiscriminator input")
        from ..autograd import discriminator_apply
        return discriminator_apply(self, self._get_runner(PatchGANRunner), [(input, second)])

    def forward_parts(self, parts):
        """Several (first, second | None) inputs of equal shape as ONE batch: returns the stacked (len(parts)*B, 1, h, w)
        patch map.  The D pass feeds its fake and its real batch through the network together this way."""
        for a, _ in parts:
            require_cuda(a, "NLayerDiscriminator input")
        from ..autograd import discriminator_apply
        return discriminator_apply(self, self._get_runner(PatchGANRunner), list(parts))


def define_D(input_nc, ndf, netD, n_layers_D=3, norm="batch", init_type="normal", init_gain=0.02, gpu_ids=[]):
    """networks.py:163-204."""
    norm_layer = get_norm_layer(norm_type=norm)
    if netD == "basic":
        net = NLayerDiscriminator(input_nc, ndf, n_layers=3, norm_layer=norm_layer)
    elif netD == "n_layers":
        net = NLayerDiscriminator(input_nc, ndf, n_layers_D, norm_layer=norm_layer)
    elif netD == "pixel":
        raise NotImplementedError("Discriminator model name [pixel] is outside the nirgan_b200 hot path")
    else:
        raise NotImplementedError("Discriminator model name [%s] is not recognized" % netD)
    return init_net(net, init_type, init_gain, gpu_ids)


# ---------------------------------------------------------------------------------------------
# GAN loss
# ---------------------------------------------------------------------------------------------
class GANLoss(nn.Module):
    """networks.py:210-276.  'lsgan' = MSE against the real/fake label buffers (kept in the
    state_dict as ``real_label`` / ``fake_label``); computed by the fused ng_lsgan_loss kernel."""

    def __init__(self, gan_mode, target_real_label=1.0, target_fake_label=0.0):
        super().__init__()
        self.register_buffer("real_label", torch.tensor(target_real_label))
        self.register_buffer("fake_label", torch.tensor(target_fake_label))
        self.gan_mode = gan_mode
        self._labels = (float(target_real_label), float(target_fake_label))
        if gan_mode == "lsgan":
            pass
        elif gan_mode in ("vanilla", "wgangp"):
            raise NotImplementedError("gan mode %s is outside the nirgan_b200 hot path (lsgan only)" % gan_mode)
        else:
            raise NotImplementedError("gan mode %s not implemented" % gan_mode)

    def get_target_tensor(self, prediction, target_is_real):
        return (self.real_label if target_is_real else self.fake_label).expand_as(prediction)

    def __call__(self, prediction, target_is_real):
        require_cuda(prediction, "GANLoss prediction")
        from ..autograd import lsgan_apply
        return lsgan_apply(prediction, self._labels[0] if target_is_real else self._labels[1])
