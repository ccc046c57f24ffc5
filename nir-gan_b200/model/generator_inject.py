"""B200 drop-in for the reference's ``model/generator_inject.py`` (SatCLIP-injected generator).

``define_G_inject(config)`` (generator_inject.py:145-200) and ``ResnetGenerator_inject`` keep their
signatures, attribute creation order (``fc`` before ``scale_param`` / ``post_correction_param`` before
``model`` => identical init RNG consumption and state_dict order: scale_param, fc.weight, fc.bias,
model.*) and quirks (square tiles only; ``scale_param`` truthiness).  The forward runs on the
sm_100a kernels: the fc projection is ``ng_linear`` and the bilinear resize + channel broadcast +
``x*(1+s*e)`` is fused into the InstanceNorm-apply kernel of down-1 (the injection point is between
IN and ReLU, generator_inject.py:107,130).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from ..engine import EngineConfig, require_cuda
from ..runners import GeneratorRunner
from .networks import _B200Module, _generator_trunk, _instance_bias, get_norm_layer, init_net


class ResnetGenerator_inject(_B200Module):
    def __init__(self, config, norm_layer, n_blocks=9):
        base, sat = config.base_configs, config.satclip
        if not base.no_dropout:
            raise NotImplementedError("nirgan_b200 ResnetGenerator_inject: dropout is outside the hot path")
        self.post_correction = sat.post_correction
        self.post_correction_init = sat.post_correction_init
        self.scaling_param = sat.scaling_param
        self.scaling_param_init = sat.scaling_param_init
        self.inject_style = sat.satclip_inject_style
        assert n_blocks >= 0
        super().__init__()
        if not _instance_bias(norm_layer):
            raise NotImplementedError("nirgan_b200 ResnetGenerator_inject: instance norm only")
        if base.output_nc != 1:
            raise NotImplementedError("nirgan_b200: the fused head kernel emits one band (output_nc=1)")
        layers = _generator_trunk(base.input_nc, base.output_nc, base.ngf, norm_layer, n_blocks)
        self.n_blocks = n_blocks
        self.b200_config = EngineConfig.from_env()
        self._runner = None
        self.embed_fc_ou_square = 128
        self.fc = nn.Linear(in_features=256, out_features=self.embed_fc_ou_square * self.embed_fc_ou_square)
        if self.scaling_param:
            print("Setting learned scale Parameter with init value: ", self.scaling_param_init)
            self.scale_param = nn.Parameter(torch.tensor(self.scaling_param_init))
        if self.post_correction:
            print("Setting Post-Correction Parameter with init value: ", self.post_correction_init)
            self.post_correction_param = nn.Parameter(torch.tensor(self.post_correction_init))
        self.model = nn.Sequential(*layers)

    def forward(self, input, embeds, wrap_pad: int = 0, reuse_token=None):
        require_cuda(input, "ResnetGenerator_inject input")
        from ..autograd import generator_apply
        return generator_apply(self, self._get_runner(GeneratorRunner), input, embeds, wrap_pad, reuse_token)


def define_G_inject(config):
    """generator_inject.py:145-200: reads config.base_configs.* / config.satclip.*; only
    netG == 'resnet_9blocks' is allowed."""
    base = config.base_configs
    norm_layer = get_norm_layer(norm_type=base.norm)
    if base.netG == "resnet_9blocks":
        net = ResnetGenerator_inject(config, norm_layer=norm_layer, n_blocks=9)
    else:
        raise NotImplementedError(
            "Generator model name [%s] is not recognized. Only resnet_9blocks for SatCLIP." % base.netG)
    return init_net(net, base.init_type, base.init_gain, [])
