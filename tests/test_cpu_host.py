"""CPU suite: host-side logic of the drop-in (no compute calls -- there is no GPU here)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import nirgan_b200
from nirgan_b200 import _lib as L
from nirgan_b200.model import networks
from nirgan_b200.model.generator_inject import define_G_inject

from nirgan_b200.config import satclip_inject_config as inject_config

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_every_declared_symbol():
    lib = L.load()
    header = open(os.path.join(ROOT, "include", "nirgan_b200.h")).read()
    declared = set(re.findall(r"\b(ng_[a-z0-9_]+)\s*\(", header))
    declared -= {"ng_conv_args"}
    assert declared, "no declarations parsed"
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/nirgan_b200.h but not exported"
    assert declared == set(L.EXPORTED_SYMBOLS), declared ^ set(L.EXPORTED_SYMBOLS)
    assert lib.ng_version() == 102


def test_conv_args_struct_layout_matches_header():
    # 20 int32 + float + 2 int32 = 92 bytes, pointers 8-aligned from offset 96; 8 pointers
    assert L.ConvArgs.x.offset == 96 and ctypes.sizeof(L.ConvArgs) == 160 and L.ConvArgs.stat_acc.offset == 152
    assert L.ConvArgs.slope.offset == 80 and L.ConvArgs.crop.offset == 84


def test_no_gpu_means_loud_failure_not_fallback():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    a = L.ConvArgs()
    st = L.load().ng_conv2d(ctypes.byref(a), None)
    assert st == -4 and "fallback" in L.last_error()       # NG_E_ARCH
    net = networks.define_G(3, 1, 64, "resnet_9blocks", "instance", False, "normal", 0.02)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        net(torch.rand(1, 3, 32, 32))


def test_state_dict_layout_and_init_match_reference(golden_dir):
    fp = np.load(f"{golden_dir}/init_fingerprints.npz")

    def check(net, tag):
        sd = net.state_dict()
        assert list(sd.keys()) == list(fp[tag + "_keys"]), tag
        for k, v in sd.items():
            ref = fp[f"{tag}.{k}"]
            got = np.array([float(v.double().sum()), float(v.double().abs().sum()), float(v.flatten()[0])])
            assert np.allclose(got, ref, rtol=0, atol=1e-9), (tag, k)   # same seed -> bit-identical weights

    torch.manual_seed(0)
    check(networks.define_G(3, 1, 64, "resnet_9blocks", "instance", False, "normal", 0.02), "G")
    torch.manual_seed(0)
    check(networks.define_D(4, 64, "basic", 3, "instance", "normal", 0.02), "D")
    torch.manual_seed(0)
    check(define_G_inject(inject_config()), "Gi")


def test_parameter_counts():
    g = networks.define_G(3, 1, 64, "resnet_9blocks", "instance")
    d = networks.define_D(4, 64, "basic", 3, "instance")
    gi = define_G_inject(inject_config())
    assert sum(p.numel() for p in g.parameters()) == 11_371_905
    assert sum(p.numel() for p in d.parameters()) == 2_765_761
    assert sum(p.numel() for p in gi.parameters()) == 15_582_594
    assert list(gi.state_dict().keys())[:3] == ["scale_param", "fc.weight", "fc.bias"]


def test_factory_errors_mirror_reference():
    with pytest.raises(NotImplementedError, match=r"Generator model name \[foo\] is not recognized"):
        networks.define_G(3, 1, 64, "foo", "instance")
    with pytest.raises(NotImplementedError, match=r"Discriminator model name \[bar\] is not recognized"):
        networks.define_D(4, 64, "bar", 3, "instance")
    with pytest.raises(NotImplementedError, match="not found"):
        networks.get_norm_layer("nope")
    with pytest.raises(NotImplementedError, match="not implemented"):
        networks.GANLoss("hinge")
    cfg = inject_config()
    cfg.base_configs.netG = "resnet_6blocks"
    with pytest.raises(NotImplementedError, match="Only resnet_9blocks for SatCLIP"):
        define_G_inject(cfg)
    gl = networks.GANLoss("lsgan")
    assert set(gl.state_dict().keys()) == {"real_label", "fake_label"}


def test_reference_checkpoint_keys_load_strict():
    import nirgan_oracle as O
    g = networks.define_G(3, 1, 64, "resnet_9blocks", "instance")
    g.load_state_dict(O.random_state_dict(O.generator_param_shapes(), seed=1), strict=True)
    d = networks.define_D(4, 64, "basic", 3, "instance")
    d.load_state_dict(O.random_state_dict(O.discriminator_param_shapes(), seed=1), strict=True)
