"""GPU parity of the optional generator losses (ng_ssim_loss, ng_emd_loss; utils/losses.py:10-29,64-78): values and
d/dpred against the reference-produced (emd) / oracle-produced (ssim) fixtures and against the oracle's autograd on
larger seeded inputs.  Tolerances (fp32 arithmetic, different summation order): SSIM loss 1e-5, SSIM gradient 1e-5 of
its max; EMD loss 2e-7 absolute -- the CDFs are O(1) fp32 running sums, so the reference's own sequential cumsum carries
~1e-7 of rounding per entry (the kernel carries the running sum in fp64 and agrees with the float64 oracle to 1e-4
relative); EMD gradient 2e-2 relative L2 -- the gradient is a step function of sign(cdf_pred - cdf_target), and where the
CDFs cross, a 1e-7 rounding difference moves a step by one element."""
import os

import numpy as np
import pytest
import torch

import postprocess_oracle as P

pytestmark = pytest.mark.gpu
Z = np.load(os.path.join(os.path.dirname(__file__), "golden", "losses_small.npz"))
CASES = sorted({k.split(".")[0] for k in Z.files})


def _run(fn, pred, target, *a):
    p = pred.cuda().requires_grad_(True)
    loss = fn(p, target.cuda(), *a)
    loss.backward()
    return loss.item(), p.grad.cpu()


def _rel(a, b):
    return float((a - b).norm() / (b.norm() + 1e-30))


@pytest.mark.parametrize("case", CASES)
def test_emd_fixture_from_the_reference(case):
    from nirgan_b200.utils.losses import emd_loss, hist_loss
    assert hist_loss is emd_loss
    v, g = _run(emd_loss, torch.from_numpy(Z[f"{case}.pred"]), torch.from_numpy(Z[f"{case}.target"]))
    assert abs(v - float(Z[f"{case}.emd"])) <= 2e-7
    assert _rel(g, torch.from_numpy(Z[f"{case}.emd_grad"])) <= 2e-2


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("win", [5, 11])
def test_ssim_fixture(case, win):
    from nirgan_b200.utils.losses import ssim_loss
    v, g = _run(ssim_loss, torch.from_numpy(Z[f"{case}.pred"]), torch.from_numpy(Z[f"{case}.target"]), win)
    ref_g = torch.from_numpy(Z[f"{case}.ssim{win}_grad"])
    assert abs(v - float(Z[f"{case}.ssim{win}"])) <= 1e-5
    assert float((g - ref_g).abs().max()) <= 1e-5 * float(ref_g.abs().max()) + 1e-9


@pytest.mark.parametrize("shape", [(4, 1, 256, 256), (2, 1, 100, 68), (1, 3, 12, 12), (2, 1, 7, 300)])
def test_ssim_against_oracle_autograd(shape):
    from nirgan_b200.losses import ssim_loss
    g = torch.Generator().manual_seed(sum(shape))
    t = torch.rand(shape, generator=g)
    p = (t + 0.2 * torch.randn(shape, generator=g)).clamp(-1, 1)
    for win in ((5, 11) if min(shape[2:]) > 5 else (5,)):
        v, gr = _run(ssim_loss, p, t, win)
        po = p.clone().requires_grad_(True)
        lo = P.ssim_loss(po, t, win)
        lo.backward()
        assert abs(v - lo.item()) <= 2e-5
        assert float((gr - po.grad).abs().max()) <= 2e-5 * float(po.grad.abs().max())


@pytest.mark.parametrize("shape", [(32, 1, 256, 256), (3, 1, 50, 30), (2, 2, 33, 31), (1, 1, 1, 5)])
def test_emd_against_oracle_autograd(shape):
    from nirgan_b200.losses import emd_loss
    g = torch.Generator().manual_seed(sum(shape))
    t = torch.rand(shape, generator=g)
    p = (t + 0.3 * torch.randn(shape, generator=g)).clamp(-1, 1)
    v, gr = _run(emd_loss, p, t)
    po = p.double().requires_grad_(True)
    lo = P.emd_loss(po, t.double())
    lo.backward()
    assert abs(v - lo.item()) <= 1e-4 * lo.item() + 1e-9
    assert _rel(gr.double(), po.grad) <= 2e-2
    assert gr.shape == p.shape


def test_losses_without_grad_and_errors():
    from nirgan_b200.losses import emd_loss, ssim_loss
    a, b = torch.rand(2, 1, 32, 32, device="cuda"), torch.rand(2, 1, 32, 32, device="cuda")
    assert abs(ssim_loss(a, a).item()) <= 1e-5 and emd_loss(a, a).item() == 0.0
    assert ssim_loss(a, b).requires_grad is False
    with pytest.raises(RuntimeError):
        ssim_loss(a.cpu(), b.cpu())
    with pytest.raises(RuntimeError):
        ssim_loss(a[..., :4, :4].contiguous(), b[..., :4, :4].contiguous(), 11)      # window larger than the image


def test_training_step_with_ssim_and_hist_terms(golden_dir):
    """G-pass loss with lambda_ssim / lambda_hist > 0 (model/pix2pix.py:231-243) equals the sum of its parts computed by
    the oracle on the prediction the model produced, and the step runs through backward + Adam."""
    from nirgan_b200.model.pix2pix import Px2Px
    from nirgan_b200.config import px2px_config as _cfg
    cfg = _cfg(lambda_rs=0.0)
    cfg.base_configs.lambda_ssim, cfg.base_configs.lambda_hist = 2.0, 50.0
    torch.manual_seed(0)
    m = Px2Px(cfg).cuda().train()
    m.netG.configure_b200(precision="fp32", impl="simt")
    m.netD.configure_b200(precision="fp32", impl="simt")
    opt_d, opt_g = m.configure_optimizers()
    batch = {"rgb": torch.rand(2, 3, 64, 64, device="cuda"), "nir": torch.rand(2, 1, 64, 64, device="cuda")}
    with torch.no_grad():
        m.eval()
        pred = m.forward(batch["rgb"]).cpu()
        m.train()
    base = _cfg(lambda_rs=0.0)
    torch.manual_seed(0)
    m0 = Px2Px(base).cuda().train()
    m0.netG.configure_b200(precision="fp32", impl="simt")
    m0.netD.configure_b200(precision="fp32", impl="simt")
    l0 = m0.training_step(batch, 0, 1).item()
    loss = m.training_step(batch, 0, 1)
    nir = batch["nir"].cpu()
    want = l0 + 2.0 * P.ssim_loss(pred, nir).item() + 50.0 * P.emd_loss(pred, nir).item()
    assert abs(loss.item() - want) <= 1e-4 * abs(want)
    opt_g.zero_grad()
    loss.backward()
    opt_g.step()
    assert all(torch.isfinite(p).all() for p in m.netG.parameters())
