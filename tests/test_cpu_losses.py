"""CPU suite for the optional-loss oracles (oracle/postprocess_oracle.py::emd_loss / ssim_loss).  The emd fixtures were
produced by the REFERENCE's utils/losses.py::emd_loss (oracle/pin_losses.py, tests/golden/PIN_REPORT_losses.txt); the ssim
fixtures by the oracle itself (kornia is not installed: unpinned) and are checked against SSIM's defining properties."""
import os

import numpy as np
import torch

import postprocess_oracle as P

Z = np.load(os.path.join(os.path.dirname(__file__), "golden", "losses_small.npz"))
CASES = sorted({k.split(".")[0] for k in Z.files})


def test_emd_oracle_reproduces_the_reference_values_and_gradients():
    for c in CASES:
        pred = torch.from_numpy(Z[f"{c}.pred"]).requires_grad_(True)
        loss = P.emd_loss(pred, torch.from_numpy(Z[f"{c}.target"]))
        loss.backward()
        assert loss.item() == float(Z[f"{c}.emd"])
        assert torch.equal(pred.grad, torch.from_numpy(Z[f"{c}.emd_grad"]))


def test_emd_is_zero_for_identical_inputs_and_shift_invariant():
    x = torch.rand(2, 1, 16, 16)
    assert P.emd_loss(x, x).item() == 0.0
    y = torch.rand(2, 1, 16, 16)
    assert abs(P.emd_loss(x + 3.0, y).item() - P.emd_loss(x, y).item()) < 1e-7      # softmax ignores a constant offset


def test_ssim_loss_properties():
    x = torch.rand(2, 1, 32, 32)
    assert abs(P.ssim_loss(x, x, 11).item()) < 1e-6
    y = torch.rand(2, 1, 32, 32)
    a, b = P.ssim_loss(x, y, 11).item(), P.ssim_loss(y, x, 11).item()
    assert abs(a - b) < 1e-6 and 0.0 < a < 2.0
    for c in CASES:
        for w in (5, 11):
            v = P.ssim_loss(torch.from_numpy(Z[f"{c}.pred"]), torch.from_numpy(Z[f"{c}.target"]), w).item()
            assert abs(v - float(Z[f"{c}.ssim{w}"])) < 1e-6


def test_ssim_map_agrees_with_an_independent_scipy_formulation():
    """kornia is not installed, so the SSIM oracle cannot be pinned against it; as a second opinion the same published
    formula is evaluated with scipy.ndimage (float64, mode='mirror' = reflect border without repeating the edge) and
    must agree with the torch restatement to float32 rounding."""
    from scipy import ndimage
    g = torch.Generator().manual_seed(9)
    a = torch.rand(1, 1, 40, 33, generator=g)
    b = (a + 0.2 * torch.randn(1, 1, 40, 33, generator=g)).clamp(0, 1)
    for win in (5, 11):
        k = P.gaussian_kernel1d(win).double().numpy()

        def filt(x):
            return ndimage.correlate1d(ndimage.correlate1d(x, k, axis=0, mode="mirror"), k, axis=1, mode="mirror")

        x, y = a[0, 0].double().numpy(), b[0, 0].double().numpy()
        m1, m2 = filt(x), filt(y)
        s1, s2, s12 = filt(x * x) - m1 * m1, filt(y * y) - m2 * m2, filt(x * y) - m1 * m2
        c1, c2 = 0.01 ** 2, 0.03 ** 2
        ref = ((2 * m1 * m2 + c1) * (2 * s12 + c2)) / ((m1 * m1 + m2 * m2 + c1) * (s1 + s2 + c2) + 1e-12)
        got = P.ssim_map(a, b, win)[0, 0].double().numpy()
        assert np.abs(got - ref).max() <= 2e-5
        assert abs((1.0 - ref.mean()) - P.ssim_loss(a, b, win).item()) <= 1e-5
