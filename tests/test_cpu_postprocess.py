"""CPU suite for the post-processing oracle (oracle/postprocess_oracle.py): the numpy restatement of
skimage.exposure.match_histograms is checked against the algorithm's defining properties, since scikit-image itself
is not available to generate golden vectors (parity unpinned, see the oracle's header)."""
import numpy as np
import torch

import postprocess_oracle as P


def _rng(seed):
    return np.random.default_rng(seed)


def test_matching_an_image_to_itself_is_the_identity():
    a = _rng(0).normal(size=(37, 41)).astype(np.float32)
    assert np.array_equal(P.match_cumulative_cdf(a, a), a)


def test_output_is_a_monotone_function_of_the_input_and_stays_in_the_reference_range():
    r = _rng(1)
    src = r.normal(size=(64, 64)).astype(np.float32)
    ref = (r.random(size=(64, 64)) ** 2 * 0.4).astype(np.float32)
    out = P.match_cumulative_cdf(src, ref)
    order = np.argsort(src.ravel(), kind="stable")
    assert np.all(np.diff(out.ravel()[order]) >= 0)
    assert out.min() >= ref.min() - 1e-7 and out.max() <= ref.max() + 1e-7
    assert out.max() == ref.max()                      # the largest source value has quantile 1 -> largest reference value


def test_quantiles_are_transferred():
    """After matching, the empirical CDF of the output at the reference's quantile points agrees with the reference's."""
    r = _rng(2)
    src = r.normal(size=(128, 128)).astype(np.float32)
    ref = r.gamma(2.0, 0.1, size=(128, 128)).astype(np.float32)
    out = P.match_cumulative_cdf(src, ref)
    for q in (0.1, 0.25, 0.5, 0.75, 0.9):
        assert abs(np.quantile(out, q) - np.quantile(ref, q)) <= 2e-3


def test_ties_get_one_value_and_small_case_by_hand():
    src = np.array([[0.0, 1.0], [1.0, 3.0]], dtype=np.float32)      # quantiles: 0 -> .25, 1 -> .75, 3 -> 1
    ref = np.array([[10.0, 20.0], [30.0, 40.0]], dtype=np.float32)  # quantile points .25,.5,.75,1 -> 10,20,30,40
    out = P.match_cumulative_cdf(src, ref)
    assert np.array_equal(out, np.array([[10.0, 30.0], [30.0, 40.0]], dtype=np.float32))
    neg = np.array([[-0.0, 0.0]], dtype=np.float32)                  # np.unique: -0.0 == 0.0
    assert np.array_equal(P.match_cumulative_cdf(neg, np.array([[5.0, 7.0]], dtype=np.float32)), [[7.0, 7.0]])


def test_reference_loop_shapes_and_dtype():
    g = torch.Generator().manual_seed(3)
    pred = torch.rand(2, 1, 32, 32, generator=g) * 2 - 1
    s2 = torch.rand(2, 1, 8, 8, generator=g) * 0.3
    out = P.postprocess(pred, s2)
    assert out.shape == (2, 1, 32, 32) and out.dtype == torch.float16
    for b in range(2):                                  # every output value is between two reference values
        assert float(out[b].min()) >= float(s2[b].min()) - 1e-3 and float(out[b].max()) <= float(s2[b].max()) + 1e-3


def test_metrics_oracle_properties():
    g = torch.Generator().manual_seed(4)
    a = torch.rand(2, 1, 40, 52, generator=g)
    m = P.calculate_metrics(a, a, "val")
    assert m["val/L1"] == 0.0 and m["val/L2"] == 0.0 and m["val/PSNR"] == float("inf")
    assert abs(m["val/SSIM"] - 1.0) <= 1e-6
    b = (a + 0.1).clamp(0, 1)
    m = P.calculate_metrics(a, b, "val")
    assert abs(m["val/PSNR"] - 10 * np.log10(1.0 / m["val/L2"])) <= 1e-4 and 0.0 < m["val/SSIM"] < 1.0
    assert abs(float(P.gaussian_kernel1d(5).sum()) - 1.0) <= 1e-6
    # symmetric in its arguments, and a constant shift of both images leaves the structure term unchanged
    assert abs(float(P.ssim_map(a, b, 5).mean()) - float(P.ssim_map(b, a, 5).mean())) <= 1e-6
