"""GPU parity of the post-processing kernels (csrc/postprocess.cu) against the CPU oracle
(oracle/postprocess_oracle.py = create_synthetic_dataset.py:34-52,111-118 with skimage's algorithm restated)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _ws(nbytes):
    return torch.empty(nbytes // 4 + 1, dtype=torch.int32, device="cuda")


@pytest.mark.parametrize("segs,n", [(1, 1), (3, 31), (2, 4096), (5, 4097), (2, 65536), (1, 262144), (7, 10000)])
def test_segmented_radix_sort_matches_torch_sort(segs, n):
    from nirgan_b200 import _lib as L
    g = torch.Generator(device="cuda").manual_seed(segs * 1000 + n)
    x = torch.randn(segs, n, generator=g, device="cuda")
    x[:, ::7] = x[:, ::7].round()                    # ties
    x[0, :3] = torch.tensor([0.0, -0.0, float("inf")], device="cuda")[: min(3, n)]
    need = segs * n * 8 + segs * ((n + 4095) // 4096) * 1024
    ws = _ws(need)
    out = torch.full_like(x, float("nan"))
    L.call("ng_sort_segments", x.data_ptr(), segs, n, out.data_ptr(), ws.data_ptr(), need,
           torch.cuda.current_stream().cuda_stream)
    ref = torch.sort(x, dim=1).values
    assert torch.equal(out + 0.0, ref + 0.0)           # +0.0 folds the sign of zero


@pytest.mark.parametrize("mode", ["nearest", "bilinear"])
@pytest.mark.parametrize("shape,size", [((2, 1, 16, 16), (64, 64)), ((1, 3, 128, 128), (138, 138)), ((2, 1, 20, 28), (20, 28)),
                                        ((1, 1, 64, 64), (37, 53))])
def test_resize_matches_interpolate(shape, size, mode):
    from nirgan_b200 import postprocess as PP
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(*shape, generator=g, device="cuda")
    got = PP.resize(x, size, mode)
    ref = F.interpolate(x, size=size, mode=mode, **({"align_corners": False} if mode == "bilinear" else {}))
    assert float((got - ref).abs().max()) <= (0.0 if mode == "nearest" else 2e-6)


def _case(kind, B, H, seed):
    g = torch.Generator().manual_seed(seed)
    if kind == "continuous":
        img = torch.tanh(torch.randn(B, 1, H, H, generator=g))
        ref = torch.rand(B, 1, H, H, generator=g) * 0.4
    elif kind == "ties":                                # quantised values on both sides, constant rows
        img = (torch.randn(B, 1, H, H, generator=g) * 4).round() / 4
        ref = (torch.rand(B, 1, H, H, generator=g) * 16).floor() / 40
        img[0, 0, 0] = 0.0
        img[0, 0, 1] = -0.0
    elif kind == "upsampled":                           # the real use: nearest x4 of a 4x coarser Sentinel-2 band
        img = torch.tanh(torch.randn(B, 1, H, H, generator=g))
        ref = F.interpolate(torch.rand(B, 1, H // 4, H // 4, generator=g) * 0.35, scale_factor=4)
    else:                                               # constant image / constant reference
        img = torch.full((B, 1, H, H), 0.25)
        ref = torch.rand(B, 1, H, H, generator=g)
        ref[-1] = 0.5
    return img, ref


@pytest.mark.parametrize("kind", ["continuous", "ties", "upsampled", "constant"])
@pytest.mark.parametrize("B,H", [(2, 32), (3, 64), (1, 256)])
def test_histogram_match_matches_oracle(kind, B, H):
    import postprocess_oracle as P
    from nirgan_b200 import postprocess as PP
    img, ref = _case(kind, B, H, seed=B * 100 + H)
    want = P.histogram_match(img, ref)
    got = PP.histogram_match(img.cuda(), ref.cuda()).cpu()
    assert got.shape == want.shape and got.dtype == torch.float32
    # the oracle interpolates in float64 and casts to float32; the kernel does the same arithmetic
    assert float((got - want).abs().max()) <= 1e-6 * max(1.0, float(want.abs().max()))


def test_postprocess_pipeline_fp16_and_reference_resize(tmp_path):
    """nearest x4 + matching + float16 (create_synthetic_dataset.py:111-116) and the .npz writer (:49-52)."""
    import postprocess_oracle as P
    from nirgan_b200 import postprocess as PP
    g = torch.Generator().manual_seed(11)
    pred = torch.tanh(torch.randn(4, 1, 128, 128, generator=g))
    s2 = torch.rand(4, 1, 32, 32, generator=g) * 0.3
    want = P.postprocess(pred, s2)
    got = PP.postprocess(pred.cuda(), s2.cuda())
    assert got.dtype == torch.float16 and got.is_cuda
    assert float((got.float().cpu() - want.float()).abs().max()) <= 2.5e-4          # one fp16 ulp at 0.3
    # a reference that is not at the image size goes through the bilinear resize first, like the reference code
    want2 = P.histogram_match(pred, s2)
    got2 = PP.histogram_match(pred.cuda(), s2.cuda()).cpu()
    assert float((got2 - want2).abs().max()) <= 2e-6
    names = [f"tile_{i:06d}" for i in range(4)]
    files = PP.save_images(got, names, str(tmp_path))
    for f, n, w in zip(files, names, got.cpu()):
        z = np.load(f)
        assert f.endswith(n + ".npz") and z["nir"].dtype == np.float16 and z["nir"].shape == (1, 128, 128)
        assert np.array_equal(z["nir"], w.numpy())


def test_full_size_properties_512():
    """BASELINE size (512x512 tiles): size-independent properties -- monotone in the input, output range = reference
    range, the reference's quantiles are transferred, matching a tile to itself is the identity."""
    from nirgan_b200 import postprocess as PP
    g = torch.Generator(device="cuda").manual_seed(5)
    img = torch.tanh(torch.randn(8, 1, 512, 512, generator=g, device="cuda"))
    ref = torch.rand(8, 1, 512, 512, generator=g, device="cuda") ** 2 * 0.4
    out = PP.histogram_match(img, ref)
    order = torch.argsort(img.view(8, -1), dim=1, stable=True)
    srt = torch.gather(out.view(8, -1), 1, order)
    assert bool((srt[:, 1:] >= srt[:, :-1]).all())
    assert torch.equal(out.amax(dim=(1, 2, 3)), ref.amax(dim=(1, 2, 3)))
    assert bool((out.amin(dim=(1, 2, 3)) >= ref.amin(dim=(1, 2, 3))).all())
    qs = torch.tensor([0.1, 0.5, 0.9], device="cuda")
    assert float((torch.quantile(out.view(8, -1), qs, dim=1) - torch.quantile(ref.view(8, -1), qs, dim=1)).abs().max()) <= 1e-3
    assert torch.equal(PP.histogram_match(img, img), img)


@pytest.mark.parametrize("shape", [(2, 1, 64, 64), (3, 1, 37, 53), (1, 3, 256, 256), (4, 1, 8, 40)])
def test_calculate_metrics_matches_oracle(shape):
    """utils/calculate_metrics.py:6-37 on the device vs the torch restatement of kornia's psnr / ssim."""
    import postprocess_oracle as P
    from nirgan_b200.utils.calculate_metrics import calculate_metrics, image_metrics
    g = torch.Generator().manual_seed(sum(shape))
    target = torch.rand(*shape, generator=g)
    pred = (target + 0.08 * torch.randn(*shape, generator=g)).clamp(0, 1)
    want = P.calculate_metrics(pred, target, "val")
    got = calculate_metrics(pred.cuda(), target.cuda(), "val")
    assert set(got) == set(want) == {"val/L1", "val/L2", "val/PSNR", "val/SSIM"}
    for k in want:
        assert abs(got[k] - want[k]) <= 2e-5 * max(1.0, abs(want[k])), (k, got[k], want[k])
    # window 11 (utils/losses.py::ssim_loss) where the image allows it
    if min(shape[-2:]) > 5:
        s11 = float(image_metrics(pred.cuda(), target.cuda(), window_size=11)[3])
        assert abs(s11 - float(P.ssim_map(pred, target, 11).mean())) <= 2e-5
    same = calculate_metrics(target.cuda(), target.cuda())
    assert same["train/L1"] == 0.0 and same["train/PSNR"] == float("inf") and abs(same["train/SSIM"] - 1.0) <= 1e-6
