"""GPU parity of the full hot path through the reference-compatible Python API.

Tolerances are the north-star's (BASELINE.json): tanh-output max-abs <= 2e-2 and mean-abs <= 2e-3 in the
low-precision (tensor-core) mode, <= 1e-4 in the fp32 verification mode; derived NDVI (eps 1e-6, clipped to
[-1,1], utils/logging_helpers.py:161-166) mean-abs <= 5e-3.  Checkers: golden fixtures produced by the real
reference modules (tests/golden, see oracle/pin_against_reference.py) and the CPU oracle on seeded inputs.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

MODES = [pytest.param("fp32", "simt", 1e-4, 2e-5, id="fp32-verify"),
         pytest.param("fp16", "simt", 2e-2, 2e-3, id="fp16-simt"),
         pytest.param("fp16", "tc", 2e-2, 2e-3, id="fp16-tc"),
         # bf16 storage (8-bit mantissa) through 23 InstanceNorm layers measures 6-7e-2 / 1.1e-2 on these fixtures
         # (profiles/r1d_precision_modes.md): 8x the fp16 error, as the mantissa widths predict.  It cannot meet the
         # north-star's 2e-2 / 2e-3, which is why fp16 operands (same tcgen05 rate) are the default fast mode; the
         # bf16 kernels are kept for range-critical use and checked against their own measured envelope.
         pytest.param("bf16", "tc", 1.2e-1, 2e-2, id="bf16-tc")]


from nirgan_b200.config import satclip_inject_config as inject_config  # noqa: E402  (the reference's SatCLIP YAML)


def make_G(sd, precision, impl, inject=False):
    import nirgan_oracle as O  # noqa: F401
    from nirgan_b200.model import networks
    from nirgan_b200.model.generator_inject import define_G_inject
    if inject:
        net = define_G_inject(inject_config())
    else:
        net = networks.define_G(3, 1, 64, "resnet_9blocks", "instance", False, "normal", 0.02)
    net.load_state_dict(sd)
    net = net.cuda().eval()
    net.configure_b200(precision=precision, impl=impl)
    return net


def _check(got, ref, tol_max, tol_mean, what):
    d = (got.float().cpu() - ref).abs()
    assert torch.isfinite(got).all(), what
    assert float(d.max()) <= tol_max, f"{what}: max-abs {float(d.max()):.3e} > {tol_max}"
    assert float(d.mean()) <= tol_mean, f"{what}: mean-abs {float(d.mean()):.3e} > {tol_mean}"


def ndvi_display(nir, red):
    return ((nir - red) / (nir + red + 1e-6)).clamp(-1, 1)


@pytest.mark.parametrize("precision,impl,tmax,tmean", MODES)
def test_generator_plain_golden(golden_dir, precision, impl, tmax, tmean):
    import nirgan_oracle as O
    g = np.load(f"{golden_dir}/g_plain_64.npz")
    sd = O.random_state_dict(O.generator_param_shapes(), seed=int(g["sd_seed"]), bias_std=float(g["bias_std"]))
    net = make_G(sd, precision, impl)
    x = torch.from_numpy(g["x"])
    with torch.no_grad():
        y = net(x.cuda())
    _check(y, torch.from_numpy(g["y"]), tmax, tmean, "G plain 64 (reference golden)")
    # wrapper: reflect-pad 10 / crop 10 fused (pix2pix.py:88-110)
    gp = np.load(f"{golden_dir}/g_plain_64_pad10.npz")
    with torch.no_grad():
        yp = net(x.cuda(), wrap_pad=10)
    _check(yp, torch.from_numpy(gp["y"]), tmax, tmean, "G plain 64 pad10 (reference golden)")


@pytest.mark.parametrize("precision,impl,tmax,tmean", MODES)
def test_generator_config1_256(golden_dir, precision, impl, tmax, tmean):
    """BASELINE.json configs[0]: one 3x256x256 tile, batch 1."""
    import nirgan_oracle as O
    g = np.load(f"{golden_dir}/g_plain_256.npz")
    sd = O.random_state_dict(O.generator_param_shapes(), seed=int(g["sd_seed"]), bias_std=float(g["bias_std"]))
    net = make_G(sd, precision, impl)
    x = torch.rand(1, 3, 256, 256, generator=torch.Generator().manual_seed(int(g["x_seed"])))
    with torch.no_grad():
        y = net(x.cuda())
    ref = torch.from_numpy(g["y"])
    _check(y, ref, tmax, tmean, "G 256 (reference golden)")
    nd = (ndvi_display(y.cpu(), x[:, 0:1]) - ndvi_display(ref, x[:, 0:1])).abs().mean()
    assert float(nd) <= (5e-3 if precision != "bf16" else 5e-2), f"derived NDVI mean-abs {float(nd):.3e}"


@pytest.mark.parametrize("precision,impl,tmax,tmean", MODES)
@pytest.mark.parametrize("scale", [0.01, 1.0])
def test_generator_inject_golden(golden_dir, precision, impl, tmax, tmean, scale):
    import nirgan_oracle as O
    g = np.load(f"{golden_dir}/g_inject_64_s{scale}.npz")
    sd = O.random_state_dict(O.generator_param_shapes(inject=True), seed=int(g["sd_seed"]), bias_std=float(g["bias_std"]))
    sd["scale_param"] = torch.tensor(float(g["scale"]))
    net = make_G(sd, precision, impl, inject=True)
    with torch.no_grad():
        y = net(torch.from_numpy(g["x"]).cuda(), torch.from_numpy(g["embeds"]).cuda())
    _check(y, torch.from_numpy(g["y"]), tmax, tmean, f"G inject 64 scale {scale}")
    if scale == 1.0:   # 84/42/21 pyramid: bilinear 128 -> 42
        gp = np.load(f"{golden_dir}/g_inject_64_pad10_s1.0.npz")
        with torch.no_grad():
            yp = net(torch.from_numpy(gp["x"]).cuda(), torch.from_numpy(gp["embeds"]).cuda(), wrap_pad=10)
        _check(yp, torch.from_numpy(gp["y"]), tmax, tmean, "G inject 64 pad10")


@pytest.mark.parametrize("precision,impl,tmax,tmean", MODES)
def test_discriminator_golden(golden_dir, precision, impl, tmax, tmean):
    import nirgan_oracle as O
    from nirgan_b200.model import networks
    g = np.load(f"{golden_dir}/d_64.npz")
    sd = O.random_state_dict(O.discriminator_param_shapes(), seed=int(g["sd_seed"]), bias_std=float(g["bias_std"]))
    net = networks.define_D(4, 64, "basic", 3, "instance", "normal", 0.02)
    net.load_state_dict(sd)
    net = net.cuda().eval().configure_b200(precision=precision, impl=impl)
    with torch.no_grad():
        y = net(torch.from_numpy(g["x"]).cuda())
    assert tuple(y.shape) == (2, 1, 6, 6)
    _check(y, torch.from_numpy(g["y"]), tmax, tmean, "PatchGAN 64")


@pytest.mark.parametrize("precision,impl,tmax,tmean", MODES)
def test_generator_vs_oracle_batch_and_sizes(precision, impl, tmax, tmean):
    """Seeded inputs, oracle computed on the host cores: ragged batch / mixed resolutions (config 5 sizes)."""
    import nirgan_oracle as O
    sd = O.random_state_dict(O.generator_param_shapes(), seed=3)
    net = make_G(sd, precision, impl)
    for B, H in ((3, 128), (1, 192), (2, 96)):
        x = torch.rand(B, 3, H, H, generator=torch.Generator().manual_seed(H))
        with torch.no_grad():
            y = net(x.cuda(), wrap_pad=10)
            ref = O.px2px_forward(sd, x, 10)
        _check(y, ref, tmax, tmean, f"G {B}x{H} pad10 vs oracle")


def test_batch_position_invariance_is_bit_exact():
    """A tile's result does not depend on which batch / which slot it is computed in (tile-sharded inference
    must reproduce the sequential loop bit-for-bit)."""
    import nirgan_oracle as O
    sd = O.random_state_dict(O.generator_param_shapes(), seed=4)
    net = make_G(sd, "fp16", "tc")
    x = torch.rand(5, 3, 64, 64, generator=torch.Generator().manual_seed(9)).cuda()
    with torch.no_grad():
        full = net(x, wrap_pad=10)
        for i in range(5):
            one = net(x[i:i + 1], wrap_pad=10)
            assert torch.equal(one[0], full[i])
        pair = net(x[[3, 1]], wrap_pad=10)
        assert torch.equal(pair[0], full[3]) and torch.equal(pair[1], full[1])


def test_cpu_tensor_raises():
    from nirgan_b200.model import networks
    net = networks.define_G(3, 1, 64, "resnet_9blocks", "instance", False, "normal", 0.02)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        net(torch.rand(1, 3, 64, 64))


def test_tile_sharded_inference_matches_sequential_loop(golden_dir):
    """create_synthetic_dataset.py:100-118 loop: 3 emulated ranks' shards == the sequential loop bit-for-bit,
    ids in sorted order, and both within tolerance of the golden produced by the real reference loop."""
    import nirgan_oracle as O
    from nirgan_b200 import synth
    g = np.load(f"{golden_dir}/synth_loop_32.npz")
    sd = O.random_state_dict(O.generator_param_shapes(), seed=int(g["sd_seed"]), bias_std=float(g["bias_std"]))
    net = make_G(sd, "fp16", "tc")
    names = [f"tile_{i:06d}.tif" for i in (3, 0, 2, 1, 4)]
    tiles = {n: torch.rand(3, 32, 32, generator=torch.Generator().manual_seed(100 + int(n[5:11]))) for n in names}
    model = lambda hr: net(hr, wrap_pad=10)
    seq = synth.run_shard(model, tiles, 0, 1, batch_size=2, device=torch.device("cuda"))
    assert list(seq.keys()) == list(g["ids"])
    merged = {}
    for r in range(3):
        merged.update(synth.run_shard(model, tiles, r, 3, batch_size=64, device=torch.device("cuda")))
    assert sorted(merged.keys()) == list(seq.keys())
    for k in seq:
        assert torch.equal(merged[k], seq[k]), k
        _check(seq[k], torch.from_numpy(g["y." + k]), 2e-2, 2e-3, f"synth loop {k}")
    # the pipelined loop (pinned staging, H2D / generator / read-back on separate streams) gives the same dictionary
    piped = synth.run_shard(model, tiles, 0, 1, batch_size=2, device=torch.device("cuda"),
                            model_async=lambda hr, ready: net.forward_async(hr, None, 10, ready), in_flight=2)
    assert list(piped.keys()) == list(seq.keys())
    for k in seq:
        assert torch.equal(piped[k], seq[k]), k
    # ... also with the loop body's post-processing (nearest x4 + histogram matching + float16) on the device
    s2 = {n: torch.rand(1, 8, 8, generator=torch.Generator().manual_seed(7 + int(n[5:11]))) * 0.4 for n in names}
    a = synth.run_shard(model, tiles, 0, 1, batch_size=2, device=torch.device("cuda"), s2_nir=s2)
    b = synth.run_shard(model, tiles, 0, 1, batch_size=2, device=torch.device("cuda"), s2_nir=s2,
                        model_async=lambda hr, ready: net.forward_async(hr, None, 10, ready))
    for k in a:
        assert a[k].dtype == torch.float16 and torch.equal(a[k], b[k]), k


def test_config2_full_size_properties():
    """BASELINE.json configs[1] at its full size (64 injected 3x256x256 tiles, the two-stream / CUDA-graph path that
    bench.py times): size-independent properties -- finite, inside tanh's range, every tile bit-identical to the same
    tile computed alone, in another slot or with a different slice split (InstanceNorm is per sample) -- and two of the
    tiles against the CPU oracle within the north-star tolerance."""
    import nirgan_oracle as O
    sd = O.random_state_dict(O.generator_param_shapes(inject=True), seed=12, scale_param=1.0)
    net = make_G(sd, "fp16", "tc", inject=True)
    g = torch.Generator().manual_seed(64)
    x = torch.rand(64, 3, 256, 256, generator=g)
    e = torch.randn(64, 256, generator=g)
    xc, ec = x.cuda(), e.cuda()
    with torch.no_grad():
        y = net(xc, ec)                       # auto: two slices of 32 on two streams
        y2 = net(xc, ec)                      # second call replays the captured CUDA graphs
        assert torch.isfinite(y).all() and float(y.abs().max()) <= 1.0
        assert torch.equal(y, y2)
        for i in (0, 31, 32, 63):
            one = net(xc[i:i + 1], ec[i:i + 1])
            assert torch.equal(one[0], y[i]), i
        perm = torch.randperm(64, generator=g)
        yp = net(xc[perm.cuda()], ec[perm.cuda()])
        assert torch.equal(yp, y[perm.cuda()])
        net.configure_b200(precision="fp16", impl="tc", streams=1)
        y1 = net(xc, ec)                      # one slice of 64, one stream
        assert torch.equal(y1, y)
        ref = O.resnet_generator_forward(sd, x[[5, 40]], embeds=e[[5, 40]])
    _check(y[[5, 40]], ref, 2e-2, 2e-3, "config 2, tiles 5 and 40 of 64 vs oracle")


def test_forward_async_matches_forward_over_consecutive_calls():
    """The streaming call overlaps consecutive steps on the generator's own streams; over a run of calls with different
    inputs (two sizes, with and without a caller-supplied ready event) every result is bit-identical to the blocking call."""
    import nirgan_oracle as O  # noqa: F401
    torch.manual_seed(0)
    from nirgan_b200.model.generator_inject import define_G_inject
    net = define_G_inject(inject_config()).cuda().eval()
    g = torch.Generator().manual_seed(5)
    xs = [torch.rand(n, 3, 64, 64, generator=g).cuda() for n in (40, 40, 6, 40, 33)]
    es = [torch.randn(x.shape[0], 256, generator=g).cuda() for x in xs]
    with torch.no_grad():
        want = [net(x, e).clone() for x, e in zip(xs, es)]
        torch.cuda.synchronize()
        outs = []
        side = torch.cuda.Stream()
        for i, (x, e) in enumerate(zip(xs, es)):
            if i % 2:
                # inputs produced on another stream, handed over by event
                with torch.cuda.stream(side):
                    x2, e2 = x.clone(), e.clone()
                    ev = torch.cuda.Event()
                    ev.record(side)
                x2.record_stream(torch.cuda.current_stream())
                outs.append(net.forward_async(x2, e2, ready=ev))
            else:
                outs.append(net.forward_async(x, e))
        main = torch.cuda.current_stream()
        for (y, done), w in zip(outs, want):
            for ev in done:
                main.wait_event(ev)
            assert torch.equal(y, w)
        # inputs that need a dtype conversion, produced on another stream: the conversion must wait for `ready` too
        with torch.cuda.stream(side):
            xh = (xs[0] * 0 + xs[3]).double()
            ev = torch.cuda.Event()
            ev.record(side)
        xh.record_stream(main)
        y, done = net.forward_async(xh, es[3], ready=ev)
        for d in done:
            main.wait_event(d)
        assert torch.equal(y, want[3])


def test_config3_full_size_512px_tiles():
    """BASELINE.json configs[2] at its full size: create_synthetic_dataset-style inference of 3x512x512 tiles behind the
    pad-10 wrapper (532 / 266 / 133 pyramid, plain generator).  Tile-sharded over 3 emulated ranks == the sequential loop
    bit for bit with ids in sorted order; a tile computed alone or in a batch of 6 has the same bits; two tiles against the
    CPU oracle within the north-star tolerance (NIR max-abs 2e-2 / mean-abs 2e-3, derived NDVI mean-abs 5e-3)."""
    import nirgan_oracle as O
    from nirgan_b200 import synth
    sd = O.random_state_dict(O.generator_param_shapes(), seed=33)
    net = make_G(sd, "fp16", "tc")
    names = [f"tile_{i:06d}.tif" for i in (4, 1, 5, 0, 3, 2)]
    tiles = {n: torch.rand(3, 512, 512, generator=torch.Generator().manual_seed(500 + int(n[5:11]))) for n in names}
    model = lambda hr: net(hr, wrap_pad=10)
    dev = torch.device("cuda")
    seq = synth.run_shard(model, tiles, 0, 1, batch_size=2, device=dev)
    assert list(seq.keys()) == [f"tile_{i:06d}" for i in range(6)]
    merged = {}
    for r in range(3):
        merged.update(synth.run_shard(model, tiles, r, 3, batch_size=2, device=dev))
    big = synth.run_shard(model, tiles, 0, 1, batch_size=6, device=dev)
    for k in seq:
        assert seq[k].shape == (1, 512, 512) and torch.isfinite(seq[k]).all()
        assert torch.equal(merged[k], seq[k]) and torch.equal(big[k], seq[k]), k
    torch.set_num_threads(max(1, (__import__("os").cpu_count() or 1)))
    for n in names[:2]:
        k = synth.tile_id(n)
        with torch.no_grad():
            ref = O.px2px_forward(sd, tiles[n][None], 10, True)[0]
        _check(seq[k], ref, 2e-2, 2e-3, f"config 3 {k}")
        red = tiles[n][0:1]
        d = (ndvi_display(seq[k].float().cpu(), red) - ndvi_display(ref, red)).abs().mean()
        assert float(d) <= 5e-3, float(d)
