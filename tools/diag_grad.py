#!/usr/bin/env python
"""Diagnostic: per-unit forward (pre-norm y) and backward (dL/dy) parity of the fp16 tensor-core generator plans against
the fp32 autograd of the same function with the fp16 storage points emulated (oracle.storage_rounding), random-probe
objective sum(w * pred).  Prints one line per unit; -> gpurun_out/diag_grad.json."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402
import nirgan_oracle as O  # noqa: E402
import nirgan_b200  # noqa: F401,E402
from nirgan_b200.model import networks  # noqa: E402


def oracle_units(sd, x, wrap):
    """resnet_generator_forward restated unit by unit (same storage points), keeping every pre-norm tensor."""
    st, inorm, rpad = O._st, O._inorm, O._rpad
    L = O.generator_layout(9)
    w = lambda i: (st(sd[f"model.{i}.weight"]), sd[f"model.{i}.bias"])
    ys = {}

    def keep(name, t):
        t.retain_grad()
        ys[name] = t
        return t

    x = F.pad(x, (wrap,) * 4, mode="reflect") if wrap else x
    h = keep("stem", st(F.conv2d(rpad(st(x), 3), *w(L["stem"]))))
    h = st(F.relu(inorm(h)))
    h = keep("d1", st(F.conv2d(h, *w(L["down1"]), stride=2, padding=1)))
    h = st(F.relu(inorm(h)))
    h = keep("d2", st(F.conv2d(h, *w(L["down2"]), stride=2, padding=1)))
    h = st(F.relu(inorm(h)))
    for b in range(9):
        i = L["block0"] + b
        r = keep(f"r{b}a", st(F.conv2d(rpad(h, 1), st(sd[f"model.{i}.conv_block.1.weight"]), sd[f"model.{i}.conv_block.1.bias"])))
        r = st(F.relu(inorm(r)))
        r = keep(f"r{b}b", st(F.conv2d(rpad(r, 1), st(sd[f"model.{i}.conv_block.5.weight"]), sd[f"model.{i}.conv_block.5.bias"])))
        h = st(h + inorm(r))
    for key, nm in (("up1", "u1"), ("up2", "u2")):
        h = keep(nm, st(F.conv_transpose2d(h, *w(L[key]), stride=2, padding=1, output_padding=1)))
        h = st(F.relu(inorm(h)))
    y = torch.tanh(F.conv2d(rpad(h, 3), *w(L["head"])))
    if wrap:
        y = y[..., wrap:-wrap, wrap:-wrap]
    return y, ys


def main():
    B, S, wrap = 2, 64, 10
    dev = torch.device("cuda:0")
    sd = O.random_state_dict(O.generator_param_shapes(), seed=61)
    g = torch.Generator().manual_seed(11)
    x = torch.rand(B, 3, S, S, generator=g)
    probe = torch.randn(B, 1, S, S, generator=g)
    res = {}
    for mode in ("storage", "fp32"):
        pg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
        if mode == "storage":
            with O.storage_rounding(torch.float16):
                y, ys = oracle_units(pg, x, wrap)
                (y * probe).sum().backward()
        else:
            y, ys = oracle_units(pg, x, wrap)
            (y * probe).sum().backward()
        res[mode] = (y.detach(), {k: (t.detach(), t.grad.detach()) for k, t in ys.items()}, {k: v.grad for k, v in pg.items()})
    net = networks.define_G(3, 1, 64, "resnet_9blocks", "instance", False, "normal", 0.02)
    net.load_state_dict(sd)
    net = net.to(dev).train().configure_b200(precision="fp16", impl="tc")
    out = net(x.to(dev), wrap_pad=wrap)
    (out * probe.to(dev)).sum().backward()
    torch.cuda.synchronize()
    runner = net._runner
    ctx = list(runner._train.values())[0]
    graph, bwd = ctx["graph"], ctx["bwd"]
    gs = bwd.records.get("gscale")
    inv = float(gs[1]) if gs is not None else 1.0
    rows = []

    def rel(a, b):
        return float((a - b).norm() / (b.norm() + 1e-30))

    # the plan's ng_in_bwd calls hold the pointers of every unit's pre-norm tensor y and of its gradient dy
    ybuf, dybuf = {}, {}
    for (fn, args, name), label in zip(bwd.ops, bwd.labels):
        if name == "ng_in_bwd":
            unit = label.split(".")[-2]
            ybuf[unit] = (args[4], args[6:10])          # y ptr, (B, H, W, C)
            dybuf[unit] = args[17]
    # simpler: the pool is one uint8 tensor; pointers are offsets into it
    pool_t = ctx["pool"].t
    base = pool_t.data_ptr()

    def view(ptr, shape):
        n = 1
        for s_ in shape:
            n *= s_
        off = ptr - base
        return pool_t[off:off + 2 * n].view(torch.float16).view(*shape)

    print(f"pred: ours vs storage-oracle max {float((out.detach().cpu() - res['storage'][0]).abs().max()):.3e} "
          f"mean {float((out.detach().cpu() - res['storage'][0]).abs().mean()):.3e} | ours vs fp32-oracle mean "
          f"{float((out.detach().cpu() - res['fp32'][0]).abs().mean()):.3e} | storage vs fp32 mean "
          f"{float((res['storage'][0] - res['fp32'][0]).abs().mean()):.3e}")
    order = ["u2", "u1"] + [f"r{b}{h}" for b in range(8, -1, -1) for h in ("b", "a")] + ["d2", "d1", "stem"]
    for unit in order:
        if unit not in ybuf:
            continue
        yptr, (b_, h_, w_, c_) = ybuf[unit]
        y_ours = view(yptr, (b_, h_, w_, c_)).permute(0, 3, 1, 2).float().cpu()
        dy_ours = view(dybuf[unit], (b_, h_, w_, c_)).permute(0, 3, 1, 2).float().cpu() * inv
        ys_, dys_ = res["storage"][1][unit]
        yf_, dyf_ = res["fp32"][1][unit]
        row = {"unit": unit, "y_vs_storage": rel(y_ours, ys_), "y_vs_fp32": rel(y_ours, yf_), "y_storage_vs_fp32": rel(ys_, yf_),
               "dy_vs_storage": rel(dy_ours, dys_), "dy_vs_fp32": rel(dy_ours, dyf_), "dy_storage_vs_fp32": rel(dys_, dyf_)}
        rows.append(row)
        print({k: (round(v, 5) if isinstance(v, float) else v) for k, v in row.items()})
    out_dir = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out_dir):
        json.dump(rows, open(os.path.join(out_dir, "diag_grad.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
