#!/usr/bin/env python
"""Summarise ncu output into profiles/ (tracked):
  launches CSV  (ncu --metrics gpu__time_duration.sum --csv)  -> per-kernel launch count / total time / share
  .ncu-rep      (ncu --set full)                              -> per-kernel duration, DRAM bytes, DRAM %, tensor-pipe %,
                                                                 registers, achieved occupancy
Usage: python tools/summarize_ncu.py --launches gpurun_out/launches_r1c.csv --rep gpurun_out/prof_inf_r1c.ncu-rep \
           --tag r1c_inference --cmd "python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
"""
import argparse
import csv
import io
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def short(name: str) -> str:
    name = re.sub(r"^void\s+", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name.replace("ng::", "")


def launches_table(path):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.reader(io.StringIO("".join(lines)))
    hdr = None
    for r in rd:
        if hdr is None:
            if "Kernel Name" in r:
                hdr = r
            continue
        if len(r) != len(hdr):
            continue
        d = dict(zip(hdr, r))
        if d.get("Metric Name") != "gpu__time_duration.sum":
            continue
        val = float(d["Metric Value"].replace(",", ""))
        unit = d.get("Metric Unit", "ns")
        us = val / 1e3 if unit in ("ns", "nsecond") else (val if unit in ("us", "usecond") else val * 1e3)
        rows.append((short(d["Kernel Name"]), us))
    agg = {}
    for k, us in rows:
        c, t = agg.get(k, (0, 0.0))
        agg[k] = (c + 1, t + us)
    tot = sum(t for _, t in agg.values())
    out = ["| kernel | launches | total us | share |", "|---|---|---|---|"]
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| `{k}` | {c} | {t:.1f} | {100 * t / tot:.1f}% |")
    return "\n".join(out), len(rows), tot


WANT = [
    ("gpu__time_duration.sum", "us"),
    ("dram__bytes_read.sum", "MB rd"),
    ("dram__bytes_write.sum", "MB wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor inst"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ %"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
]


def rep_table(path):
    if path.endswith(".csv"):
        raw = open(path).read()
    else:
        raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(io.StringIO(raw)))
    hdr, units = rd[0], rd[1]
    col = {h: i for i, h in enumerate(hdr)}
    kcol = col["Kernel Name"]
    have = [(m, lab) for m, lab in WANT if m in col]
    out = ["| kernel | n | " + " | ".join(lab for _, lab in have) + " |", "|---|---|" + "---|" * len(have)]
    groups = {}
    for r in rd[2:]:
        if len(r) != len(hdr):
            continue
        key = (short(r[kcol]), r[col["launch__grid_size"]] if "launch__grid_size" in col else "")
        groups.setdefault(key, []).append(r)
    for (k, _), rs in sorted(groups.items(), key=lambda kv: -sum(float(x[col["gpu__time_duration.sum"]].replace(",", "")) for x in kv[1])):
        cells = []
        for m, lab in have:
            vals = []
            for r in rs:
                try:
                    vals.append(float(r[col[m]].replace(",", "")))
                except ValueError:
                    pass
            v = sum(vals) / len(vals) if vals else float("nan")
            u = units[col[m]]
            if m.startswith("dram__bytes"):
                v = v / 1e6 if u == "byte" else (v / 1e3 if u == "Kbyte" else (v * 1e3 if u == "Gbyte" else v))
            if m == "gpu__time_duration.sum":
                v = v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u in ("ms", "msecond") else v)
            cells.append(f"{v:.1f}" if abs(v) < 1e6 else f"{v:.3g}")
        out.append(f"| `{k}` | {len(rs)} | " + " | ".join(cells) + " |")
    return "\n".join(out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--launches")
    ap.add_argument("--rep", action="append", help=".ncu-rep or its `--page raw --csv` export; may be repeated")
    ap.add_argument("--tag", required=True)
    ap.add_argument("--cmd", default="")
    ap.add_argument("--note", default="")
    a = ap.parse_args()
    md = [f"# ncu summary `{a.tag}`", ""]
    if a.cmd:
        md += [f"Command: `{a.cmd}`", ""]
    if a.note:
        md += [a.note, ""]
    if a.launches:
        t, n, tot = launches_table(a.launches)
        md += [f"## Launch list ({n} launches, {tot / 1e3:.2f} ms of kernel time; cold-cache, serialised: compare SHARES)",
               "", f"`ncu --metrics gpu__time_duration.sum --clock-control none` -> `{os.path.basename(a.launches)}`", "", t, ""]
    for rep in a.rep or []:
        md += [f"## `ncu --set full --clock-control none` ({os.path.basename(rep)}), mean per launch, grouped by (kernel, grid)",
               "", rep_table(rep), ""]
    path = os.path.join(ROOT, "profiles", a.tag + ".md")
    with open(path, "w") as f:
        f.write("\n".join(md))
    print("wrote", path)


if __name__ == "__main__":
    main()
