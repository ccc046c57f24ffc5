#!/usr/bin/env python
"""Recipe for ``oracle/_ref``: the reference's OWN hot-path modules, taken from where they lie under /root/reference.

The reference is pure Python, so "building" it means placing the three modules whose arithmetic is the hot path --
``model/networks.py`` (define_G / define_D / ResnetGenerator / NLayerDiscriminator / GANLoss), ``model/generator_inject.py``
(the SatCLIP-injected generator) and ``utils/remote_sensing_indices.py`` -- in an importable tree.  ``oracle/_ref/`` is
git-ignored (no reference source enters the repository's history) but travels to the GPU box with the snapshot, so
``bench.py --impl reference`` and the ``cpu_baseline`` leg can time the UNMODIFIED reference modules on the box's host
cores (``cpu_baseline.kind = "reference"``).  Run by ``__graft_entry__.build()`` whenever /root/reference exists; on the
GPU box the prebuilt tree is used as is.  TEST / MEASUREMENT INFRASTRUCTURE, never imported by the product package.
"""
import os
import shutil
import sys

REF = os.environ.get("NIRGAN_REFERENCE_ROOT", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
FILES = ["model/networks.py", "model/generator_inject.py", "utils/remote_sensing_indices.py",
         "configs/config_px2px_SatCLIP.yaml", "configs/config_px2px.yaml"]


def make() -> bool:
    if not os.path.isdir(REF):
        return os.path.isdir(DST)
    for rel in FILES:
        src, dst = os.path.join(REF, rel), os.path.join(DST, rel)
        if not os.path.exists(src):
            continue
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
    with open(os.path.join(DST, "PROVENANCE.txt"), "w") as f:
        f.write("Unmodified copies of simon-donike/NIR-GAN files (made by oracle/make_ref.py from %s):\n%s\n" % (
            REF, "\n".join(FILES)))
    return True


def load():
    """Import the reference modules from oracle/_ref (None when the tree has not been made)."""
    if not os.path.exists(os.path.join(DST, "model", "networks.py")):
        return None
    if DST not in sys.path:
        sys.path.insert(0, DST)
    import importlib
    nets = importlib.import_module("model.networks")
    inj = importlib.import_module("model.generator_inject")
    return nets, inj


if __name__ == "__main__":
    print("oracle/_ref ready" if make() else "no reference tree at %s" % REF)
