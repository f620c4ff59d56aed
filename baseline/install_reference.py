#!/usr/bin/env python
"""Install the UNMODIFIED reference (ttaoREtw/semi-tts) into baseline/_ref/ so that it travels to the GPU box.

The reference is a plain source tree without setup.py / pyproject.toml, so `pip install --target baseline/_ref
/root/reference` has nothing to build ("neither 'setup.py' nor 'pyproject.toml' found", recorded in DESIGN.md); the
equivalent of an install is a verbatim copy of the importable tree: src/, bin/, corpus/, lib/, config/, main.py and the two
data files the quantizer reads (data/phn_attr.csv, data/cmu_phn.vocab).  baseline/_ref/ is git-ignored (never part of the
repo's history) but not gpurun-ignored.  Run here, in the build container:  python baseline/install_reference.py
Used by: bench.py --impl reference (times the reference's own modules, cpu_baseline.kind = "reference"),
tests/test_gpu_reference_model.py (BASELINE config 4: the drop-in inside the reference VQVAE on a GPU).
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
SRC = os.environ.get("VQB_REFERENCE_SRC", "/root/reference")
KEEP = ["src", "bin", "corpus", "lib", "config", "main.py", "LICENSE", os.path.join("data", "phn_attr.csv"),
        os.path.join("data", "cmu_phn.vocab")]


def install(quiet=False):
    if not os.path.isfile(os.path.join(SRC, "src", "embed.py")):
        if not quiet:
            print("reference tree not present at %s: nothing installed" % SRC)
        return False
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(DST)
    for rel in KEEP:
        s, d = os.path.join(SRC, rel), os.path.join(DST, rel)
        if os.path.isdir(s):
            shutil.copytree(s, d, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
        elif os.path.isfile(s):
            os.makedirs(os.path.dirname(d), exist_ok=True)
            shutil.copy2(s, d)
    with open(os.path.join(DST, "INSTALLED_FROM"), "w") as f:
        f.write("verbatim copy of %s (no file modified); see baseline/install_reference.py\n" % SRC)
    if not quiet:
        print("installed the reference into", DST)
    return True


if __name__ == "__main__":
    sys.exit(0 if install() else 1)
