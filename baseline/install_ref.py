"""Puts the UNMODIFIED reference package under baseline/_ref/ (git-ignored, shipped to the GPU box by gpurun).

    python baseline/install_ref.py [--source /root/reference]

1. tries the contract's offline install  `pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse
   --target baseline/_ref <source>`;  the reference's build backend is poetry-core, which is not in this image, so this fails here
   (recorded in baseline/_ref/INSTALL.json and DESIGN.md);
2. falls back to what that install would have produced for a pure-Python package: a byte-for-byte copy of the `velora/` package
   directory (Python sources only; no file is edited).

Nothing under baseline/_ref/ is tracked.  It is used by `bench.py --impl reference` (the reference's own modules timed on the host
cores, and launched eagerly on the GPU as the incumbent) and by tests/test_dropin_gpu.py (the reference's own
ActorModule / CriticModule / NeuroFlowCT._train_step driving the B200 kernels).  gymnasium / sqlmodel are not installed; the
loader (oracle/ref_import.py) registers empty stand-ins for them, the reference files themselves run unmodified.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
TARGET = os.path.join(HERE, "_ref")


def tree_digest(root: str) -> str:
    h = hashlib.sha256()
    for d, _, files in sorted(os.walk(root)):
        for f in sorted(files):
            if f.endswith(".py"):
                p = os.path.join(d, f)
                h.update(os.path.relpath(p, root).encode())
                with open(p, "rb") as fh:
                    h.update(fh.read())
    return h.hexdigest()


def install(source: str = "/root/reference", quiet: bool = False) -> dict:
    src_pkg = os.path.join(source, "velora")
    if not os.path.isdir(src_pkg):
        raise SystemExit(f"no reference package at {src_pkg}")
    os.makedirs(TARGET, exist_ok=True)
    info = {"source": source, "method": None, "pip_error": None}
    cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--find-links", "/opt/wheelhouse",
           "--target", TARGET, source]
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
        if r.returncode == 0 and os.path.isfile(os.path.join(TARGET, "velora", "wiring.py")):
            info["method"] = "pip --target"
        else:
            tail = (r.stderr or r.stdout).strip().splitlines()
            info["pip_error"] = tail[-1] if tail else f"exit {r.returncode}"
    except Exception as e:  # pragma: no cover
        info["pip_error"] = repr(e)
    if info["method"] is None:
        dst = os.path.join(TARGET, "velora")
        if os.path.isdir(dst):
            shutil.rmtree(dst)
        shutil.copytree(src_pkg, dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
        info["method"] = "copy of the package directory (pip build backend poetry-core unavailable offline)"
    info["sha256_py_sources"] = tree_digest(os.path.join(TARGET, "velora"))
    info["matches_source"] = info["sha256_py_sources"] == tree_digest(src_pkg)
    with open(os.path.join(TARGET, "INSTALL.json"), "w") as f:
        json.dump(info, f, indent=1)
    if not quiet:
        print(json.dumps(info))
    return info


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--source", default="/root/reference")
    install(ap.parse_args().source)
