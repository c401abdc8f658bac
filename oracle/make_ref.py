#!/usr/bin/env python
"""Copy recipe for oracle/_ref — the UNMODIFIED reference package, shipped beside the tests.  TEST INFRASTRUCTURE ONLY.

The reference (EPFLiGHT/MultiModN) is pure Python; ``/root/reference`` exists only in the build container, not on the GPU
box.  This script copies its ``multimodn/`` package (and ``datasets/multimod_dataset.py``, the only dataset module that
imports) byte for byte into ``oracle/_ref/`` — git-ignored, so no reference source enters the history, but NOT
gpurun-ignored, so the copy travels to the GPU box with the snapshot (SURVEY.md section 8c "How the oracle ships").
``__graft_entry__.build()`` runs it whenever ``/root/reference`` is present.

Users (only tests/, smoke() and bench.py's CPU arms, through oracle/ref_live.py):
  * ``ref_asis`` — the reference as shipped, timed on the GPU box's host cores beside the vectorised port;
  * live full-size parity — the CUDA path against the reference itself at B = 65 536 (tests/test_gpu_ref_live.py).

    python oracle/make_ref.py            # (re)create oracle/_ref from /root/reference
"""
import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("MMN_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(HERE, "_ref")
COPY = [("multimodn", "multimodn"), (os.path.join("datasets", "multimod_dataset.py"), os.path.join("datasets", "multimod_dataset.py")),
        (os.path.join("datasets", "__init__.py"), os.path.join("datasets", "__init__.py"))]


def make_ref(quiet=False):
    if not os.path.isdir(os.path.join(SRC, "multimodn")):
        if not quiet:
            print(f"make_ref: {SRC}/multimodn not found (this is expected on the GPU box); nothing copied")
        return False
    digest = hashlib.sha256()
    for rel_src, rel_dst in COPY:
        s, d = os.path.join(SRC, rel_src), os.path.join(DST, rel_dst)
        if not os.path.exists(s):
            continue
        if os.path.isdir(s):
            if os.path.isdir(d):
                shutil.rmtree(d)
            shutil.copytree(s, d, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
            for root, _, files in sorted(os.walk(d)):
                for f in sorted(files):
                    with open(os.path.join(root, f), "rb") as fh:
                        digest.update(fh.read())
        else:
            os.makedirs(os.path.dirname(d), exist_ok=True)
            shutil.copyfile(s, d)
            with open(d, "rb") as fh:
                digest.update(fh.read())
    with open(os.path.join(DST, "SOURCE.txt"), "w") as fh:
        fh.write(f"verbatim copy of {SRC} (multimodn/, datasets/multimod_dataset.py) made by oracle/make_ref.py\n"
                 f"sha256 of the copied files: {digest.hexdigest()}\n")
    if not quiet:
        print(f"make_ref: copied the reference package into {DST}")
    return True


if __name__ == "__main__":
    sys.exit(0 if make_ref() else 1)
