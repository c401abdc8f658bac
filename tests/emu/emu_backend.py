"""Loads tests/emu/libmmn_emu.so — the kernel sources compiled for the host against cuda_emu.h —
behind the same ctypes surface as the real library.  TEST INFRASTRUCTURE ONLY: lets the CPU-only
build container exercise the kernels' indexing and the host logic end to end.  It is injected by
the ``emu`` fixture (tests/conftest.py); nothing in multimodn_b200 refers to it."""
import os
import subprocess

from multimodn_b200 import _lib

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SO = os.path.join(HERE, "libmmn_emu.so")
import glob
SOURCES = glob.glob(os.path.join(ROOT, "multimodn_b200", "csrc", "*.cu*"))
SOURCES += [os.path.join(ROOT, "include", "mmn.h"), os.path.join(HERE, "cuda_emu.h")]


def build_if_stale():
    newest = max(os.path.getmtime(p) for p in SOURCES)
    if not os.path.exists(SO) or os.path.getmtime(SO) < newest:
        subprocess.run(["sh", os.path.join(HERE, "build_emu.sh")], check=True)
    return SO


class EmuLibrary(_lib.Library):
    host_memory = True

    def __init__(self):
        super().__init__(build_if_stale())


_EMU = None


def get_emu_lib():
    global _EMU
    if _EMU is None:
        _EMU = EmuLibrary()
    return _EMU
