"""Import the UNMODIFIED reference package from oracle/_ref (or /root/reference when it exists).  TEST INFRASTRUCTURE
ONLY — same rules as multimodn_oracle.py: tests/, smoke() and bench.py's CPU arms may import this, the product never.

The reference imports torchmetrics / torchsummary / matplotlib, which are absent from the image; ``oracle/ref_shims``
stubs them (only the binary ConfusionMatrix does arithmetic).
"""
from __future__ import annotations

import importlib
import itertools
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def reference_root():
    """directory that holds the reference's ``multimodn`` package, or None"""
    for root in (os.path.join(HERE, "_ref"), os.environ.get("MMN_REFERENCE_ROOT", "/root/reference")):
        if root and os.path.isfile(os.path.join(root, "multimodn", "multimodn.py")):
            return root
    return None


_CACHE = None


def load_reference():
    """-> namespace with the reference's MultiModN, encoders, decoders, MultiModNHistory; None if it is not available"""
    global _CACHE
    if _CACHE is not None:
        return _CACHE
    root = reference_root()
    if root is None:
        return None
    import torch
    if not hasattr(torch._utils, "_accumulate"):
        torch._utils._accumulate = itertools.accumulate      # removed from modern torch (datasets/multimod_dataset.py:6)
    shims = os.path.join(HERE, "ref_shims")
    saved = list(sys.path)
    sys.path[:0] = [shims, root]
    try:
        mm = importlib.import_module("multimodn.multimodn")
        enc = importlib.import_module("multimodn.encoders")
        dec = importlib.import_module("multimodn.decoders")
        hist = importlib.import_module("multimodn.history")
    finally:
        sys.path[:] = saved

    class Ref:
        pass

    ref = Ref()
    ref.root = root
    ref.MultiModN = mm.MultiModN
    ref.MLPEncoder, ref.MIMIC_MLPEncoder = enc.MLPEncoder, enc.MIMIC_MLPEncoder
    ref.MLPDecoder, ref.LogisticDecoder, ref.ClassDecoder = dec.MLPDecoder, dec.LogisticDecoder, dec.ClassDecoder
    ref.MultiModNHistory = hist.MultiModNHistory
    _CACHE = ref
    return ref
