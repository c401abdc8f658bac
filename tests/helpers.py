"""Shared test helpers: golden-fixture loading and tolerant comparisons."""
import os

import numpy as np

from oracle.spec_io import spec_from_arrays

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

RTOL_FP32 = 1e-5      # BASELINE.json north_star: 1e-5 relative in fp32


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    return {k: z[k] for k in z.files}


def golden_data(fx, prefix="x"):
    out = []
    while f"{prefix}{len(out)}" in fx:
        out.append(fx[f"{prefix}{len(out)}"])
    return out


def golden_spec(fx, prefix="spec0"):
    return spec_from_arrays(fx, prefix)


def golden_grads(fx, spec, prefix="grad0"):
    g = dict(init_state=fx[f"{prefix}_init"], encoders=[], decoders=[])
    for e, enc in enumerate(spec["encoders"]):
        g["encoders"].append([(fx[f"{prefix}_e{e}_W{j}"], fx[f"{prefix}_e{e}_b{j}"]) for j in range(len(enc["layers"]))])
    for d, dec in enumerate(spec["decoders"]):
        g["decoders"].append([(fx[f"{prefix}_d{d}_W{j}"], fx[f"{prefix}_d{d}_b{j}"]) for j in range(len(dec["layers"]))])
    return g


def flat_grads(g):
    parts = [np.asarray(g["init_state"]).ravel()]
    for group in (g["encoders"], g["decoders"]):
        for layers in group:
            for W, b in layers:
                parts += [np.asarray(W).ravel(), np.asarray(b).ravel()]
    return np.concatenate(parts)


def flat_params(spec):
    parts = [np.asarray(spec["init_state"]).ravel()]
    for group in (spec["encoders"], spec["decoders"]):
        for m in group:
            for W, b in m["layers"]:
                parts += [W.ravel(), b.ravel()]
    return np.concatenate(parts)


def assert_close(actual, expected, rtol=RTOL_FP32, atol_scale=1.0, what=""):
    """relative to the magnitude of the expected ARRAY (norm-wise): the right yardstick for
    fp32 sums whose individual elements may cancel to ~0."""
    actual = np.asarray(actual, dtype=np.float64)
    expected = np.asarray(expected, dtype=np.float64)
    assert actual.shape == expected.shape, (what, actual.shape, expected.shape)
    nan_a, nan_e = np.isnan(actual), np.isnan(expected)
    assert (nan_a == nan_e).all(), f"{what}: NaN pattern differs"
    if nan_e.all():
        return
    scale = max(np.abs(expected[~nan_e]).max(), 1e-30)
    err = np.abs(actual[~nan_e] - expected[~nan_e]).max() / scale
    assert err <= rtol * atol_scale, f"{what}: max err / max|expected| = {err:.3e} > {rtol * atol_scale:.1e}"
