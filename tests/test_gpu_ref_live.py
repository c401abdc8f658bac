"""GPU: the CUDA path against the UNMODIFIED reference executed live (oracle/_ref, see oracle/make_ref.py) at the FULL
batch of BASELINE config 2 — B = 65 536 rows, MIMIC-shaped model, NaN-free so that the reference's batch-level skip rule
and the per-row select coincide — instead of only the B = 16 golden fixtures.  The reference runs on the host CPU (its
three CPython element loops make one step cost about a minute at this size); dropout is 0 on both sides because torch's
Philox stream cannot be reproduced outside torch.

Checks: every parameter gradient (norm-wise 1e-5 and per tensor), the five (E+1) x D history matrices, the state-change
vector, and `predict` bit for bit (up to arg-max ties at fp32 round-off)."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F
from torch.nn import CrossEntropyLoss

from oracle.ref_live import load_reference
from oracle.spec_io import spec_from_modules
from multimodn_b200 import MultiModNHistory
from helpers import flat_grads, assert_close
from model_utils import model_from_spec, GradTap, tapped_flat

HIST = ("loss", "accuracy", "sensitivity", "specificity", "balanced_accuracy")


class RefTap(torch.optim.Optimizer):
    def __init__(self, params):
        super().__init__(list(params), {})
        self.acc = {}

    def step(self):
        for group in self.param_groups:
            for p in group["params"]:
                if p.grad is not None:
                    self.acc[p] = p.grad.detach().clone()


def _reference_step(ref, B, feats, S, dtype):
    torch.manual_seed(11)
    encs = [ref.MIMIC_MLPEncoder(S, f, (32, 32), dropout=0.0, activation=F.relu) for f in feats]
    decs = [ref.MLPDecoder(S, (32, 32), 2) for _ in range(2)]
    rmodel = ref.MultiModN(S, encs, decs, 1.0, 0.3, device=torch.device("cpu"))
    if dtype == torch.float64:
        rmodel = rmodel.double()
    g = torch.Generator().manual_seed(5)
    xs = [torch.randn((B, f), generator=g) for f in feats]
    y = (torch.rand((B, 2), generator=g) < 0.3).to(torch.int64)
    rhist = ref.MultiModNHistory(["a", "b"])
    rtap = RefTap(rmodel.parameters())
    rmodel.train_epoch([([x.to(dtype) for x in xs], y)], rtap, CrossEntropyLoss(), rhist)

    def g_of(p):
        return rtap.acc[p].numpy()

    rgrads = dict(init_state=g_of(rmodel.init_state.state_value).reshape(-1), encoders=[], decoders=[])
    for enc in rmodel.encoders:
        lin = [m for m in enc.layers if isinstance(m, torch.nn.Linear)]
        rgrads["encoders"].append([(g_of(l.weight), g_of(l.bias)) for l in lin])
    for dec in rmodel.decoders:
        rgrads["decoders"].append([(g_of(l.weight), g_of(l.bias)) for l in dec.layers])
    return rmodel, rgrads, rhist, xs, y


def run_live(device, B, feats=(6, 99, 1024), S=64):
    ref = load_reference()
    if ref is None:
        pytest.skip("oracle/_ref is absent (run oracle/make_ref.py in the build container)")
    torch.set_num_threads(os.cpu_count() or 1)
    rmodel, rgrads, rhist, xs, y = _reference_step(ref, B, feats, S, torch.float32)
    spec = spec_from_modules(rmodel)
    with torch.no_grad():
        rpred = rmodel.predict([x[:4096] for x in xs])

    model = model_from_spec(spec, 1.0, 0.3, device, "row")
    tap = GradTap(model.parameters())
    hist = MultiModNHistory(["a", "b"])
    loader = [([x.to(device) for x in xs], y.to(device))]
    model.train_epoch(loader, tap, CrossEntropyLoss(), hist)
    got, _ = tapped_flat(model, tap)
    want = flat_grads(rgrads)
    if B < 8192:
        assert_close(got, want, rtol=1e-5, what=f"grads vs the live reference at B = {B}")
    else:
        # At tens of thousands of rows two fp32 evaluations differ by more than summation order: ~1e-6 of the ~5e7 ReLU
        # units of a step sit within round-off of their kink and switch on in one implementation and off in the other,
        # which moves a weight-gradient tensor by ~1e-4 of its magnitude.  The yardstick is therefore the reference ITSELF
        # in float64: norm-wise 1e-5 against the fp32 reference as everywhere else, and per tensor the CUDA path must be as
        # close to the float64 reference as the fp32 reference is (factor 3).
        _, rgrads64, _, _, _ = _reference_step(ref, B, feats, S, torch.float64)
        truth = flat_grads(rgrads64)
        gscale = np.abs(np.asarray(want, dtype=np.float64)).max()
        assert np.abs(got - np.asarray(want, dtype=np.float64)).max() / gscale <= 1e-5
        for name, lo, hi in want.segments:
            t64 = np.asarray(truth[lo:hi], dtype=np.float64)
            scale = np.abs(t64).max()
            e_ours = np.abs(got[lo:hi] - t64).max() / scale
            e_ref = np.abs(np.asarray(want[lo:hi], dtype=np.float64) - t64).max() / scale
            assert e_ours <= max(4e-5, 3.0 * e_ref), f"{name}: {e_ours:.2e} from the float64 reference (fp32 reference: {e_ref:.2e})"
    for n in HIST:
        assert_close(getattr(hist, n)["train"][0], getattr(rhist, n)["train"][0], rtol=1e-5, what=f"history {n}")
    assert_close(hist.state_change_loss[0], rhist.state_change_loss[0], rtol=1e-5, what="state change")
    pred = model.predict([x[:4096] for x in xs])
    assert (pred != rpred).mean() <= 1e-3


@pytest.mark.gpu
def test_c2_full_batch_against_the_live_reference():
    run_live("cuda", int(os.environ.get("MMN_LIVE_ROWS", "65536")))


def test_live_reference_small_batch_on_the_emulator(emu):
    """the same comparison at B = 96 on the emulated kernels: keeps the live-reference plumbing covered on the CPU"""
    run_live("cpu", 96, feats=(6, 19, 40), S=16)
