"""CPU, world_size 2 over gloo: the data-parallel path of MultiModN (row-sharded batches, gradient
all-reduce, epoch-metric all-reduce, batch-mode skip-flag all-reduce) reproduces the single-process
result on the concatenated batch.  Kernels run on the CPU emulator build (tests/emu); on the B200 box the
same host code drives NCCL (bench.py --gpus N)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from torch.nn import CrossEntropyLoss

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.join(HERE, "emu"))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _make(missing_mode, optimizer):
    from emu_backend import get_emu_lib
    from multimodn_b200 import MultiModN, FusedAdam
    from oracle.spec_io import random_spec, synthetic_batch
    from model_utils import model_from_spec
    lib = get_emu_lib()
    MultiModN._lib_factory = staticmethod(lambda: lib)
    rng = np.random.default_rng(21)
    feats = [6, 11, 20]
    spec = random_spec(rng, 16, feats, enc_hidden=(8, 8), n_decoders=2, dec_hidden=(8,))
    data, y = synthetic_batch(rng, feats, 2, 96, mnar=True)
    if missing_mode == "batch":
        data = [np.nan_to_num(x) for x in data]
        data[1][70, 2] = np.nan                    # lives in rank 1's shard only: rank 0 must skip too
    model = model_from_spec(spec, 0.9, 0.5, "cpu", missing_mode)
    opt = FusedAdam(model, lr=1e-2) if optimizer == "fused" else torch.optim.Adam(list(model.parameters()), 1e-2)
    return model, opt, data, y


def _run(model, opt, data, y, lo, hi):
    from multimodn_b200 import MultiModNHistory
    hist = MultiModNHistory(["a", "b"])
    loader = [([torch.from_numpy(x[lo:hi]) for x in data], torch.from_numpy(y[lo:hi]))]
    for _ in range(2):
        model.train_epoch(loader, opt, CrossEntropyLoss(), hist)
    model.test(loader, CrossEntropyLoss(), hist, tag="val")
    params = torch.cat([p.detach().reshape(-1) for p in model.parameters()]).numpy().copy()
    return params, hist


def _worker(rank, world, port, missing_mode, optimizer, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        model, opt, data, y = _make(missing_mode, optimizer)
        model.enable_data_parallel()
        n = len(y) // world
        params, hist = _run(model, opt, data, y, rank * n, (rank + 1) * n)
        if rank == 0:
            np.savez(out, params=params, loss=np.stack(hist.loss["train"]), acc=np.stack(hist.accuracy["train"]),
                     sc=np.stack(hist.state_change_loss), val=hist.loss["val"][0])
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("missing_mode,optimizer", [("row", "torch"), ("row", "fused"), ("batch", "torch")])
def test_two_ranks_equal_one_rank(tmp_path, missing_mode, optimizer):
    from helpers import assert_close
    out = str(tmp_path / "dp.npz")
    mp.spawn(_worker, args=(2, _free_port(), missing_mode, optimizer, out), nprocs=2, join=True)
    got = np.load(out)
    model, opt, data, y = _make(missing_mode, optimizer)
    params, hist = _run(model, opt, data, y, 0, len(y))
    assert_close(got["params"], params, rtol=2e-5, what="parameters after 2 DP steps")
    assert_close(got["loss"], np.stack(hist.loss["train"]), rtol=1e-5, what="train loss history")
    assert_close(got["acc"], np.stack(hist.accuracy["train"]), rtol=1e-6, what="train accuracy history")
    assert_close(got["sc"], np.stack(hist.state_change_loss), rtol=1e-5, what="state-change history")
    assert_close(got["val"], hist.loss["val"][0], rtol=1e-5, what="val loss")


def _contract_worker(rank, world, port, case, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import random
        from multimodn_b200 import MultiModNHistory
        model, opt, data, y = _make("row", "torch")
        model.enable_data_parallel()
        n = len(y) // world
        lo, hi = rank * n, (rank + 1) * n
        msg = "ok"
        if case == "ragged":                      # rank 1 feeds fewer rows than rank 0: must be reported, not averaged away
            hi -= 5 * rank
        if case == "shuffle":
            model.shuffle_mode = True
            random.seed(100 + rank)               # different local draws: rank 0's order must win everywhere
        loader = [([torch.from_numpy(x[lo:hi]) for x in data], torch.from_numpy(y[lo:hi]))]
        try:
            model.train_epoch(loader, opt, CrossEntropyLoss(), MultiModNHistory(["a", "b"]))
        except RuntimeError as exc:
            msg = str(exc)
        params = torch.cat([p.detach().reshape(-1) for p in model.parameters()]).numpy().copy()
        np.savez(out + f".{rank}.npz", params=params, msg=np.array(msg))
    finally:
        dist.destroy_process_group()


def test_ranks_with_different_batch_sizes_are_reported(tmp_path):
    out = str(tmp_path / "dp")
    mp.spawn(_contract_worker, args=(2, _free_port(), "ragged", out), nprocs=2, join=True)
    for r in range(2):
        assert "disagreed on the rows per batch" in str(np.load(out + f".{r}.npz")["msg"])


def test_shuffle_mode_under_data_parallel_uses_one_order(tmp_path):
    """shuffle_mode draws from the per-process `random`; under DP rank 0's order is broadcast, so both ranks apply the same
    sequence and end the step with identical parameters"""
    out = str(tmp_path / "dp")
    mp.spawn(_contract_worker, args=(2, _free_port(), "shuffle", out), nprocs=2, join=True)
    a, b = np.load(out + ".0.npz"), np.load(out + ".1.npz")
    assert str(a["msg"]) == "ok" and str(b["msg"]) == "ok"
    assert (a["params"] == b["params"]).all()


def test_gradient_blocks_partition_the_packed_buffer():
    """the blocks `_allreduce_grads` reduces one by one on layer-wise plans — the decoders, every encoder layer, the rest —
    cover the packed gradient buffer exactly once (alignment padding aside), whichever granularity is used"""
    from oracle.spec_io import random_spec
    from model_utils import model_from_spec
    from multimodn_b200.plan import PackedModel
    rng = np.random.default_rng(3)
    spec = random_spec(rng, 12, [5, 9, 7], enc_kind="mimic", enc_hidden=(10, 6), dropout=0.0, n_decoders=3, dec_hidden=(8,), n_classes=2)
    model = model_from_spec(spec, 1.0, 0.3, "cpu", "row")
    packed = PackedModel(model.init_state.state_value, model.encoders, model.decoders, 12)
    E, total = len(model.encoders), packed.n_params + len(model.encoders)
    ids = list(range(E))
    dec = packed.decoder_range()
    for per_layer in (True, False):
        covered = np.zeros(total, dtype=np.int32)
        covered[dec[0]:dec[1]] += 1
        for e in ids:
            blocks = packed.encoder_layer_ranges(e) if per_layer else [packed.encoder_range(e)]
            lo_e, hi_e = packed.encoder_range(e)
            for lo, hi in blocks:
                assert lo_e <= lo < hi <= hi_e
                covered[lo:hi] += 1
        for lo, hi in packed.complement_ranges(ids, total, also=[dec]):
            covered[lo:hi] += 1
        assert covered.max() == 1
        # every parameter (and the per-encoder counters behind them) is in exactly one block; only alignment padding is not
        used = np.zeros(total, dtype=bool)
        for p, off, _ in packed.slots:
            used[off:off + p.numel()] = True
        used[packed.n_params:] = True
        assert (covered[used] == 1).all()
