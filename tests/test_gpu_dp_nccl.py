"""GPU, world_size 2 over NCCL (runs only where two GPUs are visible): the data-parallel path on the real library —
row-sharded batches, gradient all-reduce (one collective for the fused fp32 kernels; per-encoder collectives behind
gradient-ready events, overlapping the remaining backward, for layer-wise bf16 plans), epoch-metric all-reduce — gives
the single-GPU result on the concatenated batch."""
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

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two GPUs")]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _make(precision, device):
    from multimodn_b200 import FusedAdam
    from oracle.spec_io import random_spec, synthetic_batch
    from model_utils import model_from_spec
    rng = np.random.default_rng(33)
    if precision == "bf16":
        S, feats, eh, dh, B = 64, [48, 80, 24], (96, 64), (32,), 512
    else:
        S, feats, eh, dh, B = 16, [6, 11, 20], (8, 8), (8,), 384
    spec = random_spec(rng, S, feats, enc_hidden=eh, dropout=0.2, n_decoders=2, dec_hidden=dh)
    data, y = synthetic_batch(rng, feats, 2, B, mnar=True)
    model = model_from_spec(spec, 0.9, 0.5, device, "row", precision=precision)
    model.runtime().dropout_base_seed = 1234
    opt = FusedAdam(model, lr=1e-2)
    return model, opt, data, y


def _run(model, opt, data, y, lo, hi, device):
    from multimodn_b200 import MultiModNHistory
    hist = MultiModNHistory(["a", "b"])
    loader = [([torch.from_numpy(x[lo:hi]).to(device) for x in data], torch.from_numpy(y[lo:hi]).to(device))]
    for _ in range(3):
        model.train_epoch(loader, opt, CrossEntropyLoss(), hist)
    torch.cuda.synchronize()
    params = torch.cat([p.detach().reshape(-1) for p in model.parameters()]).cpu().numpy().copy()
    return params, hist


def _worker(rank, world, port, precision, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        model, opt, data, y = _make(precision, dev)
        model.enable_data_parallel()
        assert (model.runtime().grad_events is not None) == (precision == "bf16")
        n = len(y) // world
        params, hist = _run(model, opt, data, y, rank * n, (rank + 1) * n, dev)
        if rank == 0:
            np.savez(out, params=params, loss=np.stack(hist.loss["train"]), sc=np.stack(hist.state_change_loss))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_two_gpus_equal_one_gpu(tmp_path, precision):
    from helpers import assert_close
    out = str(tmp_path / "dp.npz")
    mp.spawn(_worker, args=(2, _free_port(), precision, out), nprocs=2, join=True)
    got = np.load(out)
    model, opt, data, y = _make(precision, torch.device("cuda", 0))
    params, hist = _run(model, opt, data, y, 0, len(y), torch.device("cuda", 0))
    tol = 2e-5 if precision == "fp32" else 2e-3          # bf16: the shards round identical values; fp32 sums reorder
    assert_close(got["params"], params, rtol=tol, what="parameters after 3 data-parallel steps")
    assert_close(got["loss"], np.stack(hist.loss["train"]), rtol=tol, what="train loss history")
    assert_close(got["sc"], np.stack(hist.state_change_loss), rtol=tol, what="state-change history")
