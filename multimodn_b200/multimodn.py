"""MultiModN — drop-in driver of the sequential-fusion step on B200 (reference:
multimodn/multimodn.py:65-531).

Same constructor, same ``train_epoch`` / ``test`` / ``predict`` / ``get_states`` signatures, same
history matrices; underneath, each batch is ONE launch of the fused sm_100a kernel in libmmn.so
(include/mmn.h) instead of the reference's Python loop nest over encoders and decoders.

Deliberate, documented deviations from the reference:
  * ``missing_mode="row"`` (default): a NaN marks the modality missing for THAT ROW (per-row
    select) — the reference evaluated at batch size 1.  ``missing_mode="batch"`` reproduces the
    reference rule (one NaN skips the encoder for the whole batch, multimodn.py:167-169).
  * ``predict`` skips missing rows like ``test`` does (the reference propagates NaN, :449), and
    accepts numpy / list sequences (the reference needs a tensor, :518).
  * ``test`` collects, for every row, the decoder outputs at the step of encoder id E-1
    (the reference drops rows of skipped batches and mis-aligns them with the targets, :354-357).
  * an encoder id may appear at most once in an encoding sequence.
Everything else — 0.01 factor (:86), fixed D(E+1) and E divisors (:194-196), ``np.ones``
sample counters (:105,:270), unweighted batch mean (:222), CE on squashed outputs, first-max
arg-max, rows indexed by encoder id + 1, ``.grad is None`` for skipped encoders — is kept.

There is no CPU path: a non-CUDA device, a missing library, an unsupported module or criterion
raises.
"""
from __future__ import annotations

import ctypes as C
import random
from typing import Callable, Iterable, List, Optional, Tuple, Union

import numpy as np
import os
import torch
import torch.nn as nn
from torch import Tensor
from torch.optim import Optimizer
from torch.utils.data import DataLoader

from . import _lib
from .decoders.multimod_decoder import MultiModDecoder
from .encoders.multimod_encoder import MultiModEncoder
from .history import MultiModNHistory
from .metrics import get_performance_metrics, performance_metrics, to_host  # noqa: F401  (re-exported like the reference)
from .plan import PackedModel
from .state import InitState, TrainableInitState


def _check_criterion(criterion):
    """Only the criterion every reference pipeline uses is fused: default CrossEntropyLoss
    (pipelines/titanic/titanic_mlp_pipeline.py:76)."""
    ok = (isinstance(criterion, nn.CrossEntropyLoss) and criterion.weight is None
          and criterion.reduction == "mean" and criterion.ignore_index == -100
          and getattr(criterion, "label_smoothing", 0.0) == 0.0)
    if not ok:
        raise NotImplementedError("the fused step implements torch.nn.CrossEntropyLoss() with default "
                                  f"arguments only, got {criterion!r}")


class _Runtime:
    """Everything that lives next to the model on the device: plan handle, packed parameters and
    gradients, workspace, metric buffers."""

    def __init__(self, model: "MultiModN"):
        self.lib = model._lib_factory()
        self.device = model.device
        if self.device.type != "cuda" and not self.lib.host_memory:
            raise RuntimeError(f"multimodn_b200 runs on CUDA devices only (got device '{self.device}'); "
                               "there is no CPU path")
        sv = getattr(model.init_state, "state_value", None)
        if not isinstance(sv, nn.Parameter) or tuple(sv.shape) != (1, int(model.init_state.state_size)):
            raise NotImplementedError("the fused step supports TrainableInitState only (a (1, S) `state_value` parameter)")
        self.S = int(model.init_state.state_size)
        self.packed = PackedModel(model.init_state.state_value, model.encoders, model.decoders, self.S)
        self.E, self.D = len(self.packed.encoders), len(self.packed.decoders)
        self.flat = self.packed.pack(self.device)
        desc, self._keep = self.packed.model_desc(_lib.PRECISIONS[model.precision])
        handle = C.c_void_p()
        self.lib.check(self.lib.dll.mmn_plan_create(C.byref(desc), C.byref(handle)))
        self.plan = handle
        self.n_metrics = int(self.lib.dll.mmn_metrics_count(self.plan))
        self.n_grads = int(self.lib.dll.mmn_grad_count(self.plan))
        self.gflat = torch.zeros(self.n_grads, dtype=torch.float32, device=self.device)
        self.sumC = sum(m.n_classes for m in self.packed.decoders)
        self._ws = None
        self._copy_stream = None
        self._flags = torch.zeros(max(self.E, 1), dtype=torch.int32, device=self.device)
        self.step_counter = 0
        self.dropout_base_seed = int(torch.initial_seed() & 0x7FFFFFFF)
        self.layerwise = int(self.lib.dll.mmn_plan_engine(self.plan)) == 3      # bf16 plans: many launches per step
        self.grad_events = None
        self.comm_stream = None
        # CrossEntropyLoss raises on a target outside [0, C) (and ignores -100): the kernels clamp the index for memory
        # safety and set mmn_outputs.target_error; the verdict is read where the epoch's metrics are read (no sync per batch)
        self._target_err = torch.zeros(1, dtype=torch.int32, device=self.device)      # written by the kernels
        self._dp_err = torch.zeros((), dtype=torch.int64, device=self.device)

    def __del__(self):
        try:
            if getattr(self, "plan", None):
                self.lib.dll.mmn_plan_destroy(self.plan)
                self.plan = None
        except Exception:  # noqa: BLE001  (interpreter shutdown)
            pass

    def enable_grad_events(self):
        """CUDA events the library records as gradient blocks become final (mmn_plan_set_grad_events): one per encoder, one
        for the whole buffer, one for the decoders and one per encoder layer (``grad_layer_events[e][j]``)."""
        n_layers = [len(m.layers) for m in self.packed.encoders]
        n = self.E + 2 + sum(n_layers)
        events = [torch.cuda.Event() for _ in range(n)]
        for ev in events:
            ev.record(torch.cuda.current_stream(self.device))      # creates the underlying cudaEvent_t
        handles = (C.c_void_p * n)(*[ev.cuda_event for ev in events])
        self.lib.check(self.lib.dll.mmn_plan_set_grad_events(self.plan, handles, n))
        self.grad_events = events
        self.grad_layer_events, at = [], self.E + 2
        for k in n_layers:
            self.grad_layer_events.append(events[at:at + k])
            at += k
        self.comm_stream = torch.cuda.Stream(self.device)

    # ------------------------------------------------------------------------------------------
    def ensure_packed(self):
        if not self.packed.is_packed(self.flat):
            self.flat = self.packed.pack(self.device)

    def new_metrics(self) -> Tensor:
        return torch.zeros(self.n_metrics, dtype=torch.float64, device=self.device)

    def workspace(self, n_rows: int, train: bool):
        need = int(self.lib.dll.mmn_workspace_bytes(self.plan, n_rows, 1 if train else 0))
        if need < 0:
            raise _lib.MMNError("mmn_workspace_bytes failed")
        if need == 0:
            return None, 0
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._ws, need

    def stream(self):
        if self.device.type == "cuda":
            return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        return C.c_void_p(0)

    def staged(self, loader):
        """Iterate a loader one batch ahead: the host->device copies of batch i+1 (multimodn.py:132-135)
        run on a side stream while batch i computes.  Yields (data, target, encoder_sequence) with data
        and target already on the device."""
        if self.device.type != "cuda":
            for batch in loader:
                yield (list(batch) + [None])[:3]
            return
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(self.device)
        main = torch.cuda.current_stream(self.device)

        def fetch(batch):
            data, target, seq = (list(batch) + [None])[:3]
            with torch.cuda.stream(self._copy_stream):
                dev = [torch.as_tensor(t).to(device=self.device, dtype=torch.float32, non_blocking=True) for t in data]
                tgt = None if target is None else torch.as_tensor(target).to(device=self.device, dtype=torch.int64,
                                                                             non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self._copy_stream)
            return dev, tgt, seq, ev

        it = iter(loader)
        try:
            nxt = fetch(next(it))
        except StopIteration:
            return
        while nxt is not None:
            dev, tgt, seq, ev = nxt
            try:
                nxt = fetch(next(it))
            except StopIteration:
                nxt = None
            main.wait_event(ev)
            for t in dev + ([tgt] if tgt is not None else []):
                t.record_stream(main)
            yield dev, tgt, seq

    def prepare_batch(self, data: List[Tensor], target: Optional[Tensor], seq: List[Tuple[int, int]],
                      missing_mode: str, dp):
        """multimodn.py:132-135: move to the device; then fill mmn_batch."""
        if len(seq) > self.E:
            raise ValueError(f"encoding sequence has {len(seq)} steps, the model has {self.E} encoders")
        enc_ids = [e for _, e in seq]
        if len(set(enc_ids)) != len(enc_ids):
            raise ValueError("an encoder id may appear at most once in an encoding sequence")
        xs = []
        for t in data:
            t = torch.as_tensor(t)
            t = t.to(device=self.device, dtype=torch.float32, non_blocking=True)
            if t.dim() != 2:
                raise ValueError(f"modality tensors must be (batch, features), got {tuple(t.shape)}")
            if t.stride(1) != 1 or t.stride(0) < t.shape[1]:
                t = t.contiguous()
            xs.append(t)
        n_rows = xs[0].shape[0] if xs else int(target.shape[0])
        for pos, e in seq:
            if not 0 <= e < self.E:
                raise ValueError(f"encoder id {e} out of range")
            if not 0 <= pos < len(xs):
                raise ValueError(f"encoding sequence position {pos} has no data tensor")
            if xs[pos].shape[1] != self.packed.encoders[e].n_features or xs[pos].shape[0] != n_rows:
                raise ValueError(f"data[{pos}] has shape {tuple(xs[pos].shape)}, encoder {e} expects "
                                 f"({n_rows}, {self.packed.encoders[e].n_features})")
        tgt = None
        if target is not None:
            tgt = torch.as_tensor(target).to(device=self.device, dtype=torch.int64, non_blocking=True)   # :134-135
            if tgt.dim() != 2 or tgt.shape[1] != self.D or tgt.shape[0] != n_rows:
                raise ValueError(f"target must be ({n_rows}, {self.D}), got {tuple(tgt.shape)}")
            tgt = tgt.contiguous()
        npos = max(len(xs), 1)
        seq_pos = (C.c_int32 * max(len(seq), 1))(*[p for p, _ in seq])
        seq_enc = (C.c_int32 * max(len(seq), 1))(*enc_ids)
        xptr = (C.c_void_p * npos)(*[t.data_ptr() for t in xs])
        xld = (C.c_int64 * npos)(*[t.stride(0) for t in xs])
        world, rank, group = (dp[0], dp[1], dp[2]) if dp else (1, 0, None)
        n_global, row_offset = n_rows * world, rank * n_rows
        if dp and len(dp) > 3 and dp[3] is not None:        # (n_rows_global, row_offset) agreed across the ranks
            n_global, row_offset = dp[3]
        b = _lib.Batch()
        b.n_rows, b.n_rows_global, b.row_offset = n_rows, n_global, row_offset
        b.seq_len = len(seq)
        b.seq_pos = C.cast(seq_pos, C.POINTER(C.c_int32))
        b.seq_enc = C.cast(seq_enc, C.POINTER(C.c_int32))
        b.x = C.cast(xptr, C.POINTER(C.c_void_p))
        b.x_ld = C.cast(xld, C.POINTER(C.c_int64))
        b.targets = tgt.data_ptr() if tgt is not None else None
        b.skip_flags = None
        keep = (xs, tgt, seq_pos, seq_enc, xptr, xld)
        if missing_mode == "batch" and len(seq):
            self.lib.check(self.lib.dll.mmn_scan_missing(self.plan, C.byref(b), self._flags.data_ptr(), self.stream()))
            if group is not None or world > 1:
                torch.distributed.all_reduce(self._flags, op=torch.distributed.ReduceOp.MAX, group=group)
            b.skip_flags = self._flags.data_ptr()
        elif missing_mode != "row" and missing_mode != "batch":
            raise ValueError(f"missing_mode must be 'row' or 'batch', got {missing_mode!r}")
        return b, keep, n_rows

    def outputs(self, metrics=None, predictions=None, last_outputs=None, final_state=None):
        o = _lib.Outputs()
        o.metrics = metrics.data_ptr() if metrics is not None else None
        o.predictions = predictions.data_ptr() if predictions is not None else None
        o.pred_ld = predictions.shape[-1] if predictions is not None else 0
        o.last_outputs = last_outputs.data_ptr() if last_outputs is not None else None
        o.final_state = final_state.data_ptr() if final_state is not None else None
        o.target_error = self._target_err.data_ptr()
        return o

    def forward(self, batch, n_rows, **outs):
        self.ensure_packed()
        o = self.outputs(**outs)
        ws, ws_bytes = self.workspace(n_rows, False)        # bf16 plans keep their activations there; fp32 plans need none
        self.lib.check(self.lib.dll.mmn_forward(self.plan, C.byref(batch), self.flat.data_ptr(), C.byref(o),
                                                ws.data_ptr() if ws is not None else None, ws_bytes, self.stream()))

    def train_step(self, batch, n_rows, err_penalty, scp_scaled, training, metrics):
        self.ensure_packed()
        ws, ws_bytes = self.workspace(n_rows, True)
        self.step_counter += 1
        seed = (self.dropout_base_seed * 0x9E3779B1 + self.step_counter * 0x85EBCA77) & 0xFFFFFFFF
        targs = _lib.TrainArgs(err_penalty, scp_scaled, seed, 1 if training else 0)
        o = self.outputs(metrics=metrics)
        self.lib.check(self.lib.dll.mmn_train_step(self.plan, C.byref(batch), self.flat.data_ptr(), C.byref(targs),
                                                   C.byref(o), self.gflat.data_ptr(), ws.data_ptr(), ws_bytes,
                                                   self.stream()))
        return seed

    def check_targets(self):
        """raise like nn.CrossEntropyLoss does for a target outside [0, n_classes) (one small D2H; called where the
        epoch's metrics are read back anyway)"""
        if int(self._target_err.item()) != 0:
            self._target_err.zero_()
            raise IndexError("Target out of bounds: a target lies outside [0, n_classes) of its decoder "
                             "(ignore_index = -100 is not supported by the fused step)")

    def read_back_async(self, metrics: Tensor, tag):
        """device metrics -> one of two alternating pinned host buffers, without blocking the host; returns
        (tag, host tensor, event to wait for) — (tag, tensor, None) on a host-memory library"""
        if self.device.type != "cuda":
            return tag, metrics.clone(), None
        if not hasattr(self, "_rb"):
            self._rb = [torch.empty(self.n_metrics, dtype=torch.float64).pin_memory() for _ in range(2)]
            self._rb_i = 0
        host = self._rb[self._rb_i]
        self._rb_i ^= 1
        host.copy_(metrics, non_blocking=True)
        event = torch.cuda.Event()
        event.record(torch.cuda.current_stream(self.device))
        return tag, host, event

    def assign_grads(self):
        """loss.backward() epilogue (multimodn.py:203): hand each parameter a view of the packed
        gradient; an encoder that took no row keeps ``.grad = None`` (multimodn.py:168-169)."""
        touched = self.gflat[self.packed.n_params:].tolist()        # E floats (one small D2H)
        for p, off, owner in self.packed.slots:
            if owner >= 0 and touched[owner] <= 0:
                p.grad = None
            else:
                p.grad = self.gflat[off:off + p.numel()].view(p.shape)

    # ------------------------------------------------------------------------------------------
    def split_metrics(self, host: np.ndarray):
        E, D = self.E, self.D
        n = (E + 1) * D
        mats = [host[i * n:(i + 1) * n].reshape(E + 1, D).copy() for i in range(6)]
        n_present = host[6 * n:6 * n + E + 1].copy()
        sc = host[6 * n + E + 1:6 * n + 2 * E + 1].copy()
        return mats, n_present, sc


class MultiModN(nn.Module):
    _lib_factory = staticmethod(_lib.get_lib)

    def __init__(
            self,
            state_size: int,
            encoders: List[MultiModEncoder],
            decoders: List[MultiModDecoder],
            err_penalty: float,
            state_change_penalty: float,
            shuffle_mode: Optional[bool] = False,
            init_state: Optional[InitState] = None,
            device: Optional[torch.device] = None,
            missing_mode: str = "row",
            precision: str = "fp32",
    ):
        super().__init__()
        self.shuffle_mode = shuffle_mode
        self.device = torch.device(device) if device else torch.device(
            "cuda" if torch.cuda.is_available() else "cpu")
        self.init_state = TrainableInitState(state_size, self.device) if not init_state else init_state
        self.encoders = nn.ModuleList(encoders)
        self.decoders = nn.ModuleList(decoders)
        self.err_penalty = err_penalty
        self.state_change_penalty = 0.01 * state_change_penalty      # multimodn.py:86
        if missing_mode not in ("row", "batch"):
            raise ValueError(f"missing_mode must be 'row' or 'batch', got {missing_mode!r}")
        self.missing_mode = missing_mode
        if precision not in _lib.PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(_lib.PRECISIONS)}, got {precision!r}")
        self.precision = precision
        self.to(self.device)
        self._rt: Optional[_Runtime] = None
        self._dp = None                 # (world, rank, group) once data parallelism is enabled
        self._dp_hash = 0

    # -- plumbing ------------------------------------------------------------------------------
    def __getstate__(self):
        state = self.__dict__.copy()
        state["_rt"] = None             # plan handles and packed buffers are rebuilt on demand
        state["_dp"] = None
        state["_dp_hash"] = 0
        return state

    def runtime(self) -> _Runtime:
        if self._rt is None:
            self._rt = _Runtime(self)
        return self._rt

    def enable_data_parallel(self, group=None):
        """Shard every batch by rows over the ranks of ``group`` (one process per GPU): the loader
        of rank r yields its own rows; gradients (and the epoch's metric sums) are all-reduced over
        NCCL.  Parameters must start identical on every rank (same seed)."""
        import torch.distributed as dist
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        self._dp = (dist.get_world_size(group), dist.get_rank(group), group)
        rt = self.runtime()
        rt.ensure_packed()
        dist.broadcast(rt.flat, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        if self._dp[0] > 1 and rt.layerwise and self.device.type == "cuda":
            rt.enable_grad_events()
            # SMs the backward GEMMs leave to the collectives' kernels (mmn_plan_set_comm_sms): measured on B200, NCCL 2.28 —
            # 2 GPUs (P2P rings) 128 of 148, 8 GPUs (NVLS, 24 channels) 112
            rt.lib.check(rt.lib.dll.mmn_plan_set_comm_sms(rt.plan, 128 if self._dp[0] <= 2 else 112))
        return self

    def _allreduce(self, tensor):
        if self._dp and self._dp[0] > 1:
            torch.distributed.all_reduce(tensor, group=self._dp[2])

    def _allreduce_grads(self, rt, seq):
        """Gradient all-reduce of one train step.  Fused single-launch plans: one collective over the packed buffer.
        Layer-wise (bf16) plans: one collective per gradient block on a side stream behind the block's gradient-ready
        event, so that it overlaps the remaining backward GEMMs: first the decoders (finished before the
        encoders' backward), then every encoder layer in the order the default sequence finishes them (descending encoder
        id, last layer first) — the block left exposed at the end of the step is one layer — then the initial state and the
        counters.  The order does not depend on the rank-local sequence (an encoder outside the sequence waits for the
        end-of-step event), so ranks that disagree on the sequence still pair their collectives — the disagreement itself
        is reported by ``_dp_note`` / ``check_data_parallel``."""
        if not (self._dp and self._dp[0] > 1):
            return
        if not rt.grad_events:
            torch.distributed.all_reduce(rt.gflat, group=self._dp[2])
            return
        main = torch.cuda.current_stream(self.device)
        comm = rt.comm_stream
        in_seq = {e for _, e in seq}
        all_ids = list(range(rt.E - 1, -1, -1))
        with torch.cuda.stream(comm):
            dec = rt.packed.decoder_range()
            comm.wait_event(rt.grad_events[rt.E + 1])
            torch.distributed.all_reduce(rt.gflat[dec[0]:dec[1]], group=self._dp[2])
            # one collective per encoder layer shortens the exposed tail (one layer instead of one encoder) but triples the
            # number of collectives; MMN_DP_BLOCKS=layer|encoder overrides the choice below
            per_layer = os.environ.get("MMN_DP_BLOCKS", "layer" if self._dp[0] <= 2 else "encoder") != "encoder"
            for e in all_ids:
                if not per_layer:       # one collective per encoder (fewer launches, a longer exposed tail)
                    comm.wait_event(rt.grad_events[e if e in in_seq else rt.E])
                    lo, hi = rt.packed.encoder_range(e)
                    torch.distributed.all_reduce(rt.gflat[lo:hi], group=self._dp[2])
                    continue
                ranges = rt.packed.encoder_layer_ranges(e)
                for j in range(len(ranges) - 1, -1, -1):
                    comm.wait_event(rt.grad_layer_events[e][j] if e in in_seq else rt.grad_events[rt.E])
                    torch.distributed.all_reduce(rt.gflat[ranges[j][0]:ranges[j][1]], group=self._dp[2])
            comm.wait_event(rt.grad_events[rt.E])
            for lo, hi in rt.packed.complement_ranges(all_ids, rt.n_grads, also=[dec]):
                torch.distributed.all_reduce(rt.gflat[lo:hi], group=self._dp[2])
        main.wait_stream(comm)

    # Data-parallel contract: every rank feeds the same number of rows per batch (use drop_last / equal shards: the kernel
    # divides by n_rows * world and keys the dropout stream by rank * n_rows) and the same encoding sequence.  Both are
    # folded into a running hash per epoch; one tiny MAX all-reduce per epoch compares the ranks without a host sync and
    # the verdict is read with the epoch's metrics (``_finalize``) or by ``check_data_parallel()``.
    def _dp_note(self, n_rows, seq):
        if self._dp and self._dp[0] > 1:
            h = self._dp_hash
            for v in (n_rows, len(seq), *[p * 131 + e for p, e in seq]):
                h = (h * 1000003 + int(v) + 1) % 2147483629
            self._dp_hash = h

    def _dp_close_epoch(self, rt):
        if not (self._dp and self._dp[0] > 1):
            return
        h = float(self._dp_hash)
        self._dp_hash = 0
        t = torch.full((2,), h, dtype=torch.float64, device=rt.device)     # fill kernels: no host-to-device copy, no sync
        t[1].neg_()
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX, group=self._dp[2])
        rt._dp_err = torch.maximum(rt._dp_err, (t[0] + t[1] != 0).to(torch.int64))

    def check_data_parallel(self):
        """raise if the ranks of the data-parallel group disagreed on a batch size or an encoding sequence (one D2H)"""
        rt = self.runtime()
        if self._dp and self._dp[0] > 1 and int(rt._dp_err.item()) != 0:
            rt._dp_err.zero_()
            raise RuntimeError("data-parallel ranks disagreed on the rows per batch or on the encoding sequence: every rank "
                               "must feed equally sized batches (drop_last / equal shards) and the same sequence; "
                               "gradients of this epoch are not the global-batch gradients")

    # -- the encoding sequence (multimodn.py:509-531) --------------------------------------------
    def get_encoder_iterable(self, encoder_sequence, shuffle_mode: bool, train: bool) -> List[Tuple[int, int]]:
        if encoder_sequence is None:
            pairs = list(enumerate(range(len(self.encoders))))
        else:
            seq = encoder_sequence.detach().cpu().numpy() if torch.is_tensor(encoder_sequence) \
                else np.asarray(encoder_sequence)
            if seq.ndim == 2:
                if not (seq == seq[0]).all():
                    raise ValueError("Encoder sequence has different values across the batch. "
                                     "Hint: set batch size to 1 to avoid this error.")
                seq = seq[0]
            elif seq.ndim != 1:
                raise ValueError("encoder_sequence must be (batch, steps) or (steps,)")
            pairs = [(i, int(e)) for i, e in enumerate(seq)]
        if shuffle_mode and train:
            random.shuffle(pairs)                                   # multimodn.py:527-529
            if self._dp and self._dp[0] > 1:                        # one order for the whole global batch: rank 0's
                box = [pairs]
                src = torch.distributed.get_global_rank(self._dp[2], 0) if self._dp[2] is not None else 0
                torch.distributed.broadcast_object_list(box, src=src, group=self._dp[2])
                pairs = [tuple(p) for p in box[0]]
        return pairs

    # -- epoch bookkeeping (multimodn.py:222-250, 367-409) ---------------------------------------
    def _finalize(self, rt: _Runtime, metrics: Tensor, n_batches: int):
        self._allreduce(metrics)
        rt.check_targets()
        self.check_data_parallel()
        mats, n_present, sc = rt.split_metrics(metrics.cpu().numpy())
        ce, n_correct, tp, tn, fp, fn = mats
        n_samples = np.ones((rt.E + 1, 1)) + n_present.reshape(-1, 1)      # starts at ONE (:105,:270)
        nb = max(n_batches, 1)
        for d, dec in enumerate(rt.packed.decoders):
            if dec.n_classes != 2:                                         # :60-63: NaN cells once visited
                for m in (tp, tn, fp, fn):
                    m[:, d] = np.where(n_present > 0, np.nan, 0.0)
        tp32, tn32, fp32, fn32 = (m.astype(np.float32) for m in (tp, tn, fp, fn))
        with np.errstate(divide="ignore", invalid="ignore"):
            sens = np.where(tp32 + fn32 == 0, np.float32(0), tp32 / (tp32 + fn32))     # :234-236
            spec = np.where(tn32 + fp32 == 0, np.float32(0), tn32 / (tn32 + fp32))     # :238-240
        return dict(loss=ce / nb, accuracy=n_correct / n_samples, sensitivity=sens, specificity=spec,
                    balanced_accuracy=(sens + spec) / 2, state_change=sc / nb)

    # -- training (multimodn.py:89-252) ----------------------------------------------------------
    def train_epoch(
            self,
            train_loader: DataLoader,
            optimizer: Optimizer,
            criterion: Union[nn.Module, Callable],
            history: Optional[MultiModNHistory] = None,
            log_interval: Optional[int] = None,
            logger: Optional[Callable] = None,
            last_epoch: Optional[bool] = False,
    ):
        if log_interval and not logger:
            logger = print
        _check_criterion(criterion)
        self.train()
        rt = self.runtime()
        fused_opt = getattr(optimizer, "_mmn_fused", False)
        n_batches = len(train_loader)
        epoch_metrics = rt.new_metrics()
        batch_metrics = rt.new_metrics() if log_interval else None
        E, D = rt.E, rt.D

        pending_log = None

        def emit_log(entry):
            idx, host, event = entry
            if event is not None:
                event.synchronize()
            mats, _, sc = rt.split_metrics(host.numpy())
            err = mats[0].sum() / (D * (E + 1))                      # :194-196
            chg = sc.sum() / E
            loss = err * self.err_penalty + chg * self.state_change_penalty
            logger(f"Batch {idx + 1}/{n_batches}\n"
                   f"\tLoss: {loss:.4f}\n"
                   f"\tErr loss: {err:.4f}\n"
                   f"\tState change: {chg:.4f}")

        for batch_idx, (data, target, encoder_sequence) in enumerate(rt.staged(train_loader)):
            seq = self.get_encoder_iterable(encoder_sequence, shuffle_mode=self.shuffle_mode, train=True)
            optimizer.zero_grad()
            mb, keep, n_rows = rt.prepare_batch(list(data), target, seq, self.missing_mode, self._dp)
            self._dp_note(n_rows, seq)
            if batch_metrics is not None:
                batch_metrics.zero_()
            rt.train_step(mb, n_rows, float(self.err_penalty), float(self.state_change_penalty), True,
                          batch_metrics if batch_metrics is not None else epoch_metrics)
            self._allreduce_grads(rt, seq)
            if fused_opt:
                optimizer.step()                # reads the packed gradient on the device: no host sync
            else:
                rt.assign_grads()
                optimizer.step()
            del keep
            if batch_metrics is not None:
                epoch_metrics += batch_metrics
                if batch_idx % log_interval == log_interval - 1:
                    # multimodn.py:214-220.  The batch's loss is read back asynchronously (pinned buffer + event) and the
                    # line is emitted once the NEXT batch has been enqueued: same lines, same order, but the device never
                    # idles behind a host round trip per logged batch
                    bm = batch_metrics.clone()
                    self._allreduce(bm)
                    entry = rt.read_back_async(bm, batch_idx)
                    if pending_log is not None:
                        emit_log(pending_log)
                    pending_log = entry
        if pending_log is not None:
            emit_log(pending_log)

        self._dp_close_epoch(rt)
        if history is not None:
            fin = self._finalize(rt, epoch_metrics, n_batches)
            history.state_change_loss.append(fin["state_change"])
            history.append("train", fin)
        if last_epoch:
            return self.test(train_loader, criterion, history=None)

    # -- evaluation (multimodn.py:255-419) -------------------------------------------------------
    def test(
            self,
            test_loader: DataLoader,
            criterion: Union[nn.Module, Callable],
            history: Optional[MultiModNHistory] = None,
            tag: str = 'test',
            log_results: bool = False,
            logger: Optional[Callable] = None,
    ):
        if log_results and not logger:
            logger = print
        _check_criterion(criterion)
        self.eval()
        rt = self.runtime()
        n_batches = len(test_loader)
        metrics = rt.new_metrics()
        outs, tgts = [], []
        for data, target, encoder_sequence in rt.staged(test_loader):
            seq = self.get_encoder_iterable(encoder_sequence, shuffle_mode=self.shuffle_mode, train=False)
            mb, keep, n_rows = rt.prepare_batch(list(data), target, seq, self.missing_mode, self._dp)
            self._dp_note(n_rows, seq)
            last = torch.zeros((n_rows, rt.sumC), dtype=torch.float32, device=rt.device)
            rt.forward(mb, n_rows, metrics=metrics, last_outputs=last)
            outs.append(last)
            tgts.append(keep[1])
            del keep
        self._dp_close_epoch(rt)
        fin = self._finalize(rt, metrics, n_batches)
        if log_results:
            logger(f"{tag.capitalize()} results\n"
                   f"\tAverage loss: {np.mean(fin['loss']):.4f}\n"
                   f"\tAccuracy: {np.mean(fin['accuracy']):.4f}\n"
                   f"\tSensitivity: {np.nanmean(fin['sensitivity']):.4f}\n"
                   f"\tSpecificity: {np.nanmean(fin['specificity']):.4f}\n"
                   f"\tBalanced accuracy: {np.nanmean(fin['balanced_accuracy']):.4f}")
        if history is not None:
            history.append(tag, fin)
        # end-of-test metric suite on the last-encoder outputs (multimodn.py:411-419)
        results = [[] for _ in range(rt.D)]
        if outs:
            # device-side: the (N, sum C) outputs stay in HBM; sort-based curves / AUROC / F1 are torch CUDA
            # ops there and only the result tuple is copied to the host
            out_all = torch.cat(outs)
            tgt_all = torch.cat(tgts)
            col = 0
            for d, dec in enumerate(rt.packed.decoders):
                o = out_all[:, col:col + dec.n_classes]
                col += dec.n_classes
                o = o / o.sum(dim=1, keepdim=True)                          # :415
                pred = torch.max(o, dim=1)[1]                               # :416
                results[d] = to_host(get_performance_metrics(tgt_all[:, d], pred, o[:, 1] if dec.n_classes > 1 else o[:, 0]))
        return results

    # -- inference (multimodn.py:422-458) --------------------------------------------------------
    def predict(self, x: List[Tensor], encoder_sequence=None) -> np.ndarray:
        self.eval()
        rt = self.runtime()
        seq = self.get_encoder_iterable(encoder_sequence, shuffle_mode=self.shuffle_mode, train=False)
        mb, keep, n_rows = rt.prepare_batch(list(x), None, seq, self.missing_mode, None)
        preds = torch.zeros((rt.E + 1, rt.D, n_rows), dtype=torch.uint8, device=rt.device)
        rt.forward(mb, n_rows, predictions=preds)
        del keep
        if rt.device.type != "cuda":
            return preds.numpy().astype(np.float64)
        # the reference returns float64 class ids (multimodn.py:429-430): widen on the device and land in pinned host
        # memory (torch caches the pinned block), instead of a single-threaded astype over (E+1) D N values on the host
        host = torch.empty(preds.shape, dtype=torch.float64, pin_memory=True)
        host.copy_(preds.to(torch.float64))
        return host.numpy()

    def get_states(self, data_loader: DataLoader) -> List[Tensor]:
        """multimodn.py:460-492"""
        self.eval()
        rt = self.runtime()
        states = []
        for data, _, encoder_sequence in rt.staged(data_loader):
            seq = self.get_encoder_iterable(encoder_sequence, shuffle_mode=self.shuffle_mode, train=False)
            mb, keep, n_rows = rt.prepare_batch(list(data), None, seq, self.missing_mode, None)
            st = torch.empty((n_rows, rt.S), dtype=torch.float32, device=rt.device)
            rt.forward(mb, n_rows, final_state=st)
            states.append(st)
            del keep
        return list(torch.cat(states, dim=0))

    def display_arch(self, input: Optional[np.ndarray] = None):
        """multimodn.py:494-507 (torchsummary is optional here: the lowered plan is printed)."""
        rt = self.runtime()
        for name, mods in (("Encoder", rt.packed.encoders), ("Decoder", rt.packed.decoders)):
            for i, m in enumerate(mods):
                print(f"{name} {i}: " + " -> ".join(
                    f"Linear({l.in_dim}{'+S' if l.has_state else ''}, {l.out_dim}){'/' + l.act if l.act != 'identity' else ''}"
                    for l in m.layers))
