"""ctypes binding of libmmn.so (include/mmn.h) — the only way the package reaches the GPU.

There is no CPU path and no fallback: if the library has not been built, or a call fails,
this module raises.  Build with ``python -c "import __graft_entry__ as g; g.build()"`` or
``make -C multimodn_b200/csrc``.
"""
from __future__ import annotations

import ctypes as C
import os

MAX_LAYERS = 6
MAX_ENCODERS = 16
MAX_DECODERS = 16
MAX_CLASSES = 32
ACT_CODES = {"identity": 0, "relu": 1, "sigmoid": 2, "tanh": 3}
ABI_VERSION = 3

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libmmn.so")


class LayerDesc(C.Structure):
    _fields_ = [("in_dim", C.c_int32), ("out_dim", C.c_int32), ("act", C.c_int32), ("has_state", C.c_int32),
                ("w_off", C.c_int64), ("b_off", C.c_int64)]


class EncoderDesc(C.Structure):
    _fields_ = [("n_features", C.c_int32), ("n_layers", C.c_int32), ("dropout_p", C.c_float),
                ("reserved", C.c_int32), ("layers", LayerDesc * MAX_LAYERS)]


class DecoderDesc(C.Structure):
    _fields_ = [("n_classes", C.c_int32), ("n_layers", C.c_int32), ("layers", LayerDesc * MAX_LAYERS)]


class ModelDesc(C.Structure):
    _fields_ = [("state_size", C.c_int32), ("n_encoders", C.c_int32), ("n_decoders", C.c_int32),
                ("precision", C.c_int32), ("init_off", C.c_int64), ("n_params", C.c_int64),
                ("encoders", C.POINTER(EncoderDesc)), ("decoders", C.POINTER(DecoderDesc))]


class Batch(C.Structure):
    _fields_ = [("n_rows", C.c_int64), ("n_rows_global", C.c_int64), ("row_offset", C.c_int64),
                ("seq_len", C.c_int32), ("reserved", C.c_int32),
                ("seq_pos", C.POINTER(C.c_int32)), ("seq_enc", C.POINTER(C.c_int32)),
                ("x", C.POINTER(C.c_void_p)), ("x_ld", C.POINTER(C.c_int64)),
                ("targets", C.c_void_p), ("skip_flags", C.c_void_p)]


class Outputs(C.Structure):
    _fields_ = [("metrics", C.c_void_p), ("predictions", C.c_void_p), ("pred_ld", C.c_int64),
                ("last_outputs", C.c_void_p), ("final_state", C.c_void_p), ("target_error", C.c_void_p)]


class TrainArgs(C.Structure):
    _fields_ = [("err_penalty", C.c_float), ("state_change_penalty_scaled", C.c_float),
                ("dropout_seed", C.c_uint32), ("training", C.c_int32)]


PRECISIONS = {"fp32": 0, "bf16": 1}      # MMN_PRECISION_*

EXPORTS = ("mmn_last_error", "mmn_abi_version", "mmn_plan_create", "mmn_plan_destroy", "mmn_metrics_count",
           "mmn_grad_count", "mmn_workspace_bytes", "mmn_scan_missing", "mmn_forward", "mmn_train_step",
           "mmn_adam_step", "mmn_selftest_umma", "mmn_selftest_protocol", "mmn_plan_engine", "mmn_plan_forward_engine",
           "mmn_selftest_gemm_bf16", "mmn_selftest_gemm_bf16_mn", "mmn_wide_launch_count", "mmn_plan_set_grad_events", "mmn_plan_set_comm_sms",
           "mmn_selftest_fma_peak")


class MMNError(RuntimeError):
    pass


class Library:
    """Typed handle on a loaded libmmn.  ``host_memory`` is False for the real library: every
    pointer handed to it must be CUDA device memory."""

    host_memory = False

    def __init__(self, path=LIB_PATH):
        if not os.path.exists(path):
            raise MMNError(f"{path} is missing: the CUDA library has not been built "
                           "(run __graft_entry__.build() or `make -C multimodn_b200/csrc`). "
                           "multimodn_b200 has no CPU fallback.")
        self.path = path
        self.dll = C.CDLL(path)
        d = self.dll
        for name in EXPORTS:
            if not hasattr(d, name):
                raise MMNError(f"{path} does not export {name}")
        d.mmn_last_error.restype = C.c_char_p
        d.mmn_abi_version.restype = C.c_int
        d.mmn_plan_create.argtypes = [C.POINTER(ModelDesc), C.POINTER(C.c_void_p)]
        d.mmn_plan_destroy.argtypes = [C.c_void_p]
        d.mmn_plan_destroy.restype = None
        d.mmn_metrics_count.argtypes = [C.c_void_p]
        d.mmn_metrics_count.restype = C.c_int64
        d.mmn_grad_count.argtypes = [C.c_void_p]
        d.mmn_grad_count.restype = C.c_int64
        d.mmn_workspace_bytes.argtypes = [C.c_void_p, C.c_int64, C.c_int32]
        d.mmn_workspace_bytes.restype = C.c_int64
        d.mmn_scan_missing.argtypes = [C.c_void_p, C.POINTER(Batch), C.c_void_p, C.c_void_p]
        d.mmn_forward.argtypes = [C.c_void_p, C.POINTER(Batch), C.c_void_p, C.POINTER(Outputs), C.c_void_p,
                                  C.c_size_t, C.c_void_p]
        d.mmn_train_step.argtypes = [C.c_void_p, C.POINTER(Batch), C.c_void_p, C.POINTER(TrainArgs),
                                     C.POINTER(Outputs), C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        d.mmn_adam_step.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p]
        d.mmn_plan_forward_engine.argtypes = [C.c_void_p]
        d.mmn_plan_forward_engine.restype = C.c_int32
        d.mmn_plan_engine.argtypes = [C.c_void_p]
        d.mmn_plan_engine.restype = C.c_int32
        d.mmn_selftest_protocol.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        d.mmn_selftest_umma.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        d.mmn_selftest_gemm_bf16.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_longlong, C.c_void_p, C.c_longlong,
                                             C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        d.mmn_selftest_gemm_bf16_mn.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_longlong, C.c_void_p, C.c_longlong,
                                                C.c_void_p, C.c_void_p]
        d.mmn_wide_launch_count.restype = C.c_int64
        d.mmn_selftest_fma_peak.argtypes = [C.c_int, C.c_void_p, C.POINTER(C.c_double), C.c_void_p]
        d.mmn_plan_set_grad_events.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.c_int32]
        d.mmn_plan_set_comm_sms.argtypes = [C.c_void_p, C.c_int32]
        if d.mmn_abi_version() != ABI_VERSION:
            raise MMNError(f"{path}: ABI version {d.mmn_abi_version()} != {ABI_VERSION}; rebuild the library")

    def check(self, status):
        if status != 0:
            raise MMNError(self.dll.mmn_last_error().decode())

    def stream_for(self, device):
        import torch
        return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


_LIB = None


def get_lib():
    global _LIB
    if _LIB is None:
        _LIB = Library()
    return _LIB
