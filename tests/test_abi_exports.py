"""CPU: libmmn.so (the nvcc-built sm_100a library) loads without a GPU and exports every symbol that
include/mmn.h declares; struct layouts of the ctypes binding match the header's.  No compute calls."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mmn.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mmn_[a-z_0-9]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib_path():
    from multimodn_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        subprocess.run([sys.executable, os.path.join(ROOT, "__graft_entry__.py")], check=True, cwd=ROOT)
    return _lib.LIB_PATH


def test_every_declared_symbol_is_exported(lib_path):
    dll = C.CDLL(lib_path)
    names = declared_functions()
    assert len(names) >= 12
    missing = [n for n in names if not hasattr(dll, n)]
    assert not missing, f"declared in include/mmn.h but not exported: {missing}"
    from multimodn_b200 import _lib
    assert set(_lib.EXPORTS) == set(names), set(_lib.EXPORTS) ^ set(names)
    dll.mmn_abi_version.restype = C.c_int
    assert dll.mmn_abi_version() == _lib.ABI_VERSION


def test_library_is_sm100a_with_tcgen05(lib_path):
    out = subprocess.run(["cuobjdump", "-lelf", lib_path], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    sass = subprocess.run(["cuobjdump", "-sass", lib_path], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass and "LDTM" in sass         # tcgen05.mma / tcgen05.ld of the tensor-core engines
    assert "UTMALDG" in sass                            # cp.async.bulk.tensor (TMA) of the wide-regime GEMM
    assert "UTCHMMA.2CTA" in sass and "UTCBAR.2CTA.MULTICAST" in sass      # its CTA-pair (cta_group::2) variant


def test_ctypes_struct_sizes_match_header(tmp_path, lib_path):
    from multimodn_b200 import _lib
    prog = tmp_path / "sizes.c"
    prog.write_text('#include <stdio.h>\n#include "mmn.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu\\n",'
                    "sizeof(mmn_layer_desc),sizeof(mmn_encoder_desc),sizeof(mmn_decoder_desc),sizeof(mmn_model_desc),"
                    "sizeof(mmn_batch),sizeof(mmn_outputs),sizeof(mmn_train_args));return 0;}\n")
    exe = tmp_path / "sizes"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(prog), "-o", str(exe)], check=True)
    got = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True).stdout.split()]
    want = [C.sizeof(t) for t in (_lib.LayerDesc, _lib.EncoderDesc, _lib.DecoderDesc, _lib.ModelDesc, _lib.Batch,
                                  _lib.Outputs, _lib.TrainArgs)]
    assert got == want


def test_no_cpu_fallback():
    """the product path refuses non-CUDA devices and a missing library instead of falling back"""
    import torch
    from multimodn_b200 import MultiModN, _lib
    from multimodn_b200.encoders import MLPEncoder
    from multimodn_b200.decoders import LogisticDecoder
    model = MultiModN(4, [MLPEncoder(4, 3, (5,))], [LogisticDecoder(4)], 1.0, 0.0, device=torch.device("cpu"))
    with pytest.raises(RuntimeError, match="CUDA devices only"):
        model.predict([torch.zeros(2, 3)])
    with pytest.raises(_lib.MMNError, match="no CPU fallback"):
        _lib.Library(path="/nonexistent/libmmn.so")
    import inspect
    import multimodn_b200
    for mod in ("multimodn", "plan", "_lib", "optim", "history", "metrics", "state"):
        src = inspect.getsource(getattr(__import__("multimodn_b200." + mod), mod))
        assert "oracle" not in src and "emu" not in src.replace("enumerate", ""), mod


def test_unsupported_modules_raise():
    import torch
    from multimodn_b200.plan import lower_encoder, activation_name
    with pytest.raises(NotImplementedError):
        activation_name(torch.nn.functional.gelu)

    class Rnn(torch.nn.Module):
        state_size = 4
        layers = torch.nn.ModuleList([torch.nn.RNN(3, 4)])
        activation = staticmethod(torch.relu)

    with pytest.raises(NotImplementedError):
        lower_encoder(Rnn(), 4)
