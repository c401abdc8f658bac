"""tcgen05 3xTF32 GEMM self-test (mmn_selftest_umma) shared by the CPU-emulator and GPU suites."""
import ctypes as C

import numpy as np
import torch


def run_selftest(lib, device, mode, n, seed=0):
    rng = np.random.default_rng(seed + 10 * mode + n)
    if mode in (0, 3):
        a, b = rng.standard_normal((128, 32)), rng.standard_normal((n, 32))
        ref = a @ b.T
    elif mode in (1, 4):
        a, b = rng.standard_normal((128, 32)), rng.standard_normal((32, n))
        ref = a @ b
    else:
        a, b = rng.standard_normal((128, 64)), rng.standard_normal((128, 32))
        ref = a.T @ b                      # (64, 32): rows >= 64 of the 128-row output are padding
    ta = torch.tensor(a, dtype=torch.float32, device=device).contiguous()
    tb = torch.tensor(b, dtype=torch.float32, device=device).contiguous()
    out = torch.zeros((128, 32 if mode == 2 else n), dtype=torch.float32, device=device)
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream) if device != "cpu" else C.c_void_p(0)
    lib.check(lib.dll.mmn_selftest_umma(mode, n, ta.data_ptr(), tb.data_ptr(), out.data_ptr(), stream))
    if device != "cpu":
        torch.cuda.synchronize()
    got = out.cpu().numpy().astype(np.float64)
    if mode == 2:
        got = got[:64]
    ref32 = (ta.cpu().numpy().astype(np.float64), tb.cpu().numpy().astype(np.float64))
    if mode in (0, 3):
        ref = ref32[0] @ ref32[1].T
    elif mode in (1, 4):
        ref = ref32[0] @ ref32[1]
    else:
        ref = ref32[0].T @ ref32[1]
    err = np.abs(got - ref).max() / np.abs(ref).max()
    return err
