import os, sys
sys.path.insert(0, "/root/repo")
import torch
from multimodn_b200 import _lib
lib = _lib.get_lib(); dev = torch.device("cuda"); stream = torch.cuda.current_stream().cuda_stream
for (M, N, K) in [(8192, 2048, 1024), (8192, 2048, 2048)]:
    a = torch.randn(M, K, device=dev).to(torch.bfloat16); b = torch.randn(N, K, device=dev).to(torch.bfloat16)
    of = torch.empty(M, N, dtype=torch.float32, device=dev)
    ob = torch.empty(M, N, dtype=torch.bfloat16, device=dev); ot = torch.empty(N, M, dtype=torch.bfloat16, device=dev)
    for name, (pf, pb, pt) in {"bf16 + bf16^T": (None, ob, ot), "bf16 only": (None, ob, None), "bf16^T only": (None, None, ot),
                               "fp32 only": (of, None, None), "nothing": (None, None, None)}.items():
        run = lambda: lib.check(lib.dll.mmn_selftest_gemm_bf16(M, N, K, a.data_ptr(), K, b.data_ptr(), K,
                      pf.data_ptr() if pf is not None else None, pb.data_ptr() if pb is not None else None, pt.data_ptr() if pt is not None else None, stream))
        for _ in range(3): run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): run()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print(f"M={M} N={N} K={K} {name:16s}: {ms*1e3:6.1f} us {2.0*M*N*K/ms/1e9:7.1f} TFLOP/s")
