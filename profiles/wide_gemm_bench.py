"""Throughput of the wide regime's tcgen05 bf16 GEMM at config-4 layer shapes (B = 8192 rows per GPU).
usage: python profiles/wide_gemm_bench.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodn_b200 import _lib

lib = _lib.get_lib()
dev = torch.device("cuda")
stream = torch.cuda.current_stream().cuda_stream
for (M, N, K, what) in [(8192, 2048, 2048, "fwd hidden 2048->2048"), (8192, 1024, 2048, "fwd 2048->state 1024"),
                        (2048, 2048, 8192, "wgrad 2048x2048 over 8192 rows"), (8192, 2048, 1024, "decoder 1024->2048")]:
    a = torch.randn(M, K, device=dev).to(torch.bfloat16)
    b = torch.randn(N, K, device=dev).to(torch.bfloat16)
    ob = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
    ot = torch.empty(N, M, dtype=torch.bfloat16, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    def timed(with_t):
        # with_t: also store the transposed copy (a round-1 layout; the step no longer asks for it, Arena::want_t = false)
        def run():
            lib.check(lib.dll.mmn_selftest_gemm_bf16(M, N, K, a.data_ptr(), K, b.data_ptr(), K, None, ob.data_ptr(),
                                                     ot.data_ptr() if with_t else None, stream))
        for _ in range(3): run()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(20): run()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / 20
    ms, ms_t = timed(False), timed(True)
    for _ in range(3): torch.matmul(a, b.T)
    e0.record()
    for _ in range(20): torch.matmul(a, b.T)
    e1.record(); torch.cuda.synchronize()
    ms_ref = e0.elapsed_time(e1) / 20
    fl = 2.0 * M * N * K
    print(f"{what:34s} M={M} N={N} K={K}: {ms*1e3:7.1f} us  {fl/ms/1e9:7.1f} TFLOP/s   (+ transposed copy {ms_t*1e3:7.1f} us; "
          f"cuBLAS {ms_ref*1e3:7.1f} us, {fl/ms_ref/1e9:7.1f} TFLOP/s)")
