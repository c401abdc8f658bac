"""GPU: the wide regime's tcgen05 bf16 GEMM (multimodn_b200/csrc/mmn_wide.cuh) through the C ABI against
torch's fp32 matmul on the same bf16 inputs: full and ragged tiles, K tails, pitched operands, both output
orientations.  fp32 accumulation of exact bf16 products: only the summation order differs (1e-5 relative
to the row scale); the bf16 outputs are the fp32 result rounded once."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def run_gemm(M, N, K, lda=None, ldb=None, seed=0):
    from multimodn_b200 import _lib
    lib = _lib.get_lib()
    dev = torch.device("cuda")
    g = torch.Generator(device="cpu").manual_seed(seed)
    lda, ldb = lda or K, ldb or K
    a = torch.zeros(M, lda, dtype=torch.bfloat16)
    b = torch.zeros(N, ldb, dtype=torch.bfloat16)
    a[:, :K] = torch.randn(M, K, generator=g).to(torch.bfloat16)
    b[:, :K] = torch.randn(N, K, generator=g).to(torch.bfloat16)
    if lda > K:
        a[:, K:] = 7.0          # pitch padding must never be read as data
    if ldb > K:
        b[:, K:] = -3.0
    a, b = a.to(dev), b.to(dev)
    out = torch.full((M, N), float("nan"), dtype=torch.float32, device=dev)
    ob = torch.zeros(M, N, dtype=torch.bfloat16, device=dev)
    ot = torch.zeros(N, M, dtype=torch.bfloat16, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    lib.check(lib.dll.mmn_selftest_gemm_bf16(M, N, K, a.data_ptr(), lda, b.data_ptr(), ldb, out.data_ptr(), ob.data_ptr(),
                                             ot.data_ptr(), stream))
    torch.cuda.synchronize()
    want = a[:, :K].float() @ b[:, :K].float().T
    return out, ob, ot, want


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (128, 256, 256), (256, 512, 2048), (384, 768, 1024),
                                   (100, 40, 72), (8192, 2048, 2048), (130, 258, 200), (2, 2048, 2048), (2048, 2, 512),
                                   (1000, 5000, 200), (4096 + 130, 1024 + 24, 136)])     # ragged tiles on CTA pairs
def test_gemm_matches_torch(M, N, K):
    out, ob, ot, want = run_gemm(M, N, K, lda=(K + 7) // 8 * 8, ldb=(K + 7) // 8 * 8)
    scale = want.abs().max().item() + 1e-30
    assert (out - want).abs().max().item() <= 1e-5 * scale * max(1.0, (K / 64) ** 0.5)
    assert torch.equal(ob, out.to(torch.bfloat16))
    assert torch.equal(ot, out.to(torch.bfloat16).T.contiguous())


def test_gemm_pitched_operands():
    out, ob, ot, want = run_gemm(300, 520, 320, lda=512, ldb=384, seed=3)
    assert (out - want).abs().max().item() <= 1e-5 * want.abs().max().item() * 3


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (128, 256, 512), (2048, 1024, 8192), (100, 40, 72), (130, 258, 200),
                                   (2, 2048, 517), (2048, 3072, 300), (1000, 5000, 200)])
def test_gemm_mn_major_operands(M, N, K):
    """both operands contracted over their ROWS (the weight-gradient form dW = dZ^T . In without transposed copies)"""
    from multimodn_b200 import _lib
    lib = _lib.get_lib()
    dev = torch.device("cuda")
    g = torch.Generator(device="cpu").manual_seed(M + N + K)
    lda, ldb = (M + 7) // 8 * 8 + 8, (N + 7) // 8 * 8
    a = torch.full((K, lda), 5.0, dtype=torch.bfloat16)
    b = torch.full((K, ldb), -2.0, dtype=torch.bfloat16)
    a[:, :M] = torch.randn(K, M, generator=g).to(torch.bfloat16)
    b[:, :N] = torch.randn(K, N, generator=g).to(torch.bfloat16)
    a, b = a.to(dev), b.to(dev)
    out = torch.full((M, N), float("nan"), dtype=torch.float32, device=dev)
    lib.check(lib.dll.mmn_selftest_gemm_bf16_mn(M, N, K, a.data_ptr(), lda, b.data_ptr(), ldb, out.data_ptr(),
                                                torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    want = a[:, :M].float().T @ b[:, :N].float()
    assert (out - want).abs().max().item() <= 1e-5 * want.abs().max().item() * max(1.0, (K / 64) ** 0.5)

