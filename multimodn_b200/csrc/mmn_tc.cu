// mmn_tc.cu — translation unit of the tcgen05 3xTF32 self-test (mmn_tc.cuh).
#include "mmn_tc.cuh"
#include "mmn_host.h"

using namespace mmn;

// Diagnostic: one tcgen05 3xTF32 GEMM in each operand configuration of the tensor-core engine
// (mmn_tc.cuh).  a, b, out: device pointers, see mmn_tc_selftest_kernel.
extern "C" int mmn_selftest_umma(int mode, int n, const float* a, const float* b, float* out, void* stream) {
  if (mode < 0 || mode > 4 || (n != 32 && n != 64) || (mode == 2 && n != 32)) return fail("mmn_selftest_umma: bad mode / n");
  const size_t smem = 1024 + 98304 + 64;
  auto kfn = mmn_tc_selftest_kernel<0>;
  MMN_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  MMN_LAUNCH(kfn, dim3(1), dim3(256), smem, stream, mode, n, a, b, out);
  MMN_CUDA(cudaGetLastError());
  return 0;
}

