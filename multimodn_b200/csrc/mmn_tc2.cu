// mmn_tc2.cu — translation unit of the TMEM-resident tcgen05 kernel (mmn_tc2.cuh).
#include "mmn_tc2.cuh"
#include "mmn_host.h"

using namespace mmn;

bool mmn_v2_supports(const DevPlan& P) { return V2Engine::supports(P); }
size_t mmn_v2_smem(const DevPlan& P) { return V2Engine::smem_bytes(P); }

int mmn_launch_v2(const mmn_plan* plan, const StepArgs& a_in, void* stream, bool train) {
  if (train) return fail("the TMEM-resident kernel is forward-only (test / predict / get_states)");
  StepArgs a = a_in;
  const size_t smem = V2Engine::smem_bytes(plan->host);
  const int grid = grid_for(plan, MMN_ENGINE_TC2, a.n_rows);
  auto kfn = mmn_forward_kernel_v2<0>;
  static bool configured = false;
  if (!configured) {
    MMN_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan->max_smem));
    configured = true;
  }
  MMN_LAUNCH(kfn, dim3(grid), dim3(V2Engine::kBlockThreads), smem, stream, a);
  MMN_CUDA(cudaGetLastError());
  return 0;
}

// Development aid: cycles per round of the worker <-> MMA-issuer handshake (mmn_tc2.cuh).  out: device int64[2].
extern "C" int mmn_selftest_protocol(int iters, int n_mma, int flags, long long* out, void* stream) {
  auto kfn = mmn_protocol_probe_kernel<0>;
  const size_t smem = 1024 + 32768 + 64;
  MMN_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  MMN_LAUNCH(kfn, dim3(1), dim3(288), smem, stream, iters, n_mma, flags, out);
  MMN_CUDA(cudaGetLastError());
  return 0;
}


