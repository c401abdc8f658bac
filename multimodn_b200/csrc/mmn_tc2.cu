// mmn_tc2.cu — translation unit of the TMEM-resident tcgen05 kernel (mmn_tc2.cuh).
#include "mmn_tc2.cuh"
#include "mmn_host.h"

using namespace mmn;

bool mmn_v2_supports(const DevPlan& P) { return V2Engine::supports(P); }
size_t mmn_v2_smem(const DevPlan& P) { return V2Engine::smem_bytes(P); }

namespace {
template <bool TRAIN>
int launch_v2(const mmn_plan* plan, const StepArgs& a_in, void* stream) {
  StepArgs a = a_in;
  const size_t smem = V2Engine::smem_bytes(plan->host);
  const int grid = grid_for(plan, MMN_ENGINE_TC2, a.n_rows);
  auto kfn = mmn_step_kernel_v2<TRAIN>;
  MMN_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const bool dbg = getenv("MMN_DEBUG_TIMERS") != nullptr;
  if (dbg) { MMN_CUDA(cudaMalloc((void**)&a.debug_timers, sizeof(long long) * 16 * grid)); MMN_CUDA(cudaMemsetAsync(a.debug_timers, 0, sizeof(long long) * 16 * grid, (cudaStream_t)stream)); }
  MMN_LAUNCH(kfn, dim3(grid), dim3(V2Engine::kBlockThreads), smem, stream, a);
  MMN_CUDA(cudaGetLastError());
  if (dbg) {
    std::vector<long long> h(16 * (size_t)grid);
    MMN_CUDA(cudaMemcpy(h.data(), a.debug_timers, sizeof(long long) * h.size(), cudaMemcpyDeviceToHost));
    cudaFree(a.debug_timers);
    double s[16] = {0};
    for (int b = 0; b < grid; ++b) for (int i = 0; i < 16; ++i) s[i] += (double)h[b * 16 + i] / grid;
    fprintf(stderr, "[mmn v2 timers, mean cycles/CTA] total %.0f | gemms %.0f (%.0f calls, %.0f chunks) | slot-wait %.0f | W stage %.0f | prefetch+post %.0f | bias+mid %.0f | acc-wait %.0f | epilogue %.0f (dec hidden %.0f, dec metrics %.0f) | backward %.0f: colsum %.0f, wgrad %.0f, dgrad %.0f\n",
            s[15], s[6], s[8], s[7], s[0], s[1], s[2], s[3], s[4], s[5], s[9], s[10], s[14], s[11], s[12], s[13]);
  }
  return 0;
}
}  // namespace

int mmn_launch_v2(const mmn_plan* plan, const StepArgs& a, void* stream, bool train) {
  return train ? launch_v2<true>(plan, a, stream) : launch_v2<false>(plan, a, stream);
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


