// mmn_launch.cuh — launch of the engine-templated step kernel (mmn_kernels.cuh: mmn_step_kernel<ENG, TRAIN>); included by
// the translation unit of each engine.
#pragma once

#include "mmn_host.h"

namespace {
using namespace mmn;
template <class ENG, bool TRAIN>
int launch_engine(const mmn_plan* plan, const StepArgs& a_in, void* stream) {
  StepArgs a = a_in;
  const size_t smem = step_smem_bytes(plan->host, ENG::TM, ENG::stage_bytes());
  const int grid = grid_for(plan, MMN_ENGINE_FMA, a.n_rows);
  auto kfn = mmn_step_kernel<ENG, TRAIN>;
  static bool configured = false;          // per instantiation: the attributes are set once, not on every launch
  if (!configured) {
    MMN_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, plan->max_smem));
#ifndef MMN_EMU
    // two CTAs per SM need (almost) the whole 228 KB as shared memory: ask for the largest carve-out explicitly
    MMN_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
#endif
    configured = true;
  }
#ifndef MMN_EMU
  if (getenv("MMN_DEBUG_OCC")) {
    int nb = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kfn, ENG::kBlockThreads, smem);
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, kfn);
    fprintf(stderr, "[mmn occupancy] %d CTAs/SM possible (smem %zu B dynamic + %zu static, %d regs, %zu B local/thread), grid %d\n", nb, smem,
            fa.sharedSizeBytes, fa.numRegs, fa.localSizeBytes, grid);
  }
#endif
  const bool dbg = getenv("MMN_DEBUG_TIMERS") != nullptr;      // development aid: per-phase cycle counters
  if (dbg) { MMN_CUDA(cudaMalloc((void**)&a.debug_timers, sizeof(long long) * 32 * grid)); MMN_CUDA(cudaMemsetAsync(a.debug_timers, 0, sizeof(long long) * 32 * grid, (cudaStream_t)stream)); }
  MMN_LAUNCH(kfn, dim3(grid), dim3(ENG::kBlockThreads), smem, stream, a);
  MMN_CUDA(cudaGetLastError());
  if (dbg) {
    std::vector<long long> h(32 * (size_t)grid);
    MMN_CUDA(cudaMemcpy(h.data(), a.debug_timers, sizeof(long long) * h.size(), cudaMemcpyDeviceToHost));
    cudaFree(a.debug_timers);
    double s[16] = {0};
    for (int b = 0; b < grid; ++b) for (int i = 0; i < 16; ++i) s[i] += (double)h[b * 16 + i] / grid;
    fprintf(stderr, "[mmn timers, mean cycles/CTA] total %.0f | wait_done %.0f | nt %.0f (epi %.0f) | nn %.0f (epi %.0f) | tn %.0f (epi %.0f) | bwd %.0f | nt-store %.0f | post %.0f | bias %.0f | wsync %.0f | tn-dz %.0f\n",
            s[15], s[0], s[1], s[4], s[2], s[5], s[3], s[6], s[9], s[10], s[11], s[12], s[13], s[14]);
    double q[6] = {0};
    for (int b = 0; b < grid; ++b) for (int i = 0; i < 6; ++i) q[i] += (double)h[(grid + b) * 16 + i] / grid;
    fprintf(stderr, "[mmn issuer, mean/CTA] idle %.0f | issue %.0f | chain(nj<16) %.0f cycles x %.0f = %.0f each | chain(nj=16) %.0f x %.0f = %.0f each\n",
            q[0], q[1], q[2], q[3], q[3] ? q[2] / q[3] : 0.0, q[4], q[5], q[5] ? q[4] / q[5] : 0.0);
  }
  return 0;
}
}  // namespace
