// mmn_common.cuh — device-side plan, kernel argument block and build glue shared by the
// kernels (mmn_kernels.cuh) and the C ABI (mmn_api.cu).
#pragma once

#include "mmn.h"

#ifdef MMN_EMU
#include "cuda_emu.h"   // tests/emu: CPU emulation of the CUDA subset used here (test-only build)
#else
#include <cuda_runtime.h>
#define MMN_LAUNCH(kernel, grid, block, smem, stream, ...) \
  kernel<<<(grid), (block), (smem), (cudaStream_t)(stream)>>>(__VA_ARGS__)
#define MMN_DYN_SMEM(name) extern __shared__ __align__(16) char name[]
// barrier over the n worker threads only (the tensor-core engine adds one more, MMA-issuing warp)
#define MMN_WSYNC_N(n) asm volatile("bar.sync 1, %0;" ::"n"(n) : "memory")
#define MMN_CLOCK() clock64()
#endif

namespace mmn {

constexpr int kThreads = 256;   // threads per CTA: 32 row-threads x 8 column-threads
constexpr int KC = 32;          // K chunk staged through shared memory
constexpr int LDX = KC + 4;     // leading dim of staged chunks: == 4 (mod 32) -> conflict-free LDS.128
constexpr int kGroups = kThreads / 64;   // row-split groups of the weight-gradient GEMM

struct DevLayer {
  int in_dim, out_dim, act, has_state;
  int ktot;        // in_dim + has_state * S  (row length of W)
  int stash_off;   // per-row float offset of this layer's OUTPUT inside its stash group
  long long w_off, b_off;
};
struct DevEncoder {
  int F, n_layers;
  float p_drop;
  int param_lo, param_hi;   // [lo, hi) range of this encoder's parameters in the packed buffer
  DevLayer L[MMN_MAX_LAYERS];
};
struct DevDecoder {
  int C, n_layers;
  int out_off;     // column of this decoder's outputs in mmn_outputs.last_outputs
  int stash_off;   // per-row float offset of this decoder's group inside a step's decoder stash
  DevLayer L[MMN_MAX_LAYERS];
};
struct DevPlan {
  int S, E, D;
  int ldS, ldH;         // shared-memory leading dims of state-wide / hidden-wide tiles
  int enc_stash;        // floats per row per step: encoder hidden-layer outputs
  int dec_stash;        // floats per row per step: every decoder layer output
  int stash_row;        // floats per row for the whole chain
  int sumC;
  int n_metrics;
  long long init_off, n_params;
  DevEncoder enc[MMN_MAX_ENCODERS];
  DevDecoder dec[MMN_MAX_DECODERS];
};

// stash group offsets (per-row floats); a block lives at slot + off * TM as [TM x width] row-major
__host__ __device__ inline int stash_state_off(const DevPlan& p, int k) { return k * p.S; }
__host__ __device__ inline int stash_enc_off(const DevPlan& p, int k /*1-based step*/) {
  return (p.E + 1) * p.S + (k - 1) * p.enc_stash;
}
__host__ __device__ inline int stash_dec_off(const DevPlan& p, int k) {
  return (p.E + 1) * p.S + p.E * p.enc_stash + k * p.dec_stash;
}

// metrics layout (doubles), see mmn.h
__host__ __device__ inline int met_mat(const DevPlan& p, int which, int row, int d) {
  return which * (p.E + 1) * p.D + row * p.D + d;
}
__host__ __device__ inline int met_present(const DevPlan& p, int row) { return 6 * (p.E + 1) * p.D + row; }
__host__ __device__ inline int met_sc(const DevPlan& p, int e) { return 6 * (p.E + 1) * p.D + (p.E + 1) + e; }

struct StepArgs {
  const DevPlan* plan;
  const float* params;
  float* grads;
  float* stash;
  long long slot_floats;      // stash floats per CTA slot
  long long n_rows, row_offset;
  double inv_rows_global;
  int seq_len;
  int training;
  int seq_pos[MMN_MAX_ENCODERS];
  int seq_enc[MMN_MAX_ENCODERS];
  const float* x[MMN_MAX_ENCODERS];
  long long x_ld[MMN_MAX_ENCODERS];
  const long long* targets;
  const int* skip_flags;
  double* metrics;
  unsigned char* predictions;
  long long pred_ld;
  float* last_outputs;
  float* final_state;
  int* target_error;           // device flag: a target outside [0, C) was seen (the index is clamped for memory safety)
  float c_err;   // err_penalty / (D (E+1) B_global)
  float c_sc;    // 2 * state_change_penalty_scaled / (E B_global S)
  unsigned dropout_seed;
  long long* debug_timers;   // optional [grid][16] cycle counters (MMN_DEBUG_TIMERS=1), thread 0 of each CTA
};

// shared-memory footprint of the step kernel for a row tile of TM rows, given the engine's staging bytes
inline size_t step_smem_bytes(const DevPlan& p, int TM_, size_t stage_bytes) {
  const size_t TM = (size_t)TM_;
  size_t bytes = stage_bytes + (2 * TM * p.ldS + 2 * TM * p.ldH) * 4;   // staging, S/G, T, A, B
  bytes += TM * p.D * 4;                 // targets tile
  bytes += TM * 4;                       // row NaN flags
  bytes += (size_t)(p.E + 1) * 8;        // present-row counters + tile_any (ints)
  bytes = (bytes + 7) & ~(size_t)7;
  bytes += (size_t)p.n_metrics * 8;      // metric accumulators (double)
  bytes += (size_t)(p.E + 1) * TM;       // present masks
  return (bytes + 15) & ~(size_t)15;
}

}  // namespace mmn
