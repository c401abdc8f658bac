// mmn_host.h — host-side declarations shared by the translation units of libmmn.so.  Every kernel family is compiled
// in its own translation unit (mmn_fma.cu, mmn_tc.cu, mmn_tc2.cu, mmn_wide.cu): the device code nvcc generates for one
// kernel must not depend on which other kernels happen to share its compilation (inlining decisions are made per
// module — the FP32-FMA step kernel lost 40 % when the TMEM-resident backward pass joined its translation unit).
#pragma once

#include "mmn_common.cuh"

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

struct mmn_plan {
  mmn::DevPlan host;
  mmn::DevPlan* dev = nullptr;
  int n_sms = 0;
  int max_smem = 0;
  int engine = MMN_ENGINE_FMA;   // MMN_ENGINE_*
  int rm = 0;                    // FMA engine: rows per tile / 32
  int occ = 1;                   // FMA engine: CTAs per SM the kernel variant is built for
  int fwd_engine = MMN_ENGINE_FMA;   // engine of the forward-only path (test / predict / get_states)
  // wide regime (precision = bf16): bf16 copies of every weight in both orientations, refreshed every call
  void* wide_w = nullptr;
  long long wide_elems = 0;
  struct WL { long long w, wt; int ldk, ldo; };
  WL wide_enc[MMN_MAX_ENCODERS][MMN_MAX_LAYERS];
  WL wide_dec[MMN_MAX_DECODERS][MMN_MAX_LAYERS];
  // > 0: the transposed first-layer images of all decoders sit side by side in one [S x dec_cat_k] matrix (pitch dec_cat_k;
  // decoder d's columns start at dec_cat_col[d]), so sum_d dz0_d . W0_d is one GEMM with K = dec_cat_k
  int dec_cat_k = 0;
  int dec_cat_col[MMN_MAX_DECODERS] = {};
  // layer-wise plans: an internal side stream (forked from / joined to the caller's stream with events) on which the small
  // bias-gradient reductions run concurrently with the weight-gradient GEMMs
  void* side_stream = nullptr;
  void* side_fork = nullptr;     // cudaEvent_t: caller's stream -> side stream
  void* side_done = nullptr;     // cudaEvent_t: side stream -> caller's stream
  void* dec_stream = nullptr;    // second side stream: the decoders' backward chains, next to the encoders' backward GEMMs
  void* dec_fork = nullptr;
  void* dec_done[MMN_MAX_ENCODERS + 1] = {};   // cudaEvent_t per step: that step's decoder chains are done
  // bf16 tile kernel (mmn_nb.cuh; narrow models under precision = bf16): lowered plan (host / device copies) and the device
  // arena the per-step imaging kernel fills with bf16 weight images
  void* nb_host = nullptr;
  void* nb_dev = nullptr;
  void* nb_arena = nullptr;
  // optional cudaEvent_t handles recorded by mmn_train_step as gradient blocks become final (mmn_plan_set_grad_events)
  // layout: [0, E) encoder e complete; E = everything; E + 1 = decoders; E + 2 + (layers before (e, j)) = encoder e's layer j
  void* grad_events[MMN_MAX_ENCODERS + 2 + MMN_MAX_ENCODERS * MMN_MAX_LAYERS] = {};
  int n_grad_events = 0;
  int comm_sms = 0;                 // mmn_plan_set_comm_sms; 0 = default
  int grad_layer_event(int e, int j) const {        // index of the per-layer event, -1 if the caller did not ask for them
    int at = host.E + 2;
    for (int i = 0; i < e; ++i) at += host.enc[i].n_layers;
    return at + j < n_grad_events ? at + j : -1;
  }
};

int mmn_fail(const char* fmt, ...);       // records the calling thread's error message, returns 1
template <class... Args>
static inline int fail(const char* fmt, Args... args) { return mmn_fail(fmt, args...); }
#define MMN_CUDA(call)                                                            \
  do {                                                                            \
    cudaError_t e_ = (call);                                                      \
    if (e_ != cudaSuccess) return fail("%s: %s", #call, cudaGetErrorString(e_));  \
  } while (0)

inline int tile_rows(const mmn_plan* p, int engine) { return engine != MMN_ENGINE_FMA ? 128 : 32 * p->rm; }
inline int grid_for(const mmn_plan* p, int engine, int64_t n_rows) {
  const int tm = tile_rows(p, engine);
  const int64_t tiles = (n_rows + tm - 1) / tm;
  const int per_sm = engine == MMN_ENGINE_FMA ? p->occ : 1;
  return (int)std::max<int64_t>(1, std::min<int64_t>(tiles, (int64_t)p->n_sms * per_sm));
}

// per-family entry points (defined next to the kernels they launch)
size_t mmn_fma_smem(const mmn::DevPlan& P, int rm, int occ);
bool mmn_v2_supports(const mmn::DevPlan& P);
size_t mmn_v2_smem(const mmn::DevPlan& P);
int mmn_launch_fma(const mmn_plan* plan, const mmn::StepArgs& a, void* stream, bool train);
int mmn_launch_v2(const mmn_plan* plan, const mmn::StepArgs& a, void* stream, bool train);
int mmn_nb_plan_init(mmn_plan* p);
void mmn_nb_plan_free(mmn_plan* p);
int64_t mmn_nb_workspace_bytes(const mmn_plan* plan, int64_t n_rows, bool train);
int mmn_nb_step(const mmn_plan* plan, const mmn::StepArgs& a, void* ws, size_t ws_bytes, void* stream, bool train);
#ifndef MMN_EMU
int mmn_wide_plan_init(mmn_plan* p);
int64_t mmn_wide_workspace_bytes(const mmn_plan* plan, int64_t n_rows, bool train);
int mmn_wide_step(const mmn_plan* plan, const mmn::StepArgs& a, void* ws, size_t ws_bytes, void* stream, bool train);
#endif
