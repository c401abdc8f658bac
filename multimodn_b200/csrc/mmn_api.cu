// mmn_api.cu — the C ABI of libmmn.so (include/mmn.h): plan construction, launch configuration and
// argument marshalling around the kernels in mmn_kernels.cuh.  No torch types, no hidden syncs.
#include "mmn_host.h"

#include <cstdarg>
#include <string>

using namespace mmn;

namespace {
thread_local std::string g_err;

}  // namespace
int mmn_fail(const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return 1;
}
namespace {
int round32(int v) { return (v + 31) & ~31; }
size_t fma_smem(const DevPlan& P, int rm, int occ = 1) { return mmn_fma_smem(P, rm, occ); }
// largest FMA row tile (32*RM rows) whose shared-memory footprint fits
int pick_rm(const mmn_plan* p) {
  for (int rm : {4, 2, 1})
    if (fma_smem(p->host, rm) <= (size_t)p->max_smem) return rm;
  return 0;
}

int check_layer(const mmn_layer_desc& l, int S, int64_t n_params, const char* what, int idx, int j) {
  if (l.in_dim < 0 || l.out_dim <= 0) return fail("%s %d layer %d: bad dims", what, idx, j);
  if (l.act < MMN_ACT_IDENTITY || l.act > MMN_ACT_TANH) return fail("%s %d layer %d: unsupported activation", what, idx, j);
  const int64_t ktot = l.in_dim + (l.has_state ? S : 0);
  if (ktot <= 0) return fail("%s %d layer %d: empty input", what, idx, j);
  if (l.w_off < 0 || l.b_off < 0 || l.w_off + ktot * l.out_dim > n_params || l.b_off + l.out_dim > n_params)
    return fail("%s %d layer %d: parameter offsets out of range", what, idx, j);
  if ((l.w_off & 3) || (l.b_off & 3)) return fail("%s %d layer %d: offsets must be multiples of 4 floats", what, idx, j);
  return 0;
}
}  // namespace

extern "C" const char* mmn_last_error(void) { return g_err.c_str(); }
extern "C" int mmn_abi_version(void) { return MMN_ABI_VERSION; }

extern "C" int mmn_plan_create(const mmn_model_desc* desc, mmn_plan** out) {
  if (!desc || !out) return fail("mmn_plan_create: null argument");
  const int S = desc->state_size, E = desc->n_encoders, D = desc->n_decoders;
  if (S <= 0) return fail("state_size must be positive");
  if (E <= 0 || E > MMN_MAX_ENCODERS) return fail("n_encoders must be in [1, %d]", MMN_MAX_ENCODERS);
  if (D <= 0 || D > MMN_MAX_DECODERS) return fail("n_decoders must be in [1, %d]", MMN_MAX_DECODERS);
  if (desc->init_off < 0 || desc->init_off + S > desc->n_params || (desc->init_off & 3))
    return fail("init_off out of range or not a multiple of 4");
  if (desc->n_params >= (1ll << 31)) return fail("n_params too large");

  mmn_plan* p = new mmn_plan();
  DevPlan& P = p->host;
  memset(&P, 0, sizeof P);
  P.S = S; P.E = E; P.D = D;
  P.init_off = desc->init_off;
  P.n_params = desc->n_params;
  int maxH = 1;
  P.enc_stash = 0;
  for (int e = 0; e < E; ++e) {
    const mmn_encoder_desc& src = desc->encoders[e];
    DevEncoder& dst = P.enc[e];
    if (src.n_layers < 1 || src.n_layers > MMN_MAX_LAYERS) { delete p; return fail("encoder %d: n_layers must be in [1, %d]", e, MMN_MAX_LAYERS); }
    if (!(src.dropout_p >= 0.f && src.dropout_p < 1.f)) { delete p; return fail("encoder %d: dropout must be in [0, 1)", e); }
    dst.F = src.n_features; dst.n_layers = src.n_layers; dst.p_drop = src.dropout_p;
    int n_state = 0, off = 0;
    long long lo = desc->n_params, hi = 0;
    for (int j = 0; j < src.n_layers; ++j) {
      const mmn_layer_desc& l = src.layers[j];
      if (check_layer(l, S, desc->n_params, "encoder", e, j)) { delete p; return 1; }
      DevLayer& d = dst.L[j];
      d.in_dim = l.in_dim; d.out_dim = l.out_dim; d.act = l.act; d.has_state = l.has_state ? 1 : 0;
      d.ktot = l.in_dim + (l.has_state ? S : 0);
      d.w_off = l.w_off; d.b_off = l.b_off;
      n_state += d.has_state;
      if (j == 0 && l.in_dim != src.n_features) { delete p; return fail("encoder %d: layer 0 in_dim != n_features", e); }
      if (j > 0 && l.in_dim != src.layers[j - 1].out_dim) { delete p; return fail("encoder %d: layer %d in_dim mismatch", e, j); }
      if (j < src.n_layers - 1) { d.stash_off = off; off += l.out_dim; maxH = std::max(maxH, l.out_dim); }
      lo = std::min<long long>(lo, std::min(l.w_off, l.b_off));
      hi = std::max<long long>(hi, std::max<long long>(l.w_off + (long long)d.ktot * l.out_dim, l.b_off + l.out_dim));
    }
    if (src.layers[src.n_layers - 1].out_dim != S) { delete p; return fail("encoder %d: last layer must produce the state", e); }
    if (n_state != 1) { delete p; return fail("encoder %d: exactly one layer must take the state", e); }
    if (src.dropout_p > 0.f && !src.layers[0].has_state) { delete p; return fail("encoder %d: dropout needs the state on layer 0 (MIMIC_MLPEncoder)", e); }
    dst.param_lo = (int)lo; dst.param_hi = (int)hi;
    P.enc_stash = std::max(P.enc_stash, off);
  }
  for (int a = 0; a < E; ++a)
    for (int b = a + 1; b < E; ++b)
      if (P.enc[a].param_lo < P.enc[b].param_hi && P.enc[b].param_lo < P.enc[a].param_hi) {
        delete p;
        return fail("encoders %d and %d: parameter ranges overlap", a, b);
      }
  P.dec_stash = 0; P.sumC = 0;
  for (int d = 0; d < D; ++d) {
    const mmn_decoder_desc& src = desc->decoders[d];
    DevDecoder& dst = P.dec[d];
    if (src.n_layers < 1 || src.n_layers > MMN_MAX_LAYERS) { delete p; return fail("decoder %d: n_layers must be in [1, %d]", d, MMN_MAX_LAYERS); }
    if (src.n_classes < 1 || src.n_classes > MMN_MAX_CLASSES) { delete p; return fail("decoder %d: n_classes must be in [1, %d]", d, MMN_MAX_CLASSES); }
    dst.C = src.n_classes; dst.n_layers = src.n_layers; dst.out_off = P.sumC; dst.stash_off = P.dec_stash;
    int off = 0;
    for (int j = 0; j < src.n_layers; ++j) {
      const mmn_layer_desc& l = src.layers[j];
      if (check_layer(l, S, desc->n_params, "decoder", d, j)) { delete p; return 1; }
      if (l.has_state) { delete p; return fail("decoder %d: has_state is an encoder concept", d); }
      if (l.in_dim != (j == 0 ? S : src.layers[j - 1].out_dim)) { delete p; return fail("decoder %d: layer %d in_dim mismatch", d, j); }
      DevLayer& dl = dst.L[j];
      dl.in_dim = l.in_dim; dl.out_dim = l.out_dim; dl.act = l.act; dl.has_state = 0; dl.ktot = l.in_dim;
      dl.w_off = l.w_off; dl.b_off = l.b_off; dl.stash_off = off;
      off += l.out_dim;
      maxH = std::max(maxH, l.out_dim);
    }
    if (src.layers[src.n_layers - 1].out_dim != src.n_classes) { delete p; return fail("decoder %d: last layer must produce n_classes", d); }
    P.dec_stash += off;
    P.sumC += src.n_classes;
  }
  P.ldS = round32(S) + 4;
  P.ldH = round32(maxH) + 4;
  P.stash_row = (E + 1) * S + E * P.enc_stash + (E + 1) * P.dec_stash;
  P.n_metrics = 6 * (E + 1) * D + (E + 1) + E;

  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&p->n_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&p->max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess) {
    delete p;
    return fail("mmn_plan_create: no CUDA device");
  }
  if (desc->precision != MMN_PRECISION_FP32 && desc->precision != MMN_PRECISION_BF16) { delete p; return fail("precision must be MMN_PRECISION_FP32 or MMN_PRECISION_BF16"); }
  if (desc->precision == MMN_PRECISION_BF16) {
    // narrow models: the fused per-tile mma kernel (one launch per step); MMN_ENGINE=wide forces the layer-wise regime
    const char* want_bf16 = getenv("MMN_ENGINE");
    const bool force_wide = want_bf16 && !strcmp(want_bf16, "wide");
    if (!force_wide && mmn_nb_plan_init(p) == 0) {
      p->engine = p->fwd_engine = MMN_ENGINE_NB;
      if (cudaMalloc((void**)&p->dev, sizeof(DevPlan)) != cudaSuccess ||
          cudaMemcpy(p->dev, &P, sizeof(DevPlan), cudaMemcpyHostToDevice) != cudaSuccess) {
        mmn_nb_plan_free(p);
        delete p;
        return fail("mmn_plan_create: device allocation failed");
      }
      *out = p;
      return 0;
    }
    mmn_nb_plan_free(p);
#ifdef MMN_EMU
    delete p;
    return fail("precision = bf16: only the per-tile kernel (narrow models) is part of the host emulator: %s", g_err.c_str());
#else
    g_err.clear();
    p->engine = p->fwd_engine = MMN_ENGINE_WIDE;
    if (mmn_wide_plan_init(p) || cudaMalloc((void**)&p->dev, sizeof(DevPlan)) != cudaSuccess ||
        cudaMemcpy(p->dev, &P, sizeof(DevPlan), cudaMemcpyHostToDevice) != cudaSuccess) {
      if (p->wide_w) cudaFree(p->wide_w);
      delete p;
      return g_err.empty() ? fail("mmn_plan_create: device allocation failed") : 1;
    }
    *out = p;
    return 0;
#endif
  }
  p->rm = pick_rm(p);
  // two 64-row CTAs per SM (16 resident warps) when both fit: 228 KB of shared memory per SM, 1 KB reserved per CTA
  const char* occ_env = getenv("MMN_FMA_OCC");        // "1" | "2" | unset = auto
  const bool occ2_fits = fma_smem(P, 2, 2) + 1024 <= (size_t)(233472 / 2);
  if (occ2_fits && !(occ_env && !strcmp(occ_env, "1"))) { p->rm = 2; p->occ = 2; }
  // fp32 plans: the FP32-FMA kernel trains; forward-only launches (test / predict / get_states) use the TMEM-resident
  // tcgen05 3xTF32 kernel where the model qualifies (MMN_ENGINE=fma keeps them on the FMA kernel)
  const char* want = getenv("MMN_ENGINE");            // "fma" | "tc2" | unset = auto
  p->engine = MMN_ENGINE_FMA;
  const bool v2_ok = mmn_v2_supports(P) && mmn_v2_smem(P) <= (size_t)p->max_smem;
  if (want && !strcmp(want, "tc2") && !v2_ok) {
    delete p;
    return fail("MMN_ENGINE=tc2: the TMEM-resident kernel needs state <= 64, layers <= 64 wide, <= 16 classes");
  }
  p->fwd_engine = (v2_ok && !(want && !strcmp(want, "fma"))) ? MMN_ENGINE_TC2 : MMN_ENGINE_FMA;
  if (p->engine == MMN_ENGINE_FMA && p->rm == 0) {
    const size_t need = fma_smem(P, 1);
    delete p;
    return fail("model too wide for the fused fp32 step kernel (needs %zu B of shared memory per 32-row tile): use precision = bf16, "
                "the layer-wise tensor-core regime", need);
  }
  if (cudaMalloc((void**)&p->dev, sizeof(DevPlan)) != cudaSuccess ||
      cudaMemcpy(p->dev, &P, sizeof(DevPlan), cudaMemcpyHostToDevice) != cudaSuccess) {
    delete p;
    return fail("mmn_plan_create: device allocation failed");
  }
  *out = p;
  return 0;
}

extern "C" void mmn_plan_destroy(mmn_plan* plan) {
  if (!plan) return;
  if (plan->dev) cudaFree(plan->dev);
  if (plan->wide_w) cudaFree(plan->wide_w);
  mmn_nb_plan_free(plan);
#ifndef MMN_EMU
  if (plan->side_stream) cudaStreamDestroy((cudaStream_t)plan->side_stream);
  if (plan->side_fork) cudaEventDestroy((cudaEvent_t)plan->side_fork);
  if (plan->side_done) cudaEventDestroy((cudaEvent_t)plan->side_done);
  if (plan->dec_stream) cudaStreamDestroy((cudaStream_t)plan->dec_stream);
  if (plan->dec_fork) cudaEventDestroy((cudaEvent_t)plan->dec_fork);
  for (void* ev : plan->dec_done)
    if (ev) cudaEventDestroy((cudaEvent_t)ev);
#endif
  delete plan;
}

extern "C" int64_t mmn_metrics_count(const mmn_plan* plan) { return plan ? plan->host.n_metrics : -1; }
extern "C" int64_t mmn_grad_count(const mmn_plan* plan) { return plan ? plan->host.n_params + plan->host.E : -1; }

extern "C" int mmn_plan_set_grad_events(mmn_plan* plan, void* const* events, int32_t n) {
  if (!plan) return fail("mmn_plan_set_grad_events: null plan");
  const int E = plan->host.E;
  int n_layers = 0;
  for (int e = 0; e < E; ++e) n_layers += plan->host.enc[e].n_layers;
  if (n != 0 && n != E + 1 && n != E + 2 && n != E + 2 + n_layers)
    return fail("mmn_plan_set_grad_events: expected %d (one per encoder + one), %d (+ decoders) or %d (+ one per encoder layer) events, got %d",
                E + 1, E + 2, E + 2 + n_layers, n);
  if (n && !events) return fail("mmn_plan_set_grad_events: null event array");
  for (int i = 0; i < n; ++i) {
    if (!events[i]) return fail("mmn_plan_set_grad_events: event %d is null", i);
    plan->grad_events[i] = events[i];
  }
  plan->n_grad_events = n;
  return 0;
}

extern "C" int mmn_plan_set_comm_sms(mmn_plan* plan, int32_t n_sms) {
  if (!plan) return fail("mmn_plan_set_comm_sms: null plan");
  if (n_sms < 0) return fail("mmn_plan_set_comm_sms: negative SM count");
  plan->comm_sms = n_sms;
  return 0;
}

extern "C" int32_t mmn_plan_engine(const mmn_plan* plan) { return plan ? plan->engine : -1; }
extern "C" int32_t mmn_plan_forward_engine(const mmn_plan* plan) { return plan ? plan->fwd_engine : -1; }


extern "C" int64_t mmn_workspace_bytes(const mmn_plan* plan, int64_t n_rows, int32_t with_backward) {
  if (!plan || n_rows < 0) return -1;
  if (plan->engine == MMN_ENGINE_NB) return mmn_nb_workspace_bytes(plan, n_rows, with_backward != 0);
#ifndef MMN_EMU
  if (plan->engine == MMN_ENGINE_WIDE) return mmn_wide_workspace_bytes(plan, n_rows, with_backward != 0);
#endif
  if (!with_backward) return 0;
  return (int64_t)grid_for(plan, plan->engine, n_rows) * tile_rows(plan, plan->engine) * plan->host.stash_row * 4;
}

namespace {
int launch_step_impl(const mmn_plan* plan, const StepArgs& a, void* stream, bool train);
int launch_step(const mmn_plan* plan, const StepArgs& a, void* stream, bool train) { return launch_step_impl(plan, a, stream, train); }
int fill_args(const mmn_plan* plan, const mmn_batch* b, const float* params, const mmn_outputs* out, StepArgs& a) {
  const DevPlan& P = plan->host;
  if (!b || !params) return fail("null batch / params");
  if (b->n_rows <= 0) return fail("n_rows must be positive");
  if (b->seq_len < 0 || b->seq_len > P.E) return fail("seq_len must be in [0, E]");
  memset(&a, 0, sizeof a);
  a.plan = plan->dev;
  a.params = params;
  a.n_rows = b->n_rows;
  a.row_offset = b->row_offset;
  const int64_t bg = b->n_rows_global > 0 ? b->n_rows_global : b->n_rows;
  a.inv_rows_global = 1.0 / (double)bg;
  a.seq_len = b->seq_len;
  unsigned seen = 0;
  for (int k = 0; k < b->seq_len; ++k) {
    const int e = b->seq_enc[k], pos = b->seq_pos[k];
    if (e < 0 || e >= P.E) return fail("sequence step %d: encoder id %d out of range", k, e);
    if (seen & (1u << e)) return fail("sequence step %d: encoder id %d appears twice", k, e);
    seen |= 1u << e;
    if (pos < 0 || pos >= MMN_MAX_ENCODERS) return fail("sequence step %d: data position out of range", k);
    if (!b->x[pos]) return fail("sequence step %d: x[%d] is null", k, pos);
    if (b->x_ld[pos] < P.enc[e].F) return fail("sequence step %d: x[%d] row stride %lld < n_features %d", k, pos, (long long)b->x_ld[pos], P.enc[e].F);
    a.seq_enc[k] = e;
    a.seq_pos[k] = pos;
    a.x[pos] = b->x[pos];
    a.x_ld[pos] = b->x_ld[pos];
  }
  a.targets = (const long long*)b->targets;
  a.skip_flags = b->skip_flags;
  if (out) {
    a.metrics = out->metrics;
    a.predictions = out->predictions;
    a.pred_ld = out->pred_ld;
    a.last_outputs = out->last_outputs;
    a.final_state = out->final_state;
    a.target_error = out->target_error;
    if (out->predictions && out->pred_ld < b->n_rows) return fail("pred_ld < n_rows");
  }
  return 0;
}

}  // namespace

extern "C" int mmn_forward(const mmn_plan* plan, const mmn_batch* batch, const float* params,
                           const mmn_outputs* out, void* workspace, size_t workspace_bytes, void* stream) {
  if (!plan) return fail("mmn_forward: null plan");
  StepArgs a;
  if (fill_args(plan, batch, params, out, a)) return 1;
#ifndef MMN_EMU
  if (plan->engine == MMN_ENGINE_WIDE) {
    const int64_t need = mmn_workspace_bytes(plan, batch->n_rows, 0);
    if (!workspace || (int64_t)workspace_bytes < need) return fail("workspace too small: need %lld bytes", (long long)need);
    return mmn_wide_step(plan, a, workspace, workspace_bytes, stream, false);
  }
#endif
  if (plan->engine == MMN_ENGINE_NB) return mmn_nb_step(plan, a, workspace, workspace_bytes, stream, false);
  (void)workspace; (void)workspace_bytes;
  return launch_step(plan, a, stream, false);
}

extern "C" int mmn_train_step(const mmn_plan* plan, const mmn_batch* batch, const float* params,
                              const mmn_train_args* targs, const mmn_outputs* out, float* grads,
                              void* workspace, size_t workspace_bytes, void* stream) {
  if (!plan || !targs || !grads) return fail("mmn_train_step: null argument");
  const DevPlan& P = plan->host;
  StepArgs a;
  if (fill_args(plan, batch, params, out, a)) return 1;
  if (!batch->targets) return fail("mmn_train_step: targets are required");
  const int64_t need = mmn_workspace_bytes(plan, batch->n_rows, 1);
  if (!workspace || (int64_t)workspace_bytes < need) return fail("workspace too small: need %lld bytes", (long long)need);
  const double bg = 1.0 / a.inv_rows_global;
  a.grads = grads;
  a.stash = (float*)workspace;
  a.slot_floats = (long long)tile_rows(plan, plan->engine) * P.stash_row;
  a.c_err = (float)((double)targs->err_penalty / ((double)P.D * (P.E + 1) * bg));
  a.c_sc = (float)(2.0 * (double)targs->state_change_penalty_scaled / ((double)P.E * bg * P.S));
  a.dropout_seed = targs->dropout_seed;
  a.training = targs->training;
  MMN_CUDA(cudaMemsetAsync(grads, 0, sizeof(float) * (size_t)(P.n_params + P.E), (cudaStream_t)stream));
#ifndef MMN_EMU
  if (plan->engine == MMN_ENGINE_WIDE) return mmn_wide_step(plan, a, workspace, workspace_bytes, stream, true);
#endif
  if (plan->engine == MMN_ENGINE_NB) {
    if (mmn_nb_step(plan, a, workspace, workspace_bytes, stream, true)) return 1;
  } else if (launch_step(plan, a, stream, true)) return 1;
#ifndef MMN_EMU
  for (int i = 0; i < plan->n_grad_events; ++i)          // one launch: every gradient block is final at its end
    MMN_CUDA(cudaEventRecord((cudaEvent_t)plan->grad_events[i], (cudaStream_t)stream));
#endif
  return 0;
}


namespace {
int launch_step_impl(const mmn_plan* plan, const StepArgs& a, void* stream, bool train) {
  const int engine = train ? plan->engine : plan->fwd_engine;
  if (engine == MMN_ENGINE_TC2) return mmn_launch_v2(plan, a, stream, train);
  return mmn_launch_fma(plan, a, stream, train);
}
}  // namespace

#ifdef MMN_EMU
// the wide regime (mmn_wide.cu) is CUDA only: its diagnostic entry points exist in the host emulator but refuse to run
extern "C" int mmn_selftest_gemm_bf16(int, int, int, const void*, long long, const void*, long long, float*, void*, void*, void*) {
  return fail("mmn_selftest_gemm_bf16: the wide-regime GEMM is not part of the host emulator");
}
extern "C" int mmn_selftest_gemm_bf16_mn(int, int, int, const void*, long long, const void*, long long, float*, void*) {
  return fail("mmn_selftest_gemm_bf16_mn: the wide-regime GEMM is not part of the host emulator");
}
extern "C" int64_t mmn_wide_launch_count(void) { return 0; }
#endif
