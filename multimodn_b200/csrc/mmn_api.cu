// mmn_api.cu — the C ABI of libmmn.so (include/mmn.h): plan construction, launch configuration and
// argument marshalling around the kernels in mmn_kernels.cuh.  No torch types, no hidden syncs.
#include "mmn_kernels.cuh"
#include "mmn_tc.cuh"
#include "mmn_tc2.cuh"
#ifndef MMN_EMU
#include "mmn_wide_step.cuh"
#endif

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

using namespace mmn;

struct mmn_plan {
  DevPlan host;
  DevPlan* dev = nullptr;
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
};

namespace {
thread_local std::string g_err;

int fail(const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return 1;
}
#define MMN_CUDA(call)                                                            \
  do {                                                                            \
    cudaError_t e_ = (call);                                                      \
    if (e_ != cudaSuccess) return fail("%s: %s", #call, cudaGetErrorString(e_));  \
  } while (0)

int round32(int v) { return (v + 31) & ~31; }
#ifndef MMN_EMU
int wide_plan_init(mmn_plan* p);
#endif

size_t fma_smem(const DevPlan& P, int rm, int occ = 1) {
  const size_t stage = occ == 2 ? FmaEngine<2, 2>::stage_bytes()
                                : rm == 4 ? FmaEngine<4>::stage_bytes() : rm == 2 ? FmaEngine<2>::stage_bytes() : FmaEngine<1>::stage_bytes();
  return step_smem_bytes(P, 32 * rm, stage);
}
size_t tc_smem(const DevPlan& P) { return step_smem_bytes(P, TcEngine::TM, TcEngine::stage_bytes()); }
// largest FMA row tile (32*RM rows) whose shared-memory footprint fits
int pick_rm(const mmn_plan* p) {
  for (int rm : {4, 2, 1})
    if (fma_smem(p->host, rm) <= (size_t)p->max_smem) return rm;
  return 0;
}
int tile_rows(const mmn_plan* p, int engine) { return engine != MMN_ENGINE_FMA ? 128 : 32 * p->rm; }
int grid_for(const mmn_plan* p, int engine, int64_t n_rows) {
  const int tm = tile_rows(p, engine);
  const int64_t tiles = (n_rows + tm - 1) / tm;
  const int per_sm = engine == MMN_ENGINE_FMA ? p->occ : 1;
  return (int)std::max<int64_t>(1, std::min<int64_t>(tiles, (int64_t)p->n_sms * per_sm));
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
#ifdef MMN_EMU
    delete p;
    return fail("precision = bf16 (the wide regime) is not part of the host emulator");
#else
    p->engine = p->fwd_engine = MMN_ENGINE_WIDE;
    if (wide_plan_init(p) || cudaMalloc((void**)&p->dev, sizeof(DevPlan)) != cudaSuccess ||
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
  const bool tc_fits = tc_smem(P) <= (size_t)p->max_smem;
  const char* want = getenv("MMN_ENGINE");            // "tc" | "fma" | unset = auto
  if (want && !strcmp(want, "tc") && !tc_fits) {
    delete p;
    return fail("MMN_ENGINE=tc: the tensor-core engine needs %zu B of shared memory for this model (limit %d)", tc_smem(P), p->max_smem);
  }
  // default: the FP32-FMA engine (faster at the current stage of tuning, profiles/r1_engine_timers.txt);
  // MMN_ENGINE=tc opts into the tcgen05 3xTF32 engine
  p->engine = (tc_fits && want && !strcmp(want, "tc")) ? MMN_ENGINE_TC : MMN_ENGINE_FMA;
  // forward-only launches (test / predict / get_states): the TMEM-resident kernel where the model qualifies
  const bool v2_ok = V2Engine::supports(P) && V2Engine::smem_bytes(P) <= (size_t)p->max_smem;
  if (want && !strcmp(want, "tc2") && !v2_ok) {
    delete p;
    return fail("MMN_ENGINE=tc2: the TMEM-resident kernel needs state <= 64, layers <= 64 wide, <= 16 classes");
  }
  p->fwd_engine = (v2_ok && !want) || (want && !strcmp(want, "tc2")) ? MMN_ENGINE_TC2 : p->engine;
  if (want && !strcmp(want, "tc2")) p->engine = MMN_ENGINE_TC2;
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
  delete plan;
}

extern "C" int64_t mmn_metrics_count(const mmn_plan* plan) { return plan ? plan->host.n_metrics : -1; }
extern "C" int64_t mmn_grad_count(const mmn_plan* plan) { return plan ? plan->host.n_params + plan->host.E : -1; }

extern "C" int32_t mmn_plan_engine(const mmn_plan* plan) { return plan ? plan->engine : -1; }
extern "C" int32_t mmn_plan_forward_engine(const mmn_plan* plan) { return plan ? plan->fwd_engine : -1; }


// ------------------------------------------------------------------------------------------------
// wide regime (precision = bf16): layer-wise tcgen05 GEMMs (mmn_wide.cuh, mmn_wide_step.cuh)
// ------------------------------------------------------------------------------------------------
#ifndef MMN_EMU
namespace {
typedef CUresult (*TmapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
TmapEncodeFn tmap_encoder() {
  static TmapEncodeFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return (TmapEncodeFn)p;
  }();
  return fn;
}
// K-major bf16 operand [rows x k] with row pitch ld (elements): boxes of 64 k x box_rows rows, SWIZZLE_128B
int make_operand_map(CUtensorMap* m, const void* base, long long rows, long long k, long long ld, int box_rows) {
  TmapEncodeFn enc = tmap_encoder();
  if (!enc) return fail("cuTensorMapEncodeTiled is not available in this driver");
  if ((reinterpret_cast<size_t>(base) & 15) || (ld & 7)) return fail("wide GEMM operand: base must be 16-byte aligned and the row pitch a multiple of 8 elements");
  const cuuint64_t dims[2] = {(cuuint64_t)k, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  const cuuint32_t box[2] = {(cuuint32_t)wide::BK, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled failed (%d) for a %lld x %lld operand, pitch %lld", (int)r, rows, k, ld);
  return 0;
}
// MN-major bf16 operand: the contraction index runs over the ROWS of a row-major matrix [k_rows x mn] with pitch ld
// (a matrix used "transposed" without a transposed copy): boxes of 64 mn x 64 k rows, SWIZZLE_128B
int make_operand_map_mn(CUtensorMap* m, const void* base, long long mn, long long k_rows, long long ld) {
  TmapEncodeFn enc = tmap_encoder();
  if (!enc) return fail("cuTensorMapEncodeTiled is not available in this driver");
  if ((reinterpret_cast<size_t>(base) & 15) || (ld & 7)) return fail("wide GEMM operand: base must be 16-byte aligned and the row pitch a multiple of 8 elements");
  const cuuint64_t dims[2] = {(cuuint64_t)mn, (cuuint64_t)k_rows};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  const cuuint32_t box[2] = {64, (cuuint32_t)wide::BK};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled failed (%d) for an MN-major %lld x %lld operand, pitch %lld", (int)r, k_rows, mn, ld);
  return 0;
}
long long g_wide_launches = 0;       // kernels launched by the wide path (bench.py's gpu_launches)
// MMN_WIDE_TIMERS=1: CUDA events around every launch of a step, summed per category and printed (development aid)
struct WideTimers {
  bool on = false;
  std::vector<std::pair<const char*, std::pair<cudaEvent_t, cudaEvent_t>>> ev;
  cudaStream_t stream = nullptr;
  const char* cat = "other";
  void begin(const char* c) {
    cat = c;
    if (!on) return;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a, stream);
    ev.push_back({c, {a, b}});
  }
  void end() {
    if (on && !ev.empty()) cudaEventRecord(ev.back().second.second, stream);
  }
  void report() {
    if (!on) return;
    cudaStreamSynchronize(stream);
    std::vector<std::pair<std::string, std::pair<double, int>>> tot;
    double all = 0;
    for (auto& e : ev) {
      float ms = 0;
      cudaEventElapsedTime(&ms, e.second.first, e.second.second);
      cudaEventDestroy(e.second.first); cudaEventDestroy(e.second.second);
      all += ms;
      bool found = false;
      for (auto& x : tot) if (x.first == e.first) { x.second.first += ms; x.second.second++; found = true; }
      if (!found) tot.push_back({e.first, {ms, 1}});
    }
    fprintf(stderr, "[mmn wide timers] %.3f ms in %zu launches:", all, ev.size());
    for (auto& x : tot) fprintf(stderr, " %s %.3f ms (%d)", x.first.c_str(), x.second.first, x.second.second);
    fprintf(stderr, "\n");
    ev.clear();
  }
};
WideTimers g_wt;
// D[M x N] = A[M x K] . B[N x K]^T
// a_mn / b_mn = 0: the operand is [M or N rows x K] with K contiguous.  = 1: it is [K rows x M or N] with M / N contiguous.
int wide_gemm(int n_sms, const void* A, long long lda, const void* B, long long ldb, long long M, long long N, long long K,
              const wide::Epi& epi, void* stream, const char* what = "gemm", int a_mn = 0, int b_mn = 0) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  alignas(64) CUtensorMap ma, mb;
  if (a_mn ? make_operand_map_mn(&ma, A, M, K, lda) : make_operand_map(&ma, A, M, K, lda, wide::BM)) return 1;
  if (b_mn ? make_operand_map_mn(&mb, B, N, K, ldb) : make_operand_map(&mb, B, N, K, ldb, wide::BN)) return 1;
  static bool attr_set = false;
  if (!attr_set) {
    MMN_CUDA(cudaFuncSetAttribute(wide::mmn_wide_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, wide::kSmemBytes));
    attr_set = true;
  }
  const long long tiles = ((M + wide::BM - 1) / wide::BM) * ((N + wide::BN - 1) / wide::BN);
  // split-K: only for accumulating fp32 outputs (weight gradients), when the output tiles alone leave SMs idle and
  // every split still has a long K range
  int splits = 1;
  if (epi.mode == wide::EPI_ACCUM_F32 && epi.accumulate && !epi.out && !epi.out_t && tiles < n_sms) {
    const long long kb = (K + wide::BK - 1) / wide::BK;
    splits = (int)std::max<long long>(1, std::min<long long>(std::min<long long>((2 * n_sms) / tiles, kb / 16), 16));
    while (splits > 1 && (long long)(splits - 1) * ((kb + splits - 1) / splits) >= kb) --splits;     // no empty split
  }
  const int grid = (int)std::min<long long>(tiles * splits, n_sms);
  g_wt.begin(what);
  wide::mmn_wide_gemm_kernel<<<grid, wide::kThreads, wide::kSmemBytes, (cudaStream_t)stream>>>(ma, mb, (int)M, (int)N, (int)K, splits, a_mn, b_mn, epi);
  g_wt.end();
  MMN_CUDA(cudaGetLastError());
  ++g_wide_launches;
  return 0;
}

long long round8(long long v) { return (v + 7) & ~7ll; }

// bump allocator over the caller's workspace; base == nullptr only measures
struct Arena {
  char* base;
  size_t off = 0, peak = 0;
  explicit Arena(void* b) : base((char*)b) {}
  void* take(size_t bytes) {
    off = (off + 255) & ~(size_t)255;
    void* p = base ? base + off : nullptr;
    off += bytes;
    peak = std::max(peak, off);
    return p;
  }
  bool want_t = true;        // forward-only calls need no transposed copies (they only feed weight gradients)
  wide::Mat mat(long long rows, int width) {
    wide::Mat m;
    m.width = width;
    m.ld = round8(width);
    m.ldt = round8(rows);
    m.p = (wide::bf16*)take((size_t)rows * m.ld * 2);
    m.t = want_t ? (wide::bf16*)take((size_t)width * m.ldt * 2) : nullptr;
    return m;
  }
};

int wide_plan_init(mmn_plan* p) {
  const DevPlan& P = p->host;
  long long off = 0;
  auto place = [&](const DevLayer& l, mmn_plan::WL& w) {
    w.ldk = (int)round8(l.ktot);
    w.ldo = (int)round8(l.out_dim);
    w.w = off; off += ((long long)l.out_dim * w.ldk + 127) & ~127ll;
    w.wt = off; off += ((long long)l.ktot * w.ldo + 127) & ~127ll;
  };
  for (int e = 0; e < P.E; ++e) {
    if (P.enc[e].p_drop > 0.f && !P.enc[e].L[0].has_state) return fail("wide regime: dropout needs the state on layer 0");
    for (int j = 0; j < P.enc[e].n_layers; ++j) place(P.enc[e].L[j], p->wide_enc[e][j]);
  }
  for (int d = 0; d < P.D; ++d)
    for (int j = 0; j < P.dec[d].n_layers; ++j) place(P.dec[d].L[j], p->wide_dec[d][j]);
  p->wide_elems = off;
  MMN_CUDA(cudaMalloc(&p->wide_w, (size_t)off * 2));
  MMN_CUDA(cudaMemset(p->wide_w, 0, (size_t)off * 2));
  return 0;
}

// The whole step.  dry = true only sizes the workspace (no launches).
template <bool TRAIN>
int wide_step(const mmn_plan* plan, const StepArgs& a, void* ws, size_t ws_bytes, void* stream_, bool dry, size_t* need_out) {
  using namespace wide;
  const DevPlan& P = plan->host;
  const cudaStream_t stream = (cudaStream_t)stream_;
  const long long B = a.n_rows;
  const int S = P.S, E = P.E, D = P.D, L = a.seq_len;
  const int n_sms = plan->n_sms;
  Arena ar(dry ? nullptr : ws);
  ar.want_t = false;         // weight gradients read dZ and the layer inputs in place (MN-major operands)
  bf16* const wbase = (bf16*)plan->wide_w;
  const dim3 tb(256);
  auto tgrid = [&](long long rows, int width) { return dim3((unsigned)((width + 63) / 64), (unsigned)((rows + 63) / 64)); };
  g_wt.on = !dry && getenv("MMN_WIDE_TIMERS") != nullptr;
  g_wt.stream = stream;
  auto launched = [&]() -> int {
    g_wt.end();
    ++g_wide_launches;
    MMN_CUDA(cudaGetLastError());
    return 0;
  };
  Drop nodrop;
  memset(&nodrop, 0, sizeof nodrop);
  nodrop.scale = 1.f;

  // ---- 0. bf16 copies of the weights (both orientations) ----
  if (!dry) {
    auto cast = [&](const DevLayer& l, const mmn_plan::WL& w) -> int {
      g_wt.begin("cast_weight");
      wide_cast_weight_kernel<<<tgrid(l.out_dim, l.ktot), tb, 0, stream>>>(a.params + l.w_off, l.out_dim, l.ktot, wbase + w.w, w.ldk,
                                                                           wbase + w.wt, w.ldo);
      return launched();
    };
    for (int e = 0; e < E; ++e)
      for (int j = 0; j < P.enc[e].n_layers; ++j)
        if (cast(P.enc[e].L[j], plan->wide_enc[e][j])) return 1;
    for (int d = 0; d < D; ++d)
      for (int j = 0; j < P.dec[d].n_layers; ++j)
        if (cast(P.dec[d].L[j], plan->wide_dec[d][j])) return 1;
  }

  // ---- persistent buffers ----
  unsigned char* present = (unsigned char*)ar.take((size_t)(L + 1) * B);
  float* sc_sum = (float*)ar.take(sizeof(float) * (size_t)std::max(E, 1));
  int maxC = 1, maxW = S;
  for (int d = 0; d < D; ++d) {
    maxC = std::max(maxC, P.dec[d].C);
    for (int j = 0; j < P.dec[d].n_layers; ++j) maxW = std::max(maxW, P.dec[d].L[j].out_dim);
  }
  for (int e = 0; e < E; ++e)
    for (int j = 0; j < P.enc[e].n_layers; ++j) maxW = std::max(maxW, P.enc[e].L[j].out_dim);
  float* Pout = (float*)ar.take(sizeof(float) * (size_t)B * maxC);
  std::vector<Mat> Sk(L + 1);
  if (TRAIN) {
    for (int k = 0; k <= L; ++k) Sk[k] = ar.mat(B, S);
  } else {
    const Mat s0 = ar.mat(B, S), s1 = ar.mat(B, S);
    for (int k = 0; k <= L; ++k) Sk[k] = (k & 1) ? s1 : s0;
  }
  // per (step, module, layer) activations kept for the backward pass
  std::vector<Mat> enc_in((size_t)(L + 1) * MMN_MAX_LAYERS);
  std::vector<Mat> dec_h((size_t)(L + 1) * D * MMN_MAX_LAYERS);
  std::vector<Mat> dec_dz((size_t)(L + 1) * D);
  if (!dry) {
    MMN_CUDA(cudaMemsetAsync(present, 1, (size_t)(L + 1) * B, stream));
    MMN_CUDA(cudaMemsetAsync(sc_sum, 0, sizeof(float) * (size_t)std::max(E, 1), stream));
    g_wt.begin("init_state");
    wide_init_state_kernel<<<tgrid(B, S), tb, 0, stream>>>(a.params + P.init_off, B, Sk[0]);
    if (launched()) return 1;
  }
  const size_t scratch_mark = ar.off;

  auto epi0 = [] {
    Epi e;
    memset(&e, 0, sizeof e);
    e.scale = 1.f;
    return e;
  };

  // ---- decoders on s_k ----
  auto decoders_forward = [&](int k, int hist_row, bool is_last_enc, const int* skip) -> int {
    for (int d = 0; d < D; ++d) {
      const DevDecoder& dec = P.dec[d];
      Mat in = Sk[k];
      for (int j = 0; j < dec.n_layers; ++j) {
        const DevLayer& ly = dec.L[j];
        const mmn_plan::WL& w = plan->wide_dec[d][j];
        const bool last = j == dec.n_layers - 1;
        Epi e = epi0();
        e.mode = EPI_STORE; e.act = ly.act; e.bias = a.params + ly.b_off;
        Mat out;
        if (!last) {
          out = ar.mat(B, ly.out_dim);
          dec_h[((size_t)k * D + d) * MMN_MAX_LAYERS + j] = out;
          e.out = out.p; e.ld_out = out.ld; e.out_t = out.t; e.ld_out_t = out.ldt;
        } else {
          e.out_f32 = Pout; e.ld_f32 = dec.C;
        }
        if (!dry) {
          if (last && j > 0 && dec.C <= 4) {   // decoder head: skinny, bandwidth-bound kernel instead of a tensor-core tile
            g_wt.begin("head_fwd");
            wide_head_fwd_kernel<4><<<(unsigned)std::min<long long>((B + 7) / 8, 8 * n_sms), 256, 0, stream>>>(
                in, wbase + w.w, w.ldk, a.params + ly.b_off, dec.C, ly.act, B, Pout);
            if (launched()) return 1;
          } else if (wide_gemm(n_sms, in.p, in.ld, wbase + w.w, w.ldk, B, ly.out_dim, ly.ktot, e, stream, "gemm fwd")) {
            return 1;
          }
        }
        in = out;
      }
      LossArgs la;
      memset(&la, 0, sizeof la);
      la.p = Pout; la.ldp = dec.C; la.C = dec.C; la.act = dec.L[dec.n_layers - 1].act; la.D = D; la.d = d;
      la.hist_row = hist_row; la.n_mat_rows = E + 1; la.rows = B; la.targets = a.targets;
      la.present = k == 0 ? nullptr : present + (size_t)k * B;
      la.skip = skip;
      la.metrics = a.metrics; la.inv_rows_global = a.inv_rows_global;
      la.predictions = a.predictions ? a.predictions + ((long long)hist_row * D + d) * a.pred_ld : nullptr;
      if (a.last_outputs && is_last_enc) { la.last_outputs = a.last_outputs; la.ld_last = P.sumC; la.out_off = dec.out_off; }
      if (TRAIN) {
        la.coef = a.c_err;
        la.dz = ar.mat(B, dec.C);
        dec_dz[(size_t)k * D + d] = la.dz;
      }
      if (!dry) {
        if (TRAIN) {      // the pitch padding of dz is read by the TMA unit as part of full 16-byte rows: keep it finite
          MMN_CUDA(cudaMemsetAsync(la.dz.p, 0, (size_t)B * la.dz.ld * 2, stream));
        }
        g_wt.begin("decoder_loss");
        wide_decoder_loss_kernel<<<(unsigned)((B + 255) / 256), 256, 0, stream>>>(la);
        if (launched()) return 1;
      }
    }
    return 0;
  };

  if (decoders_forward(0, 0, false, nullptr)) return 1;
  if (!dry) {
    g_wt.begin("finalize");
    wide_finalize_kernel<<<1, 256, 0, stream>>>(present, B, 0, 0, 0, nullptr, nullptr, S, a.inv_rows_global,
                                                a.metrics ? a.metrics + met_present(P, 0) : nullptr, nullptr, nullptr);
    if (launched()) return 1;
  }
  if (!TRAIN) ar.off = scratch_mark;

  // ---- walk the encoding sequence ----
  for (int k = 1; k <= L; ++k) {
    const int e = a.seq_enc[k - 1], pos = a.seq_pos[k - 1];
    const DevEncoder& enc = P.enc[e];
    const int* skip = a.skip_flags ? a.skip_flags + (k - 1) : nullptr;
    unsigned char* pres = present + (size_t)k * B;
    Drop drop = nodrop;
    if (TRAIN && a.training && enc.p_drop > 0.f) {
      drop.enabled = 1;
      drop.seed_mix = a.dropout_seed ^ ((unsigned)e * 0x9E3779B9u);
      drop.thr = (unsigned)(enc.p_drop * 65536.f);
      drop.row_base = (unsigned)a.row_offset;
      drop.scale = 1.f / (1.f - enc.p_drop);
    }
    Mat in = ar.mat(B, enc.L[0].ktot);
    enc_in[(size_t)k * MMN_MAX_LAYERS + 0] = in;
    if (!dry) {
      g_wt.begin("input_x");
      wide_input_x_kernel<<<tgrid(B, enc.F), tb, 0, stream>>>(a.x[pos], a.x_ld[pos], B, enc.F, in, pres, drop);
      if (launched()) return 1;
      if (enc.L[0].has_state) {
        g_wt.begin("input_state");
        wide_input_state_kernel<<<tgrid(B, S), tb, 0, stream>>>(Sk[k - 1], B, in, enc.L[0].in_dim, drop);
        if (launched()) return 1;
      }
    }
    for (int j = 0; j < enc.n_layers; ++j) {
      const DevLayer& ly = enc.L[j];
      const mmn_plan::WL& w = plan->wide_enc[e][j];
      const bool last = j == enc.n_layers - 1;
      Epi ep = epi0();
      ep.act = ly.act; ep.bias = a.params + ly.b_off;
      Mat next;
      if (!last) {
        const DevLayer& nx = enc.L[j + 1];
        next = ar.mat(B, nx.ktot);
        enc_in[(size_t)k * MMN_MAX_LAYERS + j + 1] = next;
        ep.mode = EPI_STORE;
        ep.out = next.p; ep.ld_out = next.ld; ep.out_t = next.t; ep.ld_out_t = next.ldt;
      } else {
        ep.mode = EPI_SELECT;
        ep.aux = Sk[k - 1].p; ep.ld_aux = Sk[k - 1].ld;
        ep.present = pres; ep.skip = skip;
        ep.out = Sk[k].p; ep.ld_out = Sk[k].ld; ep.out_t = Sk[k].t; ep.ld_out_t = Sk[k].ldt;
        ep.sc_sum = TRAIN ? sc_sum + e : nullptr;
      }
      if (!dry) {
        if (wide_gemm(n_sms, in.p, in.ld, wbase + w.w, w.ldk, B, ly.out_dim, ly.ktot, ep, stream)) return 1;
        if (!last && enc.L[j + 1].has_state) {
          g_wt.begin("input_state");
          wide_input_state_kernel<<<tgrid(B, S), tb, 0, stream>>>(Sk[k - 1], B, next, enc.L[j + 1].in_dim, nodrop);
          if (launched()) return 1;
        }
      }
      in = next;
    }
    if (!dry) {
      g_wt.begin("finalize");
      wide_finalize_kernel<<<1, 256, 0, stream>>>(pres, B, k, e + 1, e, skip, TRAIN ? sc_sum + e : nullptr, S, a.inv_rows_global,
                                                  a.metrics ? a.metrics + met_present(P, 0) : nullptr,
                                                  (TRAIN && a.metrics) ? a.metrics + met_sc(P, 0) : nullptr,
                                                  TRAIN ? a.grads + P.n_params : nullptr);
      if (launched()) return 1;
    }
    if (decoders_forward(k, e + 1, e == E - 1, skip)) return 1;
    if (!TRAIN) ar.off = scratch_mark;
  }
  if (a.final_state && !dry) {
    g_wt.begin("state_out");
    wide_state_out_kernel<<<(unsigned)std::min<long long>((B * S + 255) / 256, 4096), 256, 0, stream>>>(Sk[L], B, a.final_state);
    if (launched()) return 1;
  }

  if (TRAIN) {
    // =====================================================================================================
    // backward: replay the sequence in reverse
    // =====================================================================================================
    float* G = (float*)ar.take(sizeof(float) * (size_t)B * S);
    Mat dzbuf[2] = {ar.mat(B, maxW), ar.mat(B, maxW)};
    auto view = [](const Mat& m, int width) {       // same memory, narrower logical width (pitches of the narrow matrix)
      Mat v = m;
      v.width = width;
      v.ld = round8(width);
      return v;
    };
    if (!dry) MMN_CUDA(cudaMemsetAsync(G, 0, sizeof(float) * (size_t)B * S, stream));
    // gradients of one layer given dz (both orientations) and the layer's input
    auto layer_param_grads = [&](const DevLayer& ly, const Mat& dz, const Mat& in) -> int {
      if (dry) return 0;
      g_wt.begin("bias_grad");
      wide_bias_grad_kernel<<<dim3((unsigned)((ly.out_dim + 63) / 64), (unsigned)std::max<long long>(1, std::min<long long>(32, B / 256))),
                              256, 0, stream>>>(dz.p, dz.ld, B, ly.out_dim, a.grads + ly.b_off);
      if (launched()) return 1;
      Epi e = epi0();
      e.mode = EPI_ACCUM_F32; e.accumulate = 1;
      e.out_f32 = a.grads + ly.w_off; e.ld_f32 = ly.ktot;
      return wide_gemm(n_sms, dz.p, dz.ld, in.p, in.ld, ly.out_dim, ly.ktot, B, e, stream, "gemm wgrad", 1, 1);
    };
    auto decoders_backward = [&](int k) -> int {
      for (int d = 0; d < D; ++d) {
        const DevDecoder& dec = P.dec[d];
        Mat dz = dec_dz[(size_t)k * D + d];
        int cur = 0;
        for (int j = dec.n_layers - 1; j >= 0; --j) {
          const DevLayer& ly = dec.L[j];
          const mmn_plan::WL& w = plan->wide_dec[d][j];
          const Mat in = j == 0 ? Sk[k] : dec_h[((size_t)k * D + d) * MMN_MAX_LAYERS + j - 1];
          if (j == dec.n_layers - 1 && j > 0 && dec.C <= 4) {       // decoder head (see decoders_forward)
            const Mat nz = view(dzbuf[cur], ly.in_dim);
            if (!dry) {
              g_wt.begin("head_wgrad");
              wide_head_wgrad_kernel<4><<<dim3((unsigned)((ly.in_dim + 2047) / 2048), (unsigned)std::max<long long>(1, std::min<long long>(2 * n_sms, B / 32))),
                                          256, 0, stream>>>(dz, in, dec.C, B, a.grads + ly.w_off, ly.ktot, a.grads + ly.b_off);
              if (launched()) return 1;
              g_wt.begin("head_dgrad");
              wide_head_dgrad_kernel<<<tgrid(B, ly.in_dim), tb, 0, stream>>>(dz, wbase + w.w, w.ldk, dec.C, in, dec.L[j - 1].act, B, nz);
              if (launched()) return 1;
            }
            dz = nz;
            cur ^= 1;
            continue;
          }
          if (layer_param_grads(ly, dz, in)) return 1;
          Epi e = epi0();
          if (j > 0) {
            const Mat nz = view(dzbuf[cur], ly.in_dim);
            e.mode = EPI_DACT; e.act = dec.L[j - 1].act;
            e.aux = in.p; e.ld_aux = in.ld;
            e.out = nz.p; e.ld_out = nz.ld; e.out_t = nz.t; e.ld_out_t = nz.ldt;
            if (!dry && wide_gemm(n_sms, dz.p, dz.ld, wbase + w.wt, w.ldo, B, ly.in_dim, ly.out_dim, e, stream, ly.out_dim < 64 ? "gemm dec-head dgrad" : "gemm dgrad")) return 1;
            dz = nz;
            cur ^= 1;
          } else {
            e.mode = EPI_ACCUM_F32; e.accumulate = 1;
            e.out_f32 = G; e.ld_f32 = S;
            if (!dry && wide_gemm(n_sms, dz.p, dz.ld, wbase + w.wt, w.ldo, B, S, ly.out_dim, e, stream)) return 1;
          }
        }
      }
      return 0;
    };
    for (int k = L; k >= 1; --k) {
      const int e = a.seq_enc[k - 1];
      const DevEncoder& enc = P.enc[e];
      const int* skip = a.skip_flags ? a.skip_flags + (k - 1) : nullptr;
      unsigned char* pres = present + (size_t)k * B;
      if (decoders_backward(k)) return 1;
      const int nl = enc.n_layers;
      int cur = 0;
      Mat dz = view(dzbuf[cur], S);
      cur ^= 1;
      if (!dry) {
        g_wt.begin("state_grad");
        wide_state_grad_kernel<<<tgrid(B, S), tb, 0, stream>>>(G, Sk[k], Sk[k - 1], pres, skip, a.c_sc, enc.L[nl - 1].act, B, dz);
        if (launched()) return 1;
      }
      for (int j = nl - 1; j >= 0; --j) {
        const DevLayer& ly = enc.L[j];
        const mmn_plan::WL& w = plan->wide_enc[e][j];
        const Mat in = enc_in[(size_t)k * MMN_MAX_LAYERS + j];
        if (layer_param_grads(ly, dz, in)) return 1;
        if (ly.has_state && !dry) {
          // carry into G: present rows take dz W_s (through the dropout mask), absent rows keep G; then remove u_k
          Epi ep = epi0();
          ep.mode = EPI_CARRY;
          ep.out_f32 = G; ep.ld_f32 = S;
          ep.present = pres; ep.skip = skip;
          if (j == 0 && TRAIN && a.training && enc.p_drop > 0.f) {
            ep.drop_thr = (unsigned)(enc.p_drop * 65536.f);
            ep.drop_seed = a.dropout_seed ^ ((unsigned)e * 0x9E3779B9u);
            ep.drop_row_base = (unsigned)a.row_offset;
            ep.drop_col_base = (unsigned)ly.in_dim;
            ep.scale = 1.f / (1.f - enc.p_drop);
          }
          if (wide_gemm(n_sms, dz.p, dz.ld, wbase + w.wt + (long long)ly.in_dim * w.ldo, w.ldo, B, S, ly.out_dim, ep, stream)) return 1;
          g_wt.begin("state_grad_post");
          wide_state_grad_post_kernel<<<(unsigned)std::min<long long>((B * S + 255) / 256, 4096), 256, 0, stream>>>(G, Sk[k], Sk[k - 1],
                                                                                                                   a.c_sc, B);
          if (launched()) return 1;
        }
        if (j > 0) {
          const Mat nz = view(dzbuf[cur], ly.in_dim);
          Epi ep = epi0();
          ep.mode = EPI_DACT; ep.act = enc.L[j - 1].act;
          ep.aux = in.p; ep.ld_aux = in.ld;
          ep.out = nz.p; ep.ld_out = nz.ld; ep.out_t = nz.t; ep.ld_out_t = nz.ldt;
          if (!dry && wide_gemm(n_sms, dz.p, dz.ld, wbase + w.wt, w.ldo, B, ly.in_dim, ly.out_dim, ep, stream)) return 1;
          dz = nz;
          cur ^= 1;
        }
      }
    }
    if (decoders_backward(0)) return 1;
    if (!dry) {
      g_wt.begin("colsum_f32");
      wide_colsum_f32_kernel<<<dim3((unsigned)((S + 31) / 32), 16), 256, 0, stream>>>(G, B, S, a.grads + P.init_off);
      if (launched()) return 1;
    }
  }
  g_wt.report();
  if (need_out) *need_out = ar.peak + 256;
  if (!dry && ar.peak > ws_bytes) return fail("wide regime: workspace too small (need %zu bytes, got %zu)", ar.peak, ws_bytes);
  return 0;
}

int64_t wide_workspace_bytes(const mmn_plan* plan, int64_t n_rows, bool train) {
  StepArgs a;
  memset(&a, 0, sizeof a);
  a.n_rows = n_rows;
  a.seq_len = plan->host.E;
  for (int k = 0; k < a.seq_len; ++k) { a.seq_enc[k] = k; a.seq_pos[k] = k; }
  size_t need = 0;
  const int rc = train ? wide_step<true>(plan, a, nullptr, 0, nullptr, true, &need) : wide_step<false>(plan, a, nullptr, 0, nullptr, true, &need);
  return rc ? -1 : (int64_t)need;
}
}  // namespace
#endif

extern "C" int64_t mmn_workspace_bytes(const mmn_plan* plan, int64_t n_rows, int32_t with_backward) {
  if (!plan || n_rows < 0) return -1;
#ifndef MMN_EMU
  if (plan->engine == MMN_ENGINE_WIDE) return wide_workspace_bytes(plan, n_rows, with_backward != 0);
#endif
  if (!with_backward) return 0;
  return (int64_t)grid_for(plan, plan->engine, n_rows) * tile_rows(plan, plan->engine) * plan->host.stash_row * 4;
}

namespace {
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
    if (out->predictions && out->pred_ld < b->n_rows) return fail("pred_ld < n_rows");
  }
  return 0;
}

template <class ENG, bool TRAIN>
int launch_engine(const mmn_plan* plan, const StepArgs& a_in, void* stream) {
  StepArgs a = a_in;
  const size_t smem = step_smem_bytes(plan->host, ENG::TM, ENG::stage_bytes());
  const int grid = grid_for(plan, ENG::kTensor ? MMN_ENGINE_TC : MMN_ENGINE_FMA, a.n_rows);
  auto kfn = mmn_step_kernel<ENG, TRAIN>;
  MMN_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
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
template <bool TRAIN>
int launch_step(const mmn_plan* plan, const StepArgs& a, void* stream) {
  const int engine = TRAIN ? plan->engine : plan->fwd_engine;
  if (engine == MMN_ENGINE_TC2) return launch_v2<TRAIN>(plan, a, stream);
  if (engine == MMN_ENGINE_TC) return launch_engine<TcEngine, TRAIN>(plan, a, stream);
  if (plan->occ == 2) return launch_engine<FmaEngine<2, 2>, TRAIN>(plan, a, stream);
  switch (plan->rm) {
    case 4: return launch_engine<FmaEngine<4>, TRAIN>(plan, a, stream);
    case 2: return launch_engine<FmaEngine<2>, TRAIN>(plan, a, stream);
    case 1: return launch_engine<FmaEngine<1>, TRAIN>(plan, a, stream);
    default: return fail("no tile configuration fits");
  }
}
}  // namespace

extern "C" int mmn_scan_missing(const mmn_plan* plan, const mmn_batch* b, int32_t* flags, void* stream) {
  if (!plan || !b || !flags) return fail("mmn_scan_missing: null argument");
  const DevPlan& P = plan->host;
  if (b->seq_len < 0 || b->seq_len > P.E) return fail("seq_len must be in [0, E]");
  ScanArgs s;
  memset(&s, 0, sizeof s);
  s.seq_len = b->seq_len;
  s.n_rows = b->n_rows;
  s.flags = flags;
  for (int k = 0; k < b->seq_len; ++k) {
    const int e = b->seq_enc[k], pos = b->seq_pos[k];
    if (e < 0 || e >= P.E || pos < 0 || pos >= MMN_MAX_ENCODERS || !b->x[pos]) return fail("mmn_scan_missing: bad sequence step %d", k);
    s.F[k] = P.enc[e].F;
    s.x[k] = b->x[pos];
    s.x_ld[k] = b->x_ld[pos];
  }
  MMN_CUDA(cudaMemsetAsync(flags, 0, sizeof(int32_t) * std::max(1, b->seq_len), (cudaStream_t)stream));
  const int grid = std::max(1, std::min(plan->n_sms * 4, (int)((b->n_rows + 7) / 8)));
  MMN_LAUNCH(mmn_scan_missing_kernel, dim3(grid), dim3(256), 0, stream, s);
  MMN_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int mmn_forward(const mmn_plan* plan, const mmn_batch* batch, const float* params,
                           const mmn_outputs* out, void* workspace, size_t workspace_bytes, void* stream) {
  if (!plan) return fail("mmn_forward: null plan");
  StepArgs a;
  if (fill_args(plan, batch, params, out, a)) return 1;
#ifndef MMN_EMU
  if (plan->engine == MMN_ENGINE_WIDE) {
    const int64_t need = mmn_workspace_bytes(plan, batch->n_rows, 0);
    if (!workspace || (int64_t)workspace_bytes < need) return fail("workspace too small: need %lld bytes", (long long)need);
    return wide_step<false>(plan, a, workspace, workspace_bytes, stream, false, nullptr);
  }
#endif
  (void)workspace; (void)workspace_bytes;
  return launch_step<false>(plan, a, stream);
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
  if (plan->engine == MMN_ENGINE_WIDE) return wide_step<true>(plan, a, workspace, workspace_bytes, stream, false, nullptr);
#endif
  return launch_step<true>(plan, a, stream);
}

extern "C" int mmn_adam_step(const mmn_plan* plan, float* params, const float* grads, float* exp_avg,
                             float* exp_avg_sq, int32_t* step_count, float lr, float beta1, float beta2,
                             float eps, void* stream) {
  if (!plan || !params || !grads || !exp_avg || !exp_avg_sq || !step_count) return fail("mmn_adam_step: null argument");
  MMN_LAUNCH(mmn_adam_tick_kernel, dim3(1), dim3(32), 0, stream, plan->dev, grads, step_count);
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((plan->host.n_params + 255) / 256, (int64_t)plan->n_sms * 8));
  MMN_LAUNCH(mmn_adam_kernel, dim3(grid), dim3(256), 0, stream, plan->dev, params, grads, exp_avg, exp_avg_sq,
             step_count, lr, beta1, beta2, eps);
  MMN_CUDA(cudaGetLastError());
  return 0;
}

// Diagnostic: one tcgen05 3xTF32 GEMM in each operand configuration of the tensor-core engine
// (mmn_tc.cuh).  a, b, out: device pointers, see mmn_tc_selftest_kernel.
extern "C" int mmn_selftest_umma(int mode, int n, const float* a, const float* b, float* out, void* stream) {
  if (mode < 0 || mode > 4 || (n != 32 && n != 64) || (mode == 2 && n != 32)) return fail("mmn_selftest_umma: bad mode / n");
  const size_t smem = 1024 + 98304 + 64;
  auto kfn = mmn_tc_selftest_kernel;
  MMN_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  MMN_LAUNCH(kfn, dim3(1), dim3(256), smem, stream, mode, n, a, b, out);
  MMN_CUDA(cudaGetLastError());
  return 0;
}

// Development aid: cycles per round of the worker <-> MMA-issuer handshake (mmn_tc2.cuh).  out: device int64[2].
extern "C" int mmn_selftest_protocol(int iters, int n_mma, int flags, long long* out, void* stream) {
  auto kfn = mmn_protocol_probe_kernel;
  const size_t smem = 1024 + 32768 + 64;
  MMN_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  MMN_LAUNCH(kfn, dim3(1), dim3(288), smem, stream, iters, n_mma, flags, out);
  MMN_CUDA(cudaGetLastError());
  return 0;
}


extern "C" int mmn_selftest_gemm_bf16(int M, int N, int K, const void* a, long long lda, const void* b, long long ldb,
                                      float* out_f32, void* out_bf16, void* out_bf16_t, void* stream) {
#ifdef MMN_EMU
  (void)M; (void)N; (void)K; (void)a; (void)lda; (void)b; (void)ldb; (void)out_f32; (void)out_bf16; (void)out_bf16_t; (void)stream;
  return fail("mmn_selftest_gemm_bf16: the wide-regime GEMM is not part of the host emulator");
#else
  int dev = 0, n_sms = 0;
  MMN_CUDA(cudaGetDevice(&dev));
  MMN_CUDA(cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, dev));
  wide::Epi e;
  memset(&e, 0, sizeof e);
  e.mode = wide::EPI_ACCUM_F32;
  e.scale = 1.f;
  e.out_f32 = out_f32; e.ld_f32 = N;
  e.out = (__nv_bfloat16*)out_bf16; e.ld_out = N;
  e.out_t = (__nv_bfloat16*)out_bf16_t; e.ld_out_t = M;
  return wide_gemm(n_sms, a, lda, b, ldb, M, N, K, e, stream);
#endif
}

extern "C" int mmn_selftest_gemm_bf16_mn(int M, int N, int K, const void* a, long long lda, const void* b, long long ldb,
                                         float* out_f32, void* stream) {
#ifdef MMN_EMU
  (void)M; (void)N; (void)K; (void)a; (void)lda; (void)b; (void)ldb; (void)out_f32; (void)stream;
  return fail("mmn_selftest_gemm_bf16_mn: the wide-regime GEMM is not part of the host emulator");
#else
  int dev = 0, n_sms = 0;
  MMN_CUDA(cudaGetDevice(&dev));
  MMN_CUDA(cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, dev));
  wide::Epi e;
  memset(&e, 0, sizeof e);
  e.mode = wide::EPI_ACCUM_F32;
  e.scale = 1.f;
  e.out_f32 = out_f32; e.ld_f32 = N;
  return wide_gemm(n_sms, a, lda, b, ldb, M, N, K, e, stream, "selftest", 1, 1);
#endif
}

extern "C" int64_t mmn_wide_launch_count(void) {
#ifdef MMN_EMU
  return 0;
#else
  return g_wide_launches;
#endif
}
