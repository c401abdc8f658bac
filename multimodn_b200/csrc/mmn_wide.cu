// mmn_wide.cu — translation unit of the wide regime (precision = bf16): tensor maps, the tcgen05 bf16 GEMM launcher and the
// layer-wise step orchestration (mmn_wide.cuh, mmn_wide_step.cuh).  CUDA only: not part of the host emulator.
#include "mmn_kernels.cuh"
#include <set>
#include <string>

#include "mmn_wide_step.cuh"
#include "mmn_host.h"

#include <string>

using namespace mmn;

// ------------------------------------------------------------------------------------------------
// wide regime (precision = bf16): layer-wise tcgen05 GEMMs (mmn_wide.cuh, mmn_wide_step.cuh)
// ------------------------------------------------------------------------------------------------
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
// Every kernel of the wide step is launched with programmatic stream serialisation (see pdl_entry in mmn_wide.cuh);
// MMN_WIDE_PDL=0 launches them fully serialised.
static const bool g_pdl = !(getenv("MMN_WIDE_PDL") && !strcmp(getenv("MMN_WIDE_PDL"), "0"));
template <typename... KArgs, typename... Args>
void wide_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = g_pdl ? 1 : 0;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);      // errors surface through cudaGetLastError at the call site
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
              const wide::Epi& epi_in, void* stream, const char* what = "gemm", int a_mn = 0, int b_mn = 0) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  // MMN_WIDE_EPI_DEBUG (development aid, results are wrong): none = drop the accumulators, nostore = compute but do not store
  static const int epi_debug = [] {
    const char* v = getenv("MMN_WIDE_EPI_DEBUG");
    return !v ? 0 : (!strcmp(v, "none") ? 1 : (!strcmp(v, "nostore") ? 2 : (!strcmp(v, "ldonly") ? 3 : (!strcmp(v, "aluonly") ? 4 : 0))));
  }();
  wide::Epi epi = epi_in;
  if (epi_debug == 1) epi.mode = wide::EPI_NONE;
  if (epi_debug == 3) epi.mode = wide::EPI_NONE + 1;
  if (epi_debug == 4) epi.mode = wide::EPI_NONE + 2;
  if (epi_debug == 2) { epi.out = nullptr; epi.out_t = nullptr; epi.out_f32 = nullptr; }
  if (g_wt.on) {       // timers: one category per GEMM kind, epilogue mode and shape
    static std::set<std::string> names;
    char buf[128];
    snprintf(buf, sizeof buf, "%s[epi %d%s %lldx%lldx%lld]", what, epi.mode, epi.accumulate ? "+" : "", M, N, K);
    what = names.insert(buf).first->c_str();
  }
  alignas(64) CUtensorMap ma, mb, mb_half;
  if (a_mn ? make_operand_map_mn(&ma, A, M, K, lda) : make_operand_map(&ma, A, M, K, lda, wide::BM)) return 1;
  if (b_mn ? make_operand_map_mn(&mb, B, N, K, ldb) : make_operand_map(&mb, B, N, K, ldb, wide::BN)) return 1;
  mb_half = mb;
  static bool attr_set = false;
  if (!attr_set) {
    MMN_CUDA(cudaFuncSetAttribute(wide::mmn_wide_gemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, wide::kSmemBytes));
    MMN_CUDA(cudaFuncSetAttribute(wide::mmn_wide_gemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, wide::kSmemBytes));
    attr_set = true;
  }
  const long long tiles = ((M + wide::BM - 1) / wide::BM) * ((N + wide::BN - 1) / wide::BN);
  const bool splittable = epi.mode == wide::EPI_ACCUM_F32 && epi.accumulate && !epi.out && !epi.out_t;
  // split-K factor for `n_tiles` output tiles on `n_units` CTAs (or CTA pairs): time ~ ceil(tiles * s / units) / s waves of
  // the unsplit tile, plus one more fp32 atomic pass over the output per split
  // last wave: the leftover tiles as d column slices when that shortens it; 1 = leave it.  Only halves are used: quarter
  // slices (N = 64 MMAs) shorten the wave count further on paper but measured 20 % slower on the N = 1024 GEMMs.
  auto pick_tail_div = [&](long long n_tiles, int n_units, bool pair) {
    if (n_tiles <= n_units) return 1;
    const long long rem = n_tiles % n_units;
    if (rem == 0) return 1;
    int best_d = 1;
    double best = 1.0;
    (void)pair;
    for (int d = 2; d <= 2; d *= 2) {
      const double w = (double)((rem * d + n_units - 1) / n_units) / d;
      if (w < best - 1e-9) { best = w; best_d = d; }
    }
    return best_d;
  };
  auto pick_splits = [&](long long n_tiles, int n_units) {
    int best_s = 1;
    if (!splittable || n_tiles >= n_units) return best_s;
    const long long kb = (K + wide::BK - 1) / wide::BK;
    double best = 1.0;
    for (int s = 2; s <= 16 && kb / s >= 16; ++s) {
      if ((long long)(s - 1) * ((kb + s - 1) / s) >= kb) continue;                 // no empty split
      // + the extra fp32 atomic passes over the output, measured ~3 % of a K = 8192 tile each
      const double cost = (double)((n_tiles * s + n_units - 1) / n_units) / s + 0.03 * (s - 1) * 128.0 / (double)kb;
      if (cost < best - 1e-9) { best = cost; best_s = s; }
    }
    return best_s;
  };
  // CTA pairs (cta_group::2): 256 x 256 tiles when both operands have the same orientation and the pairs can be kept busy;
  // each CTA then moves 32 KB instead of 48 KB per k-step through the L2 -> SM path that bounds the main loop.
  // MMN_WIDE_PAIR=0 keeps every GEMM on single CTAs.
  static const bool no_pair = getenv("MMN_WIDE_PAIR") && !strcmp(getenv("MMN_WIDE_PAIR"), "0");
  const long long ptiles = ((M + 2 * wide::BM - 1) / (2 * wide::BM)) * ((N + wide::BN - 1) / wide::BN);
  const int n_pairs = n_sms / 2;
  if (!no_pair && a_mn == b_mn && M >= 256 && (tiles >= n_sms || (splittable && 2 * ptiles >= n_pairs))) {
    const int psplits = pick_splits(ptiles, n_pairs);
    const int pairs = (int)std::min<long long>(ptiles * psplits, n_pairs);
    const int pair_tail = psplits == 1 ? pick_tail_div(ptiles, pairs, true) : 1;
    alignas(64) CUtensorMap mb_quarter = mb;          // K-major B: boxes of the tail slice's per-CTA share
    if (!b_mn && (make_operand_map(&mb_half, B, N, K, ldb, wide::BN / 2) ||
                  make_operand_map(&mb_quarter, B, N, K, ldb, wide::BN / std::max(pair_tail, 2) / 2)))
      return 1;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = dim3(2 * pairs); cfg.blockDim = dim3(wide::kThreads); cfg.dynamicSmemBytes = wide::kSmemBytes;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = g_pdl ? 1 : 0;
    cfg.attrs = at; cfg.numAttrs = 2;
    g_wt.begin(what);
    MMN_CUDA(cudaLaunchKernelEx(&cfg, wide::mmn_wide_gemm_kernel<true>, ma, mb_quarter, mb_half, (int)M, (int)N, (int)K, psplits, a_mn, b_mn,
                                pair_tail, epi));
    g_wt.end();
    MMN_CUDA(cudaGetLastError());
    ++g_wide_launches;
    return 0;
  }
  const int splits = pick_splits(tiles, n_sms);
  const int grid = (int)std::min<long long>(tiles * splits, n_sms);
  const int tail_halves = splits == 1 ? pick_tail_div(tiles, grid, false) : 1;
  if (tail_halves > 1 && !b_mn && make_operand_map(&mb_half, B, N, K, ldb, wide::BN / tail_halves)) return 1;
  g_wt.begin(what);
  wide_launch(wide::mmn_wide_gemm_kernel<false>, dim3(grid), dim3(wide::kThreads), wide::kSmemBytes, (cudaStream_t)stream, ma, mb, mb_half, (int)M, (int)N, (int)K, splits, a_mn, b_mn, tail_halves, epi);
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

}  // namespace
int mmn_wide_plan_init(mmn_plan* p) {

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
  // The decoders' data gradients with respect to the state, sum_d dz0_d . W0_d, as ONE GEMM: the transposed first-layer
  // images W0_d^T [S x h_d] side by side (K-concatenation).  Needs a hidden layer in every decoder (dz0_d is then a buffer of
  // ours, written as a column block of one matrix) and 16-byte aligned column blocks.
  bool cat = P.D >= 2;
  int cat_k = 0;
  for (int d = 0; d < P.D; ++d) {
    cat = cat && P.dec[d].n_layers >= 2 && P.dec[d].L[0].out_dim % 8 == 0 && P.dec[d].L[0].ktot == P.S;
    p->dec_cat_col[d] = cat_k;
    cat_k += P.dec[d].L[0].out_dim;
  }
  if (cat) {
    p->dec_cat_k = cat_k;
    const long long base = off;
    off += ((long long)P.S * cat_k + 127) & ~127ll;
    for (int d = 0; d < P.D; ++d) {
      p->wide_dec[d][0].wt = base + p->dec_cat_col[d];
      p->wide_dec[d][0].ldo = cat_k;
    }
  }
  p->wide_elems = off;
  cudaStream_t side;
  cudaEvent_t fork, done;
  MMN_CUDA(cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking));
  MMN_CUDA(cudaEventCreateWithFlags(&fork, cudaEventDisableTiming));
  MMN_CUDA(cudaEventCreateWithFlags(&done, cudaEventDisableTiming));
  p->side_stream = side; p->side_fork = fork; p->side_done = done;
  cudaStream_t dstr;
  cudaEvent_t dfork;
  MMN_CUDA(cudaStreamCreateWithFlags(&dstr, cudaStreamNonBlocking));
  MMN_CUDA(cudaEventCreateWithFlags(&dfork, cudaEventDisableTiming));
  p->dec_stream = dstr; p->dec_fork = dfork;
  for (int k = 0; k <= P.E; ++k) {
    cudaEvent_t ev;
    MMN_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    p->dec_done[k] = ev;
  }
  MMN_CUDA(cudaMalloc(&p->wide_w, (size_t)off * 2));
  MMN_CUDA(cudaMemset(p->wide_w, 0, (size_t)off * 2));
  return 0;
}

namespace {
bool ptr16(const void* p) { return (reinterpret_cast<size_t>(p) & 15) == 0; }

// decoder head, forward: picks the class-count instantiation and the straight-line variant when every row is 16-byte aligned
template <int C>
void launch_head_fwd_c(bool fast, unsigned grid, cudaStream_t st, const wide::Mat& h, const wide::bf16* Wb, long long ldk, const float* bias,
                       int act, long long rows, float* p_out) {
  if (fast) wide_launch(wide::wide_head_fwd_kernel<C, true>, dim3(grid), dim3(256), 0, st, h, Wb, ldk, bias, act, rows, p_out);
  else wide_launch(wide::wide_head_fwd_kernel<C, false>, dim3(grid), dim3(256), 0, st, h, Wb, ldk, bias, act, rows, p_out);
}
void launch_head_fwd(int C, int n_sms, cudaStream_t st, const wide::Mat& h, const wide::bf16* Wb, long long ldk, const float* bias, int act,
                     long long rows, float* p_out) {
  const bool fast = ptr16(h.p) && ptr16(Wb) && (h.ld & 7) == 0 && (ldk & 7) == 0 && (h.width & 7) == 0;
  const long long per_cta = 8 * wide::kHeadRows;
  // one resident wave (3 CTAs of 256 threads per SM at the kernel's 80 registers): the grid-stride loop then gives every warp
  // the same number of row groups instead of a second, mostly empty wave
  const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>((rows + per_cta - 1) / per_cta, 3ll * n_sms));
  switch (C) {
    case 1: launch_head_fwd_c<1>(fast, grid, st, h, Wb, ldk, bias, act, rows, p_out); break;
    case 2: launch_head_fwd_c<2>(fast, grid, st, h, Wb, ldk, bias, act, rows, p_out); break;
    case 3: launch_head_fwd_c<3>(fast, grid, st, h, Wb, ldk, bias, act, rows, p_out); break;
    default: launch_head_fwd_c<4>(fast, grid, st, h, Wb, ldk, bias, act, rows, p_out); break;
  }
}
template <int C>
void launch_head_backward_c(bool fast, dim3 grid, cudaStream_t st, const wide::Mat& dz, const wide::Mat& h, long long rows, float* gW,
                            long long ldw, float* gb, const wide::bf16* Wb, long long ldk, int act_prev, const wide::Mat& out) {
  if (fast) wide_launch(wide::wide_head_backward_kernel<C, true>, dim3(grid), dim3(256), 0, st, dz, h, rows, gW, ldw, gb, Wb, ldk, act_prev, out);
  else wide_launch(wide::wide_head_backward_kernel<C, false>, dim3(grid), dim3(256), 0, st, dz, h, rows, gW, ldw, gb, Wb, ldk, act_prev, out);
}
void launch_head_backward(int C, int n_sms, cudaStream_t st, const wide::Mat& dz, const wide::Mat& h, long long rows, float* gW, long long ldw,
                          float* gb, const wide::bf16* Wb, long long ldk, int act_prev, const wide::Mat& out) {
  const bool fast = ptr16(h.p) && ptr16(out.p) && ptr16(dz.p) && (h.ld & 7) == 0 && (out.ld & 7) == 0 && (dz.ld & 3) == 0 &&
                    (h.width & 7) == 0;
  // one wave: every CTA resident at once (kHeadBwdCtas per SM), each streaming its slice of the rows
  const long long gx = (h.width + 2047) / 2048;
  const dim3 grid((unsigned)gx, (unsigned)std::max<long long>(1, std::min<long long>((long long)wide::kHeadBwdCtas(C) * n_sms / gx, rows / 16)));
  switch (C) {
    case 1: launch_head_backward_c<1>(fast, grid, st, dz, h, rows, gW, ldw, gb, Wb, ldk, act_prev, out); break;
    case 2: launch_head_backward_c<2>(fast, grid, st, dz, h, rows, gW, ldw, gb, Wb, ldk, act_prev, out); break;
    case 3: launch_head_backward_c<3>(fast, grid, st, dz, h, rows, gW, ldw, gb, Wb, ldk, act_prev, out); break;
    default: launch_head_backward_c<4>(fast, grid, st, dz, h, rows, gW, ldw, gb, Wb, ldk, act_prev, out); break;
  }
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
  // The GEMMs are persistent (one CTA or CTA pair per SM) and statically scheduled.  In a data-parallel step the NCCL kernels
  // of the gradient all-reduce run beside the backward GEMMs; each of their CTAs takes an SM for the length of a collective,
  // the GEMM CTAs that should have run there start a whole collective later and the GEMM takes almost twice as long.
  // Backward GEMMs launched after the first gradient-ready event therefore leave SMs free for the collectives
  // (mmn_plan_set_comm_sms; profiles/dp_overlap_probe.sh on 2 GPUs: 148 SMs 3.82 ms/step, 136 3.72, 128 3.65, 120 3.80; 8 GPUs, NVLS with
  // 24 channels: 148 4.05, 112 3.96; MMN_WIDE_COMM_SMS overrides).
  static const int comm_sms_env = getenv("MMN_WIDE_COMM_SMS") ? atoi(getenv("MMN_WIDE_COMM_SMS")) : 0;
  const int comm_sms_want = comm_sms_env ? comm_sms_env : (plan->comm_sms ? plan->comm_sms : 128);
  const int comm_sms = plan->n_grad_events > 0 && comm_sms_want > 1 && comm_sms_want < n_sms ? (comm_sms_want & ~1) : n_sms;
  int gemm_sms = n_sms;            // SMs the next GEMM may use
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
  // data gradient dIn[:, col0 : col0 + n] = dZ . W[:, col0 : col0 + n] through the transposed bf16 copy of W (K-major B
  // operand: one 256-row TMA box per stage).  Using W in place as an MN-major operand (four 64 x 64 boxes per stage) works
  // when the sub-block starts on a 16-byte boundary but measured 10 % slower on config 4, so it is not used.
  auto needs_wt = [](const DevLayer&) { return true; };
  auto dgrad = [&](const Mat& dz, const DevLayer& l, const mmn_plan::WL& w, int col0, int n, const Epi& e, const char* what) -> int {
    if (needs_wt(l)) return wide_gemm(gemm_sms, dz.p, dz.ld, wbase + w.wt + (long long)col0 * w.ldo, w.ldo, B, n, l.out_dim, e, stream, what);
    return wide_gemm(gemm_sms, dz.p, dz.ld, wbase + w.w + col0, w.ldk, B, n, l.out_dim, e, stream, what, 0, 1);
  };

  void* cast_jobs_dev = ar.take(sizeof(CastJobs));
  // ---- 0. bf16 copies of the weights (both orientations) ----
  if (!dry) {
    CastJobs jobs;
    int n_jobs = 0, max_out = 0, max_k = 0;
    auto add = [&](const DevLayer& l, const mmn_plan::WL& w) {
      CastJob& j = jobs.job[n_jobs++];
      j.W = a.params + l.w_off; j.Wb = wbase + w.w; j.WbT = needs_wt(l) ? wbase + w.wt : nullptr;
      j.out = l.out_dim; j.ktot = l.ktot; j.ldk = w.ldk; j.ldo = w.ldo;
      max_out = std::max(max_out, l.out_dim); max_k = std::max(max_k, l.ktot);
    };
    for (int e = 0; e < E; ++e)
      for (int j = 0; j < P.enc[e].n_layers; ++j) add(P.enc[e].L[j], plan->wide_enc[e][j]);
    for (int d = 0; d < D; ++d)
      for (int j = 0; j < P.dec[d].n_layers; ++j) add(P.dec[d].L[j], plan->wide_dec[d][j]);
    // the job table travels through the head of the workspace (re-sent every call: params may move between calls)
    MMN_CUDA(cudaMemcpyAsync(cast_jobs_dev, &jobs, sizeof(CastJob) * n_jobs, cudaMemcpyHostToDevice, stream));
    g_wt.begin("cast_weight");
    wide_launch(wide_cast_weights_kernel, dim3((unsigned)((max_k + 63) / 64), (unsigned)((max_out + 63) / 64), (unsigned)n_jobs), dim3(256), 0, stream, 
        (const CastJobs*)cast_jobs_dev);
    if (launched()) return 1;
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
  // Training runs every decoder ONCE over the states of all steps: s_0 .. s_L are the row blocks of one [(L + 1) B x S]
  // matrix, so each decoder layer is a single GEMM with (L + 1) B rows instead of L + 1 launches that each leave a
  // partial last wave and pay their own prologue (decoders_forward_all / decoders_backward_all below).
  const long long R = (long long)(L + 1) * B;
  float* Pout = (float*)ar.take(sizeof(float) * (size_t)(TRAIN ? R : B) * maxC);
  auto rows_of = [](const Mat& m, long long r0) {      // the row block starting at r0 (row-major orientation only)
    Mat v = m;
    v.p = m.p ? m.p + r0 * m.ld : nullptr;
    v.t = nullptr;
    return v;
  };
  std::vector<Mat> Sk(L + 1);
  Mat Sall = Mat();
  if (TRAIN) {
    Sall = ar.mat(R, S);
    for (int k = 0; k <= L; ++k) Sk[k] = rows_of(Sall, (long long)k * B);
  } else {
    const Mat s0 = ar.mat(B, S), s1 = ar.mat(B, S);
    for (int k = 0; k <= L; ++k) Sk[k] = (k & 1) ? s1 : s0;
  }
  // per (step, module, layer) activations kept for the backward pass
  std::vector<Mat> enc_in((size_t)(L + 1) * MMN_MAX_LAYERS);
  if (!dry) {
    MMN_CUDA(cudaMemsetAsync(present, 1, (size_t)(L + 1) * B, stream));
    MMN_CUDA(cudaMemsetAsync(sc_sum, 0, sizeof(float) * (size_t)std::max(E, 1), stream));
    g_wt.begin("init_state");
    wide_launch(wide_init_state_kernel, dim3(tgrid(B, S)), tb, 0, stream, a.params + P.init_off, B, Sk[0]);
    if (launched()) return 1;
  }
  const size_t scratch_mark = ar.off;

  auto epi0 = [] {
    Epi e;
    memset(&e, 0, sizeof e);
    e.scale = 1.f;
    return e;
  };

  // ---- decoders on s_k, one step at a time (forward-only calls: the states are not kept) ----
  const cudaStream_t ds = stream;
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
          e.out = out.p; e.ld_out = out.ld;
        } else {
          e.out_f32 = Pout; e.ld_f32 = dec.C;
        }
        if (!dry) {
          if (last && j > 0 && dec.C <= 4) {   // decoder head: skinny, bandwidth-bound kernel instead of a tensor-core tile
            g_wt.begin("head_fwd");
            launch_head_fwd(dec.C, n_sms, ds, in, wbase + w.w, w.ldk, a.params + ly.b_off, ly.act, B, Pout);
            if (launched()) return 1;
          } else if (wide_gemm(gemm_sms, in.p, in.ld, wbase + w.w, w.ldk, B, ly.out_dim, ly.ktot, e, ds, "gemm fwd")) {
            return 1;
          }
        }
        in = out;
      }
      LossArgs la;
      memset(&la, 0, sizeof la);
      la.p = Pout; la.ldp = dec.C; la.C = dec.C; la.act = dec.L[dec.n_layers - 1].act; la.D = D; la.d = d;
      la.hist_row = hist_row; la.n_mat_rows = E + 1; la.rows = B; la.targets = a.targets; la.target_error = a.target_error;
      la.present = k == 0 ? nullptr : present + (size_t)k * B;
      la.skip = skip;
      la.metrics = a.metrics; la.inv_rows_global = a.inv_rows_global;
      la.predictions = a.predictions ? a.predictions + ((long long)hist_row * D + d) * a.pred_ld : nullptr;
      if (a.last_outputs && is_last_enc) { la.last_outputs = a.last_outputs; la.ld_last = P.sumC; la.out_off = dec.out_off; }
      if (!dry) {
        g_wt.begin("decoder_loss");
        wide_launch(wide_decoder_loss_kernel, dim3((unsigned)((B + 255) / 256)), dim3(256), 0, ds, la);
        if (launched()) return 1;
      }
    }
    return 0;
  };

  struct StepMeta { int hist_row; bool is_last_enc; const int* skip; };
  std::vector<StepMeta> meta(L + 1);
  meta[0] = {0, false, nullptr};
  if (!TRAIN && decoders_forward(0, 0, false, nullptr)) return 1;
  if (!dry && !TRAIN) {          // (training: the bookkeeping of all steps is one launch after the walk)
    g_wt.begin("finalize");
    wide_launch(wide_finalize_kernel, dim3(1), dim3(256), 0, stream, present, B, 0, 0, 0, nullptr, nullptr, S, a.inv_rows_global,
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
      if (enc.L[0].has_state) {         // features and state side by side: one launch
        g_wt.begin("input_x");
        const dim3 gx = tgrid(B, enc.F), gs = tgrid(B, S);
        wide_launch(wide_input_xs_kernel, dim3(gx.x + gs.x, gx.y), tb, 0, stream, a.x[pos], a.x_ld[pos], B, enc.F, Sk[k - 1], in,
                    enc.L[0].in_dim, pres, drop);
        if (launched()) return 1;
      } else {
        g_wt.begin("input_x");
        wide_launch(wide_input_x_kernel, dim3(tgrid(B, enc.F)), tb, 0, stream, a.x[pos], a.x_ld[pos], B, enc.F, in, pres, drop);
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
        if (wide_gemm(gemm_sms, in.p, in.ld, wbase + w.w, w.ldk, B, ly.out_dim, ly.ktot, ep, stream)) return 1;
        if (!last && enc.L[j + 1].has_state) {
          g_wt.begin("input_state");
          wide_launch(wide_input_state_kernel, dim3(tgrid(B, S)), tb, 0, stream, Sk[k - 1], B, next, enc.L[j + 1].in_dim, nodrop);
          if (launched()) return 1;
        }
      }
      in = next;
    }
    if (!dry && !TRAIN) {
      g_wt.begin("finalize");
      wide_launch(wide_finalize_kernel, dim3(1), dim3(256), 0, stream, pres, B, k, e + 1, e, skip, TRAIN ? sc_sum + e : nullptr, S, a.inv_rows_global,
                                                  a.metrics ? a.metrics + met_present(P, 0) : nullptr,
                                                  (TRAIN && a.metrics) ? a.metrics + met_sc(P, 0) : nullptr,
                                                  TRAIN ? a.grads + P.n_params : nullptr);
      if (launched()) return 1;
    }
    meta[k] = {e + 1, e == E - 1, skip};
    if (!TRAIN && decoders_forward(k, e + 1, e == E - 1, skip)) return 1;
    if (!TRAIN) ar.off = scratch_mark;
  }
  if (TRAIN && !dry) {           // per-step bookkeeping (present-row counts, state-change means), all steps in one launch
    FinalizeSteps fs;
    memset(&fs, 0, sizeof fs);
    fs.present = present; fs.rows = B; fs.sc_sum = sc_sum; fs.S = S; fs.inv_rows_global = a.inv_rows_global;
    fs.met_present = a.metrics ? a.metrics + met_present(P, 0) : nullptr;
    fs.met_sc = a.metrics ? a.metrics + met_sc(P, 0) : nullptr;
    fs.grad_tail = a.grads + P.n_params;
    for (int k = 0; k <= L; ++k) {
      fs.step[k].hist_row = meta[k].hist_row;
      fs.step[k].e = k == 0 ? 0 : a.seq_enc[k - 1];
      fs.step[k].skip = meta[k].skip;
    }
    g_wt.begin("finalize");
    wide_launch(wide_finalize_steps_kernel, dim3((unsigned)(L + 1)), dim3(256), 0, stream, fs);
    if (launched()) return 1;
  }
  // ---- training: every decoder, once, over the (L + 1) B state rows ----
  std::vector<Mat> dec_hall((size_t)D * MMN_MAX_LAYERS);     // hidden activations [R x width] of decoder d, layer j
  std::vector<Mat> dec_dzall(D);                             // loss gradient [R x C] of decoder d
  if (TRAIN) {
    for (int d = 0; d < D; ++d) {
      const DevDecoder& dec = P.dec[d];
      Mat in = Sall;
      for (int j = 0; j < dec.n_layers; ++j) {
        const DevLayer& ly = dec.L[j];
        const mmn_plan::WL& w = plan->wide_dec[d][j];
        const bool last = j == dec.n_layers - 1;
        Epi e = epi0();
        e.mode = EPI_STORE; e.act = ly.act; e.bias = a.params + ly.b_off;
        Mat out = Mat();
        if (!last) {
          out = ar.mat(R, ly.out_dim);
          dec_hall[(size_t)d * MMN_MAX_LAYERS + j] = out;
          e.out = out.p; e.ld_out = out.ld;
        } else {
          e.out_f32 = Pout; e.ld_f32 = dec.C;
        }
        if (!dry) {
          if (last && j > 0 && dec.C <= 4) {   // decoder head: skinny, bandwidth-bound kernel instead of a tensor-core tile
            g_wt.begin("head_fwd");
            launch_head_fwd(dec.C, n_sms, stream, in, wbase + w.w, w.ldk, a.params + ly.b_off, ly.act, R, Pout);
            if (launched()) return 1;
          } else if (wide_gemm(gemm_sms, in.p, in.ld, wbase + w.w, w.ldk, R, ly.out_dim, ly.ktot, e, stream, "gemm fwd")) {
            return 1;
          }
        }
        in = out;
      }
      dec_dzall[d] = ar.mat(R, dec.C);
      // the pitch padding of dz is read by the TMA unit as part of full 16-byte rows: keep it finite
      if (!dry) MMN_CUDA(cudaMemsetAsync(dec_dzall[d].p, 0, (size_t)R * dec_dzall[d].ld * 2, stream));
      {          // the per-row epilogue of every step in one launch (history row, mask and skip flag differ per step)
        LossSteps ls;
        memset(&ls, 0, sizeof ls);
        LossArgs& la = ls.base;
        la.p = Pout; la.ldp = dec.C; la.C = dec.C; la.act = dec.L[dec.n_layers - 1].act; la.D = D; la.d = d;
        la.n_mat_rows = E + 1; la.rows = B; la.targets = a.targets; la.target_error = a.target_error;
        la.metrics = a.metrics; la.inv_rows_global = a.inv_rows_global;
        la.predictions = a.predictions; ls.pred_ld = a.pred_ld;
        if (a.last_outputs) { la.last_outputs = a.last_outputs; la.ld_last = P.sumC; la.out_off = dec.out_off; }
        la.coef = a.c_err;
        la.dz = dec_dzall[d];
        ls.n_steps = L + 1;
        for (int k = 0; k <= L; ++k) {
          ls.step[k].hist_row = meta[k].hist_row;
          ls.step[k].is_last = meta[k].is_last_enc ? 1 : 0;
          ls.step[k].present = k == 0 ? nullptr : present + (size_t)k * B;
          ls.step[k].skip = meta[k].skip;
        }
        if (!dry) {
          g_wt.begin("decoder_loss");
          wide_launch(wide_decoder_loss_steps_kernel, dim3((unsigned)((B + 255) / 256), (unsigned)(L + 1)), dim3(256), 0, stream, ls);
          if (launched()) return 1;
        }
      }
    }
  }
  if (a.final_state && !dry) {
    g_wt.begin("state_out");
    wide_launch(wide_state_out_kernel, dim3((unsigned)std::min<long long>((B * S + 255) / 256, 4096)), dim3(256), 0, stream, Sk[L], B, a.final_state);
    if (launched()) return 1;
  }

  if (TRAIN) {
    // =====================================================================================================
    // backward: replay the sequence in reverse
    // =====================================================================================================
    unsigned ev_done = 0;        // encoders whose gradient-ready event has been recorded this step
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
    // the bias-gradient reductions only read dz: they run on the plan's side stream, next to the weight-gradient GEMM of
    // the same layer; the caller's stream waits for them before anything may overwrite a dz buffer (join_side)
    const cudaStream_t side = (cudaStream_t)plan->side_stream;
    const bool use_side = side && !g_wt.on;
    bool side_pending = false;
    auto join_side = [&]() -> int {
      if (side_pending) { MMN_CUDA(cudaStreamWaitEvent(stream, (cudaEvent_t)plan->side_done, 0)); side_pending = false; }
      return 0;
    };
    auto layer_param_grads = [&](const DevLayer& ly, const Mat& dz, const Mat& in) -> int {
      if (dry) return 0;
      const dim3 bgrid((unsigned)((ly.out_dim + 63) / 64), (unsigned)std::max<long long>(1, std::min<long long>(32, B / 256)));
      if (use_side) {
        MMN_CUDA(cudaEventRecord((cudaEvent_t)plan->side_fork, stream));
        MMN_CUDA(cudaStreamWaitEvent(side, (cudaEvent_t)plan->side_fork, 0));
        wide_launch(wide_bias_grad_kernel, dim3(bgrid), dim3(256), 0, side, dz.p, dz.ld, B, ly.out_dim, a.grads + ly.b_off);
        MMN_CUDA(cudaGetLastError());
        ++g_wide_launches;
        MMN_CUDA(cudaEventRecord((cudaEvent_t)plan->side_done, side));
        side_pending = true;
      } else {
        g_wt.begin("bias_grad");
        wide_launch(wide_bias_grad_kernel, dim3(bgrid), dim3(256), 0, stream, dz.p, dz.ld, B, ly.out_dim, a.grads + ly.b_off);
        if (launched()) return 1;
      }
      Epi e = epi0();
      e.mode = EPI_ACCUM_F32; e.accumulate = 1;
      e.out_f32 = a.grads + ly.w_off; e.ld_f32 = ly.ktot;
      return wide_gemm(gemm_sms, dz.p, dz.ld, in.p, in.ld, ly.out_dim, ly.ktot, B, e, stream, "gemm wgrad", 1, 1);
    };
    // Decoders, backward, over the (L + 1) B rows of all steps at once.  On the caller's stream: each decoder's chain from
    // its head down to the gradient dz_0 of its first layer, then DS (+)= dz_0 . W_0, the decoders' gradient with respect
    // to every state (row block k is added to G where the reverse replay reaches step k).  The first layers' parameter
    // gradients only read dz_0 and the states: they run on the plan's decoder stream next to the encoders' backward GEMMs.
    const bool dec_side = !dry && plan->dec_stream && !g_wt.on;
    const cudaStream_t dstream = dec_side ? (cudaStream_t)plan->dec_stream : stream;
    float* DS = (float*)ar.take(sizeof(float) * (size_t)R * S);
    {
      bool deep = false;
      for (int d = 0; d < D; ++d) deep |= P.dec[d].n_layers > 2;
      Mat dzdec[2];                               // intermediates of decoders deeper than two layers
      if (deep) { dzdec[0] = ar.mat(R, maxW); dzdec[1] = ar.mat(R, maxW); }
      std::vector<Mat> dz0(D);                    // dz of every decoder's first layer
      const bool cat = plan->dec_cat_k > 0;       // column blocks of one matrix: the data gradient below is a single GEMM
      Mat dz0cat = Mat();
      if (cat) dz0cat = ar.mat(R, plan->dec_cat_k);
      for (int d = 0; d < D; ++d) {
        if (cat) {
          dz0[d] = dz0cat;
          dz0[d].p = dz0cat.p ? dz0cat.p + plan->dec_cat_col[d] : nullptr;
          dz0[d].width = P.dec[d].L[0].out_dim;
        } else {
          dz0[d] = P.dec[d].n_layers > 1 ? ar.mat(R, P.dec[d].L[0].out_dim) : dec_dzall[d];
        }
      }
      const auto bias_grid = [&](int out_dim) {
        return dim3((unsigned)((out_dim + 63) / 64), (unsigned)std::max<long long>(1, std::min<long long>(32, R / 256)));
      };
      for (int d = 0; d < D && !dry; ++d) {
        const DevDecoder& dec = P.dec[d];
        Mat dz = dec_dzall[d];
        int cur = 0;
        for (int j = dec.n_layers - 1; j >= 1; --j) {
          const DevLayer& ly = dec.L[j];
          const mmn_plan::WL& w = plan->wide_dec[d][j];
          const Mat in = dec_hall[(size_t)d * MMN_MAX_LAYERS + j - 1];
          // dz of the layer below: the kept buffer when that layer is the first one
          const Mat nz = j == 1 ? dz0[d] : view(dzdec[cur], ly.in_dim);
          if (j == dec.n_layers - 1 && dec.C <= 4) {       // decoder head (see the forward pass)
            g_wt.begin("head_backward");
            launch_head_backward(dec.C, n_sms, stream, dz, in, R, a.grads + ly.w_off, ly.ktot, a.grads + ly.b_off, wbase + w.w, w.ldk,
                                 dec.L[j - 1].act, nz);
            if (launched()) return 1;
          } else {
            g_wt.begin("bias_grad");
            wide_launch(wide_bias_grad_kernel, dim3(bias_grid(ly.out_dim)), dim3(256), 0, stream, dz.p, dz.ld, R, ly.out_dim, a.grads + ly.b_off);
            if (launched()) return 1;
            Epi e = epi0();
            e.mode = EPI_ACCUM_F32; e.accumulate = 1;
            e.out_f32 = a.grads + ly.w_off; e.ld_f32 = ly.ktot;
            if (wide_gemm(gemm_sms, dz.p, dz.ld, in.p, in.ld, ly.out_dim, ly.ktot, R, e, stream, "gemm wgrad", 1, 1)) return 1;
            Epi ed = epi0();
            ed.mode = EPI_DACT; ed.act = dec.L[j - 1].act;
            ed.aux = in.p; ed.ld_aux = in.ld;
            ed.out = nz.p; ed.ld_out = nz.ld;
            if (wide_gemm(gemm_sms, dz.p, dz.ld, wbase + w.wt, w.ldo, R, ly.in_dim, ly.out_dim, ed, stream, "gemm dgrad")) return 1;
          }
          dz = nz;
          cur ^= 1;
        }
        if (cat) continue;
        // DS (+)= dz_0 . W_0
        const DevLayer& l0 = dec.L[0];
        const mmn_plan::WL& w0 = plan->wide_dec[d][0];
        Epi e = epi0();
        e.mode = EPI_ACCUM_F32; e.accumulate = d > 0;
        e.out_f32 = DS; e.ld_f32 = S;
        if (wide_gemm(gemm_sms, dz0[d].p, dz0[d].ld, wbase + w0.wt, w0.ldo, R, S, l0.out_dim, e, stream, "gemm dgrad")) return 1;
      }
      if (cat && !dry) {         // DS = [dz0_0 | dz0_1 | ...] . [W0_0; W0_1; ...]
        Epi e = epi0();
        e.mode = EPI_ACCUM_F32; e.accumulate = 0;
        e.out_f32 = DS; e.ld_f32 = S;
        if (wide_gemm(gemm_sms, dz0cat.p, dz0cat.ld, wbase + plan->wide_dec[0][0].wt, plan->dec_cat_k, R, S, plan->dec_cat_k, e, stream,
                      "gemm dgrad"))
          return 1;
      }
      gemm_sms = comm_sms;         // from here on collectives may be in flight
      if (!dry) {
        if (dec_side) {
          MMN_CUDA(cudaEventRecord((cudaEvent_t)plan->dec_fork, stream));
          MMN_CUDA(cudaStreamWaitEvent(dstream, (cudaEvent_t)plan->dec_fork, 0));
        }
        for (int d = 0; d < D; ++d) {             // first layers: bias and weight gradients, contraction over all R rows
          const DevLayer& l0 = P.dec[d].L[0];
          g_wt.begin("bias_grad");
          wide_launch(wide_bias_grad_kernel, dim3(bias_grid(l0.out_dim)), dim3(256), 0, dstream, dz0[d].p, dz0[d].ld, R, l0.out_dim, a.grads + l0.b_off);
          if (launched()) return 1;
          Epi e = epi0();
          e.mode = EPI_ACCUM_F32; e.accumulate = 1;
          e.out_f32 = a.grads + l0.w_off; e.ld_f32 = l0.ktot;
          if (wide_gemm(gemm_sms, dz0[d].p, dz0[d].ld, Sall.p, Sall.ld, l0.out_dim, l0.ktot, R, e, dstream, "gemm wgrad", 1, 1)) return 1;
        }
        if (dec_side) MMN_CUDA(cudaEventRecord((cudaEvent_t)plan->dec_done[0], dstream));
        // every decoder's parameter gradients are final (heads and deeper layers were finished on the caller's stream before the fork)
        if (plan->n_grad_events > E + 1) MMN_CUDA(cudaEventRecord((cudaEvent_t)plan->grad_events[E + 1], dstream));
      }
    }
    for (int k = L; k >= 1; --k) {
      const int e = a.seq_enc[k - 1];
      const DevEncoder& enc = P.enc[e];
      const int* skip = a.skip_flags ? a.skip_flags + (k - 1) : nullptr;
      unsigned char* pres = present + (size_t)k * B;
      const int nl = enc.n_layers;
      int cur = 0;
      Mat dz = view(dzbuf[cur], S);
      cur ^= 1;
      if (!dry) {
        if (join_side()) return 1;
        g_wt.begin("state_grad");
        wide_launch(wide_state_grad_kernel, dim3(tgrid(B, S)), tb, 0, stream, G, DS + (size_t)k * B * S, Sk[k], Sk[k - 1], pres, skip, a.c_sc,
                                                               enc.L[nl - 1].act, B, dz);
        if (launched()) return 1;
      }
      for (int j = nl - 1; j >= 0; --j) {
        const DevLayer& ly = enc.L[j];
        const mmn_plan::WL& w = plan->wide_enc[e][j];
        const Mat in = enc_in[(size_t)k * MMN_MAX_LAYERS + j];
        if (layer_param_grads(ly, dz, in)) return 1;
        if (!dry && plan->grad_layer_event(e, j) >= 0) {      // this layer's weight and bias gradients are final
          if (join_side()) return 1;                          // (the bias sums ran beside the GEMM and are long done)
          MMN_CUDA(cudaEventRecord((cudaEvent_t)plan->grad_events[plan->grad_layer_event(e, j)], stream));
        }
        if (ly.has_state && !dry) {
          // carry into G: present rows take dz W_s (through the dropout mask), absent rows keep G; u_k is removed in the same epilogue
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
          ep.aux = Sk[k].p; ep.ld_aux = Sk[k].ld; ep.aux2 = Sk[k - 1].p; ep.ld_aux2 = Sk[k - 1].ld; ep.c_sc = a.c_sc;
          if (dgrad(dz, ly, w, ly.in_dim, S, ep, "gemm dgrad")) return 1;
        }
        if (j > 0) {
          const Mat nz = view(dzbuf[cur], ly.in_dim);
          Epi ep = epi0();
          if (!dry && join_side()) return 1;
          ep.mode = EPI_DACT; ep.act = enc.L[j - 1].act;
          ep.aux = in.p; ep.ld_aux = in.ld;
          ep.out = nz.p; ep.ld_out = nz.ld; ep.out_t = nz.t; ep.ld_out_t = nz.ldt;
          if (!dry && dgrad(dz, ly, w, 0, ly.in_dim, ep, "gemm dgrad")) return 1;
          dz = nz;
          cur ^= 1;
        }
      }
      // encoder e's parameter gradients are final: let the caller start reducing them across ranks
      if (!dry && plan->n_grad_events) {
        if (join_side()) return 1;
        MMN_CUDA(cudaEventRecord((cudaEvent_t)plan->grad_events[e], stream));
        ev_done |= 1u << e;
      }
    }
    if (!dry) {
      if (join_side()) return 1;
      g_wt.begin("colsum_f32");
      wide_launch(wide_colsum_f32_kernel, dim3((unsigned)((S + 31) / 32), 16), dim3(256), 0, stream, G, DS, B, S, a.grads + P.init_off);
      if (launched()) return 1;
      if (dec_side) MMN_CUDA(cudaStreamWaitEvent(stream, (cudaEvent_t)plan->dec_done[0], 0));      // the decoders' parameter gradients
      for (int i = 0; i < std::min(plan->n_grad_events, E + 1); ++i)      // everything; encoders that took no step
        if (i == E || !(ev_done & (1u << i))) MMN_CUDA(cudaEventRecord((cudaEvent_t)plan->grad_events[i], stream));
      for (int e = 0; e < E; ++e)                        // layers of encoders that took no step
        for (int j = 0; j < P.enc[e].n_layers; ++j)
          if (!(ev_done & (1u << e)) && plan->grad_layer_event(e, j) >= 0)
            MMN_CUDA(cudaEventRecord((cudaEvent_t)plan->grad_events[plan->grad_layer_event(e, j)], stream));
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

int64_t mmn_wide_workspace_bytes(const mmn_plan* plan, int64_t n_rows, bool train) { return wide_workspace_bytes(plan, n_rows, train); }
int mmn_wide_step(const mmn_plan* plan, const StepArgs& a, void* ws, size_t ws_bytes, void* stream, bool train) {
  return train ? wide_step<true>(plan, a, ws, ws_bytes, stream, false, nullptr) : wide_step<false>(plan, a, ws, ws_bytes, stream, false, nullptr);
}

extern "C" int mmn_selftest_gemm_bf16(int M, int N, int K, const void* a, long long lda, const void* b, long long ldb,
                                      float* out_f32, void* out_bf16, void* out_bf16_t, void* stream) {
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
}

extern "C" int mmn_selftest_gemm_bf16_mn(int M, int N, int K, const void* a, long long lda, const void* b, long long ldb,
                                         float* out_f32, void* stream) {
  int dev = 0, n_sms = 0;
  MMN_CUDA(cudaGetDevice(&dev));
  MMN_CUDA(cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, dev));
  wide::Epi e;
  memset(&e, 0, sizeof e);
  e.mode = wide::EPI_ACCUM_F32;
  e.scale = 1.f;
  e.out_f32 = out_f32; e.ld_f32 = N;
  return wide_gemm(n_sms, a, lda, b, ldb, M, N, K, e, stream, "selftest", 1, 1);
}

extern "C" int64_t mmn_wide_launch_count(void) {
  return g_wide_launches;
}
