// mmn_nb.cu — translation unit of the bf16 tile kernel (mmn_nb.cuh): plan lowering (which models qualify, shared-memory
// arena layout), the per-step weight imaging launch and the step launch.
#include "mmn_nb.cuh"
#include "mmn_host.h"

using namespace mmn;
using namespace mmn::nb;

namespace {
int round16(int v) { return (v + 15) & ~15; }

// instantiations: (k16-steps of the state, k16-steps of the widest hidden layer)
struct Inst { int kss, ksh; };
const Inst kInsts[] = {{1, 1}, {4, 2}, {4, 4}};

// One Linear layer of the plan.  in_pad / out_pad: padded widths (multiples of 16) of the chained input / of the output inside
// the instantiation; spad: padded state width.
void fill_layer(NbLayer& o, const DevLayer& l, int in_pad, int out_pad, int spad, bool x_type, int& arena) {
  o.N = l.out_dim;
  o.ka = l.in_dim;
  o.ka_pad = in_pad;
  o.n_pad = out_pad;
  o.has_state = l.has_state;
  o.x_type = x_type ? 1 : 0;
  o.act = l.act;
  o.ktot = l.ktot;
  o.w_off = l.w_off;
  o.b_off = l.b_off;
  o.pitch = (o.ka_pad + (l.has_state ? spad : 0)) * 2 + 16;      // odd multiple of 16 bytes: conflict-free ldmatrix
  o.img_off = arena;
  arena += o.n_pad * o.pitch;
}

template <int KSS, int KSH>
int launch_inst(const mmn_plan* plan, const NbArgs& args, int grid, size_t smem, void* stream, bool train) {
  static bool configured[2] = {false, false};
  if (train) {
    auto kfn = mmn_nb_step_kernel<KSS, KSH, true>;
    if (!configured[1]) {
      MMN_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan->max_smem));
      configured[1] = true;
    }
    MMN_LAUNCH(kfn, dim3(grid), dim3(kThreadsNb), smem, stream, args);
  } else {
    auto kfn = mmn_nb_step_kernel<KSS, KSH, false>;
    if (!configured[0]) {
      MMN_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan->max_smem));
      configured[0] = true;
    }
    MMN_LAUNCH(kfn, dim3(grid), dim3(kThreadsNb), smem, stream, args);
  }
  MMN_CUDA(cudaGetLastError());
  return 0;
}
}  // namespace

// Lowers the model to the tile kernel's plan; returns false (with `why` filled) when the model does not qualify.
bool mmn_nb_build(const DevPlan& P, int max_smem, NbPlan& N, const char** why) {
  static const char* reasons[] = {"state wider than 64", "more than 3 Linear layers in an encoder or decoder",
                                  "a hidden layer wider than 64", "a decoder with more than 8 classes",
                                  "weights + staging exceed the shared memory of one SM"};
  memset(&N, 0, sizeof N);
  if (P.S > kMaxW) { *why = reasons[0]; return false; }
  N.S = P.S; N.E = P.E; N.D = P.D; N.sumC = P.sumC; N.n_metrics = P.n_metrics;
  N.init_param_off = P.init_off; N.n_params = P.n_params;
  // the instantiation: k16-steps of the state and of the widest hidden layer.  The x-fed first layer keeps 16 KSH output
  // columns in registers, so a 1-layer encoder (whose output is the state itself) needs KSH >= the state's k-steps.
  int max_hidden = 16;
  bool one_layer_encoder = false;
  for (int e = 0; e < P.E; ++e) {
    const DevEncoder& s = P.enc[e];
    if (s.n_layers > kMaxL) { *why = reasons[1]; return false; }
    one_layer_encoder |= s.n_layers == 1;
    for (int j = 0; j + 1 < s.n_layers; ++j) {
      if (s.L[j].out_dim > kMaxW) { *why = reasons[2]; return false; }
      max_hidden = std::max(max_hidden, round16(s.L[j].out_dim));
    }
  }
  for (int d = 0; d < P.D; ++d) {
    const DevDecoder& s = P.dec[d];
    if (s.n_layers > kMaxL) { *why = reasons[1]; return false; }
    if (s.C > kMaxC) { *why = reasons[3]; return false; }
    for (int j = 0; j + 1 < s.n_layers; ++j) {
      if (s.L[j].out_dim > kMaxW) { *why = reasons[2]; return false; }
      max_hidden = std::max(max_hidden, round16(s.L[j].out_dim));
    }
  }
  const int need_kss = round16(P.S) / 16, need_ksh = std::max(max_hidden / 16, one_layer_encoder ? need_kss : 0);
  const Inst* inst = nullptr;
  for (const Inst& c : kInsts)
    if (need_kss <= c.kss && need_ksh <= c.ksh) { inst = &c; break; }
  if (!inst) { *why = reasons[2]; return false; }
  N.kss = inst->kss;
  N.ksh = inst->ksh;
  const int spad = 16 * inst->kss, hpad = 16 * inst->ksh;

  int arena = 0, xs = 0;
  for (int e = 0; e < P.E; ++e) {
    const DevEncoder& s = P.enc[e];
    NbEnc& d = N.enc[e];
    d.F = s.F; d.n_layers = s.n_layers; d.p_drop = s.p_drop;
    for (int j = 0; j < s.n_layers; ++j) {
      const bool last = j == s.n_layers - 1;
      const int in_pad = j == 0 ? round16(s.F) : hpad;
      // a 1-layer encoder's image serves the forward (16 KSH output tiles) and the carry (contraction over 16 KSS rows)
      const int out_pad = !last ? hpad : (j == 0 ? std::max(spad, hpad) : spad);
      fill_layer(d.L[j], s.L[j], in_pad, out_pad, spad, j == 0, arena);
    }
    d.xs_off = xs;
    xs += round16(s.F) / 16;
  }
  N.xs_steps = xs;
  for (int dd = 0; dd < P.D; ++dd) {
    const DevDecoder& s = P.dec[dd];
    NbDec& d = N.dec[dd];
    d.C = s.C; d.n_layers = s.n_layers; d.out_off = s.out_off;
    for (int j = 0; j < s.n_layers; ++j) {
      const bool last = j == s.n_layers - 1;
      fill_layer(d.L[j], s.L[j], j == 0 ? spad : hpad, last ? 16 : hpad, spad, false, arena);
    }
  }
  // biases (fp32, one per image row) and the initial state behind the images
  arena = (arena + 15) & ~15;
  for (int e = 0; e < P.E; ++e)
    for (int j = 0; j < N.enc[e].n_layers; ++j) { N.enc[e].L[j].bias_off = arena; arena += 4 * N.enc[e].L[j].n_pad; }
  for (int dd = 0; dd < P.D; ++dd)
    for (int j = 0; j < N.dec[dd].n_layers; ++j) { N.dec[dd].L[j].bias_off = arena; arena += 4 * N.dec[dd].L[j].n_pad; }
  N.init_off = arena;
  arena += 4 * spad;
  N.arena_bytes = (arena + 15) & ~15;
  // register stash per step: the state (4 KSS registers) and up to two hidden outputs (4 KSH each)
  N.stash_step_regs = 4 * (inst->kss + 2 * inst->ksh);
  for (int e = 0; e < P.E; ++e)
    for (int j = 0; j < kMaxL; ++j) N.enc[e].stash_off[j] = 4 * (inst->kss + inst->ksh * std::min(j, 1));
  const int max_w = std::max(spad, hpad);
  N.stage_dz_off = max_w * 2;
  N.stage_pitch = 2 * max_w * 2 + 16;
  if (nb_smem_bytes(N) > (size_t)max_smem) { *why = reasons[4]; return false; }
  return true;
}

static const Inst& inst_of(const NbPlan& N) {
  for (const Inst& c : kInsts)
    if (N.kss <= c.kss && N.ksh <= c.ksh) return c;
  return kInsts[2];
}

int mmn_nb_plan_init(mmn_plan* p) {
  NbPlan* host = new NbPlan();
  const char* why = "";
  if (!mmn_nb_build(p->host, p->max_smem, *host, &why)) {
    delete host;
    return fail("the bf16 tile kernel does not take this model: %s", why);
  }
  p->nb_host = host;
  if (cudaMalloc(&p->nb_dev, sizeof(NbPlan)) != cudaSuccess ||
      cudaMemcpy(p->nb_dev, host, sizeof(NbPlan), cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMalloc(&p->nb_arena, (size_t)host->arena_bytes) != cudaSuccess)
    return fail("mmn_plan_create: device allocation failed (bf16 tile kernel)");
  return 0;
}

void mmn_nb_plan_free(mmn_plan* p) {
  if (p->nb_dev) cudaFree(p->nb_dev);
  if (p->nb_arena) cudaFree(p->nb_arena);
  delete static_cast<NbPlan*>(p->nb_host);
  p->nb_dev = p->nb_arena = p->nb_host = nullptr;
}

static int nb_grid(const mmn_plan* plan, int64_t n_rows) {
  const int64_t tiles = (n_rows + kTileRows - 1) / kTileRows;
  return (int)std::max<int64_t>(1, std::min<int64_t>((tiles + kNbGroups - 1) / kNbGroups, plan->n_sms));
}

int64_t mmn_nb_workspace_bytes(const mmn_plan* plan, int64_t n_rows, bool train) {
  if (!train) return 0;
  const NbPlan& N = *static_cast<const NbPlan*>(plan->nb_host);
  const int64_t grid = nb_grid(plan, n_rows);
  const int64_t reg_stash = grid * (kNbGroups * kWarpsPerGroup) * (int64_t)N.E * N.stash_step_regs * 32 * 4;
  const int64_t x_stash = grid * kNbGroups * (int64_t)N.xs_steps * kWarpsPerGroup * 32 * 16;
  return ((reg_stash + 255) & ~(int64_t)255) + x_stash;
}

int mmn_nb_step(const mmn_plan* plan, const StepArgs& a, void* ws, size_t ws_bytes, void* stream, bool train) {
  const NbPlan& N = *static_cast<const NbPlan*>(plan->nb_host);
  (void)ws_bytes;
  int n_layers = 0;
  for (int e = 0; e < N.E; ++e) n_layers += N.enc[e].n_layers;
  for (int d = 0; d < N.D; ++d) n_layers += N.dec[d].n_layers;
  MMN_LAUNCH(mmn_nb_prep_kernel<0>, dim3(std::min(plan->n_sms, 64)), dim3(256), 0, stream, static_cast<const NbPlan*>(plan->nb_dev), a.params,
             static_cast<unsigned char*>(plan->nb_arena));
  MMN_CUDA(cudaGetLastError());
  NbArgs args;
  args.a = a;
  args.nb_plan = static_cast<const NbPlan*>(plan->nb_dev);
  args.arena = static_cast<const unsigned char*>(plan->nb_arena);
  args.stash = static_cast<unsigned*>(ws);
  args.stash_words_per_warp = (long long)N.E * N.stash_step_regs * 32;
  const int grid = nb_grid(plan, a.n_rows);
  {
    const int64_t reg_stash = (int64_t)grid * (kNbGroups * kWarpsPerGroup) * (int64_t)N.E * N.stash_step_regs * 32 * 4;
    args.xstash = reinterpret_cast<float4*>(static_cast<char*>(ws) + ((reg_stash + 255) & ~(int64_t)255));
    args.xstash_vec_per_group = (long long)N.xs_steps * kWarpsPerGroup * 32;
  }
  const size_t smem = nb_smem_bytes(N);
  const Inst& c = inst_of(N);
  if (c.kss == 1) return launch_inst<1, 1>(plan, args, grid, smem, stream, train);
  if (c.ksh == 2) return launch_inst<4, 2>(plan, args, grid, smem, stream, train);
  return launch_inst<4, 4>(plan, args, grid, smem, stream, train);
}
