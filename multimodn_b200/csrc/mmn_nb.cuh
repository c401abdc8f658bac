// mmn_nb.cuh — the fused sequential-fusion step for NARROW models in bf16 (precision = "bf16", state <= 64, layers <= 64
// wide): one launch per train step, activations chained through REGISTERS.
//
// Why not tcgen05 here: a 32-wide layer is 128 x 32 x 32 MACs per tile — the tensor pipe is idle either way; what the round-1
// tcgen05 kernels (mmn_tc.cuh, mmn_tc2.cuh) paid for was the round trip per layer (commit -> mbarrier -> tcgen05.ld ->
// epilogue -> st.shared / tcgen05.st -> fence -> issue), ~2.5-4 us x ~100 dependent layers per tile.  With warp-level
// mma.sync.m16n8k16 the accumulator fragment of layer j IS the A fragment of layer j + 1 (same lane, same registers after a
// bf16 pack): a warp walks the whole encoder -> decoder chain of its 16 rows without a barrier, and the scheduler overlaps
// the warps.  (The wide regime, where layers are real dense contractions, stays on tcgen05: mmn_wide.cuh.)
//
// Layout of one CTA (512 threads, persistent, one per SM): 2 groups x 8 warps; a group owns a 128-row batch tile, a warp 16
// of its rows (one m16 tile).  Shared memory: every weight matrix as a bf16 row-major image [n][k] (pitch = odd multiple
// of 16 bytes: conflict-free ldmatrix; ONE image serves the forward B operand (ldmatrix) and the data-gradient B operand
// (ldmatrix.trans)), biases in fp32, and per group one staging buffer [128 rows x (a | dz)] for the weight gradients.
// The kernel is instantiated for a few (state width, hidden width) pairs in multiples of 16; a model runs on the smallest
// instantiation that holds it with zero-padded images, so every MMA loop has compile-time trip counts.
//
//   forward sweep   per step: x streamed from HBM straight into A fragments (LDG.128 ring -> NaN scan -> dropout -> cvt.bf16x2),
//                   encoder layers chained in registers, per-row missingness select, state-change sum; the state, the hidden
//                   activations (register images) and the converted x fragments are stashed (L2-resident) for the reverse sweep
//   reverse sweep   per step k = L..0: decoders forward + CE / arg-max / counters + backward on s_k (nothing of a decoder is
//                   ever stashed), G += dz . W accumulated by the MMA itself, then the encoder backward with the carry select
//   weight grads    dW = dz^T . a contracts over the 128 rows of the GROUP: each warp writes its rows of (a, dz) to the staging
//                   buffer, the 8 warps split the output tiles (ldmatrix.trans on both operands), bias gradients ride along
//                   as one extra n-tile against a synthesised ones-column, red.global.add.v2 per tile.  The x part of a first
//                   layer takes its B operand from the stashed x fragments through movmatrix — no conversion is repeated and
//                   x is read from HBM once.
//
// Rounding points = the oracle's bf16 restatement (oracle/multimodn_oracle.py, spec["precision"] = "bf16"): weights, layer
// inputs, activations, states and layer gradients bf16; accumulation, biases, G, parameter gradients fp32.
// Reference arithmetic restated: multimodn/multimodn.py:139-204, encoders/mlp_encoder.py:40-47,74-80, decoders/decoders.py:19-20,42-46.
#pragma once

#include "mmn_kernels.cuh"
#include "mmn_nb_prims.cuh"

namespace mmn {
namespace nb {

constexpr int kWarpRows = 16, kTileRows = 128;
constexpr int kNbGroups = 2, kWarpsPerGroup = kTileRows / kWarpRows, kGroupThreads = kWarpsPerGroup * 32;
constexpr int kThreadsNb = kNbGroups * kGroupThreads;
constexpr int kMaxL = 3;          // Linear layers per encoder / decoder in this engine
constexpr int kMaxW = 64;         // widest state / hidden layer
constexpr int kMaxC = 8;          // classes per decoder (one n8 tile)

struct NbLayer {
  int N;                          // true number of outputs (flush bounds, column masks)
  int ka;                         // true width of the non-state input (x or previous hidden)
  int ka_pad;                     // its padded width inside the image: round16(F) for x, 16 KSH / 16 KSS for chained inputs
  int n_pad;                      // rows of the image: 16 KSH (hidden output), 16 KSS (state output), 16 (decoder head)
  int has_state, x_type, act;
  int img_off, pitch;             // bytes: image [n_pad][ka_pad + has_state * 16 KSS] bf16, row pitch
  int bias_off;                   // byte offset of the fp32 bias [n_pad] in the arena
  int ktot;                       // row length of W / gW in floats
  long long w_off, b_off;         // offsets (floats) into the packed parameter / gradient buffers
};
struct NbEnc {
  int F, n_layers;
  float p_drop;
  int stash_off[kMaxL];           // register offset of layer j's OUTPUT inside a step's stash block (hidden layers)
  int xs_off;                     // first k16-step of this encoder's x fragments inside the group's x stash
  NbLayer L[kMaxL];
};
struct NbDec {
  int C, n_layers, out_off;
  NbLayer L[kMaxL];
};
struct NbPlan {
  int S, E, D, sumC, n_metrics;
  int kss, ksh;                   // instantiation: k16-steps of the (padded) state / hidden width
  int arena_bytes;                // weight images + biases + initial state
  int init_off;                   // byte offset of the fp32 initial state [16 kss] in the arena
  int stash_step_regs;            // registers per thread per step: state + hidden outputs
  int stage_pitch, stage_dz_off;  // staging buffer: row pitch (bytes), byte offset of the dz part inside a row
  int xs_steps;                   // k16-steps of x over all encoders (x stash size per group)
  long long init_param_off, n_params;
  NbEnc enc[MMN_MAX_ENCODERS];
  NbDec dec[MMN_MAX_DECODERS];
};

struct NbArgs {
  StepArgs a;                     // plan pointer unused (nb_plan below)
  const NbPlan* nb_plan;          // device
  const unsigned char* arena;     // device: images prepared by mmn_nb_prep_kernel for this step's parameters
  unsigned* stash;                // device: per-warp register stash
  long long stash_words_per_warp;
  float4* xstash;                 // device: per-group stash of the bf16 x fragments (16 bytes per thread per k16-step)
  long long xstash_vec_per_group;
};

// shared-memory footprint of the step kernel
inline size_t nb_smem_bytes(const NbPlan& P) {
  size_t b = (size_t)P.arena_bytes;
  b = (b + 15) & ~(size_t)15;
  b += (sizeof(NbPlan) + 15) & ~(size_t)15;
  b += (size_t)kNbGroups * kTileRows * P.stage_pitch;
  b += (size_t)kNbGroups * kTileRows * MMN_MAX_DECODERS;          // targets as bytes
  b += (size_t)kNbGroups * (MMN_MAX_ENCODERS + 1) * 4;            // tile_any
  b += (size_t)(P.E + 1) * 4;                                     // present-row counters
  b = (b + 7) & ~(size_t)7;
  b += (size_t)P.n_metrics * 8;
  return (b + 15) & ~(size_t)15;
}

// ------------------------------------------------------------------------------------------------
// weight images: fp32 parameters -> bf16 [n_pad][kpad] images (+ fp32 biases, initial state), once per step
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void nb_image_layer(const NbLayer& ly, int spad, int S, const float* __restrict__ params,
                                                unsigned char* __restrict__ arena, int tid, int nthreads) {
  const int kpad = ly.ka_pad + (ly.has_state ? spad : 0);
  unsigned short* img = reinterpret_cast<unsigned short*>(arena + ly.img_off);
  const int pitch_e = ly.pitch >> 1;
  for (int idx = tid; idx < ly.n_pad * kpad; idx += nthreads) {
    const int n = idx / kpad, c = idx - n * kpad;
    float v = 0.f;
    if (n < ly.N) {
      if (c < ly.ka_pad) {
        int actual = c;
        if (ly.x_type) {        // columns of an x-fed k16-step are permuted so that a thread's float4 (cols 4t..4t+3) lands on
          const int q = c & 15; // the fragment positions (2t, 2t+1, 2t+8, 2t+9): image col -> actual col
          actual = (c - q) + 4 * ((q & 7) >> 1) + 2 * (q >> 3) + (q & 1);
        }
        if (actual < ly.ka) v = __ldg(params + ly.w_off + (long long)n * ly.ktot + actual);
      } else {
        const int sc = c - ly.ka_pad;
        if (sc < S) v = __ldg(params + ly.w_off + (long long)n * ly.ktot + ly.ka + sc);
      }
    }
    img[n * pitch_e + c] = (unsigned short)bf16_bits(v);
  }
  float* bias = reinterpret_cast<float*>(arena + ly.bias_off);
  for (int n = tid; n < ly.n_pad; n += nthreads) bias[n] = n < ly.N ? __ldg(params + ly.b_off + n) : 0.f;
}

template <int = 0>
__global__ void __launch_bounds__(256) mmn_nb_prep_kernel(const NbPlan* plan, const float* __restrict__ params,
                                                         unsigned char* __restrict__ arena) {
  const NbPlan& P = *plan;
  // every block takes a grid-stride slice of every layer's image (the 1088-column first layer of an image encoder is
  // most of the work: one block per layer would serialise it)
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nthreads = gridDim.x * blockDim.x;
  for (int e = 0; e < P.E; ++e)
    for (int j = 0; j < P.enc[e].n_layers; ++j) nb_image_layer(P.enc[e].L[j], 16 * P.kss, P.S, params, arena, tid, nthreads);
  for (int d = 0; d < P.D; ++d)
    for (int j = 0; j < P.dec[d].n_layers; ++j) nb_image_layer(P.dec[d].L[j], 16 * P.kss, P.S, params, arena, tid, nthreads);
  float* init = reinterpret_cast<float*>(arena + P.init_off);
  for (int c = tid; c < 16 * P.kss; c += nthreads) init[c] = c < P.S ? __ldg(params + P.init_param_off + c) : 0.f;
}

// ------------------------------------------------------------------------------------------------
// small device helpers
// ------------------------------------------------------------------------------------------------
struct Lane {
  int lane, g, t, wg, gi;
};

// sigmoid / tanh through one exp2 + one reciprocal (tanh z = 2 sigmoid(2 z) - 1)
__device__ __forceinline__ float nb_squash(float z, float cin, float cm, float ca) {
  return fmaf(__fdividef(1.f, 1.f + __expf(-cin * z)), cm, ca);
}
__device__ __forceinline__ float nb_act(int act, float z) {
  if (act == MMN_ACT_RELU) return fmaxf(z, 0.f);
  if (act == MMN_ACT_SIGMOID) return nb_squash(z, 1.f, 1.f, 0.f);
  if (act == MMN_ACT_TANH) return nb_squash(z, 2.f, 2.f, -1.f);
  return z;
}
// derivative through the activation OUTPUT a, applied to two packed values
__device__ __forceinline__ void nb_dact2(int act, float a0, float a1, float& g0, float& g1) {
  if (act == MMN_ACT_RELU) { g0 = a0 > 0.f ? g0 : 0.f; g1 = a1 > 0.f ? g1 : 0.f; }
  else if (act == MMN_ACT_SIGMOID) { g0 *= a0 * (1.f - a0); g1 *= a1 * (1.f - a1); }
  else if (act == MMN_ACT_TANH) { g0 *= 1.f - a0 * a0; g1 *= 1.f - a1 * a1; }
}

// dropout keep bits of two adjacent concat columns (c, c + 1)
__device__ __forceinline__ void nb_keep2(const Drop& d, unsigned row, unsigned c, bool& k0, bool& k1) {
  if ((c & 1u) == 0) {
    const unsigned h = mmn_dropout_hash(d.seed_mix, row, c >> 1);
    k0 = (h & 0xffffu) >= d.thr;
    k1 = (h >> 16) >= d.thr;
  } else {
    k0 = (mmn_dropout_hash(d.seed_mix, row, c >> 1) >> 16) >= d.thr;
    k1 = (mmn_dropout_hash(d.seed_mix, row, (c >> 1) + 1) & 0xffffu) >= d.thr;
  }
}

// A fragments of the warp's [16 rows x 16 KS cols] activation: v[ks][i], i = 0: (row g, cols 2t, 2t+1), 1: (row g + 8, same
// cols), 2: (row g, cols 2t + 8, + 9), 3: (row g + 8, cols 2t + 8, + 9), all + 16 ks
template <int KS>
struct Frag {
  unsigned v[KS][4];
};
__device__ __forceinline__ int frag_row(const Lane& L, int i) { return L.g + 8 * (i & 1); }
__device__ __forceinline__ int frag_col(const Lane& L, int ks, int i) { return 16 * ks + 2 * L.t + 8 * (i >> 1); }

// ---- forward GEMM: acc[j] += A[ks] . W[8 j .., kcol0 + 16 ks ..]^T, all KS steps, all NT (even) n8 tiles of the image at `img` ----
template <int KS, int NT>
__device__ __forceinline__ void mma_fwd(float (&acc)[NT][4], const Frag<KS>& A, unsigned img, int pitch, int kcol0, const Lane& L) {
  static_assert(NT % 2 == 0, "tiles come in pairs (one ldmatrix.x4)");
  const unsigned lane_off = img + (unsigned)((8 * (L.lane >> 4) + (L.lane & 7)) * pitch + 16 * ((L.lane >> 3) & 1) + kcol0 * 2);
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
    for (int jp = 0; jp < NT / 2; ++jp) {
      unsigned b[4];
      ldsm_x4(b, lane_off + (unsigned)(16 * jp * pitch + 32 * ks));
      mma_bf16(acc[2 * jp], A.v[ks], b[0], b[1]);
      mma_bf16(acc[2 * jp + 1], A.v[ks], b[2], b[3]);
    }
  }
}

// ---- data gradient: acc[j] += DZ[ks] . W[16 ks .., col0 + 8 j ..]  (contraction over the layer's outputs, KS steps of 16) ----
template <int KS, int NT>
__device__ __forceinline__ void mma_dgrad(float (&acc)[NT][4], const Frag<KS>& DZ, unsigned img, int pitch, int col0, const Lane& L) {
  static_assert(NT % 2 == 0, "tiles come in pairs (one ldmatrix.x4.trans)");
  // ldmatrix.trans: matrix q = lane / 8: rows 16 ks + 8 (q & 1) + (lane & 7), cols col0 + 8 (j + (q >> 1))
  const unsigned lane_off = img + (unsigned)((8 * ((L.lane >> 3) & 1) + (L.lane & 7)) * pitch + 16 * (L.lane >> 4) + col0 * 2);
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
    for (int jp = 0; jp < NT / 2; ++jp) {
      unsigned b[4];
      ldsm_x4_t(b, lane_off + (unsigned)(16 * ks * pitch + 32 * jp));
      mma_bf16(acc[2 * jp], DZ.v[ks], b[0], b[1]);
      mma_bf16(acc[2 * jp + 1], DZ.v[ks], b[2], b[3]);
    }
  }
}

template <int NT>
__device__ __forceinline__ void acc_bias(float (&acc)[NT][4], const float* bias, const Lane& L) {
#pragma unroll
  for (int j = 0; j < NT; ++j) {
    const float2 bb = *reinterpret_cast<const float2*>(bias + 8 * j + 2 * L.t);
    acc[j][0] = bb.x; acc[j][1] = bb.y; acc[j][2] = bb.x; acc[j][3] = bb.y;
  }
}
template <int NT>
__device__ __forceinline__ void acc_zero(float (&acc)[NT][4]) {
#pragma unroll
  for (int j = 0; j < NT; ++j)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[j][c] = 0.f;
}

// activation over accumulator tiles with the kind test hoisted out of the element loops
template <int NT>
__device__ __forceinline__ void acc_act(float (&acc)[NT][4], int act) {
  if (act == MMN_ACT_RELU) {
#pragma unroll
    for (int j = 0; j < NT; ++j)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[j][c] = fmaxf(acc[j][c], 0.f);
  } else if (act == MMN_ACT_SIGMOID || act == MMN_ACT_TANH) {
    const float cin = act == MMN_ACT_TANH ? 2.f : 1.f, ca = act == MMN_ACT_TANH ? -1.f : 0.f;
#pragma unroll
    for (int j = 0; j < NT; ++j)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[j][c] = nb_squash(acc[j][c], cin, cin, ca);
  }
}

// activation + bf16 pack of NT accumulator tiles into tiles [J0, J0 + NT) of the next layer's A fragments (columns >= N -> 0)
template <int J0, int KS, int NT>
__device__ __forceinline__ void acc_to_frag(Frag<KS>& O, float (&acc)[NT][4], int act, int N, const Lane& L) {
  static_assert(J0 + NT <= 2 * KS, "tiles exceed the fragment width");
  acc_act<NT>(acc, act);
#pragma unroll
  for (int j = 0; j < NT; ++j) {
    const int jj = J0 + j;
    const int c0 = 8 * jj + 2 * L.t;
    float v0 = acc[j][0], v1 = acc[j][1], v2 = acc[j][2], v3 = acc[j][3];
    if (c0 >= N) { v0 = 0.f; v2 = 0.f; }
    if (c0 + 1 >= N) { v1 = 0.f; v3 = 0.f; }
    // tile jj -> k-step jj / 2, half jj % 2: registers (half * 2) [row g] and (half * 2 + 1) [row g + 8]
    O.v[jj >> 1][(jj & 1) * 2 + 0] = pack_bf16(v0, v1);
    O.v[jj >> 1][(jj & 1) * 2 + 1] = pack_bf16(v2, v3);
  }
}
// data gradient tiles [J0, J0 + NT) times act'(forward output held in `h`) -> bf16 fragments of dz
template <int J0, int KS, int NT>
__device__ __forceinline__ void dacc_to_frag(Frag<KS>& O, const float (&acc)[NT][4], const Frag<KS>& h, int act) {
  static_assert(J0 + NT <= 2 * KS, "tiles exceed the fragment width");
#pragma unroll
  for (int j = 0; j < NT; ++j) {
    const int jj = J0 + j;
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      const unsigned hv = h.v[jj >> 1][(jj & 1) * 2 + hh];
      float g0 = acc[j][2 * hh], g1 = acc[j][2 * hh + 1];
      nb_dact2(act, bf16_lo(hv), bf16_hi(hv), g0, g1);
      O.v[jj >> 1][(jj & 1) * 2 + hh] = pack_bf16(g0, g1);
    }
  }
}
template <int KS>
__device__ __forceinline__ void frag_zero(Frag<KS>& F) {
#pragma unroll
  for (int ks = 0; ks < KS; ++ks)
#pragma unroll
    for (int i = 0; i < 4; ++i) F.v[ks][i] = 0u;
}

// per-warp register stash in global memory: word (reg, lane) at base[reg * 32 + lane] (fully coalesced, layout-free)
template <int KS>
__device__ __forceinline__ void stash_put(unsigned* base, const Frag<KS>& F, int lane) {
#pragma unroll
  for (int ks = 0; ks < KS; ++ks)
#pragma unroll
    for (int i = 0; i < 4; ++i) __stcg(base + (ks * 4 + i) * 32 + lane, F.v[ks][i]);
}
template <int KS>
__device__ __forceinline__ void stash_get(const unsigned* base, Frag<KS>& F, int lane) {
#pragma unroll
  for (int ks = 0; ks < KS; ++ks)
#pragma unroll
    for (int i = 0; i < 4; ++i) F.v[ks][i] = __ldcg(base + (ks * 4 + i) * 32 + lane);
}

// rows of the warp's fragments written to the group's staging buffer as a row-major bf16 tile at byte column `col_off`
template <int KS>
__device__ __forceinline__ void stage_put(unsigned char* buf, int pitch, int col_off, const Frag<KS>& F, const Lane& L) {
  unsigned char* base = buf + (kWarpRows * L.wg + L.g) * pitch + col_off + 4 * L.t;
#pragma unroll
  for (int ks = 0; ks < KS; ++ks)
#pragma unroll
    for (int i = 0; i < 4; ++i)
      *reinterpret_cast<unsigned*>(base + 8 * (i & 1) * pitch + 32 * ks + 16 * (i >> 1)) = F.v[ks][i];
}

__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  return v;
}

// the initial state as A fragments (state.py:29-32: the same row for every sample), rounded to bf16 like every state
template <int KSS>
__device__ __forceinline__ void init_frag(Frag<KSS>& s, const float* init, const Lane& L) {
#pragma unroll
  for (int ks = 0; ks < KSS; ++ks) {
    const float2 a = *reinterpret_cast<const float2*>(init + 16 * ks + 2 * L.t);
    const float2 b = *reinterpret_cast<const float2*>(init + 16 * ks + 2 * L.t + 8);
    const unsigned lo = pack_bf16(a.x, a.y), hi = pack_bf16(b.x, b.y);
    s.v[ks][0] = lo; s.v[ks][1] = lo; s.v[ks][2] = hi; s.v[ks][3] = hi;
  }
}

// the state with the dropout mask of MIMIC_MLPEncoder applied: bf16(keep ? s / (1 - p) : 0) (mlp_encoder.py:41-44)
template <int KSS>
__device__ __forceinline__ void drop_state(Frag<KSS>& o, const Frag<KSS>& s, const Drop& drop, int F, const Lane& L) {
#pragma unroll
  for (int ks = 0; ks < KSS; ++ks)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const unsigned row = drop.row_base + (unsigned)(kWarpRows * L.wg + frag_row(L, i));
      bool k0, k1;
      nb_keep2(drop, row, (unsigned)(F + frag_col(L, ks, i)), k0, k1);
      const unsigned sv = s.v[ks][i];
      o.v[ks][i] = pack_bf16(k0 ? bf16_lo(sv) * drop.scale : 0.f, k1 ? bf16_hi(sv) * drop.scale : 0.f);
    }
}

// ------------------------------------------------------------------------------------------------
// weight-gradient job: gW[n][col_base + k] += sum over the group's 128 rows dz[r][n] a[r][k]   (n < N, k < K),
// gb[n] += sum_r dz[r][n].  a and dz are in the group's staging buffer (row-major bf16); work items (k-tile pair, m-tile)
// are dealt round-robin to the 8 warps; the bias column rides along with the first k-tile pair as one more MMA per k-step
// against a synthesised ones-column.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void wgrad_items(const unsigned char* buf, int pitch, int dz_off, int N, int K, float* __restrict__ gW,
                                            int ld, int col_base, float* __restrict__ gb, const Lane& L) {
  const unsigned sbuf = smem_addr(buf);
  const int MT = (N + 15) >> 4, KP = (K + 15) >> 4;
  const int n_items = MT * (KP > 0 ? KP : 1);               // K == 0: a bias-only job
  const bool v2 = ((ld & 1) == 0) && ((col_base & 1) == 0);
  // ldmatrix.trans lane offsets: A (dz^T): matrix q: rows r0 + 8 (q >> 1) + (lane & 7), cols n0 + 8 (q & 1)
  //                              B (a):    matrix q: rows r0 + 8 (q & 1) + (lane & 7), cols c0 + 8 (q >> 1)
  const int q = L.lane >> 3, lr = L.lane & 7;
  const unsigned a_lane = sbuf + (unsigned)((8 * (q >> 1) + lr) * pitch + dz_off + 16 * (q & 1));
  const unsigned b_lane = sbuf + (unsigned)((8 * (q & 1) + lr) * pitch + 16 * (q >> 1));
  const unsigned ones = L.g == 0 ? 0x3F803F80u : 0u;        // B = [1 0 0 ...]: column 0 of the extra tile sums the rows
  for (int item = L.wg; item < n_items; item += kWarpsPerGroup) {
    const int p = item / MT, m = item - p * MT;
    const bool with_bias = gb != nullptr && p == 0;         // the item of the first k-tile pair also sums dz over the rows
    float acc[2][4], accb[4];
    acc_zero<2>(acc);
    accb[0] = accb[1] = accb[2] = accb[3] = 0.f;
    const unsigned a_at = a_lane + 32 * m, b_at = b_lane + 32 * p;
    if (KP > 0) {
#pragma unroll
      for (int ks = 0; ks < kTileRows / 16; ++ks) {
        unsigned a[4], b[4];
        // the transposed loads deliver (q = 0: m 0-7 / k 0-7), (1: m 8-15 / k 0-7), (2: m 0-7 / k 8-15), (3: m 8-15 / k 8-15) = a0..a3
        ldsm_x4_t(a, a_at + (unsigned)(16 * ks * pitch));
        ldsm_x4_t(b, b_at + (unsigned)(16 * ks * pitch));
        mma_bf16(acc[0], a, b[0], b[1]);
        mma_bf16(acc[1], a, b[2], b[3]);
        if (with_bias) mma_bf16(accb, a, ones, ones);
      }
    } else {
#pragma unroll
      for (int ks = 0; ks < kTileRows / 16; ++ks) {
        unsigned a[4];
        ldsm_x4_t(a, a_at + (unsigned)(16 * ks * pitch));
        mma_bf16(accb, a, ones, ones);
      }
    }
    const int n0 = 16 * m + L.g;
    if (with_bias && L.t == 0) {
      if (n0 < N) red_add(gb + n0, accb[0]);
      if (n0 + 8 < N) red_add(gb + n0 + 8, accb[2]);
    }
    if (KP > 0) {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int k = 16 * p + 8 * j + 2 * L.t;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int n = n0 + 8 * h;
          if (n < N && k < K) {
            float* dst = gW + (long long)n * ld + col_base + k;
            if (v2 && k + 1 < K) red_add_v2(dst, acc[j][2 * h], acc[j][2 * h + 1]);
            else {
              red_add(dst, acc[j][2 * h]);
              if (k + 1 < K) red_add(dst + 1, acc[j][2 * h + 1]);
            }
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// x part of a first-layer weight gradient: gW[n][c] += sum_r dz0[r][n] x~[r][c], x~ = bf16(dropout(nan_to_0(x))).
// The forward sweep left x~ in the group's x stash exactly as its A fragments (16 bytes per thread per k16-step: rows g / g + 8
// of the warp's 16-row slab, columns 4t .. 4t + 3 of the step).  Each warp takes 16-column blocks (= forward k-steps) over ALL
// 128 rows of the group: one LDG.128 per forward warp, movmatrix.trans turns the packed column pairs into B fragments whose
// contraction index is the row.  No conversion work is repeated and x itself is read from HBM once.
// ------------------------------------------------------------------------------------------------
template <int MTX>
__device__ __forceinline__ void wgrad_x(const unsigned char* buf, int pitch, int dz_off, int N, const float4* __restrict__ xs,
                                        int F, float* __restrict__ gW, int ldw, const Lane& L) {
  const unsigned sbuf = smem_addr(buf);
  const int MT = (N + 15) >> 4;
  const bool v2 = (ldw & 1) == 0;
  const int q = L.lane >> 3, lr = L.lane & 7;
  const unsigned a_lane = sbuf + (unsigned)((8 * (q >> 1) + lr) * pitch + dz_off + 16 * (q & 1));
  const int n_blk = (F + 15) >> 4;
  // forward warp w holds rows 16 w + g (registers x, z) and 16 w + g + 8 (registers y, w) of a 16-column block; the 8 loads
  // of the warp's NEXT block are in flight while the current one is multiplied
  float4 v[kWarpsPerGroup], vn[kWarpsPerGroup];
  if (L.wg < n_blk) {
#pragma unroll
    for (int w = 0; w < kWarpsPerGroup; ++w) vn[w] = __ldcg(xs + ((long long)L.wg * kWarpsPerGroup + w) * 32 + L.lane);
  }
  for (int blk = L.wg; blk < n_blk; blk += kWarpsPerGroup) {
    const int c = 16 * blk + 4 * L.t;
    float acc[MTX][2][4];
#pragma unroll
    for (int m = 0; m < MTX; ++m) acc_zero<2>(acc[m]);
#pragma unroll
    for (int w = 0; w < kWarpsPerGroup; ++w) v[w] = vn[w];
    if (blk + kWarpsPerGroup < n_blk) {
#pragma unroll
      for (int w = 0; w < kWarpsPerGroup; ++w)
        vn[w] = __ldcg(xs + ((long long)(blk + kWarpsPerGroup) * kWarpsPerGroup + w) * 32 + L.lane);
    }
#pragma unroll
    for (int w = 0; w < kWarpsPerGroup; ++w) {
      // k16-step = the 16 rows of forward warp w: b0 from rows g (.x = cols 4t, 4t+1 -> tile 0; .z = cols 4t+2, +3 -> tile 1),
      // b1 from rows g + 8 (.y / .w)
      const unsigned b00 = movm_t(__float_as_uint(v[w].x)), b01 = movm_t(__float_as_uint(v[w].y));
      const unsigned b10 = movm_t(__float_as_uint(v[w].z)), b11 = movm_t(__float_as_uint(v[w].w));
#pragma unroll
      for (int m = 0; m < MTX; ++m) {
        if (m < MT) {
          unsigned a[4];
          ldsm_x4_t(a, a_lane + (unsigned)(16 * w * pitch + 32 * m));
          mma_bf16(acc[m][0], a, b00, b01);
          mma_bf16(acc[m][1], a, b10, b11);
        }
      }
    }
    // C fragment of tile j: rows n = 16 m + g (+ 8), "columns" 2t, 2t + 1 of the transposed block = actual x columns
    // 16 blk + 4 t + 2 j + {0, 1}
#pragma unroll
    for (int m = 0; m < MTX; ++m) {
      if (m < MT) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int k = c + 2 * j;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int n = 16 * m + L.g + 8 * h;
            if (n < N && k < F) {
              float* dst = gW + (long long)n * ldw + k;
              if (v2 && k + 1 < F) red_add_v2(dst, acc[m][j][2 * h], acc[m][j][2 * h + 1]);
              else {
                red_add(dst, acc[m][j][2 * h]);
                if (k + 1 < F) red_add(dst + 1, acc[m][j][2 * h + 1]);
              }
            }
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// x-fed first layer of an encoder, forward: streams the warp's 16 rows of x through a register ring into A fragments.
// acc (NT tiles, bias-initialised by the caller) += x~ . W0[:, x cols]^T;  nanbits: bit h set if row g + 8 h holds a NaN.
// ------------------------------------------------------------------------------------------------
template <int NT, bool TRAIN>
__device__ __forceinline__ void x_stream(float (&acc)[NT][4], const float* __restrict__ x, long long ld, int F, long long wrow0,
                                         long long n_rows, unsigned img, int pitch, const Drop& drop, float4* xs,
                                         unsigned& nanbits, const Lane& L) {
  const int ksx = (F + 15) >> 4;
  const bool vec = ((ld & 3) == 0) && ((F & 3) == 0) && ((reinterpret_cast<size_t>(x) & 15) == 0);
  long long r0 = wrow0 + L.g, r1 = wrow0 + L.g + 8;
  r0 = r0 < n_rows ? r0 : n_rows - 1;                       // rows past the batch: any valid address (they are masked)
  r1 = r1 < n_rows ? r1 : n_rows - 1;
  const float* p0 = x + r0 * ld + 4 * L.t;
  const float* p1 = x + r1 * ld + 4 * L.t;
  const int ct = 4 * L.t;
  auto load = [&](const float* p, int ks) -> float4 {
    const int c = 16 * ks + ct;
    const float* src = p + 16 * ks;
    if (vec) return c < F ? __ldg(reinterpret_cast<const float4*>(src)) : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 v;
    v.x = c + 0 < F ? __ldg(src + 0) : 0.f;
    v.y = c + 1 < F ? __ldg(src + 1) : 0.f;
    v.z = c + 2 < F ? __ldg(src + 2) : 0.f;
    v.w = c + 3 < F ? __ldg(src + 3) : 0.f;
    return v;
  };
  // NaN -> 0 (and the row is marked missing), dropout, bf16: one float4 -> two packed registers
  auto convert = [&](float4 w, int c, int h, unsigned& lo, unsigned& hi) {
    const bool bad = (w.x != w.x) | (w.y != w.y) | (w.z != w.z) | (w.w != w.w);
    if (bad) {
      nanbits |= 1u << h;
      w.x = w.x != w.x ? 0.f : w.x; w.y = w.y != w.y ? 0.f : w.y;
      w.z = w.z != w.z ? 0.f : w.z; w.w = w.w != w.w ? 0.f : w.w;
    }
    if (drop.enabled) {
      const unsigned row = drop.row_base + (unsigned)(kWarpRows * L.wg + L.g + 8 * h);
      bool k0, k1, k2, k3;
      nb_keep2(drop, row, (unsigned)c, k0, k1);
      nb_keep2(drop, row, (unsigned)c + 2, k2, k3);
      w.x = k0 ? w.x * drop.scale : 0.f; w.y = k1 ? w.y * drop.scale : 0.f;
      w.z = k2 ? w.z * drop.scale : 0.f; w.w = k3 ? w.w * drop.scale : 0.f;
    }
    lo = pack_bf16(w.x, w.y);
    hi = pack_bf16(w.z, w.w);
  };
  // ring of 4 k16-steps: the loads of step ks + 4 are issued as soon as step ks has been converted, so 8 LDG.128 per thread
  // stay in flight behind the conversion + MMAs (HBM latency)
  float4 a0 = load(p0, 0), b0 = load(p1, 0);
  float4 a1 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = a1, a2 = a1, b2 = a1, a3 = a1, b3 = a1;
  if (1 < ksx) { a1 = load(p0, 1); b1 = load(p1, 1); }
  if (2 < ksx) { a2 = load(p0, 2); b2 = load(p1, 2); }
  if (3 < ksx) { a3 = load(p0, 3); b3 = load(p1, 3); }
  const unsigned w_lane = img + (unsigned)((8 * (L.lane >> 4) + (L.lane & 7)) * pitch + 16 * ((L.lane >> 3) & 1));
#define MMN_NB_X_STEP(A_, B_, U_)                                                                         \
  {                                                                                                       \
    const int ks = ks0 + U_;                                                                              \
    if (ks < ksx) {                                                                                       \
      Frag<1> xa;                                                                                         \
      /* registers 0 / 1: rows g / g + 8, "k 2t, 2t+1"; 2 / 3: "k 2t+8, 2t+9" (image x columns permuted to match) */ \
      convert(A_, 16 * ks + ct, 0, xa.v[0][0], xa.v[0][2]);                                               \
      convert(B_, 16 * ks + ct, 1, xa.v[0][1], xa.v[0][3]);                                               \
      if (ks + 4 < ksx) { A_ = load(p0, ks + 4); B_ = load(p1, ks + 4); }                                 \
      if (TRAIN)                                                                                          \
        __stcg(xs + (long long)ks * kWarpsPerGroup * 32,                                                  \
               make_float4(__uint_as_float(xa.v[0][0]), __uint_as_float(xa.v[0][1]), __uint_as_float(xa.v[0][2]), \
                           __uint_as_float(xa.v[0][3])));                                                 \
      _Pragma("unroll") for (int jp = 0; jp < NT / 2; ++jp) {                                             \
        unsigned b[4];                                                                                    \
        ldsm_x4(b, w_lane + (unsigned)(16 * jp * pitch + 32 * ks));                                       \
        mma_bf16(acc[2 * jp], xa.v[0], b[0], b[1]);                                                       \
        mma_bf16(acc[2 * jp + 1], xa.v[0], b[2], b[3]);                                                   \
      }                                                                                                   \
    }                                                                                                     \
  }
  for (int ks0 = 0; ks0 < ksx; ks0 += 4) {
    MMN_NB_X_STEP(a0, b0, 0)
    MMN_NB_X_STEP(a1, b1, 1)
    MMN_NB_X_STEP(a2, b2, 2)
    MMN_NB_X_STEP(a3, b3, 3)
  }
#undef MMN_NB_X_STEP
}

// ------------------------------------------------------------------------------------------------
// the step kernel
// ------------------------------------------------------------------------------------------------
template <int KSS, int KSH, bool TRAIN>
__global__ void __launch_bounds__(kThreadsNb, 1) mmn_nb_step_kernel(const NbArgs args) {
  constexpr int NTS = 2 * KSS;            // n8 tiles of a state-wide output
  constexpr int NTH = 2 * KSH;            // n8 tiles of a hidden-wide output
  constexpr int PTS = NTS < 4 ? NTS : 4;  // tiles per accumulator pass (16 registers at most)
  constexpr int PTH = NTH < 4 ? NTH : 4;
  constexpr int J1S = NTS > PTS ? PTS : 0, J1H = NTH > PTH ? PTH : 0;     // first tile of a second pass (if there is one)
  const StepArgs& A = args.a;
  MMN_DYN_SMEM(smem_raw);

  Lane L;
  L.lane = threadIdx.x & 31; L.g = L.lane >> 2; L.t = L.lane & 3;
  const int warp = threadIdx.x >> 5;
  L.wg = warp % kWarpsPerGroup; L.gi = warp / kWarpsPerGroup;
  const int tid = threadIdx.x;

  // ---- carve shared memory ----
  const NbPlan& GP = *args.nb_plan;
  const int arena_bytes = (GP.arena_bytes + 15) & ~15;
  unsigned char* const smem_base = reinterpret_cast<unsigned char*>(smem_raw);
  unsigned char* arena = smem_base;
  NbPlan* Pp = reinterpret_cast<NbPlan*>(smem_base + arena_bytes);
  unsigned char* p = smem_base + arena_bytes + ((sizeof(NbPlan) + 15) & ~(size_t)15);
  const int stage_pitch = GP.stage_pitch, dz_off = GP.stage_dz_off;
  unsigned char* stage = p + (size_t)L.gi * kTileRows * stage_pitch;
  p += (size_t)kNbGroups * kTileRows * stage_pitch;
  unsigned char* ys = p + (size_t)L.gi * kTileRows * MMN_MAX_DECODERS;
  p += (size_t)kNbGroups * kTileRows * MMN_MAX_DECODERS;
  int* tile_any = reinterpret_cast<int*>(p) + L.gi * (MMN_MAX_ENCODERS + 1);
  p += (size_t)kNbGroups * (MMN_MAX_ENCODERS + 1) * 4;
  int* cnt = reinterpret_cast<int*>(p);
  p += (size_t)(GP.E + 1) * 4;
  p = smem_base + (((p - smem_base) + 7) & ~(size_t)7);
  double* met = reinterpret_cast<double*>(p);

  {   // weight images + plan into shared memory (once per CTA)
    const float4* src = reinterpret_cast<const float4*>(args.arena);
    float4* dst = reinterpret_cast<float4*>(arena);
    for (int i = tid; i < arena_bytes / 16; i += kThreadsNb) dst[i] = __ldg(src + i);
    const unsigned* ps = reinterpret_cast<const unsigned*>(args.nb_plan);
    unsigned* pd = reinterpret_cast<unsigned*>(Pp);
    for (int i = tid; i < (int)(sizeof(NbPlan) / 4); i += kThreadsNb) pd[i] = __ldg(ps + i);
    for (int i = tid; i < GP.n_metrics; i += kThreadsNb) met[i] = 0.0;
    for (int i = tid; i < GP.E + 1; i += kThreadsNb) cnt[i] = 0;
  }
  __syncthreads();
  const NbPlan& P = *Pp;
  const unsigned s_arena = smem_addr(arena);
  const int S = P.S, E = P.E, D = P.D, Ls = A.seq_len;
  const float* init = reinterpret_cast<const float*>(arena + P.init_off);

  const long long n_tiles = (A.n_rows + kTileRows - 1) / kTileRows;
  unsigned* my_stash = TRAIN ? args.stash + ((long long)blockIdx.x * (kNbGroups * kWarpsPerGroup) + warp) * args.stash_words_per_warp
                             : nullptr;
  float4* xstash_group = TRAIN ? args.xstash + ((long long)blockIdx.x * kNbGroups + L.gi) * args.xstash_vec_per_group : nullptr;
  Drop nodrop;
  nodrop.enabled = 0; nodrop.seed_mix = 0; nodrop.thr = 0; nodrop.row_base = 0; nodrop.scale = 1.f;

  for (long long tile = (long long)blockIdx.x * kNbGroups + L.gi; tile < n_tiles; tile += (long long)gridDim.x * kNbGroups) {
    const long long row0 = tile * kTileRows;                    // first row of the group's tile
    const long long wrow0 = row0 + kWarpRows * L.wg;            // first row of this warp
    group_bar(L.gi, kGroupThreads);                             // previous tile's readers of ys / tile_any / staging are done
    // ---- tile prologue: targets as bytes, flags ----
    if (A.targets) {
      for (int idx = L.wg * 32 + L.lane; idx < kTileRows * D; idx += kGroupThreads) {
        const int r = idx / D, d = idx - r * D;
        long long y = 0;
        if (row0 + r < A.n_rows) y = A.targets[(row0 + r) * D + d];
        const int C = P.dec[d].C;
        if ((y < 0 || y >= C) && A.target_error) *A.target_error = 1;   // CrossEntropyLoss would raise: reported, the caller raises
        y = y < 0 ? 0 : (y >= C ? C - 1 : y);                   // memory safety
        ys[r * MMN_MAX_DECODERS + d] = (unsigned char)y;
      }
    }
    if (L.wg == 0 && L.lane <= E) tile_any[L.lane] = L.lane == 0;
    // valid-row bits of this thread's 2 rows: bit h = row g + 8 h
    const unsigned valid = (wrow0 + L.g < A.n_rows ? 1u : 0u) | (wrow0 + L.g + 8 < A.n_rows ? 2u : 0u);
    unsigned long long pm = valid;                              // present bits: 2 per step, step k at bits [2k, 2k + 2)
    if (L.t == 0 && valid) atomicAdd(&cnt[0], __popc(valid));
    group_bar(L.gi, kGroupThreads);

    Frag<KSS> sA;
    init_frag<KSS>(sA, init, L);

    // =============================================================================================
    // decoders on the state in sA (multimodn.py:141-157, 176-191); TRAIN: followed by their backward pass, G += dLoss/ds
    // =============================================================================================
    auto decoders = [&](int hist_row, unsigned mask2, bool is_last_enc, float (&G)[NTS][4]) {
      for (int d = 0; d < D; ++d) {
        const NbDec& dec = P.dec[d];
        const int nl = dec.n_layers, C = dec.C;
        Frag<KSH> h1, h2;
        float head[2][4];
        // ---- forward chain ----
        if (nl == 1) {
          const NbLayer& l0 = dec.L[0];
          acc_bias<2>(head, reinterpret_cast<const float*>(arena + l0.bias_off), L);
          mma_fwd<KSS, 2>(head, sA, s_arena + l0.img_off, l0.pitch, 0, L);
        } else {
          {
            const NbLayer& l0 = dec.L[0];
#pragma unroll
            for (int pass = 0; pass < NTH / PTH; ++pass) {
              float acc[PTH][4];
              acc_bias<PTH>(acc, reinterpret_cast<const float*>(arena + l0.bias_off) + 8 * PTH * pass, L);
              mma_fwd<KSS, PTH>(acc, sA, s_arena + l0.img_off + 8 * PTH * pass * l0.pitch, l0.pitch, 0, L);
              if (pass == 0) acc_to_frag<0, KSH, PTH>(h1, acc, l0.act, l0.N, L);
              else acc_to_frag<J1H, KSH, PTH>(h1, acc, l0.act, l0.N, L);
            }
          }
          if (nl == 3) {
            const NbLayer& l1 = dec.L[1];
#pragma unroll
            for (int pass = 0; pass < NTH / PTH; ++pass) {
              float acc[PTH][4];
              acc_bias<PTH>(acc, reinterpret_cast<const float*>(arena + l1.bias_off) + 8 * PTH * pass, L);
              mma_fwd<KSH, PTH>(acc, h1, s_arena + l1.img_off + 8 * PTH * pass * l1.pitch, l1.pitch, 0, L);
              if (pass == 0) acc_to_frag<0, KSH, PTH>(h2, acc, l1.act, l1.N, L);
              else acc_to_frag<J1H, KSH, PTH>(h2, acc, l1.act, l1.N, L);
            }
          }
          const NbLayer& lh = dec.L[nl - 1];
          acc_bias<2>(head, reinterpret_cast<const float*>(arena + lh.bias_off), L);
          mma_fwd<KSH, 2>(head, nl == 3 ? h2 : h1, s_arena + lh.img_off, lh.pitch, 0, L);
        }
        // ---- per-row epilogue: first-max arg-max, CE on the squashed outputs, confusion cells, loss gradient ----
        const int hact = dec.L[nl - 1].act;
        Frag<1> dzh;
        frag_zero(dzh);
        float ce_sum = 0.f;
        unsigned pk1 = 0, pk2 = 0;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const bool rv = (valid >> h) & 1u, m = (mask2 >> h) & 1u;
          const int r = kWarpRows * L.wg + L.g + 8 * h;                  // row inside the group's tile
          const int c0 = 2 * L.t, c1 = 2 * L.t + 1;
          const float p0 = nb_act(hact, head[0][2 * h]), p1 = nb_act(hact, head[0][2 * h + 1]);
          const bool ok0 = c0 < C, ok1 = c1 < C;
          // arg-max with torch.max's first-maximum rule over the quad
          float bv = ok0 ? p0 : -3.4e38f;
          int bi = ok0 ? c0 : 1 << 20;
          if (ok1 && p1 > bv) { bv = p1; bi = c1; }
#pragma unroll
          for (int o = 1; o <= 2; o <<= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
          }
          const int pred = bi;
          if (!TRAIN && rv) {
            if (A.predictions && L.t == 0)
              A.predictions[((long long)hist_row * D + d) * A.pred_ld + row0 + r] = (unsigned char)pred;
            if (A.last_outputs && is_last_enc) {
              float* o = A.last_outputs + (row0 + r) * P.sumC + dec.out_off;
              if (ok0) o[c0] = p0;
              if (ok1) o[c1] = p1;
            }
          }
          if (A.targets) {
            const int y = ys[r * MMN_MAX_DECODERS + d];
            const float mx = bv;
            const float e0 = ok0 ? __expf(p0 - mx) : 0.f, e1 = ok1 ? __expf(p1 - mx) : 0.f;
            const float se = quad_sum(e0 + e1);
            const float py = quad_sum((c0 == y ? p0 : 0.f) + (c1 == y && ok1 ? p1 : 0.f));
            if (m && L.t == 0) {
              ce_sum += mx + __logf(se) - py;
              pk1 += (pred == y ? 1u : 0u);
              if (C == 2) {
                pk1 += (pred == 1 && y == 1 ? 1u << 8 : 0u) + (pred == 0 && y == 0 ? 1u << 16 : 0u) + (pred == 1 && y == 0 ? 1u << 24 : 0u);
                pk2 += (pred == 0 && y == 1 ? 1u : 0u);
              }
            }
            if (TRAIN) {
              // dLoss/dp = c_err m (softmax(p) - onehot(y)); through the output activation; rounded to bf16 (layer gradient)
              const float coef = m ? A.c_err : 0.f, inv = __fdividef(1.f, se);
              float d0 = ok0 ? coef * (e0 * inv - (c0 == y ? 1.f : 0.f)) : 0.f;
              float d1 = ok1 ? coef * (e1 * inv - (c1 == y ? 1.f : 0.f)) : 0.f;
              nb_dact2(hact, p0, p1, d0, d1);
              dzh.v[0][h] = pack_bf16(d0, d1);          // register h: row g + 8 h, k = classes 2t, 2t + 1
            }
          }
        }
        if (A.targets) {
          ce_sum = warp_sum(ce_sum);
          pk1 = warp_sum_u(pk1);
          pk2 = warp_sum_u(pk2);
          if (L.lane == 0) {
            const int base = hist_row * D + d, n = (E + 1) * D;
            atomicAdd(&met[0 * n + base], (double)ce_sum);
            atomicAdd(&met[1 * n + base], (double)(pk1 & 0xff));
            atomicAdd(&met[2 * n + base], (double)((pk1 >> 8) & 0xff));
            atomicAdd(&met[3 * n + base], (double)((pk1 >> 16) & 0xff));
            atomicAdd(&met[4 * n + base], (double)((pk1 >> 24) & 0xff));
            atomicAdd(&met[5 * n + base], (double)(pk2 & 0xff));
          }
        }
        if (!TRAIN) continue;

        // ---- backward chain of the decoder: stage (a, dz) -> data gradient -> weight gradients by the whole group ----
        float* grads = A.grads;
        const NbLayer& lh = dec.L[nl - 1];
        group_bar(L.gi, kGroupThreads);                       // staging buffer free
        if (nl == 1) stage_put<KSS>(stage, stage_pitch, 0, sA, L);
        else stage_put<KSH>(stage, stage_pitch, 0, nl == 3 ? h2 : h1, L);
        stage_put<1>(stage, stage_pitch, dz_off, dzh, L);
        Frag<KSH> dz;                                         // gradient at the previous hidden layer's pre-activation
        if (nl == 1) {
          mma_dgrad<1, NTS>(G, dzh, s_arena + lh.img_off, lh.pitch, 0, L);
        } else {
          const int pact = dec.L[nl - 2].act;
#pragma unroll
          for (int pass = 0; pass < NTH / PTH; ++pass) {
            float acc[PTH][4];
            acc_zero<PTH>(acc);
            mma_dgrad<1, PTH>(acc, dzh, s_arena + lh.img_off, lh.pitch, 8 * PTH * pass, L);
            if (pass == 0) dacc_to_frag<0, KSH, PTH>(dz, acc, nl == 3 ? h2 : h1, pact);
            else dacc_to_frag<J1H, KSH, PTH>(dz, acc, nl == 3 ? h2 : h1, pact);
          }
        }
        group_bar(L.gi, kGroupThreads);                       // staged rows of every warp are visible
        wgrad_items(stage, stage_pitch, dz_off, lh.N, lh.ka, grads + lh.w_off, lh.ktot, 0, grads + lh.b_off, L);
        if (nl == 3) {                                        // middle layer: dz is the gradient at layer 1's output
          const NbLayer& l1 = dec.L[1];
          group_bar(L.gi, kGroupThreads);
          stage_put<KSH>(stage, stage_pitch, 0, h1, L);
          stage_put<KSH>(stage, stage_pitch, dz_off, dz, L);
          Frag<KSH> dz1;
#pragma unroll
          for (int pass = 0; pass < NTH / PTH; ++pass) {
            float acc[PTH][4];
            acc_zero<PTH>(acc);
            mma_dgrad<KSH, PTH>(acc, dz, s_arena + l1.img_off, l1.pitch, 8 * PTH * pass, L);
            if (pass == 0) dacc_to_frag<0, KSH, PTH>(dz1, acc, h1, dec.L[0].act);
            else dacc_to_frag<J1H, KSH, PTH>(dz1, acc, h1, dec.L[0].act);
          }
          group_bar(L.gi, kGroupThreads);
          wgrad_items(stage, stage_pitch, dz_off, l1.N, l1.ka, grads + l1.w_off, l1.ktot, 0, grads + l1.b_off, L);
          dz = dz1;
        }
        if (nl > 1) {                                         // first layer: input = the state; its data gradient goes straight into G
          const NbLayer& l0 = dec.L[0];
          group_bar(L.gi, kGroupThreads);
          stage_put<KSS>(stage, stage_pitch, 0, sA, L);
          stage_put<KSH>(stage, stage_pitch, dz_off, dz, L);
          mma_dgrad<KSH, NTS>(G, dz, s_arena + l0.img_off, l0.pitch, 0, L);
          group_bar(L.gi, kGroupThreads);
          wgrad_items(stage, stage_pitch, dz_off, l0.N, l0.ka, grads + l0.w_off, l0.ktot, 0, grads + l0.b_off, L);
        }
      }
    };

    // ---- one chained encoder layer (input = the previous layer's fragments, + the state for MLPEncoder's last layer) ----
    auto chain_hidden = [&](const NbLayer& ly, const Frag<KSH>& in, Frag<KSH>& hout) {
#pragma unroll
      for (int pass = 0; pass < NTH / PTH; ++pass) {
        float acc[PTH][4];
        acc_bias<PTH>(acc, reinterpret_cast<const float*>(arena + ly.bias_off) + 8 * PTH * pass, L);
        mma_fwd<KSH, PTH>(acc, in, s_arena + ly.img_off + 8 * PTH * pass * ly.pitch, ly.pitch, 0, L);
        if (pass == 0) acc_to_frag<0, KSH, PTH>(hout, acc, ly.act, ly.N, L);
        else acc_to_frag<J1H, KSH, PTH>(hout, acc, ly.act, ly.N, L);
      }
    };
    auto chain_state = [&](const NbLayer& ly, const Frag<KSH>& in, Frag<KSS>& out) {
#pragma unroll
      for (int pass = 0; pass < NTS / PTS; ++pass) {
        float acc[PTS][4];
        acc_bias<PTS>(acc, reinterpret_cast<const float*>(arena + ly.bias_off) + 8 * PTS * pass, L);
        const unsigned img = s_arena + ly.img_off + 8 * PTS * pass * ly.pitch;
        mma_fwd<KSH, PTS>(acc, in, img, ly.pitch, 0, L);
        if (ly.has_state) mma_fwd<KSS, PTS>(acc, sA, img, ly.pitch, ly.ka_pad, L);
        if (pass == 0) acc_to_frag<0, KSS, PTS>(out, acc, ly.act, ly.N, L);
        else acc_to_frag<J1S, KSS, PTS>(out, acc, ly.act, ly.N, L);
      }
    };

    // =============================================================================================
    // one encoder, forward: returns the candidate state in `out` (bf16 fragments) and the NaN bits of the rows
    // =============================================================================================
    auto encoder_forward = [&](int e, int pos, const Drop& drop, Frag<KSS>& out, unsigned& nanbits, unsigned* stash_step) {
      const NbEnc& enc = P.enc[e];
      const int nl = enc.n_layers;
      Frag<KSH> h0, h1;
      nanbits = 0;
      {
        // x-fed first layer (the host guarantees 16 KSH >= its output width, the state included for 1-layer encoders)
        const NbLayer& ly = enc.L[0];
        float acc[NTH][4];
        acc_bias<NTH>(acc, reinterpret_cast<const float*>(arena + ly.bias_off), L);
        x_stream<NTH, TRAIN>(acc, A.x[pos], A.x_ld[pos], enc.F, wrow0, A.n_rows, s_arena + ly.img_off, ly.pitch, drop,
                             TRAIN ? xstash_group + ((long long)enc.xs_off * kWarpsPerGroup + L.wg) * 32 + L.lane : nullptr,
                             nanbits, L);
        if (ly.has_state) {
          if (drop.enabled) {
            Frag<KSS> sd;
            drop_state<KSS>(sd, sA, drop, enc.F, L);
            mma_fwd<KSS, NTH>(acc, sd, s_arena + ly.img_off, ly.pitch, ly.ka_pad, L);
          } else {
            mma_fwd<KSS, NTH>(acc, sA, s_arena + ly.img_off, ly.pitch, ly.ka_pad, L);
          }
        }
        acc_to_frag<0, KSH, NTH>(h0, acc, ly.act, ly.N, L);
        if (nl == 1) {
          // the output IS the state (width <= 16 KSH by the host's choice of instantiation)
          frag_zero(out);
#pragma unroll
          for (int ks = 0; ks < (KSH < KSS ? KSH : KSS); ++ks)
#pragma unroll
            for (int i = 0; i < 4; ++i) out.v[ks][i] = h0.v[ks][i];
        }
      }
      if (nl == 2) {
        if (TRAIN) stash_put<KSH>(stash_step + enc.stash_off[0] * 32, h0, L.lane);
        chain_state(enc.L[1], h0, out);
      } else if (nl == 3) {
        if (TRAIN) stash_put<KSH>(stash_step + enc.stash_off[0] * 32, h0, L.lane);
        chain_hidden(enc.L[1], h0, h1);
        if (TRAIN) stash_put<KSH>(stash_step + enc.stash_off[1] * 32, h1, L.lane);
        chain_state(enc.L[2], h1, out);
      }
      // rows share their NaN bits across the quad (each lane scanned 4 of every 16 columns)
      nanbits |= __shfl_xor_sync(0xffffffffu, nanbits, 1);
      nanbits |= __shfl_xor_sync(0xffffffffu, nanbits, 2);
    };

    // per-row select (missing rows keep their state bit for bit) + state-change sum (multimodn.py:173-174)
    auto select_state = [&](int e, const Frag<KSS>& cand, unsigned present2) {
      float sc = 0.f;
#pragma unroll
      for (int ks = 0; ks < KSS; ++ks)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if ((present2 >> (i & 1)) & 1u) {
            const unsigned o = sA.v[ks][i], n = cand.v[ks][i];
            const int c = frag_col(L, ks, i);
            const float d0 = bf16_lo(n) - bf16_lo(o), d1 = bf16_hi(n) - bf16_hi(o);
            if (c < S) sc = fmaf(d0, d0, sc);
            if (c + 1 < S) sc = fmaf(d1, d1, sc);
            sA.v[ks][i] = n;
          }
        }
      if (TRAIN) {
        sc = warp_sum(sc);
        if (L.lane == 0 && sc != 0.f) atomicAdd(&met[6 * (E + 1) * D + (E + 1) + e], (double)sc);
      }
    };

    auto make_drop = [&](int e) {
      Drop drop = nodrop;
      const NbEnc& enc = P.enc[e];
      if (TRAIN && A.training && enc.p_drop > 0.f && enc.L[0].has_state) {
        drop.enabled = 1;
        drop.seed_mix = A.dropout_seed ^ ((unsigned)e * 0x9E3779B9u);
        drop.thr = (unsigned)(enc.p_drop * 65536.f);
        drop.row_base = (unsigned)(A.row_offset + row0);
        drop.scale = 1.f / (1.f - enc.p_drop);
      }
      return drop;
    };

    float Gdummy[NTS][4];
    if (!TRAIN) decoders(0, valid, false, Gdummy);

    // ---- forward sweep over the encoding sequence (multimodn.py:159-191) ----
    for (int k = 1; k <= Ls; ++k) {
      const int e = A.seq_enc[k - 1], pos = A.seq_pos[k - 1];
      const bool skip = A.skip_flags && A.skip_flags[k - 1] != 0;      // reference batch-level rule
      unsigned present2 = 0;
      if (!skip) {
        Frag<KSS> cand;
        unsigned nanbits;
        const Drop drop = make_drop(e);
        encoder_forward(e, pos, drop, cand, nanbits, TRAIN ? my_stash + (long long)(k - 1) * P.stash_step_regs * 32 : nullptr);
        present2 = valid & ~nanbits;
        select_state(e, cand, present2);
      }
      pm |= (unsigned long long)present2 << (2 * k);
      if (L.t == 0 && present2) atomicAdd(&cnt[e + 1], __popc(present2));
      if (present2) tile_any[k] = 1;
      if (TRAIN) stash_put<KSS>(my_stash + ((long long)(k - 1) * P.stash_step_regs) * 32, sA, L.lane);
      else decoders(e + 1, present2, e == E - 1, Gdummy);
    }
    if (!TRAIN && A.final_state) {
#pragma unroll
      for (int ks = 0; ks < KSS; ++ks)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = frag_row(L, i), c = frag_col(L, ks, i);
          if (wrow0 + r < A.n_rows) {
            float* o = A.final_state + (wrow0 + r) * S;
            if (c < S) o[c] = bf16_lo(sA.v[ks][i]);
            if (c + 1 < S) o[c + 1] = bf16_hi(sA.v[ks][i]);
          }
        }
    }

    // =============================================================================================
    // reverse sweep: G = dLoss/ds_k for the warp's rows, fp32, accumulator layout
    // =============================================================================================
    if (TRAIN) {
      float G[NTS][4];
      acc_zero<NTS>(G);
      group_bar(L.gi, kGroupThreads);                        // tile_any and the x stash of every warp are visible
      float* grads = A.grads;
      for (int k = Ls; k >= 0; --k) {
        const unsigned mask2 = (unsigned)(pm >> (2 * k)) & 0x3u;
        if (k >= 1 && !tile_any[k]) continue;                 // no row of the tile took the step: s_k == s_{k-1}, nothing flows
        const int e = k >= 1 ? A.seq_enc[k - 1] : -1;
        decoders(e + 1, mask2, false, G);
        if (k == 0) break;
        const NbEnc& enc = P.enc[e];
        const int nl = enc.n_layers;
        const unsigned* stash_step = my_stash + (long long)(k - 1) * P.stash_step_regs * 32;
        // s_{k-1}: the previous step's stash, or the initial state
        Frag<KSS> sP;
        if (k >= 2) stash_get<KSS>(my_stash + ((long long)(k - 2) * P.stash_step_regs) * 32, sP, L.lane);
        else init_frag<KSS>(sP, init, L);
        // u_k = c_sc (s_k - s_{k-1}) joins G for the rows that took the step:  dz_last = present ? (G + u_k) act'(s_k) : 0
        const NbLayer& ll = enc.L[nl - 1];
        Frag<KSS> dzS;
#pragma unroll
        for (int ks = 0; ks < KSS; ++ks)
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            unsigned outv = 0;
            if ((mask2 >> (i & 1)) & 1u) {
              const unsigned a = sA.v[ks][i], b = sP.v[ks][i];
              const float a0 = bf16_lo(a), a1 = bf16_hi(a);
              const int jj = 2 * ks + (i >> 1), cc = (i & 1) * 2;       // G tile / element of (row g + 8 (i & 1); cols 2t, 2t+1 (+8))
              float z0 = G[jj][cc] + A.c_sc * (a0 - bf16_lo(b));
              float z1 = G[jj][cc + 1] + A.c_sc * (a1 - bf16_hi(b));
              nb_dact2(ll.act, a0, a1, z0, z1);
              outv = pack_bf16(z0, z1);
            }
            dzS.v[ks][i] = outv;
          }
        const Drop drop = make_drop(e);
        // The carry: rows that took the step replace G by the data gradient w.r.t. the state columns (through the dropout
        // mask) minus u_k (which belongs to s_{k-1} with the opposite sign: multimodn.py:165,174); absent rows keep G (their
        // u_k is zero: s_k == s_{k-1} bit for bit).
        auto carry = [&](auto& dzf, const NbLayer& ly, int Fcat) {
#pragma unroll
          for (int pass = 0; pass < NTS / PTS; ++pass) {
            float acc[PTS][4];
            acc_zero<PTS>(acc);
            mma_dgrad<sizeof(dzf.v) / 16, PTS>(acc, dzf, s_arena + ly.img_off, ly.pitch, ly.ka_pad + 8 * PTS * pass, L);
#pragma unroll
            for (int j = 0; j < PTS; ++j) {
              const int jj = PTS * pass + j;
#pragma unroll
              for (int hh = 0; hh < 2; ++hh) {
                if ((mask2 >> hh) & 1u) {
                  float c0 = acc[j][2 * hh], c1 = acc[j][2 * hh + 1];
                  if (drop.enabled) {
                    const unsigned row = drop.row_base + (unsigned)(kWarpRows * L.wg + L.g + 8 * hh);
                    bool k0, k1;
                    nb_keep2(drop, row, (unsigned)(Fcat + 8 * jj + 2 * L.t), k0, k1);
                    c0 = k0 ? c0 * drop.scale : 0.f;
                    c1 = k1 ? c1 * drop.scale : 0.f;
                  }
                  const unsigned a = sA.v[jj >> 1][(jj & 1) * 2 + hh], b = sP.v[jj >> 1][(jj & 1) * 2 + hh];
                  G[jj][2 * hh] = c0 - A.c_sc * (bf16_lo(a) - bf16_lo(b));
                  G[jj][2 * hh + 1] = c1 - A.c_sc * (bf16_hi(a) - bf16_hi(b));
                }
              }
            }
          }
        };
        // state columns of a layer that takes the state: second job on the already staged dz (a = s_{k-1} through the mask)
        auto state_job = [&](auto& dzf, const NbLayer& ly, bool with_bias, int Fcat) {
          group_bar(L.gi, kGroupThreads);                    // readers of the a part are done (the dz part stays)
          if (drop.enabled) {
            Frag<KSS> sd;
            drop_state<KSS>(sd, sP, drop, Fcat, L);
            stage_put<KSS>(stage, stage_pitch, 0, sd, L);
          } else {
            stage_put<KSS>(stage, stage_pitch, 0, sP, L);
          }
          carry(dzf, ly, Fcat);
          group_bar(L.gi, kGroupThreads);
          wgrad_items(stage, stage_pitch, dz_off, ly.N, S, grads + ly.w_off, ly.ktot, ly.ka, with_bias ? grads + ly.b_off : nullptr, L);
        };
        const float4* xs = xstash_group + (long long)enc.xs_off * kWarpsPerGroup * 32;
        if (nl == 1) {
          // x-fed layer that produces the state: dz = dzS (only its first KSH k-steps can be non-zero)
          const NbLayer& ly = enc.L[0];
          group_bar(L.gi, kGroupThreads);
          stage_put<KSS>(stage, stage_pitch, dz_off, dzS, L);
          group_bar(L.gi, kGroupThreads);
          if (ly.has_state) state_job(dzS, ly, true, enc.F);
          else wgrad_items(stage, stage_pitch, dz_off, ly.N, 0, grads + ly.w_off, ly.ktot, 0, grads + ly.b_off, L);
          sA = sP;
          wgrad_x<KSH>(stage, stage_pitch, dz_off, ly.N, xs, enc.F, grads + ly.w_off, ly.ktot, L);
        } else {
          // last layer: input = the stashed output of layer nl - 2 (+ the state for MLPEncoder)
          Frag<KSH> hin, dzH;
          {
            const NbLayer& ly = enc.L[nl - 1];
            stash_get<KSH>(stash_step + enc.stash_off[nl - 2] * 32, hin, L.lane);
            group_bar(L.gi, kGroupThreads);
            stage_put<KSS>(stage, stage_pitch, dz_off, dzS, L);
            stage_put<KSH>(stage, stage_pitch, 0, hin, L);
#pragma unroll
            for (int pass = 0; pass < NTH / PTH; ++pass) {
              float acc[PTH][4];
              acc_zero<PTH>(acc);
              mma_dgrad<KSS, PTH>(acc, dzS, s_arena + ly.img_off, ly.pitch, 8 * PTH * pass, L);
              if (pass == 0) dacc_to_frag<0, KSH, PTH>(dzH, acc, hin, enc.L[nl - 2].act);
              else dacc_to_frag<J1H, KSH, PTH>(dzH, acc, hin, enc.L[nl - 2].act);
            }
            group_bar(L.gi, kGroupThreads);
            wgrad_items(stage, stage_pitch, dz_off, ly.N, ly.ka, grads + ly.w_off, ly.ktot, 0, grads + ly.b_off, L);
            if (ly.has_state) state_job(dzS, ly, false, 0);
          }
          if (nl == 3) {
            const NbLayer& ly = enc.L[1];
            Frag<KSH> hin0, dz0;
            stash_get<KSH>(stash_step + enc.stash_off[0] * 32, hin0, L.lane);
            group_bar(L.gi, kGroupThreads);
            stage_put<KSH>(stage, stage_pitch, dz_off, dzH, L);
            stage_put<KSH>(stage, stage_pitch, 0, hin0, L);
#pragma unroll
            for (int pass = 0; pass < NTH / PTH; ++pass) {
              float acc[PTH][4];
              acc_zero<PTH>(acc);
              mma_dgrad<KSH, PTH>(acc, dzH, s_arena + ly.img_off, ly.pitch, 8 * PTH * pass, L);
              if (pass == 0) dacc_to_frag<0, KSH, PTH>(dz0, acc, hin0, enc.L[0].act);
              else dacc_to_frag<J1H, KSH, PTH>(dz0, acc, hin0, enc.L[0].act);
            }
            group_bar(L.gi, kGroupThreads);
            wgrad_items(stage, stage_pitch, dz_off, ly.N, ly.ka, grads + ly.w_off, ly.ktot, 0, grads + ly.b_off, L);
            dzH = dz0;
          }
          // first layer: x columns from the x stash, state columns (MIMIC_MLPEncoder) as a second job
          const NbLayer& ly = enc.L[0];
          group_bar(L.gi, kGroupThreads);
          stage_put<KSH>(stage, stage_pitch, dz_off, dzH, L);
          group_bar(L.gi, kGroupThreads);
          // (the state job first: it finalises G and ends the live ranges of s_k and dz, which leaves the x columns the
          //  registers to keep the next block's loads in flight)
          if (ly.has_state) state_job(dzH, ly, true, enc.F);
          else wgrad_items(stage, stage_pitch, dz_off, ly.N, 0, grads + ly.w_off, ly.ktot, 0, grads + ly.b_off, L);
          sA = sP;                                           // the state the next (earlier) step's decoders and u_{k-1} see
          wgrad_x<KSH>(stage, stage_pitch, dz_off, ly.N, xs, enc.F, grads + ly.w_off, ly.ktot, L);
        }
      }
      // ---- gradient of the initial state: column sums of G over the valid rows (tile backward of state.py:30) ----
#pragma unroll
      for (int jj = 0; jj < NTS; ++jj) {
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh)
          if ((valid >> hh) & 1u) { s0 += G[jj][2 * hh]; s1 += G[jj][2 * hh + 1]; }
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) {
          s0 += __shfl_xor_sync(0xffffffffu, s0, o);
          s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        }
        const int c = 8 * jj + 2 * L.t;
        if (L.g == 0) {
          if (c < S) red_add(grads + P.init_param_off + c, s0);
          if (c + 1 < S) red_add(grads + P.init_param_off + c + 1, s1);
        }
      }
    }
  }

  // ---- flush this CTA's metric partials ----
  __syncthreads();
  if (A.metrics) {
    const int nmat = 6 * (E + 1) * D;
    for (int i = tid; i < P.n_metrics; i += kThreadsNb) {
      double v;
      if (i < (E + 1) * D) v = met[i] * A.inv_rows_global;
      else if (i < nmat) v = met[i];
      else if (i < nmat + E + 1) v = (double)cnt[i - nmat];
      else v = met[i] * A.inv_rows_global / (double)S;
      if (v != 0.0) atomicAdd(A.metrics + i, v);
    }
  }
  if (TRAIN && A.grads) {
    for (int e = tid; e < E; e += kThreadsNb)
      if (cnt[e + 1]) atomicAdd(A.grads + P.n_params + e, (float)cnt[e + 1]);
  }
}

}  // namespace nb
}  // namespace mmn
