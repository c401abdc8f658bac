// mmn_tc2.cuh — the TMEM-resident tensor-core step kernel ("v2") for narrow models (state <= 64, every
// layer <= 64 wide, <= 16 classes): the batch tile's activations never visit shared memory on the forward path.
//
//   * A thread owns one row of the 128-row tile (its TMEM lane) and 16-column slices of every activation:
//     warp w serves lane quarter w & 3 and slice w >> 2.
//   * Every activation that feeds a GEMM lives in tensor memory as a (hi, lo) pair of tf32-exact column blocks;
//     tcgen05.mma reads it in place (TS form: A from TMEM), the weights come from shared memory as SWIZZLE_128B
//     (K-major) / SWIZZLE_128B_BASE32B (MN-major) hi/lo images; the accumulator is read back with tcgen05.ld,
//     bias / activation / missingness select / loss run in registers and the result is written straight back to
//     TMEM with tcgen05.st as the next layer's operand (and to the global stash for the backward pass).
//   * x is streamed from HBM row-wise: each thread fetches 16 contiguous floats of its own row per 32-column
//     chunk (full 64-byte segments), scans them for NaN, applies dropout, splits hi/lo and stores them into one of
//     two TMEM chunk slots — no shared-memory staging, no swizzle arithmetic.
//   * Weight gradients contract over the rows, which tensor memory cannot express as K: dz and the layer inputs
//     are written to shared memory as MN-major images by the epilogues / loaders and multiplied as in mmn_tc.cuh.
//   * A 9th warp issues every MMA; workers hand over opcodes through two (full, done) mbarrier pairs.
//
// TMEM column map (512 columns x 128 lanes):
//   forward : ACC 0..63 | S hi 64..127, lo 128..191 | P hi 192..255, lo 256..319 | Q hi 320..383, lo 384..447 |
//             X hi 448..479, lo 480..511   (x chunk slot 0; slot 1 borrows Q, which is idle while x streams)
//   backward: ACC 0..63 | G 64..127 (raw fp32 dLoss/ds) | T hi 128..191, lo 448..511 | P, Q as above
#pragma once

#include "mmn_tc.cuh"

namespace mmn {

constexpr unsigned V2_ACC = 0, V2_S_HI = 64, V2_S_LO = 128, V2_P_HI = 192, V2_P_LO = 256, V2_Q_HI = 320, V2_Q_LO = 384,
                   V2_X_HI = 448, V2_X_LO = 480, V2_G = 64, V2_T_HI = 128, V2_T_LO = 448;
enum { V2_A_X = 0, V2_A_S = 1, V2_A_P = 2, V2_A_Q = 3, V2_A_T = 4 };
enum { V2_OP_TS_K = 0, V2_OP_TS_MN = 1, V2_OP_TN = 2, V2_OP_QUIT = 3 };
constexpr int kV2TmemCols = 512;

// shared memory: DZ = four 16 KB images {g0 hi, g0 lo, g1 hi, g1 lo} (MN-major dz column groups, weight gradient),
// IN0 / IN1 = {hi, lo} input-chunk image pairs (weight gradient), WB = two {hi, lo} weight-block slots (8 KB images)
constexpr int kV2Dz = 4 * 4096, kV2In = 2 * 4096, kV2Wb = 2 * 2 * 2048;

struct V2Engine {
  static constexpr int TM = 128;
  static constexpr int kWorkers = 256;
  static constexpr int kBlockThreads = kWorkers + 32;
  using State = TcState;

  static size_t smem_bytes(const DevPlan& p) {
    size_t bytes = 1024 + (size_t)(kV2Dz + 2 * kV2In + kV2Wb + 256) * 4 + 64 + 2 * sizeof(TcCmd) + 16;
    bytes += (size_t)TM * p.D * 4 + TM * 4 + (size_t)(p.E + 1) * 8;
    bytes = (bytes + 7) & ~(size_t)7;
    bytes += (size_t)p.n_metrics * 8 + (size_t)(p.E + 1) * TM;
    return (bytes + 15) & ~(size_t)15;
  }
  static bool supports(const DevPlan& p) {
    if (p.S > 64 || p.ldH > 64 + 4) return false;
    for (int d = 0; d < p.D; ++d)
      if (p.dec[d].C > 16) return false;
    return true;
  }

  struct Sm {
    float *DZ, *IN0, *IN1, *WB, *RED;
    unsigned long long* bar;      // full[2], done[2]
    TcCmd* cmd;                   // [2]
    unsigned* tslot;
    int *ys, *rownan, *cnt, *tile_any;
    double* met;
    unsigned char* present;
  };
  __device__ static __forceinline__ void carve(Sm& sm, char* p, const DevPlan& P) {
    p += (1024 - (smem_u32(p) & 1023)) & 1023;
    float* f = reinterpret_cast<float*>(p);
    sm.DZ = f; f += kV2Dz;
    sm.IN0 = f; f += kV2In;
    sm.IN1 = f; f += kV2In;
    sm.WB = f; f += kV2Wb;
    sm.RED = f; f += 256;
    sm.bar = reinterpret_cast<unsigned long long*>(f);
    sm.cmd = reinterpret_cast<TcCmd*>(sm.bar + 4);
    sm.tslot = reinterpret_cast<unsigned*>(sm.cmd + 2);
    f = reinterpret_cast<float*>(sm.tslot + 4);
    sm.ys = reinterpret_cast<int*>(f); f += TM * P.D;
    sm.rownan = reinterpret_cast<int*>(f); f += TM;
    sm.cnt = reinterpret_cast<int*>(f); f += (P.E + 1);
    sm.tile_any = reinterpret_cast<int*>(f); f += (P.E + 1);
    size_t off = (reinterpret_cast<size_t>(f) + 7) & ~(size_t)7;
    sm.met = reinterpret_cast<double*>(off);
    sm.present = reinterpret_cast<unsigned char*>(sm.met + P.n_metrics);
  }

  // ---- issuer ------------------------------------------------------------------------------------------------
  template <bool BMN, unsigned AHI, unsigned ALO>
  __device__ static __forceinline__ void issue_ts(unsigned b_hi, unsigned b_lo, unsigned idesc, unsigned nj, bool first,
                                                  bool leader) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (j < (int)nj && leader) {
        const unsigned long long bh = BMN ? umma_desc_mn(b_hi, j, 4096) : umma_desc_k(b_hi, j);
        const unsigned long long bl = BMN ? umma_desc_mn(b_lo, j, 4096) : umma_desc_k(b_lo, j);
        umma_tf32_ts(V2_ACC, ALO + 8 * j, bh, idesc, (first && j == 0) ? 0u : 1u);
        umma_tf32_ts(V2_ACC, AHI + 8 * j, bl, idesc, 1u);
        umma_tf32_ts(V2_ACC, AHI + 8 * j, bh, idesc, 1u);
      }
    }
  }
  template <bool BMN>
  __device__ static __forceinline__ void issue_ts_sel(unsigned asel, bool achunk, unsigned b_hi, unsigned b_lo, bool n64,
                                                      unsigned nj, bool first, bool leader) {
    const unsigned id32 = umma_idesc_tf32(128, 32, 0, BMN ? 1 : 0), id64 = umma_idesc_tf32(128, 64, 0, BMN ? 1 : 0);
#define MMN_V2_CASE(SEL, HI, LO)                                                                        \
  if (asel == SEL) {                                                                                    \
    if (!achunk) { if (n64) issue_ts<BMN, HI, LO>(b_hi, b_lo, id64, nj, first, leader);                 \
                   else issue_ts<BMN, HI, LO>(b_hi, b_lo, id32, nj, first, leader); }                   \
    else { if (n64) issue_ts<BMN, HI + 32, LO + 32>(b_hi, b_lo, id64, nj, first, leader);               \
           else issue_ts<BMN, HI + 32, LO + 32>(b_hi, b_lo, id32, nj, first, leader); }                 \
  }
    MMN_V2_CASE(V2_A_X, V2_X_HI, V2_X_LO)
    MMN_V2_CASE(V2_A_S, V2_S_HI, V2_S_LO)
    MMN_V2_CASE(V2_A_P, V2_P_HI, V2_P_LO)
    MMN_V2_CASE(V2_A_Q, V2_Q_HI, V2_Q_LO)
    MMN_V2_CASE(V2_A_T, V2_T_HI, V2_T_LO)
#undef MMN_V2_CASE
  }
  __device__ static __forceinline__ void issuer_loop(const Sm& sm, unsigned tmem_base) {
    const unsigned dz = smem_u32(sm.DZ), in0 = smem_u32(sm.IN0), in1 = smem_u32(sm.IN1), wb = smem_u32(sm.WB);
    if (tmem_base != 0) __trap();             // the column constants above assume the CTA owns all of TMEM
    const bool leader = elect_one() != 0;
    unsigned par[2] = {0u, 0u};
    for (unsigned seq = 0;; ++seq) {
      const unsigned slot = seq & 1u;
      mbar_wait(sm.bar + slot, par[slot]);
      par[slot] ^= 1u;
      tc_fence_after();
      const unsigned op = sm.cmd[slot].op;    // [0,2) kind | [2] N=64 | [3] first | [4,9) nj | [9,12) asel | [12] achunk | [13] dz group
      const unsigned kind = op & 3u, nj = (op >> 4) & 31u, asel = (op >> 9) & 7u;
      const bool n64 = (op >> 2) & 1u, first = (op >> 3) & 1u, achunk = (op >> 12) & 1u, grp = (op >> 13) & 1u;
      if (kind == V2_OP_QUIT) break;
      const unsigned b_hi = wb + slot * 16384u;
      if (kind == V2_OP_TS_K) {
        issue_ts_sel<false>(asel, achunk, b_hi, b_hi + 8192u, n64, nj, first, leader);
      } else if (kind == V2_OP_TS_MN) {
        issue_ts_sel<true>(asel, achunk, b_hi, b_hi + 8192u, n64, nj, first, leader);
      } else {                                // weight gradient: A = dz group image, B = input chunk image
        const unsigned a_hi = dz + (grp ? 32768u : 0u), i_hi = slot ? in1 : in0;
        TcEngine::issue<true, true, 16>(a_hi, a_hi + 16384u, 0, i_hi, i_hi + 16384u, 0, umma_idesc_tf32(128, 32, 1, 1),
                                        V2_ACC + 32u * slot, 16, true, leader);
      }
      if (leader) umma_commit(sm.bar + 2 + slot);
      __syncwarp();
    }
  }

  // ---- worker side of the protocol -------------------------------------------------------------------------------
  __device__ static __forceinline__ void wait(const Sm& sm, State& es, int slot) {
    if (es.pending[slot]) {
      mbar_wait(sm.bar + 2 + slot, es.parity[slot]);
      es.parity[slot] ^= 1;
      es.pending[slot] = 0;
    }
  }
  __device__ static __forceinline__ void drain(const Sm& sm, State& es) {
    wait(sm, es, es.seq & 1);
    wait(sm, es, (es.seq & 1) ^ 1);
  }
  __device__ static __forceinline__ void post(const Sm& sm, State& es, int slot, unsigned op) {
    if (threadIdx.x == 0) sm.cmd[slot].op = op;
    fence_proxy_async();
    tc_fence_before();
    mbar_arrive(sm.bar + slot);
    es.pending[slot] = 1;
    es.seq += 1;
  }
  __device__ static __forceinline__ unsigned op_ts(bool bmn, int N, unsigned first, int nj, unsigned asel, int achunk) {
    return (bmn ? V2_OP_TS_MN : V2_OP_TS_K) | (N > 32 ? 4u : 0u) | (first ? 8u : 0u) | ((unsigned)nj << 4) | (asel << 9) |
           ((unsigned)achunk << 12);
  }
};

// per-thread view of the tile
struct V2Thread {
  int tid, lane, q, cs, r;
  unsigned lane_addr;       // TMEM lane field of this warp's quarter
};

__device__ __forceinline__ void v2_ld(const V2Thread& t, unsigned col, float (&v)[16]) {
  tmem_ld16(t.lane_addr + col + 16u * t.cs, v);
}
__device__ __forceinline__ void v2_st(const V2Thread& t, unsigned col, const float (&v)[16]) {
  tmem_st16(t.lane_addr + col + 16u * t.cs, v);
}
// split into the tf32-exact part and the remainder and store both operand blocks
__device__ __forceinline__ void v2_st_pair(const V2Thread& t, unsigned hi_col, unsigned lo_col, const float (&v)[16]) {
  float h[16], l[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) { h[i] = tf32_hi(v[i]); l[i] = v[i] - h[i]; }
  v2_st(t, hi_col, h);
  v2_st(t, lo_col, l);
}
// 16 consecutive floats of a row-major global block: row r, columns c0 .. c0+15, zero beyond `width`
template <bool COHERENT>
__device__ __forceinline__ void v2_load_row16(const float* base, long long ld, int r, int c0, int width, bool vec,
                                              float (&v)[16]) {
  const float* p = base + (long long)r * ld + c0;
  if (vec && c0 + 16 <= width) {
#pragma unroll
    for (int i = 0; i < 16; i += 4) {
      const float4 x = COHERENT ? __ldcg(reinterpret_cast<const float4*>(p + i)) : __ldg(reinterpret_cast<const float4*>(p + i));
      v[i] = x.x; v[i + 1] = x.y; v[i + 2] = x.z; v[i + 3] = x.w;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = (c0 + i < width) ? (COHERENT ? __ldcg(p + i) : __ldg(p + i)) : 0.f;
  }
}
__device__ __forceinline__ void v2_store_row16(float* base, int ld, int r, int c0, int width, const float (&v)[16]) {
  float* p = base + (long long)r * ld + c0;
  if (((ld & 3) == 0) && c0 + 16 <= width) {
#pragma unroll
    for (int i = 0; i < 16; i += 4) __stcg(reinterpret_cast<float4*>(p + i), make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]));
  } else {
#pragma unroll
    for (int i = 0; i < 16; ++i)
      if (c0 + i < width) __stcg(p + i, v[i]);
  }
}
__device__ __forceinline__ bool v2_vec_ok(const float* base, long long ld) {
  return ((ld & 3) == 0) && ((reinterpret_cast<size_t>(base) & 15) == 0);
}

}  // namespace mmn

namespace mmn {

__device__ __forceinline__ int v2_mn_off(int r, int c) {      // float offset of (row r, col c < 32) in an MN-major image
  const int c4 = c >> 2;
  return ((r >> 2) << 7) + ((r & 3) << 5) + ((((c4 >> 1) ^ (r & 3)) << 3) | ((c4 & 1) << 2)) + (c & 3);
}

struct V2Res {       // a GEMM operand segment that already lives in TMEM
  unsigned asel;
  int width, wcol;
  bool masked;       // dropout applies: staged through an x chunk slot instead of being read in place
};

// forward only (test / predict / get_states): round 2 dropped this kernel's backward pass — the FP32-FMA kernel is the fp32
// train path and the bf16 tile kernel (mmn_nb.cuh) the tensor-core one; see DESIGN.md section 4
template <int = 0>
__global__ void __launch_bounds__(V2Engine::kBlockThreads, 1) mmn_forward_kernel_v2(const StepArgs args) {
  using ENG = V2Engine;
  constexpr int TM = 128, NT = ENG::kWorkers;
  const DevPlan& P = *args.plan;
  const int tid = threadIdx.x;
  const int S = P.S, E = P.E, D = P.D, L = args.seq_len;
  const float* __restrict__ params = args.params;

  MMN_DYN_SMEM(smem_raw);
  ENG::Sm sm;
  ENG::carve(sm, smem_raw, P);
  TcState es;
  // ---- CTA setup (all threads) ----
  for (int i = tid; i < kV2Dz + 2 * kV2In + kV2Wb; i += ENG::kBlockThreads) sm.DZ[i] = 0.f;      // contiguous regions
  for (int i = tid; i < P.n_metrics; i += ENG::kBlockThreads) sm.met[i] = 0.0;
  for (int i = tid; i < E + 1; i += ENG::kBlockThreads) sm.cnt[i] = 0;
  for (int i = tid; i < TM; i += ENG::kBlockThreads) sm.rownan[i] = 0;
  if (tid == 0) {
    mbar_init(sm.bar, NT); mbar_init(sm.bar + 1, NT);
    mbar_init(sm.bar + 2, 1); mbar_init(sm.bar + 3, 1);
    mbar_fence_init();
  }
  if (tid < 32) tmem_alloc(sm.tslot, kV2TmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  es.tmem = *sm.tslot;
  es.seq = 0;
  es.pending[0] = es.pending[1] = 0;
  es.parity[0] = es.parity[1] = 0;
  for (int i = 0; i < 16; ++i) es.t[i] = 0;
  const long long t_kernel = MMN_CLOCK();
  if (tid >= NT) {
    ENG::issuer_loop(sm, es.tmem);
    return;
  }
  V2Thread t;
  t.tid = tid; t.lane = tid & 31; t.q = (tid >> 5) & 3; t.cs = tid >> 7; t.r = 32 * t.q + t.lane;
  t.lane_addr = es.tmem + ((unsigned)(32 * t.q) << 16);

  const long long n_tiles = (args.n_rows + TM - 1) / TM;
  Drop nodrop;
  nodrop.enabled = 0; nodrop.seed_mix = 0; nodrop.thr = 0; nodrop.row_base = 0; nodrop.scale = 1.f;

  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long row0 = tile * TM;
    const int rows_valid = (int)min((long long)TM, args.n_rows - row0);
    const bool valid = t.r < rows_valid;
    ENG::drain(sm, es);
    MMN_WSYNC_N(NT);
    if (args.targets) {
      for (int idx = tid; idx < TM * D; idx += NT) {
        const int rr = idx / D;
        long long y = 0;
        if (rr < rows_valid) y = args.targets[(row0 + rr) * D + (idx - rr * D)];
        if ((y < 0 || y >= P.dec[idx - rr * D].C) && args.target_error) *args.target_error = 1;
        sm.ys[idx] = (int)y;
      }
    }
    for (int i = tid; i < E + 1; i += NT) sm.tile_any[i] = i == 0;
    if (t.cs == 0) {
      sm.present[t.r] = valid;
      if (valid) atomicAdd(&sm.cnt[0], 1);
    }

    // =========================================================================================================
    // generic forward GEMM: ACC[r][n] = bias[n] + sum over x chunks and TMEM-resident segments; epi(h, n_base, v)
    // receives this thread's 16-column slices (h = 0: columns 16 cs .., h = 1: 32 + 16 cs ..) with the bias added.
    // mid() runs after every chunk has been handed to the issuer and before the accumulator is awaited.
    // =========================================================================================================
    auto gemm_fwd = [&](const float* __restrict__ W, int ldw, int N, const float* __restrict__ bias, const float* xptr,
                        long long xld, int F, const Drop& drop, const V2Res* res, int nres, auto mid, auto epi) {
      const bool xvec = xptr && v2_vec_ok(xptr, xld) && ((F & 3) == 0);
      const int nxc = xptr ? (F + 31) >> 5 : 0;
      int total = nxc;
      for (int s = 0; s < nres; ++s) total += (res[s].width + 31) >> 5;
      float wr[8], xr[16];
      // chunk c -> (segment, k0, wcol)
      auto locate = [&](int c, int& seg, int& k0, int& kw, int& wcol) {
        if (c < nxc) { seg = -1; k0 = 32 * c; kw = min(32, F - k0); wcol = k0; return; }
        c -= nxc;
        for (int s = 0; s < nres; ++s) {
          const int nc = (res[s].width + 31) >> 5;
          if (c < nc) { seg = s; k0 = 32 * c; kw = min(32, res[s].width - k0); wcol = res[s].wcol + k0; return; }
          c -= nc;
        }
        seg = -2; k0 = kw = wcol = 0;
      };
      auto prefetch = [&](int c) {
        int seg, k0, kw, wcol;
        locate(c, seg, k0, kw, wcol);
        TcEngine::w_load_k(wr, W, ldw, 0, N, wcol, kw, w_vec_ok(W, ldw, wcol));
        if (seg == -1) {
          if (valid) v2_load_row16<false>(xptr, xld, t.r, k0 + 16 * t.cs, F, xvec, xr);
          else {
#pragma unroll
            for (int i = 0; i < 16; ++i) xr[i] = 0.f;
          }
        }
      };
      const long long t_g0 = MMN_CLOCK();
      prefetch(0);
      int slot = 0;
      for (int c = 0; c < total; ++c) {
        int seg, k0, kw, wcol;
        locate(c, seg, k0, kw, wcol);
        slot = es.seq & 1;
        const long long t_w0 = MMN_CLOCK();
        ENG::wait(sm, es, slot);
        const long long t_w1 = MMN_CLOCK();
        es.t[0] += t_w1 - t_w0;
        float* wh = sm.WB + slot * 4096;
        TcEngine::w_store_k(wh, wh + 2048, wr, w_vec_ok(W, ldw, wcol));
        es.t[1] += MMN_CLOCK() - t_w1;        // weight block: global-load latency + split + swizzled stores
        es.t[7] += 1;
        unsigned asel, achunk = 0;
        const bool staged = seg == -1 || res[seg].masked;
        if (staged) {
          float v[16];
          if (seg == -1) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              v[i] = xr[i];
              if (v[i] != v[i]) { sm.rownan[t.r] = 1; v[i] = 0.f; }   // NaN: modality missing for this row
            }
          } else {                     // masked copy of a resident segment (dropout on the state columns)
            const unsigned hi = res[seg].asel == V2_A_S ? V2_S_HI : res[seg].asel == V2_A_P ? V2_P_HI : V2_Q_HI;
            const unsigned lo = res[seg].asel == V2_A_S ? V2_S_LO : res[seg].asel == V2_A_P ? V2_P_LO : V2_Q_LO;
            float l[16];
            v2_ld(t, hi + k0, v);
            v2_ld(t, lo + k0, l);
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += l[i];
          }
          if (drop.enabled) {
            const unsigned col = (unsigned)(wcol + 16 * t.cs), row = drop.row_base + (unsigned)t.r;
            if ((col & 1u) == 0) {
#pragma unroll
              for (int i = 0; i < 16; i += 2) {
                const unsigned hsh = mmn_dropout_hash(drop.seed_mix, row, (col + i) >> 1);
                v[i] = ((hsh & 0xffffu) >= drop.thr && 16 * t.cs + i < kw) ? v[i] * drop.scale : 0.f;
                v[i + 1] = ((hsh >> 16) >= drop.thr && 16 * t.cs + i + 1 < kw) ? v[i + 1] * drop.scale : 0.f;
              }
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i)
                v[i] = (mmn_dropout_keep(drop.seed_mix, row, col + i, drop.thr) && 16 * t.cs + i < kw) ? v[i] * drop.scale : 0.f;
            }
          }
          v2_st_pair(t, slot ? V2_Q_HI : V2_X_HI, slot ? V2_Q_LO : V2_X_LO, v);
          tmem_wait_st();
          asel = slot ? V2_A_Q : V2_A_X;
        } else {
          asel = res[seg].asel;
          achunk = (unsigned)(k0 >> 5);
        }
        const long long t_p0 = MMN_CLOCK();
        if (c + 1 < total) prefetch(c + 1);
        ENG::post(sm, es, slot, ENG::op_ts(false, N, c == 0, (kw + 7) >> 3, asel, (int)achunk));
        es.t[2] += MMN_CLOCK() - t_p0;        // prefetch issue + post
      }
      const long long t_b0 = MMN_CLOCK();
      // bias of this thread's two 16-column slices (offsets are 16-byte aligned)
      float4 b4[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int n = 32 * (u >> 2) + 16 * t.cs + 4 * (u & 3);
        float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
        if (n + 3 < N) b = __ldg(reinterpret_cast<const float4*>(bias + n));
        else if (n < N) { b.x = __ldg(bias + n); if (n + 1 < N) b.y = __ldg(bias + n + 1); if (n + 2 < N) b.z = __ldg(bias + n + 2); }
        b4[u] = b;
      }
      mid();
      const long long t_a0 = MMN_CLOCK();
      es.t[3] += t_a0 - t_b0;                 // bias loads + mid
      ENG::wait(sm, es, slot ^ 1);
      ENG::wait(sm, es, slot);
      tc_fence_after();
      const long long t_e0 = MMN_CLOCK();
      es.t[4] += t_e0 - t_a0;                 // waiting for the accumulator
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (32 * h < N) {
          float v[16];
          v2_ld(t, V2_ACC + 32 * h, v);
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            v[4 * u] += b4[4 * h + u].x; v[4 * u + 1] += b4[4 * h + u].y;
            v[4 * u + 2] += b4[4 * h + u].z; v[4 * u + 3] += b4[4 * h + u].w;
          }
          epi(h, 32 * h + 16 * t.cs, v);
        }
      }
      tmem_wait_st();
      es.t[5] += MMN_CLOCK() - t_e0;          // epilogue
      es.t[6] += MMN_CLOCK() - t_g0;          // whole GEMM
      es.t[8] += 1;
    };
    auto no_mid = [] {};

    // ---- decoders on the current state (A = S in TMEM) ----
    auto decoders_forward = [&](int k, int hist_row, bool is_last_enc, bool pr) {
      for (int d = 0; d < D; ++d) {
        const DevDecoder& dec = P.dec[d];
        unsigned in_sel = V2_A_S;
        int in_w = S;
        for (int j = 0; j < dec.n_layers; ++j) {
          const DevLayer& ly = dec.L[j];
          const bool last = j == dec.n_layers - 1;
          const unsigned out_sel = (j & 1) ? V2_A_Q : V2_A_P;
          const int N = ly.out_dim, act = ly.act, C = dec.C;
          V2Res rs;
          rs.asel = in_sel; rs.width = in_w; rs.wcol = 0; rs.masked = false;
          gemm_fwd(params + ly.w_off, ly.ktot, N, params + ly.b_off, nullptr, 0, 0, nodrop, &rs, 1, no_mid,
                   [&](int h, int nb, float (&v)[16]) {
                     const long long te0 = MMN_CLOCK();
                     act_fwd_n(act, v);
#pragma unroll
                     for (int i = 0; i < 16; ++i) v[i] = nb + i < N ? v[i] : 0.f;
                     if (!last) {
                       v2_st_pair(t, (out_sel == V2_A_P ? V2_P_HI : V2_Q_HI) + 32 * h, (out_sel == V2_A_P ? V2_P_LO : V2_Q_LO) + 32 * h, v);
                       es.t[9] += MMN_CLOCK() - te0;
                     } else if (t.cs == 0 && h == 0) {
                       // per-row epilogue: first-max arg-max, CE on the outputs, confusion cells (whole warps 0..3)
                       float best = v[0], mx = v[0], se, py;
                       int pred = 0, y = 0;
                       if (args.targets) {
                         y = sm.ys[t.r * D + d];
                         y = y < 0 ? 0 : (y >= C ? C - 1 : y);
                       }
                       if (C == 2) {                       // binary heads: every reference pipeline
                         pred = (v[1] > v[0] || (v[1] != v[1] && v[0] == v[0])) ? 1 : 0;
                         mx = fmaxf(v[0], v[1]);
                         se = 1.f + expf(-fabsf(v[0] - v[1]));
                         py = y ? v[1] : v[0];
                       } else {
#pragma unroll
                         for (int c = 1; c < 16; ++c)
                           if (c < C) {
                             if (v[c] > best || (v[c] != v[c] && best == best)) { best = v[c]; pred = c; }
                             mx = fmaxf(mx, v[c]);
                           }
                         se = 0.f;
                         py = v[0];
#pragma unroll
                         for (int c = 0; c < 16; ++c)
                           if (c < C) { se += expf(v[c] - mx); if (c == y) py = v[c]; }
                       }
                       if (valid) {
                         if (args.predictions) args.predictions[((long long)hist_row * D + d) * args.pred_ld + row0 + t.r] = (unsigned char)pred;
                         if (args.last_outputs && is_last_enc) {
                           float* o = args.last_outputs + (row0 + t.r) * P.sumC + dec.out_off;
#pragma unroll
                           for (int c = 0; c < 16; ++c)
                             if (c < C) o[c] = v[c];
                         }
                       }
                       if (args.targets) {
                         float ce = pr ? (mx + logf(se) - py) : 0.f;
                         unsigned pk1 = 0, pk2 = 0;
                         if (pr) {
                           pk1 = (pred == y ? 1u : 0u);
                           if (C == 2) {
                             pk1 |= (pred == 1 && y == 1 ? 1u << 8 : 0u) | (pred == 0 && y == 0 ? 1u << 16 : 0u) |
                                    (pred == 1 && y == 0 ? 1u << 24 : 0u);
                             pk2 = (pred == 0 && y == 1 ? 1u : 0u);
                           }
                         }
                         ce = warp_sum(ce);
                         pk1 = warp_sum_u(pk1);
                         pk2 = warp_sum_u(pk2);
                         if (t.lane == 0) {
                           atomicAdd(&sm.met[met_mat(P, 0, hist_row, d)], (double)ce);
                           atomicAdd(&sm.met[met_mat(P, 1, hist_row, d)], (double)(pk1 & 0xff));
                           atomicAdd(&sm.met[met_mat(P, 2, hist_row, d)], (double)((pk1 >> 8) & 0xff));
                           atomicAdd(&sm.met[met_mat(P, 3, hist_row, d)], (double)((pk1 >> 16) & 0xff));
                           atomicAdd(&sm.met[met_mat(P, 4, hist_row, d)], (double)((pk1 >> 24) & 0xff));
                           atomicAdd(&sm.met[met_mat(P, 5, hist_row, d)], (double)(pk2 & 0xff));
                         }
                       }
                       es.t[10] += MMN_CLOCK() - te0;
                     }
                   });
          in_sel = out_sel;
          in_w = N;
        }
      }
    };

    // ---- s_0 (state.py:29-32) ----
    for (int h = 0; h < ((S + 31) >> 5); ++h) {
      float v[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int c = 32 * h + 16 * t.cs + i;
        v[i] = c < S ? __ldg(params + P.init_off + c) : 0.f;
      }
      v2_st_pair(t, V2_S_HI + 32 * h, V2_S_LO + 32 * h, v);
    }
    tmem_wait_st();
    MMN_WSYNC_N(NT);
    decoders_forward(0, 0, false, valid);

    // ---- walk the encoding sequence ----
    for (int k = 1; k <= L; ++k) {
      const int e = args.seq_enc[k - 1], pos = args.seq_pos[k - 1];
      const DevEncoder& enc = P.enc[e];
      const bool skip = args.skip_flags && args.skip_flags[k - 1] != 0;
      bool pr = false;
      if (!skip) {
        const Drop drop = nodrop;                 // evaluation mode: no dropout (multimodn.py:267)
        unsigned in_sel = V2_A_P;
        int in_w = 0;
        float sc = 0.f;
        for (int j = 0; j < enc.n_layers; ++j) {
          const DevLayer& ly = enc.L[j];
          const bool last = j == enc.n_layers - 1;
          const unsigned out_sel = (j & 1) ? V2_A_Q : V2_A_P;
          const bool use_drop = drop.enabled && j == 0;
          V2Res rs[2];
          int nres = 0;
          if (j > 0) { rs[nres].asel = in_sel; rs[nres].width = in_w; rs[nres].wcol = 0; rs[nres].masked = false; ++nres; }
          if (ly.has_state) { rs[nres].asel = V2_A_S; rs[nres].width = S; rs[nres].wcol = ly.in_dim; rs[nres].masked = use_drop; ++nres; }
          const int N = ly.out_dim, act = ly.act;
          auto mid = [&] {
            if (j == 0) {           // every x chunk of the step has been scanned: fix this step's present mask
              MMN_WSYNC_N(NT);
              pr = valid && sm.rownan[t.r] == 0;
            }
          };
          gemm_fwd(params + ly.w_off, ly.ktot, N, params + ly.b_off, j == 0 ? args.x[pos] + row0 * args.x_ld[pos] : nullptr,
                   args.x_ld[pos], j == 0 ? ly.in_dim : 0, use_drop ? drop : nodrop, rs, nres, mid,
                   [&](int h, int nb, float (&v)[16]) {
                     act_fwd_n(act, v);
#pragma unroll
                     for (int i = 0; i < 16; ++i) v[i] = nb + i < N ? v[i] : 0.f;
                     if (!last) {
                       v2_st_pair(t, (out_sel == V2_A_P ? V2_P_HI : V2_Q_HI) + 32 * h, (out_sel == V2_A_P ? V2_P_LO : V2_Q_LO) + 32 * h, v);
                     } else {
                       // per-row select: missing rows keep their state bit for bit; state-change sum
                       float so[16], sl[16];
                       v2_ld(t, V2_S_HI + 32 * h, so);
                       v2_ld(t, V2_S_LO + 32 * h, sl);
#pragma unroll
                       for (int i = 0; i < 16; ++i) {
                         const float old = so[i] + sl[i];        // exact: lo = v - hi
                         const float nw = pr ? v[i] : old;
                         const float df = nw - old;
                         sc = fmaf(df, df, sc);
                         v[i] = nw;
                       }
                       v2_st_pair(t, V2_S_HI + 32 * h, V2_S_LO + 32 * h, v);   // unconditional: tcgen05.st is warp-collective
                     }
                   });
          in_sel = out_sel;
          in_w = N;
        }
        (void)sc;
      }
      // bookkeeping of the step: present mask, counters; stash s_k; reset the NaN flags
      if (t.cs == 0) {
        sm.present[k * TM + t.r] = pr;
        if (pr) { atomicAdd(&sm.cnt[e + 1], 1); sm.tile_any[k] = 1; }
      }
      if (args.final_state && k == L) {
        for (int h = 0; h < ((S + 31) >> 5); ++h) {
          float so[16], sl[16];
          v2_ld(t, V2_S_HI + 32 * h, so);
          v2_ld(t, V2_S_LO + 32 * h, sl);
#pragma unroll
          for (int i = 0; i < 16; ++i) so[i] += sl[i];
          if (args.final_state && k == L && valid) v2_store_row16(args.final_state + row0 * S, S, t.r, 32 * h + 16 * t.cs, S, so);
        }
      }
      MMN_WSYNC_N(NT);                 // every reader of this step's NaN flags is done
      if (t.cs == 0) sm.rownan[t.r] = 0;
      decoders_forward(k, e + 1, e == E - 1, pr);
    }
    if (args.final_state && L == 0 && valid) {
      for (int h = 0; h < ((S + 31) >> 5); ++h) {
        float so[16], sl[16];
        v2_ld(t, V2_S_HI + 32 * h, so);
        v2_ld(t, V2_S_LO + 32 * h, sl);
#pragma unroll
        for (int i = 0; i < 16; ++i) so[i] += sl[i];
        v2_store_row16(args.final_state + row0 * S, S, t.r, 32 * h + 16 * t.cs, S, so);
      }
    }
  }

  // ---- teardown: stop the issuer, free TMEM, flush this CTA's metric partials ----
  ENG::drain(sm, es);
  if (args.debug_timers && tid == 0) {
    es.t[15] = MMN_CLOCK() - t_kernel;
    for (int i = 0; i < 16; ++i) args.debug_timers[blockIdx.x * 16 + i] = es.t[i];
  }
  {
    const int slot = es.seq & 1;
    if (tid == 0) sm.cmd[slot].op = V2_OP_QUIT;
    mbar_arrive(sm.bar + slot);
  }
  tc_fence_before();
  MMN_WSYNC_N(NT);
  if (tid < 32) tmem_dealloc(es.tmem, kV2TmemCols);
  if (args.metrics) {
    const int nmat = 6 * (E + 1) * D;
    for (int i = tid; i < P.n_metrics; i += NT) {
      double v;
      if (i < (E + 1) * D) v = sm.met[i] * args.inv_rows_global;
      else if (i < nmat) v = sm.met[i];
      else if (i < nmat + E + 1) v = (double)sm.cnt[i - nmat];
      else v = sm.met[i] * args.inv_rows_global / (double)S;
      if (v != 0.0) atomicAdd(args.metrics + i, v);
    }
  }
}

}  // namespace mmn

namespace mmn {
// ------------------------------------------------------------------------------------------------
// latency probe of the worker <-> issuer protocol (development aid, exported as mmn_selftest_protocol):
// `iters` rounds of [workers: (fence.proxy.async) -> arrive full] -> [issuer: n_mma MMAs -> commit] ->
// [workers: wait done -> tcgen05.ld].  out[0] = cycles per round (thread 0).
//   flags bit 0: workers execute fence.proxy.async        bit 1: workers execute tcgen05.ld each round
//   flags bit 2: only warp 0 participates as worker (full barrier of 32 arrivals)
// ------------------------------------------------------------------------------------------------
template <int = 0>
__global__ void __launch_bounds__(288, 1) mmn_protocol_probe_kernel(int iters, int n_mma, int flags, long long* out) {
  MMN_DYN_SMEM(raw);
  char* base = raw + ((1024 - (smem_u32(raw) & 1023)) & 1023);
  float* img = reinterpret_cast<float*>(base);                               // 2 x 16 KB zero images
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(img + 8192);
  unsigned* slot = reinterpret_cast<unsigned*>(bar + 4);
  const int tid = threadIdx.x;
  const bool few = flags & 4;
  const int nworkers = few ? 32 : 256;
  for (int i = tid; i < 8192; i += 288) img[i] = 0.f;
  if (tid == 0) { mbar_init(bar, nworkers); mbar_init(bar + 1, 1); mbar_fence_init(); }
  if (tid < 32) tmem_alloc(slot, 64);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const unsigned tmem = *slot;
  if (tid >= 256) {                                  // issuer warp
    const bool leader = elect_one() != 0;
    const unsigned a = smem_u32(img), b = smem_u32(img + 4096);
    const unsigned id = umma_idesc_tf32(128, 32, 0, 0);
    unsigned par = 0;
    for (int it = 0; it < iters; ++it) {
      mbar_wait(bar, par);
      par ^= 1u;
      tc_fence_after();
      if (leader) {
        for (int m = 0; m < n_mma; ++m) umma_tf32(tmem, umma_desc_k(a, m & 3), umma_desc_k(b, m & 3), id, m ? 1u : 0u);
        umma_commit(bar + 1);
      }
      __syncwarp();
    }
  } else if (tid < nworkers) {
    unsigned par = 0;
    const long long t0 = MMN_CLOCK();
    float acc = 0.f;
    for (int it = 0; it < iters; ++it) {
      if (flags & 1) fence_proxy_async();
      tc_fence_before();
      mbar_arrive(bar);
      mbar_wait(bar + 1, par);
      par ^= 1u;
      tc_fence_after();
      if (flags & 2) {
        float v[16];
        tmem_ld16(tmem + ((unsigned)(32 * ((tid >> 5) & 3)) << 16), v);
        acc += v[0];
      }
    }
    if (tid == 0) { out[0] = (MMN_CLOCK() - t0) / iters; out[1] = (long long)acc; }
  }
  tc_fence_before();
  __syncthreads();
  if (tid < 32) tmem_dealloc(tmem, 64);
}
}  // namespace mmn
