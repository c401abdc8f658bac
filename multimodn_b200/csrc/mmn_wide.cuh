// mmn_wide.cuh — the wide regime (BASELINE config 4: state 1024, hidden 2048, bf16): layers that are real dense
// contractions run layer by layer as persistent tcgen05 GEMMs instead of inside the per-tile step kernel.
//
//   D[M x N] = A[M x K] . B[N x K]^T        A, B bf16, K contiguous (K-major); fp32 accumulation in tensor memory
//
// One kernel shape serves forward (A = activations, B = W), data gradient (A = dZ, B = W^T copy) and weight gradient
// (A = dZ, B = X, both read in place as MN-major operands: the contraction runs over the batch rows).
//
//   warp 0   TMA producer: cp.async.bulk.tensor.2d (SWIZZLE_128B boxes of 64 bf16 x 128 / 256 rows) into a 4-stage ring
//            (6 stages of 32 KB when two CTAs share a 256 x 256 tile, PAIR)
//   warp 1   MMA issuer: tcgen05.mma.cta_group::1 / ::2 .kind::f16, N256 K16, 4 per stage; tcgen05.commit frees the stage
//   warp 2   TMEM allocator (512 columns = two 128 x 256 fp32 accumulators: the epilogue of tile i overlaps tile i+1)
//   warps 4-11 epilogue (two per TMEM lane quarter, 128 columns each): tcgen05.ld 16 columns at a time -> bias /
//            activation / select / derivative / accumulation.  Interior tiles run epi_lean<MODE> (one straight-line loop
//            per mode, bias slice in shared memory, row operands prefetched, stores transposed through a swizzled per-warp
//            staging buffer so that each store instruction writes 8 rows x 64 contiguous bytes); edge tiles and the
//            transposed-copy output of the self-test run the generic loop below it.
//
// Out-of-range rows / columns / k are zero-filled by the TMA unit, so M, N, K need no padding (row pitches: 16 bytes).
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>

#include "mmn_tc.cuh"

namespace mmn {
namespace wide {

constexpr int BM = 128, BN = 256, BK = 64, STAGES = 4;
constexpr int kThreads = 384;          // warps 0-2: TMA / MMA / TMEM allocator; warps 4-11: epilogue
constexpr int kStageBytes = (BM + BN) * BK * 2;                  // 48 KB
constexpr int kBiasStageBytes = 2 * BN * 4;                      // the epilogue's bias slice of a tile, double-buffered
constexpr int kEpiStageBytes = 8 * 32 * 64;                      // per epilogue warp: 32 rows x 64 bytes of output being transposed
constexpr int kSmemBytes = STAGES * kStageBytes + 1024 + 256 + kBiasStageBytes + kEpiStageBytes;   // + alignment slack, barriers, bias

enum {
  EPI_STORE = 0,       // out = act(acc + bias)                                   (forward hidden / decoder layers)
  EPI_SELECT = 1,      // out = present[r] ? act(acc + bias) : aux[r][n]; sc += (out - aux)^2   (encoder output = new state)
  EPI_DACT = 2,        // out = acc * act'(aux[r][n])                             (data gradient into a hidden layer)
  EPI_ACCUM_F32 = 3,   // out_f32 (+)= acc                                         (weight gradient; state gradient G += ...)
  EPI_CARRY = 4,       // out_f32 = (present[r] ? acc : out_f32) - c_sc (aux - aux2)   (state carry through an encoder; the
                       //                                                          state-change term changes sign for s_{k-1})
  EPI_NONE = 5,        // accumulators dropped (MMN_WIDE_EPI_DEBUG=none: main-loop-only timing, development aid)
};

struct Epi {
  int mode, act, accumulate;
  const float* bias;                                  // [N] or null
  float* out_f32; long long ld_f32;                   // row-major fp32
  __nv_bfloat16* out; long long ld_out;               // row-major bf16
  __nv_bfloat16* out_t; long long ld_out_t;           // transposed bf16 [N x M]
  const __nv_bfloat16* aux; long long ld_aux;         // activation for act' / previous state for the select
  const __nv_bfloat16* aux2; long long ld_aux2; float c_sc;   // EPI_CARRY: aux = s_k, aux2 = s_{k-1}
  const unsigned char* present;                       // [M]
  const int* skip;                                    // device flag: non-zero = the step is skipped for every row
  unsigned drop_thr, drop_seed, drop_row_base, drop_col_base;   // EPI_CARRY through a dropout mask (thr 0 = none)
  float* sc_sum;                                      // state-change accumulator (one float)
  float scale;                                        // multiplies acc before everything else (EPI_ACCUM_F32)
};

__device__ __forceinline__ unsigned umma_idesc_bf16(int M, int N) {
  // kind::f16: c_format = F32 (1) @4, a_format = b_format = BF16 (1) @7 / @10, both K-major, N>>3 @17, M>>4 @24
  return (1u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(N >> 3) << 17) | ((unsigned)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(unsigned tmem_d, unsigned long long adesc, unsigned long long bdesc, unsigned idesc,
                                          unsigned accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc),
      "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, unsigned long long* bar, void* dst, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(dst)), "l"(reinterpret_cast<unsigned long long>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<unsigned long long>(map)) : "memory");
}


// ---- CTA pair (cta_group::2): two CTAs of a cluster share one M256 x N256 MMA; each holds its own 128 rows of A and one
//      half (128 rows) of B in shared memory, the leader (cluster rank 0) issues the MMAs, accumulators stay per CTA ----
__device__ __forceinline__ unsigned cluster_ctarank() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// the same shared-memory offset in the leader CTA (bit 24 of a shared::cluster address selects the CTA of the pair)
constexpr unsigned kPeerBitMask = 0xFEFFFFFFu;
__device__ __forceinline__ void tma_load_2d_pair(const CUtensorMap* map, unsigned long long* leader_bar, void* dst, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(dst)), "l"(reinterpret_cast<unsigned long long>(map)), "r"(smem_u32(leader_bar) & kPeerBitMask), "r"(c0),
               "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(unsigned tmem_d, unsigned long long adesc, unsigned long long bdesc, unsigned idesc,
                                               unsigned accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc),
      "r"(accumulate) : "memory");
}
// tcgen05.commit arriving on the same barrier offset in both CTAs of the pair
__device__ __forceinline__ void umma_commit_pair(unsigned long long* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((unsigned short)3) : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(unsigned long long* bar) {      // arrive on the LEADER CTA's copy of `bar`
  asm volatile("{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, 0;\n\t"
               "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(unsigned* slot, unsigned ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(unsigned taddr, unsigned ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// Programmatic dependent launch: every kernel of the wide step lets its successor in the stream be scheduled at once
// (launch_dependents) and waits for its own predecessor to have completed and flushed (wait) before it touches global memory.
// The successor's CTAs then start as this grid's CTAs retire — launch latency and the GEMM's prologue (barrier init, TMEM
// allocation, descriptor prefetch) hide behind the predecessor's tail.  Without the launch attribute both are no-ops.
__device__ __forceinline__ void pdl_entry() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

struct Pipe {
  int stage;
  unsigned phase;
  __device__ __forceinline__ void advance(int n_stages = STAGES) {
    if (++stage == n_stages) { stage = 0; phase ^= 1u; }
  }
};

__device__ __forceinline__ float wide_act(int act, float z) {
  switch (act) {
    case MMN_ACT_RELU: return fmaxf(z, 0.f);
    case MMN_ACT_SIGMOID: return 1.f / (1.f + __expf(-z));
    case MMN_ACT_TANH: return tanhf(z);
    default: return z;
  }
}
__device__ __forceinline__ float wide_dact(int act, float a) {      // derivative from the OUTPUT of the activation
  switch (act) {
    case MMN_ACT_RELU: return a > 0.f ? 1.f : 0.f;
    case MMN_ACT_SIGMOID: return a * (1.f - a);
    case MMN_ACT_TANH: return 1.f - a * a;
    default: return 1.f;
  }
}

// ---- the epilogue of one warp on an interior tile: every row and column exists and every row of every operand starts on a
// 16-byte boundary, so the loop over 16-column groups is straight-line per mode (the generic loop in the kernel spends ~260
// instructions per group on mode / bounds / alignment decisions and, with two epilogue warps per scheduler, that issue
// latency — not memory — is what a tile's epilogue costs).  Row-wise operands are fetched one group ahead. ----
__device__ __forceinline__ void unpack16(const uint4& a0, const uint4& a1, float (&f)[16]) {
  const __nv_bfloat162* h0 = reinterpret_cast<const __nv_bfloat162*>(&a0);
  const __nv_bfloat162* h1 = reinterpret_cast<const __nv_bfloat162*>(&a1);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f0 = __bfloat1622float2(h0[i]), f1 = __bfloat1622float2(h1[i]);
    f[2 * i] = f0.x; f[2 * i + 1] = f0.y; f[8 + 2 * i] = f1.x; f[8 + 2 * i + 1] = f1.y;
  }
}
__device__ __forceinline__ void store16_bf16(__nv_bfloat16* op, const float (&v)[16]) {
  uint4 o[2];
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(o);
#pragma unroll
  for (int i = 0; i < 8; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  reinterpret_cast<uint4*>(op)[0] = o[0];
  reinterpret_cast<uint4*>(op)[1] = o[1];
}
// Stores go through a per-warp staging buffer (32 rows x 64 bytes, 16-byte chunks XOR-swizzled by (row / 2) % 4 so that
// both the row-wise writes and the 4-lanes-per-row reads are bank-conflict free): a thread owns one accumulator ROW, and a
// direct store makes every warp instruction touch 32 lines with half-filled 32-byte sectors — measured 8.6 us of a 35 us
// K = 1024 GEMM.  Staged, one instruction writes 8 rows x 64 contiguous bytes (full sectors).
template <int MODE>
__device__ __forceinline__ void epi_lean(const Epi& epi, unsigned taddr, int c_beg, int c_end, long long r, int n0, bool pres,
                                         const float* bias_t, bool atomic, unsigned long long* acc_full_bar, unsigned parity,
                                         char* stg, float& sc) {
  constexpr bool kAux = MODE == EPI_SELECT || MODE == EPI_DACT || MODE == EPI_CARRY;
  constexpr bool kAux2 = MODE == EPI_CARRY;
  const int lane = threadIdx.x & 31;
  const long long rbase = r - lane;                       // first row of this warp's 32
  const bool f32_out = MODE == EPI_CARRY || (MODE == EPI_ACCUM_F32 && epi.out_f32);
  const __nv_bfloat16* const aux_row = kAux ? epi.aux + r * epi.ld_aux + n0 : nullptr;
  const __nv_bfloat16* const aux2_row = kAux2 ? epi.aux2 + r * epi.ld_aux2 + n0 : nullptr;
  // staging: this lane's row on the write side, (row rr + 8 j, chunk rc) on the read side
  char* const my_row = stg + lane * 64;
  const unsigned wsw = (unsigned)(lane >> 1) & 3u;
  const int rr = lane >> 2, rc = lane & 3;
  auto st_chunk = [&](int ch, const uint4& val) { *reinterpret_cast<uint4*>(my_row + (((unsigned)ch ^ wsw) << 4)) = val; };
  auto ld_chunk = [&](int j) {
    const int row = rr + 8 * j;
    return *reinterpret_cast<const uint4*>(stg + row * 64 + (((unsigned)rc ^ ((unsigned)(row >> 1) & 3u)) << 4));
  };
  // fp32 output, read side: the four rows this lane stores to, whether they take the old value (accumulation; carry: rows the
  // encoder did not touch), and that old value fetched one group ahead
  float* f32_t[4] = {nullptr, nullptr, nullptr, nullptr};
  bool add_old[4] = {false, false, false, false};
  if (f32_out) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      f32_t[j] = epi.out_f32 + (rbase + rr + 8 * j) * epi.ld_f32 + n0 + rc * 4;
      const bool pres_j = __shfl_sync(0xffffffffu, pres ? 1 : 0, rr + 8 * j) != 0;
      add_old[j] = MODE == EPI_CARRY ? !pres_j : (epi.accumulate && !atomic);
    }
  }
  __nv_bfloat16* bf_t[4] = {nullptr, nullptr, nullptr, nullptr};
  if (!f32_out) {
#pragma unroll
    for (int j = 0; j < 4; ++j) bf_t[j] = epi.out + (rbase + rr + 8 * j) * epi.ld_out + n0 + rc * 8;
  }
  const int act = epi.act;
  const float scale = epi.scale;
  uint4 aq[2] = {make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0)}, bq[2] = {aq[0], aq[0]};
  float4 oq[4] = {make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f),
                  make_float4(0.f, 0.f, 0.f, 0.f)};
  auto fetch = [&](int c0) {
    if (kAux) { aq[0] = reinterpret_cast<const uint4*>(aux_row + c0)[0]; aq[1] = reinterpret_cast<const uint4*>(aux_row + c0)[1]; }
    if (kAux2) { bq[0] = reinterpret_cast<const uint4*>(aux2_row + c0)[0]; bq[1] = reinterpret_cast<const uint4*>(aux2_row + c0)[1]; }
    if (f32_out) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (add_old[j]) oq[j] = *reinterpret_cast<const float4*>(f32_t[j] + c0);
    }
  };
  fetch(c_beg);
  mbar_wait(acc_full_bar, parity);
  tc_fence_after();
  for (int c0 = c_beg; c0 < c_end; c0 += 16) {
    const uint4 a0 = aq[0], a1 = aq[1], b0 = bq[0], b1 = bq[1];
    const float4 old0 = oq[0], old1 = oq[1], old2 = oq[2], old3 = oq[3];
    if (c0 + 16 < c_end) fetch(c0 + 16);
    float v[16];
    tmem_ld16(taddr + c0, v);
    if (MODE == EPI_STORE || MODE == EPI_SELECT) {
      if (epi.bias) {
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
          const float4 b = *reinterpret_cast<const float4*>(bias_t + c0 + i);
          v[i] += b.x; v[i + 1] += b.y; v[i + 2] += b.z; v[i + 3] += b.w;
        }
      }
      if (act == MMN_ACT_RELU) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
      } else if (act == MMN_ACT_SIGMOID) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = __fdividef(1.f, 1.f + __expf(-v[i]));
      } else if (act == MMN_ACT_TANH) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = tanhf(v[i]);
      }
      if (MODE == EPI_SELECT) {
        float aux[16];
        unpack16(a0, a1, aux);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          // what the next layer reads is the bf16-rounded state: measure the change on that
          const float nw = pres ? __bfloat162float(__float2bfloat16(v[i])) : aux[i];
          const float df = nw - aux[i];
          sc = fmaf(df, df, sc);
          v[i] = nw;
        }
      }
    } else if (MODE == EPI_DACT) {
      float aux[16];
      unpack16(a0, a1, aux);
      if (act == MMN_ACT_RELU) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = aux[i] > 0.f ? v[i] : 0.f;
      } else if (act == MMN_ACT_SIGMOID) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] *= aux[i] * (1.f - aux[i]);
      } else if (act == MMN_ACT_TANH) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] *= 1.f - aux[i] * aux[i];
      }
    } else if (MODE == EPI_ACCUM_F32) {
      if (scale != 1.f) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] *= scale;
      }
    } else {      // EPI_CARRY: the part that does not depend on the old value; absent rows contribute 0 here and keep G below
      if (epi.drop_thr) {
#pragma unroll
        for (int i = 0; i < 16; ++i)
          v[i] = mmn_dropout_keep(epi.drop_seed, epi.drop_row_base + (unsigned)r, epi.drop_col_base + (unsigned)(n0 + c0 + i), epi.drop_thr)
                     ? v[i] * scale : 0.f;
      } else if (scale != 1.f) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] *= scale;
      }
      float aux[16], b[16];
      unpack16(a0, a1, aux);
      unpack16(b0, b1, b);
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = (pres ? v[i] : 0.f) - epi.c_sc * (aux[i] - b[i]);
    }
    if (f32_out) {
      // 16 fp32 = this row's 64 staged bytes; flushed every group
#pragma unroll
      for (int i = 0; i < 4; ++i) st_chunk(i, make_uint4(__float_as_uint(v[4 * i]), __float_as_uint(v[4 * i + 1]), __float_as_uint(v[4 * i + 2]),
                                                         __float_as_uint(v[4 * i + 3])));
      __syncwarp();
      const float4 olds[4] = {old0, old1, old2, old3};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint4 u = ld_chunk(j);
        float4 o = make_float4(__uint_as_float(u.x), __uint_as_float(u.y), __uint_as_float(u.z), __uint_as_float(u.w));
        float* dst = f32_t[j] + c0;
        if (MODE == EPI_ACCUM_F32 && atomic) {
          atomicAdd(reinterpret_cast<float4*>(dst), o);
        } else {
          if (add_old[j]) { o.x += olds[j].x; o.y += olds[j].y; o.z += olds[j].z; o.w += olds[j].w; }
          *reinterpret_cast<float4*>(dst) = o;
        }
      }
      __syncwarp();
    } else {
      // 16 bf16 = half of this row's 64 staged bytes; flushed every second group
      uint4 o[2];
      __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(o);
#pragma unroll
      for (int i = 0; i < 8; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
      const int odd = ((c0 - c_beg) >> 4) & 1;
      st_chunk(2 * odd, o[0]);
      st_chunk(2 * odd + 1, o[1]);
      if (odd) {
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 4; ++j) *reinterpret_cast<uint4*>(bf_t[j] + (c0 - 16)) = ld_chunk(j);
        __syncwarp();
      }
    }
  }
}

// PAIR = true: launched as clusters of two CTAs working on one 256 x 256 output tile (K-major operands, no split-K);
// each CTA stages 16 KB of A + 16 KB of B per k-step instead of 16 + 32 and the ring is 6 stages deep.
constexpr int kPairStages = 6, kPairStageBytes = (BM + BN / 2) * BK * 2;       // 32 KB
template <bool PAIR>
__global__ void __launch_bounds__(kThreads, 1)
mmn_wide_gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                     const __grid_constant__ CUtensorMap map_b_half, int M, int N, int K, int splits, int a_mn, int b_mn,
                     int tail_div, const Epi epi) {
  constexpr int NSTG = PAIR ? kPairStages : STAGES, STG = PAIR ? kPairStageBytes : kStageBytes;
  extern __shared__ __align__(1024) char smem_raw[];
  char* base = smem_raw + ((1024 - (smem_u32(smem_raw) & 1023)) & 1023);
  unsigned long long* full = reinterpret_cast<unsigned long long*>(base + NSTG * STG);
  unsigned long long* empty = full + NSTG;
  unsigned long long* acc_full = empty + NSTG;      // [2]
  unsigned long long* acc_empty = acc_full + 2;       // [2]
  unsigned* tslot = reinterpret_cast<unsigned*>(acc_empty + 2);
  float* bias_s = reinterpret_cast<float*>(base + NSTG * STG + 256);      // [2][BN]
  char* epi_stage = base + NSTG * STG + 256 + kBiasStageBytes;            // [8 warps][32 rows][64 bytes]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned pair_rank = PAIR ? cluster_ctarank() : 0u;
  const int cta_id = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x, n_ctas = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  constexpr int TILE_M = PAIR ? 2 * BM : BM;
  // work item = (output tile, K split): split-K fills the machine when a long contraction has few output tiles (weight
  // gradients); its partial sums are added with fp32 atomics (EPI_ACCUM_F32 only)
  const int tiles_m = (M + TILE_M - 1) / TILE_M, tiles_n = (N + BN - 1) / BN, n_out_tiles = tiles_m * tiles_n;
  // last-wave balancing (tail_div = 2 or 4, only with splits == 1): the output tiles left over after the last full wave are
  // issued as tail_div column slices of BN / tail_div columns each, so that the last wave is shorter instead of emptier
  const int full_items = tail_div > 1 ? (n_out_tiles / n_ctas) * n_ctas : n_out_tiles;
  const int n_tiles = tail_div > 1 ? full_items + tail_div * (n_out_tiles - full_items) : n_out_tiles * splits;
  const int tail_w = tail_div > 1 ? BN / tail_div : BN;
  const int kb_all = (K + BK - 1) / BK, kb_per = (kb_all + splits - 1) / splits;
  // item -> (output tile, K split, column slice: -1 = the whole 256-column tile)
  auto decode = [&](int item, int& tile, int& split, int& nhalf) {
    if (item < full_items || tail_div <= 1) { tile = item % n_out_tiles; split = item / n_out_tiles; nhalf = -1; }
    else { tile = full_items + (item - full_items) / tail_div; split = 0; nhalf = (item - full_items) % tail_div; }
  };

  if (threadIdx.x == 0) {
    for (int s = 0; s < NSTG; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(acc_full + s, 1); mbar_init(acc_empty + s, PAIR ? 512 : 256); }
    mbar_fence_init();
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
  }
  if (warp == 2) {
    if (PAIR) tmem_alloc_pair(tslot, 512);
    else tmem_alloc(tslot, 512);
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all();            // barriers of both CTAs are initialised before anyone signals them
  else __syncthreads();
  tc_fence_after();
  const unsigned tmem = *tslot;
  pdl_entry();            // everything above touched only shared and tensor memory

  if (warp == 0) {
    // ===== TMA producer =====
    if (elect_one()) {
      Pipe p{0, 0};
      for (int item = cta_id; item < n_tiles; item += n_ctas) {
        int tile, split, nhalf;
        decode(item, tile, split, nhalf);
        const int m0 = (tile % tiles_m) * TILE_M + (int)pair_rank * BM, n0 = (tile / tiles_m) * BN + (nhalf > 0 ? nhalf * tail_w : 0);
        const int kb0 = split * kb_per, kb1 = min(kb_all, kb0 + kb_per);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(empty + p.stage, p.phase ^ 1u);
          char* sa = base + p.stage * STG;
          char* sb = sa + BM * BK * 2;
          if (PAIR) {         // both CTAs load their share; every byte is counted on the leader's barrier
            // (in a pair launch map_b_half has 128-row boxes = half of a 256-column tile, map_b 64-row boxes = half of a
            //  128-column tail item)
            // a pair launch gets map_b_half with 128-row boxes (half of a full tile) and map_b with tail_w / 2-row boxes
            if (pair_rank == 0) mbar_expect_tx(full + p.stage, 2 * (BM * BK * 2 + (nhalf < 0 ? BN / 2 : tail_w / 2) * BK * 2));
            if (a_mn) {
#pragma unroll
              for (int g = 0; g < BM / 64; ++g) tma_load_2d_pair(&map_a, full + p.stage, sa + g * 8192, m0 + 64 * g, kb * BK);
            } else {
              tma_load_2d_pair(&map_a, full + p.stage, sa, kb * BK, m0);
            }
            const int nb0 = n0 + (int)pair_rank * (nhalf < 0 ? BN / 2 : tail_w / 2);      // this CTA's share of the B columns
            if (b_mn) {
#pragma unroll
              for (int g = 0; g < BN / 128; ++g)
                if (nhalf < 0 || g < tail_w / 128) tma_load_2d_pair(&map_b, full + p.stage, sb + g * 8192, nb0 + 64 * g, kb * BK);
            } else if (nhalf < 0) {
              tma_load_2d_pair(&map_b_half, full + p.stage, sb, kb * BK, nb0);
            } else {
              tma_load_2d_pair(&map_b, full + p.stage, sb, kb * BK, nb0);
            }
            p.advance(NSTG);
            continue;
          }
          mbar_expect_tx(full + p.stage, nhalf < 0 ? kStageBytes : (BM + tail_w) * BK * 2);
          if (a_mn) {            // MN-major operand [k rows x mn]: one 64 x 64 box (8 KB, 128-byte rows) per 64 of M
#pragma unroll
            for (int g = 0; g < BM / 64; ++g) tma_load_2d(&map_a, full + p.stage, sa + g * 8192, m0 + 64 * g, kb * BK);
          } else {
            tma_load_2d(&map_a, full + p.stage, sa, kb * BK, m0);
          }
          if (b_mn) {
#pragma unroll
            for (int g = 0; g < BN / 64; ++g)
              if (nhalf < 0 || g < tail_w / 64) tma_load_2d(&map_b, full + p.stage, sb + g * 8192, n0 + 64 * g, kb * BK);
          } else if (nhalf < 0) {
            tma_load_2d(&map_b, full + p.stage, sb, kb * BK, n0);
          } else {
            tma_load_2d(&map_b_half, full + p.stage, sb, kb * BK, n0);
          }
          p.advance();
        }
      }
    }
  } else if (warp == 1 && pair_rank == 0) {
    // ===== MMA issuer (the leader CTA of a pair) =====
    const bool leader = elect_one() != 0;
    const unsigned major_bits = ((unsigned)(a_mn != 0) << 15) | ((unsigned)(b_mn != 0) << 16);
    const unsigned idesc_full = umma_idesc_bf16(TILE_M, BN) | major_bits, idesc_half = umma_idesc_bf16(TILE_M, tail_w) | major_bits;
    const unsigned sbase = smem_u32(base);
    Pipe p{0, 0};
    unsigned acc_phase[2] = {0u, 0u};
    int it = 0;
    for (int item = cta_id; item < n_tiles; item += n_ctas, ++it) {
      const int acc = it & 1;
      int tile_, split_, nhalf_;
      decode(item, tile_, split_, nhalf_);
      const unsigned idesc = nhalf_ < 0 ? idesc_full : idesc_half;
      const int kb_n = min(kb_all, split_ * kb_per + kb_per) - split_ * kb_per;
      mbar_wait(acc_empty + acc, acc_phase[acc] ^ 1u);
      acc_phase[acc] ^= 1u;
      tc_fence_after();
      if (kb_n <= 0) {               // empty K range (cannot happen for splits <= kb_all; keep the pipeline consistent)
        if (leader) umma_commit(acc_full + acc);
        __syncwarp();
        continue;
      }
      for (int kb = 0; kb < kb_n; ++kb) {
        mbar_wait(full + p.stage, p.phase);
        tc_fence_after();
        if (leader) {
          const unsigned sa = sbase + p.stage * STG, sb = sa + BM * BK * 2;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // K-major: 16 k = 32 bytes inside the swizzled 128-byte row, 8-row atoms 1024 B apart.
            // MN-major: 64 mn = one 128-byte row per k, 8-k atoms 1024 B apart (SBO), 64-mn groups 8 KB apart (LBO);
            //           16 k = two atoms = 2048 bytes.
            const unsigned long long da = a_mn ? umma_smem_desc(sa + 2048u * k, 8192, 1024, 2) : umma_desc_k(sa, k);
            const unsigned long long db = b_mn ? umma_smem_desc(sb + 2048u * k, 8192, 1024, 2) : umma_desc_k(sb, k);
            if (PAIR) umma_bf16_pair(tmem + acc * BN, da, db, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            else umma_bf16(tmem + acc * BN, da, db, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
          if (PAIR) {
            umma_commit_pair(empty + p.stage);
            if (kb == kb_n - 1) umma_commit_pair(acc_full + acc);
          } else {
            umma_commit(empty + p.stage);
            if (kb == kb_n - 1) umma_commit(acc_full + acc);
          }
        }
        __syncwarp();
        p.advance(NSTG);
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue =====
    // warp w may touch TMEM lanes 32 (w % 4) ..: two warps per lane quarter, each takes half of the 256 columns
    const int q = warp & 3, half = (warp - 4) >> 2;
    unsigned acc_phase[2] = {0u, 0u};
    int it = 0;
    float sc = 0.f;
    for (int item = cta_id; item < n_tiles; item += n_ctas, ++it) {
      const int acc = it & 1;
      int tile, split_e, nhalf;
      decode(item, tile, split_e, nhalf);
      const int m0 = (tile % tiles_m) * TILE_M + (int)pair_rank * BM, n0 = (tile / tiles_m) * BN + (nhalf > 0 ? nhalf * tail_w : 0);
      const int cols = nhalf < 0 ? BN : tail_w;          // accumulator columns of this item; each warp pair splits them
      const int r = m0 + 32 * q + lane;
      const bool rv = r < M;
      const bool skipped = epi.skip && *epi.skip != 0;
      const bool pres = ((epi.present && rv) ? epi.present[r] != 0 : true) && !skipped;
      const bool pair_ok = (r | 1) < M;                // rows r and r ^ 1 both exist: packed transposed stores
      if (epi.mode >= EPI_NONE) {
        mbar_wait(acc_full + acc, acc_phase[acc]);
        acc_phase[acc] ^= 1u;
        tc_fence_after();
        if (epi.mode == EPI_NONE + 1) {            // ldonly: the TMEM reads of a real epilogue, nothing else
          float keep = 0.f;
          for (int c0 = half * (cols / 2); c0 < (half + 1) * (cols / 2); c0 += 16) {
            float v[16];
            tmem_ld16(tmem + ((unsigned)(32 * q) << 16) + acc * BN + c0, v);
            keep += v[0];
          }
          if (keep == 1.2345e-30f && epi.sc_sum) *epi.sc_sum = keep;
        } else if (epi.mode == EPI_NONE + 2) {     // aluonly: about the arithmetic of a real epilogue on dummy registers
          float v[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = (float)(lane + i);
          for (int c0 = half * (cols / 2); c0 < (half + 1) * (cols / 2); c0 += 16) {
#pragma unroll
            for (int rep = 0; rep < 3; ++rep)
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = fmaf(v[i], 1.0001f, (float)c0);
          }
          float keep = 0.f;
#pragma unroll
          for (int i = 0; i < 16; ++i) keep += v[i];
          if (keep == 1.2345e-30f && epi.sc_sum) *epi.sc_sum = keep;
        }
        tc_fence_before();
        if (PAIR) mbar_arrive_leader(acc_empty + acc);
        else mbar_arrive(acc_empty + acc);
        continue;
      }
      const int c_beg = half * (cols / 2), c_end = (half + 1) * (cols / 2);
      // fast paths of the row-wise operands the epilogue reads (16 columns = two 16-byte loads per iteration)
      const bool aux_fast = epi.aux && rv && (epi.ld_aux & 7) == 0 && (reinterpret_cast<size_t>(epi.aux) & 15) == 0 && (n0 & 7) == 0;
      const bool old_fast = epi.out_f32 && rv && epi.mode == EPI_ACCUM_F32 && epi.accumulate && splits == 1 && (epi.ld_f32 & 3) == 0 &&
                            (reinterpret_cast<size_t>(epi.out_f32) & 15) == 0 && (n0 & 3) == 0;
      uint4 aq0 = make_uint4(0, 0, 0, 0), aq1 = aq0;                       // aux of the iteration about to run, fetched one ahead
      float4 oq[4] = {make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f),
                      make_float4(0.f, 0.f, 0.f, 0.f)};                   // old fp32 output, likewise
      auto fetch = [&](int c0) {
        const int n = n0 + c0;
        if (n + 16 > N) return;
        if (aux_fast) {
          const uint4* ap = reinterpret_cast<const uint4*>(epi.aux + (long long)r * epi.ld_aux + n);
          aq0 = ap[0]; aq1 = ap[1];
        }
        if (old_fast) {
          const float4* op4 = reinterpret_cast<const float4*>(epi.out_f32 + (long long)r * epi.ld_f32 + n);
#pragma unroll
          for (int i = 0; i < 4; ++i) oq[i] = op4[i];
        }
      };
      // While the tensor cores still work on this tile: the tile's bias slice -> shared memory, the lines of the row-wise
      // operands -> L2 (they were written a whole pass ago and sit in DRAM), and the first iteration's operands -> registers
      if (epi.bias) {
        const int e_tid = (int)threadIdx.x - 128, nb = n0 + e_tid;
        bias_s[(it & 1) * BN + e_tid] = nb < N ? __ldg(epi.bias + nb) : 0.f;
      }
      if (rv && n0 + c_beg < N) {
        const int n_pf = min(c_end, N - n0) - c_beg;
        auto l2_prefetch = [&](const char* ptr, int bytes) {
          for (int o = 0; o < bytes; o += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr + o));
          asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr + bytes - 1));
        };
        if (epi.aux) l2_prefetch(reinterpret_cast<const char*>(epi.aux + (long long)r * epi.ld_aux + n0 + c_beg), n_pf * 2);
        if (epi.aux2) l2_prefetch(reinterpret_cast<const char*>(epi.aux2 + (long long)r * epi.ld_aux2 + n0 + c_beg), n_pf * 2);
        if (epi.out_f32 && splits == 1 && (epi.accumulate || epi.mode == EPI_CARRY))
          l2_prefetch(reinterpret_cast<const char*>(epi.out_f32 + (long long)r * epi.ld_f32 + n0 + c_beg), n_pf * 4);
      }
      if (epi.bias) asm volatile("bar.sync 1, 256;" ::: "memory");        // the 8 epilogue warps: bias slice complete
      const float* bias_t = bias_s + (it & 1) * BN;
      {
        auto al16 = [](const void* p_) { return (reinterpret_cast<size_t>(p_) & 15) == 0; };
        const bool lean = m0 + BM <= M && n0 + c_end <= N && !epi.out_t && (!epi.out || ((epi.ld_out & 7) == 0 && al16(epi.out))) &&
                          (!epi.out_f32 || ((epi.ld_f32 & 3) == 0 && al16(epi.out_f32))) &&
                          (!epi.aux || ((epi.ld_aux & 7) == 0 && al16(epi.aux))) && (!epi.aux2 || ((epi.ld_aux2 & 7) == 0 && al16(epi.aux2))) &&
                          (epi.mode != EPI_CARRY || (epi.out_f32 && epi.aux && epi.aux2)) &&
                          ((epi.mode != EPI_SELECT && epi.mode != EPI_DACT) || epi.aux) &&
                          // one output: bf16 for activations / layer gradients, fp32 for accumulations and the carry
                          (epi.mode == EPI_CARRY || (epi.mode == EPI_ACCUM_F32 ? (epi.out != nullptr) != (epi.out_f32 != nullptr)
                                                                               : (epi.out && !epi.out_f32))) &&
                          ((c_end - c_beg) & 31) == 0;
        if (lean) {            // uniform over the CTA
          const unsigned taddr = tmem + ((unsigned)(32 * q) << 16) + acc * BN;
          const unsigned parity = acc_phase[acc];
          acc_phase[acc] ^= 1u;
          char* stg = epi_stage + (warp - 4) * (32 * 64);
          switch (epi.mode) {
            case EPI_STORE: epi_lean<EPI_STORE>(epi, taddr, c_beg, c_end, r, n0, pres, bias_t, splits > 1, acc_full + acc, parity, stg, sc); break;
            case EPI_SELECT: epi_lean<EPI_SELECT>(epi, taddr, c_beg, c_end, r, n0, pres, bias_t, splits > 1, acc_full + acc, parity, stg, sc); break;
            case EPI_DACT: epi_lean<EPI_DACT>(epi, taddr, c_beg, c_end, r, n0, pres, bias_t, splits > 1, acc_full + acc, parity, stg, sc); break;
            case EPI_ACCUM_F32: epi_lean<EPI_ACCUM_F32>(epi, taddr, c_beg, c_end, r, n0, pres, bias_t, splits > 1, acc_full + acc, parity, stg, sc); break;
            default: epi_lean<EPI_CARRY>(epi, taddr, c_beg, c_end, r, n0, pres, bias_t, splits > 1, acc_full + acc, parity, stg, sc); break;
          }
          tc_fence_before();
          if (PAIR) mbar_arrive_leader(acc_empty + acc);
          else mbar_arrive(acc_empty + acc);
          continue;
        }
      }
      fetch(c_beg);
      mbar_wait(acc_full + acc, acc_phase[acc]);
      acc_phase[acc] ^= 1u;
      tc_fence_after();
      for (int c0 = c_beg; c0 < c_end; c0 += 16) {
        const uint4 a0 = aq0, a1 = aq1;
        float4 old4[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) old4[i] = oq[i];
        if (c0 + 16 < c_end) fetch(c0 + 16);
        float v[16];
        tmem_ld16(tmem + ((unsigned)(32 * q) << 16) + acc * BN + c0, v);     // warp-collective: outside every branch
        const int n = n0 + c0;
        if (n >= N) continue;                                              // uniform
        const bool full16 = n + 16 <= N;
        float aux[16];
        if (epi.aux && rv) {
          const __nv_bfloat16* ap = epi.aux + (long long)r * epi.ld_aux + n;
          if (full16 && aux_fast) {
            const __nv_bfloat162* h0 = reinterpret_cast<const __nv_bfloat162*>(&a0);
            const __nv_bfloat162* h1 = reinterpret_cast<const __nv_bfloat162*>(&a1);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float2 f0 = __bfloat1622float2(h0[i]), f1 = __bfloat1622float2(h1[i]);
              aux[2 * i] = f0.x; aux[2 * i + 1] = f0.y; aux[8 + 2 * i] = f1.x; aux[8 + 2 * i + 1] = f1.y;
            }
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) aux[i] = n + i < N ? __bfloat162float(ap[i]) : 0.f;
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) aux[i] = 0.f;
        }
        if (epi.mode == EPI_STORE || epi.mode == EPI_SELECT) {
          if (epi.bias) {
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
              const float4 b = *reinterpret_cast<const float4*>(bias_t + c0 + i);
              v[i] += b.x; v[i + 1] += b.y; v[i + 2] += b.z; v[i + 3] += b.w;
            }
          }
          if (epi.act == MMN_ACT_RELU) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
          } else if (epi.act == MMN_ACT_SIGMOID) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = __fdividef(1.f, 1.f + __expf(-v[i]));
          } else if (epi.act == MMN_ACT_TANH) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = tanhf(v[i]);
          }
          if (epi.mode == EPI_SELECT) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              // what the next layer reads is the bf16-rounded state: measure the change on that
              const float nw = pres ? __bfloat162float(__float2bfloat16(v[i])) : aux[i];
              const float df = (rv && n + i < N) ? nw - aux[i] : 0.f;
              sc = fmaf(df, df, sc);
              v[i] = nw;
            }
          }
        } else if (epi.mode == EPI_DACT) {
          if (epi.act == MMN_ACT_RELU) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = aux[i] > 0.f ? v[i] : 0.f;
          } else if (epi.act == MMN_ACT_SIGMOID) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] *= aux[i] * (1.f - aux[i]);
          } else if (epi.act == MMN_ACT_TANH) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] *= 1.f - aux[i] * aux[i];
          }
        } else if (epi.mode == EPI_CARRY && epi.drop_thr) {
#pragma unroll
          for (int i = 0; i < 16; ++i)
            v[i] = mmn_dropout_keep(epi.drop_seed, epi.drop_row_base + (unsigned)r, epi.drop_col_base + (unsigned)(n + i), epi.drop_thr)
                       ? v[i] * epi.scale : 0.f;
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] *= epi.scale;
        }
        if (rv) {
        if (epi.out_f32) {
          float* op = epi.out_f32 + (long long)r * epi.ld_f32 + n;
          if (epi.mode == EPI_CARRY) {
            const __nv_bfloat16* bp = epi.aux2 + (long long)r * epi.ld_aux2 + n;
            float b[16];
            if (full16 && ((epi.ld_aux2 & 7) == 0)) {
              const uint4 b0 = *reinterpret_cast<const uint4*>(bp), b1 = *reinterpret_cast<const uint4*>(bp + 8);
              const __nv_bfloat162* h0 = reinterpret_cast<const __nv_bfloat162*>(&b0);
              const __nv_bfloat162* h1 = reinterpret_cast<const __nv_bfloat162*>(&b1);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float2 f0 = __bfloat1622float2(h0[i]), f1 = __bfloat1622float2(h1[i]);
                b[2 * i] = f0.x; b[2 * i + 1] = f0.y; b[8 + 2 * i] = f1.x; b[8 + 2 * i + 1] = f1.y;
              }
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i) b[i] = n + i < N ? __bfloat162float(bp[i]) : 0.f;
            }
            if (full16 && ((epi.ld_f32 & 3) == 0)) {
#pragma unroll
              for (int i = 0; i < 16; i += 4) {
                float4 g = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                if (!pres) g = *reinterpret_cast<const float4*>(op + i);
                g.x -= epi.c_sc * (aux[i] - b[i]); g.y -= epi.c_sc * (aux[i + 1] - b[i + 1]);
                g.z -= epi.c_sc * (aux[i + 2] - b[i + 2]); g.w -= epi.c_sc * (aux[i + 3] - b[i + 3]);
                *reinterpret_cast<float4*>(op + i) = g;
              }
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i)
                if (n + i < N) op[i] = (pres ? v[i] : op[i]) - epi.c_sc * (aux[i] - b[i]);
            }
          } else if (splits > 1) {
            if (full16 && ((epi.ld_f32 & 3) == 0) && ((reinterpret_cast<size_t>(epi.out_f32) & 15) == 0)) {
#pragma unroll
              for (int i = 0; i < 16; i += 4) atomicAdd(reinterpret_cast<float4*>(op + i), make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]));
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i)
                if (n + i < N) atomicAdd(op + i, v[i]);
            }
          } else if (full16 && ((epi.ld_f32 & 3) == 0)) {
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
              float4 o = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
              if (epi.accumulate) {
                const float4 old = old_fast ? old4[i / 4] : *reinterpret_cast<const float4*>(op + i);
                o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
              }
              *reinterpret_cast<float4*>(op + i) = o;
            }
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (n + i < N) op[i] = epi.accumulate ? op[i] + v[i] : v[i];
          }
        }
        if (epi.out) {
          __nv_bfloat16* op = epi.out + (long long)r * epi.ld_out + n;
          if (full16 && ((epi.ld_out & 7) == 0)) {
            uint4 o[2];
            __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(o);
#pragma unroll
            for (int i = 0; i < 8; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
            *reinterpret_cast<uint4*>(op) = o[0];
            *reinterpret_cast<uint4*>(op + 8) = o[1];
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (n + i < N) op[i] = __float2bfloat16(v[i]);
          }
        }
        }
        if (epi.out_t) {
          // transposed copy: lanes r, r ^ 1 exchange so that each writes one 4-byte pair {row even, row odd} of a column
          // (even lanes the even columns, odd lanes the odd ones) instead of two 2-byte stores
#pragma unroll
          for (int i = 0; i < 16; i += 2) {
            const float o0 = __shfl_xor_sync(0xffffffffu, v[i], 1), o1 = __shfl_xor_sync(0xffffffffu, v[i + 1], 1);
            if (pair_ok) {
              const int col = n + i + (lane & 1);
              const __nv_bfloat162 pr = (lane & 1) ? __floats2bfloat162_rn(o1, v[i + 1]) : __floats2bfloat162_rn(v[i], o0);
              if (col < N) *reinterpret_cast<__nv_bfloat162*>(epi.out_t + (long long)col * epi.ld_out_t + (r & ~1)) = pr;
            } else if (rv) {
              if (n + i < N) epi.out_t[(long long)(n + i) * epi.ld_out_t + r] = __float2bfloat16(v[i]);
              if (n + i + 1 < N) epi.out_t[(long long)(n + i + 1) * epi.ld_out_t + r] = __float2bfloat16(v[i + 1]);
            }
          }
        }
        __syncwarp();                 // the next tcgen05.ld is warp-collective
      }
      tc_fence_before();
      if (PAIR) mbar_arrive_leader(acc_empty + acc);       // the leader's issuer waits for both CTAs' epilogues
      else mbar_arrive(acc_empty + acc);
    }
    if (epi.sc_sum) {
#pragma unroll
      for (int o = 16; o; o >>= 1) sc += __shfl_xor_sync(0xffffffffu, sc, o);
      if (lane == 0 && sc != 0.f) atomicAdd(epi.sc_sum, sc);
    }
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all();            // no CTA leaves while its peer may still signal its barriers / read its smem
  else __syncthreads();
  if (warp == 2) {
    if (PAIR) tmem_dealloc_pair(tmem, 512);
    else tmem_dealloc(tmem, 512);
  }
}

}  // namespace wide
}  // namespace mmn
