// mmn_nb_prims.cuh — warp-level primitives of the bf16 tile kernel (mmn_nb.cuh): bf16 packing, ldmatrix, movmatrix,
// mma.sync.m16n8k16 (bf16 x bf16 -> fp32), vector reds.  Under -DMMN_EMU (tests/emu, CPU only) each instruction is replaced
// by a lane-exact functional model, so the fragment bookkeeping of the kernel is exercised on the CPU; the real instructions
// are validated by the -m gpu tests.
//
// Fragment conventions (PTX ISA, mma.m16n8k16 with .bf16 operands), lane = 4 g + t:
//   A (16 x 16, row):  a0 = (row g,     k 2t, 2t+1)   a1 = (row g + 8, k 2t, 2t+1)
//                      a2 = (row g,     k 2t+8, +9)   a3 = (row g + 8, k 2t+8, +9)
//   B (16 x 8,  col):  b0 = (k 2t, 2t+1; col g)       b1 = (k 2t+8, 2t+9; col g)
//   C (16 x 8):        c0, c1 = (row g; cols 2t, 2t+1)    c2, c3 = (row g + 8; cols 2t, 2t+1)
//   ldmatrix (8 x 8 b16 per matrix, rows addressed by lanes 8 i .. 8 i + 7 for matrix i): lane T receives
//       (row T / 4, cols 2 (T % 4), +1), with .trans (rows 2 (T % 4), +1; col T / 4)
//   movmatrix.trans: the same exchange between registers.
#pragma once

#include "mmn_common.cuh"

namespace mmn {
namespace nb {

// round-to-nearest-even fp32 -> bf16 (NaN stays NaN), two values packed: lo in bits [0,16), hi in [16,32)
__host__ __device__ inline unsigned bf16_bits(float v) {
  unsigned u;
#if defined(__CUDA_ARCH__)
  u = __float_as_uint(v);
#else
  memcpy(&u, &v, 4);
#endif
  if ((u & 0x7F800000u) == 0x7F800000u && (u & 0x007FFFFFu)) return (u >> 16) | 0x40u;   // NaN
  return (u + 0x7FFFu + ((u >> 16) & 1u)) >> 16;
}

#ifndef MMN_EMU
__device__ __forceinline__ unsigned pack_bf16(float lo, float hi) {
  unsigned d;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
__device__ __forceinline__ float bf16_lo(unsigned p) { return __uint_as_float(p << 16); }
__device__ __forceinline__ float bf16_hi(unsigned p) { return __uint_as_float(p & 0xFFFF0000u); }
__device__ __forceinline__ unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void ldsm_x4(unsigned (&r)[4], unsigned saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(saddr) : "memory");
}
__device__ __forceinline__ void ldsm_x4_t(unsigned (&r)[4], unsigned saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(saddr) : "memory");
}
__device__ __forceinline__ unsigned movm_t(unsigned v) {
  unsigned d;
  asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(d) : "r"(v));
  return d;
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const unsigned (&a)[4], unsigned b0, unsigned b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void red_add_v2(float* p, float a, float b) {
  asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void red_add(float* p, float a) { asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(a) : "memory"); }
// barrier over the threads of one group (ids 1, 2: id 0 is __syncthreads)
__device__ __forceinline__ void group_bar(int gi, int n_threads) { asm volatile("bar.sync %0, %1;" ::"r"(gi + 1), "r"(n_threads) : "memory"); }
#else
static inline unsigned pack_bf16(float lo, float hi) { return bf16_bits(lo) | (bf16_bits(hi) << 16); }
static inline float bf16_lo(unsigned p) { return __uint_as_float(p << 16); }
static inline float bf16_hi(unsigned p) { return __uint_as_float(p & 0xFFFF0000u); }
static inline unsigned smem_addr(const void* p) { return (unsigned)((const char*)p - emu::st().dyn_smem); }
namespace emuprim {
inline const char* sptr(unsigned saddr) { return emu::st().dyn_smem + saddr; }
inline void ldsm(unsigned (&r)[4], unsigned saddr, bool trans) {
  const unsigned T = threadIdx.x & 31;
  uint64_t w = saddr;
  emu::WarpExchange::publish(&w, 1);
  for (int i = 0; i < 4; ++i) {
    if (!trans) {
      const char* row = sptr((unsigned)emu::WarpExchange::peer(8 * i + T / 4, 0));
      memcpy(&r[i], row + 4 * (T % 4), 4);
    } else {
      unsigned short lo, hi;
      memcpy(&lo, sptr((unsigned)emu::WarpExchange::peer(8 * i + 2 * (T % 4), 0)) + 2 * (T / 4), 2);
      memcpy(&hi, sptr((unsigned)emu::WarpExchange::peer(8 * i + 2 * (T % 4) + 1, 0)) + 2 * (T / 4), 2);
      r[i] = (unsigned)lo | ((unsigned)hi << 16);
    }
  }
  emu::WarpExchange::done();
}
inline float bf(unsigned packed, int half) { return __uint_as_float(half ? (packed & 0xFFFF0000u) : (packed << 16)); }
}  // namespace emuprim
static inline void ldsm_x4(unsigned (&r)[4], unsigned saddr) { emuprim::ldsm(r, saddr, false); }
static inline void ldsm_x4_t(unsigned (&r)[4], unsigned saddr) { emuprim::ldsm(r, saddr, true); }
static inline unsigned movm_t(unsigned v) {
  const unsigned T = threadIdx.x & 31;
  uint64_t w = v;
  emu::WarpExchange::publish(&w, 1);
  // element M[r][c] lives in lane 4 r + c / 2, half c % 2; the output lane T holds (M[2 (T % 4)][T / 4], M[2 (T % 4) + 1][T / 4])
  auto elem = [&](unsigned r, unsigned c) {
    const unsigned p = (unsigned)emu::WarpExchange::peer(4 * r + c / 2, 0);
    return (c & 1) ? (p >> 16) : (p & 0xFFFFu);
  };
  const unsigned out = elem(2 * (T % 4), T / 4) | (elem(2 * (T % 4) + 1, T / 4) << 16);
  emu::WarpExchange::done();
  return out;
}
static inline void mma_bf16(float (&c)[4], const unsigned (&a)[4], unsigned b0, unsigned b1) {
  const unsigned T = threadIdx.x & 31, g = T / 4, t = T % 4;
  uint64_t w[3] = {(uint64_t)a[0] | ((uint64_t)a[1] << 32), (uint64_t)a[2] | ((uint64_t)a[3] << 32), (uint64_t)b0 | ((uint64_t)b1 << 32)};
  emu::WarpExchange::publish(w, 3);
  auto A = [&](unsigned row, unsigned k) {      // row < 16, k < 16
    const unsigned lane = 4 * (row & 7) + (k & 7) / 2;
    const uint64_t word = emu::WarpExchange::peer(lane, k >= 8 ? 1 : 0);
    const unsigned reg = (unsigned)(row >= 8 ? (word >> 32) : word);
    return emuprim::bf(reg, k & 1);
  };
  auto B = [&](unsigned k, unsigned col) {      // k < 16, col < 8
    const unsigned lane = 4 * col + (k & 7) / 2;
    const uint64_t word = emu::WarpExchange::peer(lane, 2);
    const unsigned reg = (unsigned)(k >= 8 ? (word >> 32) : word);
    return emuprim::bf(reg, k & 1);
  };
  float out[4];
  for (int i = 0; i < 4; ++i) {
    const unsigned row = g + (i >= 2 ? 8 : 0), col = 2 * t + (i & 1);
    float s = c[i];
    for (unsigned k = 0; k < 16; ++k) s += A(row, k) * B(k, col);
    out[i] = s;
  }
  emu::WarpExchange::done();
  for (int i = 0; i < 4; ++i) c[i] = out[i];
}
static inline void red_add_v2(float* p, float a, float b) { p[0] += a; p[1] += b; }
static inline void red_add(float* p, float a) { *p += a; }
static inline void group_bar(int gi, int n_threads) { emu::named_barrier_id(gi + 1, n_threads); }
#endif

}  // namespace nb
}  // namespace mmn
