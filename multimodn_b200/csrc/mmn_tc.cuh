// mmn_tc.cuh — tcgen05 (5th-gen tensor core) GEMM engine of the fused step, fp32-accurate via 3xTF32.
//
// Same three GEMM shapes as the FMA engine (mmn_kernels.cuh), same staging/epilogue structure, but the
// multiply runs on the tensor cores with the accumulator in TMEM:
//
//   * every operand chunk is staged into shared memory in the UMMA canonical SWIZZLE_128B layout
//     ([rows][32 fp32] = 128-byte rows, 8-row / 1024-byte atoms, 16-byte chunk index XOR (row & 7)),
//     split into hi = tf32-representable part and lo = v - hi, so that
//         a*b ~= hi_a*hi_b + lo_a*hi_b + hi_a*lo_b          (error ~2^-21 relative)
//     and every tcgen05.mma.kind::tf32 consumes operands it represents exactly (hi) or to 2^-11 (lo);
//   * ONE [rows x 32] image serves as a K-major operand (rows = M or N, contraction along the 32
//     columns: forward activations, weights W[n][k]) or as an MN-major operand (contraction along
//     the rows: weights for the data gradient, both operands of the weight gradient) — no transposes;
//   * one elected thread issues the MMAs (M = 128 batch rows, N = 32/64, K = 8 per instruction) and
//     commits them to an mbarrier; all 8 warps then read their 32-lane quarter of the accumulator
//     with tcgen05.ld and run the epilogue (bias, activation, masks, gradient reds).
//
// Under -DMMN_EMU (tests/emu, CPU only) the tcgen05 / mbarrier instructions are replaced by a
// functional model with the same descriptor arithmetic, so the control flow and the layout code are
// exercised on the CPU; the real instruction semantics are validated by the -m gpu tests.
#pragma once

#include "mmn_common.cuh"

namespace mmn {

constexpr int kTcStageXB = 2 * 2 * 128 * 32;   // floats: 2 buffers x (hi, lo) x [128 x 32]      = 64 KB
constexpr int kTcStageWB = 2 * 2 * 64 * 32;    // floats: 2 buffers x (hi, lo) x [64 x 32]       = 32 KB
constexpr int kTcTmemCols = 64;

// float offset of element (row, col<32) inside a SWIZZLE_128B image
__device__ __forceinline__ int sw128(int row, int col) {
  return ((row >> 3) << 8) + ((row & 7) << 5) + ((((col >> 2) ^ (row & 7)) << 2) | (col & 3));
}
__device__ __forceinline__ float tf32_hi(float v) { return __uint_as_float(__float_as_uint(v) & 0xFFFFE000u); }

// ---- descriptors (cute/arch/mma_sm100_desc.hpp: UMMA::SmemDescriptor / InstrDescriptor bit layout) ----
// smem descriptor: [0,14) start>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version=1 | [61,64) layout (2 = SWIZZLE_128B)
__device__ __forceinline__ unsigned long long umma_smem_desc(unsigned saddr_bytes, unsigned lbo_bytes, unsigned sbo_bytes,
                                                             unsigned layout_type) {
  return (unsigned long long)((saddr_bytes >> 4) & 0x3FFFu) | ((unsigned long long)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((unsigned long long)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46) | ((unsigned long long)layout_type << 61);
}
// K-major operand image ([rows][32 fp32], SWIZZLE_128B): rows = M or N, 8-row atoms 1024 B apart;
// k_slice selects 8 of the 32 columns (32 bytes inside the swizzled row)
__device__ __forceinline__ unsigned long long umma_desc_k(unsigned img_bytes, int k_slice) {
  return umma_smem_desc(img_bytes + 32u * k_slice, 0, 1024, 2);
}
// MN-major operand image ([K rows][32 fp32 along MN], SWIZZLE_128B_BASE32B — the only MN-major layout
// for 32-bit operands): 4-row atoms 512 B apart, successive 32-element MN groups group_bytes apart;
// k_slice selects 8 rows
__device__ __forceinline__ unsigned long long umma_desc_mn(unsigned img_bytes, int k_slice, unsigned group_bytes) {
  return umma_smem_desc(img_bytes + 1024u * k_slice, group_bytes, 512, 1);
}
// instruction descriptor, kind::tf32, fp32 accumulate: c_format=1 @4, a/b_format=2 (TF32) @7/@10,
// a_major @15, b_major @16 (0 = K-major, 1 = MN-major), N>>3 @17, M>>4 @24
__device__ __forceinline__ unsigned umma_idesc_tf32(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)a_mn_major << 15) | ((unsigned)b_mn_major << 16) |
         ((unsigned)(N >> 3) << 17) | ((unsigned)(M >> 4) << 24);
}

struct TcState {        // identical in every thread of the CTA
  unsigned tmem;        // TMEM base address of the accumulator columns
  unsigned pending[2];  // an uncollected tcgen05.commit is outstanding on mbarrier b
  unsigned parity[2];
};

#ifndef MMN_EMU
// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  // bounded spin: a protocol bug traps instead of hanging the GPU
  for (unsigned spins = 0;; ++spins) {
    unsigned done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (done) return;
    if (spins > (1u << 24)) __trap();
  }
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(unsigned* slot, unsigned ncols) {   // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(unsigned taddr, unsigned ncols) {  // the same warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_tf32(unsigned tmem_d, unsigned long long adesc, unsigned long long bdesc,
                                          unsigned idesc, unsigned accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc),
      "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(unsigned long long* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// this thread's TMEM lane (row), 16 consecutive fp32 columns starting at taddr's column
__device__ __forceinline__ void tmem_ld16(unsigned taddr, float (&v)[16]) {
  unsigned r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
#else
// ------------------------------------------------------------------------------------------------
// functional model for the CPU emulator build (tests/emu)
// ------------------------------------------------------------------------------------------------
namespace tcemu {
inline float (&tmem())[128][512] { static float t[128][512]; return t; }
inline float trunc_tf32(float v) { return __uint_as_float(__float_as_uint(v) & 0xFFFFE000u); }
inline float operand(unsigned long long desc, int mn_major, int mn, int k) {
  const unsigned start = (unsigned)(desc & 0x3FFF) << 4, lbo = (unsigned)((desc >> 16) & 0x3FFF) << 4,
                 sbo = (unsigned)((desc >> 32) & 0x3FFF) << 4;
  unsigned addr;
  const unsigned layout = (unsigned)(desc >> 61);
  if (!mn_major) {                       // K-major, SWIZZLE_128B: 8-row atoms along MN, Swizzle<3,4,3>
    if (layout != 2) abort();
    addr = start + (mn >> 3) * sbo + (mn & 7) * 128 + k * 4;
    addr ^= ((addr >> 7) & 7u) << 4;
  } else {                               // MN-major, SWIZZLE_128B_BASE32B: 4-row atoms along K, Swizzle<2,5,2>
    if (layout != 1) abort();
    addr = start + (mn >> 5) * lbo + (k >> 2) * sbo + (k & 3) * 128 + (mn & 31) * 4;
    addr ^= ((addr >> 7) & 3u) << 5;
  }
  float v; memcpy(&v, emu::st().dyn_smem + addr, 4);
  return trunc_tf32(v);
}
}  // namespace tcemu
static inline unsigned smem_u32(const void* p) { return (unsigned)((const char*)p - emu::st().dyn_smem); }
// mbarrier model: the word counts completed phases; a phase with parity P is complete once the count's
// low bit differs from P (tcgen05.commit completes its phase immediately: the model's MMAs are synchronous)
static inline void mbar_init(unsigned long long* bar, unsigned) { *bar = 0; }
static inline void mbar_fence_init() {}
static inline void mbar_wait(unsigned long long* bar, unsigned parity) {
  while (((*bar) & 1ull) == (unsigned long long)parity) emu::yield();
}
static inline void fence_proxy_async() {}
static inline void tc_fence_before() {}
static inline void tc_fence_after() {}
static inline void tmem_alloc(unsigned* slot, unsigned) { *slot = 0; }
static inline void tmem_dealloc(unsigned, unsigned) {}
static inline void umma_tf32(unsigned tmem_d, unsigned long long adesc, unsigned long long bdesc, unsigned idesc,
                             unsigned accumulate) {
  const int M = ((idesc >> 24) & 31) << 4, N = ((idesc >> 17) & 63) << 3;
  const int amn = (idesc >> 15) & 1, bmn = (idesc >> 16) & 1;
  const int col0 = tmem_d & 0xFFFF;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      float s = accumulate ? tcemu::tmem()[m][col0 + n] : 0.f;
      for (int k = 0; k < 8; ++k) s += tcemu::operand(adesc, amn, m, k) * tcemu::operand(bdesc, bmn, n, k);
      tcemu::tmem()[m][col0 + n] = s;
    }
}
static inline void umma_commit(unsigned long long* bar) { *bar += 1; }
static inline void tmem_ld16(unsigned taddr, float (&v)[16]) {
  const int lane = (taddr >> 16) + (threadIdx.x & 31), col = taddr & 0xFFFF;
  for (int i = 0; i < 16; ++i) v[i] = tcemu::tmem()[lane][col + i];
}
#endif

}  // namespace mmn

namespace mmn {

// write one float4 (row r, columns 4*c4 .. 4*c4+3 of a 32-wide chunk) into the hi / lo images;
// MN = false: K-major image (SWIZZLE_128B), MN = true: MN-major image (SWIZZLE_128B_BASE32B)
template <bool MN>
__device__ __forceinline__ void tc_store_quad(float* hi_img, float* lo_img, int r, int c4, float4 v) {
  const int o = MN ? ((r >> 2) << 7) + ((r & 3) << 5) + ((((c4 >> 1) ^ (r & 3)) << 3) | ((c4 & 1) << 2))
                   : ((r >> 3) << 8) + ((r & 7) << 5) + ((c4 ^ (r & 7)) << 2);
  float4 h = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
  *reinterpret_cast<float4*>(hi_img + o) = h;
  *reinterpret_cast<float4*>(lo_img + o) = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
}

// ------------------------------------------------------------------------------------------------
// self-test of the three operand configurations the engine uses (validated on the GPU by
// tests/test_gpu_tc_selftest.py): mode 0  D[r][n] = sum_k A[r][k] B[n][k]   (A, B K-major)
//                                 mode 1  D[r][j] = sum_n A[r][n] W[n][j]   (A K-major, W MN-major)
//                                 mode 2  D[n][k] = sum_r Z[r][n] X[r][k]   (Z, X MN-major, K = 128 rows)
// A: [128 x 32], B: [N x 32], W: [32 x N], Z: [128 x 64], X: [128 x 32]; out: [128 x N] (mode 2: N = 32,
// rows n < 64 meaningful).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 1) mmn_tc_selftest_kernel(int mode, int N, const float* __restrict__ A,
                                                                const float* __restrict__ B, float* __restrict__ out) {
  MMN_DYN_SMEM(raw);
  char* base = raw + ((1024 - (smem_u32(raw) & 1023)) & 1023);
  float* XB = reinterpret_cast<float*>(base);                 // 64 KB
  float* WB = XB + 16384;                                     // 32 KB
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(WB + 8192);
  unsigned* slot = reinterpret_cast<unsigned*>(bar + 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) { mbar_init(bar, 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc(slot, kTcTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const unsigned tmem = *slot;
  float *a_hi, *a_lo, *b_hi, *b_lo;
  if (mode == 2) { a_hi = XB; a_lo = XB + 8192; b_hi = WB; b_lo = WB + 4096; }
  else { a_hi = XB; a_lo = XB + 4096; b_hi = WB; b_lo = WB + 2048; }
  // stage operands
  if (mode == 2) {
    for (int idx = tid; idx < 128 * 16; idx += 256) {      // Z: two 32-column groups
      const int r = idx >> 4, c4 = idx & 15;
      const float4 v = *reinterpret_cast<const float4*>(A + r * 64 + 4 * c4);
      tc_store_quad<true>(a_hi + (c4 >> 3) * 4096, a_lo + (c4 >> 3) * 4096, r, c4 & 7, v);
    }
    for (int idx = tid; idx < 128 * 8; idx += 256) {
      const int r = idx >> 3, c4 = idx & 7;
      tc_store_quad<true>(b_hi, b_lo, r, c4, *reinterpret_cast<const float4*>(B + r * 32 + 4 * c4));
    }
  } else {
    for (int idx = tid; idx < 128 * 8; idx += 256) {
      const int r = idx >> 3, c4 = idx & 7;
      tc_store_quad<false>(a_hi, a_lo, r, c4, *reinterpret_cast<const float4*>(A + r * 32 + 4 * c4));
    }
    if (mode == 0) {
      for (int idx = tid; idx < N * 8; idx += 256) {
        const int r = idx >> 3, c4 = idx & 7;
        tc_store_quad<false>(b_hi, b_lo, r, c4, *reinterpret_cast<const float4*>(B + r * 32 + 4 * c4));
      }
    } else {
      const int q = N >> 2;                               // W [32 x N]: groups of 32 columns, 1024 floats each
      for (int idx = tid; idx < 32 * q; idx += 256) {
        const int r = idx / q, c4 = idx - r * q;
        tc_store_quad<true>(b_hi + (c4 >> 3) * 1024, b_lo + (c4 >> 3) * 1024, r, c4 & 7,
                      *reinterpret_cast<const float4*>(B + r * N + 4 * c4));
      }
    }
  }
  fence_proxy_async();
  __syncthreads();
  if (tid == 0) {
    tc_fence_after();
    const unsigned ah = smem_u32(a_hi), al = smem_u32(a_lo), bh = smem_u32(b_hi), bl = smem_u32(b_lo);
    if (mode == 0) {
      const unsigned id = umma_idesc_tf32(128, N, 0, 0);
      for (int j = 0; j < 4; ++j) {
        umma_tf32(tmem, umma_desc_k(al, j), umma_desc_k(bh, j), id, j > 0);
        umma_tf32(tmem, umma_desc_k(ah, j), umma_desc_k(bl, j), id, 1);
        umma_tf32(tmem, umma_desc_k(ah, j), umma_desc_k(bh, j), id, 1);
      }
    } else if (mode == 1) {
      const unsigned id = umma_idesc_tf32(128, N, 0, 1);
      for (int j = 0; j < 4; ++j) {
        umma_tf32(tmem, umma_desc_k(al, j), umma_desc_mn(bh, j, 4096), id, j > 0);
        umma_tf32(tmem, umma_desc_k(ah, j), umma_desc_mn(bl, j, 4096), id, 1);
        umma_tf32(tmem, umma_desc_k(ah, j), umma_desc_mn(bh, j, 4096), id, 1);
      }
    } else {
      const unsigned id = umma_idesc_tf32(128, 32, 1, 1);
      for (int j = 0; j < 16; ++j) {
        umma_tf32(tmem, umma_desc_mn(al, j, 16384), umma_desc_mn(bh, j, 16384), id, j > 0);
        umma_tf32(tmem, umma_desc_mn(ah, j, 16384), umma_desc_mn(bl, j, 16384), id, 1);
        umma_tf32(tmem, umma_desc_mn(ah, j, 16384), umma_desc_mn(bh, j, 16384), id, 1);
      }
    }
    umma_commit(bar);
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  const int Nout = mode == 2 ? 32 : N;
  const int q = warp & 3, half = warp >> 2;
  for (int c0 = half * 16; c0 < Nout; c0 += 32) {
    float v[16];
    tmem_ld16(tmem + ((unsigned)(q * 32) << 16) + c0, v);
    for (int i = 0; i < 16; ++i) out[(q * 32 + lane) * Nout + c0 + i] = v[i];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, kTcTmemCols);
}

}  // namespace mmn
