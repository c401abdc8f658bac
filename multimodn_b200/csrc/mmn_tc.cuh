// mmn_tc.cuh — tcgen05 (5th-gen tensor core) building blocks for fp32-accurate GEMMs via 3xTF32: PTX wrappers (mbarrier,
// tcgen05.alloc / mma / ld / st / commit), UMMA shared-memory and instruction descriptors, the SWIZZLE_128B (K-major) and
// SWIZZLE_128B_BASE32B (MN-major) image writers with the hi / lo split
//         a*b ~= hi_a*hi_b + lo_a*hi_b + hi_a*lo_b          (error ~2^-21 relative),
// and a one-tile self-test kernel for every operand configuration.  Users: the TMEM-resident forward kernel (mmn_tc2.cuh)
// and, for the PTX wrappers, the wide regime (mmn_wide.cuh).
//
// Under -DMMN_EMU (tests/emu, CPU only) the tcgen05 / mbarrier instructions are replaced by a functional model with the
// same descriptor arithmetic, so the control flow and the layout code are exercised on the CPU; the real instruction
// semantics are validated by the -m gpu tests (tests/test_gpu_tc_selftest.py checks the model against the hardware).
#pragma once

#include "mmn_kernels.cuh"

#include "mmn_common.cuh"

namespace mmn {

constexpr int kTcStageXB = 2 * 2 * 128 * 32;   // floats: 2 buffers x (hi, lo) x [128 x 32]      = 64 KB
constexpr int kTcStageWB = 2 * 2 * 64 * 32;    // floats: 2 buffers x (hi, lo) x [64 x 32]       = 32 KB
constexpr int kTcTmemCols = 128;

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

struct TcState {        // identical in every worker thread of the CTA
  unsigned tmem;        // TMEM base address of the accumulator columns
  unsigned seq;         // chunks posted so far; chunk i uses staging slot / mbarrier pair i & 1
  unsigned pending[2];  // an uncollected tcgen05.commit is outstanding on done[slot]
  unsigned parity[2];
  long long t[16];      // debug cycle counters (thread 0): 0 wait-done, 1 nt, 2 nn, 3 tn, 4 nt-epi, 5 nn-epi, 6 tn-epi, 7 stage
};
// One staged chunk handed to the MMA-issuing warp: nj k-slices, three MMAs each (lo*hi, hi*lo, hi*hi).
// Only an opcode crosses shared memory; the issuer rebuilds every MMA operand from warp-uniform values
// (staging-buffer base addresses, the slot, immediates) so that they live in uniform registers.
enum { TC_OP_NT = 0, TC_OP_NN = 1, TC_OP_TN = 2, TC_OP_QUIT = 3 };
struct TcCmd {
  unsigned op;      // [0,2) kind | [2] N == 64 | [3] first (overwrite the accumulator) | [4,9) nj
  unsigned pad[3];
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
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
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
// A operand from TMEM (K-major: lane = row, 8 consecutive columns = one k-slice), B from shared memory
__device__ __forceinline__ void umma_tf32_ts(unsigned tmem_d, unsigned tmem_a, unsigned long long bdesc, unsigned idesc,
                                             unsigned accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}" ::"r"(tmem_d), "r"(tmem_a),
      "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
// this thread's TMEM lane, 16 consecutive fp32 columns
__device__ __forceinline__ void tmem_st16(unsigned taddr, const float (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
      "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
      "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
      "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(unsigned long long* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ unsigned elect_one() {     // one lane of the (converged) warp
  unsigned pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred;
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
__device__ __forceinline__ void tmem_ld8(unsigned taddr, float (&v)[8]) {
  unsigned r[8];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
      : "r"(taddr) : "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
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
// word layout: [0,32) completed phases, [32,48) expected arrivals, [48,64) arrivals of the current phase
static inline void mbar_init(unsigned long long* bar, unsigned count) { *bar = (unsigned long long)count << 32; }
static inline void mbar_arrive(unsigned long long* bar) {
  unsigned long long w = *bar;
  const unsigned expected = (unsigned)((w >> 32) & 0xFFFF), arrived = (unsigned)(w >> 48) + 1;
  if (arrived == expected) w = ((w & 0x0000FFFFFFFFFFFFull) + 1);       // phase complete, arrivals reset
  else w = (w & 0x0000FFFFFFFFFFFFull) | ((unsigned long long)arrived << 48);
  *bar = w;
}
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
static inline void umma_tf32_ts(unsigned tmem_d, unsigned tmem_a, unsigned long long bdesc, unsigned idesc,
                                unsigned accumulate) {
  const int M = ((idesc >> 24) & 31) << 4, N = ((idesc >> 17) & 63) << 3;
  const int bmn = (idesc >> 16) & 1;
  if ((idesc >> 15) & 1) abort();                      // A from TMEM cannot be MN-major
  const int col0 = tmem_d & 0xFFFF, acol = tmem_a & 0xFFFF;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      float s = accumulate ? tcemu::tmem()[m][col0 + n] : 0.f;
      for (int k = 0; k < 8; ++k) s += tcemu::trunc_tf32(tcemu::tmem()[m][acol + k]) * tcemu::operand(bdesc, bmn, n, k);
      tcemu::tmem()[m][col0 + n] = s;
    }
}
static inline void tmem_st16(unsigned taddr, const float (&v)[16]) {
  emu::warp_collective();
  const int lane = (taddr >> 16) + (threadIdx.x & 31), col = taddr & 0xFFFF;
  for (int i = 0; i < 16; ++i) tcemu::tmem()[lane][col + i] = v[i];
}
static inline void tmem_wait_st() {}
static inline void umma_commit(unsigned long long* bar) { mbar_arrive(bar); }
static inline unsigned elect_one() { return (threadIdx.x & 31) == 0; }
static inline void tmem_ld16(unsigned taddr, float (&v)[16]) {
  emu::warp_collective();
  const int lane = (taddr >> 16) + (threadIdx.x & 31), col = taddr & 0xFFFF;
  for (int i = 0; i < 16; ++i) v[i] = tcemu::tmem()[lane][col + i];
}
static inline void tmem_ld8(unsigned taddr, float (&v)[8]) {
  emu::warp_collective();
  const int lane = (taddr >> 16) + (threadIdx.x & 31), col = taddr & 0xFFFF;
  for (int i = 0; i < 8; ++i) v[i] = tcemu::tmem()[lane][col + i];
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

// element-wise variant of tc_store_quad (ragged sources staged one float per thread)
template <bool MN>
__device__ __forceinline__ void tc_store_elem(float* hi_img, float* lo_img, int r, int c, float v) {
  const int c4 = c >> 2;
  const int o = (MN ? ((r >> 2) << 7) + ((r & 3) << 5) + ((((c4 >> 1) ^ (r & 3)) << 3) | ((c4 & 1) << 2))
                    : ((r >> 3) << 8) + ((r & 7) << 5) + ((c4 ^ (r & 7)) << 2)) + (c & 3);
  const float h = tf32_hi(v);
  hi_img[o] = h;
  lo_img[o] = v - h;
}

// ------------------------------------------------------------------------------------------------
// Shared helpers of the tcgen05 3xTF32 kernels: the MMA issue sequence of one staged chunk and the staging of a weight
// block as a (hi, lo) K-major image pair.  (Round 1 also had a complete step engine on shared-memory-staged operands here;
// it lost to the FP32-FMA kernel on every configuration and was removed in round 2 — DESIGN.md section 4.)
// ------------------------------------------------------------------------------------------------
struct TcEngine {
  static constexpr int kWorkers = 256;                    // worker threads of the TMEM-resident kernel (mmn_tc2.cuh)
  static constexpr int NW = kWorkers;
  static constexpr int QW = 512 / NW;                     // float4 per thread of a weight block
  using State = TcState;

  // one chunk: NJ k-slices x (lo*hi, hi*lo, hi*hi); every operand is a function of uniform values
  template <bool AMN, bool BMN, int NJ>
  __device__ static __forceinline__ void issue(unsigned a_hi, unsigned a_lo, unsigned a_grp, unsigned b_hi, unsigned b_lo,
                                               unsigned b_grp, unsigned idesc, unsigned tmem, unsigned nj, bool first,
                                               bool leader) {
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      if (j < (int)nj && leader) {
        const unsigned long long ah = AMN ? umma_desc_mn(a_hi, j, a_grp) : umma_desc_k(a_hi, j);
        const unsigned long long al = AMN ? umma_desc_mn(a_lo, j, a_grp) : umma_desc_k(a_lo, j);
        const unsigned long long bh = BMN ? umma_desc_mn(b_hi, j, b_grp) : umma_desc_k(b_hi, j);
        const unsigned long long bl = BMN ? umma_desc_mn(b_lo, j, b_grp) : umma_desc_k(b_lo, j);
        umma_tf32(tmem, al, bh, idesc, (first && j == 0) ? 0u : 1u);
        umma_tf32(tmem, ah, bl, idesc, 1u);
        umma_tf32(tmem, ah, bh, idesc, 1u);
      }
    }
  }
  // ---- weight block: [nrows <= 64][ncols <= 32] of row-major W -> K-major image pair (rows = n) ----
  __device__ static __forceinline__ void w_load_k(float (&w)[4 * QW], const float* __restrict__ W, int ldw, int row0, int nrows,
                                                  int col0, int ncols, bool vec) {
    const int t = threadIdx.x;
    if (vec) {
      const int c4 = (t & 7) * 4;
#pragma unroll
      for (int i = 0; i < QW; ++i) {
        const int r = (t >> 3) + (NW / 8) * i;
        float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < nrows && c4 < ncols) {
          const float* p = W + (long long)(row0 + r) * ldw + col0 + c4;
          if (c4 + 3 < ncols) q = __ldg(reinterpret_cast<const float4*>(p));
          else { q.x = __ldg(p); if (c4 + 1 < ncols) q.y = __ldg(p + 1); if (c4 + 2 < ncols) q.z = __ldg(p + 2); }
        }
        w[4 * i] = q.x; w[4 * i + 1] = q.y; w[4 * i + 2] = q.z; w[4 * i + 3] = q.w;
      }
    } else {
      const int c = t & 31;
#pragma unroll
      for (int i = 0; i < 4 * QW; ++i) {
        const int r = (t >> 5) + (NW / 32) * i;
        w[i] = (r < nrows && c < ncols) ? __ldg(W + (long long)(row0 + r) * ldw + col0 + c) : 0.f;
      }
    }
  }
  __device__ static __forceinline__ void w_store_k(float* hi, float* lo, const float (&w)[4 * QW], bool vec) {
    const int t = threadIdx.x;
    if (vec) {
#pragma unroll
      for (int i = 0; i < QW; ++i)
        tc_store_quad<false>(hi, lo, (t >> 3) + (NW / 8) * i, t & 7, make_float4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]));
    } else {
#pragma unroll
      for (int i = 0; i < 4 * QW; ++i) tc_store_elem<false>(hi, lo, (t >> 5) + (NW / 32) * i, t & 31, w[i]);
    }
  }
};


// ------------------------------------------------------------------------------------------------
// self-test of the three operand configurations the engine uses (validated on the GPU by
// tests/test_gpu_tc_selftest.py): mode 0  D[r][n] = sum_k A[r][k] B[n][k]   (A, B K-major)
//                                 mode 1  D[r][j] = sum_n A[r][n] W[n][j]   (A K-major, W MN-major)
//                                 mode 2  D[n][k] = sum_r Z[r][n] X[r][k]   (Z, X MN-major, K = 128 rows)
// A: [128 x 32], B: [N x 32], W: [32 x N], Z: [128 x 64], X: [128 x 32]; out: [128 x N] (mode 2: N = 32,
// rows n < 64 meaningful).
// ------------------------------------------------------------------------------------------------
template <int = 0>
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
  const bool a_tmem = mode >= 3;                      // modes 3 / 4 = modes 0 / 1 with A staged in TMEM columns 64.. / 96..
  if (a_tmem) {
    const int q4 = warp & 3, cs = warp >> 2;
    float v[16], h[16], l[16];
    for (int i = 0; i < 16; i += 4) {
      const float4 x = *reinterpret_cast<const float4*>(A + (q4 * 32 + lane) * 32 + 16 * cs + i);
      v[i] = x.x; v[i + 1] = x.y; v[i + 2] = x.z; v[i + 3] = x.w;
    }
    for (int i = 0; i < 16; ++i) { h[i] = tf32_hi(v[i]); l[i] = v[i] - h[i]; }
    tmem_st16(tmem + ((unsigned)(q4 * 32) << 16) + 64 + 16 * cs, h);
    tmem_st16(tmem + ((unsigned)(q4 * 32) << 16) + 96 + 16 * cs, l);
    tmem_wait_st();
    mode -= 3;
  }
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
    if (!a_tmem)
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
  tc_fence_before();
  __syncthreads();
  if (tid == 0) {
    tc_fence_after();
    const unsigned ah = smem_u32(a_hi), al = smem_u32(a_lo), bh = smem_u32(b_hi), bl = smem_u32(b_lo);
    if (a_tmem) {
      const unsigned id = umma_idesc_tf32(128, N, 0, mode == 1 ? 1 : 0);
      for (int j = 0; j < 4; ++j) {
        const unsigned long long dbh = mode == 1 ? umma_desc_mn(bh, j, 4096) : umma_desc_k(bh, j);
        const unsigned long long dbl = mode == 1 ? umma_desc_mn(bl, j, 4096) : umma_desc_k(bl, j);
        umma_tf32_ts(tmem, tmem + 96 + 8 * j, dbh, id, j > 0);
        umma_tf32_ts(tmem, tmem + 64 + 8 * j, dbl, id, 1);
        umma_tf32_ts(tmem, tmem + 64 + 8 * j, dbh, id, 1);
      }
    } else if (mode == 0) {
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
