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
// Tensor-core engine: same three GEMM entry points as FmaEngine, 128-row tiles only.
// Shared-memory staging (1024-byte aligned, first in the dynamic allocation):
//   XB 64 KB = 2 buffers x {hi, lo} x [128 x 32] activation chunk images (16 KB each)
//   WB 32 KB = 2 buffers x {hi, lo} x [64 x 32] weight block images (8 KB each)
// The weight-gradient GEMM re-purposes them: XB = dz images {hi g0, hi g1, lo g0, lo g1}, WB = one
// input-chunk image pair {hi, lo}.
// ------------------------------------------------------------------------------------------------
struct TcEngine {
  static constexpr int RM = 4;
  static constexpr int TM = 128;
  static constexpr bool kTensor = true;
  static constexpr int kWorkers = 256;                    // worker warps: kWorkers / 128 per TMEM lane quarter
  static constexpr int NW = kWorkers;
  static constexpr int QA = 1024 / NW;                    // float4 per thread of a [128 x 32] chunk
  static constexpr int QW = 512 / NW;                     // float4 per thread of a weight block
  static constexpr int CS = NW / 128;                     // column slices of the accumulator (one per warp of a quarter)
  static constexpr int kBlockThreads = kWorkers + 32;     // + the MMA-issuing warp
  static constexpr int kMinBlocks = 1;
  using State = TcState;
  static size_t stage_bytes() { return 1024 + (size_t)(kTcStageXB + kTcStageWB + 512) * 4 + 64 + 2 * sizeof(TcCmd); }

  // mbarriers: full[2] (one arrival per worker: chunk staged) then done[2] (1 arrival: tcgen05.commit)
  __device__ static __forceinline__ unsigned long long* full_bar(const Smem& sm, int slot) { return sm.bar + slot; }
  __device__ static __forceinline__ unsigned long long* done_bar(const Smem& sm, int slot) { return sm.bar + 2 + slot; }
  __device__ static __forceinline__ TcCmd* cmd_slot(const Smem& sm, int slot) {
    return reinterpret_cast<TcCmd*>(sm.bar + 4) + slot;
  }

  __device__ static __forceinline__ char* carve(Smem& sm, char* p) {
    p += (1024 - (smem_u32(p) & 1023)) & 1023;
    float* f = reinterpret_cast<float*>(p);
    sm.XB = f; f += kTcStageXB;
    sm.WB = f; f += kTcStageWB;
    sm.RED = f; f += 512;
    sm.bar = reinterpret_cast<unsigned long long*>(f);
    char* q = reinterpret_cast<char*>(sm.bar + 4) + 2 * sizeof(TcCmd);
    sm.tslot = reinterpret_cast<unsigned*>(q);
    return q + 16;
  }
  __device__ static __forceinline__ void init(const Smem& sm, State& es) {      // every thread of the CTA
    const int tid = threadIdx.x;
    for (int i = tid; i < kTcStageXB + kTcStageWB; i += kBlockThreads) sm.XB[i] = 0.f;   // XB and WB are contiguous
    if (tid == 0) {
      mbar_init(full_bar(sm, 0), kWorkers); mbar_init(full_bar(sm, 1), kWorkers);
      mbar_init(done_bar(sm, 0), 1); mbar_init(done_bar(sm, 1), 1);
      mbar_fence_init();
    }
    if (tid < 32) tmem_alloc(sm.tslot, kTcTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    es.tmem = *sm.tslot;
    es.seq = 0;
    es.pending[0] = es.pending[1] = 0;
    es.parity[0] = es.parity[1] = 0;
    for (int i = 0; i < 16; ++i) es.t[i] = 0;
  }
  __device__ static __forceinline__ void fini(const Smem& sm, State& es) {      // workers only
    drain(sm, es);
    const int slot = es.seq & 1;
    if (threadIdx.x == 0) cmd_slot(sm, slot)->op = TC_OP_QUIT;
    mbar_arrive(full_bar(sm, slot));
    tc_fence_before();
    MMN_WSYNC_N(kWorkers);
    if (threadIdx.x < 32) tmem_dealloc(es.tmem, kTcTmemCols);
  }
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
  // the extra warp: wait for a staged chunk, issue its MMAs, commit them to the slot's done barrier
  __device__ static __forceinline__ void issuer_loop(const Smem& sm, State& es, long long* dbg = nullptr) {
    const unsigned xb = smem_u32(sm.XB), wb = smem_u32(sm.WB);     // warp-uniform
    const unsigned tmem0 = 0;                                      // the CTA's only TMEM allocation starts at column 0
    if (es.tmem != 0) __trap();
    const bool leader = elect_one() != 0;
    unsigned par[2] = {0u, 0u};
    long long t_idle = 0, t_issue = 0;
    for (unsigned seq = 0;; ++seq) {
      const unsigned slot = seq & 1u;
      const long long t0 = MMN_CLOCK();
      mbar_wait(full_bar(sm, slot), par[slot]);
      par[slot] ^= 1u;
      tc_fence_after();
      const unsigned op = cmd_slot(sm, slot)->op;    // mbar_wait above is an acquire + compiler barrier
      const unsigned kind = op & 3u, nj = (op >> 4) & 31u;
      const bool n64 = (op >> 2) & 1u, first = (op >> 3) & 1u;
      if (kind == TC_OP_QUIT) break;
      const long long t1 = MMN_CLOCK();
      const unsigned a_hi = xb + slot * 32768u, b_hi = wb + slot * 16384u;
      if (kind == TC_OP_NT) {
        if (n64) issue<false, false, 4>(a_hi, a_hi + 16384u, 0, b_hi, b_hi + 8192u, 0, umma_idesc_tf32(128, 64, 0, 0), tmem0, nj, first, leader);
        else issue<false, false, 4>(a_hi, a_hi + 16384u, 0, b_hi, b_hi + 8192u, 0, umma_idesc_tf32(128, 32, 0, 0), tmem0, nj, first, leader);
      } else if (kind == TC_OP_NN) {
        if (n64) issue<false, true, 4>(a_hi, a_hi + 16384u, 0, b_hi, b_hi + 8192u, 4096, umma_idesc_tf32(128, 64, 0, 1), tmem0, nj, first, leader);
        else issue<false, true, 4>(a_hi, a_hi + 16384u, 0, b_hi, b_hi + 8192u, 4096, umma_idesc_tf32(128, 32, 0, 1), tmem0, nj, first, leader);
      } else {           // TN: A = dz images at XB, B = input chunk at WB (slot 0) or the upper half of XB (slot 1)
        const unsigned in_hi = slot ? xb + 32768u : wb;
        issue<true, true, 16>(xb, xb + 16384u, 0, in_hi, in_hi + 16384u, 0, umma_idesc_tf32(128, 32, 1, 1), tmem0 + 32u * slot, 16, true, leader);
      }
      if (leader) umma_commit(done_bar(sm, slot));
      __syncwarp();
      const long long t2 = MMN_CLOCK();
      t_idle += t1 - t0;
      t_issue += t2 - t1;
    }
    if (dbg && leader) { dbg[0] = t_idle; dbg[1] = t_issue; }
  }
  __device__ static __forceinline__ void wait(const Smem& sm, State& es, int slot) {
    if (es.pending[slot]) {
      const long long t0 = MMN_CLOCK();
      mbar_wait(done_bar(sm, slot), es.parity[slot]);
      es.t[0] += MMN_CLOCK() - t0;
      es.parity[slot] ^= 1;
      es.pending[slot] = 0;
    }
  }
  __device__ static __forceinline__ void drain(const Smem& sm, State& es) {
    wait(sm, es, es.seq & 1);          // older commit first (in-order completion)
    wait(sm, es, (es.seq & 1) ^ 1);
  }
  // hand the chunk staged in `slot` to the issuer (the operand images sit at fixed places of the slot)
  __device__ static __forceinline__ void post(const Smem& sm, State& es, int slot, unsigned kind, int N, int nj,
                                              unsigned first) {
    if (threadIdx.x == 0) cmd_slot(sm, slot)->op = kind | (N == 64 ? 4u : 0u) | (first ? 8u : 0u) | ((unsigned)nj << 4);
    fence_proxy_async();               // this thread's image stores -> visible to the tensor-core (async) proxy
    tc_fence_before();                 // this thread's earlier tcgen05.ld of the accumulator are ordered before
    mbar_arrive(full_bar(sm, slot));
    es.pending[slot] = 1;
    es.seq += 1;
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
  // ---- weight block for the data gradient: rows n (contraction) x up to 64 output columns j ->
  //      MN-major image pairs, one 4 KB image per group of 32 columns ----
  __device__ static __forceinline__ void w_load_mn(float (&w)[4 * QW], const float* __restrict__ W, int ldw, int row0, int nrows,
                                                   int col0, int ncols, bool vec) {
    const int t = threadIdx.x;
    if (vec) {
      const int c4 = (t & 15) * 4;
#pragma unroll
      for (int i = 0; i < QW; ++i) {
        const int r = (t >> 4) + (NW / 16) * i;
        float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < nrows && c4 < ncols) {
          const float* p = W + (long long)(row0 + r) * ldw + col0 + c4;
          if (c4 + 3 < ncols) q = __ldg(reinterpret_cast<const float4*>(p));
          else { q.x = __ldg(p); if (c4 + 1 < ncols) q.y = __ldg(p + 1); if (c4 + 2 < ncols) q.z = __ldg(p + 2); }
        }
        w[4 * i] = q.x; w[4 * i + 1] = q.y; w[4 * i + 2] = q.z; w[4 * i + 3] = q.w;
      }
    } else {
      const int c = t & 63;
#pragma unroll
      for (int i = 0; i < 4 * QW; ++i) {
        const int r = (t >> 6) + (NW / 64) * i;
        w[i] = (r < nrows && c < ncols) ? __ldg(W + (long long)(row0 + r) * ldw + col0 + c) : 0.f;
      }
    }
  }
  __device__ static __forceinline__ void w_store_mn(float* hi, float* lo, const float (&w)[4 * QW], bool vec) {
    const int t = threadIdx.x;
    if (vec) {
      const int c4 = t & 15, g = c4 >> 3;
#pragma unroll
      for (int i = 0; i < QW; ++i)
        tc_store_quad<true>(hi + g * 1024, lo + g * 1024, (t >> 4) + (NW / 16) * i, c4 & 7,
                            make_float4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]));
    } else {
      const int c = t & 63, g = c >> 5;
#pragma unroll
      for (int i = 0; i < 4 * QW; ++i) tc_store_elem<true>(hi + g * 1024, lo + g * 1024, (t >> 6) + (NW / 64) * i, c & 31, w[i]);
    }
  }

  // ---- activation chunk [128 x 32]: global sources are fetched into 8 registers per thread early ...
  __device__ static __forceinline__ void a_load(float (&v)[4 * QA], const ASeg& sg, int k0, int kw, int rows_valid, bool vec) {
    if (sg.kind != SEG_X && sg.kind != SEG_STASH) return;
    const int t = threadIdx.x;
    if (vec) {
      const int c4 = (t & 7) * 4;
#pragma unroll
      for (int i = 0; i < QA; ++i) {
        const int r = (t >> 3) + (NW / 8) * i;
        float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < rows_valid && c4 < kw) {
          const float4* p = reinterpret_cast<const float4*>(sg.ptr + (long long)r * sg.ld + k0 + c4);
          q = sg.kind == SEG_X ? __ldg(p) : __ldcg(p);
        }
        v[4 * i] = q.x; v[4 * i + 1] = q.y; v[4 * i + 2] = q.z; v[4 * i + 3] = q.w;
      }
    } else {
      const int c = t & 31;
#pragma unroll
      for (int i = 0; i < 4 * QA; ++i) {
        const int r = (t >> 5) + (NW / 32) * i;
        float x = 0.f;
        if (r < rows_valid && c < kw) {
          const float* p = sg.ptr + (long long)r * sg.ld + k0 + c;
          x = sg.kind == SEG_X ? __ldg(p) : __ldcg(p);   // stash: written earlier by this CTA, L2-coherent load
        }
        v[i] = x;
      }
    }
  }
  // ---- ... and written as an image pair here; shared-memory tiles are read here.  NaN scan / sanitise
  //      for x, dropout when enabled. ----
  template <bool MN>
  __device__ static __forceinline__ void a_store(float* hi, float* lo, float (&v)[4 * QA], const Smem& sm, const ASeg& sg,
                                                 int k0, int kw, const Drop& drop, bool scan_nan, bool vec) {
    const int t = threadIdx.x;
    const bool from_smem = sg.kind == SEG_SMEM || sg.kind == SEG_SMEM_STAGED;
    if (vec || from_smem) {
      const int c4 = t & 7;
#pragma unroll
      for (int i = 0; i < QA; ++i) {
        const int r = (t >> 3) + (NW / 8) * i;
        float4 q;
        if (from_smem) q = *reinterpret_cast<const float4*>(sg.ptr + (long long)r * sg.ld + k0 + 4 * c4);   // zero-padded tile
        else q = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        if (sg.kind == SEG_X) {
          bool bad = false;
          if (q.x != q.x) { bad = true; q.x = 0.f; }
          if (q.y != q.y) { bad = true; q.y = 0.f; }
          if (q.z != q.z) { bad = true; q.z = 0.f; }
          if (q.w != q.w) { bad = true; q.w = 0.f; }
          if (bad && scan_nan) sm.rownan[r] = 1;
        }
        if (drop.enabled) {
          const unsigned col = (unsigned)(sg.wcol + k0 + 4 * c4), row = drop.row_base + (unsigned)r;
          if ((col & 1u) == 0) {
            const unsigned h0 = mmn_dropout_hash(drop.seed_mix, row, col >> 1);
            const unsigned h1 = mmn_dropout_hash(drop.seed_mix, row, (col >> 1) + 1);
            q.x = (h0 & 0xffffu) >= drop.thr ? q.x * drop.scale : 0.f;
            q.y = (h0 >> 16) >= drop.thr ? q.y * drop.scale : 0.f;
            q.z = (h1 & 0xffffu) >= drop.thr ? q.z * drop.scale : 0.f;
            q.w = (h1 >> 16) >= drop.thr ? q.w * drop.scale : 0.f;
          } else {
            q.x = mmn_dropout_keep(drop.seed_mix, row, col + 0, drop.thr) ? q.x * drop.scale : 0.f;
            q.y = mmn_dropout_keep(drop.seed_mix, row, col + 1, drop.thr) ? q.y * drop.scale : 0.f;
            q.z = mmn_dropout_keep(drop.seed_mix, row, col + 2, drop.thr) ? q.z * drop.scale : 0.f;
            q.w = mmn_dropout_keep(drop.seed_mix, row, col + 3, drop.thr) ? q.w * drop.scale : 0.f;
          }
          if (4 * c4 + 0 >= kw) q.x = 0.f;
          if (4 * c4 + 1 >= kw) q.y = 0.f;
          if (4 * c4 + 2 >= kw) q.z = 0.f;
          if (4 * c4 + 3 >= kw) q.w = 0.f;
        }
        tc_store_quad<MN>(hi, lo, r, c4, q);
      }
    } else {
      const int c = t & 31;
#pragma unroll
      for (int i = 0; i < 4 * QA; ++i) {
        const int r = (t >> 5) + (NW / 32) * i;
        float x = v[i];
        if (sg.kind == SEG_X && x != x) {
          if (scan_nan) sm.rownan[r] = 1;
          x = 0.f;
        }
        if (drop.enabled && c < kw)
          x = mmn_dropout_keep(drop.seed_mix, drop.row_base + (unsigned)r, (unsigned)(sg.wcol + k0 + c), drop.thr)
                  ? x * drop.scale : 0.f;
        tc_store_elem<MN>(hi, lo, r, c, x);
      }
    }
  }
  // this thread's accumulator slice: lane quarter q = warp & 3 (row 32 q + lane), column slice warp >> 2
  static_assert(CS == 2, "the epilogues below give each thread 16-column slices (2 warps per lane quarter)");
  // this thread's 16-column slice h of an N = 32 (h = 0) or N = 64 (h = 0, 1) accumulator: columns 32 h + 16 cs ..
  __device__ static __forceinline__ void acc_load(const State& es, int q, int col, float (&v)[16]) {
    tmem_ld16(es.tmem + ((unsigned)(32 * q) << 16) + col, v);
  }

  // ------------------------------------------------------------------------------------------------
  // out[r][n] = bias[n] + sum_seg sum_k a[r][k] W[n][wcol + k]      (A, W K-major; N pass of 32 / 64)
  // ------------------------------------------------------------------------------------------------
  template <class Epi>
  __device__ static __forceinline__ void gemm_nt(const Smem& sm, State& es, const float* __restrict__ W, int ldw, int N,
                                                 const float* __restrict__ bias, int act, const ASeg* segs, int nseg,
                                                 const Drop& drop, int rows_valid, bool scan_nan, Epi epi) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, q = warp & 3, cs = warp >> 2;
    const long long t_begin = MMN_CLOCK();
    bool avec[2];
    avec[0] = seg_vec_ok(segs[0]);
    avec[1] = nseg > 1 ? seg_vec_ok(segs[1]) : false;
    for (int n0 = 0; n0 < N; n0 += 64) {
      const int nrows = min(64, N - n0);
      const int Np = nrows <= 32 ? 32 : 64;
      float wr[4 * QW], ar[4 * QA];
      ChunkIt it{0, 0};
      bool wv = w_vec_ok(W, ldw, segs[0].wcol);
      w_load_k(wr, W, ldw, n0, nrows, segs[0].wcol, min(KC, segs[0].width), wv);
      a_load(ar, segs[0], 0, min(KC, segs[0].width), rows_valid, avec[0]);
      unsigned first = 1;
      int slot = 0;
      while (it.valid(nseg)) {
        const ASeg sg = segs[it.s];
        const int kw = min(KC, sg.width - it.k0);
        slot = es.seq & 1;
        float* xh = sm.XB + slot * 8192;
        float* wh = sm.WB + slot * 4096;
        wait(sm, es, slot);                       // MMAs that read this slot two chunks ago are done
        w_store_k(wh, wh + 2048, wr, wv);
        a_store<false>(xh, xh + 4096, ar, sm, sg, it.k0, kw, drop, scan_nan && n0 == 0, avec[it.s]);
        ChunkIt nx = it;
        nx.next(segs);
        if (nx.valid(nseg)) {
          const ASeg& ns = segs[nx.s];
          const int nkw = min(KC, ns.width - nx.k0);
          wv = w_vec_ok(W, ldw, ns.wcol + nx.k0);
          w_load_k(wr, W, ldw, n0, nrows, ns.wcol + nx.k0, nkw, wv);
          a_load(ar, ns, nx.k0, nkw, rows_valid, avec[nx.s]);
        }
        post(sm, es, slot, TC_OP_NT, Np, (kw + 7) >> 3, first);
        first = 0;
        it = nx;
      }
      auto load_bias = [&](int n, float (&bj)[16]) {
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
          float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (n + i + 3 < N) b4 = __ldg(reinterpret_cast<const float4*>(bias + n + i));     // bias offsets are 16-byte aligned
          else {
            if (n + i < N) b4.x = __ldg(bias + n + i);
            if (n + i + 1 < N) b4.y = __ldg(bias + n + i + 1);
            if (n + i + 2 < N) b4.z = __ldg(bias + n + i + 2);
          }
          bj[i] = b4.x; bj[i + 1] = b4.y; bj[i + 2] = b4.z; bj[i + 3] = b4.w;
        }
      };
      float bj[16];
      load_bias(n0 + 16 * cs, bj);
      wait(sm, es, slot ^ 1);     // older commit first (in-order completion), then the last one
      wait(sm, es, slot);
      tc_fence_after();
      const long long t_epi = MMN_CLOCK();
      const int r = 32 * q + lane;
      for (int h = 0; h < (Np >> 5); ++h) {
        const int cb = 32 * h + 16 * cs;
        if (h) load_bias(n0 + cb, bj);
        float v[16];
        acc_load(es, q, cb, v);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] += bj[i];
        act_fwd_n(act, v);
#pragma unroll
        for (int i = 0; i < 16; ++i) epi(r, n0 + cb + i, v[i]);
      }
      es.t[4] += MMN_CLOCK() - t_epi;
    }
    MMN_WSYNC_N(kWorkers);        // outputs (and the row NaN flags) visible to every worker
    es.t[1] += MMN_CLOCK() - t_begin;
  }

  // ------------------------------------------------------------------------------------------------
  // out[r][j] = sum_n dz[r][n] W[n][col0 + j]      (dz K-major from its shared tile, W MN-major)
  // ------------------------------------------------------------------------------------------------
  template <class Pre, class Epi>
  __device__ static __forceinline__ void gemm_nn(const Smem& sm, State& es, const float* dz, int ldd, int N,
                                                 const float* __restrict__ W, int ldw, int col0, int J, Pre pre, Epi epi) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, q = warp & 3, cs = warp >> 2;
    const int r = 32 * q + lane;
    const long long t_begin = MMN_CLOCK();
    Drop nodrop;
    nodrop.enabled = 0; nodrop.seed_mix = 0; nodrop.thr = 0; nodrop.row_base = 0; nodrop.scale = 1.f;
    for (int j0 = 0; j0 < J; j0 += 64) {
      const int jw = min(64, J - j0);
      const int Np = jw <= 32 ? 32 : 64;
      float pv[16];
#pragma unroll
      for (int i = 0; i < 16; i += 4) {
        const float4 p4 = pre(r, j0 + 16 * cs + i);
        pv[i] = p4.x; pv[i + 1] = p4.y; pv[i + 2] = p4.z; pv[i + 3] = p4.w;
      }
      const bool wv = w_vec_ok(W, ldw, col0 + j0);
      float wr[4 * QW], dummy[4 * QA];
      w_load_mn(wr, W, ldw, 0, min(32, N), col0 + j0, jw, wv);
      unsigned first = 1;
      int slot = 0;
      for (int n0 = 0; n0 < N; n0 += 32) {
        const int nw = min(32, N - n0);
        slot = es.seq & 1;
        float* xh = sm.XB + slot * 8192;
        float* wh = sm.WB + slot * 4096;
        wait(sm, es, slot);
        w_store_mn(wh, wh + 2048, wr, wv);
        ASeg sg;
        sg.ptr = dz; sg.ld = ldd; sg.width = N; sg.kind = SEG_SMEM; sg.wcol = 0;
        a_store<false>(xh, xh + 4096, dummy, sm, sg, n0, nw, nodrop, false, false);
        if (n0 + 32 < N) w_load_mn(wr, W, ldw, n0 + 32, min(32, N - n0 - 32), col0 + j0, jw, wv);
        post(sm, es, slot, TC_OP_NN, Np, (nw + 7) >> 3, first);
        first = 0;
      }
      wait(sm, es, slot ^ 1);
      wait(sm, es, slot);
      tc_fence_after();
      const long long t_epi = MMN_CLOCK();
      for (int h = 0; h < (Np >> 5); ++h) {
        const int cb = 32 * h + 16 * cs;
        if (h) {
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            const float4 p4 = pre(r, j0 + cb + i);
            pv[i] = p4.x; pv[i + 1] = p4.y; pv[i + 2] = p4.z; pv[i + 3] = p4.w;
          }
        }
        float v[16];
        acc_load(es, q, cb, v);
#pragma unroll
        for (int i = 0; i < 16; ++i) epi(r, j0 + cb + i, v[i], pv[i]);
      }
      es.t[5] += MMN_CLOCK() - t_epi;
    }
    MMN_WSYNC_N(kWorkers);
    es.t[2] += MMN_CLOCK() - t_begin;
  }

  // ------------------------------------------------------------------------------------------------
  // gW[n][wcol + k] += sum_r dz[r][n] a[r][k]      (both operands MN-major, contraction = the 128 rows)
  // D[n][k]: lanes = output rows n of one 32-column group of dz (all four lane quarters see the same
  // group: group stride 0), columns = the 32 k of a chunk.  Staging: XB[0,32K) = dz {hi, lo};
  // input chunks alternate between WB and XB[32K,64K); accumulator columns alternate with them, so the
  // epilogue (TMEM -> red.global) of chunk c overlaps the MMAs of chunk c + 1.
  // ------------------------------------------------------------------------------------------------
  __device__ static __forceinline__ void gemm_tn(const Smem& sm, State& es, const float* dz, int ldd, int N, const ASeg& sg,
                                                 const Drop& drop, int rows_valid, float* __restrict__ gW, int ldw) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, q = warp & 3, cs = warp >> 2;
    const bool vec_ok = ((ldw & 3) == 0) && ((sg.wcol & 3) == 0) && ((reinterpret_cast<size_t>(gW) & 15) == 0);
    const bool avec = seg_vec_ok(sg);
    const long long t_begin = MMN_CLOCK();
    auto epilogue = [&](int slot, int nb0, int k0) {
      wait(sm, es, slot);
      tc_fence_after();
      const long long t_epi = MMN_CLOCK();
      constexpr int TC = 16;               // chunk columns per thread
      float v[16];
      acc_load(es, q, 32 * slot + TC * cs, v);
      const int n = nb0 + lane;
      if (q == 0 && n < N) {
        const int kc = k0 + TC * cs;
        float* dst = gW + (long long)n * ldw + sg.wcol + kc;
#pragma unroll
        for (int i = 0; i < TC; i += 4) {
          if (vec_ok && kc + i + 3 < sg.width) {
            atomicAdd(reinterpret_cast<float4*>(dst + i), make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]));
          } else {
#pragma unroll
            for (int u = 0; u < 4; ++u)
              if (kc + i + u < sg.width) atomicAdd(dst + i + u, v[i + u]);
          }
        }
      }
      es.t[6] += MMN_CLOCK() - t_epi;
    };
    for (int nb0 = 0; nb0 < N; nb0 += 32) {
      float ar[4 * QA];
      a_load(ar, sg, 0, min(KC, sg.width), rows_valid, avec);
      drain(sm, es);
      MMN_WSYNC_N(kWorkers);       // every worker is past its reads of the staging buffers / dz tile writes
      for (int idx = tid; idx < 128 * 8; idx += kWorkers) {      // dz column group -> MN-major A image pair
        const int rr = idx >> 3, c4 = idx & 7;
        const float4 v = *reinterpret_cast<const float4*>(dz + rr * ldd + nb0 + 4 * c4);
        tc_store_quad<true>(sm.XB, sm.XB + 4096, rr, c4, v);
      }
      int prev_slot = -1, prev_k0 = 0;
      for (int k0 = 0; k0 < sg.width; k0 += KC) {
        const int kw = min(KC, sg.width - k0);
        const int slot = es.seq & 1;
        float* ih = slot ? sm.XB + 8192 : sm.WB;
        wait(sm, es, slot);        // (already collected by the epilogue two chunks ago)
        a_store<true>(ih, ih + 4096, ar, sm, sg, k0, kw, drop, false, avec);
        if (k0 + KC < sg.width) a_load(ar, sg, k0 + KC, min(KC, sg.width - k0 - KC), rows_valid, avec);
        post(sm, es, slot, TC_OP_TN, 32, 16, 1u);
        if (prev_slot >= 0) epilogue(prev_slot, nb0, prev_k0);
        prev_slot = slot;
        prev_k0 = k0;
      }
      if (prev_slot >= 0) epilogue(prev_slot, nb0, prev_k0);
    }
    MMN_WSYNC_N(kWorkers);
    es.t[3] += MMN_CLOCK() - t_begin;
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
