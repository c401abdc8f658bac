// cuda_emu.h — a tiny single-threaded CUDA execution-model emulator.  TEST INFRASTRUCTURE ONLY.
//
// The build container has nvcc but no GPU.  To debug the indexing / masking / bookkeeping logic
// of the kernels before spending GPU minutes, tests/emu/build_emu.sh compiles the SAME kernel
// sources (multimodn_b200/csrc/*.cu) with g++ against this header into tests/emu/libmmn_emu.so.
// Every CUDA thread of a block is a ucontext fiber; __syncthreads and the warp shuffles are
// cooperative barriers between fibers; blocks run one after another.  "Device" pointers are
// host pointers.  It cannot find races (fibers are deterministic) and says nothing about
// performance; compute-sanitizer and the -m gpu tests on the B200 do that.
//
// The product package never loads this library: multimodn_b200/_lib.py only opens libmmn.so
// (nvcc, sm_100a) and MultiModN refuses non-CUDA devices.
#pragma once
#ifndef MMN_EMU
#error "cuda_emu.h is only for the -DMMN_EMU host build"
#endif

#include <ucontext.h>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __shared__ static        // blocks run one after another, the threads of a block are fibers of one OS thread
#define __restrict__ __restrict

struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(8) float2 { float x, y; };
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }

typedef void* cudaStream_t;
typedef int cudaError_t;
enum { cudaSuccess = 0 };
enum { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
enum { cudaDevAttrMultiProcessorCount = 16, cudaDevAttrMaxSharedMemoryPerBlockOptin = 97 };

namespace emu {
struct Fiber {
  ucontext_t ctx;
  std::vector<char> stack;
  bool done = false;
  unsigned tid = 0;
};
struct State {
  ucontext_t sched;
  std::vector<Fiber> fibers;
  Fiber* cur = nullptr;
  unsigned nthreads = 0;
  // block barrier
  unsigned bar_count = 0;
  unsigned bar_gen = 0;
  int bar_or = 0, bar_or_result = 0;
  // warp barriers
  std::vector<unsigned> wcount, wgen;
  std::vector<uint64_t> wslot;  // [warp][lane]
  std::function<void()> body;
  char* dyn_smem = nullptr;
  size_t dyn_smem_bytes = 0;
};
inline State& st() { static State s; return s; }
inline dim3& tIdx() { static dim3 v; return v; }
inline dim3& bIdx() { static dim3 v; return v; }
inline dim3& bDim() { static dim3 v; return v; }
inline dim3& gDim() { static dim3 v; return v; }

inline void yield() {
  State& s = st();
  Fiber* f = s.cur;
  swapcontext(&f->ctx, &s.sched);
  tIdx() = dim3(f->tid);   // restored after resume
}
inline void fiber_entry() {
  State& s = st();
  s.body();
  s.cur->done = true;
  swapcontext(&s.cur->ctx, &s.sched);
}
inline void run_block(unsigned nthreads, const std::function<void()>& body) {
  State& s = st();
  s.body = body;
  s.nthreads = nthreads;
  s.bar_count = 0; s.bar_gen = 0; s.bar_or = 0;
  unsigned nwarps = (nthreads + 31) / 32;
  s.wcount.assign(nwarps, 0); s.wgen.assign(nwarps, 0); s.wslot.assign(nwarps * 32, 0);
  if (s.fibers.size() < nthreads) s.fibers.resize(nthreads);
  for (unsigned t = 0; t < nthreads; ++t) {
    Fiber& f = s.fibers[t];
    if (f.stack.empty()) f.stack.resize(256 * 1024);
    f.done = false; f.tid = t;
    getcontext(&f.ctx);
    f.ctx.uc_stack.ss_sp = f.stack.data();
    f.ctx.uc_stack.ss_size = f.stack.size();
    f.ctx.uc_link = &s.sched;
    makecontext(&f.ctx, (void (*)())fiber_entry, 0);
  }
  unsigned remaining = nthreads;
  while (remaining) {
    for (unsigned t = 0; t < nthreads; ++t) {
      Fiber& f = s.fibers[t];
      if (f.done) continue;
      s.cur = &f;
      tIdx() = dim3(t);
      swapcontext(&s.sched, &f.ctx);
      if (f.done) --remaining;
    }
  }
}
inline void block_barrier(int pred, int* or_out) {
  State& s = st();
  unsigned gen = s.bar_gen;
  s.bar_or |= pred;
  if (++s.bar_count == s.nthreads) {
    s.bar_count = 0; s.bar_or_result = s.bar_or; s.bar_or = 0; ++s.bar_gen;
  } else {
    while (s.bar_gen == gen) yield();
  }
  if (or_out) *or_out = s.bar_or_result;
}
// barrier over threads [0, count) (bar.sync id, count); one id in use
inline void named_barrier(unsigned count) {
  State& s = st();
  static unsigned n_arrived = 0, gen_ = 0;
  unsigned gen = gen_;
  if (++n_arrived == count) { n_arrived = 0; ++gen_; }
  else while (gen_ == gen) yield();
}
inline void warp_barrier() {
  State& s = st();
  unsigned w = tIdx().x / 32;
  unsigned lanes = (w == s.wcount.size() - 1 && s.nthreads % 32) ? s.nthreads % 32 : 32;
  unsigned gen = s.wgen[w];
  if (++s.wcount[w] == lanes) { s.wcount[w] = 0; ++s.wgen[w]; }
  else {
    unsigned long spins = 0;
    while (s.wgen[w] == gen) {
      yield();
      if (++spins > 50000000ul) { fprintf(stderr, "cuda_emu: warp-collective op reached by only part of warp %u (divergence)\n", w); abort(); }
    }
  }
}
// a .sync.aligned instruction: every lane of the warp must execute it together
inline void warp_collective() { warp_barrier(); }
// barrier `id` over `count` threads (bar.sync id, count); any number of ids, each with its own arrival counter
inline void named_barrier_id(unsigned id, unsigned count) {
  static unsigned n_arrived[16] = {0}, gen_[16] = {0};
  const unsigned gen = gen_[id];
  if (++n_arrived[id] == count) { n_arrived[id] = 0; ++gen_[id]; }
  else while (gen_[id] == gen) yield();
}
// warp-wide exchange: every lane publishes N 64-bit words, then reads any lane's words (mma.sync / ldmatrix / movmatrix
// models).  Usage: publish(...); ...read peer()...; done();
struct WarpExchange {
  static constexpr int kWords = 8;
  static uint64_t (&buf())[64][32][kWords] { static uint64_t b[64][32][kWords]; return b; }
  static void publish(const uint64_t* words, int n) {
    const unsigned w = tIdx().x / 32, lane = tIdx().x % 32;
    for (int i = 0; i < n; ++i) buf()[w][lane][i] = words[i];
    warp_barrier();
  }
  static uint64_t peer(unsigned lane, int i) { return buf()[tIdx().x / 32][lane & 31][i]; }
  static void done() { warp_barrier(); }
};
template <class T>
inline T shfl(T v, unsigned src_lane) {
  static_assert(sizeof(T) <= 8, "shfl payload");
  State& s = st();
  unsigned w = tIdx().x / 32, lane = tIdx().x % 32;
  uint64_t bits = 0; memcpy(&bits, &v, sizeof(T));
  s.wslot[w * 32 + lane] = bits;
  warp_barrier();
  uint64_t got = s.wslot[w * 32 + (src_lane & 31)];
  warp_barrier();
  T out; memcpy(&out, &got, sizeof(T));
  return out;
}
}  // namespace emu

#define threadIdx (emu::tIdx())
#define blockIdx (emu::bIdx())
#define blockDim (emu::bDim())
#define gridDim (emu::gDim())

static inline void __syncthreads() { emu::block_barrier(0, nullptr); }
#define MMN_WSYNC_N(n) emu::named_barrier(n)
#define MMN_CLOCK() 0ll
static inline int __syncthreads_or(int p) { int r; emu::block_barrier(p != 0, &r); return r; }
static inline void __syncwarp(unsigned = 0xffffffffu) { emu::warp_barrier(); }
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int m) { return emu::shfl(v, (threadIdx.x % 32) ^ m); }
template <class T> static inline T __shfl_down_sync(unsigned, T v, int d) {
  unsigned l = threadIdx.x % 32; return emu::shfl(v, l + d < 32 ? l + d : l);
}
template <class T> static inline T __shfl_sync(unsigned, T v, int src) { return emu::shfl(v, src); }
static inline unsigned __ballot_sync(unsigned, int pred) {
  uint64_t w = pred ? 1 : 0;
  emu::WarpExchange::publish(&w, 1);
  unsigned m = 0;
  for (unsigned l = 0; l < 32; ++l) m |= (unsigned)(emu::WarpExchange::peer(l, 0) & 1) << l;
  emu::WarpExchange::done();
  return m;
}
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline void __threadfence() {}
static inline void __trap() { abort(); }

template <class T> static inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
static inline void atomicAdd(float4* p, float4 v) { p->x += v.x; p->y += v.y; p->z += v.z; p->w += v.w; }
static inline int atomicOr(int* p, int v) { int o = *p; *p |= v; return o; }
static inline int atomicMax(int* p, int v) { int o = *p; if (v > o) *p = v; return o; }

template <class T> static inline T __ldg(const T* p) { return *p; }
template <class T> static inline T __ldcg(const T* p) { return *p; }
template <class T> static inline T __ldcs(const T* p) { return *p; }
template <class T> static inline void __stcg(T* p, T v) { *p = v; }
template <class T> static inline void __stcs(T* p, T v) { *p = v; }
static inline float __fdividef(float a, float b) { return a / b; }
#define __expf(a) expf(a)      /* glibc declares __expf / __logf itself */
#define __logf(a) logf(a)
static inline float __saturatef(float a) { return a < 0 ? 0 : (a > 1 ? 1 : a); }
static inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
using std::isnan;
#include <algorithm>
using std::min;
using std::max;

// ---- runtime API subset (device memory == host memory) ----
static inline cudaError_t cudaMalloc(void** p, size_t n) { *p = malloc(n ? n : 1); return *p ? 0 : 2; }
static inline cudaError_t cudaFree(void* p) { free(p); return 0; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, int) { memcpy(d, s, n); return 0; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, int, cudaStream_t) { memcpy(d, s, n); return 0; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { memset(d, v, n); return 0; }
static inline cudaError_t cudaGetLastError() { return 0; }
static inline cudaError_t cudaPeekAtLastError() { return 0; }
static inline const char* cudaGetErrorString(cudaError_t) { return "emu error"; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return 0; }
static inline cudaError_t cudaDeviceGetAttribute(int* v, int attr, int) {
  *v = (attr == cudaDevAttrMultiProcessorCount) ? 3 : 232448;   // 3 "SMs": exercises multi-tile CTAs
  return 0;
}
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, int, int) { return 0; }

// launch: blocks run sequentially, threads of a block as fibers
#define MMN_LAUNCH(kernel, grid, block, smem, stream, ...)                          \
  do {                                                                              \
    dim3 g_ = (grid), b_ = (block);                                                 \
    emu::gDim() = g_; emu::bDim() = b_;                                             \
    std::vector<char> smem_((size_t)(smem) + 1024);                                   \
    emu::st().dyn_smem = (char*)(((uintptr_t)smem_.data() + 1023) & ~(uintptr_t)1023);  \
    emu::st().dyn_smem_bytes = (smem);                                              \
    for (unsigned bx_ = 0; bx_ < g_.x; ++bx_) {                                     \
      emu::bIdx() = dim3(bx_);                                                      \
      emu::run_block(b_.x, [&]() { kernel(__VA_ARGS__); });                         \
    }                                                                               \
  } while (0)
#define MMN_DYN_SMEM(name) char* name = emu::st().dyn_smem
