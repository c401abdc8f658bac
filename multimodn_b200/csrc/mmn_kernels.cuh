// mmn_kernels.cuh — the fused sequential-fusion step for sm_100a (fp32 FMA path).
//
// One persistent CTA per batch tile of TM = 32*RM rows walks the whole encoding sequence with the
// running state resident in shared memory (reference loop nest: multimodn/multimodn.py:139-191):
//
//   s_0 = init state                                   state.py:29-32
//   every decoder on s_0: CE, arg-max, counters        multimodn.py:141-157
//   for each (position, encoder id) in the sequence:   multimodn.py:159-163
//       stream x tile from HBM, per-row NaN scan       multimodn.py:168 (per row instead of per batch)
//       s^ = Encoder(s, x)                             mlp_encoder.py:40-47 / 74-80
//       s  = present ? s^ : s ; state-change sum       multimodn.py:173-174
//       every decoder on s                             multimodn.py:176-191
//   (training) replay the sequence in reverse for loss.backward()   multimodn.py:194-203
//
// Every Linear layer is a small GEMM over the tile, computed with 4x4 (RM x 4) register tiles by
// 32 row-threads x 8 column-threads.  Three shapes are needed and each has one routine:
//   gemm_nt : out[r][n]  = sum_k a[r][k]  W[n][k]   forward
//   gemm_nn : din[r][j]  = sum_n dz[r][n] W[n][j]   data gradient
//   gemm_tn : dW[n][k]  += sum_r dz[r][n] a[r][k]   weight gradient (rows split over 4 thread groups,
//                                                   reduced in shared memory, one red.global per tile)
// Shared-memory tiles are row-major with leading dimension == 4 (mod 32) and rows / columns are
// assigned to threads with stride (r = ty + 32 i, n = tx + 8 j) so that every LDS.128 of a warp
// touches 32 distinct banks.  Invariant: columns [width, roundup32(width)) of every activation tile
// are zero, so K loops may run in whole float4 steps.
//
// The backward pass needs the forward activations of the tile; they are stashed in a per-CTA slot of
// global memory that is reused tile after tile (it stays L2-resident) and read back with ld.global.cg.
#pragma once

#include "mmn_common.cuh"

namespace mmn {

// ------------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float act_fwd(int act, float z) {
  switch (act) {
    case MMN_ACT_RELU: return z > 0.f ? z : 0.f;
    case MMN_ACT_SIGMOID: return __fdividef(1.f, 1.f + expf(-z));
    case MMN_ACT_TANH: return tanhf(z);
    default: return z;
  }
}
// derivative through the activation OUTPUT
__device__ __forceinline__ float act_bwd(int act, float out) {
  switch (act) {
    case MMN_ACT_RELU: return out > 0.f ? 1.f : 0.f;
    case MMN_ACT_SIGMOID: return out * (1.f - out);
    case MMN_ACT_TANH: return 1.f - out * out;
    default: return 1.f;
  }
}

// activation over a small register array with the kind test hoisted out of the element loop
// (a per-element switch unrolls into a branch ladder plus the slow-path division call, 16x per epilogue)
template <int N>
__device__ __forceinline__ void act_fwd_n(int act, float (&v)[N]) {
  if (act == MMN_ACT_RELU) {
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = fmaxf(v[i], 0.f);
  } else if (act == MMN_ACT_SIGMOID) {
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = __fdividef(1.f, 1.f + expf(-v[i]));
  } else if (act == MMN_ACT_TANH) {
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = tanhf(v[i]);
  }
}
// derivative through the activation OUTPUT as branch-free arithmetic: d = relu ? (a > 0) : c0 + c1 a + c2 a^2
struct ActBwd {
  float c0, c1, c2;
  bool relu;
  __device__ __forceinline__ explicit ActBwd(int act) {
    relu = act == MMN_ACT_RELU;
    c0 = (act == MMN_ACT_SIGMOID) ? 0.f : 1.f;
    c1 = (act == MMN_ACT_SIGMOID) ? 1.f : 0.f;
    c2 = (act == MMN_ACT_SIGMOID || act == MMN_ACT_TANH) ? -1.f : 0.f;
  }
  __device__ __forceinline__ float operator()(float a) const {
    const float poly = fmaf(a, fmaf(a, c2, c1), c0);
    return relu ? (a > 0.f ? 1.f : 0.f) : poly;
  }
};

// dropout keep decision: counter-based hash of (seed, encoder, global row, column pair); one 32-bit
// hash serves two adjacent columns (16 bits each).  The oracle (oracle/multimodn_oracle.py:
// dropout_keep) computes the same bits.
__device__ __forceinline__ unsigned mmn_dropout_hash(unsigned seed_mix, unsigned row, unsigned colpair) {
  unsigned x = row * 0x85EBCA6Bu + colpair * 0xC2B2AE35u + seed_mix;
  x ^= x >> 16; x *= 0x7FEB352Du; x ^= x >> 15; x *= 0x846CA68Bu; x ^= x >> 16;
  return x;
}
__device__ __forceinline__ bool mmn_dropout_keep(unsigned seed_mix, unsigned row, unsigned col, unsigned thr16) {
  const unsigned h = mmn_dropout_hash(seed_mix, row, col >> 1);
  return ((col & 1u) ? (h >> 16) : (h & 0xffffu)) >= thr16;
}

struct Drop {
  int enabled;
  unsigned seed_mix, thr, row_base;
  float scale;
};

struct Smem {
  float *S, *T, *A, *B, *XB, *WB, *RED;   // XB, WB, RED are double-buffered
  int* ys;
  int* rownan;
  int* cnt;          // [E+1] present rows per history row, summed over this CTA's tiles
  int* tile_any;     // [E+1] does the current tile have a present row at step k
  double* met;
  unsigned char* present;   // [(E+1)][TM]
  unsigned long long* bar;  // tensor-core engine: two mbarriers
  unsigned* tslot;          // tensor-core engine: TMEM base address slot
};

// One K-segment of a GEMM's activation operand.
enum { SEG_SMEM = 0, SEG_X = 1, SEG_STASH = 2, SEG_SMEM_STAGED = 3 };
struct ASeg {
  const float* ptr;   // SEG_SMEM*: tile base in shared memory; SEG_X/SEG_STASH: global, row 0 of the tile
  long long ld;
  int width;
  int kind;
  int wcol;           // first weight column this segment multiplies (== its column in [a || state])
};
__device__ __forceinline__ bool seg_vec_ok(const ASeg& sg) {
  return (sg.kind == SEG_X || sg.kind == SEG_STASH) && ((sg.ld & 3) == 0) && ((sg.width & 3) == 0) &&
         ((reinterpret_cast<size_t>(sg.ptr) & 15) == 0);
}

template <int RM>
struct Cfg {
  static constexpr int TM = 32 * RM;
};

// ---- 32x32 block of a row-major weight matrix -> WB[32][LDX], zero-filled outside (nrows, ncols).
// Split in a load half (global -> registers, issued early) and a store half (registers -> smem).
__device__ __forceinline__ bool w_vec_ok(const float* W, int ldw, int col0) {
  return ((ldw & 3) == 0) && ((col0 & 3) == 0) && ((reinterpret_cast<size_t>(W) & 15) == 0);
}
__device__ __forceinline__ void w_block_load(float (&w)[4], const float* __restrict__ W, int ldw, int row0,
                                             int nrows, int col0, int ncols, bool vec) {
  const int t = threadIdx.x;
  if (vec) {
    const int c4 = (t & 7) * 4, r = t >> 3;
    w[0] = w[1] = w[2] = w[3] = 0.f;
    if (r < nrows) {
      const float* p = W + (long long)(row0 + r) * ldw + col0 + c4;
      if (c4 + 3 < ncols) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(p));
        w[0] = q.x; w[1] = q.y; w[2] = q.z; w[3] = q.w;
      } else {
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (c4 + u < ncols) w[u] = __ldg(p + u);
      }
    }
  } else {
    const int c = t & 31, r0 = t >> 5;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = r0 + 8 * i;
      w[i] = (r < nrows && c < ncols) ? __ldg(W + (long long)(row0 + r) * ldw + col0 + c) : 0.f;
    }
  }
}
__device__ __forceinline__ void w_block_store(float* WB, const float (&w)[4], bool vec) {
  const int t = threadIdx.x;
  if (vec) {
    *reinterpret_cast<float4*>(WB + (t >> 3) * LDX + (t & 7) * 4) = make_float4(w[0], w[1], w[2], w[3]);
  } else {
    const int c = t & 31, r0 = t >> 5;
#pragma unroll
    for (int i = 0; i < 4; ++i) WB[(r0 + 8 * i) * LDX + c] = w[i];
  }
}

// ---- TM x 32 chunk of an activation source -> XB[TM][LDX]: zero-fill, NaN scan + sanitise for x,
// dropout mask when enabled.  Load half: 4*RM values per thread in flight (LDG.128 when aligned).
template <int RM>
__device__ __forceinline__ void a_chunk_load(float (&v)[4 * RM], const ASeg& sg, int k0, int kw, int rows_valid,
                                             bool vec) {
  if (sg.kind != SEG_X && sg.kind != SEG_STASH) return;
  const int t = threadIdx.x;
  if (vec) {
    const int c4 = (t & 7) * 4, r0 = t >> 3;
#pragma unroll
    for (int i = 0; i < RM; ++i) {
      const int r = r0 + 32 * i;
      float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < rows_valid && c4 < kw) {
        const float4* p = reinterpret_cast<const float4*>(sg.ptr + (long long)r * sg.ld + k0 + c4);
        q = sg.kind == SEG_X ? __ldg(p) : __ldcg(p);
      }
      v[4 * i + 0] = q.x; v[4 * i + 1] = q.y; v[4 * i + 2] = q.z; v[4 * i + 3] = q.w;
    }
  } else {
    const int c = t & 31, r0 = t >> 5;
#pragma unroll
    for (int i = 0; i < 4 * RM; ++i) {
      const int r = r0 + 8 * i;
      float x = 0.f;
      if (r < rows_valid && c < kw) {
        const float* p = sg.ptr + (long long)r * sg.ld + k0 + c;
        x = sg.kind == SEG_X ? __ldg(p) : __ldcg(p);   // stash: written earlier by this CTA, L2-coherent load
      }
      v[i] = x;
    }
  }
}
template <int RM>
__device__ __forceinline__ void a_chunk_store(float* XB, float (&v)[4 * RM], const Smem& sm, const ASeg& sg, int k0,
                                              int kw, const Drop& drop, bool scan_nan, bool vec) {
  const int t = threadIdx.x;
  if (vec) {
    const int c4 = (t & 7) * 4, r0 = t >> 3;
#pragma unroll
    for (int i = 0; i < RM; ++i) {
      const int r = r0 + 32 * i;
      if (sg.kind == SEG_X) {
        bool bad = false;
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (v[4 * i + u] != v[4 * i + u]) { bad = true; v[4 * i + u] = 0.f; }   // never let NaN reach arithmetic
        if (bad && scan_nan) sm.rownan[r] = 1;              // NaN marks the modality missing for this row
      }
      if (drop.enabled) {
        const unsigned col = (unsigned)(sg.wcol + k0 + c4), row = drop.row_base + (unsigned)r;
        if ((col & 1u) == 0) {
          const unsigned h0 = mmn_dropout_hash(drop.seed_mix, row, col >> 1);
          const unsigned h1 = mmn_dropout_hash(drop.seed_mix, row, (col >> 1) + 1);
          v[4 * i + 0] = (h0 & 0xffffu) >= drop.thr ? v[4 * i + 0] * drop.scale : 0.f;
          v[4 * i + 1] = (h0 >> 16) >= drop.thr ? v[4 * i + 1] * drop.scale : 0.f;
          v[4 * i + 2] = (h1 & 0xffffu) >= drop.thr ? v[4 * i + 2] * drop.scale : 0.f;
          v[4 * i + 3] = (h1 >> 16) >= drop.thr ? v[4 * i + 3] * drop.scale : 0.f;
        } else {
#pragma unroll
          for (int u = 0; u < 4; ++u)
            v[4 * i + u] = mmn_dropout_keep(drop.seed_mix, row, col + u, drop.thr) ? v[4 * i + u] * drop.scale : 0.f;
        }
      }
      *reinterpret_cast<float4*>(XB + r * LDX + c4) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    }
  } else {
    const int c = t & 31, r0 = t >> 5;
#pragma unroll
    for (int i = 0; i < 4 * RM; ++i) {
      const int r = r0 + 8 * i;
      float x = v[i];
      if (sg.kind == SEG_SMEM_STAGED) x = c < kw ? sg.ptr[(long long)r * sg.ld + k0 + c] : 0.f;
      if (sg.kind == SEG_X && x != x) {
        if (scan_nan) sm.rownan[r] = 1;
        x = 0.f;
      }
      if (drop.enabled && c < kw)
        x = mmn_dropout_keep(drop.seed_mix, drop.row_base + (unsigned)r, (unsigned)(sg.wcol + k0 + c), drop.thr)
                ? x * drop.scale : 0.f;
      XB[r * LDX + c] = x;
    }
  }
}

// iterator over the (segment, k0) chunks of a GEMM's K dimension
struct ChunkIt {
  int s, k0;
  __device__ __forceinline__ bool valid(int nseg) const { return s < nseg; }
  __device__ __forceinline__ void next(const ASeg* segs) {
    k0 += KC;
    if (k0 >= segs[s].width) { ++s; k0 = 0; }
  }
};

// ------------------------------------------------------------------------------------------------
// gemm_nt: out[r][n] = bias[n] + sum over segments sum_k a[r][k] * W[n][wcol + k];  epi(r, n, value)
// for every (r, n) of each 32-column pass, including the zero-pad columns n >= N.
// Chunks are double-buffered: the global loads of chunk c+1 are in flight while chunk c is multiplied
// (one __syncthreads per chunk).
// ------------------------------------------------------------------------------------------------
template <int RM, class Epi>
__device__ __forceinline__ void fma_gemm_nt(const Smem& sm, const float* __restrict__ W, int ldw, int N,
                                        const float* __restrict__ bias, int act, const ASeg* segs, int nseg,
                                        const Drop& drop, int rows_valid, bool scan_nan, Epi epi) {
  constexpr int TM = Cfg<RM>::TM;
  const int tid = threadIdx.x, tx = tid & 7, ty = tid >> 3;
  bool avec[2];
  avec[0] = seg_vec_ok(segs[0]);
  avec[1] = nseg > 1 ? seg_vec_ok(segs[1]) : false;
  for (int n0 = 0; n0 < N; n0 += 32) {
    const int nrows = min(32, N - n0);
    float acc[RM][4], bj[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) bj[j] = (n0 + tx + 8 * j) < N ? __ldg(bias + n0 + tx + 8 * j) : 0.f;
#pragma unroll
    for (int i = 0; i < RM; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    float wr[4], ar[4 * RM];
    ChunkIt it{0, 0};
    bool wv = w_vec_ok(W, ldw, segs[0].wcol);
    w_block_load(wr, W, ldw, n0, nrows, segs[0].wcol, min(KC, segs[0].width), wv);
    a_chunk_load<RM>(ar, segs[0], 0, min(KC, segs[0].width), rows_valid, avec[0]);
    int buf = 0;
    MMN_WSYNC_N(kThreads);                    // previous users of the staging buffers are done
    while (it.valid(nseg)) {
      const ASeg sg = segs[it.s];
      const int kw = min(KC, sg.width - it.k0);
      float* WBb = sm.WB + buf * (32 * LDX);
      float* XBb = sm.XB + buf * (TM * LDX);
      w_block_store(WBb, wr, wv);
      const float* a;
      int lda;
      if (sg.kind == SEG_SMEM) {
        a = sg.ptr + it.k0;
        lda = (int)sg.ld;
      } else {
        a_chunk_store<RM>(XBb, ar, sm, sg, it.k0, kw, drop, scan_nan && n0 == 0, avec[it.s]);
        a = XBb;
        lda = LDX;
      }
      ChunkIt nx = it;
      nx.next(segs);
      if (nx.valid(nseg)) {
        const ASeg& ns = segs[nx.s];
        const int nkw = min(KC, ns.width - nx.k0);
        wv = w_vec_ok(W, ldw, ns.wcol + nx.k0);
        w_block_load(wr, W, ldw, n0, nrows, ns.wcol + nx.k0, nkw, wv);
        a_chunk_load<RM>(ar, ns, nx.k0, nkw, rows_valid, avec[nx.s]);
      }
      MMN_WSYNC_N(kThreads);
      const int nq = (kw + 3) >> 2;
      const float* ap[RM];
#pragma unroll
      for (int i = 0; i < RM; ++i) ap[i] = a + (ty + 32 * i) * lda;
      const float* bp = WBb + tx * LDX;
      auto kstep = [&](int q) {          // q-th float4 of the chunk: RM + 4 LDS.128 feed 16 RM FFMA
        float4 av[RM], bv[4];
#pragma unroll
        for (int i = 0; i < RM; ++i) av[i] = *reinterpret_cast<const float4*>(ap[i] + 4 * q);
#pragma unroll
        for (int j = 0; j < 4; ++j) bv[j] = *reinterpret_cast<const float4*>(bp + 8 * j * LDX + 4 * q);
#pragma unroll
        for (int i = 0; i < RM; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i].x, bv[j].x, acc[i][j]);
#pragma unroll
        for (int i = 0; i < RM; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i].y, bv[j].y, acc[i][j]);
#pragma unroll
        for (int i = 0; i < RM; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i].z, bv[j].z, acc[i][j]);
#pragma unroll
        for (int i = 0; i < RM; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i].w, bv[j].w, acc[i][j]);
      };
      if (RM <= 2 && nq == 8) {          // full chunk: every shared-memory offset is an immediate
#pragma unroll
        for (int q = 0; q < 8; ++q) kstep(q);
      } else {
#pragma unroll 2
        for (int q = 0; q < nq; ++q) kstep(q);
      }
      it = nx;
      buf ^= 1;
    }
#pragma unroll
    for (int i = 0; i < RM; ++i) {
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] += bj[j];
      act_fwd_n(act, acc[i]);
#pragma unroll
      for (int j = 0; j < 4; ++j) epi(ty + 32 * i, n0 + tx + 8 * j, acc[i][j]);
    }
  }
  MMN_WSYNC_N(kThreads);
}

// ------------------------------------------------------------------------------------------------
// gemm_nn: out[r][j] = sum_n dz[r][n] * W[n][col0 + j], j < J;  epi(r, j, acc, pre(r, j)) for every
// (r, j) of each 32-column pass including pad columns j >= J.  dz: shared tile, zero-padded to 32
// columns.  pre(r, j4) -> float4 (columns j4 .. j4+3, j4 % 4 == 0) is evaluated BEFORE the K loop so that
// its global loads (stashed activations) are hidden behind the multiply.
// ------------------------------------------------------------------------------------------------
template <int RM, class Pre, class Epi>
__device__ __forceinline__ void fma_gemm_nn(const Smem& sm, const float* dz, int ldd, int N,
                                        const float* __restrict__ W, int ldw, int col0, int J, Pre pre, Epi epi) {
  const int tid = threadIdx.x, tx = tid & 7, ty = tid >> 3;
  for (int j0 = 0; j0 < J; j0 += 32) {
    const int jw = min(32, J - j0);
    float acc[RM][4];
    float4 pv[RM];
#pragma unroll
    for (int i = 0; i < RM; ++i) {
      pv[i] = pre(ty + 32 * i, j0 + 4 * tx);
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[i][c] = 0.f;
    }
    const bool wv = w_vec_ok(W, ldw, col0 + j0);
    float wr[4];
    w_block_load(wr, W, ldw, 0, min(32, N), col0 + j0, jw, wv);
    int buf = 0;
    MMN_WSYNC_N(kThreads);
    for (int n0 = 0; n0 < N; n0 += 32) {
      const int nw = min(32, N - n0);
      float* WBb = sm.WB + buf * (32 * LDX);
      w_block_store(WBb, wr, wv);
      if (n0 + 32 < N) w_block_load(wr, W, ldw, n0 + 32, min(32, N - n0 - 32), col0 + j0, jw, wv);
      MMN_WSYNC_N(kThreads);
      const int nq = (nw + 3) >> 2;
      const float* ap[RM];
#pragma unroll
      for (int i = 0; i < RM; ++i) ap[i] = dz + (ty + 32 * i) * ldd + n0;
      const float* bp = WBb + 4 * tx;
      auto kstep = [&](int q) {
        float4 av[RM], bv[4];
#pragma unroll
        for (int i = 0; i < RM; ++i) av[i] = *reinterpret_cast<const float4*>(ap[i] + 4 * q);
#pragma unroll
        for (int u = 0; u < 4; ++u) bv[u] = *reinterpret_cast<const float4*>(bp + (4 * q + u) * LDX);
#pragma unroll
        for (int i = 0; i < RM; ++i) {
          acc[i][0] = fmaf(av[i].x, bv[0].x, acc[i][0]); acc[i][1] = fmaf(av[i].x, bv[0].y, acc[i][1]);
          acc[i][2] = fmaf(av[i].x, bv[0].z, acc[i][2]); acc[i][3] = fmaf(av[i].x, bv[0].w, acc[i][3]);
        }
#pragma unroll
        for (int i = 0; i < RM; ++i) {
          acc[i][0] = fmaf(av[i].y, bv[1].x, acc[i][0]); acc[i][1] = fmaf(av[i].y, bv[1].y, acc[i][1]);
          acc[i][2] = fmaf(av[i].y, bv[1].z, acc[i][2]); acc[i][3] = fmaf(av[i].y, bv[1].w, acc[i][3]);
        }
#pragma unroll
        for (int i = 0; i < RM; ++i) {
          acc[i][0] = fmaf(av[i].z, bv[2].x, acc[i][0]); acc[i][1] = fmaf(av[i].z, bv[2].y, acc[i][1]);
          acc[i][2] = fmaf(av[i].z, bv[2].z, acc[i][2]); acc[i][3] = fmaf(av[i].z, bv[2].w, acc[i][3]);
        }
#pragma unroll
        for (int i = 0; i < RM; ++i) {
          acc[i][0] = fmaf(av[i].w, bv[3].x, acc[i][0]); acc[i][1] = fmaf(av[i].w, bv[3].y, acc[i][1]);
          acc[i][2] = fmaf(av[i].w, bv[3].z, acc[i][2]); acc[i][3] = fmaf(av[i].w, bv[3].w, acc[i][3]);
        }
      };
      if (RM <= 2 && nq == 8) {
#pragma unroll
        for (int q = 0; q < 8; ++q) kstep(q);
      } else {
#pragma unroll 2
        for (int q = 0; q < nq; ++q) kstep(q);
      }
      buf ^= 1;
    }
#pragma unroll
    for (int i = 0; i < RM; ++i)
#pragma unroll
      for (int c = 0; c < 4; ++c)
        epi(ty + 32 * i, j0 + 4 * tx + c, acc[i][c], c == 0 ? pv[i].x : c == 1 ? pv[i].y : c == 2 ? pv[i].z : pv[i].w);
  }
  MMN_WSYNC_N(kThreads);
}

// ------------------------------------------------------------------------------------------------
// gemm_tn: gW[n][wcol + k] += sum_r dz[r][n] * a[r][k]   (weight gradient of one K-segment).
// The TM rows are split over kGroups groups of 64 threads; each group owns the whole 32x32 output
// block for its rows (4x4 outputs per thread), the groups' partial blocks are summed through
// shared memory and the tile's contribution is added to global memory with one red per element.
// Input chunks and the reduction scratch are double-buffered: one __syncthreads per 32x32 block.
// ------------------------------------------------------------------------------------------------
template <int RM, int REDBUFS>
__device__ __forceinline__ void fma_gemm_tn(const Smem& sm, const float* dz, int ldd, int N, const ASeg& sg,
                                        const Drop& drop, int rows_valid, float* __restrict__ gW, int ldw) {
  constexpr int TM = Cfg<RM>::TM;
  constexpr int RPG = 8 * RM;     // rows per group
  const int tid = threadIdx.x, g = tid >> 6, u = tid & 63, nt = u >> 3, kt = u & 7;
  const bool vec_ok = ((ldw & 3) == 0) && ((sg.wcol & 3) == 0) && ((reinterpret_cast<size_t>(gW) & 15) == 0);
  const bool avec = seg_vec_ok(sg);
  const int nnb = (N + 31) >> 5;
  float ar[4 * RM];
  a_chunk_load<RM>(ar, sg, 0, min(KC, sg.width), rows_valid, avec);
  MMN_WSYNC_N(kThreads);                 // previous users of XB / RED are done
  int xbuf = 0, rbuf = 0;
  // software pipeline over (chunk, n-block) items: the reduction of item i-1 runs after the sync of item i
  int pend_n0 = -1, pend_k0 = 0, pend_rbuf = 0;
  auto flush = [&](int n0, int k0, int rb) {
    const int n = tid >> 3, kq = tid & 7;    // 256 threads x 4 outputs = the 32x32 block
    const float* red = sm.RED + rb * (kGroups * 1024);
    float4 s = *reinterpret_cast<const float4*>(red + n * 32 + 4 * kq);
#pragma unroll
    for (int gg = 1; gg < kGroups; ++gg) {
      const float4 o = *reinterpret_cast<const float4*>(red + gg * 1024 + n * 32 + 4 * kq);
      s.x += o.x; s.y += o.y; s.z += o.z; s.w += o.w;
    }
    if (n0 + n < N) {
      const int kc = k0 + 4 * kq;
      float* dst = gW + (long long)(n0 + n) * ldw + sg.wcol + kc;
      if (vec_ok && kc + 3 < sg.width) {
        atomicAdd(reinterpret_cast<float4*>(dst), s);
      } else {
        if (kc + 0 < sg.width) atomicAdd(dst + 0, s.x);
        if (kc + 1 < sg.width) atomicAdd(dst + 1, s.y);
        if (kc + 2 < sg.width) atomicAdd(dst + 2, s.z);
        if (kc + 3 < sg.width) atomicAdd(dst + 3, s.w);
      }
    }
  };
  for (int k0 = 0; k0 < sg.width; k0 += KC) {
    const int kw = min(KC, sg.width - k0);
    const float* a;
    int lda;
    if (sg.kind == SEG_SMEM) {
      a = sg.ptr + k0;
      lda = (int)sg.ld;
    } else {
      float* XBb = sm.XB + xbuf * (TM * LDX);
      a_chunk_store<RM>(XBb, ar, sm, sg, k0, kw, drop, false, avec);
      a = XBb;
      lda = LDX;
      xbuf ^= 1;
    }
    if (k0 + KC < sg.width) a_chunk_load<RM>(ar, sg, k0 + KC, min(KC, sg.width - k0 - KC), rows_valid, avec);
    for (int nb = 0; nb < nnb; ++nb) {
      const int n0 = nb * 32;
      MMN_WSYNC_N(kThreads);             // chunk visible; RED of the pending item complete
      if (pend_n0 >= 0) flush(pend_n0, pend_k0, pend_rbuf);
      float acc[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
      const float* dp = dz + g * RPG * ldd + n0 + 4 * nt;
      const float* ip = a + g * RPG * lda + 4 * kt;
#pragma unroll (RM <= 2 ? RPG : 4)
      for (int rr = 0; rr < RPG; ++rr) {
        const float4 dv = *reinterpret_cast<const float4*>(dp + rr * ldd);
        const float4 iv = *reinterpret_cast<const float4*>(ip + rr * lda);
        acc[0][0] = fmaf(dv.x, iv.x, acc[0][0]); acc[0][1] = fmaf(dv.x, iv.y, acc[0][1]);
        acc[0][2] = fmaf(dv.x, iv.z, acc[0][2]); acc[0][3] = fmaf(dv.x, iv.w, acc[0][3]);
        acc[1][0] = fmaf(dv.y, iv.x, acc[1][0]); acc[1][1] = fmaf(dv.y, iv.y, acc[1][1]);
        acc[1][2] = fmaf(dv.y, iv.z, acc[1][2]); acc[1][3] = fmaf(dv.y, iv.w, acc[1][3]);
        acc[2][0] = fmaf(dv.z, iv.x, acc[2][0]); acc[2][1] = fmaf(dv.z, iv.y, acc[2][1]);
        acc[2][2] = fmaf(dv.z, iv.z, acc[2][2]); acc[2][3] = fmaf(dv.z, iv.w, acc[2][3]);
        acc[3][0] = fmaf(dv.w, iv.x, acc[3][0]); acc[3][1] = fmaf(dv.w, iv.y, acc[3][1]);
        acc[3][2] = fmaf(dv.w, iv.z, acc[3][2]); acc[3][3] = fmaf(dv.w, iv.w, acc[3][3]);
      }
      if (REDBUFS == 1) MMN_WSYNC_N(kThreads);      // single scratch: every reader of the pending block is done
      float* red = sm.RED + rbuf * (kGroups * 1024) + g * 1024;
#pragma unroll
      for (int i = 0; i < 4; ++i)
        *reinterpret_cast<float4*>(red + (4 * nt + i) * 32 + 4 * kt) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
      pend_n0 = n0; pend_k0 = k0; pend_rbuf = rbuf;
      if (REDBUFS == 2) rbuf ^= 1;
    }
  }
  MMN_WSYNC_N(kThreads);
  if (pend_n0 >= 0) flush(pend_n0, pend_k0, pend_rbuf);
}

// ------------------------------------------------------------------------------------------------
// FMA engine: the three GEMM shapes on the FP32 pipe (any tile height RM in {1, 2, 4})
// ------------------------------------------------------------------------------------------------
struct NoState { long long t[16]; };
// MINB = CTAs per SM the kernel is compiled for (2 halves the register budget and single-buffers the
// weight-gradient scratch so that two 64-row tiles share an SM: 16 resident warps instead of 8)
template <int RM_, int MINB_ = 1>
struct FmaEngine {
  static constexpr int RM = RM_;
  static constexpr int TM = 32 * RM_;
  static constexpr bool kTensor = false;
  static constexpr int kWorkers = kThreads;
  static constexpr int kBlockThreads = kThreads;
  static constexpr int kMinBlocks = MINB_;
  static constexpr int kRedBufs = MINB_ >= 2 ? 1 : 2;
  using State = NoState;
  __device__ static __forceinline__ void issuer_loop(const Smem&, State&, long long* = nullptr) {}
  static size_t stage_bytes() { return (size_t)(2 * (TM * LDX + 32 * LDX) + kRedBufs * kGroups * 1024) * 4; }
  __device__ static __forceinline__ char* carve(Smem& sm, char* p) {
    float* f = reinterpret_cast<float*>(p);
    sm.XB = f; f += 2 * TM * LDX;
    sm.WB = f; f += 2 * 32 * LDX;
    sm.RED = f; f += kRedBufs * kGroups * 1024;
    return reinterpret_cast<char*>(f);
  }
  __device__ static __forceinline__ void init(const Smem&, State&) {}
  __device__ static __forceinline__ void fini(const Smem&, State&) {}
  template <class Epi>
  __device__ static __forceinline__ void gemm_nt(const Smem& sm, State&, const float* __restrict__ W, int ldw, int N,
                                                 const float* __restrict__ bias, int act, const ASeg* segs, int nseg,
                                                 const Drop& drop, int rows_valid, bool scan_nan, Epi epi) {
    fma_gemm_nt<RM>(sm, W, ldw, N, bias, act, segs, nseg, drop, rows_valid, scan_nan, epi);
  }
  template <class Pre, class Epi>
  __device__ static __forceinline__ void gemm_nn(const Smem& sm, State&, const float* dz, int ldd, int N,
                                                 const float* __restrict__ W, int ldw, int col0, int J, Pre pre, Epi epi) {
    fma_gemm_nn<RM>(sm, dz, ldd, N, W, ldw, col0, J, pre, epi);
  }
  __device__ static __forceinline__ void gemm_tn(const Smem& sm, State&, const float* dz, int ldd, int N, const ASeg& sg,
                                                 const Drop& drop, int rows_valid, float* __restrict__ gW, int ldw) {
    fma_gemm_tn<RM, kRedBufs>(sm, dz, ldd, N, sg, drop, rows_valid, gW, ldw);
  }
};

// bias gradient: gb[n] += sum_r dz[r][n]
template <int TM, int NT>
__device__ __forceinline__ void colsum_red(const Smem& sm, const float* dz, int ldd, int N, float* __restrict__ gb) {
  constexpr int PARTS = NT / 32;
  const int tid = threadIdx.x, c = tid & 31, part = tid >> 5;
  for (int n0 = 0; n0 < N; n0 += 32) {
    float s = 0.f;
    for (int r = part; r < TM; r += PARTS) s += dz[r * ldd + n0 + c];
    MMN_WSYNC_N(NT);
    sm.RED[part * 32 + c] = s;
    MMN_WSYNC_N(NT);
    if (tid < 32 && n0 + tid < N) {
      float tot = 0.f;
#pragma unroll
      for (int p = 0; p < PARTS; ++p) tot += sm.RED[p * 32 + tid];
      atomicAdd(gb + n0 + tid, tot);
    }
  }
  MMN_WSYNC_N(NT);
}

// walks idx = tid, tid + kThreads, ... of a [rows x w] index space as (r, c) without a division per step
template <int NT>
struct RowCol {
  int r, c, q, rem, w;
  __device__ __forceinline__ RowCol(int w_) : w(w_) {
    r = (int)threadIdx.x / w_; c = (int)threadIdx.x - r * w_;
    q = NT / w_; rem = NT - q * w_;
  }
  __device__ __forceinline__ void next() {
    r += q; c += rem;
    if (c >= w) { c -= w; ++r; }
  }
};

// shared tile [TM x width] -> global row-major [TM x width] (float4 when the width allows)
template <int TM, int NT>
__device__ __forceinline__ void stash_store(float* dst, const float* buf, int ld, int width) {
  if ((width & 3) == 0) {
    const int w4 = width >> 2;
    for (RowCol<NT> it(w4); it.r < TM; it.next())
      __stcg(reinterpret_cast<float4*>(dst) + it.r * w4 + it.c,
             *reinterpret_cast<const float4*>(buf + it.r * ld + 4 * it.c));
  } else {
    for (RowCol<NT> it(width); it.r < TM; it.next()) __stcg(dst + it.r * width + it.c, buf[it.r * ld + it.c]);
  }
}

// 4 consecutive columns (c4 % 4 == 0) of row r of a row-major [rows x width] global block, L2-coherent;
// zero beyond width
__device__ __forceinline__ float4 ldcg4(const float* base, int r, int width, int c4) {
  const float* p = base + (long long)r * width + c4;
  if (((width & 3) == 0) && ((reinterpret_cast<size_t>(base) & 15) == 0) && c4 + 3 < width)
    return __ldcg(reinterpret_cast<const float4*>(p));
  float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c4 + 0 < width) q.x = __ldcg(p + 0);
  if (c4 + 1 < width) q.y = __ldcg(p + 1);
  if (c4 + 2 < width) q.z = __ldcg(p + 2);
  if (c4 + 3 < width) q.w = __ldcg(p + 3);
  return q;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ unsigned warp_sum_u(unsigned v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------------------------------------
// the step kernel
// ------------------------------------------------------------------------------------------------
template <class ENG, bool TRAIN>
__global__ void __launch_bounds__(ENG::kBlockThreads, ENG::kMinBlocks) mmn_step_kernel(const StepArgs args) {
  constexpr int RM = ENG::RM;
  constexpr int TM = ENG::TM;
  constexpr int NT = ENG::kWorkers;
  const DevPlan& P = *args.plan;
  const int tid = threadIdx.x;
  const int S = P.S, E = P.E, D = P.D, ldS = P.ldS, ldH = P.ldH;
  const int L = args.seq_len;
  const float* __restrict__ params = args.params;

  MMN_DYN_SMEM(smem_raw);
  Smem sm;
  typename ENG::State es;
  {
    float* f = reinterpret_cast<float*>(ENG::carve(sm, smem_raw));   // engine staging first (alignment), then tiles
    sm.S = f; f += TM * ldS;
    sm.T = f; f += TM * ldS;
    sm.A = f; f += TM * ldH;
    sm.B = f; f += TM * ldH;
    sm.ys = reinterpret_cast<int*>(f); f += TM * D;
    sm.rownan = reinterpret_cast<int*>(f); f += TM;
    sm.cnt = reinterpret_cast<int*>(f); f += (E + 1);
    sm.tile_any = reinterpret_cast<int*>(f); f += (E + 1);
    size_t off = (reinterpret_cast<char*>(f) - smem_raw + 7) & ~(size_t)7;
    sm.met = reinterpret_cast<double*>(smem_raw + off);
    sm.present = reinterpret_cast<unsigned char*>(sm.met + P.n_metrics);
  }
  for (int i = tid; i < P.n_metrics; i += NT) sm.met[i] = 0.0;
  for (int i = tid; i < E + 1; i += NT) sm.cnt[i] = 0;
  ENG::init(sm, es);
  const long long t_kernel = MMN_CLOCK();

  const long long n_tiles = (args.n_rows + TM - 1) / TM;
  float* slot = TRAIN ? args.stash + (long long)blockIdx.x * args.slot_floats : nullptr;
  Drop nodrop;
  nodrop.enabled = 0; nodrop.seed_mix = 0; nodrop.thr = 0; nodrop.row_base = 0; nodrop.scale = 1.f;

  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long row0 = tile * TM;
    const int rows_valid = (int)min((long long)TM, args.n_rows - row0);
    MMN_WSYNC_N(NT);

    // ---- tile prologue: targets, initial state (state.py:29-32), masks ----
    if (args.targets) {
      for (int idx = tid; idx < TM * D; idx += NT) {
        const int r = idx / D;
        long long y = 0;
        if (r < rows_valid) y = args.targets[(row0 + r) * D + (idx - r * D)];
        if ((y < 0 || y >= P.dec[idx - r * D].C) && args.target_error) *args.target_error = 1;   // CrossEntropyLoss would raise
        sm.ys[idx] = (int)y;
      }
    }
    for (int idx = tid; idx < TM * ldS; idx += NT) {
      const int c = idx % ldS;
      sm.S[idx] = c < S ? __ldg(params + P.init_off + c) : 0.f;
    }
    for (int r = tid; r < TM; r += NT) sm.present[r] = r < rows_valid;
    for (int i = tid; i < E + 1; i += NT) sm.tile_any[i] = i == 0;
    if (tid < TM && tid < rows_valid) atomicAdd(&sm.cnt[0], 1);
    MMN_WSYNC_N(NT);
    if (TRAIN) stash_store<TM, NT>(slot + (long long)stash_state_off(P, 0) * TM, sm.S, ldS, S);

    // ---- decoders on the current state (multimodn.py:141-157, 176-191) ----
    auto decoders_forward = [&](int k, int hist_row, bool is_last_enc) {
      for (int d = 0; d < D; ++d) {
        const DevDecoder& dec = P.dec[d];
        const float* in = sm.S;
        int ldin = ldS;
        for (int j = 0; j < dec.n_layers; ++j) {
          const DevLayer& ly = dec.L[j];
          float* out = (j & 1) ? sm.B : sm.A;
          ASeg seg;
          seg.ptr = in; seg.ld = ldin; seg.width = ly.in_dim; seg.kind = SEG_SMEM; seg.wcol = 0;
          const int N = ly.out_dim, act = ly.act;
          ENG::gemm_nt(sm, es, params + ly.w_off, ly.ktot, N, params + ly.b_off, act, &seg, 1, nodrop, rows_valid, false,
                      [&](int r, int n, float z) { out[r * ldH + n] = n < N ? z : 0.f; });
          if (TRAIN)
            stash_store<TM, NT>(slot + (long long)(stash_dec_off(P, k) + dec.stash_off + ly.stash_off) * TM, out, ldH, N);
          in = out;
          ldin = ldH;
        }
        // per-row epilogue: first-max arg-max, CE on the outputs, confusion cells
        const int C = dec.C;
        if (tid < TM) {                       // whole warps: TM is a multiple of 32
          const int r = tid;
          const float* p = in + r * ldH;
          const bool m = sm.present[k * TM + r] != 0;
          float best = p[0];
          int pred = 0;
          for (int c = 1; c < C; ++c) {
            const float v = p[c];
            if (v > best || (v != v && best == best)) { best = v; pred = c; }   // NaN wins, like torch.max
          }
          if (r < rows_valid) {
            if (args.predictions) args.predictions[((long long)hist_row * D + d) * args.pred_ld + row0 + r] = (unsigned char)pred;
            if (args.last_outputs && is_last_enc) {
              float* o = args.last_outputs + (row0 + r) * P.sumC + dec.out_off;
              for (int c = 0; c < C; ++c) o[c] = p[c];
            }
          }
          if (args.targets) {
            int y = sm.ys[r * D + d];
            y = y < 0 ? 0 : (y >= C ? C - 1 : y);
            float mx = p[0];
            for (int c = 1; c < C; ++c) mx = fmaxf(mx, p[c]);
            float se = 0.f;
            for (int c = 0; c < C; ++c) se += expf(p[c] - mx);
            float ce = m ? (mx + logf(se) - p[y]) : 0.f;
            unsigned pk1 = 0, pk2 = 0;
            if (m) {
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
            if ((tid & 31) == 0) {
              atomicAdd(&sm.met[met_mat(P, 0, hist_row, d)], (double)ce);
              atomicAdd(&sm.met[met_mat(P, 1, hist_row, d)], (double)(pk1 & 0xff));
              atomicAdd(&sm.met[met_mat(P, 2, hist_row, d)], (double)((pk1 >> 8) & 0xff));
              atomicAdd(&sm.met[met_mat(P, 3, hist_row, d)], (double)((pk1 >> 16) & 0xff));
              atomicAdd(&sm.met[met_mat(P, 4, hist_row, d)], (double)((pk1 >> 24) & 0xff));
              atomicAdd(&sm.met[met_mat(P, 5, hist_row, d)], (double)(pk2 & 0xff));
            }
          }
        }
        MMN_WSYNC_N(NT);
      }
    };
    decoders_forward(0, 0, false);

    // ---- walk the encoding sequence (multimodn.py:159-191) ----
    for (int k = 1; k <= L; ++k) {
      const int e = args.seq_enc[k - 1], pos = args.seq_pos[k - 1];
      const DevEncoder& enc = P.enc[e];
      const bool skip = args.skip_flags && args.skip_flags[k - 1] != 0;   // reference batch-level rule
      for (int r = tid; r < TM; r += NT) sm.rownan[r] = 0;
      MMN_WSYNC_N(NT);
      if (!skip) {
        Drop drop = nodrop;
        if (TRAIN && args.training && enc.p_drop > 0.f && enc.L[0].has_state) {
          drop.enabled = 1;
          drop.seed_mix = args.dropout_seed ^ ((unsigned)e * 0x9E3779B9u);
          drop.thr = (unsigned)(enc.p_drop * 65536.f);
          drop.row_base = (unsigned)(args.row_offset + row0);
          drop.scale = 1.f / (1.f - enc.p_drop);
        }
        const float* in = nullptr;
        int ldin = 0;
        for (int j = 0; j < enc.n_layers; ++j) {
          const DevLayer& ly = enc.L[j];
          const bool last = j == enc.n_layers - 1;
          float* out = last ? sm.T : ((j & 1) ? sm.B : sm.A);
          const int ldo = last ? ldS : ldH;
          ASeg segs[2];
          int nseg = 0;
          if (j == 0) {
            segs[0].ptr = args.x[pos] + row0 * args.x_ld[pos];
            segs[0].ld = args.x_ld[pos]; segs[0].width = ly.in_dim; segs[0].kind = SEG_X; segs[0].wcol = 0;
          } else {
            segs[0].ptr = in; segs[0].ld = ldin; segs[0].width = ly.in_dim; segs[0].kind = SEG_SMEM; segs[0].wcol = 0;
          }
          nseg = 1;
          const bool use_drop = drop.enabled && j == 0;
          if (ly.has_state) {
            segs[1].ptr = sm.S; segs[1].ld = ldS; segs[1].width = S;
            segs[1].kind = use_drop ? SEG_SMEM_STAGED : SEG_SMEM;
            segs[1].wcol = ly.in_dim;
            nseg = 2;
          }
          const int N = ly.out_dim, act = ly.act;
          ENG::gemm_nt(sm, es, params + ly.w_off, ly.ktot, N, params + ly.b_off, act, segs, nseg, use_drop ? drop : nodrop,
                      rows_valid, j == 0,
                      [&](int r, int n, float z) { out[r * ldo + n] = n < N ? z : 0.f; });
          if (TRAIN && !last)
            stash_store<TM, NT>(slot + (long long)(stash_enc_off(P, k) + ly.stash_off) * TM, out, ldH, N);
          in = out;
          ldin = ldo;
        }
      }
      // per-row select (missing rows keep their state bit for bit) + state-change sum (multimodn.py:174)
      if (tid < TM) {
        const bool pr = !skip && tid < rows_valid && sm.rownan[tid] == 0;
        sm.present[k * TM + tid] = pr;
        if (pr) { atomicAdd(&sm.cnt[e + 1], 1); sm.tile_any[k] = 1; }
      }
      MMN_WSYNC_N(NT);
      const int any = sm.tile_any[k];
      if (any) {
        float sc = 0.f;
        for (RowCol<NT> it(S); it.r < TM; it.next()) {
          const int r = it.r, c = it.c;
          if (sm.present[k * TM + r]) {
            const float o = sm.S[r * ldS + c], nw = sm.T[r * ldS + c], df = nw - o;
            sc = fmaf(df, df, sc);
            sm.S[r * ldS + c] = nw;
          }
        }
        sc = warp_sum(sc);
        if ((tid & 31) == 0 && TRAIN) atomicAdd(&sm.met[met_sc(P, e)], (double)sc);
      }
      MMN_WSYNC_N(NT);
      if (TRAIN) stash_store<TM, NT>(slot + (long long)stash_state_off(P, k) * TM, sm.S, ldS, S);
      decoders_forward(k, e + 1, e == E - 1);
    }
    if (args.final_state) {
      for (RowCol<NT> it(S); it.r < rows_valid; it.next())
        args.final_state[(row0 + it.r) * S + it.c] = sm.S[it.r * ldS + it.c];
    }

    // =============================================================================================
    // backward: replay the sequence in reverse.  G (aliasing the state tile)
    // holds dLoss/ds_k for the tile.
    // =============================================================================================
    if (TRAIN) {
      MMN_WSYNC_N(NT);
      const long long t_bwd = MMN_CLOCK();
      float* G = sm.S;
      float* grads = args.grads;
      for (int idx = tid; idx < TM * ldS; idx += NT) G[idx] = 0.f;
      MMN_WSYNC_N(NT);

      auto decoders_backward = [&](int k) {
        const float* sk = slot + (long long)stash_state_off(P, k) * TM;
        for (int d = 0; d < D; ++d) {
          const DevDecoder& dec = P.dec[d];
          const int C = dec.C, nl = dec.n_layers;
          const long long dbase = (long long)(stash_dec_off(P, k) + dec.stash_off) * TM;
          // dz of the last layer from the stashed outputs p: CE-on-outputs gradient through out_act
          {
            const float* pst = slot + dbase + (long long)dec.L[nl - 1].stash_off * TM;
            const ActBwd dact(dec.L[nl - 1].act);
            const int Cpad = (C + 31) & ~31;
            if (tid < TM) {
              const int r = tid;
              const bool m = sm.present[k * TM + r] != 0;
              float p[MMN_MAX_CLASSES];
              float mx = -3.4e38f;
              for (int c = 0; c < C; ++c) { p[c] = __ldcg(pst + r * C + c); mx = fmaxf(mx, p[c]); }
              float se = 0.f;
              for (int c = 0; c < C; ++c) se += expf(p[c] - mx);
              int y = sm.ys[r * D + d];
              y = y < 0 ? 0 : (y >= C ? C - 1 : y);
              const float coef = m ? args.c_err : 0.f;
              const float inv = 1.f / se;
              for (int c = 0; c < Cpad; ++c) {
                float v = 0.f;
                if (c < C) v = coef * (expf(p[c] - mx) * inv - (c == y ? 1.f : 0.f)) * dact(p[c]);
                sm.A[r * ldH + c] = v;
              }
            }
            MMN_WSYNC_N(NT);
          }
          float* cur = sm.A;
          for (int j = nl - 1; j >= 0; --j) {
            const DevLayer& ly = dec.L[j];
            colsum_red<TM, NT>(sm, cur, ldH, ly.out_dim, grads + ly.b_off);
            ASeg seg;
            seg.kind = SEG_STASH; seg.wcol = 0; seg.width = ly.in_dim; seg.ld = ly.in_dim;
            seg.ptr = j == 0 ? sk : slot + dbase + (long long)dec.L[j - 1].stash_off * TM;
            ENG::gemm_tn(sm, es, cur, ldH, ly.out_dim, seg, nodrop, TM, grads + ly.w_off, ly.ktot);
            if (j > 0) {
              float* other = cur == sm.A ? sm.B : sm.A;
              const float* ast = slot + dbase + (long long)dec.L[j - 1].stash_off * TM;
              const int J = ly.in_dim;
              const ActBwd dact(dec.L[j - 1].act);
              ENG::gemm_nn(sm, es, cur, ldH, ly.out_dim, params + ly.w_off, ly.ktot, 0, J,
                          [&](int r, int jc) { return ldcg4(ast, r, J, jc); },
                          [&](int r, int jc, float acc, float a) {
                            other[r * ldH + jc] = jc < J ? acc * dact(a) : 0.f;
                          });
              cur = other;
            } else {
              ENG::gemm_nn(sm, es, cur, ldH, ly.out_dim, params + ly.w_off, ly.ktot, 0, S,
                          [&](int, int) { return make_float4(0.f, 0.f, 0.f, 0.f); },
                          [&](int r, int jc, float acc, float) {
                            if (jc < S) G[r * ldS + jc] += acc;
                          });
            }
          }
        }
      };

      for (int k = L; k >= 1; --k) {
        if (!sm.tile_any[k]) continue;     // no row of this tile took the step: s_k == s_{k-1}, nothing flows
        const int e = args.seq_enc[k - 1], pos = args.seq_pos[k - 1];
        const DevEncoder& enc = P.enc[e];
        decoders_backward(k);
        const float* sk = slot + (long long)stash_state_off(P, k) * TM;
        const float* skm1 = slot + (long long)stash_state_off(P, k - 1) * TM;
        const int nl = enc.n_layers;
        // G += u_k ; dz_last = present ? G * act'(s_k) : 0
        {
          const ActBwd dact(enc.L[nl - 1].act);
          const int Spad = (S + 31) & ~31;
          for (RowCol<NT> it(Spad); it.r < TM; it.next()) {
            const int r = it.r, c = it.c;
            float dzv = 0.f;
            if (c < S) {
              const float a = __ldcg(sk + r * S + c), b = __ldcg(skm1 + r * S + c);
              const float g = G[r * ldS + c] + args.c_sc * (a - b);
              G[r * ldS + c] = g;
              if (sm.present[k * TM + r]) dzv = g * dact(a);
            }
            sm.T[r * ldS + c] = dzv;
          }
          MMN_WSYNC_N(NT);
        }
        Drop drop = nodrop;
        if (args.training && enc.p_drop > 0.f && enc.L[0].has_state) {
          drop.enabled = 1;
          drop.seed_mix = args.dropout_seed ^ ((unsigned)e * 0x9E3779B9u);
          drop.thr = (unsigned)(enc.p_drop * 65536.f);
          drop.row_base = (unsigned)(args.row_offset + row0);
          drop.scale = 1.f / (1.f - enc.p_drop);
        }
        float* cur = sm.T;
        int ldc = ldS;
        const long long ebase = (long long)stash_enc_off(P, k) * TM;
        for (int j = nl - 1; j >= 0; --j) {
          const DevLayer& ly = enc.L[j];
          const bool use_drop = drop.enabled && j == 0;
          colsum_red<TM, NT>(sm, cur, ldc, ly.out_dim, grads + ly.b_off);
          ASeg seg;
          seg.wcol = 0; seg.width = ly.in_dim;
          if (j == 0) {
            seg.kind = SEG_X; seg.ptr = args.x[pos] + row0 * args.x_ld[pos]; seg.ld = args.x_ld[pos];
          } else {
            seg.kind = SEG_STASH; seg.ptr = slot + ebase + (long long)enc.L[j - 1].stash_off * TM; seg.ld = ly.in_dim;
          }
          ENG::gemm_tn(sm, es, cur, ldc, ly.out_dim, seg, use_drop ? drop : nodrop, rows_valid, grads + ly.w_off, ly.ktot);
          if (ly.has_state) {
            ASeg s2;
            s2.kind = SEG_STASH; s2.ptr = skm1; s2.ld = S; s2.width = S; s2.wcol = ly.in_dim;
            ENG::gemm_tn(sm, es, cur, ldc, ly.out_dim, s2, use_drop ? drop : nodrop, TM, grads + ly.w_off, ly.ktot);
          }
          float* other = nullptr;
          if (j > 0) {
            other = cur == sm.A ? sm.B : sm.A;
            const float* ast = slot + ebase + (long long)enc.L[j - 1].stash_off * TM;
            const int J = ly.in_dim;
            const ActBwd dact(enc.L[j - 1].act);
            ENG::gemm_nn(sm, es, cur, ldc, ly.out_dim, params + ly.w_off, ly.ktot, 0, J,
                        [&](int r, int jc) { return ldcg4(ast, r, J, jc); },
                        [&](int r, int jc, float acc, float a) {
                          other[r * ldH + jc] = jc < J ? acc * dact(a) : 0.f;
                        });
          }
          if (ly.has_state) {
            // carry into G: present rows take dz W_s (through the dropout mask), absent rows keep G;
            // then remove u_k, which belongs to s_{k-1} with the opposite sign (multimodn.py:165,174)
            const int in_dim = ly.in_dim;
            ENG::gemm_nn(sm, es, cur, ldc, ly.out_dim, params + ly.w_off, ly.ktot, in_dim, S,
                        [&](int r, int jc) {      // u_k, loaded before the multiply
                          const float4 a = ldcg4(sk, r, S, jc), b = ldcg4(skm1, r, S, jc);
                          return make_float4(args.c_sc * (a.x - b.x), args.c_sc * (a.y - b.y), args.c_sc * (a.z - b.z),
                                             args.c_sc * (a.w - b.w));
                        },
                        [&](int r, int jc, float acc, float u) {
                          if (jc < S) {
                            float carry = acc;
                            if (use_drop)
                              carry = mmn_dropout_keep(drop.seed_mix, drop.row_base + (unsigned)r, (unsigned)(in_dim + jc), drop.thr)
                                          ? carry * drop.scale : 0.f;
                            const float g = sm.present[k * TM + r] ? carry : G[r * ldS + jc];
                            G[r * ldS + jc] = g - u;
                          }
                        });
          }
          if (j > 0) { cur = other; ldc = ldH; }
        }
      }
      decoders_backward(0);
      colsum_red<TM, NT>(sm, G, ldS, S, grads + P.init_off);      // tile backward of state.py:30
      es.t[9] += MMN_CLOCK() - t_bwd;
    }
  }

  // ---- flush this CTA's metric partials ----
  MMN_WSYNC_N(NT);
  ENG::fini(sm, es);
  if (args.debug_timers && tid == 0) {
    es.t[15] = MMN_CLOCK() - t_kernel;
    for (int i = 0; i < 16; ++i) args.debug_timers[blockIdx.x * 16 + i] = es.t[i];
  }
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
  if (TRAIN && args.grads) {
    for (int e = tid; e < E; e += NT)
      if (sm.cnt[e + 1]) atomicAdd(args.grads + P.n_params + e, (float)sm.cnt[e + 1]);
  }
}

// ------------------------------------------------------------------------------------------------
// any(isnan) per modality tensor — the reference's batch-level missingness test (multimodn.py:168)
// ------------------------------------------------------------------------------------------------
struct ScanArgs {
  int seq_len;
  long long n_rows;
  int F[MMN_MAX_ENCODERS];
  const float* x[MMN_MAX_ENCODERS];
  long long x_ld[MMN_MAX_ENCODERS];
  int* flags;
};
template <int = 0>
__global__ void __launch_bounds__(256) mmn_scan_missing_kernel(const ScanArgs a) {
  for (int k = 0; k < a.seq_len; ++k) {
    const long long n = a.n_rows * a.F[k];
    int found = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
      const long long r = i / a.F[k];
      const float v = __ldg(a.x[k] + r * a.x_ld[k] + (i - r * a.F[k]));
      if (v != v) found = 1;
    }
    if (__syncthreads_or(found) && threadIdx.x == 0) a.flags[k] = 1;
  }
}

// ------------------------------------------------------------------------------------------------
// Adam on the packed buffers (torch.optim.Adam defaults; multimodn.py:204)
// ------------------------------------------------------------------------------------------------
template <int = 0>
__global__ void mmn_adam_tick_kernel(const DevPlan* plan, const float* grads, int* step_count) {
  const int i = threadIdx.x;
  if (i == 0) step_count[0] += 1;
  if (i >= 1 && i <= plan->E && grads[plan->n_params + (i - 1)] > 0.f) step_count[i] += 1;
}
template <int = 0>
__global__ void __launch_bounds__(256) mmn_adam_kernel(const DevPlan* plan, float* params, const float* grads,
                                                        float* m, float* v, const int* step_count, float lr,
                                                        float b1, float b2, float eps) {
  const DevPlan& P = *plan;
  // per-owner constants (owner 0 = decoders + initial state, owner e + 1 = encoder e): bias corrections in double as
  // torch.optim.Adam computes them, once per block instead of once per element
  __shared__ float s_step[MMN_MAX_ENCODERS + 1], s_sqrt_bc2[MMN_MAX_ENCODERS + 1];
  __shared__ int s_live[MMN_MAX_ENCODERS + 1];
  if ((int)threadIdx.x <= P.E) {
    const int owner = threadIdx.x;
    const int t = step_count[owner];
    const double bc1 = 1.0 - pow((double)b1, (double)t), bc2 = 1.0 - pow((double)b2, (double)t);
    s_step[owner] = (float)((double)lr / bc1);
    s_sqrt_bc2[owner] = (float)sqrt(bc2);
    s_live[owner] = owner == 0 || grads[P.n_params + owner - 1] > 0.f;    // `.grad is None`: untouched
  }
  __syncthreads();
  auto owner_of = [&](long long i) {
    int owner = 0;
    for (int e = 0; e < P.E; ++e)
      if (i >= P.enc[e].param_lo && i < P.enc[e].param_hi) owner = e + 1;
    return owner;
  };
  auto update = [&](float& p, float g, float& mi, float& vi, int owner) {
    mi = mi + (g - mi) * (1.f - b1);
    vi = vi * b2 + g * g * (1.f - b2);
    const float denom = sqrtf(vi) / s_sqrt_bc2[owner] + eps;
    p = p - s_step[owner] * (mi / denom);
  };
  // blocks of four: parameter blocks start on multiples of four floats, so a quad never spans two owners
  // (the floats between a block's end and the next multiple of four are padding with zero gradient)
  const long long n4 = P.n_params >> 2;
  const bool vec = (((size_t)params | (size_t)grads | (size_t)m | (size_t)v) & 15) == 0;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, stride = (long long)gridDim.x * blockDim.x;
  if (vec) {
    for (long long q = tid; q < n4; q += stride) {
      const int owner = owner_of(q << 2);
      if (!s_live[owner]) continue;
      float4 p4 = reinterpret_cast<float4*>(params)[q], m4 = reinterpret_cast<float4*>(m)[q], v4 = reinterpret_cast<float4*>(v)[q];
      const float4 g4 = reinterpret_cast<const float4*>(grads)[q];
      update(p4.x, g4.x, m4.x, v4.x, owner);
      update(p4.y, g4.y, m4.y, v4.y, owner);
      update(p4.z, g4.z, m4.z, v4.z, owner);
      update(p4.w, g4.w, m4.w, v4.w, owner);
      reinterpret_cast<float4*>(params)[q] = p4;
      reinterpret_cast<float4*>(m)[q] = m4;
      reinterpret_cast<float4*>(v)[q] = v4;
    }
  }
  for (long long i = (vec ? (n4 << 2) : 0) + tid; i < P.n_params; i += stride) {
    const int owner = owner_of(i);
    if (!s_live[owner]) continue;
    float pi = params[i], mi = m[i], vi = v[i];
    update(pi, grads[i], mi, vi, owner);
    params[i] = pi; m[i] = mi; v[i] = vi;
  }
}

}  // namespace mmn
