// mmn_wide_step.cuh — the wide regime's step: the sequential-fusion chain run layer by layer on the tcgen05 bf16 GEMM
// of mmn_wide.cuh (precision = bf16: bf16 weights / activations / layer gradients, fp32 accumulation, fp32 master
// weights, fp32 state gradient and parameter gradients).  Included by mmn_api.cu (needs fail(), MMN_CUDA, mmn_plan).
//
// One GEMM kernel serves the three contractions; operands are K-major ([rows x K], K contiguous) or MN-major (a row-major
// matrix contracted over its ROWS, fetched as 64 x 64 TMA boxes), so no activation is ever transposed in memory:
//   forward   Y[B x out]   = In[B x k]      . W[out x k]^T            A, B K-major
//   dgrad     dIn[B x in]  = dZ[B x out]    . W^T[in x out]^T         A, B K-major (W^T: a second bf16 copy of the weights)
//   wgrad     dW[out x k]  = sum_r dZ[r][out] In[r][k]                A = dZ, B = In, both MN-major
// The GEMM epilogues fuse bias + activation, the per-row missingness select + state-change sum, act' of the data
// gradient, the fp32 accumulation of weight / state gradients and the state carry through an encoder.
#pragma once

#include "mmn_wide.cuh"

namespace mmn {
namespace wide {

typedef __nv_bfloat16 bf16;

struct Mat {             // one [rows x width] bf16 matrix in both orientations
  bf16* p; long long ld;        // row-major, pitch ld (multiple of 8)
  bf16* t; long long ldt;       // transposed [width x rows], pitch ldt = round8(rows)
  int width;
};

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const __nv_bfloat162* hp = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) { const float2 x = __bfloat1622float2(hp[i]); f[2 * i] = x.x; f[2 * i + 1] = x.y; }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 u;
  __nv_bfloat162* hp = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) hp[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return u;
}
// 16-byte read-only load the compiler cannot split or fence behind a branch (the streaming kernels below issue several per
// thread before the first use)
__device__ __forceinline__ uint4 ldg16(const void* p) {
  uint4 v;
  asm("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ bool aligned16(const void* p) { return (reinterpret_cast<size_t>(p) & 15) == 0; }
__device__ __forceinline__ void load8(const bf16* p, int n_valid, float (&f)[8]) {      // 8 bf16, zero beyond n_valid
  if (n_valid >= 8 && (reinterpret_cast<size_t>(p) & 15) == 0) {
    unpack8(*reinterpret_cast<const uint4*>(p), f);
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = i < n_valid ? __bfloat162float(p[i]) : 0.f;
  }
}
__device__ __forceinline__ void load8(const float* p, int n_valid, float (&f)[8]) {     // 8 fp32, zero beyond n_valid
  if (n_valid >= 8 && (reinterpret_cast<size_t>(p) & 15) == 0) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = i < n_valid ? p[i] : 0.f;
  }
}

// 64 x 64 tile written in both orientations with 16-byte accesses: value8(r, c, n_valid, v) fills the 8 values of row r,
// columns c .. c+7 (n_valid of them exist) exactly once; they are stored row-major and, through shared memory, transposed.
// Launch: 256 threads, grid (ceil(width / 64), ceil(rows / 64)).
template <typename F>
__device__ __forceinline__ void tile64_emit(F value8, long long rows, int width, bf16* dst, long long ld, bf16* dst_t, long long ldt,
                                            int colofs, int col_tile = -1) {
  __shared__ __align__(16) bf16 tile[64][72];
  const int t = threadIdx.x;
  const long long r0 = (long long)blockIdx.y * 64;
  const int c0 = (col_tile < 0 ? (int)blockIdx.x : col_tile) * 64;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int rr = (t >> 3) + 32 * i, cc = (t & 7) * 8;
    const long long r = r0 + rr;
    const int c = c0 + cc;
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = 0.f;
    const int nv = min(8, width - c);
    if (r < rows && nv > 0) value8(r, c, nv, v);
    const uint4 pk = pack8(v);
    if (dst && r < rows && nv > 0) {
      bf16* op = dst + r * ld + colofs + c;
      if (nv == 8 && (reinterpret_cast<size_t>(op) & 15) == 0) {
        *reinterpret_cast<uint4*>(op) = pk;
      } else {
        const bf16* pv = reinterpret_cast<const bf16*>(&pk);
        for (int u = 0; u < nv; ++u) op[u] = pv[u];
      }
    }
    *reinterpret_cast<uint4*>(&tile[rr][cc]) = pk;
  }
  __syncthreads();
  if (dst_t) {
    const int kk = t >> 2, rc = (t & 3) * 16;
    const int c = c0 + kk;
    const long long r = r0 + rc;
    if (c < width && r < rows) {
      __align__(16) bf16 col[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) col[j] = tile[rc + j][kk];
      bf16* op = dst_t + (long long)(colofs + c) * ldt + r;
      if (r + 16 <= rows && (reinterpret_cast<size_t>(op) & 15) == 0) {
        reinterpret_cast<uint4*>(op)[0] = reinterpret_cast<const uint4*>(col)[0];
        reinterpret_cast<uint4*>(op)[1] = reinterpret_cast<const uint4*>(col)[1];
      } else {
        for (int j = 0; j < 16 && r + j < rows; ++j) op[j] = col[j];
      }
    }
  }
}

// fp32 master weights -> bf16 W (pitch ldk) and W^T [ktot x out] (pitch ldo), every layer of the model in ONE launch:
// blockIdx.z selects the layer, blocks outside its extent leave at once
struct CastJob { const float* W; bf16* Wb; bf16* WbT; int out, ktot; long long ldk, ldo; };
constexpr int kMaxCastJobs = (MMN_MAX_ENCODERS + MMN_MAX_DECODERS) * MMN_MAX_LAYERS;
struct CastJobs { CastJob job[kMaxCastJobs]; };
__global__ void __launch_bounds__(256) wide_cast_weights_kernel(const CastJobs* __restrict__ jobs) {
  pdl_entry();
  const CastJob j = jobs->job[blockIdx.z];
  if ((int)blockIdx.x * 64 >= j.ktot || (int)blockIdx.y * 64 >= j.out) return;
  tile64_emit([&](long long r, int c, int nv, float (&v)[8]) { load8(j.W + r * j.ktot + c, nv, v); }, j.out, j.ktot, j.Wb, j.ldk,
              j.WbT, j.ldo, 0);
}
// features of one modality -> columns [0, F) of the first layer's input; NaN -> 0 and the row is marked absent
__global__ void __launch_bounds__(256) wide_input_x_kernel(const float* __restrict__ x, long long x_ld, long long rows, int F, Mat in,
                                                           unsigned char* present, Drop drop) {
  pdl_entry();
  tile64_emit([&](long long r, int c, int nv, float (&v)[8]) {
    load8(x + r * x_ld + c, nv, v);
    bool nan = false;
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (v[i] != v[i]) { nan = true; v[i] = 0.f; }
    if (nan) present[r] = 0;
    if (drop.enabled) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        v[i] = mmn_dropout_keep(drop.seed_mix, drop.row_base + (unsigned)r, (unsigned)(c + i), drop.thr) ? v[i] * drop.scale : 0.f;
    }
  }, rows, F, in.p, in.ld, in.t, in.ldt, 0);
}
// the running state -> columns [colofs, colofs + S) of a layer input (torch.cat([x, state]), mlp_encoder.py:41,78)
__global__ void __launch_bounds__(256) wide_input_state_kernel(Mat s, long long rows, Mat in, int colofs, Drop drop) {
  pdl_entry();
  tile64_emit([&](long long r, int c, int nv, float (&v)[8]) {
    load8(s.p + r * s.ld + c, nv, v);
    if (drop.enabled) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        v[i] = mmn_dropout_keep(drop.seed_mix, drop.row_base + (unsigned)r, (unsigned)(colofs + c + i), drop.thr) ? v[i] * drop.scale : 0.f;
    }
  }, rows, s.width, in.p, in.ld, in.t, in.ldt, colofs);
}
// both parts of an encoder's first-layer input in one launch: column tiles [0, ceil(F / 64)) stage the features, the rest the state
__global__ void __launch_bounds__(256) wide_input_xs_kernel(const float* __restrict__ x, long long x_ld, long long rows, int F, Mat s,
                                                            Mat in, int colofs, unsigned char* present, Drop drop) {
  pdl_entry();
  const int x_tiles = (F + 63) / 64;
  if ((int)blockIdx.x < x_tiles) {
    tile64_emit([&](long long r, int c, int nv, float (&v)[8]) {
      load8(x + r * x_ld + c, nv, v);
      bool nan = false;
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (v[i] != v[i]) { nan = true; v[i] = 0.f; }
      if (nan) present[r] = 0;
      if (drop.enabled) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
          v[i] = mmn_dropout_keep(drop.seed_mix, drop.row_base + (unsigned)r, (unsigned)(c + i), drop.thr) ? v[i] * drop.scale : 0.f;
      }
    }, rows, F, in.p, in.ld, in.t, in.ldt, 0);
  } else {
    tile64_emit([&](long long r, int c, int nv, float (&v)[8]) {
      load8(s.p + r * s.ld + c, nv, v);
      if (drop.enabled) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
          v[i] = mmn_dropout_keep(drop.seed_mix, drop.row_base + (unsigned)r, (unsigned)(colofs + c + i), drop.thr) ? v[i] * drop.scale : 0.f;
      }
    }, rows, s.width, in.p, in.ld, in.t, in.ldt, colofs, (int)blockIdx.x - x_tiles);
  }
}
// s_0 = tile(state_value) (state.py:29-32)
__global__ void __launch_bounds__(256) wide_init_state_kernel(const float* __restrict__ init, long long rows, Mat s) {
  pdl_entry();
  tile64_emit([&](long long, int c, int nv, float (&v)[8]) { load8(init + c, nv, v); }, rows, s.width, s.p, s.ld, s.t, s.ldt, 0);
}
// G += add_k + u_k ; dz = present ? G * act'(s_k) : 0      (u_k = c_sc (s_k - s_{k-1}); add_k = the decoders' gradient
// with respect to s_k when it was computed for all steps at once, else null)
__global__ void __launch_bounds__(256) wide_state_grad_kernel(float* G, const float* __restrict__ add, Mat sk, Mat skm1,
                                                              const unsigned char* present, const int* skip, float c_sc, int act,
                                                              long long rows, Mat dz) {
  pdl_entry();
  const bool skipped = skip && *skip != 0;
  tile64_emit([&](long long r, int c, int nv, float (&v)[8]) {
    float a[8], b[8], g[8];
    load8(sk.p + r * sk.ld + c, nv, a);
    load8(skm1.p + r * skm1.ld + c, nv, b);
    float* gp = G + r * sk.width + c;
    load8(gp, nv, g);
    if (add) {
      float d[8];
      load8(add + r * sk.width + c, nv, d);
#pragma unroll
      for (int i = 0; i < 8; ++i) g[i] += d[i];
    }
    const bool live = present[r] && !skipped;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      g[i] += c_sc * (a[i] - b[i]);
      v[i] = live ? g[i] * wide_dact(act, a[i]) : 0.f;
    }
    if (nv == 8 && (reinterpret_cast<size_t>(gp) & 15) == 0) {
      reinterpret_cast<float4*>(gp)[0] = make_float4(g[0], g[1], g[2], g[3]);
      reinterpret_cast<float4*>(gp)[1] = make_float4(g[4], g[5], g[6], g[7]);
    } else {
      for (int i = 0; i < nv; ++i) gp[i] = g[i];
    }
  }, rows, sk.width, dz.p, dz.ld, dz.t, dz.ldt, 0);
}
// bias gradient: gb[n] += sum_r dz[r][n]          CTA = 64 columns x a slice of the rows; thread = 8 columns, every 32nd row
__global__ void __launch_bounds__(256) wide_bias_grad_kernel(const bf16* __restrict__ dz, long long ld, long long rows, int n_out, float* gb) {
  pdl_entry();
  __shared__ float red[32][65];
  const int chunk = threadIdx.x & 7, rl = threadIdx.x >> 3;
  const int n0 = blockIdx.x * 64 + chunk * 8;
  const long long per = (rows + gridDim.y - 1) / gridDim.y, r0 = blockIdx.y * per, r1 = min(rows, r0 + per);
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  if (n0 < n_out) {
    long long r = r0 + rl;
    if (n0 + 8 <= n_out && aligned16(dz + n0) && (ld & 7) == 0) {     // straight-line: 8 loads in flight per thread
      const bf16* src = dz + n0;
      for (; r + 7 * 32 < r1; r += 8 * 32) {
        uint4 u[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) u[j] = ldg16(src + (r + 32 * j) * ld);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float v[8];
          unpack8(u[j], v);
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[i] += v[i];
        }
      }
    }
    for (; r < r1; r += 32) {
      float v[8];
      load8(dz + r * ld + n0, n_out - n0, v);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] += v[i];
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) red[rl][chunk * 8 + i] = acc[i];
  __syncthreads();
  if (threadIdx.x < 64) {
    float tot = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) tot += red[j][threadIdx.x];
    const int n = blockIdx.x * 64 + threadIdx.x;
    if (n < n_out && tot != 0.f) atomicAdd(gb + n, tot);
  }
}
// column sums of the fp32 state gradient -> gradient of state_value (tile backward, state.py:30)
__global__ void wide_colsum_f32_kernel(const float* __restrict__ G, const float* __restrict__ add, long long rows, int S, float* out) {
  pdl_entry();
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  const int part = threadIdx.x >> 5, parts = blockDim.x >> 5;
  __shared__ float red[8][32];
  float s = 0.f;
  if (c < S)
    for (long long r = part + (long long)blockIdx.y * parts; r < rows; r += (long long)parts * gridDim.y)
      s += G[r * S + c] + (add ? add[r * S + c] : 0.f);
  red[part][threadIdx.x & 31] = s;
  __syncthreads();
  if (part == 0 && c < S) {
    float tot = 0.f;
    for (int w = 0; w < parts; ++w) tot += red[w][threadIdx.x];
    atomicAdd(out + c, tot);
  }
}
// bf16 state -> fp32 final_state
__global__ void wide_state_out_kernel(Mat s, long long rows, float* out) {
  pdl_entry();
  const long long n = rows * s.width;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / s.width;
    out[i] = __bfloat162float(s.p[r * s.ld + (i - r * s.width)]);
  }
}


// ---- decoder head: the last Linear of an MLP decoder (out = n_classes <= 32) is far too skinny for a 128 x 256 tensor-core
// tile; two bandwidth-bound kernels replace its forward and its backward (data + weight gradient) GEMMs ----
// p[r][c] = act(b[c] + sum_k h[r][k] W[c][k])       fp32 accumulation of bf16 products.
// A warp owns kHeadRows consecutive rows at a time and strides over k 256 columns per trip, so each weight chunk is fetched
// once per kHeadRows rows and kHeadRows 16-byte loads of h are in flight per lane and trip (x2 with the unroll).
// fast = every row of h and W starts 16-byte aligned and the width is a multiple of 8 (checked by the launcher).
constexpr int kHeadRows = 4;
template <int C, bool FAST>
__global__ void __launch_bounds__(256, 3) wide_head_fwd_kernel(Mat h, const bf16* __restrict__ Wb, long long ldk, const float* __restrict__ bias,
                                                            int act, long long rows, float* __restrict__ p_out) {
  pdl_entry();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (long long rb = ((long long)blockIdx.x * 8 + warp) * kHeadRows; rb < rows; rb += (long long)gridDim.x * 8 * kHeadRows) {
    float acc[kHeadRows][C];
    const bf16* hr[kHeadRows];
#pragma unroll
    for (int q = 0; q < kHeadRows; ++q) {
      hr[q] = h.p + min(rb + q, rows - 1) * h.ld;      // rows past the end re-read the last one; their result is dropped
#pragma unroll
      for (int c = 0; c < C; ++c) acc[q][c] = 0.f;
    }
    if (FAST) {
#pragma unroll 2
      for (int k0 = lane * 8; k0 < h.width; k0 += 256) {
        uint4 hu[kHeadRows], wu[C];
#pragma unroll
        for (int q = 0; q < kHeadRows; ++q) hu[q] = ldg16(hr[q] + k0);
#pragma unroll
        for (int c = 0; c < C; ++c) wu[c] = ldg16(Wb + (long long)c * ldk + k0);
#pragma unroll
        for (int c = 0; c < C; ++c) {
          float wv[8];
          unpack8(wu[c], wv);
#pragma unroll
          for (int q = 0; q < kHeadRows; ++q) {
            float hv[8];
            unpack8(hu[q], hv);
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[q][c] = fmaf(hv[i], wv[i], acc[q][c]);
          }
        }
      }
    } else {
      for (int k0 = lane * 8; k0 < h.width; k0 += 256) {
#pragma unroll
        for (int c = 0; c < C; ++c) {
          float wv[8];
          load8(Wb + (long long)c * ldk + k0, h.width - k0, wv);
#pragma unroll
          for (int q = 0; q < kHeadRows; ++q) {
            float hv[8];
            load8(hr[q] + k0, h.width - k0, hv);
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[q][c] = fmaf(hv[i], wv[i], acc[q][c]);
          }
        }
      }
    }
#pragma unroll
    for (int q = 0; q < kHeadRows; ++q) {
#pragma unroll
      for (int c = 0; c < C; ++c) {
        float s = acc[q][c];
#pragma unroll
        for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0 && rb + q < rows) p_out[(rb + q) * C + c] = wide_act(act, s + bias[c]);
      }
    }
  }
}
// The head's whole backward in one pass over its input h:  dW[c][k] += sum_r dz[r][c] h[r][k],  db[c] += sum_r dz[r][c]  and
// dh[r][k] = act'(h[r][k]) * sum_c dz[r][c] W[c][k].   thread = 8 consecutive k, blockIdx.y = a slice of the rows.
// fast (see above, plus dz and dh rows 16-byte aligned): four rows of h in flight per thread, one 8-byte load of a dz row.
constexpr int kHeadBwdCtas(int C) { return C <= 2 ? 3 : 2; }      // resident CTAs per SM the register budget is set for
template <int C> struct DzWord { typedef uint2 type; };      // one row of dz (C <= 4 bf16) in one load
template <> struct DzWord<1> { typedef unsigned type; };
template <> struct DzWord<2> { typedef unsigned type; };
template <int C, bool FAST>
__global__ void __launch_bounds__(256, kHeadBwdCtas(C)) wide_head_backward_kernel(Mat dz, Mat h, long long rows, float* gW, long long ldw, float* gb,
                                                                    const bf16* __restrict__ Wb, long long ldk, int act_prev, Mat out) {
  pdl_entry();
  const int k0 = (blockIdx.x * 256 + threadIdx.x) * 8;
  const long long per = (rows + gridDim.y - 1) / gridDim.y, r0 = blockIdx.y * per, r1 = min(rows, r0 + per);
  float acc[C][8], wv[C][8];
#pragma unroll
  for (int c = 0; c < C; ++c) {
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[c][i] = wv[c][i] = 0.f;
    if (k0 < h.width) load8(Wb + (long long)c * ldk + k0, h.width - k0, wv[c]);
  }
  if (k0 < h.width) {
    const int nv = min(8, h.width - k0);
    // one row: dh and the weight-gradient partial sums
    auto row = [&](long long r, const float (&hv)[8], const float (&d)[C]) {
      float dh[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) dh[i] = 0.f;
#pragma unroll
      for (int c = 0; c < C; ++c) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          acc[c][i] = fmaf(d[c], hv[i], acc[c][i]);
          dh[i] = fmaf(d[c], wv[c][i], dh[i]);
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) dh[i] *= wide_dact(act_prev, hv[i]);
      bf16* op = out.p + r * out.ld + k0;
      if (FAST) {
        *reinterpret_cast<uint4*>(op) = pack8(dh);
      } else if (nv == 8 && aligned16(op)) {
        *reinterpret_cast<uint4*>(op) = pack8(dh);
      } else {
        for (int i = 0; i < nv; ++i) op[i] = __float2bfloat16(dh[i]);
      }
    };
    long long r = r0;
    if (FAST) {
      for (; r + 4 <= r1; r += 4) {
        uint4 hu[4];
        typename DzWord<C>::type du[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          hu[j] = ldg16(h.p + (r + j) * h.ld + k0);
          du[j] = *reinterpret_cast<const typename DzWord<C>::type*>(dz.p + (r + j) * dz.ld);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float hv[8], d[C];
          unpack8(hu[j], hv);
          const bf16* dp = reinterpret_cast<const bf16*>(&du[j]);
#pragma unroll
          for (int c = 0; c < C; ++c) d[c] = __bfloat162float(dp[c]);
          row(r + j, hv, d);
        }
      }
    }
    for (; r < r1; ++r) {
      float hv[8], d[C];
      load8(h.p + r * h.ld + k0, nv, hv);
#pragma unroll
      for (int c = 0; c < C; ++c) d[c] = __bfloat162float(dz.p[r * dz.ld + c]);
      row(r, hv, d);
    }
#pragma unroll
    for (int c = 0; c < C; ++c) {
      float* dst = gW + (long long)c * ldw + k0;
      if (k0 + 8 <= h.width && aligned16(dst)) {      // two red.global.add.v4.f32
        atomicAdd(reinterpret_cast<float4*>(dst), make_float4(acc[c][0], acc[c][1], acc[c][2], acc[c][3]));
        atomicAdd(reinterpret_cast<float4*>(dst) + 1, make_float4(acc[c][4], acc[c][5], acc[c][6], acc[c][7]));
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (k0 + i < h.width && acc[c][i] != 0.f) atomicAdd(dst + i, acc[c][i]);
      }
    }
  }
  if (blockIdx.x == 0 && (int)threadIdx.x < C) {
    float s = 0.f;
    for (long long r = r0; r < r1; ++r) s += __bfloat162float(dz.p[r * dz.ld + threadIdx.x]);
    if (s != 0.f) atomicAdd(gb + threadIdx.x, s);
  }
}

struct LossArgs {
  const float* p; long long ldp;          // decoder outputs [rows x C] fp32
  int C, act, D, d, hist_row, n_mat_rows; // n_mat_rows = E + 1
  long long rows;
  const long long* targets;               // [rows x D] or null
  int* target_error;                      // device flag: a target outside [0, C) was seen
  const unsigned char* present;           // [rows] or null (initial state: every row counts)
  const int* skip;
  double* metrics; double inv_rows_global;
  unsigned char* predictions;             // already offset to (hist_row, d, 0), or null
  float* last_outputs; long long ld_last; int out_off;   // or null
  float coef;                             // err_penalty / (D (E+1) B_global); 0 = forward only
  Mat dz;                                 // TRAIN: gradient w.r.t. the last layer's pre-activation
};
// per-row decoder epilogue (multimodn.py:144-157,179-191): first-max arg-max, CE on the outputs, counters, and the
// gradient of the loss through the output activation
__device__ __forceinline__ void decoder_loss_rows(const LossArgs& a) {
  const long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const bool rv = r < a.rows;
  const bool skipped = a.skip && *a.skip != 0;
  const bool m = rv && (a.present ? a.present[r] != 0 : true) && !skipped;
  float p[MMN_MAX_CLASSES];
  float mx = -3.4e38f, best = 0.f;
  int pred = 0, y = 0;
  float ce = 0.f;
  unsigned ok = 0, tp = 0, tn = 0, fp = 0, fn = 0;
  if (rv) {
    for (int c = 0; c < a.C; ++c) p[c] = a.p[r * a.ldp + c];
    best = p[0];
    for (int c = 0; c < a.C; ++c) {
      if (c && (p[c] > best || (p[c] != p[c] && best == best))) { best = p[c]; pred = c; }
      mx = fmaxf(mx, p[c]);
    }
    if (a.predictions) a.predictions[r] = (unsigned char)pred;
    if (a.last_outputs)
      for (int c = 0; c < a.C; ++c) a.last_outputs[r * a.ld_last + a.out_off + c] = p[c];
    if (a.targets) {
      y = (int)a.targets[r * a.D + a.d];
      if ((y < 0 || y >= a.C) && a.target_error) *a.target_error = 1;
      y = y < 0 ? 0 : (y >= a.C ? a.C - 1 : y);
      float se = 0.f;
      for (int c = 0; c < a.C; ++c) se += expf(p[c] - mx);
      if (m) {
        ce = mx + logf(se) - p[y];
        ok = pred == y;
        if (a.C == 2) { tp = pred == 1 && y == 1; tn = pred == 0 && y == 0; fp = pred == 1 && y == 0; fn = pred == 0 && y == 1; }
      }
      if (a.coef != 0.f) {
        const float inv = 1.f / se, coef = m ? a.coef : 0.f;
        for (int c = 0; c < a.C; ++c) {
          const float g = coef * (expf(p[c] - mx) * inv - (c == y ? 1.f : 0.f)) * wide_dact(a.act, p[c]);
          a.dz.p[r * a.dz.ld + c] = __float2bfloat16(g);
          if (a.dz.t) a.dz.t[(long long)c * a.dz.ldt + r] = __float2bfloat16(g);
        }
      }
    }
  }
  if (!a.targets || !a.metrics) return;
  __shared__ float s_ce[8];
  __shared__ unsigned s_cnt[8][5];
  unsigned cnt[5] = {ok, tp, tn, fp, fn};
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    ce += __shfl_xor_sync(0xffffffffu, ce, o);
#pragma unroll
    for (int i = 0; i < 5; ++i) cnt[i] += __shfl_xor_sync(0xffffffffu, cnt[i], o);
  }
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) {
    s_ce[w] = ce;
    for (int i = 0; i < 5; ++i) s_cnt[w][i] = cnt[i];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double tce = 0.0;
    unsigned tc[5] = {0, 0, 0, 0, 0};
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) {
      tce += s_ce[i];
      for (int j = 0; j < 5; ++j) tc[j] += s_cnt[i][j];
    }
    const int stride = a.n_mat_rows * a.D, at = a.hist_row * a.D + a.d;
    if (tce != 0.0) atomicAdd(a.metrics + at, tce * a.inv_rows_global);
    for (int j = 0; j < 5; ++j)
      if (tc[j]) atomicAdd(a.metrics + (j + 1) * stride + at, (double)tc[j]);
  }
}
__global__ void wide_decoder_loss_kernel(const LossArgs a) {
  pdl_entry(); decoder_loss_rows(a); }
// The same epilogue for one decoder after EVERY step in one launch (training: the decoder ran once over all steps' states):
// blockIdx.y = step k, whose rows are the k-th block of `rows` rows of p / dz.
struct LossSteps {
  LossArgs base;                           // p, dz: step 0; hist_row / present / skip / predictions / last_outputs: per step below
  int n_steps;
  long long pred_ld;                       // predictions: [(E + 1) x D x pred_ld], base.predictions = the array's start (or null)
  struct Step { int hist_row, is_last; const unsigned char* present; const int* skip; } step[MMN_MAX_ENCODERS + 1];
};
__global__ void wide_decoder_loss_steps_kernel(const __grid_constant__ LossSteps s) {
  pdl_entry();
  const int k = blockIdx.y;
  LossArgs a = s.base;
  a.p += (long long)k * a.rows * a.ldp;
  a.dz.p += (long long)k * a.rows * a.dz.ld;
  a.dz.t = nullptr;
  a.hist_row = s.step[k].hist_row;
  a.present = s.step[k].present;
  a.skip = s.step[k].skip;
  a.predictions = s.base.predictions ? s.base.predictions + ((long long)a.hist_row * a.D + a.d) * s.pred_ld : nullptr;
  if (!s.step[k].is_last) a.last_outputs = nullptr;
  decoder_loss_rows(a);
}
// per-step bookkeeping: present-row counts, state-change means, the gradient buffer's "encoder took a row" tail
__device__ __forceinline__ void finalize_step(const unsigned char* present, long long rows, int k, int hist_row, int e, const int* skip,
                                              const float* sc_sum, int S, double inv_rows_global, double* met_present, double* met_sc,
                                              float* grad_tail) {
  __shared__ unsigned red[8];
  unsigned n = 0;
  const bool skipped = skip && *skip != 0;
  for (long long r = threadIdx.x; r < rows; r += blockDim.x) n += (k == 0 || (present[r] && !skipped)) ? 1u : 0u;
#pragma unroll
  for (int o = 16; o; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = n;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned tot = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += red[w];
    if (met_present && tot) atomicAdd(met_present + hist_row, (double)tot);
    if (k > 0) {
      if (met_sc && sc_sum && *sc_sum != 0.f) atomicAdd(met_sc + e, (double)*sc_sum * inv_rows_global / (double)S);
      if (grad_tail && tot) atomicAdd(grad_tail + e, (float)tot);
    }
  }
}
__global__ void wide_finalize_kernel(const unsigned char* present, long long rows, int k, int hist_row, int e, const int* skip,
                                     const float* sc_sum, int S, double inv_rows_global, double* met_present, double* met_sc,
                                     float* grad_tail) {
  pdl_entry();
  finalize_step(present, rows, k, hist_row, e, skip, sc_sum, S, inv_rows_global, met_present, met_sc, grad_tail);
}
// every step of a training call in one launch: blockIdx.x = step k (present: [(L + 1) x rows], sc_sum: one float per encoder)
struct FinalizeSteps {
  const unsigned char* present; long long rows;
  const float* sc_sum; int S; double inv_rows_global;
  double* met_present; double* met_sc; float* grad_tail;
  struct Step { int hist_row, e; const int* skip; } step[MMN_MAX_ENCODERS + 1];
};
__global__ void wide_finalize_steps_kernel(const __grid_constant__ FinalizeSteps s) {
  pdl_entry();
  const int k = blockIdx.x;
  finalize_step(s.present + (long long)k * s.rows, s.rows, k, s.step[k].hist_row, s.step[k].e, s.step[k].skip,
                k > 0 && s.sc_sum ? s.sc_sum + s.step[k].e : nullptr, s.S, s.inv_rows_global, s.met_present, s.met_sc, s.grad_tail);
}

}  // namespace wide
}  // namespace mmn
