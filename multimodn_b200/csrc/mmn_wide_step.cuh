// mmn_wide_step.cuh — the wide regime's step: the sequential-fusion chain run layer by layer on the tcgen05 bf16 GEMM
// of mmn_wide.cuh (precision = bf16: bf16 weights / activations / layer gradients, fp32 accumulation, fp32 master
// weights, fp32 state gradient and parameter gradients).  Included by mmn_api.cu (needs fail(), MMN_CUDA, mmn_plan).
//
// Every activation and every layer gradient is kept in BOTH orientations ([rows x width] and [width x rows]) so that
// forward, data-gradient and weight-gradient GEMMs are all "A[M x K] . B[N x K]^T with K contiguous":
//   forward   Y[B x out]   = In[B x k]      . W[out x k]^T
//   dgrad     dIn[B x in]  = dZ[B x out]    . W^T[in x out]^T
//   wgrad     dW[out x k]  = dZ^T[out x B]  . In^T[k x B]^T
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

// 32 x 32 tile: value(r, c) computed once per element, written row-major and (through shared memory) transposed
template <typename F>
__device__ __forceinline__ void tile32_emit(F value, long long rows, int width, bf16* dst, long long ld, bf16* dst_t, long long ldt,
                                            int colofs) {
  __shared__ float tile[32][33];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const long long r0 = (long long)blockIdx.y * 32;
  const int c0 = blockIdx.x * 32;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long r = r0 + ty + 8 * i;
    const int c = c0 + tx;
    float v = 0.f;
    if (r < rows && c < width) {
      v = value(r, c);
      if (dst) dst[r * ld + colofs + c] = __float2bfloat16(v);
    }
    tile[ty + 8 * i][tx] = v;
  }
  __syncthreads();
  if (dst_t) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = c0 + ty + 8 * i;
      const long long r = r0 + tx;
      if (r < rows && c < width) dst_t[(long long)(colofs + c) * ldt + r] = __float2bfloat16(tile[tx][ty + 8 * i]);
    }
  }
}

// fp32 master weight [out x ktot] -> bf16 W (pitch ldk) and W^T [ktot x out] (pitch ldo)
__global__ void wide_cast_weight_kernel(const float* __restrict__ W, int out, int ktot, bf16* Wb, long long ldk, bf16* WbT, long long ldo) {
  tile32_emit([&](long long r, int c) { return W[r * ktot + c]; }, out, ktot, Wb, ldk, WbT, ldo, 0);
}
// features of one modality -> columns [0, F) of the first layer's input; NaN -> 0 and the row is marked absent
__global__ void wide_input_x_kernel(const float* __restrict__ x, long long x_ld, long long rows, int F, Mat in, unsigned char* present,
                                    Drop drop) {
  tile32_emit([&](long long r, int c) {
    float v = x[r * x_ld + c];
    if (v != v) { present[r] = 0; v = 0.f; }
    if (drop.enabled) v = mmn_dropout_keep(drop.seed_mix, drop.row_base + (unsigned)r, (unsigned)c, drop.thr) ? v * drop.scale : 0.f;
    return v;
  }, rows, F, in.p, in.ld, in.t, in.ldt, 0);
}
// the running state -> columns [colofs, colofs + S) of a layer input (torch.cat([x, state]), mlp_encoder.py:41,78)
__global__ void wide_input_state_kernel(Mat s, long long rows, Mat in, int colofs, Drop drop) {
  tile32_emit([&](long long r, int c) {
    float v = __bfloat162float(s.p[r * s.ld + c]);
    if (drop.enabled) v = mmn_dropout_keep(drop.seed_mix, drop.row_base + (unsigned)r, (unsigned)(colofs + c), drop.thr) ? v * drop.scale : 0.f;
    return v;
  }, rows, s.width, in.p, in.ld, in.t, in.ldt, colofs);
}
// s_0 = tile(state_value) (state.py:29-32)
__global__ void wide_init_state_kernel(const float* __restrict__ init, long long rows, Mat s) {
  tile32_emit([&](long long, int c) { return init[c]; }, rows, s.width, s.p, s.ld, s.t, s.ldt, 0);
}
// G += u_k ; dz = present ? G * act'(s_k) : 0      (u_k = c_sc (s_k - s_{k-1}))
__global__ void wide_state_grad_kernel(float* G, Mat sk, Mat skm1, const unsigned char* present, const int* skip, float c_sc, int act,
                                       long long rows, Mat dz) {
  const bool skipped = skip && *skip != 0;
  tile32_emit([&](long long r, int c) {
    const float a = __bfloat162float(sk.p[r * sk.ld + c]), b = __bfloat162float(skm1.p[r * skm1.ld + c]);
    const float g = G[r * sk.width + c] + c_sc * (a - b);
    G[r * sk.width + c] = g;
    return (present[r] && !skipped) ? g * wide_dact(act, a) : 0.f;
  }, rows, sk.width, dz.p, dz.ld, dz.t, dz.ldt, 0);
}
// G -= u_k (the state-change term reaches s_{k-1} with the opposite sign)
__global__ void wide_state_grad_post_kernel(float* G, Mat sk, Mat skm1, float c_sc, long long rows) {
  const long long n = rows * sk.width;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / sk.width;
    const int c = (int)(i - r * sk.width);
    G[i] -= c_sc * (__bfloat162float(sk.p[r * sk.ld + c]) - __bfloat162float(skm1.p[r * skm1.ld + c]));
  }
}
// bias gradient: gb[n] += sum_r dz^T[n][r]          (one CTA per output)
__global__ void wide_bias_grad_kernel(const bf16* __restrict__ dzt, long long ldt, long long rows, float* gb) {
  __shared__ float red[8];
  const bf16* row = dzt + (long long)blockIdx.x * ldt;
  float s = 0.f;
  for (long long r = threadIdx.x; r < rows; r += blockDim.x) s += __bfloat162float(row[r]);
#pragma unroll
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += red[w];
    atomicAdd(gb + blockIdx.x, tot);
  }
}
// column sums of the fp32 state gradient -> gradient of state_value (tile backward, state.py:30)
__global__ void wide_colsum_f32_kernel(const float* __restrict__ G, long long rows, int S, float* out) {
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  const int part = threadIdx.x >> 5, parts = blockDim.x >> 5;
  __shared__ float red[8][32];
  float s = 0.f;
  if (c < S)
    for (long long r = part + (long long)blockIdx.y * parts; r < rows; r += (long long)parts * gridDim.y) s += G[r * S + c];
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
  const long long n = rows * s.width;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / s.width;
    out[i] = __bfloat162float(s.p[r * s.ld + (i - r * s.width)]);
  }
}

struct LossArgs {
  const float* p; long long ldp;          // decoder outputs [rows x C] fp32
  int C, act, D, d, hist_row, n_mat_rows; // n_mat_rows = E + 1
  long long rows;
  const long long* targets;               // [rows x D] or null
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
__global__ void wide_decoder_loss_kernel(const LossArgs a) {
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
          a.dz.t[(long long)c * a.dz.ldt + r] = __float2bfloat16(g);
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
// per-step bookkeeping: present-row counts, state-change means, the gradient buffer's "encoder took a row" tail
__global__ void wide_finalize_kernel(const unsigned char* present, long long rows, int k, int hist_row, int e, const int* skip,
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

}  // namespace wide
}  // namespace mmn
