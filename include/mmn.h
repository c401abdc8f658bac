/*
 * mmn.h — C ABI of libmmn.so: the B200 (sm_100a) implementation of MultiModN's
 * sequential-fusion step (train / test / predict / get_states).
 *
 * The reference (EPFLiGHT/MultiModN) is pure Python/PyTorch and has no FFI of its own; the
 * "plugin API" of this path is the Python object protocol
 *     MultiModN.train_epoch / test / predict / get_states   (multimodn/multimodn.py:89-492)
 *     MultiModEncoder.forward(state, x)                     (multimodn/encoders/multimod_encoder.py:15-17)
 *     MultiModDecoder.forward(state)                        (multimodn/decoders/multimod_decoder.py:14-16)
 * Each entry point below names the reference lines it replaces.  A maintainer of the reference
 * binds them with ctypes (INTEGRATION.md shows the stub); multimodn_b200/_lib.py is that
 * binding.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; mmn_last_error() returns the
 *     message of the calling thread's last failure.  No C++ exception crosses the boundary.
 *   - the caller (PyTorch) owns every buffer; the library owns only the opaque plan.
 *   - all work is ordered on the caller's stream (a cudaStream_t passed as void*); no call
 *     synchronises the device.  bf16 plans fork two plan-owned streams from the caller's stream
 *     inside mmn_train_step (cudaEventRecord / cudaStreamWaitEvent) and join them before the call
 *     returns: whatever the caller enqueues next on its stream sees every result.
 *   - "device" pointers are CUDA device memory; "host" pointers are ordinary host memory read
 *     before the call returns.
 *   - one host thread per plan at a time.
 */
#ifndef MMN_H_
#define MMN_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MMN_ABI_VERSION 3   /* 2: mmn_model_desc.precision, bf16 plans, gradient-ready events; 3: mmn_outputs.target_error */
#define MMN_MAX_LAYERS 6      /* Linear layers per encoder / decoder */
#define MMN_MAX_ENCODERS 16
#define MMN_MAX_DECODERS 16
#define MMN_MAX_CLASSES 32

/* Arithmetic of a plan.  FP32: the fused per-tile step kernels, 1e-5 relative to the reference.  BF16: bf16 weights /
 * activations / layer gradients, fp32 accumulation, fp32 master weights and gradients; 1e-2 relative to the reference.
 * Narrow models (state <= 64, layers <= 64 wide, <= 3 Linear layers per module, <= 8 classes) run the fused per-tile
 * mma kernel (MMN_ENGINE_NB, one launch per step); everything else the wide regime (BASELINE config 4): every layer a
 * tcgen05 GEMM over the whole batch (MMN_ENGINE_WIDE). */
enum { MMN_PRECISION_FP32 = 0, MMN_PRECISION_BF16 = 1 };

/* activation applied after a Linear layer (mlp_encoder.py:46,76; decoders.py:20,44-45) */
enum { MMN_ACT_IDENTITY = 0, MMN_ACT_RELU = 1, MMN_ACT_SIGMOID = 2, MMN_ACT_TANH = 3 };

/* One nn.Linear: y = act(W [a || state?] + b), W row-major [out_dim x (in_dim + has_state*S)],
 * state columns LAST (torch.cat([x, state]) at mlp_encoder.py:41,78). Offsets are in floats
 * into the packed parameter / gradient buffers and must be multiples of 4 (16-byte aligned). */
typedef struct mmn_layer_desc {
  int32_t in_dim;     /* width of the non-state input */
  int32_t out_dim;
  int32_t act;        /* MMN_ACT_* */
  int32_t has_state;  /* 1: the running state is concatenated to this layer's input */
  int64_t w_off;
  int64_t b_off;
} mmn_layer_desc;

/* MLPEncoder family (mlp_encoder.py:49-94, slp_encoders.py:5-34): has_state on the last layer,
 * act = IDENTITY on it.  MIMIC_MLPEncoder (mlp_encoder.py:9-47): has_state on layer 0, act after
 * every layer, dropout_p applied to [x || state] in training. */
typedef struct mmn_encoder_desc {
  int32_t n_features;
  int32_t n_layers;
  float dropout_p;
  int32_t reserved;
  mmn_layer_desc layers[MMN_MAX_LAYERS];
} mmn_encoder_desc;

/* ClassDecoder / LogisticDecoder / MLPDecoder (decoders.py:9-53). */
typedef struct mmn_decoder_desc {
  int32_t n_classes;
  int32_t n_layers;
  mmn_layer_desc layers[MMN_MAX_LAYERS];
} mmn_decoder_desc;

/* MultiModN(state_size, encoders, decoders, ...) (multimodn.py:66-87). */
typedef struct mmn_model_desc {
  int32_t state_size;
  int32_t n_encoders;
  int32_t n_decoders;
  int32_t precision;  /* MMN_PRECISION_* */
  int64_t init_off;   /* TrainableInitState.state_value (1,S) (state.py:25-27) */
  int64_t n_params;   /* length of the packed parameter buffer in floats */
  const mmn_encoder_desc* encoders; /* host */
  const mmn_decoder_desc* decoders; /* host */
} mmn_model_desc;

/* One batch = the (data, target, encoder_sequence) tuple of multimodn.py:119, already on the
 * device (multimodn.py:132-135).  Rows may be a data-parallel shard of a global batch. */
typedef struct mmn_batch {
  int64_t n_rows;         /* rows in this call */
  int64_t n_rows_global;  /* divisor of the batch means (== n_rows on one GPU) */
  int64_t row_offset;     /* global index of row 0 (keys the dropout stream) */
  int32_t seq_len;        /* steps in the encoding sequence (multimodn.py:509-531) */
  int32_t reserved;
  const int32_t* seq_pos; /* host [seq_len]: index into x[] (position in the data list) */
  const int32_t* seq_enc; /* host [seq_len]: encoder id; ids must be distinct */
  const float* const* x;  /* host array of device pointers, x[pos] = (n_rows, F) fp32 row-major */
  const int64_t* x_ld;    /* host [len(x)]: row stride of x[pos] in floats */
  const int64_t* targets; /* device (n_rows, D) int64 row-major, or NULL (predict/get_states) */
  const int32_t* skip_flags; /* device [seq_len] or NULL.  Non-zero: the step is skipped for every
                                row — the reference's batch-level rule (multimodn.py:167-169),
                                filled by mmn_scan_missing.  NULL: per-row select only. */
} mmn_batch;

/* What a step emits; every pointer is device memory and may be NULL. */
typedef struct mmn_outputs {
  /* accumulated (+=) so one buffer can collect an epoch (multimodn.py:206-212,360-365).
   * mmn_metrics_count() doubles laid out as 6 matrices (E+1)xD row-major, then 2 vectors:
   *   [0] ce          sum_b m_b * CE_b / n_rows_global     (multimodn.py:146,181)
   *   [1] n_correct   (multimodn.py:147,183)
   *   [2] tp [3] tn [4] fp [5] fn   cm[target][pred] cells (multimodn.py:51-58)
   *   n_present (E+1)  rows that took the step (multimodn.py:121,171)
   *   state_change (E) sum (s_k - s_{k-1})^2 / (n_rows_global*S), by ENCODER ID (multimodn.py:174)
   * matrix row 0 = initial state, row e+1 = after encoder id e (multimodn.py:181). */
  double* metrics;
  uint8_t* predictions;  /* [(E+1) x D x pred_ld] first-max class ids (multimodn.py:144,455) */
  int64_t pred_ld;
  float* last_outputs;   /* (n_rows, sum_d C_d): decoder outputs at the step of encoder id E-1
                            (multimodn.py:354-357) */
  float* final_state;    /* (n_rows, S) (multimodn.py:488-492) */
  int32_t* target_error; /* device int32[1] or NULL: set to 1 when a target lies outside [0, n_classes) of its decoder —
                            nn.CrossEntropyLoss raises there (multimodn.py:146); the kernels clamp the index for memory
                            safety and report, the caller raises */
} mmn_outputs;

typedef struct mmn_train_args {
  float err_penalty;                 /* multimodn.py:85 */
  float state_change_penalty_scaled; /* already multiplied by 0.01 (multimodn.py:86) */
  uint32_t dropout_seed;             /* stream key of this step's dropout masks */
  int32_t training;                  /* 1: apply dropout (self.train(), multimodn.py:102) */
} mmn_train_args;

typedef struct mmn_plan mmn_plan;

const char* mmn_last_error(void);
int mmn_abi_version(void);

/* Builds the device-side plan for a model (replaces nn.ModuleList traversal, multimodn.py:83-84,
 * 141,159-163,176). */
int mmn_plan_create(const mmn_model_desc* desc, mmn_plan** out);
void mmn_plan_destroy(mmn_plan* plan);

/* Which kernel family a plan uses:
 *   MMN_ENGINE_FMA   fp32 plans, mmn_train_step (and mmn_forward when the model does not qualify for TC2): FP32-FMA
 *                    register-tile GEMMs, one launch per step
 *   MMN_ENGINE_TC2   fp32 plans, mmn_forward: tcgen05 3xTF32 with the activations resident in tensor memory; needs
 *                    state <= 64, layers <= 64 wide, <= 16 classes (MMN_ENGINE=fma opts out)
 *   MMN_ENGINE_NB    bf16 plans, narrow models (state <= 64, layers <= 64 wide, <= 3 Linear layers per module, <= 8 classes):
 *                    fused per-tile mma.sync kernel, one launch per step, forward and train
 *   MMN_ENGINE_WIDE  bf16 plans, everything else (and MMN_ENGINE=wide): every layer a tcgen05 GEMM over the whole batch
 * (MMN_ENGINE_TC, round 1's shared-memory-staged tcgen05 step engine, no longer exists; the value is kept reserved.) */
enum { MMN_ENGINE_FMA = 0, MMN_ENGINE_TC = 1 /* reserved */, MMN_ENGINE_TC2 = 2, MMN_ENGINE_WIDE = 3, MMN_ENGINE_NB = 4 };
int32_t mmn_plan_engine(const mmn_plan* plan);           /* engine of mmn_train_step */

/* Data-parallel overlap (the gradient all-reduce "issued per encoder block in reverse order to overlap with
 * the remaining backward").  events: cudaEvent_t handles owned by the caller (n = 0 clears).  n = E + 1: every
 * mmn_train_step records events[e] as soon as encoder e's parameter gradients are final and events[E] when the whole
 * gradient buffer is (decoders, initial state, present counts; encoders outside the sequence).  n = E + 2 adds
 * events[E + 1] = every decoder's parameter gradients are final (layer-wise plans finish them before the encoders').
 * n = E + 2 + (number of encoder Linear layers) adds one event per encoder layer, in (encoder, layer) order: that
 * layer's weight and bias gradients are final — the last block to finish is then one layer, not one encoder.
 * Events are recorded on streams ordered with the caller's stream; wait on them from any stream.  Layer-wise (bf16)
 * plans record them between their launches; the single-launch fp32 kernels record all of them after the launch. */
int mmn_plan_set_grad_events(mmn_plan* plan, void* const* events, int32_t n);
/* Layer-wise (bf16) plans: SMs the backward GEMMs may use while gradient collectives are in flight (between the first
 * gradient-ready event of a step and its end).  The GEMMs are persistent and statically scheduled: a collective's CTAs that
 * take SMs from a running GEMM make it wait a whole collective for its last tiles, so a data-parallel caller leaves the
 * collective its SMs up front (NCCL on B200: ~20 CTAs on 2 GPUs, 24 NVLS channels on 8).  0 = all SMs (default without
 * grad events); with grad events the default is 128.  Ignored by the single-launch engines. */
int mmn_plan_set_comm_sms(mmn_plan* plan, int32_t n_sms);
int32_t mmn_plan_forward_engine(const mmn_plan* plan);   /* engine of mmn_forward */

int64_t mmn_metrics_count(const mmn_plan* plan);            /* doubles in mmn_outputs.metrics */
int64_t mmn_grad_count(const mmn_plan* plan);               /* floats in the gradient buffer:
                                                               n_params + E (tail: per-encoder
                                                               present-row counts of the step) */
/* Bytes of scratch for a call on n_rows rows: the activation stash of the resident batch tiles (fp32 plans, backward
 * only) or every layer's activations in both orientations (bf16 plans; mmn_forward needs a workspace too). */
int64_t mmn_workspace_bytes(const mmn_plan* plan, int64_t n_rows, int32_t with_backward);

/* Reference batch-level missingness test, `any(data_encoder.isnan().flatten())`
 * (multimodn.py:168,331,485): flags[k] = 1 if x[seq_pos[k]] holds any NaN.  flags: device
 * int32[seq_len], overwritten. */
int mmn_scan_missing(const mmn_plan* plan, const mmn_batch* batch, int32_t* flags, void* stream);

/* Forward chain: test() / predict() / get_states() inner loop (multimodn.py:301-357,434-455,
 * 476-490): init state, per step encoder + per-row missingness select, every decoder, masked
 * CE / counters / predictions. */
int mmn_forward(const mmn_plan* plan, const mmn_batch* batch, const float* params,
                const mmn_outputs* out, void* workspace, size_t workspace_bytes, void* stream);

/* One train step without the optimizer: forward as above (plus the state-change term) and
 * loss.backward() (multimodn.py:139-203) in one launch.  grads: device float[mmn_grad_count],
 * overwritten (zero_grad + backward, multimodn.py:137,203). */
int mmn_train_step(const mmn_plan* plan, const mmn_batch* batch, const float* params,
                   const mmn_train_args* args, const mmn_outputs* out, float* grads,
                   void* workspace, size_t workspace_bytes, void* stream);

/* torch.optim.Adam.step() on the packed buffers (optimizer.step(), multimodn.py:204; defaults of
 * pipelines/titanic/titanic_mlp_pipeline.py:74).  A parameter block whose encoder took no row
 * this step (grads tail == 0) is left untouched, moments and step count included — what torch
 * does for `.grad is None`.  step_count: device int32[1 + E] (decoders/init, then per encoder). */
int mmn_adam_step(const mmn_plan* plan, float* params, const float* grads, float* exp_avg,
                  float* exp_avg_sq, int32_t* step_count, float lr, float beta1, float beta2,
                  float eps, void* stream);

/* Diagnostic (tests only): one 128-row tcgen05 3xTF32 GEMM in each operand configuration of the
 * tensor-core engine.  mode 0: out[r][j] = sum_k a[r][k] b[j][k]; mode 1: out[r][j] = sum_k a[r][k] b[k][j];
 * mode 2: out[i][j] = sum_r a[r][i] b[r][j] (a: 128 x 64, rows i < 64 meaningful); modes 3 / 4 = modes 0 / 1 with
 * the A operand in tensor memory (tcgen05.st + TS-form MMA).  k = 32, n in {32, 64}. */
int mmn_selftest_umma(int mode, int n, const float* a, const float* b, float* out, void* stream);

/* Diagnostic: cycles per round of the worker <-> MMA-issuer mbarrier handshake (out: device int64[2]). */
int mmn_selftest_protocol(int iters, int n_mma, int flags, long long* out, void* stream);

/* Diagnostic (tests, profiles): the wide regime's GEMM on its own — out[M x N] = a[M x K] . b[N x K]^T, a and b bf16
 * with K contiguous (row pitches lda / ldb in elements, multiples of 8), fp32 accumulation in tensor memory.  Any of
 * out_f32 (M x N fp32), out_bf16 (M x N) and out_bf16_t (N x M, the transposed copy every wide-regime epilogue also
 * writes) may be NULL. */
int mmn_selftest_gemm_bf16(int M, int N, int K, const void* a, long long lda, const void* b, long long ldb,
                           float* out_f32, void* out_bf16, void* out_bf16_t, void* stream);

/* The same GEMM with both operands used "transposed" in place, as the weight-gradient GEMMs do: out[M x N] =
 * sum_k a[k][m] b[k][n], a: [K x M] and b: [K x N] bf16 row-major (pitches multiples of 8), out_f32: M x N. */
int mmn_selftest_gemm_bf16_mn(int M, int N, int K, const void* a, long long lda, const void* b, long long ldb,
                              float* out_f32, void* stream);

/* Diagnostic (bench.py): FP32-FMA micro-benchmark — every SM runs 2048 threads of 8 independent FFMA chains, `iters` rounds
 * each.  out: device float[4] (keeps the chains alive); *flops (host) receives the floating-point operations of the launch.
 * Timed by the caller with CUDA events on `stream`: the measured FP32-FMA peak the fp32 step kernels are held against
 * (SURVEY.md section 8d). */
int mmn_selftest_fma_peak(int iters, float* out, double* flops, void* stream);

/* Kernels launched so far by bf16 (wide-regime) plans in this process: launch accounting for benchmarks. */
int64_t mmn_wide_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* MMN_H_ */
