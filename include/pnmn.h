/* pnmn.h — C ABI of the B200-native probnmn-clevr hot path.
 *
 * The reference (kdexd/probnmn-clevr) has no FFI: its "plugin API" for this path is the python
 * nn.Module surface of NeuralModuleNetwork / ProgramGenerator.  The python classes in
 * probnmn_clevr_b200/ keep that surface and call the entry points below through ctypes with raw
 * device pointers, sizes and a cudaStream_t — no torch types cross this boundary.  Every entry
 * point cites the reference code it replaces.  Conventions:
 *   - all functions returning int: 0 = ok, non-zero = error; pnmn_last_error() has the text
 *   - the product entry points allocate NO device memory: the caller owns every device buffer (sizes are reported by
 *     pnmn_plan_sizes / pnmn_model_packed_floats / pnmn_pg_workspace_bytes).  What the library does own: a small pool of
 *     page-locked HOST staging buffers for task tables (recycled, bounded) and, per device, a handful of non-blocking
 *     streams and events: one side stream with two events for the module network (bias gradients run underneath the
 *     weight-gradient kernel); for the seq2seq passes a capture stream, four branch streams with twelve events (independent
 *     launches of a pass run as parallel branches) and the instantiated CUDA graphs of the passes seen so far (at most 256
 *     keys; a graph's device-side footprint is the driver's).  Only the pnmn_debug_* bring-up entry points allocate (and
 *     free) device scratch of their own.
 *   - every launch goes to the device that is current on the calling thread; the python layer makes the tensors'
 *     device current around each call
 *   - no CPU fallback exists: without a CUDA device every compute entry point fails
 *   - thread safety: a plan/model is used by one thread at a time; different plans are independent
 */
#ifndef PNMN_H_
#define PNMN_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PNMN_VERSION 1

typedef struct pnmn_model pnmn_model;
typedef struct pnmn_plan pnmn_plan;

int pnmn_version(void);
/* 1 when the library was built with the CUDA-core bring-up twins of the tensor-core kernels (make BRINGUP=1); the release
 * build has none and the impl_simt / PNMN_PG_SIMT / PNMN_EXEC=levels switches fail with "not supported" */
int pnmn_has_bringup_kernels(void);
const char* pnmn_last_error(void);

/* How NeuralModuleNetwork.__init__/forward classify a program token
 * (probnmn/models/nmn.py:90-111 and :207-229). */
enum pnmn_token_kind {
  PNMN_TOK_SKIP = 0,      /* @@PADDING@@ @@UNKNOWN@@ @start@ @end@ unique            nmn.py:207-208 */
  PNMN_TOK_SCENE = 1,     /* saved = out; out = ones                                  nmn.py:211-217 */
  PNMN_TOK_AND = 2,       /* intersect -> AndModule  (torch.min)                      nmn.py:98-99   */
  PNMN_TOK_OR = 3,        /* union     -> OrModule   (torch.max)                      nmn.py:100-101 */
  PNMN_TOK_COMPARE = 4,   /* equal_*, less_than, greater_than -> ComparisonModule     nmn.py:102-103 */
  PNMN_TOK_QUERY = 5,     /* query_*, exist, count -> QueryModule                     nmn.py:104-105 */
  PNMN_TOK_RELATE = 6,    /* relate[*] -> RelateModule                                nmn.py:106-107 */
  PNMN_TOK_SAME = 7,      /* same_* -> SameModule                                     nmn.py:108-109 */
  PNMN_TOK_ATTENTION = 8  /* everything else (filter_*) -> AttentionModule            nmn.py:110-111 */
};

#define PNMN_MAX_MODULE_PARAMS 12

/* Static description of a NeuralModuleNetwork (replaces the module table built in
 * probnmn/models/nmn.py:86-115 and the stem of :67-72).  All parameters live in ONE flat fp32
 * buffer; offsets are in floats.  token_param_off[v*12 + 2*i], [.. + 2*i + 1] are weight / bias
 * of the i-th conv of token v's module in forward order (conv1..conv6 | projection,conv1,conv2 |
 * conv), -1 when absent.  stem_param_off = {stem.0.weight, stem.0.bias, stem.2.weight,
 * stem.2.bias}.  module_channels is fixed at 128 and the feature map at 14x14. */
pnmn_model* pnmn_model_create(int vocab_size, const int32_t* token_kind, const int64_t* token_param_off,
                              const int64_t* stem_param_off, int in_channels);
void pnmn_model_destroy(pnmn_model* m);
/* floats (4-byte units) of scratch needed for the fp16-packed forward + dgrad weight tiles */
int64_t pnmn_model_packed_floats(const pnmn_model* m);

/* Program compiler: replaces the per-sample python interpreter with its B device->host syncs
 * (probnmn/models/nmn.py:191-238).  programs_host: int64 [B][L] prefix-order token ids (host).
 * Validity follows the reference's bare-except semantics exactly. */
pnmn_plan* pnmn_plan_create(const pnmn_model* m, const int64_t* programs_host, int batch, int length,
                            int need_grad);
/* The same with flags.  PNMN_PLAN_INPUT_BY_ROW: the stem input of sample n is unit n of pnmn_buffers.ain whether or not its
 * program is valid (default: the units of the valid samples are packed) -- for a forward pass whose program-independent part
 * ran ahead of the programs (pnmn_nmn_prestage). */
#define PNMN_PLAN_INPUT_BY_ROW 1
/* PNMN_PLAN_FORWARD_HALF (with need_grad = 0): the plan stands in for the FORWARD pass of the need_grad = 1 plan of the same
 * programs and flags -- every forward tensor has the same place in the arenas in both, and pnmn_plan_sizes of this plan
 * covers what the full plan's backward pass will allocate -- so a caller whose programs only become known inside the step
 * (joint training: they are sampled by the generator, modules/elbo.py:230-239) can launch pnmn_nmn_forward from this plan,
 * which compiles in half the time, compile the full plan meanwhile on another thread, and run pnmn_nmn_backward from that
 * one on the same buffers, `blob` pointing to the full plan's uploaded tables. */
#define PNMN_PLAN_FORWARD_HALF 2
pnmn_plan* pnmn_plan_create_ex(const pnmn_model* m, const int64_t* programs_host, int batch, int length, int need_grad,
                               int flags);
void pnmn_plan_destroy(pnmn_plan* p);
/* Cap the executor's persistent grid for this plan (0 = two CTAs per SM, the default).  A batch may be compiled as several
 * independent plans -- on several host threads, which is what shortens the time between knowing the programs and launching
 * the executor -- whose passes then run side by side on different streams and share the device's CTA slots. */
int pnmn_plan_set_exec_ctas(pnmn_plan* p, int max_ctas);
/* Optional: copy the plan's task tables into `device_blob` (PNMN_SZ_BLOB bytes, caller-owned) on `stream` ahead of time;
 * a pnmn_nmn_forward whose pnmn_buffers.blob is the same pointer then skips its own upload (the caller orders the
 * forward's stream after this one).  Lets an input pipeline keep every host -> device copy off the compute stream. */
int pnmn_plan_upload(pnmn_plan* p, void* device_blob, void* stream);
/* valid[b] = 1 if the reference would execute program b without raising (nmn.py:231-238) */
int pnmn_plan_valid(const pnmn_plan* p, uint8_t* valid);

enum pnmn_size_slot {
  PNMN_SZ_ARENA16 = 0, /* floats, P16 plane arena (must be zero-filled when (re)allocated); units hold
                          32 fp32 planes + 16 fp16 shadow half planes */
  PNMN_SZ_ARENA18 = 1, /* floats, P18 plane arena (zero-filled) */
  PNMN_SZ_ARENA22 = 2, /* floats, P22 plane arena (zero-filled) */
  PNMN_SZ_MAPS = 3,    /* floats, attention maps (zero-filled) */
  PNMN_SZ_DMAPS = 4,   /* floats, attention-map gradients */
  PNMN_SZ_IDX = 5,     /* int32, SameModule argmax slots */
  PNMN_SZ_BLOB = 6,    /* bytes, device scratch for task tables */
  PNMN_SZ_AIN = 7,     /* floats, stem-input arena (features as planes + fp16 shadow; zero-filled) */
  PNMN_SZ_COUNT = 8
};
int pnmn_plan_sizes(const pnmn_plan* p, int64_t* sizes /* [PNMN_SZ_COUNT] */);
/* statistics: [0] #valid programs, [1] #3x3 module conv instances, [2] #module tokens executed,
 * [3] forward launches, [4] backward launches, [5] algorithmic forward FLOPs of the module convs,
 * [6] forward conv CTAs, [7] wgrad CTAs, [8]/[9] forward FLOPs run by the conv kernel variants <2,2>/<1,3>
 * (stem included), [10]/[11] the same for backward (dgrad), [12] wgrad FLOPs, [13] elementwise CTAs,
 * [14] backward executor tasks, [15] byte offset inside the blob of the per-sample stem-input table
 * (int64 per sample, < 0 = invalid program: lets the caller build the validity mask on the device) */
int pnmn_plan_stats(const pnmn_plan* p, int64_t* stats /* [16] */);

typedef struct pnmn_buffers {
  float* arena16;
  float* arena18;
  float* arena22;
  float* maps;
  float* dmaps;
  int32_t* idx;
  void* blob;        /* device */
  float* packed;     /* device, pnmn_model_packed_floats() floats */
  const float* params; /* device, flat fp32 parameters */
  float* grads;      /* device, flat fp32 gradients (same offsets), may be NULL for forward */
  float* ain;        /* device, PNMN_SZ_AIN floats */
  float* scratch;    /* device, >= 64 floats (loss scale of the backward pass) */
} pnmn_buffers;

/* Forward of stem + module executor: replaces NeuralModuleNetwork.forward up to the tensor that
 * enters the classifier (probnmn/models/nmn.py:183-241).  features: device fp32 NCHW
 * [B][in_channels][14][14].  final_out: device fp32 NCHW [B][128][14][14] (zeros for invalid
 * programs, nmn.py:236).  stream: cudaStream_t. */
int pnmn_nmn_forward(pnmn_plan* p, const pnmn_buffers* bufs, const float* features, float* final_out,
                     void* stream);
/* The same with features that are already fp16 [B][C][14][14] -- the operand precision the executor keeps of its input
 * (tf32 rounding, then a saturating fp16 copy): features produced by pnmn_round_features_f16 (dst[i] = that value of
 * src[i], n % 4 == 0; e.g. a device-resident feature cache filled once, data/readers.py:63-108) give results identical to
 * the fp32 call, at half the bytes per step. */
/* Program-independent part of the forward pass, for callers that have a batch's features before its programs (in the
 * joint-training step the programs are SAMPLED by the generator's forward pass, modules/elbo.py:230-239; the weights and
 * the features are known when the step starts): packs the current weights into `packed` and lays out the features of all
 * `batch` rows in `ain` (pnmn_model_ain_floats(m, batch) floats, zero-filled once).  pack_table_dev: device copy of the
 * model's static pack-task table (pnmn_model_pack_table_bytes bytes, fetched once with pnmn_model_pack_table).  The forward
 * pass itself is then pnmn_nmn_forward(plan created with PNMN_PLAN_INPUT_BY_ROW, same packed / ain, features = NULL). */
int64_t pnmn_model_pack_table_bytes(const pnmn_model* m);
int pnmn_model_pack_table(const pnmn_model* m, void* host_out);
int64_t pnmn_model_ain_floats(const pnmn_model* m, int batch);
int pnmn_nmn_prestage(const pnmn_model* m, const void* pack_table_dev, const float* params, void* packed,
                      const void* features, int features_half, float* ain, int batch, void* stream);
int pnmn_nmn_forward_f16(pnmn_plan* plan, const pnmn_buffers* bufs, const void* features_f16, float* final_out, void* stream);
int pnmn_round_features_f16(const float* src, void* dst, int64_t n, void* stream);
/* Backward of the same: grad_final_out is d(loss)/d(final_out) [B][128][14][14]; gradients of all
 * stem and module parameters are ACCUMULATED into bufs->grads (autograd semantics). */
int pnmn_nmn_backward(pnmn_plan* p, const pnmn_buffers* bufs, const float* grad_final_out, void* stream);

/* Classifier helper: ReLU -> MaxPool2d(2,2) -> flatten (probnmn/models/nmn.py:77-79, nmn_modules.py:250-251) on the
 * channels-last output y [B][14*14][C] of the 1x1 convolution (device, fp32, bias included): pooled [B][C*7*7] in the
 * reference's (C, 7, 7) flatten order, code [B][C*7*7] bytes (argmax position | 4 if active) for the backward pass, which
 * writes gy [B][14*14][C] from g [B][C*7*7].  C must be a multiple of 64. */
int pnmn_relu_pool_fwd(const float* y, float* pooled, void* code, int64_t B, int64_t C, void* stream);
int pnmn_relu_pool_bwd(const float* g, const void* code, float* gy, int64_t B, int64_t C, void* stream);
/* pnmn_relu_pool_fwd with the 1x1 conv's bias[C] added to y first (y = the bias-free GEMM output) */
int pnmn_relu_pool_fwd_bias(const float* y, const float* bias, float* pooled, void* code, int64_t B, int64_t C, void* stream);
/* Classifier GEMMs (probnmn/models/nmn.py:75-83 and their gradients): C[M][N] (row stride ldc) = (accumulate ? C : 0) +
 * sum_k A(m, k) * B(n, k) (+ bias[n]) with A(m, k) = A[m*a_rs + k*a_ks], B(n, k) = B[n*b_rs + k*b_ks] -- any strides, so that
 * x.w^T, g.w and g^T.x are the same entry point -- fp32 in, fp32 out, computed on the tensor cores from bf16 (hi, lo) splits
 * (three tcgen05 MMAs per k step, ~16 mantissa bits per operand); csrc/gemm.cu.  Replaces the cuBLAS / cuDNN calls under
 * nn.Conv2d(k=1) and nn.Linear.  workspace: pnmn_gemm_split_workspace(M, N, K) floats, 16-byte aligned (packed operands +
 * the partial results of a split contraction, which are added in a fixed order: results are deterministic). */
int64_t pnmn_gemm_split_workspace(int M, int N, int K);
int pnmn_gemm_split(const float* A, int64_t a_rs, int64_t a_ks, const float* B, int64_t b_rs, int64_t b_ks, float* C,
                    int64_t ldc, int M, int N, int K, const float* bias, int accumulate, float* workspace,
                    int64_t workspace_floats, void* stream);

/* Answer head behind the classifier (probnmn/models/nmn.py:245-269) in one launch: per row the predicted answer (first
 * maximum of the logits; unknown_index for a row whose program is invalid), the loss (cross entropy against answers[b], or
 * -max log-probability when answers is NULL; the constant 3.33 for an invalid row) and, when correct != NULL, the number of
 * rows with prediction == answer ADDED to *correct (int64, device).  xin = the plan's per-sample stem-input table inside its
 * task-table blob (int64 [batch] at byte offset pnmn_plan_stats[15]; < 0 = invalid program); invalid [batch] receives the
 * 0 / 1 mask (the table itself is recycled with the plan).  The backward entry point writes
 * d(sum_b grad_loss[b] * loss[b]) / d(logits) [batch][num_answers]; invalid rows get zeros (their loss is a constant: the
 * in-place writes of nmn.py:258,268). */
int pnmn_answer_loss_forward(const float* logits, const int64_t* answers, const int64_t* xin, int batch, int num_answers,
                             int64_t unknown_index, int64_t* predictions, float* loss, uint8_t* invalid, int64_t* correct,
                             void* stream);
int pnmn_answer_loss_backward(const float* logits, const int64_t* answers, const uint8_t* invalid, const int64_t* predictions,
                              const float* grad_loss, int batch, int num_answers, float* grad_logits, void* stream);

/* SM partition for concurrent streams (no counterpart in the reference, which runs its models one after the other:
 * modules/elbo.py:230-275).  The module executor (pnmn_nmn_forward / _backward: exec_kernel, wgrad_tc_kernel) is
 * persistent and would otherwise own every SM for the length of a pass; with n > 0 its CTAs keep off the n highest-numbered
 * SMs, which stay available to kernels that other streams launch meanwhile (the LSTM passes of the joint-training step,
 * probnmn_clevr_b200/joint.py).  n is clamped to 3/4 of the device; 0 (default) = the executor uses the whole device.
 * Process-wide; returns the previous value.  Results do not depend on it. */
int pnmn_set_reserved_sms(int n);

/* kernels launched by the library so far (reset != 0 clears the counter); bench.py reports it as "gpu_launches" */
long long pnmn_launch_count(int reset);

/* Optional per-launch device timing (CUDA events on the launching stream), used by bench.py for the
 * roofline of the dominant kernel.  ms / launches have 8 slots: {elementwise, conv<2 samples x 2 tiles>,
 * conv<1 x 3>, wgrad, bias_grad, weight pack, feature layout, other}.  Reading synchronises and clears. */
int pnmn_profile_enable(int on);
int pnmn_profile_read(double* ms, int64_t* launches);

/* ---- LSTM seq2seq: ProgramGenerator (probnmn/models/program_generator.py:27-59 on top of
 * probnmn/modules/seq2seq_base.py:49-341 and AllenNLP 0.9.0 SimpleSeq2Seq) ---------------------------
 * All parameters live in one flat fp32 buffer; the offsets (in floats) name the reference's state-dict
 * entries.  hidden == input size == 256, a 2-layer encoder, vocabularies of at most 128 entries. */
typedef struct pnmn_pg_desc {
  int32_t vocab_src, vocab_tgt, hidden, num_layers;
  int64_t src_embed;                                          /* _source_embedder.token_embedder_tokens.weight (Vs,H) */
  int64_t enc_w_ih[2], enc_w_hh[2], enc_b_ih[2], enc_b_hh[2]; /* _encoder._module.{weight,bias}_{ih,hh}_l{0,1}        */
  int64_t tgt_embed;                                          /* _target_embedder.weight (Vt,H)                       */
  int64_t dec_w_ih, dec_w_hh, dec_b_ih, dec_b_hh;             /* _decoder_cell.*; weight_ih (4H,2H) = [attended|embed] */
  int64_t out_w, out_b;                                       /* _output_projection_layer.{weight (Vt,H), bias}       */
} pnmn_pg_desc;

/* bytes of device scratch for one forward (+ backward when need_grad); must be ZERO-FILLED whenever it is
 * (re)allocated or any of (batch, tq, tp, steps, need_grad) changes; -1 on bad arguments */
int64_t pnmn_pg_workspace_bytes(const pnmn_pg_desc* m, int batch, int tq, int tp, int steps, int need_grad);

/* Seq2SeqBase.forward (seq2seq_base.py:101-155) + _forward_loop (:157-276): boundary tokens, 2-layer LSTM encoder,
 * attention decoder, greedy (sampling = 0, torch.max) or categorical (sampling = 1, pad/unk/start never drawn) token
 * choice, _trim_predictions (:278-293) and the per-row loss: teacher-forced sequence cross entropy when `target`
 * is given (:247-254, :333-341; steps must be tp + 1), else the negated length-normalised log-probability of the
 * chosen tokens (:235-244; steps = max_decoding_steps).  source: device int64 [batch][tq] zero-padded question
 * tokens WITHOUT boundaries; target: device int64 [batch][tp] or NULL.  Outputs (device): raw_predictions and
 * predictions int64 [batch][steps] (before / after trimming), loss fp32 [batch], logits fp32
 * [batch][steps][vocab_tgt] (optional, may be NULL).  seed: Philox key of this call's sampling stream. */
int pnmn_pg_forward(const pnmn_pg_desc* m, const float* params, const int64_t* source, const int64_t* target, int batch,
                    int tq, int tp, int steps, int sampling, uint64_t seed, int need_grad, void* workspace,
                    int64_t* raw_predictions, int64_t* predictions, float* loss, float* logits, void* stream);
/* The same pass over teacher-forced AND free-running rows at once (the reference calls the generator once for the rows
 * without program supervision, free-running, and once for the rows with it, teacher-forced:
 * trainers/joint_training_trainer.py:139-144,164-168 / modules/elbo.py:230-233): rows with row_teacher[b] != 0 (device,
 * [batch] bytes) are teacher-forced on target[b]; the others decode freely by categorical sampling for free_steps steps and
 * take the sampled-sequence loss (their target row is ignored, predictions beyond free_steps are padding).  Decoding runs
 * tp + 1 steps; outputs are [batch][tp + 1].  Per row the results equal those of the two separate calls; the backward pass
 * is pnmn_pg_backward with steps = tp + 1, teacher = 1. */
int pnmn_pg_forward_mixed(const pnmn_pg_desc* m, const float* params, const int64_t* source, const int64_t* target,
                          const uint8_t* row_teacher, int free_steps, int batch, int tq, int tp, uint64_t seed, int need_grad,
                          void* workspace, int64_t* raw_predictions, int64_t* predictions, float* loss, float* logits,
                          void* stream);
/* autograd of the above: grad_loss is d(objective)/d(loss) [batch]; parameter gradients are ACCUMULATED into grads
 * (same offsets as params).  Must follow a pnmn_pg_forward with need_grad = 1 on the same workspace and sizes. */
int pnmn_pg_backward(const pnmn_pg_desc* m, const float* params, float* grads, const float* grad_loss, int batch, int tq,
                     int tp, int steps, int teacher, void* workspace, void* stream);

/* ---- ProgramPrior (probnmn/models/program_prior.py:16-155): 2-layer LSTM language model over program tokens with tied
 * input / output embeddings.  Offsets (in floats) name the reference's state-dict entries inside one flat fp32 buffer. */
typedef struct pnmn_prior_desc {
  int32_t vocab, hidden, num_layers, pad_;
  int64_t embed;                              /* _embedder.token_embedder_programs.weight == _output_layer.weight (V,H) */
  int64_t w_ih[2], w_hh[2], b_ih[2], b_hh[2]; /* _encoder._module.{weight,bias}_{ih,hh}_l{0,1}                          */
  int64_t proj;                               /* _projection_layer.weight (input_size, hidden_size), no bias            */
} pnmn_prior_desc;
/* bytes of ZERO-FILLED device scratch for one forward; -1 on bad arguments */
int64_t pnmn_prior_workspace_bytes(const pnmn_prior_desc* m, int batch, int length);
/* ProgramPrior.forward (program_prior.py:80-155): programs: device int64 [batch][length] zero-padded WITHOUT boundary
 * tokens.  Outputs (device): loss fp32 [batch] = teacher-forced sequence cross entropy of @start@ p_1..p_m @end@
 * (sum over tokens / (count + 1e-13)), predictions int64 [batch][length + 1] = one categorical draw per position from
 * the next-token distribution with pad / unk / start zeroed, times the mask (:119-139), logits fp32
 * [batch][length + 1][vocab] (optional).  Forward only: in the joint-training step the prior is frozen and only feeds the
 * REINFORCE reward (trainers/joint_training_trainer.py:107-112, modules/elbo.py:256). */
int pnmn_prior_forward(const pnmn_prior_desc* m, const float* params, const int64_t* programs, int batch, int length,
                       uint64_t seed, void* workspace, int64_t* predictions, float* loss, float* logits, void* stream);

/* ---- optimiser side of the joint-training step ------------------------------------------------------------------------
 * Element-wise gradient clamp to [-clamp, clamp] (trainers/joint_training_trainer.py:182-188; clamp <= 0 disables it)
 * fused with one Adam step (torch.optim.Adam, amsgrad off, as built in trainers/_trainer.py:103-108 and stepped at :193)
 * over n consecutive floats: params / grads / exp_avg / exp_avg_sq are 16-byte-aligned device arrays, `step` is the
 * 1-based step count of these parameters.  write_clamped_grad != 0 also stores the clamped gradient (the reference clamps
 * parameter.grad in place). */
int pnmn_clamp_adam(float* params, float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, int64_t step, double lr,
                    double beta1, double beta2, double eps, double weight_decay, double clamp, int write_clamped_grad,
                    void* stream);
/* REINFORCE reward, moving-average baseline and ELBO scalars of modules/elbo.py in one launch, nothing leaves the device
 * (the reference reads `centered_reward.mean().item()` every step, elbo.py:33).  Inputs: per-row losses (device fp32 [n];
 * nmn_loss NULL = question coding, elbo.py:128-163).  mode 0 = "ours" (:256-275), 1 = "baseline" (:241-251).
 * baseline: device fp32 [1], read then updated (b += decay * mean(reward - b)); centered: device fp32 [n] = reward - b
 * (before the update); stats: device fp32 [5] = {reconstruction_likelihood, kl_divergence, elbo, reinforce_reward,
 * mean nmn loss}. */
int pnmn_elbo_glue(const float* pg_loss, const float* qr_loss, const float* prior_loss, const float* nmn_loss, int n,
                   float beta, float gamma, float baseline_decay, int mode, float* baseline, float* centered, float* stats,
                   void* stream);

/* ---- bring-up entry points used by tests/ (kernel-level parity against torch) ---------------- */
/* byte offsets of {encoder outputs, layer-0 h, layer-1 (+decoder) h, layer-1 c, decoder c, source table, target table,
 * attention probabilities, logits} inside a pnmn_pg workspace, then {state slot bytes, Ts, padded batch} */
int pnmn_pg_debug_layout(const pnmn_pg_desc* m, int batch, int tq, int tp, int steps, int need_grad, int64_t* out /* [12] */);
int pnmn_debug_launch_conv(const void* tasks_host, int n_tasks, const void* cfgs_host, int n_cfgs,
                           int variant, int impl_simt, void* stream);
int pnmn_debug_launch_wgrad(const void* tasks_host, int n_tasks, const void* insts_host, int n_insts,
                            int impl_simt, void* stream);
int pnmn_debug_pack(const void* pack_tasks_host, int n_tasks, int total_tiles, const float* params,
                    void* packed, void* stream);
int pnmn_debug_nchw_to_planes(const float* src, float* dst, int batch, int channels,
                              int64_t dst_sample_stride, void* stream);
int pnmn_debug_launch_elt(const void* tasks_host, int n_tasks, void* stream);
/* per-task timestamps of the persistent executor (32 int64 per task, layout in csrc/exec.cu: fetched, ready, body done, published,
 * SM id, type|n_samp<<8|n_mt<<16, MMA k-steps, flags); forward tasks first, then backward; NULL disables */
int pnmn_debug_set_trace(void* device_buffer, int64_t capacity_tasks);
/* task metadata of the persistent executor lists (pass 0 forward / 1 backward): 16 int32 per task
 * {type, n_deps, deps[10], conv: n_samp, n_mt, MMAs per accumulator, flags | elementwise: op, part, 0, 0};
 * returns the number of tasks (host only, no device work) */
int64_t pnmn_debug_plan_meta(const pnmn_plan* p, int pass, int32_t* out, int64_t cap_tasks);
/* Raw records of a plan: pass 0 / 1 = the 128-byte task records of the forward / backward list (device pointers still
 * symbolic), pass 2 = the 48-byte convolution configurations.  Returns the count; `out` may be NULL. */
int64_t pnmn_debug_plan_records(const pnmn_plan* p, int pass, void* out, int64_t cap_records);
/* attention maps (1-channel module outputs; probnmn/modules/nmn_modules.py:82-87,160-168,200-208 and the 1-channel results
 * of And / Or, :25-27,43-45) of a plan, for parity tests: 4 int32 per record {sample, index of the module call inside the
 * sample's program in execution order, token id, map unit}; after pnmn_nmn_forward map unit u is the top-left 14 x 14 block
 * of the 16 x 16 fp32 grid at pnmn_buffers.maps + 256 * u.  Returns the number of records (host only). */
int64_t pnmn_debug_plan_maps(const pnmn_plan* p, int32_t* out, int64_t cap_records);
/* CUDA-graph cache of the LSTM passes: {cached keys, keys with an instantiated graph, keys whose capture failed (plain
 * launches are used for those), graph launches so far} */
int pnmn_debug_graph_stats(int64_t* out /* [4] */);
/* accumulated host-side milliseconds spent in {pnmn_plan_create, pnmn_nmn_forward, pnmn_nmn_backward} and the
 * number of plans created since the last call (reading clears) */
int pnmn_debug_host_times(double* ms);

#ifdef __cplusplus
}
#endif
#endif /* PNMN_H_ */
