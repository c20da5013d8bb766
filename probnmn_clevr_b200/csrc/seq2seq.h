// Internal types of the LSTM seq2seq (ProgramGenerator) path.  Nothing here is part of the C ABI
// (include/pnmn.h); the entry points live in seq2seq_api.cu.
//
// Reference: probnmn/modules/seq2seq_base.py:101-276 on top of AllenNLP 0.9.0 SimpleSeq2Seq
// (2-layer nn.LSTM encoder, dot-product attention, LSTMCell decoder, Linear projection).
//
// "Operand format": every matrix that feeds a tensor-core GEMM (hidden states, attended vectors,
// gate gradients, weights) is stored as TWO fp16 matrices hi + lo (x ~= hi + lo, 22 mantissa bits)
// in the unswizzled tcgen05 core-matrix order
//        buf[row tile][k / 8][row % tile][k % 8]          (tile = 128 rows for activations,
//                                                           64 rows for packed weights)
// so that a 64-deep K chunk of one tile is ONE contiguous bulk copy and is at once a K-major
// operand (8 rows x 16 B core matrices, LBO = tile*16 B, SBO = 128 B) and — read along the rows —
// an MN-major operand for the weight-gradient GEMMs (LBO = 128 B, SBO = tile*16 B).
// A GEMM is three MMAs per k-step: hi*hi + lo*hi + hi*lo with fp32 accumulation in TMEM, i.e. fp32-class
// products on the fp16 tensor pipe: greedy argmax tokens must match the reference's fp32 path bit for bit
// (BASELINE.json north_star), which a single 11-bit operand cannot deliver.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace pnmn {

constexpr int kSH = 256;        // hidden size == embedding size (configs/*.yml: INPUT_SIZE = HIDDEN_SIZE = 256)
constexpr int kSG = 4 * kSH;    // gate width
constexpr int kSMaxV = 128;     // target vocabulary bound of the row kernels (programs: 44, questions: ~93)
constexpr int kSMaxT = 64;      // source length bound (questions <= 45 tokens + @end@)

// halves offset of element (r, k) of an activation operand with K features (128-row tiles)
__host__ __device__ inline size_t op_off(int r, int k, int K) {
  return (static_cast<size_t>(r >> 7) * (K >> 3) + (k >> 3)) * 1024 + static_cast<size_t>(r & 127) * 8 + (k & 7);
}
// halves offset of element (n, k) of a packed weight with K features (64-row tiles)
__host__ __device__ inline size_t wp_off(int n, int k, int K) {
  return (static_cast<size_t>(n >> 6) * (K >> 3) + (k >> 3)) * 512 + static_cast<size_t>(n & 63) * 8 + (k & 7);
}

// Programmatic dependent launch: the ~110 dependent per-step kernels of a pass are launched with
// cudaLaunchAttributeProgrammaticStreamSerialization; each one lets its successor start at once
// (pdl_launch_dependents) and blocks (pdl_wait) only where it first touches data of its predecessor, so the launch
// gap, barrier / TMEM set-up and the weight prefetch of step t+1 overlap with step t.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

#ifdef __CUDACC__
template <class... KArgs, class... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl,
                              Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, args...);
}
#endif
bool seq_use_pdl();   // PNMN_PG_NOPDL=1 disables it (diagnostics)

__device__ __forceinline__ void split_f16(float x, __half& hi, __half& lo) {
  x = fminf(fmaxf(x, -65504.f), 65504.f);
  hi = __float2half_rn(x);
  lo = __float2half_rn(x - __half2float(hi));
}

// ---- tensor-core step GEMM ------------------------------------------------------------------------
//   acc[b][n] = sum_k A[b][k] * W[n][k]        b: batch rows (128 per CTA), n: 64 per CTA
// A is the concatenation of up to two operand buffers along K; W is a packed weight.
enum GemmEpilogue : int { EPI_LSTM = 0, EPI_DGRAD = 1 };

struct GemmArgs {
  const __half* a[2];     // operand buffers (hi; lo at + a_lo[i])
  int64_t a_lo[2];        // halves
  int a_K[2];             // features of each buffer (tile stride = a_K*128 halves)
  int chunks_per_src;     // 64-wide K chunks taken from a[0] before a[1]
  int K;                  // total K (multiple of 64)
  const __half* w;        // packed weight [N/64][K/8][64][8] (hi; lo at + w_lo)
  int64_t w_lo;
  int B;                  // real rows (rows >= B are never written)
  int t;                  // time step (mask = t < len[b]); len == nullptr: every row is valid
  const int* len;
  // ---- EPI_LSTM: n = gate*16 + u inside the CTA's tile nt <-> hidden unit j = nt*16 + u
  const float* table;     // [V][1024] additive term (input projection + biases), row tok[b*tok_stride]
  const int* tok;         // nullptr: row 0
  int tok_stride;
  const float* h_prev;    // fp32 [Bp][256]
  const float* c_prev;
  float* h_out;           // fp32 [Bp][256]
  float* c_out;
  __half* h_op;           // operand copy of h_out (K = 256); lo at + h_op_lo
  int64_t h_op_lo;
  float* gates;           // [Bp][1024] activated gates i,f,g,o (saved for backward) or nullptr
  float* out_f;           // masked output (0 beyond the length): element (b, j) at out_f[b*out_stride + j], or nullptr
  int64_t out_stride;
  __half* out_op;         // operand copy of the masked output (K = 256) or nullptr
  int64_t out_op_lo;
  // ---- EPI_DGRAD: column n of the output belongs to out[n / 256][b][n % 256]
  float* out[2];
  int keep_masked[2];     // masked rows: 1 = leave out[i] untouched (carried state), 0 = write zeros
  const float* scale;     // {loss scale, 1 / loss scale}: results are multiplied by scale[1]
};

cudaError_t launch_step_gemm(const GemmArgs& g, int epilogue, int n_tiles, int m_tiles, bool simt, cudaStream_t st);
// Up to two independent step GEMMs in ONE launch (blockIdx.y = which * m_tiles + m tile): step k of LSTM layer 0 together
// with step k-1 of layer 1 (a wavefront over the two encoder layers: Ts + 1 dependent launches instead of 2 Ts), and the
// same for their data-gradient GEMMs in the backward pass.
struct GemmPair {
  GemmArgs g[2];
  int n_tiles[2];
  int m_tiles;
  int count;   // 1 or 2
};
cudaError_t launch_step_gemm_pair(const GemmPair& p, int epilogue, bool simt, cudaStream_t st);

// ---- weight-gradient GEMM (contraction over batch rows and time) ------------------------------------
//   dW[g][k] += scale[1] * sum_{t, b} dG[t][b][g] * X[t][b][k]
struct WgradSeqArgs {
  const __half* dg;       // [T][operand Bp x 1024] (hi; lo at + dg_lo), step stride dg_step (halves)
  int64_t dg_lo, dg_step;
  const __half* x;        // [T][operand Bp x 256]
  int64_t x_lo, x_step;
  int T, m_tiles;         // time steps, 128-row tiles per step
  float* dw;              // fp32 [1024][ld] (+ column offset already applied)
  int ld;
  const float* scale;
};
cudaError_t launch_wgrad_seq(const WgradSeqArgs& a, bool simt, cudaStream_t st);

// ---- CUDA-core kernels (seq2seq_rows.cu) -----------------------------------------------------------
struct PackJob {
  int N, K;               // packed matrix shape (rows n, features k)
  int mode;               // 0: forward (rows gate-interleaved per 16 hidden units), 1: transposed for dgrad
  int split;              // forward: k < split from src0 else src1; transposed: n < split from src0 else src1
  int64_t src0, src1;     // float offsets into the flat parameter buffer
  int ld0, ld1;           // row strides of the sources
  int64_t dst;            // halves offset into the packed buffer (hi; lo at + N*K)
};
struct PackJobs { PackJob j[8]; };   // passed by value as a kernel argument
cudaError_t launch_pack_seq(const PackJobs& jobs, int n_jobs, const float* params, __half* packed, cudaStream_t st);

// C[m][n] (+)= alpha * sum_k A[m*sam + k*sak] * B[k*sbk + n*sbn] (+ bias0[n] + bias1[n])       (plain fp32 FMAs)
struct SimtGemm {
  const float* A; const float* B; float* C;
  int M, N, K;
  int64_t sam, sak, sbk, sbn, ldc;
  const float* bias0; const float* bias1;
  float alpha;
  int accumulate;
};
cudaError_t launch_simt_gemm(const SimtGemm& g, cudaStream_t st);

struct SeqDims {
  int B, Bp, Tq, Ts, Tp, S, Vs, Vt;
  int teacher;   // targets given
  int sampling;  // 1 = categorical sampling, 0 = greedy
  int free_S;    // decoding steps of FREE-RUNNING rows (== S unless teacher-forced and free-running rows share one call)
};

// token preparation (AllenNLP add_sentence_boundary_token_ids + seq2seq_base.py:128-141)
// d.B rows are prepared; rows >= `rows` (padding up to the tile size) become empty sequences.  Also stores the call's
// Philox key in device memory (seed_out), so that the launches that follow do not depend on it (CUDA graph replay).
// row_teacher (device, [rows] bytes, may be nullptr = every row follows d.teacher) -> row_mode[d.B] ints in the workspace:
// 1 = the row is teacher-forced, 0 = it decodes freely for d.free_S steps.
cudaError_t launch_prepare_tokens(const int64_t* source, const int64_t* target, const uint8_t* row_teacher, SeqDims d, int rows,
                                  unsigned long long seed, unsigned long long* seed_out, int* src, int* src_len, int* tgt,
                                  int* row_mode, cudaStream_t st);

struct DecRowArgs {
  SeqDims d;
  int t;                  // step whose attention is computed (== S: only the logits of step S-1)
  const float* h_dec;     // fp32 [S+1][Bp][256]; slot 0 = initial decoder state, slot t+1 = h_t
  const float* enc;       // fp32 [B][Ts][256]
  const int* src_len;
  const int* tgt;         // [B][Tp+2] or nullptr
  const int* row_mode;    // [B] 1 = teacher-forced row, 0 = free-running row
  const float* out_w; const float* out_b;   // [Vt][256], [Vt]
  float* logits;          // [S][B][Vt]
  float* lse;             // [S][B]
  int* pred;              // [S][B] raw predictions
  float* logp;            // [S][B] log-probability of the prediction
  int* inp;               // [S][B] decoder input token of each step
  float* attn_p;          // [S][B][Ts]
  __half* att_op;         // [S][operand Bp x 256] (hi; lo at + att_lo), step stride att_step
  int64_t att_lo, att_step;
  const unsigned long long* seed;   // device: Philox key of this call
  int defer_out;          // 1: the per-step launches skip the output projection (launch_dec_out computes all steps at once)
};
cudaError_t launch_dec_row(const DecRowArgs& a, cudaStream_t st);
cudaError_t launch_dec_out(const DecRowArgs& a, cudaStream_t st);
cudaError_t launch_dec_out_post(const DecRowArgs& a, cudaStream_t st);   // logits already in a.logits (tensor-core GEMM)

struct FinalizeArgs {
  SeqDims d;
  int rows;               // real batch rows (<= d.B); only these are written to the caller's outputs
  const int* pred; const float* logp; const float* logits; const float* lse; const int* tgt;
  const int* row_mode;    // [B] as in DecRowArgs
  int64_t* raw_out; int64_t* pred_out; float* loss; float* logits_out;
  float* coef;            // [S][B] d(loss_b)/d(-logprob or nll at step t), before the incoming gradient
  int* label;             // [S][B] class whose one-hot enters dlogits
};
cudaError_t launch_finalize(const FinalizeArgs& a, cudaStream_t st);

struct DecBwdRowArgs {
  SeqDims d;
  int t;                  // step whose logits/cell backward runs (-1: only the attention backward of step 0)
  int do_attn;            // also run the attention backward of step t+1
  const float* grad_loss; // [B]
  const float* coef; const int* label;
  const float* logits; const float* lse;
  const float* out_w;
  const float* h_dec;     // [S+1][Bp][256]
  const float* c_dec;     // [S+1][Bp][256]
  const float* gates;     // [S][Bp][1024]
  const float* enc; const int* src_len;
  const float* attn_p;
  float* dlogits;         // [S][B][Vt]
  float* dhp;             // [S][Bp][256] d(h_t) through the output projection (launch_dec_bwd_proj, all steps at once)
  float* dh; float* dc;   // carried gradients [Bp][256]
  const float* datt;      // [Bp][256] (written by the dgrad GEMM of step t+1)
  float* denc;            // [B][Ts][256]
  __half* dg_op;          // [S][operand Bp x 1024]
  int64_t dg_lo, dg_step;
  const float* scale;
};
cudaError_t launch_dec_bwd_row(const DecBwdRowArgs& a, cudaStream_t st);
cudaError_t launch_dec_bwd_proj(const DecBwdRowArgs& a, cudaStream_t st);

struct EncCellBwdArgs {
  int B, Bp, t, Ts;
  const int* src_len;
  const float* gates;     // [Bp][1024] of this step
  const float* c_prev;    // [Bp][256]
  const float* c_cur;
  const float* dext;      // extra gradient on h_t: element (b, j) at dext[b*dext_stride + j] (nullptr: none)
  int64_t dext_stride;
  float* dh; float* dc;   // carried
  __half* dg_op; int64_t dg_lo;   // operand of this step
  const float* scale;
};
cudaError_t launch_enc_cell_bwd(const EncCellBwdArgs& a, cudaStream_t st);
struct EncCellBwdPair { EncCellBwdArgs a[2]; int count; };   // blockIdx.y selects: the two encoder layers in one launch
cudaError_t launch_enc_cell_bwd_pair(const EncCellBwdPair& p, cudaStream_t st);

// dP[v][g] (+)= scale[1] * sum_{(t,b): tok == v} dG[t][b][g];  tok == nullptr: everything goes to row 0
struct TableGradArgs {
  const __half* dg; int64_t dg_lo, dg_step;
  const int* tok; int64_t tok_step; int tok_stride;   // token of (t, b) = tok[t*tok_step + b*tok_stride]
  int T, B, V;
  float* dP;              // [V][1024]
  const float* scale;
};
cudaError_t launch_table_grad(const TableGradArgs& a, cudaStream_t st);
// out[g] += sum_v dP[v][g] for two destinations (bias_ih, bias_hh)
cudaError_t launch_bias_from_table(const float* dP, int V, float* db0, float* db1, cudaStream_t st);
// scale[0] = 2^k with max|g| * 2^k in [2^9, 2^10), scale[1] = 1 / scale[0]
cudaError_t launch_seq_loss_scale(const float* grad_loss, int B, float* scale, cudaStream_t st);
// staged[b] = b < rows ? grad_loss[b] : 0 for b < padded (the padding rows of a tile carry no gradient)
cudaError_t launch_stage_grad_loss(const float* grad_loss, int rows, int padded, float* staged, cudaStream_t st);
// dst[i] += src[i]
cudaError_t launch_accumulate(float* dst, const float* src, int64_t n, cudaStream_t st);

// CUDA-graph cache of a pass (seq2seq_api.cu): `body` issues the launches of the pass on the stream it is given; after one
// plain run per key they are captured once and replayed with a single cudaGraphLaunch on `st`.  pass: 0 forward, 1 backward,
// 2 / 3 the CUDA-core twins, 4 ProgramPrior forward.
struct GraphKey {
  const void* ws; const void* params; int device, pass, Bp, Tq, Tp, S, sampling, teacher, need_grad, Vs, Vt;
};

}  // namespace pnmn

#ifdef __CUDACC__
#include <functional>
namespace pnmn {
int run_graphed_pass(const GraphKey& key, cudaStream_t st, const std::function<int(cudaStream_t)>& body);
}
#endif
