// ProgramPrior.forward (probnmn/models/program_prior.py:80-155): a 2-layer LSTM language model over program tokens with
// tied input / output embeddings, evaluated with teacher forcing; the per-row sequence cross entropy is the
// log p(z) term of the REINFORCE reward of the joint-training step (modules/elbo.py:256,265-270).
//
// The packed 2-layer LSTM runs on the SAME tcgen05 step-GEMM kernels as the seq2seq encoder (seq2seq_gemm.cu:
// split-fp16 operands, gates fused in the epilogue); what is new here is the token preparation (both boundary tokens
// stay: the model reads @start@ p_1..p_m @end@), the folded output head
//     logits_t = E (W_p h_t)  =  (E W_p) h_t          E: tied embedding (V,256), W_p: _projection_layer (256,256)
// (one V x 256 matrix per call instead of two GEMVs per position), and a row kernel for log-softmax / cross entropy /
// the categorical "predictions" (:119-137; pad / unk / start are never drawn).
#include <cstring>
#include <string>

#include "../../include/pnmn.h"
#include "seq2seq.h"

using namespace pnmn;

namespace pnmn { void set_last_error(const std::string& s); void count_launches(int n); }

namespace {

int fail(const std::string& s) {
  pnmn::set_last_error(s);
  return 1;
}
#define CUDA_OK(x)                                                                        \
  do {                                                                                    \
    cudaError_t e_ = (x);                                                                 \
    if (e_ != cudaSuccess) return fail(std::string(#x) + ": " + cudaGetErrorString(e_)); \
  } while (0)

constexpr int kPad = 0, kStart = 2, kEnd = 3;

struct PriorLayout {
  int B, Bp, T, Ts, V;
  int64_t slotf, slotop;
  int64_t tok, len, P0, P1, fold, packed, pk_hh0, pk_1, h0f, c0f, h0op, out0op, h1f, c1f, h1op, enc, total;
};

PriorLayout prior_layout(const pnmn_prior_desc* m, int B, int T) {
  PriorLayout L;
  std::memset(&L, 0, sizeof(L));
  L.B = B; L.Bp = (B + 127) / 128 * 128; L.T = T; L.Ts = T + 2; L.V = m->vocab;
  L.slotf = static_cast<int64_t>(L.Bp) * kSH;
  L.slotop = 2 * L.slotf;
  int64_t o = 0;
  auto take = [&](int64_t bytes) { const int64_t r = o; o += (bytes + 255) / 256 * 256; return r; };
  L.tok = take(4ll * L.Bp * L.Ts); L.len = take(4ll * L.Bp);
  L.P0 = take(4ll * L.V * kSG); L.P1 = take(4ll * kSG); L.fold = take(4ll * L.V * kSH);
  int64_t ph = 0;
  auto takeh = [&](int64_t n, int64_t k) { const int64_t r = ph; ph += 2 * n * k; return r; };
  L.pk_hh0 = takeh(kSG, kSH); L.pk_1 = takeh(kSG, 2 * kSH);
  L.packed = take(2 * ph);
  L.h0f = take(4 * L.slotf * (L.Ts + 1)); L.c0f = take(4 * L.slotf * (L.Ts + 1));
  L.h0op = take(2 * L.slotop * (L.Ts + 1)); L.out0op = take(2 * L.slotop * L.Ts);
  L.h1f = take(4 * L.slotf * (L.Ts + 1)); L.c1f = take(4 * L.slotf * (L.Ts + 1));
  L.h1op = take(2 * L.slotop * (L.Ts + 1));
  L.enc = take(4ll * L.Bp * L.Ts * kSH);
  L.total = o;
  return L;
}

int check(const pnmn_prior_desc* m, int B, int T) {
  if (!m) return fail("pnmn_prior: NULL model description");
  if (m->hidden != kSH) return fail("pnmn_prior: the B200 LSTM kernels are built for input_size = hidden_size = 256");
  if (m->num_layers != 2) return fail("pnmn_prior: the LSTM must have 2 layers");
  if (m->vocab < 4 || m->vocab > kSMaxV) return fail("pnmn_prior: vocabulary must have 4..128 entries");
  if (B < 1) return fail("pnmn_prior: empty batch");
  if (T < 0 || T + 2 > kSMaxT) return fail("pnmn_prior: programs longer than 62 tokens are not supported");
  return 0;
}

template <class T>
T* at(void* ws, int64_t off) { return reinterpret_cast<T*>(static_cast<uint8_t*>(ws) + off); }

// tokens: [B][T] zero-padded -> [B][T+2] = @start@ p_1..p_m @end@ 0..  (AllenNLP add_sentence_boundary_token_ids,
// program_prior.py:104-107); len[b] = m + 2 (the LSTM runs over the whole boundary-added sequence, :112-117)
__global__ void prior_tokens_kernel(const int64_t* __restrict__ programs, int B, int rows, int T, int V,
                                    unsigned long long seed, unsigned long long* __restrict__ seed_out,
                                    int* __restrict__ tok, int* __restrict__ len) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b == 0) seed_out[0] = seed;
  if (b >= B) return;
  const bool real = b < rows;   // padding rows of the last 128-row tile: empty programs, never read back
  const int Ts = T + 2;
  int m = 0;
  for (int s = 0; real && s < T; ++s) m += programs[static_cast<size_t>(b) * T + s] != kPad;
  int* row = tok + static_cast<size_t>(b) * Ts;
  row[0] = kStart;
  for (int s = 0; s < T; ++s) {
    const int64_t v = real ? programs[static_cast<size_t>(b) * T + s] : 0;
    row[1 + s] = (v < 0 || v >= V) ? 1 : static_cast<int>(v);
  }
  row[T + 1] = kPad;
  row[m + 1] = kEnd;
  int n = 0;
  for (int s = 0; s < Ts; ++s) n += row[s] != kPad;
  len[b] = n;
}

__device__ __forceinline__ float wsum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float wmax(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
  const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
  const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}
__device__ __forceinline__ float philox_uniform(unsigned long long seed, uint32_t row, uint32_t step) {
  uint32_t c[4] = {row, step, 0x9E3779B9u, 1u};
  uint32_t k0 = static_cast<uint32_t>(seed), k1 = static_cast<uint32_t>(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return static_cast<float>(c[0] >> 8) * (1.0f / 16777216.0f);
}

// One CTA per (row, position t < T+1): logits = fold . h_t, log-softmax, cross entropy against token t+1 (:139-146),
// categorical prediction (:119-137).  grid = (T + 1, B); 4 warps, warp w handles vocabulary entries w, w+4, ...
__global__ void __launch_bounds__(128) prior_head_kernel(const float* __restrict__ enc, const float* __restrict__ fold,
                                                         const int* __restrict__ tok, int B, int Ts, int V,
                                                         const unsigned long long* __restrict__ seed_ptr, float* __restrict__ nll,
                                                         int64_t* __restrict__ predictions, float* __restrict__ logits_out) {
  __shared__ float sh[kSH], slg[kSMaxV], sp[kSMaxV];
  const int t = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int W = Ts - 1;   // positions with a next token
  const float* h = enc + (static_cast<size_t>(b) * Ts + t) * kSH;
  sh[tid] = h[tid]; sh[tid + 128] = h[tid + 128];
  __syncthreads();
  for (int v = warp; v < V; v += 4) {
    const float4 w0 = *reinterpret_cast<const float4*>(fold + static_cast<size_t>(v) * kSH + lane * 8);
    const float4 w1 = *reinterpret_cast<const float4*>(fold + static_cast<size_t>(v) * kSH + lane * 8 + 4);
    const float* x = sh + lane * 8;
    float acc = w0.x * x[0];
    acc = fmaf(w0.y, x[1], acc); acc = fmaf(w0.z, x[2], acc); acc = fmaf(w0.w, x[3], acc);
    acc = fmaf(w1.x, x[4], acc); acc = fmaf(w1.y, x[5], acc); acc = fmaf(w1.z, x[6], acc); acc = fmaf(w1.w, x[7], acc);
    acc = wsum(acc);
    if (lane == 0) slg[v] = acc;
  }
  __syncthreads();
  if (warp != 0) return;
  float mx = -INFINITY;
  for (int v = lane; v < V; v += 32) mx = fmaxf(mx, slg[v]);
  mx = wmax(mx);
  float sum = 0.f;
  for (int v = lane; v < V; v += 32) { const float e = expf(slg[v] - mx); sp[v] = e; sum += e; }
  sum = wsum(sum);
  const float lse = mx + logf(sum);
  const int target = tok[static_cast<size_t>(b) * Ts + t + 1];
  if (logits_out)
    for (int v = lane; v < V; v += 32) logits_out[(static_cast<size_t>(b) * W + t) * V + v] = slg[v];
  __syncwarp();
  if (lane == 0) {
    nll[static_cast<size_t>(b) * W + t] = target != kPad ? lse - slg[target] : 0.f;
    // multinomial over softmax with pad / unk / start zeroed, then "* mask" (:139)
    float total = 0.f;
    for (int v = kStart + 1; v < V; ++v) total += sp[v];
    const float u = philox_uniform(seed_ptr[0], static_cast<uint32_t>(b), static_cast<uint32_t>(t)) * total;
    float cum = 0.f;
    int pred = V - 1;
    for (int v = kStart + 1; v < V; ++v) {
      cum += sp[v];
      if (cum > u) { pred = v; break; }
    }
    predictions[static_cast<size_t>(b) * W + t] = target != kPad ? pred : 0;
  }
}

// loss[b] = sum_t nll / (count + 1e-13): sequence_cross_entropy_with_logits(average=None) (:142-147)
__global__ void prior_loss_kernel(const float* __restrict__ nll, const int* __restrict__ tok, int B, int Ts,
                                  float* __restrict__ loss) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int W = Ts - 1;
  float tot = 0.f, n = 0.f;
  for (int t = 0; t < W; ++t) {
    const float m = tok[static_cast<size_t>(b) * Ts + t + 1] != kPad ? 1.f : 0.f;
    tot += nll[static_cast<size_t>(b) * W + t] * m;
    n += m;
  }
  loss[b] = tot / (n + 1e-13f);
}

}  // namespace

extern "C" int64_t pnmn_prior_workspace_bytes(const pnmn_prior_desc* m, int batch, int length) {
  if (check(m, batch, length)) return -1;
  const PriorLayout L = prior_layout(m, batch, length);
  return L.total + 4ll * L.Bp * (L.Ts - 1) + 256;   // + per-position nll + the Philox key
}

extern "C" int pnmn_prior_forward(const pnmn_prior_desc* m, const float* params, const int64_t* programs, int batch,
                                  int length, uint64_t seed, void* ws, int64_t* predictions, float* loss,
                                  float* logits_out, void* stream) {
  if (check(m, batch, length)) return 1;
  if (!params || !programs || !ws || !predictions || !loss) return fail("pnmn_prior_forward: NULL buffer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const PriorLayout L = prior_layout(m, batch, length);
  const int MT = L.Bp / 128;
  __half* packed = at<__half>(ws, L.packed);
  float* nll = at<float>(ws, L.total);

  unsigned long long* seed_dev = reinterpret_cast<unsigned long long*>(nll + static_cast<size_t>(L.Bp) * (L.Ts - 1));
  prior_tokens_kernel<<<(L.Bp + 127) / 128, 128, 0, st>>>(programs, L.Bp, L.B, L.T, L.V, seed, seed_dev, at<int>(ws, L.tok),
                                                          at<int>(ws, L.len));
  CUDA_OK(cudaGetLastError());
  // the LSTM over whole 128-row tiles: one CUDA graph per (workspace, parameters, shape), see seq2seq_api.cu
  GraphKey key{ws, params, 0, 4, L.Bp, L.T, 0, 0, 0, 0, 0, L.V, L.V};
  const int rc = run_graphed_pass(key, st, [&](cudaStream_t st) -> int {
  {
    PackJobs J;
    std::memset(&J, 0, sizeof(J));
    auto job = [](int N, int K, int mode, int split, int64_t s0, int ld0, int64_t s1, int ld1, int64_t dst) {
      PackJob j; j.N = N; j.K = K; j.mode = mode; j.split = split; j.src0 = s0; j.src1 = s1; j.ld0 = ld0; j.ld1 = ld1; j.dst = dst;
      return j;
    };
    J.j[0] = job(kSG, kSH, 0, kSH, m->w_hh[0], kSH, m->w_hh[0], kSH, L.pk_hh0);
    J.j[1] = job(kSG, 2 * kSH, 0, kSH, m->w_ih[1], kSH, m->w_hh[1], kSH, L.pk_1);
    CUDA_OK(launch_pack_seq(J, 2, params, packed, st));
  }
  {
    SimtGemm g;
    std::memset(&g, 0, sizeof(g));
    g.alpha = 1.f; g.sak = 1; g.sbk = 1;
    // P0[v] = E[v] . W_ih0^T + b_ih0 + b_hh0
    g.A = params + m->embed; g.sam = kSH; g.M = L.V; g.K = kSH; g.N = kSG; g.ldc = kSG;
    g.B = params + m->w_ih[0]; g.sbn = kSH;
    g.bias0 = params + m->b_ih[0]; g.bias1 = params + m->b_hh[0]; g.C = at<float>(ws, L.P0);
    CUDA_OK(launch_simt_gemm(g, st));
    g.M = 1; g.K = 0; g.bias0 = params + m->b_ih[1]; g.bias1 = params + m->b_hh[1]; g.C = at<float>(ws, L.P1);
    CUDA_OK(launch_simt_gemm(g, st));
    // fold[v][j] = sum_e E[v][e] * W_p[e][j]      (W_p: _projection_layer.weight (input_size, hidden_size))
    std::memset(&g, 0, sizeof(g));
    g.alpha = 1.f;
    g.A = params + m->embed; g.sam = kSH; g.sak = 1; g.M = L.V; g.K = kSH; g.N = kSH; g.ldc = kSH;
    g.B = params + m->proj; g.sbk = kSH; g.sbn = 1; g.C = at<float>(ws, L.fold);
    CUDA_OK(launch_simt_gemm(g, st));
  }
  GemmArgs g;
  std::memset(&g, 0, sizeof(g));
  g.B = L.Bp; g.chunks_per_src = 4; g.a_K[0] = g.a_K[1] = kSH; g.a_lo[0] = g.a_lo[1] = L.slotf; g.h_op_lo = L.slotf;
  g.out_op_lo = L.slotf;
  g.len = at<int>(ws, L.len);
  auto l0 = [&](int t) {
    GemmArgs a = g;
    a.t = t; a.K = kSH;
    a.a[0] = at<__half>(ws, L.h0op) + t * L.slotop; a.a[1] = nullptr;
    a.w = packed + L.pk_hh0; a.w_lo = static_cast<int64_t>(kSG) * kSH;
    a.table = at<float>(ws, L.P0); a.tok = at<int>(ws, L.tok) + t; a.tok_stride = L.Ts;
    a.h_prev = at<float>(ws, L.h0f) + t * L.slotf; a.c_prev = at<float>(ws, L.c0f) + t * L.slotf;
    a.h_out = at<float>(ws, L.h0f) + (t + 1) * L.slotf; a.c_out = at<float>(ws, L.c0f) + (t + 1) * L.slotf;
    a.h_op = at<__half>(ws, L.h0op) + (t + 1) * L.slotop;
    a.gates = nullptr; a.out_f = nullptr; a.out_op = at<__half>(ws, L.out0op) + t * L.slotop;
    return a;
  };
  auto l1 = [&](int t) {
    GemmArgs a = g;
    a.t = t; a.K = 2 * kSH;
    a.a[0] = at<__half>(ws, L.out0op) + t * L.slotop; a.a[1] = at<__half>(ws, L.h1op) + t * L.slotop;
    a.w = packed + L.pk_1; a.w_lo = static_cast<int64_t>(kSG) * 2 * kSH;
    a.table = at<float>(ws, L.P1); a.tok = nullptr; a.tok_stride = 0;
    a.h_prev = at<float>(ws, L.h1f) + t * L.slotf; a.c_prev = at<float>(ws, L.c1f) + t * L.slotf;
    a.h_out = at<float>(ws, L.h1f) + (t + 1) * L.slotf; a.c_out = at<float>(ws, L.c1f) + (t + 1) * L.slotf;
    a.h_op = at<__half>(ws, L.h1op) + (t + 1) * L.slotop;
    a.gates = nullptr;
    a.out_f = at<float>(ws, L.enc) + static_cast<int64_t>(t) * kSH; a.out_stride = static_cast<int64_t>(L.Ts) * kSH;
    a.out_op = nullptr;
    return a;
  };
  for (int k = 0; k <= L.Ts; ++k) {   // wavefront over the two layers (seq2seq_api.cu)
    GemmPair pr;
    pr.m_tiles = MT; pr.n_tiles[0] = pr.n_tiles[1] = kSG / 64;
    if (k < L.Ts && k >= 1) { pr.count = 2; pr.g[0] = l0(k); pr.g[1] = l1(k - 1); }
    else { pr.count = 1; pr.g[0] = k < L.Ts ? l0(k) : l1(k - 1); pr.g[1] = pr.g[0]; }
    CUDA_OK(launch_step_gemm_pair(pr, EPI_LSTM, false, st));
  }
    return 0;
  });
  if (rc) return rc;
  prior_head_kernel<<<dim3(L.Ts - 1, L.B), 128, 0, st>>>(at<float>(ws, L.enc), at<float>(ws, L.fold), at<int>(ws, L.tok), L.B,
                                                         L.Ts, L.V, seed_dev, nll, predictions, logits_out);
  CUDA_OK(cudaGetLastError());
  prior_loss_kernel<<<(L.B + 127) / 128, 128, 0, st>>>(nll, at<int>(ws, L.tok), L.B, L.Ts, loss);
  CUDA_OK(cudaGetLastError());
  pnmn::count_launches(1 + 1 + 3 + (L.Ts + 1) + 2);
  return 0;
}
