// CUDA-core kernels of the LSTM seq2seq path: everything that is row-local (one batch row per CTA) or
// tiny — token boundaries, the decoder's output projection / softmax / token choice / dot-product attention,
// their backward, the LSTM cell backward, table (embedding-projection) gradients, weight packing and a
// small strided fp32 GEMM for the |V| x 1024 tables.
//
// Reference: probnmn/modules/seq2seq_base.py:101-341 (in-repo logic) and AllenNLP 0.9.0
// SimpleSeq2Seq._prepare_output_projections / nn.util.masked_softmax / add_sentence_boundary_token_ids
// (restated in oracle/seq2seq_oracle.py).
#include "seq2seq.h"

namespace pnmn {

namespace {

constexpr int kPad = 0, kStart = 2, kEnd = 3;  // identical in every padded namespace (seq2seq_base.py:61-65)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ uint4 pack8h(const __half* h) {
  uint4 r;
  r.x = static_cast<uint32_t>(__half_as_ushort(h[0])) | (static_cast<uint32_t>(__half_as_ushort(h[1])) << 16);
  r.y = static_cast<uint32_t>(__half_as_ushort(h[2])) | (static_cast<uint32_t>(__half_as_ushort(h[3])) << 16);
  r.z = static_cast<uint32_t>(__half_as_ushort(h[4])) | (static_cast<uint32_t>(__half_as_ushort(h[5])) << 16);
  r.w = static_cast<uint32_t>(__half_as_ushort(h[6])) | (static_cast<uint32_t>(__half_as_ushort(h[7])) << 16);
  return r;
}
// 8 consecutive features [k0, k0+8) of row b -> operand buffer (hi; lo at + lo_off)
__device__ __forceinline__ void store_op8(__half* op, int64_t lo_off, int b, int k0, int K, const float* x, float mul) {
  __half hi[8], lo[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) split_f16(x[e] * mul, hi[e], lo[e]);
  const size_t off = op_off(b, k0, K);
  *reinterpret_cast<uint4*>(op + off) = pack8h(hi);
  *reinterpret_cast<uint4*>(op + lo_off + off) = pack8h(lo);
}

// Philox4x32-10 (counter-based; one draw per (row, step), reproducible and independent of launch geometry)
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
  const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
  const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}
__device__ __forceinline__ float philox_uniform(unsigned long long seed, uint32_t row, uint32_t step) {
  uint32_t c[4] = {row, step, 0x9E3779B9u, 0u};
  uint32_t k0 = static_cast<uint32_t>(seed), k1 = static_cast<uint32_t>(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return static_cast<float>(c[0] >> 8) * (1.0f / 16777216.0f);  // [0, 1)
}

}  // namespace

// =====================================================================================================
// tokens: AllenNLP add_sentence_boundary_token_ids, then the source drops its leading @start@
// (seq2seq_base.py:128-141).  src: [B][Tq+1] = w_1..w_n @end@ 0..;  tgt: [B][Tp+2] = @start@ p_1..p_m @end@ 0..
// =====================================================================================================
// out-of-vocabulary ids map to @@UNKNOWN@@ (the reference's embedding lookup would raise instead)
__device__ __forceinline__ int clamp_tok(int64_t v, int V) { return (v < 0 || v >= V) ? 1 : static_cast<int>(v); }

__global__ void prepare_tokens_kernel(const int64_t* __restrict__ source, const int64_t* __restrict__ target,
                                      const uint8_t* __restrict__ row_teacher, SeqDims d, int rows,
                                      unsigned long long seed, unsigned long long* __restrict__ seed_out,
                                      int* __restrict__ src, int* __restrict__ src_len, int* __restrict__ tgt,
                                      int* __restrict__ row_mode) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b == 0) seed_out[0] = seed;
  if (b >= d.B) return;
  const bool real = b < rows;
  row_mode[b] = (real && row_teacher) ? (row_teacher[b] != 0) : d.teacher;   // padding rows of the last tile: empty source / target (their loss is never read, their
                                // incoming gradient is zero, so they contribute nothing)
  int n = 0;
  for (int s = 0; real && s < d.Tq; ++s) n += source[static_cast<size_t>(b) * d.Tq + s] != kPad;
  for (int s = 0; s < d.Ts; ++s)
    src[static_cast<size_t>(b) * d.Ts + s] = (real && s < d.Tq) ? clamp_tok(source[static_cast<size_t>(b) * d.Tq + s], d.Vs) : kPad;
  src[static_cast<size_t>(b) * d.Ts + n] = kEnd;
  int len = 0;
  for (int s = 0; s < d.Ts; ++s) len += src[static_cast<size_t>(b) * d.Ts + s] != kPad;
  src_len[b] = len;
  if (target) {
    const int W = d.Tp + 2;
    int m = 0;
    for (int s = 0; real && s < d.Tp; ++s) m += target[static_cast<size_t>(b) * d.Tp + s] != kPad;
    tgt[static_cast<size_t>(b) * W] = kStart;
    for (int s = 0; s < d.Tp; ++s)
      tgt[static_cast<size_t>(b) * W + 1 + s] = real ? clamp_tok(target[static_cast<size_t>(b) * d.Tp + s], d.Vt) : kPad;
    tgt[static_cast<size_t>(b) * W + d.Tp + 1] = kPad;
    tgt[static_cast<size_t>(b) * W + m + 1] = kEnd;
  }
}
cudaError_t launch_prepare_tokens(const int64_t* source, const int64_t* target, const uint8_t* row_teacher, SeqDims d, int rows,
                                  unsigned long long seed, unsigned long long* seed_out, int* src, int* src_len, int* tgt,
                                  int* row_mode, cudaStream_t st) {
  prepare_tokens_kernel<<<(d.B + 127) / 128, 128, 0, st>>>(source, target, row_teacher, d, rows, seed, seed_out, src, src_len, tgt,
                                                          row_mode);
  return cudaGetLastError();
}

__global__ void stage_grad_loss_kernel(const float* __restrict__ g, int rows, int padded, float* __restrict__ staged) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < padded) staged[b] = b < rows ? g[b] : 0.f;
}
cudaError_t launch_stage_grad_loss(const float* grad_loss, int rows, int padded, float* staged, cudaStream_t st) {
  stage_grad_loss_kernel<<<(padded + 127) / 128, 128, 0, st>>>(grad_loss, rows, padded, staged);
  return cudaGetLastError();
}

__global__ void accumulate_kernel(float* __restrict__ dst, const float* __restrict__ src, int64_t n4) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 a = reinterpret_cast<float4*>(dst)[i];
    const float4 b = reinterpret_cast<const float4*>(src)[i];
    a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    reinterpret_cast<float4*>(dst)[i] = a;
  }
}
cudaError_t launch_accumulate(float* dst, const float* src, int64_t n, cudaStream_t st) {
  // n is a multiple of 64 (flat parameter layout), both buffers 16-byte aligned
  const int64_t n4 = n / 4;
  const int blocks = static_cast<int>(n4 / 256 + 1 < 148 * 4 ? n4 / 256 + 1 : 148 * 4);
  accumulate_kernel<<<blocks, 256, 0, st>>>(dst, src, n4);
  return cudaGetLastError();
}

// =====================================================================================================
// decoder, row-local part of one step (one CTA per batch row, 256 threads = one per hidden unit):
//   (a) t > 0: output projection of step t-1, softmax / log-softmax, greedy max or categorical sampling
//       (seq2seq_base.py:201-220), (b) t < S: choose the input token of step t (:188-198), dot-product
//       attention over the encoder outputs with AllenNLP's masked_softmax, attended vector -> operand copy.
// =====================================================================================================
// softmax / log-softmax over the Vt logits in `slg`, the chosen token (greedy max or categorical sampling,
// seq2seq_base.py:201-220) and its log-probability, by ONE warp; `satt` = Vt floats of scratch
__device__ __forceinline__ void pick_token(const DecRowArgs& a, const SeqDims& d, int b, int tp, const float* slg, float* satt,
                                           int* s_pred, int lane, bool write_logits) {
      float m = -INFINITY;
      for (int v = lane; v < d.Vt; v += 32) m = fmaxf(m, slg[v]);
      m = warp_max(m);
      float e[kSMaxV / 32], sum = 0.f;
#pragma unroll
      for (int i = 0; i < kSMaxV / 32; ++i) {
        const int v = lane + 32 * i;
        e[i] = v < d.Vt ? expf(slg[v] - m) : 0.f;
        sum += e[i];
      }
      sum = warp_sum(sum);
      int pred;
      if (!d.sampling) {
        // torch.max(class_probabilities, 1): largest probability, lowest index on ties (seq2seq_base.py:209)
        float bv = -1.f; int bi = 0x7fffffff;
#pragma unroll
        for (int i = 0; i < kSMaxV / 32; ++i) {
          const int v = lane + 32 * i;
          const float p = e[i] / sum;
          if (v < d.Vt && p > bv) { bv = p; bi = v; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
          const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
          if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        pred = bi;
      } else {
        // never sample @@PADDING@@ / @@UNKNOWN@@ / @start@ (indices 0..2), then torch.multinomial (:212-215)
        // probabilities go through shared memory (reuse satt) for the sequential inverse-CDF walk
#pragma unroll
        for (int i = 0; i < kSMaxV / 32; ++i) {
          const int v = lane + 32 * i;
          if (v < d.Vt) satt[v] = v <= kStart ? 0.f : e[i] / sum;
        }
        __syncwarp();
        // inverse-CDF walk over the vocabulary in index order, warp-parallel: lane l owns entries [4l, 4l+4)
        float p4[4], run = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int v = 4 * lane + i;
          p4[i] = v < d.Vt ? satt[v] : 0.f;
          run += p4[i];
        }
        float incl = run;   // inclusive scan of the lanes' sums
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const float up = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += up;
        }
        const float total = __shfl_sync(0xffffffffu, incl, 31);
        const float u = philox_uniform(a.seed[0], static_cast<uint32_t>(b), static_cast<uint32_t>(tp)) * total;
        float cum = incl - run;
        int hit = -1, last = -1;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (p4[i] > 0.f) {
            last = 4 * lane + i;
            cum += p4[i];
            if (hit < 0 && cum > u) hit = 4 * lane + i;
          }
        }
        const unsigned hits = __ballot_sync(0xffffffffu, hit >= 0);
        const unsigned lasts = __ballot_sync(0xffffffffu, last >= 0);
        if (hits) pred = __shfl_sync(0xffffffffu, hit, __ffs(hits) - 1);
        else if (lasts) pred = __shfl_sync(0xffffffffu, last, 31 - __clz(lasts));   // rounding: the last drawable token
        else pred = kEnd;
      }
      const float lse = m + logf(sum);
      if (write_logits)
        for (int v = lane; v < d.Vt; v += 32) a.logits[(static_cast<size_t>(tp) * d.B + b) * d.Vt + v] = slg[v];
      if (lane == 0) {
        a.lse[static_cast<size_t>(tp) * d.B + b] = lse;
        a.pred[static_cast<size_t>(tp) * d.B + b] = pred;
        a.logp[static_cast<size_t>(tp) * d.B + b] = slg[pred] - lse;
        *s_pred = pred;
      }
}

// output projection of decoding step tp from its hidden state `sh` (shared, 256 floats): logits, log-sum-exp, the chosen
// token and its log-probability.  Called by a whole CTA.
__device__ __forceinline__ void dec_output_step(const DecRowArgs& a, const SeqDims& d, int b, int tp, const float* sh,
                                                float* slg, float* satt, int* s_pred, int warp, int lane) {
    // four vocabulary entries per round, their eight 16-byte loads issued before the first reduction
    for (int v0 = warp; v0 < d.Vt; v0 += 32) {
      float4 w0[4], w1[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int v = v0 + 8 * k;
        if (v < d.Vt) {
          w0[k] = __ldg(reinterpret_cast<const float4*>(a.out_w + static_cast<size_t>(v) * kSH + lane * 8));
          w1[k] = __ldg(reinterpret_cast<const float4*>(a.out_w + static_cast<size_t>(v) * kSH + lane * 8 + 4));
        }
      }
      const float* h = sh + lane * 8;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int v = v0 + 8 * k;
        if (v < d.Vt) {
          float acc = w0[k].x * h[0];
          acc = fmaf(w0[k].y, h[1], acc); acc = fmaf(w0[k].z, h[2], acc); acc = fmaf(w0[k].w, h[3], acc);
          acc = fmaf(w1[k].x, h[4], acc); acc = fmaf(w1[k].y, h[5], acc); acc = fmaf(w1[k].z, h[6], acc); acc = fmaf(w1[k].w, h[7], acc);
          acc = warp_sum(acc);
          if (lane == 0) slg[v] = acc + __ldg(a.out_b + v);
        }
      }
    }
    __syncthreads();
    if (warp == 0) pick_token(a, d, b, tp, slg, satt, s_pred, lane, true);
    __syncthreads();
}

__global__ void __launch_bounds__(256) dec_row_kernel(const DecRowArgs a) {
  extern __shared__ __align__(16) float s_enc[];   // [Ts][256]: the row's encoder outputs, fetched once for scores AND context
  __shared__ float sh[kSH], satt[kSH], slg[kSMaxV], ssc[kSMaxT], sp[kSMaxT];
  __shared__ int s_pred;
  const SeqDims& d = a.d;
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int t = a.t;
  pdl_launch_dependents();
  pdl_wait();
  sh[tid] = a.h_dec[(static_cast<size_t>(t) * d.Bp + b) * kSH + tid];
  __syncthreads();

  // the output of step t-1 feeds the recurrence only through a free-running row's next input token: a call whose rows are
  // all teacher-forced computes every step's output after the loop, in one launch over (step, row) (dec_out_kernel)
  if (t > 0 && !a.defer_out) dec_output_step(a, d, b, t - 1, sh, slg, satt, &s_pred, warp, lane);
  if (t >= d.S) return;

  // ---- input token of step t: gold token under teacher forcing, else the previous prediction ----------
  if (tid == 0) {
    const int tok = a.row_mode[b] ? a.tgt[static_cast<size_t>(b) * (d.Tp + 2) + t] : (t == 0 ? kStart : s_pred);
    a.inp[static_cast<size_t>(t) * d.B + b] = tok;
  }
  // ---- dot-product attention ---------------------------------------------------------------------------
  const float* enc = a.enc + static_cast<size_t>(b) * d.Ts * kSH;
  const int len = a.src_len[b];
  // warp w takes positions w, w+8, ...: all of its (<= 8) row loads are issued before the first dot product, and the rows
  // are kept in shared memory for the context vector below (they used to be re-read from L2 by a loop of dependent loads:
  // ~3 us of this ~13 us kernel)
  {
    float4 e0[kSMaxT / 8], e1[kSMaxT / 8];
#pragma unroll
    for (int k = 0; k < kSMaxT / 8; ++k) {
      const int s = warp + 8 * k;
      if (s < d.Ts) {
        e0[k] = __ldg(reinterpret_cast<const float4*>(enc + static_cast<size_t>(s) * kSH + lane * 8));
        e1[k] = __ldg(reinterpret_cast<const float4*>(enc + static_cast<size_t>(s) * kSH + lane * 8 + 4));
      }
    }
    const float* h = sh + lane * 8;
#pragma unroll
    for (int k = 0; k < kSMaxT / 8; ++k) {
      const int s = warp + 8 * k;
      if (s < d.Ts) {
        *reinterpret_cast<float4*>(s_enc + s * kSH + lane * 8) = e0[k];
        *reinterpret_cast<float4*>(s_enc + s * kSH + lane * 8 + 4) = e1[k];
        float acc = e0[k].x * h[0];
        acc = fmaf(e0[k].y, h[1], acc); acc = fmaf(e0[k].z, h[2], acc); acc = fmaf(e0[k].w, h[3], acc);
        acc = fmaf(e1[k].x, h[4], acc); acc = fmaf(e1[k].y, h[5], acc); acc = fmaf(e1[k].z, h[6], acc); acc = fmaf(e1[k].w, h[7], acc);
        acc = warp_sum(acc);
        if (lane == 0) ssc[s] = acc;
      }
    }
  }
  __syncthreads();
  if (warp == 0) {
    // masked_softmax (allennlp 0.9.0): p = softmax(scores * mask) * mask; p /= (sum(p) + 1e-13)
    float z[kSMaxT / 32], m = -INFINITY;
#pragma unroll
    for (int i = 0; i < kSMaxT / 32; ++i) {
      const int s = lane + 32 * i;
      z[i] = s < d.Ts ? (s < len ? ssc[s] : 0.f) : -INFINITY;
      m = fmaxf(m, z[i]);
    }
    m = warp_max(m);
    float u[kSMaxT / 32], Z = 0.f;
#pragma unroll
    for (int i = 0; i < kSMaxT / 32; ++i) {
      const int s = lane + 32 * i;
      u[i] = s < d.Ts ? expf(z[i] - m) : 0.f;
      Z += u[i];
    }
    Z = warp_sum(Z);
    float r[kSMaxT / 32], R = 0.f;
#pragma unroll
    for (int i = 0; i < kSMaxT / 32; ++i) {
      const int s = lane + 32 * i;
      r[i] = s < len ? u[i] / Z : 0.f;
      R += r[i];
    }
    R = warp_sum(R) + 1e-13f;
#pragma unroll
    for (int i = 0; i < kSMaxT / 32; ++i) {
      const int s = lane + 32 * i;
      if (s < d.Ts) {
        const float p = r[i] / R;
        sp[s] = p;
        a.attn_p[(static_cast<size_t>(t) * d.B + b) * d.Ts + s] = p;
      }
    }
  }
  __syncthreads();
  {
    float acc = 0.f;
    for (int s = 0; s < len; ++s) acc = fmaf(sp[s], s_enc[s * kSH + tid], acc);
    satt[tid] = acc;
  }
  __syncthreads();
  if (tid < kSH / 8) store_op8(a.att_op + static_cast<size_t>(t) * a.att_step, a.att_lo, b, tid * 8, kSH, satt + tid * 8, 1.f);
}
cudaError_t launch_dec_row(const DecRowArgs& a, cudaStream_t st) {
  const size_t smem = static_cast<size_t>(a.d.Ts) * kSH * sizeof(float);
  static size_t attr_smem = 0;
  if (smem > attr_smem) {
    cudaError_t e = cudaFuncSetAttribute(dec_row_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kSMaxT * kSH * sizeof(float)));
    if (e != cudaSuccess) return e;
    attr_smem = kSMaxT * kSH * sizeof(float);
  }
  return launch_pdl(dec_row_kernel, dim3(a.d.B), dim3(256), smem, st, seq_use_pdl(), a);
}

// every step's output of a teacher-forced pass at once: grid = (rows, steps)
__global__ void __launch_bounds__(256) dec_out_kernel(const DecRowArgs a) {
  __shared__ float sh[kSH], satt[kSH], slg[kSMaxV];
  __shared__ int s_pred;
  const SeqDims& d = a.d;
  const int b = blockIdx.x, tp = blockIdx.y, tid = threadIdx.x;
  sh[tid] = a.h_dec[(static_cast<size_t>(tp + 1) * d.Bp + b) * kSH + tid];
  __syncthreads();
  dec_output_step(a, d, b, tp, sh, slg, satt, &s_pred, tid >> 5, tid & 31);
}
cudaError_t launch_dec_out(const DecRowArgs& a, cudaStream_t st) {
  dec_out_kernel<<<dim3(a.d.B, a.d.S), 256, 0, st>>>(a);
  return cudaGetLastError();
}

// the same with the logits of every step already computed (one tensor-core GEMM over all (step, row) pairs,
// pnmn_gemm_split in pnmn_pg_forward): one WARP per (row, step) does the softmax, the prediction and its log-probability
__global__ void __launch_bounds__(256) dec_out_post_kernel(const DecRowArgs a) {
  __shared__ float slg[8][kSMaxV], satt[8][kSMaxV];
  __shared__ int s_pred[8];
  const SeqDims& d = a.d;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int item = blockIdx.x * 8 + warp;
  if (item >= d.B * d.S) return;
  const int tp = item / d.B, b = item - tp * d.B;
  for (int v = lane; v < d.Vt; v += 32) slg[warp][v] = a.logits[(static_cast<size_t>(tp) * d.B + b) * d.Vt + v];
  __syncwarp();
  pick_token(a, d, b, tp, slg[warp], satt[warp], &s_pred[warp], lane, false);
}
cudaError_t launch_dec_out_post(const DecRowArgs& a, cudaStream_t st) {
  dec_out_post_kernel<<<(a.d.B * a.d.S + 7) / 8, 256, 0, st>>>(a);
  return cudaGetLastError();
}

// =====================================================================================================
// trimming (seq2seq_base.py:278-293), per-row losses (:235-254, :333-341) and the dlogits coefficients
// =====================================================================================================
// one WARP per row (lanes over the decoding steps; a thread per row walked its <= 46 steps serially: 46-70 us per call)
__global__ void __launch_bounds__(256) finalize_kernel(const FinalizeArgs a) {
  const SeqDims& d = a.d;
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (b >= d.B) return;
  const bool real = b < a.rows;   // (padding rows still get their coef / label entries: the backward pass reads them)
  const bool tf = a.row_mode[b] != 0;
  const int Se = tf ? d.S : d.free_S;   // a free-running row of a mixed call stops after free_S steps
  int first_end = 0x7fffffff;
  for (int t = lane; t < Se; t += 32)
    if (a.pred[static_cast<size_t>(t) * d.B + b] == kEnd) { first_end = t; break; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) first_end = min(first_end, __shfl_xor_sync(0xffffffffu, first_end, o));
  if (first_end == 0x7fffffff) first_end = -1;
  float lp_sum = 0.f, cnt = 0.f;
  for (int t = lane; t < d.S; t += 32) {
    const int raw = t < Se ? a.pred[static_cast<size_t>(t) * d.B + b] : kPad;
    int keep;
    if (t >= Se) keep = kPad;
    else if (first_end < 0) keep = raw;                // no @end@: row unchanged
    else if (first_end == 0) keep = kPad;              // @end@ first: the whole row becomes padding
    else keep = t <= first_end ? raw : kPad;
    if (real) {
      a.raw_out[static_cast<size_t>(b) * d.S + t] = raw;
      a.pred_out[static_cast<size_t>(b) * d.S + t] = keep;
    }
    const float pm = keep != kPad ? 1.f : 0.f;
    lp_sum += a.logp[static_cast<size_t>(t) * d.B + b] * pm;
    cnt += pm;
  }
  lp_sum = warp_sum(lp_sum);
  cnt = warp_sum(cnt);
  if (!tf) {
    if (real && lane == 0) a.loss[b] = -(lp_sum / (cnt + 1e-12f));
    for (int t = lane; t < d.S; t += 32) {
      const int raw = a.pred[static_cast<size_t>(t) * d.B + b];
      const bool kept = t < Se && (first_end < 0 ? raw != kPad : (first_end > 0 && t <= first_end && raw != kPad));
      a.coef[static_cast<size_t>(t) * d.B + b] = kept ? 1.f / (cnt + 1e-12f) : 0.f;
      a.label[static_cast<size_t>(t) * d.B + b] = raw;
    }
  } else {
    // sequence_cross_entropy_with_logits(average=None): sum(nll * mask) / (sum(mask) + 1e-13) per row
    const int W = d.Tp + 2;
    float n = 0.f, tot = 0.f;
    for (int t = lane; t < d.S; t += 32) {
      const int lab = a.tgt[static_cast<size_t>(b) * W + t + 1];
      const float m = lab != kPad ? 1.f : 0.f;
      const float nll = a.lse[static_cast<size_t>(t) * d.B + b] - a.logits[(static_cast<size_t>(t) * d.B + b) * d.Vt + lab];
      tot += nll * m;
      n += m;
    }
    n = warp_sum(n);
    tot = warp_sum(tot);
    for (int t = lane; t < d.S; t += 32) {
      const int lab = a.tgt[static_cast<size_t>(b) * W + t + 1];
      a.coef[static_cast<size_t>(t) * d.B + b] = (lab != kPad ? 1.f : 0.f) / (n + 1e-13f);
      a.label[static_cast<size_t>(t) * d.B + b] = lab;
    }
    if (real && lane == 0) a.loss[b] = tot / (n + 1e-13f);
  }
  if (a.logits_out && real)
    for (int i = lane; i < d.S * d.Vt; i += 32) {
      const int t = i / d.Vt, v = i - t * d.Vt;
      a.logits_out[(static_cast<size_t>(b) * d.S + t) * d.Vt + v] = a.logits[(static_cast<size_t>(t) * d.B + b) * d.Vt + v];
    }
}
cudaError_t launch_finalize(const FinalizeArgs& a, cudaStream_t st) {
  finalize_kernel<<<(a.d.B * 32 + 255) / 256, 256, 0, st>>>(a);
  return cudaGetLastError();
}

// =====================================================================================================
// decoder backward, row-local part (one CTA per row, thread = hidden unit)
// =====================================================================================================
__device__ __forceinline__ void cell_bwd(float dh, float dc_in, float i_, float f_, float g_, float o_, float c_cur,
                                         float c_prev, float (&da)[4], float& dc_prev) {
  const float tc = tanhf(c_cur);
  const float do_ = dh * tc;
  const float dc = dc_in + dh * o_ * (1.f - tc * tc);
  da[0] = dc * g_ * i_ * (1.f - i_);
  da[1] = dc * c_prev * f_ * (1.f - f_);
  da[2] = dc * i_ * (1.f - g_ * g_);
  da[3] = do_ * o_ * (1.f - o_);
  dc_prev = dc * f_;
}

__global__ void __launch_bounds__(256) dec_bwd_row_kernel(const DecBwdRowArgs a) {
  extern __shared__ __align__(16) float s_enc[];   // [Ts][256] (as in dec_row_kernel)
  __shared__ float s_datt[kSH], s_dp[kSMaxT], s_ds[kSMaxT], s_p[kSMaxT];
  __shared__ __align__(16) float s_dg[kSG];
  const SeqDims& d = a.d;
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int t = a.t;
  pdl_launch_dependents();
  pdl_wait();
  if (b >= d.B) {
    // padding row of the 128-row tile: its gate-gradient operand and dlogits row are contracted over by the weight-gradient
    // GEMMs, so they must be zero whatever an earlier, larger batch left there
    if (t >= 0) {
      if (tid < kSG / 8) {
        const uint4 z = make_uint4(0u, 0u, 0u, 0u);
        __half* op = a.dg_op + static_cast<size_t>(t) * a.dg_step;
        const size_t off = op_off(b, tid * 8, kSG);
        *reinterpret_cast<uint4*>(op + off) = z;
        *reinterpret_cast<uint4*>(op + a.dg_lo + off) = z;
      }
    }
    return;
  }
  float dh = a.dh[static_cast<size_t>(b) * kSH + tid];
  // dh / datt are accumulation targets of the next data-gradient GEMM (split-K partial sums): leave them zeroed
  if (t >= 0) a.dh[static_cast<size_t>(b) * kSH + tid] = 0.f;

  if (a.do_attn) {
    // attention of step ta = t+1 used h_t (slot ta): scores_s = enc_s . h, p = masked_softmax, att = sum p_s enc_s
    const int ta = t + 1;
    const float* enc = a.enc + static_cast<size_t>(b) * d.Ts * kSH;
    float* denc = a.denc + static_cast<size_t>(b) * d.Ts * kSH;
    const int len = a.src_len[b];
    const float datt = a.datt[static_cast<size_t>(b) * kSH + tid];
    const_cast<float*>(a.datt)[static_cast<size_t>(b) * kSH + tid] = 0.f;
    const float hq = a.h_dec[(static_cast<size_t>(ta) * d.Bp + b) * kSH + tid];
    s_datt[tid] = datt;
    if (tid < d.Ts) s_p[tid] = a.attn_p[(static_cast<size_t>(ta) * d.B + b) * d.Ts + tid];
    __syncthreads();
    {
      float4 e0[kSMaxT / 8], e1[kSMaxT / 8];
#pragma unroll
      for (int k = 0; k < kSMaxT / 8; ++k) {
        const int s = warp + 8 * k;
        if (s < len) {
          e0[k] = __ldg(reinterpret_cast<const float4*>(enc + static_cast<size_t>(s) * kSH + lane * 8));
          e1[k] = __ldg(reinterpret_cast<const float4*>(enc + static_cast<size_t>(s) * kSH + lane * 8 + 4));
        }
      }
      const float* g = s_datt + lane * 8;
#pragma unroll
      for (int k = 0; k < kSMaxT / 8; ++k) {
        const int s = warp + 8 * k;
        if (s < len) {
          *reinterpret_cast<float4*>(s_enc + s * kSH + lane * 8) = e0[k];
          *reinterpret_cast<float4*>(s_enc + s * kSH + lane * 8 + 4) = e1[k];
          float acc = e0[k].x * g[0];
          acc = fmaf(e0[k].y, g[1], acc); acc = fmaf(e0[k].z, g[2], acc); acc = fmaf(e0[k].w, g[3], acc);
          acc = fmaf(e1[k].x, g[4], acc); acc = fmaf(e1[k].y, g[5], acc); acc = fmaf(e1[k].z, g[6], acc); acc = fmaf(e1[k].w, g[7], acc);
          acc = warp_sum(acc);
          if (lane == 0) s_dp[s] = acc;
        }
      }
    }
    __syncthreads();
    if (warp == 0) {
      float G = 0.f;
      for (int s = lane; s < len; s += 32) G += s_p[s] * s_dp[s];
      G = warp_sum(G);
      for (int s = lane; s < d.Ts; s += 32) s_ds[s] = s < len ? s_p[s] * (s_dp[s] - G) : 0.f;
    }
    __syncthreads();
    // eight positions at a time, every load before the first store: `denc` may alias `enc` as far as the compiler knows,
    // and with one read-modify-write per iteration the loop was a chain of up to Ts dependent L2 round trips
    float acc = 0.f;
    for (int s0 = 0; s0 < len; s0 += 8) {
      float dv[8];
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (s0 + k < len) dv[k] = denc[static_cast<size_t>(s0 + k) * kSH + tid];
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (s0 + k < len) {
          acc = fmaf(s_ds[s0 + k], s_enc[(s0 + k) * kSH + tid], acc);
          denc[static_cast<size_t>(s0 + k) * kSH + tid] = dv[k] + s_p[s0 + k] * datt + s_ds[s0 + k] * hq;
        }
    }
    dh += acc;
  }
  if (t < 0) {
    a.dh[static_cast<size_t>(b) * kSH + tid] = dh;
    return;
  }

  // ---- output projection + (log-)softmax backward of step t: precomputed for every step (dec_bwd_proj_kernel) ----------
  dh += a.dhp[(static_cast<size_t>(t) * d.Bp + b) * kSH + tid];
  // ---- LSTM cell backward ---------------------------------------------------------------------------------
  const float* gt = a.gates + (static_cast<size_t>(t) * d.Bp + b) * kSG;
  float da[4], dc_prev;
  cell_bwd(dh, a.dc[static_cast<size_t>(b) * kSH + tid], gt[tid], gt[kSH + tid], gt[2 * kSH + tid], gt[3 * kSH + tid],
           a.c_dec[(static_cast<size_t>(t + 1) * d.Bp + b) * kSH + tid], a.c_dec[(static_cast<size_t>(t) * d.Bp + b) * kSH + tid], da,
           dc_prev);
  a.dc[static_cast<size_t>(b) * kSH + tid] = dc_prev;
  s_dg[tid] = da[0]; s_dg[kSH + tid] = da[1]; s_dg[2 * kSH + tid] = da[2]; s_dg[3 * kSH + tid] = da[3];
  __syncthreads();
  if (tid < kSG / 8) store_op8(a.dg_op + static_cast<size_t>(t) * a.dg_step, a.dg_lo, b, tid * 8, kSG, s_dg + tid * 8, a.scale[0]);
}
// d(loss)/d(logits) of EVERY step and its image under the output projection, dhp[t][b][j] = sum_v W_o[v][j] * dlogits[t][b][v]:
// neither depends on the backward recurrence (the coefficients, logits and labels are the forward pass's), so they are
// computed in one launch over (row, step) instead of inside each of the S dependent row kernels (~2 us of each step)
__global__ void __launch_bounds__(256) dec_bwd_proj_kernel(const DecBwdRowArgs a) {
  __shared__ float s_dl[kSMaxV];
  const SeqDims& d = a.d;
  const int b = blockIdx.x, t = blockIdx.y, tid = threadIdx.x;
  const float cf = b < d.B ? a.grad_loss[b] * a.coef[static_cast<size_t>(t) * d.B + b] : 0.f;
  if (tid < d.Vt) {
    float dl = 0.f;
    if (cf != 0.f) {
      const float lg = a.logits[(static_cast<size_t>(t) * d.B + b) * d.Vt + tid];
      const float p = expf(lg - a.lse[static_cast<size_t>(t) * d.B + b]);
      dl = cf * (p - (tid == a.label[static_cast<size_t>(t) * d.B + b] ? 1.f : 0.f));
    }
    s_dl[tid] = dl;
    a.dlogits[(static_cast<size_t>(t) * d.Bp + b) * d.Vt + tid] = dl;
  }
  __syncthreads();
  float acc = 0.f;
  if (cf != 0.f)
    for (int v = 0; v < d.Vt; ++v) acc = fmaf(__ldg(a.out_w + static_cast<size_t>(v) * kSH + tid), s_dl[v], acc);
  a.dhp[(static_cast<size_t>(t) * d.Bp + b) * kSH + tid] = acc;
}
cudaError_t launch_dec_bwd_proj(const DecBwdRowArgs& a, cudaStream_t st) {
  dec_bwd_proj_kernel<<<dim3(a.d.Bp, a.d.S), 256, 0, st>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_dec_bwd_row(const DecBwdRowArgs& a, cudaStream_t st) {
  const size_t smem = static_cast<size_t>(a.d.Ts) * kSH * sizeof(float);
  static size_t attr_smem = 0;
  if (smem > attr_smem) {
    cudaError_t e = cudaFuncSetAttribute(dec_bwd_row_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kSMaxT * kSH * sizeof(float)));
    if (e != cudaSuccess) return e;
    attr_smem = kSMaxT * kSH * sizeof(float);
  }
  return launch_pdl(dec_bwd_row_kernel, dim3(a.d.Bp), dim3(256), smem, st, seq_use_pdl(), a);
}

// encoder: LSTM cell backward of one (layer, step); rows beyond their length carry dh / dc through unchanged
__global__ void __launch_bounds__(256) enc_cell_bwd_kernel(const EncCellBwdPair pr) {
  __shared__ __align__(16) float s_dg[kSG];
  const EncCellBwdArgs& a = pr.a[blockIdx.y];
  const int b = blockIdx.x, tid = threadIdx.x;
  pdl_launch_dependents();
  pdl_wait();
  const bool valid = b < a.B && a.t < a.src_len[b];   // (b >= B: padding row of the tile, zero gradient operand)
  float da[4] = {0.f, 0.f, 0.f, 0.f};
  if (valid) {
    float dh = a.dh[static_cast<size_t>(b) * kSH + tid];
    a.dh[static_cast<size_t>(b) * kSH + tid] = 0.f;   // the data-gradient GEMM of this step accumulates the new value
    if (a.dext) dh += a.dext[static_cast<size_t>(b) * a.dext_stride + tid];
    const float* gt = a.gates + static_cast<size_t>(b) * kSG;
    float dc_prev;
    cell_bwd(dh, a.dc[static_cast<size_t>(b) * kSH + tid], gt[tid], gt[kSH + tid], gt[2 * kSH + tid], gt[3 * kSH + tid],
             a.c_cur[static_cast<size_t>(b) * kSH + tid], a.c_prev[static_cast<size_t>(b) * kSH + tid], da, dc_prev);
    a.dc[static_cast<size_t>(b) * kSH + tid] = dc_prev;
  }
  s_dg[tid] = da[0]; s_dg[kSH + tid] = da[1]; s_dg[2 * kSH + tid] = da[2]; s_dg[3 * kSH + tid] = da[3];
  __syncthreads();
  if (tid < kSG / 8) store_op8(a.dg_op, a.dg_lo, b, tid * 8, kSG, s_dg + tid * 8, a.scale[0]);
}
cudaError_t launch_enc_cell_bwd_pair(const EncCellBwdPair& p, cudaStream_t st) {
  return launch_pdl(enc_cell_bwd_kernel, dim3(p.a[0].Bp, p.count), dim3(256), 0, st, seq_use_pdl(), p);
}
cudaError_t launch_enc_cell_bwd(const EncCellBwdArgs& a, cudaStream_t st) {
  EncCellBwdPair p;
  p.a[0] = a; p.a[1] = a; p.count = 1;
  return launch_enc_cell_bwd_pair(p, st);
}

// =====================================================================================================
// table gradient: dP[v][g] += scale[1] * sum over (t, b) with token v of dG[t][b][g]
// CTA = 32 gate columns x a slice of the time range.  128 rows x 32 columns of gradients are staged in shared memory with
// coalesced 16-byte loads (many in flight); the scatter then runs out of shared memory with lane = column: rows are OWNED
// by warps through their token (token % 8 == warp), so a warp is the only writer of its tokens' partial sums -- no atomics,
// no bank conflicts.  Without a token table (the layer-1 bias: every row goes to entry 0) the warps sum rows in registers.
// (The first version let all 256 threads atomicAdd into shared memory, thread = (row, 8 columns): rows with the same
// token -- padding, frequent words -- hit the same addresses, 32-way conflicts; 70-290 us per launch in profiles/r2.)
// =====================================================================================================
constexpr int kTGRows = 128;                 // rows staged per pass
constexpr int kTGStride = 33;                // padded row stride of the staged tile (conflict-free column access)
__global__ void __launch_bounds__(256) table_grad_kernel(const TableGradArgs a) {
  __shared__ float acc[kSMaxV * 32];
  __shared__ float tile[kTGRows * kTGStride];
  __shared__ int stok[kTGRows];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < a.V * 32; i += 256) acc[i] = 0.f;
  const int g0 = blockIdx.x * 32;
  const int t0 = static_cast<int>((static_cast<long long>(a.T) * blockIdx.y) / gridDim.y);
  const int t1 = static_cast<int>((static_cast<long long>(a.T) * (blockIdx.y + 1)) / gridDim.y);
  float r = 0.f;   // (no token table: column sums in registers)
  for (int t = t0; t < t1; ++t) {
    const __half* dg = a.dg + static_cast<size_t>(t) * a.dg_step;
    for (int b0 = 0; b0 < a.B; b0 += kTGRows) {
      __syncthreads();   // the previous tile has been consumed (and acc is zeroed on the first pass)
      // stage 128 rows x 32 columns: thread = (row, 8-column group), 16-byte loads that are contiguous across a warp's rows
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int i = tid + 256 * k, row = i & (kTGRows - 1), grp = i >> 7;   // grp in 0..3
        const int b = b0 + row;
        float x[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (b < a.B) {
          const size_t off = op_off(b, g0 + grp * 8, kSG);
          const uint4 hi = *reinterpret_cast<const uint4*>(dg + off);
          const uint4 lo = *reinterpret_cast<const uint4*>(dg + a.dg_lo + off);
          const __half2* h2 = reinterpret_cast<const __half2*>(&hi);
          const __half2* l2 = reinterpret_cast<const __half2*>(&lo);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 hf = __half22float2(h2[e]), lf = __half22float2(l2[e]);
            x[2 * e] = hf.x + lf.x; x[2 * e + 1] = hf.y + lf.y;
          }
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) tile[row * kTGStride + grp * 8 + e] = x[e];
      }
      if (tid < kTGRows) {
        const int b = b0 + tid;
        stok[tid] = (a.tok && b < a.B) ? a.tok[static_cast<size_t>(t) * a.tok_step + static_cast<size_t>(b) * a.tok_stride] : 0;
      }
      __syncthreads();
      if (!a.tok) {
        for (int row = warp; row < kTGRows; row += 8) r += tile[row * kTGStride + lane];
      } else {
        // scatter by token; warp w is the only writer of the tokens with v % 8 == w.  Padding tokens are skipped: a padded
        // position carries no gradient (masked encoder steps; decoder steps behind the end of the target have weight 0)
#pragma unroll
        for (int q = 0; q < kTGRows / 32; ++q) {
          const int v = stok[q * 32 + lane];
          unsigned mine = __ballot_sync(0xffffffffu, v > 0 && (v & 7) == warp);
          while (mine) {
            const int s0 = __ffs(mine) - 1;
            mine &= mine - 1;
            const int v0 = __shfl_sync(0xffffffffu, v, s0);
            acc[v0 * 32 + lane] += tile[(q * 32 + s0) * kTGStride + lane];
          }
        }
      }
    }
  }
  if (!a.tok) atomicAdd(&acc[lane], r);
  __syncthreads();
  const float unscale = a.scale[1];
  for (int i = tid; i < a.V * 32; i += 256) {
    const float x = acc[i];
    if (x != 0.f) atomicAdd(a.dP + static_cast<size_t>(i >> 5) * kSG + g0 + (i & 31), x * unscale);
  }
}
cudaError_t launch_table_grad(const TableGradArgs& a, cudaStream_t st) {
  if (a.T <= 0) return cudaSuccess;
  const int split = a.T < 8 ? a.T : 8;
  table_grad_kernel<<<dim3(kSG / 32, split), 256, 0, st>>>(a);
  return cudaGetLastError();
}

__global__ void bias_from_table_kernel(const float* __restrict__ dP, int V, float* __restrict__ db0, float* __restrict__ db1) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= kSG) return;
  float s = 0.f;
  for (int v = 0; v < V; ++v) s += dP[static_cast<size_t>(v) * kSG + g];
  db0[g] += s;
  db1[g] += s;
}
cudaError_t launch_bias_from_table(const float* dP, int V, float* db0, float* db1, cudaStream_t st) {
  bias_from_table_kernel<<<kSG / 256, 256, 0, st>>>(dP, V, db0, db1);
  return cudaGetLastError();
}

__global__ void seq_loss_scale_kernel(const float* __restrict__ g, int B, float* __restrict__ scale) {
  __shared__ float red[32];
  float m = 0.f;
  for (int i = threadIdx.x; i < B; i += blockDim.x) m = fmaxf(m, fabsf(g[i]));
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < static_cast<int>(blockDim.x >> 5); ++i) m = fmaxf(m, red[i]);
    float s = 1.f;
    if (m > 0.f && isfinite(m)) {
      int e;
      frexpf(m, &e);            // m = f * 2^e, f in [0.5, 1)
      int k = 10 - e;           // m * 2^k in [2^9, 2^10)
      k = k > 60 ? 60 : (k < -60 ? -60 : k);
      s = ldexpf(1.f, k);
    }
    scale[0] = s;
    scale[1] = 1.f / s;
    scale[2] = 1.f;
  }
}
cudaError_t launch_seq_loss_scale(const float* grad_loss, int B, float* scale, cudaStream_t st) {
  seq_loss_scale_kernel<<<1, 256, 0, st>>>(grad_loss, B, scale);
  return cudaGetLastError();
}

// =====================================================================================================
// weight packing: fp32 parameters -> split fp16 tiles [n/64][k/8][64][8] (hi, then lo)
// =====================================================================================================
__global__ void __launch_bounds__(256) pack_seq_kernel(const PackJobs jobs, const float* __restrict__ params,
                                                       __half* __restrict__ packed) {
  const PackJob& j = jobs.j[blockIdx.y];
  const int groups = j.N * (j.K >> 3);
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= groups) return;
  // idx enumerates (n tile, k group, row) in storage order so that the 16-byte stores coalesce
  const int r = idx & 63, kg = (idx >> 6) % (j.K >> 3), nt = (idx >> 6) / (j.K >> 3);
  const int n = nt * 64 + r;
  __half hi[8], lo[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int k = kg * 8 + e;
    float v;
    if (j.mode == 0) {
      const int gr = (r >> 4) * kSH + nt * 16 + (r & 15);
      v = k < j.split ? params[j.src0 + static_cast<int64_t>(gr) * j.ld0 + k]
                      : params[j.src1 + static_cast<int64_t>(gr) * j.ld1 + (k - j.split)];
    } else {
      v = n < j.split ? params[j.src0 + static_cast<int64_t>(k) * j.ld0 + n]
                      : params[j.src1 + static_cast<int64_t>(k) * j.ld1 + (n - j.split)];
    }
    split_f16(v, hi[e], lo[e]);
  }
  const size_t off = j.dst + wp_off(n, kg * 8, j.K);
  *reinterpret_cast<uint4*>(packed + off) = pack8h(hi);
  *reinterpret_cast<uint4*>(packed + off + static_cast<size_t>(j.N) * j.K) = pack8h(lo);
}
cudaError_t launch_pack_seq(const PackJobs& jobs, int n_jobs, const float* params, __half* packed, cudaStream_t st) {
  // the largest job is 1024 x 512: 65536 groups of 8
  pack_seq_kernel<<<dim3(256, n_jobs), 256, 0, st>>>(jobs, params, packed);
  return cudaGetLastError();
}

// =====================================================================================================
// small strided fp32 GEMM (tables, embedding / projection gradients): 32x32 tile per CTA, split-K over
// blockIdx.z with atomics when gridDim.z > 1 (C must then be an accumulation target)
// =====================================================================================================
__global__ void __launch_bounds__(256) simt_gemm_kernel(const SimtGemm g) {
  __shared__ float sa[32][33], sb[32][33];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
  const int kz = gridDim.z;
  const int k_lo = static_cast<int>((static_cast<long long>(g.K) * blockIdx.z) / kz);
  const int k_hi = static_cast<int>((static_cast<long long>(g.K) * (blockIdx.z + 1)) / kz);
  float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
  for (int k0 = k_lo; k0 < k_hi; k0 += 32) {
    for (int i = threadIdx.x; i < 1024; i += 256) {
      const int r = i >> 5, c = i & 31;  // r: m (or n) inside the tile, c: k
      const int k = k0 + c;
      sa[r][c] = (m0 + r < g.M && k < k_hi) ? g.A[static_cast<int64_t>(m0 + r) * g.sam + static_cast<int64_t>(k) * g.sak] : 0.f;
      sb[r][c] = (n0 + r < g.N && k < k_hi) ? g.B[static_cast<int64_t>(k) * g.sbk + static_cast<int64_t>(n0 + r) * g.sbn] : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int c = 0; c < 32; ++c) {
      const float a0 = sa[ty][c], a1 = sa[ty + 16][c], b0 = sb[tx][c], b1 = sb[tx + 16][c];
      acc[0][0] = fmaf(a0, b0, acc[0][0]); acc[0][1] = fmaf(a0, b1, acc[0][1]);
      acc[1][0] = fmaf(a1, b0, acc[1][0]); acc[1][1] = fmaf(a1, b1, acc[1][1]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int jn = 0; jn < 2; ++jn) {
      const int m = m0 + ty + 16 * i, n = n0 + tx + 16 * jn;
      if (m >= g.M || n >= g.N) continue;
      float v = acc[i][jn] * g.alpha;
      if (blockIdx.z == 0) {
        if (g.bias0) v += g.bias0[n];
        if (g.bias1) v += g.bias1[n];
      }
      float* c = g.C + static_cast<int64_t>(m) * g.ldc + n;
      if (kz > 1) atomicAdd(c, v);
      else if (g.accumulate) *c += v;
      else *c = v;
    }
}
cudaError_t launch_simt_gemm(const SimtGemm& g, cudaStream_t st) {
  // few output tiles and a long contraction: split K over blockIdx.z (atomics), needs an accumulation target
  int kz = 1;
  if (g.accumulate && g.K >= 512) kz = g.K >= 2048 ? 16 : g.K / 128;
  simt_gemm_kernel<<<dim3((g.N + 31) / 32, (g.M + 31) / 32, kz), 256, 0, st>>>(g);
  return cudaGetLastError();
}

}  // namespace pnmn
