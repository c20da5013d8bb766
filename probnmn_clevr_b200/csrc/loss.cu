// Answer head of the module network after the classifier (probnmn/models/nmn.py:245-269): log-softmax over the answer
// logits, the predicted answer (torch.max: first maximum), its replacement by @@UNKNOWN@@ and the constant loss 3.33 for
// rows whose program could not be executed, the per-row cross entropy (or the negated log-probability of the prediction
// when no answers are given), the number of correct predictions for the accuracy metric -- and the gradient of the per-row
// losses with respect to the logits.  One warp per row, one launch each way (the eager version was ~10 small launches:
// log_softmax, max, two masked_fill, cross_entropy, ==, sum and their backward nodes).  Latency-bound, a few KB of data.
#include <cuda_runtime.h>
#include <cmath>
#include <cstdint>
#include <string>

#include "../../include/pnmn.h"

namespace pnmn { void set_last_error(const std::string& s); void count_launches(int n); }

namespace {

int fail(const std::string& s) {
  pnmn::set_last_error(s);
  return 1;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// xin[b] < 0 marks a row whose program is invalid (the per-sample stem-input table of the plan, pnmn_plan_stats[15])
__global__ void __launch_bounds__(256) answer_loss_fwd_kernel(const float* __restrict__ logits, const int64_t* __restrict__ answers,
                                                              const int64_t* __restrict__ xin, int B, int A, int64_t unknown,
                                                              int64_t* __restrict__ predictions, float* __restrict__ loss,
                                                              uint8_t* __restrict__ invalid_out, int64_t* __restrict__ correct) {
  const int b = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (b >= B) return;
  const float* row = logits + static_cast<size_t>(b) * A;
  float m = -INFINITY;
  int arg = 0x7fffffff;
  for (int a = lane; a < A; a += 32) {
    const float v = row[a];
    if (v > m) { m = v; arg = a; }          // first maximum within the lane's strided subsequence
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, m, o);
    const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
    if (om > m || (om == m && oa < arg)) { m = om; arg = oa; }   // lowest index on ties (torch.max)
  }
  float s = 0.f;
  for (int a = lane; a < A; a += 32) s += expf(row[a] - m);
  s = warp_sum(s);
  const float lse = m + logf(s);
  if (lane == 0) {
    const bool invalid = xin[b] < 0;
    invalid_out[b] = invalid ? 1 : 0;
    const int64_t pred = invalid ? unknown : static_cast<int64_t>(arg);
    predictions[b] = pred;
    float l;
    if (answers) {
      const int64_t t = answers[b];
      l = (t >= 0 && t < A) ? lse - row[t] : NAN;   // (the reference's cross_entropy raises on an out-of-range class)
      if (correct && pred == t) atomicAdd(reinterpret_cast<unsigned long long*>(correct), 1ull);
    } else {
      l = lse - m;
    }
    loss[b] = invalid ? 3.33f : l;
  }
}

// dlogits[b][a] = g[b] * (softmax(logits[b])[a] - [a == label_b]) for valid rows, 0 for invalid ones; label = the answer, or
// the predicted class when no answers were given (loss = -max log-probability)
__global__ void __launch_bounds__(256) answer_loss_bwd_kernel(const float* __restrict__ logits, const int64_t* __restrict__ answers,
                                                              const uint8_t* __restrict__ invalid, const int64_t* __restrict__ predictions,
                                                              const float* __restrict__ gloss, int B, int A,
                                                              float* __restrict__ dlogits) {
  const int b = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (b >= B) return;
  const float* row = logits + static_cast<size_t>(b) * A;
  float* drow = dlogits + static_cast<size_t>(b) * A;
  const float g = gloss[b];
  if (invalid[b] || g == 0.f) {
    for (int a = lane; a < A; a += 32) drow[a] = 0.f;
    return;
  }
  float m = -INFINITY;
  for (int a = lane; a < A; a += 32) m = fmaxf(m, row[a]);
  m = warp_max(m);
  float s = 0.f;
  for (int a = lane; a < A; a += 32) s += expf(row[a] - m);
  s = warp_sum(s);
  const int64_t label = answers ? answers[b] : predictions[b];
  const float inv = 1.f / s;
  for (int a = lane; a < A; a += 32) drow[a] = g * (expf(row[a] - m) * inv - (a == label ? 1.f : 0.f));
}

}  // namespace

extern "C" int pnmn_answer_loss_forward(const float* logits, const int64_t* answers, const int64_t* xin, int batch, int num_answers,
                                        int64_t unknown_index, int64_t* predictions, float* loss, uint8_t* invalid, int64_t* correct,
                                        void* stream) {
  if (!logits || !xin || !predictions || !loss || !invalid) return fail("pnmn_answer_loss_forward: NULL buffer");
  if (batch < 1 || num_answers < 1) return fail("pnmn_answer_loss_forward: empty batch / no answers");
  answer_loss_fwd_kernel<<<(batch + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(logits, answers, xin, batch, num_answers,
                                                                                       unknown_index, predictions, loss, invalid, correct);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(std::string("answer_loss_fwd_kernel: ") + cudaGetErrorString(e));
  pnmn::count_launches(1);
  return 0;
}

extern "C" int pnmn_answer_loss_backward(const float* logits, const int64_t* answers, const uint8_t* invalid, const int64_t* predictions,
                                         const float* grad_loss, int batch, int num_answers, float* grad_logits, void* stream) {
  if (!logits || !invalid || !predictions || !grad_loss || !grad_logits) return fail("pnmn_answer_loss_backward: NULL buffer");
  if (batch < 1 || num_answers < 1) return fail("pnmn_answer_loss_backward: empty batch / no answers");
  answer_loss_bwd_kernel<<<(batch + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(logits, answers, invalid, predictions, grad_loss,
                                                                                       batch, num_answers, grad_logits);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(std::string("answer_loss_bwd_kernel: ") + cudaGetErrorString(e));
  pnmn::count_launches(1);
  return 0;
}
