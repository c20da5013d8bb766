// Optimiser-side kernels of the joint-training step:
//   * clamp_adam_kernel: element-wise gradient clamp (joint_training_trainer.py:182-188) fused with the Adam update
//     (torch.optim.Adam as constructed in trainers/_trainer.py:103-108) over one flat parameter range;
//   * elbo_glue_kernel: the REINFORCE reward, the moving-average baseline and the ELBO scalars of
//     probnmn/modules/elbo.py:28-34,61-89,256-275 in one launch, with the baseline kept on the device (the reference
//     synchronises once per step on `centered_reward.mean().item()`, elbo.py:33).
// Both are HBM- / latency-bound CUDA-core kernels: 16-byte vector accesses, grid sized to the SM count.
#include <cuda_runtime.h>
#include <algorithm>
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

struct AdamScalars {
  float step_size;     // lr / (1 - beta1^t)
  float bc2_sqrt;      // sqrt(1 - beta2^t)
  float beta1, beta2, eps, weight_decay;
  float clamp;         // > 0: g = min(max(g, -clamp), clamp) first
  int write_grad;      // store the clamped gradient back (parameter.grad.clamp_ is in place in the reference)
};

__device__ __forceinline__ void adam_one(float& p, float& g, float& m, float& v, const AdamScalars& a) {
  if (a.clamp > 0.f) g = fminf(fmaxf(g, -a.clamp), a.clamp);
  float gg = g;
  if (a.weight_decay != 0.f) gg = gg + a.weight_decay * p;     // grad.add(p, alpha=weight_decay)
  m = m * a.beta1 + (1.f - a.beta1) * gg;                       // exp_avg.mul_(beta1).add_(grad, alpha=1-beta1)
  v = v * a.beta2 + (1.f - a.beta2) * gg * gg;                  // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1-beta2)
  const float denom = sqrtf(v) / a.bc2_sqrt + a.eps;            // (exp_avg_sq.sqrt() / sqrt(bias_correction2)).add_(eps)
  p = p - a.step_size * (m / denom);                            // p.addcdiv_(exp_avg, denom, value=-step_size)
}

__global__ void __launch_bounds__(256) clamp_adam_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m,
                                                         float* __restrict__ v, int64_t n, const AdamScalars a) {
  const int64_t n4 = n >> 2;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 P = reinterpret_cast<float4*>(p)[i], G = reinterpret_cast<const float4*>(g)[i];
    float4 M = reinterpret_cast<float4*>(m)[i], V = reinterpret_cast<float4*>(v)[i];
    adam_one(P.x, G.x, M.x, V.x, a); adam_one(P.y, G.y, M.y, V.y, a);
    adam_one(P.z, G.z, M.z, V.z, a); adam_one(P.w, G.w, M.w, V.w, a);
    reinterpret_cast<float4*>(p)[i] = P;
    reinterpret_cast<float4*>(m)[i] = M;
    reinterpret_cast<float4*>(v)[i] = V;
    if (a.write_grad) reinterpret_cast<float4*>(g)[i] = G;
  }
  // tail (n not a multiple of 4)
  for (int64_t i = (n4 << 2) + static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    float P = p[i], G = g[i], M = m[i], V = v[i];
    adam_one(P, G, M, V, a);
    p[i] = P; m[i] = M; v[i] = V;
    if (a.write_grad) g[i] = G;
  }
}

// ---- ELBO / REINFORCE glue: one CTA ---------------------------------------------------------------------------------
__device__ __forceinline__ float block_sum(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float s = 0.f;
  for (int i = 0; i < static_cast<int>(blockDim.x >> 5); ++i) s += red[i];
  return s;
}

// mode 0 ("ours", elbo.py:256-275 / 150-161): reward = lp_rec + beta*lp_prior - beta*lp_gen (+ gamma*lp_ans)
// mode 1 ("baseline", elbo.py:241-251): reward = lp_ans, elbo = mean(pg_loss * centered)
__global__ void __launch_bounds__(256) elbo_glue_kernel(const float* __restrict__ pg_loss, const float* __restrict__ qr_loss,
                                                        const float* __restrict__ prior_loss, const float* __restrict__ nmn_loss,
                                                        int n, float beta, float gamma, float decay, int mode,
                                                        float* __restrict__ baseline, float* __restrict__ centered,
                                                        float* __restrict__ stats) {
  __shared__ float red[8];
  const float b0 = baseline[0];
  float s_rec = 0.f, s_kl = 0.f, s_elbo = 0.f, s_rew = 0.f, s_cent = 0.f, s_nmn = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float lp_gen = -pg_loss[i];
    const float lp_ans = nmn_loss ? -nmn_loss[i] : 0.f;
    float reward, c, elbo_i, kl_i = 0.f, rec_i = 0.f;
    if (mode == 0) {
      const float lp_rec = -qr_loss[i], lp_prior = -prior_loss[i];
      reward = lp_rec + beta * lp_prior - beta * lp_gen;
      if (nmn_loss) reward += gamma * lp_ans;
      c = reward - b0;
      kl_i = lp_gen * c - beta * lp_gen;
      rec_i = lp_rec;
      elbo_i = lp_rec - kl_i;
    } else {
      reward = lp_ans;
      c = reward - b0;
      elbo_i = pg_loss[i] * c;
    }
    centered[i] = c;
    s_rec += rec_i; s_kl += kl_i; s_elbo += elbo_i; s_rew += reward; s_cent += c; s_nmn += nmn_loss ? nmn_loss[i] : 0.f;
  }
  const float inv = n > 0 ? 1.f / static_cast<float>(n) : 0.f;
  s_rec = block_sum(s_rec, red); s_kl = block_sum(s_kl, red); s_elbo = block_sum(s_elbo, red);
  s_rew = block_sum(s_rew, red); s_cent = block_sum(s_cent, red); s_nmn = block_sum(s_nmn, red);
  if (threadIdx.x == 0) {
    stats[0] = s_rec * inv; stats[1] = s_kl * inv; stats[2] = s_elbo * inv; stats[3] = s_rew * inv; stats[4] = s_nmn * inv;
    baseline[0] = b0 + decay * (s_cent * inv);   // self._reinforce_baseline += decay * centered_reward.mean()
  }
}

}  // namespace

extern "C" int pnmn_clamp_adam(float* params, float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, int64_t step,
                               double lr, double beta1, double beta2, double eps, double weight_decay, double clamp,
                               int write_clamped_grad, void* stream) {
  if (!params || !grads || !exp_avg || !exp_avg_sq || n < 0 || step < 1) return fail("pnmn_clamp_adam: bad arguments");
  if (n == 0) return 0;
  if ((reinterpret_cast<uintptr_t>(params) | reinterpret_cast<uintptr_t>(grads) | reinterpret_cast<uintptr_t>(exp_avg) |
       reinterpret_cast<uintptr_t>(exp_avg_sq)) & 15)
    return fail("pnmn_clamp_adam: buffers must be 16-byte aligned");
  AdamScalars a;
  // scalar arithmetic in double on the host, as torch.optim.Adam does in python
  const double bc1 = 1.0 - std::pow(beta1, static_cast<double>(step));
  const double bc2 = 1.0 - std::pow(beta2, static_cast<double>(step));
  a.step_size = static_cast<float>(lr / bc1);
  a.bc2_sqrt = static_cast<float>(std::sqrt(bc2));
  a.beta1 = static_cast<float>(beta1); a.beta2 = static_cast<float>(beta2); a.eps = static_cast<float>(eps);
  a.weight_decay = static_cast<float>(weight_decay); a.clamp = static_cast<float>(clamp);
  a.write_grad = write_clamped_grad;
  const int64_t work = (n + 3) / 4;
  int blocks = static_cast<int>(std::min<int64_t>((work + 255) / 256, 148 * 8));
  clamp_adam_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(params, grads, exp_avg, exp_avg_sq, n, a);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(std::string("clamp_adam_kernel: ") + cudaGetErrorString(e));
  pnmn::count_launches(1);
  return 0;
}

extern "C" int pnmn_elbo_glue(const float* pg_loss, const float* qr_loss, const float* prior_loss, const float* nmn_loss,
                              int n, float beta, float gamma, float baseline_decay, int mode, float* baseline,
                              float* centered, float* stats, void* stream) {
  if (!pg_loss || !baseline || !centered || !stats || n < 0) return fail("pnmn_elbo_glue: bad arguments");
  if (mode == 0 && (!qr_loss || !prior_loss)) return fail("pnmn_elbo_glue: the full objective needs the reconstruction and prior losses");
  if (mode == 1 && !nmn_loss) return fail("pnmn_elbo_glue: the baseline objective needs the answer loss");
  if (mode != 0 && mode != 1) return fail("pnmn_elbo_glue: mode must be 0 (ours) or 1 (baseline)");
  elbo_glue_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(pg_loss, qr_loss, prior_loss, nmn_loss, n, beta, gamma,
                                                                     baseline_decay, mode, baseline, centered, stats);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(std::string("elbo_glue_kernel: ") + cudaGetErrorString(e));
  pnmn::count_launches(1);
  return 0;
}
