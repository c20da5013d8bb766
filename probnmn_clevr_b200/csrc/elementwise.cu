// CUDA-core pieces of the NMN module executor: everything that is not a convolution.
// One CTA per EltTask (one sample, one op); all plane operands here are P16 (see executor.h).
// Every kernel writes VALID slots only, so the permanent-zero padding of the arenas survives.
//
// Reference semantics (file:line in /root/reference):
//   ATTEND      feats * attn.repeat(1,C,1,1)           probnmn/modules/nmn_modules.py:83,120,161
//   SAME        argmax / index_select / 1x1 / sigmoid  probnmn/modules/nmn_modules.py:200-208
//   MINMAX      torch.min / torch.max (broadcasting)   probnmn/modules/nmn_modules.py:25-27,43-45
//   DOTSIG_BWD  backward of sigmoid(conv1x1(relu(.)))  probnmn/modules/nmn_modules.py:86,167
//   GATHER      torch.cat(final_module_outputs) / zeros for invalid programs   probnmn/models/nmn.py:233-241
#include <cuda_fp16.h>

#include "elt_body.cuh"

namespace pnmn {

__global__ void __launch_bounds__(256) elt_kernel(const EltTask* __restrict__ tasks) {
  __shared__ EltSmem sm;
  const EltTask t = tasks[blockIdx.x];
  elt_task_body(t, threadIdx.x, sm);
}

// Loss scale of one backward pass: the largest power of two that brings max|d(final)| to <= 2^10,
// so that the fp16 shadow copies of the gradients stay in the normal fp16 range (fp32 planes carry
// the same scaled values; every parameter gradient is multiplied by scale[1] = 1/scale on its way out).
__global__ void __launch_bounds__(256) amax_kernel(const float* __restrict__ x, size_t n, unsigned int* amax_bits) {
  float m = 0.f;
  // 16-byte loads, four in flight per thread (x is a 16-byte aligned tensor; the tail is read element-wise)
  const size_t n4 = n / 4, stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  const float4* x4 = reinterpret_cast<const float4*>(x);
  size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  for (; i + 3 * stride < n4; i += 4 * stride) {
    const float4 a = x4[i], b = x4[i + stride], c = x4[i + 2 * stride], d = x4[i + 3 * stride];
    m = fmaxf(m, fmaxf(fmaxf(fmaxf(fabsf(a.x), fabsf(a.y)), fmaxf(fabsf(a.z), fabsf(a.w))),
                       fmaxf(fmaxf(fabsf(b.x), fabsf(b.y)), fmaxf(fabsf(b.z), fabsf(b.w)))));
    m = fmaxf(m, fmaxf(fmaxf(fmaxf(fabsf(c.x), fabsf(c.y)), fmaxf(fabsf(c.z), fabsf(c.w))),
                       fmaxf(fmaxf(fabsf(d.x), fabsf(d.y)), fmaxf(fabsf(d.z), fabsf(d.w)))));
  }
  for (; i < n4; i += stride) {
    const float4 a = x4[i];
    m = fmaxf(m, fmaxf(fmaxf(fabsf(a.x), fabsf(a.y)), fmaxf(fabsf(a.z), fabsf(a.w))));
  }
  for (size_t j = n4 * 4 + blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; j < n; j += stride) m = fmaxf(m, fabsf(x[j]));
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.f && m < INFINITY) atomicMax(amax_bits, __float_as_uint(m));
}
__global__ void scale_kernel(float* scale /* [0]=scale [1]=1/scale, [2]=amax bits */) {
  const float amax = __uint_as_float(reinterpret_cast<unsigned int*>(scale)[2]);
  float s = 1.f;
  if (amax > 0.f) {
    int e;
    frexpf(amax, &e);  // amax = f * 2^e, f in [0.5, 1)
    s = ldexpf(1.f, 10 - e);
  }
  scale[0] = s;
  scale[1] = 1.f / s;
}

cudaError_t launch_loss_scale(const float* grad, size_t n, float* scale, cudaStream_t stream) {
  cudaError_t e = cudaMemsetAsync(scale, 0, 16, stream);
  if (e != cudaSuccess) return e;
  amax_kernel<<<148 * 4, 256, 0, stream>>>(grad, n, reinterpret_cast<unsigned int*>(scale) + 2);
  scale_kernel<<<1, 1, 0, stream>>>(scale);
  return cudaGetLastError();
}

cudaError_t launch_elt(const EltTask* d_tasks, int n_tasks, cudaStream_t stream) {
  if (n_tasks <= 0) return cudaSuccess;
  elt_kernel<<<n_tasks, 256, 0, stream>>>(d_tasks);
  return cudaGetLastError();
}

}  // namespace pnmn
