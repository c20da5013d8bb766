// CUDA-core twin of the executor's convolution task (bring-up / debugging only).
//
//   out[slot, n] = epilogue( sum_{kb, tap, k} in[(kb,k)-th channel][slot + shift(tap)] * W[kb][tap][k][n] )
//
// Same ConvTask / ConvCfg semantics, same fp16 operands (half planes, packed fp16 weight tiles) and
// the same fused epilogue as the tcgen05 path in exec.cu, but plain fp32 FMAs on CUDA cores.  It is
// selected only by PNMN_EXEC=levels (one launch per level, used by tests to separate tensor-core
// descriptor problems from scheduling / layout problems); the product path is the persistent
// tcgen05 executor (exec.cu).
//
// Reference semantics: nn.Conv2d(128,128,3,padding=d,dilation=d) + F.relu and the 1x1 + sigmoid
// heads of probnmn/modules/nmn_modules.py:82-87,119-123,160-168,239-244; stem nmn.py:67-72.
#include <cuda_fp16.h>

#include "executor.h"
#include "layout.h"
#include "tcgen05.cuh"

namespace pnmn {

// CUDA-core twin of the tensor-core convolution tasks: bring-up / cross-check builds only (make BRINGUP=1).  The release
// library carries no second implementation of a hot op.
#ifdef PNMN_BRINGUP

__device__ __forceinline__ int tap_shift(const ConvCfg& c, int tap) {
  return c.ntaps == 9 ? ((tap / 3 - 1) * c.S_in + (tap % 3 - 1)) * c.dil : 0;
}
__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
  const __half2 h = __floats2half2_rn(fminf(fmaxf(a, -65504.f), 65504.f), fminf(fmaxf(b, -65504.f), 65504.f));
  return *reinterpret_cast<const uint32_t*>(&h);
}

// Fused epilogue for one output row (pixel slot) and 4 consecutive output channels.
__device__ __forceinline__ void epilogue4(const ConvTask& t, const ConvCfg& c, int s, int kc, int so,
                                          int sx, float4& o, float& dot) {
  const int n0 = kc * 4;
  if (c.flags & F_BIAS) {
    const float4 b = __ldg(reinterpret_cast<const float4*>(t.bias + n0));
    o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
  }
  if (c.flags & F_RELU) {
    o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
  }
  if (c.flags & F_MASK) {
    const float4 m =
        *reinterpret_cast<const float4*>(t.aux[s] + (static_cast<size_t>(kc) * c.P_aux + sx) * 4);
    o.x = m.x > 0.f ? o.x : 0.f; o.y = m.y > 0.f ? o.y : 0.f;
    o.z = m.z > 0.f ? o.z : 0.f; o.w = m.w > 0.f ? o.w : 0.f;
  }
  float4* dst = reinterpret_cast<float4*>(t.out[s] + (static_cast<size_t>(kc) * c.P_out + so) * 4);
  if (c.flags & F_ACCUM) {
    const float4 old = *dst;
    o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
  }
  if (c.flags & F_DOTSIG) {
    const float4 w = __ldg(reinterpret_cast<const float4*>(t.w3 + n0));
    dot = fmaf(o.x, w.x, dot); dot = fmaf(o.y, w.y, dot);
    dot = fmaf(o.z, w.z, dot); dot = fmaf(o.w, w.w, dot);
  }
  o.x = to_tf32(o.x); o.y = to_tf32(o.y); o.z = to_tf32(o.z); o.w = to_tf32(o.w);
  if (c.flags & F_STORE) *dst = o;
}

__global__ void __launch_bounds__(256)
conv_simt_kernel(const ConvTask* __restrict__ tasks, const ConvCfg* __restrict__ cfgs) {
  const ConvTask t = tasks[blockIdx.x];
  const ConvCfg c = cfgs[t.cfg];
  const int nmt = c.P_in == 484 ? 3 : 2;
  const int rows = nmt * 128;
  // one thread per (sample, row, kc-out); 4 output channels each
  const int total = t.n_samp * rows * kKC;
  __shared__ float dots[NSMAX * 512];
  for (int i = threadIdx.x; i < NSMAX * 512; i += blockDim.x) dots[i] = 0.f;
  __syncthreads();
  for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
    const int kc_out = idx % kKC;
    const int r = (idx / kKC) % rows;
    const int s = idx / (kKC * rows);
    const int y = r / c.S_in, x = r - y * c.S_in;
    if (y >= kHW || x >= kHW || r < t.mt0 * 128 || r >= (t.mt0 + t.n_mt) * 128) continue;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int kb = 0; kb < c.n_kb; ++kb) {
      const int which = kb / c.kb_per_in, kbl = kb % c.kb_per_in;
      const __half* in = static_cast<const __half*>(t.in[which][s]);
      for (int tap = 0; tap < c.ntaps; ++tap) {
        const int rr = r + tap_shift(c, tap);
        const __half* wt = static_cast<const __half*>(t.w) + (static_cast<size_t>(kb) * c.ntaps + tap) * 2048;
        for (int kc = 0; kc < 2; ++kc) {
          const int plane = kbl * 2 + kc;
          if (plane == 0 && rr < 0) continue;  // staged lead gap == zeros
          const __half* a = in + (static_cast<ptrdiff_t>(plane) * c.P_in + rr) * 8;
          for (int e = 0; e < 8; ++e) {
            const float av = __half2float(a[e]);
            for (int j = 0; j < 4; ++j)
              acc[j] = fmaf(av, __half2float(wt[(kc * 128 + kc_out * 4 + j) * 8 + e]), acc[j]);
          }
        }
      }
    }
    float dot = 0.f;
    float4 o = make_float4(acc[0], acc[1], acc[2], acc[3]);
    const int so = y * c.S_out + x;
    epilogue4(t, c, s, kc_out, so, y * c.S_aux + x, o, dot);
    if (c.flags & F_HALF) {
      uint8_t* hb = reinterpret_cast<uint8_t*>(t.out[s]) + shadow_bytes(c.P_out);
      *reinterpret_cast<uint2*>(hb + (static_cast<size_t>(kc_out >> 1) * c.P_out + so) * 16 + (kc_out & 1) * 8) =
          make_uint2(pack_half2(o.x, o.y), pack_half2(o.z, o.w));
    }
    if (c.flags & F_DOTSIG) atomicAdd(&dots[s * 512 + r], dot);
  }
  __syncthreads();
  if (c.flags & F_DOTSIG) {
    for (int i = threadIdx.x; i < t.n_samp * rows; i += blockDim.x) {
      const int s = i / rows, r = i % rows;
      const int y = r / c.S_in, x = r - y * c.S_in;
      if (y < kHW && x < kHW && r >= t.mt0 * 128 && r < (t.mt0 + t.n_mt) * 128)
        t.map_out[s][y * 16 + x] = 1.f / (1.f + expf(-(dots[s * 512 + r] + __ldg(t.b3))));
    }
  }
}

cudaError_t launch_conv_simt(const ConvTask* d_tasks, int n_tasks, const ConvCfg* d_cfgs, cudaStream_t stream) {
  if (n_tasks <= 0) return cudaSuccess;
  conv_simt_kernel<<<n_tasks, 256, 0, stream>>>(d_tasks, d_cfgs);
  return cudaGetLastError();
}
#else
cudaError_t launch_conv_simt(const ConvTask* d_tasks, int n_tasks, const ConvCfg* d_cfgs, cudaStream_t stream) {
  (void)d_tasks; (void)n_tasks; (void)d_cfgs; (void)stream;
  return cudaErrorNotSupported;   // built without the bring-up kernels
}
#endif

}  // namespace pnmn
