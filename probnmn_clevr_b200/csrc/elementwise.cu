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

#include "elt.h"
#include "tcgen05.cuh"

namespace pnmn {

__device__ __forceinline__ int valid_slot16(int i) {  // i in [0,196) -> P16 slot
  return (i / kHW) * 16 + (i % kHW);
}

__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
  const __half2 h = __floats2half2_rn(fminf(fmaxf(a, -65504.f), 65504.f), fminf(fmaxf(b, -65504.f), 65504.f));
  return *reinterpret_cast<const uint32_t*>(&h);
}
// fp16 shadow of 4 channels (plane kc) of P16 slot s, stored behind the 32 fp32 planes of `base`
__device__ __forceinline__ void st_half4(float* base, int kc, int s, float4 v) {
  uint8_t* hb = reinterpret_cast<uint8_t*>(base) + shadow_bytes(256);
  *reinterpret_cast<uint2*>(hb + (static_cast<size_t>(kc >> 1) * 256 + s) * 16 + (kc & 1) * 8) =
      make_uint2(pack_half2(v.x, v.y), pack_half2(v.z, v.w));
}

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

__device__ __forceinline__ float block_sum(float v, float* red) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  float s = 0.f;
  for (int i = 0; i < (blockDim.x >> 5); ++i) s += red[i];
  return s;
}

__global__ void __launch_bounds__(256) elt_kernel(const EltTask* __restrict__ tasks) {
  const EltTask t = tasks[blockIdx.x];
  const int tid = threadIdx.x;
  __shared__ float sh[256];
  __shared__ float red[8];
  __shared__ int sh_idx;

  switch (t.op) {
    case OP_ATTEND: {  // o = a(feat) * b(map)
      for (int i = tid; i < kKC * 196; i += 256) {
        const int kc = i / 196, s = valid_slot16(i % 196);
        const float m = t.b[s];
        float4 v = ld4(t.a + (kc * 256 + s) * 4);
        v.x = to_tf32(v.x * m); v.y = to_tf32(v.y * m); v.z = to_tf32(v.z * m); v.w = to_tf32(v.w * m);
        st4(t.o + (kc * 256 + s) * 4, v);
        if (t.flags & EF_HALF) st_half4(t.o, kc, s, v);
      }
    } break;

    case OP_ATTEND_BWD: {  // a = dX0, b = feat, c = map; o = dmap (+=), o2 = dfeat (write / +=)
      for (int i = tid; i < kKC * 196; i += 256) {
        const int kc = i / 196, s = valid_slot16(i % 196);
        const float m = t.c[s];
        float4 g = ld4(t.a + (kc * 256 + s) * 4);
        g.x *= m; g.y *= m; g.z *= m; g.w *= m;
        float* d = t.o2 + (kc * 256 + s) * 4;
        if (t.flags & EF_ACCUM) {
          const float4 old = ld4(d);
          g.x += old.x; g.y += old.y; g.z += old.z; g.w += old.w;
        }
        st4(d, g);
      }
      if (t.o != nullptr && tid < 196) {
        const int s = valid_slot16(tid);
        float acc = 0.f;
        for (int kc = 0; kc < kKC; ++kc) {
          const float4 g = ld4(t.a + (kc * 256 + s) * 4), f = ld4(t.b + (kc * 256 + s) * 4);
          acc += g.x * f.x + g.y * f.y + g.z * f.z + g.w * f.w;
        }
        t.o[s] += acc;
      }
    } break;

    case OP_SAME: {  // a = feat, b = map, w = [129], c = bias[1]; o = out map, idx = argmax slot
      // argmax over valid pixels, first maximum in row-major order wins (F.max_pool2d indices)
      float bv = -INFINITY;
      int bi = 1 << 30;
      if (tid < 196) { bv = t.b[valid_slot16(tid)]; bi = tid; }
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
      }
      __shared__ float wv[8];
      __shared__ int wi[8];
      if ((tid & 31) == 0) { wv[tid >> 5] = bv; wi[tid >> 5] = bi; }
      __syncthreads();
      if (tid == 0) {
        for (int i = 1; i < 8; ++i)
          if (wv[i] > bv || (wv[i] == bv && wi[i] < bi)) { bv = wv[i]; bi = wi[i]; }
        sh_idx = valid_slot16(bi);
        if (t.idx) t.idx[0] = sh_idx;
      }
      __syncthreads();
      const int is = sh_idx;
      if (tid < 128) sh[tid] = t.a[((tid >> 2) * 256 + is) * 4 + (tid & 3)] * t.w[tid];  // v_c * w_c
      __syncthreads();
      if (tid < 196) {
        const int s = valid_slot16(tid);
        float acc = 0.f;
        for (int kc = 0; kc < kKC; ++kc) {
          const float4 f = ld4(t.a + (kc * 256 + s) * 4);
          acc += f.x * sh[kc * 4] + f.y * sh[kc * 4 + 1] + f.z * sh[kc * 4 + 2] + f.w * sh[kc * 4 + 3];
        }
        acc += t.b[s] * t.w[128] + t.c[0];
        t.o[s] = 1.f / (1.f + expf(-acc));
      }
    } break;

    case OP_SAME_BWD: {
      // a = feat, b = in map, c = out map (sigmoid), g = d(out map), w = [129];
      // o = d(in map) (+=), o2 = dfeat (write / +=), dw = [129] (atomic), dw2 = dbias (atomic)
      const int is = t.idx[0];
      float gp = 0.f;  // g[p] * out*(1-out)
      int s = 0;
      if (tid < 196) {
        s = valid_slot16(tid);
        const float o = t.c[s];
        gp = t.g[s] * o * (1.f - o);
        if (t.o) t.o[s] += gp * t.w[128];
      }
      sh[tid] = gp;
      const float sum_g = block_sum(gp, red);
      const float sum_ga = block_sum(tid < 196 ? gp * t.b[s] : 0.f, red);
      const float unscale = t.scale[1];
      if (tid == 0) { atomicAdd(t.dw2, sum_g * unscale); atomicAdd(t.dw + 128, sum_ga * unscale); }
      __syncthreads();
      if (tid < 128) {
        // channel tid: q = sum_p g[p]*feat[c][p];  dw_c += q*v_c;  dfeat[c][p] (+)= g[p]*w_c*v_c; dfeat[c][is] += q*w_c
        const int kc = tid >> 2, e = tid & 3;
        const float v = t.a[(kc * 256 + is) * 4 + e], wc = t.w[tid];
        float q = 0.f;
        for (int i = 0; i < 196; ++i) {
          const int sl = valid_slot16(i);
          const float gi = sh[i];
          q = fmaf(gi, t.a[(kc * 256 + sl) * 4 + e], q);
          float* d = t.o2 + (kc * 256 + sl) * 4 + e;
          const float val = gi * wc * v;
          *d = (t.flags & EF_ACCUM) ? *d + val : val;
        }
        atomicAdd(t.dw + tid, q * v * unscale);
        t.o2[(kc * 256 + is) * 4 + e] += q * wc;
      }
    } break;

    case OP_MINMAX: {  // a, b -> o ; EF_MAX, EF_A_MAP, EF_B_MAP
      const bool mx = t.flags & EF_MAX, am = t.flags & EF_A_MAP, bm = t.flags & EF_B_MAP;
      if (am && bm) {
        if (tid < 196) {
          const int s = valid_slot16(tid);
          t.o[s] = mx ? fmaxf(t.a[s], t.b[s]) : fminf(t.a[s], t.b[s]);
        }
      } else {
        for (int i = tid; i < kKC * 196; i += 256) {
          const int kc = i / 196, s = valid_slot16(i % 196);
          float4 x, y;
          if (am) { const float m = t.a[s]; x = make_float4(m, m, m, m); } else x = ld4(t.a + (kc * 256 + s) * 4);
          if (bm) { const float m = t.b[s]; y = make_float4(m, m, m, m); } else y = ld4(t.b + (kc * 256 + s) * 4);
          float4 r;
          r.x = mx ? fmaxf(x.x, y.x) : fminf(x.x, y.x); r.y = mx ? fmaxf(x.y, y.y) : fminf(x.y, y.y);
          r.z = mx ? fmaxf(x.z, y.z) : fminf(x.z, y.z); r.w = mx ? fmaxf(x.w, y.w) : fminf(x.w, y.w);
          st4(t.o + (kc * 256 + s) * 4, r);
          if (t.flags & EF_HALF) st_half4(t.o, kc, s, r);
        }
      }
    } break;

    case OP_MINMAX_BWD: {
      // g = d(out); a, b = forward operands; o = d(a), o2 = d(b) (nullptr = no grad needed).
      // torch.minimum/maximum backward: the selected operand gets g, ties split g/2 each.
      // map-typed gradients are accumulated (+=); plane-typed use EF_ACCUM (o) / EF_ACCUM2 (o2).
      const bool mx = t.flags & EF_MAX, am = t.flags & EF_A_MAP, bm = t.flags & EF_B_MAP;
      if (tid < 196) {
        const int s = valid_slot16(tid);
        float ga_map = 0.f, gb_map = 0.f;
        const int nkc = (am && bm) ? 1 : kKC;
        for (int kc = 0; kc < nkc; ++kc) {
          float xa[4], xb[4], gg[4], ra[4], rb[4];
          if (am) { xa[0] = xa[1] = xa[2] = xa[3] = t.a[s]; } else { const float4 v = ld4(t.a + (kc * 256 + s) * 4); xa[0] = v.x; xa[1] = v.y; xa[2] = v.z; xa[3] = v.w; }
          if (bm) { xb[0] = xb[1] = xb[2] = xb[3] = t.b[s]; } else { const float4 v = ld4(t.b + (kc * 256 + s) * 4); xb[0] = v.x; xb[1] = v.y; xb[2] = v.z; xb[3] = v.w; }
          if (am && bm) { gg[0] = t.g[s]; gg[1] = gg[2] = gg[3] = 0.f; } else { const float4 v = ld4(t.g + (kc * 256 + s) * 4); gg[0] = v.x; gg[1] = v.y; gg[2] = v.z; gg[3] = v.w; }
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const bool a_sel = mx ? (xa[e] > xb[e]) : (xa[e] < xb[e]);
            const bool tie = xa[e] == xb[e];
            ra[e] = tie ? 0.5f * gg[e] : (a_sel ? gg[e] : 0.f);
            rb[e] = tie ? 0.5f * gg[e] : (a_sel ? 0.f : gg[e]);
          }
          if (am) ga_map += ra[0] + ra[1] + ra[2] + ra[3];
          else if (t.o) {
            float* d = t.o + (kc * 256 + s) * 4;
            float4 v = make_float4(ra[0], ra[1], ra[2], ra[3]);
            if (t.flags & EF_ACCUM) { const float4 old = ld4(d); v.x += old.x; v.y += old.y; v.z += old.z; v.w += old.w; }
            st4(d, v);
          }
          if (bm) gb_map += rb[0] + rb[1] + rb[2] + rb[3];
          else if (t.o2) {
            float* d = t.o2 + (kc * 256 + s) * 4;
            float4 v = make_float4(rb[0], rb[1], rb[2], rb[3]);
            if (t.flags & EF_ACCUM2) { const float4 old = ld4(d); v.x += old.x; v.y += old.y; v.z += old.z; v.w += old.w; }
            st4(d, v);
          }
        }
        if (am && t.o) t.o[s] += ga_map;
        if (bm && t.o2) t.o2[s] += gb_map;
      }
    } break;

    case OP_DOTSIG_BWD: {
      // g = d(map out), c = map out (sigmoid), a = Y (relu output feeding the 1x1), w = w3[128]
      // o = dZ (write) = g*s*(1-s)*w3[n] masked by Y>0 ; dw[n] += sum_p gp*Y[n][p] ; dw2 += sum gp
      float gp = 0.f;
      if (tid < 196) {
        const int s = valid_slot16(tid);
        const float o = t.c[s];
        gp = t.g[s] * o * (1.f - o);
      }
      sh[tid] = gp;
      const float sum_g = block_sum(gp, red);
      const float unscale = t.scale[1];
      if (tid == 0) atomicAdd(t.dw2, sum_g * unscale);
      __syncthreads();
      for (int i = tid; i < kKC * 196; i += 256) {
        const int kc = i / 196, p = i % 196, s = valid_slot16(p);
        const float4 y = ld4(t.a + (kc * 256 + s) * 4);
        const float4 w = ld4(t.w + kc * 4);
        const float gi = sh[p];
        float4 d;
        d.x = y.x > 0.f ? to_tf32(gi * w.x) : 0.f; d.y = y.y > 0.f ? to_tf32(gi * w.y) : 0.f;
        d.z = y.z > 0.f ? to_tf32(gi * w.z) : 0.f; d.w = y.w > 0.f ? to_tf32(gi * w.w) : 0.f;
        st4(t.o + (kc * 256 + s) * 4, d);
        st_half4(t.o, kc, s, d);
      }
      if (tid < 128) {
        const int kc = tid >> 2, e = tid & 3;
        float q = 0.f;
        for (int p = 0; p < 196; ++p) q = fmaf(sh[p], t.a[(kc * 256 + valid_slot16(p)) * 4 + e], q);
        atomicAdd(t.dw + tid, q * unscale);
      }
    } break;

    case OP_RELU_MASK: {  // o = a (dY) masked by b (Y) > 0
      for (int i = tid; i < kKC * 196; i += 256) {
        const int kc = i / 196, s = valid_slot16(i % 196);
        float4 g = ld4(t.a + (kc * 256 + s) * 4);
        const float4 y = ld4(t.b + (kc * 256 + s) * 4);
        g.x = y.x > 0.f ? to_tf32(g.x) : 0.f; g.y = y.y > 0.f ? to_tf32(g.y) : 0.f;
        g.z = y.z > 0.f ? to_tf32(g.z) : 0.f; g.w = y.w > 0.f ? to_tf32(g.w) : 0.f;
        st4(t.o + (kc * 256 + s) * 4, g);
        st_half4(t.o, kc, s, g);
      }
    } break;

    case OP_SCATTER: {  // a = NCHW [128][196] grad (times loss scale), b = Y planes (mask, optional) -> o planes
      const float sc = t.scale[0];
      for (int i = tid; i < kKC * 196; i += 256) {
        const int kc = i / 196, p = i % 196, s = valid_slot16(p);
        float4 g = make_float4(sc * t.a[(kc * 4) * 196 + p], sc * t.a[(kc * 4 + 1) * 196 + p],
                               sc * t.a[(kc * 4 + 2) * 196 + p], sc * t.a[(kc * 4 + 3) * 196 + p]);
        float* d = t.o + (kc * 256 + s) * 4;
        if (t.flags & EF_ACCUM) { const float4 old = ld4(d); g.x += old.x; g.y += old.y; g.z += old.z; g.w += old.w; }
        if (t.b) {
          const float4 y = ld4(t.b + (kc * 256 + s) * 4);
          g.x = y.x > 0.f ? to_tf32(g.x) : 0.f; g.y = y.y > 0.f ? to_tf32(g.y) : 0.f;
          g.z = y.z > 0.f ? to_tf32(g.z) : 0.f; g.w = y.w > 0.f ? to_tf32(g.w) : 0.f;
        }
        st4(d, g);
        st_half4(t.o, kc, s, g);
      }
    } break;

    case OP_GATHER: {  // a = planes (nullptr -> zeros) -> o NCHW [128][196]
      for (int i = tid; i < 128 * 196; i += 256) {
        const int c = i / 196, p = i % 196;
        t.o[i] = t.a ? t.a[((c >> 2) * 256 + valid_slot16(p)) * 4 + (c & 3)] : 0.f;
      }
    } break;

    default: break;
  }
}

// Loss scale of one backward pass: the largest power of two that brings max|d(final)| to <= 2^10,
// so that the fp16 shadow copies of the gradients stay in the normal fp16 range (fp32 planes carry
// the same scaled values; every parameter gradient is multiplied by scale[1] = 1/scale on its way out).
__global__ void __launch_bounds__(256) amax_kernel(const float* __restrict__ x, size_t n, unsigned int* amax_bits) {
  float m = 0.f;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x)
    m = fmaxf(m, fabsf(x[i]));
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
  amax_kernel<<<296, 256, 0, stream>>>(grad, n, reinterpret_cast<unsigned int*>(scale) + 2);
  scale_kernel<<<1, 1, 0, stream>>>(scale);
  return cudaGetLastError();
}

cudaError_t launch_elt(const EltTask* d_tasks, int n_tasks, cudaStream_t stream) {
  if (n_tasks <= 0) return cudaSuccess;
  elt_kernel<<<n_tasks, 256, 0, stream>>>(d_tasks);
  return cudaGetLastError();
}

}  // namespace pnmn
