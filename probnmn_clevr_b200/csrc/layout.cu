// Layout kernels: reference-layout parameters / NCHW features  <->  executor formats.
#include <cuda_fp16.h>

#include "executor.h"
#include <cuda_bf16.h>
#include "layout.h"
#include "tcgen05.cuh"

namespace pnmn {

// Packed weight tile (the tcgen05 B operand, K-major, no swizzle), fp16, 16 k x 128 n = 4 KB:
//   dst[((kb*ntaps + tap)*2 + kc)*128*8 + n*8 + e] = half( src[(k_off + kb*16 + kc*8 + e)*k_stride
//                                                            + (n_off + n)*n_stride + tap'*tap_stride] )
// with tap' = flip ? ntaps-1-tap : tap.  Forward conv: k = cin, n = cout; dgrad: k = cout, n = cin and
// the taps are mirrored (a correlation with the transposed, flipped kernel).  fp16 keeps the same 10-bit
// mantissa as tf32 (weights below 6.1e-5 in magnitude lose relative precision, which is immaterial).
__global__ void __launch_bounds__(256) pack_weights_kernel(const PackTask* __restrict__ tasks,
                                                           const float* __restrict__ params,
                                                           __half* __restrict__ packed, int n_tasks) {
  // blockIdx.x enumerates tiles of all tasks; tasks carry their first global tile index (sorted)
  int task = 0;
  int tile = blockIdx.x;
  {
    int lo = 0, hi = n_tasks;
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (tasks[mid].first_tile <= tile) lo = mid; else hi = mid;
    }
    task = lo;
  }
  const PackTask t = tasks[task];
  tile -= t.first_tile;
  const int kb = tile / t.ntaps, tap = tile % t.ntaps;
  const int tp = t.flip ? t.ntaps - 1 - tap : tap;
  const float* src = params + t.src_off;
  __half* dst = packed + t.dst_off + static_cast<size_t>(tile) * 2048;
  for (int i = threadIdx.x; i < 2048; i += 256) {
    const int e = i & 7, n = (i >> 3) & 127, kc = i >> 10;
    const int k = t.k_off + kb * 16 + kc * 8 + e;
    const float v = src[static_cast<size_t>(k) * t.k_stride + static_cast<size_t>(t.n_off + n) * t.n_stride +
                        static_cast<size_t>(tp) * t.tap_stride];
    dst[i] = __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f));
  }
}

cudaError_t launch_pack(const PackTask* d_tasks, int n_tasks, int total_tiles, const float* params,
                        void* packed, cudaStream_t stream) {
  if (total_tiles <= 0) return cudaSuccess;
  pack_weights_kernel<<<total_tiles, 256, 0, stream>>>(d_tasks, params, static_cast<__half*>(packed), n_tasks);
  return cudaGetLastError();
}

// NCHW fp32 features [B][C][14][14]  ->  the fp16 half planes [C/8][P16][8] of a stem-input unit (valid slots only; the
// padding slots are the arena's permanent zeros).  A unit is laid out like every activation -- C/4 fp32 planes followed
// by their fp16 shadow -- but nothing ever reads the fp32 planes of the INPUT features (the stem's first conv and its
// weight gradient take the shadow), so only the shadow is written: 103 MB instead of 371 MB per 256 samples.
// dst_off[b] = float offset of sample b's unit inside `dst`, or < 0 to skip the sample (invalid program: its stem is
// never executed).  One block per (sample, 8 channels): 16-byte stores, one per pixel slot.
// T = float: features as the reference passes them.  T = __half: features that already went through
// round_features_f16_kernel (a device-resident feature cache, probnmn_clevr_b200/feed.py): the values are taken as they are,
// which makes the two paths bit-identical.
// dst_off == nullptr: sample b goes to unit b (off = guard_floats + b * unit_floats), whatever its program (pnmn_nmn_prestage).
template <class T>
__global__ void __launch_bounds__(256) nchw_to_planes_kernel(const T* __restrict__ src, float* __restrict__ dst,
                                                             int C, const int64_t* __restrict__ dst_off, int64_t guard_floats,
                                                             int64_t unit_floats) {
  const int b = blockIdx.y;
  const int hp = blockIdx.x;  // half plane = 8 channels
  const int64_t off = dst_off ? dst_off[b] : guard_floats + b * unit_floats;
  if (off < 0) return;
  const T* s = src + (static_cast<size_t>(b) * C + hp * 8) * 196;
  __shared__ float tile[8 * 196];
  if (sizeof(T) == 4) {
    for (int i = threadIdx.x; i < 8 * 196 / 4; i += 256)
      reinterpret_cast<float4*>(tile)[i] = reinterpret_cast<const float4*>(s)[i];  // (8 * 196 floats, 16-byte aligned)
  } else {
    for (int i = threadIdx.x; i < 8 * 196 / 8; i += 256) {                          // (8 * 196 halves, 16-byte aligned)
      const uint4 v = reinterpret_cast<const uint4*>(s)[i];
      const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = __half22float2(h[e]);
        tile[8 * i + 2 * e] = f.x; tile[8 * i + 2 * e + 1] = f.y;
      }
    }
  }
  __syncthreads();
  if (threadIdx.x < 196) {
    const int p = threadIdx.x, slot = (p / kHW) * 16 + (p % kHW);
    uint32_t h[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      // same value chain as every other activation: tf32 rounding first, then the (saturating) fp16 operand copy
      const __half2 v = __floats2half2_rn(fminf(to_tf32(tile[(2 * e) * 196 + p]), 65504.f),
                                          fminf(to_tf32(tile[(2 * e + 1) * 196 + p]), 65504.f));
      h[e] = *reinterpret_cast<const uint32_t*>(&v);
    }
    uint8_t* hb = reinterpret_cast<uint8_t*>(dst + off) + static_cast<size_t>(C / 4) * 256 * 16;
    *reinterpret_cast<uint4*>(hb + (static_cast<size_t>(hp) * 256 + slot) * 16) = make_uint4(h[0], h[1], h[2], h[3]);
  }
}

cudaError_t launch_nchw_to_planes(const void* src, int src_is_half, float* dst, int B, int C, const int64_t* dst_off,
                                  int64_t guard_floats, int64_t unit_floats, cudaStream_t stream) {
  if (B <= 0) return cudaSuccess;
  if (src_is_half)
    nchw_to_planes_kernel<__half><<<dim3(C / 8, B), 256, 0, stream>>>(static_cast<const __half*>(src), dst, C, dst_off, guard_floats, unit_floats);
  else
    nchw_to_planes_kernel<float><<<dim3(C / 8, B), 256, 0, stream>>>(static_cast<const float*>(src), dst, C, dst_off, guard_floats, unit_floats);
  return cudaGetLastError();
}

// dst[i] = the fp16 operand value the executor derives from an fp32 feature (tf32 rounding, then the saturating fp16 copy):
// what a feature cache stores, 2 bytes per value, so that cached and freshly uploaded features give identical results
__global__ void round_features_f16_kernel(const float* __restrict__ src, __half* __restrict__ dst, int64_t n4) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n4; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(src)[i];
    const __half2 a = __floats2half2_rn(fminf(to_tf32(v.x), 65504.f), fminf(to_tf32(v.y), 65504.f));
    const __half2 b = __floats2half2_rn(fminf(to_tf32(v.z), 65504.f), fminf(to_tf32(v.w), 65504.f));
    uint2 o;
    o.x = *reinterpret_cast<const uint32_t*>(&a); o.y = *reinterpret_cast<const uint32_t*>(&b);
    reinterpret_cast<uint2*>(dst)[i] = o;
  }
}
cudaError_t launch_round_features_f16(const float* src, void* dst, int64_t n, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  const int64_t n4 = n / 4;
  const int blocks = static_cast<int>(n4 / 256 + 1 < 148 * 16 ? n4 / 256 + 1 : 148 * 16);
  round_features_f16_kernel<<<blocks, 256, 0, stream>>>(src, static_cast<__half*>(dst), n4);
  return cudaGetLastError();
}


// ---- classifier: ReLU + 2x2 max-pool + (C,7,7) flatten of the channels-last 1x1-conv output (nmn.py:77-79, nmn_modules.py:250)
// y: [B][196][C] fp32 (bias already added), pooled: [B][C*49] fp32, code: [B][C*49] bytes = argmax position in the
// window (bits 0-1: dy*2 + dx, first maximum in scan order like ATen's max_pool2d) | 4 if the pooled value is > 0.
// One block per (sample, 64 channels): every global access is a contiguous 256-byte (reads) or 12.5 KB (writes) run.
constexpr int kPoolCB = 64;
// `bias` (optional, [C]) is added to y first -- max(v_i) + b == max(v_i + b) exactly, so the conv bias costs nothing here,
// where the library GEMM would first broadcast it into its 205 MB output.
__global__ void __launch_bounds__(256) relu_pool_fwd_kernel(const float* __restrict__ y, const float* __restrict__ bias,
                                                            float* __restrict__ pooled, uint8_t* __restrict__ code, int C) {
  __shared__ float tile[kPoolCB * 49];
  __shared__ uint8_t ctile[kPoolCB * 49];
  const int b = blockIdx.y, c0 = blockIdx.x * kPoolCB;
  const int cw = threadIdx.x % kPoolCB, wq = threadIdx.x / kPoolCB;
  const float bb = bias ? bias[c0 + cw] : 0.f;
  const float* yb = y + (static_cast<size_t>(b) * 196) * C + c0 + cw;
  for (int w = wq; w < 49; w += 4) {
    const int py = w / 7, px = w - py * 7;
    const float* p00 = yb + static_cast<size_t>((2 * py) * 14 + 2 * px) * C;
    const float v0 = p00[0] + bb, v1 = p00[C] + bb, v2 = p00[static_cast<size_t>(14) * C] + bb, v3 = p00[static_cast<size_t>(15) * C] + bb;
    float m = v0; int a = 0;
    if (v1 > m) { m = v1; a = 1; }
    if (v2 > m) { m = v2; a = 2; }
    if (v3 > m) { m = v3; a = 3; }
    tile[cw * 49 + w] = fmaxf(m, 0.f);
    ctile[cw * 49 + w] = static_cast<uint8_t>(a | (m > 0.f ? 4 : 0));
  }
  __syncthreads();
  const size_t o = static_cast<size_t>(b) * C * 49 + static_cast<size_t>(c0) * 49;
  for (int i = threadIdx.x; i < kPoolCB * 49; i += 256) { pooled[o + i] = tile[i]; code[o + i] = ctile[i]; }
}
__global__ void __launch_bounds__(256) relu_pool_bwd_kernel(const float* __restrict__ g, const uint8_t* __restrict__ code,
                                                            float* __restrict__ gy, int C) {
  __shared__ float tile[kPoolCB * 49];
  __shared__ uint8_t ctile[kPoolCB * 49];
  const int b = blockIdx.y, c0 = blockIdx.x * kPoolCB;
  const size_t o = static_cast<size_t>(b) * C * 49 + static_cast<size_t>(c0) * 49;
  for (int i = threadIdx.x; i < kPoolCB * 49; i += 256) { tile[i] = g[o + i]; ctile[i] = code[o + i]; }
  __syncthreads();
  const int cw = threadIdx.x % kPoolCB, pq = threadIdx.x / kPoolCB;
  float* gb = gy + (static_cast<size_t>(b) * 196) * C + c0 + cw;
  for (int p = pq; p < 196; p += 4) {
    const int yy = p / 14, xx = p - yy * 14;
    const int w = (yy >> 1) * 7 + (xx >> 1), pos = (yy & 1) * 2 + (xx & 1);
    const int cd = ctile[cw * 49 + w];
    gb[static_cast<size_t>(p) * C] = (cd == (pos | 4)) ? tile[cw * 49 + w] : 0.f;
  }
}
cudaError_t launch_relu_pool_fwd(const float* y, const float* bias, float* pooled, uint8_t* code, int B, int C, cudaStream_t st) {
  if (B <= 0) return cudaSuccess;
  relu_pool_fwd_kernel<<<dim3(C / kPoolCB, B), 256, 0, st>>>(y, bias, pooled, code, C);
  return cudaGetLastError();
}
cudaError_t launch_relu_pool_bwd(const float* g, const uint8_t* code, float* gy, int B, int C, cudaStream_t st) {
  if (B <= 0) return cudaSuccess;
  relu_pool_bwd_kernel<<<dim3(C / kPoolCB, B), 256, 0, st>>>(g, code, gy, C);
  return cudaGetLastError();
}

}  // namespace pnmn
