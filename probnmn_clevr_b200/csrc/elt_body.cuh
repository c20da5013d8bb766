// Device body of the CUDA-core (non-convolution) executor ops, shared by the stand-alone elt_kernel
// (bring-up / level-synchronous path) and the persistent executor kernel (exec.cu).
// Exactly 256 threads run a task; `ELT_SYNC()` is a named barrier over those 256 threads.
//
// Reference semantics (file:line in /root/reference):
//   ATTEND      feats * attn.repeat(1,C,1,1)           probnmn/modules/nmn_modules.py:83,120,161
//   SAME        argmax / index_select / 1x1 / sigmoid  probnmn/modules/nmn_modules.py:200-208
//   MINMAX      torch.min / torch.max (broadcasting)   probnmn/modules/nmn_modules.py:25-27,43-45
//   DOTSIG_BWD  backward of sigmoid(conv1x1(relu(.)))  probnmn/modules/nmn_modules.py:86,167
//   GATHER      torch.cat(final_module_outputs) / zeros for invalid programs   probnmn/models/nmn.py:233-241
#pragma once
#include <cuda_fp16.h>

#include "elt.h"
#include "tcgen05.cuh"

namespace pnmn {

#define ELT_SYNC() asm volatile("bar.sync 1, 256;" ::: "memory")

struct EltSmem {
  float sh[256];
  float red[8];
  float wv[8];
  int wi[8];
  int idx;
};

__device__ __forceinline__ int valid_slot16(int i) {  // i in [0,196) -> P16 slot
  return (i / kHW) * 16 + (i % kHW);
}
__device__ __forceinline__ uint32_t elt_pack_half2(float a, float b) {
  const __half2 h = __floats2half2_rn(fminf(fmaxf(a, -65504.f), 65504.f), fminf(fmaxf(b, -65504.f), 65504.f));
  return *reinterpret_cast<const uint32_t*>(&h);
}
// fp16 shadow of 4 channels (plane kc) of P16 slot s, stored behind the 32 fp32 planes of `base`
__device__ __forceinline__ void st_half4(float* base, int kc, int s, float4 v) {
  uint8_t* hb = reinterpret_cast<uint8_t*>(base) + shadow_bytes(256);
  *reinterpret_cast<uint2*>(hb + (static_cast<size_t>(kc >> 1) * 256 + s) * 16 + (kc & 1) * 8) =
      make_uint2(elt_pack_half2(v.x, v.y), elt_pack_half2(v.z, v.w));
}
// Explicit global-space 16-byte accesses.  Every pointer of a task record is loaded from memory, so the compiler can only
// emit GENERIC loads / stores (LD.E / ST.E) for plain dereferences; the .global forms keep the LSU on its fast path.
__device__ __forceinline__ uint4 ldg128(const void* p) {
  uint4 v;
  asm volatile("ld.global.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void stg128(void* p, uint4 v) {
  asm volatile("st.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ float4 ld4(const float* p) {
  const uint4 v = ldg128(p);
  return make_float4(__uint_as_float(v.x), __uint_as_float(v.y), __uint_as_float(v.z), __uint_as_float(v.w));
}
__device__ __forceinline__ void st4(float* p, float4 v) {
  stg128(p, make_uint4(__float_as_uint(v.x), __float_as_uint(v.y), __float_as_uint(v.z), __float_as_uint(v.w)));
}

__device__ __forceinline__ float elt_block_sum(float v, float* red, int tid) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  ELT_SYNC();
  if ((tid & 31) == 0) red[tid >> 5] = v;
  ELT_SYNC();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += red[i];
  return s;
}

// Plane-parallel loops run over the 32*196 (plane, pixel) float4 items, 4 items per thread in flight.
constexpr int kEltItems = kKC * 196;

__device__ __forceinline__ void elt_task_body(const EltTask& t, const int tid, EltSmem& sm) {
  // plane range of this part (ATTEND, ATTEND_BWD, RELU_MASK, SCATTER, GATHER, DOTSIG_BWD may be split)
  const int n_parts = t.n_parts > 0 ? t.n_parts : 1;
  const int kc_lo = kKC * t.part / n_parts, kc_hi = kKC * (t.part + 1) / n_parts;
  const int items = (kc_hi - kc_lo) * 196;
  switch (t.op) {
    case OP_ATTEND: {  // o = a(feat) * b(map)
      for (int i0 = tid; i0 < items; i0 += 4 * 256) {
        float4 v[4];
        float m[4];
        int off[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int i = i0 + u * 256;
          if (i < items) {
            const int kc = kc_lo + i / 196, s = valid_slot16(i % 196);
            off[u] = (kc * 256 + s) * 4;
            v[u] = ld4(t.a + off[u]);
            m[u] = t.b[s];
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int i = i0 + u * 256;
          if (i < items) {
            float4 r = v[u];
            r.x = to_tf32(r.x * m[u]); r.y = to_tf32(r.y * m[u]); r.z = to_tf32(r.z * m[u]); r.w = to_tf32(r.w * m[u]);
            st4(t.o + off[u], r);
            if (t.flags & EF_HALF) st_half4(t.o, off[u] >> 10, (off[u] >> 2) & 255, r);
          }
        }
      }
    } break;

    case OP_ATTEND_BWD: {  // a = dX0, b = feat, c = map; o = dmap (+=), o2 = dfeat (write / +=)
      const bool acc = t.flags & EF_ACCUM;
      for (int i0 = tid; i0 < items; i0 += 4 * 256) {
        float4 g[4], old[4];
        float m[4];
        int off[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int i = i0 + u * 256;
          if (i < items) {
            const int kc = kc_lo + i / 196, s = valid_slot16(i % 196);
            off[u] = (kc * 256 + s) * 4;
            g[u] = ld4(t.a + off[u]);
            m[u] = t.c[s];
            if (acc) old[u] = ld4(t.o2 + off[u]);
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int i = i0 + u * 256;
          if (i < items) {
            float4 r = g[u];
            r.x *= m[u]; r.y *= m[u]; r.z *= m[u]; r.w *= m[u];
            if (acc) { r.x += old[u].x; r.y += old[u].y; r.z += old[u].z; r.w += old[u].w; }
            st4(t.o2 + off[u], r);
          }
        }
      }
      if (t.o != nullptr && tid < 196) {
        const int s = valid_slot16(tid);
        float acc0 = 0.f, acc1 = 0.f;
#pragma unroll 4
        for (int kc = kc_lo; kc < kc_hi; kc += 2) {
          const float4 g0 = ld4(t.a + (kc * 256 + s) * 4), f0 = ld4(t.b + (kc * 256 + s) * 4);
          const float4 g1 = ld4(t.a + ((kc + 1) * 256 + s) * 4), f1 = ld4(t.b + ((kc + 1) * 256 + s) * 4);
          acc0 += g0.x * f0.x + g0.y * f0.y + g0.z * f0.z + g0.w * f0.w;
          acc1 += g1.x * f1.x + g1.y * f1.y + g1.z * f1.z + g1.w * f1.w;
        }
        if (n_parts > 1) red_add_f32(t.o + s, acc0 + acc1);
        else t.o[s] += acc0 + acc1;
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
      if ((tid & 31) == 0) { sm.wv[tid >> 5] = bv; sm.wi[tid >> 5] = bi; }
      ELT_SYNC();
      if (tid == 0) {
        for (int i = 1; i < 8; ++i)
          if (sm.wv[i] > bv || (sm.wv[i] == bv && sm.wi[i] < bi)) { bv = sm.wv[i]; bi = sm.wi[i]; }
        sm.idx = valid_slot16(bi);
        if (t.idx) t.idx[0] = sm.idx;
      }
      ELT_SYNC();
      const int is = sm.idx;
      if (tid < 128) sm.sh[tid] = t.a[((tid >> 2) * 256 + is) * 4 + (tid & 3)] * t.w[tid];  // v_c * w_c
      ELT_SYNC();
      if (tid < 196) {
        const int s = valid_slot16(tid);
        float acc = 0.f;
#pragma unroll 8
        for (int kc = 0; kc < kKC; ++kc) {
          const float4 f = ld4(t.a + (kc * 256 + s) * 4);
          acc += f.x * sm.sh[kc * 4] + f.y * sm.sh[kc * 4 + 1] + f.z * sm.sh[kc * 4 + 2] + f.w * sm.sh[kc * 4 + 3];
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
      sm.sh[tid] = gp;
      const float sum_g = elt_block_sum(gp, sm.red, tid);
      const float sum_ga = elt_block_sum(tid < 196 ? gp * t.b[s] : 0.f, sm.red, tid);
      const float unscale = t.scale[1];
      if (tid == 0) { red_add_f32(t.dw2, sum_g * unscale); red_add_f32(t.dw + 128, sum_ga * unscale); }
      ELT_SYNC();
      {
        // thread = (plane kc, part): 4 channels, pixels part, part+8, ...
        // q_c = sum_p g[p]*feat[c][p];  dw_c += q_c*v_c;  dfeat[c][p] (+)= g[p]*w_c*v_c;  dfeat[c][is] += q_c*w_c
        const int kc = tid >> 3, part = tid & 7;
        const float4 v = ld4(t.a + (kc * 256 + is) * 4), wc = ld4(t.w + kc * 4);
        const bool acc = t.flags & EF_ACCUM;
        float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int p = part; p < 196; p += 8) {
          const int sl = valid_slot16(p);
          const float gi = sm.sh[p];
          const float4 f = ld4(t.a + (kc * 256 + sl) * 4);
          q.x = fmaf(gi, f.x, q.x); q.y = fmaf(gi, f.y, q.y); q.z = fmaf(gi, f.z, q.z); q.w = fmaf(gi, f.w, q.w);
          float4 d = make_float4(gi * wc.x * v.x, gi * wc.y * v.y, gi * wc.z * v.z, gi * wc.w * v.w);
          float* dp = t.o2 + (kc * 256 + sl) * 4;
          if (acc) { const float4 old = ld4(dp); d.x += old.x; d.y += old.y; d.z += old.z; d.w += old.w; }
          st4(dp, d);
        }
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
          q.x += __shfl_xor_sync(0xffffffffu, q.x, o); q.y += __shfl_xor_sync(0xffffffffu, q.y, o);
          q.z += __shfl_xor_sync(0xffffffffu, q.z, o); q.w += __shfl_xor_sync(0xffffffffu, q.w, o);
        }
        ELT_SYNC();  // all dfeat writes of slot `is` are done before the extra term is added
        if (part == 0) {
          red_add_f32(t.dw + kc * 4 + 0, q.x * v.x * unscale); red_add_f32(t.dw + kc * 4 + 1, q.y * v.y * unscale);
          red_add_f32(t.dw + kc * 4 + 2, q.z * v.z * unscale); red_add_f32(t.dw + kc * 4 + 3, q.w * v.w * unscale);
          float* dp = t.o2 + (kc * 256 + is) * 4;
          float4 d = ld4(dp);
          d.x += q.x * wc.x; d.y += q.y * wc.y; d.z += q.z * wc.z; d.w += q.w * wc.w;
          st4(dp, d);
        }
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
        for (int i = tid; i < kEltItems; i += 256) {
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
      sm.sh[tid] = gp;
      const float sum_g = elt_block_sum(gp, sm.red, tid);
      const float unscale = t.scale[1];
      if (tid == 0 && t.part == 0) red_add_f32(t.dw2, sum_g * unscale);
      ELT_SYNC();
      {
        // thread = (plane kc, lane-in-plane sub): pixels sub, sub+tpk, ... ; writes dZ and accumulates dw3 in one pass
        const int tpk = 256 / (kc_hi - kc_lo);  // threads per plane: 8 (whole tensor) .. 32 (a quarter)
        const int kc = kc_lo + tid / tpk, sub = tid % tpk;
        const float4 w = ld4(t.w + kc * 4);
        float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int p0 = sub; p0 < 196; p0 += 4 * tpk) {
          float4 y[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int p = p0 + tpk * u;
            if (p < 196) y[u] = ld4(t.a + (kc * 256 + valid_slot16(p)) * 4);
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int p = p0 + tpk * u;
            if (p < 196) {
              const int s = valid_slot16(p);
              const float gi = sm.sh[p];
              q.x = fmaf(gi, y[u].x, q.x); q.y = fmaf(gi, y[u].y, q.y); q.z = fmaf(gi, y[u].z, q.z); q.w = fmaf(gi, y[u].w, q.w);
              float4 d;
              d.x = y[u].x > 0.f ? to_tf32(gi * w.x) : 0.f; d.y = y[u].y > 0.f ? to_tf32(gi * w.y) : 0.f;
              d.z = y[u].z > 0.f ? to_tf32(gi * w.z) : 0.f; d.w = y[u].w > 0.f ? to_tf32(gi * w.w) : 0.f;
              st4(t.o + (kc * 256 + s) * 4, d);
              st_half4(t.o, kc, s, d);
            }
          }
        }
        for (int o = 1; o < tpk; o <<= 1) {
          q.x += __shfl_xor_sync(0xffffffffu, q.x, o); q.y += __shfl_xor_sync(0xffffffffu, q.y, o);
          q.z += __shfl_xor_sync(0xffffffffu, q.z, o); q.w += __shfl_xor_sync(0xffffffffu, q.w, o);
        }
        if (sub == 0) {
          red_add_f32(t.dw + kc * 4 + 0, q.x * unscale); red_add_f32(t.dw + kc * 4 + 1, q.y * unscale);
          red_add_f32(t.dw + kc * 4 + 2, q.z * unscale); red_add_f32(t.dw + kc * 4 + 3, q.w * unscale);
        }
      }
    } break;

    case OP_RELU_MASK: {  // o = a (dY) [+ c, the second d(feat) accumulator of a two-strand backward] masked by b (Y) > 0
      for (int i0 = tid; i0 < items; i0 += 4 * 256) {
        float4 g[4], y[4];
        int off[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int i = i0 + u * 256;
          if (i < items) {
            off[u] = ((kc_lo + i / 196) * 256 + valid_slot16(i % 196)) * 4;
            g[u] = ld4(t.a + off[u]);
            y[u] = ld4(t.b + off[u]);
            if (t.c) {
              const float4 g2 = ld4(t.c + off[u]);
              g[u].x += g2.x; g[u].y += g2.y; g[u].z += g2.z; g[u].w += g2.w;
            }
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int i = i0 + u * 256;
          if (i < items) {
            float4 r;
            r.x = y[u].x > 0.f ? to_tf32(g[u].x) : 0.f; r.y = y[u].y > 0.f ? to_tf32(g[u].y) : 0.f;
            r.z = y[u].z > 0.f ? to_tf32(g[u].z) : 0.f; r.w = y[u].w > 0.f ? to_tf32(g[u].w) : 0.f;
            st4(t.o + off[u], r);
            st_half4(t.o, off[u] >> 10, (off[u] >> 2) & 255, r);
          }
        }
      }
    } break;

    case OP_SCATTER: {  // a = NCHW [128][196] grad (times loss scale), b = Y planes (mask, optional) -> o planes
      const float sc = t.scale[0];
      for (int i = tid; i < items; i += 256) {
        const int kc = kc_lo + i / 196, p = i % 196, s = valid_slot16(p);
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
      for (int i = kc_lo * 4 * 196 + tid; i < kc_hi * 4 * 196; i += 256) {
        const int c = i / 196, p = i % 196;
        t.o[i] = t.a ? t.a[((c >> 2) * 256 + valid_slot16(p)) * 4 + (c & 3)] : 0.f;
      }
    } break;

    default: break;
  }
}

}  // namespace pnmn
