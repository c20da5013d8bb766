// Weight gradient of the shift-GEMM convolutions:
//
//   dW[cout][cin][ty][tx] += sum_{instances} sum_{slot} dZ[cout][slot] * X[cin][slot + shift(ty,tx)]
//
// "Instance" = one (sample, step) at which this conv ran in the forward pass; a module that was
// used by 37 samples of the batch contributes 37 instances to the same dW (reference: autograd
// accumulating into one nn.Conv2d.weight.grad, probnmn/models/nmn.py:114-115,229).
//
// GEMM view per tap: M = cout (128, from dZ), N = cin (128, from X), K = pixel slots.  Both
// operands are "MN-major" (channels contiguous per slot).  tcgen05 has no unswizzled MN-major
// layout for tf32 (measured: such an MMA returns zeros), so the operands are the fp16 shadow
// copies ("half planes", 8 channels = 16 B per slot; executor.h) and the MMA is kind::f16 with
// fp32 accumulation: same 10-bit mantissa as tf32, twice the tensor rate, half the bytes.  A tap is
// again just a different start address of the X descriptor.
// One CTA owns one tap ROW (3 taps -> 3 accumulators of 128 TMEM columns) of one weight tensor
// and walks a list of instances, streaming 128-slot chunks of dZ and X through an mbarrier ring.
#include <cuda_fp16.h>

#include "executor.h"
#include "tcgen05.cuh"

namespace pnmn {

constexpr int kWgThreads = 256;
constexpr int kWgChunk = 128;       // slots (K) per stage: 2 KB per half plane and copy (the copy engine is issue-bound on
                                    // small copies: 1 KB pieces ran at ~10 GB/s per SM, ncu profiles/r1)
constexpr int kWgStages = 3;
constexpr int kWgHeader = 1024;
constexpr int kWgSmemTotal = 227 * 1024;
constexpr int kHP = kC / 8;              // fp16 half planes per 128-channel tensor
constexpr int kWgMaxXS = kWgChunk + 16;  // chunk + 2*dil halo, dil <= 8
constexpr int kWgStageBytes = kHP * kWgChunk * 16 + kHP * kWgMaxXS * 16;  // 16 KB + 20 KB
static_assert(kWgHeader + kWgStages * kWgStageBytes <= kWgSmemTotal, "wgrad smem");

struct WgSmemHeader {
  uint64_t full[kWgStages];
  uint64_t empty[kWgStages];
  uint64_t tmem_full;
  uint32_t tmem_base;
  int task_idx[2];   // double-buffered: a slot is rewritten two barriers after its last reader
};

__device__ __forceinline__ int wg_smid() {
  int v;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(v));
  return v;
}

// counter == nullptr: CTA i runs task i (bring-up entry point).  Otherwise every CTA PULLS tasks from *counter until the list
// is exhausted, and CTAs that land on an SM >= sm_limit leave at once: the kernel keeps off the SMs reserved for other
// streams (pnmn_set_reserved_sms) exactly like the executor.  The grid still has one CTA per task: a CTA needs a whole SM
// (227 KB of shared memory, all 512 TMEM columns), and when kernels of other streams come and go on the device only a
// steady supply of pending CTAs gets hold of every SM that falls free (one persistent CTA per SM, placed once at launch,
// ended up on a fraction of the SMs); CTAs that start after the list has run dry return before setting anything up.
__global__ void __launch_bounds__(kWgThreads, 1)
wgrad_tc_kernel(const WgradTask* __restrict__ tasks, int n_tasks, int* __restrict__ counter, int sm_limit) {
  if (counter && wg_smid() >= sm_limit) return;
  extern __shared__ __align__(1024) uint8_t smem[];
  WgSmemHeader* hdr = reinterpret_cast<WgSmemHeader*>(smem);
  uint8_t* ring = smem + kWgHeader;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  int task_idx = static_cast<int>(blockIdx.x);
  if (counter) {
    if (threadIdx.x == 0) hdr->task_idx[0] = atomicAdd(counter, 1);
    __syncthreads();
    task_idx = hdr->task_idx[0];
    if (task_idx >= n_tasks) return;
  }

  if (threadIdx.x == 0) {
    for (int i = 0; i < kWgStages; ++i) {
      mbar_init(smem_u32(&hdr->full[i]), 1);
      mbar_init(smem_u32(&hdr->empty[i]), 1);
    }
    mbar_init(smem_u32(&hdr->tmem_full), 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<512>(smem_u32(&hdr->tmem_base));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = hdr->tmem_base;

  // ring iterations / tasks this CTA has completed so far (the same in every role): barrier phases carry over
  uint32_t git = 0, n_done = 0;
  while (task_idx < n_tasks) {
  const WgradTask t = tasks[task_idx];

  const int halo = t.ntaps_x == 3 ? t.dil : 0;
  const int xs = kWgChunk + 2 * halo;                 // X slots per half plane per stage
  const int n_valid = kHW * t.S;                      // slots that can carry a non-zero dZ
  const int n_chunks = (n_valid + kWgChunk - 1) / kWgChunk;
  const int row_shift = t.ntaps_x == 3 ? (t.tap_row - 1) * t.dil * t.S : 0;
  const uint32_t dz_bytes = kHP * kWgChunk * 16;
  const int total = t.n_inst * n_chunks;

  if (warp == 0) {
    // producer: lane = half plane index (lanes 16..31 idle).  The instance table is read 32 entries at a time into
    // registers (one entry per lane) and broadcast with shuffles: a dependent global load per ring stage would put an
    // L2 round trip on the issue path of every stage.
    unsigned long long my_dz = 0, my_x = 0;
    for (int it = 0; it < total; ++it) {
      const int st = (git + it) % kWgStages;
      const uint32_t ph = ((git + it) / kWgStages) & 1;
      const int inst = it / n_chunks, ch = it % n_chunks;
      if (ch == 0 && (inst & 31) == 0) {
        const int mine = inst + lane;
        if (mine < t.n_inst) {
          const WgradInst wi = t.inst[mine];
          my_dz = reinterpret_cast<unsigned long long>(wi.dz);
          my_x = reinterpret_cast<unsigned long long>(wi.x);
        }
      }
      const uint8_t* gdz = reinterpret_cast<const uint8_t*>(__shfl_sync(0xffffffffu, my_dz, inst & 31));
      const uint8_t* gx = reinterpret_cast<const uint8_t*>(__shfl_sync(0xffffffffu, my_x, inst & 31));
      const uint32_t bar = smem_u32(&hdr->full[st]);
      if (lane == 0) {
        mbar_wait(smem_u32(&hdr->empty[st]), ph ^ 1);
        mbar_arrive_expect_tx(bar, dz_bytes + kHP * xs * 16);
      }
      __syncwarp();
      {
        // lanes 0..15 fetch the dZ half planes, lanes 16..31 the X half planes
        const int c0 = ch * kWgChunk, hp = lane & (kHP - 1);
        uint8_t* sdz = ring + st * kWgStageBytes;
        uint8_t* sx = sdz + dz_bytes;
        if (lane < kHP)
          bulk_g2s(smem_u32(sdz + hp * kWgChunk * 16), gdz + (static_cast<ptrdiff_t>(hp) * t.P + c0) * 16, kWgChunk * 16, bar);
        else
          bulk_g2s(smem_u32(sx + hp * xs * 16), gx + (static_cast<ptrdiff_t>(hp) * t.P + c0 + row_shift - halo) * 16, xs * 16,
                   bar);
      }
    }
  } else if (warp == 2) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_f16(128, 128, 1, 1);
      // MN-major, no swizzle: LBO = next 8-slot K group (128 B), SBO = next half plane.  Descriptor updates
      // are one 32-bit add on the start-address field (16-byte units = slots).
      const uint32_t a_hi = static_cast<uint32_t>(make_smem_desc(0, 0, kWgChunk * 16u) >> 32);
      const uint32_t b_hi = static_cast<uint32_t>(make_smem_desc(0, 0, static_cast<uint32_t>(xs) * 16u) >> 32);
      const uint32_t lbo = (128u >> 4) << 16;
      const uint32_t ring_lo = smem_u32(ring) >> 4;
      for (int it = 0; it < total; ++it) {
        const int st = (git + it) % kWgStages;
        mbar_wait(smem_u32(&hdr->full[st]), ((git + it) / kWgStages) & 1);
        tc_fence_after();
        const uint32_t a_lo0 = ring_lo + st * (kWgStageBytes >> 4) + lbo;
        const uint32_t b_lo0 = a_lo0 + (dz_bytes >> 4);
        const int c0 = (it % n_chunks) * kWgChunk;
        int nk = (n_valid - c0 + 15) / 16;  // 16-slot MMA steps that can still see a non-zero dZ
        nk = nk > kWgChunk / 16 ? kWgChunk / 16 : nk;
        for (int tx = 0; tx < t.ntaps_x; ++tx) {
          const uint32_t b_lo = b_lo0 + static_cast<uint32_t>(halo * tx);
          const uint32_t d = tmem_base + tx * 128;
#pragma unroll
          for (int k16 = 0; k16 < kWgChunk / 16; ++k16) {
            if (k16 < nk)
              umma_f16(d, (static_cast<uint64_t>(a_hi) << 32) | (a_lo0 + k16 * 16),
                       (static_cast<uint64_t>(b_hi) << 32) | (b_lo + k16 * 16), idesc, (it | k16) != 0);
          }
        }
        umma_commit(smem_u32(&hdr->empty[st]));
      }
      umma_commit(smem_u32(&hdr->tmem_full));
    }
  } else if (warp >= 4) {
    const int q = warp & 3;
    const int cout = q * 32 + lane;
    mbar_wait(smem_u32(&hdr->tmem_full), n_done & 1);
    tc_fence_after();
    const int kk = t.ksize * t.ksize;
    const float unscale = t.scale ? __ldg(t.scale + 1) : 1.f;
    for (int tx = 0; tx < t.ntaps_x; ++tx) {
      const int tap = t.ksize == 3 ? t.tap_row * 3 + tx : 0;
      for (int chunk = 0; chunk < 4; ++chunk) {
        uint32_t v[32];
        tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + tx * 128 + chunk * 32, v);
        tmem_ld_wait();
        if (total > 0) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int cin = t.cin0 + chunk * 32 + j;
            red_add_f32(t.dw + (static_cast<size_t>(cout) * t.cin_total + cin) * kk + tap,
                      __uint_as_float(v[j]) * unscale);
          }
        }
      }
    }
  }
  git += static_cast<uint32_t>(total);
  ++n_done;
  if (!counter) break;
  // (thread 0 may be a whole -- short -- task ahead of a thread that has not read the previous index yet: two slots)
  if (threadIdx.x == 0) hdr->task_idx[n_done & 1] = atomicAdd(counter, 1);
  // the epilogue has read the accumulators before the next task's first MMA overwrites them
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  task_idx = hdr->task_idx[n_done & 1];
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<512>(tmem_base);
}

#ifdef PNMN_BRINGUP
// CUDA-core twin with identical task semantics (bring-up / debugging only, PNMN_CONV_IMPL=simt).
__global__ void __launch_bounds__(256) wgrad_simt_kernel(const WgradTask* __restrict__ tasks) {
  const WgradTask t = tasks[blockIdx.x];
  const int kk = t.ksize * t.ksize;
  const int nslots = kHW * t.S;
  const float unscale = t.scale ? t.scale[1] : 1.f;
  for (int idx = threadIdx.x; idx < t.ntaps_x * 128 * 128; idx += blockDim.x) {
    const int cin = idx % 128, cout = (idx / 128) % 128, tx = idx / (128 * 128);
    const int shift = t.ntaps_x == 3 ? ((t.tap_row - 1) * t.S + (tx - 1)) * t.dil : 0;
    float acc = 0.f;
    for (int i = 0; i < t.n_inst; ++i) {
      const WgradInst wi = t.inst[i];
      const __half* dz = static_cast<const __half*>(wi.dz) + static_cast<size_t>(cout / 8) * t.P * 8 + (cout % 8);
      const __half* x = static_cast<const __half*>(wi.x) + static_cast<ptrdiff_t>(cin / 8) * t.P * 8 + (cin % 8);
      for (int p = 0; p < nslots; ++p) {
        const float g = __half2float(dz[p * 8]);
        if (g != 0.f) acc = fmaf(g, __half2float(x[static_cast<ptrdiff_t>(p + shift) * 8]), acc);
      }
    }
    const int tap = t.ksize == 3 ? t.tap_row * 3 + tx : 0;
    if (t.n_inst > 0)
      red_add_f32(t.dw + (static_cast<size_t>(cout) * t.cin_total + t.cin0 + cin) * kk + tap, acc * unscale);
  }
}

#endif

// db[n] += sum over instances and slots of dZ[n][slot]      (one CTA per weight tensor)
struct BiasGradTask {
  const WgradInst* inst;
  int n_inst;
  int P;
  float* db;
  const float* scale;
};
__global__ void __launch_bounds__(128) bias_grad_kernel(const BiasGradTask* __restrict__ tasks) {
  // thread = (half plane hp = tid/8, slot lane = tid%8): 16-byte loads, 8 adjacent threads read 128 contiguous bytes
  const BiasGradTask t = tasks[blockIdx.x];
  const int hp = threadIdx.x >> 3, sl = threadIdx.x & 7;
  const int i0 = blockIdx.y, di = gridDim.y;
  if (i0 >= t.n_inst) return;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int i = i0; i < t.n_inst; i += di) {
    const uint4* dz = reinterpret_cast<const uint4*>(t.inst[i].dz) + static_cast<size_t>(hp) * t.P;
    for (int p = sl; p < t.P; p += 32) {
      uint4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = (p + 8 * u < t.P) ? dz[p + 8 * u] : make_uint4(0, 0, 0, 0);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const __half2* h = reinterpret_cast<const __half2*>(&v[u]);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = __half22float2(h[e]);
          acc[2 * e] += f.x; acc[2 * e + 1] += f.y;
        }
      }
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], 1);
    acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], 2);
    acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], 4);
  }
  if (sl == 0) {
    const float unscale = t.scale ? t.scale[1] : 1.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) red_add_f32(t.db + hp * 8 + e, acc[e] * unscale);
  }
}

int reserved_sms();   // exec.cu

cudaError_t launch_wgrad(const WgradTask* d_tasks, int n_tasks, int impl_simt, int* d_counter, cudaStream_t stream) {
  if (n_tasks <= 0) return cudaSuccess;
  if (impl_simt) {
#ifdef PNMN_BRINGUP
    wgrad_simt_kernel<<<n_tasks, 256, 0, stream>>>(d_tasks);
    return cudaGetLastError();
#else
    return cudaErrorNotSupported;   // built without the bring-up kernels (make BRINGUP=1)
#endif
  }
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgSmemTotal);
    if (e != cudaSuccess) return e;
    attr_done = true;
  }
  if (!d_counter) {
    wgrad_tc_kernel<<<n_tasks, kWgThreads, kWgSmemTotal, stream>>>(d_tasks, n_tasks, nullptr, 0);
    return cudaGetLastError();
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int r = reserved_sms();
  const int sm_limit = sms - (r < sms * 3 / 4 ? r : sms * 3 / 4);
  wgrad_tc_kernel<<<n_tasks, kWgThreads, kWgSmemTotal, stream>>>(d_tasks, n_tasks, d_counter, sm_limit);
  return cudaGetLastError();
}

cudaError_t launch_bias_grad(const void* d_tasks, int n_tasks, int split, cudaStream_t stream) {
  if (n_tasks <= 0) return cudaSuccess;
  bias_grad_kernel<<<dim3(n_tasks, split), 128, 0, stream>>>(static_cast<const BiasGradTask*>(d_tasks));
  return cudaGetLastError();
}

}  // namespace pnmn
