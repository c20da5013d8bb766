// Microbenchmark: what does the per-stage protocol of the executor's mainloop cost?
//   V0 tight MMA loop | V1 + tcgen05.commit every G MMAs | V2 + tcgen05.fence::after_thread_sync
//   V3 + mbarrier try_wait on a completed phase | V4 = V1 + another warp streaming cp.async.bulk into smem
#include <cstdio>
#include <cuda_runtime.h>
#include "tcgen05.cuh"
using namespace pnmn;

template <int V, int G>
__global__ void __launch_bounds__(128, 1) bench(long long* out, int n_stage, const uint8_t* gsrc) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bars[8], done_bar, cbar[4];
  __shared__ uint32_t tmem_ptr;
  __shared__ volatile int stop;
  __shared__ int flag;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) mbar_init(smem_u32(&bars[i]), 1);
    for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&cbar[i]), 1);
    mbar_init(smem_u32(&done_bar), 1);
    stop = 0; flag = 1;
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<512>(smem_u32(&tmem_ptr));
  fence_proxy_async();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = tmem_ptr;
  if (threadIdx.x == 32) {
    const uint32_t a0 = smem_u32(smem) >> 4, b0 = (smem_u32(smem) + 64 * 1024) >> 4;
    const uint32_t idesc = make_idesc_f16(128, 128, 0, 0);
    const uint64_t hi = make_smem_desc(0, 0, 128) & 0xFFFFFFFF00000000ull;
    const uint32_t lbo = 256u << 16;
    // a completed phase to wait on (V3)
    mbar_arrive(smem_u32(&done_bar));
    const long long t0 = clock64();
    for (int st = 0; st < n_stage; ++st) {
      if (V == 3 || V == 4) mbar_wait(smem_u32(&done_bar), 0);
      if (V >= 5) { while (*reinterpret_cast<volatile int*>(&flag) < 1) {} }
      if (V >= 2) tc_fence_after();
#pragma unroll
      for (int i = 0; i < G; ++i)
        umma_f16(tmem + (i & 3) * 128, hi | (a0 + lbo + (i & 7) * 16), hi | (b0 + lbo + (st & 3) * 256), idesc, 1);
      if (V >= 1) umma_commit(smem_u32(&bars[st & 7]));
    }
    umma_commit(smem_u32(&cbar[0]));
    const long long t1 = clock64();
    mbar_wait(smem_u32(&cbar[0]), 0);
    const long long t2 = clock64();
    stop = 1;
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  } else if (V == 6 && threadIdx.x == 96) {
    // a watcher thread hammering mbarrier waits on a completed phase while the issuer polls a plain shared flag
    while (!stop) mbar_wait(smem_u32(&done_bar), 0);
  } else if (V == 4 && threadIdx.x == 64) {
    // stream 12 KB bulk copies into the upper smem region as fast as one barrier round trip allows
    uint32_t ph = 0;
    int k = 0;
    while (!stop) {
      const uint32_t bar = smem_u32(&cbar[1 + (k & 1)]);
      mbar_arrive_expect_tx(bar, 12288);
      bulk_g2s(smem_u32(smem) + 128 * 1024 + (k & 3) * 12288, gsrc + (size_t)((k * 148 + blockIdx.x) & 4095) * 12288, 12288, bar);
      if (k & 1) { mbar_wait(smem_u32(&cbar[1]), ph); mbar_wait(smem_u32(&cbar[2]), ph); ph ^= 1; }
      ++k;
    }
    if (k & 1) mbar_wait(smem_u32(&cbar[1]), ph);
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

template <int V, int G>
void run(const char* name, long long* d, const uint8_t* g) {
  cudaFuncSetAttribute(bench<V, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int n_stage = 1024;
  bench<V, G><<<148, 128, 200 * 1024>>>(d, n_stage, g);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("%-46s G=%2d: issue %.1f cyc/mma, complete %.1f cyc/mma, %.0f cyc/stage (%s)\n", name, G, double(h[0]) / (n_stage * G),
         double(h[1]) / (n_stage * G), double(h[1]) / n_stage, cudaGetErrorString(e));
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  uint8_t* g; cudaMalloc(&g, (size_t)4096 * 12288); cudaMemset(g, 0, (size_t)4096 * 12288);
  run<0, 3>("V0 tight loop", d, g);               run<0, 12>("V0 tight loop", d, g);
  run<1, 3>("V1 + commit per stage", d, g);       run<1, 12>("V1 + commit per stage", d, g);
  run<2, 3>("V2 + fence::after_thread_sync", d, g); run<2, 12>("V2 + fence::after_thread_sync", d, g);
  run<3, 3>("V3 + mbarrier wait (completed phase)", d, g); run<3, 12>("V3 + mbarrier wait (completed phase)", d, g);
  run<4, 3>("V4 = V1 + concurrent cp.async.bulk stream", d, g); run<4, 12>("V4 = V1 + concurrent cp.async.bulk stream", d, g);
  run<5, 3>("V5 = V2 + poll a volatile shared flag", d, g); run<5, 1>("V5 = V2 + poll a volatile shared flag", d, g);
  run<6, 3>("V6 = V5 + a watcher warp doing mbarrier waits", d, g); run<6, 1>("V6 = V5 + a watcher warp doing mbarrier waits", d, g);
  run<3, 1>("V3 + mbarrier wait (completed phase)", d, g); run<3, 6>("V3 + mbarrier wait (completed phase)", d, g);
  return 0;
}
