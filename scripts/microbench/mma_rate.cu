// Microbenchmark: issue rate of tcgen05.mma with the executor's unswizzled smem layouts (no loads).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../probnmn_clevr_b200/csrc mma_rate.cu -o mma_rate
#include <cstdio>
#include <cuda_runtime.h>
#include "tcgen05.cuh"
using namespace pnmn;

// mode 0: tf32 K-major (A rows = slots, LBO = 4 KB plane)   mode 1: f16 K-major (same geometry, K=16)
// mode 2: f16 MN-major (wgrad geometry)
template <int MODE, int NACC>
__global__ void __launch_bounds__(128, 1) mma_rate(long long* out, int n_mma) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc<512>(smem_u32(&tmem_ptr));
  fence_proxy_async();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = tmem_ptr;
  if (threadIdx.x == 32) {
    const uint32_t a0 = smem_u32(smem) >> 4, b0 = (smem_u32(smem) + 96 * 1024) >> 4;
    uint32_t idesc, a_hi, b_hi, lbo;
    if (MODE == 0) { idesc = make_idesc_tf32(128, 128, 0, 0); a_hi = b_hi = uint32_t(make_smem_desc(0, 0, 128) >> 32); lbo = 256u << 16; }
    else if (MODE == 1) { idesc = make_idesc_f16(128, 128, 0, 0); a_hi = b_hi = uint32_t(make_smem_desc(0, 0, 128) >> 32); lbo = 256u << 16; }
    else { idesc = make_idesc_f16(128, 128, 1, 1); a_hi = b_hi = uint32_t(make_smem_desc(0, 0, 1024) >> 32); lbo = 8u << 16; }
    const long long t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < n_mma; ++i) {
      const uint32_t d = tmem + (i & (NACC - 1)) * 128;
      const uint32_t a_lo = a0 + lbo + (i & 7) * 16, b_lo = b0 + lbo + (i & 3) * 128;
      if (MODE == 0) umma_tf32(d, (uint64_t(a_hi) << 32) | a_lo, (uint64_t(b_hi) << 32) | b_lo, idesc, 1);
      else umma_f16(d, (uint64_t(a_hi) << 32) | a_lo, (uint64_t(b_hi) << 32) | b_lo, idesc, 1);
    }
    umma_commit(smem_u32(&bar));
    const long long t1 = clock64();
    mbar_wait(smem_u32(&bar), 0);
    const long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

template <int MODE, int NACC>
void run1(const char* name, int grid, long long* d) {
  const int n_acc = NACC;
  {
    cudaFuncSetAttribute(mma_rate<MODE, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int n = 4096;
    mma_rate<MODE, NACC><<<grid, 128, 200 * 1024>>>(d, n);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("%-28s grid %3d acc %d: issue %.1f cyc/mma, complete %.1f cyc/mma (%s)\n", name, grid, n_acc, double(h[0]) / n, double(h[1]) / n, cudaGetErrorString(e));
  }
}
template <int MODE>
void run(const char* name, int grid) {
  long long* d; cudaMalloc(&d, 16);
  run1<MODE, 1>(name, grid, d);
  run1<MODE, 4>(name, grid, d);
  cudaFree(d);
}

int main() {
  for (int grid : {1, 148}) {
    run<0>("tf32 K-major M128 N128 K8", grid);
    run<1>("f16  K-major M128 N128 K16", grid);
    run<2>("f16  MN-major M128 N128 K16", grid);
  }
  return 0;
}
