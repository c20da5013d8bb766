// Microbenchmark: throughput of an L2-resident cp.async.bulk stream into a shared-memory ring as a function of the
// stage size, the ring depth, the number of CTAs per SM, the number of SMs streaming at once, and WHO issues the copies:
//   n_thr  warps, each with its own ring and barriers (lane 0 issues)                -> is the limit per thread or per CTA?
//   n_lane lanes of one warp, each issuing 1/n_lane of every stage on the same barrier -> can one warp pipeline several copies?
// Answers: is the executor's weight stream bound by latency x ring capacity, by L2 bandwidth, or by the issue rate?
#include <cstdio>
#include <cuda_runtime.h>
#include "tcgen05.cuh"
using namespace pnmn;

__global__ void __launch_bounds__(128) stream(const uint8_t* __restrict__ src, size_t window, int n_windows, int stage_bytes,
                                              int depth, int n_iter, long long* out, int n_thr, int n_lane) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full_all[128];
  if (threadIdx.x == 0) {
    for (int i = 0; i < 128; ++i) mbar_init(smem_u32(&full_all[i]), 1);
    fence_mbar_init();
  }
  fence_proxy_async();
  __syncthreads();
  const int k = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (k < n_thr) {
    uint64_t* full = full_all + 32 * k;
    uint8_t* ring = smem + k * depth * stage_bytes;
    const uint8_t* base = src + static_cast<size_t>((blockIdx.x * 7919u) % n_windows) * window;
    const int per_window = static_cast<int>(window / stage_bytes);
    const int piece = stage_bytes / n_lane;
    long long t0 = clock64();
    for (int it = 0; it < n_iter + depth; ++it) {
      const int s = it % depth;
      if (it >= depth && lane == 0) mbar_wait(smem_u32(&full[s]), ((it / depth) - 1) & 1);
      if (it == depth) t0 = clock64();  // ring primed
      if (it < n_iter) {
        if (lane == 0) mbar_arrive_expect_tx(smem_u32(&full[s]), stage_bytes);
        __syncwarp();
        if (lane < n_lane)
          bulk_g2s(smem_u32(ring) + s * stage_bytes + lane * piece,
                   base + static_cast<size_t>((it * n_thr + k) % per_window) * stage_bytes + lane * piece, piece, smem_u32(&full[s]));
      }
      __syncwarp();
    }
    const long long t1 = clock64();
    if (lane == 0) out[blockIdx.x * 4 + k] = t1 - t0;
  }
}

int main() {
  const size_t window = 288 * 1024;
  const int n_windows = 76;  // 22 MB, L2-resident like the packed weights
  uint8_t* g;
  cudaMalloc(&g, window * n_windows);
  cudaMemset(g, 1, window * n_windows);
  long long* d;
  cudaMalloc(&d, 8 * 2048);
  long long h[2048];
  cudaFuncSetAttribute(stream, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int cfgs[][2] = {{4096, 4}, {12288, 2}, {12288, 4}, {24576, 2}, {24576, 4}, {49152, 2}};
  const int who[][2] = {{1, 1}, {2, 1}, {4, 1}, {1, 2}, {1, 3}, {1, 6}};
  for (auto& w : who)
    for (int per_sm = 1; per_sm <= 2; ++per_sm)
      for (int grid_sms : {8, 148})
        for (auto& c : cfgs) {
          const int n_thr = w[0], n_lane = w[1];
          const int stage = c[0], depth = c[1];
          const int smem = per_sm == 1 ? 200 * 1024 : 100 * 1024;
          if (stage * depth * n_thr > smem || stage % (n_lane * 16) != 0) continue;
          if ((n_thr > 1 || n_lane > 1) && per_sm == 1) continue;
          const int n_iter = static_cast<int>(8 * window / stage);
          const int grid = grid_sms * per_sm;
          for (int rep = 0; rep < 2; ++rep) stream<<<grid, 128, smem>>>(g, window, n_windows, stage, depth, n_iter, d, n_thr, n_lane);
          cudaError_t e = cudaDeviceSynchronize();
          cudaMemcpy(h, d, grid * 4 * 8, cudaMemcpyDeviceToHost);
          double mean = 0;
          for (int i = 0; i < grid; ++i)
            for (int k = 0; k < n_thr; ++k) mean += h[i * 4 + k];
          mean /= grid * n_thr;
          const double bytes = static_cast<double>(n_iter - depth) * stage;
          const double bpc = bytes / mean * n_thr;
          printf("warps %d lanes %d ctas/SM %d SMs %3d stage %5d B x depth %d (%5.1f KB in flight): %6.1f B/clk/CTA %6.1f B/clk/SM chip %6.0f B/clk, "
                 "%5.0f clk per stage per issuer (%s)\n",
                 n_thr, n_lane, per_sm, grid_sms, stage, depth, n_thr * stage * depth / 1024.0, bpc, bpc * per_sm, bpc * grid,
                 mean / (n_iter - depth), cudaGetErrorString(e));
        }
  return 0;
}
