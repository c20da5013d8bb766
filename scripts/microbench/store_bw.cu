// Microbenchmark: cost of the executor's epilogue store pattern.  128 threads (4 warps) per CTA, each thread owns one
// pixel slot and writes, per "chunk", 8 x 16 B into fp32 planes (4 KB apart) and 4 x 16 B into fp16 half planes, exactly
// like exec.cu's epilogue.  Variants: CTAs per SM, how many SMs are active, fp32+fp16 vs fp16 only, default vs .cs/.wt stores,
// and a load variant (8 x LDG.128 per chunk from an L2-resident / DRAM-resident buffer).
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>

template <int MODE>  // 0: STG default fp32+fp16, 1: fp16 only, 2: st.global.cs, 3: loads (8 x LDG.128 per chunk) + 4 stores
__global__ void __launch_bounds__(128) epi(uint8_t* __restrict__ buf, size_t unit_bytes, int n_units, int n_tiles, long long* out) {
  const int tid = threadIdx.x;
  long long t0 = clock64();
  float acc = 0.f;
  for (int t = 0; t < n_tiles; ++t) {
    const size_t unit = (static_cast<size_t>(blockIdx.x) * 131 + t * 17) % n_units;
    uint8_t* base = buf + unit * unit_bytes;             // 32 fp32 planes x 256 slots x 16 B, then 16 half planes x 256 x 16 B
    uint8_t* hb = base + 32 * 256 * 16;
    const int so = (t & 1) * 128 + tid;
#pragma unroll 1
    for (int chunk = 0; chunk < 4; ++chunk) {
      float4 v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = make_float4(tid + j, chunk, t, acc);
      if (MODE == 3) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 a = *reinterpret_cast<const float4*>(base + (static_cast<size_t>(chunk * 8 + j) * 256 + so) * 16);
          v[j].x += a.x; v[j].y += a.y; acc += a.z;
        }
      }
      if (MODE == 0 || MODE == 2) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float4* p = reinterpret_cast<float4*>(base + (static_cast<size_t>(chunk * 8 + j) * 256 + so) * 16);
          if (MODE == 2) __stcs(p, v[j]); else *p = v[j];
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float4* p = reinterpret_cast<float4*>(hb + (static_cast<size_t>(chunk * 4 + j) * 256 + so) * 16);
        if (MODE == 2) __stcs(p, v[j]); else *p = v[j];
      }
    }
  }
  __threadfence();
  const long long t1 = clock64();
  if (tid == 0) out[blockIdx.x] = t1 - t0;
  if (acc == 12345.f) out[0] = 0;
}

template <int MODE>
void run(const char* name, uint8_t* g, size_t unit, int n_units, long long* d) {
  long long h[1024];
  for (int per_sm = 1; per_sm <= 4; per_sm *= 2)
    for (int sms : {8, 148}) {
      const int grid = sms * per_sm, n_tiles = 64;
      epi<MODE><<<grid, 128, per_sm == 1 ? 200 * 1024 : (per_sm == 2 ? 100 * 1024 : 50 * 1024)>>>(g, unit, n_units, n_tiles, d);
      epi<MODE><<<grid, 128, per_sm == 1 ? 200 * 1024 : (per_sm == 2 ? 100 * 1024 : 50 * 1024)>>>(g, unit, n_units, n_tiles, d);
      cudaError_t e = cudaDeviceSynchronize();
      cudaMemcpy(h, d, grid * 8, cudaMemcpyDeviceToHost);
      double mean = 0;
      for (int i = 0; i < grid; ++i) mean += h[i];
      mean /= grid;
      const double bytes = (MODE == 1 ? 32.0 : (MODE == 3 ? 32.0 + 64.0 : 96.0)) * 1024 * n_tiles;
      printf("%-34s ctas/SM %d SMs %3d: %7.0f clk per 128x128 tile, %6.1f B/clk/CTA, %6.1f B/clk/SM (%s)\n", name, per_sm, sms,
             mean / n_tiles, bytes / mean, bytes / mean * per_sm, cudaGetErrorString(e));
    }
}

int main() {
  const size_t unit = 32 * 256 * 16 + 16 * 256 * 16;  // 192 KB
  long long* d;
  cudaMalloc(&d, 8 * 1024);
  for (int big = 0; big < 2; ++big) {
    const int n_units = big ? 16384 : 256;  // 3 GB (DRAM-resident) vs 48 MB (L2-resident)
    uint8_t* g;
    if (cudaMalloc(&g, unit * n_units) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    cudaMemset(g, 0, unit * n_units);
    printf("---- buffer %.0f MB\n", unit * n_units / 1048576.0);
    cudaFuncSetAttribute(epi<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(epi<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(epi<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(epi<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    run<0>("fp32 planes + fp16 shadow (STG)", g, unit, n_units, d);
    run<1>("fp16 shadow only", g, unit, n_units, d);
    run<2>("fp32 + fp16, st.global.cs", g, unit, n_units, d);
    run<3>("8 LDG.128 + 4 STG.128 per chunk", g, unit, n_units, d);
    cudaFree(g);
  }
  return 0;
}
