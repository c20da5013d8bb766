// Host-visible declarations of the layout / packing kernels (layout.cu) and kernel launchers.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include "executor.h"
#include "elt.h"

namespace pnmn {

struct PackTask {
  int64_t src_off;   // offset (floats) of the reference-layout weight inside the flat parameter buffer
  int64_t dst_off;   // offset (halves) inside the packed-weight buffer
  int first_tile;    // global index of this task's first 16x128 tile
  int n_kb, ntaps, flip;
  int k_off, n_off;
  int k_stride, n_stride, tap_stride;
  int pad_;
};

struct BiasGradTaskH {  // mirrors BiasGradTask in wgrad.cu
  const WgradInst* inst;
  int n_inst;
  int P;
  float* db;
  const float* scale;
};

cudaError_t launch_pack(const PackTask* d_tasks, int n_tasks, int total_tiles, const float* params,
                        void* packed, cudaStream_t stream);
cudaError_t launch_nchw_to_planes(const void* src, int src_is_half, float* dst, int B, int C, const int64_t* dst_off,
                                  int64_t guard_floats, int64_t unit_floats, cudaStream_t stream);
cudaError_t launch_round_features_f16(const float* src, void* dst, int64_t n, cudaStream_t stream);
cudaError_t launch_conv_simt(const ConvTask* d_tasks, int n_tasks, const ConvCfg* d_cfgs, cudaStream_t stream);
// d_counter: zeroed device int (tasks are then pulled by one persistent CTA per SM), or nullptr (one CTA per task)
cudaError_t launch_wgrad(const WgradTask* d_tasks, int n_tasks, int impl_simt, int* d_counter, cudaStream_t stream);
cudaError_t launch_bias_grad(const void* d_tasks, int n_tasks, int split, cudaStream_t stream);
cudaError_t launch_elt(const EltTask* d_tasks, int n_tasks, cudaStream_t stream);
cudaError_t launch_loss_scale(const float* grad, size_t n, float* scale, cudaStream_t stream);

// classifier: ReLU + 2x2/2 max-pool + (C,7,7) flatten of a channels-last [B][14*14][C] tensor, and its backward (layout.cu)
cudaError_t launch_relu_pool_fwd(const float* y, const float* bias, float* pooled, uint8_t* code, int B, int C, cudaStream_t st);
cudaError_t launch_relu_pool_bwd(const float* g, const uint8_t* code, float* gy, int B, int C, cudaStream_t st);

}  // namespace pnmn
