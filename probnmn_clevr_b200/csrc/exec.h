// Task metadata of the persistent executor kernel (exec.cu).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include "executor.h"

namespace pnmn {

constexpr int kMaxDeps = 10;  // two paired samples x up to four parts (split elementwise op / M tiles) of the previous stage

constexpr int kTraceW = 32;  // int64 entries per task in the optional trace buffer (exec.cu)

enum TaskType : int { TASK_CONV = 0, TASK_ELT = 1 };

struct TaskMeta {
  int type;
  int n_deps;
  int deps[kMaxDeps];  // indices of the tasks that must have published `done` first, -1 = unused
};
static_assert(sizeof(TaskMeta) == 48, "TaskMeta must stay 48 bytes");

// max_ctas > 0 caps the persistent grid (several task lists running side by side share the SMs' CTA slots)
cudaError_t launch_exec(const uint8_t* d_tasks, const TaskMeta* d_meta, int n_tasks, const ConvCfg* d_cfgs,
                        int* d_counter, int* d_done, long long* d_trace, int max_ctas, cudaStream_t stream);

}  // namespace pnmn
