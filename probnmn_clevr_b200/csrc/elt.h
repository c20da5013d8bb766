// Task record of the CUDA-core (non-convolution) executor ops.  See elementwise.cu.
#pragma once
#include "executor.h"

namespace pnmn {

enum EltOp : int {
  OP_ATTEND = 1,
  OP_ATTEND_BWD,
  OP_SAME,
  OP_SAME_BWD,
  OP_MINMAX,
  OP_MINMAX_BWD,
  OP_DOTSIG_BWD,
  OP_RELU_MASK,
  OP_SCATTER,
  OP_GATHER,
};

enum EltFlags : int {
  EF_ACCUM = 1,    // plane-typed output `o`/`o2` (first) is accumulated instead of written
  EF_ACCUM2 = 2,   // same for `o2` of OP_MINMAX_BWD
  EF_MAX = 4,      // OP_MINMAX*: max instead of min
  EF_A_MAP = 8,    // operand a is a 1-channel map (broadcast over channels)
  EF_B_MAP = 16,   // operand b is a 1-channel map
  EF_HALF = 32,    // also write the fp16 shadow behind the plane-typed output `o`
};

struct EltTask {
  int op;
  int flags;
  const float* a;
  const float* b;
  const float* c;
  const float* g;
  float* o;
  float* o2;
  const float* w;
  float* dw;
  float* dw2;
  int* idx;
  const float* scale;  // {loss scale, 1/loss scale} of this backward pass (backward ops only)
  int part, n_parts;   // plane-parallel ops are split over n_parts CTAs (planes [32*part/n, 32*(part+1)/n)); 0 = 1
  int64_t pad_[3];
};
static_assert(sizeof(EltTask) == 128, "EltTask must stay 128 bytes");

}  // namespace pnmn
