// Internal types shared by the host-side program compiler / scheduler and the CUDA kernels of
// the NMN module executor.  Nothing here is part of the public C ABI (see include/pnmn.h).
//
// Activation format ("planes"): a C-channel 14x14 map is stored as C/4 planes; plane kc holds
// channels 4kc..4kc+3 for every pixel slot, 16 bytes per slot:   buf[kc][slot][4] (fp32).
// A slot is a position of a zero-padded SxR grid, slot = y*S + x; only y,x < 14 carry data,
// every other slot is a permanent zero (written once when the arena is created).  The zero
// slots are what makes a 3x3 (dilated) convolution a plain *row shift* of the MMA A operand:
// tap (ty,tx) of dilation d reads slot + (ty*S + tx)*d, and every out-of-image read lands on a
// zero slot of the same plane, of the previous plane, or of the staging buffer's lead gap.
//   P16: S=16, 256 slots  (dilation 1, 2)      P18: S=18, 324 slots (dilation 4)
//   P22: S=22, 484 slots  (dilation 8)
//
// fp16 shadow ("half planes"): a buffer that takes part in a weight gradient (conv inputs X and
// pre-activation gradients dZ) is followed in memory by an fp16 copy hbuf[c/8][slot][8] (16 bytes
// per slot again, so taps stay row shifts).  The wgrad GEMM contracts over SLOTS, which makes both
// operands "MN-major"; tcgen05 offers no unswizzled MN-major layout for tf32, but it does for
// 16-bit types, and fp16 has the same 10-bit mantissa as tf32.  Gradients are kept in fp16 range by
// one power-of-two loss scale per backward pass (see scale_kernel in elementwise.cu).
#pragma once
#include <cstdint>

namespace pnmn {

constexpr int kHW = 14;          // spatial size of the ResNet-101 stage-3 feature map
constexpr int kC = 128;          // module channels
constexpr int kKC = kC / 4;      // planes per 128-channel activation
constexpr int NSMAX = 2;         // samples per conv CTA (share one weight stream)

struct PlaneFmt {
  int S;  // row stride in slots
  int P;  // slots per plane
};
constexpr PlaneFmt kP16{16, 256};
constexpr PlaneFmt kP18{18, 324};
constexpr PlaneFmt kP22{22, 484};

inline PlaneFmt fmt_for_dilation(int d) {
  return d <= 2 ? PlaneFmt{16, 256} : (d == 4 ? PlaneFmt{18, 324} : PlaneFmt{22, 484});
}

// ---- conv (forward / dgrad) --------------------------------------------------------------------
enum ConvFlags : int {
  F_BIAS = 1,     // y = acc + bias[n]
  F_RELU = 2,     // y = max(y, 0)
  F_STORE = 4,    // write y (tf32-rounded) to out[]
  F_DOTSIG = 8,   // map_out[slot] = sigmoid(sum_n y[n]*w3[n] + b3)   (fused 1x1 conv + sigmoid)
  F_MASK = 16,    // y = aux[n] > 0 ? y : 0                            (ReLU backward)
  F_ACCUM = 32,   // y += out[] (existing contents)
  F_HALF = 64,    // also write the fp16 shadow copy behind out[]
  // persistent executor only (exec.cu) -- the elementwise op that would follow is folded into the epilogue:
  F_ATTEND = 128, // with F_DOTSIG: also write the NEXT module's attended input  x0 = feat * map  (fp16 shadow only);
                  // aux[s] = feat (P16 fp32 planes), in[1][s] = shadow of x0          (nmn_modules.py:83,120,161)
  F_ATTBWD = 256, // backward of that product on the data gradient g this conv computes:  dmap[p] += sum_c g[c][p]*feat[c][p],
                  // out[s] (= dfeat) (+)= g*map;  aux[s] = feat, map_out[s] = map (read), in[1][s] = dmap (P16 fp32 map)
  F_MASK16 = 512, // like F_MASK, but aux[s] points to the fp16 half planes of the forward activation (persistent executor:
                  // activations that only feed other convs are stored as fp16 planes alone, no F_STORE)
};

// byte offset of the fp16 shadow behind a 128-channel fp32 plane buffer of P slots per plane
__host__ __device__ inline size_t shadow_bytes(int P) { return static_cast<size_t>(kKC) * P * 16; }

struct ConvCfg {
  int n_kb;        // number of 16-channel k-blocks (K = n_kb*16*ntaps)
  int kb_per_in;   // k-blocks taken from each input tensor (n_kb / kb_per_in inputs, <= 2)
  int ntaps;       // 9 (3x3) or 1 (1x1)
  int dil;         // dilation
  int S_in, P_in;  // plane format of the input(s)
  int S_out, P_out;
  int S_aux, P_aux;
  int flags;
  int lead;        // zero lead gap (slots) in front of each sample's staged planes
};

struct ConvTask {
  const void* in[2][NSMAX];   // [input][sample] -> fp16 half plane 0 of that input (the shadow copy)
  float* out[NSMAX];
  const float* aux[NSMAX];    // ReLU-mask source
  float* map_out[NSMAX];      // P16 attention map (F_DOTSIG)
  const void* w;              // packed fp16 weight tiles: [kb][tap][kc(2)][n(128)][8]
  const float* bias;          // [128]
  const float* w3;            // [128]
  const float* b3;            // [1]
  int cfg;
  int n_samp;
  int mt0;    // first 128-row M tile this CTA computes
  int n_mt;   // number of M tiles (the scheduler splits a sample over CTAs when a level has few tasks)
};
static_assert(sizeof(ConvTask) == 128, "ConvTask must stay 128 bytes");

// ---- wgrad ---------------------------------------------------------------------------------------
struct WgradInst {
  const void* dz;  // fp16 half planes [16][P][8] of the (loss-scaled) pre-activation gradient
  const void* x;   // fp16 half planes of the conv's input (half plane 0 of the 128-channel cin tile)
};
struct WgradTask {
  const WgradInst* inst;  // device array
  int n_inst;
  int tap_row;      // 0..2 (ty+1) for 3x3, 0 for 1x1
  int ntaps_x;      // 3 or 1
  int dil;
  int S, P;         // plane format of dz and x
  int cin_total;    // Cin of the weight tensor (row stride of dW in the reference layout)
  int cin0;         // first input channel of this 128-wide tile
  int ksize;        // 3 or 1
  float* dw;        // reference-layout gradient: [cout][cin_total][k][k]
  const float* scale;  // {loss scale, 1/loss scale}; dW is accumulated times scale[1]
};

}  // namespace pnmn
