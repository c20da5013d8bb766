// Split-precision GEMM of the classifier (probnmn/models/nmn.py:75-83: Conv2d 1x1 128 -> 1024 over the B*196 pixels,
// Linear 50176 -> 1024, Linear 1024 -> 28, and the six gradient products behind them):
//
//     C[m][n] (+)= sum_k A(m, k) * B(n, k) (+ bias[n])        A(m, k) = A[m*a_rs + k*a_ks],  B(n, k) = B[n*b_rs + k*b_ks]
//
// with fp32 operands of ANY row / k strides (so that x.w^T, g.w and g^T.x are the same code), fp32 result, and fp32-class
// products on the tensor cores: every operand value is split into two bf16 numbers, x ~= hi + lo (16 mantissa bits), and each
// k16 step issues three tcgen05.mma kind::f16 (hi.hi + lo.hi + hi.lo) into one fp32 TMEM accumulator.
//
// Two kernels.  gemm_pack_kernel reads an operand in its home layout (16-byte loads along k when k is the contiguous
// dimension, lane = row when the row is), splits it and writes bf16 tiles in the tensor core's canonical K-major core-matrix
// layout, one contiguous 32 KB block (hi tile, lo tile) per (128 rows, 64 k): the operand of a ring stage is then ONE bulk
// async copy.  gemm_split_tc_kernel is persistent (one CTA per SM walks the (split, tile) list): one thread streams the
// blocks through a three-stage mbarrier ring, one thread issues the MMAs (12 per stage), four warps drain the accumulator
// with tcgen05.ld (bias, 16-byte stores); two TMEM accumulators alternate so that the epilogue of a tile overlaps the next
// tile's operand stream.  A contraction that is split over CTAs (few output tiles, long K) leaves its partial tiles in a
// workspace that a second kernel adds up in a fixed order: results are deterministic.
//
// (First version, measured and replaced: eight producer warps converted the fp32 operands inside the GEMM kernel, straight
// from their home layout.  It was bound by the barrier hand-over of its small 32-k stages (~700 cycles per stage with
// nothing else in the loop) plus the exposed latency of one register-held chunk per thread; cp.async staging four chunks
// deep added two passes over shared memory and was no faster.)
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdlib>
#include <string>

#include "../../include/pnmn.h"
#include "tcgen05.cuh"

namespace pnmn { void set_last_error(const std::string& s); void count_launches(int n); }

namespace {

using namespace pnmn;

int fail(const std::string& s) {
  pnmn::set_last_error(s);
  return 1;
}

constexpr int kTM = 128, kTN = 128, kTK = 64;         // CTA tile, k chunk per ring stage
constexpr int kStages = 3;
constexpr int kOpBytes = kTM * kTK * 2;               // one bf16 tile of one operand (hi or lo): 16 KB
constexpr int kBlockBytes = 2 * kOpBytes;             // packed block of one operand: hi tile, lo tile
constexpr int kStageBytes = 2 * kBlockBytes;          // A block, B block: 64 KB
constexpr int kHeader = 1024;
constexpr int kSmem = kHeader + kStages * kStageBytes;
static_assert(kSmem <= 227 * 1024, "gemm smem");
constexpr int kThreads = 6 * 32;                      // warp 0: copies, warp 1: TMEM + MMAs, warps 2-5: epilogue

struct Header {
  uint64_t full[kStages], empty[kStages];
  uint64_t acc_full[2], acc_empty[2];                 // the two 128-column accumulators alternate between tiles
  uint32_t tmem_base;
};

// ---- operand packing ----------------------------------------------------------------------------------------------------
// dst block (rt, kc) at ((rt * n_kc + kc) * kBlockBytes): hi tile then lo tile, each [k/8 (8)][128 rows][8 bf16]; rows beyond
// `rows` and k beyond K are zero.  One thread per (row, 8 k): lanes run along the rows (16-byte stores coalesce).
__global__ void __launch_bounds__(256) gemm_pack_kernel(const float* __restrict__ src, int64_t rs, int64_t ks, int rows, int K,
                                                        int n_kc, int vec, uint8_t* __restrict__ dst, int64_t items) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < items; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    // i = ((rt * n_kc + kc) * 8 + kg) * 128 + r
    const int r = static_cast<int>(i & 127), kg = static_cast<int>((i >> 7) & 7);
    const int64_t blk = i >> 10;
    const int kc = static_cast<int>(blk % n_kc), rt = static_cast<int>(blk / n_kc);
    const int row = rt * kTM + r, k0 = kc * kTK + kg * 8;
    float v[8];
    if (row < rows && k0 < K) {
      const float* p = src + static_cast<int64_t>(row) * rs;
      if (vec && k0 + 8 <= K) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(p + k0)), b = __ldg(reinterpret_cast<const float4*>(p + k0 + 4));
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = k0 + e < K ? __ldg(p + static_cast<int64_t>(k0 + e) * ks) : 0.f;
      }
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = 0.f;
    }
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      // two values per conversion instruction (cvt.rn.bf16x2.f32); bf16 -> fp32 is a shift / mask of the bit pattern
      const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
      const uint32_t hb = *reinterpret_cast<const uint32_t*>(&h);
      const __nv_bfloat162 l = __floats2bfloat162_rn(v[2 * e] - __uint_as_float(hb << 16), v[2 * e + 1] - __uint_as_float(hb & 0xffff0000u));
      hi[e] = hb;
      lo[e] = *reinterpret_cast<const uint32_t*>(&l);
    }
    uint8_t* d = dst + blk * kBlockBytes + (kg * kTM + r) * 16;
    *reinterpret_cast<uint4*>(d) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(d + kOpBytes) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}

struct GemmArgs {
  const uint8_t* A;       // packed operands (gemm_pack_kernel)
  const uint8_t* B;
  float* C; int64_t ldc;
  int64_t split_stride;   // floats between the outputs of consecutive splits (0: every split writes / adds into C itself)
  const float* bias;      // added by split 0 (only without a separate reduction pass)
  int M, N, K;
  int tiles_m, tiles_n, splits, n_kc;
  int chunks_per_split;   // k chunks (of kTK) per split
  int atomic;             // 1: red.add into C, 0: store
};

// Persistent: CTA c works on tiles c, c + gridDim.x, ... of the (split, tile_m, tile_n) list.  The operand ring and its
// barrier phases run on across tiles; the accumulator alternates between two TMEM regions so that the epilogue warps drain
// tile i while the copies and the MMAs of tile i + 1 are already under way (products with a short contraction -- the 1x1
// conv's K = 128, the big Linear's weight gradient with K = batch -- are otherwise all prologue and epilogue).
__global__ void __launch_bounds__(kThreads, 1) gemm_split_tc_kernel(const GemmArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  Header* hdr = reinterpret_cast<Header*>(smem);
  uint8_t* ring = smem + kHeader;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_tiles = a.splits * a.tiles_m * a.tiles_n;

  if (tid == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(smem_u32(&hdr->full[i]), 1);
      mbar_init(smem_u32(&hdr->empty[i]), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&hdr->acc_full[i]), 1);
      mbar_init(smem_u32(&hdr->acc_empty[i]), 4);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<2 * kTN>(smem_u32(&hdr->tmem_base));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = hdr->tmem_base;

  if (warp == 0) {
    if (lane == 0) {
      // ---- copies: one 32 KB block per operand and stage ----
      uint32_t git = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const int tn = t % a.tiles_n, tm = (t / a.tiles_n) % a.tiles_m, sp = t / (a.tiles_n * a.tiles_m);
        const int c_lo = sp * a.chunks_per_split;
        const int n_chunks = min(a.n_kc, c_lo + a.chunks_per_split) - c_lo;
        const uint8_t* pa = a.A + (static_cast<int64_t>(tm) * a.n_kc + c_lo) * kBlockBytes;
        const uint8_t* pb = a.B + (static_cast<int64_t>(tn) * a.n_kc + c_lo) * kBlockBytes;
        for (int it = 0; it < n_chunks; ++it, ++git) {
          const int st = git % kStages;
          const uint32_t bar = smem_u32(&hdr->full[st]);
          mbar_wait(smem_u32(&hdr->empty[st]), ((git / kStages) & 1) ^ 1);
          mbar_arrive_expect_tx(bar, kStageBytes);
          const uint32_t dst = smem_u32(ring + st * kStageBytes);
          bulk_g2s(dst, pa + static_cast<int64_t>(it) * kBlockBytes, kBlockBytes, bar);
          bulk_g2s(dst + kBlockBytes, pb + static_cast<int64_t>(it) * kBlockBytes, kBlockBytes, bar);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ---- MMA issuer ----
      const uint32_t idesc = make_idesc_f16(kTM, kTN, 0, 0) | (1u << 7) | (1u << 10);   // bf16 x bf16 -> fp32
      // K-major, no swizzle: LBO = distance between 16-byte k slices (one 8-deep group of all rows), SBO = 8 rows
      const uint64_t d_hi = make_smem_desc(0, kTM * 16, 128) & 0xFFFFFFFFFFFFC000ull;
      uint32_t git = 0, n_done = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++n_done) {
        const int sp = t / (a.tiles_n * a.tiles_m);
        const int c_lo = sp * a.chunks_per_split;
        const int n_chunks = min(a.n_kc, c_lo + a.chunks_per_split) - c_lo;
        const uint32_t buf = n_done & 1;
        mbar_wait(smem_u32(&hdr->acc_empty[buf]), ((n_done >> 1) & 1) ^ 1);   // the epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t acc = tmem_base + buf * kTN;
        for (int it = 0; it < n_chunks; ++it, ++git) {
          const int st = git % kStages;
          mbar_wait(smem_u32(&hdr->full[st]), (git / kStages) & 1);
          tc_fence_after();
          const uint32_t sb = smem_u32(ring + st * kStageBytes);
#pragma unroll
          for (int i = 0; i < kTK / 16; ++i) {
            const uint32_t off = i * 2 * kTM * 16;
            const uint64_t ah = d_hi | (((sb + off) >> 4) & 0x3FFF);
            const uint64_t al = d_hi | (((sb + kOpBytes + off) >> 4) & 0x3FFF);
            const uint64_t bh = d_hi | (((sb + kBlockBytes + off) >> 4) & 0x3FFF);
            const uint64_t bl = d_hi | (((sb + kBlockBytes + kOpBytes + off) >> 4) & 0x3FFF);
            umma_f16(acc, ah, bh, idesc, (it | i) != 0);
            umma_f16(acc, al, bh, idesc, 1);
            umma_f16(acc, ah, bl, idesc, 1);
          }
          umma_commit(smem_u32(&hdr->empty[st]));
        }
        umma_commit(smem_u32(&hdr->acc_full[buf]));
      }
    }
  } else {
    // ---- epilogue: warp w owns TMEM lane quarter w % 4 (rows), all 128 columns ----
    const int q = warp & 3;
    uint32_t n_done = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++n_done) {
      const int tn = t % a.tiles_n, tm = (t / a.tiles_n) % a.tiles_m, sp = t / (a.tiles_n * a.tiles_m);
      const int m0 = tm * kTM, n0 = tn * kTN;
      const uint32_t buf = n_done & 1;
      if (lane == 0) mbar_wait(smem_u32(&hdr->acc_full[buf]), (n_done >> 1) & 1);
      __syncwarp();
      tc_fence_after();
      const int gm = m0 + q * 32 + lane;
      const uint32_t taddr = tmem_base + buf * kTN + (static_cast<uint32_t>(q * 32) << 16);
      float* cbase = a.C + static_cast<int64_t>(sp) * a.split_stride;
      const bool vec_out = (a.ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(cbase) & 15) == 0) && (n0 + kTN <= a.N);
      const bool add_bias = a.bias != nullptr && sp == 0;
#pragma unroll 1
      for (int part = 0; part < kTN / 32; ++part) {
        uint32_t v[32];
        tmem_ld32(taddr + part * 32, v);
        tmem_ld_wait();
        if (gm < a.M) {
          const int gn0 = n0 + part * 32;
          float* crow = cbase + static_cast<int64_t>(gm) * a.ldc + gn0;
          if (vec_out) {
#pragma unroll
            for (int c = 0; c < 32; c += 4) {
              float x0 = __uint_as_float(v[c]), x1 = __uint_as_float(v[c + 1]), x2 = __uint_as_float(v[c + 2]), x3 = __uint_as_float(v[c + 3]);
              if (add_bias) {
                const float4 bb = __ldg(reinterpret_cast<const float4*>(a.bias + gn0 + c));
                x0 += bb.x; x1 += bb.y; x2 += bb.z; x3 += bb.w;
              }
              if (a.atomic) red_add_f32x4(crow + c, x0, x1, x2, x3);
              else *reinterpret_cast<float4*>(crow + c) = make_float4(x0, x1, x2, x3);
            }
          } else {
#pragma unroll
            for (int c = 0; c < 32; ++c) {
              if (gn0 + c < a.N) {
                float x = __uint_as_float(v[c]);
                if (add_bias) x += __ldg(a.bias + gn0 + c);
                if (a.atomic) red_add_f32(crow + c, x);
                else crow[c] = x;
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&hdr->acc_empty[buf]));   // 4 arrivals: the MMA warp may overwrite this accumulator
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<2 * kTN>(tmem_base);
}

// C[m][n] = sum over splits of part[s][m][n] (+ bias[n]), in split order: the deterministic end of a split contraction
__global__ void __launch_bounds__(256) gemm_reduce_splits_kernel(const float* __restrict__ part, int64_t split_stride, int splits,
                                                                 const float* __restrict__ bias, float* __restrict__ C, int64_t ldc,
                                                                 int M, int N, int accumulate) {
  const int64_t total = static_cast<int64_t>(M) * N;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int m = static_cast<int>(i / N), n = static_cast<int>(i % N);
    float s = bias ? bias[n] : 0.f;
    for (int k = 0; k < splits; ++k) s += part[k * split_stride + i];
    float* c = C + static_cast<int64_t>(m) * ldc + n;
    *c = accumulate ? *c + s : s;
  }
}

}  // namespace

namespace {
int plan_splits(int M, int N, int K, int sms, int* chunks_per_split) {
  const int tiles = ((M + kTM - 1) / kTM) * ((N + kTN - 1) / kTN);
  const int chunks = (K + kTK - 1) / kTK;
  int splits = 1;
  if (tiles < sms) {
    splits = (2 * sms + tiles - 1) / tiles;             // ~two waves of work items
    const int max_splits = (chunks + 3) / 4;            // at least 4 ring stages (256 k) of work per item
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
  }
  const int cps = (chunks + splits - 1) / splits;
  if (chunks_per_split) *chunks_per_split = cps;
  return (chunks + cps - 1) / cps;
}
int device_sms() {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaGetLastError();
  return sms;
}
int64_t packed_bytes(int rows, int K) {
  return static_cast<int64_t>((rows + kTM - 1) / kTM) * ((K + kTK - 1) / kTK) * kBlockBytes;
}
}  // namespace

// floats of scratch pnmn_gemm_split needs: the two packed operands (bf16 hi / lo tiles, padded to 128 rows x 64 k) and, when the
// contraction is split over CTAs, one partial result per split (added up in a fixed order)
extern "C" int64_t pnmn_gemm_split_workspace(int M, int N, int K) {
  const int splits = plan_splits(M, N, K, device_sms(), nullptr);
  const int64_t part = splits > 1 ? static_cast<int64_t>(splits) * M * N : 0;
  return (packed_bytes(M, K) + packed_bytes(N, K)) / 4 + part;
}

// C [M][N] (row stride ldc) = (accumulate ? C : 0) + A . B^T (+ bias), see the header of this file.
extern "C" int pnmn_gemm_split(const float* A, int64_t a_rs, int64_t a_ks, const float* B, int64_t b_rs, int64_t b_ks, float* C,
                               int64_t ldc, int M, int N, int K, const float* bias, int accumulate, float* workspace,
                               int64_t workspace_floats, void* stream) {
  if (!A || !B || !C || !workspace) return fail("pnmn_gemm_split: NULL buffer");
  if (M < 1 || N < 1 || K < 1) return fail("pnmn_gemm_split: empty product");
  if (workspace_floats < pnmn_gemm_split_workspace(M, N, K)) return fail("pnmn_gemm_split: workspace too small");
  if (reinterpret_cast<uintptr_t>(workspace) & 15) return fail("pnmn_gemm_split: workspace must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  static bool attr_done = false;
  if (!attr_done) {
    const cudaError_t e = cudaFuncSetAttribute(gemm_split_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    if (e != cudaSuccess) return fail(std::string("pnmn_gemm_split: ") + cudaGetErrorString(e));
    attr_done = true;
  }
  const int sms = device_sms();
  uint8_t* pa = reinterpret_cast<uint8_t*>(workspace);
  uint8_t* pb = pa + packed_bytes(M, K);
  float* part = reinterpret_cast<float*>(pb + packed_bytes(N, K));
  const int n_kc = (K + kTK - 1) / kTK;
  int launches = 0;
  auto pack = [&](const float* src, int64_t rs, int64_t ks, int rows, uint8_t* dst) -> cudaError_t {
    const int64_t items = packed_bytes(rows, K) / 32;   // one (row, 8 k) item writes 16 B hi + 16 B lo
    const int vec = ks == 1 && rs % 4 == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0;
    const int64_t want = (items + 255) / 256;
    const int blocks = static_cast<int>(want < static_cast<int64_t>(sms) * 16 ? want : static_cast<int64_t>(sms) * 16);
    gemm_pack_kernel<<<blocks, 256, 0, st>>>(src, rs, ks, rows, K, n_kc, vec, dst, items);
    ++launches;
    return cudaGetLastError();
  };
  cudaError_t e = pack(A, a_rs, a_ks, M, pa);
  if (e == cudaSuccess) e = pack(B, b_rs, b_ks, N, pb);
  if (e != cudaSuccess) return fail(std::string("gemm_pack_kernel: ") + cudaGetErrorString(e));

  GemmArgs g;
  g.A = pa; g.B = pb; g.C = C; g.ldc = ldc; g.bias = bias; g.M = M; g.N = N; g.K = K; g.split_stride = 0; g.n_kc = n_kc;
  g.tiles_m = (M + kTM - 1) / kTM; g.tiles_n = (N + kTN - 1) / kTN;
  g.splits = plan_splits(M, N, K, sms, &g.chunks_per_split);
  const bool staged = g.splits > 1;
  if (staged) {
    // every split stores its partial tile into its own slab; a second kernel adds the slabs up in split order
    g.C = part; g.ldc = N; g.split_stride = static_cast<int64_t>(M) * N; g.bias = nullptr; g.atomic = 0;
  } else {
    g.atomic = accumulate ? 1 : 0;
  }
  const int work = g.splits * g.tiles_m * g.tiles_n;
  gemm_split_tc_kernel<<<work < sms ? work : sms, kThreads, kSmem, st>>>(g);
  ++launches;
  e = cudaGetLastError();
  if (e != cudaSuccess) return fail(std::string("gemm_split_tc_kernel: ") + cudaGetErrorString(e));
  if (staged) {
    const int64_t total = static_cast<int64_t>(M) * N;
    const int blocks = static_cast<int>(total / 256 + 1 < sms * 8 ? total / 256 + 1 : sms * 8);
    gemm_reduce_splits_kernel<<<blocks, 256, 0, st>>>(part, g.split_stride, g.splits, bias, C, ldc, M, N, accumulate);
    ++launches;
    e = cudaGetLastError();
    if (e != cudaSuccess) return fail(std::string("gemm_reduce_splits_kernel: ") + cudaGetErrorString(e));
  }
  pnmn::count_launches(launches);
  return 0;
}
