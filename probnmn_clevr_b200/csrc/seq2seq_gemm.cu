// Tensor-core GEMMs of the LSTM seq2seq path (ProgramGenerator): one time step of an LSTM layer / the
// decoder cell with the gate non-linearities fused into the epilogue, the matching data-gradient GEMM,
// and the weight-gradient GEMM that contracts over batch rows and time.
//
// Reference arithmetic: nn.LSTM / nn.LSTMCell inside AllenNLP's SimpleSeq2Seq as used by
// probnmn/modules/seq2seq_base.py:77-92,201 (gate order i, f, g, o; c' = f*c + i*g; h' = o*tanh(c')).
//
// All three kernels use tcgen05.mma kind::f16 on split fp16 operands (hi*hi + lo*hi + hi*lo, fp32
// accumulators in TMEM; see seq2seq.h), fed by 1-D bulk async copies into an mbarrier ring, and read
// the accumulators back with tcgen05.ld for the fused epilogue.  Warp roles: 0 = copy producer,
// 1 = MMA issuer, 2 = TMEM allocation, 4..7 = epilogue (one TMEM lane quarter each).
// Each kernel has a CUDA-core twin with the same epilogue (PNMN_PG_SIMT=1, bring-up only).
#include <cstdlib>

#include "seq2seq.h"
#include "tcgen05.cuh"

namespace pnmn {

bool seq_use_pdl() {
  static const bool v = std::getenv("PNMN_PG_NOPDL") == nullptr;
  return v;
}

constexpr int kGThreads = 256;
// Ring depth.  With 2 stages (97 KB) two CTAs fit on an SM (the kernel stays below 128 registers); measured on the
// joint-training step that halves the SMs a pass occupies but costs 17 % per pass (1.42 -> 1.67 ms forward), so 4 it is.
#ifndef PNMN_PG_STAGES
#define PNMN_PG_STAGES 4
#endif
constexpr int kGStages = PNMN_PG_STAGES;
constexpr int kGABytes = 128 * 64 * 2;   // one 64-deep K chunk of a 128-row activation tile (hi or lo)
constexpr int kGWBytes = 64 * 64 * 2;    // the same of a 64-row weight tile
constexpr int kGStage = 2 * kGABytes + 2 * kGWBytes;
constexpr int kGHeader = 1024;
constexpr int kGSmem = kGHeader + kGStages * kGStage;
static_assert(kGSmem <= 227 * 1024, "step GEMM smem");

// ---- optional per-launch phase stamps (make TRACE=1; scripts/pg_step_trace.py): globaltimer ns of CTA (0, 0, 0) ----
#ifdef PNMN_PG_TRACE
constexpr int kPgTrW = 16, kPgTrCap = 8192;
__device__ long long g_pgtr[kPgTrW * kPgTrCap];
__device__ int g_pgtr_n;
__device__ __forceinline__ long long pg_gtime() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define PGT(slot) do { if (trace_on && tr_slot < kPgTrCap) g_pgtr[tr_slot * kPgTrW + (slot)] = pg_gtime(); } while (0)
#else
#define PGT(slot) do { } while (0)
#endif

struct GemmHeader {
  uint64_t full[kGStages], empty[kGStages];
  uint64_t tmem_full;
  uint32_t tmem_base;
  int tr_slot;
};

// Gate non-linearities from the fast exponential and reciprocal (MUFU.EX2 / MUFU.RCP, ~2 ulp each): absolute error below
// 2e-7, against a parity bar of 1e-3.  The IEEE expf / division / tanhf versions cost 2.4 us of a 13 us step (27 slow-path
// calls and 35 divergence regions in the SASS of the epilogue; scripts/pg_step_trace.py).
__device__ __forceinline__ float sigmoidf_(float x) {
  x = fminf(fmaxf(x, -30.f), 30.f);
  return __fdividef(1.f, 1.f + __expf(-x));
}
__device__ __forceinline__ float tanhf_(float x) {
  // |x| < 0.2: odd Taylor polynomial (next term 62/2835 x^9: < 6e-8 relative), else 1 - 2 / (1 + e^2x)
  const float xc = fminf(fmaxf(x, -15.f), 15.f);
  const float big = 1.f - __fdividef(2.f, 1.f + __expf(2.f * xc));
  const float x2 = x * x;
  const float small = x * fmaf(x2, fmaf(x2, fmaf(x2, -17.f / 315.f, 2.f / 15.f), -1.f / 3.f), 1.f);
  return fabsf(x) < 0.2f ? small : big;
}

__device__ __forceinline__ uint4 pack8(const __half* h) {
  uint4 r;
  r.x = static_cast<uint32_t>(__half_as_ushort(h[0])) | (static_cast<uint32_t>(__half_as_ushort(h[1])) << 16);
  r.y = static_cast<uint32_t>(__half_as_ushort(h[2])) | (static_cast<uint32_t>(__half_as_ushort(h[3])) << 16);
  r.z = static_cast<uint32_t>(__half_as_ushort(h[4])) | (static_cast<uint32_t>(__half_as_ushort(h[5])) << 16);
  r.w = static_cast<uint32_t>(__half_as_ushort(h[6])) | (static_cast<uint32_t>(__half_as_ushort(h[7])) << 16);
  return r;
}

// write 8 consecutive features [j0, j0+8) of row b into an operand buffer with K = 256: one 16-byte group of the hi and of
// the lo operand
__device__ __forceinline__ void store_op8(__half* op, int64_t lo_off, int b, int j0, const float* x) {
  __half hi[8], lo[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) split_f16(x[u], hi[u], lo[u]);
  const size_t off = op_off(b, j0, kSH);
  *reinterpret_cast<uint4*>(op + off) = pack8(hi);
  *reinterpret_cast<uint4*>(op + lo_off + off) = pack8(lo);
}

// ---- epilogues (shared by the tensor-core kernel and its CUDA-core twin) ----------------------------
// LSTM cell of 8 hidden units j = nt*16 + hf*8 + e of row b.  Everything the cell reads besides the accumulator -- the
// additive table row of the row's token (a dependent load behind the token), the previous cell / hidden state -- is
// fetched by lstm_preload BEFORE the accumulator is waited for, so that these loads (two dependent L2 round trips) hide
// under the operand stream and the MMAs; measured per launch (scripts/pg_step_trace.py): the epilogue took 7.4 us of a
// 13 us step when four warps did loads, math and stores for 16 units each after the accumulator had arrived.
struct LstmPre {
  float tab[4][8];   // gate (i, f, g, o) x unit
  float c[8], h[8];
  bool valid;
};
__device__ __forceinline__ void ld8(const float* p, float* d) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  d[0] = a.x; d[1] = a.y; d[2] = a.z; d[3] = a.w; d[4] = b.x; d[5] = b.y; d[6] = b.z; d[7] = b.w;
}
__device__ __forceinline__ void st8(float* p, const float* d) {
  *reinterpret_cast<float4*>(p) = make_float4(d[0], d[1], d[2], d[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(d[4], d[5], d[6], d[7]);
}
__device__ __forceinline__ void lstm_preload(const GemmArgs& g, int b, int nt, int hf, LstmPre& p) {
  if (b >= g.B) return;
  const int j0 = nt * 16 + hf * 8;
  const float* tab = g.table + static_cast<size_t>(g.tok ? g.tok[static_cast<size_t>(b) * g.tok_stride] : 0) * kSG + j0;
  p.valid = !g.len || g.t < g.len[b];
#pragma unroll
  for (int gate = 0; gate < 4; ++gate) ld8(tab + gate * kSH, p.tab[gate]);
  ld8(g.c_prev + static_cast<size_t>(b) * kSH + j0, p.c);
  if (!p.valid) ld8(g.h_prev + static_cast<size_t>(b) * kSH + j0, p.h);   // only carried through a masked step
}
// acc[gate][e]: pre-activation (recurrent part) of gate `gate` of unit e.  (Staging the fp32 results in shared memory and
// writing them with four lanes per 64-byte row segment was tried: slower, 2.5 us for the store loop against ~1.1 us for
// the direct thread = row stores below.)
__device__ __forceinline__ void epi_lstm_half(const GemmArgs& g, int b, int nt, int hf, const LstmPre& p,
                                              const float (&acc)[4][8]) {
  if (b >= g.B) return;
  const int j0 = nt * 16 + hf * 8;
  float gi[8], gf[8], gg[8], go[8], hn[8], cn[8], ov[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const float i_ = sigmoidf_(acc[0][e] + p.tab[0][e]);
    const float f_ = sigmoidf_(acc[1][e] + p.tab[1][e]);
    const float g_ = tanhf_(acc[2][e] + p.tab[2][e]);
    const float o_ = sigmoidf_(acc[3][e] + p.tab[3][e]);
    const float c2 = f_ * p.c[e] + i_ * g_;
    const float h2 = o_ * tanhf_(c2);
    gi[e] = i_; gf[e] = f_; gg[e] = g_; go[e] = o_;
    hn[e] = p.valid ? h2 : p.h[e];
    cn[e] = p.valid ? c2 : p.c[e];
    ov[e] = p.valid ? h2 : 0.f;
  }
  // the operand copies are [k/8][row][8]: consecutive rows are consecutive 16-byte groups, thread = row is coalesced
  store_op8(g.h_op, g.h_op_lo, b, j0, hn);
  if (g.out_op) store_op8(g.out_op, g.out_op_lo, b, j0, ov);
  st8(g.h_out + static_cast<size_t>(b) * kSH + j0, hn);
  st8(g.c_out + static_cast<size_t>(b) * kSH + j0, cn);
  if (g.out_f) st8(g.out_f + static_cast<size_t>(b) * g.out_stride + j0, ov);
  if (g.gates) {
    float* gd = g.gates + static_cast<size_t>(b) * kSG + j0;
    st8(gd, gi); st8(gd + kSH, gf); st8(gd + 2 * kSH, gg); st8(gd + 3 * kSH, go);
  }
}
#ifdef PNMN_BRINGUP
// whole 64-column tile of one row (CUDA-core twin): acc[n], n = gate*16 + u
__device__ __forceinline__ void epi_lstm(const GemmArgs& g, int b, int nt, const float* acc) {
#pragma unroll
  for (int hf = 0; hf < 2; ++hf) {
    LstmPre p;
    lstm_preload(g, b, nt, hf, p);
    float a4[4][8];
#pragma unroll
    for (int gate = 0; gate < 4; ++gate)
#pragma unroll
      for (int e = 0; e < 8; ++e) a4[gate][e] = acc[gate * 16 + hf * 8 + e];
    epi_lstm_half(g, b, nt, hf, p, a4);
  }
}

#endif

// acc[c], c in [0, 64): column n = nt*64 + c of [dA1 | dA2].  The K range is split over blockIdx.z, so the partial
// sums are ADDED (fire-and-forget reductions) into buffers whose consumers left them zeroed; masked rows (beyond the
// sequence length) contribute nothing: carried state stays, per-step outputs stay zero.
__device__ __forceinline__ void epi_dgrad_cols(const GemmArgs& g, int b, int nt, int c0, int nc, const float* acc) {
  if (b >= g.B) return;
  if (g.len && g.t >= g.len[b]) return;
  const int n0 = nt * 64, which = n0 >> 8, col0 = n0 & 255;
  const float s = g.scale[1];
  float* od = g.out[which] + static_cast<size_t>(b) * kSH + col0 + c0;
#pragma unroll
  for (int c = 0; c < 32; c += 4)
    if (c < nc) red_add_f32x4(od + c, acc[c] * s, acc[c + 1] * s, acc[c + 2] * s, acc[c + 3] * s);
}
#ifdef PNMN_BRINGUP
__device__ __forceinline__ void epi_dgrad(const GemmArgs& g, int b, int nt, const float* acc) {
  epi_dgrad_cols(g, b, nt, 0, 32, acc);
  epi_dgrad_cols(g, b, nt, 32, 32, acc + 32);
}
#endif

// ---- tensor-core kernel ------------------------------------------------------------------------------
template <int EPI>
__global__ void __launch_bounds__(kGThreads, kGStages <= 2 ? 2 : 1) step_gemm_tc_kernel(const GemmPair pr) {
  extern __shared__ __align__(1024) uint8_t smem[];
  GemmHeader* hdr = reinterpret_cast<GemmHeader*>(smem);
  uint8_t* ring = smem + kGHeader;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int which = blockIdx.y >= pr.m_tiles ? 1 : 0;
  const int nt = blockIdx.x, mt = blockIdx.y - which * pr.m_tiles;
  if (nt >= pr.n_tiles[which]) {   // the two GEMMs of a pair may have different widths
    pdl_launch_dependents();
    return;
  }
  const GemmArgs& g = pr.g[which];
  // K chunks [kc_lo, kc_hi) of this CTA (EPI_DGRAD splits K over blockIdx.z; EPI_LSTM runs with gridDim.z == 1)
  const int kc_lo = (g.K / 64) * blockIdx.z / gridDim.z, kc_hi = (g.K / 64) * (blockIdx.z + 1) / gridDim.z;
  const int n_chunks = kc_hi - kc_lo;
  pdl_launch_dependents();
#ifdef PNMN_PG_TRACE
  const bool trace_on = blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
  if (trace_on && threadIdx.x == 0) {
    hdr->tr_slot = atomicAdd(&g_pgtr_n, 1);
    if (hdr->tr_slot < kPgTrCap) {
      g_pgtr[hdr->tr_slot * kPgTrW + 0] = pg_gtime();
      g_pgtr[hdr->tr_slot * kPgTrW + 15] = EPI * 1000000 + g.K * 1000 + gridDim.x * gridDim.y * gridDim.z;
    }
  }
#endif

  if (threadIdx.x == 0) {
    for (int i = 0; i < kGStages; ++i) {
      mbar_init(smem_u32(&hdr->full[i]), 1);
      mbar_init(smem_u32(&hdr->empty[i]), 1);
    }
    mbar_init(smem_u32(&hdr->tmem_full), 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<64>(smem_u32(&hdr->tmem_base));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = hdr->tmem_base;
#ifdef PNMN_PG_TRACE
  const int tr_slot = hdr->tr_slot;
  if (threadIdx.x == 0) PGT(1);
#endif

  // Roles: lane 0 of warp 0 streams the operands, lane 0 of warp 1 issues the MMAs; then ALL eight warps run the epilogue:
  // warp w owns TMEM lane quarter q = w % 4 (rows q*32 + lane) and column half hf = w / 4 (EPI_LSTM: hidden units
  // hf*8 .. hf*8+7 of the tile, EPI_DGRAD: columns hf*32 .. hf*32+31).
  const int q = warp & 3, hf = warp >> 2;
  const int b = mt * 128 + q * 32 + lane;
  if (warp == 0 && lane == 0) {
    // the packed weights were written long before the previous step: prefetch them for the first ring pass, THEN wait
    // for the previous kernel (whose epilogue produces this step's activation operand)
    const int pre = n_chunks < kGStages ? n_chunks : kGStages;
    for (int it = 0; it < pre; ++it) {
      const uint32_t bar = smem_u32(&hdr->full[it]);
      mbar_arrive_expect_tx(bar, kGStage);
      const __half* w = g.w + (static_cast<size_t>(nt) * (g.K >> 3) + (kc_lo + it) * 8) * 512;
      const uint32_t dst = smem_u32(ring + it * kGStage);
      bulk_g2s(dst + 2 * kGABytes, w, kGWBytes, bar);
      bulk_g2s(dst + 2 * kGABytes + kGWBytes, w + g.w_lo, kGWBytes, bar);
    }
  }
  pdl_wait();   // everything below reads what the previous step's kernel wrote
  LstmPre lp;
  // (the producer thread first gets the operand stream going, then fetches its own row's epilogue inputs)
  if (EPI == EPI_LSTM && !(warp == 0 && lane == 0)) lstm_preload(g, b, nt, hf, lp);
  if (warp == 0 && lane == 0) {
    PGT(2);
    const int pre = n_chunks < kGStages ? n_chunks : kGStages;
    for (int it = 0; it < n_chunks; ++it) {
      const int kc = kc_lo + it;
      const int st = it % kGStages;
      const uint32_t bar = smem_u32(&hdr->full[st]);
      const uint32_t dst = smem_u32(ring + st * kGStage);
      if (it >= pre) {
        mbar_wait(smem_u32(&hdr->empty[st]), ((it / kGStages) & 1) ^ 1);
        mbar_arrive_expect_tx(bar, kGStage);
        const __half* w = g.w + (static_cast<size_t>(nt) * (g.K >> 3) + kc * 8) * 512;
        bulk_g2s(dst + 2 * kGABytes, w, kGWBytes, bar);
        bulk_g2s(dst + 2 * kGABytes + kGWBytes, w + g.w_lo, kGWBytes, bar);
      }
      const int si = kc / g.chunks_per_src, cj = kc % g.chunks_per_src;
      const __half* a = g.a[si] + (static_cast<size_t>(mt) * (g.a_K[si] >> 3) + cj * 8) * 1024;
      bulk_g2s(dst, a, kGABytes, bar);
      bulk_g2s(dst + kGABytes, a + g.a_lo[si], kGABytes, bar);
      if (EPI == EPI_LSTM && it + 1 == (n_chunks < kGStages ? n_chunks : kGStages)) lstm_preload(g, b, nt, hf, lp);
    }
    PGT(3);
  } else if (warp == 1 && lane == 0) {
    const uint32_t idesc = make_idesc_f16(128, 64, 0, 0);
    // K-major, no swizzle: LBO = distance between the two 8-deep halves of a k16 step (one plane), SBO = 128 B
    const uint64_t a_hi = make_smem_desc(0, 128 * 16, 128) & 0xFFFFFFFFFFFFC000ull;
    const uint64_t w_hi = make_smem_desc(0, 64 * 16, 128) & 0xFFFFFFFFFFFFC000ull;
    for (int kc = 0; kc < n_chunks; ++kc) {   // kc counts this CTA's chunks
      const int st = kc % kGStages;
      mbar_wait(smem_u32(&hdr->full[st]), (kc / kGStages) & 1);
      tc_fence_after();
      if (kc == 0) PGT(4);
      if (kc == n_chunks - 1) PGT(5);
      const uint32_t base = smem_u32(ring + st * kGStage);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint64_t ah = a_hi | (((base + i * 2 * 2048) >> 4) & 0x3FFF);
        const uint64_t al = a_hi | (((base + kGABytes + i * 2 * 2048) >> 4) & 0x3FFF);
        const uint64_t wh = w_hi | (((base + 2 * kGABytes + i * 2 * 1024) >> 4) & 0x3FFF);
        const uint64_t wl = w_hi | (((base + 2 * kGABytes + kGWBytes + i * 2 * 1024) >> 4) & 0x3FFF);
        umma_f16(tmem_base, ah, wh, idesc, (kc | i) != 0);
        umma_f16(tmem_base, al, wh, idesc, 1);
        umma_f16(tmem_base, ah, wl, idesc, 1);
      }
      umma_commit(smem_u32(&hdr->empty[st]));
    }
    umma_commit(smem_u32(&hdr->tmem_full));
    PGT(6);
  }
  __syncwarp();
  mbar_wait(smem_u32(&hdr->tmem_full), 0);
  tc_fence_after();
  if (threadIdx.x == 128) PGT(7);
  const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
  if (EPI == EPI_LSTM) {
    uint32_t v[4][8];
#pragma unroll
    for (int gate = 0; gate < 4; ++gate) tmem_ld8(lane_addr + gate * 16 + hf * 8, v[gate]);
    tmem_ld_wait();
    if (threadIdx.x == 128) PGT(8);
    float acc[4][8];
#pragma unroll
    for (int gate = 0; gate < 4; ++gate)
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[gate][e] = __uint_as_float(v[gate][e]);
    epi_lstm_half(g, b, nt, hf, lp, acc);
  } else {
    uint32_t v[32];
    tmem_ld32(lane_addr + hf * 32, v);
    tmem_ld_wait();
    if (threadIdx.x == 128) PGT(8);
    float acc[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[i] = __uint_as_float(v[i]);
    epi_dgrad_cols(g, b, nt, hf * 32, 32, acc);
  }
  if (threadIdx.x == 128) PGT(9);
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) PGT(10);
  if (warp == 2) tmem_dealloc<64>(tmem_base);
}

#ifdef PNMN_PG_TRACE
}  // namespace pnmn
// trace builds only: copies the stamps of the launches recorded so far ([n][16] int64) and clears the counter
extern "C" int pnmn_debug_pg_trace(long long* out, int cap) {
  int n = 0;
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(&n, pnmn::g_pgtr_n, sizeof(int));
  if (n > pnmn::kPgTrCap) n = pnmn::kPgTrCap;
  if (n > cap) n = cap;
  if (n > 0) cudaMemcpyFromSymbol(out, pnmn::g_pgtr, sizeof(long long) * pnmn::kPgTrW * n);
  const int zero = 0;
  cudaMemcpyToSymbol(pnmn::g_pgtr_n, &zero, sizeof(int));
  return n;
}
namespace pnmn {
#endif

#ifdef PNMN_BRINGUP
// ---- CUDA-core twin (bring-up): thread = batch row, 64 fp32 accumulators ------------------------------
template <int EPI>
__global__ void __launch_bounds__(128) step_gemm_simt_kernel(const GemmPair pr) {
  const int which = blockIdx.y >= pr.m_tiles ? 1 : 0;
  const int nt = blockIdx.x, mt = blockIdx.y - which * pr.m_tiles;
  if (nt >= pr.n_tiles[which]) return;
  const GemmArgs& g = pr.g[which];
  const int b = mt * 128 + threadIdx.x;
  float acc[64];
#pragma unroll
  for (int n = 0; n < 64; ++n) acc[n] = 0.f;
  for (int k = 0; k < g.K; ++k) {
    const int kc = k / 64, si = kc / g.chunks_per_src;
    const int kk = k - si * g.chunks_per_src * 64;
    const size_t ao = op_off(b, kk, g.a_K[si]);
    const float av = __half2float(g.a[si][ao]) + __half2float(g.a[si][g.a_lo[si] + ao]);
#pragma unroll 8
    for (int n = 0; n < 64; ++n) {
      const size_t wo = wp_off(nt * 64 + n, k, g.K);
      acc[n] = fmaf(av, __half2float(g.w[wo]) + __half2float(g.w[g.w_lo + wo]), acc[n]);
    }
  }
  if (EPI == EPI_LSTM) epi_lstm(g, b, nt, acc);
  else epi_dgrad(g, b, nt, acc);
}

#endif

cudaError_t launch_step_gemm_pair(const GemmPair& p, int epilogue, bool simt, cudaStream_t st) {
  // the data-gradient GEMMs have K = 1024 and few output tiles: split K four ways (partial sums are reduced with atomics)
  const int nx = p.count > 1 && p.n_tiles[1] > p.n_tiles[0] ? p.n_tiles[1] : p.n_tiles[0];
  static const int ksplit = std::getenv("PNMN_PG_DGRAD_SPLIT") ? std::atoi(std::getenv("PNMN_PG_DGRAD_SPLIT")) : 4;
  const dim3 grid(nx, p.m_tiles * p.count, (epilogue == EPI_DGRAD && !simt) ? ksplit : 1);
  if (simt) {
#ifdef PNMN_BRINGUP
    if (epilogue == EPI_LSTM) step_gemm_simt_kernel<EPI_LSTM><<<grid, 128, 0, st>>>(p);
    else step_gemm_simt_kernel<EPI_DGRAD><<<grid, 128, 0, st>>>(p);
    return cudaGetLastError();
#else
    return cudaErrorNotSupported;   // built without the bring-up kernels (make BRINGUP=1)
#endif
  }
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(step_gemm_tc_kernel<EPI_LSTM>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGSmem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(step_gemm_tc_kernel<EPI_DGRAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGSmem);
    if (e != cudaSuccess) return e;
    attr_done = true;
  }
  const bool pdl = seq_use_pdl();
  if (epilogue == EPI_LSTM) return launch_pdl(step_gemm_tc_kernel<EPI_LSTM>, grid, dim3(kGThreads), kGSmem, st, pdl, p);
  return launch_pdl(step_gemm_tc_kernel<EPI_DGRAD>, grid, dim3(kGThreads), kGSmem, st, pdl, p);
}

cudaError_t launch_step_gemm(const GemmArgs& g, int epilogue, int n_tiles, int m_tiles, bool simt, cudaStream_t st) {
  GemmPair p;
  p.g[0] = g; p.g[1] = g;
  p.n_tiles[0] = p.n_tiles[1] = n_tiles; p.m_tiles = m_tiles; p.count = 1;
  return launch_step_gemm_pair(p, epilogue, simt, st);
}

// =====================================================================================================
// weight gradient: dW[g][k] += scale[1] * sum_{t,b} dG[t][b][g] * X[t][b][k]
// GEMM view: M = 128 gate rows (blockIdx.x), N = 128 features (blockIdx.y), K = batch rows x time; both
// operands are read MN-major straight from the operand buffers (64 batch rows per ring stage); the time
// range is split over blockIdx.z and the partial sums are combined with fp32 atomics.
// =====================================================================================================
constexpr int kWSStages = 3;
constexpr int kWSRows = 64;
constexpr int kWSPlane = kWSRows * 16;             // 1 KB: 8 channels of 64 batch rows
constexpr int kWSOperand = 16 * kWSPlane;          // 128 channels (hi or lo)
constexpr int kWSStage = 4 * kWSOperand;           // dG hi, dG lo, X hi, X lo
constexpr int kWSSmem = kGHeader + kWSStages * kWSStage;
static_assert(kWSSmem <= 227 * 1024, "wgrad_seq smem");

struct WSHeader {
  uint64_t full[kWSStages], empty[kWSStages];
  uint64_t tmem_full;
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(kGThreads, 1) wgrad_seq_tc_kernel(const WgradSeqArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  WSHeader* hdr = reinterpret_cast<WSHeader*>(smem);
  uint8_t* ring = smem + kGHeader;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gt = blockIdx.x, nt = blockIdx.y;
  const int t0 = static_cast<int>((static_cast<long long>(a.T) * blockIdx.z) / gridDim.z);
  const int t1 = static_cast<int>((static_cast<long long>(a.T) * (blockIdx.z + 1)) / gridDim.z);
  const int per_t = a.m_tiles * 2;                 // 64-row chunks per time step
  const int total = (t1 - t0) * per_t;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kWSStages; ++i) {
      mbar_init(smem_u32(&hdr->full[i]), 1);
      mbar_init(smem_u32(&hdr->empty[i]), 1);
    }
    mbar_init(smem_u32(&hdr->tmem_full), 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<128>(smem_u32(&hdr->tmem_base));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = hdr->tmem_base;

  if (warp == 0) {
    for (int it = 0; it < total; ++it) {
      const int st = it % kWSStages;
      const uint32_t bar = smem_u32(&hdr->full[st]);
      if (lane == 0) {
        mbar_wait(smem_u32(&hdr->empty[st]), ((it / kWSStages) & 1) ^ 1);
        mbar_arrive_expect_tx(bar, kWSStage);
      }
      __syncwarp();
      if (lane < 16) {
        const int t = t0 + it / per_t, c = it % per_t, mt = c >> 1, r0 = (c & 1) * kWSRows;
        const __half* dg = a.dg + static_cast<size_t>(t) * a.dg_step +
                           (static_cast<size_t>(mt) * (kSG >> 3) + gt * 16 + lane) * 1024 + r0 * 8;
        const __half* x = a.x + static_cast<size_t>(t) * a.x_step +
                          (static_cast<size_t>(mt) * (kSH >> 3) + nt * 16 + lane) * 1024 + r0 * 8;
        const uint32_t dst = smem_u32(ring + st * kWSStage) + lane * kWSPlane;
        bulk_g2s(dst, dg, kWSPlane, bar);
        bulk_g2s(dst + kWSOperand, dg + a.dg_lo, kWSPlane, bar);
        bulk_g2s(dst + 2 * kWSOperand, x, kWSPlane, bar);
        bulk_g2s(dst + 3 * kWSOperand, x + a.x_lo, kWSPlane, bar);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_f16(128, 128, 1, 1);
      // MN-major, no swizzle: LBO = next 8-row K group (128 B), SBO = next 8-channel plane
      const uint64_t d_hi = make_smem_desc(0, 128, kWSPlane) & 0xFFFFFFFFFFFFC000ull;
      for (int it = 0; it < total; ++it) {
        const int st = it % kWSStages;
        mbar_wait(smem_u32(&hdr->full[st]), (it / kWSStages) & 1);
        tc_fence_after();
        const uint32_t base = smem_u32(ring + st * kWSStage);
#pragma unroll
        for (int k16 = 0; k16 < kWSRows / 16; ++k16) {
          const uint64_t gh = d_hi | (((base + k16 * 256) >> 4) & 0x3FFF);
          const uint64_t gl = d_hi | (((base + kWSOperand + k16 * 256) >> 4) & 0x3FFF);
          const uint64_t xh = d_hi | (((base + 2 * kWSOperand + k16 * 256) >> 4) & 0x3FFF);
          const uint64_t xl = d_hi | (((base + 3 * kWSOperand + k16 * 256) >> 4) & 0x3FFF);
          umma_f16(tmem_base, gh, xh, idesc, (it | k16) != 0);
          umma_f16(tmem_base, gl, xh, idesc, 1);
          umma_f16(tmem_base, gh, xl, idesc, 1);
        }
        umma_commit(smem_u32(&hdr->empty[st]));
      }
      umma_commit(smem_u32(&hdr->tmem_full));
    }
  } else if (warp >= 4) {
    const int q = warp & 3;
    const int grow = gt * 128 + q * 32 + lane;
    mbar_wait(smem_u32(&hdr->tmem_full), 0);
    tc_fence_after();
    const float unscale = a.scale[1];
    for (int chunk = 0; chunk < 4; ++chunk) {
      uint32_t v[32];
      tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + chunk * 32, v);
      tmem_ld_wait();
      if (total > 0) {
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          red_add_f32x4(a.dw + static_cast<size_t>(grow) * a.ld + nt * 128 + chunk * 32 + j, __uint_as_float(v[j]) * unscale,
                        __uint_as_float(v[j + 1]) * unscale, __uint_as_float(v[j + 2]) * unscale, __uint_as_float(v[j + 3]) * unscale);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<128>(tmem_base);
}

#ifdef PNMN_BRINGUP
__global__ void __launch_bounds__(256) wgrad_seq_simt_kernel(const WgradSeqArgs a) {
  const int k = blockIdx.x * 16 + (threadIdx.x & 15);
  const int grow = blockIdx.y * 16 + (threadIdx.x >> 4);
  float acc = 0.f;
  for (int t = 0; t < a.T; ++t) {
    const __half* dg = a.dg + static_cast<size_t>(t) * a.dg_step;
    const __half* x = a.x + static_cast<size_t>(t) * a.x_step;
    for (int b = 0; b < a.m_tiles * 128; ++b) {
      const size_t go = op_off(b, grow, kSG), xo = op_off(b, k, kSH);
      acc = fmaf(__half2float(dg[go]) + __half2float(dg[a.dg_lo + go]), __half2float(x[xo]) + __half2float(x[a.x_lo + xo]), acc);
    }
  }
  atomicAdd(a.dw + static_cast<size_t>(grow) * a.ld + k, acc * a.scale[1]);
}

#endif

cudaError_t launch_wgrad_seq(const WgradSeqArgs& a, bool simt, cudaStream_t st) {
  if (a.T <= 0) return cudaSuccess;
  if (simt) {
#ifdef PNMN_BRINGUP
    wgrad_seq_simt_kernel<<<dim3(kSH / 16, kSG / 16), 256, 0, st>>>(a);
    return cudaGetLastError();
#else
    return cudaErrorNotSupported;
#endif
  }
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_seq_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWSSmem);
    if (e != cudaSuccess) return e;
    attr_done = true;
  }
  const int split = a.T < 9 ? a.T : 9;   // 8 x 2 x 9 = 144 CTAs on 148 SMs
  wgrad_seq_tc_kernel<<<dim3(kSG / 128, kSH / 128, split), kGThreads, kWSSmem, st>>>(a);
  return cudaGetLastError();
}

}  // namespace pnmn
