// Persistent program executor: ONE launch walks every sample's compiled program.
//
// The host-side program compiler (nmn_executor.cu) turns a batch of programs into a task list:
// tensor-core convolution tasks (shift-GEMM, see conv.cu for the formulation) and CUDA-core tasks
// (attend, same, min/max, the backward pieces; elt_body.cuh).  Every task names the tasks that
// produce its inputs.  296 resident CTAs (two per SM) pull tasks from a global counter in list
// order, spin on the `done` flags of their predecessors, run the task and publish their own flag.
// Because predecessors always sit earlier in the list and tasks are fetched in order, a waiting
// CTA can only wait for a task that is already running: no deadlock, no host round trips, no
// per-level launch latency, and chains of different samples overlap freely across the SMs.  The list
// order is the schedule: the compiler sorts it by remaining chain time (critical path first).
//
// Inside a CTA a convolution task is warp-specialised: warp 3 fetches the task record; warp 0 polls the
// dependency flags and then streams activation k-blocks, warps 1 and 3 stream weight stages with cp.async.bulk
// into mbarrier rings (the first ring-full of weights, the configuration and the bias staging do not depend on
// the producers and are issued while warp 0 still polls), warp 2 issues tcgen05.mma (kind::f16 on fp16 operands,
// fp32 accumulators in TMEM), and all 8 warps run the fused epilogue (warps 4-7 take accumulator columns 0..63
// of their TMEM lane quarter, warps 0-3 columns 64..127).  TMEM, barriers and ring phases persist across tasks.
//
// TWO CTAs are resident per SM (256 threads, <= 128 registers, 107 KB of shared memory and 256 TMEM
// columns each): a task is a strictly serial chain (wait for predecessors -> first operands -> MMAs ->
// epilogue -> publish; profiles/r1: a 72-MMA task spends 2.4 us of its 20 us in the tensor pipe), so
// the second CTA's mainloop fills the tensor pipe while the first one is in its epilogue or waiting.
// A task therefore owns at most two accumulators (one sample x two M tiles, or two samples x one).
#include <cuda_fp16.h>
#include <cstdlib>
#include <type_traits>

#include "elt_body.cuh"
#include "executor.h"
#include "exec.h"
#include "tcgen05.cuh"

namespace pnmn {

constexpr int kExThreads = 256;
constexpr int kExCtasPerSM = 2;
constexpr int kExMaxAcc = 2;               // accumulators (128 TMEM columns each) per task
constexpr int kExWTile = 16 * 128 * 2;     // 4 KB: 16 k x 128 n fp16
// Weight ring: a stage holds SIX taps (24 KB) of the linear (k-block, tap) weight stream of a 3x3 conv -- two k-blocks are
// three stages -- or one tap of a 1x1 conv.  The issuing thread pays ~175 cycles per barrier wait + commit and ~48 per
// MMA (scripts/microbench/mma_stage.cu), so with 3-tap stages a one-accumulator task spent 320 cycles of issue work per
// 192 cycles of tensor work; six taps bring the two level.  The ring is THREE stages deep when the activation ring leaves
// room (one sample, P16 / P18 planes), else two.
constexpr int kExWTapsPerStage = 6;
constexpr int kExWStage = kExWTapsPerStage * kExWTile;
constexpr int kExWStagesMax = 3;
constexpr int kExAStagesMax = 4;          // activation ring: 2-4 stages of one k-block, see the ring split in exec_kernel
constexpr int kExHeader = 5 * 1024;
constexpr int kExSmem = 107 * 1024;        // header | weight ring | activation ring | guard for the positive tap shifts
static_assert(kExCtasPerSM * (kExSmem + 1024) <= 228 * 1024, "executor smem: two CTAs per SM");

struct ExHeader {
  uint64_t full_a[kExAStagesMax], empty_a[kExAStagesMax];
  uint64_t full_w[kExWStagesMax], empty_w[kExWStagesMax];
  uint64_t tmem_full;
  uint32_t tmem_base;
  int task_idx;
  TaskMeta meta;
  alignas(16) uint8_t task[128];
  alignas(16) float bias[128];
  alignas(16) float w3[128];
  float dotp[kExMaxAcc][2][128];  // partial 64-channel dot products of the two column halves of an accumulator
  EltSmem elt;
};
static_assert(sizeof(ExHeader) <= kExHeader, "executor header");

__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ int ld_relaxed(const int* p) {
  int v;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

// two fp32 -> packed fp16x2 (a in the low half), round to nearest, saturating to +-65504: one F2FP instruction
__device__ __forceinline__ uint32_t pack_half2_sat(float a, float b) {
  uint32_t u;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(u) : "f"(b), "f"(a));
  return u;
}

// fp32 -> tf32 (round to nearest, ties away: cvt.rna.tf32.f32), result as an fp32 bit pattern with 13 zero low bits.
// Two integer instructions instead of the four ptxas emits for the cvt (no special case is needed: inf stays inf, the
// largest finite values round to inf as they should).
__device__ __forceinline__ uint32_t round_tf32(float x) { return (__float_as_uint(x) + 0x1000u) & 0xffffe000u; }
// same with ReLU folded into the conversion
__device__ __forceinline__ uint32_t pack_half2_relu_sat(float a, float b) {
  uint32_t u;
  asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(u) : "f"(b), "f"(a));
  return u;
}
// h where the fp16 pair m is > 0, else 0 (ReLU backward on packed halves)
__device__ __forceinline__ uint32_t mask_half2_gt0(uint32_t h, uint32_t m) {
  const __half2 hv = *reinterpret_cast<const __half2*>(&h);
  const __half2 mv = *reinterpret_cast<const __half2*>(&m);
  const __half2 r = __hmul2(hv, __hgt2(mv, __float2half2_rn(0.f)));
  return *reinterpret_cast<const uint32_t*>(&r);
}

__device__ __forceinline__ int ex_tap_shift(const ConvCfg& c, int tap) {
  return c.ntaps == 9 ? ((tap / 3 - 1) * c.S_in + (tap % 3 - 1)) * c.dil : 0;
}

__device__ __forceinline__ long long gtime() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ int smid() {
  int v;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(v));
  return v;
}

// trace (optional, debugging / profiling): 16 int64 per task (globaltimer, ns):
//   [0] fetched  [1] dependencies satisfied  [2] producer done  [3] published
//   [4] SM id    [5] type | n_samp<<8 | n_mt<<16   [6] MMAs per (sample, M tile)   [7] cfg flags / elt op
//   [8] roles start (after barrier B)  [9] first operands landed  [10] last MMA issued
//   [11] accumulators complete  [12] epilogue stores issued  [13] epilogue fenced
template <bool kTrace>
__global__ void __launch_bounds__(kExThreads, kExCtasPerSM)
exec_kernel(const uint8_t* __restrict__ tasks, const TaskMeta* __restrict__ metas, int n_tasks,
            const ConvCfg* __restrict__ cfgs, int* __restrict__ counter, int* __restrict__ done,
            long long* __restrict__ trace, int dbg, int sm_limit) {
  // SM partition (pnmn_set_reserved_sms): a CTA that lands on one of the reserved SMs leaves at once -- tasks are pulled
  // from a global counter, so the CTAs on the other SMs simply run the whole list -- and the SM stays free for the
  // kernels of other streams (the LSTM passes of the joint-training step, which cannot share an SM with two executor CTAs)
  if (smid() >= sm_limit) return;
  extern __shared__ __align__(1024) uint8_t smem[];
  ExHeader* hdr = reinterpret_cast<ExHeader*>(smem);
  uint8_t* w_ring = smem + kExHeader;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int i = 0; i < kExAStagesMax; ++i) {
      mbar_init(smem_u32(&hdr->full_a[i]), 1);
      mbar_init(smem_u32(&hdr->empty_a[i]), 1);
    }
    for (int i = 0; i < kExWStagesMax; ++i) {
      mbar_init(smem_u32(&hdr->full_w[i]), 1);
      mbar_init(smem_u32(&hdr->empty_w[i]), 1);
    }
    mbar_init(smem_u32(&hdr->tmem_full), 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<kExMaxAcc * 128>(smem_u32(&hdr->tmem_base));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = hdr->tmem_base;

  // Barrier phases persist across tasks.  Every task starts at ring stage 0 (all stages are free once the previous
  // task's accumulators completed), with a task-dependent number of stages, so each role thread tracks the phase parity
  // of every barrier it waits on as one bit per stage: "full" bits (the MMA issuer) start at 0, "empty" bits (the
  // producers) at 1 (= passes on a fresh barrier); a bit flips each time its stage is used.
  uint32_t pf_a = 0, pf_w = 0, pe_a = 0xfu, pe_w = 0x7u, n_conv = 0;
  // Tried and dropped in round 2 (both parity-green, both measured on the 256-row workload, 1.278 ms before):
  //  * task prefetch -- a convolution task claims its successor (atomicAdd on the list counter) and pulls its record into a
  //    second record buffer with a bulk copy while its own MMAs drain, taking the counter round trip and the dependent
  //    record load (~1 us) out of the CTA's serial chain: 1.325 ms.  A task claimed ~4 us early sits behind its claimer's
  //    epilogue while another CTA is idle, and its dependency-independent prologue (first ring-full of weights) starts later.
  //  * two MMA-issuing warps (2 and 4; two accumulators: one each; one accumulator: the six MMAs of a weight stage split
  //    three / three into TMEM slots 0 and 1 that the epilogue adds; "empty" barriers with two arrivals): 1.305 ms.  The
  //    issuing thread's ~110 cycles per MMA are not what bounds the issue phase -- the tensor pipe is shared with the
  //    second resident CTA, and the extra waits / commits of a second issuer cost more than its parallel issue saves.

  for (;;) {
    // ---------------- scheduler: fetch the next task ----------------
    if (warp == 3) {
      int idx = 0;
      if (lane == 0) idx = atomicAdd(counter, 1);
      idx = __shfl_sync(0xffffffffu, idx, 0);
      if (idx < n_tasks) {
        if (kTrace && lane == 0) trace[idx * kTraceW + 0] = gtime();
        reinterpret_cast<uint32_t*>(hdr->task)[lane] = reinterpret_cast<const uint32_t*>(tasks + static_cast<size_t>(idx) * 128)[lane];
        if (lane < static_cast<int>(sizeof(TaskMeta) / 4))
          reinterpret_cast<uint32_t*>(&hdr->meta)[lane] = reinterpret_cast<const uint32_t*>(metas + idx)[lane];
      }
      if (lane == 0) hdr->task_idx = idx;
    }
    __syncthreads();  // (A) the task record is in place; its producers may still be running
    const int idx = hdr->task_idx;
    if (idx >= n_tasks) break;
    // Wait for the producers (warp 0, the warp that streams the activations).  Everything of a task that does NOT read
    // a producer's output -- configuration, bias / head weights, the zero lead gaps and above all the first stages of the
    // WEIGHT stream -- is done by the other warps meanwhile, so a task on the critical path (which always arrives here
    // before its producer has published) starts its MMAs as soon as its activations can be fetched.
    auto wait_deps = [&]() {
      if (lane < kMaxDeps) {
        const int d = hdr->meta.deps[lane];
        if (d >= 0) {
          // relaxed polls (an acquire load invalidates the SM's whole L1 on EVERY poll, which hurts the other CTA of
          // this SM), then ONE acquire fence once every producer has published
          while (ld_relaxed(done + d) == 0) __nanosleep(32);
        }
      }
      __syncwarp();
      asm volatile("fence.acq_rel.gpu;" ::: "memory");
      if (kTrace && lane == 0) { trace[idx * kTraceW + 1] = gtime(); trace[idx * kTraceW + 4] = smid(); }
    };

    if (hdr->meta.type == TASK_CONV) {
      // Everything the roles need is copied into registers up front: the PTX wrappers carry "memory"
      // clobbers, so anything left in shared / local memory would be re-read around every MMA.
      // They are also passed through a lane-0 shuffle: that is how the compiler learns they are warp-uniform,
      // which lets the MMA issuer build its descriptors in uniform registers (no R2UR / elect loops per MMA).
#define UNI(x) __shfl_sync(0xffffffffu, (x), 0)
      const ConvTask* tp = reinterpret_cast<const ConvTask*>(hdr->task);
      const int cfg_id = UNI(tp->cfg), n_samp = UNI(tp->n_samp), mt0 = UNI(tp->mt0), n_mt = UNI(tp->n_mt);
      const ConvCfg* cp = cfgs + cfg_id;
      const int n_kb = UNI(cp->n_kb), kb_per_in = UNI(cp->kb_per_in), ntaps = UNI(cp->ntaps), dil = UNI(cp->dil);
      const int S_in = UNI(cp->S_in), P_in = UNI(cp->P_in), S_out = UNI(cp->S_out), P_out = UNI(cp->P_out);
      const int S_aux = UNI(cp->S_aux), P_aux = UNI(cp->P_aux);
      const int flags = UNI(cp->flags), lead = UNI(cp->lead);
#undef UNI
      const uint32_t plane_bytes = static_cast<uint32_t>(P_in) * 16u;
      const uint32_t samp_bytes = static_cast<uint32_t>(lead) * 16u + 2u * plane_bytes;  // 2 half planes = 16 ch
      const int tps = ntaps == 9 ? kExWTapsPerStage : 1;  // taps per weight stage
      const int n_ws = n_kb * ntaps / tps;                 // weight stages of this task (n_kb is even)
      const uint32_t a_stage = (static_cast<uint32_t>(n_samp) * samp_bytes + 127u) & ~127u;
      // ring split of this task: both streams are latency-bound (a stage can only be refilled once its MMAs completed), so
      // the shared memory left after the header and the guard for the largest positive tap shift goes to as many stages
      // as fit -- (weight, activation) stages = (3,3) for one P16 sample, (2,4) for one P18 sample, (2,3) for two P16
      // samples, else (2,2)
      const uint32_t guard = ntaps == 9 ? ((static_cast<uint32_t>((S_in + 1) * dil) * 16u + 255u) & ~127u) : 128u;
      const uint32_t ring_avail = static_cast<uint32_t>(kExSmem - kExHeader) - guard;
      int n_wst = 2, n_ast = 2;
      if (3u * a_stage + 3u * kExWStage <= ring_avail) { n_wst = 3; n_ast = 3; }
      else if (4u * a_stage + 2u * kExWStage <= ring_avail) n_ast = 4;
      else if (3u * a_stage + 2u * kExWStage <= ring_avail) n_ast = 3;
      uint8_t* a_ring = w_ring + n_wst * kExWStage;
      // ---------------- activation producer (warp 0, lane 0): k-blocks [a_kb, upto) ----------------
      int a_kb = 0, a_sa = 0;
      auto produce_a = [&](int upto) {
        const uint8_t* in00 = static_cast<const uint8_t*>(tp->in[0][0]);
        const uint8_t* in01 = static_cast<const uint8_t*>(tp->in[0][1]);
        const uint8_t* in10 = static_cast<const uint8_t*>(tp->in[1][0]);
        const uint8_t* in11 = static_cast<const uint8_t*>(tp->in[1][1]);
        const uint32_t a_base = smem_u32(a_ring) + lead * 16u;
        const uint32_t kb_bytes = 2u * plane_bytes;
        fence_proxy_async_all();
        for (; a_kb < upto; ++a_kb) {
          mbar_wait(smem_u32(&hdr->empty_a[a_sa]), (pe_a >> a_sa) & 1u);
          pe_a ^= 1u << a_sa;
          const uint32_t bar = smem_u32(&hdr->full_a[a_sa]);
          mbar_arrive_expect_tx(bar, n_samp * kb_bytes);
          const bool second = a_kb >= kb_per_in;
          const size_t off = static_cast<size_t>(second ? a_kb - kb_per_in : a_kb) * kb_bytes;
          bulk_g2s(a_base + a_sa * a_stage, (second ? in10 : in00) + off, kb_bytes, bar);
          if (n_samp > 1) bulk_g2s(a_base + a_sa * a_stage + samp_bytes, (second ? in11 : in01) + off, kb_bytes, bar);
          a_sa = a_sa + 1 == n_ast ? 0 : a_sa + 1;
        }
      };
      // ---------------- weight producers (lane 0 of warps 1 and 3): stages [w_it, upto) ----------------
      // TWO issuing threads take alternate stages: one thread can only start a bulk copy every ~500 cycles whatever its
      // size (scripts/microbench/bulk_stream.cu).  Both walk the whole stage sequence so that their phase bits follow
      // every use of every stage.
      int w_it = 0, w_sw = 0;
      auto produce_w = [&](int upto) {
        const uint8_t* wsrc = static_cast<const uint8_t*>(tp->w);
        const uint32_t bytes = static_cast<uint32_t>(tps) * kExWTile;
        const uint32_t w_base = smem_u32(w_ring);
        const int mine = warp == 3 ? 1 : 0;
        fence_proxy_async_all();
        for (; w_it < upto; ++w_it) {
          if ((w_it & 1) == mine) {
            mbar_wait(smem_u32(&hdr->empty_w[w_sw]), (pe_w >> w_sw) & 1u);
            const uint32_t bar = smem_u32(&hdr->full_w[w_sw]);
            mbar_arrive_expect_tx(bar, bytes);
            bulk_g2s(w_base + w_sw * kExWStage, wsrc + static_cast<size_t>(w_it) * bytes, bytes, bar);
          }
          pe_w ^= 1u << w_sw;
          w_sw = w_sw + 1 == n_wst ? 0 : w_sw + 1;
        }
      };
      // ---- before the producers of this task have published: the first ring-full of weights, the zero lead gaps of
      // this plane format, bias / 1x1 head weights; warp 0 polls the dependency flags and then starts the activations
      if (warp == 0) {
        wait_deps();
        if (lane == 0) produce_a(n_ast < n_kb ? n_ast : n_kb);
      } else if (warp == 1 || warp == 3) {
        if (lane == 0) produce_w(n_wst < n_ws ? n_wst : n_ws);
      } else if (warp >= 4) {
        const int t = tid - 128;
        for (int st = 0; st < n_ast; ++st)
          for (int s = 0; s < n_samp; ++s) {
            float4* g = reinterpret_cast<float4*>(a_ring + st * a_stage + s * samp_bytes);
            for (int i = t; i < lead; i += 128) g[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          }
        hdr->bias[t] = (flags & F_BIAS) ? __ldg(tp->bias + t) : 0.f;
        hdr->w3[t] = (flags & F_DOTSIG) ? __ldg(tp->w3 + t) : 0.f;
      }
      fence_proxy_async();
      __syncthreads();  // (B) producers have published (warp 0's acquire fence is cumulative over this barrier)

      if (warp == 0) {
        if (lane == 0) produce_a(n_kb);
      } else if (warp == 1 || warp == 3) {
        if (lane == 0) produce_w(n_ws);
      } else if (warp == 2) {
        // ---------------- MMA issuer ----------------
        // The whole warp runs the loop (uniform control flow and operands); one elected lane issues.
        {
          tc_fence_after();
          if (kTrace && lane == 0) trace[idx * kTraceW + 8] = gtime();
          const uint32_t idesc = make_idesc_f16(128, 128, 0, 0);
          const uint64_t d_hi = make_smem_desc(0, 0, 128u) & 0xFFFFFFFF00000000ull;
          // K-major, no swizzle: LBO = distance between the two 8-channel halves of a 16-deep k-block
          const uint32_t a_lo0 = (smem_u32(a_ring) >> 4) + static_cast<uint32_t>(lead) + (static_cast<uint32_t>(P_in) << 16);
          const uint32_t b_lo0 = (smem_u32(w_ring) >> 4) + ((2048u >> 4) << 16);
          const uint32_t samp16 = samp_bytes >> 4, a_stage16 = a_stage >> 4;
          const int row_shift = ntaps == 9 ? S_in * dil : 0;  // slots between tap rows
          const int col_shift = ntaps == 9 ? dil : 0;
          const uint32_t tm = __shfl_sync(0xffffffffu, tmem_base, 0);
          // accumulators of this task: up to kExMaxAcc (sample, M tile) pairs, TMEM slot = running index
          uint32_t aoff[kExMaxAcc], doff[kExMaxAcc];
          int n_acc = 0;
#pragma unroll
          for (int k = 0; k < kExMaxAcc; ++k) { aoff[k] = 0; doff[k] = 0; }
#pragma unroll
          for (int s = 0; s < 2; ++s)
#pragma unroll
            for (int m = 0; m < 3; ++m)
              if (s < n_samp && m >= mt0 && m < mt0 + n_mt && n_acc < kExMaxAcc) {
                const uint32_t av = s * samp16 + m * 128, dv = tm + n_acc * 128;
                if (n_acc == 0) { aoff[0] = av; doff[0] = dv; }
                else { aoff[1] = av; doff[1] = dv; }
                ++n_acc;
              }
          uint32_t fa = __shfl_sync(0xffffffffu, pf_a, 0), fw = __shfl_sync(0xffffffffu, pf_w, 0);
          long long wait_a = 0, wait_w = 0;
          constexpr bool tr = kTrace;
          const uint32_t bar_fa = smem_u32(&hdr->full_a[0]), bar_ea = smem_u32(&hdr->empty_a[0]);
          const uint32_t bar_fw = smem_u32(&hdr->full_w[0]), bar_ew = smem_u32(&hdr->empty_w[0]);
          const uint32_t d_hi32 = static_cast<uint32_t>(d_hi >> 32);
          int sw = 0, sa = 0;
          auto wait_full_a = [&](int sa) {
            const long long c0 = tr ? clock64() : 0;
            mbar_wait(bar_fa + sa * 8, (fa >> sa) & 1u);
            fa ^= 1u << sa;
            if (tr) wait_a += clock64() - c0;
          };
          auto wait_full_w = [&]() {
            const long long c0 = tr ? clock64() : 0;
            mbar_wait(bar_fw + sw * 8, (fw >> sw) & 1u);
            fw ^= 1u << sw;
            if (tr) wait_w += clock64() - c0;
          };
          // 3x3 conv: two k-blocks (18 taps) = three weight stages A, B, C per trip, everything unrolled so that every
          // descriptor is one add away from a loop-invariant value:
          //   A: taps 0-5 of k-block 0 | B: taps 6-8 of k-block 0, taps 0-2 of k-block 1 | C: taps 3-8 of k-block 1
          auto run9 = [&](auto nacc_c) {
            constexpr int NACC = decltype(nacc_c)::value;
            int sh[9];
#pragma unroll
            for (int t = 0; t < 9; ++t) sh[t] = (t / 3 - 1) * row_shift + (t % 3 - 1) * col_shift;
            for (int kb = 0; kb < n_kb; kb += 2) {
              const int sa0 = sa, sa1 = sa + 1 == n_ast ? 0 : sa + 1;  // activation stages of k-blocks kb, kb + 1
              sa = sa1 + 1 == n_ast ? 0 : sa1 + 1;
              const uint32_t a_lo_kb[2] = {a_lo0 + sa0 * a_stage16, a_lo0 + sa1 * a_stage16};
              const uint32_t bar_ea_kb[2] = {bar_ea + sa0 * 8, bar_ea + sa1 * 8};
#pragma unroll
              for (int st = 0; st < 3; ++st) {
                if (st < 2) wait_full_a(st == 0 ? sa0 : sa1);
                wait_full_w();
                tc_fence_after();
                if (tr && lane == 0 && kb == 0 && st == 0) trace[idx * kTraceW + 9] = gtime();
                const uint32_t b_lo_st = b_lo0 + sw * (kExWStage >> 4);
                if (elect_one()) {
#pragma unroll
                  for (int j = 0; j < 6; ++j) {
                    const int t18 = st * 6 + j, kbi = t18 / 9, tap = t18 % 9;
                    const uint32_t a_lo = a_lo_kb[kbi] + static_cast<uint32_t>(sh[tap]);
#pragma unroll
                    for (int k = 0; k < NACC; ++k)
                      umma_f16_2x32(doff[k], a_lo + aoff[k], d_hi32, b_lo_st + j * (kExWTile >> 4), d_hi32, idesc,
                                    t18 == 0 ? (kb == 0 ? 0u : 1u) : 1u);
                    if (tap == 8) umma_commit(bar_ea_kb[kbi]);
                  }
                  umma_commit(bar_ew + sw * 8);
                }
                __syncwarp();
                sw = sw + 1 == n_wst ? 0 : sw + 1;
              }
            }
          };
          // 1x1 conv: one tap per k-block and per weight stage
          auto run1 = [&](auto nacc_c) {
            constexpr int NACC = decltype(nacc_c)::value;
            for (int kb = 0; kb < n_kb; ++kb) {
              wait_full_a(sa);
              wait_full_w();
              tc_fence_after();
              if (tr && lane == 0 && kb == 0) trace[idx * kTraceW + 9] = gtime();
              const uint32_t a_lo = a_lo0 + sa * a_stage16, b_lo = b_lo0 + sw * (kExWStage >> 4);
              if (elect_one()) {
#pragma unroll
                for (int k = 0; k < NACC; ++k) umma_f16_2x32(doff[k], a_lo + aoff[k], d_hi32, b_lo, d_hi32, idesc, kb == 0 ? 0u : 1u);
                umma_commit(bar_ew + sw * 8);
                umma_commit(bar_ea + sa * 8);
              }
              __syncwarp();
              sw = sw + 1 == n_wst ? 0 : sw + 1;
              sa = sa + 1 == n_ast ? 0 : sa + 1;
            }
          };
          using I1 = std::integral_constant<int, 1>; using I2 = std::integral_constant<int, 2>;
          if (ntaps == 9) {
            if (n_acc == 1) run9(I1{}); else run9(I2{});
          } else {
            if (n_acc == 1) run1(I1{}); else run1(I2{});
          }
          if (elect_one()) umma_commit(smem_u32(&hdr->tmem_full));
          __syncwarp();
          pf_a = fa; pf_w = fw;
          if (kTrace && lane == 0) { trace[idx * kTraceW + 10] = gtime(); trace[idx * kTraceW + 14] = wait_a; trace[idx * kTraceW + 15] = wait_w; }
        }
      }
      __syncwarp();
      {
        // ---------------- epilogue: ALL 8 warps (the role warps join once their loops are done) ----------------
        // A warp may only read the TMEM lane quarter (warp % 4); warps 4-7 take accumulator columns 0..63 of their
        // quarter, warps 0-3 columns 64..127, so a thread owns one pixel row x 64 channels.  (Measured, profiles/r1:
        // with 4 epilogue warps a 32-column chunk took 0.7 us even with every store removed -- one warp per scheduler
        // running a dependent ALU chain -- and a warp can only issue ~8 B/clk of 16-byte stores.)
        const int q = warp & 3;
        const int half = warp < 4 ? 1 : 0;
        const int prow = q * 32 + lane;  // pixel row of this thread inside the 128-row tile
        const float b3 = (flags & F_DOTSIG) ? __ldg(tp->b3) : 0.f;
        mbar_wait(smem_u32(&hdr->tmem_full), n_conv & 1);
        tc_fence_after();
        if (kTrace && tid == 128) { trace[idx * kTraceW + 11] = gtime(); trace[idx * kTraceW + 20] = clock64(); }
        int acc_slot = 0;
        // (deliberately NOT unrolled over the samples: the kernel's code footprint decides how well it runs)
#pragma unroll 1
        for (int s = 0; s < n_samp; ++s) {
          float* const outp = tp->out[s];
          const float* const aux_ss = tp->aux[s];
          float* const map_ss = tp->map_out[s];
          void* const in1_ss = const_cast<void*>(tp->in[1][s]);
          const float* const auxp = (flags & F_MASK) ? aux_ss : outp;  // F_MASK and F_ACCUM never combine
          uint8_t* const hb = reinterpret_cast<uint8_t*>(outp) + shadow_bytes(P_out);
          for (int mt = mt0; mt < mt0 + n_mt && acc_slot < kExMaxAcc; ++mt, ++acc_slot) {
            const int r = mt * 128 + prow;
            const int y = r / S_in, x = r - y * S_in;
            const bool valid = (y < kHW) && (x < kHW);
            const int so = y * S_out + x;
            const int sx = (flags & F_MASK) ? y * S_aux + x : so;
            const int Px = (flags & F_MASK) ? P_aux : P_out;
            float dot = 0.f;
            // lean tasks: the output only feeds other convs (and wgrad / a ReLU mask): fp16 planes alone
            const bool lean = (flags & F_HALF) && !(flags & (F_STORE | F_DOTSIG | F_ATTBWD | F_ACCUM | F_MASK));
            // 16 accumulator columns per pass (4 fp32 planes / 2 fp16 half planes): everything a pass keeps live fits in
            // registers -- a spilled value costs an L2 round trip here, because the other CTA's acquire / release
            // traffic keeps invalidating the SM's L1
#pragma unroll 1
            for (int chunk = 4 * half; chunk < 4 * half + 4; ++chunk) {
              uint32_t v[16];
              tmem_ld16(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc_slot * 128 + chunk * 16, v);
              if (lean) {
                uint4 mh[2];
                if (valid && (flags & F_MASK16)) {
                  const uint8_t* mb = reinterpret_cast<const uint8_t*>(aux_ss) + (static_cast<size_t>(chunk * 2) * P_aux + (y * S_aux + x)) * 16;
                  mh[0] = ldg128(mb);
                  mh[1] = ldg128(mb + static_cast<size_t>(P_aux) * 16);
                }
                tmem_ld_wait();
                if (kTrace && tid == 128 && acc_slot == 0 && chunk == 0) trace[idx * kTraceW + 16] = clock64();
                if (valid) {
                  if (flags & F_BIAS) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                      const float4 bz = *reinterpret_cast<const float4*>(&hdr->bias[chunk * 16 + j * 4]);
                      v[4 * j] = __float_as_uint(__uint_as_float(v[4 * j]) + bz.x);
                      v[4 * j + 1] = __float_as_uint(__uint_as_float(v[4 * j + 1]) + bz.y);
                      v[4 * j + 2] = __float_as_uint(__uint_as_float(v[4 * j + 2]) + bz.z);
                      v[4 * j + 3] = __float_as_uint(__uint_as_float(v[4 * j + 3]) + bz.w);
                    }
                  }
                  uint8_t* ob = hb + (static_cast<size_t>(chunk * 2) * P_out + so) * 16;
#pragma unroll
                  for (int j = 0; j < 2; ++j) {
                    uint4 h;
                    if (flags & F_RELU) {
                      h = make_uint4(pack_half2_relu_sat(__uint_as_float(v[8 * j]), __uint_as_float(v[8 * j + 1])),
                                     pack_half2_relu_sat(__uint_as_float(v[8 * j + 2]), __uint_as_float(v[8 * j + 3])),
                                     pack_half2_relu_sat(__uint_as_float(v[8 * j + 4]), __uint_as_float(v[8 * j + 5])),
                                     pack_half2_relu_sat(__uint_as_float(v[8 * j + 6]), __uint_as_float(v[8 * j + 7])));
                    } else {
                      h = make_uint4(pack_half2_sat(__uint_as_float(v[8 * j]), __uint_as_float(v[8 * j + 1])),
                                     pack_half2_sat(__uint_as_float(v[8 * j + 2]), __uint_as_float(v[8 * j + 3])),
                                     pack_half2_sat(__uint_as_float(v[8 * j + 4]), __uint_as_float(v[8 * j + 5])),
                                     pack_half2_sat(__uint_as_float(v[8 * j + 6]), __uint_as_float(v[8 * j + 7])));
                    }
                    if (flags & F_MASK16) {
                      h.x = mask_half2_gt0(h.x, mh[j].x); h.y = mask_half2_gt0(h.y, mh[j].y);
                      h.z = mask_half2_gt0(h.z, mh[j].z); h.w = mask_half2_gt0(h.w, mh[j].w);
                    }
                    if (!(dbg & 2)) stg128(ob + static_cast<size_t>(j) * P_out * 16, h);
                  }
                }
                if (kTrace && tid == 128 && acc_slot == 0 && chunk == 0) trace[idx * kTraceW + 17] = clock64();
                continue;
              }
              float4 ax[4];
              if (valid && (flags & (F_MASK | F_ACCUM))) {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                  ax[j] = ld4(auxp + (static_cast<size_t>(chunk * 4 + j) * Px + sx) * 4);
              }
              float4 fx[4];
              if (valid && (flags & F_ATTBWD)) {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                  fx[j] = ld4(aux_ss + (static_cast<size_t>(chunk * 4 + j) * 256 + (y * 16 + x)) * 4);
              }
              tmem_ld_wait();
              if (kTrace && tid == 128 && acc_slot == 0 && chunk == 0) trace[idx * kTraceW + 16] = clock64();
              if (valid && (flags & F_ATTBWD)) {
                // g = this conv's output (gradient w.r.t. feat * map): dmap += <g, feat>, dfeat (+)= g * map
                const float mval = map_ss[y * 16 + x];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const size_t off = (static_cast<size_t>(chunk * 4 + j) * 256 + (y * 16 + x)) * 4;
                  const float4 f = fx[j];
                  const float gx = __uint_as_float(v[4 * j]), gy = __uint_as_float(v[4 * j + 1]);
                  const float gz = __uint_as_float(v[4 * j + 2]), gw = __uint_as_float(v[4 * j + 3]);
                  dot = fmaf(gx, f.x, dot); dot = fmaf(gy, f.y, dot); dot = fmaf(gz, f.z, dot); dot = fmaf(gw, f.w, dot);
                  float4 o = make_float4(gx * mval, gy * mval, gz * mval, gw * mval);
                  if (flags & F_ACCUM) { o.x += ax[j].x; o.y += ax[j].y; o.z += ax[j].z; o.w += ax[j].w; }
                  st4(outp + off, o);
                }
              } else if (valid) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const int n0 = (chunk * 4 + j) * 4;
                  const float4 bz = *reinterpret_cast<const float4*>(&hdr->bias[n0]);
                  float4 o = make_float4(__uint_as_float(v[4 * j]) + bz.x, __uint_as_float(v[4 * j + 1]) + bz.y,
                                         __uint_as_float(v[4 * j + 2]) + bz.z, __uint_as_float(v[4 * j + 3]) + bz.w);
                  if (flags & F_RELU) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                  if (flags & F_MASK) {
                    o.x = ax[j].x > 0.f ? o.x : 0.f; o.y = ax[j].y > 0.f ? o.y : 0.f;
                    o.z = ax[j].z > 0.f ? o.z : 0.f; o.w = ax[j].w > 0.f ? o.w : 0.f;
                  }
                  if (flags & F_ACCUM) { o.x += ax[j].x; o.y += ax[j].y; o.z += ax[j].z; o.w += ax[j].w; }
                  if (flags & F_DOTSIG) {
                    const float4 wz = *reinterpret_cast<const float4*>(&hdr->w3[n0]);
                    dot = fmaf(o.x, wz.x, dot); dot = fmaf(o.y, wz.y, dot); dot = fmaf(o.z, wz.z, dot); dot = fmaf(o.w, wz.w, dot);
                  }
                  v[4 * j] = round_tf32(o.x); v[4 * j + 1] = round_tf32(o.y);
                  v[4 * j + 2] = round_tf32(o.z); v[4 * j + 3] = round_tf32(o.w);
                }
                if ((flags & F_STORE) && !(dbg & 1)) {
#pragma unroll
                  for (int j = 0; j < 4; ++j)
                    stg128(outp + (static_cast<size_t>(chunk * 4 + j) * P_out + so) * 4,
                           make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]));
                }
                if ((flags & F_HALF) && !(dbg & 2)) {
#pragma unroll
                  for (int j = 0; j < 2; ++j) {
                    const uint4 h = make_uint4(
                        pack_half2_sat(__uint_as_float(v[8 * j]), __uint_as_float(v[8 * j + 1])),
                        pack_half2_sat(__uint_as_float(v[8 * j + 2]), __uint_as_float(v[8 * j + 3])),
                        pack_half2_sat(__uint_as_float(v[8 * j + 4]), __uint_as_float(v[8 * j + 5])),
                        pack_half2_sat(__uint_as_float(v[8 * j + 6]), __uint_as_float(v[8 * j + 7])));
                    stg128(hb + (static_cast<size_t>(chunk * 2 + j) * P_out + so) * 16, h);
                  }
                }
              }
              if (kTrace && tid == 128 && acc_slot == 0 && chunk == 0) trace[idx * kTraceW + 17] = clock64();
            }
            if (kTrace && tid == 128 && acc_slot == 0) trace[idx * kTraceW + 18] = clock64();
            if (flags & (F_ATTBWD | F_DOTSIG)) {
              // the two column halves of a pixel row live in two warps: combine their 64-channel partial dot products
              hdr->dotp[acc_slot][half][prow] = dot;
              __syncthreads();
              dot = hdr->dotp[acc_slot][0][prow] + hdr->dotp[acc_slot][1][prow];
            }
            if ((flags & F_ATTBWD) && valid && half == 0) {
              float* dmap = reinterpret_cast<float*>(in1_ss);
              dmap[y * 16 + x] += dot;   // one writer per pixel: a sample's backward chain is serial
            }
            // the whole 128-channel dot product of the pixel: 1x1 head + sigmoid (nmn_modules.py:86,167)
            if ((flags & F_DOTSIG) && valid) {
              const float m = 1.f / (1.f + expf(-(dot + b3)));
              if (half == 0) map_ss[y * 16 + x] = m;
              if ((flags & F_ATTEND) && !(dbg & 4)) {
                // the next module's first conv reads feat * map: write its fp16 operand planes right here
                // (this thread: 8 of the 16 half planes; all 16 feature loads are issued before the first use)
                const float* __restrict__ feat = aux_ss;
                uint8_t* __restrict__ x0h = reinterpret_cast<uint8_t*>(in1_ss);
                const int s16 = y * 16 + x;
#pragma unroll 1
                for (int b4 = 0; b4 < 2; ++b4) {
                  float4 f[8];
#pragma unroll
                  for (int i = 0; i < 8; ++i)
                    f[i] = ld4(feat + (static_cast<size_t>(16 * half + 8 * b4 + i) * 256 + s16) * 4);
#pragma unroll
                  for (int i = 0; i < 4; ++i) {
                    const float4 f0 = f[2 * i], f1 = f[2 * i + 1];
                    const uint4 h = make_uint4(pack_half2_sat(f0.x * m, f0.y * m), pack_half2_sat(f0.z * m, f0.w * m),
                                               pack_half2_sat(f1.x * m, f1.y * m), pack_half2_sat(f1.z * m, f1.w * m));
                    stg128(x0h + (static_cast<size_t>(8 * half + 4 * b4 + i) * 256 + s16) * 16, h);
                  }
                }
              }
            }
            if (kTrace && tid == 128 && acc_slot == 0) trace[idx * kTraceW + 19] = clock64();
          }
        }
        if (kTrace && tid == 128) trace[idx * kTraceW + 12] = gtime();
        tc_fence_before();
        if (kTrace && tid == 128) trace[idx * kTraceW + 13] = gtime();
      }
      if (kTrace && tid == 0) {
        trace[idx * kTraceW + 5] = TASK_CONV | (n_samp << 8) | (n_mt << 16);
        trace[idx * kTraceW + 6] = n_kb * ntaps;  // one K=16 MMA per (k-block, tap, sample, M tile)
        trace[idx * kTraceW + 7] = flags;
      }
      ++n_conv;
    } else {
      // CUDA-core task: all 256 threads of the CTA (the tensor-core roles have nothing to do meanwhile)
      if (warp == 0) wait_deps();
      __syncthreads();  // (B)
      const EltTask& t = *reinterpret_cast<const EltTask*>(hdr->task);
#ifndef PNMN_NO_ELT  // (timing experiment: how much of the conv path's time is instruction-cache pressure from this code?)
      elt_task_body(t, tid, hdr->elt);
#endif
      if (kTrace && tid == 128) { trace[idx * kTraceW + 5] = TASK_ELT; trace[idx * kTraceW + 6] = 0; trace[idx * kTraceW + 7] = t.op; }
    }
    if (kTrace && tid == 0) trace[idx * kTraceW + 2] = gtime();
    __syncthreads();  // (C) every store of this task has been issued; the release below is cumulative over them (bar.sync
                      // orders the CTA's writes before thread 0's st.release.gpu -- the cutlass::Semaphore::release pattern)
    if (tid == 0) {
      st_release(done + idx, 1);
      if (trace) trace[idx * kTraceW + 3] = gtime();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<kExMaxAcc * 128>(tmem_base);
}

static int g_reserved_sms = std::getenv("PNMN_RESERVE_SMS") ? std::atoi(std::getenv("PNMN_RESERVE_SMS")) : 0;
void set_reserved_sms(int n) { g_reserved_sms = n < 0 ? 0 : n; }
int reserved_sms() { return g_reserved_sms; }

cudaError_t launch_exec(const uint8_t* d_tasks, const TaskMeta* d_meta, int n_tasks, const ConvCfg* d_cfgs,
                        int* d_counter, int* d_done, long long* d_trace, int max_ctas, cudaStream_t stream) {
  if (n_tasks <= 0) return cudaSuccess;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(exec_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kExSmem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(exec_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kExSmem);
    if (e != cudaSuccess) return e;
    attr_done = true;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  // PNMN_EXEC_CTAS caps the persistent grid (default: two CTAs per SM).  With fewer CTAs than 2 x SMs some SMs keep enough
  // shared memory free for the kernels of OTHER streams (the LSTM passes of the joint-training step) to run alongside.
  static const int cap = std::getenv("PNMN_EXEC_CTAS") ? std::atoi(std::getenv("PNMN_EXEC_CTAS")) : 0;
  int grid = n_tasks < kExCtasPerSM * sms ? n_tasks : kExCtasPerSM * sms;
  if (cap > 0 && grid > cap) grid = cap;
  if (max_ctas > 0 && grid > max_ctas) grid = max_ctas;
  static const int dbg = std::getenv("PNMN_EXEC_DBG") ? std::atoi(std::getenv("PNMN_EXEC_DBG")) : 0;  // timing experiments only
  // at least a quarter of the SMs always stay with the executor
  const int sm_limit = sms - (g_reserved_sms < sms * 3 / 4 ? g_reserved_sms : sms * 3 / 4);
  if (d_trace) exec_kernel<true><<<grid, kExThreads, kExSmem, stream>>>(d_tasks, d_meta, n_tasks, d_cfgs, d_counter, d_done, d_trace, dbg, sm_limit);
  else exec_kernel<false><<<grid, kExThreads, kExSmem, stream>>>(d_tasks, d_meta, n_tasks, d_cfgs, d_counter, d_done, d_trace, dbg, sm_limit);
  return cudaGetLastError();
}

}  // namespace pnmn
