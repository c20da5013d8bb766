// Host side of the NMN module executor: model description, program compiler, level scheduler and
// the C ABI (include/pnmn.h).
//
// The reference walks every sample's program in python, one tiny batch-1 kernel at a time, with a
// device->host sync per sample (probnmn/models/nmn.py:191-238).  Here the whole batch is compiled
// once on the host into task tables: every sample becomes a chain of stages (elementwise op, conv
// on the tensor cores), chains are aligned into global steps, conv stages of the same step that
// use the SAME weights are paired into one CTA (shared weight stream), and each step is at most
// three launches.  The backward chain is derived from the same symbolic execution; weight
// gradients are batched per weight tensor over all its instances in the batch.
#include <algorithm>
#include <chrono>
#include <array>
#include <cstdio>
#include <cstdlib>
#include <cstddef>
#include <cstring>
#include <malloc.h>
#include <map>
#include <string>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/pnmn.h"
#include "exec.h"
#include "layout.h"

using namespace pnmn;

namespace {

thread_local std::string g_err;
int fail(const std::string& s) {
  g_err = s;
  return 1;
}
#define CUDA_OK(x)                                                                  \
  do {                                                                              \
    cudaError_t e_ = (x);                                                           \
    if (e_ != cudaSuccess) return fail(std::string(#x) + ": " + cudaGetErrorString(e_)); \
  } while (0)

// ---- symbolic pointers: (arena id << 56) | byte offset, resolved at launch time ----------------
enum ArenaId : int {
  AR_NULL = 0, AR_A16, AR_A18, AR_A22, AR_MAPS, AR_DMAPS, AR_IDX, AR_PACKED, AR_PARAMS, AR_GRADS,
  AR_BLOB, AR_FINAL, AR_GRADOUT, AR_AIN, AR_SCRATCH, AR_COUNT
};
template <class T>
T* sym(int arena, int64_t byte_off) {
  return reinterpret_cast<T*>((static_cast<uint64_t>(arena) << 56) | static_cast<uint64_t>(byte_off));
}
template <class T>
void resolve(T*& p, const uint64_t* base) {
  const uint64_t v = reinterpret_cast<uint64_t>(p);
  if (v == 0) return;
  p = reinterpret_cast<T*>(base[v >> 56] + (v & ((1ull << 56) - 1)));
}

constexpr int64_t kGuard = 16384;  // zero bytes in front of unit 0 of every plane arena
// arena unit = 32 fp32 planes followed by their fp16 shadow (16 half planes), see executor.h
constexpr int64_t kUnit16 = 48ll * 256 * 16;
constexpr int64_t kUnit18 = 48ll * 324 * 16;
constexpr int64_t kUnit22 = 48ll * 484 * 16;
constexpr int kInstChunk = 16;     // instances per wgrad CTA
constexpr int kNumSMs = 148;
constexpr int kEltParts = 4;
constexpr int kBiasSplit = 32;

struct ConvW {
  int64_t w_off = -1, b_off = -1;  // floats into the flat parameter buffer
  int cin = 128, ksize = 3;
  int64_t pk_fwd = -1, pk_bwd = -1;  // halves into the packed (fp16) weight-tile buffer
};
struct ModuleDesc {
  int kind = PNMN_TOK_SKIP;
  std::vector<int> convs;          // ids into pnmn_model::convs (MMA convs, forward order)
  int64_t head_w = -1, head_b = -1;  // 1x1 -> 1 head (Attention/Relate) or SameModule.conv
};

}  // namespace

struct pnmn_model {
  int V = 0, in_ch = 1024;
  std::vector<ModuleDesc> mods;
  std::vector<ConvW> convs;
  int stem1 = -1, stem2 = -1;
  std::vector<PackTask> pack;
  int total_tiles = 0;
  int64_t packed_floats = 0;
};

namespace {

int add_conv_w(pnmn_model* m, int64_t w, int64_t b, int cin, int ksize, bool need_bwd) {
  ConvW c;
  c.w_off = w; c.b_off = b; c.cin = cin; c.ksize = ksize;
  const int kk = ksize * ksize;
  auto add_pack = [&](int n_kb, int flip, int k_off, int n_off, int ks, int ns, int ts) {
    PackTask t{};
    t.src_off = w; t.dst_off = m->packed_floats; t.first_tile = m->total_tiles;
    t.n_kb = n_kb; t.ntaps = kk; t.flip = flip; t.k_off = k_off; t.n_off = n_off;
    t.k_stride = ks; t.n_stride = ns; t.tap_stride = ts;
    m->pack.push_back(t);
    const int64_t off = m->packed_floats;
    m->total_tiles += n_kb * kk;
    m->packed_floats += static_cast<int64_t>(n_kb) * kk * 2048;
    return off;
  };
  // forward: k = cin, n = cout.  W[n][k][tap]
  c.pk_fwd = add_pack(cin / 16, 0, 0, 0, kk, cin * kk, kk == 1 ? 0 : 1);
  if (need_bwd) {
    // dgrad: k = cout (128), n = cin tile; one 8-k-block pack per 128 input channels
    c.pk_bwd = m->packed_floats;
    for (int h = 0; h < cin / 128; ++h) add_pack(8, 1, 0, 128 * h, cin * kk, kk, kk == 1 ? 0 : 1);
  }
  m->convs.push_back(c);
  return static_cast<int>(m->convs.size()) - 1;
}

}  // namespace

namespace pnmn {
void set_last_error(const std::string& s) { g_err = s; }
static long long g_launch_count = 0;   // kernels launched by the library (bench.py "gpu_launches")
void count_launches(int n) { g_launch_count += n; }
}
namespace pnmn { void set_reserved_sms(int n); int reserved_sms(); }
extern "C" int pnmn_set_reserved_sms(int n) {
  const int prev = pnmn::reserved_sms();
  pnmn::set_reserved_sms(n);
  return prev;
}

extern "C" long long pnmn_launch_count(int reset) {
  const long long v = pnmn::g_launch_count;
  if (reset) pnmn::g_launch_count = 0;
  return v;
}

extern "C" int pnmn_version(void) { return PNMN_VERSION; }
extern "C" int pnmn_has_bringup_kernels(void) {
#ifdef PNMN_BRINGUP
  return 1;
#else
  return 0;
#endif
}
extern "C" const char* pnmn_last_error(void) { return g_err.c_str(); }

extern "C" pnmn_model* pnmn_model_create(int vocab_size, const int32_t* token_kind,
                                         const int64_t* token_param_off, const int64_t* stem_param_off,
                                         int in_channels) {
  if (in_channels % 128 != 0) { g_err = "in_channels must be a multiple of 128"; return nullptr; }
  // The program compiler builds several MB of task tables per forward.  glibc serves such blocks with mmap and returns
  // them on free, so every plan would page-fault its tables in again (~half of pnmn_plan_create's time); keep them on
  // the heap instead.
  static const bool heap_tuned = [] {
    mallopt(M_MMAP_THRESHOLD, 1 << 30);
    mallopt(M_TRIM_THRESHOLD, 1 << 30);
    return true;
  }();
  (void)heap_tuned;
  auto* m = new pnmn_model();
  m->V = vocab_size;
  m->in_ch = in_channels;
  m->mods.resize(vocab_size);
  m->stem1 = add_conv_w(m, stem_param_off[0], stem_param_off[1], in_channels, 3, false);
  m->stem2 = add_conv_w(m, stem_param_off[2], stem_param_off[3], 128, 3, true);
  for (int v = 0; v < vocab_size; ++v) {
    ModuleDesc& d = m->mods[v];
    d.kind = token_kind[v];
    const int64_t* po = token_param_off + static_cast<size_t>(v) * PNMN_MAX_MODULE_PARAMS;
    auto conv3 = [&](int i) { d.convs.push_back(add_conv_w(m, po[2 * i], po[2 * i + 1], 128, 3, true)); };
    switch (d.kind) {
      case PNMN_TOK_ATTENTION: conv3(0); conv3(1); d.head_w = po[4]; d.head_b = po[5]; break;
      case PNMN_TOK_QUERY: conv3(0); conv3(1); break;
      case PNMN_TOK_RELATE: for (int i = 0; i < 5; ++i) conv3(i); d.head_w = po[10]; d.head_b = po[11]; break;
      case PNMN_TOK_SAME: d.head_w = po[0]; d.head_b = po[1]; break;
      case PNMN_TOK_COMPARE:
        d.convs.push_back(add_conv_w(m, po[0], po[1], 256, 1, true));
        conv3(1); conv3(2);
        break;
      default: break;
    }
  }
  return m;
}
extern "C" void pnmn_model_destroy(pnmn_model* m) { delete m; }
extern "C" int64_t pnmn_model_packed_floats(const pnmn_model* m) { return (m->packed_floats + 1) / 2; }  // fp16 tiles

// =================================================================================================
// Plan
// =================================================================================================
namespace {

enum ValKind { VK_FEAT, VK_ONES, VK_MAP, VK_BUF };
struct Val {
  int kind = VK_BUF;
  int ch = 128;
  int unit = -1;        // A16 unit (FEAT/BUF) or map index (MAP/ONES)
  bool relu_out = false;
  int ncons = 0;        // live consumers (backward)
  int gunit = -1;       // A16 unit of the gradient (128-ch values)
  bool gwritten = false;
  bool live = false;
  bool cons_minmax = false;
  int dotsig_proto = -1;   // forward conv proto (index into Sched::protos) whose fused 1x1 head produced this map
  bool attend_fused = false;
  int strand = -1;         // forward strand (Sched) that produces this value; -1 = the constant all-ones map
};
struct OpRec {
  int kind = 0, tok = 0;
  int in0 = -1, in1 = -1, out = -1;  // value ids
  bool x0_is_feat = false;
  int x0_unit = -1;                  // attended features (A16)
  int nconv = 0;
  int y_unit[5] = {-1, -1, -1, -1, -1};  // conv outputs (arena of the NEXT conv's input format)
  int idx_slot = -1;
  bool attend_fused = false;  // feat * map is produced by the previous module's head conv (F_ATTEND), no elementwise task
  int strand = -1;            // forward strand of this op's stages
  int secondary = -1;         // binary ops: forward strand of the other input when it differs (joined here), else -1
};

enum LaunchKind { LK_ELT = 0, LK_CONV0 = 1, LK_CONV1 = 2, LK_WGRAD = 3, LK_BIAS = 4 };
struct LaunchItem { int kind; int64_t off; int count; };

struct ConvProto {
  ConvTask t;
  int sample;
};

struct TaskRec { uint8_t b[128]; };

// Stages are placed on STRANDS.  A strand is a serial chain of stages (every stage waits for all tasks of the strand's
// previous stage).  A sample starts with one strand (the stem); with the persistent executor the independent sub-chains of
// a program -- everything between a `scene` token and the binary module that consumes it (nmn.py:216-222: `scene` parks the
// running output in `saved` and restarts from the all-ones attention) -- get their own strand: fork() starts a strand
// behind the current stage of its parent, join() makes the next stage of a strand wait for another strand as well.  The
// two branches of a comparison / union / intersection then run concurrently instead of back to back, which shortens the
// longest dependency chain of the batch -- the quantity that bounds the executor's span (scripts/sim_sched.py).
struct Sched {
  std::vector<int> step, last;                // per strand
  std::vector<int> fork_parent, fork_step;    // per strand: parent (-1 = root) and the step of its first stage
  std::vector<int> strand_sample;             // per strand: the sample it belongs to
  std::vector<uint8_t> critical;              // per strand: its sample has one of the longest chains (mark_critical)
  struct Join { int step, primary, secondary; };
  std::vector<Join> joins;
  std::vector<std::array<std::vector<int>, 3>> buckets;  // indices into elts / protos
  std::vector<EltTask> elts;
  std::vector<int> elt_sample;                // strand of every elementwise task
  std::vector<ConvProto> protos;              // ConvProto::sample = strand
  bool unified;  // persistent executor: one stage per strand per step (dependencies are explicit)
  bool dep_overflow = false;
  Sched(int B, bool uni) : unified(uni) {
    step.reserve(static_cast<size_t>(B) * 3); last.reserve(static_cast<size_t>(B) * 3);
    fork_parent.reserve(static_cast<size_t>(B) * 3); fork_step.reserve(static_cast<size_t>(B) * 3);
    elts.reserve(static_cast<size_t>(B) * 64);
    elt_sample.reserve(static_cast<size_t>(B) * 64);
    protos.reserve(static_cast<size_t>(B) * 48);
    buckets.reserve(128);
  }
  int new_strand(int sample) {
    step.push_back(0); last.push_back(-1); fork_parent.push_back(-1); fork_step.push_back(0);
    strand_sample.push_back(sample);
    return static_cast<int>(step.size()) - 1;
  }
  // Samples whose chain is long enough to bound the pass: their convolutions are never paired with another sample and
  // always split over the M tiles (two CTAs, ~11 us per stage instead of ~20 us for a paired task), everything else is
  // grouped for throughput.  A sample is critical when its last stage lies beyond `frac` of the batch's deepest one.
  void mark_critical(int n_samples, float frac) {
    critical.assign(step.size(), 0);
    if (!unified || frac >= 1.f) return;
    std::vector<int> end(n_samples, 0);
    int deepest = 0;
    for (size_t s = 0; s < step.size(); ++s) {
      const int e = last[s] >= 0 ? step[s] + 1 : 0;
      end[strand_sample[s]] = std::max(end[strand_sample[s]], e);
      deepest = std::max(deepest, e);
    }
    const int thr = std::max(8, static_cast<int>(frac * static_cast<float>(deepest)));
    for (size_t s = 0; s < step.size(); ++s) critical[s] = end[strand_sample[s]] >= thr;
  }
  int next_step(int s) const { return last[s] >= 0 ? step[s] + 1 : step[s]; }
  // a strand whose first stage waits for the parent's CURRENT stage (later stages of the parent do not matter to it)
  int fork(int parent) {
    const int start = next_step(parent);
    const int s = new_strand(strand_sample[parent]);
    step[s] = start; fork_parent[s] = parent; fork_step[s] = start;
    return s;
  }
  // the next stage placed on `primary` also waits for the current stage of `secondary`
  void join(int primary, int secondary) {
    const int st = std::max(next_step(primary), next_step(secondary));
    step[primary] = last[primary] >= 0 ? st - 1 : st;
    joins.push_back(Join{st, primary, secondary});
  }
  int place(int s, int kind) {
    if (unified ? last[s] >= 0 : kind <= last[s]) step[s]++;
    last[s] = kind;
    if (static_cast<int>(buckets.size()) <= step[s]) buckets.resize(step[s] + 1);
    return step[s];
  }
  void add_elt(int s, const EltTask& t) {
    const int st = place(s, LK_ELT);
    // plane-parallel ops are split over 4 CTAs: shorter critical path, more memory parallelism
    const bool splittable = t.op == OP_ATTEND || t.op == OP_ATTEND_BWD || t.op == OP_RELU_MASK || t.op == OP_SCATTER ||
                            t.op == OP_GATHER || t.op == OP_DOTSIG_BWD;
    const int n_parts = splittable ? kEltParts : 1;
    for (int part = 0; part < n_parts; ++part) {
      EltTask e = t;
      e.part = part; e.n_parts = n_parts;
      buckets[st][LK_ELT].push_back(static_cast<int>(elts.size()));
      elts.push_back(e);
      elt_sample.push_back(s);
    }
  }
  int add_conv(int s, const ConvTask& t, int variant) {
    const int kind = variant == 0 ? LK_CONV0 : LK_CONV1;
    const int st = place(s, kind);
    buckets[st][kind].push_back(static_cast<int>(protos.size()));
    protos.push_back(ConvProto{t, s});
    return static_cast<int>(protos.size()) - 1;
  }
  // Conv tasks of one step: pair samples that share (cfg, weights) on wide levels, split M tiles on narrow
  // ones.  Calls emit(task, samples[], n_samples) for every CTA-level task.
  template <class Emit>
  void group_convs(std::vector<int>& v, int kind, Emit emit) {
    if (v.empty()) return;
    // stable order by (cfg, weights): sort packed 64-bit keys (cfg | weight offset / 16 | position) instead of chasing
    // the 136-byte prototypes in a comparator -- this sort was a quarter of pnmn_plan_create
    {
      std::vector<uint64_t> keys(v.size());
      for (size_t i = 0; i < v.size(); ++i) {
        const ConvTask& a = protos[v[i]].t;
        const uint64_t w16 = (reinterpret_cast<uint64_t>(a.w) & ((1ull << 36) - 1)) >> 4;  // symbolic pointer: offset in the packed-weight arena
        keys[i] = static_cast<uint64_t>(a.cfg) << 56 | (w16 & 0xFFFFFFFull) << 28 | static_cast<uint64_t>(i);
      }
      std::sort(keys.begin(), keys.end());
      std::vector<int> sorted(v.size());
      for (size_t i = 0; i < v.size(); ++i) sorted[i] = v[keys[i] & 0xFFFFFFFull];
      v.swap(sorted);
    }
    // Task granularity adapts to the level's width (one CTA per SM, 148 SMs): wide levels pair two
    // samples per CTA (shared weight stream), narrow levels split a sample's M tiles over several CTAs.
    const int U = static_cast<int>(v.size());
    const int nmt = kind == LK_CONV0 ? 2 : 3;
    static const bool pair_always = std::getenv("PNMN_PAIR_ALWAYS") != nullptr;  // diagnostics
    static const int pair_min = std::getenv("PNMN_PAIR_MIN") ? std::atoi(std::getenv("PNMN_PAIR_MIN")) : kNumSMs;
    static const int split_max = std::getenv("PNMN_SPLIT_MAX") ? std::atoi(std::getenv("PNMN_SPLIT_MAX")) : kNumSMs + kNumSMs / 2;
    const int cap = (kind == LK_CONV0 && (U > pair_min || pair_always)) ? NSMAX : 1;
    static const bool no_split = std::getenv("PNMN_NOSPLIT") != nullptr;  // diagnostics
    const int split = (!no_split && U * nmt <= split_max) ? nmt : 1;
    const bool have_crit = !critical.empty();
    size_t i = 0;
    while (i < v.size()) {
      ConvTask t = protos[v[i]].t;
      int samples[NSMAX] = {protos[v[i]].sample, -1};
      t.n_samp = 1;
      const bool crit = have_crit && critical[samples[0]];
      size_t j = i + 1;
      while (j < v.size() && !crit && t.n_samp < cap && protos[v[j]].t.cfg == t.cfg && protos[v[j]].t.w == t.w &&
             !(have_crit && critical[protos[v[j]].sample])) {
        const ConvTask& o = protos[v[j]].t;
        const int k = t.n_samp++;
        samples[k] = protos[v[j]].sample;
        t.in[0][k] = o.in[0][0]; t.in[1][k] = o.in[1][0];
        t.out[k] = o.out[0]; t.aux[k] = o.aux[0]; t.map_out[k] = o.map_out[0];
        ++j;
      }
      // a task owns at most TWO accumulators (two executor CTAs share an SM's 512 TMEM columns, exec.cu)
      if (split > 1 || t.n_samp > 1 || crit) {
        for (int m = 0; m < nmt; ++m) { t.mt0 = m; t.n_mt = 1; emit(t, samples, t.n_samp); }
      } else {
        for (int m = 0; m < nmt; m += 2) { t.mt0 = m; t.n_mt = std::min(2, nmt - m); emit(t, samples, t.n_samp); }
      }
      i = j;
    }
  }

  // Level-synchronous order (one launch per kind per step).
  void flatten(std::vector<EltTask>& out_elt, std::vector<ConvTask>& out_conv, std::vector<LaunchItem>& launches) {
    for (auto& b : buckets) {
      if (!b[LK_ELT].empty()) {
        launches.push_back({LK_ELT, static_cast<int64_t>(out_elt.size()), static_cast<int>(b[LK_ELT].size())});
        for (int i : b[LK_ELT]) out_elt.push_back(elts[i]);
      }
      for (int kind = LK_CONV0; kind <= LK_CONV1; ++kind) {
        if (b[kind].empty()) continue;
        const int64_t off = static_cast<int64_t>(out_conv.size());
        group_convs(b[kind], kind, [&](const ConvTask& t, const int*, int) { out_conv.push_back(t); });
        launches.push_back({kind, off, static_cast<int>(out_conv.size() - off)});
      }
    }
  }

  // Persistent-executor order: one list, every task names the tasks of its strands' previous stage (plus the stages that
  // joins and forks bring in).  The list is then sorted by each task's REMAINING CHAIN TIME, longest first (the classic
  // critical-path list-scheduling priority; a producer's remaining time exceeds its consumers', so producers stay in
  // front and the executor's in-order fetch stays deadlock-free).  With the step-aligned order the long chains' stages
  // sat behind whole levels of short-chain work; replaying the recorded trace (scripts/sim_sched.py) puts the forward
  // pass of the bench workload at 767 us step-aligned vs 704 us in this order (critical path 695 us), the backward at
  // 839 vs 754 us.  PNMN_LIST_ORDER=level keeps the step-aligned list.  Durations come from a cost model fitted to the
  // trace (profiles/r1b_summary.md).
  void flatten_persistent(std::vector<TaskRec>& out, std::vector<TaskMeta>& meta, const std::vector<ConvCfg>& cfgs) {
    static const bool cp_order = [] { const char* e = std::getenv("PNMN_LIST_ORDER"); return !(e && std::string(e) == "level"); }();
    struct Latest { int n = 0; int ids[8] = {0, 0, 0, 0, 0, 0, 0, 0}; int stamp = -1; };
    std::vector<Latest> latest(step.size()), next(step.size());
    // forks / joins by the step at which they take effect
    std::vector<std::vector<int>> forks_at(buckets.size()), joins_at(buckets.size());
    for (size_t s = 0; s < step.size(); ++s)
      if (fork_parent[s] >= 0 && fork_step[s] < static_cast<int>(buckets.size())) forks_at[fork_step[s]].push_back(static_cast<int>(s));
    for (size_t j = 0; j < joins.size(); ++j)
      if (joins[j].step < static_cast<int>(buckets.size())) joins_at[joins[j].step].push_back(static_cast<int>(j));
    std::vector<float> dur;
    size_t total = elts.size();
    for (auto& b : buckets) total += b[LK_CONV0].size() * 2 + b[LK_CONV1].size() * 3;
    out.reserve(total);
    meta.reserve(total);
    dur.reserve(total);
    int cur = 0;
    static const bool no_deps = std::getenv("PNMN_NODEPS") != nullptr;  // diagnostics: throughput without dependencies (results are garbage)
    auto push = [&](const void* rec, int type, const int* strands, int ns, float d) {
      out.push_back(*static_cast<const TaskRec*>(rec));
      meta.emplace_back();
      dur.push_back(d);
      TaskMeta& m = meta.back();
      m.type = type; m.n_deps = 0;
      for (int k = 0; k < kMaxDeps; ++k) m.deps[k] = -1;
      const int id = static_cast<int>(out.size()) - 1;
      for (int k = 0; k < ns; ++k) {
        const Latest& l = latest[strands[k]];
        for (int j = 0; j < (no_deps ? 0 : l.n); ++j) {
          bool dup = false;  // two strands forked from the same stage share their producers
          for (int q = 0; q < m.n_deps; ++q) dup = dup || m.deps[q] == l.ids[j];
          if (dup) continue;
          if (m.n_deps < kMaxDeps) m.deps[m.n_deps++] = l.ids[j];
          else dep_overflow = true;
        }
        Latest& nx = next[strands[k]];
        if (nx.stamp != cur) { nx.stamp = cur; nx.n = 0; }
        if (nx.n < 8) nx.ids[nx.n++] = id;
        else dep_overflow = true;
      }
    };
    // body + publish times in us, fitted to the per-task trace of the bench workload (one-accumulator 3x3 conv: 12 us)
    auto conv_us = [&](const ConvTask& t) {
      const ConvCfg& c = cfgs[t.cfg];
      const float mm = static_cast<float>(c.n_kb * c.ntaps);
      const float nacc = static_cast<float>(t.n_samp * t.n_mt);
      float extra = 0.f;
      if (c.flags & F_ATTBWD) extra += 2.8f;
      if (c.flags & F_MASK16) extra += 1.6f;
      if (c.flags & (F_MASK | F_ACCUM)) extra += 2.3f;
      if (c.flags & F_ATTEND) extra += 3.3f;
      return 7.f + nacc * (1.f + (mm <= 100.f ? 0.055f : 0.09f) * mm + extra) + 2.7f * static_cast<float>(t.n_samp - 1);
    };
    auto elt_us = [](const EltTask& e) {
      switch (e.op) {
        case OP_SAME: return 14.f;
        case OP_SAME_BWD: return 21.f;
        case OP_MINMAX: return 2.5f;
        case OP_MINMAX_BWD: return 3.7f;
        case OP_DOTSIG_BWD: return 6.f;
        case OP_ATTEND: return 7.f;
        default: return 8.5f;
      }
    };
    for (auto& b : buckets) {
      // a forked strand starts from its parent's latest stage (strand ids grow from parent to child); a join adds the other
      // strand's latest stage to the producers of this step's stage
      for (int s : forks_at[cur]) { latest[s] = latest[fork_parent[s]]; latest[s].stamp = -1; }
      for (int j : joins_at[cur]) {
        Latest& l = latest[joins[j].primary];
        const Latest& o = latest[joins[j].secondary];
        for (int k = 0; k < o.n; ++k) {
          bool dup = false;
          for (int q = 0; q < l.n; ++q) dup = dup || l.ids[q] == o.ids[k];
          if (dup) continue;
          if (l.n < 8) l.ids[l.n++] = o.ids[k];
          else dep_overflow = true;
        }
      }
      for (int kind = LK_CONV0; kind <= LK_CONV1; ++kind)
        group_convs(b[kind], kind, [&](const ConvTask& t, const int* strands, int ns) { push(&t, TASK_CONV, strands, ns, conv_us(t)); });
      for (int i : b[LK_ELT]) push(&elts[i], TASK_ELT, &elt_sample[i], 1, elt_us(elts[i]));
      // a strand has exactly one stage per step: its `latest` set becomes this step's task ids
      for (int kind = 0; kind < 3; ++kind)
        for (int i : b[kind]) {
          const int smp = kind == LK_ELT ? elt_sample[i] : protos[i].sample;
          if (next[smp].stamp == cur) { latest[smp] = next[smp]; next[smp].stamp = -1; }
        }
      ++cur;
    }
    if (!cp_order || out.empty() || dep_overflow) return;
    static const bool timing = std::getenv("PNMN_PLAN_TIMING") != nullptr;
    const auto t_sort0 = std::chrono::steady_clock::now();
    struct SortTimer {
      bool on; std::chrono::steady_clock::time_point t0; size_t n;
      ~SortTimer() {
        if (on) std::fprintf(stderr, "plan:   critical-path order of %zu tasks: %.2f ms\n", n,
                             std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
      }
    } sort_timer{timing, t_sort0, out.size()};
    const int n = static_cast<int>(out.size());
    // remaining chain time of every task (its own duration + the longest chain of consumers behind it)
    std::vector<float> rem(dur);
    for (int i = n - 1; i >= 0; --i)
      for (int k = 0; k < meta[i].n_deps; ++k) {
        const int j = meta[i].deps[k];
        rem[j] = std::max(rem[j], dur[j] + 1.f + rem[i]);
      }
    // Stable counting sort by remaining time, longest first, quantised to 1/8 us.  The list is topologically ordered and
    // the sort is stable, so a producer (whose remaining time is at least its consumer's) stays in front on ties.
    float rmax = 0.f;
    for (int i = 0; i < n; ++i) rmax = std::max(rmax, rem[i]);
    const int n_bins = std::min(1 << 16, static_cast<int>(rmax * 8.f) + 2);
    std::vector<int> bin(n), start(n_bins + 1, 0), newpos(n);
    for (int i = 0; i < n; ++i) {
      bin[i] = n_bins - 1 - std::min(n_bins - 1, static_cast<int>(rem[i] * 8.f));
      ++start[bin[i] + 1];
    }
    for (int k = 0; k < n_bins; ++k) start[k + 1] += start[k];
    for (int i = 0; i < n; ++i) newpos[i] = start[bin[i]]++;
    for (int i = 0; i < n; ++i)
      for (int k = 0; k < meta[i].n_deps; ++k)
        if (newpos[meta[i].deps[k]] >= newpos[i]) return;  // cannot happen (see above); keep the step-aligned list if it does
    std::vector<TaskRec> out2(n);
    std::vector<TaskMeta> meta2(n);
    for (int i = 0; i < n; ++i) {
      const int d = newpos[i];
      out2[d] = out[i];
      meta2[d] = meta[i];
      for (int k = 0; k < meta2[d].n_deps; ++k) meta2[d].deps[k] = newpos[meta2[d].deps[k]];
    }
    out.swap(out2);
    meta.swap(meta2);
  }
};

}  // namespace

struct pnmn_plan {
  const pnmn_model* m = nullptr;
  int B = 0, L = 0;
  bool need_grad = false;
  std::vector<uint8_t> valid;
  int64_t n16 = 0, n18 = 0, n22 = 0, nmaps = 0, nidx = 0;
  std::vector<ConvCfg> cfgs;
  std::vector<ConvTask> fconv, bconv;
  std::vector<EltTask> felt, belt;
  std::vector<LaunchItem> flaunch, blaunch;
  std::vector<WgradInst> insts;
  std::vector<WgradTask> wtasks;
  std::vector<BiasGradTaskH> btasks;
  std::vector<int> xin_unit;  // per sample, -1 if the stem is skipped (invalid program)
  std::vector<int32_t> map_trace;  // {sample, module call index, token, map unit} of every 1-channel module output
  bool persistent = true;     // one persistent executor launch per pass (exec.cu) vs one launch per level
  int exec_ctas = 0;          // > 0: cap of the executor's persistent grid for this plan (pnmn_plan_set_exec_ctas)
  bool input_by_row = false;  // stem-input unit of sample n is n (pnmn_nmn_prestage), not the running count of valid samples
  // PNMN_PLAN_FORWARD_HALF: a need_grad = 0 plan that stands in for the forward pass of the need_grad = 1 plan of the same
  // programs; its arena sizes include an upper bound of what that plan's backward pass allocates (units of P16 / P18 / P22)
  bool forward_half = false;
  int64_t bwd_bound16 = 0, bwd_bound18 = 0, bwd_bound22 = 0;
  void* uploaded_to = nullptr;  // device buffer that already holds this plan's task tables (pnmn_plan_upload)
  std::vector<TaskRec> ftask, btask;
  std::vector<TaskMeta> fmeta, bmeta;
  int64_t off_ftask = 0, off_fmeta = 0, off_fsync = 0, off_btask = 0, off_bmeta = 0, off_bsync = 0, off_wsync = 0;
  // blob layout (bytes)
  int64_t off_cfg = 0, off_xin = 0, off_pack = 0, off_fconv = 0, off_felt = 0, off_bconv = 0, off_belt = 0, off_inst = 0,
          off_wt = 0, off_bt = 0, blob_bytes = 0;
  int64_t stats[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  std::vector<uint8_t> host_blob;   // level-by-level path only (resolved on the host)
  struct PinnedBlob* pin = nullptr;  // persistent path: unresolved records, resolved on the device after upload
};

namespace {

struct Builder {
  pnmn_plan& p;
  const pnmn_model& m;
  std::vector<Val> vals;
  std::vector<std::vector<WgradInst>> conv_insts;  // per model conv id
  std::vector<PlaneFmt> conv_inst_fmt;
  std::vector<int> conv_inst_dil;

  Builder(pnmn_plan& plan) : p(plan), m(*plan.m), conv_insts(plan.m->convs.size()),
                             conv_inst_fmt(plan.m->convs.size(), kP16), conv_inst_dil(plan.m->convs.size(), 1) {}

  int alloc16(int n = 1) { const int u = static_cast<int>(p.n16); p.n16 += n; return u; }
  int alloc_fmt(PlaneFmt f) {
    if (f.P == kP16.P) return alloc16();
    if (f.P == kP18.P) return static_cast<int>(p.n18++);
    return static_cast<int>(p.n22++);
  }
  float* pfmt(PlaneFmt f, int unit) const {
    if (f.P == kP16.P) return sym<float>(AR_A16, kGuard + unit * kUnit16);
    if (f.P == kP18.P) return sym<float>(AR_A18, kGuard + unit * kUnit18);
    return sym<float>(AR_A22, kGuard + unit * kUnit22);
  }
  float* p16(int unit) const { return pfmt(kP16, unit); }
  // fp16 shadow behind a 128-channel plane buffer (symbolic pointer arithmetic on the byte offset)
  static const void* shadow(const float* p, PlaneFmt f) {
    return reinterpret_cast<const void*>(reinterpret_cast<uint64_t>(p) + shadow_bytes(f.P));
  }
  const float* scalep() const { return sym<const float>(AR_SCRATCH, 0); }
  int64_t ain_unit_bytes() const { return static_cast<int64_t>(m.in_ch / 4) * 256 * 16 * 3 / 2; }
  float* ainp(int k) const { return sym<float>(AR_AIN, kGuard + k * ain_unit_bytes()); }
  float* mapp(int i) const { return sym<float>(AR_MAPS, static_cast<int64_t>(i) * 1024); }
  float* dmapp(int i) const { return sym<float>(AR_DMAPS, static_cast<int64_t>(i) * 1024); }
  const float* param(int64_t off) const { return sym<const float>(AR_PARAMS, off * 4); }
  float* grad(int64_t off) const { return sym<float>(AR_GRADS, off * 4); }
  const void* packed(int64_t off) const { return sym<const void>(AR_PACKED, off * 2); }

  // configurations are looked up ~13k times per plan: compare one packed 64-bit key instead of the 48-byte records
  std::vector<uint64_t> cfg_keys;
  static uint64_t cfg_key(const ConvCfg& c) {
    return static_cast<uint64_t>(c.n_kb) | static_cast<uint64_t>(c.kb_per_in) << 8 | static_cast<uint64_t>(c.ntaps) << 16 |
           static_cast<uint64_t>(c.dil) << 20 | static_cast<uint64_t>(c.S_in) << 24 | static_cast<uint64_t>(c.S_out) << 32 |
           static_cast<uint64_t>(c.S_aux) << 40 | static_cast<uint64_t>(c.flags) << 48;
  }
  int cfg_id(const ConvCfg& c) {
    const uint64_t key = cfg_key(c);
    for (size_t i = 0; i < cfg_keys.size(); ++i)
      if (cfg_keys[i] == key) return static_cast<int>(i);
    p.cfgs.push_back(c);
    cfg_keys.push_back(key);
    return static_cast<int>(p.cfgs.size()) - 1;
  }
  int make_cfg(int n_kb, int kb_per_in, int ntaps, int dil, PlaneFmt in, PlaneFmt out, PlaneFmt aux, int flags) {
    ConvCfg c;
    std::memset(&c, 0, sizeof(c));
    c.n_kb = n_kb; c.kb_per_in = kb_per_in; c.ntaps = ntaps; c.dil = dil;
    c.S_in = in.S; c.P_in = in.P; c.S_out = out.S; c.P_out = out.P; c.S_aux = aux.S; c.P_aux = aux.P;
    c.flags = flags;
    c.lead = ntaps == 9 ? ((dil * in.S + dil + 7) / 8) * 8 : 8;
    return cfg_id(c);
  }

  int cfg_with_flags(int id, int extra) {
    ConvCfg c = p.cfgs[id];
    c.flags |= extra;
    return cfg_id(c);
  }

  int new_val(int kind, int ch, int unit, bool relu_out) {
    Val v; v.kind = kind; v.ch = ch; v.unit = unit; v.relu_out = relu_out;
    vals.push_back(v);
    return static_cast<int>(vals.size()) - 1;
  }
};

// ---- pinned staging buffers for the task tables --------------------------------------------------
// A plan's records are written ONCE (with symbolic pointers) into page-locked memory by pnmn_plan_create,
// uploaded with one truly asynchronous copy and resolved to device addresses by resolve_kernel; the host
// never walks the records again.  Buffers are recycled; the event says when the last upload has been read.
}  // namespace
struct PinnedBlob {
  uint8_t* p = nullptr;
  size_t cap = 0;
  bool pinned = false, pending = false;
  cudaEvent_t ev = nullptr;
};
namespace {
std::vector<PinnedBlob*> g_pin_free;
int g_pin_total = 0;
std::mutex g_pin_mutex;  // plans may be compiled on a helper thread (NeuralModuleNetwork.precompile) and destroyed on another
PinnedBlob* pin_acquire(size_t bytes) {
  std::lock_guard<std::mutex> lock(g_pin_mutex);
  PinnedBlob* best = nullptr;
  // prefer a buffer whose last upload has already been consumed by the GPU (the host may run ahead of the device)
  for (int pass = 0; pass < 2 && !best; ++pass) {
    // everything is in flight: grow the pool up to eight buffers (two look-ahead compiles + the host running two steps ahead of the
    // device; cudaHostAlloc costs milliseconds and synchronises, so the pool must not keep growing), else wait for the oldest
    if (pass == 1 && g_pin_total < 8) break;
    for (size_t i = 0; i < g_pin_free.size(); ++i) {
      PinnedBlob* c = g_pin_free[i];
      if (c->cap < bytes) continue;
      if (pass == 0 && c->pending && cudaEventQuery(c->ev) != cudaSuccess) continue;
      best = c; g_pin_free.erase(g_pin_free.begin() + i); break;
    }
  }
  cudaGetLastError();  // cudaEventQuery's cudaErrorNotReady is not an error
  if (!best) {
    best = new PinnedBlob();
    best->cap = bytes + bytes / 2 + 4096;
    void* q = nullptr;
    if (cudaHostAlloc(&q, best->cap, cudaHostAllocDefault) == cudaSuccess) {
      best->p = static_cast<uint8_t*>(q); best->pinned = true;
      cudaEventCreateWithFlags(&best->ev, cudaEventDisableTiming);
      ++g_pin_total;
    } else {
      cudaGetLastError();  // no device (CPU-only validity checks): plain memory, never uploaded
      best->p = static_cast<uint8_t*>(std::malloc(best->cap));
    }
  }
  if (best->pending) { cudaEventSynchronize(best->ev); best->pending = false; }
  return best;
}
void pin_release(PinnedBlob* b) {
  if (!b) return;
  std::lock_guard<std::mutex> lock(g_pin_mutex);
  g_pin_free.push_back(b);
}

struct BaseTable { uint64_t b[AR_COUNT]; };
// one thread per 8-byte word of a record array; `mask` marks the words that hold (symbolic) pointers.
// type_off >= 0: per-record int at blob[type_off + rec*type_stride] selects mask (0) or mask_alt (!= 0).
__global__ void resolve_kernel(uint8_t* blob, int64_t off, int n_rec, int words, uint32_t mask, uint32_t mask_alt,
                               int64_t type_off, int type_stride, BaseTable bases) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t rec = i / words;
  const int w = static_cast<int>(i % words);
  if (rec >= n_rec) return;
  uint32_t m = mask;
  if (type_off >= 0 && *reinterpret_cast<const int*>(blob + type_off + rec * type_stride) != 0) m = mask_alt;
  if (!((m >> w) & 1u)) return;
  uint64_t* p = reinterpret_cast<uint64_t*>(blob + off) + rec * words + w;
  const uint64_t v = *p;
  if (v >> 56) *p = bases.b[v >> 56] + (v & ((1ull << 56) - 1));
}
constexpr uint32_t kMaskConv = 0x3FFFu;        // ConvTask: 14 pointers in words 0..13
constexpr uint32_t kMaskElt = 0xFFEu;          // EltTask: 11 pointers in words 1..11
constexpr uint32_t kMaskInst = 0x3u;           // WgradInst: dz, x
constexpr uint32_t kMaskWgrad = 0xC1u;         // WgradTask: inst (word 0), dw (6), scale (7)
constexpr uint32_t kMaskBias = 0xDu;           // BiasGradTaskH: inst (0), db (2), scale (3)
static_assert(sizeof(WgradTask) == 64 && sizeof(BiasGradTaskH) == 32 && sizeof(WgradInst) == 16, "record sizes");
static_assert(offsetof(ConvTask, cfg) == 112 && offsetof(EltTask, a) == 8 && offsetof(EltTask, scale) == 88, "pointer words");
static_assert(offsetof(WgradTask, dw) == 48 && offsetof(BiasGradTaskH, db) == 16, "pointer words");

cudaError_t launch_resolve(uint8_t* blob, int64_t off, size_t n_rec, int rec_bytes, uint32_t mask, uint32_t mask_alt,
                           int64_t type_off, int type_stride, const BaseTable& bt, cudaStream_t st) {
  if (n_rec == 0) return cudaSuccess;
  const int words = rec_bytes / 8;
  const int64_t total = static_cast<int64_t>(n_rec) * words;
  resolve_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(blob, off, static_cast<int>(n_rec), words, mask,
                                                                              mask_alt, type_off, type_stride, bt);
  return cudaGetLastError();
}

static const int kRelateDil[5] = {1, 2, 4, 8, 1};

std::mutex g_host_ms_mutex;
double g_host_ms[4] = {0, 0, 0, 0};  // plan_create, forward (host part), backward (host part), calls
struct HostTimer {
  int slot; std::chrono::steady_clock::time_point t0;
  explicit HostTimer(int s) : slot(s), t0(std::chrono::steady_clock::now()) {}
  ~HostTimer() {
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    std::lock_guard<std::mutex> lock(g_host_ms_mutex);
    g_host_ms[slot] += ms;
  }
};
long long* g_trace = nullptr;   // optional device buffer for per-task timestamps (pnmn_debug_set_trace)
int64_t g_trace_cap = 0;        // capacity in tasks

// one persistent executor launch per pass (exec.cu).  Bring-up builds (make BRINGUP=1) can also run the task lists level by
// level on the CUDA-core twin kernels (PNMN_EXEC=levels); the release library has no second implementation.
bool exec_persistent() {
#ifdef PNMN_BRINGUP
  const char* e = std::getenv("PNMN_EXEC");
  return !(e && std::string(e) == "levels");
#else
  return true;
#endif
}

}  // namespace

static pnmn_plan* plan_create_impl(const pnmn_model* m, const int64_t* programs, int B, int L, int need_grad, int flags,
                                   bool allow_strands, bool* overflow) {
  HostTimer timer(0);
  { std::lock_guard<std::mutex> lock(g_host_ms_mutex); g_host_ms[3] += 1; }
  auto* plan = new pnmn_plan();
  pnmn_plan& p = *plan;
  p.m = m; p.B = B; p.L = L; p.need_grad = need_grad != 0;
  p.input_by_row = (flags & PNMN_PLAN_INPUT_BY_ROW) != 0;
  p.forward_half = (flags & PNMN_PLAN_FORWARD_HALF) != 0 && !p.need_grad;
  p.valid.assign(B, 0);
  p.xin_unit.assign(B, -1);
  Builder bd(p);
  p.persistent = exec_persistent();
  Sched fs(B, p.persistent), bs(B, p.persistent);
  int n_ain = 0;  // valid samples so far (index into the stem-input arena)

  // fold feat * map (and its backward) into the neighbouring conv epilogues; only the persistent executor implements it
  static const bool no_fuse = std::getenv("PNMN_NOFUSE") != nullptr;
  const bool fuse_attend = p.persistent && !no_fuse;
  // independent sub-chains of a program on their own strands (Sched); PNMN_NOSTRANDS=1 restores one chain per sample
  static const bool no_strands = std::getenv("PNMN_NOSTRANDS") != nullptr;
  const bool strands_on = p.persistent && allow_strands && !no_strands;
  // (Tried and dropped, profiles/r1_notes: storing conv outputs that only feed other convs as fp16 planes alone, with
  // fp16 ReLU masks in the backward, cuts the executor's DRAM traffic by a third but not its time: the epilogue is
  // latency-bound, not store-bound, and the extra mask path cost registers.)
  const int HF = F_HALF;    // every conv / attend output is the fp16 operand of the next conv (and of wgrad)
  // Persistent executor: activations / gradients that only feed other convs (next conv, wgrad, a ReLU mask) are stored as
  // fp16 planes alone -- the epilogue is instruction-bound (profiles/r1_summary.md), and the fp16-only path needs a
  // fifth of the instructions and a third of the bytes of the fp32 + tf32-rounding + fp16 one.
  static const bool no_lean = std::getenv("PNMN_NOLEAN") != nullptr;
  const bool lean = p.persistent && !no_lean;
  const int ST_INT = lean ? 0 : F_STORE;                       // store flag of such an internal tensor
  const int MASK_INT = lean ? F_MASK16 : F_MASK;               // ReLU mask read from such an internal forward activation
  auto mask_src = [&](const float* act, PlaneFmt f) -> const float* {
    return lean ? static_cast<const float*>(Builder::shadow(act, f)) : act;
  };
  const int EHF = EF_HALF;
  p.nmaps = 1;  // map 0 = the constant all-ones attention of `scene` (nmn.py:216)
  const int ones_val = bd.new_val(VK_ONES, 1, 0, false);
  bool ones_emitted = false;

  std::vector<int> feat_unit(B, -1), y1s_unit(B, -1), dfeat_unit(B, -1);
  int64_t n_conv3 = 0, n_tokens = 0, flops = 0;
  // The forward stages of ALL samples are emitted first, the backward stages in a second pass: every forward tensor then
  // has the same arena unit (and every forward configuration the same id) whether or not the plan has a backward pass, so
  // that a forward pass run from a need_grad = 0 plan (PNMN_PLAN_FORWARD_HALF, compiled in half the time) can be followed
  // by the backward pass of the full plan of the same programs, compiled meanwhile.
  struct SampleRec { int n; std::vector<OpRec> ops; int out, feat_val, xin; bool par; };
  std::vector<SampleRec> emitted;
  if (p.need_grad) emitted.reserve(B);

  for (int n = 0; n < B; ++n) {
    // ---------------- symbolic execution (nmn.py:198-233) ----------------
    const size_t val_mark = bd.vals.size();
    const int feat_val = bd.new_val(VK_FEAT, 128, -1, true);
    std::vector<OpRec> ops;
    int out = feat_val, saved = -1;
    int n_scene = 0, n_binary = 0;
    bool ok = true;
    for (int i = L - 1; i >= 0 && ok; --i) {
      const int64_t tok = programs[static_cast<size_t>(n) * L + i];
      if (tok < 0 || tok >= m->V) { ok = false; break; }  // vocabulary lookup raises -> invalid
      const ModuleDesc& md = m->mods[tok];
      switch (md.kind) {
        case PNMN_TOK_SKIP: break;
        case PNMN_TOK_SCENE: saved = out; out = ones_val; ++n_scene; break;
        case PNMN_TOK_AND:
        case PNMN_TOK_OR: {
          if (saved < 0) { ok = false; break; }
          OpRec r; r.kind = md.kind; r.tok = static_cast<int>(tok); r.in0 = out; r.in1 = saved;
          const int ch = std::max(bd.vals[out].ch, bd.vals[saved].ch);
          r.out = bd.new_val(ch == 1 ? VK_MAP : VK_BUF, ch, -1, false);
          ops.push_back(r); out = r.out; ++n_binary;
        } break;
        case PNMN_TOK_COMPARE: {
          if (saved < 0 || bd.vals[out].ch != 128 || bd.vals[saved].ch != 128) { ok = false; break; }
          OpRec r; r.kind = md.kind; r.tok = static_cast<int>(tok); r.in0 = out; r.in1 = saved; r.nconv = 3;
          r.out = bd.new_val(VK_BUF, 128, -1, true);
          ops.push_back(r); out = r.out; ++n_binary;
        } break;
        case PNMN_TOK_QUERY:
        case PNMN_TOK_RELATE:
        case PNMN_TOK_SAME:
        case PNMN_TOK_ATTENTION: {
          if (bd.vals[out].ch != 1) { ok = false; break; }
          OpRec r; r.kind = md.kind; r.tok = static_cast<int>(tok); r.in0 = out;
          r.nconv = md.kind == PNMN_TOK_RELATE ? 5 : (md.kind == PNMN_TOK_SAME ? 0 : 2);
          r.out = md.kind == PNMN_TOK_QUERY ? bd.new_val(VK_BUF, 128, -1, true) : bd.new_val(VK_MAP, 1, -1, false);
          ops.push_back(r); out = r.out;
        } break;
        default: ok = false; break;
      }
    }
    if (ok && bd.vals[out].ch != 128) ok = false;  // nmn.py:231-232
    const int st_f = fs.new_strand(n);  // the sample's stem strand
    if (!ok) {
      bd.vals.resize(val_mark);
      EltTask g{}; g.op = OP_GATHER; g.a = nullptr;
      g.o = sym<float>(AR_FINAL, static_cast<int64_t>(n) * 128 * 196 * 4);
      fs.add_elt(st_f, g);
      continue;
    }
    // Independent sub-chains on their own strands (see Sched).  Restricted to the well-formed shape -- at most two `scene`
    // tokens and one binary module, so every value has at most one consumer: the backward pass then has at most two
    // concurrent strands, each value's gradient has a single writer, and only d(feat) needs a second accumulator.
    const bool par = strands_on && n_scene <= 2 && n_binary <= 1;
    bd.vals[feat_val].strand = st_f;
    p.valid[n] = 1;
    p.stats[0]++;

    // ---------------- stem (nmn.py:67-72,183) ----------------
    const int xin = p.input_by_row ? n : n_ain++;   // by row: the features were laid out before the programs were known
    p.xin_unit[n] = xin;
    y1s_unit[n] = bd.alloc16();
    feat_unit[n] = bd.alloc16();
    bd.vals[feat_val].unit = feat_unit[n];
    if (!ones_emitted) {
      ones_emitted = true;  // the all-ones map is (re)written by the first valid sample's chain
    }
    {
      const ConvW& c1 = m->convs[m->stem1];
      ConvTask t{};
      t.cfg = bd.make_cfg(m->in_ch / 16, m->in_ch / 16, 9, 1, kP16, kP16, kP16, F_BIAS | F_RELU | ST_INT | HF);
      t.in[0][0] = reinterpret_cast<const void*>(reinterpret_cast<uint64_t>(bd.ainp(xin)) + static_cast<uint64_t>(m->in_ch / 4) * 256 * 16);
      t.out[0] = bd.p16(y1s_unit[n]);
      t.w = bd.packed(c1.pk_fwd); t.bias = bd.param(c1.b_off);
      fs.add_conv(st_f, t, 0);
      const ConvW& c2 = m->convs[m->stem2];
      ConvTask u{};
      u.cfg = bd.make_cfg(8, 8, 9, 1, kP16, kP16, kP16, F_BIAS | F_RELU | F_STORE | HF);
      u.in[0][0] = Builder::shadow(bd.p16(y1s_unit[n]), kP16); u.out[0] = bd.p16(feat_unit[n]);
      u.w = bd.packed(c2.pk_fwd); u.bias = bd.param(c2.b_off);
      fs.add_conv(st_f, u, 0);
    }
    const float* featp = bd.p16(feat_unit[n]);

    // ---------------- forward stages ----------------
    int op_index = -1;
    for (OpRec& r : ops) {
      const ModuleDesc& md = m->mods[r.tok];
      n_tokens++;
      ++op_index;
      // strand of this op's stages
      int sf = st_f;
      if (par) {
        const int pa = bd.vals[r.in0].strand, pb = r.in1 >= 0 ? bd.vals[r.in1].strand : -1;
        const bool binary = r.kind == PNMN_TOK_AND || r.kind == PNMN_TOK_OR || r.kind == PNMN_TOK_COMPARE;
        if (pa >= 0 && pa != st_f) sf = pa;                 // continue the running output's chain
        else if (binary && pb >= 0 && pb != st_f) sf = pb;  // the running output is the stem / all-ones: continue `saved`'s chain
        else sf = fs.fork(st_f);                            // first module after `scene`: a new chain behind the stem
        if (binary && pb >= 0 && pb != st_f && pb != sf) { r.secondary = pb; fs.join(sf, pb); }
      }
      r.strand = sf;
      bd.vals[r.out].strand = sf;
      switch (r.kind) {
        case PNMN_TOK_AND:
        case PNMN_TOK_OR: {
          Val& vo = bd.vals[r.out];
          const Val& a = bd.vals[r.in0]; const Val& b = bd.vals[r.in1];
          EltTask e{}; e.op = OP_MINMAX;
          e.flags = EHF | (r.kind == PNMN_TOK_OR ? EF_MAX : 0) | (a.ch == 1 ? EF_A_MAP : 0) | (b.ch == 1 ? EF_B_MAP : 0);
          e.a = a.ch == 1 ? bd.mapp(a.unit) : bd.p16(a.unit);
          e.b = b.ch == 1 ? bd.mapp(b.unit) : bd.p16(b.unit);
          if (vo.ch == 1) { vo.unit = static_cast<int>(p.nmaps++); e.o = bd.mapp(vo.unit); }
          else { vo.unit = bd.alloc16(); e.o = bd.p16(vo.unit); }
          fs.add_elt(sf, e);
        } break;
        case PNMN_TOK_SAME: {
          Val& vo = bd.vals[r.out];
          vo.unit = static_cast<int>(p.nmaps++);
          r.idx_slot = static_cast<int>(p.nidx++);
          EltTask e{}; e.op = OP_SAME;
          e.a = featp; e.b = bd.mapp(bd.vals[r.in0].unit);
          e.w = bd.param(md.head_w); e.c = bd.param(md.head_b);
          e.o = bd.mapp(vo.unit); e.idx = sym<int>(AR_IDX, static_cast<int64_t>(r.idx_slot) * 4);
          fs.add_elt(sf, e);
        } break;
        case PNMN_TOK_COMPARE: {
          Val& vo = bd.vals[r.out];
          const ConvW& pj = m->convs[md.convs[0]];
          r.y_unit[0] = bd.alloc16(); r.y_unit[1] = bd.alloc16(); r.y_unit[2] = bd.alloc16();
          vo.unit = r.y_unit[2];
          ConvTask t{};
          t.cfg = bd.make_cfg(16, 8, 1, 1, kP16, kP16, kP16, F_BIAS | F_RELU | ST_INT | HF);
          t.in[0][0] = Builder::shadow(bd.p16(bd.vals[r.in0].unit), kP16);
          t.in[1][0] = Builder::shadow(bd.p16(bd.vals[r.in1].unit), kP16);
          t.out[0] = bd.p16(r.y_unit[0]); t.w = bd.packed(pj.pk_fwd); t.bias = bd.param(pj.b_off);
          fs.add_conv(sf, t, 0);
          flops += 2ll * 196 * 128 * 256;
          for (int i = 1; i < 3; ++i) {
            const ConvW& cw = m->convs[md.convs[i]];
            ConvTask u{};
            u.cfg = bd.make_cfg(8, 8, 9, 1, kP16, kP16, kP16, F_BIAS | F_RELU | (i == 2 ? F_STORE : ST_INT) | HF);
            u.in[0][0] = Builder::shadow(bd.p16(r.y_unit[i - 1]), kP16); u.out[0] = bd.p16(r.y_unit[i]);
            u.w = bd.packed(cw.pk_fwd); u.bias = bd.param(cw.b_off);
            fs.add_conv(sf, u, 0);
            n_conv3++;
          }
        } break;
        default: {  // ATTENTION / QUERY / RELATE
          const Val& a = bd.vals[r.in0];
          const float* x = featp;
          if (a.kind == VK_ONES) {
            r.x0_is_feat = true;  // feats * ones == feats: feed the stem output straight in
          } else if (fuse_attend && a.kind == VK_MAP && a.dotsig_proto >= 0 && !a.attend_fused) {
            // the map comes out of a conv epilogue (fused 1x1 head): let that epilogue write feat * map as well
            r.x0_unit = bd.alloc16();
            r.attend_fused = true;
            bd.vals[r.in0].attend_fused = true;
            ConvProto& pr = fs.protos[a.dotsig_proto];
            pr.t.aux[0] = featp;
            pr.t.in[1][0] = Builder::shadow(bd.p16(r.x0_unit), kP16);
            pr.t.cfg = bd.cfg_with_flags(pr.t.cfg, F_ATTEND);
            x = bd.p16(r.x0_unit);
          } else {
            r.x0_unit = bd.alloc16();
            EltTask e{}; e.op = OP_ATTEND; e.flags = EHF; e.a = featp; e.b = bd.mapp(a.unit); e.o = bd.p16(r.x0_unit);
            fs.add_elt(sf, e);
            x = bd.p16(r.x0_unit);
          }
          const bool head = r.kind != PNMN_TOK_QUERY;
          Val& vo = bd.vals[r.out];
          for (int i = 0; i < r.nconv; ++i) {
            const int d = r.kind == PNMN_TOK_RELATE ? kRelateDil[i] : 1;
            const int dn = (r.kind == PNMN_TOK_RELATE && i + 1 < r.nconv) ? kRelateDil[i + 1] : 1;
            const PlaneFmt fin = fmt_for_dilation(d), fout = fmt_for_dilation(dn);
            const ConvW& cw = m->convs[md.convs[i]];
            const bool last = i + 1 == r.nconv;
            r.y_unit[i] = bd.alloc_fmt(fout);
            ConvTask t{};
            t.cfg = bd.make_cfg(8, 8, 9, d, fin, fout, fout, F_BIAS | F_RELU | (last ? F_STORE : ST_INT) | HF | ((last && head) ? F_DOTSIG : 0));
            t.in[0][0] = Builder::shadow(x, fin); t.out[0] = bd.pfmt(fout, r.y_unit[i]);
            t.w = bd.packed(cw.pk_fwd); t.bias = bd.param(cw.b_off);
            if (last && head) {
              vo.unit = static_cast<int>(p.nmaps++);
              t.map_out[0] = bd.mapp(vo.unit);
              t.w3 = bd.param(md.head_w); t.b3 = bd.param(md.head_b);
              flops += 2ll * 196 * 128;
            }
            const int proto = fs.add_conv(sf, t, fin.P == kP22.P ? 1 : 0);
            if (last && head) bd.vals[r.out].dotsig_proto = proto;
            x = t.out[0];
            n_conv3++;
          }
          if (!head) vo.unit = r.y_unit[r.nconv - 1];
        } break;
      }
      if (bd.vals[r.out].ch == 1) {
        const int32_t rec[4] = {n, op_index, r.tok, bd.vals[r.out].unit};
        p.map_trace.insert(p.map_trace.end(), rec, rec + 4);
      }
    }
    {
      EltTask g{}; g.op = OP_GATHER; g.a = bd.p16(bd.vals[out].unit);
      g.o = sym<float>(AR_FINAL, static_cast<int64_t>(n) * 128 * 196 * 4);
      fs.add_elt(bd.vals[out].strand >= 0 ? bd.vals[out].strand : st_f, g);
    }
    if (p.forward_half) {
      // upper bound of the units the backward stages of this sample will allocate (checked against the full plan in
      // tests/test_plan_cpu.py): d(feat) twice, the stem's dZ, the output's gradient; per module the dZ / dX buffers of its
      // convolutions and one gradient buffer per 128-channel value
      p.bwd_bound16 += 4;
      for (const OpRec& r : ops) {
        switch (r.kind) {
          case PNMN_TOK_AND: case PNMN_TOK_OR: p.bwd_bound16 += 3; break;
          case PNMN_TOK_COMPARE: p.bwd_bound16 += 5; break;
          case PNMN_TOK_RELATE: p.bwd_bound16 += 4; p.bwd_bound18 += 1; p.bwd_bound22 += 1; break;
          case PNMN_TOK_SAME: break;
          default: p.bwd_bound16 += 4; break;   // ATTENTION / QUERY
        }
      }
    }
    if (!p.need_grad) continue;
    emitted.push_back(SampleRec{n, std::move(ops), out, feat_val, xin, par});
  }

  for (SampleRec& sr : emitted) {
    const int n = sr.n, out = sr.out, feat_val = sr.feat_val, xin = sr.xin;
    const bool par = sr.par;
    std::vector<OpRec>& ops = sr.ops;
    const float* featp = bd.p16(feat_unit[n]);
    // ---------------- liveness / consumer counts ----------------
    bd.vals[out].live = true; bd.vals[out].ncons = 1;
    for (int k = static_cast<int>(ops.size()) - 1; k >= 0; --k) {
      const OpRec& r = ops[k];
      if (!bd.vals[r.out].live) continue;
      for (int in : {r.in0, r.in1}) {
        if (in < 0 || bd.vals[in].kind == VK_ONES) continue;
        bd.vals[in].live = true;
        bd.vals[in].ncons++;
        if (r.kind == PNMN_TOK_AND || r.kind == PNMN_TOK_OR) bd.vals[in].cons_minmax = true;
      }
    }
    dfeat_unit[n] = bd.alloc16();
    bd.vals[feat_val].gunit = dfeat_unit[n];
    const float* dfeatp = bd.p16(dfeat_unit[n]);
    auto fused_mask = [&](const Val& v) { return v.relu_out && v.kind == VK_BUF && v.ncons == 1 && !v.cons_minmax; };
    auto gbuf = [&](Val& v) -> float* {
      if (v.gunit < 0) v.gunit = bd.alloc16();
      return bd.p16(v.gunit);
    };
    auto add_inst = [&](int conv_id, const float* dz, const float* x, PlaneFmt f, int dil) {
      bd.conv_insts[conv_id].push_back(WgradInst{Builder::shadow(dz, f), Builder::shadow(x, f)});
      bd.conv_inst_fmt[conv_id] = f;
      bd.conv_inst_dil[conv_id] = dil;
    };

    // ---------------- backward stages ----------------
    // Backward strands mirror the forward ones: the root strand carries the output's chain; at the backward of a binary
    // module the chain of its other input forks off.  The forked chain accumulates its share of d(feat) in a buffer of
    // its own (two concurrent read-modify-write streams into one buffer would race); the stem's backward adds the two.
    const int rb = bs.new_strand(n);
    int sec_f = -1, sec_b = -1;          // forward / backward strand of the forked chain
    int dfeat2_unit = -1;
    bool dfeat2_written = false;
    auto bstrand = [&](const OpRec& r) { return (par && r.strand == sec_f && sec_b >= 0) ? sec_b : rb; };
    // (pointer, written flag) of the d(feat) accumulator a backward strand uses
    auto dfeat_of = [&](int sb, bool*& written) -> float* {
      if (sb == sec_b && sec_b >= 0) {
        if (dfeat2_unit < 0) dfeat2_unit = bd.alloc16();
        written = &dfeat2_written;
        return bd.p16(dfeat2_unit);
      }
      written = &bd.vals[feat_val].gwritten;
      return const_cast<float*>(dfeatp);
    };
    {  // d(final module output) arrives as NCHW from the classifier
      Val& v = bd.vals[out];
      EltTask e{}; e.op = OP_SCATTER; e.scale = bd.scalep();
      e.a = sym<const float>(AR_GRADOUT, static_cast<int64_t>(n) * 128 * 196 * 4);
      e.b = fused_mask(v) ? bd.p16(v.unit) : nullptr;
      e.o = gbuf(v); e.flags = v.gwritten ? EF_ACCUM : 0;
      v.gwritten = true;
      bs.add_elt(rb, e);
    }
    for (int k = static_cast<int>(ops.size()) - 1; k >= 0; --k) {
      OpRec& r = ops[k];
      Val& vo = bd.vals[r.out];
      if (!vo.live) continue;
      const ModuleDesc& md = m->mods[r.tok];
      const int sb = bstrand(r);
      // 128-channel outputs whose gradient was accumulated unmasked need the ReLU mask now
      if (vo.kind == VK_BUF && vo.relu_out && !fused_mask(vo)) {
        EltTask e{}; e.op = OP_RELU_MASK; e.a = bd.p16(vo.gunit); e.b = bd.p16(vo.unit); e.o = bd.p16(vo.gunit);
        bs.add_elt(sb, e);
      }
      switch (r.kind) {
        case PNMN_TOK_AND:
        case PNMN_TOK_OR: {
          Val& a = bd.vals[r.in0]; Val& b = bd.vals[r.in1];
          EltTask e{}; e.op = OP_MINMAX_BWD;
          e.flags = (r.kind == PNMN_TOK_OR ? EF_MAX : 0) | (a.ch == 1 ? EF_A_MAP : 0) | (b.ch == 1 ? EF_B_MAP : 0);
          e.a = a.ch == 1 ? bd.mapp(a.unit) : bd.p16(a.unit);
          e.b = b.ch == 1 ? bd.mapp(b.unit) : bd.p16(b.unit);
          e.g = vo.ch == 1 ? bd.dmapp(vo.unit) : bd.p16(vo.gunit);
          if (a.kind != VK_ONES) {
            if (a.ch == 1) e.o = bd.dmapp(a.unit);
            else { e.o = gbuf(a); if (a.gwritten) e.flags |= EF_ACCUM; a.gwritten = true; }
          }
          if (b.kind != VK_ONES) {
            if (b.ch == 1) e.o2 = bd.dmapp(b.unit);
            else { e.o2 = gbuf(b); if (b.gwritten) e.flags |= EF_ACCUM2; b.gwritten = true; }
          }
          bs.add_elt(sb, e);
          if (par && r.secondary >= 0 && sec_b < 0) { sec_f = r.secondary; sec_b = bs.fork(sb); }
        } break;
        case PNMN_TOK_SAME: {
          const Val& a = bd.vals[r.in0];
          Val& fv = bd.vals[feat_val];
          EltTask e{}; e.op = OP_SAME_BWD; e.scale = bd.scalep();
          e.a = featp; e.b = bd.mapp(a.unit); e.c = bd.mapp(vo.unit); e.g = bd.dmapp(vo.unit);
          e.o = a.kind == VK_ONES ? nullptr : bd.dmapp(a.unit);
          bool* dfw = nullptr;
          e.o2 = dfeat_of(sb, dfw); e.flags = *dfw ? EF_ACCUM : 0; *dfw = true;
          (void)fv;
          e.w = bd.param(md.head_w); e.dw = bd.grad(md.head_w); e.dw2 = bd.grad(md.head_b);
          e.idx = sym<int>(AR_IDX, static_cast<int64_t>(r.idx_slot) * 4);
          bs.add_elt(sb, e);
        } break;
        case PNMN_TOK_COMPARE: {
          // dZ of conv2 is the (masked) gradient of the output value
          const float* dz = bd.p16(vo.gunit);
          for (int i = 2; i >= 1; --i) {
            const ConvW& cw = m->convs[md.convs[i]];
            add_inst(md.convs[i], dz, bd.p16(r.y_unit[i - 1]), kP16, 1);
            const int du = bd.alloc16();
            ConvTask t{};
            t.cfg = bd.make_cfg(8, 8, 9, 1, kP16, kP16, kP16, ST_INT | MASK_INT | F_HALF);
            t.in[0][0] = Builder::shadow(dz, kP16); t.out[0] = bd.p16(du);
            t.aux[0] = mask_src(bd.p16(r.y_unit[i - 1]), kP16);
            t.w = bd.packed(cw.pk_bwd);
            bs.add_conv(sb, t, 0);
            dz = bd.p16(du);
          }
          const ConvW& pj = m->convs[md.convs[0]];
          const int ins[2] = {r.in0, r.in1};
          for (int h = 0; h < 2; ++h) {
            Val& v = bd.vals[ins[h]];
            // projection wgrad: dW[:, 128h:128h+128] += dZp (x) in_h
            bd.conv_insts[md.convs[0]].push_back(
                WgradInst{Builder::shadow(dz, kP16), Builder::shadow(bd.p16(v.unit), kP16)});  // even = in0, odd = in1
            const bool fm = fused_mask(v);
            ConvTask t{};
            t.cfg = bd.make_cfg(8, 8, 1, 1, kP16, kP16, kP16, F_STORE | (fm ? (F_MASK | F_HALF) : 0) | (v.gwritten ? F_ACCUM : 0));
            t.in[0][0] = Builder::shadow(dz, kP16); t.out[0] = gbuf(v); t.aux[0] = fm ? bd.p16(v.unit) : nullptr;
            t.w = bd.packed(pj.pk_bwd + static_cast<int64_t>(h) * 8 * 2048);
            v.gwritten = true;
            bs.add_conv(sb, t, 0);   // (the two projection halves are consecutive stages of this strand)
          }
          if (par && r.secondary >= 0 && sec_b < 0) { sec_f = r.secondary; sec_b = bs.fork(sb); }
        } break;
        default: {  // ATTENTION / QUERY / RELATE
          const Val& a = bd.vals[r.in0];
          const bool head = r.kind != PNMN_TOK_QUERY;
          const int nc = r.nconv;
          auto dil_of = [&](int i) { return r.kind == PNMN_TOK_RELATE ? kRelateDil[i] : 1; };
          const float* dz;
          if (head) {
            const int du = bd.alloc16();
            EltTask e{}; e.op = OP_DOTSIG_BWD; e.scale = bd.scalep();
            e.g = bd.dmapp(vo.unit); e.c = bd.mapp(vo.unit); e.a = bd.p16(r.y_unit[nc - 1]);
            e.w = bd.param(md.head_w); e.dw = bd.grad(md.head_w); e.dw2 = bd.grad(md.head_b);
            e.o = bd.p16(du);
            bs.add_elt(sb, e);
            dz = bd.p16(du);
          } else {
            dz = bd.p16(vo.gunit);
          }
          for (int i = nc - 1; i >= 0; --i) {
            const int d = dil_of(i);
            const PlaneFmt f = fmt_for_dilation(d);  // format of dZ_i and of conv i's input
            const ConvW& cw = m->convs[md.convs[i]];
            const float* xin_i = i == 0 ? (r.x0_is_feat ? featp : bd.p16(r.x0_unit)) : bd.pfmt(f, r.y_unit[i - 1]);
            add_inst(md.convs[i], dz, xin_i, f, d);
            ConvTask t{};
            t.in[0][0] = Builder::shadow(dz, f); t.w = bd.packed(cw.pk_bwd);
            if (i > 0) {
              const PlaneFmt fprev = fmt_for_dilation(dil_of(i - 1));
              const int du = bd.alloc_fmt(fprev);
              t.cfg = bd.make_cfg(8, 8, 9, d, f, fprev, f, ST_INT | MASK_INT | F_HALF);
              t.out[0] = bd.pfmt(fprev, du);
              t.aux[0] = mask_src(xin_i, f);
              bs.add_conv(sb, t, f.P == kP22.P ? 1 : 0);
              dz = t.out[0];
            } else if (r.x0_is_feat) {
              bool* dfw = nullptr;
              float* dfp = dfeat_of(sb, dfw);
              t.cfg = bd.make_cfg(8, 8, 9, d, f, kP16, kP16, F_STORE | (*dfw ? F_ACCUM : 0));
              t.out[0] = dfp;
              *dfw = true;
              bs.add_conv(sb, t, 0);
            } else if (fuse_attend) {
              // dX0 never touches memory: the epilogue turns it into dmap += <dX0, feat> and dfeat (+)= dX0 * map
              bool* dfw = nullptr;
              float* dfp = dfeat_of(sb, dfw);
              t.cfg = bd.make_cfg(8, 8, 9, d, f, kP16, kP16, F_ATTBWD | (*dfw ? F_ACCUM : 0));
              t.out[0] = dfp;
              t.aux[0] = featp;
              t.map_out[0] = bd.mapp(a.unit);
              t.in[1][0] = bd.dmapp(a.unit);
              *dfw = true;
              bs.add_conv(sb, t, 0);
            } else {
              const int du = bd.alloc16();
              t.cfg = bd.make_cfg(8, 8, 9, d, f, kP16, kP16, F_STORE);
              t.out[0] = bd.p16(du);
              bs.add_conv(sb, t, 0);
              bool* dfw = nullptr;
              float* dfp = dfeat_of(sb, dfw);
              EltTask e{}; e.op = OP_ATTEND_BWD;
              e.a = bd.p16(du); e.b = featp; e.c = bd.mapp(a.unit);
              e.o = bd.dmapp(a.unit); e.o2 = dfp;
              e.flags = *dfw ? EF_ACCUM : 0; *dfw = true;
              bs.add_elt(sb, e);
            }
          }
        } break;
      }
    }
    // ---------------- stem backward ----------------
    if (bd.vals[feat_val].gwritten || dfeat2_written) {
      // d(feat) = the root strand's accumulator (+ the forked chain's), masked by the stem's ReLU
      const bool main_w = bd.vals[feat_val].gwritten;
      EltTask e{}; e.op = OP_RELU_MASK; e.b = featp; e.o = const_cast<float*>(dfeatp);
      e.a = main_w ? dfeatp : bd.p16(dfeat2_unit);
      e.c = (main_w && dfeat2_written) ? bd.p16(dfeat2_unit) : nullptr;
      if (sec_b >= 0) bs.join(rb, sec_b);
      bs.add_elt(rb, e);
      const int dz1 = bd.alloc16();
      const ConvW& c2 = m->convs[m->stem2];
      ConvTask t{};
      t.cfg = bd.make_cfg(8, 8, 9, 1, kP16, kP16, kP16, ST_INT | MASK_INT | F_HALF);
      t.in[0][0] = Builder::shadow(dfeatp, kP16); t.out[0] = bd.p16(dz1);
      t.aux[0] = mask_src(bd.p16(y1s_unit[n]), kP16);
      t.w = bd.packed(c2.pk_bwd);
      bs.add_conv(rb, t, 0);
      add_inst(m->stem2, dfeatp, bd.p16(y1s_unit[n]), kP16, 1);
      bd.conv_insts[m->stem1].push_back(WgradInst{
          Builder::shadow(bd.p16(dz1), kP16),
          reinterpret_cast<const void*>(reinterpret_cast<uint64_t>(bd.ainp(xin)) + static_cast<uint64_t>(m->in_ch / 4) * 256 * 16)});
    }
  }

  // ---------------- flatten ----------------
  const auto t_emit = std::chrono::steady_clock::now();
  // the forward list, the backward list and the weight-gradient tasks are independent: the backward list (the longest)
  // is flattened on a helper thread while this thread does the other two
  std::thread bwd_flatten;
  {
    // thresholds from replaying the recorded trace over the bench workload (scripts/sim_sched.py) and confirmed on the GPU;
    // the backward pass has ~1.8x the tasks of the forward pass, so splitting costs it more slot time
    static const float crit_fwd = std::getenv("PNMN_CRIT") ? static_cast<float>(std::atof(std::getenv("PNMN_CRIT"))) : 0.7f;
    static const float crit_bwd = std::getenv("PNMN_CRIT_BWD") ? static_cast<float>(std::atof(std::getenv("PNMN_CRIT_BWD"))) : 0.8f;
    fs.mark_critical(B, crit_fwd);
    if (p.need_grad) bs.mark_critical(B, crit_bwd);
  }
  if (p.persistent) {
    static const bool plan_timing_f = std::getenv("PNMN_PLAN_TIMING") != nullptr;
    if (p.need_grad) bwd_flatten = std::thread([&] {
      const auto a = std::chrono::steady_clock::now();
      bs.flatten_persistent(p.btask, p.bmeta, p.cfgs);
      if (plan_timing_f) std::fprintf(stderr, "plan:   backward flatten (helper thread) %.2f ms\n",
                                      std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - a).count());
    });
    {
      const auto a = std::chrono::steady_clock::now();
      fs.flatten_persistent(p.ftask, p.fmeta, p.cfgs);
      if (plan_timing_f) std::fprintf(stderr, "plan:   forward flatten %.2f ms\n",
                                      std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - a).count());
    }
  } else {
    fs.flatten(p.felt, p.fconv, p.flaunch);
  }
  struct Joiner { std::thread& t; ~Joiner() { if (t.joinable()) t.join(); } } joiner{bwd_flatten};
  if (p.need_grad) {
    if (!p.persistent) bs.flatten(p.belt, p.bconv, p.blaunch);
    // wgrad / bias-grad tasks per weight tensor
    const int64_t inst_base = 0;
    (void)inst_base;
    for (size_t c = 0; c < m->convs.size(); ++c) {
      const auto& v = bd.conv_insts[c];
      if (v.empty()) continue;
      const ConvW& cw = m->convs[c];
      const PlaneFmt f = bd.conv_inst_fmt[c];
      if (cw.ksize == 1) {
        // projection: instances alternate (dz,in0),(dz,in1); split into the two 128-channel halves
        for (int h = 0; h < 2; ++h) {
          const int64_t first = static_cast<int64_t>(p.insts.size());
          for (size_t i = h; i < v.size(); i += 2) p.insts.push_back(v[i]);
          const int n_inst = static_cast<int>(p.insts.size() - first);
          for (int i0 = 0; i0 < n_inst; i0 += kInstChunk) {
            WgradTask t{};
            t.inst = sym<const WgradInst>(AR_BLOB, (first + i0) * static_cast<int64_t>(sizeof(WgradInst)));
            t.n_inst = std::min(kInstChunk, n_inst - i0);
            t.tap_row = 0; t.ntaps_x = 1; t.dil = 1; t.S = f.S; t.P = f.P;
            t.cin_total = 256; t.cin0 = 128 * h; t.ksize = 1; t.dw = bd.grad(cw.w_off); t.scale = bd.scalep();
            p.wtasks.push_back(t);
          }
          if (h == 0) {
            BiasGradTaskH b{};
            b.inst = sym<const WgradInst>(AR_BLOB, first * static_cast<int64_t>(sizeof(WgradInst)));
            b.n_inst = n_inst; b.P = f.P; b.db = bd.grad(cw.b_off); b.scale = bd.scalep();
            p.btasks.push_back(b);
          }
        }
        continue;
      }
      const int tiles = cw.cin / 128;
      for (int tile = 0; tile < tiles; ++tile) {
        const int64_t first = static_cast<int64_t>(p.insts.size());
        for (const WgradInst& wi : v) {
          WgradInst x = wi;
          x.x = reinterpret_cast<const void*>(reinterpret_cast<uint64_t>(wi.x) + static_cast<uint64_t>(tile) * 16 * 256 * 16);
          p.insts.push_back(x);
        }
        const int n_inst = static_cast<int>(v.size());
        for (int i0 = 0; i0 < n_inst; i0 += kInstChunk)
          for (int ty = 0; ty < 3; ++ty) {
            WgradTask t{};
            t.inst = sym<const WgradInst>(AR_BLOB, (first + i0) * static_cast<int64_t>(sizeof(WgradInst)));
            t.n_inst = std::min(kInstChunk, n_inst - i0);
            t.tap_row = ty; t.ntaps_x = 3; t.dil = bd.conv_inst_dil[c]; t.S = f.S; t.P = f.P;
            t.cin_total = cw.cin; t.cin0 = 128 * tile; t.ksize = 3; t.dw = bd.grad(cw.w_off); t.scale = bd.scalep();
            p.wtasks.push_back(t);
          }
        if (tile == 0) {
          BiasGradTaskH b{};
          b.inst = sym<const WgradInst>(AR_BLOB, first * static_cast<int64_t>(sizeof(WgradInst)));
          b.n_inst = n_inst; b.P = f.P; b.db = bd.grad(cw.b_off); b.scale = bd.scalep();
          p.btasks.push_back(b);
        }
      }
    }
    p.blaunch.push_back({LK_WGRAD, 0, static_cast<int>(p.wtasks.size())});
    p.blaunch.push_back({LK_BIAS, 0, static_cast<int>(p.btasks.size())});
  }

  if (bwd_flatten.joinable()) bwd_flatten.join();
  const auto t_flat = std::chrono::steady_clock::now();
  static const bool plan_timing = std::getenv("PNMN_PLAN_TIMING") != nullptr;
  auto ms_between = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
  if (fs.dep_overflow || bs.dep_overflow) {
    g_err = "internal error: a task has more predecessors than kMaxDeps";
    *overflow = true;
    delete plan;
    return nullptr;
  }
  // ---------------- blob layout ----------------
  auto align = [](int64_t x) { return (x + 255) / 256 * 256; };
  int64_t o = 0;
  p.off_cfg = o; o = align(o + static_cast<int64_t>(p.cfgs.size() * sizeof(ConvCfg)));
  p.off_xin = o; o = align(o + static_cast<int64_t>(B) * 8);
  // the model's (static) weight-pack tasks travel with every plan: the library then owns no device memory at all
  p.off_pack = o; o = align(o + static_cast<int64_t>(m->pack.size() * sizeof(PackTask)));
  p.off_fconv = o; o = align(o + static_cast<int64_t>(p.fconv.size() * sizeof(ConvTask)));
  p.off_felt = o; o = align(o + static_cast<int64_t>(p.felt.size() * sizeof(EltTask)));
  p.off_ftask = o; o = align(o + static_cast<int64_t>(p.ftask.size() * sizeof(TaskRec)));
  p.off_fmeta = o; o = align(o + static_cast<int64_t>(p.fmeta.size() * sizeof(TaskMeta)));
  p.off_bconv = o; o = align(o + static_cast<int64_t>(p.bconv.size() * sizeof(ConvTask)));
  p.off_belt = o; o = align(o + static_cast<int64_t>(p.belt.size() * sizeof(EltTask)));
  p.off_inst = o; o = align(o + static_cast<int64_t>(p.insts.size() * sizeof(WgradInst)));
  p.off_wt = o; o = align(o + static_cast<int64_t>(p.wtasks.size() * sizeof(WgradTask)));
  p.off_bt = o; o = align(o + static_cast<int64_t>(p.btasks.size() * sizeof(BiasGradTaskH)));
  p.off_btask = o; o = align(o + static_cast<int64_t>(p.btask.size() * sizeof(TaskRec)));
  p.off_bmeta = o; o = align(o + static_cast<int64_t>(p.bmeta.size() * sizeof(TaskMeta)));
  p.off_fsync = o; o = align(o + 4 * static_cast<int64_t>(p.ftask.size() + 1));
  p.off_bsync = o; o = align(o + 4 * static_cast<int64_t>(p.btask.size() + 1));
  p.off_wsync = o; o = align(o + 4);   // task counter of the weight-gradient kernel
  p.blob_bytes = std::max<int64_t>(o, 256);
  // instance pointers inside wgrad/bias tasks were relative to the instance table
  for (auto& t : p.wtasks)
    t.inst = reinterpret_cast<const WgradInst*>(reinterpret_cast<uint64_t>(t.inst) + static_cast<uint64_t>(p.off_inst));
  for (auto& t : p.btasks)
    t.inst = reinterpret_cast<const WgradInst*>(reinterpret_cast<uint64_t>(t.inst) + static_cast<uint64_t>(p.off_inst));

  if (p.persistent) {
    p.pin = pin_acquire(static_cast<size_t>(p.blob_bytes));
    uint8_t* hb = p.pin->p;
    auto put = [&](int64_t off, const void* src, size_t bytes) { if (bytes) std::memcpy(hb + off, src, bytes); };
    put(p.off_cfg, p.cfgs.data(), p.cfgs.size() * sizeof(ConvCfg));
    put(p.off_pack, m->pack.data(), m->pack.size() * sizeof(PackTask));
    put(p.off_ftask, p.ftask.data(), p.ftask.size() * sizeof(TaskRec));
    put(p.off_fmeta, p.fmeta.data(), p.fmeta.size() * sizeof(TaskMeta));
    put(p.off_btask, p.btask.data(), p.btask.size() * sizeof(TaskRec));
    put(p.off_bmeta, p.bmeta.data(), p.bmeta.size() * sizeof(TaskMeta));
    put(p.off_inst, p.insts.data(), p.insts.size() * sizeof(WgradInst));
    put(p.off_wt, p.wtasks.data(), p.wtasks.size() * sizeof(WgradTask));
    put(p.off_bt, p.btasks.data(), p.btasks.size() * sizeof(BiasGradTaskH));
    int64_t* x = reinterpret_cast<int64_t*>(hb + p.off_xin);
    const int64_t ain_unit = static_cast<int64_t>(m->in_ch / 4) * 256 * 16 * 3 / 2;
    for (int n = 0; n < B; ++n) x[n] = p.xin_unit[n] < 0 ? -1 : (kGuard + p.xin_unit[n] * ain_unit) / 4;
    // the dependency flags + task counters live at the tail of the blob: they are uploaded as zeros
    std::memset(hb + p.off_fsync, 0, static_cast<size_t>(p.blob_bytes - p.off_fsync));
  }
  const auto t_blob = std::chrono::steady_clock::now();
  p.stats[1] = n_conv3;
  p.stats[2] = n_tokens;
  p.stats[3] = static_cast<int64_t>(p.flaunch.size());
  p.stats[4] = static_cast<int64_t>(p.blaunch.size());
  p.stats[5] = flops + n_conv3 * 2ll * 196 * 128 * 128 * 9;
  p.stats[6] = static_cast<int64_t>(p.fconv.size());
  p.stats[7] = static_cast<int64_t>(p.wtasks.size());
  // algorithmic FLOPs (dense convention 2*196*N*K, padding taps counted) per kernel family
  auto conv_flops = [&](const std::vector<ConvTask>& tasks, const std::vector<LaunchItem>& ls, int kind) {
    int64_t f = 0;
    for (const LaunchItem& l : ls)
      if (l.kind == kind)
        for (int i = 0; i < l.count; ++i) {
          const ConvTask& t = tasks[l.off + i];
          const ConvCfg& c = p.cfgs[t.cfg];
          f += static_cast<int64_t>(t.n_samp) * 2 * 196 * 128 * (16ll * c.n_kb * c.ntaps);
        }
    return f;
  };
  p.stats[8] = conv_flops(p.fconv, p.flaunch, LK_CONV0);
  p.stats[9] = conv_flops(p.fconv, p.flaunch, LK_CONV1);
  p.stats[10] = conv_flops(p.bconv, p.blaunch, LK_CONV0);
  p.stats[11] = conv_flops(p.bconv, p.blaunch, LK_CONV1);
  if (p.persistent) {
    auto task_flops = [&](const std::vector<TaskRec>& ts, const std::vector<TaskMeta>& ms) {
      int64_t f = 0;
      for (size_t i = 0; i < ts.size(); ++i)
        if (ms[i].type == TASK_CONV) {
          const ConvTask& t = *reinterpret_cast<const ConvTask*>(ts[i].b);
          const ConvCfg& c = p.cfgs[t.cfg];
          const int nmt = c.P_in == 484 ? 3 : 2;
          f += static_cast<int64_t>(t.n_samp) * 2 * 196 * 128 * (16ll * c.n_kb * c.ntaps) * t.n_mt / nmt;
        }
      return f;
    };
    p.stats[8] = task_flops(p.ftask, p.fmeta);
    p.stats[10] = task_flops(p.btask, p.bmeta);
    p.stats[3] = 1; p.stats[4] = 3;
    p.stats[6] = static_cast<int64_t>(p.ftask.size());
    p.stats[14] = static_cast<int64_t>(p.btask.size());
  }
  for (const WgradTask& t : p.wtasks) p.stats[12] += static_cast<int64_t>(t.n_inst) * 2 * 196 * 128 * 128 * t.ntaps_x;
  p.stats[13] = static_cast<int64_t>(p.felt.size() + p.belt.size());
  p.stats[15] = p.off_xin;  // byte offset, inside the task-table blob, of the per-sample stem-input table (int64, < 0 = invalid program)
  if (plan_timing)
    std::fprintf(stderr, "plan: emit %.2f ms, flatten+wgrad %.2f ms, blob %.2f ms, stats %.2f ms, tasks fwd %zu bwd %zu\n",
                 ms_between(timer.t0, t_emit), ms_between(t_emit, t_flat), ms_between(t_flat, t_blob),
                 ms_between(t_blob, std::chrono::steady_clock::now()), p.ftask.size(), p.btask.size());
  return plan;
}

extern "C" pnmn_plan* pnmn_plan_create_ex(const pnmn_model* m, const int64_t* programs, int B, int L, int need_grad, int flags) {
  bool overflow = false;
  pnmn_plan* p = plan_create_impl(m, programs, B, L, need_grad, flags, true, &overflow);
  // joins of concurrent strands can exceed a task's dependency slots on exotic batches: compile those with one chain per sample
  if (!p && overflow) p = plan_create_impl(m, programs, B, L, need_grad, flags, false, &overflow);
  return p;
}
extern "C" pnmn_plan* pnmn_plan_create(const pnmn_model* m, const int64_t* programs, int B, int L, int need_grad) {
  return pnmn_plan_create_ex(m, programs, B, L, need_grad, 0);
}

// ---- program-independent part of the forward pass ---------------------------------------------------------------------
extern "C" int64_t pnmn_model_pack_table_bytes(const pnmn_model* m) { return static_cast<int64_t>(m->pack.size() * sizeof(PackTask)); }
extern "C" int pnmn_model_pack_table(const pnmn_model* m, void* host_out) {
  std::memcpy(host_out, m->pack.data(), m->pack.size() * sizeof(PackTask));
  return 0;
}
extern "C" int64_t pnmn_model_ain_floats(const pnmn_model* m, int batch) {
  return (2 * kGuard + std::max<int64_t>(batch, 1) * (static_cast<int64_t>(m->in_ch / 4) * 256 * 16 * 3 / 2)) / 4;
}
extern "C" int pnmn_nmn_prestage(const pnmn_model* m, const void* pack_table_dev, const float* params, void* packed,
                                 const void* features, int features_half, float* ain, int batch, void* stream) {
  if (!m || !pack_table_dev || !params || !packed || !features || !ain || batch < 1) return fail("pnmn_nmn_prestage: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CUDA_OK(launch_pack(static_cast<const PackTask*>(pack_table_dev), static_cast<int>(m->pack.size()), m->total_tiles, params, packed, st));
  const int64_t unit = static_cast<int64_t>(m->in_ch / 4) * 256 * 16 * 3 / 2;
  CUDA_OK(launch_nchw_to_planes(features, features_half, ain, batch, m->in_ch, nullptr, kGuard / 4, unit / 4, st));
  pnmn::count_launches(2);
  return 0;
}

// Optional early upload of a plan's task tables (input pipelines): copies the page-locked tables into `device_blob`
// (PNMN_SZ_BLOB bytes, caller-owned) on `stream`.  A later pnmn_nmn_forward whose buffers name the same `blob` skips its own
// upload -- the caller makes the forward's stream wait for this one.  Why: an upload issued inside forward is ordered behind
// the previous step's kernels, and by the time it may start the host -> device copy engine is usually busy with the NEXT
// batch's features (3.7 ms for 205 MB), which stalls the executor; uploaded ahead of time the tables are simply there.
extern "C" int pnmn_plan_upload(pnmn_plan* pp, void* device_blob, void* stream) {
  if (!pp || !device_blob) return fail("pnmn_plan_upload: bad arguments");
  pnmn_plan& p = *pp;
  if (!p.persistent) return fail("pnmn_plan_upload: only the persistent executor keeps its tables in one blob");
  if (!p.pin || !p.pin->pinned) return fail("plan has no page-locked task table (created without a CUDA device?)");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CUDA_OK(cudaMemcpyAsync(device_blob, p.pin->p, static_cast<size_t>(p.blob_bytes), cudaMemcpyHostToDevice, st));
  CUDA_OK(cudaEventRecord(p.pin->ev, st));
  p.pin->pending = true;
  p.uploaded_to = device_blob;
  return 0;
}

extern "C" int pnmn_plan_set_exec_ctas(pnmn_plan* p, int max_ctas) {
  if (!p) return fail("pnmn_plan_set_exec_ctas: NULL plan");
  p->exec_ctas = max_ctas > 0 ? max_ctas : 0;
  return 0;
}

extern "C" void pnmn_plan_destroy(pnmn_plan* p) {
  if (p) pin_release(p->pin);
  delete p;
}

extern "C" int pnmn_plan_valid(const pnmn_plan* p, uint8_t* valid) {
  std::memcpy(valid, p->valid.data(), p->valid.size());
  return 0;
}

extern "C" int pnmn_plan_sizes(const pnmn_plan* p, int64_t* s) {
  for (int i = 0; i < PNMN_SZ_COUNT; ++i) s[i] = 0;
  s[PNMN_SZ_ARENA16] = (2 * kGuard + std::max<int64_t>(p->n16 + p->bwd_bound16, 1) * kUnit16) / 4;
  s[PNMN_SZ_ARENA18] = (2 * kGuard + std::max<int64_t>(p->n18 + p->bwd_bound18, 1) * kUnit18) / 4;
  s[PNMN_SZ_ARENA22] = (2 * kGuard + std::max<int64_t>(p->n22 + p->bwd_bound22, 1) * kUnit22) / 4;
  s[PNMN_SZ_MAPS] = p->nmaps * 256;
  s[PNMN_SZ_DMAPS] = p->nmaps * 256;
  s[PNMN_SZ_IDX] = std::max<int64_t>(p->nidx, 1);
  s[PNMN_SZ_BLOB] = p->blob_bytes;
  {
    int64_t n_valid = 0;
    for (uint8_t v : p->valid) n_valid += v;
    if (p->input_by_row) n_valid = p->B;
    s[PNMN_SZ_AIN] = (2 * kGuard + std::max<int64_t>(n_valid, 1) * (static_cast<int64_t>(p->m->in_ch / 4) * 256 * 16 * 3 / 2)) / 4;
  }
  return 0;
}

extern "C" int pnmn_plan_stats(const pnmn_plan* p, int64_t* stats) {
  std::memcpy(stats, p->stats, sizeof(p->stats));
  return 0;
}

namespace {

void fill_bases(uint64_t* base, const pnmn_buffers* b, const void* final_out, const void* grad_out) {
  std::memset(base, 0, sizeof(uint64_t) * AR_COUNT);
  base[AR_A16] = reinterpret_cast<uint64_t>(b->arena16);
  base[AR_A18] = reinterpret_cast<uint64_t>(b->arena18);
  base[AR_A22] = reinterpret_cast<uint64_t>(b->arena22);
  base[AR_MAPS] = reinterpret_cast<uint64_t>(b->maps);
  base[AR_DMAPS] = reinterpret_cast<uint64_t>(b->dmaps);
  base[AR_IDX] = reinterpret_cast<uint64_t>(b->idx);
  base[AR_PACKED] = reinterpret_cast<uint64_t>(b->packed);
  base[AR_PARAMS] = reinterpret_cast<uint64_t>(b->params);
  base[AR_GRADS] = reinterpret_cast<uint64_t>(b->grads);
  base[AR_BLOB] = reinterpret_cast<uint64_t>(b->blob);
  base[AR_FINAL] = reinterpret_cast<uint64_t>(final_out);
  base[AR_GRADOUT] = reinterpret_cast<uint64_t>(grad_out);
  base[AR_AIN] = reinterpret_cast<uint64_t>(b->ain);
  base[AR_SCRATCH] = reinterpret_cast<uint64_t>(b->scratch);
}

// ---- optional per-launch timing (bench.py roofline): CUDA events on the launching stream ----------
struct ProfRec { int kind; cudaEvent_t a, b; };
bool g_prof_on = false;
std::vector<ProfRec> g_prof;
enum ProfKind { PK_ELT = 0, PK_CONV0, PK_CONV1, PK_WGRAD, PK_BIAS, PK_PACK, PK_LAYOUT, PK_OTHER, PK_COUNT };
struct ProfScope {
  cudaStream_t st; bool on; ProfRec r;
  ProfScope(int kind, cudaStream_t s) : st(s), on(g_prof_on) {
    if (!on) return;
    r.kind = kind; cudaEventCreate(&r.a); cudaEventCreate(&r.b); cudaEventRecord(r.a, st);
  }
  ~ProfScope() { if (on) { cudaEventRecord(r.b, st); g_prof.push_back(r); } }
};

void resolve_conv(ConvTask& t, const uint64_t* base) {
  for (int i = 0; i < 2; ++i)
    for (int s = 0; s < NSMAX; ++s) resolve(t.in[i][s], base);
  for (int s = 0; s < NSMAX; ++s) { resolve(t.out[s], base); resolve(t.aux[s], base); resolve(t.map_out[s], base); }
  resolve(t.w, base); resolve(t.bias, base); resolve(t.w3, base); resolve(t.b3, base);
}
void resolve_elt(EltTask& t, const uint64_t* base) {
  resolve(t.a, base); resolve(t.b, base); resolve(t.c, base); resolve(t.g, base);
  resolve(t.o, base); resolve(t.o2, base); resolve(t.w, base); resolve(t.dw, base); resolve(t.dw2, base);
  resolve(t.idx, base); resolve(t.scale, base);
}

__global__ void fill_ones_map_kernel(float* map) {
  const int i = threadIdx.x;
  if (i < 256) map[i] = ((i >> 4) < kHW && (i & 15) < kHW) ? 1.f : 0.f;
}

// Side stream (one per device) for work that only has to finish by the end of the backward call: the bias-gradient kernel
// (128-thread CTAs, no shared memory, HBM-bound) runs underneath wgrad_tc_kernel, which keeps one 256-thread CTA per SM busy
// waiting on its operand ring (profiles/r1c: 12 % of the warp slots active).
struct SideStream { cudaStream_t s = nullptr; cudaEvent_t fork = nullptr, join = nullptr; };
std::mutex g_side_mutex;
std::map<int, SideStream> g_side;
SideStream* side_stream() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
  std::lock_guard<std::mutex> lock(g_side_mutex);
  SideStream& ss = g_side[dev];
  if (!ss.s) {
    if (cudaStreamCreateWithFlags(&ss.s, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&ss.fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ss.join, cudaEventDisableTiming) != cudaSuccess) {
      cudaGetLastError();
      ss = SideStream();
      return nullptr;
    }
  }
  return &ss;
}

int run_launches(const pnmn_plan& p, const std::vector<LaunchItem>& ls, const uint8_t* blob, bool bwd, cudaStream_t st) {
  const ConvCfg* cfgs = reinterpret_cast<const ConvCfg*>(blob + p.off_cfg);
  static const bool no_side = std::getenv("PNMN_NO_SIDE_STREAM") != nullptr;
  if (bwd && p.persistent && !g_prof_on && !no_side && ls.size() == 2 && ls[0].kind == LK_WGRAD && ls[1].kind == LK_BIAS) {
    if (SideStream* ss = side_stream()) {
      CUDA_OK(cudaEventRecord(ss->fork, st));
      CUDA_OK(cudaStreamWaitEvent(ss->s, ss->fork, 0));
      CUDA_OK(launch_bias_grad(blob + p.off_bt, ls[1].count, kBiasSplit, ss->s));
      CUDA_OK(cudaEventRecord(ss->join, ss->s));
      CUDA_OK(launch_wgrad(reinterpret_cast<const WgradTask*>(blob + p.off_wt), ls[0].count, 0,
                           reinterpret_cast<int*>(const_cast<uint8_t*>(blob) + p.off_wsync), st));
      CUDA_OK(cudaStreamWaitEvent(st, ss->join, 0));
      return 0;
    }
  }
  const ConvTask* conv = reinterpret_cast<const ConvTask*>(blob + (bwd ? p.off_bconv : p.off_fconv));
  const EltTask* elt = reinterpret_cast<const EltTask*>(blob + (bwd ? p.off_belt : p.off_felt));
  for (const LaunchItem& l : ls) {
    ProfScope prof(l.kind, st);  // LaunchKind values coincide with ProfKind 0..4
    switch (l.kind) {
      case LK_ELT: CUDA_OK(launch_elt(elt + l.off, l.count, st)); break;
      case LK_CONV0:
      case LK_CONV1: CUDA_OK(launch_conv_simt(conv + l.off, l.count, cfgs, st)); break;
      case LK_WGRAD: CUDA_OK(launch_wgrad(reinterpret_cast<const WgradTask*>(blob + p.off_wt), l.count, p.persistent ? 0 : 1,
                                          p.persistent ? reinterpret_cast<int*>(const_cast<uint8_t*>(blob) + p.off_wsync) : nullptr, st)); break;
      case LK_BIAS: CUDA_OK(launch_bias_grad(blob + p.off_bt, l.count, kBiasSplit, st)); break;
      default: break;
    }
  }
  return 0;
}

}  // namespace

static int nmn_forward_impl(pnmn_plan* pp, const pnmn_buffers* bufs, const void* features, int features_half, float* final_out,
                            void* stream);
extern "C" int pnmn_nmn_forward(pnmn_plan* pp, const pnmn_buffers* bufs, const float* features, float* final_out,
                                void* stream) {
  return nmn_forward_impl(pp, bufs, features, 0, final_out, stream);
}
extern "C" int pnmn_nmn_forward_f16(pnmn_plan* pp, const pnmn_buffers* bufs, const void* features_f16, float* final_out,
                                    void* stream) {
  return nmn_forward_impl(pp, bufs, features_f16, 1, final_out, stream);
}
extern "C" int pnmn_round_features_f16(const float* src, void* dst, int64_t n, void* stream) {
  if (n % 4) return fail("pnmn_round_features_f16: n must be a multiple of 4");
  CUDA_OK(launch_round_features_f16(src, dst, n, static_cast<cudaStream_t>(stream)));
  pnmn::count_launches(1);
  return 0;
}
static int nmn_forward_impl(pnmn_plan* pp, const pnmn_buffers* bufs, const void* features, int features_half, float* final_out,
                            void* stream) {
  HostTimer timer(1);
  pnmn_plan& p = *pp;
  const pnmn_model& m = *p.m;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint64_t base[AR_COUNT];
  fill_bases(base, bufs, final_out, nullptr);
  if (p.persistent) {
    // one asynchronous upload of the whole (unresolved) blob, forward + backward records and zeroed flags
    if (!p.pin || !p.pin->pinned) return fail("plan has no page-locked task table (created without a CUDA device?)");
    if (p.uploaded_to == nullptr || p.uploaded_to != bufs->blob) {
      CUDA_OK(cudaMemcpyAsync(bufs->blob, p.pin->p, static_cast<size_t>(p.blob_bytes), cudaMemcpyHostToDevice, st));
      CUDA_OK(cudaEventRecord(p.pin->ev, st));
      p.pin->pending = true;
    }
    BaseTable bt;
    std::memcpy(bt.b, base, sizeof(bt.b));
    CUDA_OK(launch_resolve(static_cast<uint8_t*>(bufs->blob), p.off_ftask, p.ftask.size(), 128, kMaskConv, kMaskElt,
                           p.off_fmeta, sizeof(TaskMeta), bt, st));
  } else {
  // resolve + upload the forward part of the blob (cfgs, conv tasks, elt tasks, pack tasks)
  const int64_t fwd_bytes = p.off_bconv;
  if (p.host_blob.size() < static_cast<size_t>(p.blob_bytes)) p.host_blob.resize(static_cast<size_t>(p.blob_bytes));
  std::memcpy(p.host_blob.data() + p.off_cfg, p.cfgs.data(), p.cfgs.size() * sizeof(ConvCfg));
  std::memcpy(p.host_blob.data() + p.off_pack, m.pack.data(), m.pack.size() * sizeof(PackTask));
  {
    ConvTask* d = reinterpret_cast<ConvTask*>(p.host_blob.data() + p.off_fconv);
    for (size_t i = 0; i < p.fconv.size(); ++i) { d[i] = p.fconv[i]; resolve_conv(d[i], base); }
    EltTask* e = reinterpret_cast<EltTask*>(p.host_blob.data() + p.off_felt);
    for (size_t i = 0; i < p.felt.size(); ++i) { e[i] = p.felt[i]; resolve_elt(e[i], base); }
    TaskRec* ft = reinterpret_cast<TaskRec*>(p.host_blob.data() + p.off_ftask);
    for (size_t i = 0; i < p.ftask.size(); ++i) {
      ft[i] = p.ftask[i];
      if (p.fmeta[i].type == TASK_CONV) resolve_conv(*reinterpret_cast<ConvTask*>(ft[i].b), base);
      else resolve_elt(*reinterpret_cast<EltTask*>(ft[i].b), base);
    }
    if (!p.fmeta.empty()) std::memcpy(p.host_blob.data() + p.off_fmeta, p.fmeta.data(), p.fmeta.size() * sizeof(TaskMeta));
    int64_t* x = reinterpret_cast<int64_t*>(p.host_blob.data() + p.off_xin);
    const int64_t ain_unit = static_cast<int64_t>(m.in_ch / 4) * 256 * 16 * 3 / 2;
    for (int n = 0; n < p.B; ++n) x[n] = p.xin_unit[n] < 0 ? -1 : (kGuard + p.xin_unit[n] * ain_unit) / 4;
  }
  CUDA_OK(cudaMemcpyAsync(bufs->blob, p.host_blob.data(), static_cast<size_t>(fwd_bytes), cudaMemcpyHostToDevice, st));
  }
  if (!features && !p.input_by_row)
    return fail("pnmn_nmn_forward: features == NULL needs a plan created with PNMN_PLAN_INPUT_BY_ROW (after pnmn_nmn_prestage)");
  // pack weights (fp16 MMA tile order); the pack-task table is part of the plan's blob (uploaded above)
  if (features) {
    ProfScope prof(PK_PACK, st);
    CUDA_OK(launch_pack(reinterpret_cast<const PackTask*>(static_cast<const uint8_t*>(bufs->blob) + p.off_pack),
                        static_cast<int>(m.pack.size()), m.total_tiles, bufs->params, bufs->packed, st));
  }
  fill_ones_map_kernel<<<1, 256, 0, st>>>(bufs->maps);
  // features -> planes for every valid sample
  if (features) {
    ProfScope prof_layout(PK_LAYOUT, st);
    CUDA_OK(launch_nchw_to_planes(features, features_half, bufs->ain, p.B, m.in_ch,
                                  reinterpret_cast<const int64_t*>(static_cast<const uint8_t*>(bufs->blob) + p.off_xin), 0, 0, st));
  }
  if (p.persistent) {
    uint8_t* blob = static_cast<uint8_t*>(bufs->blob);
    pnmn::count_launches(5);   // resolve, pack_weights, fill_ones_map, nchw_to_planes, exec_kernel
    ProfScope prof(PK_CONV0, st);
    CUDA_OK(launch_exec(blob + p.off_ftask, reinterpret_cast<const TaskMeta*>(blob + p.off_fmeta),
                        static_cast<int>(p.ftask.size()), reinterpret_cast<const ConvCfg*>(blob + p.off_cfg),
                        reinterpret_cast<int*>(blob + p.off_fsync), reinterpret_cast<int*>(blob + p.off_fsync) + 1,
                        (g_trace && static_cast<int64_t>(p.ftask.size()) <= g_trace_cap) ? g_trace : nullptr, p.exec_ctas, st));
    return 0;
  }
  return run_launches(p, p.flaunch, static_cast<const uint8_t*>(bufs->blob), false, st);
}

extern "C" int pnmn_nmn_backward(pnmn_plan* pp, const pnmn_buffers* bufs, const float* grad_final_out, void* stream) {
  HostTimer timer(2);
  pnmn_plan& p = *pp;
  if (!p.need_grad) return fail("plan was created with need_grad = 0");
  if (!bufs->grads) return fail("grads buffer is NULL");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint64_t base[AR_COUNT];
  fill_bases(base, bufs, nullptr, grad_final_out);
  if (p.persistent) {
    BaseTable bt;
    std::memcpy(bt.b, base, sizeof(bt.b));
    uint8_t* blob = static_cast<uint8_t*>(bufs->blob);
    CUDA_OK(launch_resolve(blob, p.off_btask, p.btask.size(), 128, kMaskConv, kMaskElt, p.off_bmeta, sizeof(TaskMeta), bt, st));
    CUDA_OK(launch_resolve(blob, p.off_inst, p.insts.size(), sizeof(WgradInst), kMaskInst, kMaskInst, -1, 0, bt, st));
    CUDA_OK(launch_resolve(blob, p.off_wt, p.wtasks.size(), sizeof(WgradTask), kMaskWgrad, kMaskWgrad, -1, 0, bt, st));
    CUDA_OK(launch_resolve(blob, p.off_bt, p.btasks.size(), sizeof(BiasGradTaskH), kMaskBias, kMaskBias, -1, 0, bt, st));
  } else {
  {
    ConvTask* d = reinterpret_cast<ConvTask*>(p.host_blob.data() + p.off_bconv);
    for (size_t i = 0; i < p.bconv.size(); ++i) { d[i] = p.bconv[i]; resolve_conv(d[i], base); }
    EltTask* e = reinterpret_cast<EltTask*>(p.host_blob.data() + p.off_belt);
    for (size_t i = 0; i < p.belt.size(); ++i) { e[i] = p.belt[i]; resolve_elt(e[i], base); }
    TaskRec* bt2 = reinterpret_cast<TaskRec*>(p.host_blob.data() + p.off_btask);
    for (size_t i = 0; i < p.btask.size(); ++i) {
      bt2[i] = p.btask[i];
      if (p.bmeta[i].type == TASK_CONV) resolve_conv(*reinterpret_cast<ConvTask*>(bt2[i].b), base);
      else resolve_elt(*reinterpret_cast<EltTask*>(bt2[i].b), base);
    }
    if (!p.bmeta.empty()) std::memcpy(p.host_blob.data() + p.off_bmeta, p.bmeta.data(), p.bmeta.size() * sizeof(TaskMeta));
    WgradInst* wi = reinterpret_cast<WgradInst*>(p.host_blob.data() + p.off_inst);
    for (size_t i = 0; i < p.insts.size(); ++i) { wi[i] = p.insts[i]; resolve(wi[i].dz, base); resolve(wi[i].x, base); }
    WgradTask* wt = reinterpret_cast<WgradTask*>(p.host_blob.data() + p.off_wt);
    for (size_t i = 0; i < p.wtasks.size(); ++i) { wt[i] = p.wtasks[i]; resolve(wt[i].inst, base); resolve(wt[i].dw, base); resolve(wt[i].scale, base); }
    BiasGradTaskH* bt = reinterpret_cast<BiasGradTaskH*>(p.host_blob.data() + p.off_bt);
    for (size_t i = 0; i < p.btasks.size(); ++i) { bt[i] = p.btasks[i]; resolve(bt[i].inst, base); resolve(bt[i].db, base); resolve(bt[i].scale, base); }
  }
  CUDA_OK(cudaMemcpyAsync(static_cast<uint8_t*>(bufs->blob) + p.off_bconv, p.host_blob.data() + p.off_bconv,
                          static_cast<size_t>(p.blob_bytes - p.off_bconv), cudaMemcpyHostToDevice, st));
  }
  CUDA_OK(cudaMemsetAsync(bufs->dmaps, 0, static_cast<size_t>(p.nmaps) * 1024, st));
  CUDA_OK(launch_loss_scale(grad_final_out, static_cast<size_t>(p.B) * 128 * 196, bufs->scratch, st));
  if (p.persistent) {
    uint8_t* blob = static_cast<uint8_t*>(bufs->blob);
    pnmn::count_launches(9);   // 4 x resolve, amax + loss scale, exec_kernel, wgrad_tc, bias_grad
    {
      ProfScope prof(PK_CONV0, st);
      CUDA_OK(launch_exec(blob + p.off_btask, reinterpret_cast<const TaskMeta*>(blob + p.off_bmeta),
                          static_cast<int>(p.btask.size()), reinterpret_cast<const ConvCfg*>(blob + p.off_cfg),
                          reinterpret_cast<int*>(blob + p.off_bsync), reinterpret_cast<int*>(blob + p.off_bsync) + 1,
                          (g_trace && static_cast<int64_t>(p.ftask.size() + p.btask.size()) <= g_trace_cap)
                              ? g_trace + kTraceW * p.ftask.size() : nullptr, p.exec_ctas, st));
    }
  }
  return run_launches(p, p.blaunch, static_cast<const uint8_t*>(bufs->blob), true, st);
}

// per-task timestamps of the persistent executor: 8 int64 per task, forward tasks first, then backward
extern "C" int pnmn_debug_set_trace(void* device_buffer, int64_t capacity_tasks) {
  g_trace = static_cast<long long*>(device_buffer);
  g_trace_cap = capacity_tasks;
  return 0;
}

// task metadata of the persistent executor's lists (offline schedule analysis, scripts/sim_sched.py):
// pass 0 = forward, 1 = backward; out receives 16 int32 per task {type, n_deps, deps[10], then for a conv task
// n_samp, n_mt, MMAs per accumulator, cfg flags / for an elementwise task op, part, 0, 0}; returns the task count
extern "C" int64_t pnmn_debug_plan_meta(const pnmn_plan* p, int pass, int32_t* out, int64_t cap_tasks) {
  const std::vector<TaskMeta>& m = pass == 0 ? p->fmeta : p->bmeta;
  const std::vector<TaskRec>& r = pass == 0 ? p->ftask : p->btask;
  const int64_t n = static_cast<int64_t>(m.size());
  for (int64_t i = 0; out && i < std::min(n, cap_tasks); ++i) {
    int32_t* o = out + i * 16;
    std::memcpy(o, &m[i], sizeof(TaskMeta));
    if (m[i].type == TASK_CONV) {
      const ConvTask& t = *reinterpret_cast<const ConvTask*>(r[i].b);
      const ConvCfg& c = p->cfgs[t.cfg];
      o[12] = t.n_samp; o[13] = t.n_mt; o[14] = c.n_kb * c.ntaps; o[15] = c.flags;
    } else {
      const EltTask& t = *reinterpret_cast<const EltTask*>(r[i].b);
      o[12] = t.op; o[13] = t.part; o[14] = 0; o[15] = 0;
    }
  }
  return n;
}

// The raw (unresolved) records of a plan: pass 0 / 1 = the 128-byte task records of the forward / backward list, pass 2 =
// the convolution configurations.  Returns the record count; `out` (may be NULL) receives min(count, cap) records.  Tests
// use it to check that a PNMN_PLAN_FORWARD_HALF plan and the full plan of the same programs agree record for record.
extern "C" int64_t pnmn_debug_plan_records(const pnmn_plan* p, int pass, void* out, int64_t cap_records) {
  if (pass == 2) {
    const int64_t n = static_cast<int64_t>(p->cfgs.size());
    if (out) std::memcpy(out, p->cfgs.data(), static_cast<size_t>(std::min(n, cap_records)) * sizeof(ConvCfg));
    return n;
  }
  const std::vector<TaskRec>& r = pass ? p->btask : p->ftask;
  const int64_t n = static_cast<int64_t>(r.size());
  if (out) std::memcpy(out, r.data(), static_cast<size_t>(std::min(n, cap_records)) * sizeof(TaskRec));
  return n;
}

// 1-channel module outputs (attention maps) of a plan: 4 int32 per record {sample, index of the module call inside the
// sample's program (execution order, `scene` / skipped tokens not counted), token id, map unit}; map unit u lives at
// pnmn_buffers.maps + 256 * u as a 16 x 16 grid whose top-left 14 x 14 block is the map.  Returns the record count.
extern "C" int64_t pnmn_debug_plan_maps(const pnmn_plan* p, int32_t* out, int64_t cap_records) {
  const int64_t n = static_cast<int64_t>(p->map_trace.size() / 4);
  if (out) std::memcpy(out, p->map_trace.data(), static_cast<size_t>(std::min(n, cap_records)) * 4 * sizeof(int32_t));
  return n;
}

// accumulated host-side milliseconds: {plan_create, forward, backward, #plans}; reading clears
extern "C" int pnmn_debug_host_times(double* ms) {
  for (int i = 0; i < 4; ++i) { ms[i] = g_host_ms[i]; g_host_ms[i] = 0; }
  return 0;
}

extern "C" int pnmn_profile_enable(int on) {
  g_prof_on = on != 0;
  return 0;
}
// ms[k], launches[k] for k in {elt, conv<2,2>, conv<1,3>, wgrad, bias_grad, pack, layout, other}; clears the log
extern "C" int pnmn_profile_read(double* ms, int64_t* launches) {
  for (int k = 0; k < PK_COUNT; ++k) { ms[k] = 0; launches[k] = 0; }
  CUDA_OK(cudaDeviceSynchronize());
  for (ProfRec& r : g_prof) {
    float t = 0.f;
    cudaEventElapsedTime(&t, r.a, r.b);
    ms[r.kind] += t; launches[r.kind] += 1;
    cudaEventDestroy(r.a); cudaEventDestroy(r.b);
  }
  g_prof.clear();
  return 0;
}

// bf16 hi/lo split of an fp32 matrix for the classifier's split-precision library GEMMs (see layout.cu)

// ReLU + 2x2 max-pool + flatten between the classifier's two GEMMs (nmn.py:77-79) and its backward (see layout.cu)
extern "C" int pnmn_relu_pool_fwd(const float* y, float* pooled, void* code, int64_t B, int64_t C, void* stream) {
  if (!y || !pooled || !code || B < 0 || C <= 0 || C % 64 != 0) return fail("pnmn_relu_pool_fwd: bad arguments (C must be a multiple of 64)");
  CUDA_OK(launch_relu_pool_fwd(y, nullptr, pooled, static_cast<uint8_t*>(code), static_cast<int>(B), static_cast<int>(C), static_cast<cudaStream_t>(stream)));
  pnmn::count_launches(1);
  return 0;
}
// same with the 1x1 conv's bias added to y first (y is then the bias-free GEMM output)
extern "C" int pnmn_relu_pool_fwd_bias(const float* y, const float* bias, float* pooled, void* code, int64_t B, int64_t C, void* stream) {
  if (!y || !bias || !pooled || !code || B < 0 || C <= 0 || C % 64 != 0) return fail("pnmn_relu_pool_fwd_bias: bad arguments (C must be a multiple of 64)");
  CUDA_OK(launch_relu_pool_fwd(y, bias, pooled, static_cast<uint8_t*>(code), static_cast<int>(B), static_cast<int>(C), static_cast<cudaStream_t>(stream)));
  pnmn::count_launches(1);
  return 0;
}
extern "C" int pnmn_relu_pool_bwd(const float* g, const void* code, float* gy, int64_t B, int64_t C, void* stream) {
  if (!g || !gy || !code || B < 0 || C <= 0 || C % 64 != 0) return fail("pnmn_relu_pool_bwd: bad arguments (C must be a multiple of 64)");
  CUDA_OK(launch_relu_pool_bwd(g, static_cast<const uint8_t*>(code), gy, static_cast<int>(B), static_cast<int>(C), static_cast<cudaStream_t>(stream)));
  pnmn::count_launches(1);
  return 0;
}

// the same backward with the gradient written as the bf16 (hi, lo) pair [2][B*196][C] of the split-precision GEMMs
// bf16 (hi, lo) pair [2][n] of an fp32 array (n a multiple of 4)

// ---- bring-up entry points -----------------------------------------------------------------------
namespace {
template <class T>
int upload(const void* host, size_t n, T** dev) {
  CUDA_OK(cudaMalloc(dev, std::max<size_t>(n, 1) * sizeof(T)));
  CUDA_OK(cudaMemcpy(*dev, host, n * sizeof(T), cudaMemcpyHostToDevice));
  return 0;
}
}  // namespace

extern "C" int pnmn_debug_launch_conv(const void* tasks_host, int n_tasks, const void* cfgs_host, int n_cfgs,
                                      int variant, int impl_simt, void* stream) {
  (void)variant;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ConvTask* dt = nullptr; ConvCfg* dc = nullptr;
  if (upload(tasks_host, n_tasks, &dt)) return 1;
  if (upload(cfgs_host, n_cfgs, &dc)) return 1;
  if (impl_simt) {
    CUDA_OK(launch_conv_simt(dt, n_tasks, dc, st));
  } else {
    // run the tasks through the persistent executor kernel without dependencies; the kernel handles at most two
    // accumulators per task, so wider test tasks are cut into per-M-tile pieces first
    {
      std::vector<ConvTask> host(static_cast<const ConvTask*>(tasks_host), static_cast<const ConvTask*>(tasks_host) + n_tasks);
      std::vector<ConvTask> cut;
      for (const ConvTask& t : host) {
        if (t.n_samp * t.n_mt <= 2) { cut.push_back(t); continue; }
        for (int m = t.mt0; m < t.mt0 + t.n_mt; ++m) { ConvTask u = t; u.mt0 = m; u.n_mt = 1; cut.push_back(u); }
      }
      cudaFree(dt);
      dt = nullptr;
      n_tasks = static_cast<int>(cut.size());
      if (upload(cut.data(), cut.size(), &dt)) return 1;
    }
    std::vector<TaskMeta> metas(n_tasks);
    for (auto& mm : metas) { mm.type = TASK_CONV; mm.n_deps = 0; for (int k = 0; k < kMaxDeps; ++k) mm.deps[k] = -1; }
    TaskMeta* dm = nullptr; int* sync = nullptr;
    if (upload(metas.data(), metas.size(), &dm)) return 1;
    CUDA_OK(cudaMalloc(&sync, 4 * (n_tasks + 1)));
    CUDA_OK(cudaMemsetAsync(sync, 0, 4 * (n_tasks + 1), st));
    CUDA_OK(launch_exec(reinterpret_cast<const uint8_t*>(dt), dm, n_tasks, dc, sync, sync + 1, nullptr, 0, st));
    CUDA_OK(cudaStreamSynchronize(st));
    cudaFree(dm); cudaFree(sync);
  }
  CUDA_OK(cudaStreamSynchronize(st));
  cudaFree(dt); cudaFree(dc);
  return 0;
}
extern "C" int pnmn_debug_launch_wgrad(const void* tasks_host, int n_tasks, const void* insts_host, int n_insts,
                                       int impl_simt, void* stream) {
  // task.inst holds an INDEX into insts_host; it is rebased onto the uploaded table here
  WgradInst* di = nullptr;
  if (upload(insts_host, n_insts, &di)) return 1;
  std::vector<WgradTask> t(static_cast<const WgradTask*>(tasks_host), static_cast<const WgradTask*>(tasks_host) + n_tasks);
  for (auto& x : t) x.inst = di + reinterpret_cast<uint64_t>(x.inst);
  WgradTask* dt = nullptr;
  if (upload(t.data(), t.size(), &dt)) return 1;
  CUDA_OK(launch_wgrad(dt, n_tasks, impl_simt, nullptr, static_cast<cudaStream_t>(stream)));
  CUDA_OK(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
  cudaFree(dt); cudaFree(di);
  return 0;
}
extern "C" int pnmn_debug_pack(const void* pack_tasks_host, int n_tasks, int total_tiles, const float* params,
                               void* packed, void* stream) {
  PackTask* dt = nullptr;
  if (upload(pack_tasks_host, n_tasks, &dt)) return 1;
  CUDA_OK(launch_pack(dt, n_tasks, total_tiles, params, packed, static_cast<cudaStream_t>(stream)));
  CUDA_OK(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
  cudaFree(dt);
  return 0;
}
extern "C" int pnmn_debug_nchw_to_planes(const float* src, float* dst, int batch, int channels,
                                         int64_t dst_sample_stride, void* stream) {
  std::vector<int64_t> off(batch);
  for (int b = 0; b < batch; ++b) off[b] = b * dst_sample_stride;
  int64_t* d = nullptr;
  if (upload(off.data(), off.size(), &d)) return 1;
  CUDA_OK(launch_nchw_to_planes(src, 0, dst, batch, channels, d, 0, 0, static_cast<cudaStream_t>(stream)));
  CUDA_OK(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
  cudaFree(d);
  return 0;
}
extern "C" int pnmn_debug_launch_elt(const void* tasks_host, int n_tasks, void* stream) {
  EltTask* dt = nullptr;
  if (upload(tasks_host, n_tasks, &dt)) return 1;
  CUDA_OK(launch_elt(dt, n_tasks, static_cast<cudaStream_t>(stream)));
  CUDA_OK(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
  cudaFree(dt);
  return 0;
}
