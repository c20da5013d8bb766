// C ABI of the LSTM seq2seq path (include/pnmn.h: pnmn_pg_*): workspace layout and the launch sequences of
// ProgramGenerator.forward / backward.
//
// Replaces Seq2SeqBase.forward + _forward_loop + _trim_predictions + _get_loss
// (probnmn/modules/seq2seq_base.py:101-341) together with AllenNLP's _encode / _init_decoder_state /
// _prepare_output_projections underneath it, and the autograd pass through them.  Every launch goes to the
// caller's stream; nothing synchronises with the host (the reference syncs once per decoding step and once per
// row, seq2seq_base.py:188,286).
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <tuple>
#include <algorithm>
#include <functional>

#include "../../include/pnmn.h"
#include "seq2seq.h"

using namespace pnmn;

extern "C" const char* pnmn_last_error(void);
namespace pnmn { void set_last_error(const std::string& s); void count_launches(int n); }

namespace {

int fail(const std::string& s) {
  pnmn::set_last_error(s);
  return 1;
}
#define CUDA_OK(x)                                                                        \
  do {                                                                                    \
    cudaError_t e_ = (x);                                                                 \
    if (e_ != cudaSuccess) return fail(std::string(#x) + ": " + cudaGetErrorString(e_)); \
  } while (0)

// CUDA-core twins of the LSTM GEMMs exist in bring-up builds only (make BRINGUP=1, then PNMN_PG_SIMT=1 selects them)
bool use_simt() {
#ifdef PNMN_BRINGUP
  static const bool v = std::getenv("PNMN_PG_SIMT") != nullptr;
  return v;
#else
  return false;
#endif
}

// byte offsets into the caller's workspace (all 256-byte aligned)
struct Layout {
  SeqDims d;
  int64_t slotf, slotop, slotg, slotdg;   // per-step sizes: fp32 state (floats), operand K=256 (halves, hi+lo),
                                          // gates (floats), gate-gradient operand (halves, hi+lo)
  int64_t src, src_len, tgt, inp, pred, label, logp, lse, coef, logits, attn_p, dlogits;
  int64_t P0, Pd, P1, dP0, dPd, dP1;
  int64_t pk_hh0, pk_1, pk_d, pkT_0, pkT_1, pkT_d;   // packed weights (halves offsets inside `packed`)
  int64_t packed;
  int64_t zeros;
  int64_t h0f, c0f, h0op, out0op, g0;      // encoder layer 0
  int64_t h1f, c1f, h1op, g1, enc;         // encoder layer 1 (+ decoder slots appended to h1f / h1op)
  int64_t cdf, attop, gd;                  // decoder
  int64_t dh, dc, dh0, dc0, datt, denc, dout0, dgd, dg1, dg0, scale, dhp;
  int64_t seed, gstage, gws, extent, rowmode, gemm_ws, gemm_ws_floats;   // Philox key; incoming gradient padded to Bp; parameter gradients of this call
  int64_t total;
};

int64_t param_extent(const pnmn_pg_desc* m);

Layout make_layout(const pnmn_pg_desc* m, int B, int Tq, int Tp, int S, bool need_grad) {
  Layout L;
  std::memset(&L, 0, sizeof(L));
  SeqDims& d = L.d;
  d.B = B; d.Bp = (B + 127) / 128 * 128; d.Tq = Tq; d.Ts = Tq + 1; d.Tp = Tp; d.S = S;
  d.Vs = m->vocab_src; d.Vt = m->vocab_tgt;
  L.slotf = static_cast<int64_t>(d.Bp) * kSH;
  L.slotop = 2 * L.slotf;
  L.slotg = static_cast<int64_t>(d.Bp) * kSG;
  L.slotdg = 2 * L.slotg;
  int64_t o = 0;
  auto take = [&](int64_t bytes) { const int64_t r = o; o += (bytes + 255) / 256 * 256; return r; };
  // every size is a function of the PADDED batch (Bp) only: a caller whose batch size varies from call to call (the
  // supervised / unsupervised split of the joint-training step) keeps one zero-filled workspace per padded size; the
  // kernels index with the real B, the backward row kernels zero the gradient operands of the padding rows
  const int64_t Bq = d.Bp;
  const int64_t SB = static_cast<int64_t>(S) * Bq;
  L.src = take(4ll * Bq * d.Ts); L.src_len = take(4ll * Bq); L.tgt = take(4ll * Bq * (Tp + 2));
  L.inp = take(4 * SB); L.pred = take(4 * SB); L.label = take(4 * SB);
  L.logp = take(4 * SB); L.lse = take(4 * SB); L.coef = take(4 * SB);
  L.logits = take(4 * SB * d.Vt); L.attn_p = take(4 * SB * d.Ts);
  L.P0 = take(4ll * d.Vs * kSG); L.Pd = take(4ll * d.Vt * kSG); L.P1 = take(4ll * kSG);
  // packed weights: forward tiles, then the transposed tiles of the data-gradient GEMMs
  int64_t ph = 0;
  auto takeh = [&](int64_t n, int64_t k) { const int64_t r = ph; ph += 2 * n * k; return r; };
  L.pk_hh0 = takeh(kSG, kSH); L.pk_1 = takeh(kSG, 2 * kSH); L.pk_d = takeh(kSG, 2 * kSH);
  L.pkT_0 = takeh(kSH, kSG); L.pkT_1 = takeh(2 * kSH, kSG); L.pkT_d = takeh(2 * kSH, kSG);
  L.packed = take(2 * ph);
  L.zeros = take(4 * L.slotf);
  L.h0f = take(4 * L.slotf * (d.Ts + 1)); L.c0f = take(4 * L.slotf * (d.Ts + 1));
  L.h0op = take(2 * L.slotop * (d.Ts + 1)); L.out0op = take(2 * L.slotop * d.Ts);
  // layer 1 state slots 0..Ts, directly followed by the decoder's slots 1..S (decoder slot 0 IS layer-1 slot Ts)
  L.h1f = take(4 * L.slotf * (d.Ts + 1 + S)); L.c1f = take(4 * L.slotf * (d.Ts + 1));
  L.h1op = take(2 * L.slotop * (d.Ts + 1 + S));
  L.enc = take(4ll * Bq * d.Ts * kSH);
  L.cdf = take(4 * L.slotf * (S + 1)); L.attop = take(2 * L.slotop * S);
  L.scale = take(256);
  L.seed = take(256);
  L.rowmode = take(4ll * d.Bp);
  // scratch of the one tensor-core GEMM that computes the logits of every step of a teacher-forced pass (pnmn_gemm_split)
  L.gemm_ws_floats = pnmn_gemm_split_workspace(S * d.Bp, d.Vt, kSH);
  L.gemm_ws = take(4 * L.gemm_ws_floats);
  L.extent = param_extent(m);
  if (need_grad) {
    L.gstage = take(4ll * d.Bp);
    L.gws = take(4 * L.extent);
    L.g0 = take(4 * L.slotg * d.Ts); L.g1 = take(4 * L.slotg * d.Ts); L.gd = take(4 * L.slotg * S);
    L.dlogits = take(4ll * S * d.Bp * d.Vt);
    L.dhp = take(4 * L.slotf * S);
    L.dP0 = take(4ll * d.Vs * kSG); L.dPd = take(4ll * d.Vt * kSG); L.dP1 = take(4ll * kSG);
    L.dh = take(4 * L.slotf); L.dc = take(4 * L.slotf); L.datt = take(4 * L.slotf);
    L.dh0 = take(4 * L.slotf); L.dc0 = take(4 * L.slotf);
    L.denc = take(4ll * Bq * d.Ts * kSH); L.dout0 = take(4 * L.slotf * d.Ts);
    L.dgd = take(2 * L.slotdg * S); L.dg1 = take(2 * L.slotdg * d.Ts); L.dg0 = take(2 * L.slotdg * d.Ts);
  }
  L.total = o;
  return L;
}

int check_dims(const pnmn_pg_desc* m, int B, int Tq, int Tp, int S, bool teacher) {
  if (!m) return fail("pnmn_pg: NULL model description");
  if (m->hidden != kSH) return fail("pnmn_pg: the B200 seq2seq kernels are built for input_size = hidden_size = 256");
  if (m->num_layers != 2) return fail("pnmn_pg: the encoder must have 2 layers");
  if (m->vocab_tgt < 4 || m->vocab_tgt > kSMaxV) return fail("pnmn_pg: target vocabulary must have 4..128 entries");
  if (m->vocab_src < 4 || m->vocab_src > kSMaxV) return fail("pnmn_pg: source vocabulary must have 4..128 entries");
  if (B < 1) return fail("pnmn_pg: empty batch");
  if (Tq < 0 || Tq + 1 > kSMaxT) return fail("pnmn_pg: source sequences longer than 63 tokens are not supported");
  if (S < 1) return fail("pnmn_pg: at least one decoding step");
  if (teacher && S != Tp + 1) return fail("pnmn_pg: with targets the number of decoding steps must be target_length + 1");
  return 0;
}

template <class T>
T* at(void* ws, int64_t off) { return reinterpret_cast<T*>(static_cast<uint8_t*>(ws) + off); }

// floats spanned by the model's parameters inside the flat buffer (the layout is the caller's; only the offsets are known)
int64_t param_extent(const pnmn_pg_desc* m) {
  const int64_t H = m->hidden, G = 4 * H;
  int64_t e = 0;
  auto up = [&](int64_t off, int64_t n) { e = std::max(e, off + n); };
  up(m->src_embed, m->vocab_src * H); up(m->tgt_embed, m->vocab_tgt * H);
  for (int l = 0; l < 2; ++l) { up(m->enc_w_ih[l], G * H); up(m->enc_w_hh[l], G * H); up(m->enc_b_ih[l], G); up(m->enc_b_hh[l], G); }
  up(m->dec_w_ih, G * 2 * H); up(m->dec_w_hh, G * H); up(m->dec_b_ih, G); up(m->dec_b_hh, G);
  up(m->out_w, m->vocab_tgt * H); up(m->out_b, m->vocab_tgt);
  return (e + 63) / 64 * 64;
}

// ---- CUDA graphs --------------------------------------------------------------------------------------------------------
// A pass is ~190 (forward) / ~250 (backward) dependent launches of 4-16 us kernels: issued one by one they cost the host
// ~3.5 us each, and the joint-training step (four such passes per iteration) becomes host-bound.  Everything between the
// token preparation and the finalisation touches only the workspace and the flat parameter buffer and depends on the
// call only through data in the workspace (tokens, lengths, the Philox key, the staged incoming gradient), so the launch
// sequence is captured ONCE per (workspace, parameters, shape) into a CUDA graph -- on a private stream, with the same
// programmatic-dependent-launch edges -- and replayed with one cudaGraphLaunch on the caller's stream afterwards.
// PNMN_PG_NOGRAPH=1 disables it; a failed capture falls back to plain launches for that key.
struct KeyLess {
  bool operator()(const GraphKey& a, const GraphKey& o) const {
    return std::tie(a.ws, a.params, a.device, a.pass, a.Bp, a.Tq, a.Tp, a.S, a.sampling, a.teacher, a.need_grad, a.Vs, a.Vt) <
           std::tie(o.ws, o.params, o.device, o.pass, o.Bp, o.Tq, o.Tp, o.S, o.sampling, o.teacher, o.need_grad, o.Vs, o.Vt);
  }
};
struct GraphEntry { int calls = 0; cudaGraphExec_t exec = nullptr; bool failed = false; };

std::mutex g_graph_mutex;
std::map<GraphKey, GraphEntry, KeyLess> g_graphs;
std::map<int, cudaStream_t> g_capture_streams;
long long g_graph_launches = 0;

// ---- independent branches of a pass ---------------------------------------------------------------------------------------
// The tail of the backward pass (five weight-gradient GEMMs, three table gradients, the small embedding / projection GEMMs) and
// the table / packing prologue of the forward pass are mutually independent launches of 15-60 us that do not fill the device
// one at a time.  They are issued on library-owned side streams forked from / joined back into the pass's stream with events:
// captured, that makes parallel branches of the CUDA graph; issued directly (first call, PNMN_PG_NOGRAPH=1) it is plain
// multi-stream concurrency.  PNMN_PG_NOFORK=1 keeps everything on the caller's stream.
constexpr int kForkStreams = 4, kForkEvents = 12;
struct ForkSet {
  cudaStream_t s[kForkStreams];
  cudaEvent_t ev[kForkEvents];
};
std::map<int, ForkSet*> g_forks;
ForkSet* fork_set() {
  static const bool off = std::getenv("PNMN_PG_NOFORK") != nullptr;
  if (off) return nullptr;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
  std::lock_guard<std::mutex> lock(g_graph_mutex);
  auto it = g_forks.find(dev);
  if (it != g_forks.end()) return it->second;
  ForkSet* f = new ForkSet();
  bool ok = true;
  for (int i = 0; i < kForkStreams; ++i) ok = ok && cudaStreamCreateWithFlags(&f->s[i], cudaStreamNonBlocking) == cudaSuccess;
  for (int i = 0; i < kForkEvents; ++i) ok = ok && cudaEventCreateWithFlags(&f->ev[i], cudaEventDisableTiming) == cudaSuccess;
  if (!ok) { cudaGetLastError(); delete f; f = nullptr; }
  g_forks[dev] = f;
  return f;
}
// side stream i continues from everything issued on `st` so far / `st` continues after everything issued on side stream i
struct Forker {
  ForkSet* f;
  cudaStream_t st;
  int next_ev = 0;
  cudaStream_t fork(int i) {
    if (!f || next_ev >= kForkEvents) return st;
    cudaEvent_t e = f->ev[next_ev++];
    if (cudaEventRecord(e, st) != cudaSuccess || cudaStreamWaitEvent(f->s[i], e, 0) != cudaSuccess) return st;
    return f->s[i];
  }
  int join(cudaStream_t side) {
    if (side == st) return 0;
    if (!f || next_ev >= kForkEvents) return 1;
    cudaEvent_t e = f->ev[next_ev++];
    if (cudaEventRecord(e, side) != cudaSuccess || cudaStreamWaitEvent(st, e, 0) != cudaSuccess) return 1;
    return 0;
  }
};

std::mutex g_capture_mutex;   // one capture at a time (the capture stream and the fork streams are per device, not per caller)

bool graphs_enabled() {
  static const bool v = std::getenv("PNMN_PG_NOGRAPH") == nullptr;
  return v;
}

int run_graphed(GraphKey key, cudaStream_t st, const std::function<int(cudaStream_t)>& body) {
  if (!graphs_enabled()) return body(st);
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return body(st);
  key.device = dev;
  GraphEntry* e;
  cudaStream_t cap;
  {
    std::lock_guard<std::mutex> lock(g_graph_mutex);
    if (g_graphs.size() > 256) {   // bounded: forget everything (workspaces of dead shapes), graphs are re-captured on demand
      for (auto& kv : g_graphs) if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
      g_graphs.clear();
    }
    e = &g_graphs[key];
    cudaStream_t& c = g_capture_streams[dev];
    if (!c && cudaStreamCreateWithFlags(&c, cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); e->failed = true; }
    cap = c;
  }
  if (e->failed) return body(st);
  if (e->exec) {
    CUDA_OK(cudaGraphLaunch(e->exec, st));
    ++g_graph_launches;
    return 0;
  }
  if (e->calls++ == 0) return body(st);   // first call: plain launches (one-time function attributes are set here)
  std::lock_guard<std::mutex> capture_lock(g_capture_mutex);
  if (cudaStreamBeginCapture(cap, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
    cudaGetLastError();
    e->failed = true;
    return body(st);
  }
  const int rc = body(cap);
  cudaGraph_t graph = nullptr;
  const cudaError_t err = cudaStreamEndCapture(cap, &graph);
  if (rc != 0 || err != cudaSuccess || !graph) {
    cudaGetLastError();
    if (graph) cudaGraphDestroy(graph);
    e->failed = true;
    return body(st);
  }
  const cudaError_t ierr = cudaGraphInstantiate(&e->exec, graph, 0);
  cudaGraphDestroy(graph);
  if (ierr != cudaSuccess) {
    cudaGetLastError();
    e->exec = nullptr;
    e->failed = true;
    return body(st);
  }
  CUDA_OK(cudaGraphLaunch(e->exec, st));
  return 0;
}

}  // namespace

// {cached keys, keys with an instantiated graph, keys whose capture failed, graph launches so far}
extern "C" int pnmn_debug_graph_stats(int64_t* out) {
  std::lock_guard<std::mutex> lock(g_graph_mutex);
  out[0] = static_cast<int64_t>(g_graphs.size()); out[1] = 0; out[2] = 0; out[3] = g_graph_launches;
  for (auto& kv : g_graphs) { out[1] += kv.second.exec != nullptr; out[2] += kv.second.failed; }
  return 0;
}

namespace pnmn {
int run_graphed_pass(const GraphKey& key, cudaStream_t st, const std::function<int(cudaStream_t)>& body) {
  return run_graphed(key, st, body);
}
}  // namespace pnmn

extern "C" int64_t pnmn_pg_workspace_bytes(const pnmn_pg_desc* m, int batch, int tq, int tp, int steps, int need_grad) {
  if (check_dims(m, batch, tq, tp, steps, false)) return -1;
  return make_layout(m, batch, tq, tp, steps, need_grad != 0).total;
}

// byte offsets of a few workspace buffers, for the parity tests: {enc (fp32 [B][Ts][256]), h0f, h1f, c1f, cdf, P0, Pd,
// attn_p, logits, slot bytes of the fp32 state arrays, Ts, Bp}
extern "C" int pnmn_pg_debug_layout(const pnmn_pg_desc* m, int batch, int tq, int tp, int steps, int need_grad, int64_t* out) {
  if (check_dims(m, batch, tq, tp, steps, false)) return 1;
  const Layout L = make_layout(m, batch, tq, tp, steps, need_grad != 0);
  const int64_t v[12] = {L.enc, L.h0f, L.h1f, L.c1f, L.cdf, L.P0, L.Pd, L.attn_p, L.logits, 4 * L.slotf, L.d.Ts, L.d.Bp};
  for (int i = 0; i < 12; ++i) out[i] = v[i];
  return 0;
}

static int pg_forward_impl(const pnmn_pg_desc* m, const float* params, const int64_t* source, const int64_t* target,
                           const uint8_t* row_teacher, int free_steps, int batch, int tq, int tp, int steps, int sampling,
                           uint64_t seed, int need_grad, void* ws, int64_t* raw_predictions, int64_t* predictions, float* loss,
                           float* logits_out, void* stream) {
  const bool teacher = target != nullptr;
  if (check_dims(m, batch, tq, tp, steps, teacher)) return 1;
  if (!params || !source || !ws || !raw_predictions || !predictions || !loss) return fail("pnmn_pg_forward: NULL buffer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Layout L = make_layout(m, batch, tq, tp, steps, need_grad != 0);
  SeqDims& d = L.d;
  d.teacher = teacher ? 1 : 0;
  d.sampling = sampling ? 1 : 0;
  d.free_S = row_teacher ? free_steps : d.S;
  const bool simt = use_simt();
  const int MT = d.Bp / 128;
  __half* packed = at<__half>(ws, L.packed);

  const int rows = batch;
  d.B = d.Bp;   // every kernel runs on whole 128-row tiles; rows >= `rows` are empty sequences (launch_prepare_tokens)
  CUDA_OK(launch_prepare_tokens(source, target, row_teacher, d, rows, seed, at<unsigned long long>(ws, L.seed), at<int>(ws, L.src),
                                at<int>(ws, L.src_len), at<int>(ws, L.tgt), at<int>(ws, L.rowmode), st));
  // (a mixed call launches different kernels than a purely teacher-forced one of the same shape: part of the key)
  GraphKey key{ws, params, 0, 0, d.Bp, d.Tq, d.Tp, d.S, d.sampling | (row_teacher ? 2 : 0), d.teacher, need_grad != 0, d.Vs, d.Vt};
  if (simt) key.pass = 2;
  const int rc = run_graphed(key, st, [&](cudaStream_t st) -> int {
  Forker fk{fork_set(), st};
  cudaStream_t s_p0 = fk.fork(0), s_pd = fk.fork(1);

  // ---- weights -> split fp16 tiles ------------------------------------------------------------------------
  {
    PackJobs J;
    std::memset(&J, 0, sizeof(J));
    auto job = [](int N, int K, int mode, int split, int64_t s0, int ld0, int64_t s1, int ld1, int64_t dst) {
      PackJob j; j.N = N; j.K = K; j.mode = mode; j.split = split; j.src0 = s0; j.src1 = s1; j.ld0 = ld0; j.ld1 = ld1; j.dst = dst;
      return j;
    };
    J.j[0] = job(kSG, kSH, 0, kSH, m->enc_w_hh[0], kSH, m->enc_w_hh[0], kSH, L.pk_hh0);
    J.j[1] = job(kSG, 2 * kSH, 0, kSH, m->enc_w_ih[1], kSH, m->enc_w_hh[1], kSH, L.pk_1);
    J.j[2] = job(kSG, 2 * kSH, 0, kSH, m->dec_w_ih, 2 * kSH, m->dec_w_hh, kSH, L.pk_d);   // cat(attended, embedding): attended first
    J.j[3] = job(kSH, kSG, 1, kSH, m->enc_w_hh[0], kSH, m->enc_w_hh[0], kSH, L.pkT_0);
    J.j[4] = job(2 * kSH, kSG, 1, kSH, m->enc_w_ih[1], kSH, m->enc_w_hh[1], kSH, L.pkT_1);
    J.j[5] = job(2 * kSH, kSG, 1, kSH, m->dec_w_ih, 2 * kSH, m->dec_w_hh, kSH, L.pkT_d);
    CUDA_OK(launch_pack_seq(J, need_grad ? 6 : 3, params, packed, st));
  }
  // ---- input-projection tables: P[v] = Emb[v] . W_ih^T + b_ih + b_hh (side branches, next to the weight packing; the
  // decoder's table is only needed after the encoder) ---------------------------------------------------------------
  {
    SimtGemm g;
    std::memset(&g, 0, sizeof(g));
    g.alpha = 1.f; g.N = kSG; g.ldc = kSG; g.sak = 1; g.sbk = 1;
    g.A = params + m->src_embed; g.sam = kSH; g.M = d.Vs; g.K = kSH;
    g.B = params + m->enc_w_ih[0]; g.sbn = kSH;
    g.bias0 = params + m->enc_b_ih[0]; g.bias1 = params + m->enc_b_hh[0]; g.C = at<float>(ws, L.P0);
    CUDA_OK(launch_simt_gemm(g, s_p0));
    g.A = params + m->tgt_embed; g.M = d.Vt;
    g.B = params + m->dec_w_ih + kSH; g.sbn = 2 * kSH;
    g.bias0 = params + m->dec_b_ih; g.bias1 = params + m->dec_b_hh; g.C = at<float>(ws, L.Pd);
    CUDA_OK(launch_simt_gemm(g, s_pd));
    g.M = 1; g.K = 0; g.bias0 = params + m->enc_b_ih[1]; g.bias1 = params + m->enc_b_hh[1]; g.C = at<float>(ws, L.P1);
    CUDA_OK(launch_simt_gemm(g, s_p0));
  }

  if (fk.join(s_p0)) return fail("pnmn_pg_forward: stream join failed");
  GemmArgs g;
  std::memset(&g, 0, sizeof(g));
  g.B = d.B; g.chunks_per_src = 4; g.a_K[0] = g.a_K[1] = kSH; g.a_lo[0] = g.a_lo[1] = L.slotf; g.h_op_lo = L.slotf;
  g.out_op_lo = L.slotf;
  g.len = at<int>(ws, L.src_len);
  // ---- encoder: the two layers as a wavefront -- tick k runs step k of layer 0 and step k-1 of layer 1 in ONE launch
  // (layer 1 at step k-1 needs layer 0's output of step k-1 and its own state of step k-2, both written by earlier ticks)
  auto enc_l0 = [&](int t) {
    GemmArgs a = g;
    a.t = t; a.K = kSH;
    a.a[0] = at<__half>(ws, L.h0op) + t * L.slotop; a.a[1] = nullptr;
    a.w = packed + L.pk_hh0; a.w_lo = static_cast<int64_t>(kSG) * kSH;
    a.table = at<float>(ws, L.P0); a.tok = at<int>(ws, L.src) + t; a.tok_stride = d.Ts;
    a.h_prev = at<float>(ws, L.h0f) + t * L.slotf; a.c_prev = at<float>(ws, L.c0f) + t * L.slotf;
    a.h_out = at<float>(ws, L.h0f) + (t + 1) * L.slotf; a.c_out = at<float>(ws, L.c0f) + (t + 1) * L.slotf;
    a.h_op = at<__half>(ws, L.h0op) + (t + 1) * L.slotop;
    a.gates = need_grad ? at<float>(ws, L.g0) + t * L.slotg : nullptr;
    a.out_f = nullptr; a.out_op = at<__half>(ws, L.out0op) + t * L.slotop;
    return a;
  };
  auto enc_l1 = [&](int t) {
    GemmArgs a = g;
    a.t = t; a.K = 2 * kSH;
    a.a[0] = at<__half>(ws, L.out0op) + t * L.slotop; a.a[1] = at<__half>(ws, L.h1op) + t * L.slotop;
    a.w = packed + L.pk_1; a.w_lo = static_cast<int64_t>(kSG) * 2 * kSH;
    a.table = at<float>(ws, L.P1); a.tok = nullptr; a.tok_stride = 0;
    a.h_prev = at<float>(ws, L.h1f) + t * L.slotf; a.c_prev = at<float>(ws, L.c1f) + t * L.slotf;
    a.h_out = at<float>(ws, L.h1f) + (t + 1) * L.slotf; a.c_out = at<float>(ws, L.c1f) + (t + 1) * L.slotf;
    a.h_op = at<__half>(ws, L.h1op) + (t + 1) * L.slotop;
    a.gates = need_grad ? at<float>(ws, L.g1) + t * L.slotg : nullptr;
    a.out_f = at<float>(ws, L.enc) + static_cast<int64_t>(t) * kSH; a.out_stride = static_cast<int64_t>(d.Ts) * kSH;
    a.out_op = nullptr;
    return a;
  };
  for (int k = 0; k <= d.Ts; ++k) {
    GemmPair pr;
    pr.m_tiles = MT; pr.n_tiles[0] = pr.n_tiles[1] = kSG / 64;
    if (k < d.Ts && k >= 1) { pr.count = 2; pr.g[0] = enc_l0(k); pr.g[1] = enc_l1(k - 1); }
    else { pr.count = 1; pr.g[0] = k < d.Ts ? enc_l0(k) : enc_l1(k - 1); pr.g[1] = pr.g[0]; }
    CUDA_OK(launch_step_gemm_pair(pr, EPI_LSTM, simt, st));
  }
  // ---- decoder: h_0 = encoder output at the last valid position = frozen final layer-1 state, c_0 = 0 -------
  if (fk.join(s_pd)) return fail("pnmn_pg_forward: stream join failed");
  DecRowArgs r;
  std::memset(&r, 0, sizeof(r));
  r.d = d;
  r.h_dec = at<float>(ws, L.h1f) + d.Ts * L.slotf;
  r.enc = at<float>(ws, L.enc); r.src_len = at<int>(ws, L.src_len); r.tgt = teacher ? at<int>(ws, L.tgt) : nullptr;
  r.row_mode = at<int>(ws, L.rowmode);
  r.out_w = params + m->out_w; r.out_b = params + m->out_b;
  r.logits = at<float>(ws, L.logits); r.lse = at<float>(ws, L.lse); r.pred = at<int>(ws, L.pred);
  r.logp = at<float>(ws, L.logp); r.inp = at<int>(ws, L.inp); r.attn_p = at<float>(ws, L.attn_p);
  r.att_op = at<__half>(ws, L.attop); r.att_lo = L.slotf; r.att_step = L.slotop;
  r.seed = at<unsigned long long>(ws, L.seed);
  r.defer_out = (teacher && !row_teacher) ? 1 : 0;   // no free-running rows: outputs of all steps in one launch after the loop
  g.len = nullptr; g.out_f = nullptr; g.out_op = nullptr;
  for (int t = 0; t <= d.S; ++t) {
    r.t = t;
    if (t == d.S && r.defer_out) {
      // every step's logits in ONE tensor-core GEMM, [S * Bp, 256] x [Vt, 256]^T + bias (the per-(row, step) projection
      // kernel re-read the whole output matrix in each of its 10 k CTAs: 180 us for the reconstructor's 41 steps)
      if (pnmn_gemm_split(r.h_dec + L.slotf, kSH, 1, r.out_w, kSH, 1, r.logits, d.Vt, d.S * d.B, d.Vt, kSH, r.out_b, 0,
                          at<float>(ws, L.gemm_ws), L.gemm_ws_floats, st))
        return 1;
      CUDA_OK(launch_dec_out_post(r, st));
      break;
    }
    CUDA_OK(launch_dec_row(r, st));
    if (t == d.S) break;
    g.t = t; g.K = 2 * kSH;
    g.a[0] = at<__half>(ws, L.attop) + t * L.slotop; g.a[1] = at<__half>(ws, L.h1op) + (d.Ts + t) * L.slotop;
    g.w = packed + L.pk_d; g.w_lo = static_cast<int64_t>(kSG) * 2 * kSH;
    g.table = at<float>(ws, L.Pd); g.tok = at<int>(ws, L.inp) + static_cast<int64_t>(t) * d.B; g.tok_stride = 1;
    g.h_prev = at<float>(ws, L.h1f) + (d.Ts + t) * L.slotf; g.c_prev = at<float>(ws, L.cdf) + t * L.slotf;
    g.h_out = at<float>(ws, L.h1f) + (d.Ts + t + 1) * L.slotf; g.c_out = at<float>(ws, L.cdf) + (t + 1) * L.slotf;
    g.h_op = at<__half>(ws, L.h1op) + (d.Ts + t + 1) * L.slotop;
    g.gates = need_grad ? at<float>(ws, L.gd) + t * L.slotg : nullptr;
    CUDA_OK(launch_step_gemm(g, EPI_LSTM, kSG / 64, MT, simt, st));
  }
    return 0;
  });
  if (rc) return rc;
  FinalizeArgs f;
  std::memset(&f, 0, sizeof(f));
  f.d = d; f.rows = rows;
  f.pred = at<int>(ws, L.pred); f.logp = at<float>(ws, L.logp); f.logits = at<float>(ws, L.logits); f.lse = at<float>(ws, L.lse);
  f.tgt = teacher ? at<int>(ws, L.tgt) : nullptr;
  f.row_mode = at<int>(ws, L.rowmode);
  f.raw_out = raw_predictions; f.pred_out = predictions; f.loss = loss; f.logits_out = logits_out;
  f.coef = at<float>(ws, L.coef); f.label = at<int>(ws, L.label);
  CUDA_OK(launch_finalize(f, st));
  pnmn::count_launches(6 + (d.Ts + 1) + 2 * d.S + 1);   // prepare, pack, 3 tables, encoder ticks, decoder row + step kernels, finalize
  return 0;
}

extern "C" int pnmn_pg_forward(const pnmn_pg_desc* m, const float* params, const int64_t* source, const int64_t* target,
                               int batch, int tq, int tp, int steps, int sampling, uint64_t seed, int need_grad,
                               void* ws, int64_t* raw_predictions, int64_t* predictions, float* loss, float* logits_out,
                               void* stream) {
  return pg_forward_impl(m, params, source, target, nullptr, steps, batch, tq, tp, steps, sampling, seed, need_grad, ws,
                         raw_predictions, predictions, loss, logits_out, stream);
}

// One pass over teacher-forced AND free-running rows (the supervised and unsupervised rows of a joint-training batch:
// joint_training_trainer.py:139-144 and :164-168 call the generator once for each kind): rows with row_teacher[b] != 0 are
// teacher-forced on target[b], the others decode freely (categorical sampling) for free_steps steps and take the
// sampled-sequence loss; per row the results are those of the two separate calls.  steps must be tp + 1 >= free_steps.
extern "C" int pnmn_pg_forward_mixed(const pnmn_pg_desc* m, const float* params, const int64_t* source, const int64_t* target,
                                     const uint8_t* row_teacher, int free_steps, int batch, int tq, int tp, uint64_t seed,
                                     int need_grad, void* ws, int64_t* raw_predictions, int64_t* predictions, float* loss,
                                     float* logits_out, void* stream) {
  if (!target || !row_teacher) return fail("pnmn_pg_forward_mixed: target and row_teacher are required");
  if (free_steps < 1 || free_steps > tp + 1) return fail("pnmn_pg_forward_mixed: free_steps must be in [1, tp + 1]");
  return pg_forward_impl(m, params, source, target, row_teacher, free_steps, batch, tq, tp, tp + 1, 1, seed, need_grad, ws,
                         raw_predictions, predictions, loss, logits_out, stream);
}

extern "C" int pnmn_pg_backward(const pnmn_pg_desc* m, const float* params, float* grads, const float* grad_loss, int batch,
                                int tq, int tp, int steps, int teacher, void* ws, void* stream) {
  if (check_dims(m, batch, tq, tp, steps, teacher != 0)) return 1;
  if (!params || !grads || !grad_loss || !ws) return fail("pnmn_pg_backward: NULL buffer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Layout L = make_layout(m, batch, tq, tp, steps, true);
  SeqDims& d = L.d;
  d.teacher = teacher ? 1 : 0;
  const bool simt = use_simt();
  const int MT = d.Bp / 128;
  const __half* packed = at<__half>(ws, L.packed);
  float* scale = at<float>(ws, L.scale);

  const int rows = batch;
  d.B = d.Bp;
  float* gstage = at<float>(ws, L.gstage);
  float* gws = at<float>(ws, L.gws);
  CUDA_OK(launch_stage_grad_loss(grad_loss, rows, d.Bp, gstage, st));
  GraphKey key{ws, params, 0, 1, d.Bp, d.Tq, d.Tp, d.S, 0, d.teacher, 1, d.Vs, d.Vt};
  if (simt) key.pass = 3;
  const int rc = run_graphed(key, st, [&](cudaStream_t st) -> int {
  Forker fk{fork_set(), st};
  CUDA_OK(launch_seq_loss_scale(gstage, d.B, scale, st));
  CUDA_OK(cudaMemsetAsync(gws, 0, 4 * L.extent, st));
  CUDA_OK(cudaMemsetAsync(at<float>(ws, L.dh), 0, 4 * L.slotf, st));
  CUDA_OK(cudaMemsetAsync(at<float>(ws, L.dc), 0, 4 * L.slotf, st));
  CUDA_OK(cudaMemsetAsync(at<float>(ws, L.datt), 0, 4 * L.slotf, st));
  CUDA_OK(cudaMemsetAsync(at<float>(ws, L.dout0), 0, 4 * L.slotf * d.Ts, st));
  CUDA_OK(cudaMemsetAsync(at<float>(ws, L.denc), 0, 4ll * d.B * d.Ts * kSH, st));
  CUDA_OK(cudaMemsetAsync(at<float>(ws, L.dP0), 0, 4ll * d.Vs * kSG, st));
  CUDA_OK(cudaMemsetAsync(at<float>(ws, L.dPd), 0, 4ll * d.Vt * kSG, st));
  CUDA_OK(cudaMemsetAsync(at<float>(ws, L.dP1), 0, 4ll * kSG, st));

  GemmArgs g;
  std::memset(&g, 0, sizeof(g));
  g.B = d.B; g.K = kSG; g.chunks_per_src = kSG / 64; g.a_K[0] = kSG; g.a_lo[0] = L.slotg; g.scale = scale;

  // ---- decoder ------------------------------------------------------------------------------------------------
  DecBwdRowArgs r;
  std::memset(&r, 0, sizeof(r));
  r.d = d;
  r.grad_loss = gstage; r.coef = at<float>(ws, L.coef); r.label = at<int>(ws, L.label);
  r.logits = at<float>(ws, L.logits); r.lse = at<float>(ws, L.lse); r.out_w = params + m->out_w;
  r.h_dec = at<float>(ws, L.h1f) + d.Ts * L.slotf; r.c_dec = at<float>(ws, L.cdf); r.gates = at<float>(ws, L.gd);
  r.enc = at<float>(ws, L.enc); r.src_len = at<int>(ws, L.src_len); r.attn_p = at<float>(ws, L.attn_p);
  r.dlogits = at<float>(ws, L.dlogits); r.dhp = at<float>(ws, L.dhp); r.dh = at<float>(ws, L.dh); r.dc = at<float>(ws, L.dc);
  r.datt = at<float>(ws, L.datt); r.denc = at<float>(ws, L.denc);
  r.dg_op = at<__half>(ws, L.dgd); r.dg_lo = L.slotg; r.dg_step = L.slotdg; r.scale = scale;
  g.len = nullptr;
  g.w = packed + L.pkT_d; g.w_lo = static_cast<int64_t>(2 * kSH) * kSG;
  g.out[0] = at<float>(ws, L.datt); g.out[1] = at<float>(ws, L.dh);
  CUDA_OK(launch_dec_bwd_proj(r, st));
  for (int t = d.S - 1; t >= 0; --t) {
    r.t = t; r.do_attn = t < d.S - 1 ? 1 : 0;
    CUDA_OK(launch_dec_bwd_row(r, st));
    g.t = t; g.a[0] = at<__half>(ws, L.dgd) + t * L.slotdg;
    CUDA_OK(launch_step_gemm(g, EPI_DGRAD, 2 * kSH / 64, MT, simt, st));
  }
  // ---- decoder-side parameter gradients: everything they read is final once the decoder loop is through; they run on a
  // side branch underneath the encoder's time loop (which occupies 32-64 SMs)
  WgradSeqArgs wbase;
  std::memset(&wbase, 0, sizeof(wbase));
  wbase.dg_lo = L.slotg; wbase.dg_step = L.slotdg; wbase.x_lo = L.slotf; wbase.x_step = L.slotop; wbase.m_tiles = MT; wbase.scale = scale;
  TableGradArgs tgbase;
  std::memset(&tgbase, 0, sizeof(tgbase));
  tgbase.dg_lo = L.slotg; tgbase.dg_step = L.slotdg; tgbase.B = d.B; tgbase.scale = scale;
  cudaStream_t s_dec = fk.fork(0);
  {
    WgradSeqArgs w = wbase;
    w.dg = at<__half>(ws, L.dgd); w.T = d.S;
    w.x = at<__half>(ws, L.attop); w.dw = gws + m->dec_w_ih; w.ld = 2 * kSH;
    CUDA_OK(launch_wgrad_seq(w, simt, s_dec));
    w.x = at<__half>(ws, L.h1op) + d.Ts * L.slotop; w.dw = gws + m->dec_w_hh; w.ld = kSH;
    CUDA_OK(launch_wgrad_seq(w, simt, s_dec));
    TableGradArgs tg = tgbase;
    tg.dg = at<__half>(ws, L.dgd); tg.T = d.S; tg.V = d.Vt; tg.tok = at<int>(ws, L.inp); tg.tok_step = d.B; tg.tok_stride = 1;
    tg.dP = at<float>(ws, L.dPd);
    CUDA_OK(launch_table_grad(tg, s_dec));
    CUDA_OK(launch_bias_from_table(at<float>(ws, L.dPd), d.Vt, gws + m->dec_b_ih, gws + m->dec_b_hh, s_dec));
    SimtGemm sg;
    std::memset(&sg, 0, sizeof(sg));
    sg.alpha = 1.f; sg.accumulate = 1;
    // dEmb_tgt[v][e] += sum_g dPd[v][g] * W_dec_ih[g][256 + e]
    sg.A = at<float>(ws, L.dPd); sg.sam = kSG; sg.sak = 1; sg.M = d.Vt; sg.K = kSG;
    sg.B = params + m->dec_w_ih + kSH; sg.sbk = 2 * kSH; sg.sbn = 1; sg.N = kSH; sg.C = gws + m->tgt_embed; sg.ldc = kSH;
    CUDA_OK(launch_simt_gemm(sg, s_dec));
    // dW_dec_ih[g][256 + e] += sum_v dPd[v][g] * Emb_tgt[v][e]
    sg.A = at<float>(ws, L.dPd); sg.sam = 1; sg.sak = kSG; sg.M = kSG; sg.K = d.Vt;
    sg.B = params + m->tgt_embed; sg.sbk = kSH; sg.sbn = 1; sg.N = kSH; sg.C = gws + m->dec_w_ih + kSH; sg.ldc = 2 * kSH;
    CUDA_OK(launch_simt_gemm(sg, s_dec));
    // output projection: dW_o[v][j] += sum_{t,b} dlogits[t][b][v] * h_t[b][j];  db_o[v] += sum dlogits
    sg.A = at<float>(ws, L.dlogits); sg.sam = 1; sg.sak = d.Vt; sg.M = d.Vt; sg.K = d.S * d.Bp;
    sg.B = at<float>(ws, L.h1f) + (d.Ts + 1) * L.slotf; sg.sbk = kSH; sg.sbn = 1; sg.N = kSH; sg.C = gws + m->out_w; sg.ldc = kSH;
    CUDA_OK(launch_simt_gemm(sg, s_dec));
    sg.B = scale + 2; sg.sbk = 0; sg.sbn = 0; sg.N = 1; sg.C = gws + m->out_b; sg.ldc = 1;
    CUDA_OK(launch_simt_gemm(sg, s_dec));
  }
  r.t = -1; r.do_attn = 1;
  CUDA_OK(launch_dec_bwd_row(r, st));   // attention of step 0 -> d(initial decoder state), d(encoder outputs)

  // ---- encoder, both layers as a wavefront: tick j runs layer 1 at t = Ts-1-j and layer 0 at t = Ts-j (which needs
  // d(layer-0 output) of step Ts-j, written by layer 1's data-gradient GEMM in tick j-1); two launches per tick.
  // Layer 1: dh carries d(final state) from the decoder, dc restarts (the decoder's c_0 is a constant); layer 0 has its own
  // carried dh / dc.
  CUDA_OK(cudaMemsetAsync(at<float>(ws, L.dc), 0, 4 * L.slotf, st));
  CUDA_OK(cudaMemsetAsync(at<float>(ws, L.dh0), 0, 4 * L.slotf, st));
  CUDA_OK(cudaMemsetAsync(at<float>(ws, L.dc0), 0, 4 * L.slotf, st));
  EncCellBwdArgs e;
  std::memset(&e, 0, sizeof(e));
  e.B = d.B; e.Bp = d.Bp; e.Ts = d.Ts; e.src_len = at<int>(ws, L.src_len);
  e.dg_lo = L.slotg; e.scale = scale;
  g.len = at<int>(ws, L.src_len);
  auto cell1 = [&](int t) {
    EncCellBwdArgs a = e;
    a.dh = at<float>(ws, L.dh); a.dc = at<float>(ws, L.dc);
    a.t = t; a.gates = at<float>(ws, L.g1) + t * L.slotg;
    a.c_prev = at<float>(ws, L.c1f) + t * L.slotf; a.c_cur = at<float>(ws, L.c1f) + (t + 1) * L.slotf;
    a.dext = at<float>(ws, L.denc) + static_cast<int64_t>(t) * kSH; a.dext_stride = static_cast<int64_t>(d.Ts) * kSH;
    a.dg_op = at<__half>(ws, L.dg1) + t * L.slotdg;
    return a;
  };
  auto cell0 = [&](int t) {
    EncCellBwdArgs a = e;
    a.dh = at<float>(ws, L.dh0); a.dc = at<float>(ws, L.dc0);
    a.t = t; a.gates = at<float>(ws, L.g0) + t * L.slotg;
    a.c_prev = at<float>(ws, L.c0f) + t * L.slotf; a.c_cur = at<float>(ws, L.c0f) + (t + 1) * L.slotf;
    a.dext = at<float>(ws, L.dout0) + t * L.slotf; a.dext_stride = kSH;
    a.dg_op = at<__half>(ws, L.dg0) + t * L.slotdg;
    return a;
  };
  auto dgrad1 = [&](int t) {
    GemmArgs a = g;
    a.w = packed + L.pkT_1; a.w_lo = static_cast<int64_t>(2 * kSH) * kSG;
    a.out[0] = at<float>(ws, L.dout0) + t * L.slotf; a.out[1] = at<float>(ws, L.dh);
    a.keep_masked[0] = 0; a.keep_masked[1] = 1;
    a.t = t; a.a[0] = at<__half>(ws, L.dg1) + t * L.slotdg;
    return a;
  };
  auto dgrad0 = [&](int t) {
    GemmArgs a = g;
    a.w = packed + L.pkT_0; a.w_lo = static_cast<int64_t>(kSH) * kSG;
    a.out[0] = at<float>(ws, L.dh0); a.out[1] = nullptr; a.keep_masked[0] = 1; a.keep_masked[1] = 1;
    a.t = t; a.a[0] = at<__half>(ws, L.dg0) + t * L.slotdg;
    return a;
  };
  for (int j = 0; j <= d.Ts; ++j) {
    const int t1 = d.Ts - 1 - j, t0 = d.Ts - j;
    const bool has1 = j < d.Ts, has0 = j >= 1;
    EncCellBwdPair cp;
    GemmPair gp;
    gp.m_tiles = MT;
    if (has1 && has0) {
      cp.count = gp.count = 2;
      cp.a[0] = cell1(t1); cp.a[1] = cell0(t0);
      gp.g[0] = dgrad1(t1); gp.g[1] = dgrad0(t0); gp.n_tiles[0] = 2 * kSH / 64; gp.n_tiles[1] = kSH / 64;
    } else if (has1) {
      cp.count = gp.count = 1;
      cp.a[0] = cp.a[1] = cell1(t1);
      gp.g[0] = gp.g[1] = dgrad1(t1); gp.n_tiles[0] = gp.n_tiles[1] = 2 * kSH / 64;
    } else {
      cp.count = gp.count = 1;
      cp.a[0] = cp.a[1] = cell0(t0);
      gp.g[0] = gp.g[1] = dgrad0(t0); gp.n_tiles[0] = gp.n_tiles[1] = kSH / 64;
    }
    CUDA_OK(launch_enc_cell_bwd_pair(cp, st));
    CUDA_OK(launch_step_gemm_pair(gp, EPI_DGRAD, simt, st));
  }

  // ---- encoder-side parameter gradients: three weight-gradient GEMMs, two table gradients and their small GEMMs, all
  // independent of each other: main stream + three side branches (the decoder-side ones already run on a fourth branch,
  // forked right behind the decoder loop, underneath the encoder's time loop)
  {
    cudaStream_t s1 = fk.fork(1), s2 = fk.fork(2), s3 = fk.fork(3);
    WgradSeqArgs w = wbase;
    w.dg = at<__half>(ws, L.dg0); w.T = d.Ts;
    w.x = at<__half>(ws, L.h0op); w.dw = gws + m->enc_w_hh[0]; w.ld = kSH;
    CUDA_OK(launch_wgrad_seq(w, simt, st));
    w.dg = at<__half>(ws, L.dg1);
    w.x = at<__half>(ws, L.out0op); w.dw = gws + m->enc_w_ih[1];
    CUDA_OK(launch_wgrad_seq(w, simt, s1));
    w.x = at<__half>(ws, L.h1op); w.dw = gws + m->enc_w_hh[1];
    CUDA_OK(launch_wgrad_seq(w, simt, s2));

    TableGradArgs tg = tgbase;
    tg.dg = at<__half>(ws, L.dg0); tg.T = d.Ts; tg.V = d.Vs; tg.tok = at<int>(ws, L.src); tg.tok_step = 1; tg.tok_stride = d.Ts;
    tg.dP = at<float>(ws, L.dP0);
    CUDA_OK(launch_table_grad(tg, s3));
    CUDA_OK(launch_bias_from_table(at<float>(ws, L.dP0), d.Vs, gws + m->enc_b_ih[0], gws + m->enc_b_hh[0], s3));
    SimtGemm sg;
    std::memset(&sg, 0, sizeof(sg));
    sg.alpha = 1.f; sg.accumulate = 1;
    // dEmb_src[v][e] += sum_g dP0[v][g] * W_ih0[g][e]
    sg.A = at<float>(ws, L.dP0); sg.sam = kSG; sg.sak = 1; sg.M = d.Vs; sg.K = kSG;
    sg.B = params + m->enc_w_ih[0]; sg.sbk = kSH; sg.sbn = 1; sg.N = kSH; sg.C = gws + m->src_embed; sg.ldc = kSH;
    CUDA_OK(launch_simt_gemm(sg, s3));
    // dW_ih0[g][e] += sum_v dP0[v][g] * Emb_src[v][e]
    sg.A = at<float>(ws, L.dP0); sg.sam = 1; sg.sak = kSG; sg.M = kSG; sg.K = d.Vs;
    sg.B = params + m->src_embed; sg.sbk = kSH; sg.sbn = 1; sg.N = kSH; sg.C = gws + m->enc_w_ih[0]; sg.ldc = kSH;
    CUDA_OK(launch_simt_gemm(sg, s3));
    tg.dg = at<__half>(ws, L.dg1); tg.V = 1; tg.tok = nullptr; tg.dP = at<float>(ws, L.dP1);
    CUDA_OK(launch_table_grad(tg, s3));
    CUDA_OK(launch_bias_from_table(at<float>(ws, L.dP1), 1, gws + m->enc_b_ih[1], gws + m->enc_b_hh[1], s3));
    if (fk.join(s1) || fk.join(s2) || fk.join(s3) || fk.join(s_dec)) return fail("pnmn_pg_backward: stream join failed");
  }
    return 0;
  });
  if (rc) return rc;
  CUDA_OK(launch_accumulate(grads, gws, L.extent, st));
  pnmn::count_launches(3 + 2 * d.S + 1 + 2 * (d.Ts + 1) + 5 + 3 + 3 + 6 + 1);   // stage + scale, decoder, encoder ticks, wgrads, tables, biases, small GEMMs, accumulate
  return 0;
}
