"""GPU diagnostic: per-task timeline of the persistent executor for one bench-sized step."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
from probnmn_clevr_b200 import _lib as L
from probnmn_clevr_b200.nmn import NeuralModuleNetwork
from probnmn_clevr_b200.synthetic import ProgramSampler, make_answers, make_features, make_nmn_state_dict
from probnmn_clevr_b200.vocabulary import Vocabulary
B = int(os.environ.get("B", 256))
vocab = Vocabulary.clevr()
m = NeuralModuleNetwork(vocab); m.load_state_dict(make_nmn_state_dict(vocab, 0)); m = m.cuda().train()
feats = make_features(B, 0).cuda(); progs = ProgramSampler(vocab, seed=0).sample(B, 40); ans = make_answers(B, 0).cuda()
def step():
    m.zero_grad(set_to_none=True)
    out = m(feats, progs.cuda(), ans); out["loss"].mean().backward()
for _ in range(3): step()
cap = 60000
trace = torch.zeros(cap * 32, dtype=torch.int64, device="cuda")
L.lib().pnmn_debug_set_trace(trace.data_ptr(), cap)
step(); torch.cuda.synchronize()
L.lib().pnmn_debug_set_trace(None, 0)
st = m.last_plan_stats
nf, nb = st[6], st[14]
tr = trace.cpu().numpy().reshape(-1, 32)
for name, t in (("forward", tr[:nf]), ("backward", tr[nf:nf + nb])):
    t0 = t[:, 0].min(); span = (t[:, 3].max() - t0) / 1e3
    typ = t[:, 5] & 0xff; ns = (t[:, 5] >> 8) & 0xff; nmt = (t[:, 5] >> 16) & 0xff
    wait = (t[:, 1] - t[:, 0]) / 1e3; body = (t[:, 2] - t[:, 1]) / 1e3; pub = (t[:, 3] - t[:, 2]) / 1e3
    print(f"== {name}: {len(t)} tasks, span {span:.1f} us, SMs used {len(set(t[:,4]))}")
    busy = body.sum() / (span * 148)
    print(f"   sum(body) {body.sum():.0f} us -> {100*busy:.1f}% of 148 SMs busy; sum(wait on deps) {wait.sum():.0f} us; sum(publish) {pub.sum():.0f} us")
    conv = typ == 0
    for key in sorted(set(zip(ns[conv], nmt[conv], t[conv, 6]))):
        sel = conv & (ns == key[0]) & (nmt == key[1]) & (t[:, 6] == key[2])
        mma_us = key[0] * key[1] * key[2] * 64 / 1.9e3
        ph = lambda a, b: ((t[sel, a] - t[sel, b]) / 1e3).mean()
        print(f"   conv n_samp={key[0]} n_mt={key[1]} mmas/tile={key[2]:4d}: {sel.sum():5d} tasks | total {ph(3,0):6.1f} us = fetch+deps {ph(1,0):5.1f} + setup {ph(8,1):5.1f} + first operands {ph(9,8):5.1f} + issue {ph(10,9):6.1f} + drain {ph(11,10):5.1f} + epilogue {ph(12,11):5.1f} + fence {ph(13,12):5.1f} + publish {ph(3,13):5.1f} | MMA floor {mma_us:.1f} us | issuer waits: A {t[sel,14].mean()/1.9e3:5.1f} us, W {t[sel,15].mean()/1.9e3:5.1f} us")
    # epilogue of accumulator 0 (clock64 deltas, cycles -> us at 1.9 GHz), by the flag bits that change the epilogue's work
    for fl in sorted(set(t[conv, 7])):
        sel = conv & (t[:, 7] == fl) & (t[:, 6] == 72)
        if sel.sum() < 20: continue
        c = lambda a, b: ((t[sel, a] - t[sel, b]) / 1.9e3).mean()
        print(f"   epilogue flags={int(fl):#05x} n={sel.sum():5d}: first tmem ld {c(16,20):5.2f} | chunk0 body {c(17,16):5.2f} | chunks 1-3 {c(18,17):5.2f} | tail {c(19,18):5.2f} | (acc0 total {c(19,20):5.2f}; all {((t[sel,12]-t[sel,11])/1e3).mean():5.2f})")
    for op in sorted(set(t[~conv, 7])):
        sel = (~conv) & (t[:, 7] == op)
        tot = (t[sel, 3] - t[sel, 1]) / 1e3
        print(f"   elt op {op}: {sel.sum():5d} tasks, ready->published mean {tot.mean():6.1f} us (max {tot.max():.1f})")
    # timeline: active tasks over time
    edges = np.linspace(0, span, 21)
    mid = ((t[:, 1] + t[:, 2]) / 2 - t0) / 1e3
    hist, _ = np.histogram(mid, bins=edges, weights=body)
    print("   busy SMs per 5% of the span:", [int(round(h / (span / 20))) for h in hist])
np.save("gpurun_out/trace.npy", tr[:nf + nb])
