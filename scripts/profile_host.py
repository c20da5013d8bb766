"""GPU diagnostic: where the HOST spends its time while issuing a bench step (the step is host-bound once the device work
drops below ~5.5 ms)."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
from probnmn_clevr_b200.nmn import NeuralModuleNetwork
from probnmn_clevr_b200.synthetic import ProgramSampler, make_answers, make_features, make_nmn_state_dict
from probnmn_clevr_b200.vocabulary import Vocabulary

B = 256
vocab = Vocabulary.clevr()
m = NeuralModuleNetwork(vocab); m.load_state_dict(make_nmn_state_dict(vocab, 0)); m = m.cuda().train()
feats = make_features(B, 0).cuda(); progs = ProgramSampler(vocab, seed=0).sample(B, 40).pin_memory(); ans = make_answers(B, 0).cuda()

LOOKAHEAD = os.environ.get("LOOKAHEAD") == "1"  # plans compiled two steps ahead on the helper threads (as bench.py does)

def step():
    m.zero_grad(set_to_none=True)
    out = m(feats, progs, ans)
    if LOOKAHEAD:
        m.precompile(progs)
    out["loss"].mean().backward()

for _ in range(5): step()
torch.cuda.synchronize()
# coarse phases
T = {"zero_grad": 0.0, "forward": 0.0, "loss.mean": 0.0, "backward": 0.0}
N = 20
for _ in range(N):
    torch.cuda.synchronize()
    t0 = time.perf_counter(); m.zero_grad(set_to_none=True)
    t1 = time.perf_counter(); out = m(feats, progs, ans)
    if LOOKAHEAD: m.precompile(progs)
    t2 = time.perf_counter(); l = out["loss"].mean()
    t3 = time.perf_counter(); l.backward()
    t4 = time.perf_counter()
    T["zero_grad"] += t1 - t0; T["forward"] += t2 - t1; T["loss.mean"] += t3 - t2; T["backward"] += t4 - t3
print("cpus:", os.cpu_count(), len(os.sched_getaffinity(0)), open("/sys/fs/cgroup/cpu.max").read().strip() if os.path.exists("/sys/fs/cgroup/cpu.max") else "-")
print("host ms per step (device idle at the start of every step):", {k: round(v / N * 1e3, 3) for k, v in T.items()})
import cProfile, pstats
pr = cProfile.Profile()
torch.cuda.synchronize()
pr.enable()
for _ in range(10): step()
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr); st.sort_stats("tottime").print_stats(28)
