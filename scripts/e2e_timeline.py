"""GPU diagnostic: device-side timeline of the end-to-end leg (bench.py e2e_step): where does a step wait?"""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
from probnmn_clevr_b200.feed import DevicePrefetcher
from probnmn_clevr_b200.nmn import NeuralModuleNetwork
from probnmn_clevr_b200.synthetic import ProgramSampler, make_answers, make_features, make_nmn_state_dict
from probnmn_clevr_b200.vocabulary import Vocabulary

B, N = 256, int(os.environ.get("STEPS", 24))
dev = torch.device("cuda:0")
vocab = Vocabulary.clevr()
m = NeuralModuleNetwork(vocab); m.load_state_dict(make_nmn_state_dict(vocab, 0)); m = m.to(dev).train()
host = [(make_features(B, s).pin_memory(), ProgramSampler(vocab, seed=s).sample(B, 40).pin_memory(), make_answers(B, s).pin_memory()) for s in range(2)]
feed = DevicePrefetcher(dev)
loss_host = torch.zeros(N, dtype=torch.float32).pin_memory()
loss_ev = [torch.cuda.Event() for _ in range(N)]
E = lambda: [torch.cuda.Event(enable_timing=True) for _ in range(N)]
e_start, e_feat, e_fwd, e_bwd = E(), E(), E(), E()
h_issue = []
MODE = os.environ.get("MODE", "e2e")


def run(n):
    for i in range(n):
        t0 = time.perf_counter()
        e_start[i].record()
        if MODE == "e2e":
            if feed.pending() == 0:
                feed.submit(i, (host[i % 2][0], host[i % 2][2]))
            f, a = feed.get(i)
        else:
            f, a = res[i % 2]
        e_feat[i].record()
        m.zero_grad(set_to_none=True)
        out = m(f, host[i % 2][1], a)
        if i + 2 < n:
            m.precompile(host[(i + 2) % 2][1])
        if MODE == "e2e" and i + 1 < n:
            feed.submit(i + 1, (host[(i + 1) % 2][0], host[(i + 1) % 2][2]))
        e_fwd[i].record()
        loss = out["loss"].mean()
        loss.backward()
        e_bwd[i].record()
        loss_host[i:i + 1].copy_(loss.detach().reshape(1), non_blocking=True)
        loss_ev[i].record()
        t1 = time.perf_counter()
        if i > 0:
            loss_ev[i - 1].synchronize()
        h_issue.append((t1 - t0) * 1e3)
    loss_ev[n - 1].synchronize()


res = [(h[0].to(dev), h[2].to(dev)) for h in host]
run(8)
torch.cuda.synchronize()
h_issue.clear()
t0 = time.perf_counter()
run(N)
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) * 1e3
print(f"mode {MODE}: wall {wall / N:.2f} ms/step; host issue {sum(h_issue) / N:.2f} ms/step")
print(" step | start->start | wait for features | forward | backward | idle before start (prev bwd end -> start)")
for i in range(2, N):
    print(f"  {i:3d} | {e_start[i - 1].elapsed_time(e_start[i]):6.2f} | {e_start[i].elapsed_time(e_feat[i]):6.2f} | {e_feat[i].elapsed_time(e_fwd[i]):6.2f} | "
          f"{e_fwd[i].elapsed_time(e_bwd[i]):6.2f} | {e_bwd[i - 1].elapsed_time(e_start[i]):6.2f}   host issue {h_issue[i]:.2f}")
