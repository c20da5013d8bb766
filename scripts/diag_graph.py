"""diagnostics: are the seq2seq passes replayed as CUDA graphs, and what does a pass cost on host and device"""
import ctypes, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from probnmn_clevr_b200 import _lib as L
from probnmn_clevr_b200.seq2seq import ProgramGenerator
from probnmn_clevr_b200.synthetic import make_joint_batch
from probnmn_clevr_b200.vocabulary import Vocabulary

vocab = Vocabulary.clevr()
pg = ProgramGenerator(vocab).cuda().train()
b = make_joint_batch(vocab, 123, seed=0, with_images=False)
q, p = b["question"].cuda(), b["program"].cuda()
stats = (ctypes.c_int64 * 4)()
for mode in ("sampling", "teacher"):
    for it in range(6):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        pg.zero_grad()
        out = pg(q, p if mode == "teacher" else None, decoding_strategy="sampling")
        t1 = time.perf_counter()
        out["loss"].mean().backward()
        t2 = time.perf_counter()
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        L.lib().pnmn_debug_graph_stats(stats)
        print(f"{mode} it {it}: host fwd {1e3*(t1-t0):.2f} ms, host bwd {1e3*(t2-t1):.2f} ms, total {1e3*(t3-t0):.2f} ms, graph stats {list(stats)}", flush=True)
