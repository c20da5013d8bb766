"""GPU diagnostic (trace build: make -C probnmn_clevr_b200/csrc clean all TRACE=1): phase stamps of the LSTM step kernels of
one ProgramGenerator pass, CTA (0,0,0) of every launch.  Columns (us since the kernel's first instruction): set-up done,
producer past griddepcontrol.wait, all copies issued, first / last ring stage landed, last MMA issued, accumulator complete,
accumulator in registers, epilogue stores issued, CTA done; then the gap to the next launch's start."""
import ctypes, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from probnmn_clevr_b200 import _lib as L
from probnmn_clevr_b200.seq2seq import ProgramGenerator
from probnmn_clevr_b200.synthetic import make_joint_batch
from probnmn_clevr_b200.vocabulary import Vocabulary

vocab = Vocabulary.clevr()
pg = ProgramGenerator(vocab).cuda().train()
bt = make_joint_batch(vocab, 256, seed=0, with_images=False)
q = bt["question"].cuda()
lib = L.lib()
lib.pnmn_debug_pg_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
buf = np.zeros((8192, 16), dtype=np.int64)
for _ in range(4):
    pg.zero_grad(); pg(q, decoding_strategy="sampling")["loss"].mean().backward()
torch.cuda.synchronize()
lib.pnmn_debug_pg_trace(buf.ctypes.data, 8192)
pg.zero_grad(); pg(q, decoding_strategy="sampling")["loss"].mean().backward()
n = lib.pnmn_debug_pg_trace(buf.ctypes.data, 8192)
tr = buf[:n]
tr = tr[np.argsort(tr[:, 0])]
print("launches traced:", n, "graphs:", os.environ.get("PNMN_PG_NOGRAPH") is None)
names = ["setup", "pdl_wait", "issued", "stage0", "stageN", "mma_done_issue", "acc_full", "tmem_ld", "epi_done", "cta_done"]
kinds = {}
for i in range(n):
    key = int(tr[i, 15])   # EPI * 1e6 + K * 1e3 + CTAs of the launch (K = 1024 of the data-gradient GEMMs overflows into EPI)
    rel = (tr[i, 1:11] - tr[i, 0]) / 1e3
    gap = (tr[i + 1, 0] - tr[i, 0]) / 1e3 if i + 1 < n else np.nan
    kinds.setdefault(key, []).append(np.concatenate([rel, [gap]]))
for key, rows in sorted(kinds.items()):
    a = np.array(rows)
    med = np.nanmedian(a, axis=0)
    print(f"EPI {key // 1000000} K {key // 1000 % 1000} CTAs {key % 1000}: {len(rows)} launches")
    print("   " + "  ".join(f"{nm} {v:5.2f}" for nm, v in zip(names + ["next_start"], med)))
