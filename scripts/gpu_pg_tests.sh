#!/bin/bash
# seq2seq: parity tests (CUDA-core twins, then tcgen05) and the time of one ProgramGenerator step
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
PNMN_PG_SIMT=1 timeout 600 python -m pytest tests/test_seq2seq_gpu.py -q 2>&1 | tail -3
timeout 600 python -m pytest tests/test_seq2seq_gpu.py -q -s 2>&1 | grep -E "passed|failed|FAILED|rel err [0-9.e-]+$" | tail -8 | cut -c1-200
timeout 300 python - <<'PY'
import os, sys, time, torch
sys.path.insert(0, os.getcwd())
from probnmn_clevr_b200.seq2seq import ProgramGenerator
from probnmn_clevr_b200.synthetic import ProgramSampler, make_questions, make_seq2seq_state_dict
from probnmn_clevr_b200.vocabulary import Vocabulary
vocab = Vocabulary.clevr()
pg = ProgramGenerator(vocab); pg.load_state_dict(make_seq2seq_state_dict(93, 44, seed=0)); pg = pg.cuda().train()
B = 256
q = make_questions(B, 93, seed=0, max_length=40).cuda(); p = ProgramSampler(vocab, seed=0).sample(B, 26).cuda()
def mix():
    pg.zero_grad(set_to_none=True)
    a = pg(q[:128], p[:128], decoding_strategy="sampling"); b = pg(q[128:], decoding_strategy="sampling")
    (a["loss"].mean() + b["loss"].mean()).backward()
def samp():
    pg.zero_grad(set_to_none=True)
    pg(q, decoding_strategy="sampling")["loss"].mean().backward()
def fwd_only():
    with torch.no_grad(): pg(q, decoding_strategy="greedy")
for name, fn in (("mix 128 teacher + 128 sampled fwd+bwd", mix), ("256 sampled fwd+bwd", samp), ("256 greedy fwd only", fwd_only)):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10): fn()
    torch.cuda.synchronize(); print(f"{name}: {(time.perf_counter()-t0)*100:.2f} ms")
PY
