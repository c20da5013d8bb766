"""GPU diagnostic: one ProgramGenerator (or QuestionReconstructor with QR=1) pass, batch 256, for an ncu launch list:
ncu --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file out.csv python scripts/pg_pass_ncu.py"""
import os, sys
import torch
import torch.cuda.profiler as cp
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from probnmn_clevr_b200.seq2seq import ProgramGenerator, QuestionReconstructor
from probnmn_clevr_b200.synthetic import make_joint_batch
from probnmn_clevr_b200.vocabulary import Vocabulary

vocab = Vocabulary.clevr()
bt = make_joint_batch(vocab, 256, seed=0, with_images=False)
q, p = bt["question"].cuda(), bt["program"].cuda()
if os.environ.get("QR"):
    m = QuestionReconstructor(vocab).cuda().train()
    run = lambda: m(p, q, decoding_strategy="sampling")["loss"].mean().backward()
else:
    m = ProgramGenerator(vocab).cuda().train()
    run = lambda: m(q, decoding_strategy="sampling")["loss"].mean().backward()
for _ in range(3):
    m.zero_grad(); run()
torch.cuda.synchronize()
cp.start(); m.zero_grad(); run(); torch.cuda.synchronize(); cp.stop()
