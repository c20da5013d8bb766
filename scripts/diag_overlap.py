"""diagnostics: do two LSTM passes on two streams overlap?  (alone vs concurrent, forward and backward)"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from probnmn_clevr_b200.seq2seq import ProgramGenerator, QuestionReconstructor
from probnmn_clevr_b200.synthetic import make_joint_batch
from probnmn_clevr_b200.vocabulary import Vocabulary

vocab = Vocabulary.clevr()
pg = ProgramGenerator(vocab).cuda().train()
qr = QuestionReconstructor(vocab).cuda().train()
bt = make_joint_batch(vocab, 256, seed=0, with_images=False)
q, p = bt["question"].cuda(), bt["program"].cuda()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
g = torch.full((256,), 1.0 / 256, device="cuda")


def run(which, phase):
    """returns elapsed ms of the phase ('fwd' or 'bwd') with the passes in `which` running concurrently"""
    tot = 0.0
    for it in range(8):
        pg.zero_grad(); qr.zero_grad()
        torch.cuda.synchronize()
        outs = {}
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if phase == "fwd":
            e0.record()
            s1.wait_stream(torch.cuda.current_stream()); s2.wait_stream(torch.cuda.current_stream())
        if "pg" in which:
            with torch.cuda.stream(s1):
                outs["pg"] = pg(q, p, decoding_strategy="sampling")["loss"]
        if "qr" in which:
            with torch.cuda.stream(s2):
                outs["qr"] = qr(p, q, decoding_strategy="sampling")["loss"]
        if phase == "fwd":
            torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
            e1.record()
        torch.cuda.synchronize()
        if phase == "bwd":
            e0.record()
            s1.wait_stream(torch.cuda.current_stream()); s2.wait_stream(torch.cuda.current_stream())
        if "pg" in which:
            with torch.cuda.stream(s1):
                torch.autograd.backward([outs["pg"]], [g])
        if "qr" in which:
            with torch.cuda.stream(s2):
                torch.autograd.backward([outs["qr"]], [g])
        if phase == "bwd":
            torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
            e1.record()
        torch.cuda.synchronize()
        if it >= 3:
            tot += e0.elapsed_time(e1)
    return tot / 5


for phase in ("fwd", "bwd"):
    print(phase, " ".join(f"{w}: {run(w, phase):.2f} ms" for w in (("pg",), ("qr",), ("pg", "qr"))), flush=True)
