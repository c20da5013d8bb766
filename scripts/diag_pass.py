"""diagnostics: device time of one LSTM pass (forward, backward) as a function of the batch size"""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from probnmn_clevr_b200.seq2seq import ProgramGenerator, QuestionReconstructor
from probnmn_clevr_b200.synthetic import make_joint_batch
from probnmn_clevr_b200.vocabulary import Vocabulary

vocab = Vocabulary.clevr()
pg = ProgramGenerator(vocab).cuda().train()
qr = QuestionReconstructor(vocab).cuda().train()


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    f = b = 0.0
    for _ in range(n):
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record(); loss = fn(backward=False); e1.record(); loss.mean().backward(); e2.record()
        torch.cuda.synchronize()
        f += e0.elapsed_time(e1); b += e1.elapsed_time(e2)
    return f / n, b / n


SIZES = [int(x) for x in sys.argv[1:]] or [64, 123, 128, 200, 256]
for B in SIZES:
    bt = make_joint_batch(vocab, B, seed=0, with_images=False)
    q, p = bt["question"].cuda(), bt["program"].cuda()
    rows = torch.zeros(B, dtype=torch.uint8, device="cuda"); rows[B // 2:] = 1

    def pg_free(backward=True):
        pg.zero_grad(); out = pg(q, decoding_strategy="sampling")["loss"]
        if backward: out.mean().backward()
        return out

    def pg_teacher(backward=True):
        pg.zero_grad(); out = pg(q, p, decoding_strategy="sampling")["loss"]
        if backward: out.mean().backward()
        return out

    def pg_mixed(backward=True):
        pg.zero_grad(); out = pg.forward_mixed(q, p, rows)["loss"]
        if backward: out.mean().backward()
        return out

    def qr_teacher(backward=True):
        qr.zero_grad(); out = qr(p, q, decoding_strategy="sampling")["loss"]
        if backward: out.mean().backward()
        return out

    for name, fn in (("pg free", pg_free), ("pg teacher", pg_teacher), ("pg mixed", pg_mixed), ("qr teacher", qr_teacher)):
        f, b = timeit(fn)
        print(f"B={B:3d} {name:10s}: forward {f:.3f} ms, backward {b:.3f} ms", flush=True)
