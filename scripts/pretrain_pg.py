#!/usr/bin/env python
"""Trains the ProgramGenerator on the synthetic question -> program mapping (probnmn_clevr_b200/synthetic.py) with the
repo's own CUDA forward / backward and FusedClampAdam, and stores the weights as fp16 in
probnmn_clevr_b200/assets/pg_synthetic_fp16.npz (run on the GPU box: `gpurun -- python scripts/pretrain_pg.py`, then copy
gpurun_out/pg_synthetic_fp16.npz into the assets directory).

Why: joint_training_ours.yml starts from question_coding / module_training checkpoints (CHECKPOINTS.*, :27-31); there are no
CLEVR checkpoints in this environment, and a random-init generator samples almost only non-executable programs, for which
the module network does no work -- a benchmark of that would time an empty executor.  Both bench arms (ours and the CPU
reference) load this file, so they sample from the same distribution.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from probnmn_clevr_b200.nmn import NeuralModuleNetwork  # noqa: E402
from probnmn_clevr_b200.optim import FusedClampAdam  # noqa: E402
from probnmn_clevr_b200.seq2seq import ProgramGenerator  # noqa: E402
from probnmn_clevr_b200.synthetic import ProgramSampler, make_seq2seq_state_dict, questions_for_programs  # noqa: E402
from probnmn_clevr_b200.vocabulary import Vocabulary  # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
    torch.manual_seed(0)
    vocab = Vocabulary.clevr()
    vq, vp = vocab.get_vocab_size("questions"), vocab.get_vocab_size("programs")
    pg = ProgramGenerator(vocab)
    pg.load_state_dict(make_seq2seq_state_dict(vq, vp, seed=0))
    pg = pg.cuda().train()
    opt = FusedClampAdam(pg.parameters(), lr=2e-3, clamp=5.0, modules=[pg])
    sampler = ProgramSampler(vocab, seed=12345)
    for it in range(steps):
        if it == steps * 6 // 10:
            opt.param_groups[0]["lr"] = 5e-4
        if it == steps * 9 // 10:
            opt.param_groups[0]["lr"] = 1e-4
        programs = sampler.sample(256, 26)
        questions = questions_for_programs(programs, vq, seed=it)
        opt.zero_grad()
        out = pg(questions.cuda(), programs.cuda())
        out["loss"].mean().backward()
        opt.step()
        if it % 250 == 0 or it == steps - 1:
            print(f"step {it}: teacher-forced loss {float(out['loss'].detach().mean()):.4f}", flush=True)
    # how many SAMPLED programs are executable (the number that matters for the benchmark)
    nmn = NeuralModuleNetwork(vocab).cuda().eval()
    programs = ProgramSampler(vocab, seed=999).sample(512, 26)
    questions = questions_for_programs(programs, vq, seed=999).cuda()
    with torch.no_grad():
        for strategy in ("greedy", "sampling"):
            pred = pg(questions, decoding_strategy=strategy)["predictions"]
            exact = float((pred.cpu()[:, :26] == programs).all(1).float().mean())
            feats = torch.zeros(512, 1024, 14, 14, device="cuda")
            valid = float((nmn(feats, pred)["predictions"] != 28).float().mean())
            print(f"{strategy}: exact programs {exact:.3f}, executable programs {valid:.3f}")
    out_dir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    path = os.path.join(out_dir, "pg_synthetic_fp16.npz")
    np.savez_compressed(path, **{k: v.detach().cpu().numpy().astype(np.float16) for k, v in pg.state_dict().items()})
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
