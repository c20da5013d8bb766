"""GPU, >= 2 devices (run with `gpurun --gpus 2 -- python -m pytest tests/test_dist_gpu.py -m gpu`; skipped on a one-GPU box):
data parallelism must reproduce the single-process global batch (the reference has no working multi-GPU
path to compare with, so N-rank gradients are validated against the single-process gradients of the same global batch)."""
import os
import tempfile

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _worker(rank, world, path, out_path):
    from probnmn_clevr_b200.dist import shard_rows
    from probnmn_clevr_b200.nmn import NeuralModuleNetwork
    from probnmn_clevr_b200.seq2seq import ProgramGenerator
    from probnmn_clevr_b200.synthetic import make_joint_batch, make_nmn_state_dict, make_seq2seq_state_dict
    from probnmn_clevr_b200.vocabulary import Vocabulary
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", init_method=f"file://{path}", rank=rank, world_size=world, device_id=dev)
    try:
        vocab = Vocabulary.clevr()
        nmn = NeuralModuleNetwork(vocab)
        nmn.load_state_dict(make_nmn_state_dict(vocab, 0))
        pg = ProgramGenerator(vocab)
        pg.load_state_dict(make_seq2seq_state_dict(93, 44, seed=0))
        nmn, pg = nmn.to(dev).train(), pg.to(dev).train()
        batch = make_joint_batch(vocab, 24, seed=5)
        B = 24
        lo, hi = shard_rows(B, rank, world)

        def grads(rows, overlap):
            nmn.zero_grad(); pg.zero_grad()
            sl = slice(*rows)
            out = nmn(batch["image"][sl].to(dev), batch["program"][sl].to(dev), batch["answer"][sl].to(dev))
            gen = pg(batch["question"][sl].to(dev), batch["program"][sl].to(dev))
            # each rank's share of the GLOBAL mean: sum over its rows / B, times world (the all-reduce averages over ranks)
            scale = world if rows != (0, B) else 1
            ((out["loss"].sum() + gen["loss"].sum()) / B * scale).backward()
            if rows != (0, B):
                nmn.allreduce_gradients()
                from probnmn_clevr_b200.dist import allreduce_gradients
                allreduce_gradients([pg])
            return torch.cat([p.grad.flatten() for m in (nmn, pg) for p in m.parameters()]).clone()

        g_dp = grads((lo, hi), False)
        nmn.enable_gradient_overlap()
        g_dp_overlap = grads((lo, hi), True)
        nmn._grad_overlap.remove(); nmn._grad_overlap = None
        g_full = grads((0, B), False)
        torch.cuda.synchronize()
        scale = float(g_full.abs().max())
        e1 = float((g_dp - g_full).abs().max()) / scale
        e2 = float((g_dp_overlap - g_full).abs().max()) / scale
        if rank == 0:
            with open(out_path, "w") as f:
                f.write(f"{e1} {e2}")
        # every rank must hold the same averaged gradient
        mine = g_dp.clone()
        dist.broadcast(mine, src=0)
        assert torch.equal(mine, g_dp)
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_rank_gradients_equal_single_process_global_batch():
    with tempfile.TemporaryDirectory() as d:
        out = os.path.join(d, "err.txt")
        mp.spawn(_worker, args=(2, os.path.join(d, "rendezvous"), out), nprocs=2, join=True)
        e1, e2 = (float(x) for x in open(out).read().split())
    print(f"2-rank vs single-process gradient: max-norm rel err {e1:.2e} (plain), {e2:.2e} (overlapped all-reduce)")
    # identical arithmetic per row; only the summation order of the atomically accumulated weight gradients differs
    assert e1 < 1e-4 and e2 < 1e-4
