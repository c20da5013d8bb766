"""world_size-2 gloo tests (CPU) of the data-parallel host logic: batch sharding and the gradient average.  The GPU box
runs the same code over NCCL (bench.py --gpus N)."""
import os
import tempfile

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from probnmn_clevr_b200.dist import GradientOverlap, allreduce_gradients, gradient_buckets, shard_rows


def test_shard_rows_cover_the_batch():
    for n in (1, 7, 256, 2048):
        for world in (1, 2, 3, 8):
            spans = [shard_rows(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1


class _FlatGradModel(torch.nn.Module):
    """Mimics the drop-ins: parameters whose gradients are views into ONE flat buffer, plus a free-standing one."""

    def __init__(self):
        super().__init__()
        self.a = torch.nn.Parameter(torch.zeros(5, 3))
        self.b = torch.nn.Parameter(torch.zeros(7))
        self.c = torch.nn.Parameter(torch.zeros(2, 2))

    def fake_backward(self, value):
        flat = torch.full((64,), float(value))
        self.a.grad = flat[0:15].view(5, 3)
        self.b.grad = flat[16:23]
        self.c.grad = torch.full((2, 2), float(value) * 10)


def _worker(rank, world, path, rows):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    dist.init_process_group("gloo", init_method=f"file://{path}", rank=rank, world_size=world)
    try:
        model = _FlatGradModel()
        model.fake_backward(rank + 1)
        assert len(gradient_buckets([model])) == 2          # one flat bucket + one plain tensor
        n = allreduce_gradients([model])
        assert n == 2
        want = sum(r + 1 for r in range(world)) / world
        assert torch.allclose(model.a.grad, torch.full((5, 3), want))
        assert torch.allclose(model.b.grad, torch.full((7,), want))
        assert torch.allclose(model.c.grad, torch.full((2, 2), want * 10))

        # unequal shards: weighting each rank's mean by its share reproduces the global-batch mean exactly
        torch.manual_seed(0)
        x = torch.randn(rows, 4)
        w = torch.nn.Linear(4, 1)
        torch.manual_seed(1)
        with torch.no_grad():
            w.weight.copy_(torch.randn(1, 4)); w.bias.zero_()
        ref = torch.nn.Linear(4, 1)
        ref.load_state_dict(w.state_dict())
        ref(x).pow(2).mean().backward()
        b, e = shard_rows(rows, rank, world)
        w(x[b:e]).pow(2).mean().backward()
        allreduce_gradients([w], weight=(e - b) / rows * world)
        assert torch.allclose(w.weight.grad, ref.weight.grad, atol=1e-6)
        assert torch.allclose(w.bias.grad, ref.bias.grad, atol=1e-6)

        # overlapped variant: hooked parameters start their all-reduce inside backward, finish() reduces the rest (here:
        # the flat-buffer model) -- same averages as the plain path
        torch.manual_seed(2 + rank)
        net = torch.nn.Sequential(torch.nn.Linear(4, 8), torch.nn.ReLU(), torch.nn.Linear(8, 1))
        torch.manual_seed(5)
        with torch.no_grad():
            for p in net.parameters():
                p.copy_(torch.randn_like(p))
        twin = torch.nn.Sequential(torch.nn.Linear(4, 8), torch.nn.ReLU(), torch.nn.Linear(8, 1))
        twin.load_state_dict(net.state_dict())
        xs = torch.randn(6, 4)
        for min_numel, n_coll in ((1, 4 + 2), (16, 1 + 1 + 2)):
            # min_numel = 16: only the (8, 4) weight starts its own collective inside backward; the other three hooked
            # tensors travel together in one flat buffer at finish()
            overlap = GradientOverlap(net.parameters(), min_numel=min_numel)
            for _ in range(2):  # two steps: the pending lists are cleared by finish()
                net.zero_grad(set_to_none=True); twin.zero_grad(set_to_none=True)
                net(xs).pow(2).mean().backward()
                twin(xs).pow(2).mean().backward()
                model.fake_backward(rank + 1)
                used = overlap.finish([net, model])
                assert used == n_coll, used      # hooked tensors (+ their flat buffer) + the flat bucket + the plain tensor
                allreduce_gradients([twin])
                for p, q in zip(net.parameters(), twin.parameters()):
                    assert torch.allclose(p.grad, q.grad, atol=1e-7)
                assert torch.allclose(model.a.grad, torch.full((5, 3), want))
            overlap.remove()
    finally:
        dist.destroy_process_group()


def test_gradient_allreduce_world_size_2():
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker, args=(2, os.path.join(d, "rdzv"), 7), nprocs=2, join=True)
