"""GPU: the joint-training step through the C ABI against the CPU oracles and the
golden vectors of the reference's elbo.py.

Tolerances: per-row losses rtol/atol 1e-3 (north_star); ELBO scalars 1e-3; seq2seq gradients relative L2 <= 5e-3 per
tensor (split-fp16 GEMMs are fp32-class; the REINFORCE coefficient carries the NMN loss's 1e-3); NMN gradients global
relative L2 <= 6e-2 against the fp32 oracle (tf32-class operands, see test_nmn_gpu.py); optimizer: |p - p_torch| <= 2.5e-7
(two ulps of a parameter of magnitude 1..2) after four steps with lr = 1e-3, i.e. 1e-4 of the distance moved.
"""
import os

import numpy as np
import pytest
import torch

from oracle import joint_oracle, prior_oracle
from probnmn_clevr_b200.elbo import JointTrainingElbo, QuestionCodingElbo
from probnmn_clevr_b200.joint import JointTrainingStep, split_batch
from probnmn_clevr_b200.nmn import NeuralModuleNetwork
from probnmn_clevr_b200.optim import FusedClampAdam
from probnmn_clevr_b200.program_prior import ProgramPrior
from probnmn_clevr_b200.seq2seq import ProgramGenerator, QuestionReconstructor
from probnmn_clevr_b200.synthetic import (ProgramSampler, make_joint_batch, make_nmn_state_dict, make_prior_state_dict,
                                          make_seq2seq_state_dict)
from probnmn_clevr_b200.vocabulary import Vocabulary

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "elbo_golden.npz")


def _rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / (b.norm() + 1e-30))


# ---------------------------------------------------------------------------------------------------------------------
# ELBO / REINFORCE glue kernel against the reference's elbo.py (golden)
# ---------------------------------------------------------------------------------------------------------------------
class _Stub(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.queue = []

    def forward(self, *a, **k):
        return self.queue.pop(0)


@pytest.mark.parametrize("mode", ["joint_ours", "joint_baseline", "question_coding"])
def test_elbo_glue_matches_reference_golden(mode):
    g = np.load(GOLDEN)
    beta, gamma, decay = (float(x) for x in g["hyper"])
    for n in (1, 7, 128, 131):
        pg, qr, prior, nmn = _Stub(), _Stub(), _Stub(), _Stub()
        if mode == "question_coding":
            elbo = QuestionCodingElbo(pg, qr, prior, beta=beta, baseline_decay=decay)
        else:
            elbo = JointTrainingElbo(pg, qr, prior, nmn, beta=beta, gamma=gamma, baseline_decay=decay,
                                     objective="ours" if mode == "joint_ours" else "baseline", concurrent=(n == 128))
        for k in range(3):
            tag = f"{mode}.n{n}.call{k}"
            leaves = {key: torch.from_numpy(g[f"{tag}.in.{key}"]).cuda().requires_grad_(True) for key in ("pg", "qr", "prior", "nmn")}
            dummy = torch.zeros(n, 3, dtype=torch.long, device="cuda")
            for model, key in ((pg, "pg"), (qr, "qr"), (prior, "prior"), (nmn, "nmn")):
                model.queue.append({"predictions": dummy, "loss": leaves[key]})
            if mode == "question_coding":
                out = elbo(dummy)
                objective = -out["elbo"]
            else:
                out = elbo(dummy, torch.zeros(n, 1, device="cuda"), torch.zeros(n, dtype=torch.long, device="cuda"))
                objective = gamma * out["nmn_loss"] - out["elbo"]
            objective.backward()
            for key in [k2[len(tag) + 5:] for k2 in g.files if k2.startswith(tag + ".out.")]:
                np.testing.assert_allclose(out[key].item(), g[f"{tag}.out.{key}"], rtol=2e-5, atol=2e-6, err_msg=f"{tag} {key}")
            for key, leaf in leaves.items():
                mine = leaf.grad.cpu().numpy() if leaf.grad is not None else np.zeros(n, np.float32)
                np.testing.assert_allclose(mine, g[f"{tag}.grad.{key}"], rtol=2e-5, atol=1e-8, err_msg=f"{tag} grad {key}")
            np.testing.assert_allclose(elbo._reinforce._reinforce_baseline, g[f"{tag}.baseline_after"], rtol=2e-5, atol=1e-6)
            for model in (pg, qr, prior, nmn):
                model.queue.clear()


# ---------------------------------------------------------------------------------------------------------------------
# ProgramPrior
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("batch", [5, 128, 200])
def test_program_prior_forward_matches_oracle(batch):
    vocab = Vocabulary.clevr()
    sd = make_prior_state_dict(44, seed=1)
    prior = ProgramPrior(vocab)
    missing = prior.load_state_dict(sd, strict=True)
    prior = prior.cuda().eval()
    prior.return_logits = True
    sampler = ProgramSampler(vocab, seed=batch)
    # (+ a program that fills all 26 positions and an empty one; padding is always trailing, as in the CLEVR token files)
    full = torch.from_numpy(np.random.default_rng(batch).integers(4, 44, size=(1, 26), dtype=np.int64))
    programs = torch.cat([sampler.sample(batch - 2, 26), full, torch.zeros(1, 26, dtype=torch.int64)])
    with torch.no_grad():
        ref = prior_oracle.prior_forward(sd, programs, torch.Generator().manual_seed(0))
        for _ in range(2):   # twice: the second call reuses the workspace
            out = prior(programs.cuda())
    assert out["loss"].grad_fn is None
    np.testing.assert_allclose(out["loss"].cpu().numpy(), ref["loss"].numpy(), rtol=1e-4, atol=1e-5)
    mask = torch.arange(27)[None] <= (programs != 0).sum(1, keepdim=True)   # positions whose next token exists
    e = float(((out["logits"].cpu() - ref["logits"]).abs() * mask[:, :, None]).max() / ref["logits"].abs().max())
    print(f"prior B={batch}: logits max-norm rel err {e:.2e}")
    assert e < 1e-4
    pred = out["predictions"].cpu()
    assert pred.shape == (batch, 27) and ((pred == 0) == ~mask).all() and (pred[pred != 0] > 2).all()
    assert set(prior.state_dict()) == set(sd) and prior.get_metrics()["perplexity"] > 1.0


# ---------------------------------------------------------------------------------------------------------------------
# fused clamp + Adam against torch.optim.Adam
# ---------------------------------------------------------------------------------------------------------------------
def test_fused_clamp_adam_matches_torch_adam():
    vocab = Vocabulary.clevr()
    nmn = NeuralModuleNetwork(vocab)
    nmn.load_state_dict(make_nmn_state_dict(vocab, 0))
    pg = ProgramGenerator(vocab)
    nmn, pg = nmn.cuda().train(), pg.cuda().train()
    batch = make_joint_batch(vocab, 12, seed=4)
    out = nmn(batch["image"].cuda(), batch["program"].cuda(), batch["answer"].cuda())
    (1e3 * out["loss"].mean()).backward()                      # (x 1000: some gradients exceed the clamp)
    gen = pg(batch["question"].cuda(), batch["program"].cuda())
    (30 * gen["loss"].mean()).backward()
    params = list(pg.parameters()) + list(nmn.parameters())
    assert all(p.grad is not None for p in params)
    grads = [p.grad.detach().clone() for p in params]
    assert max(float(g.abs().max()) for g in grads) > 5.0
    ref_params = [torch.nn.Parameter(p.detach().clone()) for p in params]
    ours = FusedClampAdam(params, lr=1e-3, clamp=5.0, modules=[pg, nmn])
    ref = torch.optim.Adam(ref_params, lr=1e-3)
    for it in range(3):
        for p, q, g in zip(params, ref_params, grads):
            q.grad = g.clone().clamp_(min=-5, max=5)
            assert torch.equal(p.grad, g)                        # our step leaves .grad alone unless asked
        ours.step()
        ref.step()
    assert ours.launches_last_step <= 10, ours.launches_last_step   # pg: 1 range; nmn: stem, modules, 6 classifier tensors
    worst = max(float((p.detach() - q.detach()).abs().max()) for p, q in zip(params, ref_params))
    print(f"fused clamp+Adam vs torch.optim.Adam after 3 steps: max |dp| = {worst:.2e}, launches {ours.launches_last_step}")
    assert worst <= 2.5e-7
    # state interchange: torch.optim.Adam continues from our state dict, we continue from its
    ref2 = torch.optim.Adam(ref_params, lr=1e-3)
    import copy
    ref2.load_state_dict(copy.deepcopy(ours.state_dict()))   # (a checkpoint round trip copies; load_state_dict alone aliases)
    sd_ref = ref.state_dict()
    for k in (0, len(params) - 1):
        assert float(sd_ref["state"][k]["step"]) == 3.0 == float(ours.state_dict()["state"][k]["step"])
        assert float((sd_ref["state"][k]["exp_avg"] - ours.state_dict()["state"][k]["exp_avg"]).abs().max()) < 1e-7
    # a missing gradient: that tensor is skipped (torch semantics), the rest of its range takes the per-tensor path
    skip = 5
    params[skip].grad = None
    ref_params[skip].grad = None
    before = params[skip].detach().clone()
    for p, q, g in zip(params, ref_params, grads):
        if p.grad is not None:
            q.grad = g.clone().clamp_(min=-5, max=5)
    ours.step()
    ref2.step()
    assert torch.equal(params[skip].detach(), before)
    worst = max(float((p.detach() - q.detach()).abs().max()) for p, q in zip(params, ref_params))
    assert worst <= 2.5e-7, worst
    assert float(ours.state_dict()["state"][skip]["step"]) == 3.0 and float(ours.state_dict()["state"][0]["step"]) == 4.0


# ---------------------------------------------------------------------------------------------------------------------
# the whole iteration
# ---------------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def trained():
    """models of the joint step; the program generator is first trained (teacher forcing, our kernels + FusedClampAdam) on
    the synthetic question -> program mapping until its samples are mostly executable programs"""
    torch.manual_seed(0)
    vocab = Vocabulary.clevr()
    vq, vp = vocab.get_vocab_size("questions"), vocab.get_vocab_size("programs")
    pg = ProgramGenerator(vocab)
    pg.load_state_dict(make_seq2seq_state_dict(vq, vp, seed=0))
    qr = QuestionReconstructor(vocab)
    qr.load_state_dict(make_seq2seq_state_dict(vp, vq, seed=1))
    nmn = NeuralModuleNetwork(vocab)
    nmn.load_state_dict(make_nmn_state_dict(vocab, 0))
    prior = ProgramPrior(vocab)
    prior.load_state_dict(make_prior_state_dict(vp, seed=0))
    pg, qr, nmn, prior = pg.cuda().train(), qr.cuda().train(), nmn.cuda().train(), prior.cuda().eval()
    opt = FusedClampAdam(pg.parameters(), lr=2e-3, clamp=5.0, modules=[pg])
    losses = []
    for it in range(800):
        b = make_joint_batch(vocab, 256, seed=1000 + it % 40, with_images=False)
        opt.zero_grad()
        out = pg(b["question"].cuda(), b["program"].cuda())
        out["loss"].mean().backward()
        opt.step()
        if it % 100 == 0 or it == 799:
            losses.append(float(out["loss"].mean()))
    print("program generator pre-training loss:", " ".join(f"{x:.3f}" for x in losses))
    assert losses[-1] < 0.5 * losses[0], "teacher-forced training with the CUDA backward + fused Adam does not learn"
    return vocab, pg, qr, nmn, prior, losses


def _state(module):
    return {k: v.detach().cpu().clone().requires_grad_(True) for k, v in module.state_dict().items()}


@pytest.mark.parametrize("fused", [True, False], ids=["fused_passes", "one_pass_per_call"])
def test_joint_iteration_matches_oracle(trained, fused):
    vocab, pg, qr, nmn, prior, _ = trained
    step = JointTrainingStep(pg, qr, nmn, prior, alpha=100.0, beta=0.1, gamma=1.0, delta=0.99, lr=1e-6, concurrent=True,
                             fused=fused)
    pg.return_logits = True
    # a larger batch first: the compared iteration then runs in workspaces whose padding rows hold stale data
    step.optimizer.zero_grad()
    step.do_iteration({k: v for k, v in make_joint_batch(vocab, 72, seed=7).items()})
    batch = make_joint_batch(vocab, 28, seed=8)
    sds = {"program_generator": _state(pg), "question_reconstructor": _state(qr), "nmn": _state(nmn), "program_prior": _state(prior)}
    sds["program_prior"]["_output_layer.weight"] = sds["program_prior"]["_embedder.token_embedder_programs.weight"]
    baseline0 = step.elbo._reinforce._reinforce_baseline
    step.optimizer.zero_grad()
    out = step.do_iteration(batch)
    torch.cuda.synchronize()
    last = step.elbo.last_outputs
    raw = last["program_generator"]["raw_predictions"].cpu()
    pg.return_logits = False
    state = joint_oracle.ElboState(baseline0)
    torch.set_num_threads(os.cpu_count())
    ref = joint_oracle.joint_iteration(sds, vocab, batch, state, alpha=100.0, beta=0.1, gamma=1.0, delta=0.99, forced_programs=raw)
    assert torch.equal(last["program_generator"]["predictions"].cpu(), ref["sampled_programs"])
    n_valid = int(ref["rows"]["nmn_valid"].sum())
    print(f"unsupervised rows {raw.shape[0]}, executable sampled programs {n_valid}")
    assert n_valid >= 3, "the pre-trained generator should sample some executable programs"
    rows = {"pg_loss": last["program_generator"]["loss"], "qr_loss": last["question_reconstructor"]["loss"],
            "prior_loss": last["program_prior"]["loss"], "nmn_loss": last["nmn"]["loss"]}
    for k, v in rows.items():
        np.testing.assert_allclose(v.detach().cpu().numpy(), ref["rows"][k].numpy(), rtol=1e-3, atol=1e-3, err_msg=k)
    assert torch.equal(last["nmn"]["predictions"].cpu() == 28, ref["rows"]["nmn_valid"] == 0)
    for k, v in ref["elbo"].items():
        np.testing.assert_allclose(out["elbo"][k].item(), v.item(), rtol=1e-3, atol=1e-3, err_msg=k)
    for k, v in ref["loss"].items():
        np.testing.assert_allclose(out["loss"][k].item(), v.item(), rtol=1e-3, atol=1e-3, err_msg=k)
    np.testing.assert_allclose(out["objective"].item(), ref["objective"].item(), rtol=1e-3, atol=1e-2)
    np.testing.assert_allclose(step.elbo._reinforce._reinforce_baseline, state.baseline, rtol=1e-3, atol=1e-3)
    # gradients
    for name, model in (("program_generator", pg), ("question_reconstructor", qr)):
        worst = 0.0
        for k, p in model.named_parameters():
            r = sds[name][k].grad
            worst = max(worst, _rel(p.grad.cpu(), r))
        print(f"{name}: worst per-tensor gradient rel L2 err {worst:.2e}")
        assert worst < 5e-3
    num = den = 0.0
    for k, p in nmn.named_parameters():
        r = sds["nmn"][k].grad
        r = torch.zeros_like(sds["nmn"][k]) if r is None else r
        mine = torch.zeros_like(r) if p.grad is None else p.grad.cpu()
        num += float((mine.double() - r.double()).pow(2).sum()); den += float(r.double().pow(2).sum())
    print(f"nmn: global gradient rel L2 err {(num / den) ** 0.5:.2e}")
    assert (num / den) ** 0.5 < 6e-2
    # the optimizer step moves every parameter that has a gradient by at most lr (Adam's first steps) and nothing else
    before = [p.detach().clone() for p in step.optimizer.param_groups[0]["params"]]
    step.optimizer.step()
    delta = max(float((p.detach() - b).abs().max()) for p, b in zip(step.optimizer.param_groups[0]["params"], before))
    assert 0 < delta <= 1.1e-6


def test_concurrent_streams_and_handover_change_nothing(trained):
    """side streams (concurrent=True) and the pinned-host hand-over of the sampled programs are scheduling only"""
    vocab, pg, qr, nmn, prior, _ = trained
    batch = make_joint_batch(vocab, 40, seed=21)
    results = []
    for fused in (False, True):
        _compare_variants(vocab, pg, qr, nmn, prior, batch, fused)


def _compare_variants(vocab, pg, qr, nmn, prior, batch, fused):
    results = []
    for concurrent, handover in ((True, True), (False, True), (False, False)):
        step = JointTrainingStep(pg, qr, nmn, prior, lr=1e-6, concurrent=concurrent, fused=fused)
        pg.handover_predictions = handover
        pg._calls = qr._calls = prior._calls = pg._teacher_calls = qr._teacher_calls = 100   # same Philox keys for every variant
        step.optimizer.zero_grad()
        out = step.do_iteration(batch)
        torch.cuda.synchronize()
        flat = torch.cat([p.grad.flatten() for m in (pg, qr, nmn) for p in m.parameters() if p.grad is not None]).clone()
        results.append((out, step.elbo.last_outputs["program_generator"]["predictions"].clone(), flat))
    pg.handover_predictions = True
    (o0, p0, g0) = results[0]
    for o, p, g in results[1:]:
        assert torch.equal(p, p0)
        for k in o0["elbo"]:
            assert torch.equal(o["elbo"][k], o0["elbo"][k]), k
        assert torch.equal(o["objective"], o0["objective"])
        assert float((g - g0).abs().max()) <= 2e-5 * float(g0.abs().max())   # atomics: summation order varies


def test_full_step_runs_and_learns(trained):
    """a few complete steps (zero_grad, iteration, fused clamp + Adam) at a practical learning rate: the supervised
    question-reconstruction loss falls"""
    vocab, pg, qr, nmn, prior, _ = trained
    step = JointTrainingStep(pg, qr, nmn, prior, lr=3e-4)
    first = last = None
    for it in range(60):
        out = step.step(split_batch(make_joint_batch(vocab, 64, seed=300 + it % 5)))
        v = float(out["loss"]["question_reconstruction_gt"])
        first = v if first is None else first
        last = v
        assert np.isfinite(float(out["objective"]))
    print(f"question reconstruction (supervised rows): {first:.3f} -> {last:.3f} in 60 steps")
    assert last < 0.85 * first


def test_deferred_module_network_backward_changes_nothing(trained):
    """JointTrainingStep(defer_nmn=True): the module network's backward pass + update of step i is issued with step i + 1
    (flush() issues the last one).  Same objective every step and the same models afterwards as with every step complete in
    itself: the sampling streams are replayed (same call counters), so both runs see the same programs."""
    vocab, pg, qr, nmn, prior, _ = trained
    models = (pg, qr, nmn)
    start = [{k: v.detach().clone() for k, v in m.state_dict().items()} for m in models]
    counters = [(m._calls, m._teacher_calls) for m in (pg, qr)]
    batches = [make_joint_batch(vocab, 40, seed=30 + i) for i in range(3)]

    def run(defer):
        for m, sd in zip(models, start):
            m.load_state_dict(sd)
        for m, (c, t) in zip((pg, qr), counters):
            m._calls, m._teacher_calls = c, t
        step = JointTrainingStep(pg, qr, nmn, prior, alpha=100.0, beta=0.1, gamma=1.0, delta=0.99, lr=1e-4, defer_nmn=defer)
        step.defer_qr = defer       # (the reconstructor's backward pass + update likewise, PNMN_JOINT_DEFER_QR)
        objectives = [float(step.step(b)["objective"]) for b in batches]
        assert (step._pending_nmn is not None) == defer and (step._pending_qr is not None) == defer
        step.flush()
        assert step._pending_nmn is None and step._pending_qr is None
        torch.cuda.synchronize()
        nmn_steps = {int(step.optimizer.state_dict()["state"][i]["step"]) for i in step.optimizer.state_dict()["state"]}
        return objectives, [{k: v.detach().clone() for k, v in m.state_dict().items()} for m in models], nmn_steps

    try:
        obj_a, sd_a, steps_a = run(False)
        obj_b, sd_b, steps_b = run(True)
    finally:
        for m, sd in zip(models, start):
            m.load_state_dict(sd)
    assert steps_a == steps_b == {3}                       # every parameter of every model was stepped three times
    np.testing.assert_allclose(obj_a, obj_b, rtol=2e-4)
    for name, a, b, s0 in zip(("program_generator", "question_reconstructor", "nmn"), sd_a, sd_b, start):
        # (L2 over all parameters: Adam's first steps move every element by ~lr * sign(gradient), and the weight-gradient
        # kernels' atomics make the sign of a near-zero gradient entry a coin toss from run to run)
        moved = sum(float((a[k] - s0[k]).double().pow(2).sum()) for k in a) ** 0.5
        diff = sum(float((a[k] - b[k]).double().pow(2).sum()) for k in a) ** 0.5
        print(f"{name}: parameters moved by {moved:.3e} (L2), deferred vs complete steps differ by {diff:.3e}")
        assert moved > 0 and diff <= 0.05 * moved


def test_device_prefetcher_slot_reuse_and_ordering():
    """feed.DevicePrefetcher: batches come back in order with the right contents while copies of later batches are in
    flight; a slot is only overwritten after its tenant was released; re-allocation on a shape change"""
    from probnmn_clevr_b200.feed import DevicePrefetcher
    dev = torch.device("cuda", 0)
    feed = DevicePrefetcher(dev, depth=3)
    host = [(torch.full((1 << 20,), float(i)).pin_memory(), torch.arange(i, i + 8).pin_memory()) for i in range(8)]
    sums = []
    feed.submit(0, host[0])
    for i in range(8):
        if i + 1 < 8:
            feed.submit(i + 1, host[i + 1])
        x, y = feed.get(i)
        # slow consumer: the copy of batch i+2 must not land in this slot before the work queued here has run
        acc = x.clone()
        for _ in range(20):
            acc = acc * 1.0 + 0.0
        sums.append((acc.sum(), y.clone()))
    with pytest.raises(RuntimeError):
        for k in range(100, 104):
            feed.submit(k, host[0])
    torch.cuda.synchronize()
    for i, (s, y) in enumerate(sums):
        assert float(s) == float(i) * (1 << 20) and torch.equal(y.cpu(), host[i][1])
    with pytest.raises(ValueError):
        feed.submit("pageable", (torch.zeros(4),))
    feed2 = DevicePrefetcher(dev, depth=2)
    feed2.submit("a", (torch.ones(16).pin_memory(),))
    (a,) = feed2.get("a")
    feed2.submit("b", ((torch.ones(32) * 2).pin_memory(),))   # other shape: slot re-allocated
    (b,) = feed2.get("b")
    assert float(a.sum()) == 16 and float(b.sum()) == 64
