"""GPU: the CUDA NeuralModuleNetwork (through the C ABI) against the reference's golden vectors and the
CPU oracle, forward and backward.

Tolerances (north_star: "within 1e-3 fp32 relative tolerance (answer logits, attention maps)"):
  logits / final module outputs : max|a-b| <= 1e-3 * max|b|   (tensor cores run tf32: 10-bit mantissa)
  loss                          : rtol 1e-3, atol 1e-3
  parameter gradients           : max|a-b| <= 5e-3 * max|b| per tensor (fp16/tf32 products in dgrad/wgrad)
"""
import os

import numpy as np
import pytest
import torch

from oracle import nmn_oracle
from probnmn_clevr_b200.nmn import NeuralModuleNetwork
from probnmn_clevr_b200.synthetic import ProgramSampler, make_answers, make_features, make_nmn_state_dict
from probnmn_clevr_b200.vocabulary import Vocabulary

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "nmn_golden.npz")
BIG_SUBSAMPLE = 997


@pytest.fixture(scope="module")
def model():
    vocab = Vocabulary.clevr()
    m = NeuralModuleNetwork(vocab)
    m.load_state_dict(make_nmn_state_dict(vocab, 0))
    return m.cuda()


def _relmax(a, b):
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


def _run(model, programs, answers, feats, train=True):
    box = {}
    h = model.classifier.register_forward_hook(lambda m, i, o: box.update(final=i[0].detach(), logits=o.detach()))
    model.train(train)
    model.zero_grad()
    out = model(feats.cuda(), programs.cuda(), None if answers is None else answers.cuda())
    h.remove()
    return out, box


@pytest.mark.parametrize("name", ["semantic", "sampled", "garbage"])
def test_forward_matches_reference_golden(model, name):
    g = np.load(GOLDEN)
    programs = torch.from_numpy(g[f"{name}.programs"])
    answers = torch.from_numpy(g[f"{name}.answers"])
    feats = make_features(programs.shape[0], 0)
    out, box = _run(model, programs, answers, feats)
    pred = out["predictions"].cpu().numpy()
    valid = g[f"{name}.valid"]
    assert ((pred == 28) == (valid == 0)).all()
    e_logits = _relmax(box["logits"].cpu().numpy(), g[f"{name}.logits"])
    e_final = _relmax(box["final"].sum(dim=(1, 2, 3)).cpu().numpy(), g[f"{name}.final_sum"])
    e_f0 = _relmax(box["final"][min(2, len(valid) - 1)].cpu().numpy(), g[f"{name}.final_sample0"])
    print(f"[{name}] logits rel err {e_logits:.2e}  final-sum rel err {e_final:.2e}  final[sample] rel err {e_f0:.2e}")
    assert e_logits < 1e-3 and e_f0 < 1e-3
    np.testing.assert_allclose(out["loss"].detach().cpu().numpy(), g[f"{name}.loss"], rtol=1e-3, atol=1e-3)
    # bit-exact argmax wherever the reference's top-2 margin exceeds the tolerance
    lg = g[f"{name}.logits"]
    top2 = np.sort(lg, axis=1)[:, -2:]
    safe = (top2[:, 1] - top2[:, 0] > 2e-3 * np.abs(lg).max()) & (valid == 1)
    assert (pred[safe] == g[f"{name}.predictions"][safe]).all()
    out_na, _ = _run(model, programs, None, feats, train=False)
    np.testing.assert_allclose(out_na["loss"].detach().cpu().numpy(), g[f"{name}.loss_noanswer"], rtol=1e-3, atol=1e-3)
    assert "metrics" not in out_na and "metrics" in out


def _grad_errors(model, ref_grads, verbose=True):
    """per-tensor relative L2 error (tensors carrying a non-negligible share of the gradient) and the global one"""
    rows, num, den = [], 0.0, 0.0
    for k, p in model.named_parameters():
        ref = ref_grads[k]
        mine = torch.zeros_like(ref) if p.grad is None else p.grad.detach().cpu()
        e2 = float((mine.double() - ref.double()).pow(2).sum())
        r2 = float(ref.double().pow(2).sum())
        num, den = num + e2, den + r2
        if r2 == 0:
            assert e2 == 0, k
        rows.append((e2 ** 0.5, r2 ** 0.5, k))
    biggest = max(r for _, r, _ in rows)
    rows.sort(reverse=True)
    if verbose:
        for e, r, k in rows[:6]:
            print(f"    |err| {e:.3e}  |ref| {r:.3e}  rel {e / (r + 1e-30):.2e}  {k}")
    worst = max(((e / r, k) for e, r, k in rows if r > 1e-3 * biggest), default=(0.0, ""))
    return worst, (num / den) ** 0.5


def _smooth_state_dict(sd):
    """A weight set on which the network is SMOOTH: every ReLU input is positive (small conv weights, bias +1)
    and the sigmoid heads are unsaturated.  On the He-initialised weights the gradient is a discontinuous
    function of the forward pass: a one-ulp tf32 difference after a store (unavoidable between any two tf32
    implementations: accumulation order decides the rounding direction) flips ~2e-4 of the ReLU masks per
    layer, i.e. ~1.5% (L2) of that layer's gradient; the fp32 and tf32-rounded ORACLES differ by 3e-2 for
    that reason.  With the kinks out of the way the backward arithmetic itself can be checked tightly."""
    out = {}
    for k, v in sd.items():
        if k.startswith("classifier."):
            out[k] = v.clone()
        elif k.endswith(("conv3.weight", "conv6.weight", ".conv.weight")) and v.shape[0] == 1:
            out[k] = v * 0.03
        elif k.endswith(("conv3.bias", "conv6.bias", ".conv.bias")) and v.shape[0] == 1:
            out[k] = torch.zeros_like(v)
        elif k.endswith(".weight"):
            out[k] = v * (0.2 if k == "stem.0.weight" else 0.1)
        else:
            out[k] = torch.ones_like(v)
    return out


@pytest.mark.parametrize("name", ["semantic", "sampled"])
def test_backward_matches_oracle_on_smooth_network(name):
    """Parameter gradients of the CUDA path against autograd through the fp32 oracle (see _smooth_state_dict).
    Bounds: relative L2 error 1e-2 for every tensor holding > 0.1% of the largest tensor-gradient norm,
    3e-3 for the whole gradient (tf32 forward/dgrad operands, fp16 wgrad operands, fp32 accumulation)."""
    g = np.load(GOLDEN)
    vocab = Vocabulary.clevr()
    sd0 = _smooth_state_dict(make_nmn_state_dict(vocab, 0))
    model = NeuralModuleNetwork(vocab)
    model.load_state_dict(sd0)
    model = model.cuda()
    programs = torch.from_numpy(g[f"{name}.programs"])
    answers = torch.from_numpy(g[f"{name}.answers"])
    if name == "semantic":
        # rows 6, 12, 23 take min/max of two 128-channel tensors that are both ~1 on this weight set: another
        # kink (which operand wins flips with a 1-ulp difference); they stay in the He-init tests above/below
        keep = [i for i in range(programs.shape[0]) if i not in (6, 12, 23)]
        programs, answers = programs[keep], answers[keep]
    feats = make_features(programs.shape[0], 0)
    out, box = _run(model, programs, answers, feats)
    out["loss"].mean().backward()
    sd = {k: v.clone().requires_grad_(True) for k, v in sd0.items()}
    # the classifier (its ReLUs and max-pool are kinks too) is evaluated at the CUDA path's module outputs
    ref = nmn_oracle.nmn_forward(sd, vocab, feats, programs, answers, final_override=box["final"].cpu())
    ref["loss"].mean().backward()
    ref_grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in sd.items()}
    e_fwd = _relmax(box["final"].cpu().numpy(), nmn_oracle.nmn_forward(sd0, vocab, feats, programs)["final"].numpy())
    worst, l2 = _grad_errors(model, ref_grads)
    print(f"[{name}] smooth network: module outputs rel err {e_fwd:.2e}; worst per-tensor grad L2 rel err {worst[0]:.2e} at "
          f"{worst[1]}; global L2 rel err {l2:.2e}")
    assert e_fwd < 1e-3
    assert worst[0] < 1e-2 and l2 < 3e-3


@pytest.mark.parametrize("name", ["semantic", "sampled"])
def test_backward_close_to_fp32_reference_golden(model, name):
    """Against the reference's own fp32 gradients (golden file): global agreement only (per-tensor deviations
    of saturated-sigmoid heads under tf32 are large for ANY tf32 implementation; measured with the oracle)."""
    g = np.load(GOLDEN)
    programs = torch.from_numpy(g[f"{name}.programs"])
    answers = torch.from_numpy(g[f"{name}.answers"])
    feats = make_features(programs.shape[0], 0)
    out, _ = _run(model, programs, answers, feats)
    out["loss"].mean().backward()
    num = den = 0.0
    for k, p in model.named_parameters():
        grad = np.zeros(p.shape, np.float32) if p.grad is None else p.grad.detach().cpu().numpy()
        if f"{name}.grad.{k}" in g:
            ref, mine = g[f"{name}.grad.{k}"], grad
        else:
            ref, mine = g[f"{name}.gradsub.{k}"], grad.reshape(-1)[::BIG_SUBSAMPLE]
        if np.abs(ref).max() == 0:
            assert np.abs(mine).max() == 0, k
        num += float(((mine.astype(np.float64) - ref) ** 2).sum())
        den += float((ref.astype(np.float64) ** 2).sum())
    l2 = (num / den) ** 0.5
    print(f"[{name}] vs fp32 reference golden: global L2 rel err of the (sub-sampled) gradient {l2:.2e}")
    assert l2 < 6e-2  # the tf32-rounded oracle itself sits at 3e-2 from the fp32 reference on these batches


def test_attention_maps_and_bigger_batch_against_oracle(model):
    """a fresh seeded batch (not in the golden file): oracle on CPU vs CUDA path, incl. garbage rows"""
    vocab = model.vocabulary
    sampler = ProgramSampler(vocab, seed=11)
    programs = torch.cat([sampler.sample(24, 26), sampler.garbage(8, 26)])
    B = programs.shape[0]
    feats, answers = make_features(B, 3), make_answers(B, 3)
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    with torch.no_grad():
        ref = nmn_oracle.nmn_forward(sd, vocab, feats, programs, answers)
    out, box = _run(model, programs, answers, feats)
    assert (out["predictions"].cpu() == 28).eq(ref["valid"] == 0).all()
    e = _relmax(box["logits"].cpu().numpy(), ref["logits"].numpy())
    ef = _relmax(box["final"].cpu().numpy(), ref["final"].numpy())
    print(f"oracle batch: logits rel err {e:.2e}, final rel err {ef:.2e}")
    assert e < 1e-3 and ef < 1e-3
    np.testing.assert_allclose(out["loss"].detach().cpu().numpy(), ref["loss"].numpy(), rtol=1e-3, atol=1e-3)


def test_precompile_lookahead_changes_nothing(model):
    """NeuralModuleNetwork.precompile (plan compiled ahead of time on a helper thread) is purely a host-side pipeline:
    the step that picks the plan up returns the same bits as an inline compile, a plan compiled for other programs is
    discarded, and the training-mode metrics are the lazily evaluated dict."""
    vocab = model.vocabulary
    sampler = ProgramSampler(vocab, seed=23)
    programs = torch.cat([sampler.sample(12, 26), sampler.garbage(4, 26)])
    other = sampler.sample(16, 26)
    feats, answers = make_features(16, 5).cuda(), make_answers(16, 5).cuda()
    model.train()

    def step(pre):
        model.zero_grad()
        if pre is not None:
            model.precompile(pre)
        out = model(feats, programs, answers)
        out["loss"].mean().backward()
        g = torch.cat([p.grad.flatten() for p in model.parameters() if p.grad is not None])
        return out, g.clone()

    ref, gref = step(None)
    hit, ghit = step(programs)           # plan picked up
    miss, gmiss = step(other)            # stale plan discarded, inline compile
    for out, g in ((hit, ghit), (miss, gmiss)):
        assert torch.equal(out["predictions"], ref["predictions"]) and torch.equal(out["loss"], ref["loss"])
        assert out["metrics"] == ref["metrics"] and isinstance(out["metrics"], dict)
    # gradients: the weight-gradient kernels accumulate with atomics (order varies run to run), so equality is to rounding
    for g in (ghit, gmiss):
        assert float((g - gref).abs().max()) <= 1e-5 * float(gref.abs().max())


def test_split_compile_changes_nothing(model):
    """Training-mode forward without a precompiled plan: the forward pass launches from the forward-half plan
    (PNMN_PLAN_FORWARD_HALF) and the backward pass runs from the full plan compiled meanwhile on a helper thread.  Same
    bits forward, same gradients to rounding as one plan compiled inline; with a prestaged input as well; and a forward
    pass whose backward never runs still collects its full plan."""
    vocab = model.vocabulary
    sampler = ProgramSampler(vocab, seed=29)
    programs = torch.cat([sampler.sample(44, 26), sampler.garbage(4, 26)])
    feats, answers = make_features(48, 6).cuda(), make_answers(48, 6).cuda()
    model.train()

    def step(split, prestage=False):
        model.split_compile = split
        model.zero_grad()
        if prestage:
            model.prestage(feats)
        out = model(feats, programs, answers)
        stats_fwd = list(model.last_plan_stats)
        out["loss"].mean().backward()
        g = torch.cat([p.grad.flatten() for p in model.parameters() if p.grad is not None])
        return out, g.clone(), stats_fwd, list(model.last_plan_stats)

    try:
        ref, gref, _, stats_ref = step(False)
        for prestage in (False, True):
            out, g, stats_fwd, stats_full = step(True, prestage)
            assert torch.equal(out["predictions"], ref["predictions"]) and torch.equal(out["loss"], ref["loss"])
            assert float((g - gref).abs().max()) <= 1e-5 * float(gref.abs().max())
            assert stats_fwd[10] == 0 and stats_full[8] == stats_ref[8] and stats_full[10] == stats_ref[10] > 0
        model.split_compile = True
        out = model(feats, programs, answers)     # never backpropagated
        del out
        out, g, _, _ = step(True)
        assert torch.equal(out["loss"], ref["loss"])
    finally:
        model.split_compile = True


# ---------------------------------------------------------------------------------------------------------------------
# attention maps (north_star: "within 1e-3 ... (answer logits, attention maps)")
# ---------------------------------------------------------------------------------------------------------------------
MAP_TOL = 1e-3   # max |a - b| over a map; maps are sigmoid outputs / min / max of such, i.e. values in [0, 1]


def _captured_maps(model, programs, answers, feats, train):
    model.capture_attention_maps = True
    try:
        out, box = _run(model, programs, answers, feats, train=train)
        torch.cuda.synchronize()
        return out, {(n, j): m for n, j, _tok, m in model.last_attention_maps}
    finally:
        model.capture_attention_maps = False
        model.last_attention_maps = None


@pytest.mark.parametrize("train", [True, False], ids=["train_plan", "eval_plan"])
@pytest.mark.parametrize("name", ["semantic", "sampled"])
def test_attention_maps_match_reference_golden(model, name, train):
    """Every 1-channel module output (AttentionModule / RelateModule / SameModule outputs, 1-channel And / Or results:
    nmn_modules.py:25-27,43-45,82-87,160-168,200-208) of the CUDA executor against the maps the VERBATIM reference produced
    (forward hooks on its modules, oracle/make_golden.py), element-wise: max |a - b| <= 1e-3 on values in [0, 1]."""
    g = np.load(GOLDEN)
    programs = torch.from_numpy(g[f"{name}.programs"])
    answers = torch.from_numpy(g[f"{name}.answers"])
    feats = make_features(programs.shape[0], 0)
    if train:
        _, mine = _captured_maps(model, programs, answers, feats, True)
    else:
        with torch.no_grad():
            _, mine = _captured_maps(model, programs, None, feats, False)
    gold = g[f"{name}.attention_maps"]
    assert len(gold) > 0 and len(mine) == len(gold), (len(mine), len(gold))
    worst = 0.0
    for row in gold:
        key = (int(row[0]), int(row[1]))
        assert key in mine, f"the executor produced no map for sample {key[0]}, module call {key[1]}"
        worst = max(worst, float(np.abs(mine[key].numpy().reshape(-1) - row[2:]).max()))
    print(f"[{name}/{'train' if train else 'eval'}] {len(gold)} attention maps, max |cuda - reference| = {worst:.2e}")
    assert worst <= MAP_TOL


# ---------------------------------------------------------------------------------------------------------------------
# parity at the benchmarked configuration (bench.py: batch 256, programs <= 40 tokens, seeds 0 / 1 / 100)
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("seed,length,rows", [(0, 40, 256), (1, 40, 256), (100, 40, 256), (0, 26, 128)],
                         ids=["bench_seed0", "bench_seed1", "bench_rank1", "joint_128x26"])
def test_bench_configuration_against_oracle(model, seed, length, rows):
    """The batches bench.py times (pairing of samples, critical-sample splitting, strands and the dependency-slot
    fallback only engage at this size): logits, module outputs, loss, validity and attention maps against the CPU oracle,
    for the training plan (need_grad) and the evaluation plan (torch.no_grad)."""
    vocab = model.vocabulary
    programs = ProgramSampler(vocab, seed=seed).sample(256, length)[:rows]
    feats, answers = make_features(rows, seed), make_answers(rows, seed)
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    torch.set_num_threads(os.cpu_count())
    with torch.no_grad():
        ref = nmn_oracle.nmn_forward(sd, vocab, feats, programs, answers, want=("traces",))
    ref_maps = {(n, j): o[0, 0] for n, tr in enumerate(ref["traces"]) if int(ref["valid"][n]) == 1
                for j, (_t, o) in enumerate(tr) if o.shape[1] == 1}
    for train in (True, False):
        if train:
            out, mine = _captured_maps(model, programs, answers, feats, True)
            box = None
        else:
            with torch.no_grad():
                out, mine = _captured_maps(model, programs, answers, feats, False)
        # _captured_maps ran _run, whose hook results are gone: run the classifier hook again on the same inputs
        with torch.set_grad_enabled(train):
            out2, box = _run(model, programs, answers, feats, train=train)
        assert torch.equal(out["predictions"], out2["predictions"])
        assert (out["predictions"].cpu() == 28).eq(ref["valid"] == 0).all()
        e = _relmax(box["logits"].cpu().numpy(), ref["logits"].numpy())
        ef = _relmax(box["final"].cpu().numpy(), ref["final"].numpy())
        np.testing.assert_allclose(out["loss"].detach().cpu().numpy(), ref["loss"].numpy(), rtol=1e-3, atol=1e-3)
        assert set(mine) == set(ref_maps)
        em = max(float((mine[k] - ref_maps[k]).abs().max()) for k in ref_maps)
        print(f"[seed {seed} L {length} B {rows} {'train' if train else 'eval'}] logits {e:.2e}  final {ef:.2e}  "
              f"maps {em:.2e} ({len(ref_maps)} maps)  valid {int(ref['valid'].sum())}")
        assert e < 1e-3 and ef < 1e-3 and em <= MAP_TOL
        if train:
            out2["loss"].mean().backward()   # releases the plan / workspace like a training step


def test_precompile_under_no_grad_reads_its_own_validity_mask(model):
    """Evaluation pipeline with look-ahead compiles: the validity mask of batch i must come from batch i's task
    tables even when the helper thread is already uploading batch i+1's tables into a recycled buffer."""
    vocab = model.vocabulary
    sampler = ProgramSampler(vocab, seed=31)
    batches = [torch.cat([sampler.sample(10, 26), sampler.garbage(6, 26)])[torch.randperm(16, generator=torch.Generator().manual_seed(k))]
               for k in range(6)]
    feats, answers = make_features(16, 9).cuda(), make_answers(16, 9).cuda()
    model.eval()
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    expect = [nmn_oracle.nmn_forward(sd, vocab, feats.cpu(), p, answers.cpu())["valid"] == 0 for p in batches]
    with torch.no_grad():
        for rounds in range(3):
            model.precompile(batches[0], need_grad=False)
            for i, p in enumerate(batches):
                if i + 1 < len(batches):
                    model.precompile(batches[i + 1], need_grad=False)
                out = model(feats, p, answers)
                assert torch.equal(out["predictions"].cpu() == 28, expect[i]), (rounds, i)
                assert torch.equal(out["loss"].cpu() == 3.33, expect[i]), (rounds, i)
    model.train()


def test_backward_tighter_than_tf32_rounding_noise():
    """He-initialised network (ReLU kinks and saturated heads included): the CUDA gradient must sit CLOSER to the
    tf32-operand oracle than that oracle sits to the fp32 one -- i.e. whatever differs from the reference's fp32 gradient
    is the operand precision north_star asks for (tensor cores), not an error of the backward pass.  Errors are relative
    L2 norms, per tensor (tensors holding > 1 % of the largest tensor-gradient norm) and over the whole gradient."""
    g = np.load(GOLDEN)
    vocab = Vocabulary.clevr()
    sd0 = make_nmn_state_dict(vocab, 0)
    model = NeuralModuleNetwork(vocab)
    model.load_state_dict(sd0)
    model = model.cuda()
    programs = torch.from_numpy(g["sampled.programs"])
    answers = torch.from_numpy(g["sampled.answers"])
    feats = make_features(programs.shape[0], 0)
    out, box = _run(model, programs, answers, feats)
    out["loss"].mean().backward()
    torch.set_num_threads(os.cpu_count())

    def oracle_grads(mode):
        sd = {k: v.clone().requires_grad_(True) for k, v in sd0.items()}
        with nmn_oracle.operand_rounding(mode):
            ref = nmn_oracle.nmn_forward(sd, vocab, feats, programs, answers)
            ref["loss"].mean().backward()
        return {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in sd.items()}

    g_tf32, g_fp32 = oracle_grads("tf32"), oracle_grads("fp32")
    worst_c, l2_c = _grad_errors(model, g_tf32, verbose=False)

    class _Holder:   # _grad_errors reads named_parameters() with .grad
        def __init__(self, grads):
            self.items = [(k, type("P", (), {"grad": v})()) for k, v in grads.items()]

        def named_parameters(self):
            return self.items
    worst_o, l2_o = _grad_errors(_Holder(g_tf32), g_fp32, verbose=False)
    print(f"He-init: cuda vs tf32 oracle: global {l2_c:.2e}, worst tensor {worst_c[0]:.2e} ({worst_c[1]}); "
          f"tf32 oracle vs fp32 oracle: global {l2_o:.2e}, worst tensor {worst_o[0]:.2e} ({worst_o[1]})")
    assert l2_c < l2_o, "the CUDA gradient is further from the tf32 oracle than tf32 rounding itself moves the gradient"
    assert l2_c < 3e-2


def test_feature_cache_gives_identical_results():
    """feed.ImageFeatureCache (device-resident fp16 features keyed by image index, readers.py:63-108) + the fp16 entry point:
    logits, predictions, losses and gradients are IDENTICAL to the ones for the fp32 features of the same images"""
    from probnmn_clevr_b200.feed import ImageFeatureCache
    from probnmn_clevr_b200.synthetic import ProgramSampler, make_answers, make_features, make_nmn_state_dict
    images = make_features(12, 11)                                   # 12 "images"
    cache = ImageFeatureCache(images.double().numpy(), "cuda:0", split="train", chunk_images=5)
    assert len(cache) == 12 and cache.split == "train" and cache.features.dtype == torch.float16
    index = torch.tensor([3, 3, 0, 11, 7, 3, 5, 0, 9, 1])           # questions share images
    vocab = Vocabulary.clevr()
    model = NeuralModuleNetwork(vocab)
    model.load_state_dict(make_nmn_state_dict(vocab, 0))
    model = model.cuda().train()
    programs = ProgramSampler(vocab, seed=5).sample(10, 26).cuda()
    answers = make_answers(10, 5).cuda()
    outs, grads, logits = [], [], []
    hook = model.classifier.register_forward_hook(lambda m, i, o: logits.append(o.detach().clone()))
    for feats in (images[index].cuda(), cache.gather(index)):
        model.zero_grad()
        out = model(feats, programs, answers)
        out["loss"].mean().backward()
        outs.append(out)
        grads.append(torch.cat([p.grad.flatten() for p in model.stem.parameters()]).clone())
    hook.remove()
    assert torch.equal(logits[0], logits[1])
    assert torch.equal(outs[0]["predictions"], outs[1]["predictions"]) and torch.equal(outs[0]["loss"], outs[1]["loss"])
    # (weight gradients are summed with fp32 atomics: equal up to the summation order)
    assert float((grads[0] - grads[1]).abs().max()) <= 1e-5 * float(grads[0].abs().max())
    item = cache[3]
    assert item.dtype.name == "float32" and item.shape == (1024, 14, 14)
    assert float(torch.as_tensor(item).sub(images[3]).abs().max()) <= 2e-3 * float(images[3].abs().max())


def test_prestage_changes_nothing():
    """NeuralModuleNetwork.prestage (weights packed and features laid out before the programs are known, stem inputs indexed
    by row) followed by forward gives the same logits / predictions / losses / gradients as forward alone, invalid programs
    included; a prestage for other features is ignored"""
    vocab = Vocabulary.clevr()
    model = NeuralModuleNetwork(vocab)
    model.load_state_dict(make_nmn_state_dict(vocab, 0))
    model = model.cuda().train()
    sampler = ProgramSampler(vocab, seed=9)
    programs = torch.cat([sampler.sample(5, 26), sampler.garbage(2, 26), sampler.sample(6, 26)]).cuda()
    feats, answers = make_features(13, 9).cuda(), make_answers(13, 9).cuda()
    other = make_features(13, 10).cuda()
    results = []
    for mode in ("plain", "prestaged", "stale"):
        logits = []
        hook = model.classifier.register_forward_hook(lambda m, i, o: logits.append(o.detach().clone()))
        model.zero_grad()
        if mode == "prestaged":
            model.prestage(feats)
        elif mode == "stale":
            model.prestage(other)           # not the features of the forward that follows: must be dropped
        out = model(feats, programs, answers)
        out["loss"].mean().backward()
        hook.remove()
        g = torch.cat([p.grad.flatten() for p in model.stem.parameters()]).clone()
        results.append((logits[0], out["predictions"].clone(), out["loss"].detach().clone(), g))
    for r in results[1:]:
        assert torch.equal(r[0], results[0][0]) and torch.equal(r[1], results[0][1]) and torch.equal(r[2], results[0][2])
        assert float((r[3] - results[0][3]).abs().max()) <= 1e-5 * float(results[0][3].abs().max())
    assert int((results[0][1] == 28).sum()) == 2


def test_compile_chunks_change_nothing():
    """a batch compiled as 4 independent plans on 4 host threads and executed on 4 streams (NeuralModuleNetwork.compile_chunks)
    gives the same logits / predictions / losses as the single plan, and the same gradients up to summation order; invalid
    programs land in different chunks"""
    vocab = Vocabulary.clevr()
    sampler = ProgramSampler(vocab, seed=21)
    B = 70
    programs = sampler.sample(B, 26)
    programs[[3, 40, 69]] = sampler.garbage(3, 26)
    programs = programs.cuda()
    feats, answers = make_features(B, 21).cuda(), make_answers(B, 21).cuda()
    results = []
    for chunks in (1, 4):
        model = NeuralModuleNetwork(vocab)
        model.load_state_dict(make_nmn_state_dict(vocab, 0))
        model = model.cuda().train()
        model.compile_chunks = chunks
        logits = []
        hook = model.classifier.register_forward_hook(lambda m, i, o: logits.append(o.detach().clone()))
        for _ in range(2):                      # second pass: recycled workspaces
            logits.clear()
            model.zero_grad()
            out = model(feats, programs, answers)
            out["loss"].mean().backward()
        hook.remove()
        g = torch.cat([p.grad.flatten() for n, p in model.named_parameters() if not n.startswith("classifier.")]).clone()
        results.append((logits[0], out["predictions"].clone(), out["loss"].detach().clone(), g, model.last_plan_stats[0]))
    a, b = results
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and torch.equal(a[2], b[2])
    assert float((a[3] - b[3]).abs().max()) <= 2e-5 * float(a[3].abs().max())
    assert a[4] == b[4] == B - 3 and int((a[1] == 28).sum()) == 3
