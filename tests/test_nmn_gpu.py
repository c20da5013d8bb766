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


def _grad_errors(model, ref_grads):
    """per-tensor max|a-b|/max|b| and the global relative L2 error over all parameters"""
    worst, num, den = (0.0, ""), 0.0, 0.0
    for k, p in model.named_parameters():
        ref = ref_grads[k]
        mine = torch.zeros_like(ref) if p.grad is None else p.grad.detach().cpu()
        num += float((mine.double() - ref.double()).pow(2).sum())
        den += float(ref.double().pow(2).sum())
        if ref.abs().max() == 0:
            assert mine.abs().max() == 0, k
            continue
        worst = max(worst, (float((mine - ref).abs().max() / ref.abs().max()), k))
    return worst, (num / den) ** 0.5


@pytest.mark.parametrize("name", ["semantic", "sampled"])
def test_backward_matches_tf32_oracle(model, name):
    """Parameter gradients against the oracle run with tf32 operand rounding at the same points (see
    oracle/nmn_oracle.py: gradients of this net are ill-conditioned w.r.t. 1e-3 forward perturbations, the
    fp32 reference itself moves by up to ~20% per tensor under tf32 rounding).
    Bound: 5e-3 of each tensor's max |grad| (tf32 dgrad operands, fp16 wgrad operands, fp32 accumulation)."""
    g = np.load(GOLDEN)
    vocab = model.vocabulary
    programs = torch.from_numpy(g[f"{name}.programs"])
    answers = torch.from_numpy(g[f"{name}.answers"])
    feats = make_features(programs.shape[0], 0)
    sd = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in model.state_dict().items()}
    with nmn_oracle.operand_rounding("tf32"):
        ref = nmn_oracle.nmn_forward(sd, vocab, feats, programs, answers)
    ref["loss"].mean().backward()
    ref_grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in sd.items()}
    out, _ = _run(model, programs, answers, feats)
    out["loss"].mean().backward()
    worst, l2 = _grad_errors(model, ref_grads)
    print(f"[{name}] vs tf32 oracle: worst per-tensor grad rel err {worst[0]:.2e} at {worst[1]}; global L2 rel err {l2:.2e}")
    assert worst[0] < 5e-3 and l2 < 2e-3


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
