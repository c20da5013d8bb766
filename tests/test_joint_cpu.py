"""CPU: the joint-training oracles (oracle/joint_oracle.py, oracle/prior_oracle.py) against the golden vectors recorded
from the reference's own probnmn/modules/elbo.py (tests/golden/elbo_golden.npz, oracle/make_elbo_golden.py) and against
independent torch modules; host-side helpers of the joint step."""
import os

import numpy as np
import pytest
import torch

from oracle import joint_oracle, prior_oracle, seq2seq_oracle
from probnmn_clevr_b200.synthetic import (ProgramSampler, make_joint_batch, make_prior_state_dict, questions_for_programs)
from probnmn_clevr_b200.vocabulary import Vocabulary

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "elbo_golden.npz")
SIZES = (1, 7, 128, 131)


@pytest.mark.parametrize("mode", ["joint_ours", "joint_baseline", "question_coding"])
def test_elbo_restatement_matches_reference_golden(mode):
    """outputs, d(objective)/d(per-row losses) and the moving-average baseline over three successive calls"""
    g = np.load(GOLDEN)
    beta, gamma, decay = g["hyper"]
    for n in SIZES:
        state = joint_oracle.ElboState()
        for k in range(3):
            tag = f"{mode}.n{n}.call{k}"
            leaves = {key: torch.from_numpy(g[f"{tag}.in.{key}"]).clone().requires_grad_(True) for key in ("pg", "qr", "prior", "nmn")}
            if mode == "question_coding":
                out = joint_oracle.question_coding_elbo(state, leaves["pg"], leaves["qr"], leaves["prior"], beta, decay)
                objective = -out["elbo"]
            else:
                out = joint_oracle.joint_elbo(state, leaves["pg"], leaves["qr"], leaves["prior"], leaves["nmn"], beta, gamma,
                                              decay, "ours" if mode == "joint_ours" else "baseline")
                objective = gamma * out["nmn_loss"] - out["elbo"]
            objective.backward()
            for key, v in out.items():
                np.testing.assert_allclose(v.item(), g[f"{tag}.out.{key}"], rtol=1e-6, atol=1e-7, err_msg=f"{tag} {key}")
            for key, leaf in leaves.items():
                mine = leaf.grad.numpy() if leaf.grad is not None else np.zeros(n, np.float32)
                np.testing.assert_allclose(mine, g[f"{tag}.grad.{key}"], rtol=1e-6, atol=1e-9, err_msg=f"{tag} grad {key}")
            np.testing.assert_allclose(state.baseline, g[f"{tag}.baseline_after"], rtol=1e-9)


def test_elbo_golden_regenerates_from_the_reference_when_present():
    """pins the golden file itself: where /root/reference exists (build container) the reference's elbo.py, run again,
    must reproduce the committed vectors"""
    if not os.path.exists("/root/reference/probnmn/modules/elbo.py"):
        pytest.skip("/root/reference is not available here")
    from oracle.make_elbo_golden import StubModel, cases, load_reference_elbo
    ref = load_reference_elbo()
    g = np.load(GOLDEN)
    n, calls = cases()[2]
    models = [StubModel() for _ in range(4)]
    elbo = ref.JointTrainingElbo(*models, beta=0.1, gamma=1.5, baseline_decay=0.99, objective="ours")
    for k, c in enumerate(calls):
        for m, key in zip(models, ("pg", "qr", "prior", "nmn")):
            m.queue.append({"predictions": torch.zeros(n, 3, dtype=torch.long), "loss": c[key]})
        out = elbo(torch.zeros(n, 3, dtype=torch.long), torch.zeros(n, 1), torch.zeros(n, dtype=torch.long))
        assert np.float32(out["elbo"].item()) == g[f"joint_ours.n{n}.call{k}.out.elbo"]


def test_prior_oracle_against_torch_modules():
    """the restated ProgramPrior arithmetic against nn.Embedding + nn.LSTM (packed) + tied Linear layers"""
    vocab = Vocabulary.clevr()
    sd = make_prior_state_dict(44, seed=3)
    programs = ProgramSampler(vocab, seed=5).sample(6, 26)
    with torch.no_grad():
        mine = prior_oracle.prior_forward(sd, programs, torch.Generator().manual_seed(0))
    lstm = torch.nn.LSTM(256, 256, num_layers=2, batch_first=True)
    lstm.load_state_dict({k.split("._module.")[1]: v for k, v in sd.items() if "_module" in k})
    tokens = seq2seq_oracle.add_sentence_boundary_token_ids(programs)
    lengths = (tokens != 0).sum(1)
    x = torch.nn.functional.embedding(tokens, sd["_embedder.token_embedder_programs.weight"])
    packed = torch.nn.utils.rnn.pack_padded_sequence(x, lengths, batch_first=True, enforce_sorted=False)
    with torch.no_grad():
        y, _ = torch.nn.utils.rnn.pad_packed_sequence(lstm(packed)[0], batch_first=True, total_length=tokens.shape[1])
        logits = (y @ sd["_projection_layer.weight"].t()) @ sd["_output_layer.weight"].t()
        nll = torch.nn.functional.cross_entropy(logits[:, :-1].reshape(-1, 44), tokens[:, 1:].reshape(-1), reduction="none").view(6, -1)
        m = (tokens[:, 1:] != 0).float()
        loss = (nll * m).sum(1) / (m.sum(1) + 1e-13)
    np.testing.assert_allclose(mine["loss"].numpy(), loss.numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(mine["logits"].numpy(), logits[:, :-1].numpy(), rtol=1e-4, atol=1e-5)
    pred = mine["predictions"]
    assert pred.shape == (6, 27) and ((pred == 0) == (tokens[:, 1:] == 0)).all() and (pred[pred != 0] > 2).all()


def test_synthetic_joint_batch_and_host_split():
    from probnmn_clevr_b200.joint import split_batch
    vocab = Vocabulary.clevr()
    batch = make_joint_batch(vocab, 64, seed=2)
    assert batch["question"].shape == (64, 40) and batch["program"].shape == (64, 26) and batch["image"].shape == (64, 1024, 14, 14)
    assert int(batch["question"].max()) < vocab.get_vocab_size("questions")
    # the question determines the program: same program -> same content prefix
    q2 = questions_for_programs(batch["program"], 93, seed=99)
    n = (batch["program"] != 0).sum(1)
    for b in range(64):
        assert torch.equal(q2[b, : n[b]], batch["question"][b, : n[b]])
    parts = split_batch(batch)
    sup = batch["supervision"].bool()
    assert torch.equal(parts["unsup"]["question"], batch["question"][~sup]) and torch.equal(parts["sup"]["program"], batch["program"][sup])
    assert parts["unsup"]["image"].shape[0] == int((~sup).sum()) and "image" not in parts["sup"]
    assert 16 < int(sup.sum()) < 48


# ---- the whole iteration against the reference's own trainer code (tests/golden/joint_golden.npz) ---------------------------
JOINT_GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "joint_golden.npz")


def test_joint_iteration_restatement_matches_reference_golden():
    """oracle/joint_oracle.joint_iteration against what the reference's OWN ``JointTrainingTrainer._do_iteration`` computed
    (joint_training_trainer.py:128-198 over the reference's JointTrainingElbo, ProgramGenerator, QuestionReconstructor,
    ProgramPrior and NeuralModuleNetwork, all verbatim: oracle/make_joint_golden.py) on the same seeded weights and
    batches, replaying the programs the reference sampled: output dictionary, REINFORCE baseline over two iterations,
    clamped gradients of all three trained models."""
    from oracle import joint_oracle
    from oracle.make_joint_golden import BATCH, BATCH_SEEDS, CLAMP, HYPER, SUB, state_dicts
    from probnmn_clevr_b200.synthetic import make_joint_batch
    from probnmn_clevr_b200.vocabulary import Vocabulary
    g = np.load(JOINT_GOLDEN)
    vocab = Vocabulary.clevr()
    torch.set_num_threads(os.cpu_count())
    sds = {name: {k: v.clone().requires_grad_(True) for k, v in sd.items()} for name, sd in state_dicts(vocab).items()}
    sds["program_prior"]["_output_layer.weight"] = sds["program_prior"]["_embedder.token_embedder_programs.weight"]
    state = joint_oracle.ElboState(0.0)
    for it, seed in enumerate(BATCH_SEEDS):
        batch = make_joint_batch(vocab, BATCH, seed=seed)
        for p in joint_oracle.trained_parameters(sds):
            p.grad = None
        raw = torch.from_numpy(g[f"it{it}.raw_programs"])
        out = joint_oracle.joint_iteration(sds, vocab, batch, state, objective="ours", forced_programs=raw, **HYPER)
        assert int(out["rows"]["nmn_valid"].sum()) == raw.shape[0]      # the pre-trained generator's samples are executable
        for p in joint_oracle.trained_parameters(sds):
            if p.grad is not None:
                p.grad.clamp_(min=-CLAMP, max=CLAMP)
        for group in ("elbo", "loss"):
            for k, v in out[group].items():
                want = float(g[f"it{it}.{group}.{k}"])
                assert abs(v.item() - want) <= 1e-5 * max(1.0, abs(want)), (it, group, k)
        assert abs(out["objective"].item() - float(g[f"it{it}.objective"])) <= 1e-5 * abs(float(g[f"it{it}.objective"]))
        assert abs(state.baseline - float(g[f"it{it}.baseline"])) <= 1e-5 * abs(float(g[f"it{it}.baseline"]))
        for name in ("program_generator", "question_reconstructor", "nmn"):
            for k, p in sds[name].items():
                got = torch.zeros_like(p) if p.grad is None else p.grad
                norm = float(g[f"it{it}.gradnorm.{name}.{k}"])
                assert abs(float(got.double().norm()) - norm) <= 2e-4 * norm + 1e-9, (it, name, k)
                want = torch.from_numpy(g[f"it{it}.gradsub.{name}.{k}"])
                sub = got.reshape(-1)[::(SUB if got.numel() > 4096 else 1)]
                assert float((sub - want).abs().max()) <= 2e-4 * float(want.abs().max()) + 1e-9, (it, name, k)
