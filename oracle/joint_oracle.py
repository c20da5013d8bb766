"""ORACLE (test infrastructure, not product code) — CPU restatement of one ``joint_training`` iteration:
``probnmn/modules/elbo.py`` (REINFORCE baseline, ELBO) and ``JointTrainingTrainer._do_iteration`` + the optimizer step
(``probnmn/trainers/joint_training_trainer.py:128-198``, ``probnmn/trainers/_trainer.py:103-108,193``) over the model
oracles of this directory.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may import this file.

Parity status: the ELBO / REINFORCE arithmetic is PINNED -- ``oracle/make_elbo_golden.py`` runs the reference's own
``probnmn/modules/elbo.py`` (unmodified source, loaded by path over a stub of ``probnmn.models``) on seeded per-row
losses and stores inputs, outputs, gradients and the baseline trajectory in ``tests/golden/elbo_golden.npz``;
``tests/test_joint_cpu.py`` checks this restatement against it.  The whole iteration is PINNED as well:
``oracle/make_joint_golden.py`` calls the reference's own ``JointTrainingTrainer._do_iteration``
(joint_training_trainer.py:128-198, loaded by path, unmodified) over the reference's own elbo.py and four model classes
on seeded weights / batches and records ``tests/golden/joint_golden.npz`` (sampled programs, output dictionary, baseline
over two iterations, clamped gradients of the three trained models); ``joint_iteration`` below reproduces it to 3e-6
when it replays the sampled programs.
The model oracles keep their own status (NMN pinned; seq2seq / prior pinned against the reference's own files run over a
shim of the absent AllenNLP, whose ~60 restated lines stay unpinned).
"""
from typing import Dict, Optional

import torch

from . import nmn_oracle, prior_oracle, seq2seq_oracle


class ElboState:
    """The moving-average REINFORCE baseline (elbo.py:24-26)."""

    def __init__(self, baseline: float = 0.0):
        self.baseline = baseline


def reinforce(state: ElboState, inputs, reward, decay):
    """elbo.py:28-34"""
    centered = reward.detach() - state.baseline
    state.baseline += decay * centered.mean().item()
    return inputs * centered


def elbo_terms(state: ElboState, lp_gen, lp_rec, reward, beta, decay):
    """_ElboWithReinforce._forward (elbo.py:61-89)"""
    kl = reinforce(state, lp_gen, reward, decay) - beta * lp_gen
    elbo = lp_rec - kl
    return {"reconstruction_likelihood": lp_rec.mean(), "kl_divergence": kl.mean(), "elbo": elbo.mean(),
            "reinforce_reward": reward.mean()}


def joint_elbo(state: ElboState, pg_loss, qr_loss, prior_loss, nmn_loss, beta, gamma, decay, objective="ours"):
    """JointTrainingElbo.forward after the model calls (elbo.py:241-282)."""
    if objective == "baseline":
        reward = -nmn_loss
        out = {"elbo": reinforce(state, pg_loss, reward, decay).mean(), "reinforce_reward": reward.mean()}
    else:
        lp_rec, lp_gen, lp_prior, lp_ans = -qr_loss, -pg_loss, -prior_loss, -nmn_loss
        reward = lp_rec + beta * lp_prior - beta * lp_gen + gamma * lp_ans
        out = elbo_terms(state, lp_gen, lp_rec, reward, beta, decay)
    out["nmn_loss"] = nmn_loss.mean()
    return out


def question_coding_elbo(state: ElboState, pg_loss, qr_loss, prior_loss, beta, decay):
    """QuestionCodingElbo.forward after the model calls (elbo.py:143-163)."""
    lp_rec, lp_gen, lp_prior = -qr_loss, -pg_loss, -prior_loss
    return elbo_terms(state, lp_gen, lp_rec, lp_rec + beta * (lp_prior - lp_gen), beta, decay)


def joint_iteration(sds: Dict[str, Dict[str, torch.Tensor]], vocab, batch, state: ElboState, alpha=100.0, beta=0.1,
                    gamma=1.0, delta=0.99, objective="ours", forced_programs: Optional[torch.Tensor] = None,
                    generator: Optional[torch.Generator] = None):
    """JointTrainingTrainer._do_iteration (joint_training_trainer.py:128-198) up to and including ``backward()``; the
    state dicts' tensors must require grad.  ``forced_programs`` replays the programs another implementation sampled
    for the unsupervised rows (RNG streams cannot be matched).  Returns the reference's iteration dictionary plus the
    sampled programs."""
    sup = batch["supervision"]
    isup, iu = sup.nonzero().flatten(), (1 - sup).nonzero().flatten()          # :131-132
    q_u, f_u, a_u = batch["question"][iu], batch["image"][iu], batch["answer"][iu]
    pg = seq2seq_oracle.seq2seq_forward(sds["program_generator"], q_u, None, "sampling", 26, generator, forced_programs)
    programs = pg["predictions"]                                                # elbo.py:233
    qr = seq2seq_oracle.seq2seq_forward(sds["question_reconstructor"], programs, q_u, "sampling", 45, generator)
    nmn = nmn_oracle.nmn_forward(sds["nmn"], vocab, f_u, programs, a_u)         # elbo.py:239
    with torch.no_grad():
        prior = prior_oracle.prior_forward(sds["program_prior"], programs, generator)
    elbo = joint_elbo(state, pg["loss"], qr["loss"], prior["loss"], nmn["loss"], beta, gamma, delta, objective)
    nmn_loss = elbo.pop("nmn_loss")                                             # :145
    loss = gamma * nmn_loss - elbo["elbo"]                                      # :146
    out = {"loss": {"nmn": nmn_loss.detach()}, "elbo": {k: v.detach() for k, v in elbo.items()}}
    if objective == "ours":                                                     # :148-176
        p_s, q_s = batch["program"][isup], batch["question"][isup]
        pg_s = seq2seq_oracle.seq2seq_forward(sds["program_generator"], q_s, p_s, "sampling", 26, generator)
        qr_s = seq2seq_oracle.seq2seq_forward(sds["question_reconstructor"], p_s, q_s, "sampling", 45, generator)
        pg_l, qr_l = pg_s["loss"].mean(), qr_s["loss"].mean()
        loss = loss + alpha * (pg_l + qr_l)
        out["loss"].update({"question_reconstruction_gt": qr_l.detach(), "program_generation_gt": pg_l.detach()})
    loss.backward()                                                             # :178
    out["objective"] = loss.detach()
    out["sampled_programs"] = programs
    out["rows"] = {"pg_loss": pg["loss"].detach(), "qr_loss": qr["loss"].detach(), "prior_loss": prior["loss"].detach(),
                   "nmn_loss": nmn["loss"].detach(), "nmn_valid": nmn["valid"]}
    return out


def trained_parameters(sds):
    """parameter order of the reference's optimizer (trainers/_trainer.py:103-108)"""
    return [p for name in ("program_generator", "question_reconstructor", "nmn") for p in sds[name].values()]


def clamp_and_step(sds, optimizer, clamp=5.0):
    """joint_training_trainer.py:182-188 + trainers/_trainer.py:193"""
    for p in trained_parameters(sds):
        if p.grad is not None:
            p.grad.clamp_(min=-clamp, max=clamp)
    optimizer.step()
