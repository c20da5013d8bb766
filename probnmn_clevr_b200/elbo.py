"""ELBO / REINFORCE glue of the question-coding and joint-training objectives (reference: probnmn/modules/elbo.py) over
the CUDA drop-ins, with the same class names, constructor arguments, ``forward`` signatures and output dictionaries.

What differs from the reference:

* the reward, the centred reward, the moving-average baseline update and the five scalars of the output dictionary are
  ONE kernel launch (``pnmn_elbo_glue``, ``csrc/optim.cu``) inside one autograd node, and the baseline lives on the
  device: the reference reads ``centered_reward.mean().item()`` -- a device synchronisation -- in every step
  (elbo.py:33).  ``Reinforce._reinforce_baseline`` still reads as a python float (it synchronises when read);
* ``JointTrainingElbo.forward`` may run the question reconstructor and the program prior on a side stream while the
  host compiles the sampled programs for the module network (``concurrent=True``, the default); results are identical.

The reference's own ``probnmn/modules/elbo.py`` also runs unmodified over the drop-ins; it then keeps
its per-step synchronisation and its ~25 small element-wise launches.
"""
import ctypes
from typing import Dict, Optional

import torch
from torch import nn

from . import _lib as L


class _ElboGlueFn(torch.autograd.Function):
    """(pg_loss, qr_loss, prior_loss, nmn_loss) -> (stats[5], centered): see ``pnmn_elbo_glue`` in include/pnmn.h.
    Differentiable outputs: stats[2] = elbo (w.r.t. pg_loss and qr_loss) and stats[4] = mean nmn loss."""

    @staticmethod
    def forward(ctx, pg_loss, qr_loss, prior_loss, nmn_loss, baseline, beta, gamma, decay, mode):
        n = pg_loss.numel()
        dev = pg_loss.device
        f32 = lambda t: None if t is None else t.detach().contiguous().float()
        pg, qr, pr, nm = f32(pg_loss), f32(qr_loss), f32(prior_loss), f32(nmn_loss)
        stats = torch.empty(5, dtype=torch.float32, device=dev)
        centered = torch.empty(n, dtype=torch.float32, device=dev)
        ptr = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
        with torch.cuda.device(dev):
            stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            L.check(L.lib().pnmn_elbo_glue(ptr(pg), ptr(qr), ptr(pr), ptr(nm), n, float(beta), float(gamma), float(decay),
                                           int(mode), ptr(baseline), ptr(centered), ptr(stats), stream), "pnmn_elbo_glue")
        ctx.save_for_backward(centered)
        ctx.beta, ctx.mode, ctx.n = float(beta), int(mode), n
        ctx.has = (qr_loss is not None, nmn_loss is not None)
        ctx.mark_non_differentiable(centered)
        return stats, centered

    @staticmethod
    def backward(ctx, g_stats, _g_centered):
        (centered,) = ctx.saved_tensors
        inv = 1.0 / max(ctx.n, 1)
        g_elbo, g_nmn = g_stats[2], g_stats[4]
        if ctx.mode == 0:
            # elbo = mean(-qr_loss + pg_loss * centered - beta * pg_loss)      (lp = -loss; elbo.py:61-89)
            d_pg = g_elbo * inv * (centered - ctx.beta)
            d_qr = (-g_elbo * inv).expand(ctx.n) if ctx.has[0] else None
        else:
            d_pg = g_elbo * inv * centered                     # elbo = mean(pg_loss * centered)   (elbo.py:245-250)
            d_qr = None
        d_nmn = (g_nmn * inv).expand(ctx.n) if ctx.has[1] else None
        return d_pg, d_qr, None, d_nmn, None, None, None, None, None


def elbo_glue(pg_loss, qr_loss, prior_loss, nmn_loss, baseline, beta, gamma, decay, mode=0):
    """``pnmn_elbo_glue`` without autograd: returns ``(stats[5], centered[n])`` and updates ``baseline`` in place.  For callers
    that apply the closed-form gradients of the objective themselves (``joint.JointTrainingStep``): with lp = -loss,
    d(-elbo)/d(pg_loss) = (beta - centered) / n, d(-elbo)/d(qr_loss) = 1 / n (elbo.py:61-89)."""
    return _ElboGlueFn.apply(pg_loss.detach(), None if qr_loss is None else qr_loss.detach(),
                             None if prior_loss is None else prior_loss.detach(),
                             None if nmn_loss is None else nmn_loss.detach(), baseline, beta, gamma, decay, mode)


class Reinforce(nn.Module):
    r"""REINFORCE with a decaying moving-average baseline (elbo.py:12-34): ``forward(inputs, reward)`` returns
    ``inputs * (reward.detach() - baseline)`` and then moves the baseline by ``decay * mean(reward - baseline)`` -- not a
    textbook exponential average; kept as the reference has it.  The baseline is a device scalar."""

    def __init__(self, baseline_decay: float = 0.99):
        super().__init__()
        self._baseline_decay = baseline_decay
        self.register_buffer("_baseline", torch.zeros(1), persistent=False)

    @property
    def _reinforce_baseline(self) -> float:
        return float(self._baseline.item())

    @_reinforce_baseline.setter
    def _reinforce_baseline(self, value: float) -> None:
        self._baseline.fill_(float(value))

    def baseline_on(self, device) -> torch.Tensor:
        if self._baseline.device != device:
            self._baseline = self._baseline.to(device)
        return self._baseline

    def forward(self, inputs, reward):
        baseline = self.baseline_on(inputs.device)
        centered_reward = reward.detach() - baseline
        baseline.add_(self._baseline_decay * centered_reward.mean())   # no .item(): nothing leaves the device
        return inputs * centered_reward


class _ElboWithReinforce(nn.Module):
    r"""Fully Monte Carlo evidence lower bound from per-row losses (elbo.py:37-89)."""

    def __init__(self, beta: float = 0.1, baseline_decay: float = 0.99):
        super().__init__()
        self._reinforce = Reinforce(baseline_decay=baseline_decay)
        self._beta = beta

    def _glue(self, pg_loss, qr_loss, prior_loss, nmn_loss, gamma: float, mode: int) -> Dict[str, torch.Tensor]:
        baseline = self._reinforce.baseline_on(pg_loss.device)
        stats, _ = _ElboGlueFn.apply(pg_loss, qr_loss, prior_loss, nmn_loss, baseline, self._beta, gamma,
                                     self._reinforce._baseline_decay, mode)
        if mode == 0:
            out = {"reconstruction_likelihood": stats[0].detach(), "kl_divergence": stats[1].detach(), "elbo": stats[2],
                   "reinforce_reward": stats[3].detach()}
        else:
            out = {"elbo": stats[2], "reinforce_reward": stats[3].detach()}
        if nmn_loss is not None:
            out["nmn_loss"] = stats[4]
        return out


class QuestionCodingElbo(_ElboWithReinforce):
    r"""ELBO for questions without program supervision (elbo.py:92-163); same arguments and outputs."""

    def __init__(self, program_generator, question_reconstructor, program_prior, beta: float = 0.1,
                 baseline_decay: float = 0.99):
        super().__init__(beta, baseline_decay)
        self._program_generator = program_generator
        self._question_reconstructor = question_reconstructor
        self._program_prior = program_prior

    def forward(self, question_tokens: torch.LongTensor):
        pg = self._program_generator(question_tokens, decoding_strategy="sampling")
        sampled_programs = pg["predictions"]
        qr = self._question_reconstructor(sampled_programs, question_tokens, decoding_strategy="sampling")
        prior = self._program_prior(sampled_programs)
        return self._glue(pg["loss"], qr["loss"], prior["loss"], None, 0.0, 0)


class JointTrainingElbo(_ElboWithReinforce):
    r"""ELBO with the answer log-likelihood term (elbo.py:166-282); same arguments and outputs (``"nmn_loss"`` included).

    ``concurrent``: run the question reconstructor and the program prior on a side stream while this thread waits for
    the sampled programs and compiles them for the module network (host work the reference does not have).
    """

    def __init__(self, program_generator, question_reconstructor, program_prior, nmn, beta: float = 0.1,
                 gamma: float = 10, baseline_decay: float = 0.99, objective: str = "ours", concurrent: bool = True):
        super().__init__(beta, baseline_decay)
        self._program_generator = program_generator
        self._question_reconstructor = question_reconstructor
        self._program_prior = program_prior
        self._nmn = nmn
        self._gamma = gamma
        self._objective = objective
        self._concurrent = concurrent
        self._side: Optional[torch.cuda.Stream] = None

    def forward(self, question_tokens: torch.LongTensor, image_features: torch.FloatTensor,
                answer_tokens: torch.LongTensor):
        dev = question_tokens.device
        pg = self._program_generator(question_tokens, decoding_strategy="sampling")
        sampled_programs = pg["predictions"]
        if self._concurrent and dev.type == "cuda":
            main = torch.cuda.current_stream(dev)
            if self._side is None or self._side.device != dev:
                self._side = torch.cuda.Stream(dev)
            side = self._side
            side.wait_stream(main)
            with torch.cuda.stream(side):
                qr = self._question_reconstructor(sampled_programs, question_tokens, decoding_strategy="sampling")
                prior = self._program_prior(sampled_programs) if self._objective != "baseline" else None
            nmn = self._nmn(image_features, sampled_programs, answer_tokens)
            main.wait_stream(side)
            for t in (qr["loss"], qr["predictions"]) + ((prior["loss"], prior["predictions"]) if prior else ()):
                t.record_stream(main)
        else:
            qr = self._question_reconstructor(sampled_programs, question_tokens, decoding_strategy="sampling")
            nmn = self._nmn(image_features, sampled_programs, answer_tokens)
            prior = self._program_prior(sampled_programs) if self._objective != "baseline" else None
        if self._objective == "baseline":
            out = self._glue(pg["loss"], None, None, nmn["loss"], self._gamma, 1)
        else:
            out = self._glue(pg["loss"], qr["loss"], prior["loss"], nmn["loss"], self._gamma, 0)
        self.last_outputs = {"program_generator": pg, "question_reconstructor": qr, "nmn": nmn, "program_prior": prior}
        return out
