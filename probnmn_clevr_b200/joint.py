"""One iteration of the ``joint_training`` phase over the CUDA drop-ins.

Mirrors the reference's ``_Trainer.step`` for this phase -- ``optimizer.zero_grad()`` (trainers/_trainer.py:193),
``JointTrainingTrainer._do_iteration`` (trainers/joint_training_trainer.py:128-198: unsupervised rows through
``JointTrainingElbo``, supervised rows teacher-forced through the program generator and the question reconstructor,
``loss_objective.backward()``, element-wise gradient clamp) and ``optimizer.step()`` (:193) -- with the same objective,
the same hyper-parameter names (ALPHA, BETA, GAMMA, DELTA, OBJECTIVE of the config) and the same returned dictionary.
The trainer's data loading, logging, checkpointing and LR scheduling stay the reference's (out of scope).

What is different is only HOW the work is issued:

* the supervised / unsupervised split is taken on the host when ``batch["supervision"]`` is a host tensor (the reference
  calls ``.nonzero()`` on the device: one synchronisation, joint_training_trainer.py:131-132);
* independent passes run on separate CUDA streams: the teacher-forced passes over the supervised rows on one, the
  question reconstructor + program prior over the sampled programs on another, while this thread compiles the sampled
  programs for the module executor; autograd replays each pass on its own stream;
* rows of both kinds share passes (``fused=True``, the default): the program generator runs ONCE over all rows -- the
  unsupervised ones decode freely, the supervised ones are teacher-forced (``Seq2SeqBase.forward_mixed``) -- and the
  question reconstructor runs ONCE over the sampled and the ground-truth programs together (both calls are teacher-forced
  on the questions).  An LSTM pass is a chain of ~450 dependent kernels whose duration hardly depends on the number of
  rows, so two passes instead of four halve the recurrent part of the step; per row the arithmetic is unchanged;
* the gradient clamp is fused into the optimizer (``optim.FusedClampAdam``); with more than one process the gradients are
  averaged over NCCL before it (clamp AFTER the reduction, as ``nn.DataParallel`` + ``clamp_`` does in the reference).
"""
import os
from typing import Any, Dict, Optional

import torch
import torch.distributed as dist

from .elbo import JointTrainingElbo, elbo_glue
from .optim import FusedClampAdam


def split_batch(batch: Dict[str, torch.Tensor]) -> Dict[str, Dict[str, torch.Tensor]]:
    """Host-side split of a reference-style batch (keys ``question``, ``answer``, ``program``, ``image``,
    ``supervision``; probnmn/data/datasets.py:209-228) into the rows without program supervision (``"unsup"``: question,
    image, answer) and the rows with it (``"sup"``: question, program) -- joint_training_trainer.py:131-138,160-161.
    An input pipeline calls this BEFORE the host -> device copy: the features of supervised rows are never used by the
    step, so half of the 205 MB batch does not have to travel."""
    sup = batch["supervision"].to(torch.bool)
    iu = (~sup).nonzero().flatten()
    isup = sup.nonzero().flatten()
    return {"unsup": {"question": batch["question"][iu], "image": batch["image"][iu], "answer": batch["answer"][iu]},
            "sup": {"question": batch["question"][isup], "program": batch["program"][isup]}}


class JointTrainingStep:
    r"""
    Parameters
    ----------
    program_generator, question_reconstructor, nmn: trained models (``probnmn_clevr_b200`` drop-ins), as in
        ``JointTrainingTrainer.__init__`` (joint_training_trainer.py:81-105)
    program_prior: frozen ``ProgramPrior`` (:107-112); put in ``eval()`` mode here
    alpha, beta, gamma, delta, objective: ALPHA / BETA / GAMMA / DELTA / OBJECTIVE of the config
        (configs/joint_training_ours.yml:6-16)
    lr, weight_decay: OPTIM.LR_INITIAL / OPTIM.WEIGHT_DECAY (:18-25); clamp: the [-5, 5] gradient clamp (:187-188)
    concurrent: issue independent passes on side streams (results do not depend on it)
    defer_nmn: software-pipeline the module network's backward pass + update into the next ``step`` (see ``flush``)
    group: process group for data-parallel gradient averaging (default group when ``torch.distributed`` is initialised)
    """

    def __init__(self, program_generator, question_reconstructor, nmn, program_prior, alpha: float = 100.0,
                 beta: float = 0.1, gamma: float = 1.0, delta: float = 0.99, objective: str = "ours", lr: float = 1e-6,
                 weight_decay: float = 0.0, clamp: Optional[float] = 5.0, concurrent: bool = True, fused: bool = True,
                 group=None, reserved_sms: Optional[int] = None, defer_nmn: Optional[bool] = None):
        self.program_generator, self.question_reconstructor = program_generator, question_reconstructor
        self.nmn, self.program_prior = nmn, program_prior
        program_prior.eval()
        self.alpha, self.gamma, self.objective = alpha, gamma, objective
        self.elbo = JointTrainingElbo(program_generator, question_reconstructor, program_prior, nmn, beta=beta, gamma=gamma,
                                      baseline_decay=delta, objective=objective, concurrent=concurrent)
        trained = [program_generator, question_reconstructor, nmn]
        # same parameter order as the reference's optimizer (trainers/_trainer.py:103-108: models in dict order)
        params = [p for m in trained for p in m.parameters()]
        self.optimizer = FusedClampAdam(params, lr=lr, weight_decay=weight_decay, clamp=clamp, modules=trained)
        self.concurrent = concurrent
        self.fused = fused
        # SMs the module executor's persistent kernels leave to the LSTM passes running next to them
        # (pnmn_set_reserved_sms, include/pnmn.h).  Measured with bench.py (two runs each, 1 x B200): 0 SMs 6.83 / 6.86 ms
        # per step, 8: 6.74 / 6.72, 12: 6.69 / 6.69, 16: 6.65 / 6.66, 20: 6.68 / 6.69, 32: 6.95, 40: 7.36.
        if reserved_sms is None:
            reserved_sms = int(os.environ.get("PNMN_JOINT_RESERVE_SMS", "16"))
        self.reserved_sms = reserved_sms if concurrent else 0
        self.prestage = os.environ.get("PNMN_JOINT_PRESTAGE", "1") != "0"
        # inside step(), every model is updated (clamp + Adam) right behind its own backward pass (and gradient average), on
        # that pass's stream, instead of all three after the whole backward phase: the module network's 257 MB update, bound
        # by HBM, then runs underneath the tail of the generator's backward pass, which is bound by latency (measured:
        # 6.04-6.26 against 6.13-6.27 ms per step, end to end 6.14-6.18 against 6.19-6.29).  PNMN_JOINT_EARLY_ADAM=0: one
        # optimizer.step() at the end
        self.early_adam = os.environ.get("PNMN_JOINT_EARLY_ADAM", "1") != "0"
        self._stepping = False
        # defer_nmn (opt-in; PNMN_JOINT_DEFER_NMN=1): inside step(), the module network's backward pass, gradient average and
        # update of step i are ISSUED at the start of step i + 1 -- next to the generator's forward pass of that step, which
        # needs the generator's new weights only and leaves most of the device idle -- instead of next to the two LSTM
        # backward passes of step i.  Same arithmetic, same order of operations per model; flush() issues what is pending
        # (call it before reading the module network's parameters or gradients, evaluating, or saving a checkpoint).
        # Measured (bench.py): 6.12 -> 5.75-5.82 ms per step.
        if defer_nmn is None:
            defer_nmn = os.environ.get("PNMN_JOINT_DEFER_NMN", "0") == "1"
        self.defer_nmn = bool(defer_nmn) and fused
        self._pending_nmn = None
        # PNMN_JOINT_DEFER_QR=1 (with defer_nmn): the question reconstructor's backward pass + update are deferred the same
        # way, onto the reconstructor's stream at the start of the next step (its next forward pass queues behind them).
        # Same results (tests/test_joint_gpu.py), no gain: 5.94 against 5.82 ms per step -- the reconstructor's next
        # forward pass, which the objective waits for, then starts later.  Default off.
        self.defer_qr = self.defer_nmn and os.environ.get("PNMN_JOINT_DEFER_QR", "0") == "1"
        self._pending_qr = None
        # issue order of the backward passes (experiment, see _do_iteration_fused): 0 = each right behind its forward pass
        self.order = int(os.environ.get("PNMN_JOINT_ORDER", "0"))
        # PNMN_JOINT_CHUNKS=k: the sampled programs are compiled as k independent plans on k host threads (the compile sits
        # between the generator's forward pass and the module network's: NeuralModuleNetwork._chunked_runs).  Measured at
        # k = 4: the module network's forward pass ends 0.23 ms earlier, the step is no faster (6.77 vs 6.94 ms) -- the step is
        # bound by the sum of its device work, not by this latency.  Default off.
        chunks = int(os.environ.get("PNMN_JOINT_CHUNKS", "1"))
        if chunks > 1 and getattr(nmn, "compile_chunks", None) == 1:
            nmn.compile_chunks = chunks
        self._qr_stream: Optional[torch.cuda.Stream] = None
        if concurrent:
            # passes on side streams accumulate into parameters whose AccumulateGrad node lives on another stream: intended
            try:
                torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)
            except AttributeError:
                pass
        self.group = group
        self._reduced: set = set()
        self._sup_stream: Optional[torch.cuda.Stream] = None
        self.iteration = -1
        self.trace = None   # diagnostics: a list collects (label, CUDA event) marks of one step (scripts/joint_timeline.py)

    def _streams(self, dev):
        if getattr(self, "_side_streams", None) is None or self._side_streams[0].device != dev:
            # (PNMN_JOINT_PRIORITY=1 gives the LSTM passes' streams high priority -- a pending CTA of a step kernel is then
            # placed before pending CTAs of the module network's bulk kernels.  Measured: the passes finish earlier, the
            # module network later, the step 0.2-0.3 ms slower: 6.87 vs 7.08 ms.  Default off.)
            mode = os.environ.get("PNMN_JOINT_PRIORITY", "0")   # "2": the generator's stream only (its backward pass ends the step)
            prios = {"1": (-1, -1, -1), "2": (-1, 0, 0)}.get(mode, (0, 0, 0))
            self._side_streams = tuple(torch.cuda.Stream(dev, priority=pr) for pr in prios)
        return self._side_streams

    def _mark(self, label: str) -> None:
        if self.trace is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self.trace.append((label, ev))

    @classmethod
    def from_config(cls, config, program_generator, question_reconstructor, nmn, program_prior, **kw):
        _C = config
        return cls(program_generator, question_reconstructor, nmn, program_prior, alpha=_C.ALPHA, beta=_C.BETA,
                   gamma=_C.GAMMA, delta=_C.DELTA, objective=_C.OBJECTIVE, lr=_C.OPTIM.LR_INITIAL,
                   weight_decay=_C.OPTIM.WEIGHT_DECAY, **kw)

    # ---- joint_training_trainer.py:128-198 ------------------------------------------------------------------------------
    def do_iteration(self, batch: Dict[str, Any]) -> Dict[str, Any]:
        if "unsup" in batch:
            parts = batch
        elif not batch["supervision"].is_cuda:
            parts = split_batch(batch)
        else:
            sup = batch["supervision"]
            isup, iu = sup.nonzero().flatten(), (1 - sup).nonzero().flatten()   # (device: synchronises, like the reference)
            parts = {"unsup": {k: batch[k][iu] for k in ("question", "image", "answer")},
                     "sup": {"question": batch["question"][isup], "program": batch["program"][isup]}}
        un, su = parts["unsup"], parts["sup"]
        dev = self.nmn.stem[0].weight.device
        to = lambda t: t if t.device == dev else t.to(dev, non_blocking=True)
        ours = self.objective == "ours"
        main = torch.cuda.current_stream(dev)
        if self.fused and ours and un["question"].shape[0] > 0 and su["question"].shape[0] > 0:
            return self._do_iteration_fused(un, su, dev, to, main)
        # (a module-network backward pass + update deferred by the previous step: this path runs the passes one by one,
        # nothing to overlap it with -- issue it first)
        self._issue_pending_qr(main)
        self._issue_pending_nmn()

        sup_out = None
        if ours:
            q_sup, p_sup = to(su["question"]), to(su["program"])
            if self.concurrent:
                if self._sup_stream is None or self._sup_stream.device != dev:
                    self._sup_stream = torch.cuda.Stream(dev)
                side = self._sup_stream
                side.wait_stream(main)
                with torch.cuda.stream(side):
                    sup_out = self._supervised(q_sup, p_sup)
            # (sequential order otherwise: after the ELBO, as in the reference)

        elbo_output_dict = self.elbo(to(un["question"]), to(un["image"]), to(un["answer"]))
        nmn_loss = elbo_output_dict.pop("nmn_loss")
        loss_objective = self.gamma * nmn_loss - elbo_output_dict["elbo"]

        if ours:
            if sup_out is None:
                sup_out = self._supervised(q_sup, p_sup)
            else:
                main.wait_stream(self._sup_stream)
                for t in sup_out:
                    t.record_stream(main)
            pg_sup, qr_sup = sup_out
            loss_objective = loss_objective + self.alpha * (pg_sup + qr_sup)

        loss_objective.backward()
        # (clamp of every gradient to [-5, 5], :182-188: fused into self.optimizer.step())

        out = {"loss": {"nmn": nmn_loss.detach()}, "elbo": {k: v.detach() for k, v in elbo_output_dict.items()}}
        if ours:
            out["loss"].update({"question_reconstruction_gt": qr_sup.detach(), "program_generation_gt": pg_sup.detach()})
        out["objective"] = loss_objective.detach()
        return out

    def _do_iteration_fused(self, un, su, dev, to, main) -> Dict[str, Any]:
        """The same objective with the rows of both kinds sharing one generator pass and one reconstructor pass."""
        from .seq2seq import _handover
        pg_m, qr_m = self.program_generator, self.question_reconstructor
        q_u, img, ans = to(un["question"]), to(un["image"]), to(un["answer"])
        q_s, p_s = to(su["question"]), to(su["program"])
        nu, ns = q_u.shape[0], q_s.shape[0]
        free = pg_m._max_decoding_steps
        width = max(p_s.shape[1], free - 1)
        q_all = torch.cat([q_u, q_s])
        targets = torch.zeros(nu + ns, width, dtype=torch.int64, device=dev)
        targets[nu:, : p_s.shape[1]] = p_s
        teacher_rows = torch.zeros(nu + ns, dtype=torch.uint8, device=dev)
        teacher_rows[nu:] = 1

        # (diagnostics, PNMN_DIAG_STALE_PROGRAMS=1: the module network runs an EARLIER step's sampled programs, compiled
        # ahead of time -- what the step's timeline would be if the program compiler took no time.  Not the objective.)
        stale = None
        if os.environ.get("PNMN_DIAG_STALE_PROGRAMS") == "1":
            if getattr(self, "_stale", None) is None:
                self._stale = {}
            stale = self._stale.get(nu)
            if stale is not None:
                self.nmn.precompile(stale)
        self._mark("start")
        # The objective is  gamma * mean(nmn_loss) - elbo + alpha * (mean(pg_loss_sup) + mean(qr_loss_sup))  with
        # -elbo = mean(qr_loss_u + (beta - centered) * pg_loss_u)  over the unsupervised rows (lp = -loss, elbo.py:61-89) and
        # centered = detached reward - baseline.  Its gradient w.r.t. every per-row loss is therefore known in closed form:
        #     d/d nmn_loss[b] = gamma / nu        d/d qr_loss[b] = 1 / nu (unsupervised), alpha / ns (supervised)
        #     d/d pg_loss[b]  = (beta - centered[b]) / nu (unsupervised), alpha / ns (supervised)
        # Only the generator's depends on the other models' outputs, so the reconstructor's and the module network's
        # backward passes are started right behind their forward passes instead of after the whole forward phase.
        # Streams: generator, reconstructor and prior each get their own, the module network stays on the caller's; autograd
        # replays every pass on the stream its forward ran on.
        nu_f, ns_f = float(nu), float(ns)
        order = self.order
        coef_qr = torch.empty(nu + ns, dtype=torch.float32, device=dev)
        coef_qr[:nu] = 1.0 / nu_f
        coef_qr[nu:] = self.alpha / ns_f
        coef_nmn = torch.full((nu,), self.gamma / nu_f, dtype=torch.float32, device=dev)
        pwidth = max(free, p_s.shape[1])
        streams = self._streams(dev) if self.concurrent else (main, main, main)
        s_pg, s_qr, s_prior = streams
        if self.concurrent:
            s_pg.wait_stream(main)
        with torch.cuda.stream(s_pg):
            pg = pg_m.forward_mixed(q_all, targets, teacher_rows, free_steps=free)       # elbo.py:230-233 + trainer :164-168
            sampled = pg["predictions"][:nu, :free].contiguous()
            if pg_m.handover_predictions:
                _handover(sampled)                                                     # the module network compiles on the host
            programs_all = torch.zeros(nu + ns, pwidth, dtype=torch.int64, device=dev)
            programs_all[:nu, :free] = sampled
            programs_all[nu:, : p_s.shape[1]] = p_s
            self._mark("pg_fwd_end(pg)")
        # (deferred backward passes of the previous step: the reconstructor's on its own stream, the module network's on the
        # caller's; the generator's forward pass waits for neither)
        self._issue_pending_qr(s_qr)
        self._issue_pending_nmn()
        # the module network's weights and features are known now, its programs only after the generator's forward pass:
        # weight packing and feature layout run here, on the caller's stream, next to that pass
        if self.prestage and nu > 0:
            self.nmn.prestage(img)
        if self.concurrent:
            s_qr.wait_stream(s_pg)
            s_prior.wait_stream(s_pg)
            main.wait_stream(s_pg)
        with torch.cuda.stream(s_qr):
            qr = qr_m(programs_all, q_all, decoding_strategy="sampling")              # elbo.py:236-238 + trainer :169-173
            qr_loss = qr["loss"].detach()
            qr_fwd_done = torch.cuda.Event()
            qr_fwd_done.record()
            self._mark("qr_fwd_end(qr)")
            if order == 0 and self._stepping and self.defer_qr:
                self._pending_qr = (qr["loss"], coef_qr)
                self._reduced.add(id(qr_m))
                self._updated.add(id(qr_m))
            elif order == 0:
                torch.autograd.backward([qr["loss"]], [coef_qr])
                self._mark("qr_bwd_end(qr)")
                self._reduce_early([qr_m])
                self._update_early(qr_m)
        with torch.cuda.stream(s_prior):
            prior = self.program_prior(sampled)                                        # elbo.py:256
            self._mark("prior_fwd_end(prior)")
        nmn = self.nmn(img, sampled if stale is None else stale, ans)                  # elbo.py:239
        self._mark("nmn_fwd_end")
        if order:
            # the reconstructor's backward pass waits for the module network's forward pass: an LSTM pass and the
            # persistent executor only take turns on the SMs, while two LSTM passes run side by side
            nmn_fwd_done = torch.cuda.Event()
            nmn_fwd_done.record()
            if order == 2:
                torch.autograd.backward([nmn["loss"]], [coef_nmn])
                self._mark("nmn_bwd_end")
                self._reduce_early([self.nmn])
                nmn_fwd_done = torch.cuda.Event()
                nmn_fwd_done.record()
            with torch.cuda.stream(s_qr):
                if self.concurrent:
                    s_qr.wait_event(nmn_fwd_done)
                torch.autograd.backward([qr["loss"]], [coef_qr])
                self._mark("qr_bwd_end(qr)")
                self._reduce_early([qr_m])
        nmn_loss_rows = nmn["loss"].detach()
        if self.concurrent:
            main.wait_event(qr_fwd_done)
            main.wait_stream(s_prior)
        pg_loss = pg["loss"].detach()
        pg_loss_u, qr_loss_u = pg_loss[:nu], qr_loss[:nu]
        baseline = self.elbo._reinforce.baseline_on(dev)
        stats, centered = elbo_glue(pg_loss_u, qr_loss_u, prior["loss"], nmn_loss_rows, baseline, self.elbo._beta, self.gamma,
                                    self.elbo._reinforce._baseline_decay, 0)
        coef_pg = torch.empty(nu + ns, dtype=torch.float32, device=dev)
        coef_pg[:nu] = (self.elbo._beta - centered) / nu_f
        coef_pg[nu:] = self.alpha / ns_f
        self._mark("objective")
        # the generator's backward pass first (it is the longest chain still to run), then the module network's
        if self.concurrent:
            s_pg.wait_stream(main)
        with torch.cuda.stream(s_pg):
            torch.autograd.backward([pg["loss"]], [coef_pg])
            self._mark("pg_bwd_end(pg)")
        if order == 1 and self.concurrent:
            main.wait_stream(s_pg)
            main.wait_stream(s_qr)
        defer = self._stepping and self.defer_nmn and order == 0
        if defer:
            self._pending_nmn = (nmn["loss"], coef_nmn)
            self._reduced.add(id(self.nmn))
            self._updated.add(id(self.nmn))
        elif order != 2:
            torch.autograd.backward([nmn["loss"]], [coef_nmn])
            self._mark("nmn_bwd_end")
        # gradient averaging, issued in the order in which the gradients become final (NCCL runs a communicator's collectives
        # in issue order): reconstructor (above), module network -- its classifier gradients are already travelling, started
        # by hooks inside its backward pass --, generator last (its backward pass is the last to finish)
        if order != 2 and not defer:
            self._reduce_early([self.nmn])
            self._update_early(self.nmn)
        with torch.cuda.stream(s_pg):
            self._reduce_early([pg_m])
            self._update_early(pg_m)
        pg_sup, qr_sup = pg_loss[nu:].mean(), qr_loss[nu:].mean()
        loss_objective = self.gamma * stats[4] - stats[2] + self.alpha * (pg_sup + qr_sup)
        if self.concurrent:
            main.wait_stream(s_pg)
            main.wait_stream(s_qr)
            for t in (pg["loss"], pg["predictions"], sampled, programs_all, qr["loss"], prior["loss"], coef_pg, coef_qr, pg_loss,
                      qr_loss):
                t.record_stream(main)
            programs_all.record_stream(s_qr)
            coef_qr.record_stream(s_qr)
            coef_pg.record_stream(s_pg)
            sampled.record_stream(s_prior)
        self._mark("backward_end")
        if os.environ.get("PNMN_DIAG_STALE_PROGRAMS") == "1" and getattr(sampled, "_pnmn_host", None) is not None:
            sampled._pnmn_host[1].synchronize()
            self._stale[nu] = sampled._pnmn_host[0].clone()

        self.elbo.last_outputs = {
            "program_generator": {"predictions": sampled, "loss": pg_loss_u,
                                  "raw_predictions": pg["raw_predictions"][:nu, :free] if "raw_predictions" in pg else None},
            "question_reconstructor": {"loss": qr_loss_u}, "nmn": nmn, "program_prior": prior}
        return {"loss": {"nmn": stats[4], "question_reconstruction_gt": qr_sup, "program_generation_gt": pg_sup},
                "elbo": {"reconstruction_likelihood": stats[0], "kl_divergence": stats[1], "elbo": stats[2],
                         "reinforce_reward": stats[3]},
                "objective": loss_objective}

    def _supervised(self, questions, programs):
        """alpha * (log q(z'|x') + log p(x'|z')) over the rows with ground-truth programs (:152-176)."""
        pg = self.program_generator(questions, programs, decoding_strategy="sampling")
        qr = self.question_reconstructor(programs, questions, decoding_strategy="sampling")
        return pg["loss"].mean(), qr["loss"].mean()

    def _issue_pending_nmn(self) -> None:
        """Backward pass, gradient average and clamp + Adam of the module network for the step that deferred them, on the
        current stream."""
        pending, self._pending_nmn = self._pending_nmn, None
        if pending is None:
            return
        loss, coef = pending
        torch.autograd.backward([loss], [coef])
        self._mark("nmn_bwd_end(deferred)")
        if self._distributed():
            self.nmn.allreduce_gradients(group=self.group)
        self.optimizer.step(only=self.nmn)
        self.nmn.zero_grad(set_to_none=True)

    def _issue_pending_qr(self, stream) -> None:
        """Backward pass, gradient average and clamp + Adam of the question reconstructor for the step that deferred them,
        on ``stream``."""
        pending, self._pending_qr = self._pending_qr, None
        if pending is None:
            return
        loss, coef = pending
        qr_m = self.question_reconstructor
        with torch.cuda.stream(stream):
            torch.autograd.backward([loss], [coef])
            self._mark("qr_bwd_end(deferred)")
            if self._distributed():
                from .dist import allreduce_gradients
                allreduce_gradients([qr_m], group=self.group)
            self.optimizer.step(only=qr_m)
            qr_m.zero_grad(set_to_none=True)

    def flush(self) -> None:
        """Issue whatever step() deferred (``defer_nmn``).  After it the models and their optimizer state are what the
        reference's step leaves behind."""
        if self._pending_qr is not None:
            dev = self.nmn.stem[0].weight.device
            with torch.cuda.device(dev):
                main = torch.cuda.current_stream(dev)
                s_qr = self._streams(dev)[1] if self.concurrent else main
                self._issue_pending_qr(s_qr)
                main.wait_stream(s_qr)
        if self._pending_nmn is not None:
            dev = self.nmn.stem[0].weight.device
            with torch.cuda.device(dev):
                if self.reserved_sms:
                    from . import _lib as L
                    prev = L.lib().pnmn_set_reserved_sms(self.reserved_sms)
                    try:
                        self._issue_pending_nmn()
                    finally:
                        L.lib().pnmn_set_reserved_sms(prev)
                else:
                    self._issue_pending_nmn()

    def _update_early(self, model) -> None:
        """clamp + Adam of one model on the current stream, right behind its backward pass (and its gradient average)"""
        if self._stepping and self.early_adam and self.order == 0:
            self.optimizer.step(only=model)
            self._updated.add(id(model))

    def _distributed(self) -> bool:
        return dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1

    def _reduce_early(self, models) -> None:
        """Data-parallel runs: average the gradients of ``models`` NOW, on the current stream (which waits for the
        collective; the host does not), instead of after the whole backward phase.  Remembered so that ``step`` does not
        reduce them again."""
        if not self._distributed():
            return
        from .dist import allreduce_gradients
        for m in models:
            if m is self.nmn:
                self.nmn.allreduce_gradients(group=self.group)
            else:
                allreduce_gradients([m], group=self.group)
            self._reduced.add(id(m))

    def allreduce_gradients(self) -> None:
        """Average whatever ``do_iteration`` has not averaged already (everything, on the unfused path)."""
        from .dist import allreduce_gradients
        if id(self.nmn) not in self._reduced:
            self.nmn.allreduce_gradients(group=self.group)
        rest = [m for m in (self.program_generator, self.question_reconstructor) if id(m) not in self._reduced]
        if rest:
            allreduce_gradients(rest, group=self.group)
        self._reduced.clear()

    def step(self, batch: Dict[str, Any]) -> Dict[str, Any]:
        """``_Trainer.step`` (trainers/_trainer.py:172-196) without the dataloader / tensorboard parts."""
        if self._pending_nmn is None and self._pending_qr is None:
            self.optimizer.zero_grad(set_to_none=True)
        else:
            # a deferred model's gradients of the previous step are still to be computed and applied (_issue_pending_*
            # clears them afterwards); the others start from zero as usual
            self.program_generator.zero_grad(set_to_none=True)
            if self._pending_qr is None:
                self.question_reconstructor.zero_grad(set_to_none=True)
            if self._pending_nmn is None:
                self.nmn.zero_grad(set_to_none=True)
        self._stepping, self._updated = True, set()
        self.optimizer.launches_last_step = 0
        try:
            return self._step(batch)
        finally:
            self._stepping = False

    def _step(self, batch: Dict[str, Any]) -> Dict[str, Any]:
        if self.reserved_sms:
            from . import _lib as L
            prev = L.lib().pnmn_set_reserved_sms(self.reserved_sms)
            try:
                out = self.do_iteration(batch)
            finally:
                L.lib().pnmn_set_reserved_sms(prev)
        else:
            out = self.do_iteration(batch)
        if self._distributed():
            self.allreduce_gradients()
            self._mark("allreduce_end")
        trained = (self.program_generator, self.question_reconstructor, self.nmn)
        if not self._updated:
            self.optimizer.step()
        else:
            for m in trained:
                if id(m) not in self._updated:
                    self.optimizer.step(only=m)
        self._mark("optimizer_end")
        self.iteration += 1
        return out
