"""Drop-in ``Seq2SeqBase`` / ``ProgramGenerator`` / ``QuestionReconstructor`` (reference:
probnmn/modules/seq2seq_base.py, probnmn/models/program_generator.py, probnmn/models/question_reconstructor.py)
whose whole forward and backward run in hand-written sm_100a CUDA behind ``pnmn_pg_forward`` /
``pnmn_pg_backward`` (``include/pnmn.h``).

Same constructor arguments, ``from_config``, ``forward(source_tokens, target_tokens=None,
decoding_strategy="sampling") -> {"predictions", "loss"}``, ``get_metrics``, ``decode`` and the state-dict
keys of the AllenNLP 0.9.0 ``SimpleSeq2Seq`` the reference builds on, so
``JointTrainingTrainer`` / ``QuestionCodingTrainer``, ``CheckpointManager.load`` and Adam work unchanged.

What differs from the reference: no per-step / per-row host synchronisation (seq2seq_base.py:188,286), tokens are
chosen on the device (greedy: ``torch.max`` semantics, lowest index on ties; sampling: a counter-based Philox
stream keyed by ``torch.initial_seed()`` and a per-module call counter instead of the global generator that
``torch.multinomial`` advances), and ``"loss"`` is produced by one autograd node whose backward is the CUDA
backward pass.  There is no CPU or eager fallback.
"""
import ctypes
import math
from collections import Counter
from typing import Dict, List, Optional, Tuple

import torch
from torch import nn

from . import _lib as L
from .vocabulary import Vocabulary


class _LstmParameters(nn.Module):
    """Parameter holder with ``nn.LSTM``'s names (``weight_ih_l0`` ...) and default initialisation."""

    def __init__(self, input_size: int, hidden_size: int, num_layers: int):
        super().__init__()
        k = 1.0 / math.sqrt(hidden_size)
        for layer in range(num_layers):
            cin = input_size if layer == 0 else hidden_size
            for name, shape in (("weight_ih", (4 * hidden_size, cin)), ("weight_hh", (4 * hidden_size, hidden_size)),
                                ("bias_ih", (4 * hidden_size,)), ("bias_hh", (4 * hidden_size,))):
                self.register_parameter(f"{name}_l{layer}", nn.Parameter(torch.empty(shape).uniform_(-k, k)))


class _Holder(nn.Module):
    pass


class _Workspaces:
    """Zero-filled device scratch per (device, sizes); a workspace is busy from forward until its backward ran (or the
    graph was dropped).  The library requires zero-filled memory whenever the sizes change, hence one per key."""

    def __init__(self, limit: int = 12):
        self.free: Dict[tuple, List[torch.Tensor]] = {}
        self.order: List[tuple] = []
        self.limit = limit
        self.sizes: Dict[tuple, int] = {}

    def acquire(self, key, nbytes, device):
        lst = self.free.setdefault(key, [])
        if key in self.order:
            self.order.remove(key)
        self.order.append(key)
        if lst:
            return lst.pop()
        while len(self.order) > self.limit:
            self.free.pop(self.order.pop(0), None)
        return torch.zeros(nbytes, dtype=torch.uint8, device=device)

    def release(self, key, ws):
        if key in self.order:
            self.free.setdefault(key, []).append(ws)


_WS = _Workspaces()


class _Run:
    def __init__(self, key, ws, args):
        self.key, self.ws, self.args = key, ws, args

    def close(self):
        if self.ws is not None:
            _WS.release(self.key, self.ws)
            self.ws = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class _HostCopies:
    """Pinned staging buffers + one copy stream per device for ``_handover``."""

    def __init__(self):
        self.streams: Dict[torch.device, torch.cuda.Stream] = {}
        self.ring: Dict[torch.dtype, list] = {}   # dtype -> [items, next index]

    def stream(self, device):
        s = self.streams.get(device)
        if s is None:
            s = self.streams[device] = torch.cuda.Stream(device)
        return s

    def buffer(self, shape, dtype):
        """Round-robin over 8 flat pinned buffers per dtype (grown on demand, viewed with the requested shape): a buffer is
        overwritten only after 8 later hand-overs (two training steps), long after its reader -- which clones it -- is
        done.  Keyed by dtype, not shape: the sub-batch sizes of the joint-training step change every step and page-locked
        allocations cost milliseconds."""
        numel = 1
        for d in shape:
            numel *= int(d)
        ring = self.ring.setdefault(dtype, [[], 0])
        items, nxt = ring
        if len(items) < 8:
            item = [torch.empty(max(numel, 1 << 15), dtype=dtype).pin_memory(), None, None]
            items.append(item)
        else:
            item = items[nxt % 8]
            ring[1] = nxt + 1
            if item[1] is not None:
                item[1].synchronize()
            if item[0].numel() < numel:
                item[0] = torch.empty(numel * 2, dtype=dtype).pin_memory()
        item[2] = item[0][:numel].view(shape)
        return item


_HOST = _HostCopies()


def _handover(tensor: torch.Tensor) -> None:
    """Attach ``tensor._pnmn_host = (pinned copy, event)``: the copy runs on a side stream as soon as everything queued so
    far on the current stream (the producing forward pass) is done, independent of what the caller queues next."""
    dev = tensor.device
    current = torch.cuda.current_stream(dev)
    side = _HOST.stream(dev)
    item = _HOST.buffer(tensor.shape, tensor.dtype)
    ready = torch.cuda.Event()
    ready.record(current)
    side.wait_event(ready)
    with torch.cuda.stream(side):
        item[2].copy_(tensor, non_blocking=True)
        done = torch.cuda.Event()
        done.record(side)
    item[1] = done
    tensor._pnmn_host = (item[2], done)


class _Seq2SeqFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, run, flat, slices, loss, *params):
        ctx.run, ctx.flat, ctx.slices = run, flat, slices
        return loss.clone()

    @staticmethod
    def backward(ctx, grad_loss):
        run, flat = ctx.run, ctx.flat
        if run.ws is None:
            raise RuntimeError("seq2seq backward called twice (the workspace was already released)")
        desc, B, Tq, Tp, S, teacher = run.args
        gflat = torch.zeros_like(flat)
        grad_loss = grad_loss.contiguous().float()
        with torch.cuda.device(flat.device):
            stream = ctypes.c_void_p(torch.cuda.current_stream(flat.device).cuda_stream)
            L.check(L.lib().pnmn_pg_backward(ctypes.byref(desc), ctypes.c_void_p(flat.data_ptr()),
                                             ctypes.c_void_p(gflat.data_ptr()), ctypes.c_void_p(grad_loss.data_ptr()), B, Tq, Tp,
                                             S, teacher, ctypes.c_void_p(run.ws.data_ptr()), stream), "pnmn_pg_backward")
        run.close()
        grads = tuple(gflat[o:o + n].view(shape) for (o, n, shape) in ctx.slices)
        return (None, None, None, None) + grads


class Seq2SeqBase(nn.Module):
    r"""
    LSTM encoder-decoder with dot-product attention, greedy / categorical decoding and per-sequence losses
    (reference: probnmn/modules/seq2seq_base.py:16-375).

    Parameters
    ----------
    vocabulary: object with AllenNLP's ``Vocabulary`` lookup methods
    source_namespace, target_namespace: str
    input_size, hidden_size: int, optional (default = 256) -- the CUDA kernels are built for 256 / 256
    num_layers: int, optional (default = 2)
    dropout: float, optional (default = 0.0) -- only 0.0 is supported (all reference configs use 0.0)
    max_decoding_steps: int, optional (default = 30)
    """

    _instances = 0

    def __init__(self, vocabulary, source_namespace: str, target_namespace: str, input_size: int = 256,
                 hidden_size: int = 256, num_layers: int = 2, dropout: float = 0.0, max_decoding_steps: int = 30):
        super().__init__()
        if input_size != 256 or hidden_size != 256 or num_layers != 2:
            raise ValueError("the B200 seq2seq kernels are built for input_size = hidden_size = 256 and num_layers = 2")
        if dropout != 0.0:
            raise ValueError("dropout != 0 is not supported by the B200 seq2seq kernels (the reference configs use 0.0)")
        self.vocab = vocabulary
        self._source_namespace, self._target_namespace = source_namespace, target_namespace
        self._pad_index = vocabulary.get_token_index("@@PADDING@@", namespace=source_namespace)
        self._unk_index = vocabulary.get_token_index("@@UNKNOWN@@", namespace=source_namespace)
        self._end_index = vocabulary.get_token_index("@end@", namespace=source_namespace)
        self._start_index = vocabulary.get_token_index("@start@", namespace=source_namespace)
        if (self._pad_index, self._unk_index, self._start_index, self._end_index) != (0, 1, 2, 3):
            raise ValueError("special tokens must sit at indices 0..3 (build_vocabulary.py:114-119)")
        vs = vocabulary.get_vocab_size(namespace=source_namespace)
        vt = vocabulary.get_vocab_size(namespace=target_namespace)
        if max(vs, vt) > 128:
            raise ValueError("vocabularies of more than 128 entries are not supported by the B200 seq2seq kernels")
        self._vs, self._vt, self._hidden = vs, vt, hidden_size
        self._max_decoding_steps = max_decoding_steps
        self._scheduled_sampling_ratio = 0.0

        # AllenNLP SimpleSeq2Seq's sub-modules, by name (state-dict contract)
        self._source_embedder = _Holder()
        self._source_embedder.token_embedder_tokens = nn.Embedding(vs, input_size, padding_idx=self._pad_index)
        nn.init.xavier_uniform_(self._source_embedder.token_embedder_tokens.weight)
        with torch.no_grad():
            self._source_embedder.token_embedder_tokens.weight[self._pad_index].zero_()
        self._encoder = _Holder()
        self._encoder._module = _LstmParameters(input_size, hidden_size, num_layers)
        self._target_embedder = nn.Embedding(vt, input_size)
        nn.init.xavier_uniform_(self._target_embedder.weight)
        self._decoder_cell = nn.LSTMCell(input_size + hidden_size, hidden_size)
        self._output_projection_layer = nn.Linear(hidden_size, vt)

        self._flat: Optional[torch.Tensor] = None
        self._layout: Optional[List[Tuple[str, int, int, torch.Size]]] = None
        self._desc = None
        self._calls = 0
        self._teacher_calls = 0
        # per-instance salt of the sampling stream: construction order within the process (deterministic for a given script,
        # unlike id(self)), so that two runs with the same torch.manual_seed draw the same samples
        Seq2SeqBase._instances += 1
        self._salt = Seq2SeqBase._instances
        self.return_logits = False   # tests: also return "logits" (B, steps, V) and "raw_predictions"
        # start an asynchronous device -> pinned-host copy of "predictions" on a side stream right behind the forward pass:
        # a NeuralModuleNetwork that is handed this tensor (modules/elbo.py:233-239, joint_training_evaluator.py:98-103)
        # compiles the programs on the host and would otherwise synchronise the whole compute stream to read them
        self.handover_predictions = True
        self._pending_metrics: list = []
        self._metrics = {"loss_sum": 0.0, "loss_n": 0, "seq_correct": 0, "seq_n": 0, "recall_sum": 0.0, "recall_n": 0,
                         "bleu_match": Counter(), "bleu_total": Counter(), "bleu_pred_len": 0, "bleu_ref_len": 0}

    # ---- flat parameter buffer -------------------------------------------------------------------------------------
    def _ensure_flat(self):
        named = list(self.named_parameters())
        dev = named[0][1].device
        ok = self._flat is not None and self._flat.device == dev
        if ok:
            base = self._flat.data_ptr()
            for (name, off, n, _), (_, p) in zip(self._layout, named):
                if p.data_ptr() != base + 4 * off or p.device != dev or not p.is_contiguous():
                    ok = False
                    break
        if ok:
            return
        layout, off = [], 0
        for name, p in named:
            layout.append((name, off, p.numel(), p.shape))
            off += (p.numel() + 63) // 64 * 64
        flat = torch.zeros(off, dtype=torch.float32, device=dev)
        for (name, o, n, shape), (_, p) in zip(layout, named):
            flat[o:o + n].copy_(p.data.reshape(-1))
            p.data = flat[o:o + n].view(shape)
        self._flat, self._layout = flat, layout
        offs = {name: o for name, o, _, _ in layout}
        d = L.PgDesc()
        d.vocab_src, d.vocab_tgt, d.hidden, d.num_layers = self._vs, self._vt, self._hidden, 2
        d.src_embed = offs["_source_embedder.token_embedder_tokens.weight"]
        for layer in range(2):
            d.enc_w_ih[layer] = offs[f"_encoder._module.weight_ih_l{layer}"]
            d.enc_w_hh[layer] = offs[f"_encoder._module.weight_hh_l{layer}"]
            d.enc_b_ih[layer] = offs[f"_encoder._module.bias_ih_l{layer}"]
            d.enc_b_hh[layer] = offs[f"_encoder._module.bias_hh_l{layer}"]
        d.tgt_embed = offs["_target_embedder.weight"]
        d.dec_w_ih, d.dec_w_hh = offs["_decoder_cell.weight_ih"], offs["_decoder_cell.weight_hh"]
        d.dec_b_ih, d.dec_b_hh = offs["_decoder_cell.bias_ih"], offs["_decoder_cell.bias_hh"]
        d.out_w, d.out_b = offs["_output_projection_layer.weight"], offs["_output_projection_layer.bias"]
        self._desc = d

    # ---- forward ---------------------------------------------------------------------------------------------------
    def forward(self, source_tokens: torch.Tensor, target_tokens: Optional[torch.Tensor] = None,
                decoding_strategy: str = "sampling") -> Dict[str, torch.Tensor]:
        r"""
        Same contract as the reference (seq2seq_base.py:101-155): ``source_tokens`` (B, T_src) and optional
        ``target_tokens`` (B, T_tgt) are zero-padded and carry NO ``@start@`` / ``@end@``; returns ``predictions``
        (B, steps) trimmed after the first ``@end@`` and ``loss`` (B,): teacher-forced sequence cross entropy when
        targets are given, else the negated length-normalised log-probability of the decoded tokens.
        """
        if decoding_strategy not in ("sampling", "greedy"):
            raise ValueError(f"decoding_strategy must be 'sampling' or 'greedy', got {decoding_strategy!r}")
        if not source_tokens.is_cuda:
            raise RuntimeError("Seq2SeqBase (B200) needs CUDA tensors; there is no CPU fallback")
        with torch.cuda.device(source_tokens.device):  # the library launches on the thread's current device
            return self._forward(source_tokens, target_tokens, decoding_strategy)

    def forward_mixed(self, source_tokens: torch.Tensor, target_tokens: torch.Tensor, teacher_rows: torch.Tensor,
                      free_steps: Optional[int] = None) -> Dict[str, torch.Tensor]:
        r"""
        ONE pass over rows of both kinds (no counterpart in the reference, which calls the model once per kind --
        trainers/joint_training_trainer.py:139-144,164-168): rows with ``teacher_rows[b]`` true are teacher-forced on
        ``target_tokens[b]`` exactly as ``forward(source, target)`` would, the others decode freely by categorical
        sampling for ``free_steps`` (default ``max_decoding_steps``) steps exactly as ``forward(source)`` would, their
        target row is ignored.  Returns ``predictions`` (B, T_tgt + 1) -- columns beyond ``free_steps`` of a free-running
        row are padding -- and ``loss`` (B,).  Per row the values equal those of the separate calls (rows are
        independent); the sampled tokens differ because the Philox stream is indexed by the row's position in the call.
        """
        if not source_tokens.is_cuda:
            raise RuntimeError("Seq2SeqBase (B200) needs CUDA tensors; there is no CPU fallback")
        free_steps = self._max_decoding_steps if free_steps is None else int(free_steps)
        if target_tokens.shape[1] + 1 < free_steps:
            raise ValueError("target_tokens must have at least free_steps - 1 columns (pad with zeros)")
        with torch.cuda.device(source_tokens.device):
            return self._forward(source_tokens, target_tokens, "sampling",
                                 teacher_rows.to(source_tokens.device, torch.uint8).contiguous(), free_steps)

    def _forward(self, source_tokens, target_tokens, decoding_strategy, teacher_rows=None, free_steps=0):
        lib = L.lib()
        self._ensure_flat()
        dev = source_tokens.device
        source = source_tokens.detach().to(torch.int64).contiguous()
        B, Tq = source.shape
        teacher = target_tokens is not None
        if teacher:
            target = target_tokens.detach().to(dev, torch.int64).contiguous()
            Tp = target.shape[1]
            S = Tp + 1                                      # seq2seq_base.py:168-175
        else:
            target, Tp, S = None, 0, self._max_decoding_steps  # :177
        need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        # the workspace layout depends on the batch size only through its multiple of 128 (csrc/seq2seq_api.cu), so the
        # varying sub-batch sizes of the joint-training step (supervised / unsupervised split) share zero-filled buffers
        Bp = (B + 127) // 128 * 128
        key = (dev, Bp, Tq, Tp, S, need_grad, self._vs, self._vt)
        nbytes = _WS.sizes.get(key)
        if nbytes is None:
            nbytes = lib.pnmn_pg_workspace_bytes(ctypes.byref(self._desc), Bp, Tq, Tp, S, 1 if need_grad else 0)
            if nbytes < 0:
                raise RuntimeError("pnmn_pg_workspace_bytes failed: " + lib.pnmn_last_error().decode())
            _WS.sizes[key] = nbytes
        ws = _WS.acquire(key, nbytes, dev)
        run = _Run(key, ws, (self._desc, B, Tq, Tp, S, 1 if teacher else 0))

        raw = torch.empty(B, S, dtype=torch.int64, device=dev)
        predictions = torch.empty(B, S, dtype=torch.int64, device=dev)
        loss = torch.empty(B, dtype=torch.float32, device=dev)
        logits = torch.empty(B, S, self._vt, dtype=torch.float32, device=dev) if self.return_logits else None
        # Philox key of this call: (torch.manual_seed value, construction-order salt of the module, per-module call counter).
        # Free-running and teacher-forced calls count separately (the draws of a teacher-forced call only decide its
        # "predictions", never its loss), so the samples of the free-running pass do not depend on whether the supervised
        # pass of the same step was issued before or after it.  The counters advance on EVERY forward (validation passes
        # included), as the global generator does in the reference.
        if teacher:
            self._teacher_calls += 1
            counter = 2 * self._teacher_calls + 1
        else:
            self._calls += 1
            counter = 2 * self._calls
        seed = (torch.initial_seed() * 0x9E3779B97F4A7C15 + counter * 0xD1B54A32D192ED03 + self._salt * 0x2545F4914F6CDD1D) % (1 << 64)
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        if teacher_rows is not None:
            L.check(lib.pnmn_pg_forward_mixed(
                ctypes.byref(self._desc), ctypes.c_void_p(self._flat.data_ptr()), ctypes.c_void_p(source.data_ptr()),
                ctypes.c_void_p(target.data_ptr()), ctypes.c_void_p(teacher_rows.data_ptr()), free_steps, B, Tq, Tp,
                ctypes.c_uint64(seed), 1 if need_grad else 0, ctypes.c_void_p(ws.data_ptr()), ctypes.c_void_p(raw.data_ptr()),
                ctypes.c_void_p(predictions.data_ptr()), ctypes.c_void_p(loss.data_ptr()),
                ctypes.c_void_p(logits.data_ptr()) if logits is not None else None, stream), "pnmn_pg_forward_mixed")
        else:
            L.check(lib.pnmn_pg_forward(
                ctypes.byref(self._desc), ctypes.c_void_p(self._flat.data_ptr()), ctypes.c_void_p(source.data_ptr()),
                ctypes.c_void_p(target.data_ptr()) if teacher else None, B, Tq, Tp, S,
                1 if decoding_strategy == "sampling" else 0, ctypes.c_uint64(seed), 1 if need_grad else 0,
                ctypes.c_void_p(ws.data_ptr()), ctypes.c_void_p(raw.data_ptr()), ctypes.c_void_p(predictions.data_ptr()),
                ctypes.c_void_p(loss.data_ptr()), ctypes.c_void_p(logits.data_ptr()) if logits is not None else None, stream),
                "pnmn_pg_forward")
        if need_grad:
            params = [p for _, p in self.named_parameters()]
            slices = [(o, n, shape) for _, o, n, shape in self._layout]
            loss = _Seq2SeqFn.apply(run, self._flat, slices, loss, *params)
        else:
            run.close()

        if self.handover_predictions and teacher_rows is None:
            _handover(predictions)
        output_dict = {"predictions": predictions, "loss": loss}
        if self.return_logits:
            output_dict["logits"], output_dict["raw_predictions"] = logits, raw
        if teacher and not self.training and teacher_rows is None:
            # seq2seq_base.py:258-274.  The statistics are host-side (n-gram counts); the batch is only queued here and
            # processed when a metric is read, so that a validation loop does not synchronise once per batch
            self._pending_metrics.append((predictions.detach(), target, loss.detach()))
        return output_dict

    # ---- metrics (validation only; host side) -------------------------------------------------------------------------
    def _record_metrics(self, predictions, target, loss):
        m = self._metrics
        m["loss_sum"] += float(loss.detach().mean())
        m["loss_n"] += 1
        B, Tp = target.shape
        # relevant targets = boundary-added targets without @start@: p_1..p_m @end@ 0...
        rel = torch.zeros(B, Tp + 1, dtype=torch.int64)
        tgt = target.cpu()
        rel[:, :Tp] = tgt
        lengths = (tgt != self._pad_index).sum(1)
        rel[torch.arange(B), lengths] = self._end_index
        pred = predictions.detach().cpu()[:, : Tp + 1]
        mask = rel != self._pad_index
        m["seq_correct"] += int(((pred == rel) | ~mask).all(1).sum())
        m["seq_n"] += B
        exclude = {self._pad_index, self._end_index, self._start_index}
        for p_row, r_row, k_row in zip(pred.tolist(), rel.tolist(), mask.tolist()):
            gold = [t for t, k in zip(r_row, k_row) if k]
            hyp = set(p_row)
            m["recall_sum"] += sum(1 for t in gold if t in hyp) / max(len(gold), 1)
            m["recall_n"] += 1
            # BLEU statistics (allennlp.training.metrics.BLEU: n-grams 1..4, specials excluded)
            h = [t for t in p_row if t not in exclude]
            r = [t for t in r_row if t not in exclude]
            m["bleu_pred_len"] += len(h)
            m["bleu_ref_len"] += len(r)
            for n in range(1, 5):
                hc = Counter(tuple(h[i:i + n]) for i in range(len(h) - n + 1))
                rc = Counter(tuple(r[i:i + n]) for i in range(len(r) - n + 1))
                m["bleu_match"][n] += sum(min(c, rc[g]) for g, c in hc.items())
                m["bleu_total"][n] += sum(hc.values())

    def get_metrics(self, reset: bool = True) -> Dict[str, float]:
        """``{"BLEU", "perplexity", "sequence_accuracy", "word_error_rate"}`` in evaluation mode, ``{}`` while training
        (seq2seq_base.py:343-375)."""
        out: Dict[str, float] = {}
        pending, self._pending_metrics = self._pending_metrics, []
        for predictions, target, loss in pending:
            self._record_metrics(predictions, target, loss)
        if not self.training:
            m = self._metrics
            if m["bleu_pred_len"] == 0 or any(m["bleu_match"][n] == 0 for n in range(1, 5)):
                bleu = 0.0
            else:
                logp = sum(0.25 * math.log(m["bleu_match"][n] / m["bleu_total"][n]) for n in range(1, 5))
                bp = 1.0 if m["bleu_pred_len"] > m["bleu_ref_len"] else math.exp(1.0 - m["bleu_ref_len"] / m["bleu_pred_len"])
                bleu = bp * math.exp(logp)
            avg = m["loss_sum"] / m["loss_n"] if m["loss_n"] else 0.0
            out = {"BLEU": bleu, "perplexity": 2 ** avg,
                   "sequence_accuracy": m["seq_correct"] / m["seq_n"] if m["seq_n"] else 0.0,
                   "word_error_rate": 1 - (m["recall_sum"] / m["recall_n"] if m["recall_n"] else 0.0)}
            if reset:
                self._metrics = {"loss_sum": 0.0, "loss_n": 0, "seq_correct": 0, "seq_n": 0, "recall_sum": 0.0,
                                 "recall_n": 0, "bleu_match": Counter(), "bleu_total": Counter(), "bleu_pred_len": 0,
                                 "bleu_ref_len": 0}
        return out

    def decode(self, output_dict: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        """AllenNLP ``SimpleSeq2Seq.decode``: indices -> token strings up to (excluding) the first ``@end@``."""
        rows = output_dict["predictions"]
        rows = rows.detach().cpu().tolist() if isinstance(rows, torch.Tensor) else rows
        all_tokens = []
        for row in rows:
            if self._end_index in row:
                row = row[: row.index(self._end_index)]
            all_tokens.append([self.vocab.get_token_from_index(i, namespace=self._target_namespace) for i in row])
        output_dict["predicted_tokens"] = all_tokens
        return output_dict


class ProgramGenerator(Seq2SeqBase):
    r"""Questions -> programs (reference: probnmn/models/program_generator.py:10-59); ``max_decoding_steps = 26``, the
    longest program of the CLEVR v1.0 train split."""

    def __init__(self, vocabulary, input_size: int = 256, hidden_size: int = 256, num_layers: int = 2, dropout: float = 0.0):
        super().__init__(vocabulary, source_namespace="questions", target_namespace="programs", input_size=input_size,
                         hidden_size=hidden_size, num_layers=num_layers, dropout=dropout, max_decoding_steps=26)

    @classmethod
    def from_config(cls, config):
        _C = config
        return cls(vocabulary=Vocabulary.from_files(_C.DATA.VOCABULARY), input_size=_C.PROGRAM_GENERATOR.INPUT_SIZE,
                   hidden_size=_C.PROGRAM_GENERATOR.HIDDEN_SIZE, num_layers=_C.PROGRAM_GENERATOR.NUM_LAYERS,
                   dropout=_C.PROGRAM_GENERATOR.DROPOUT)


class QuestionReconstructor(Seq2SeqBase):
    r"""Programs -> questions (reference: probnmn/models/question_reconstructor.py:11-61); ``max_decoding_steps = 45``.
    Same kernels with the namespaces swapped."""

    def __init__(self, vocabulary, input_size: int = 256, hidden_size: int = 256, num_layers: int = 2, dropout: float = 0.0):
        super().__init__(vocabulary, source_namespace="programs", target_namespace="questions", input_size=input_size,
                         hidden_size=hidden_size, num_layers=num_layers, dropout=dropout, max_decoding_steps=45)

    @classmethod
    def from_config(cls, config):
        _C = config
        return cls(vocabulary=Vocabulary.from_files(_C.DATA.VOCABULARY), input_size=_C.QUESTION_RECONSTRUCTOR.INPUT_SIZE,
                   hidden_size=_C.QUESTION_RECONSTRUCTOR.HIDDEN_SIZE, num_layers=_C.QUESTION_RECONSTRUCTOR.NUM_LAYERS,
                   dropout=_C.QUESTION_RECONSTRUCTOR.DROPOUT)
