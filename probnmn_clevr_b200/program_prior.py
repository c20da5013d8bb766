"""Drop-in ``ProgramPrior`` (reference: probnmn/models/program_prior.py) whose forward pass runs in hand-written sm_100a
CUDA behind ``pnmn_prior_forward`` (``include/pnmn.h``; ``csrc/prior.cu`` on top of the tcgen05 LSTM step kernels of the
seq2seq path).

Same constructor arguments, ``from_config``, ``forward(program_tokens) -> {"predictions", "loss"}``, ``get_metrics`` and
state-dict keys as the reference (AllenNLP names: ``_embedder.token_embedder_programs.weight``,
``_encoder._module.{weight,bias}_{ih,hh}_l{0,1}``, ``_projection_layer.weight``, ``_output_layer.weight`` tied to the
embedding), so ``CheckpointManager(program_prior=...).load`` (trainers/joint_training_trainer.py:107-112) works unchanged.

Scope: the FORWARD pass, which is what the joint-training and question-coding steps use -- the
prior is frozen there (``.eval()``, not in the optimizer) and its loss only enters the detached REINFORCE reward
(modules/elbo.py:151,256).  ``"loss"`` therefore carries no autograd graph; training the prior itself (the reference's
``program_prior`` phase) and ``sample()`` are outside the hot path and not provided.  There is no CPU or eager fallback.
"""
import ctypes
from typing import Dict, List, Optional, Tuple

import torch
from torch import nn

from . import _lib as L
from .seq2seq import _Holder, _LstmParameters
from .vocabulary import Vocabulary


class ProgramPrior(nn.Module):
    r"""
    A language model over CLEVR program sequences (reference: probnmn/models/program_prior.py:16-78).

    Parameters
    ----------
    vocabulary: object with AllenNLP's ``Vocabulary`` lookup methods, namespace "programs"
    input_size, hidden_size: int -- the CUDA kernels are built for 256 / 256 (every reference config uses these;
        note the reference's own keyword default for ``hidden_size`` is 128)
    num_layers: int, optional (default = 2)
    dropout: float, optional (default = 0.0) -- only 0.0 is supported
    """

    _instances = 0

    def __init__(self, vocabulary, input_size: int = 256, hidden_size: int = 256, num_layers: int = 2,
                 dropout: float = 0.0):
        super().__init__()
        if input_size != 256 or hidden_size != 256 or num_layers != 2:
            raise ValueError("the B200 LSTM kernels are built for input_size = hidden_size = 256 and num_layers = 2")
        if dropout != 0.0:
            raise ValueError("dropout != 0 is not supported (the reference configs use 0.0)")
        self._start_index = vocabulary.get_token_index("@start@", namespace="programs")
        self._end_index = vocabulary.get_token_index("@end@", namespace="programs")
        self._pad_index = vocabulary.get_token_index("@@PADDING@@", namespace="programs")
        self._unk_index = vocabulary.get_token_index("@@UNKNOWN@@", namespace="programs")
        if (self._pad_index, self._unk_index, self._start_index, self._end_index) != (0, 1, 2, 3):
            raise ValueError("special tokens must sit at indices 0..3 (build_vocabulary.py:114-119)")
        vocab_size = vocabulary.get_vocab_size(namespace="programs")
        if vocab_size > 128:
            raise ValueError("vocabularies of more than 128 entries are not supported by the B200 LSTM kernels")
        self._vocab_size = vocab_size

        # AllenNLP sub-module names (state-dict contract)
        self._embedder = _Holder()
        self._embedder.token_embedder_programs = nn.Embedding(vocab_size, input_size, padding_idx=self._pad_index)
        nn.init.xavier_uniform_(self._embedder.token_embedder_programs.weight)
        with torch.no_grad():
            self._embedder.token_embedder_programs.weight[self._pad_index].zero_()
        self._encoder = _Holder()
        self._encoder._module = _LstmParameters(input_size, hidden_size, num_layers)
        # project and tie input and output embeddings (program_prior.py:59-62)
        self._projection_layer = nn.Linear(hidden_size, input_size, bias=False)
        self._output_layer = nn.Linear(input_size, vocab_size, bias=False)
        self._output_layer.weight = self._embedder.token_embedder_programs.weight

        self._flat: Optional[torch.Tensor] = None
        self._layout: Optional[List[Tuple[str, int, int, torch.Size]]] = None
        self._desc = None
        self._ws: Dict[tuple, torch.Tensor] = {}
        self._calls = 0
        ProgramPrior._instances += 1
        self._salt = ProgramPrior._instances
        self.return_logits = False     # tests: also return "logits" (B, T + 1, V)
        self._log2_perplexity_total, self._log2_perplexity_count, self._pending = 0.0, 0, []

    @classmethod
    def from_config(cls, config):
        _C = config
        return cls(vocabulary=Vocabulary.from_files(_C.DATA.VOCABULARY), input_size=_C.PROGRAM_PRIOR.INPUT_SIZE,
                   hidden_size=_C.PROGRAM_PRIOR.HIDDEN_SIZE, num_layers=_C.PROGRAM_PRIOR.NUM_LAYERS,
                   dropout=_C.PROGRAM_PRIOR.DROPOUT)

    def _ensure_flat(self):
        named = list(self.named_parameters())   # the tied weight appears once
        dev = named[0][1].device
        if self._flat is not None and self._flat.device == dev:
            base = self._flat.data_ptr()
            if all(p.data_ptr() == base + 4 * off for (_, off, _, _), (_, p) in zip(self._layout, named)):
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
        d = L.PriorDesc()
        d.vocab, d.hidden, d.num_layers = self._vocab_size, 256, 2
        d.embed = offs["_embedder.token_embedder_programs.weight"]
        for layer in range(2):
            d.w_ih[layer] = offs[f"_encoder._module.weight_ih_l{layer}"]
            d.w_hh[layer] = offs[f"_encoder._module.weight_hh_l{layer}"]
            d.b_ih[layer] = offs[f"_encoder._module.bias_ih_l{layer}"]
            d.b_hh[layer] = offs[f"_encoder._module.bias_hh_l{layer}"]
        d.proj = offs["_projection_layer.weight"]
        self._desc = d
        self._ws = {}

    def forward(self, program_tokens: torch.Tensor) -> Dict[str, torch.Tensor]:
        r"""
        ``program_tokens`` (B, T): zero-padded program sequences without boundary tokens.  Returns ``predictions``
        (B, T + 1) -- one categorical draw per position from the model's next-token distribution, masked
        (program_prior.py:119-139) -- and ``loss`` (B,): the teacher-forced sequence cross entropy of
        ``@start@ p_1..p_m @end@`` (:141-147).
        """
        if not program_tokens.is_cuda:
            raise RuntimeError("ProgramPrior (B200) needs CUDA tensors; there is no CPU fallback")
        with torch.cuda.device(program_tokens.device):
            return self._forward(program_tokens)

    def _forward(self, program_tokens):
        lib = L.lib()
        self._ensure_flat()
        dev = program_tokens.device
        tokens = program_tokens.detach().to(torch.int64).contiguous()
        B, T = tokens.shape
        Bp = (B + 127) // 128 * 128
        key = (dev, Bp, T)
        ws = self._ws.get(key)
        if ws is None:
            nbytes = lib.pnmn_prior_workspace_bytes(ctypes.byref(self._desc), Bp, T)
            if nbytes < 0:
                raise RuntimeError("pnmn_prior_workspace_bytes failed: " + lib.pnmn_last_error().decode())
            if len(self._ws) > 8:
                self._ws.clear()
            ws = self._ws[key] = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
        predictions = torch.empty(B, T + 1, dtype=torch.int64, device=dev)
        loss = torch.empty(B, dtype=torch.float32, device=dev)
        logits = torch.empty(B, T + 1, self._vocab_size, dtype=torch.float32, device=dev) if self.return_logits else None
        self._calls += 1
        seed = (torch.initial_seed() * 0x9E3779B97F4A7C15 + self._calls * 0xD1B54A32D192ED03
                + self._salt * 0x94D049BB133111EB) % (1 << 64)
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        L.check(lib.pnmn_prior_forward(ctypes.byref(self._desc), ctypes.c_void_p(self._flat.data_ptr()),
                                       ctypes.c_void_p(tokens.data_ptr()), B, T, ctypes.c_uint64(seed),
                                       ctypes.c_void_p(ws.data_ptr()), ctypes.c_void_p(predictions.data_ptr()),
                                       ctypes.c_void_p(loss.data_ptr()),
                                       ctypes.c_void_p(logits.data_ptr()) if logits is not None else None, stream),
                "pnmn_prior_forward")
        if not self.training:
            # the reference records sequence_cross_entropy.mean().item() here (:149-150): kept on the device until read
            self._pending.append(loss.mean())
        out = {"predictions": predictions, "loss": loss}
        if logits is not None:
            out["logits"] = logits
        return out

    def get_metrics(self, reset: bool = True) -> Dict[str, float]:
        """``{"perplexity"}`` = 2 ** (average of the per-batch mean losses seen in evaluation mode) (:152-169)."""
        total = self._log2_perplexity_total + sum(float(t.item()) for t in self._pending)
        count = self._log2_perplexity_count + len(self._pending)
        if reset:
            self._log2_perplexity_total, self._log2_perplexity_count, self._pending = 0.0, 0, []
        else:
            self._log2_perplexity_total, self._log2_perplexity_count, self._pending = total, count, []
        return {"perplexity": 2 ** (total / count if count else 0.0)}

    def sample(self, num_samples: int = 1, max_sequence_length: int = 28):
        raise NotImplementedError("ProgramPrior.sample (program_prior.py:171-280) is an inspection utility outside the "
                                  "B200 hot path; load the checkpoint into the reference model to use it")
