"""ORACLE (test infrastructure, not product code) — CPU restatement of the reference's Neural Module
Network forward pass in plain fp32 PyTorch.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import this file; the product path (``probnmn_clevr_b200``) never does.

Parity status: PINNED.  ``oracle/make_golden.py`` imports the reference's own
``probnmn/models/nmn.py`` + ``probnmn/modules/nmn_modules.py`` verbatim from ``/root/reference``
(over a stub of the absent ``allennlp``/``yacs`` packages, ``oracle/ref_shim``), runs both on the
same seeded inputs, asserts agreement and stores the reference's outputs in ``tests/golden/``;
``tests/test_oracle.py`` re-checks this restatement against those vectors on every run.

Semantic notes:
  * ``SameModule`` follows the pinned torch==1.4.0 meaning of ``LongTensor / int`` = floor division
    (``requirements.txt:6``; nmn_modules.py:203); under torch 2.x the unpatched reference raises there
    and silently marks every ``same_*`` program invalid.
  * the reference's bare ``except`` (nmn.py:235) is restated as an explicit validity predicate.

Each function cites the reference lines it restates.  The restatement is functional (a state dict
with the reference's parameter names), not a copy of the reference's module classes.
"""
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

SKIP_TOKENS = {"@@PADDING@@", "@start@", "@end@", "@@UNKNOWN@@", "unique"}  # nmn.py:207
BINARY_TOKENS = {"intersect", "union", "less_than", "greater_than"}          # nmn.py:219-224
INVALID_LOSS = 3.33                                                            # nmn.py:259,268


def token_class(token: str) -> str:
    """nmn.py:90-111 (order of the substring tests matters)."""
    if token in SKIP_TOKENS:
        return "skip"
    if token == "scene":
        return "scene"
    if token == "intersect":
        return "and"
    if token == "union":
        return "or"
    if "equal" in token or token in ("less_than", "greater_than"):
        return "comparison"
    if "query" in token or token in ("exist", "count"):
        return "query"
    if "relate" in token:
        return "relate"
    if "same" in token:
        return "same"
    return "attention"


# Operand rounding mode.  "fp32" is the reference's arithmetic.  "tf32" rounds the two operands of every
# 128-output-channel convolution (activations and weights) to tf32 (cvt.rna, 10-bit mantissa) with a
# straight-through gradient, fp32 accumulation: the arithmetic of a tf32 tensor-core implementation (and of
# the reference itself on an Ampere-or-later GPU, where torch.backends.cudnn.allow_tf32 defaults to True).
# In "tf32" mode the stem output and the 128-channel module outputs are rounded as well, because the CUDA
# path STORES every tensor-core output tf32-rounded (a conv output is always the next conv's operand).
# Gradients of this network are ill-conditioned w.r.t. 1e-4-level forward perturbations (saturated sigmoids:
# the tf32 oracle's gradient sits 3e-2 (global L2) from the fp32 one while its logits agree to 6e-4), so
# gradient parity of the CUDA path is asserted against the "tf32" oracle; outputs against the fp32 one.
_ROUNDING = "fp32"


class operand_rounding:
    def __init__(self, mode):
        assert mode in ("fp32", "tf32")
        self.mode = mode

    def __enter__(self):
        global _ROUNDING
        self.prev, _ROUNDING = _ROUNDING, self.mode

    def __exit__(self, *a):
        global _ROUNDING
        _ROUNDING = self.prev


def _round_tf32(t):
    bits = t.detach().contiguous().view(torch.int32)
    return ((bits + 0x1000) & ~0x1FFF).view(torch.float32)


def _r(t):
    if _ROUNDING == "fp32":
        return t
    return t + (_round_tf32(t) - t.detach())


def _conv(sd, name, x, dilation=1, k3=True):
    pad = dilation if k3 else 0
    w = sd[name + ".weight"]
    if w.shape[0] == 1:  # 1-channel heads: fp32 everywhere
        return F.conv2d(x, w, sd[name + ".bias"], padding=pad, dilation=dilation)
    return F.conv2d(_r(x), _r(w), sd[name + ".bias"], padding=pad, dilation=dilation)


def attention_module(sd, tok, feats, attn):
    """nmn_modules.py:82-87"""
    x = feats * attn
    x = F.relu(_conv(sd, f"{tok}.conv1", x))
    x = F.relu(_conv(sd, f"{tok}.conv2", x))
    return torch.sigmoid(_conv(sd, f"{tok}.conv3", x, k3=False))


def query_module(sd, tok, feats, attn):
    """nmn_modules.py:119-123"""
    x = feats * attn
    x = F.relu(_conv(sd, f"{tok}.conv1", x))
    return _r(F.relu(_conv(sd, f"{tok}.conv2", x)))


def relate_module(sd, tok, feats, attn):
    """nmn_modules.py:160-168 (dilations 1,2,4,8,1)"""
    x = feats * attn
    for i, d in enumerate((1, 2, 4, 8, 1), start=1):
        x = F.relu(_conv(sd, f"{tok}.conv{i}", x, dilation=d))
    return torch.sigmoid(_conv(sd, f"{tok}.conv6", x, k3=False))


def same_module(sd, tok, feats, attn):
    """nmn_modules.py:200-208; argmax = first maximum in row-major order, row = idx // W (torch 1.4)."""
    h, w = attn.shape[2], attn.shape[3]
    flat = attn[0, 0].reshape(-1)
    idx = int((flat == flat.max()).nonzero()[0])  # first maximum in row-major order
    row, col = idx // w, idx % w
    vec = feats[:, :, row:row + 1, col:col + 1]
    x = torch.cat([feats * vec, attn], dim=1)
    return torch.sigmoid(_conv(sd, f"{tok}.conv", x, k3=False))


def comparison_module(sd, tok, in1, in2):
    """nmn_modules.py:239-244"""
    x = F.relu(_conv(sd, f"{tok}.projection", torch.cat([in1, in2], 1), k3=False))
    x = F.relu(_conv(sd, f"{tok}.conv1", x))
    return _r(F.relu(_conv(sd, f"{tok}.conv2", x)))


def run_program(sd, vocabulary, feat_input, token_ids, module_channels=128, trace=None):
    """One sample of nmn.py:198-238.  Returns (final tensor or None when invalid)."""
    out, saved = feat_input, None
    for i in reversed([int(t) for t in token_ids]):
        try:
            tok = vocabulary.get_token_from_index(i, namespace="programs")
        except KeyError:
            return None
        cls = token_class(tok)
        if cls == "skip":
            continue
        if cls == "scene":  # nmn.py:211-217
            saved = out
            out = torch.ones_like(feat_input)[:, :1]
            continue
        if cls in ("and", "or"):  # nmn_modules.py:25-27,43-45: TypeError when saved is None
            if saved is None:
                return None
            out = torch.min(out, saved) if cls == "and" else torch.max(out, saved)
        elif cls == "comparison":  # cat must give 2*C channels for the projection
            if saved is None or out.shape[1] != module_channels or saved.shape[1] != module_channels:
                return None
            out = comparison_module(sd, tok, out, saved)
        else:  # unary: attn.repeat(1,C,1,1) * feats needs a 1-channel attention
            if out.shape[1] != 1:
                return None
            fn = {"attention": attention_module, "query": query_module, "relate": relate_module,
                  "same": same_module}[cls]
            out = fn(sd, tok, feat_input, out)
        if trace is not None:
            trace.append((tok, out))
    if out.shape[1] != module_channels:  # nmn.py:231-232
        return None
    return out


def classifier(sd, x):
    """nmn.py:75-83"""
    x = F.relu(F.conv2d(x, sd["classifier.0.weight"], sd["classifier.0.bias"]))
    x = F.max_pool2d(x, kernel_size=2, stride=2)
    x = x.reshape(x.shape[0], -1)
    x = F.relu(F.linear(x, sd["classifier.4.weight"], sd["classifier.4.bias"]))
    return F.linear(x, sd["classifier.6.weight"], sd["classifier.6.bias"])


def stem(sd, features):
    """nmn.py:67-72,183"""
    x = F.relu(F.conv2d(_r(features), _r(sd["stem.0.weight"]), sd["stem.0.bias"], padding=1))
    return _r(F.relu(F.conv2d(_r(x), _r(sd["stem.2.weight"]), sd["stem.2.bias"], padding=1)))


def nmn_forward(sd: Dict[str, torch.Tensor], vocabulary, features: torch.Tensor, programs: torch.Tensor,
                answers: Optional[torch.Tensor] = None, want: Tuple[str, ...] = (),
                final_override: Optional[torch.Tensor] = None):
    """NeuralModuleNetwork.forward (nmn.py:139-275) without the metric objects.

    Returns a dict with "predictions", "loss" and additionally "logits", "valid", "final" (module
    outputs entering the classifier) and, if requested through ``want``, "traces"."""
    feat = stem(sd, features)
    B, C = feat.shape[0], feat.shape[1]
    finals, valid, traces = [], [], []
    for n in range(B):
        tr: List = [] if "traces" in want else None
        out = run_program(sd, vocabulary, feat[n:n + 1], programs[n].tolist(), C, tr)
        if out is None:
            out = torch.zeros_like(feat[n:n + 1])  # nmn.py:236
            valid.append(0)
        else:
            valid.append(1)
        finals.append(out)
        traces.append(tr)
    final = torch.cat(finals, 0)
    if final_override is not None:
        # evaluate the classifier AT the given module outputs (straight-through to this graph): lets a test
        # compare backward passes without the classifier's ReLU / max-pool kinks amplifying 1-ulp differences
        final = final + (final_override - final).detach()
    logits = classifier(sd, final)
    logprobs = F.log_softmax(logits, dim=-1)
    best_lp, pred = torch.max(logprobs, dim=1)
    valid_t = torch.tensor(valid)
    unk = vocabulary.get_token_index("@@UNKNOWN@@", namespace="answers")
    pred = pred.clone()
    pred[valid_t == 0] = unk  # nmn.py:250-253
    if answers is not None:
        loss = F.cross_entropy(logits, answers, reduction="none")
    else:
        loss = -best_lp
    # in-place overwrite in the reference (nmn.py:259,268): the constant carries no gradient
    loss = torch.where(valid_t.to(loss.device) == 0, torch.full_like(loss, INVALID_LOSS), loss)
    out = {"predictions": pred, "loss": loss, "logits": logits, "valid": valid_t, "final": final}
    if "traces" in want:
        out["traces"] = traces
    return out
