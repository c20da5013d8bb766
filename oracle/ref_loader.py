"""ORACLE tooling — import the reference's NMN verbatim from /root/reference (this container only).

``/root/reference`` does not exist on the GPU box; nothing under ``tests/ -m gpu``, ``smoke()`` or
``bench.py`` may call this.  It is used by ``oracle/make_golden.py`` (fixture generation) and by the
CPU-side test that pins the restatement when the reference happens to be present.
"""
import importlib
import importlib.util
import inspect
import os
import sys
import textwrap

HERE = os.path.dirname(os.path.abspath(__file__))
SHIM = os.path.join(HERE, "ref_shim")
REPO = os.path.dirname(HERE)


def reference_available(root: str = "/root/reference") -> bool:
    return os.path.exists(os.path.join(root, "probnmn", "models", "nmn.py"))


def load_reference_nmn(root: str = "/root/reference"):
    """Returns the reference's ``NeuralModuleNetwork`` class (probnmn/models/nmn.py, unmodified source).

    ``probnmn/models/__init__.py`` pulls in the AllenNLP seq2seq models, so ``nmn.py`` is loaded by file
    path; ``probnmn.config`` and ``probnmn.modules.nmn_modules`` import normally from ``root``.
    ``SameModule.forward`` is re-compiled from its own source with ``/`` -> ``//`` on the index line
    (nmn_modules.py:203): the pinned torch==1.4.0 floor-divides a LongTensor by an int, torch 2.x does not
    (SURVEY.md §8c)."""
    for p in (REPO, SHIM, root):
        if p not in sys.path:
            sys.path.insert(0, p)
    mods = importlib.import_module("probnmn.modules.nmn_modules")
    if not getattr(mods.SameModule, "_pnmn_floor_div", False):   # (idempotent: the loaders below call this one too)
        src = textwrap.dedent(inspect.getsource(mods.SameModule.forward))
        assert "the_idx[0, 0, 0, 0] / size" in src, "reference SameModule.forward changed"
        ns = {}
        exec(compile(src.replace("the_idx[0, 0, 0, 0] / size", "the_idx[0, 0, 0, 0] // size"), "<SameModule.forward //>", "exec"),
             vars(mods), ns)
        mods.SameModule.forward = ns["forward"]
        mods.SameModule._pnmn_floor_div = True
    spec = importlib.util.spec_from_file_location("probnmn_reference_nmn", os.path.join(root, "probnmn", "models", "nmn.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.NeuralModuleNetwork


def load_reference_seq2seq(root: str = "/root/reference"):
    """Returns the reference's ``(ProgramGenerator, QuestionReconstructor, ProgramPrior)`` classes, imported normally
    from ``root`` -- ``probnmn/modules/seq2seq_base.py``, ``probnmn/models/{program_generator,question_reconstructor,
    program_prior}.py`` run VERBATIM -- over ``oracle/ref_shim/allennlp``, a restatement of the AllenNLP 0.9.0 classes
    those files build on (``SimpleSeq2Seq._encode / _init_decoder_state / _prepare_output_projections``,
    ``PytorchSeq2SeqWrapper`` around the real ``torch.nn.LSTM`` on packed sequences, ``DotProductAttention`` +
    ``masked_softmax``, ``Embedding``, ``add_sentence_boundary_token_ids``, ``sequence_cross_entropy_with_logits``)."""
    for p in (REPO, SHIM, root):
        if p not in sys.path:
            sys.path.insert(0, p)
    load_reference_nmn(root)   # (patches SameModule.forward for torch 2.x before probnmn.models imports nmn.py)
    pg = importlib.import_module("probnmn.models.program_generator")
    qr = importlib.import_module("probnmn.models.question_reconstructor")
    pr = importlib.import_module("probnmn.models.program_prior")
    return pg.ProgramGenerator, qr.QuestionReconstructor, pr.ProgramPrior
