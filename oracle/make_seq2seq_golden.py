"""ORACLE tooling -- generate tests/golden/seq2seq_golden.npz from the reference's OWN seq2seq models (run in the build
container, where /root/reference exists):

    python oracle/make_seq2seq_golden.py

``probnmn/modules/seq2seq_base.py`` and ``probnmn/models/{program_generator,question_reconstructor,program_prior}.py``
are imported VERBATIM (oracle/ref_loader.load_reference_seq2seq).  What they build on -- ``allennlp==0.9.0``
(requirements.txt:1), absent and not installable here -- is supplied by ``oracle/ref_shim/allennlp``: the LSTM arithmetic
is torch's own ``nn.LSTM`` on packed sequences and ``nn.LSTMCell``; ``SimpleSeq2Seq._encode / _init_decoder_state /
_prepare_output_projections``, ``DotProductAttention`` + ``masked_softmax``, ``add_sentence_boundary_token_ids`` and
``sequence_cross_entropy_with_logits`` are restated from AllenNLP 0.9.0's published source.  So these vectors pin
everything that lives in the reference repository (boundary handling, the decoding loop, the sampling mask, trimming,
both losses, the prior's forward) and torch's LSTM; the ~60 restated AllenNLP lines remain a restatement.

Weights and inputs are NOT stored (regenerated from their seeds: probnmn_clevr_b200.synthetic).  The restatement
oracle/seq2seq_oracle.py / oracle/prior_oracle.py must agree with the reference on every case before the file is written.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, REPO)

from oracle import prior_oracle, seq2seq_oracle  # noqa: E402
from oracle.ref_loader import load_reference_seq2seq  # noqa: E402
from probnmn_clevr_b200.synthetic import (ProgramSampler, make_prior_state_dict, make_questions,  # noqa: E402
                                          make_seq2seq_state_dict, questions_for_programs)
from probnmn_clevr_b200.vocabulary import Vocabulary  # noqa: E402

BIG_SUBSAMPLE = 997   # stride of the stored slice of large gradient tensors
END = 3

# (case name, model, weight seed, output-projection gain, input seed, rows)
CASES = [
    ("pg", "pg", 0, 1.0, 1, 12),
    ("pg_sharp", "pg", 3, 4.0, 5, 12),      # realistic decision margins: free-running tokens are well separated
    ("qr", "qr", 1, 1.0, 2, 10),
    ("qr_sharp", "qr", 2, 4.0, 3, 10),
]


def case_inputs(vocab, model, seed, rows):
    """(source tokens, target tokens) incl. the edge rows the reference handles: an empty source, an empty target,
    a full-length source / target."""
    programs = ProgramSampler(vocab, seed=seed).sample(rows, 26)
    questions = make_questions(rows, vocab.get_vocab_size("questions"), seed=seed, max_length=40)
    questions[0] = 0                                        # empty question: the encoder sees a lone @end@
    programs[1] = 0                                         # empty program
    questions[2] = torch.randint(4, 93, (40,), generator=torch.Generator().manual_seed(seed))   # full length
    return (questions, programs) if model == "pg" else (programs, questions)


def grads_record(golden, prefix, named_grads):
    for k, g in named_grads.items():
        if g.numel() <= 4096:
            golden[f"{prefix}.grad.{k}"] = g.numpy()
        else:
            golden[f"{prefix}.gradsub.{k}"] = g.reshape(-1)[::BIG_SUBSAMPLE].numpy().copy()
        golden[f"{prefix}.gradnorm.{k}"] = np.float64(g.double().norm().item())


def reference_grads(model, loss, weights):
    model.zero_grad()
    (loss * weights).sum().backward()
    return {k: (p.grad.detach().clone() if p.grad is not None else torch.zeros_like(p)) for k, p in model.named_parameters()}


def oracle_grads(sd, loss, weights):
    for p in sd.values():
        p.grad = None
    (loss * weights).sum().backward()
    return {k: (p.grad.detach().clone() if p.grad is not None else torch.zeros_like(p)) for k, p in sd.items()}


def check(name, a, b, tol=1e-6):
    err = float((a - b).abs().max() / (b.abs().max() + 1e-12)) if a.numel() else 0.0
    assert err <= tol, (name, err)
    return err


def main():
    torch.set_num_threads(os.cpu_count())
    vocab = Vocabulary.clevr()
    RefPG, RefQR, RefPrior = load_reference_seq2seq()
    golden = {}
    for name, kind, wseed, gain, iseed, rows in CASES:
        cls = RefPG if kind == "pg" else RefQR
        vs = vocab.get_vocab_size("questions" if kind == "pg" else "programs")
        vt = vocab.get_vocab_size("programs" if kind == "pg" else "questions")
        steps_free = 26 if kind == "pg" else 45
        sd = make_seq2seq_state_dict(vs, vt, seed=wseed, gain=gain)
        ref = cls(vocab)
        print(name, "state dict:", ref.load_state_dict(sd, strict=True))
        sd_req = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
        src, tgt = case_inputs(vocab, kind, iseed, rows)
        w = torch.linspace(0.5, 1.5, rows)   # non-uniform upstream gradient
        logits_box = []
        hook = ref._output_projection_layer.register_forward_hook(lambda m, i, o: logits_box.append(o.detach()))
        raw_box = []
        trim = ref._trim_predictions
        ref._trim_predictions = lambda p: (raw_box.append(p.detach().clone()), trim(p))[1]   # records its input only

        # ---- teacher forcing, training mode (seq2seq_base.py:101-160,247-254,333-341) ----
        ref.train()
        torch.manual_seed(100 + iseed)
        logits_box.clear()
        out = ref(src, tgt, decoding_strategy="sampling")
        logits = torch.stack(logits_box, 1)
        g_ref = reference_grads(ref, out["loss"], w)
        mine = seq2seq_oracle.seq2seq_forward(sd_req, src, tgt, "greedy")
        e1 = check(name + " tf logits", mine["logits"].detach(), logits)
        e2 = check(name + " tf loss", mine["loss"].detach(), out["loss"].detach())
        g_mine = oracle_grads(sd_req, mine["loss"], w)
        e3 = max(check(f"{name} tf grad {k}", g_mine[k], g_ref[k], 1e-5) for k in g_ref)
        print(f"[{name}] teacher-forced: restatement vs reference  logits {e1:.1e}  loss {e2:.1e}  grads {e3:.1e}")
        golden[f"{name}.source"], golden[f"{name}.target"] = src.numpy(), tgt.numpy()
        golden[f"{name}.tf.logits"] = logits.numpy()
        golden[f"{name}.tf.loss"] = out["loss"].detach().numpy()
        golden[f"{name}.weights"] = w.numpy()
        grads_record(golden, f"{name}.tf", g_ref)

        # ---- teacher forcing, greedy predictions in eval mode (the validation path, :208-209,256-274) ----
        ref.eval()
        with torch.no_grad():
            out = ref(src, tgt, decoding_strategy="greedy")
        assert torch.equal(out["predictions"], seq2seq_oracle.seq2seq_forward(sd, src, tgt, "greedy")["predictions"])
        golden[f"{name}.tf.greedy_predictions"] = out["predictions"].numpy()

        # ---- free-running greedy decoding (evaluation: :175-177,197-198,208-209,230-244) ----
        raw_box.clear(); logits_box.clear()
        with torch.no_grad():
            out = ref(src, None, decoding_strategy="greedy")
        mine = seq2seq_oracle.seq2seq_forward(sd, src, None, "greedy", steps_free)
        assert torch.equal(raw_box[0], mine["raw_predictions"]) and torch.equal(out["predictions"], mine["predictions"])
        check(name + " greedy loss", mine["loss"], out["loss"])
        golden[f"{name}.greedy.raw_predictions"] = raw_box[0].numpy()
        golden[f"{name}.greedy.predictions"] = out["predictions"].numpy()
        golden[f"{name}.greedy.loss"] = out["loss"].numpy()
        golden[f"{name}.greedy.logits"] = torch.stack(logits_box, 1).numpy()

        # ---- free-running sampling, training mode (REINFORCE: :210-215,230-244); the draws are the reference's ----
        ref.train()
        torch.manual_seed(200 + iseed)
        raw_box.clear()
        out = ref(src, None, decoding_strategy="sampling")
        raw = raw_box[0]
        assert int((raw <= 2).sum()) == 0                       # pad / unk / start are never drawn (:212-214)
        g_ref = reference_grads(ref, out["loss"], w)
        mine = seq2seq_oracle.seq2seq_forward(sd_req, src, None, "sampling", steps_free, forced_choices=raw)
        assert torch.equal(out["predictions"], mine["predictions"])
        e2 = check(name + " sampled loss", mine["loss"].detach(), out["loss"].detach())
        g_mine = oracle_grads(sd_req, mine["loss"], w)
        e3 = max(check(f"{name} sampled grad {k}", g_mine[k], g_ref[k], 1e-5) for k in g_ref)
        print(f"[{name}] sampled: restatement vs reference  loss {e2:.1e}  grads {e3:.1e};"
              f" rows ending early {(out['predictions'] == 0).any(1).sum().item()}/{rows}")
        golden[f"{name}.sampled.raw_predictions"] = raw.numpy()
        golden[f"{name}.sampled.predictions"] = out["predictions"].numpy()
        golden[f"{name}.sampled.loss"] = out["loss"].detach().numpy()
        grads_record(golden, f"{name}.sampled", g_ref)
        hook.remove()

    # ---- a model whose first greedy token is @end@: the trimmed row is empty, the loss 0 (:285-289, :244) ----
    sd = make_seq2seq_state_dict(93, 44, seed=4)
    sd["_output_projection_layer.bias"] = sd["_output_projection_layer.bias"].clone()
    sd["_output_projection_layer.bias"][END] = 50.0
    ref = RefPG(vocab)
    ref.load_state_dict(sd, strict=True)
    ref.eval()
    src, _ = case_inputs(vocab, "pg", 7, 4)
    with torch.no_grad():
        out = ref(src, None, decoding_strategy="greedy")
    mine = seq2seq_oracle.seq2seq_forward(sd, src, None, "greedy", 26)
    assert torch.equal(out["predictions"], mine["predictions"]) and int(out["predictions"].abs().sum()) == 0
    check("end-first loss", mine["loss"], out["loss"])
    golden["end_first.source"] = src.numpy()
    golden["end_first.predictions"] = out["predictions"].numpy()
    golden["end_first.loss"] = out["loss"].numpy()

    # ---- ProgramPrior.forward (program_prior.py:80-155), hidden size 256 as in the joint-training configuration ----
    sdp = make_prior_state_dict(44, hidden=256, seed=0)
    prior = RefPrior(vocab, input_size=256, hidden_size=256)
    print("prior state dict:", prior.load_state_dict(sdp, strict=True))
    prior.eval()
    programs = ProgramSampler(vocab, seed=9).sample(16, 26)
    programs[3] = 0
    torch.manual_seed(300)
    with torch.no_grad():
        out = prior(programs)
    mine = prior_oracle.prior_forward(sdp, programs)
    e = check("prior loss", mine["loss"], out["loss"])
    print(f"[prior] restatement vs reference loss {e:.1e}")
    golden["prior.programs"] = programs.numpy()
    golden["prior.loss"] = out["loss"].numpy()
    golden["prior.predictions_shape"] = np.asarray(out["predictions"].shape)

    path = os.path.join(REPO, "tests", "golden", "seq2seq_golden.npz")
    np.savez_compressed(path, **golden)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
