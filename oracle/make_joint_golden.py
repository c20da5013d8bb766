"""ORACLE tooling -- golden vectors for the whole joint-training iteration from the reference's OWN code (build container
only: /root/reference does not exist on the GPU box):

    python oracle/make_joint_golden.py        ->  tests/golden/joint_golden.npz

``JointTrainingTrainer._do_iteration`` (probnmn/trainers/joint_training_trainer.py:128-198, loaded by path, unmodified;
its trainer base class, dataset and checkpoint imports are replaced by empty stubs, none of them is reached by
``_do_iteration``) is called with a stand-in ``self`` that carries exactly what the method reads: the reference's own
``JointTrainingElbo`` (probnmn/modules/elbo.py) over the reference's own ``ProgramGenerator``, ``QuestionReconstructor``,
``ProgramPrior`` and ``NeuralModuleNetwork`` (all imported verbatim; AllenNLP comes from ``oracle/ref_shim``, see
``oracle/make_seq2seq_golden.py`` for what that does and does not pin) and the ALPHA / GAMMA / OBJECTIVE values of
``configs/joint_training_ours.yml``.  Two iterations in a row (``zero_grad`` in between, as ``_Trainer.step`` does,
trainers/_trainer.py:193) exercise the moving-average REINFORCE baseline.

Stored: the programs the reference sampled (raw multinomial draws, captured at the entrance of ``_trim_predictions``), the
iteration's output dictionary, the baseline after each iteration and, for every trained parameter, the norm of its
clamped gradient plus a slice of it.  Weights and inputs are regenerated from their seeds (the generator's from
the committed pre-trained asset ``probnmn_clevr_b200/assets/pg_synthetic_fp16.npz``).  The restatement
``oracle/joint_oracle.py`` must reproduce all of it (replaying the sampled programs) before the file is written.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, REPO)

from oracle import joint_oracle  # noqa: E402
from oracle.ref_loader import load_reference_nmn, load_reference_seq2seq  # noqa: E402
from probnmn_clevr_b200.synthetic import (make_joint_batch, make_nmn_state_dict, make_prior_state_dict,  # noqa: E402
                                          make_seq2seq_state_dict)
from probnmn_clevr_b200.vocabulary import Vocabulary  # noqa: E402

HYPER = dict(alpha=100.0, beta=0.1, gamma=1.0, delta=0.99)     # configs/joint_training_ours.yml:6-16
BATCH, BATCH_SEEDS = 10, (21, 22)
SUB = 9973   # stride of the stored slice of large gradient tensors
CLAMP = 5.0


def load_reference_trainer(root="/root/reference"):
    """``probnmn/trainers/joint_training_trainer.py`` by path; the modules it imports but ``_do_iteration`` never touches
    (datasets -> h5py, checkpointing -> loguru, the trainer base -> tensorboardX) are empty stubs."""
    stubs = {
        "probnmn.data.datasets": {"JointTrainingDataset": object},
        "probnmn.data.samplers": {"SupervisionWeightedRandomSampler": object},
        "probnmn.utils.checkpointing": {"CheckpointManager": object},
        "probnmn.trainers": {},
        "probnmn.trainers._trainer": {"_Trainer": object},
    }
    saved = {k: sys.modules.get(k) for k in stubs}
    for name, attrs in stubs.items():
        mod = types.ModuleType(name)
        mod.__dict__.update(attrs)
        if name == "probnmn.trainers":
            mod.__path__ = []
        sys.modules[name] = mod
    try:
        spec = importlib.util.spec_from_file_location("probnmn.trainers.joint_training_trainer",
                                                      os.path.join(root, "probnmn", "trainers", "joint_training_trainer.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod.JointTrainingTrainer


PG_ASSET = os.path.join(REPO, "probnmn_clevr_b200", "assets", "pg_synthetic_fp16.npz")


def state_dicts(vocab):
    """The generator is the one pre-trained on the synthetic question -> program mapping (scripts/pretrain_pg.py; the
    asset bench.py uses), so that its samples are mostly executable programs and the module network's branch of the
    objective carries gradients; everything else is seeded."""
    z = np.load(PG_ASSET)
    return {"program_generator": {k: torch.from_numpy(z[k].astype("float32")) for k in z.files},
            "question_reconstructor": make_seq2seq_state_dict(44, 93, seed=12),
            "nmn": make_nmn_state_dict(vocab, 13),
            "program_prior": make_prior_state_dict(44, hidden=256, seed=14)}


def batches(vocab):
    return [make_joint_batch(vocab, BATCH, seed=s) for s in BATCH_SEEDS]


def clamped_grads(named_params):
    return {k: (p.grad.detach().clone() if p.grad is not None else torch.zeros_like(p)) for k, p in named_params}


def main():
    torch.set_num_threads(os.cpu_count())
    vocab = Vocabulary.clevr()
    RefPG, RefQR, RefPrior = load_reference_seq2seq()
    RefNMN = load_reference_nmn()
    import probnmn.modules.elbo as ref_elbo                     # the reference's elbo.py, imported normally
    Trainer = load_reference_trainer()
    sds = state_dicts(vocab)
    models = {"program_generator": RefPG(vocab), "question_reconstructor": RefQR(vocab), "nmn": RefNMN(vocab),
              "program_prior": RefPrior(vocab, input_size=256, hidden_size=256)}
    for k, m in models.items():
        print(k, m.load_state_dict(sds[k], strict=True))
        m.train()
    models["program_prior"].eval()                              # joint_training_trainer.py:112
    elbo = ref_elbo.JointTrainingElbo(program_generator=models["program_generator"],
                                      question_reconstructor=models["question_reconstructor"], nmn=models["nmn"],
                                      program_prior=models["program_prior"], beta=HYPER["beta"], gamma=HYPER["gamma"],
                                      baseline_decay=HYPER["delta"], objective="ours")
    stand_in = types.SimpleNamespace(
        _C=types.SimpleNamespace(GAMMA=HYPER["gamma"], ALPHA=HYPER["alpha"], OBJECTIVE="ours"), _elbo=elbo,
        _program_generator=models["program_generator"], _question_reconstructor=models["question_reconstructor"],
        _nmn=models["nmn"])
    # record the generator's raw draws: _trim_predictions (seq2seq_base.py:278) is handed them before trimming
    raw_box = []
    pg = models["program_generator"]
    trim = pg._trim_predictions
    pg._trim_predictions = lambda p: (raw_box.append(p.detach().clone()), trim(p))[1]

    sds_req = {name: {k: v.clone().requires_grad_(True) for k, v in sd.items()} for name, sd in sds.items()}
    sds_req["program_prior"]["_output_layer.weight"] = sds_req["program_prior"]["_embedder.token_embedder_programs.weight"]
    state = joint_oracle.ElboState(0.0)
    golden = {}
    trained = ("program_generator", "question_reconstructor", "nmn")
    for it, batch in enumerate(batches(vocab)):
        for name in trained:
            models[name].zero_grad()
        torch.manual_seed(500 + it)
        raw_box.clear()
        out = Trainer._do_iteration(stand_in, batch)            # forward, backward and the [-5, 5] clamp (:128-198)
        raw = raw_box[0]                                         # first _trim_predictions call = the unsupervised rows (elbo.py:230)
        n_unsup = int((1 - batch["supervision"]).sum())
        assert raw.shape == (n_unsup, 26)
        objective = (HYPER["gamma"] * out["loss"]["nmn"] - out["elbo"]["elbo"]
                     + HYPER["alpha"] * (out["loss"]["program_generation_gt"] + out["loss"]["question_reconstruction_gt"]))
        # ---- the restatement must agree ----
        for name in trained:
            for p in sds_req[name].values():
                p.grad = None
        mine = joint_oracle.joint_iteration(sds_req, vocab, batch, state, objective="ours", forced_programs=raw, **HYPER)
        for p in joint_oracle.trained_parameters(sds_req):
            if p.grad is not None:
                p.grad.clamp_(min=-CLAMP, max=CLAMP)
        for k, v in out["elbo"].items():
            assert abs(v.item() - mine["elbo"][k].item()) <= 1e-5 * max(1.0, abs(v.item())), (it, k, v.item(), mine["elbo"][k].item())
        for k, v in out["loss"].items():
            assert abs(v.item() - mine["loss"][k].item()) <= 1e-5 * max(1.0, abs(v.item())), (it, k)
        assert abs(state.baseline - elbo._reinforce._reinforce_baseline) <= 1e-5 * max(1.0, abs(state.baseline))
        worst = 0.0
        for name in trained:
            ref_g = clamped_grads(models[name].named_parameters())
            for k, g in ref_g.items():
                mg = sds_req[name][k].grad
                mg = torch.zeros_like(g) if mg is None else mg
                err = float((mg - g).abs().max() / (g.abs().max() + 1e-12)) if float(g.abs().max()) > 0 else float(mg.abs().max())
                worst = max(worst, err)
                assert err < 2e-4, (it, name, k, err)
                golden[f"it{it}.gradnorm.{name}.{k}"] = np.float64(g.double().norm().item())
                golden[f"it{it}.gradsub.{name}.{k}"] = g.reshape(-1)[::(SUB if g.numel() > 4096 else 1)].numpy().copy()
        valid = int(mine["rows"]["nmn_valid"].sum())
        print(f"iteration {it}: {n_unsup} unsupervised rows, {valid} executable sampled programs, objective {objective.item():.4f},"
              f" baseline {elbo._reinforce._reinforce_baseline:.4f}; restatement vs reference: worst gradient entry {worst:.1e}")
        golden[f"it{it}.raw_programs"] = raw.numpy()
        golden[f"it{it}.baseline"] = np.float64(elbo._reinforce._reinforce_baseline)
        golden[f"it{it}.objective"] = np.float64(objective.item())
        for k, v in out["elbo"].items():
            golden[f"it{it}.elbo.{k}"] = np.float64(v.item())
        for k, v in out["loss"].items():
            golden[f"it{it}.loss.{k}"] = np.float64(v.item())
    path = os.path.join(REPO, "tests", "golden", "joint_golden.npz")
    np.savez_compressed(path, **golden)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
