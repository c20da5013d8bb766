"""ORACLE tooling — golden vectors for the ELBO / REINFORCE glue from the reference's OWN ``probnmn/modules/elbo.py``
(this container only: /root/reference does not exist on the GPU box).

The file is loaded by path, unmodified; ``probnmn.models`` (which would pull in AllenNLP) is replaced by a stub module
that only provides the four class names elbo.py imports for its type annotations.  The four models are replaced by stubs
returning seeded per-row losses (leaf tensors, so that the gradient of the objective w.r.t. every loss is recorded too).
Three successive calls per case exercise the moving-average baseline.

    python oracle/make_elbo_golden.py        ->  tests/golden/elbo_golden.npz
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def load_reference_elbo(root="/root/reference"):
    stub = types.ModuleType("probnmn.models")
    for name in ("ProgramGenerator", "ProgramPrior", "QuestionReconstructor", "NeuralModuleNetwork"):
        setattr(stub, name, type(name, (torch.nn.Module,), {}))
    pkg = types.ModuleType("probnmn")
    pkg.__path__ = []
    saved = {k: sys.modules.get(k) for k in ("probnmn", "probnmn.models")}
    sys.modules["probnmn"], sys.modules["probnmn.models"] = pkg, stub
    try:
        spec = importlib.util.spec_from_file_location("probnmn_reference_elbo", os.path.join(root, "probnmn", "modules", "elbo.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod


class StubModel(torch.nn.Module):
    """Returns the next prepared {"predictions", "loss"} dictionary, whatever it is called with."""

    def __init__(self):
        super().__init__()
        self.queue = []

    def forward(self, *args, **kwargs):
        return self.queue.pop(0)


def cases():
    g = torch.Generator().manual_seed(1234)
    out = []
    for n in (1, 7, 128, 131):
        calls = []
        for _ in range(3):
            calls.append({"pg": torch.rand(n, generator=g) * 3 + 0.1, "qr": torch.rand(n, generator=g) * 4 + 0.2,
                          "prior": torch.rand(n, generator=g) * 2 + 0.3,
                          "nmn": torch.where(torch.rand(n, generator=g) < 0.2, torch.full((n,), 3.33), torch.rand(n, generator=g) * 5)})
        out.append((n, calls))
    return out


def main():
    ref = load_reference_elbo()
    golden = {}
    for mode in ("joint_ours", "joint_baseline", "question_coding"):
        for n, calls in cases():
            pg, qr, prior, nmn = StubModel(), StubModel(), StubModel(), StubModel()
            beta, gamma, decay = 0.1, 1.5, 0.99
            if mode == "question_coding":
                elbo = ref.QuestionCodingElbo(pg, qr, prior, beta=beta, baseline_decay=decay)
            else:
                elbo = ref.JointTrainingElbo(pg, qr, prior, nmn, beta=beta, gamma=gamma, baseline_decay=decay,
                                             objective="ours" if mode == "joint_ours" else "baseline")
            for k, c in enumerate(calls):
                leaves = {key: v.clone().requires_grad_(True) for key, v in c.items()}
                dummy = torch.zeros(n, 3, dtype=torch.long)
                for model, key in ((pg, "pg"), (qr, "qr"), (prior, "prior"), (nmn, "nmn")):
                    model.queue.append({"predictions": dummy, "loss": leaves[key]})
                if mode == "question_coding":
                    out = elbo(dummy)
                    objective = -out["elbo"]
                else:
                    out = elbo(dummy, torch.zeros(n, 1), torch.zeros(n, dtype=torch.long))
                    objective = gamma * out["nmn_loss"] - out["elbo"]        # joint_training_trainer.py:145-146
                objective.backward()
                tag = f"{mode}.n{n}.call{k}"
                for key, v in c.items():
                    golden[f"{tag}.in.{key}"] = v.numpy()
                for key, v in out.items():
                    golden[f"{tag}.out.{key}"] = np.float32(v.item())
                golden[f"{tag}.objective"] = np.float32(objective.item())
                for key, leaf in leaves.items():
                    golden[f"{tag}.grad.{key}"] = (leaf.grad if leaf.grad is not None else torch.zeros(n)).numpy()
                golden[f"{tag}.baseline_after"] = np.float64(elbo._reinforce._reinforce_baseline)
            for model in (pg, qr, prior, nmn):
                model.queue.clear()
    golden["hyper"] = np.array([0.1, 1.5, 0.99], np.float64)
    path = os.path.join(REPO, "tests", "golden", "elbo_golden.npz")
    np.savez_compressed(path, **golden)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB,", len(golden), "arrays")


if __name__ == "__main__":
    main()
