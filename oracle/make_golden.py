"""ORACLE tooling — generate tests/golden/nmn_golden.npz from the REAL reference (run in the build
container, where /root/reference exists):

    python oracle/make_golden.py

It imports the reference's NeuralModuleNetwork verbatim (oracle/ref_loader.py), loads seeded weights
(probnmn_clevr_b200.synthetic.make_nmn_state_dict) into it, runs it on seeded features and on three
program sets (hand-built interpreter-semantics cases of SURVEY.md appendix A, grammar-sampled valid
CLEVR programs, uniform-random garbage), asserts that the restatement oracle/nmn_oracle.py agrees, and
stores the reference's outputs.  Weights and features are NOT stored (regenerated from their seeds).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, REPO)

from oracle import nmn_oracle  # noqa: E402
from oracle.ref_loader import load_reference_nmn  # noqa: E402
from probnmn_clevr_b200.synthetic import (ProgramSampler, make_answers, make_features, make_nmn_state_dict,  # noqa: E402
                                          programs_from_tokens)
from probnmn_clevr_b200.vocabulary import Vocabulary  # noqa: E402

WEIGHT_SEED, FEATURE_SEED = 0, 0

# interpreter semantics + one program per module family
SEMANTIC_CASES = [
    [],
    ["@start@", "@end@"],
    ["query_color", "scene"],
    ["filter_color[red]", "scene"],
    ["count", "count", "scene"],
    ["count", "filter_shape[cube]"],
    ["intersect", "scene"],
    ["count", "intersect", "filter_size[large]", "scene"],
    ["equal_color", "query_color", "scene"],
    ["equal_color", "scene", "scene"],
    ["count", "scene", "scene"],
    ["count", "@@UNKNOWN@@", "filter_color[blue]", "unique", "scene"],
    ["union", "scene"],
    ["exist", "filter_material[metal]", "filter_color[cyan]", "scene"],
    ["query_shape", "unique", "relate[left]", "unique", "filter_color[red]", "scene"],
    ["count", "same_color", "unique", "filter_shape[sphere]", "scene"],
    ["query_size", "unique", "same_shape", "unique", "filter_size[small]", "relate[behind]", "unique",
     "filter_material[rubber]", "scene"],
    ["count", "union", "filter_color[red]", "scene", "filter_shape[cube]", "scene"],
    ["exist", "intersect", "relate[front]", "unique", "filter_color[green]", "scene", "filter_size[large]", "scene"],
    ["equal_integer", "count", "filter_color[red]", "scene", "count", "filter_shape[cube]", "scene"],
    ["less_than", "count", "filter_color[gray]", "scene", "count", "relate[right]", "unique", "filter_shape[cylinder]",
     "scene"],
    ["equal_material", "query_material", "unique", "filter_color[purple]", "scene", "query_material", "unique",
     "filter_shape[sphere]", "scene"],
    ["greater_than", "count", "scene", "count", "filter_size[small]", "scene"],
    ["union", "query_color", "scene"],
]

BIG_SUBSAMPLE = 997  # stride of the stored slice of large gradient tensors


def run_reference(ref, features, programs, answers):
    ref.train()
    out = ref(features, programs, answers)
    return out


def main():
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    vocab = Vocabulary.clevr()
    RefNMN = load_reference_nmn()
    ref = RefNMN(vocab)
    sd = make_nmn_state_dict(vocab, WEIGHT_SEED)
    missing = ref.load_state_dict(sd, strict=True)
    print("reference state dict loaded:", missing, "tensors:", len(sd))

    sampler = ProgramSampler(vocab, seed=1)
    sets = {
        "semantic": programs_from_tokens(vocab, SEMANTIC_CASES, 26),
        "sampled": sampler.sample(8, 26),
        "garbage": sampler.garbage(8, 26),
    }
    golden = {}
    sd_req = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    for name, programs in sets.items():
        B = programs.shape[0]
        feats = make_features(B, FEATURE_SEED)
        answers = make_answers(B, FEATURE_SEED)
        # ---- the reference itself ----
        logits_box = {}
        h = ref.classifier.register_forward_hook(lambda m, i, o: logits_box.update(logits=o, final=i[0]))
        ref.zero_grad()
        out = run_reference(ref, feats, programs, answers)
        h.remove()
        out["loss"].mean().backward()
        ref_grads = {k: p.grad.detach().clone() if p.grad is not None else torch.zeros_like(p) for k, p in ref.named_parameters()}
        with torch.no_grad():
            out_na = ref(feats, programs)
        logits = logits_box["logits"].detach()
        final = logits_box["final"].detach()
        valid = (out["predictions"] != 28).long() if True else None
        # validity cannot be read off predictions alone if the net predicts... it cannot predict 28 (nmn.py:57-63)
        # ---- the restatement must agree ----
        for p in sd_req.values():
            p.grad = None
        mine = nmn_oracle.nmn_forward(sd_req, vocab, feats, programs, answers, want=("traces",))
        mine["loss"].mean().backward()
        assert torch.equal(mine["valid"], valid), (name, mine["valid"], valid)
        assert torch.equal(mine["predictions"], out["predictions"]), name
        assert torch.allclose(mine["logits"], logits, rtol=1e-5, atol=1e-5), (name, (mine["logits"] - logits).abs().max())
        assert torch.allclose(mine["loss"], out["loss"].detach(), rtol=1e-5, atol=1e-5), name
        assert torch.allclose(mine["final"], final, rtol=1e-5, atol=1e-5), name
        mine_na = nmn_oracle.nmn_forward(sd, vocab, feats, programs)
        assert torch.allclose(mine_na["loss"], out_na["loss"], rtol=1e-5, atol=1e-5), name
        gerr = 0.0
        for k, g in ref_grads.items():
            mg = sd_req[k].grad if sd_req[k].grad is not None else torch.zeros_like(sd_req[k])
            scale = g.abs().max().item() + 1e-12
            gerr = max(gerr, (mg - g).abs().max().item() / scale) if g.abs().max() > 0 else gerr
            assert torch.allclose(mg, g, rtol=1e-4, atol=1e-5 * scale + 1e-9), (name, k)
        print(f"[{name}] B={B} valid={valid.tolist()} max rel grad diff restatement-vs-reference {gerr:.2e}")

        golden[f"{name}.programs"] = programs.numpy()
        golden[f"{name}.answers"] = answers.numpy()
        golden[f"{name}.valid"] = valid.numpy()
        golden[f"{name}.logits"] = logits.numpy()
        golden[f"{name}.loss"] = out["loss"].detach().numpy()
        golden[f"{name}.loss_noanswer"] = out_na["loss"].numpy()
        golden[f"{name}.predictions"] = out["predictions"].numpy()
        golden[f"{name}.final_sum"] = final.sum(dim=(1, 2, 3)).numpy()
        golden[f"{name}.final_abs_sum"] = final.abs().sum(dim=(1, 2, 3)).numpy()
        golden[f"{name}.final_sample0"] = final[min(2, B - 1)].numpy()
        # attention maps of the reference: one sample at a time with hooks on every module
        maps = []
        for n in range(B):
            if not valid[n]:
                continue
            rec = []
            hooks = [m.register_forward_hook(lambda mod, i, o, rec=rec: rec.append(o.detach()))
                     for nm, m in ref.named_children() if nm not in ("stem", "classifier", "_loss")]
            with torch.no_grad():
                ref(feats[n:n + 1], programs[n:n + 1])
            for hk in hooks:
                hk.remove()
            mine_tr = [o for _, o in mine["traces"][n]]
            assert len(rec) == len(mine_tr), (name, n, len(rec), len(mine_tr))
            for a, b in zip(rec, mine_tr):
                assert torch.allclose(a, b.detach(), rtol=1e-5, atol=1e-5)
            for j, a in enumerate(rec):
                if a.shape[1] == 1:
                    maps.append(np.concatenate([[n, j], a.reshape(-1).numpy()]))
        golden[f"{name}.attention_maps"] = np.stack(maps).astype(np.float32) if maps else np.zeros((0, 198), np.float32)
        for k, g in ref_grads.items():
            if g.numel() <= 4096:
                golden[f"{name}.grad.{k}"] = g.numpy()
            else:
                golden[f"{name}.gradsub.{k}"] = g.reshape(-1)[::BIG_SUBSAMPLE].numpy().copy()
            golden[f"{name}.gradnorm.{k}"] = np.float64(g.double().norm().item())
    path = os.path.join(REPO, "tests", "golden", "nmn_golden.npz")
    np.savez_compressed(path, **golden)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
