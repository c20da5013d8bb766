"""GPU diagnostic: the executor's forward pass has no atomics, so repeated runs must agree bit for bit (a mismatch would
mean a missing fence between a producer task and its consumers); backward gradients are compared to rounding."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
from probnmn_clevr_b200.nmn import NeuralModuleNetwork
from probnmn_clevr_b200.synthetic import ProgramSampler, make_answers, make_features, make_nmn_state_dict
from probnmn_clevr_b200.vocabulary import Vocabulary
B = int(os.environ.get("B", 256)); R = int(os.environ.get("R", 8))
vocab = Vocabulary.clevr()
m = NeuralModuleNetwork(vocab); m.load_state_dict(make_nmn_state_dict(vocab, 0)); m = m.cuda().train()
feats = make_features(B, 0).cuda(); progs = ProgramSampler(vocab, seed=0).sample(B, 40); ans = make_answers(B, 0).cuda()
box = {}
m.classifier.register_forward_hook(lambda mod, i, o: box.update(final=i[0].detach().clone(), logits=o.detach().clone()))
ref = None; bad_f = bad_l = 0; gmax = 0.0
for r in range(R):
    m.zero_grad(set_to_none=True)
    out = m(feats, progs, ans); out["loss"].mean().backward()
    g = torch.cat([p.grad.flatten() for p in m.parameters() if p.grad is not None])
    cur = (box["final"], box["logits"], g.clone())
    if ref is None: ref = cur; continue
    nf = int((cur[0] != ref[0]).sum()); nl = int((cur[1] != ref[1]).sum())
    bad_f += nf > 0; bad_l += nl > 0
    gmax = max(gmax, float((cur[2] - ref[2]).abs().max() / ref[2].abs().max()))
    if nf: print(f"run {r}: {nf} final elements differ, max abs diff {float((cur[0]-ref[0]).abs().max()):.3e}")
print(f"forward bitwise mismatches: final {bad_f}/{R-1} runs, logits {bad_l}/{R-1} runs; gradient max rel diff {gmax:.2e}")
