"""GPU diagnostic: per-sample error of the module outputs against the tf32-rounding oracle."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
from oracle import nmn_oracle
from probnmn_clevr_b200.nmn import NeuralModuleNetwork
from probnmn_clevr_b200.synthetic import make_features, make_nmn_state_dict, make_answers
from probnmn_clevr_b200.vocabulary import Vocabulary

vocab = Vocabulary.clevr()
m = NeuralModuleNetwork(vocab); m.load_state_dict(make_nmn_state_dict(vocab, 0)); m = m.cuda()
g = np.load("tests/golden/nmn_golden.npz")
for name in ("semantic", "sampled"):
    programs = torch.from_numpy(g[f"{name}.programs"]); answers = torch.from_numpy(g[f"{name}.answers"])
    feats = make_features(programs.shape[0], 0)
    sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    with torch.no_grad(), nmn_oracle.operand_rounding("tf32"):
        ref = nmn_oracle.nmn_forward(sd, vocab, feats, programs, answers)
        reff = nmn_oracle.stem(sd, feats)
    with torch.no_grad():
        ref32 = nmn_oracle.nmn_forward(sd, vocab, feats, programs, answers)
    box = {}
    h = m.classifier.register_forward_hook(lambda mod, i, o: box.update(final=i[0].detach().cpu(), logits=o.detach().cpu()))
    with torch.no_grad():
        m(feats.cuda(), programs.cuda(), answers.cuda())
    h.remove()
    cls = nmn_oracle.classifier(sd, box["final"])
    print(f"== {name} impl={os.environ.get('PNMN_CONV_IMPL','tc')}: classifier(gpu) vs classifier(cpu) on the same input: "
          f"{float((cls - box['logits']).abs().max() / box['logits'].abs().max()):.2e}")
    for n in range(programs.shape[0]):
        toks = [vocab.get_token_from_index(int(t), 'programs') for t in programs[n] if int(t) != 0]
        a, b, c = box["final"][n], ref["final"][n], ref32["final"][n]
        e = float((a - b).abs().max() / (b.abs().max() + 1e-30))
        e32 = float((a - c).abs().max() / (c.abs().max() + 1e-30))
        nz = float((a != b).float().mean())
        print(f"  sample {n:2d} vs tf32-oracle {e:.2e} (frac differing {nz:.3f})  vs fp32 {e32:.2e}  {' '.join(toks)}")
