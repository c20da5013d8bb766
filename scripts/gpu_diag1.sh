#!/bin/bash
# round-1 diagnostics: executor timelines under three scheduling variants + ncu full capture of the executor kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in default nosplit nodeps; do
  unset PNMN_NOSPLIT PNMN_NODEPS
  [ $v = nosplit ] && export PNMN_NOSPLIT=1
  [ $v = nodeps ] && export PNMN_NODEPS=1
  timeout 300 python scripts/trace_exec.py 2>&1 | grep -E "conv n_samp|elt op|==|sum|busy" | cut -c1-330 > gpurun_out/trace_$v.txt
  head -3 gpurun_out/trace_$v.txt
done
unset PNMN_NOSPLIT PNMN_NODEPS
cat > /tmp/one_step.py <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd())
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
from probnmn_clevr_b200.nmn import NeuralModuleNetwork
from probnmn_clevr_b200.synthetic import ProgramSampler, make_answers, make_features, make_nmn_state_dict
from probnmn_clevr_b200.vocabulary import Vocabulary
vocab = Vocabulary.clevr()
m = NeuralModuleNetwork(vocab); m.load_state_dict(make_nmn_state_dict(vocab, 0)); m = m.cuda().train()
B = 256
feats = make_features(B, 0).cuda(); progs = ProgramSampler(vocab, seed=0).sample(B, 40); ans = make_answers(B, 0).cuda()
def step():
    m.zero_grad(set_to_none=True)
    out = m(feats, progs.cuda(), ans); out["loss"].mean().backward()
for _ in range(3): step()
torch.cuda.synchronize()
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'exec_kernel|wgrad_tc' -s 6 -c 3 -f -o gpurun_out/prof_exec_r1 python /tmp/one_step.py > gpurun_out/ncu_exec.log 2>&1
tail -3 gpurun_out/ncu_exec.log
ls -la gpurun_out/*.ncu-rep
