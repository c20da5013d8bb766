#!/bin/bash
# ncu launch list of ONE bench step (per-launch device time; cold-cache, serialised: compare shares)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cat > /tmp/one_step.py <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd())
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
from probnmn_clevr_b200.nmn import NeuralModuleNetwork
from probnmn_clevr_b200.synthetic import ProgramSampler, make_answers, make_features, make_nmn_state_dict
from probnmn_clevr_b200.vocabulary import Vocabulary
import torch.cuda.profiler as cp
vocab = Vocabulary.clevr()
m = NeuralModuleNetwork(vocab); m.load_state_dict(make_nmn_state_dict(vocab, 0)); m = m.cuda().train()
B = 256
feats = make_features(B, 0).cuda(); progs = ProgramSampler(vocab, seed=0).sample(B, 40); ans = make_answers(B, 0).cuda()
def step():
    m.zero_grad(set_to_none=True)
    out = m(feats, progs.cuda(), ans); out["loss"].mean().backward()
for _ in range(2): step()
torch.cuda.synchronize()
cp.start(); step(); torch.cuda.synchronize(); cp.stop()
PY
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,launch__grid_size --clock-control none --csv --log-file gpurun_out/launches_${TAG:-r1}.csv python /tmp/one_step.py > gpurun_out/ncu_list.log 2>&1
tail -3 gpurun_out/ncu_list.log
python - <<'PY'
import csv, collections
import os
rows=[r for r in csv.reader(open('gpurun_out/launches_%s.csv' % os.environ.get('TAG','r1'))) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); mi=hdr.index('Metric Name'); vi=hdr.index('Metric Value'); ii=hdr.index('ID')
d=collections.OrderedDict()
for r in rows[1:]:
    d.setdefault(r[ii],{'k':r[ki]})[r[mi]]=float(r[vi].replace(',',''))
agg=collections.defaultdict(lambda:[0,0.0])
for v in d.values():
    agg[v['k'][:60]][0]+=1; agg[v['k'][:60]][1]+=v.get('gpu__time_duration.sum',0)/1e3
tot=sum(a[1] for a in agg.values())
print('total us',tot)
for k,a in sorted(agg.items(), key=lambda x:-x[1][1])[:16]: print(f'{a[1]:10.1f} us {a[0]:5d} launches {100*a[1]/tot:5.1f}%  {k}')
print('--- conv_tc launches (grid, us) in order')
print([ (int(v.get('launch__grid_size',0)), round(v['gpu__time_duration.sum']/1e3,1)) for v in d.values() if 'conv_tc' in v['k']])
print('--- elt launches (grid, us)')
print([ (int(v.get('launch__grid_size',0)), round(v['gpu__time_duration.sum']/1e3,1)) for v in d.values() if 'elt_kernel' in v['k']])
PY
