#!/bin/bash
# launch list (ncu gpu__time_duration) of one ProgramGenerator step: 128 rows teacher-forced + 128 rows sampled, fwd+bwd
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cat > /tmp/pg_step.py <<'PY'
import os, sys, time, torch
sys.path.insert(0, os.getcwd())
import torch.cuda.profiler as cp
from probnmn_clevr_b200.seq2seq import ProgramGenerator
from probnmn_clevr_b200.synthetic import ProgramSampler, make_questions, make_seq2seq_state_dict
from probnmn_clevr_b200.vocabulary import Vocabulary
vocab = Vocabulary.clevr()
pg = ProgramGenerator(vocab); pg.load_state_dict(make_seq2seq_state_dict(93, 44, seed=0)); pg = pg.cuda().train()
B = 256
q = make_questions(B, 93, seed=0, max_length=40).cuda(); p = ProgramSampler(vocab, seed=0).sample(B, 26).cuda()
def step():
    pg.zero_grad(set_to_none=True)
    a = pg(q[:128], p[:128], decoding_strategy="sampling"); b = pg(q[128:], decoding_strategy="sampling")
    (a["loss"].mean() + b["loss"].mean()).backward()
for _ in range(3): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10): step()
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"pg step: host launch {(t1-t0)*100:.2f} ms, with drain {(t2-t0)*100:.2f} ms")
cp.start(); step(); torch.cuda.synchronize(); cp.stop()
PY
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,launch__grid_size --clock-control none --csv --log-file gpurun_out/launches_pg_r1.csv python /tmp/pg_step.py 2>&1 | grep "pg step"
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/launches_pg_r1.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); mi=hdr.index('Metric Name'); vi=hdr.index('Metric Value'); ii=hdr.index('ID')
d=collections.OrderedDict()
for r in rows[1:]:
    d.setdefault(r[ii],{'k':r[ki]})[r[mi]]=float(r[vi].replace(',',''))
agg=collections.defaultdict(lambda:[0,0.0])
for v in d.values():
    agg[v['k'][:70]][0]+=1; agg[v['k'][:70]][1]+=v.get('gpu__time_duration.sum',0)/1e3
tot=sum(a[1] for a in agg.values())
print('total us',tot, 'launches', len(d))
for k,a in sorted(agg.items(), key=lambda x:-x[1][1])[:20]: print(f'{a[1]:10.1f} us {a[0]:5d} launches {a[1]/a[0]:8.1f} us each {100*a[1]/tot:5.1f}%  {k}')
PY
