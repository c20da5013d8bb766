"""diagnostics: device-side timeline of one fused joint-training step (marks recorded on the streams)"""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from probnmn_clevr_b200.joint import JointTrainingStep, split_batch
from probnmn_clevr_b200.nmn import NeuralModuleNetwork
from probnmn_clevr_b200.program_prior import ProgramPrior
from probnmn_clevr_b200.seq2seq import ProgramGenerator, QuestionReconstructor
from probnmn_clevr_b200.synthetic import make_joint_batch
from probnmn_clevr_b200.vocabulary import Vocabulary

import torch.distributed as dist
WORLD = int(os.environ.get("WORLD_SIZE", "1"))
RANK = int(os.environ.get("RANK", "0"))
LOCAL = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(LOCAL)
if WORLD > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", LOCAL))
dev = torch.device("cuda", LOCAL)
vocab = Vocabulary.clevr()
sds = bench.joint_state_dicts(vocab)
models = {}
for name, cls in (("program_generator", ProgramGenerator), ("question_reconstructor", QuestionReconstructor),
                  ("nmn", NeuralModuleNetwork), ("program_prior", ProgramPrior)):
    m = cls(vocab); m.load_state_dict(sds[name]); models[name] = m.to(dev).train()
if WORLD > 1 and os.environ.get("PNMN_NO_GRAD_OVERLAP") is None:
    models["nmn"].enable_gradient_overlap()
js = JointTrainingStep(models["program_generator"], models["question_reconstructor"], models["nmn"], models["program_prior"],
                       concurrent=os.environ.get("PNMN_NO_STREAMS") is None, **bench.JOINT)
parts = []
for i in range(2):
    p = split_batch(make_joint_batch(vocab, 256, seed=100 * RANK + i))
    parts.append({k: {kk: vv.to(dev) for kk, vv in v.items()} for k, v in p.items()})
for i in range(10):
    js.step(parts[i % 2])
torch.cuda.synchronize()
acc = {}
N = 10
for i in range(N):
    js.trace = []
    t0 = time.perf_counter()
    js.step(parts[i % 2])
    host_ms = (time.perf_counter() - t0) * 1e3
    torch.cuda.synchronize()
    base = js.trace[0][1]
    for label, ev in js.trace:
        acc.setdefault(label, []).append(base.elapsed_time(ev))
    acc.setdefault("host_issue_ms", []).append(host_ms)
js.trace = None
if RANK == 0:
    for k, v in acc.items():
        print(f"{k:28s} {sum(v) / len(v):7.3f} ms")
if WORLD > 1:
    dist.destroy_process_group()
