"""diagnostics: kineto (CUPTI) trace of a few joint-training steps -> gpurun_out/joint_trace.json (chrome trace)"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from probnmn_clevr_b200.joint import JointTrainingStep, split_batch
from probnmn_clevr_b200.nmn import NeuralModuleNetwork
from probnmn_clevr_b200.program_prior import ProgramPrior
from probnmn_clevr_b200.seq2seq import ProgramGenerator, QuestionReconstructor
from probnmn_clevr_b200.synthetic import make_joint_batch
from probnmn_clevr_b200.vocabulary import Vocabulary
from torch.profiler import ProfilerActivity, profile

dev = torch.device("cuda", 0)
vocab = Vocabulary.clevr()
sds = bench.joint_state_dicts(vocab)
models = {}
for name, cls in (("program_generator", ProgramGenerator), ("question_reconstructor", QuestionReconstructor),
                  ("nmn", NeuralModuleNetwork), ("program_prior", ProgramPrior)):
    m = cls(vocab); m.load_state_dict(sds[name]); models[name] = m.to(dev).train()
js = JointTrainingStep(models["program_generator"], models["question_reconstructor"], models["nmn"], models["program_prior"], **bench.JOINT)
parts = []
for i in range(2):
    p = split_batch(make_joint_batch(vocab, 256, seed=i))
    parts.append({k: {kk: vv.to(dev) for kk, vv in v.items()} for k, v in p.items()})
for i in range(10):
    js.step(parts[i % 2])
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for i in range(3):
        js.step(parts[i % 2])
    torch.cuda.synchronize()
out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "joint_trace.json")
prof.export_chrome_trace(out)
print("wrote", out, os.path.getsize(out) >> 20, "MiB")
