"""GPU diagnostic: where the HOST spends its time while issuing one fused joint-training step (cProfile + coarse stamps)."""
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

dev = torch.device("cuda", 0)
vocab = Vocabulary.clevr()
sds = bench.joint_state_dicts(vocab)
models = {}
for name, cls in (("program_generator", ProgramGenerator), ("question_reconstructor", QuestionReconstructor),
                  ("nmn", NeuralModuleNetwork), ("program_prior", ProgramPrior)):
    m = cls(vocab); m.load_state_dict(sds[name]); models[name] = m.to(dev).train()
js = JointTrainingStep(models["program_generator"], models["question_reconstructor"], models["nmn"], models["program_prior"],
                       **bench.JOINT)
parts = []
for i in range(2):
    p = split_batch(make_joint_batch(vocab, 256, seed=i))
    parts.append({k: {kk: vv.to(dev) for kk, vv in v.items()} for k, v in p.items()})
for i in range(10):
    js.step(parts[i % 2])
torch.cuda.synchronize()
N = 30
t0 = time.perf_counter()
for i in range(N):
    js.step(parts[i % 2])
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host issue {1e3 * (t1 - t0) / N:.3f} ms/step, incl. drain {1e3 * (t2 - t0) / N:.3f} ms/step")
import cProfile, pstats
pr = cProfile.Profile()
pr.enable()
for i in range(20):
    js.step(parts[i % 2])
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(35)
st.sort_stats("cumulative").print_stats(45)
