"""diagnostics (N > 1, torchrun): when do the NCCL kernels of one joint-training step run relative to the module network's
persistent kernels, the LSTM passes' last kernels and the optimizer?  kineto trace of 2 steps on rank 0, compact text summary."""
import json, os, sys, tempfile
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from probnmn_clevr_b200.joint import JointTrainingStep, split_batch
from probnmn_clevr_b200.nmn import NeuralModuleNetwork
from probnmn_clevr_b200.program_prior import ProgramPrior
from probnmn_clevr_b200.seq2seq import ProgramGenerator, QuestionReconstructor
from probnmn_clevr_b200.synthetic import make_joint_batch
from probnmn_clevr_b200.vocabulary import Vocabulary
from torch.profiler import ProfilerActivity, profile

WORLD, RANK, LOCAL = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
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
js = JointTrainingStep(models["program_generator"], models["question_reconstructor"], models["nmn"], models["program_prior"], **bench.JOINT)
parts = []
for i in range(2):
    p = split_batch(make_joint_batch(vocab, 256, seed=100 * RANK + i))
    parts.append({k: {kk: vv.to(dev) for kk, vv in v.items()} for k, v in p.items()})
for i in range(10):
    js.step(parts[i % 2])
torch.cuda.synchronize()
if WORLD > 1:
    dist.barrier()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for i in range(2):
        js.step(parts[i % 2])
        torch.cuda.synchronize()
if RANK == 0:
    path = os.path.join(tempfile.gettempdir(), "joint_nccl_trace.json")
    prof.export_chrome_trace(path)
    ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") == "kernel"]
    ev.sort(key=lambda e: e["ts"])
    # second step = kernels after the largest gap
    gaps = [(ev[i + 1]["ts"] - (ev[i]["ts"] + ev[i]["dur"]), i) for i in range(len(ev) - 1)]
    cut = max(gaps)[1] + 1
    step = ev[cut:]
    t0 = step[0]["ts"]
    print(f"step: {len(step)} kernels, {(step[-1]['ts'] + step[-1]['dur'] - t0) / 1e3:.3f} ms")
    keys = ("nccl", "exec_kernel", "wgrad_tc_kernel", "clamp_adam", "gemm_split_tc", "elbo_glue", "finalize")
    for e in step:
        n = e["name"]
        if any(k in n for k in keys) and (e["dur"] > 15 or "nccl" in n or "elbo" in n):
            print(f"  {(e['ts'] - t0) / 1e3:7.3f} ms  +{e['dur'] / 1e3:6.3f} ms  stream {e['args'].get('stream')}  grid {e['args'].get('grid')}  {n[:70]}")
    last = {}
    for e in step:
        last[e["args"].get("stream")] = (e["ts"] + e["dur"] - t0) / 1e3
    print("last kernel end per stream:", {k: round(v, 3) for k, v in last.items()})
if WORLD > 1:
    dist.destroy_process_group()
