#!/bin/bash
# Round-2 evidence run (tag r2): ncu launch list of ONE joint-training step, ncu --set full of exec_kernel (forward + backward
# launch of that step) and of wgrad_tc_kernel.  Everything lands in gpurun_out/; `python scripts/summarize_profiles.py r2`
# copies the summaries to profiles/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cat > /tmp/one_joint_step.py <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd())
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
import bench
from probnmn_clevr_b200.joint import JointTrainingStep, split_batch
from probnmn_clevr_b200.nmn import NeuralModuleNetwork
from probnmn_clevr_b200.program_prior import ProgramPrior
from probnmn_clevr_b200.seq2seq import ProgramGenerator, QuestionReconstructor
from probnmn_clevr_b200.synthetic import make_joint_batch
from probnmn_clevr_b200.vocabulary import Vocabulary
import torch.cuda.profiler as cp
dev = torch.device("cuda", 0)
vocab = Vocabulary.clevr()
sds = bench.joint_state_dicts(vocab)
models = {}
for name, cls in (("program_generator", ProgramGenerator), ("question_reconstructor", QuestionReconstructor),
                  ("nmn", NeuralModuleNetwork), ("program_prior", ProgramPrior)):
    m = cls(vocab); m.load_state_dict(sds[name]); models[name] = m.to(dev).train()
js = JointTrainingStep(models["program_generator"], models["question_reconstructor"], models["nmn"], models["program_prior"], **bench.JOINT)
p = split_batch(make_joint_batch(vocab, 256, seed=0))
parts = {k: {kk: vv.to(dev) for kk, vv in v.items()} for k, v in p.items()}
for _ in range(4): js.step(parts)
torch.cuda.synchronize()
cp.start(); js.step(parts); torch.cuda.synchronize(); cp.stop()
print("plan stats", models["nmn"].last_plan_stats)
PY
TAG=${TAG:-r2}
timeout 1200 ncu --profile-from-start off --metrics gpu__time_duration.sum,launch__grid_size --clock-control none --csv \
    --log-file gpurun_out/launches_${TAG}.csv python /tmp/one_joint_step.py > gpurun_out/ncu_list.log 2>&1
tail -2 gpurun_out/ncu_list.log
[ -n "$ONLY_LIST" ] && exit 0
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:exec_kernel -c 2 \
    -f -o gpurun_out/exec_${TAG} python /tmp/one_joint_step.py > gpurun_out/ncu_exec.log 2>&1; tail -2 gpurun_out/ncu_exec.log
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:wgrad_tc -c 1 \
    -f -o gpurun_out/wgradtc_${TAG} python /tmp/one_joint_step.py > gpurun_out/ncu_wgrad.log 2>&1; tail -2 gpurun_out/ncu_wgrad.log
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:step_gemm_tc -s 40 -c 2 \
    -f -o gpurun_out/stepgemm_${TAG} python /tmp/one_joint_step.py > gpurun_out/ncu_stepgemm.log 2>&1; tail -2 gpurun_out/ncu_stepgemm.log
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_split_tc -c 3 \
    -f -o gpurun_out/gemm_${TAG} python /tmp/one_joint_step.py > gpurun_out/ncu_gemm.log 2>&1; tail -2 gpurun_out/ncu_gemm.log
ls -la gpurun_out/*.ncu-rep
