#!/bin/bash
# Round-1 evidence run: host/device split of a step, executor timeline, ncu launch list, ncu --set full of the
# persistent executor and the wgrad kernel.  Everything lands in gpurun_out/ (summaries are copied to profiles/).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python scripts/profile_step.py > gpurun_out/profile_step.txt 2>&1; tail -45 gpurun_out/profile_step.txt | cut -c1-200
timeout 300 python scripts/trace_exec.py 2>&1 | grep -E "conv n_samp|elt op|==|sum|busy" | cut -c1-400 > gpurun_out/trace.txt; cat gpurun_out/trace.txt | cut -c1-300
bash scripts/gpu_ncu_list.sh 2>&1 | tail -30
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:exec_kernel -c 2 \
    -f -o gpurun_out/exec_r1 python /tmp/one_step.py > gpurun_out/ncu_exec.log 2>&1; tail -3 gpurun_out/ncu_exec.log
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:wgrad -c 1 \
    -f -o gpurun_out/wgrad_r1 python /tmp/one_step.py > gpurun_out/ncu_wgrad.log 2>&1; tail -3 gpurun_out/ncu_wgrad.log
ls -la gpurun_out
