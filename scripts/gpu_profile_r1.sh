#!/bin/bash
# Round-1 evidence run (final state): host/device split of a step, executor timeline, ncu launch lists (NMN step and
# ProgramGenerator step), ncu --set full of exec_kernel (2 launches), wgrad_tc_kernel and the LSTM step GEMM.
# Everything lands in gpurun_out/; scripts/summarize_profiles.py copies the summaries to profiles/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python scripts/profile_step.py > gpurun_out/profile_step.txt 2>&1; tail -34 gpurun_out/profile_step.txt | cut -c1-60,120-215
timeout 300 python scripts/trace_exec.py 2>&1 | grep -E "conv n_samp|elt op|==|sum|busy" | cut -c1-400 > gpurun_out/trace.txt
bash scripts/gpu_ncu_list.sh 2>&1 | tail -22
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:exec_kernel -c 2 \
    -f -o gpurun_out/exec_r1 python /tmp/one_step.py > gpurun_out/ncu_exec.log 2>&1; tail -2 gpurun_out/ncu_exec.log
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:wgrad_tc -c 1 \
    -f -o gpurun_out/wgradtc_r1 python /tmp/one_step.py > gpurun_out/ncu_wgrad.log 2>&1; tail -2 gpurun_out/ncu_wgrad.log
bash scripts/gpu_pg_profile.sh 2>&1 | tail -16
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:step_gemm_tc -s 60 -c 2 \
    -f -o gpurun_out/lstm_r1 python /tmp/pg_step.py > gpurun_out/ncu_lstm.log 2>&1; tail -2 gpurun_out/ncu_lstm.log
rm -f gpurun_out/wgrad_r1.ncu-rep
ls -la gpurun_out | head -30
