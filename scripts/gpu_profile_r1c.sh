#!/bin/bash
# End-of-round-1 evidence run (tag r1c): smoke + all GPU tests, default bench line, executor timeline, ncu launch list of
# one NMN step, ncu --set full of exec_kernel (forward + backward launch) and of wgrad_tc_kernel.
# Everything lands in gpurun_out/; `python scripts/summarize_profiles.py r1c` copies the summaries to profiles/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err || tail -5 gpurun_out/bench.err
timeout 300 python scripts/profile_step.py > gpurun_out/profile_step.txt 2>&1
timeout 300 python scripts/trace_exec.py 2>&1 | grep -E "conv n_samp|elt op|==|sum|busy|epilogue" | cut -c1-400 > gpurun_out/trace.txt
TAG=r1c bash scripts/gpu_ncu_list.sh 2>&1 | tail -30
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:exec_kernel -c 2 \
    -f -o gpurun_out/exec_r1c python /tmp/one_step.py > gpurun_out/ncu_exec.log 2>&1; tail -2 gpurun_out/ncu_exec.log
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:wgrad_tc -c 1 \
    -f -o gpurun_out/wgradtc_r1c python /tmp/one_step.py > gpurun_out/ncu_wgrad.log 2>&1; tail -2 gpurun_out/ncu_wgrad.log
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print('value', round(d['value']), 'ms/step', round(d['ms_per_step'],2), 'e2e', d['e2e'], 'roofline', {k: d['roofline'][k] for k in ('achieved','peak','frac','traffic')}, 'clocks', d['clocks'])
print({k: round(v,2) for k,v in d['kernel_ms_per_step'].items()}, d.get('host_ms_per_step'), 'launches', d['gpu_launches'])
print('pg', d.get('pg'), '\njoint', d.get('joint'), '\ncpu', d.get('cpu_baseline'))
PY
ls -la gpurun_out | head -40
