#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
bash -c 'sed -n "/^cat > \/tmp\/one_step.py/,/^PY$/p" scripts/gpu_ncu_list.sh | sed "1d;\$d" > /tmp/one_step.py'
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:wgrad_tc -c 1 \
    -f -o gpurun_out/wgradtc_r1 python /tmp/one_step.py > gpurun_out/ncu_wgradtc.log 2>&1; tail -3 gpurun_out/ncu_wgradtc.log
