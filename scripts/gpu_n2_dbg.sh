#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
PNMN_DEBUG_INPUTS=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 12 --warmup 8 --no-cpu-baseline --no-extras > gpurun_out/dbg.json 2> gpurun_out/dbg.err
echo rc=$?; grep -m3 "RuntimeError\|Assertion" gpurun_out/dbg.err | cut -c1-400; head -c 300 gpurun_out/dbg.json
