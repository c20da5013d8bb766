#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() {
  name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 12 --warmup 8 --no-cpu-baseline --no-extras > gpurun_out/bis_$name.json 2> gpurun_out/bis_$name.err
  rc=$?
  echo "$name rc=$rc $(head -c 230 gpurun_out/bis_$name.json | cut -c60-230) $(grep -m1 -o 'Assertion.*failed' gpurun_out/bis_$name.err | head -1)"
}
run noprecompile PNMN_NO_PRECOMPILE=1 PNMN_NO_GRAD_OVERLAP=1
run unfused PNMN_CLASSIFIER_FUSED=0 PNMN_NO_GRAD_OVERLAP=1
run default PNMN_X=1
