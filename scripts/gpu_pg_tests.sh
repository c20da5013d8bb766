#!/bin/bash
# seq2seq bring-up: CUDA-core twins first (logic), then the tcgen05 kernels (descriptors)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
PNMN_PG_SIMT=1 timeout 600 python -m pytest tests/test_seq2seq_gpu.py -q -s 2>&1 | grep -vE "^$" | tail -70 | cut -c1-220 > gpurun_out/pg_simt.txt
echo "=== SIMT twins"; cat gpurun_out/pg_simt.txt
timeout 600 python -m pytest tests/test_seq2seq_gpu.py -q -s 2>&1 | grep -vE "^$" | tail -70 | cut -c1-220 > gpurun_out/pg_tc.txt
echo "=== tcgen05"; cat gpurun_out/pg_tc.txt
