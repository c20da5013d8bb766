#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_nmn_gpu.py tests/test_kernels_gpu.py -q 2>&1 | grep -E "passed|failed|FAILED" | tail -5
timeout 600 python scripts/trace_exec.py 2>&1 | grep -E "conv n_samp|elt op|==|sum" | cut -c1-330 | tee gpurun_out/trace.txt
