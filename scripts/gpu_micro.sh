#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 scripts/microbench/bulk_stream 2>&1 | tee gpurun_out/bulk_stream2.txt | grep -v "SMs   8 " | tail -80
