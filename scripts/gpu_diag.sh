#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for impl in simt tc; do PNMN_CONV_IMPL=$impl timeout 600 python scripts/diag_forward.py 2>&1 | tail -70; done | tee gpurun_out/diag.log
