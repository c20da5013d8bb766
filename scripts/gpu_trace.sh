#!/bin/bash
# per-task timeline of the persistent executor; PNMN_EXEC_DBG variants are timing experiments (results are wrong on purpose)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for dbg in ${DBGS:-0}; do
  echo "#### PNMN_EXEC_DBG=$dbg"
  PNMN_EXEC_DBG=$dbg timeout 600 python scripts/trace_exec.py 2>&1 | grep -E "conv n_samp|==|sum|epilogue" | grep -E "==|mmas/tile=  72|epilogue" | cut -c1-330 | tee gpurun_out/trace_dbg$dbg.txt
done
