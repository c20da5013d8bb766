#!/bin/bash
# compute-sanitizer over the smoke run (module network forward + backward, ProgramGenerator passes, one joint-training
# iteration: every product kernel launches at least once); logs land in gpurun_out/, summaries are copied to profiles/
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  PNMN_PG_NOGRAPH=1 timeout 1200 compute-sanitizer --tool $tool --print-limit 20 \
    python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_$tool.log 2>&1
  echo "== $tool: exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|smoke ok|Error|hazard" gpurun_out/sanitizer_$tool.log | head -12
done
