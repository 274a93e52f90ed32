#!/bin/bash
# compute-sanitizer over the small all-kernel workload; summaries land in gpurun_out/ (copy to profiles/).
R=${1:-r02}
mkdir -p gpurun_out
compute-sanitizer --version > gpurun_out/${R}_sanitize_version.log 2>&1; echo "version rc=$?"; head -3 gpurun_out/${R}_sanitize_version.log
timeout 600 python tools/sanitize_workload.py > gpurun_out/${R}_sanitize_plain.log 2>&1; echo "plain rc=$?"; tail -3 gpurun_out/${R}_sanitize_plain.log
for TOOL in memcheck racecheck synccheck; do
  timeout 1200 compute-sanitizer --tool $TOOL --print-limit 20 --error-exitcode 7 python tools/sanitize_workload.py > gpurun_out/${R}_sanitize_$TOOL.log 2>&1
  echo "$TOOL rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|^ok |done|Error|hazard|rror" gpurun_out/${R}_sanitize_$TOOL.log | tail -14
  head -c 20000 gpurun_out/${R}_sanitize_$TOOL.log > gpurun_out/${R}_sanitize_$TOOL.head; tail -c 20000 gpurun_out/${R}_sanitize_$TOOL.log > gpurun_out/${R}_sanitize_$TOOL.tail
  rm -f gpurun_out/${R}_sanitize_$TOOL.log
done
