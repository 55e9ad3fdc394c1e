#!/bin/bash
# compute-sanitizer over tests/tools/sanitize_case.py; TOOL=memcheck|racecheck|synccheck|initcheck (one tool per gpurun call)
mkdir -p gpurun_out
TOOL=${TOOL:-memcheck}
timeout 120 python tests/tools/sanitize_case.py > gpurun_out/sanitize_plain.log 2>&1 || { tail -5 gpurun_out/sanitize_plain.log; exit 1; }
tail -1 gpurun_out/sanitize_plain.log | cut -c1-200
timeout 1500 compute-sanitizer --tool $TOOL --print-limit 30 python tests/tools/sanitize_case.py > gpurun_out/sanitize_$TOOL.log 2>&1
echo "rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize case ok|Error|hazard" gpurun_out/sanitize_$TOOL.log | head -20
