#!/bin/bash
TOOL=${1:-memcheck}
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python tools/sanitize_target.py > gpurun_out/san_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/san_plain.log; exit 1; }
timeout 1500 compute-sanitizer --tool $TOOL --launch-timeout 0 --print-limit 20 python tools/sanitize_target.py > gpurun_out/r2_sanitizer_$TOOL.log 2>&1
echo "sanitizer $TOOL exit $?"
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize target ok|Error|hazard" gpurun_out/r2_sanitizer_$TOOL.log | head -20
