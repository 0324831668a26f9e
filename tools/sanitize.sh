#!/bin/bash
# compute-sanitizer passes over tools/sanitize_driver.py.   bash tools/sanitize.sh <tag>
TAG=${1:-sanitize}; OUT=gpurun_out/$TAG; mkdir -p $OUT
for tool in ${TOOLS:-memcheck racecheck synccheck initcheck}; do
  echo "== $tool"
  timeout ${TOOL_TIMEOUT:-900} compute-sanitizer --tool $tool --error-exitcode 99 python tools/sanitize_driver.py > $OUT/$tool.log 2>&1
  echo "rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize driver ok|Error|error" $OUT/$tool.log | head -8
done
