#!/bin/bash
# run under gpurun: tools/sanitize_tc.sh <tag>
TAG=${1:-r02}
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool python tools/sanitize_tc.py > gpurun_out/${tool}_tc_${TAG}.txt 2>&1
  echo "== $tool: $(grep -c 'Error\|Hazard\|error:' gpurun_out/${tool}_tc_${TAG}.txt) flagged lines"; tail -4 gpurun_out/${tool}_tc_${TAG}.txt
done
