#!/bin/bash
# compute-sanitizer over tools/sanitize_case.py; usage: tools/sanitize.sh [tool ...]   (default: all four)
mkdir -p gpurun_out
tools_=${@:-memcheck racecheck initcheck synccheck}
out=gpurun_out/r02_sanitizer_$(echo $tools_ | tr ' ' '_').txt
: > $out
for t in $tools_; do
  echo "--- $t" >> $out
  timeout 600 compute-sanitizer --tool $t --print-limit 6 python tools/sanitize_case.py 2>&1 | grep -E "SANITIZE_CASE_DONE|SUMMARY|Uninit|Error|error" | head -12 >> $out
done
cat $out
