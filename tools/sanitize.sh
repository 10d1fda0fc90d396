#!/bin/bash
# compute-sanitizer: memcheck, racecheck, synccheck, initcheck on a small
# run of every kernel family.  Logs -> gpurun_out/sanitizer_<tool>.log (copy to profiles/).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 420 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_target.py \
    > gpurun_out/sanitizer_$tool.log 2>&1
  echo "$tool exit $?" >> gpurun_out/sanitizer_$tool.log
  tail -4 gpurun_out/sanitizer_$tool.log
done
