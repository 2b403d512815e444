#!/bin/bash
# compute-sanitizer passes over the tensor-path kernels on small shapes.
# Usage (on the GPU box):  bash tools/sanitize.sh [outdir]     -> <outdir>/sanitizer_{memcheck,racecheck,synccheck}.log
out=${1:-gpurun_out}
mkdir -p "$out"
for tool in memcheck synccheck racecheck; do
  echo "== compute-sanitizer --tool $tool" | tee "$out/sanitizer_$tool.log"
  timeout ${SAN_TIMEOUT:-420} compute-sanitizer --tool $tool --print-limit 20 --launch-timeout 0 \
      python tools/sanitize_driver.py >> "$out/sanitizer_$tool.log" 2>&1
  echo "exit code $?" >> "$out/sanitizer_$tool.log"
  tail -4 "$out/sanitizer_$tool.log"
done
