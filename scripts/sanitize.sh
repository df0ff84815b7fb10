#!/bin/bash
# compute-sanitizer over the sweep kernels (the reference has no race detection; this is ours).
# usage (under gpurun): bash scripts/sanitize.sh TAG  -> gpurun_out/sanitize_TAG.log
tag=${1:-r02}
out=gpurun_out/sanitize_$tag.log
mkdir -p gpurun_out; : > $out
for tool in memcheck racecheck synccheck; do
  for k in tiled partial cut2 resident colour; do
    echo "=== compute-sanitizer --tool $tool : $k" >> $out
    timeout 600 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_case.py $k 2>&1 | grep -v "^$" | tail -25 >> $out
  done
done
grep -E "^===|ERROR SUMMARY|RACECHECK SUMMARY|ok:" $out
