#!/bin/bash
# round 2, GPU call B: full test-suite with the new host path, sanitizer on the tiled kernel, bench lines of every workload
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --durations=10 > gpurun_out/pytest_gpu_r02b.log 2>&1; tail -25 gpurun_out/pytest_gpu_r02b.log
for k in tiled; do for tool in racecheck synccheck; do echo "== $tool $k"; timeout 600 compute-sanitizer --tool $tool --print-limit 5 python scripts/sanitize_case.py $k 2>&1 | tail -12; done; done > gpurun_out/sanitize_r02b.log 2>&1; grep -E "==|SUMMARY|ok:" gpurun_out/sanitize_r02b.log
timeout 900 python bench.py --steps 3 --warmup 3 2> gpurun_out/bench_r02b.err | grep "^{" > gpurun_out/bench_r02b_c5.json; head -c 3000 gpurun_out/bench_r02b_c5.json; tail -5 gpurun_out/bench_r02b.err
for wl in c2 c3 c4; do
  timeout 600 python bench.py --workload $wl --steps 3 --warmup 3 2> gpurun_out/bench_r02b_$wl.err | grep "^{" > gpurun_out/bench_r02b_$wl.json; head -c 1500 gpurun_out/bench_r02b_$wl.json; echo; tail -3 gpurun_out/bench_r02b_$wl.err
done
