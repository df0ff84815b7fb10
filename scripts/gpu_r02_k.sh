#!/bin/bash
# round 2, GPU call K: packed-FP32 (FFMA2 / FADD2) tiled kernel vs v11 -- timing A/B, state hash, audit + sweep parity tests, ncu capture
tag=${1:-r02k}
mkdir -p gpurun_out
timeout 600 python scripts/exp_time.py 512x512x512 5 build/exp/lib_v11.so default > gpurun_out/exp_$tag.txt 2>&1
timeout 300 python scripts/exp_time.py 128x128x128 20 build/exp/lib_v11.so default >> gpurun_out/exp_$tag.txt 2>&1
cat gpurun_out/exp_$tag.txt
timeout 1200 python -m pytest tests/test_gpu_audit.py tests/test_gpu_sweep.py -q -m gpu -x > gpurun_out/pytest_gpu_$tag.log 2>&1; tail -5 gpurun_out/pytest_gpu_$tag.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_$tag.err | grep "^{" > gpurun_out/bench_${tag}_c5.json
python - <<PY
import json
d = json.load(open("gpurun_out/bench_${tag}_c5.json"))
print(d["value"], d["e2e"]["value"], d["state_hash"], d["accept_ratio"], d["roofline"]["frac"])
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sn_tiled_kernel -s 1 -c 1 -f -o gpurun_out/prof_tiled_$tag python scripts/prof_one.py 512x512x512 2 > gpurun_out/prof_tiled_$tag.log 2>&1; tail -2 gpurun_out/prof_tiled_$tag.log
