#!/bin/bash
# round 2, GPU call H: split-layout tiled kernel -- parity tests, time per sweep, bench line (state_hash must not change)
tag=${1:-r02h}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x --durations=5 > gpurun_out/pytest_gpu_$tag.log 2>&1; tail -12 gpurun_out/pytest_gpu_$tag.log
timeout 300 python scripts/exp_time.py 512x512x512 5 default > gpurun_out/exp_$tag.txt 2>&1; cat gpurun_out/exp_$tag.txt
timeout 300 python scripts/exp_time.py 128x128x128 20 default >> gpurun_out/exp_$tag.txt 2>&1; tail -1 gpurun_out/exp_$tag.txt
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_$tag.err | grep "^{" > gpurun_out/bench_${tag}_c5.json
python - <<PY
import json
d = json.load(open("gpurun_out/bench_${tag}_c5.json"))
print(d["value"], d["e2e"]["value"], d["state_hash"], d["accept_ratio"], d["roofline"]["frac"])
PY
tail -3 gpurun_out/bench_$tag.err
