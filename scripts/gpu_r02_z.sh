#!/bin/bash
# round 2, GPU call Z: third worker role (two B roles) for lattices without species -- timing of the split variants, parity tests, bench line
tag=${1:-r02z}
mkdir -p gpurun_out
timeout 900 python scripts/exp_time.py 512x512x512 5 build/exp/lib_nb1.so default build/exp/lib_b1_2.so build/exp/lib_b1_4.so build/exp/lib_b1_3gf.so > gpurun_out/exp_$tag.txt 2>&1
timeout 300 python scripts/exp_time.py 128x128x128 20 build/exp/lib_nb1.so default >> gpurun_out/exp_$tag.txt 2>&1
cat gpurun_out/exp_$tag.txt
timeout 1500 python -m pytest tests/test_gpu_audit.py tests/test_gpu_sweep.py tests/test_gpu_multi.py -q -m gpu -x > gpurun_out/pytest_gpu_$tag.log 2>&1; tail -4 gpurun_out/pytest_gpu_$tag.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_$tag.err | grep "^{" > gpurun_out/bench_${tag}_c5.json
python - <<PY
import json
d = json.load(open("gpurun_out/bench_${tag}_c5.json"))
print(d["value"], d["e2e"]["value"], d["state_hash"], d["accept_ratio"], d["energy_per_site"], d["roofline"]["frac"])
PY
tail -3 gpurun_out/bench_$tag.err
