#!/bin/bash
# One-GPU evidence run: GPU test-suite, bench line, ncu launch list of the bench command, ncu --set full of the sweep kernel.
# usage (under gpurun): bash scripts/gpu_profile.sh TAG
tag=${1:-r01}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_$tag.log 2>&1; tail -3 gpurun_out/pytest_gpu_$tag.log
timeout 600 python bench.py --steps 3 --warmup 3 2> gpurun_out/bench_$tag.err | grep "^{" > gpurun_out/bench_${tag}_n1.json; head -c 600 gpurun_out/bench_${tag}_n1.json; echo
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/launches_$tag.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sn_tiled_kernel -s 1 -c 1 -f -o gpurun_out/prof_tiled_$tag \
    python bench.py --steps 1 --warmup 1 --sweeps-per-step 2 --no-cpu-baseline > gpurun_out/prof_tiled_$tag.log 2>&1
ls -la gpurun_out | tail -8
