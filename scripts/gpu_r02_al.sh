#!/bin/bash
# round 2, GPU call AL: partial tiles (any X, Y >= 32, Z >= 32 a multiple of 4 on the tiled kernel) -- parity tests, timing
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_audit.py tests/test_gpu_sweep.py tests/test_gpu_multi.py tests/test_driver.py -q -m gpu -x > gpurun_out/pytest_gpu_r02al.log 2>&1; tail -5 gpurun_out/pytest_gpu_r02al.log
timeout 600 python scripts/exp_time.py 512x512x512 5 default 2>&1 | tee gpurun_out/exp_r02al.txt
for s in 100x100x100 200x200x200 100x100x28 250x250x252; do timeout 300 python scripts/exp_time.py $s 10 default 2>&1 | tee -a gpurun_out/exp_r02al.txt; done
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | grep "^{" | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('c5 10 steps:', d['value'], d['state_hash'], '(expected 022f795848a0c3fd)')"
