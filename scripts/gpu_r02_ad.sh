#!/bin/bash
# round 2, GPU call AD: hand-over slots alternate (racecheck found the control warp overwriting the slot early) -- sanitizer, tests, timing
mkdir -p gpurun_out
bash scripts/sanitize.sh r02final
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu_r02ad.log 2>&1; tail -3 gpurun_out/pytest_gpu_r02ad.log
timeout 600 python scripts/exp_time.py 512x512x512 5 default 2>&1 | tee gpurun_out/exp_r02ad.txt
timeout 300 python scripts/exp_time.py 128x128x128 20 default 2>&1 | tee -a gpurun_out/exp_r02ad.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | grep "^{" | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('c5 10 steps:', d['value'], d['state_hash'], '(expected 022f795848a0c3fd)')"
