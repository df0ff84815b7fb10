#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --durations=8 > gpurun_out/pytest_gpu_r02d.log 2>&1; tail -22 gpurun_out/pytest_gpu_r02d.log
timeout 900 bash scripts/sanitize.sh r02d
