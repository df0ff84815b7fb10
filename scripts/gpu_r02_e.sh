#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x --durations=8 > gpurun_out/pytest_gpu_r02e.log 2>&1; tail -22 gpurun_out/pytest_gpu_r02e.log
timeout 300 python scripts/time_observables.py > gpurun_out/observables_r02e.txt 2>&1; cat gpurun_out/observables_r02e.txt
