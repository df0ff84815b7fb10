#!/bin/bash
# round 2, GPU call A: test-suite (audit, KAT, slabs on one device), sanitizer, kernel time decomposition, FFMA2 probe
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x --durations=15 > gpurun_out/pytest_gpu_r02a.log 2>&1; tail -25 gpurun_out/pytest_gpu_r02a.log
./scripts/ffma2_probe > gpurun_out/ffma2_probe.txt 2>&1; cat gpurun_out/ffma2_probe.txt
timeout 600 python scripts/exp_time.py 512x512x512 5 default build/exp/lib_noload.so build/exp/lib_nofp.so build/exp/lib_nochain.so build/exp/lib_noga.so build/exp/lib_nogb.so > gpurun_out/exp_r02a.txt 2>&1; cat gpurun_out/exp_r02a.txt
timeout 1500 bash scripts/sanitize.sh r02a
