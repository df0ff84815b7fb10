#!/bin/bash
# debug: 2-slab tests on one device; synccheck with barrier.sync
mkdir -p gpurun_out
export SN_SPIN_TIMEOUT_S=6
timeout 600 python -m pytest tests/test_gpu_multi.py -q -x -k "two_slabs" -s > gpurun_out/multi_r02c.log 2>&1; tail -30 gpurun_out/multi_r02c.log
CUDA_DEVICE_MAX_CONNECTIONS=32 timeout 600 python -m pytest tests/test_gpu_multi.py -q -x -k "two_slabs" -s > gpurun_out/multi_r02c_conn32.log 2>&1; tail -5 gpurun_out/multi_r02c_conn32.log
timeout 600 compute-sanitizer --tool synccheck --print-limit 3 python scripts/sanitize_case.py tiled 2>&1 | tail -30 > gpurun_out/sync_r02c.log; cat gpurun_out/sync_r02c.log | cut -c1-200
