#!/bin/bash
# round 2, GPU call L: tiled-kernel variants (control-warp polling pause, role B draws first, scalar vs packed FP32, loads two pairs ahead)
tag=${1:-r02l}
mkdir -p gpurun_out
timeout 900 python scripts/exp_time.py 512x512x512 5 build/exp/lib_v11.so default build/exp/lib_poll100.so build/exp/lib_drawfirst.so build/exp/lib_scalar.so build/exp/lib_scalar_drawfirst.so build/exp/lib_depth2.so > gpurun_out/exp_$tag.txt 2>&1
timeout 300 python scripts/exp_time.py 128x128x128 20 build/exp/lib_v11.so default build/exp/lib_scalar.so >> gpurun_out/exp_$tag.txt 2>&1
cat gpurun_out/exp_$tag.txt
nvcc -arch=sm_100a -O3 -o /tmp/lds_probe scripts/lds_bcast_probe.cu && /tmp/lds_probe > gpurun_out/lds_probe_$tag.txt 2>&1; cat gpurun_out/lds_probe_$tag.txt
