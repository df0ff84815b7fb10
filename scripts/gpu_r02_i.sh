#!/bin/bash
# round 2, GPU call I: time decomposition of the v11 tiled kernel + ncu --set full capture
tag=${1:-r02i}
mkdir -p gpurun_out
timeout 600 python scripts/exp_time.py 512x512x512 5 default build/exp/lib_noload.so build/exp/lib_nofp.so build/exp/lib_nochain.so build/exp/lib_noga.so build/exp/lib_nogb.so > gpurun_out/exp_$tag.txt 2>&1; cat gpurun_out/exp_$tag.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sn_tiled_kernel -s 1 -c 1 -f -o gpurun_out/prof_tiled_$tag python scripts/prof_one.py 512x512x512 2 > gpurun_out/prof_tiled_$tag.log 2>&1; tail -2 gpurun_out/prof_tiled_$tag.log
ls -la gpurun_out/*.ncu-rep
