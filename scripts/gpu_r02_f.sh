#!/bin/bash
mkdir -p gpurun_out
for k in sn_rdf_tiled_kernel sn_potential_tiled_kernel; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -f -o gpurun_out/prof_$k python scripts/prof_obs.py 128 > gpurun_out/prof_$k.log 2>&1; tail -2 gpurun_out/prof_$k.log
done
ls -la gpurun_out/*.ncu-rep
