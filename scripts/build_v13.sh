#!/bin/bash
# scratch: build the v13 experiment (sources under build/v13/csrc, made from profiles/experiments/r02_tiled_v13_segment_per_thread.patch)
# as build/exp/lib_v13$1.so; extra nvcc flags after the name suffix
set -e
cd "$(dirname "$0")/../build/v13/csrc"
sfx=$1; shift || true
mkdir -p ../../exp
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -I../../../include "$@" -c sn_lib.cu -o ../../exp/sn_lib_v13$sfx.o
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -fmad=false -I../../../include -c sn_energy_exact.cu -o ../../exp/sn_exact_v13.o
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../exp/lib_v13$sfx.so ../../exp/sn_lib_v13$sfx.o ../../exp/sn_exact_v13.o
echo built build/exp/lib_v13$sfx.so
