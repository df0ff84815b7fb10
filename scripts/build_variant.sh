#!/bin/bash
# scratch: build a kernel-experiment variant of the library: scripts/build_variant.sh NAME -DSN_EXP_...
set -e
cd "$(dirname "$0")/../starrynight_b200/csrc"
name=$1; shift
mkdir -p ../../build/exp
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden "$@" -c sn_lib.cu -o ../../build/exp/sn_lib_$name.o
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../build/exp/lib_$name.so ../../build/exp/sn_lib_$name.o ../../build/csrc/sn_energy_exact.o
echo built build/exp/lib_$name.so
