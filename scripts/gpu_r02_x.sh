#!/bin/bash
# round 2, GPU call X (2 GPUs): thin slabs (512x512x128 on 2 GPUs) with single cross-GPU ingredients switched off (timing only: results are wrong)
mkdir -p gpurun_out; : > gpurun_out/thin_r02x.txt
for lib in default build/exp/lib_noremote.so build/exp/lib_gpufence.so build/exp/lib_gpuacq.so build/exp/lib_allthree.so; do
  if [ $lib = default ]; then unset SN_B200_LIB; else export SN_B200_LIB=$PWD/$lib; fi
  SN_SPIN_TIMEOUT_S=20 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --shape 512,512,128 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2> gpurun_out/bench_r02x_thin.err | grep "^{" > gpurun_out/bench_r02x_thin.json
  python -c "
import json; d=json.load(open('gpurun_out/bench_r02x_thin.json')); print('$lib', 'thin slabs N=2:', d['value'], d['ms_per_step'])" | tee -a gpurun_out/thin_r02x.txt
done
