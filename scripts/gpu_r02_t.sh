#!/bin/bash
# round 2, GPU call T (2 GPUs): (a) time decomposition of v12 on one GPU; (b) thin slabs (512x512x128 on 2 GPUs = the per-GPU shape of the 8-GPU run) with two polling pauses; (c) N=1 line with 10 steps for the hash of the 8-GPU run
tag=${1:-r02t}
mkdir -p gpurun_out
timeout 600 python scripts/exp_time.py 512x512x512 5 default build/exp/lib_noload.so build/exp/lib_nofp.so build/exp/lib_nochain.so > gpurun_out/exp_$tag.txt 2>&1; cat gpurun_out/exp_$tag.txt
for lib in default build/exp/lib_poll100.so; do
  if [ $lib = default ]; then unset SN_B200_LIB; else export SN_B200_LIB=$PWD/$lib; fi
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --shape 512,512,128 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2> gpurun_out/bench_${tag}_thin.err | grep "^{" > gpurun_out/bench_${tag}_thin.json
  python -c "
import json; d=json.load(open('gpurun_out/bench_${tag}_thin.json')); print('$lib', 'thin slabs N=2:', d['value'], d['ms_per_step'])"
done
unset SN_B200_LIB
timeout 600 python bench.py --shape 512,512,128 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | grep "^{" | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('N=1 512x512x128:', d['value'], d['ms_per_step'])"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | grep "^{" > gpurun_out/bench_${tag}_c5_n1_10steps.json; python -c "
import json; d=json.load(open('gpurun_out/bench_${tag}_c5_n1_10steps.json')); print('N=1 c5 10 steps:', d['value'], d['state_hash'], d['accept_ratio'])"
