#!/bin/bash
# round 2, GPU call Y (2 GPUs): one fence per publication -- slab parity tests, thin-slab and cube timing with state hash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_driver.py tests/test_gpu_sweep.py -q -m gpu -x > gpurun_out/multi_r02y.log 2>&1; tail -4 gpurun_out/multi_r02y.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --shape 512,512,128 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2> gpurun_out/bench_r02y_thin.err | grep "^{" > gpurun_out/bench_r02y_thin.json
python -c "
import json; d=json.load(open('gpurun_out/bench_r02y_thin.json')); print('thin slabs N=2:', d['value'], d['ms_per_step'], d['state_hash'])"
timeout 300 python bench.py --shape 512,512,128 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | grep "^{" | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('N=1 512x512x128:', d['value'], d['ms_per_step'], d['state_hash'])"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_r02y_c5_n2.err | grep "^{" > gpurun_out/bench_r02y_c5_n2.json
python -c "
import json; d=json.load(open('gpurun_out/bench_r02y_c5_n2.json')); print('c5 N=2:', d['value'], d['e2e']['value'], d['state_hash'], '(N=1, 10 steps: 022f795848a0c3fd)')"
