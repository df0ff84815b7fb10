#!/bin/bash
# round 2, GPU call P (N GPUs): slab parity tests across real GPUs, driver test, bench at N with state_hash (c5, obs)
n=${1:-2}; tag=${2:-r02p}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/multi_${tag}_n$n.log
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_driver.py -q -m gpu -rs --durations=5 >> gpurun_out/multi_${tag}_n$n.log 2>&1; tail -12 gpurun_out/multi_${tag}_n$n.log
for w in c5 obs c5b; do
  if [ $w = c5b ] && [ $n -lt 8 ]; then continue; fi
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --workload $w --steps 3 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_${tag}_${w}_n$n.err | grep "^{" > gpurun_out/bench_${tag}_${w}_n$n.json
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_${tag}_${w}_n$n.json"))
    print("$w", $n, d["value"], d["e2e"]["value"], d.get("state_hash"), d.get("accept_ratio"), d["roofline"]["frac"])
except Exception as e:
    print("$w", $n, "failed", e)
PY
  tail -2 gpurun_out/bench_${tag}_${w}_n$n.err
done
