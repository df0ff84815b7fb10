#!/bin/bash
# round 2, GPU call R (8 GPUs): bench c5 / c5b / obs at N = 8 (state_hash must equal the 1-GPU string)
n=${1:-8}; tag=${2:-r02r}
mkdir -p gpurun_out
for w in c5 c5b obs; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --workload $w --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_${tag}_${w}_n$n.err | grep "^{" > gpurun_out/bench_${tag}_${w}_n$n.json
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_${tag}_${w}_n$n.json"))
    print("$w", $n, d["value"], d["e2e"]["value"], d.get("state_hash"), d.get("accept_ratio"), d["roofline"]["frac"], d.get("parts_ms"))
except Exception as e:
    print("$w", $n, "failed", e)
PY
  grep -v "OMP_NUM\|^\*\*\*" gpurun_out/bench_${tag}_${w}_n$n.err | tail -3
done
