#!/bin/bash
# round 2, GPU call G (2 GPUs): slab parity tests across real GPUs, 2-GPU driver test, bench N=1 and N=2 with state_hash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/multi_r02g.log
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_driver.py -q -m gpu -rs --durations=8 >> gpurun_out/multi_r02g.log 2>&1; tail -25 gpurun_out/multi_r02g.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_r02g_n1.err | grep "^{" > gpurun_out/bench_r02g_n1.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_r02g_n2.err | grep "^{" > gpurun_out/bench_r02g_n2.json
python - <<'PY'
import json
for n in (1, 2):
    try:
        d = json.load(open(f"gpurun_out/bench_r02g_n{n}.json"))
        print(n, d["value"], d["e2e"]["value"], d.get("state_hash"), d.get("accept_ratio"))
    except Exception as e:
        print(n, "failed", e)
PY
tail -5 gpurun_out/bench_r02g_n2.err
