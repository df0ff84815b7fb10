#!/bin/bash
# round 2, GPU call AI (2 GPUs): species flags resolved through an event on the upload -- full test-suite, e2e at N = 1 and 2
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu_r02ai.log 2>&1; tail -3 gpurun_out/pytest_gpu_r02ai.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2> gpurun_out/bench_r02ai_c5_n1.err | grep "^{" > gpurun_out/bench_r02ai_c5_n1.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline 2> gpurun_out/bench_r02ai_c5_n2.err | grep "^{" > gpurun_out/bench_r02ai_c5_n2.json
timeout 600 python bench.py --workload c4 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | grep "^{" > gpurun_out/bench_r02ai_c4_n1.json
python - <<PY
import json
for f in ("c5_n1", "c5_n2", "c4_n1"):
    d = json.load(open("gpurun_out/bench_r02ai_%s.json" % f))
    print(f, d["value"], "e2e", d["e2e"]["value"], "serial", d["e2e"]["serial"]["value"], d["state_hash"])
PY
