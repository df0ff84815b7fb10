#!/bin/bash
# round 2, GPU call AC (2 GPUs): RDF kernel v3 (FP64 box, 4x8x8 tile, 512 threads) -- parity tests, obs bench at N = 1 and 2
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_observables.py tests/test_gpu_multi.py tests/test_driver.py -q -m gpu -x > gpurun_out/pytest_obs_r02ac.log 2>&1; tail -4 gpurun_out/pytest_obs_r02ac.log
timeout 600 python bench.py --workload obs --steps 5 --warmup 3 2> gpurun_out/bench_r02ac_obs.err | grep "^{" > gpurun_out/bench_r02ac_obs_n1.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload obs --steps 5 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_r02ac_obs2.err | grep "^{" > gpurun_out/bench_r02ac_obs_n2.json
python - <<PY
import json
for n in (1, 2):
    d = json.load(open("gpurun_out/bench_r02ac_obs_n%d.json" % n))
    print(n, d["value"], d["e2e"]["value"], d["parts_ms"], d["roofline"]["achieved"], d["roofline"]["frac"])
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sn_rdf_tiled_kernel -c 1 -f -o gpurun_out/prof_sn_rdf_tiled_kernel_r02ac python scripts/prof_obs.py 128 > gpurun_out/prof_rdf_r02ac.log 2>&1; tail -1 gpurun_out/prof_rdf_r02ac.log
