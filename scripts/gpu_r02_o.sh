#!/bin/bash
# round 2, GPU call O: observable kernels v2 (constant-memory RDF table, conflict-free FP64 boxes, tables in shared memory): parity tests, bench line, ncu
tag=${1:-r02o}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_observables.py tests/test_gpu_multi.py tests/test_driver.py -q -m gpu -x > gpurun_out/pytest_obs_$tag.log 2>&1; tail -4 gpurun_out/pytest_obs_$tag.log
timeout 600 python bench.py --workload obs --steps 3 --warmup 2 2> gpurun_out/bench_${tag}_obs.err | grep "^{" > gpurun_out/bench_${tag}_obs.json
python - <<PY
import json
d = json.load(open("gpurun_out/bench_${tag}_obs.json"))
print(d["value"], d["e2e"]["value"], d["parts_ms"], d["roofline"]["achieved"], d["roofline"]["peak"])
PY
tail -3 gpurun_out/bench_${tag}_obs.err
for k in sn_rdf_tiled_kernel sn_potential_tiled_kernel sn_efield_tiled_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -f -o gpurun_out/prof_${k}_$tag python scripts/prof_obs.py 128 > gpurun_out/prof_${k}_$tag.log 2>&1; tail -1 gpurun_out/prof_${k}_$tag.log
done
