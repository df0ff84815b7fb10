#!/bin/bash
# round 2, GPU call N: observables -- bench line (--workload obs), ncu --set full of the three tiled stencils at 128^3
tag=${1:-r02n}
mkdir -p gpurun_out
timeout 600 python bench.py --workload obs --steps 3 --warmup 2 2> gpurun_out/bench_${tag}_obs.err | grep "^{" > gpurun_out/bench_${tag}_obs.json
python - <<PY
import json
d = json.load(open("gpurun_out/bench_${tag}_obs.json"))
print(d["value"], d["e2e"]["value"], d["parts_ms"], d["roofline"]["achieved"], d["roofline"]["peak"], d.get("cpu_baseline"))
PY
tail -3 gpurun_out/bench_${tag}_obs.err
for k in sn_rdf_tiled_kernel sn_potential_tiled_kernel sn_efield_tiled_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -f -o gpurun_out/prof_${k}_$tag python scripts/prof_obs.py 128 > gpurun_out/prof_${k}_$tag.log 2>&1; tail -1 gpurun_out/prof_${k}_$tag.log
done
ls -la gpurun_out/*.ncu-rep
