#!/bin/bash
# round 2, GPU call J: confirm restored tree -- full GPU test-suite, default bench line, other workloads, observable timings
tag=${1:-r02j}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x --durations=5 > gpurun_out/pytest_gpu_$tag.log 2>&1; tail -6 gpurun_out/pytest_gpu_$tag.log
timeout 600 python bench.py --steps 3 --warmup 3 2> gpurun_out/bench_${tag}_c5.err | grep "^{" > gpurun_out/bench_${tag}_c5.json
for w in c2 c3 c4 c5b; do
  timeout 600 python bench.py --workload $w --steps 3 --warmup 3 --cpu-seconds 6 2> gpurun_out/bench_${tag}_$w.err | grep "^{" > gpurun_out/bench_${tag}_$w.json
done
python - <<PY
import json
for w in ("c5", "c2", "c3", "c4", "c5b"):
    try:
        d = json.load(open("gpurun_out/bench_${tag}_%s.json" % w))
        print(w, "%.4e" % d["value"], "e2e %.4e" % d["e2e"]["value"], "serial %.4e" % d["e2e"]["serial"]["value"], d["state_hash"], "%.4f" % d["accept_ratio"], "frac %.3f" % d["roofline"]["frac"], d["roofline"]["kernel"], "cpu", d.get("cpu_baseline", {}).get("value"))
    except Exception as e:
        print(w, "failed", e)
PY
tail -3 gpurun_out/bench_${tag}_c5.err
timeout 300 python scripts/time_observables.py 256 > gpurun_out/obs_$tag.txt 2>&1; cat gpurun_out/obs_$tag.txt
timeout 300 python scripts/exp_time.py 128x128x128 20 default >> gpurun_out/exp_$tag.txt 2>&1
timeout 300 python scripts/exp_time.py 256x256x256 10 default >> gpurun_out/exp_$tag.txt 2>&1
timeout 300 python scripts/exp_time.py 96x96x96 20 default >> gpurun_out/exp_$tag.txt 2>&1
cat gpurun_out/exp_$tag.txt
