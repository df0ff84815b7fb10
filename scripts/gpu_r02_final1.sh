#!/bin/bash
# round 2, final 1-GPU evidence: smoke, full GPU test-suite, launch list of the bench command, bench lines of every workload
tag=${1:-r02f}
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$tag.log 2>&1; tail -3 gpurun_out/smoke_$tag.log
timeout 1500 python -m pytest tests -q -m gpu --durations=5 > gpurun_out/pytest_gpu_$tag.log 2>&1; tail -8 gpurun_out/pytest_gpu_$tag.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/launches_$tag.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 2> gpurun_out/bench_${tag}_c5.err | grep "^{" > gpurun_out/bench_${tag}_c5.json
for w in c2 c3 c4 c5b obs; do
  timeout 600 python bench.py --workload $w --steps 5 --warmup 3 --cpu-seconds 6 2> gpurun_out/bench_${tag}_$w.err | grep "^{" > gpurun_out/bench_${tag}_$w.json
done
python - <<PY
import json
for w in ("c5", "c2", "c3", "c4", "c5b", "obs"):
    try:
        d = json.load(open("gpurun_out/bench_${tag}_%s.json" % w))
        print(w, "%.4e" % d["value"], "e2e %.4e" % d["e2e"]["value"], d.get("state_hash"), "frac %.3f" % d["roofline"]["frac"], d["roofline"]["kernel"], "cpu", d.get("cpu_baseline", {}).get("value"))
    except Exception as e:
        print(w, "failed", e)
PY
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 --cpu-seconds 4 2>/dev/null | cut -c1-300
