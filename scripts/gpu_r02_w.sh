#!/bin/bash
# round 2, GPU call W: where do thin slabs lose 12 %?  one GPU, periodic handles of 64 / 128 / 512 planes (same x, y)
mkdir -p gpurun_out
for shp in 512,512,64 512,512,128 512,512,512; do
  timeout 600 python bench.py --shape $shp --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | grep "^{" | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('N=1 $shp:', d['value'], d['ms_per_step'])"
done | tee gpurun_out/thin_r02w.txt
