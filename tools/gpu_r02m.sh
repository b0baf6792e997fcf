#!/bin/bash
# round 2, call M: deferred token emit; caller's stream test
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu -k "manifest or fresh or mixed or config or many_streams or extreme or campaign or callers_stream or full_size" 2>&1 | tail -4
for wl in vp8_1080p_q75_m4_1part_simple_rgba vp8_1080p_q75_m4_8part_normal_rgba vp8_256x256_q80_rgbA; do
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --distinct 64 --e2e-steps 0 --no-others --workload $wl > gpurun_out/r02m_bench_$wl.json 2>&1
python - <<PY
import json
d=json.loads(open('gpurun_out/r02m_bench_$wl.json').read().strip().splitlines()[-1])
print("$wl value",d["value"],"ms",d["ms_per_step"],{k:v["ms"] for k,v in d["kernels"].items()})
PY
done
