#!/bin/bash
# round 2, call K: input ring (cp.async.bulk) on / off
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu -k "manifest or fresh or mixed or config3 or many_streams" 2>&1 | tail -3
for r in 1 0; do
for wl in vp8_1080p_q75_m4_1part_simple_rgba vp8_1080p_q75_m4_8part_normal_rgba vp8_256x256_q80_rgbA; do
WEBP_B200_TOKEN_RING=$r timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --distinct 64 --e2e-steps 0 --no-others --workload $wl > gpurun_out/r02k_bench_ring${r}_$wl.json 2>&1
python - <<PY
import json
d=json.loads(open('gpurun_out/r02k_bench_ring${r}_$wl.json').read().strip().splitlines()[-1])
print("ring=$r $wl value",d["value"],"ms",d["ms_per_step"],{k:v["ms"] for k,v in d["kernels"].items()})
PY
done
done
