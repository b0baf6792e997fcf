#!/bin/bash
# GPU box: one short bench line per workload (the other BASELINE configs' shapes + lossless), into gpurun_out/${TAG}_workloads.log
TAG=${1:-r01s}
mkdir -p gpurun_out
: > gpurun_out/${TAG}_workloads.log
run() { python bench.py --steps 2 --e2e-steps 1 --no-cpu-baseline "$@" 2> gpurun_out/wl.err | grep '^{' | tail -1 >> gpurun_out/${TAG}_workloads.log || tail -3 gpurun_out/wl.err; }
run --workload vp8_1080p_q75_m4_8part_normal_rgba --distinct 64
run --workload vp8_256x256_q80_rgbA --distinct 512
run --workload vp8_4096x4096_q90_alpha_rgba --distinct 8 --batch 128
run --workload vp8l_1080p_lossless_rgba --distinct 16 --batch 512
python - gpurun_out/${TAG}_workloads.log <<'PY'
import json,sys
for l in open(sys.argv[1]):
    d=json.loads(l); print(d["config"]["workload"], d["config"]["batch_per_gpu"], d["value"], "Mpix/s", d["ms_per_step"], "ms", {k:v["ms"] for k,v in d["kernels"].items()}, "e2e", d["e2e"] and d["e2e"]["value"])
PY
