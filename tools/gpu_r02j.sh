#!/bin/bash
# round 2, call J: compressed bytes through per-stream shared-memory rings filled by cp.async.bulk
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "manifest or fresh or full_size or mixed or config or extreme or campaign or many_streams or alpha_batch" 2>&1 | tail -6 > gpurun_out/r02j_pytest_gpu.log; cat gpurun_out/r02j_pytest_gpu.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --distinct 64 --e2e-steps 0 --no-others > gpurun_out/r02j_bench.json 2> gpurun_out/r02j_bench.err; tail -3 gpurun_out/r02j_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02j_bench.json').read().strip().splitlines()[-1])
print("value",d["value"],"ms",d["ms_per_step"],{k:v["ms"] for k,v in d["kernels"].items()})
PY
for wl in vp8_1080p_q75_m4_8part_normal_rgba vp8_256x256_q80_rgbA; do
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --distinct 64 --e2e-steps 0 --no-others --workload $wl > gpurun_out/r02j_bench_$wl.json 2>&1
python - <<PY
import json
d=json.loads(open('gpurun_out/r02j_bench_$wl.json').read().strip().splitlines()[-1])
print("$wl value",d["value"],"ms",d["ms_per_step"],{k:v["ms"] for k,v in d["kernels"].items()})
PY
done
