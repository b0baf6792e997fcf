#!/bin/bash
# round 2, call I: K3 with its HBM loads two macroblocks ahead; mapping tests; animation batch
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu -k "manifest or fresh or full_size or mixed or config or extreme or campaign or anim or token_mapping or row_bands or dither or crop" 2>&1 | tail -6 > gpurun_out/r02i_pytest_gpu.log; cat gpurun_out/r02i_pytest_gpu.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --distinct 64 --e2e-steps 0 --no-others > gpurun_out/r02i_bench.json 2> gpurun_out/r02i_bench.err; tail -3 gpurun_out/r02i_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02i_bench.json').read().strip().splitlines()[-1])
print("value",d["value"],"ms",d["ms_per_step"],{k:v["ms"] for k,v in d["kernels"].items()})
PY
