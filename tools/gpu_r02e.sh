#!/bin/bash
# round 2, call E: fp parser v2 timing, e2e host-side timing, ncu source view of the fp parser
mkdir -p gpurun_out
python bench.py --steps 4 --warmup 3 --no-cpu-baseline --distinct 64 --no-others > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err; tail -3 gpurun_out/r02e_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02e_bench.json').read().strip().splitlines()[-1])
print("value",d["value"],"ms",d["ms_per_step"],"with_h2d",d["value_with_h2d"])
print("e2e",d["e2e"])
print({k:v["ms"] for k,v in d["kernels"].items()}, d["parse"]["cycles_per_decode"] if d["parse"] else None)
PY
CMD="python bench.py --distinct 32 --steps 1 --warmup 3 --e2e-steps 0 --no-cpu-baseline --no-others"
ncu --set full --clock-control none --import-source on -k regex:k_parse_tokens_fp -s 3 -c 1 -o gpurun_out/r02e_tokens_fp $CMD > gpurun_out/r02e_ncu.log 2>&1
tail -3 gpurun_out/r02e_ncu.log
