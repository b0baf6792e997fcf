#!/bin/bash
# round 2, call F: fp parser v3 (loads after the bit), status words without the copy engine
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "manifest or fresh or full_size or mixed or config or submit or alpha_batch or extreme" 2>&1 | tail -5 > gpurun_out/r02f_pytest_gpu.log; cat gpurun_out/r02f_pytest_gpu.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --distinct 64 --no-others > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err; tail -3 gpurun_out/r02f_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02f_bench.json').read().strip().splitlines()[-1])
print("value",d["value"],"ms",d["ms_per_step"],"with_h2d",d["value_with_h2d"])
print("e2e",d["e2e"])
print({k:v["ms"] for k,v in d["kernels"].items()}, d["parse"]["cycles_per_decode"] if d["parse"] else None)
PY
CMD="python bench.py --distinct 32 --steps 1 --warmup 3 --e2e-steps 0 --no-cpu-baseline --no-others"
ncu --set full --clock-control none --import-source on -k regex:k_parse_tokens_fp -s 3 -c 1 -o gpurun_out/r02f_tokens_fp $CMD > gpurun_out/r02f_ncu.log 2>&1
tail -2 gpurun_out/r02f_ncu.log
