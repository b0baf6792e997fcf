#!/bin/bash
# round 2, call D: first run of the fp token parser + the copier thread
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "manifest or fresh or full_size or mixed or config or submit or waves or resident or extreme or campaign" 2>&1 | tail -8 > gpurun_out/r02d_pytest_gpu.log; cat gpurun_out/r02d_pytest_gpu.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --distinct 64 > gpurun_out/r02d_bench.json 2> gpurun_out/r02d_bench.err; tail -3 gpurun_out/r02d_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02d_bench.json').read().strip().splitlines()[-1])
print("value",d["value"],"ms",d["ms_per_step"],"with_h2d",d["value_with_h2d"])
e=d["e2e"]; print("e2e",e["value"],e["ms_per_step"],"blocking",e["blocking_ms"],"floor",e["d2h_floor_ms"],e["bit_exact_spot_check"])
print({k:v["ms"] for k,v in d["kernels"].items()}, d["parse"]["cycles_per_decode"] if d["parse"] else None)
for k,v in d["other_workloads"].items():
    if "error" in v: print(k,v); continue
    print(k,"value",v["value"],"ms",v["ms_per_step"],"e2e",v["e2e"]["value"],v["e2e"]["ms_per_step"],"blocking",v["e2e"]["blocking_ms"],"floor",v["e2e"]["d2h_floor_ms"],v["e2e"]["bit_exact_spot_check"],{kk:vv["ms"] for kk,vv in v["kernels"].items()})
PY
