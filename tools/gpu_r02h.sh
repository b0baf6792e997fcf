#!/bin/bash
# round 2, call H (2 GPUs): whole GPU suite incl. the in-process two-device test, bench N=1 with every workload, bench N=2
mkdir -p gpurun_out
nvidia-smi -L
timeout 1800 python -m pytest tests -x -q -m gpu --durations=6 2>&1 | tail -14 > gpurun_out/r02h_pytest_gpu.log; cat gpurun_out/r02h_pytest_gpu.log
python bench.py --steps 6 --warmup 3 --cpu-seconds 6 > gpurun_out/r02h_bench.json 2> gpurun_out/r02h_bench.err; tail -3 gpurun_out/r02h_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02h_bench.json').read().strip().splitlines()[-1])
print("value",d["value"],"ms",d["ms_per_step"],"with_h2d",d["value_with_h2d"], "cpu", d["cpu_baseline"]["value"])
e=d["e2e"]; print("e2e",e["value"],e["ms_per_step"],"blocking",e["blocking_ms"],"floor",e["d2h_floor_ms"],e["bit_exact_spot_check"],e["host_wait_ms"])
print({k:v["ms"] for k,v in d["kernels"].items()}, d["parse"])
for k,v in d["other_workloads"].items():
    if "error" in v: print(k,v); continue
    print(k,"value",v["value"],"ms",v["ms_per_step"],"e2e",v["e2e"]["value"],v["e2e"]["ms_per_step"],"blocking",v["e2e"]["blocking_ms"],"floor",v["e2e"]["d2h_floor_ms"],v["e2e"]["bit_exact_spot_check"],"cpu",v["cpu_baseline"]["value"] if v["cpu_baseline"] else None,{kk:vv["ms"] for kk,vv in v["kernels"].items()})
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r02h_bench_n2.json 2> gpurun_out/r02h_bench_n2.err; tail -2 gpurun_out/r02h_bench_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02h_bench_n2.json').read().strip().splitlines()[-1])
print("N=2 value",d["value"],"ms",d["ms_per_step"],"e2e",d["e2e"]["value"],d["e2e"]["ms_per_step"],d["e2e"].get("d2h_GBps_bare"), {k:(v["value"],v["e2e"]["value"]) for k,v in d["other_workloads"].items() if "error" not in v})
PY
