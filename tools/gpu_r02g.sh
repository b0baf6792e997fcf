#!/bin/bash
# round 2, call G: fp parser, straight-line groups (default) against a branch per decode
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "manifest or fresh or full_size or mixed or config or extreme or campaign" 2>&1 | tail -5 > gpurun_out/r02g_pytest_gpu.log; cat gpurun_out/r02g_pytest_gpu.log
for g in 1 0; do
WEBP_B200_TOKEN_GROUPED=$g python bench.py --steps 3 --warmup 3 --no-cpu-baseline --distinct 64 --e2e-steps 0 --no-others > gpurun_out/r02g_bench_g$g.json 2> gpurun_out/r02g_bench_g$g.err; tail -3 gpurun_out/r02g_bench_g$g.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02g_bench_g$g.json').read().strip().splitlines()[-1])
print("grouped=$g value",d["value"],"ms",d["ms_per_step"],{k:v["ms"] for k,v in d["kernels"].items()}, d["parse"]["cycles_per_decode"] if d["parse"] else None)
PY
done
WEBP_B200_TOKEN_GROUPED=1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --distinct 64 --e2e-steps 0 --no-others --workload vp8_1080p_q75_m4_8part_normal_rgba > gpurun_out/r02g_bench_8p.json 2>&1
python - <<PY
import json
d=json.loads(open('gpurun_out/r02g_bench_8p.json').read().strip().splitlines()[-1])
print("8part value",d["value"],"ms",d["ms_per_step"],{k:v["ms"] for k,v in d["kernels"].items()}, d["parse"]["cycles_per_decode"] if d["parse"] else None)
PY
CMD="python bench.py --distinct 32 --steps 1 --warmup 3 --e2e-steps 0 --no-cpu-baseline --no-others"
ncu --set full --clock-control none --import-source on -k regex:k_parse_tokens_fp -s 3 -c 1 -o gpurun_out/r02g_tokens_fp $CMD > gpurun_out/r02g_ncu.log 2>&1
tail -2 gpurun_out/r02g_ncu.log
