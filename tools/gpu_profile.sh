#!/bin/bash
# Runs on the GPU box (via gpurun): parity tests, a short bench, the ncu launch list and one full capture of
# the token-parse kernel. Outputs land in gpurun_out/ (copied to profiles/ by hand afterwards).
TAG=${1:-r01}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 > gpurun_out/${TAG}_pytest_gpu.log
CMD="python bench.py --batch 512 --distinct 32 --steps 2 --warmup 3 --e2e-steps 0 --no-cpu-baseline"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu1.log 2>&1
$CMD > gpurun_out/${TAG}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_parse_tokens -s 3 -c 1 -o gpurun_out/${TAG}_tokens $CMD > gpurun_out/${TAG}_ncu2.log 2>&1
tail -2 gpurun_out/${TAG}_pytest_gpu.log; tail -c 600 gpurun_out/${TAG}_plain.log; tail -3 gpurun_out/${TAG}_ncu2.log
