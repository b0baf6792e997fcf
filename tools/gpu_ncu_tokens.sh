#!/bin/bash
# GPU box: one full ncu capture of the token-parse kernel at the bench batch size.
TAG=${1:-r01b}
shift
CMD="python bench.py --batch ${BATCH:-4096} --distinct 32 --steps 1 --warmup 3 --e2e-steps 0 --no-cpu-baseline $*"
mkdir -p gpurun_out
$CMD > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_parse_tokens -s 3 -c 1 -o gpurun_out/${TAG}_tokens $CMD > gpurun_out/${TAG}_ncu.log 2>&1
tail -c 400 gpurun_out/${TAG}_plain.log; tail -3 gpurun_out/${TAG}_ncu.log
