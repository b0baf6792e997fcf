#!/bin/bash
# GPU box: full ncu captures of the three pixel-stage kernels at the bench batch size (one launch each).
TAG=${1:-r01g}
CMD="python bench.py --batch ${BATCH:-4096} --distinct 32 --steps 1 --warmup 3 --e2e-steps 0 --no-cpu-baseline"
mkdir -p gpurun_out
$CMD > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_reconstruct|k_loop_filter|k_emit' -s 9 -c 3 -o gpurun_out/${TAG}_pixels $CMD > gpurun_out/${TAG}_ncu.log 2>&1
tail -c 300 gpurun_out/${TAG}_plain.log; tail -3 gpurun_out/${TAG}_ncu.log
