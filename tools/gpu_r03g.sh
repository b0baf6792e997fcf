#!/bin/bash
# round 2, second session, call G: full ncu captures (source on) of the reworked pixel kernels, 1024 full-HD images, one launch each
mkdir -p gpurun_out
CMD="python bench.py --batch 1024 --distinct 32 --steps 1 --warmup 3 --e2e-steps 0 --no-cpu-baseline --no-others"
$CMD > gpurun_out/r03g_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_reconstruct|k_loop_filter|k_emit' -s 9 -c 3 -o gpurun_out/r03g_pixels $CMD > gpurun_out/r03g_ncu.log 2>&1
tail -c 400 gpurun_out/r03g_plain.log | grep -o '"kernels.*"clocks' | cut -c1-400; grep -i "passes\|error" gpurun_out/r03g_ncu.log | cut -c1-160 | tail -4
