#!/bin/bash
# round 2, second session, final call: the whole GPU suite on the round's last code, the default bench line (all workloads), the
# ncu launch list of a bench run and full captures of the pixel kernels (1024 full-HD images, one launch each)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r03z_pytest_gpu.log 2>&1; tail -3 gpurun_out/r03z_pytest_gpu.log
( time python bench.py > gpurun_out/r03z_bench.json 2> gpurun_out/r03z_bench.err ) 2> gpurun_out/r03z_bench_time.txt; tail -1 gpurun_out/r03z_bench.json | cut -c1-300; tail -3 gpurun_out/r03z_bench_time.txt
CMD="python bench.py --steps 2 --warmup 3 --e2e-steps 0 --no-cpu-baseline --no-others --distinct 64"
$CMD > gpurun_out/r03z_plain.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r03z_ncu_launches.csv $CMD > gpurun_out/r03z_ncu_launches.log 2>&1
CMD2="python bench.py --batch 1024 --distinct 32 --steps 1 --warmup 3 --e2e-steps 0 --no-cpu-baseline --no-others"
$CMD2 > gpurun_out/r03z_plain2.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_reconstruct|k_loop_filter|k_emit' -s 9 -c 3 -o gpurun_out/r03z_pixels $CMD2 > gpurun_out/r03z_ncu_full.log 2>&1
ls -la gpurun_out/ | grep r03z
