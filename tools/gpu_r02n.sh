#!/bin/bash
# round 2, call N: the ncu rows round 1 lacked -- launch list, then full captures of k_parse_modes, k_reconstruct, k_loop_filter, k_emit
# (config 2) and of k_alpha_pixels / k_alpha_finish (config 5, small batch)
mkdir -p gpurun_out
timeout 300 python -m pytest tests -x -q -m gpu -k "callers_stream" 2>&1 | tail -2
CMD="python bench.py --distinct 32 --steps 2 --warmup 3 --e2e-steps 0 --no-cpu-baseline --no-others"
$CMD > gpurun_out/r02n_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02n_launches.csv $CMD > gpurun_out/r02n_ncu1.log 2>&1
tail -1 gpurun_out/r02n_ncu1.log | cut -c1-200
ncu --set full --clock-control none --import-source on -k regex:'k_parse_modes|k_reconstruct|k_loop_filter|k_emit' -s 12 -c 4 -o gpurun_out/r02n_pixels $CMD > gpurun_out/r02n_ncu2.log 2>&1
tail -1 gpurun_out/r02n_ncu2.log | cut -c1-200
CMD5="python bench.py --workload vp8_4096x4096_q90_alpha_rgba --batch 32 --distinct 4 --steps 1 --warmup 3 --e2e-steps 0 --no-cpu-baseline --no-others"
$CMD5 > gpurun_out/r02n_plain5.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'k_alpha_pixels|k_alpha_finish' -s 6 -c 2 -o gpurun_out/r02n_alpha $CMD5 > gpurun_out/r02n_ncu3.log 2>&1
tail -1 gpurun_out/r02n_ncu3.log | cut -c1-200
