#!/bin/bash
# round 2, second session, call N: the damage campaign on the GPU with the round's last kernels, then the default bench line
mkdir -p gpurun_out
timeout 600 python tools/fuzz_gpu.py --seconds 100 --batch 2048 --seed 20 > gpurun_out/r03n_fuzz_gpu.log 2>&1; tail -3 gpurun_out/r03n_fuzz_gpu.log | cut -c1-700
python bench.py > gpurun_out/r03n_bench.json 2> gpurun_out/r03n_bench.err; tail -1 gpurun_out/r03n_bench.json | cut -c1-200
