#!/bin/bash
# round 2, call C: the new host pipeline (submit/wait, shared scratch, per-device attributes) -- tests, then bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu --durations=8 2>&1 | tail -25 > gpurun_out/r02c_pytest_gpu.log; cat gpurun_out/r02c_pytest_gpu.log
python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/r02c_bench_reference.json 2> gpurun_out/r02c_bench_reference.err; tail -c 400 gpurun_out/r02c_bench_reference.json
WEBP_B200_TRACE=1 python bench.py --steps 6 --warmup 3 > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err; tail -c 6000 gpurun_out/r02c_bench.json; tail -5 gpurun_out/r02c_bench.err
