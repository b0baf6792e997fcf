#!/bin/bash
# round 2, call R: literal reader for partitions that start with 0xFF (k_parse_literal): parity tests, damage campaign with forced 0xFF, bench sanity
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "starting_with_ff or both_damaged or damage_campaign or manifest or mixed" > gpurun_out/r02r_pytest.log 2>&1; tail -3 gpurun_out/r02r_pytest.log
timeout 600 python tools/fuzz_gpu.py --seconds 150 --batch 2048 --seed 9 > gpurun_out/r02r_fuzz_gpu.log 2>&1; tail -3 gpurun_out/r02r_fuzz_gpu.log | cut -c1-600
python bench.py --distinct 32 --steps 3 --warmup 3 --e2e-steps 0 --no-cpu-baseline --no-others > gpurun_out/r02r_bench.log 2>&1; tail -1 gpurun_out/r02r_bench.log | grep -o '"kernels.*"clocks' | cut -c1-400
