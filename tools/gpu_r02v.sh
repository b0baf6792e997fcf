#!/bin/bash
# round 2, call V: stage-level test on the device; the default bench line (all workloads, config 5 at 512 images); 4096 distinct images once
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu -k "parse_stages or slow_memory or crafted" > gpurun_out/r02v_pytest.log 2>&1; tail -3 gpurun_out/r02v_pytest.log
python bench.py > gpurun_out/r02v_bench.json 2> gpurun_out/r02v_bench.err; tail -1 gpurun_out/r02v_bench.json | cut -c1-400
python bench.py --distinct 4096 --steps 3 --warmup 3 --e2e-steps 0 --no-cpu-baseline --no-others > gpurun_out/r02v_bench_distinct4096.json 2> gpurun_out/r02v_bench_distinct4096.err; tail -1 gpurun_out/r02v_bench_distinct4096.json | cut -c1-600
