#!/bin/bash
# round 2, call O: failing-row bookkeeping (both chunks damaged) on the device; damage campaign without that tolerated class; bench sanity
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "both_damaged or alpha or damage_campaign or callers_stream or every_token_mapping or incremental" > gpurun_out/r02o_pytest.log 2>&1; tail -5 gpurun_out/r02o_pytest.log
python bench.py --distinct 32 --steps 3 --warmup 3 --e2e-steps 0 --no-cpu-baseline --no-others > gpurun_out/r02o_bench.log 2>&1; tail -1 gpurun_out/r02o_bench.log | grep -o '"kernels.*"clocks' | cut -c1-400
