#!/bin/bash
# round 2, second session, last call: smoke(), a parity subset and a second seed of the GPU damage campaign on the tree as committed
mkdir -p gpurun_out
python __graft_entry__.py smoke > gpurun_out/r03q_smoke.log 2>&1; tail -1 gpurun_out/r03q_smoke.log
timeout 600 python -m pytest tests -x -q -m gpu -k "manifest or fresh_corpora or mixed_sizes or parse_stages or dithering or config4" > gpurun_out/r03q_pytest.log 2>&1; tail -2 gpurun_out/r03q_pytest.log
timeout 300 python tools/fuzz_gpu.py --seconds 75 --batch 2048 --seed 21 > gpurun_out/r03q_fuzz_gpu.log 2>&1; tail -1 gpurun_out/r03q_fuzz_gpu.log | cut -c1-400
