#!/bin/bash
# GPU box, short budget: the tests that touch this session's device change (VP8L pixel loop), smoke, then the bench line.
TAG=${1:-r01u}
mkdir -p gpurun_out
timeout 170 python -m pytest tests -x -q -m gpu -k "lossless or manifest or mixed_sizes" --durations=4 2>&1 | tail -10 > gpurun_out/${TAG}_pytest_gpu_subset.log; cat gpurun_out/${TAG}_pytest_gpu_subset.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; tail -1 gpurun_out/${TAG}_smoke.log
timeout 170 python bench.py > gpurun_out/${TAG}_bench.log 2>&1; tail -1 gpurun_out/${TAG}_bench.log | cut -c1-1500
