#!/bin/bash
# GPU box: one test (or -k expression) of the GPU suite, output kept under gpurun_out/.
K=${1:?-k expression}; TAG=${2:-r01u}
mkdir -p gpurun_out
timeout ${3:-200} python -m pytest tests -x -q -m gpu -k "$K" --durations=3 2>&1 | tail -25 > gpurun_out/${TAG}_pytest_one.log; cat gpurun_out/${TAG}_pytest_one.log
