#!/bin/bash
# round 2, call W: the whole GPU suite on the current code
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/r02w_pytest_gpu.log 2>&1; tail -4 gpurun_out/r02w_pytest_gpu.log
