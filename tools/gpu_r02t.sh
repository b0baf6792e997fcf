#!/bin/bash
# round 2, call T: request / bitstream order (host header probe, unusable slow-memory buffer reported after the decode)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu -k "slow_memory or scaling or crop or internal_memory or mixed or incremental or callers or dwebp or anim" > gpurun_out/r02t_pytest.log 2>&1; tail -3 gpurun_out/r02t_pytest.log
