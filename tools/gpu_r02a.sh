#!/bin/bash
# round 2, call A: chain-floor microbenchmarks + parser-vs-copy interference (isolated)
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o /tmp/chain_floor tools/chain_floor.cu && /tmp/chain_floor > gpurun_out/r02a_chain_floor.json 2> gpurun_out/r02a_chain_floor.err
cat gpurun_out/r02a_chain_floor.json | cut -c1-3000
timeout 900 python tools/exp_interference.py > gpurun_out/r02a_interference.log 2>&1; cat gpurun_out/r02a_interference.log | tail -8
