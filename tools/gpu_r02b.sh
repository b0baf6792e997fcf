#!/bin/bash
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o /tmp/chain_floor tools/chain_floor.cu && /tmp/chain_floor > gpurun_out/r02b_chain_floor.json 2> gpurun_out/r02b_chain_floor.err
tail -c 600 gpurun_out/r02b_chain_floor.json; cat gpurun_out/r02b_chain_floor.err
free -g | head -2; nproc
