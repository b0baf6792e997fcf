#!/bin/bash
# round 2, second session, call K: end to end with the page-locked buffers placed next to the GPU (host_numa in the line)
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r03k_topo.txt 2>&1; lscpu | grep -i "numa\|socket\|model name" >> gpurun_out/r03k_topo.txt
python bench.py --distinct 64 --steps 4 --warmup 3 --no-cpu-baseline --no-others > gpurun_out/r03k_bench.json 2> gpurun_out/r03k_bench.err
tail -1 gpurun_out/r03k_bench.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['d2h_GBps_bare'], d['e2e']['d2h_floor_ms'], d['host_numa'], d['kernels'])"
cat gpurun_out/r03k_topo.txt | head -20
