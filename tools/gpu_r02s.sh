#!/bin/bash
# round 2, call S: VP8L limits lifted (group remap, palette at any position), scaling beyond 16383, ImgDesc layout change: parity + bench sanity
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu -k "crafted or scaling or lossless or alpha or starting_with_ff or damage_campaign or manifest or extreme" > gpurun_out/r02s_pytest.log 2>&1; tail -3 gpurun_out/r02s_pytest.log
timeout 600 python tools/fuzz_gpu.py --seconds 120 --batch 2048 --seed 10 > gpurun_out/r02s_fuzz_gpu.log 2>&1; tail -3 gpurun_out/r02s_fuzz_gpu.log | cut -c1-600
python bench.py --distinct 32 --steps 3 --warmup 3 --e2e-steps 0 --no-cpu-baseline --no-others > gpurun_out/r02s_bench.log 2>&1; tail -1 gpurun_out/r02s_bench.log | grep -o '"kernels.*"clocks' | cut -c1-400
