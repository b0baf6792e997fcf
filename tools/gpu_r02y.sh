#!/bin/bash
# round 2, call Y: backward references of the VP8L / ALPH pixel loop carried out by the whole warp: parity (alpha, lossless, damage), config 5 and VP8L timings
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu -k "alpha or lossless or crafted or damage_campaign or config5 or both_damaged or many_small" > gpurun_out/r02y_pytest.log 2>&1; tail -3 gpurun_out/r02y_pytest.log
timeout 600 python tools/fuzz_gpu.py --seconds 100 --batch 2048 --seed 12 > gpurun_out/r02y_fuzz_gpu.log 2>&1; tail -1 gpurun_out/r02y_fuzz_gpu.log | cut -c1-400
for w in vp8_4096x4096_q90_alpha_rgba vp8l_1080p_lossless_rgba; do
  timeout 900 python bench.py --workload $w --steps 2 --warmup 3 --e2e-steps 0 --no-cpu-baseline --no-others 2>/dev/null | tail -1 | grep -o '"value": [0-9.]*\|"kernels.*"clocks' | tr '\n' ' ' | cut -c1-420 | sed "s/^/$w /"; echo
done | tee gpurun_out/r02y_alpha_timings.log
