#!/bin/bash
# round 2, call Y2: queue-full test off the per-symbol path: parity subset, VP8L and config 5 timings again
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "alpha or lossless or crafted or config5" > gpurun_out/r02y2_pytest.log 2>&1; tail -2 gpurun_out/r02y2_pytest.log
for w in vp8l_1080p_lossless_rgba vp8_4096x4096_q90_alpha_rgba; do
  timeout 900 python bench.py --workload $w --steps 2 --warmup 3 --e2e-steps 0 --no-cpu-baseline --no-others 2>/dev/null | tail -1 | grep -o '"value": [0-9.]*\|"alpha": {"ms": [0-9.]*' | tr '\n' ' ' | sed "s/^/$w /"; echo
done | tee gpurun_out/r02y2_timings.log
timeout 300 python tools/fuzz_gpu.py --seconds 60 --batch 2048 --seed 13 2>&1 | tail -1 | cut -c1-300
