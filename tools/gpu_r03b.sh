#!/bin/bash
# round 2, second session, call B: K3 as one warp per macroblock row (progress counters, next macroblock's MbInfo/MbTok fetched ahead)
# against the anti-diagonal wavefront with a block-wide barrier, at 8 and 16 warps per image; parity subset on the new default.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "manifest or fresh_corpora or mixed_sizes or dithering or row_bands or config3 or full_size_batch or extreme" > gpurun_out/r03b_pytest.log 2>&1; tail -3 gpurun_out/r03b_pytest.log
for r in 1 0; do for w in 8 16 4; do
  export WEBP_B200_RECON_WARPS=$w WEBP_B200_RECON_ROWS=$r
  python bench.py --distinct 64 --steps 3 --warmup 3 --e2e-steps 0 --no-cpu-baseline --no-others 2>&1 | tail -1 | grep -o '"value": [0-9.]*\|"recon": {"ms": [0-9.]*' | tr '\n' ' ' | sed "s/^/rows=$r warps=$w /"; echo
done; done | tee gpurun_out/r03b_recon_rows.log
