#!/bin/bash
# round 2, second session, call I: 4 warps per image in K3 / K4 for small pictures (65536 thumbnails of 16 x 16 macroblocks)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "manifest or fresh_corpora or mixed_sizes or dithering or config4 or many_small or extreme" > gpurun_out/r03i_pytest.log 2>&1; tail -3 gpurun_out/r03i_pytest.log
for w in 8 4; do
  export WEBP_B200_PIXEL_WARPS=$w
  python bench.py --workload vp8_256x256_q80_rgbA --steps 3 --warmup 3 --e2e-steps 0 --no-cpu-baseline --no-others 2>&1 | tail -1 | grep -o '"value": [0-9.]*\|"recon": {"ms": [0-9.]*\|"filter": {"ms": [0-9.]*\|"emit": {"ms": [0-9.]*' | tr '\n' ' ' | sed "s/^/thumbnails pixel_warps=$w /"; echo
done | tee gpurun_out/r03i_pixel_warps.log
