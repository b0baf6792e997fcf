#!/bin/bash
# round 2, second session, call R: 2 warps per image in K3 / K4 for 16-row thumbnails against 4
mkdir -p gpurun_out
for w in 4 2; do
  export WEBP_B200_PIXEL_WARPS=$w
  python bench.py --workload vp8_256x256_q80_rgbA --steps 3 --warmup 3 --e2e-steps 0 --no-cpu-baseline --no-others 2>&1 | tail -1 | grep -o '"value": [0-9.]*\|"recon": {"ms": [0-9.]*\|"filter": {"ms": [0-9.]*' | tr '\n' ' ' | sed "s/^/thumbnails pixel_warps=$w /"; echo
done | tee gpurun_out/r03r_pixel_warps2.log
WEBP_B200_PIXEL_WARPS=2 timeout 300 python -m pytest tests -x -q -m gpu -k "manifest or mixed_sizes or config4 or dithering" 2>&1 | tail -1 | tee -a gpurun_out/r03r_pixel_warps2.log
