#!/bin/bash
# round 2, call U: banded probability rows (2 KB per image) for launches with very many small images: parity, thumbnails with and without, full-HD sanity
mkdir -p gpurun_out
python bench.py --distinct 32 --steps 3 --warmup 3 --e2e-steps 0 --no-cpu-baseline --no-others 2>&1 | tail -1 | grep -o '"kernels.*"clocks' | cut -c1-330 | tee gpurun_out/r02u_fullhd.log
for b in 0 1 auto; do
  if [ $b = auto ]; then unset WEBP_B200_TOKEN_BAND; else export WEBP_B200_TOKEN_BAND=$b; fi
  python bench.py --workload vp8_256x256_q80_rgbA --steps 3 --warmup 3 --e2e-steps 0 --no-cpu-baseline --no-others 2>&1 | tail -1 | grep -o '"value": [0-9.]*\|"kernels.*"clocks' | tr '\n' ' ' | cut -c1-420 | sed "s/^/band=$b /"; echo
done | tee gpurun_out/r02u_thumbnails.log
unset WEBP_B200_TOKEN_BAND
timeout 1500 python -m pytest tests -x -q -m gpu -k "every_token_mapping or config4 or manifest or many_streams" > gpurun_out/r02u_pytest.log 2>&1; tail -3 gpurun_out/r02u_pytest.log
