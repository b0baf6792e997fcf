#!/bin/bash
# round 2, second session, call A: K3 (i4x4 sub-blocks read their neighbours straight from the tile, unrolled wavefront,
# token loads in flight beside the neighbour phase) and K4 (simple filter: two edges per pass): parity subset, then kernel times
# at 4 / 8 / 16 warps per image for K3, and config 3 (normal filter).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "manifest or fresh_corpora or mixed_sizes or dithering or row_bands or config3 or full_size_batch or parse_stages" > gpurun_out/r03a_pytest.log 2>&1; tail -3 gpurun_out/r03a_pytest.log
for w in 8 4 16; do
  export WEBP_B200_RECON_WARPS=$w
  python bench.py --distinct 64 --steps 3 --warmup 3 --e2e-steps 0 --no-cpu-baseline --no-others 2>&1 | tail -1 | grep -o '"value": [0-9.]*\|"kernels.*"clocks' | tr '\n' ' ' | cut -c1-460 | sed "s/^/recon_warps=$w /"; echo
done | tee gpurun_out/r03a_recon_warps.log
unset WEBP_B200_RECON_WARPS
python bench.py --workload vp8_1080p_q75_m4_8part_normal_rgba --distinct 64 --steps 3 --warmup 3 --e2e-steps 0 --no-cpu-baseline --no-others 2>&1 | tail -1 | grep -o '"value": [0-9.]*\|"kernels.*"clocks' | tr '\n' ' ' | cut -c1-460 | tee gpurun_out/r03a_config3.log
