#!/bin/bash
# round 2, second session, call C: K1 as lockstep lanes (vp8_modes_lockstep.h) against one image per warp; K3 with the next
# macroblock's tokens fetched ahead. Parity subset on the new defaults, then the A/B on configs 2, 3, 4.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "manifest or fresh_corpora or mixed_sizes or row_bands or config3 or config4 or full_size_batch or extreme or parse_stages or damage_campaign or both_damaged or starting_with_ff or many_small" > gpurun_out/r03c_pytest.log 2>&1; tail -3 gpurun_out/r03c_pytest.log
for wl in vp8_1080p_q75_m4_1part_simple_rgba vp8_1080p_q75_m4_8part_normal_rgba vp8_256x256_q80_rgbA; do for m in lockstep warp; do
  export WEBP_B200_MODES=$m
  python bench.py --workload $wl --distinct 64 --steps 3 --warmup 3 --e2e-steps 0 --no-cpu-baseline --no-others 2>&1 | tail -1 | grep -o '"value": [0-9.]*\|"modes": {"ms": [0-9.]*\|"recon": {"ms": [0-9.]*' | tr '\n' ' ' | sed "s/^/$wl modes=$m /"; echo
done; done | tee gpurun_out/r03c_modes_lockstep.log
