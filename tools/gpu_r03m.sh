#!/bin/bash
# round 2, second session, call M: K1 with the sub-block contexts rotating through their words (no indexed byte inserts)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "manifest or fresh_corpora or mixed_sizes or parse_stages or damage_campaign or both_damaged or starting_with_ff or config4" > gpurun_out/r03m_pytest.log 2>&1; tail -3 gpurun_out/r03m_pytest.log
for wl in vp8_1080p_q75_m4_1part_simple_rgba vp8_256x256_q80_rgbA; do
  python bench.py --workload $wl --distinct 64 --steps 3 --warmup 3 --e2e-steps 0 --no-cpu-baseline --no-others 2>&1 | tail -1 | grep -o '"value": [0-9.]*\|"modes": {"ms": [0-9.]*' | tr '\n' ' ' | sed "s/^/$wl /"; echo
done | tee gpurun_out/r03m_modes_rotating.log
