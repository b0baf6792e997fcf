#!/bin/bash
# round 2, second session, call J: K3 with the short form for macroblocks whose blocks carry at most a DC level
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "manifest or fresh_corpora or mixed_sizes or full_size_batch or config3 or parse_stages" > gpurun_out/r03j_pytest.log 2>&1; tail -3 gpurun_out/r03j_pytest.log
for wl in vp8_1080p_q75_m4_1part_simple_rgba vp8_256x256_q80_rgbA; do
  python bench.py --workload $wl --distinct 64 --steps 3 --warmup 3 --e2e-steps 0 --no-cpu-baseline --no-others 2>&1 | tail -1 | grep -o '"value": [0-9.]*\|"recon": {"ms": [0-9.]*\|"filter": {"ms": [0-9.]*' | tr '\n' ' ' | sed "s/^/$wl /"; echo
done | tee gpurun_out/r03j_dc_only.log
