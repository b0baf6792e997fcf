#!/bin/bash
# round 2, second session, call H: K4 tile rows 36 bytes apart (bank-conflict-free vertical edges) at 5 blocks per SM; K3 with the
# neighbour context interleaved per column / per row and the 16x16 / chroma mode looked at once per run: parity subset, times
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "manifest or fresh_corpora or mixed_sizes or dithering or row_bands or config3 or full_size_batch or extreme or crop_and_flip" > gpurun_out/r03h_pytest.log 2>&1; tail -3 gpurun_out/r03h_pytest.log
for wl in vp8_1080p_q75_m4_1part_simple_rgba vp8_1080p_q75_m4_8part_normal_rgba vp8_256x256_q80_rgbA; do
  python bench.py --workload $wl --distinct 64 --steps 3 --warmup 3 --e2e-steps 0 --no-cpu-baseline --no-others 2>&1 | tail -1 | grep -o '"value": [0-9.]*\|"recon": {"ms": [0-9.]*\|"filter": {"ms": [0-9.]*\|"emit": {"ms": [0-9.]*' | tr '\n' ' ' | sed "s/^/$wl /"; echo
done | tee gpurun_out/r03h_pixel_kernels.log
