#!/bin/bash
# round 2, second session, call P: resident blocks per SM of K3 (4 / 5: 64 / 48 registers) and K4 (5 / 6: 48 / 40 registers)
mkdir -p gpurun_out
cp libwebp_b200/libwebpdecoder_b200.so /tmp/cur.so
for v in k3b4_k4b5 k3b4_k4b6 k3b5_k4b6; do
  cp libwebp_b200/csrc/build/variants/$v.so libwebp_b200/libwebpdecoder_b200.so
  for wl in vp8_1080p_q75_m4_1part_simple_rgba vp8_1080p_q75_m4_8part_normal_rgba; do
    python bench.py --workload $wl --distinct 64 --steps 3 --warmup 3 --e2e-steps 0 --no-cpu-baseline --no-others 2>&1 | tail -1 | grep -o '"value": [0-9.]*\|"recon": {"ms": [0-9.]*\|"filter": {"ms": [0-9.]*' | tr '\n' ' ' | sed "s/^/$v $wl /"; echo
  done
done | tee gpurun_out/r03p_blocks_per_sm.log
cp /tmp/cur.so libwebp_b200/libwebpdecoder_b200.so
