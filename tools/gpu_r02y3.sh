#!/bin/bash
# round 2, call Y3: whole-picture VP8L and config 5 with the library before / after the warp-cooperative copies, same box
mkdir -p gpurun_out
cp libwebp_b200/libwebpdecoder_b200.so /tmp/cur.so
for v in head warp; do
  cp libwebp_b200/csrc/build/variants/$v.so libwebp_b200/libwebpdecoder_b200.so
  for w in vp8l_1080p_lossless_rgba vp8_4096x4096_q90_alpha_rgba; do
    timeout 900 python bench.py --workload $w --steps 2 --warmup 3 --e2e-steps 0 --no-cpu-baseline --no-others 2>/dev/null | tail -1 | grep -o '"value": [0-9.]*\|"alpha": {"ms": [0-9.]*' | tr '\n' ' ' | sed "s/^/$v $w /"; echo
  done
done | tee gpurun_out/r02y3_vp8l_ab.log
cp /tmp/cur.so libwebp_b200/libwebpdecoder_b200.so
