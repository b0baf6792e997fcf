#!/bin/bash
# round 2, second session, call T: K5 with the chroma upsampling of a row pair sharing its sums (upsample8_both), at 5 and 4 blocks per SM
mkdir -p gpurun_out
cp libwebp_b200/libwebpdecoder_b200.so /tmp/cur.so
for v in emit_base emit_both_b5 emit_both_b4; do
  cp libwebp_b200/csrc/build/variants/$v.so libwebp_b200/libwebpdecoder_b200.so
  python bench.py --distinct 32 --steps 3 --warmup 3 --e2e-steps 0 --no-cpu-baseline --no-others 2>&1 | tail -1 | grep -o '"value": [0-9.]*\|"emit": {"ms": [0-9.]*' | tr '\n' ' ' | sed "s/^/$v /"; echo
done | tee gpurun_out/r03t_emit_variants.log
cp /tmp/cur.so libwebp_b200/libwebpdecoder_b200.so
