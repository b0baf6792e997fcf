#!/bin/bash
# round 2, call Q: vote-interval variants of the token parser; config 5 at 512 images per batch; ncu row of k_loop_filter (smaller batch)
mkdir -p gpurun_out
cp libwebp_b200/libwebpdecoder_b200.so /tmp/cur.so
for v in g8 g4 g16 g8 g4 g16; do
  cp libwebp_b200/csrc/build/variants/$v.so libwebp_b200/libwebpdecoder_b200.so
  python bench.py --distinct 32 --steps 3 --warmup 3 --e2e-steps 0 --no-cpu-baseline --no-others 2>&1 | tail -1 | grep -o '"tokens": {"ms": [0-9.]*' | sed "s/^/$v /"
done | tee gpurun_out/r02q_variants.log
cp /tmp/cur.so libwebp_b200/libwebpdecoder_b200.so
timeout 600 python bench.py --workload vp8_4096x4096_q90_alpha_rgba --batch 512 --distinct 8 --steps 1 --warmup 1 --e2e-steps 0 --no-cpu-baseline --no-others > gpurun_out/r02q_config5_b512.log 2>&1; tail -1 gpurun_out/r02q_config5_b512.log | cut -c1-200; grep -o '"kernels.*"clocks' gpurun_out/r02q_config5_b512.log | cut -c1-500
CMD="python bench.py --batch 1024 --distinct 32 --steps 1 --warmup 3 --e2e-steps 0 --no-cpu-baseline --no-others"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_loop_filter' -s 3 -c 1 -o gpurun_out/r02q_filter $CMD > gpurun_out/r02q_ncu.log 2>&1
grep -i "passes\|error" gpurun_out/r02q_ncu.log | cut -c1-200
