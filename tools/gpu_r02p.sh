#!/bin/bash
# round 2, call P: which part of the failing-row bookkeeping costs the token parser its 4 % (variants of the library, same box)
mkdir -p gpurun_out
cp libwebp_b200/libwebpdecoder_b200.so /tmp/cur.so
for v in v3_old v5 v3_old v5; do
  cp libwebp_b200/csrc/build/variants/$v.so libwebp_b200/libwebpdecoder_b200.so
  python bench.py --distinct 32 --steps 3 --warmup 3 --e2e-steps 0 --no-cpu-baseline --no-others 2>&1 | tail -1 | grep -o '"tokens": {"ms": [0-9.]*' | sed "s/^/$v /"
done | tee gpurun_out/r02p_variants.log
cp /tmp/cur.so libwebp_b200/libwebpdecoder_b200.so
timeout 900 python -m pytest tests -x -q -m gpu -k "both_damaged or alpha or damage_campaign or every_token_mapping or incremental" > gpurun_out/r02p_pytest.log 2>&1; tail -3 gpurun_out/r02p_pytest.log
timeout 600 python tools/fuzz_gpu.py --seconds 60 --batch 2048 --seed 8 > gpurun_out/r02p_fuzz_gpu.log 2>&1; tail -1 gpurun_out/r02p_fuzz_gpu.log | cut -c1-400
