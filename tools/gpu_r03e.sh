#!/bin/bash
# round 2, second session, call E: the fp token parser with the 4096 streams of config 2 spread over more warps per sub-partition
# (fewer lanes per warp: fewer steps with a block end in them; two to four warps share one issue port)
mkdir -p gpurun_out
for cfg in "4 7" "8 4" "12 3" "16 2"; do
  set -- $cfg
  export WEBP_B200_TOKEN_CW=$1 WEBP_B200_TOKEN_LPW=$2
  python bench.py --distinct 64 --steps 3 --warmup 3 --e2e-steps 0 --no-cpu-baseline --no-others 2>&1 | tail -1 | grep -o '"value": [0-9.]*\|"tokens": {"ms": [0-9.]*' | tr '\n' ' ' | sed "s/^/warps_per_block=$1 lanes_per_warp=$2 /"; echo
done | tee gpurun_out/r03e_tokens_warps_lanes.log
