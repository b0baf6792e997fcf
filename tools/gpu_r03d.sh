#!/bin/bash
# round 2, second session, call D: lockstep K1 at 1..5 lanes per warp (more warps per sub-partition hide the serialised leaf paths?)
mkdir -p gpurun_out
export WEBP_B200_MODES=lockstep
for ll in 1 2 3 4 5; do
  export WEBP_B200_MODES_LANES=$ll
  python bench.py --distinct 64 --steps 3 --warmup 3 --e2e-steps 0 --no-cpu-baseline --no-others 2>&1 | tail -1 | grep -o '"value": [0-9.]*\|"modes": {"ms": [0-9.]*' | tr '\n' ' ' | sed "s/^/lockstep lanes=$ll /"; echo
done | tee gpurun_out/r03d_modes_lockstep_lanes.log
