#!/bin/bash
# GPU box: the three ways of running the lockstep lanes per workload (numbers go to gpurun_out/ only).
mkdir -p gpurun_out
show() { python - "$1" <<'PY'
import json,sys
l=[x for x in open(sys.argv[1]) if x.startswith("{")]
d=json.loads(l[-1]) if l else None
print(sys.argv[1], d and (d["value"], d["ms_per_step"], {k:v["ms"] for k,v in d["kernels"].items()}))
PY
}
B="python bench.py --distinct 64 --steps 2 --e2e-steps 0 --no-cpu-baseline"
run() { name=$1; shift; env "$@" $B $EXTRA > gpurun_out/$name.log 2>&1; show gpurun_out/$name.log; }
EXTRA=""
for g in ${STYLES:-0 1 2}; do run ls4_h_g$g WEBP_B200_TOKEN_GROUPED=$g; done
EXTRA="--workload vp8_256x256_q80_rgbA --distinct 512"
for g in ${STYLES:-0 1 2}; do run ls4_t_g$g WEBP_B200_TOKEN_GROUPED=$g; done
EXTRA="--workload vp8_1080p_q75_m4_8part_normal_rgba"
for g in ${STYLES:-0 1 2}; do run ls4_p8_g$g WEBP_B200_TOKEN_GROUPED=$g; done
