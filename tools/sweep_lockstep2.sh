#!/bin/bash
# GPU box: lockstep parser geometry on the headline workload (numbers go to gpurun_out/ only).
mkdir -p gpurun_out
show() { python - "$1" <<'PY'
import json,sys
l=[x for x in open(sys.argv[1]) if x.startswith("{")]
d=json.loads(l[-1]) if l else None
print(sys.argv[1], d and (d["value"], d["ms_per_step"], {k:v["ms"] for k,v in d["kernels"].items()}))
PY
}
WEBP_B200_TOKEN_MAP=k timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
B="python bench.py --distinct 64 --steps 2 --e2e-steps 0 --no-cpu-baseline"
run() { name=$1; shift; env "$@" $B $EXTRA > gpurun_out/$name.log 2>&1; show gpurun_out/$name.log; }
EXTRA=""
for g in ${GEOMS:-4:0 8:0 7:4}; do
  run ls2_h_k${g%%:*}l${g##*:} WEBP_B200_TOKEN_MAP=k WEBP_B200_TOKEN_CW=${g%%:*} WEBP_B200_TOKEN_LPW=${g##*:}
done
if [ -z "$ONLY_HD" ]; then
EXTRA="--workload vp8_256x256_q80_rgbA --distinct 512"
run ls2_t WEBP_B200_X=1
EXTRA="--workload vp8_1080p_q75_m4_8part_normal_rgba"
run ls2_p8 WEBP_B200_X=1
fi
