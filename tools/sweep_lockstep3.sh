#!/bin/bash
# GPU box: the two ways of running the lockstep lanes (block ends on the spot / grouped event points) per workload.
mkdir -p gpurun_out
show() { python - "$1" <<'PY'
import json,sys
l=[x for x in open(sys.argv[1]) if x.startswith("{")]
d=json.loads(l[-1]) if l else None
print(sys.argv[1], d and (d["value"], d["ms_per_step"], {k:v["ms"] for k,v in d["kernels"].items()}))
PY
}
WEBP_B200_TOKEN_MAP_INNER=1 timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
B="python bench.py --distinct 64 --steps 2 --e2e-steps 0 --no-cpu-baseline"
run() { name=$1; shift; env "$@" $B $EXTRA > gpurun_out/$name.log 2>&1; show gpurun_out/$name.log; }
EXTRA=""
run ls3_h_default WEBP_B200_X=1
EXTRA="--workload vp8_256x256_q80_rgbA --distinct 512"
run ls3_t_default WEBP_B200_X=1
run ls3_t_k16g WEBP_B200_TOKEN_CW=16 WEBP_B200_TOKEN_GROUPED=1
run ls3_t_k16i WEBP_B200_TOKEN_CW=16 WEBP_B200_TOKEN_GROUPED=0
EXTRA="--workload vp8_1080p_q75_m4_8part_normal_rgba"
run ls3_p8_default WEBP_B200_X=1
run ls3_p8_k8g WEBP_B200_TOKEN_CW=8 WEBP_B200_TOKEN_GROUPED=1
