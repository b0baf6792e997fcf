#!/bin/bash
# GPU box: parity tests with the lockstep token parser forced, then A/B timing of its geometry against the
# warp-per-partition mapping (not a bench line: numbers go to gpurun_out/ only).
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
run ls_h_warp WEBP_B200_TOKEN_MAP=warp
run ls_h_k4 WEBP_B200_TOKEN_MAP=k WEBP_B200_TOKEN_CW=4
run ls_h_k8 WEBP_B200_TOKEN_MAP=k WEBP_B200_TOKEN_CW=8
run ls_h_k4l4 WEBP_B200_TOKEN_MAP=k WEBP_B200_TOKEN_CW=4 WEBP_B200_TOKEN_LPW=4
run ls_h_k2 WEBP_B200_TOKEN_MAP=k WEBP_B200_TOKEN_CW=2
EXTRA="--workload vp8_256x256_q80_rgbA --distinct 512"
run ls_t_k4 WEBP_B200_TOKEN_MAP=k WEBP_B200_TOKEN_CW=4
run ls_t_k8 WEBP_B200_TOKEN_MAP=k WEBP_B200_TOKEN_CW=8
EXTRA="--workload vp8_1080p_q75_m4_8part_normal_rgba"
run ls_p8_k4 WEBP_B200_TOKEN_MAP=k WEBP_B200_TOKEN_CW=4
run ls_p8_k8 WEBP_B200_TOKEN_MAP=k WEBP_B200_TOKEN_CW=8
