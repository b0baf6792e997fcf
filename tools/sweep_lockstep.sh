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
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
WEBP_B200_TOKEN_MAP=k timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
B="python bench.py --distinct 64 --steps 2 --e2e-steps 0 --no-cpu-baseline"
run() { name=$1; shift; env "$@" $B $EXTRA > gpurun_out/$name.log 2>&1; show gpurun_out/$name.log; }
EXTRA=""
run ls_h_warp28 WEBP_B200_TOKEN_MAP=warp
run ls_h_warp1 WEBP_B200_TOKEN_MAP=warp WEBP_B200_TOKEN_IPB=1
run ls_h_warp14 WEBP_B200_TOKEN_MAP=warp WEBP_B200_TOKEN_IPB=14
run ls_h_k14l2 WEBP_B200_TOKEN_MAP=k WEBP_B200_TOKEN_CW=14 WEBP_B200_TOKEN_LPW=2
run ls_h_k10l3 WEBP_B200_TOKEN_MAP=k WEBP_B200_TOKEN_CW=10 WEBP_B200_TOKEN_LPW=3
run ls_h_k7l4 WEBP_B200_TOKEN_MAP=k WEBP_B200_TOKEN_CW=7 WEBP_B200_TOKEN_LPW=4
EXTRA="--workload vp8_256x256_q80_rgbA --distinct 512"
run ls_t_warp WEBP_B200_TOKEN_MAP=warp
run ls_t_k16 WEBP_B200_TOKEN_MAP=k WEBP_B200_TOKEN_CW=16
EXTRA="--workload vp8_1080p_q75_m4_8part_normal_rgba"
run ls_p8_warp WEBP_B200_TOKEN_MAP=warp
run ls_p8_k16 WEBP_B200_TOKEN_MAP=k WEBP_B200_TOKEN_CW=16
