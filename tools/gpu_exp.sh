#!/bin/bash
# GPU box: A/B of the memset overlap and the two-stream pixel stages (numbers to gpurun_out/ only).
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu -k "concurrent or manifest or full_size or many_streams" 2>&1 | tail -3
show() { python - "$1" <<'PY'
import json,sys
l=[x for x in open(sys.argv[1]) if x.startswith("{")]
d=json.loads(l[-1]) if l else None
print(sys.argv[1], d and (d["value"], d["ms_per_step"], {k:v["ms"] for k,v in d["kernels"].items()}, d["e2e"] and d["e2e"]["ms_per_step"]))
PY
}
B="python bench.py --distinct 64 --steps 3 --no-cpu-baseline"
run() { name=$1; shift; env "$@" $B $EXTRA > gpurun_out/$name.log 2>&1; show gpurun_out/$name.log; }
EXTRA="--e2e-steps 2"
run x_base WEBP_B200_MEMSET_OVERLAP=0
run x_memset WEBP_B200_MEMSET_OVERLAP=1
EXTRA="--e2e-steps 0"
run x_ps2 WEBP_B200_MEMSET_OVERLAP=1 WEBP_B200_PIXEL_STREAMS=2
EXTRA="--e2e-steps 0 --workload vp8_1080p_q75_m4_8part_normal_rgba"
run x_p8_base WEBP_B200_MEMSET_OVERLAP=0
run x_p8_memset WEBP_B200_MEMSET_OVERLAP=1
run x_p8_ps2 WEBP_B200_MEMSET_OVERLAP=1 WEBP_B200_PIXEL_STREAMS=2
