#!/bin/bash
# GPU box: parity tests + one short bench of each workload (numbers to gpurun_out/ only).
mkdir -p gpurun_out
show() { python - "$1" <<'PY'
import json,sys
l=[x for x in open(sys.argv[1]) if x.startswith("{")]
d=json.loads(l[-1]) if l else None
print(sys.argv[1], d and (d["value"], d["ms_per_step"], {k:v["ms"] for k,v in d["kernels"].items()}, d["e2e"] and d["e2e"]["value"]))
PY
}
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
B="python bench.py --distinct 64 --steps 2 --e2e-steps ${E2E:-0} --no-cpu-baseline"
$B > gpurun_out/q_hd.log 2>&1; show gpurun_out/q_hd.log
if [ -z "$ONLY_HD" ]; then
$B --workload vp8_256x256_q80_rgbA --distinct 512 > gpurun_out/q_thumb.log 2>&1; show gpurun_out/q_thumb.log
$B --workload vp8_1080p_q75_m4_8part_normal_rgba > gpurun_out/q_p8.log 2>&1; show gpurun_out/q_p8.log
fi
