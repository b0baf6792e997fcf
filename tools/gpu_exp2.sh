#!/bin/bash
# GPU box: images per warp in the mode parse (numbers to gpurun_out/ only).
mkdir -p gpurun_out
show() { python - "$1" <<'PY'
import json,sys
l=[x for x in open(sys.argv[1]) if x.startswith("{")]
d=json.loads(l[-1]) if l else None
print(sys.argv[1], d and (d["value"], d["ms_per_step"], {k:v["ms"] for k,v in d["kernels"].items()}))
PY
}
B="python bench.py --distinct 64 --steps 2 --no-cpu-baseline --e2e-steps 0"
run() { name=$1; shift; env "$@" $B $EXTRA > gpurun_out/$name.log 2>&1; show gpurun_out/$name.log; }
for L in 2 4 7; do
WEBP_B200_MODES_LANES=$L timeout 300 python -m pytest tests -x -q -m gpu -k "manifest or full_size" 2>&1 | tail -1
EXTRA=""; run m_h_l$L WEBP_B200_MODES_LANES=$L
EXTRA="--workload vp8_1080p_q75_m4_8part_normal_rgba"; run m_p8_l$L WEBP_B200_MODES_LANES=$L
done
EXTRA="--workload vp8_256x256_q80_rgbA --distinct 512"; run m_t_l1 WEBP_B200_MODES_LANES=1
EXTRA="--workload vp8_256x256_q80_rgbA --distinct 512"; run m_t_l4 WEBP_B200_MODES_LANES=4
