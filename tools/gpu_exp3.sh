#!/bin/bash
# GPU box: images per warp in the mode parse, thumbnails at two batch sizes (numbers to gpurun_out/ only).
mkdir -p gpurun_out
show() { python - "$1" <<'PY'
import json,sys
l=[x for x in open(sys.argv[1]) if x.startswith("{")]
d=json.loads(l[-1]) if l else None
print(sys.argv[1], d and (d["value"], d["ms_per_step"], {k:v["ms"] for k,v in d["kernels"].items()}))
PY
}
B="python bench.py --steps 2 --no-cpu-baseline --e2e-steps 0 --workload vp8_256x256_q80_rgbA --distinct 512"
run() { name=$1; shift; env "$@" $B $EXTRA > gpurun_out/$name.log 2>&1; show gpurun_out/$name.log; }
EXTRA=""
for L in 8 16 32; do run m_t_l$L WEBP_B200_MODES_LANES=$L; done
EXTRA="--batch 16384"
for L in 1 2 4 8; do run m_t16k_l$L WEBP_B200_MODES_LANES=$L; done
EXTRA="--batch 8192"
for L in 1 2 4; do run m_t8k_l$L WEBP_B200_MODES_LANES=$L; done
WEBP_B200_MODES_LANES=16 timeout 300 python -m pytest tests -x -q -m gpu -k "manifest or full_size or mixed" 2>&1 | tail -1
