#!/bin/bash
# GPU box: end-to-end (host buffers in/out) timing against pipeline waves and pixel-stage chunk size.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
run() { name=$1; shift
  env WEBP_B200_TRACE=1 "$@" python bench.py --distinct 64 --steps 1 --e2e-steps 3 --no-cpu-baseline > gpurun_out/$name.log 2> gpurun_out/$name.err
  echo "== $name"; grep trace gpurun_out/$name.err | tail -1 | cut -c1-330
  python - gpurun_out/$name.log <<'PY'
import json,sys
l=[x for x in open(sys.argv[1]) if x.startswith("{")]
d=json.loads(l[-1]); print(d["value"], d["e2e"]["value"], d["e2e"]["ms_per_step"])
PY
}
run e2e_w1 WEBP_B200_HOST_WAVES=1
run e2e_w2 WEBP_B200_HOST_WAVES=2
run e2e_w3 WEBP_B200_HOST_WAVES=3
run e2e_w2_c1024 WEBP_B200_HOST_WAVES=2 WEBP_B200_CHUNK_MB=1024
