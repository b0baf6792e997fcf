mkdir -p gpurun_out
show() { python - "$1" <<'PY'
import json,sys
l=[x for x in open(sys.argv[1]) if x.startswith("{")]
d=json.loads(l[-1]) if l else None
print(sys.argv[1], d and (d["value"], d["ms_per_step"], {k:v["ms"] for k,v in d["kernels"].items()}))
PY
}
timeout 300 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
B="python bench.py --distinct 64 --steps 3 --e2e-steps 0 --no-cpu-baseline"
WEBP_B200_TOKEN_MAP=lanes WEBP_B200_TOKEN_LPW=4 $B > gpurun_out/h_l4.log 2>&1; show gpurun_out/h_l4.log
WEBP_B200_TOKEN_MAP=lanes WEBP_B200_TOKEN_LPW=8 $B > gpurun_out/h_l8.log 2>&1; show gpurun_out/h_l8.log
T="$B --workload vp8_256x256_q80_rgbA --distinct 512"
WEBP_B200_TOKEN_MAP=warp $T > gpurun_out/t_warp.log 2>&1; show gpurun_out/t_warp.log
for l in 8 16 32; do WEBP_B200_TOKEN_MAP=lanes WEBP_B200_TOKEN_LPW=$l $T > gpurun_out/t_l$l.log 2>&1; show gpurun_out/t_l$l.log; done
W="$B --workload vp8_1080p_q75_m4_8part_normal_rgba"
WEBP_B200_TOKEN_MAP=warp $W > gpurun_out/p8_warp.log 2>&1; show gpurun_out/p8_warp.log
for l in 8 32; do WEBP_B200_TOKEN_MAP=lanes WEBP_B200_TOKEN_LPW=$l $W > gpurun_out/p8_l$l.log 2>&1; show gpurun_out/p8_l$l.log; done
tail -3 gpurun_out/t_l32.log | cut -c1-300
