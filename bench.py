#!/usr/bin/env python3
"""bench.py -- WebP lossy decode throughput (decoded Mpix/s to RGBA) on N B200s, next to the reference on the
box's host cores. Contract: see the task description; one JSON line on stdout from rank 0.

Workload (BASELINE.json configs[1]): a batch of 4096 synthetic 1920x1080 images, reference encoder q75 -m 4,
1 token partition, 1 segment, simple loop filter, decoded to MODE_RGBA. The images are procedural
(oracle/reftool.c, seeded) and encoded by the UNMODIFIED reference encoder from oracle/_ref at start-up;
`--distinct D` images are encoded per rank and tiled to the batch size (image i of the batch = i mod D), which
keeps start-up inside the time budget on a 16-core host. Every GPU decodes its own batch (shard by index, no
collective): weak scaling.

  value : device-resident throughput -- compressed bytes already in HBM, pixels left in HBM; one step = one
          WebPBatchDecode over the whole batch; timed with CUDA events on the library's own stream.
  e2e   : the C-ABI call a user makes, WebPDecodeBatch(host buffers in, host buffers out): H2D of the
          compressed files and D2H of all RGBA pixels inside the timed region (pinned memory).
  --impl reference : the reference's own WebPDecode, one image per thread on all host cores (oracle/_ref).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (width, height, default batch, encoder config factory name, colourspace name)
    "vp8_1080p_q75_m4_1part_simple_rgba": (1920, 1080, 4096, "cfg_simple_1part", "RGBA"),
    "vp8_1080p_q75_m4_8part_normal_rgba": (1920, 1080, 4096, "cfg_normal_8part", "RGBA"),
    "vp8_256x256_q80_rgbA": (256, 256, 65536, "cfg_default", "rgbA"),
    "vp8_4096x4096_q90_alpha_rgba": (4096, 4096, 256, "cfg_alpha_q90", "RGBA"),
    # not a BASELINE config: whole-picture VP8L through the same entry point, one thread per picture
    "vp8l_1080p_lossless_rgba": (1920, 1080, 1024, "cfg_lossless", "RGBA"),
}
METRIC = "webp_lossy_decode_mpix_per_s_rgba"
UNIT = "Mpix/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="vp8_1080p_q75_m4_1part_simple_rgba", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="images per GPU (0 = the workload's default)")
    ap.add_argument("--distinct", type=int, default=256, help="distinct images encoded per rank, tiled to --batch")
    ap.add_argument("--e2e-steps", type=int, default=-1, help="end-to-end steps (-1 = same as --steps)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target duration of the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def make_corpus(workload, distinct, rank, threads):
    """`distinct` seeded synthetic images encoded by the reference encoder (oracle/_ref)."""
    from oracle import refwebp as R
    w, h, _, cfgname, _ = WORKLOADS[workload]
    cfg = getattr(R, cfgname)()
    return R.encode_corpus(distinct, w, h, cfg, seed0=1 + 100000 * rank, alpha=("alpha" in workload), nthreads=threads)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.gpu, self.lines, self.proc = gpu, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        sm, mx, reasons = [], 0, set()
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1])); mx = max(mx, float(p[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def host_memory_budget():
    """Bytes of host memory this job may still take: min(MemAvailable, cgroup limit - usage)."""
    avail = None
    try:
        for ln in open("/proc/meminfo"):
            if ln.startswith("MemAvailable:"):
                avail = int(ln.split()[1]) * 1024
    except OSError:
        pass
    for lim_f, use_f in (("/sys/fs/cgroup/memory.max", "/sys/fs/cgroup/memory.current"),
                         ("/sys/fs/cgroup/memory/memory.limit_in_bytes", "/sys/fs/cgroup/memory/memory.usage_in_bytes")):
        try:
            lim = open(lim_f).read().strip()
            if lim != "max":
                room = int(lim) - int(open(use_f).read().strip())
                avail = room if avail is None else min(avail, room)
        except (OSError, ValueError):
            pass
    return avail


def cpu_baseline(datas, csp_ref, seconds, threads):
    """Reference WebPDecode, one image per thread on `threads` host threads, bounded sample (oracle/_ref)."""
    from oracle import refwebp as R
    probe = R.decode_bench(datas, threads, csp_ref, passes=1)
    per_pass = max(probe["seconds"], 1e-3)
    passes = max(1, min(200, int(seconds / per_pass)))
    r = R.decode_bench(datas, threads, csp_ref, passes=passes)
    return {"value": round(r["mpix_s"], 1), "unit": UNIT, "cores": threads, "kind": "reference",
            "sample": f"{len(datas)} distinct images x {passes} passes, {r['seconds']:.1f}s wall, "
                      f"one image per thread, SIMD on, external RGBA buffers, errors={r['errors']}"}


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation on all host cores; rank 0 only."""
    if rank != 0:
        return
    from oracle import refwebp as R
    if not R.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libwebp_ref.so missing (built where /root/reference exists)"}))
        return
    w, h, default_batch, _, cspname = WORKLOADS[args.workload]
    threads = os.cpu_count() or 1
    distinct = min(args.distinct, args.batch or default_batch)
    datas = make_corpus(args.workload, distinct, 0, threads)
    csp = getattr(R, "MODE_" + cspname)
    # each step = one pass over the bounded sample; W warm-up passes are folded into decode_bench's own warm-up
    probe = R.decode_bench(datas, threads, csp, passes=max(1, args.warmup))
    r = R.decode_bench(datas, threads, csp, passes=max(1, args.steps))
    ms = r["seconds"] / max(1, args.steps) * 1e3
    val = round(r["mpix_s"], 1)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": args.workload, "batch_per_gpu": args.batch or default_batch, "distinct_images": distinct,
                   "colorspace": cspname, "step": f"one pass over {distinct} distinct images on {threads} host threads"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "reference",
                         "sample": f"{distinct} distinct images x {args.steps} passes, one image per thread, SIMD on, errors={r['errors']}"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "warmup_mpix_s": round(probe["mpix_s"], 1),
    }
    print(json.dumps(line))


def main():
    args = parse_args()
    rank, local_rank, world = dist_env()
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    import numpy as np
    import torch
    import libwebp_b200 as W
    from oracle import refwebp as R   # corpus generation + CPU baseline only (never the product path)

    if W.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device; libwebp_b200 has no CPU path")
    use_dist = world > 1
    if use_dist:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    device = local_rank
    torch.cuda.set_device(device)

    w, h, default_batch, _, cspname = WORKLOADS[args.workload]
    batch_n = args.batch or default_batch
    distinct = min(args.distinct, batch_n)
    csp = getattr(W, "MODE_" + cspname)
    host_threads = max(1, (os.cpu_count() or 1) // world)

    t0 = time.time()
    corpus = make_corpus(args.workload, distinct, rank, host_threads)
    t_corpus = time.time() - t0
    datas = [corpus[i % distinct] for i in range(batch_n)]
    file_bytes = sum(len(d) for d in datas)
    mpix = batch_n * w * h * 1e-6

    def barrier():
        torch.cuda.synchronize()
        if use_dist:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(x):
        if not use_dist:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{device}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ------------------------------------------------------------------ device-resident throughput (`value`)
    res = W.Batch(datas, csp, device=device, output=W.WEBP_BATCH_DEVICE, pinned=True)
    st = res.create()
    if st != 0:
        raise SystemExit(f"WebPBatchCreate failed: {st} {W.last_error()}")
    for _ in range(max(args.warmup, 3)):
        if res.decode() != 0:
            raise SystemExit(f"decode failed: {res.statuses()[:8]} {W.last_error()}")
    sampler = ClockSampler(device)
    barrier()
    sampler.start()
    t_wall0 = time.perf_counter()
    dev_ms, stage = 0.0, {}
    launches = 0
    for _ in range(args.steps):
        if res.decode() != 0:
            raise SystemExit("decode failed in the timed region")
        t = res.timings()
        dev_ms += t["total_ms"]
        launches += t["launches"]
        for k in ("modes_ms", "tokens_ms", "recon_ms", "filter_ms", "emit_ms", "alpha_ms"):
            stage[k] = stage.get(k, 0.0) + t[k]
    barrier()
    wall_ms = (time.perf_counter() - t_wall0) * 1e3
    clocks = sampler.stop()
    dev_ms_step = reduce_max(dev_ms / args.steps)
    wall_ms_step = reduce_max(wall_ms / args.steps)
    res.close()

    # ------------------------------------------------------------------ end to end through the C ABI (`e2e`)
    e2e_steps = args.steps if args.e2e_steps < 0 else args.e2e_steps
    e2e = None
    if e2e_steps > 0:
        # page-locked input + output buffers of every rank must fit the host: shrink the e2e batch if they would not
        e2e_n = batch_n
        per_img = 4 * w * h + file_bytes // batch_n + 1024
        room = host_memory_budget()
        if room is not None:
            fit = int(0.5 * room / world / per_img)
            if fit < e2e_n:
                e2e_n = max(distinct, fit // distinct * distinct)
        e2e_datas = datas[:e2e_n]
        e2e_mpix = e2e_n * w * h * 1e-6
        hb = W.Batch(e2e_datas, csp, device=device, output=W.WEBP_BATCH_HOST, pinned=True)
        if hb.decode_oneshot() != 0:     # warm-up (also faults the pinned pages in)
            raise SystemExit(f"WebPDecodeBatch failed: {W.last_error()}")
        barrier()
        t1 = time.perf_counter()
        for _ in range(e2e_steps):
            if hb.decode_oneshot() != 0:
                raise SystemExit("WebPDecodeBatch failed in the timed region")
        barrier()
        e2e_ms_step = reduce_max((time.perf_counter() - t1) * 1e3 / e2e_steps)
        # spot-check the bytes that came back against the reference (not timed)
        ok = True
        for i in (0, e2e_n // 2, e2e_n - 1):
            s_ref, want = R.decode(e2e_datas[i], getattr(R, "MODE_" + cspname), 0)
            ok &= bool(s_ref == 0 and np.array_equal(hb.host_output(i), want))
        e2e = {"value": round(e2e_mpix * world / (e2e_ms_step * 1e-3), 1), "unit": UNIT, "h2d_bytes_per_step": hb.h2d_bytes,
               "d2h_bytes_per_step": hb.d2h_bytes, "ms_per_step": round(e2e_ms_step, 3), "steps": e2e_steps,
               "batch_per_gpu": e2e_n, "bit_exact_spot_check": ok,
               "api": "WebPDecodeBatch(host buffers in, host buffers out), pinned memory"}
        hb.close()

    if rank != 0:
        if use_dist:
            dist.destroy_process_group()
        return

    # ------------------------------------------------------------------ roofline of the dominant kernel + stages
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        hbm_peak, peak_src = float(peaks["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (copy, burst)"
    except Exception:
        hbm_peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    n_mb = batch_n * ((w + 15) // 16) * ((h + 15) // 16)
    px = batch_n * w * h
    per = {k: v / args.steps for k, v in stage.items()}
    # algorithmic bytes per launch (DESIGN.md "Kernels"): what each stage must move at minimum
    alg = {
        "modes_ms": file_bytes * 0.08 + 16 * n_mb,                    # partition 0 (~8% of the file) in, MbInfo out
        "tokens_ms": file_bytes + 2 * 8 * n_mb,                       # token partitions in, nz codes in/out
        "recon_ms": 1.5 * px + 16 * n_mb,                             # planes out (+ the non-zero coefficients in)
        "filter_ms": 3.0 * px,                                        # planes read + written
        "emit_ms": 5.5 * px,                                          # 1.5 B/px in, 4 B/px out
        # ALPH: one alpha byte per pixel out (the chunk itself is tiny); whole-picture VP8L: file in, 4 B/px out
        "alpha_ms": 1.0 * px if "alpha" in args.workload else (file_bytes + 4.0 * px) if "vp8l" in args.workload else 0.0,
    }
    kernels = {k[:-3]: {"ms": round(per[k], 3), "alg_GBps": round(alg[k] / (per[k] * 1e-3) / 1e9, 1) if per[k] > 0 else None,
                        "share": round(per[k] / max(sum(per.values()), 1e-9), 3)} for k in per}
    dom = max(per, key=per.get)
    achieved = alg[dom] / (per[dom] * 1e-3) / 1e9
    # DRAM traffic of the dominant kernel from the committed ncu capture of this workload (per launch), if one exists
    traffic, traffic_src = None, None
    try:
        for t in json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))):
            if t["kernel"] == "k_" + dom[:-3] and t["workload"] == args.workload and t["batch"] == batch_n:
                traffic, traffic_src = t["dram_bytes"], t["source"]
    except (OSError, ValueError, KeyError):
        pass
    sm_mhz = clocks.get("sm_mhz") or 1965.0
    bits_per_cycle_sm = (file_bytes * 8) / (per["tokens_ms"] * 1e-3 * sm_mhz * 1e6 * 148) if per["tokens_ms"] > 0 else None
    step_bytes = file_bytes + 4.0 * px
    line = {
        "metric": METRIC, "value": round(mpix * world / (dev_ms_step * 1e-3), 1), "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": round(dev_ms_step, 3),
        "wall_ms_per_step": round(wall_ms_step, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {"workload": args.workload, "batch_per_gpu": batch_n, "distinct_images": distinct, "width": w, "height": h,
                   "colorspace": cspname, "compressed_bytes_per_gpu": file_bytes, "bpp": round(file_bytes * 8 / px, 3),
                   "l2": "inputs_larger_than_l2", "sharding": "by image index, no collective",
                   "corpus_seconds": round(t_corpus, 1)},
        "e2e": e2e, "gpu_launches": launches,
        "roofline": {"kernel": "k_" + dom[:-3], "bound": "hbm", "achieved": round(achieved, 2), "peak": hbm_peak, "unit": "GB/s",
                     "frac": round(achieved / hbm_peak, 5), "traffic": traffic, "traffic_source": traffic_src,
                     "algorithmic_bytes": int(alg[dom]), "peak_source": peak_src,
                     "note": "serial boolean decoding: bound by the latency of one dependent chain per stream, not by HBM; see parse_bits_per_cycle_per_sm"},
        "roofline_step": {"bound": "hbm", "achieved": round(step_bytes / (dev_ms_step * 1e-3) / 1e9, 1), "peak": hbm_peak,
                          "unit": "GB/s", "frac": round(step_bytes / (dev_ms_step * 1e-3) / 1e9 / hbm_peak, 4),
                          "bytes": "compressed file + RGBA output per image"},
        "parse_bits_per_cycle_per_sm": round(bits_per_cycle_sm, 4) if bits_per_cycle_sm else None,
        "kernels": kernels, "clocks": clocks, "host_cores": os.cpu_count(),
    }
    if not args.no_cpu_baseline and world == 1 and R.available():
        line["cpu_baseline"] = cpu_baseline(corpus, getattr(R, "MODE_" + cspname), args.cpu_seconds, os.cpu_count() or 1)
    else:
        line["cpu_baseline"] = None
    print(json.dumps(line))
    if use_dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
