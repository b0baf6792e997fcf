#!/usr/bin/env python3
"""bench.py -- WebP lossy decode throughput (decoded Mpix/s to RGBA) on N B200s, next to the reference on the
box's host cores. Contract: see the task description; one JSON line on stdout from rank 0.

Headline workload (BASELINE.json configs[1]): a batch of 4096 synthetic 1920x1080 images, reference encoder q75 -m 4,
1 token partition, 1 segment, simple loop filter, decoded to MODE_RGBA. The images are procedural
(oracle/reftool.c, seeded) and encoded by the UNMODIFIED reference encoder from oracle/_ref at start-up;
`--distinct D` images are encoded per rank and tiled to the batch size (image i of the batch = i mod D), which
keeps start-up inside the time budget on a 16-core host. Every GPU decodes its own batch (shard by index, no
collective): weak scaling.

  value          device-resident throughput: compressed bytes already in HBM, pixels left in HBM; one step = one
                 WebPBatchDecode over the whole batch; timed with CUDA events on the library's own stream.
  value_with_h2d the same with the compressed files starting in page-locked HOST memory: WebPBatchCreate
                 (upload) + WebPBatchDecode per step, pixels left in HBM, wall clock.
  e2e            the C-ABI a caller uses for throughput, host buffers in and out: WebPBatchSubmit / WebPBatchWait with
                 two batches in flight (the pixels of batch k travel while batch k+1 is parsed), every step uploads its
                 compressed files and downloads all its pixels; wall clock over K batches including pipeline fill and
                 drain. `blocking_ms` beside it is one WebPDecodeBatch call on its own.
  other_workloads  BASELINE configs 3, 4, 5 (and whole-picture VP8L) in the same run, a few steps each.
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
    # name: (width, height, default batch, encoder config factory name, colourspace name, BASELINE.json config)
    "vp8_1080p_q75_m4_1part_simple_rgba": (1920, 1080, 4096, "cfg_simple_1part", "RGBA", 2),
    "vp8_1080p_q75_m4_8part_normal_rgba": (1920, 1080, 4096, "cfg_normal_8part", "RGBA", 3),
    "vp8_1080p_q75_m4_8part_normal_yuv": (1920, 1080, 4096, "cfg_normal_8part", "YUV", 3),
    "vp8_256x256_q80_rgbA": (256, 256, 65536, "cfg_default", "rgbA", 4),
    # 512 images: the token time of a batch is the time of ONE chain (22 M decodes here), so throughput grows with the batch until
    # every SM sub-partition has a stream (592); 512 is what 180 GB hold (1.96 KB of scratch per macroblock + the ARGB plane of
    # the alpha decode + the output: 279 MB per image)
    "vp8_4096x4096_q90_alpha_rgba": (4096, 4096, 512, "cfg_alpha_q90", "RGBA", 5),
    # not a BASELINE config: whole-picture VP8L through the same entry point
    "vp8l_1080p_lossless_rgba": (1920, 1080, 1024, "cfg_lossless", "RGBA", None),
}
# the end-to-end leg keeps TWO batches in flight (their outputs and alpha planes twice in HBM). Where one batch fills the device
# (4096x4096 + ALPH: 279 MB per image) it runs ONE batch at a time instead: halving the batch would halve the throughput of the
# parse (its time is the time of one chain, whatever the batch), which costs more than the overlap of the download gains.
E2E_SLOTS = {"vp8_4096x4096_q90_alpha_rgba": 1}
HEADLINE = "vp8_1080p_q75_m4_1part_simple_rgba"
OTHERS = ["vp8_1080p_q75_m4_8part_normal_rgba", "vp8_1080p_q75_m4_8part_normal_yuv", "vp8_256x256_q80_rgbA",
          "vp8_4096x4096_q90_alpha_rgba", "vp8l_1080p_lossless_rgba"]
METRIC = "webp_lossy_decode_mpix_per_s_rgba"
UNIT = "Mpix/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=HEADLINE, choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="images per GPU (0 = the workload's default)")
    ap.add_argument("--distinct", type=int, default=256, help="distinct images encoded per rank, tiled to --batch")
    ap.add_argument("--e2e-steps", type=int, default=-1, help="end-to-end steps (-1 = same as --steps)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target duration of the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-others", action="store_true", help="skip the other BASELINE configs (other_workloads)")
    ap.add_argument("--others-steps", type=int, default=2)
    return ap.parse_args()


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def workload_distinct(workload, distinct, batch_n):
    """Distinct images actually encoded: the big shapes take fewer so that start-up stays within seconds."""
    w, h = WORKLOADS[workload][:2]
    if w * h >= 4096 * 4096:
        distinct = min(distinct, 8)
    return max(1, min(distinct, batch_n))


def make_corpus(workload, distinct, rank, threads):
    """`distinct` seeded synthetic images encoded by the reference encoder (oracle/_ref)."""
    from oracle import refwebp as R
    w, h, _, cfgname, _, _ = WORKLOADS[workload]
    cfg = getattr(R, cfgname)()
    return R.encode_corpus(distinct, w, h, cfg, seed0=1 + 100000 * rank, alpha=("alpha" in workload), nthreads=threads)


def workload_config(workload, batch_n, distinct, file_bytes):
    """The `config` object: identical for both arms (the driver compares them)."""
    w, h, _, _, cspname, cfgno = WORKLOADS[workload]
    return {"workload": workload, "baseline_config": cfgno, "batch_per_gpu": batch_n, "distinct_images": distinct, "width": w,
            "height": h, "colorspace": cspname, "compressed_bytes_per_gpu": file_bytes,
            "bpp": round(file_bytes * 8 / (batch_n * w * h), 3), "l2": "inputs_larger_than_l2",
            "sharding": "by image index, no collective"}


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
                      f"one image per thread, SIMD on, external buffers, errors={r['errors']}"}


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation on all host cores; rank 0 only."""
    if rank != 0:
        return
    from oracle import refwebp as R
    if not R.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libwebp_ref.so missing (built where /root/reference exists)"}))
        return
    w, h, default_batch, _, cspname, _ = WORKLOADS[args.workload]
    threads = os.cpu_count() or 1
    batch_n = args.batch or default_batch
    distinct = workload_distinct(args.workload, args.distinct, batch_n)
    datas = make_corpus(args.workload, distinct, 0, threads)
    file_bytes = sum(len(datas[i % distinct]) for i in range(batch_n))
    csp = getattr(R, "MODE_" + cspname)
    # each step = one pass over the bounded sample; W warm-up passes first
    probe = R.decode_bench(datas, threads, csp, passes=max(1, args.warmup))
    r = R.decode_bench(datas, threads, csp, passes=max(1, args.steps))
    ms = r["seconds"] / max(1, args.steps) * 1e3
    val = round(r["mpix_s"], 1)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": workload_config(args.workload, batch_n, distinct, file_bytes),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "reference",
                         "sample": f"each step = one pass over the {distinct} distinct images of the workload on {threads} host threads, "
                                   f"{args.steps} passes, one image per thread, SIMD on, errors={r['errors']}"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "warmup_mpix_s": round(probe["mpix_s"], 1),
    }
    print(json.dumps(line))


# Shortest dependent chain of ONE boolean decode measured on this pool's B200 by tools/chain_floor.cu
# (profiles/r02b_chain_floor.json): the floor the token parse is held against (cycles per decode of one stream).
def chain_floor():
    try:
        r = json.load(open(os.path.join(ROOT, "profiles", "chain_floor.json")))
        return float(r["floor_cycles_per_decode"]), r["source"]
    except Exception:
        return 45.4, "fallback: booldec_fp.w1.l7 of profiles/r02b_chain_floor.json"


_ARENA = [None]   # one page-locked arena for the end-to-end buffers of every workload of the run


_NUMA = {}        # what gpu_numa_cpus() found, for the JSON line


def gpu_numa_cpus(device):
    """CPUs of the NUMA node this rank's GPU hangs off (None when the box has one node or does not say). Page-locked
    memory is placed where the allocating thread runs, and a device-to-host copy into the other socket's memory crosses
    the inter-socket link: with several ranks copying at once that link is shared on top of the PCIe uplinks."""
    try:
        import torch
        p = torch.cuda.get_device_properties(device)
        bus = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        nodes = [d for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit()]
        _NUMA.update({"gpu_pci": bus, "node": node, "nodes": len(nodes)})
        if node < 0 or len(nodes) < 2:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        _NUMA["cpus"] = len(cpus)
        return cpus or None
    except Exception as ex:   # never fail the bench over placement
        _NUMA["error"] = str(ex)[:80]
        return None


def host_arena(W, nbytes, device=0):
    a = _ARENA[0]
    if a is not None and a.buf.nbytes >= nbytes:
        a.reset()
        return a
    if a is not None:
        a.free()
    # allocate (= fault in and lock) the pages from a CPU next to the GPU, then give the thread its CPUs back
    old, near = None, gpu_numa_cpus(device)
    try:
        if near:
            old = os.sched_getaffinity(0)
            os.sched_setaffinity(0, near)
        _ARENA[0] = W.HostArena(nbytes)
    finally:
        if old is not None:
            os.sched_setaffinity(0, old)
    return _ARENA[0]


def measure(args, workload, rank, world, device, steps, warmup, e2e_steps, cpu_seconds, torch, dist, full):
    """One workload on this rank's device. Returns the result dict (rank 0) or None."""
    import numpy as np
    import libwebp_b200 as W
    from oracle import refwebp as R   # corpus generation + CPU baseline only (never the product path)
    from oracle import portwebp as P  # chain lengths (boolean decodes per stream) for the parse roofline
    use_dist = world > 1
    w, h, default_batch, _, cspname, _ = WORKLOADS[workload]
    batch_n = (args.batch if workload == args.workload else 0) or default_batch
    distinct = workload_distinct(workload, args.distinct, batch_n)
    csp = getattr(W, "MODE_" + cspname)
    host_threads = max(1, (os.cpu_count() or 1) // world)

    t0 = time.time()
    corpus = make_corpus(workload, distinct, rank, host_threads)
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
    for _ in range(max(warmup, 3)):
        if res.decode() != 0:
            raise SystemExit(f"decode failed: {res.statuses()[:8]} {W.last_error()}")
    sampler = ClockSampler(device)
    barrier()
    sampler.start()
    t_wall0 = time.perf_counter()
    dev_ms, stage = 0.0, {}
    launches = 0
    for _ in range(steps):
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
    dev_ms_step = reduce_max(dev_ms / steps)
    wall_ms_step = reduce_max(wall_ms / steps)
    res.destroy()

    # ------------------------------------------------------------------ the same with the inputs starting in host memory
    h2d_steps = max(1, min(steps, 5))
    barrier()
    t1 = time.perf_counter()
    h2d_each = []
    for _ in range(h2d_steps):
        ts = time.perf_counter()
        if res.create() != 0 or res.decode() != 0:
            raise SystemExit("create+decode failed in the timed region")
        res.destroy()
        h2d_each.append(round((time.perf_counter() - ts) * 1e3, 1))
    barrier()
    h2d_ms_step = reduce_max((time.perf_counter() - t1) * 1e3 / h2d_steps)
    res.close()

    # ------------------------------------------------------------------ end to end through the C ABI (`e2e`)
    e2e = None
    if e2e_steps > 0:
        # two batches in flight, each with its own page-locked input + output buffers: shrink the batch if the host cannot hold them
        e2e_n = batch_n
        nslots = E2E_SLOTS.get(workload, 2)
        per_img = nslots * (W.out_bytes(csp, w, h) + file_bytes // batch_n + 1024)
        room = host_memory_budget()
        if room is not None:
            if _ARENA[0] is not None:
                room += _ARENA[0].buf.nbytes   # already ours: reused by this workload
            fit = int(0.6 * room / world / per_img)
            if fit < e2e_n:
                e2e_n = max(distinct, fit // distinct * distinct)
        e2e_datas = datas[:e2e_n]
        e2e_mpix = e2e_n * w * h * 1e-6
        slot_bytes = e2e_n * (((W.out_bytes(csp, w, h) + 255) & ~255)) + sum((len(d) + 15) & ~15 for d in e2e_datas) + (1 << 20)
        arena = host_arena(W, nslots * slot_bytes, device)
        slots = [W.Batch(e2e_datas, csp, device=device, output=W.WEBP_BATCH_HOST, pinned=True, arena=arena) for _ in range(nslots)]
        # warm-up: one blocking call per slot (also faults the pinned pages in), then the blocking latency on its own
        for sl in slots:
            if sl.decode_oneshot() != 0:
                raise SystemExit(f"WebPDecodeBatch failed: {W.last_error()}")
        barrier()
        tb = time.perf_counter()
        if slots[0].decode_oneshot() != 0:
            raise SystemExit("WebPDecodeBatch failed")
        blocking_ms = reduce_max((time.perf_counter() - tb) * 1e3)
        barrier()
        t1 = time.perf_counter()
        inflight = []
        host_submit_ms, host_wait_ms = [], []
        for k in range(e2e_steps):
            sl = slots[k % nslots]
            if len(inflight) == nslots:
                tw = time.perf_counter()
                if inflight.pop(0).wait() != 0:
                    raise SystemExit("WebPBatchWait failed in the timed region")
                host_wait_ms.append(round((time.perf_counter() - tw) * 1e3, 1))
            tw = time.perf_counter()
            if sl.submit() != 0:
                raise SystemExit(f"WebPBatchSubmit failed in the timed region: {W.last_error()}")
            host_submit_ms.append(round((time.perf_counter() - tw) * 1e3, 1))
            inflight.append(sl)
        for sl in inflight:
            tw = time.perf_counter()
            if sl.wait() != 0:
                raise SystemExit("WebPBatchWait failed in the timed region")
            host_wait_ms.append(round((time.perf_counter() - tw) * 1e3, 1))
        barrier()
        e2e_ms_step = reduce_max((time.perf_counter() - t1) * 1e3 / e2e_steps)
        # spot-check the bytes that came back against the reference (not timed)
        ok = True
        for sl in slots[:min(2, e2e_steps)]:
            for i in (0, e2e_n // 2, e2e_n - 1):
                s_ref, want = R.decode(e2e_datas[i], getattr(R, "MODE_" + cspname), 0)
                ok &= bool(s_ref == 0 and np.array_equal(sl.host_output(i), want))
        e2e = {"value": round(e2e_mpix * world / (e2e_ms_step * 1e-3), 1), "unit": UNIT, "h2d_bytes_per_step": slots[0].h2d_bytes,
               "d2h_bytes_per_step": slots[0].d2h_bytes, "ms_per_step": round(e2e_ms_step, 3), "steps": e2e_steps,
               "batch_per_gpu": e2e_n, "bit_exact_spot_check": ok, "blocking_ms": round(blocking_ms, 3),
               "host_submit_ms": host_submit_ms[:8], "host_wait_ms": host_wait_ms[:8],
               "d2h_floor_ms": None,
               "api": "WebPBatchSubmit/WebPBatchWait, %s, host buffers in and out (pinned); "
                      "blocking_ms = one WebPDecodeBatch call" % ("two batches in flight" if nslots == 2 else "one batch at a time (one fills the device)")}
        # bare D2H ceiling of this box at this N: the same bytes, nothing else running (read against e2e at N = 1/2/4/8)
        try:
            nb = min(slots[0].d2h_bytes, 8 << 30)
            dev_t = torch.empty(nb, dtype=torch.uint8, device=f"cuda:{device}")
            host_t = torch.from_numpy(slots[0]._out_arr[:nb])
            barrier()
            tc = time.perf_counter()
            host_t.copy_(dev_t, non_blocking=True)
            barrier()
            gbs = nb / (time.perf_counter() - tc) / 1e9
            gbs = 1.0 / reduce_max(1.0 / gbs)
            e2e["d2h_GBps_bare"] = round(gbs, 1)
            e2e["d2h_floor_ms"] = round(slots[0].d2h_bytes / gbs / 1e6, 1)
            del dev_t
        except Exception as ex:   # never fail the bench over the ceiling probe
            e2e["d2h_GBps_bare"] = f"probe failed: {ex}"
        for sl in slots:
            sl.close()

    if rank != 0:
        return None

    # ------------------------------------------------------------------ per-kernel figures
    n_mb = batch_n * ((w + 15) // 16) * ((h + 15) // 16)
    px = batch_n * w * h
    per = {k: v / steps for k, v in stage.items()}
    lossless = "vp8l" in workload
    # algorithmic bytes per launch (DESIGN.md "Kernels"): what each stage must move at minimum
    out_bpp = W.out_bytes(csp, w, h) / (w * h)
    alg = {
        "modes_ms": file_bytes * 0.08 + 16 * n_mb,                    # partition 0 (~8% of the file) in, MbInfo out
        "tokens_ms": file_bytes + 2 * 8 * n_mb,                       # token partitions in, nz codes in/out
        "recon_ms": 1.5 * px + 16 * n_mb,                             # planes out (+ the non-zero coefficients in)
        "filter_ms": 3.0 * px,                                        # planes read + written
        "emit_ms": (1.5 + out_bpp) * px,                              # 1.5 B/px in, the output out
        # ALPH: one alpha byte per pixel out (the chunk itself is tiny); whole-picture VP8L: file in, 4 B/px out
        "alpha_ms": 1.0 * px if "alpha" in workload else (file_bytes + 4.0 * px) if lossless else 0.0,
    }
    if lossless:
        for k in ("modes_ms", "tokens_ms", "recon_ms", "filter_ms", "emit_ms"):
            alg[k] = 0.0
    kernels = {k[:-3]: {"ms": round(per[k], 3), "alg_GBps": round(alg[k] / (per[k] * 1e-3) / 1e9, 1) if per[k] > 0.05 else None,
                        "share": round(per[k] / max(sum(per.values()), 1e-9), 3)} for k in per}
    sm_mhz = clocks.get("sm_mhz") or 1965.0
    # the token parse against its own bound: cycles per boolean decode of the longest chain (streams advance in lockstep)
    parse = None
    if not lossless and per["tokens_ms"] > 0.05:
        chains = []
        for d in corpus[:min(distinct, 16)]:
            stc, p0, parts = P.count_decodes(d)
            if stc == 0 and parts:
                chains.append((p0, max(parts), sum(parts)))
        if chains:
            longest = sum(c[1] for c in chains) / len(chains)
            floor, floor_src = chain_floor()
            cyc = per["tokens_ms"] * 1e-3 * sm_mhz * 1e6 / longest
            parse = {"decodes_per_chain": int(longest), "decodes_per_image": int(sum(c[2] for c in chains) / len(chains)),
                     "first_partition_decodes": int(sum(c[0] for c in chains) / len(chains)),
                     "cycles_per_decode": round(cyc, 1), "floor_cycles_per_decode": floor, "frac_of_floor": round(floor / cyc, 4),
                     "floor_source": floor_src,
                     "bits_per_cycle_per_sm": round((file_bytes * 8) / (per["tokens_ms"] * 1e-3 * sm_mhz * 1e6 * 148), 4)}
    out = {
        "value": round(mpix * world / (dev_ms_step * 1e-3), 1), "ms_per_step": round(dev_ms_step, 3),
        "wall_ms_per_step": round(wall_ms_step, 3),
        "value_with_h2d": round(mpix * world / (h2d_ms_step * 1e-3), 1), "h2d_ms_per_step": round(h2d_ms_step, 3), "h2d_ms_each": h2d_each,
        "config": workload_config(workload, batch_n, distinct, file_bytes), "e2e": e2e, "gpu_launches": launches,
        "kernels": kernels, "parse": parse, "clocks": clocks, "corpus_seconds": round(t_corpus, 1),
        "_per": per, "_alg": alg, "_corpus": corpus, "_csp_ref": getattr(R, "MODE_" + cspname),
    }
    if not args.no_cpu_baseline and world == 1 and R.available() and cpu_seconds > 0:
        out["cpu_baseline"] = cpu_baseline(corpus, getattr(R, "MODE_" + cspname), cpu_seconds, os.cpu_count() or 1)
    else:
        out["cpu_baseline"] = None
    return out


def main():
    args = parse_args()
    rank, local_rank, world = dist_env()
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    import torch
    import libwebp_b200 as W

    if W.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device; libwebp_b200 has no CPU path")
    use_dist = world > 1
    dist = None
    if use_dist:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    device = local_rank
    torch.cuda.set_device(device)
    # a throughput caller lets the library keep its device blocks between batches (default: a quarter of the device)
    W.set_cache_limit(int(0.85 * torch.cuda.get_device_properties(device).total_memory), device)

    e2e_steps = args.steps if args.e2e_steps < 0 else args.e2e_steps
    main_r = measure(args, args.workload, rank, world, device, args.steps, args.warmup, e2e_steps, args.cpu_seconds, torch, dist, True)
    others = {}
    if not args.no_others and args.workload == HEADLINE:
        for wl in OTHERS:
            if world > 1 and WORKLOADS[wl][5] != 4:
                continue   # beyond one GPU only the thumbnail config (BASELINE config 4 names 1/2/4/8 GPUs) rides along
            W.trim_cache(device)
            try:
                r = measure(args, wl, rank, world, device, args.others_steps, 3, min(e2e_steps, max(2, args.others_steps)),
                            min(args.cpu_seconds, 4.0), torch, dist, False)
            except SystemExit as ex:
                r = {"error": str(ex)} if rank == 0 else None
            if r is not None:
                others[wl] = {k: v for k, v in r.items() if not k.startswith("_") and k != "clocks"}
    if rank != 0:
        if use_dist:
            dist.destroy_process_group()
        return

    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        hbm_peak, peak_src = float(peaks["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (copy, burst)"
    except Exception:
        hbm_peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    r = main_r
    per, alg = r["_per"], r["_alg"]
    dom = max(per, key=per.get)
    batch_n = r["config"]["batch_per_gpu"]
    w, h = r["config"]["width"], r["config"]["height"]
    file_bytes = r["config"]["compressed_bytes_per_gpu"]
    # DRAM traffic of the dominant kernel from the committed ncu capture of this workload (per launch), if one exists
    traffic, traffic_src = None, None
    try:
        for t in json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))):
            if t["kernel"] == "k_" + dom[:-3] and t["workload"] == args.workload and t["batch"] == batch_n:
                traffic, traffic_src = t["dram_bytes"], t["source"]
    except (OSError, ValueError, KeyError):
        pass
    hbm_achieved = alg[dom] / (per[dom] * 1e-3) / 1e9
    if dom == "tokens_ms" and r["parse"]:
        # the serial boolean decode is bound by the latency of its dependent chain, not by bytes: held against the
        # shortest chain of one decode measured on this hardware (tools/chain_floor.cu)
        roof = {"kernel": "k_parse_tokens", "bound": "latency", "achieved": r["parse"]["cycles_per_decode"],
                "peak": r["parse"]["floor_cycles_per_decode"], "unit": "cycles per boolean decode of one stream (lower is better)",
                "frac": r["parse"]["frac_of_floor"], "peak_source": r["parse"]["floor_source"],
                "traffic": traffic, "traffic_source": traffic_src, "algorithmic_bytes": int(alg[dom]),
                "hbm": {"achieved_GBps": round(hbm_achieved, 2), "peak_GBps": hbm_peak, "frac": round(hbm_achieved / hbm_peak, 5),
                        "peak_source": peak_src}}
    else:
        roof = {"kernel": "k_" + dom[:-3], "bound": "hbm", "achieved": round(hbm_achieved, 2), "peak": hbm_peak, "unit": "GB/s",
                "frac": round(hbm_achieved / hbm_peak, 5), "traffic": traffic, "traffic_source": traffic_src,
                "algorithmic_bytes": int(alg[dom]), "peak_source": peak_src}
    step_bytes = file_bytes + 4.0 * batch_n * w * h
    line = {
        "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": r["ms_per_step"],
        "wall_ms_per_step": r["wall_ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic", "config": r["config"],
        "value_with_h2d": r["value_with_h2d"], "h2d_ms_per_step": r["h2d_ms_per_step"], "h2d_ms_each": r.get("h2d_ms_each"),
        "e2e": r["e2e"], "gpu_launches": r["gpu_launches"], "roofline": roof,
        "roofline_step": {"bound": "hbm", "achieved": round(step_bytes / (r["ms_per_step"] * 1e-3) / 1e9, 1), "peak": hbm_peak,
                          "unit": "GB/s", "frac": round(step_bytes / (r["ms_per_step"] * 1e-3) / 1e9 / hbm_peak, 4),
                          "bytes": "compressed file + RGBA output per image"},
        "parse": r["parse"], "parse_bits_per_cycle_per_sm": r["parse"]["bits_per_cycle_per_sm"] if r["parse"] else None,
        "kernels": r["kernels"], "clocks": r["clocks"], "host_cores": os.cpu_count(), "corpus_seconds": r["corpus_seconds"],
        "host_numa": dict(_NUMA) or None,   # where rank 0's page-locked end-to-end buffers were placed (gpu_numa_cpus)
        "cpu_baseline": r["cpu_baseline"], "other_workloads": others,
    }
    print(json.dumps(line))
    if use_dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
