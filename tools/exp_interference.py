#!/usr/bin/env python3
"""Does the token parser slow down beside a device->host copy?  ("what's weak" 7: the claim in
vp8_batch.cu was taken from the row-band experiment, which also relaunches the parser and lets the host block in
cudaMemcpy2DAsync between launches.)  Here the parser runs exactly as in the bench (one launch, whole batch, pixels left
in HBM) while an UNRELATED pinned-memory copy runs on another stream, issued by another host thread:

  alone        decode only
  beside_d2h   decode while 55 GB/s worth of device->host copies are in flight
  beside_h2d   the same with host->device copies

Prints per-stage device times of WebPBatchDecode for the three cases.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--distinct", type=int, default=64)
    ap.add_argument("--copy-gb", type=float, default=4.0)
    args = ap.parse_args()
    import torch
    import libwebp_b200 as W
    from oracle import refwebp as R
    corpus = R.encode_corpus(args.distinct, 1920, 1080, R.cfg_simple_1part(), seed0=1, alpha=False, nthreads=os.cpu_count() or 1)
    datas = [corpus[i % args.distinct] for i in range(args.batch)]
    torch.cuda.set_device(0)
    res = W.Batch(datas, W.MODE_RGBA, device=0, output=W.WEBP_BATCH_DEVICE, pinned=True)
    if res.create() != 0:
        raise SystemExit("create failed")
    for _ in range(3):
        res.decode()
    n = int(args.copy_gb * (1 << 30))
    dev = torch.empty(n, dtype=torch.uint8, device="cuda")
    host = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    cs = torch.cuda.Stream()
    stop = threading.Event()
    moved = [0]

    def pump(direction):
        with torch.cuda.stream(cs):
            while not stop.is_set():
                if direction == "d2h":
                    host.copy_(dev, non_blocking=True)
                else:
                    dev.copy_(host, non_blocking=True)
                cs.synchronize()
                moved[0] += n

    out = {}
    for case in ("alone", "beside_d2h", "beside_h2d", "alone_again"):
        stop.clear(); moved[0] = 0
        th = None
        if case.startswith("beside"):
            th = threading.Thread(target=pump, args=(case[-3:],)); th.start()
            time.sleep(0.3)
        t0 = time.perf_counter()
        acc = {}
        for _ in range(3):
            if res.decode() != 0:
                raise SystemExit("decode failed")
            for k, v in res.timings().items():
                acc[k] = acc.get(k, 0.0) + v / 3
        dt = time.perf_counter() - t0
        stop.set()
        if th is not None:
            th.join()
        acc = {k: round(v, 2) for k, v in acc.items()}
        acc["copy_GBps_during"] = round(moved[0] / dt / 1e9, 1) if th is not None else 0.0
        out[case] = acc
        print(case, json.dumps(acc), flush=True)
    res.close()


if __name__ == "__main__":
    main()
