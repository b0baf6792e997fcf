#!/usr/bin/env python
"""Damage campaign on the GPU (test infrastructure): the mutants of tools/fuzz_emu.py go through the product itself --
WebPDecodeBatch over the C ABI, the real kernels -- in batches, and through the compiled reference (oracle/_ref) on the
host. Per-item status must equal the reference's and the pixels must be equal wherever the reference decodes; one damaged
file must never disturb its neighbours in the batch (every batch also carries intact files). The known damaged-file
deviations of DESIGN.md section 5 are counted separately.

    python tools/fuzz_gpu.py --seconds 40 --batch 2048 [--seed 1]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=40)
    ap.add_argument("--batch", type=int, default=2048)
    ap.add_argument("--seed", type=int, default=1)
    a = ap.parse_args()
    import libwebp_b200 as W
    from oracle import refwebp as R
    import fuzz_emu as F
    S = F.seeds(["lossy", "alpha", "lossless"])
    rng = np.random.default_rng(a.seed)
    hist, bad = {}, []
    cases = batches = 0
    t0 = time.time()
    while time.time() - t0 < a.seconds:
        items = []
        while len(items) < a.batch:
            kind, name, data = S[int(rng.integers(0, len(S)))]
            b = data if len(items) % 8 == 0 else F.mutate(rng, data)      # every eighth file is intact
            sf, f = R.features(b)
            if sf == 0 and f["width"] * f["height"] > F.MAX_PIXELS:
                continue
            items.append((name, data, b, sf, f))
        sts, outs = W.decode_batch([it[2] for it in items], W.MODE_RGBA)
        batches += 1
        for i, (name, data, b, sf, f) in enumerate(items):
            s_ref, want = R.decode(b, R.MODE_RGBA, 0)
            s_gpu = int(sts[i])
            cases += 1
            hist[s_ref] = hist.get(s_ref, 0) + 1
            same = s_gpu == s_ref and (s_ref != 0 or np.array_equal(want.reshape(-1), outs[i].reshape(-1)))
            if not same and s_ref == 0 and s_gpu == 0:
                _, want_c = R.decode(b, R.MODE_RGBA, 0, simd=False)     # the reference's C dsp path (class 1)
                if np.array_equal(want_c.reshape(-1), outs[i].reshape(-1)):
                    hist["simd_vs_c"] = hist.get("simd_vs_c", 0) + 1
                    same = True
            if same:
                continue
            tag = "gpu_%s_%d_%d" % (name, batches, i)
            bad.append((tag, s_ref, s_gpu))
            os.makedirs(os.path.join(ROOT, "gpurun_out", "fuzz"), exist_ok=True)
            open(os.path.join(ROOT, "gpurun_out", "fuzz", tag + ".webp"), "wb").write(b)
    for t in bad[:50]:
        print("MISMATCH file=%s ref=%d gpu=%d" % t)
    print(json.dumps({"cases": cases, "batches": batches, "batch": a.batch, "mismatches": len(bad),
                      "reference_status_histogram": {str(k): v for k, v in sorted(hist.items(), key=lambda kv: str(kv[0]))},
                      "seed": a.seed, "seconds": a.seconds}))
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
