#!/usr/bin/env python
"""Encoder-space campaign (CPU, test infrastructure): pictures of random size and content encoded by the reference encoder
with random settings (quality 0..100, method 0..6, 1..4 segments, both loop filters at every strength and sharpness, 1..8
token partitions, sns, alpha quality / filtering, lossless) are decoded by the compiled reference (oracle/_ref) and by the
host build of the device code (tests/emu) with random decoding options: every colourspace, crop window, flip,
no_fancy_upsampling, bypass_filtering, each token-parse run style, use_scaling (up, down, one axis derived, 1x1, on a crop
window), dithering_strength and alpha_dithering_strength. Status and bytes must be equal. Exit code 1 otherwise.

    python tools/fuzz_encode.py --seconds 300 --jobs 7 [--seed 1]
"""
import argparse
import ctypes as C
import json
import multiprocessing as mp
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
EMU_DIR = os.path.join(ROOT, "tests", "emu")
CSPS = (1, 7, 0, 2, 3, 8, 4, 9, 5, 6, 10, 11, 12)


def picture(R, rng, w, h, alpha):
    kind = int(rng.integers(0, 5))
    pix = np.zeros((h, w, 4), np.uint8)
    if kind == 0:
        pix[..., :3] = R.synth(w, h, int(rng.integers(0, 1 << 30)))
    elif kind == 1:      # noise: every token category, large coefficients at high quality
        pix[..., :3] = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    elif kind == 2:      # flat areas and hard edges: skipped macroblocks, i16 modes
        pix[..., :3] = int(rng.integers(0, 256))
        for _ in range(int(rng.integers(1, 6))):
            x0, y0 = int(rng.integers(0, w)), int(rng.integers(0, h))
            pix[y0:y0 + int(rng.integers(1, h + 1)), x0:x0 + int(rng.integers(1, w + 1)), :3] = rng.integers(0, 256, 3, dtype=np.uint8)
    elif kind == 3:      # smooth ramps
        y, x = np.mgrid[0:h, 0:w].astype(np.float64)
        for c in range(3):
            pix[..., c] = np.clip(128 + 127 * np.sin(x * rng.uniform(0.005, 0.2) + y * rng.uniform(0.005, 0.2) + c), 0, 255)
    else:                # extremes: saturated colours against each other (clipping in the transforms and the colour conversion)
        pix[..., :3] = rng.choice(np.array([0, 255], np.uint8), (h, w, 3))
        bs = int(rng.integers(1, 9))
        pix[..., :3] = np.kron(pix[::bs, ::bs, :3], np.ones((bs, bs, 1), np.uint8))[:h, :w]
    if alpha:
        ak = int(rng.integers(0, 3))
        if ak == 0:
            pix[..., 3] = rng.integers(0, 256, (h, w), dtype=np.uint8)
        elif ak == 1:
            y, x = np.mgrid[0:h, 0:w]
            pix[..., 3] = np.clip((x * 255) // max(1, w - 1) + rng.integers(-20, 20), 0, 255)
        else:
            pix[..., 3] = rng.choice(np.array([0, 255, 128], np.uint8), (h, w))
    else:
        pix[..., 3] = 255
    return pix


def worker(args):
    wid, seed, seconds = args
    from oracle import refwebp as R
    L = C.CDLL(os.path.join(EMU_DIR, "libvp8_emu.so"))
    L.emu_decode.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p]
    L.emu_decode_window.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int,
                                    C.c_int, C.c_int]
    L.emu_decode_scaled.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int,
                                    C.c_int, C.c_int, C.c_int, C.c_int]
    L.emu_decode_dithered.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int,
                                      C.c_int, C.c_int, C.c_int]
    L.emu_set_alpha_dithering.argtypes = [C.c_int]
    rng = np.random.default_rng(seed * 1000 + wid)
    t0 = time.time()
    files = decodes = 0
    bad = []
    while time.time() - t0 < seconds:
        w = int(rng.integers(1, 40)) if rng.random() < 0.3 else int(rng.integers(1, 420))
        h = int(rng.integers(1, 40)) if rng.random() < 0.3 else int(rng.integers(1, 300))
        alpha = rng.random() < 0.3
        lossless = rng.random() < 0.15
        kw = dict(segments=int(rng.integers(1, 5)), filter_type=int(rng.integers(0, 2)), filter_strength=int(rng.integers(0, 101)),
                  filter_sharpness=int(rng.integers(0, 8)), partitions=int(rng.integers(0, 4)), sns_strength=int(rng.integers(0, 101)))
        if alpha:
            kw.update(alpha_filtering=int(rng.integers(0, 3)), alpha_quality=int(rng.choice([100, 100, 90, 50, 10, 0])))
        if lossless:
            kw = dict(lossless=1)
        q = float(rng.choice([0, 1, 5, 20, 50, 75, 90, 98, 100])) if rng.random() < 0.5 else float(rng.integers(0, 101))
        cfg = R.EncCfg(q, int(rng.integers(0, 7)), **kw)
        data = R.encode(picture(R, rng, w, h, alpha), cfg)
        files += 1
        # options.use_scaling (on a crop window half of the time) and options.dithering_strength / alpha_dithering_strength
        for it in range(2):
            csp = CSPS[int(rng.integers(0, len(CSPS)))]
            crop = None
            if rng.random() < 0.5:
                cw, ch = int(rng.integers(1, w + 1)), int(rng.integers(1, h + 1))
                crop = (int(rng.integers(0, w - cw + 1)), int(rng.integers(0, h - ch + 1)), cw, ch)
            cw, ch = (crop[2], crop[3]) if crop else (w, h)
            c = crop or (0, 0, 0, 0)
            flip = int(rng.integers(0, 2))
            if it == 0:
                req = [(int(rng.integers(1, 2 * cw + 2)), int(rng.integers(1, 2 * ch + 2))), (0, int(rng.integers(1, 2 * ch + 2))),
                       (int(rng.integers(1, 2 * cw + 2)), 0), (1, 1), (int(rng.integers(1, 9)), int(rng.integers(1, 9))),
                       (cw, ch), (int(rng.integers(1, 6 * cw + 2)), int(rng.integers(1, 3)))][int(rng.integers(0, 7))]
                s_ref, (sw, sh), want = R.decode_scaled(data, csp, 8 if flip else 0, crop, req)
                bpp = 1 if csp in (11, 12) else R.BPP[csp]
                n = (sw * sh + 2 * ((sw + 1) // 2) * ((sh + 1) // 2) + (sw * sh if csp == 12 else 0)) if csp in (11, 12) else sw * sh * bpp
                out = np.zeros(max(n, 16), np.uint8)
                s_emu = L.emu_decode_scaled(data, len(data), csp, 4 if flip else 0, out.ctypes.data, out.size, sw * bpp,
                                            c[0], c[1], c[2], c[3], sw, sh) if s_ref == 0 else s_ref
                what = "scale%s" % (req,)
            else:
                if lossless:
                    continue
                strength, astrength = int(rng.choice([50, 100, 30, 1])), int(rng.choice([0, 100, 50])) if alpha else 0
                s_ref, want = R.decode_dithered(data, csp, 8 if flip else 0, crop, strength, astrength)
                bpp = 1 if csp in (11, 12) else R.BPP[csp]
                n = want.size if s_ref == 0 else 16
                out = np.zeros(max(n, 16), np.uint8)
                L.emu_set_alpha_dithering(astrength)
                s_emu = L.emu_decode_dithered(data, len(data), csp, 4 if flip else 0, out.ctypes.data, out.size, cw * bpp, strength,
                                              c[0], c[1], c[2], c[3])
                L.emu_set_alpha_dithering(0)
                what = "dither(%d,%d)" % (strength, astrength)
            decodes += 1
            if s_emu != s_ref or s_ref != 0 or not np.array_equal(want.reshape(-1)[:n], out[:n]):
                tag = "enc_w%d_%d_%s" % (wid, files, "s" if it == 0 else "d")
                bad.append((tag, w, h, q, cfg.method, json.dumps(kw) + " " + what, csp, str(crop), flip, 0, 0, 0, s_ref, s_emu))
                os.makedirs(os.path.join(ROOT, "gpurun_out", "fuzz"), exist_ok=True)
                open(os.path.join(ROOT, "gpurun_out", "fuzz", tag + ".webp"), "wb").write(data)
        for it in range(4):
            csp = CSPS[int(rng.integers(0, len(CSPS)))]
            crop = None
            if it >= 2:
                cw, ch = int(rng.integers(1, w + 1)), int(rng.integers(1, h + 1))
                crop = (int(rng.integers(0, w - cw + 1)), int(rng.integers(0, h - ch + 1)), cw, ch)
            flip, nofancy, bypass = int(rng.integers(0, 2)), int(rng.integers(0, 2)), int(rng.random() < 0.2)
            s_ref, want = R.decode_window(data, csp, (8 if flip else 0) | (2 if nofancy else 0) | bypass, crop)
            ow, oh = (crop[2], crop[3]) if crop else (w, h)
            bpp = 1 if csp in (11, 12) else R.BPP[csp]
            n = (ow * oh + 2 * ((ow + 1) // 2) * ((oh + 1) // 2) + (ow * oh if csp == 12 else 0)) if csp in (11, 12) else ow * oh * bpp
            out = np.zeros(max(n, 16), np.uint8)
            c = crop or (0, 0, 0, 0)
            dev_flags = (4 if flip else 0) | (2 if nofancy else 0) | bypass
            variant = 0
            if crop is None and not flip and csp in (1, 11) and not lossless:   # the run styles of the token parse (whole picture only)
                variant = (0, 1, 2, 4, 8, 9, 24, 25, 56)[int(rng.integers(0, 9))]
                s_emu = L.emu_decode(data, len(data), csp, dev_flags, out.ctypes.data, out.size, ow * bpp, variant, None)
            else:
                s_emu = L.emu_decode_window(data, len(data), csp, dev_flags, out.ctypes.data, out.size, ow * bpp, c[0], c[1], c[2], c[3])
            decodes += 1
            if s_emu != s_ref or s_ref != 0 or not np.array_equal(want.reshape(-1)[:n], out[:n]):
                tag = "enc_w%d_%d_%d" % (wid, files, it)
                bad.append((tag, w, h, q, cfg.method, json.dumps(kw), csp, str(crop), flip, nofancy, bypass, variant, s_ref, s_emu))
                os.makedirs(os.path.join(ROOT, "gpurun_out", "fuzz"), exist_ok=True)
                open(os.path.join(ROOT, "gpurun_out", "fuzz", tag + ".webp"), "wb").write(data)
    return files, decodes, bad


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=60)
    ap.add_argument("--jobs", type=int, default=max(1, (os.cpu_count() or 2) - 1))
    ap.add_argument("--seed", type=int, default=1)
    a = ap.parse_args()
    subprocess.check_call(["make", "-s", "-C", EMU_DIR])
    with mp.get_context("spawn").Pool(a.jobs) as p:
        res = p.map(worker, [(i, a.seed, a.seconds) for i in range(a.jobs)])
    bad = [t for r in res for t in r[2]]
    for t in bad:
        print("MISMATCH %s %dx%d q%.0f m%d %s csp=%d crop=%s flip=%d nofancy=%d bypass=%d variant=%d ref=%d emu=%d" % t)
    print(json.dumps({"files": sum(r[0] for r in res), "decodes": sum(r[1] for r in res), "mismatches": len(bad),
                      "seed": a.seed, "jobs": a.jobs, "seconds": a.seconds}))
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
