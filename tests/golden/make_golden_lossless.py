#!/usr/bin/env python3
"""Generates the lossless (VP8L) fixtures tests/golden/lossless_*.webp and manifest_lossless.json from the UNMODIFIED reference
(oracle/_ref/libwebp_ref.so), next to make_golden.py's lossy ones and with the same manifest layout: sha256 of the reference
WebPDecode output per (colourspace, flags) combination; SIMD on and off must agree."""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import refwebp as R  # noqa: E402

COMBOS = [(R.MODE_RGBA, 0), (R.MODE_rgbA, 0), (R.MODE_BGR, 0), (R.MODE_Argb, 0), (R.MODE_RGB_565, 0), (R.MODE_rgbA_4444, 0),
          (R.MODE_YUV, 0), (R.MODE_YUVA, 0)]


def pictures():
    rng = np.random.default_rng(2024)
    photo = np.zeros((75, 100, 4), np.uint8)            # predictor + cross-colour + subtract-green, colour cache
    photo[..., :3] = R.synth(100, 75, 31)
    photo[..., 3] = 255
    yield "lossless_photo_100x75", photo, dict(quality=75, method=4)
    alpha = np.zeros((61, 83, 4), np.uint8)             # translucent
    alpha[..., :3] = R.synth(83, 61, 32)
    alpha[..., 3] = np.clip(rng.integers(0, 400, (61, 83)), 0, 255)
    yield "lossless_alpha_83x61", alpha, dict(quality=100, method=6)
    pal = rng.integers(0, 256, (5, 4), dtype=np.uint8)  # 5 colours: palette with 4-bit bundling
    y, x = np.mgrid[0:50, 0:121]
    yield "lossless_palette5_121x50", pal[((x // 4 + y // 3) + rng.integers(0, 2, (50, 121))) % 5], dict(quality=50, method=3)
    yield "lossless_1x1", np.array([[[10, 200, 30, 128]]], np.uint8), dict(quality=75, method=4)


def main():
    manifest = []
    for name, pix, kw in pictures():
        data = R.encode(np.ascontiguousarray(pix), R.EncCfg(lossless=1, **kw))
        with open(os.path.join(HERE, name + ".webp"), "wb") as f:
            f.write(data)
        st, feat = R.features(data)
        outs = {}
        for csp, fl in COMBOS:
            s1, a = R.decode(data, csp, fl, simd=True)
            s2, b = R.decode(data, csp, fl, simd=False)
            assert s1 == 0 and s2 == 0 and (a == b).all(), (name, csp, fl)
            outs[f"{csp}:{fl}"] = hashlib.sha256(a.tobytes()).hexdigest()
        manifest.append(dict(file=name + ".webp", bytes=len(data), features=feat, meta=dict(enc=kw), sha256=outs))
        print(name, len(data), feat)
    with open(os.path.join(HERE, "manifest_lossless.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
