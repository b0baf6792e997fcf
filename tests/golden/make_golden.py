#!/usr/bin/env python3
"""Generates tests/golden/*.webp and manifest.json from the UNMODIFIED reference (oracle/_ref/libwebp_ref.so).

Run where /root/reference exists (`make -C oracle ref` first). The fixtures are small on purpose; they travel
with the repo so the GPU-box tests have reference answers even if oracle/_ref were absent there.
Each manifest entry: file, encoder config, and sha256 of the reference WebPDecode output for every
(colourspace, flags) combination the parity tests use (flags: 1 bypass_filtering, 2 no_fancy_upsampling).
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import refwebp as R  # noqa: E402

CASES = [  # name, w, h, seed, EncCfg kwargs
    ("simple_1part_320x200", 320, 200, 3, dict(quality=75, method=4, segments=1, filter_type=0, partitions=0)),
    ("normal_8part_400x300", 400, 300, 4, dict(quality=75, method=4, segments=4, filter_type=1, partitions=3, low_memory=1)),
    ("default_q80_256x256", 256, 256, 5, dict(quality=80, method=4)),
    ("odd_255x127_q50", 255, 127, 6, dict(quality=50, method=4)),
    ("tiny_17x16_4part", 17, 16, 7, dict(quality=30, method=4, partitions=2, low_memory=1)),
    ("one_pixel", 1, 1, 8, dict(quality=80, method=4)),
    ("sharp_130x97_q95", 130, 97, 9, dict(quality=95, method=4, filter_sharpness=5)),
    ("lowq_2part_200x333", 200, 333, 10, dict(quality=5, method=2, partitions=1, segments=2)),
    ("strongfilter_96x64", 96, 64, 11, dict(quality=40, method=4, filter_strength=100, filter_type=1)),
    ("nofilter_64x48", 64, 48, 12, dict(quality=70, method=4, filter_strength=0)),
]
# Images with an ALPH chunk (config 5 shape, small): name, w, h, seed, EncCfg kwargs. The ALPH header byte they end
# up with is recorded in the manifest (method | filter << 2 | pre-processing << 4).
ALPHA_CASES = [
    ("alpha_gradient_130x97", 130, 97, 21, dict(quality=75, method=4, alpha_filtering=2)),        # gradient / palette
    ("alpha_predictor_320x200", 320, 200, 4, dict(quality=75, method=4, alpha_filtering=0)),      # VP8L predictor transform
    ("alpha_fast_255x127_q60", 255, 127, 22, dict(quality=75, method=4, alpha_filtering=1, alpha_quality=60)),
    ("alpha_tiny_17x16", 17, 16, 5, dict(quality=75, method=4)),
    ("alpha_raw_1x1", 1, 1, 7, dict(quality=75, method=4)),                                        # method 0 (raw)
    ("alpha_lowq_200x150", 200, 150, 23, dict(quality=40, method=2, alpha_quality=20, alpha_filtering=2)),
]
ALPHA_COMBOS = [(R.MODE_RGBA, 0), (R.MODE_rgbA, 0), (R.MODE_ARGB, 0), (R.MODE_Argb, 0), (R.MODE_BGRA, 0), (R.MODE_bgrA, 0),
                (R.MODE_RGB, 0), (R.MODE_RGBA, 3), (R.MODE_YUV, 0)]
COMBOS = [(R.MODE_RGBA, 0), (R.MODE_RGBA, 1), (R.MODE_RGBA, 2), (R.MODE_RGB, 0), (R.MODE_BGRA, 0), (R.MODE_ARGB, 0),
          (R.MODE_BGR, 3), (R.MODE_rgbA, 0), (R.MODE_YUV, 0), (R.MODE_YUV, 1)]


def sha(a):
    return hashlib.sha256(a.tobytes()).hexdigest()


def main():
    manifest = []
    ref_test = "/root/reference/examples/test.webp"
    files = []
    if os.path.exists(ref_test):   # config 1 of BASELINE.json (4880 bytes, the reference's only lossy fixture)
        shutil.copy(ref_test, os.path.join(HERE, "ref_examples_test.webp"))
        files.append(("ref_examples_test", dict(source="examples/test.webp")))
    for name, w, h, seed, kw in CASES:
        data = R.encode(R.synth(w, h, seed), R.EncCfg(**kw))
        with open(os.path.join(HERE, name + ".webp"), "wb") as f:
            f.write(data)
        files.append((name, dict(width=w, height=h, seed=seed, enc=kw)))
    for name, meta in files:
        data = open(os.path.join(HERE, name + ".webp"), "rb").read()
        st, feat = R.features(data)
        outs = {}
        for csp, fl in COMBOS:
            s1, a = R.decode(data, csp, fl, simd=True)
            s2, b = R.decode(data, csp, fl, simd=False)
            assert s1 == 0 and s2 == 0 and (a == b).all(), (name, csp, fl)   # SIMD on == SIMD off
            outs[f"{csp}:{fl}"] = sha(a)
        manifest.append(dict(file=name + ".webp", bytes=len(data), features=feat, meta=meta, sha256=outs))
        print(name, len(data), feat)
    with open(os.path.join(HERE, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    amanifest = []
    for name, w, h, seed, kw in ALPHA_CASES:
        data = R.encode(R.synth(w, h, seed, alpha=True), R.EncCfg(**kw))
        with open(os.path.join(HERE, name + ".webp"), "wb") as f:
            f.write(data)
        st, feat = R.features(data)
        i = data.find(b"ALPH")
        outs = {}
        for csp, fl in ALPHA_COMBOS:
            s1, a = R.decode(data, csp, fl, simd=True)
            s2, b = R.decode(data, csp, fl, simd=False)
            assert s1 == 0 and s2 == 0 and (a == b).all(), (name, csp, fl)
            outs[f"{csp}:{fl}"] = sha(a)
        amanifest.append(dict(file=name + ".webp", bytes=len(data), features=feat, alph_header=data[i + 8],
                              meta=dict(width=w, height=h, seed=seed, enc=kw), sha256=outs))
        print(name, len(data), feat, hex(data[i + 8]))
    with open(os.path.join(HERE, "manifest_alpha.json"), "w") as f:
        json.dump(amanifest, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
