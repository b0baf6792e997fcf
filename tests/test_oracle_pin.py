"""Pins the oracle: (1) the compiled reference reproduces the hashes recorded in BASELINE.md for the reference's
own fixture, (2) the plain-C restatement equals the compiled reference byte for byte on every golden fixture,
every colourspace/flag combination and on fresh seeded corpora, (3) both equal the committed manifest."""
import hashlib

import numpy as np
import pytest

from conftest import sha

COMBOS = [(1, 0), (1, 1), (1, 2), (0, 0), (3, 0), (4, 0), (2, 3), (7, 0), (11, 0), (11, 1)]

# sha256 of `dwebp examples/test.webp <fmt>` from the reference build
PINNED_PPM = "db448ba15096dd0941cacb7e7cc8f0bf5461226c423c330fa73d0591bd8ac980"
PINNED_YUV = "c6f5e29437bb5a96d2a64f71250b36f6c387c9d9246b7961c552f1951a1aa7d6"
PINNED_NOFANCY_PPM = "e7f43a0c18130d4a669e3e069324596c9c4e15cd445e5b76c3f94157da1d4c12"


def _ppm(rgb):
    h, s = rgb.shape
    return hashlib.sha256(b"P6\n%d %d\n255\n" % (s // 3, h) + rgb.tobytes()).hexdigest()


def test_reference_build_matches_pinned_hashes(ref, manifest):
    data = next(e for e in manifest if e["file"] == "ref_examples_test.webp")["data"]
    for simd in (True, False):
        st, rgb = ref.decode(data, ref.MODE_RGB, 0, simd=simd)
        assert st == 0 and _ppm(rgb) == PINNED_PPM
        st, yuv = ref.decode(data, ref.MODE_YUV, 0, simd=simd)
        assert st == 0 and sha(yuv) == PINNED_YUV
        st, rgb = ref.decode(data, ref.MODE_RGB, ref.FLAG_NO_FANCY, simd=simd)
        assert st == 0 and _ppm(rgb) == PINNED_NOFANCY_PPM


def test_port_matches_pinned_hashes(port, manifest):
    data = next(e for e in manifest if e["file"] == "ref_examples_test.webp")["data"]
    st, rgb = port.decode(data, port.RGB, 0)
    assert st == 0 and _ppm(rgb) == PINNED_PPM
    st, yuv = port.decode(data, port.YUV, 0)
    assert st == 0 and sha(yuv) == PINNED_YUV


def test_port_matches_manifest(port, manifest):
    for e in manifest:
        for key, want in e["sha256"].items():
            csp, fl = map(int, key.split(":"))
            st, out = port.decode(e["data"], csp, fl)
            assert st == 0, (e["file"], key)
            assert sha(out) == want, (e["file"], key)
        st, f = port.features(e["data"])
        assert st == 0 and f == e["features"]


def test_reference_matches_manifest(ref, manifest):
    for e in manifest:
        for key, want in e["sha256"].items():
            csp, fl = map(int, key.split(":"))
            st, out = ref.decode(e["data"], csp, fl)
            assert st == 0 and sha(out) == want, (e["file"], key)


@pytest.mark.parametrize("w,h,kind,seed", [(1920, 1080, "simple", 21), (1920, 1080, "8part", 22), (640, 361, "default", 23),
                                           (48, 1000, "8part", 24), (1000, 48, "default", 25)])
def test_port_equals_reference_on_fresh_corpora(ref, port, w, h, kind, seed):
    cfg = {"simple": ref.cfg_simple_1part(), "8part": ref.cfg_normal_8part(), "default": ref.cfg_default()}[kind]
    data = ref.encode(ref.synth(w, h, seed), cfg)
    for csp, fl in COMBOS:
        s1, a = ref.decode(data, csp, fl)
        s2, b = port.decode(data, csp, fl)
        assert s1 == s2 == 0
        assert np.array_equal(a, b), (kind, csp, fl)


def test_port_equals_reference_on_damaged_files(ref, port, manifest):
    """Status codes on truncated / corrupted input (vp8_dec.c:651-659, webp_dec.c:761-767)."""
    data = next(e for e in manifest if e["file"] == "normal_8part_400x300.webp")["data"]
    cases = [data[:n] for n in (0, 5, 11, 12, 19, 20, 29, 30, 40, 200, 3000, len(data) // 2, len(data) - 1)]
    rng = np.random.default_rng(5)
    for _ in range(12):
        b = bytearray(data)
        for _ in range(3):
            b[int(rng.integers(30, len(b)))] ^= int(rng.integers(1, 256))
        cases.append(bytes(b))
    cases.append(b"RIFF" + data[4:8] + b"WEBX" + data[12:])
    cases.append(data[12:])          # bare "VP8 " chunk without RIFF
    cases.append(data[20:])          # bare VP8 frame
    for c in cases:
        s1, a = ref.decode(c, ref.MODE_RGBA, 0)
        s2, b = port.decode(c, port.RGBA, 0)
        assert s1 == s2, (len(c), s1, s2)
        if s1 == 0:
            assert np.array_equal(a, b)
        f1, f2 = ref.features(c), port.features(c)
        assert f1[0] == f2[0] and (f1[0] != 0 or f1[1] == f2[1]), (len(c), f1, f2)


def test_reference_matches_lossless_manifest(ref, lmanifest):
    for e in lmanifest:
        st, f = ref.features(e["data"])
        assert st == 0 and f == e["features"]
        for key, want in e["sha256"].items():
            csp, fl = map(int, key.split(":"))
            st, out = ref.decode(e["data"], csp, fl)
            assert st == 0 and sha(out) == want, (e["file"], key)
