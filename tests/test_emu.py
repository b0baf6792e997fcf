"""Device logic on the host (tests/emu): the product's parse / reconstruct / filter / emit code compiled with
-DVP8_EMU (a warp phase = a loop over 32 lanes, a wavefront step = a loop over its macroblocks, visited in
both orders) must reproduce the reference byte for byte. This is a unit test of the kernels' arithmetic and
dependency analysis where no GPU exists; the GPU parity tests proper are tests/test_gpu_parity.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, sha, smooth_image

EMU_DIR = os.path.join(ROOT, "tests", "emu")


@pytest.fixture(scope="module")
def emu():
    subprocess.check_call(["make", "-s", "-C", EMU_DIR])
    L = C.CDLL(os.path.join(EMU_DIR, "libvp8_emu.so"))
    L.emu_decode.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p]

    def run(data, w, h, csp, flags=0, reverse=0, unfiltered=None):
        if csp == 11:
            out = np.zeros(w * h + 2 * ((w + 1) // 2) * ((h + 1) // 2), np.uint8)
            stride = w
        else:
            bpp = 3 if csp in (0, 2) else 4
            out = np.zeros((h, w * bpp), np.uint8)
            stride = w * bpp
        st = L.emu_decode(data, len(data), csp, flags, out.ctypes.data, out.size, stride, reverse,
                          unfiltered.ctypes.data if unfiltered is not None else None)
        return st, out
    return run


def test_emu_matches_manifest(emu, manifest):
    for e in manifest:
        w, h = e["features"]["width"], e["features"]["height"]
        for key, want in e["sha256"].items():
            csp, fl = map(int, key.split(":"))
            for rev in (0, 1, 2, 3, 4, 8, 9, 24, 25, 56, 64, 65, 80, 192, 208, 320, 336):   # +256 = the lockstep mode parser; +128 = the banded probability rows; 64 = the fp parser (fp32 boolean decoder, token stream; 80 = with a branch per decode); 2 = row-at-a-time token parser, 4 = lazy ring producer (lanes find the ring dry), 8 = lockstep lanes, 24 = lockstep lanes with grouped event points, 56 = the same as straight-line groups
                st, out = emu(e["data"], w, h, csp, fl, rev)
                assert st == 0 and sha(out) == want, (e["file"], key, rev)


def test_emu_stage_dumps_match_port(emu, port, manifest):
    """Unfiltered reconstruction (K3's output) against the oracle's stage dump."""
    for e in manifest:
        st, d = port.dump(e["data"])
        assert st == 0
        unf = np.zeros(d["mb_w"] * d["mb_h"] * 384, np.uint8)
        st, _ = emu(e["data"], d["width"], d["height"], 11, 0, 0, unf)
        assert st == 0 and np.array_equal(unf, d["unfiltered"]), e["file"]


def test_emu_status_on_damaged_files(emu, port, manifest):
    data = next(e for e in manifest if e["file"] == "normal_8part_400x300.webp")["data"]
    cases = [data[:n] for n in (200, 3000, len(data) // 2, len(data) - 1)]
    rng = np.random.default_rng(7)
    for _ in range(20):
        b = bytearray(data)
        for _ in range(3):
            b[int(rng.integers(30, len(b)))] ^= int(rng.integers(1, 256))
        cases.append(bytes(b))
    data1 = next(e for e in manifest if e["file"] == "simple_1part_320x200.webp")["data"]
    cases += [data1[:n] for n in (100, 2000, len(data1) - 1)]
    for c in cases:
        s_ref, a = port.decode(c, port.RGBA, 0)
        w, h = port.features(c)[1]["width"], port.features(c)[1]["height"]
        for variant in (0, 2, 8, 24, 56, 64, 80, 192, 208, 320, 336):   # lane state machine / row-at-a-time parser / lockstep lanes (block ends on the spot, grouped) / fp parser / + 256: the lockstep mode parser
            s_emu, b = emu(c, max(w, 1), max(h, 1), 1, 0, variant)
            assert s_emu == s_ref, (len(c), s_ref, s_emu, variant)
            if s_ref == 0:
                assert np.array_equal(a, b)


@pytest.mark.parametrize("kind", ["simple", "8part"])
def test_emu_full_hd(emu, ref, kind):
    cfg = ref.cfg_simple_1part() if kind == "simple" else ref.cfg_normal_8part()
    data = ref.encode(ref.synth(1920, 1080, 31), cfg)
    for csp in (1, 11):
        st, want = ref.decode(data, csp, 0)
        st2, got = emu(data, 1920, 1080, csp, 0, 8 if csp == 1 else 24)
        assert st == st2 == 0 and np.array_equal(want.reshape(-1), got.reshape(-1))
        st2, got = emu(data, 1920, 1080, csp, 0, 320)   # the default parsers of the product (lockstep mode parse, fp token parse)
        assert st == st2 == 0 and np.array_equal(want.reshape(-1), got.reshape(-1))


def test_emu_alpha_matches_manifest(emu, amanifest):
    """ALPH chunks (raw, palette + every row filter, VP8L predictor transform) through the device code's host build:
    alpha byte, premultiplied modes, alpha-first modes, and alpha ignored where the output has no alpha channel."""
    for e in amanifest:
        w, h = e["features"]["width"], e["features"]["height"]
        for key, want in e["sha256"].items():
            csp, fl = map(int, key.split(":"))
            st, out = emu(e["data"], w, h, csp, fl, 0)
            assert st == 0 and sha(out) == want, (e["file"], key)


def test_emu_alpha_status_on_damaged_chunks(emu, ref, amanifest):
    """A damaged ALPH payload must end exactly like the reference: same status (header failures surface as
    VP8_STATUS_OUT_OF_MEMORY there, pixel-loop failures as VP8_STATUS_BITSTREAM_ERROR), same pixels when it decodes."""
    rng = np.random.default_rng(3)
    for e in amanifest:
        data = e["data"]
        i = data.find(b"ALPH")
        size = int.from_bytes(data[i + 4:i + 8], "little")
        w, h = e["features"]["width"], e["features"]["height"]
        for _ in range(8):
            b = bytearray(data)
            b[i + 8 + int(rng.integers(0, size))] ^= int(rng.integers(1, 256))
            b = bytes(b)
            s_ref, want = ref.decode(b, ref.MODE_RGBA, 0)
            s_emu, got = emu(b, w, h, 1, 0, 0)
            assert s_emu == s_ref, (e["file"], s_ref, s_emu)
            if s_ref == 0:
                assert np.array_equal(want.reshape(-1), got.reshape(-1)), e["file"]


def test_emu_alpha_and_vp8_both_damaged(emu, ref, amanifest):
    """Both chunks damaged: the status is the one of the failure the reference's row loop meets first (alpha rows are decoded as
    the macroblock rows above them finish). The fp parser (variants 64, 80) records the failing rows for that."""
    rng = np.random.default_rng(29)
    seen = set()
    for e in amanifest:
        data = e["data"]
        if len(data) < 400:
            continue
        i = data.find(b"ALPH")
        asz = int.from_bytes(data[i + 4:i + 8], "little")
        v = data.find(b"VP8 ", i + 8 + asz)
        vsz = int.from_bytes(data[v + 4:v + 8], "little")
        w, h = e["features"]["width"], e["features"]["height"]
        for k in range(60):
            b = bytearray(data)
            b[i + 8 + int(rng.integers(0, asz))] ^= int(rng.integers(1, 256))
            if k & 1:
                b[v + 8 + 10 + int(rng.integers(0, vsz - 10))] ^= int(rng.integers(1, 256))
            else:
                keep = int(rng.integers(vsz // 8, vsz))
                b[v + 8 + keep:v + 8 + vsz] = bytes(vsz - keep)
            b = bytes(b)
            s_ref, want = ref.decode(b, ref.MODE_RGBA, 0)
            seen.add(s_ref)
            for variant in (64, 80, 192, 320):
                s_emu, got = emu(b, w, h, 1, 0, variant)
                assert s_emu == s_ref, (e["file"], k, s_ref, s_emu, variant)
                if s_ref == 0 and not np.array_equal(want.reshape(-1), got.reshape(-1)):
                    _, want = ref.decode(b, ref.MODE_RGBA, 0, simd=False)   # SSE2 vs C transforms on out-of-range coefficients
                    assert np.array_equal(want.reshape(-1), got.reshape(-1)), e["file"]
    assert {3, 7} <= seen, seen



def _ff_mutants(rng, data, count):
    """Copies of a lossy file in which the first partition or the (single) token partition starts with byte 0xFF -- what no
    encoder writes and what makes the reference's reader leave its range (vp8_literal.h) -- alone or with more damage."""
    i = data.find(b"VP8 ")
    fo = i + 8
    part0 = int.from_bytes(data[fo:fo + 3], "little") >> 5
    out = []
    for k in range(count):
        b = bytearray(data)
        at = fo + 10 if k % 2 == 0 else fo + 10 + part0
        b[min(at, len(b) - 1)] = 0xFF
        if k >= 2:
            for _ in range(int(rng.integers(0, 3))):
                b[int(rng.integers(fo + 10, len(b)))] ^= int(rng.integers(1, 256))
        out.append(bytes(b))
    return out


def test_emu_partition_starting_with_ff(emu, ref, port, manifest, amanifest):
    """The literal reader (vp8_parse_core.h:RefBits, variants 64 and 80 = the default parser): status and pixels of the reference
    on files whose partitions start with 0xFF; the oracle's restatement reads the same way."""
    rng = np.random.default_rng(37)
    seen = {}
    for e in list(manifest) + list(amanifest):
        data = e["data"]
        if data.find(b"VP8 ") < 0 or len(data) < 200:
            continue
        w, h = e["features"]["width"], e["features"]["height"]
        for b in _ff_mutants(rng, data, 24):
            if ref.features(b)[0] != 0:
                continue
            s_ref, want = ref.decode(b, ref.MODE_RGBA, 0)
            seen[s_ref] = seen.get(s_ref, 0) + 1
            for variant in (64, 80, 192, 320):
                s_emu, got = emu(b, w, h, 1, 0, variant)
                assert s_emu == s_ref, (e["file"], s_ref, s_emu, variant)
                if s_ref == 0 and not np.array_equal(want.reshape(-1), got.reshape(-1)):
                    _, want = ref.decode(b, ref.MODE_RGBA, 0, simd=False)   # SSE2 vs C transforms on out-of-range coefficients
                    assert np.array_equal(want.reshape(-1), got.reshape(-1)), e["file"]
            if not e["features"].get("has_alpha"):
                s_port, got = port.decode(b, port.RGBA, 0)
                assert s_port == s_ref, (e["file"], s_ref, s_port)
                if s_ref == 0:
                    assert np.array_equal(want.reshape(-1), got.reshape(-1)), e["file"]
    assert seen.get(0, 0) > 20 and seen.get(7, 0) > 20, seen


def test_emu_crop_and_flip_match_reference(ref, manifest, amanifest):
    """options.use_cropping / options.flip: the window is upsampled as if it were the picture, rows below it are never
    decoded (so data missing down there goes unnoticed), the 8-bit alpha path restarts its horizontal unfilter at the
    window -- all as the reference does (frame_dec.c:430-490,571-596, vp8l_dec.c:887-912)."""
    subprocess.check_call(["make", "-s", "-C", EMU_DIR])
    L = C.CDLL(os.path.join(EMU_DIR, "libvp8_emu.so"))
    L.emu_decode_window.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int,
                                    C.c_int, C.c_int]

    def emu_window(data, csp, dev_flags, crop, W, H):
        w, h = (crop[2], crop[3]) if crop else (W, H)
        n = (w * h + 2 * ((w + 1) // 2) * ((h + 1) // 2) + (w * h if csp == 12 else 0)) if csp in (11, 12) else w * h * ref.BPP[csp]
        out = np.zeros(max(n, 16), np.uint8)
        c = crop or (0, 0, 0, 0)
        st = L.emu_decode_window(data, len(data), csp, dev_flags, out.ctypes.data, out.size, w if csp in (11, 12) else w * ref.BPP[csp],
                                 c[0], c[1], c[2], c[3])
        return st, out[:n]

    rng = np.random.default_rng(9)
    for e in list(manifest) + list(amanifest):
        W, H = e["features"]["width"], e["features"]["height"]
        for it in range(4):
            crop = None
            if it > 0:
                cw, ch = int(rng.integers(1, W + 1)), int(rng.integers(1, H + 1))
                crop = (int(rng.integers(0, W - cw + 1)), int(rng.integers(0, H - ch + 1)), cw, ch)
            flip, nofancy = int(rng.integers(0, 2)), int(rng.integers(0, 2)) if it % 2 else 0
            for csp in (1, 7, 0, 11, 5, 6, 10, 12):   # incl. the 16-bit colourspaces and MODE_YUVA
                s_ref, want = ref.decode_window(e["data"], csp, (8 if flip else 0) | (2 if nofancy else 0), crop)
                s_emu, got = emu_window(e["data"], csp, (4 if flip else 0) | (2 if nofancy else 0), crop, W, H)
                assert s_emu == s_ref, (e["file"], crop, flip, csp, s_ref, s_emu)
                if s_ref == 0:
                    assert np.array_equal(want, got), (e["file"], crop, flip, nofancy, csp)
    # a file cut short: a window near the top still decodes, one that reaches the missing rows does not
    data = next(e for e in manifest if e["file"] == "simple_1part_320x200.webp")["data"]
    cut = data[: len(data) * 2 // 3]
    for crop in ((0, 0, 320, 32), (16, 16, 100, 40), (0, 120, 320, 80), (0, 0, 320, 200)):
        s_ref, want = ref.decode_window(cut, 1, 0, crop)
        s_emu, got = emu_window(cut, 1, 0, crop, 320, 200)
        assert s_emu == s_ref, (crop, s_ref, s_emu)
        if s_ref == 0:
            assert np.array_equal(want, got)
    # windows that do not fit are refused the same way
    for crop in ((300, 0, 40, 10), (0, 199, 10, 3), (310, 190, 12, 12)):
        s_ref, _ = ref.decode_window(data, 1, 0, crop)
        s_emu, _ = emu_window(data, 1, 0, crop, 320, 200)
        assert s_ref == 2   # the product refuses these on the host (plan_item), the emulation harness never sees them
        del s_emu


def test_emu_scaling_matches_reference(ref, manifest, amanifest):
    """options.use_scaling (io_dec.c:239-556, src/dsp/rescaler.c): up and down, one axis each way, ratio-preserving
    requests, crop + scale, planar and packed colourspaces, flip; the loop filter is dropped for large downscaling ratios
    exactly like WebPIoInitFromOptions does. The device code's host build against the reference, byte for byte."""
    subprocess.check_call(["make", "-s", "-C", EMU_DIR])
    L = C.CDLL(os.path.join(EMU_DIR, "libvp8_emu.so"))
    L.emu_decode_scaled.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int,
                                    C.c_int, C.c_int, C.c_int, C.c_int]
    rng = np.random.default_rng(21)
    for e in list(manifest) + list(amanifest):
        has_alpha = bool(e["features"].get("has_alpha"))
        W, H = e["features"]["width"], e["features"]["height"]
        for it in range(6 + (2 if W * H <= 4096 else 0)):   # two requests beyond 16383 on the small pictures
            crop = None
            if it >= 4 and it < 6:
                cw, ch = int(rng.integers(1, W + 1)), int(rng.integers(1, H + 1))
                crop = (int(rng.integers(0, W - cw + 1)), int(rng.integers(0, H - ch + 1)), cw, ch)
            w, h = (crop[2], crop[3]) if crop else (W, H)
            req = [(max(1, w // 2), max(1, h // 3)), (w * 2 + 1, h + 7), (max(1, w - 1), h * 3), (0, max(1, h // 2)), (w + 5, 0),
                   (int(rng.integers(1, 2 * w + 2)), int(rng.integers(1, 2 * h + 2))), (20001, max(1, h // 2)), (3, 16500)][it]
            flip = int(rng.integers(0, 2))
            # files with an ALPH chunk: the alpha plane is rescaled too, premultiplied modes multiply after scaling, MODE_YUVA
            # multiplies the luma before and divides it after (io_dec.c:252-300, 414-470)
            for csp in ((1, 7, 12, 10, 5, 9, 0) if has_alpha else (1, 0, 11, 6, 12, 4)):
                s_ref, (sw, sh), want = ref.decode_scaled(e["data"], csp, 8 if flip else 0, crop, req)
                assert s_ref == 0, (e["file"], req, crop, s_ref)
                n = (sw * sh + 2 * ((sw + 1) // 2) * ((sh + 1) // 2) + (sw * sh if csp == 12 else 0)) if csp in (11, 12) else sw * sh * ref.BPP[csp]
                out = np.zeros(max(n, 16), np.uint8)
                c = crop or (0, 0, 0, 0)
                st = L.emu_decode_scaled(e["data"], len(e["data"]), csp, 4 if flip else 0, out.ctypes.data, out.size,
                                         sw if csp in (11, 12) else sw * ref.BPP[csp], c[0], c[1], c[2], c[3], sw, sh)
                assert st == 0, (e["file"], req, crop, csp, st)
                assert np.array_equal(out[:n], want), (e["file"], req, crop, flip, csp, (sw, sh))


def test_emu_dithering_matches_reference(ref):
    """options.dithering_strength (frame_dec.c:319-386): one pseudo-random sequence per picture, consumed in raster order by
    the macroblocks without AC chroma coefficients inside the crop window's macroblock range, added after a row has been
    filtered and before the next row filters across the common edge. Smooth pictures at fine quantisers, both loop filters,
    1-4 segments, strengths 50 (dwebp's default) and 100, crop windows; the device code's host build against the reference."""
    subprocess.check_call(["make", "-s", "-C", EMU_DIR])
    L = C.CDLL(os.path.join(EMU_DIR, "libvp8_emu.so"))
    L.emu_decode_dithered.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int,
                                      C.c_int, C.c_int, C.c_int]
    rng = np.random.default_rng(31)
    changed = 0
    for k, (q, segs, flt, w, h) in enumerate(((100, 1, 0, 200, 136), (99, 4, 1, 177, 93), (97, 4, 1, 64, 200), (95, 2, 0, 320, 48), (90, 4, 1, 130, 130))):
        data = ref.encode(smooth_image(w, h, 40 + k), ref.EncCfg(q, 4, segments=segs, filter_type=flt, filter_strength=40))
        for strength in (50, 100):
            for it in range(3):
                crop = None
                if it > 0:
                    cw, ch = int(rng.integers(1, w + 1)), int(rng.integers(1, h + 1))
                    crop = (int(rng.integers(0, w - cw + 1)), int(rng.integers(0, h - ch + 1)), cw, ch)
                ow, oh = (crop[2], crop[3]) if crop else (w, h)
                for csp in (1, 11):
                    s_ref, want = ref.decode_dithered(data, csp, 0, crop, strength)
                    _, plain = ref.decode_dithered(data, csp, 0, crop, 0)
                    changed += int((want != plain).sum())
                    n = want.size
                    out = np.zeros(max(n, 16), np.uint8)
                    c = crop or (0, 0, 0, 0)
                    st = L.emu_decode_dithered(data, len(data), csp, 0, out.ctypes.data, out.size, ow if csp == 11 else ow * 4,
                                               strength, c[0], c[1], c[2], c[3])
                    assert st == s_ref == 0 and np.array_equal(out[:n], want), (q, segs, flt, strength, crop, csp)
    assert changed > 10000   # the cases do exercise the dithering


def test_emu_alpha_dithering_matches_reference(ref, amanifest):
    """options.alpha_dithering_strength (dwebp's default is 100) on ALPH planes whose levels the encoder quantised
    (alpha_q < 100: ALPH header pre-processing = 1, alpha_dec.c:200-230): WebPDequantizeLevels on the crop window of the
    finished plane (quant_levels_dec_utils.c:262-291). Smooth alpha ramps so that the box average does move pixels; several
    strengths (radius 0..4), crop windows smaller than the filter, premultiplied and plain output, MODE_YUVA."""
    subprocess.check_call(["make", "-s", "-C", EMU_DIR])
    L = C.CDLL(os.path.join(EMU_DIR, "libvp8_emu.so"))
    L.emu_decode_window.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int,
                                    C.c_int, C.c_int]
    L.emu_set_alpha_dithering.argtypes = [C.c_int]
    rng = np.random.default_rng(77)
    cases = [e["data"] for e in amanifest]
    for k, (w, h, aq) in enumerate(((160, 120, 30), (97, 203, 60), (64, 64, 10), (300, 40, 80), (33, 9, 50))):
        y, x = np.mgrid[0:h, 0:w].astype(np.float64)
        pix = np.zeros((h, w, 4), np.uint8)
        pix[..., :3] = ref.synth(w, h, 500 + k)
        pix[..., 3] = np.clip(128 + 120 * np.sin(x * 0.03 + k) * np.cos(y * 0.025), 0, 255)
        cases.append(ref.encode(pix, ref.EncCfg(80, 4, alpha_quality=aq, alpha_filtering=k % 3)))
    changed = 0
    try:
        for data in cases:
            st0, f = ref.features(data)
            W, H = f["width"], f["height"]
            for strength in (100, 50, 30, 20, 150):
                L.emu_set_alpha_dithering(strength)
                for it in range(3):
                    crop = None
                    if it > 0:
                        cw, ch = int(rng.integers(1, W + 1)), int(rng.integers(1, H + 1))
                        if it == 2:
                            cw, ch = min(cw, 7), min(ch, 5)
                        crop = (int(rng.integers(0, W - cw + 1)), int(rng.integers(0, H - ch + 1)), cw, ch)
                    ow, oh = (crop[2], crop[3]) if crop else (W, H)
                    for csp in (1, 7, 12):
                        s_ref, want = ref.decode_dithered(data, csp, 0, crop, 0, strength)
                        _, plain = ref.decode_dithered(data, csp, 0, crop, 0, 0)
                        changed += int((want != plain).sum())
                        n = want.size
                        out = np.zeros(max(n, 16), np.uint8)
                        c = crop or (0, 0, 0, 0)
                        st = L.emu_decode_window(data, len(data), csp, 0, out.ctypes.data, out.size, ow if csp == 12 else ow * 4,
                                                 c[0], c[1], c[2], c[3])
                        assert st == s_ref == 0 and np.array_equal(out[:n], want), (len(data), strength, crop, csp)
    finally:
        L.emu_set_alpha_dithering(0)
    assert changed > 10000   # the cases do exercise the de-banding


def lossless_cases(ref):
    """Lossless (VP8L) pictures that between them use every transform, the colour cache, meta-Huffman groups and all pixel
    bundlings: photo-like synthetic pictures (predictor + cross-colour + subtract-green), palettes of 2 / 4 / 16 / 40 / 256
    colours, translucent alpha, odd sizes down to 1x1, encoder methods 0..6."""
    rng = np.random.default_rng(123)
    cases = []
    for k, (w, h, q, m) in enumerate(((64, 48, 75, 4), (131, 77, 100, 6), (200, 150, 20, 0), (33, 91, 50, 2), (1, 1, 75, 4), (7, 3, 90, 3))):
        pix = np.zeros((h, w, 4), np.uint8)
        pix[..., :3] = ref.synth(w, h, 900 + k)
        pix[..., 3] = 255 if k % 2 == 0 else np.clip(rng.integers(0, 400, (h, w)), 0, 255)
        cases.append(ref.encode(pix, ref.EncCfg(q, m, lossless=1)))
    for k, ncol in enumerate((2, 4, 16, 40, 256)):
        w, h = 97 + 10 * k, 61 + 7 * k
        pal = rng.integers(0, 256, (ncol, 4), dtype=np.uint8)
        if k % 2 == 0:
            pal[:, 3] = 255
        y, x = np.mgrid[0:h, 0:w]
        idx = ((x // 5 + y // 3) + rng.integers(0, 2, (h, w))) % ncol
        cases.append(ref.encode(pal[idx], ref.EncCfg(75, 4 + k % 3, lossless=1)))
    big = np.zeros((300, 400, 4), np.uint8)
    big[..., :3] = ref.synth(400, 300, 77)
    big[..., 3] = 255
    big[100:200, 50:350, :3] = 40          # flat area: LZ77 runs + colour cache
    cases.append(ref.encode(big, ref.EncCfg(100, 6, lossless=1)))
    return cases


def test_emu_lossless_matches_reference(ref):
    """Whole-picture VP8L through the device code's host build: every RGB-family colourspace (premultiplied and 16-bit ones
    included) and MODE_YUV / MODE_YUVA, crop windows at odd offsets (not snapped for lossless), flip; damaged and truncated
    files end with the reference's status."""
    subprocess.check_call(["make", "-s", "-C", EMU_DIR])
    L = C.CDLL(os.path.join(EMU_DIR, "libvp8_emu.so"))
    L.emu_decode_window.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int,
                                    C.c_int, C.c_int]

    def emu_window(data, csp, dev_flags, crop, W, H):
        w, h = (crop[2], crop[3]) if crop else (W, H)
        bpp = 1 if csp in (11, 12) else ref.BPP[csp]
        n = (w * h + 2 * ((w + 1) // 2) * ((h + 1) // 2) + (w * h if csp == 12 else 0)) if csp in (11, 12) else w * h * bpp
        out = np.zeros(max(n, 16), np.uint8)
        c = crop or (0, 0, 0, 0)
        st = L.emu_decode_window(data, len(data), csp, dev_flags, out.ctypes.data, out.size, w * bpp, c[0], c[1], c[2], c[3])
        return st, out[:n]

    rng = np.random.default_rng(5)
    for data in lossless_cases(ref):
        st0, f = ref.features(data)
        assert st0 == 0 and f["format"] == 2
        W, H = f["width"], f["height"]
        for it in range(3):
            crop = None
            if it > 0:
                cw, ch = int(rng.integers(1, W + 1)), int(rng.integers(1, H + 1))
                crop = (int(rng.integers(0, W - cw + 1)), int(rng.integers(0, H - ch + 1)), cw, ch)
            flip = int(rng.integers(0, 2))
            for csp in (1, 7, 0, 2, 3, 8, 4, 9, 5, 6, 10, 11, 12):
                s_ref, want = ref.decode_window(data, csp, 8 if flip else 0, crop)
                s_emu, got = emu_window(data, csp, 4 if flip else 0, crop, W, H)
                assert s_emu == s_ref == 0, (len(data), crop, flip, csp, s_ref, s_emu)
                assert np.array_equal(want, got), (len(data), W, H, crop, flip, csp)
        # damage: flipped bytes and truncation
        for k in range(12):
            b = bytearray(data)
            if k % 3 == 2 and len(b) > 40:
                b = b[: int(rng.integers(30, len(b)))]
            else:
                b[int(rng.integers(20, len(b)))] ^= int(rng.integers(1, 256))
            b = bytes(b)
            s_ref, want = ref.decode(b, 1, 0)
            sf, fb = ref.features(b)     # the damage may have hit the dimensions
            s_emu, got = emu_window(b, 1, 0, None, fb["width"] if sf == 0 else W, fb["height"] if sf == 0 else H)
            assert s_emu == s_ref, (len(data), k, s_ref, s_emu)
            if s_ref == 0:
                assert np.array_equal(want.reshape(-1), got)


def test_emu_crafted_vp8l_corners(ref, amanifest):
    """Hand-made streams (tests/vp8l_craft.py) for what no encoder of the reference writes but its decoder accepts: group numbers
    beyond 1000 / beyond the pixel count (remapped to the ones in use, vp8l_dec.c:399-424), a colour-indexing transform that is
    not the first transform (rows widened in place); whole pictures in every output family and inside ALPH chunks."""
    import vp8l_craft
    subprocess.check_call(["make", "-s", "-C", EMU_DIR])
    L = C.CDLL(os.path.join(EMU_DIR, "libvp8_emu.so"))
    L.emu_decode_window.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int,
                                    C.c_int, C.c_int]
    host = [e["data"] for e in amanifest if e["file"] == "alpha_tiny_17x16.webp"][0]
    for name, data in vp8l_craft.crafted_cases(host):
        st0, f = ref.features(data)
        assert st0 == 0, name
        w, h = f["width"], f["height"]
        for csp in (1, 7, 0, 5, 11, 12):
            s_ref, want = ref.decode_window(data, csp, 0, None)
            assert s_ref == 0, (name, csp, s_ref)      # the reference does decode these
            bpp = 1 if csp in (11, 12) else ref.BPP[csp]
            n = (w * h + 2 * ((w + 1) // 2) * ((h + 1) // 2) + (w * h if csp == 12 else 0)) if csp in (11, 12) else w * h * bpp
            out = np.zeros(max(n, 16), np.uint8)
            st = L.emu_decode_window(data, len(data), csp, 0, out.ctypes.data, out.size, w * bpp, 0, 0, 0, 0)
            assert st == 0, (name, csp, st)
            assert np.array_equal(want, out[:n]), (name, csp)
        for k in range(1, 6):   # cut short: same status
            b = data[: len(data) - 2 * k]
            if ref.features(b)[0] != 0:
                continue
            s_ref, _ = ref.decode(b, 1, 0)
            out = np.zeros(max(w * h * 4, 16), np.uint8)
            st = L.emu_decode_window(b, len(b), 1, 0, out.ctypes.data, out.size, w * 4, 0, 0, 0, 0)
            assert st == s_ref, (name, k, s_ref, st)


def test_emu_lossless_palette_picture_losing_its_last_bits(ref, lmanifest):
    """A whole VP8L picture always runs the reference's 32-bit pixel loop (VP8LDecodeImage -> DecodeImageData, vp8l_dec.c:1761-1765);
    only an ALPH payload may take the 8-bit one. They differ in when running out of data is an error: the 32-bit loop fails as
    soon as the reader is past the end, even with every pixel decoded. Palette pictures (the 8-bit loop's shape) with the tail
    cut off or damaged must therefore end like the reference. (Found by tools/fuzz_emu.py.)"""
    subprocess.check_call(["make", "-s", "-C", EMU_DIR])
    L = C.CDLL(os.path.join(EMU_DIR, "libvp8_emu.so"))
    L.emu_decode_window.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int,
                                    C.c_int, C.c_int]
    rng = np.random.default_rng(17)
    seen = {0: 0, 3: 0}
    golden = next(e["data"] for e in lmanifest if e["file"] == "lossless_palette5_121x50.webp")
    known = []
    for at, val in ((704, 0x97), (1476, 0xfd), (755, 0xd5), (1436, 0xd7), (235, 0xb0)):   # decoded to the last pixel, reader past the end
        b = bytearray(golden)
        b[at] = val
        known.append(bytes(b))
    for data in [golden] + lossless_cases(ref)[6:11]:      # palette pictures
        _, f = ref.features(data)
        W, H = f["width"], f["height"]
        cases = [data[:-k] for k in range(1, 12)] + (known if data is golden else [])
        for _ in range(200 if data is golden else 60):
            b = bytearray(data)
            b[int(rng.integers(len(b) // 4, len(b)))] ^= int(rng.integers(1, 256))
            cases.append(bytes(b))
        for b in cases:
            s_ref, want = ref.decode(b, 1, 0)
            out = np.zeros(W * H * 4, np.uint8)
            s_emu = L.emu_decode_window(b, len(b), 1, 0, out.ctypes.data, out.size, W * 4, 0, 0, 0, 0)
            if ref.features(b)[0] == 7:
                assert s_ref == 3          # WebPDecode's mapping of the probe's NOT_ENOUGH_DATA (the harness has no probe)
                continue
            assert s_emu == s_ref, (len(data), len(b), s_ref, s_emu)
            seen[s_ref] = seen.get(s_ref, 0) + 1
            if s_ref == 0:
                assert np.array_equal(want.reshape(-1), out)
    assert seen[0] > 20 and seen[3] > 20


def test_emu_lossless_scaling_matches_reference(ref):
    """options.use_scaling on lossless pictures (vp8l_dec.c:560-737): premultiply, four-channel rescaler, un-premultiply with
    the reference's spill-over, then the colourspace conversion or ConvertToYUVA on the scaled rows; up and down, crop + scale,
    flip, translucent and opaque pictures."""
    subprocess.check_call(["make", "-s", "-C", EMU_DIR])
    L = C.CDLL(os.path.join(EMU_DIR, "libvp8_emu.so"))
    L.emu_decode_scaled.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int,
                                    C.c_int, C.c_int, C.c_int, C.c_int]
    rng = np.random.default_rng(22)
    for data in lossless_cases(ref)[:9]:
        _, f = ref.features(data)
        W, H = f["width"], f["height"]
        for it in range(6 + (2 if W * H <= 4096 else 0)):   # two requests beyond 16383 on the small pictures
            crop = None
            if it >= 4 and it < 6:
                cw, ch = int(rng.integers(1, W + 1)), int(rng.integers(1, H + 1))
                crop = (int(rng.integers(0, W - cw + 1)), int(rng.integers(0, H - ch + 1)), cw, ch)
            w, h = (crop[2], crop[3]) if crop else (W, H)
            req = [(max(1, w // 2), max(1, h // 3)), (w * 2 + 1, h + 7), (max(1, w - 1), h * 3), (0, max(1, h // 2)), (w + 5, 0),
                   (int(rng.integers(1, 2 * w + 2)), int(rng.integers(1, 2 * h + 2))), (20001, max(1, h // 2)), (3, 16500)][it]
            flip = int(rng.integers(0, 2))
            for csp in (1, 7, 12, 10, 9, 0, 11, 6):
                s_ref, (sw, sh), want = ref.decode_scaled(data, csp, 8 if flip else 0, crop, req)
                assert s_ref == 0, (len(data), req, crop, s_ref)
                n = (sw * sh + 2 * ((sw + 1) // 2) * ((sh + 1) // 2) + (sw * sh if csp == 12 else 0)) if csp in (11, 12) else sw * sh * ref.BPP[csp]
                out = np.zeros(max(n, 16), np.uint8)
                c = crop or (0, 0, 0, 0)
                st = L.emu_decode_scaled(data, len(data), csp, 4 if flip else 0, out.ctypes.data, out.size,
                                         sw if csp in (11, 12) else sw * ref.BPP[csp], c[0], c[1], c[2], c[3], sw, sh)
                assert st == 0, (len(data), req, crop, csp, st)
                assert np.array_equal(out[:n], want), (len(data), W, H, req, crop, flip, csp, (sw, sh))


def test_emu_lossless_matches_manifest(lmanifest):
    """The committed lossless fixtures through the device code's host build: sha256 of every (colourspace, flags) combination
    equals the reference's recorded one (no reference library needed at run time)."""
    import hashlib
    subprocess.check_call(["make", "-s", "-C", EMU_DIR])
    L = C.CDLL(os.path.join(EMU_DIR, "libvp8_emu.so"))
    L.emu_decode_window.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int,
                                    C.c_int, C.c_int]
    bpps = {0: 3, 1: 4, 2: 3, 3: 4, 4: 4, 5: 2, 6: 2, 7: 4, 8: 4, 9: 4, 10: 2}
    for e in lmanifest:
        w, h = e["features"]["width"], e["features"]["height"]
        for key, want in e["sha256"].items():
            csp, fl = map(int, key.split(":"))
            n = (w * h + 2 * ((w + 1) // 2) * ((h + 1) // 2) + (w * h if csp == 12 else 0)) if csp in (11, 12) else w * h * bpps[csp]
            out = np.zeros(max(n, 16), np.uint8)
            st = L.emu_decode_window(e["data"], len(e["data"]), csp, 0, out.ctypes.data, out.size, w if csp in (11, 12) else w * bpps[csp],
                                     0, 0, 0, 0)
            assert st == 0 and hashlib.sha256(out[:n].tobytes()).hexdigest() == want, (e["file"], key)


@pytest.mark.parametrize("tool,args", [("fuzz_emu.py", ["--seed", "3"]), ("fuzz_encode.py", ["--seed", "4"]),
                                       ("fuzz_emu.py", ["--seed", "5", "--port", "--kinds", "lossy"])])
def test_campaign_prefix(ref, tool, args):
    """A few seconds of each parity campaign under tools/ (seeds whose long runs are logged in profiles/r01u_fuzz_*.log): mutated
    files and random encoder settings x decoding options, compiled reference against the device code's host build; the
    oracle's C restatement against the compiled reference. The case sequence per worker is fixed by the seed; only how far it
    gets depends on the machine."""
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", tool), "--seconds", "6", "--jobs", "2"] + args,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-1000:])
