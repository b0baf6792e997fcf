"""GPU parity tests: the CUDA path, called through the C ABI (WebPDecode / WebPDecodeBatch / the resident
batch), must be byte-identical to the reference. Checkers: the committed golden manifest (reference hashes), the
plain-C oracle, and the compiled reference when oracle/_ref travelled to the box. Integer/byte work: the bar
is bit-exact everywhere."""
import ctypes as C

import os

import numpy as np
import pytest

from conftest import sha

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def W(product):
    assert product.device_count() > 0, "no CUDA device: the product has no CPU path"
    return product


def test_webpdecode_matches_manifest(W, manifest):
    for e in manifest:
        for key, want in e["sha256"].items():
            csp, fl = map(int, key.split(":"))
            st, out = W.WebPDecode(e["data"], csp, bypass_filtering=fl & 1, no_fancy_upsampling=fl & 2)
            assert st == 0, (e["file"], key, st, W.last_error())
            assert sha(out) == want, (e["file"], key)


def test_internal_memory_and_strides(W, port, manifest):
    e = next(m for m in manifest if m["file"] == "odd_255x127_q50.webp")
    _, want = port.decode(e["data"], port.RGBA, 0)
    st, out = W.WebPDecode(e["data"], W.MODE_RGBA, external=False)          # library-allocated buffer
    assert st == 0 and np.array_equal(out, want)
    st, out = W.WebPDecode(e["data"], W.MODE_RGBA, stride=255 * 4 + 20)      # padded external rows
    assert st == 0 and np.array_equal(out[:, :255 * 4], want) and not out[:, 255 * 4:].any()
    st, out = W.WebPDecode(e["data"], W.MODE_YUV, external=False)
    _, wanty = port.decode(e["data"], port.YUV, 0)
    assert st == 0 and np.array_equal(out, wanty)
    # simple API
    L = W.lib()
    w, h = C.c_int(), C.c_int()
    p = L.WebPDecodeRGBA(e["data"], len(e["data"]), C.byref(w), C.byref(h))
    assert p and (w.value, h.value) == (255, 127)
    got = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), (127, 255 * 4)).copy()
    L.WebPFree(p)
    assert np.array_equal(got, want)


def test_batch_mixed_sizes_modes_and_failures(W, port, manifest):
    """One batch holding every fixture plus damaged files: per-item status and pixels as WebPDecode would give,
    a bad image never poisons its neighbours."""
    good = [e["data"] for e in manifest]
    d = next(m for m in manifest if m["file"] == "normal_8part_400x300.webp")["data"]
    rng = np.random.default_rng(11)
    bad = [d[:200], d[:3000], d[: len(d) // 2], d[:-1], d[:11], b"junkjunkjunkjunk"]
    for _ in range(10):
        b = bytearray(d)
        for _ in range(3):
            b[int(rng.integers(30, len(b)))] ^= int(rng.integers(1, 256))
        bad.append(bytes(b))
    datas = []
    for i in range(max(len(good), len(bad))):
        if i < len(good):
            datas.append(good[i])
        if i < len(bad):
            datas.append(bad[i])
    for csp in (W.MODE_RGBA, W.MODE_YUV, W.MODE_BGR):
        sts, outs = W.decode_batch(datas, csp)
        for i, data in enumerate(datas):
            want_st, want = port.decode(data, csp, 0)
            assert sts[i] == want_st, (i, len(data), sts[i], want_st)
            if want_st == 0:
                assert np.array_equal(outs[i].reshape(-1), want.reshape(-1)), (i, csp)


def test_fresh_corpora_against_compiled_reference(W, ref):
    """Seeded corpora encoded on the spot by the reference encoder: simple/normal filter, 1/4 segments,
    1..8 partitions, odd sizes; RGBA, rgbA, YUV; SIMD on and off on the reference side."""
    cases = [(640, 360, ref.cfg_simple_1part(), 41), (640, 360, ref.cfg_normal_8part(), 42), (256, 256, ref.cfg_default(), 43),
             (1920, 1080, ref.cfg_simple_1part(), 44), (1920, 1080, ref.cfg_normal_8part(), 45),
             (333, 77, ref.EncCfg(60, 4, partitions=1, low_memory=1, segments=3), 46), (16, 16, ref.cfg_default(30), 47),
             (2048, 64, ref.cfg_default(90), 48), (64, 2048, ref.cfg_normal_8part(20), 49)]
    datas = [ref.encode(ref.synth(w, h, seed), cfg) for (w, h, cfg, seed) in cases]
    for csp in (W.MODE_RGBA, W.MODE_rgbA, W.MODE_YUV):
        for fl in (0, 1, 2):
            if csp == W.MODE_YUV and fl == 2:
                continue
            sts, outs = W.decode_batch(datas, csp, bypass_filtering=fl & 1, no_fancy_upsampling=fl & 2)
            for i, data in enumerate(datas):
                for simd in (True, False):
                    st, want = ref.decode(data, csp, fl, simd=simd)
                    assert st == 0 and sts[i] == 0
                    assert np.array_equal(outs[i].reshape(-1), want.reshape(-1)), (cases[i][:2], csp, fl, simd)


def test_resident_batch_device_outputs_and_repeat(W, port, manifest):
    """Resident API: decode twice from the same uploaded inputs, identical results; timings populated."""
    datas = [e["data"] for e in manifest] * 3
    b = W.Batch(datas, W.MODE_RGBA)
    try:
        assert b.create() == 0
        assert b.decode() == 0
        t = b.timings()
        assert t["launches"] >= 5 and t["total_ms"] > 0
        assert b.download() == 0
        first = [b.host_output(i).copy() for i in range(b.n)]
        assert b.decode() == 0 and b.download() == 0
        for i in range(b.n):
            assert np.array_equal(first[i], b.host_output(i))
            _, want = port.decode(datas[i], port.RGBA, 0)
            assert np.array_equal(first[i], want)
        p = b.device_output(0)
        assert p is not None and p.y_or_rgba and p.width == b.dims[0][0]
    finally:
        b.close()


def test_waves_small_scratch(W, port, manifest):
    """Force several waves through a tiny scratch budget: same bytes."""
    datas = [e["data"] for e in manifest] * 4
    b = W.Batch(datas, W.MODE_RGBA, scratch_bytes=700 * 1200)
    try:
        assert b.decode_oneshot() == 0
        for i in range(b.n):
            _, want = port.decode(datas[i], port.RGBA, 0)
            assert np.array_equal(b.host_output(i), want), i
    finally:
        b.close()


def test_dithering(W, ref, manifest, amanifest):
    """options.dithering_strength on the device (dwebp's default is 50): identical to the reference where it changes
    pixels (smooth pictures at fine quantisers, both loop filters, crop windows) and where it does not (the manifest);
    options.alpha_dithering_strength (dwebp: 100) is a no-op unless the encoder quantised the alpha levels; where it
    did, the plane is de-banded on the device exactly like WebPDequantizeLevels (k_alpha_smooth)."""
    from conftest import smooth_image
    for e in manifest:
        st, out = W.WebPDecode(e["data"], W.MODE_RGBA, dithering_strength=50)
        s_ref, want = ref.decode_dithered(e["data"], W.MODE_RGBA, 0, None, 50)
        assert st == s_ref == 0 and np.array_equal(out.reshape(-1), want), e["file"]
    rng = np.random.default_rng(32)
    changed = 0
    for k, (q, segs, flt, w, h) in enumerate(((100, 1, 0, 200, 136), (99, 4, 1, 177, 93), (97, 4, 1, 64, 200), (95, 2, 0, 320, 48), (90, 4, 1, 1920, 1080))):
        data = ref.encode(smooth_image(w, h, 40 + k), ref.EncCfg(q, 4, segments=segs, filter_type=flt, filter_strength=40))
        for strength in (50, 100):
            for it in range(3):
                crop = None
                if it > 0:
                    cw, ch = int(rng.integers(1, w + 1)), int(rng.integers(1, h + 1))
                    crop = (int(rng.integers(0, w - cw + 1)), int(rng.integers(0, h - ch + 1)), cw, ch)
                for csp in (W.MODE_RGBA, W.MODE_YUV):
                    s_ref, want = ref.decode_dithered(data, csp, 0, crop, strength)
                    _, plain = ref.decode_dithered(data, csp, 0, crop, 0)
                    changed += int((want != plain).sum())
                    st, out = W.WebPDecode(data, csp, dithering_strength=strength, crop=crop)
                    assert st == s_ref == 0, (q, strength, crop, csp, st, W.last_error())
                    assert np.array_equal(out.reshape(-1)[:want.size], want), (q, segs, flt, strength, crop, csp)
    assert changed > 10000
    for e in amanifest:
        i = e["data"].find(b"ALPH")
        quantised = ((e["data"][i + 8] >> 4) & 3) == 1
        for csp, crop, ad in ((W.MODE_RGBA, None, 100), (W.MODE_rgbA, (3, 5, 90, 60), 50), (W.MODE_YUVA, None, 30)):
            if crop is not None and (e["features"]["width"] < 100 or e["features"]["height"] < 70):
                crop = None
            st, out = W.WebPDecode(e["data"], csp, dithering_strength=50, alpha_dithering_strength=ad, crop=crop)
            s_ref, want = ref.decode_dithered(e["data"], csp, 0, crop, 50, ad)
            assert st == s_ref == 0 and np.array_equal(out.reshape(-1)[:want.size], want), (e["file"], csp, crop, ad)
            if quantised and csp == W.MODE_RGBA:
                _, plain = ref.decode_dithered(e["data"], csp, 0, crop, 50, 0)
                assert (want != plain).any(), e["file"]     # the de-banding does move pixels on these planes


def test_full_size_batch_properties(W, ref):
    """BASELINE config 2 at reduced count (64 distinct 1080p images, simple filter, 1 partition) and config 4 shape
    (256x256 q80), decoded as one batch; every image compared with the reference; plus size-independent
    properties: idempotence of repeated decodes and a checksum of checksums equal to the reference's."""
    import hashlib
    d2 = ref.encode_corpus(16, 1920, 1080, ref.cfg_simple_1part(), seed0=1000)
    d3 = ref.encode_corpus(8, 1920, 1080, ref.cfg_normal_8part(), seed0=2000)
    d4 = ref.encode_corpus(128, 256, 256, ref.cfg_default(), seed0=3000)
    datas = d2 + d3 + d4
    sts, outs = W.decode_batch(datas, W.MODE_RGBA)
    h_gpu, h_ref = hashlib.sha256(), hashlib.sha256()
    for i, data in enumerate(datas):
        st, want = ref.decode(data, ref.MODE_RGBA, 0)
        assert st == 0 and sts[i] == 0
        assert np.array_equal(outs[i], want), i
        h_gpu.update(sha(outs[i]).encode()); h_ref.update(sha(want).encode())
    assert h_gpu.hexdigest() == h_ref.hexdigest()
    sts2, outs2 = W.decode_batch(datas, W.MODE_RGBA)
    assert all(np.array_equal(a, b) for a, b in zip(outs, outs2))


def test_alpha_matches_manifest(W, amanifest):
    """ALPH chunks decoded on the device (VP8L subset + row unfilters), alpha merged by the output kernel."""
    for e in amanifest:
        for key, want in e["sha256"].items():
            csp, fl = map(int, key.split(":"))
            st, out = W.WebPDecode(e["data"], csp, bypass_filtering=fl & 1, no_fancy_upsampling=fl & 2)
            assert st == 0, (e["file"], key, st, W.last_error())
            assert sha(out) == want, (e["file"], key)


def test_alpha_batch_with_opaque_and_damaged(W, ref, manifest, amanifest):
    """One batch mixing opaque images, images with alpha and damaged ALPH payloads; decoded twice on the resident path
    (the alpha work areas are planned once and reused)."""
    rng = np.random.default_rng(17)
    datas = [e["data"] for e in manifest[:4]] + [e["data"] for e in amanifest]
    for e in amanifest[:4]:
        d = e["data"]
        i = d.find(b"ALPH")
        size = int.from_bytes(d[i + 4:i + 8], "little")
        for _ in range(3):
            b = bytearray(d)
            b[i + 8 + int(rng.integers(0, size))] ^= int(rng.integers(1, 256))
            datas.append(bytes(b))
    for csp in (W.MODE_RGBA, W.MODE_rgbA, W.MODE_Argb):
        sts, outs = W.decode_batch(datas, csp, device=0)
        for d, st, out in zip(datas, sts, outs):
            s_ref, want = ref.decode(d, csp, 0)
            assert st == s_ref, (len(d), st, s_ref)
            if s_ref == 0:
                assert np.array_equal(out.reshape(-1), want.reshape(-1))


def test_alpha_large_gradient(W, ref):
    """Config-5 shape at a size the CPU checker finishes quickly: 1024x1024 q90 with a gradient-filtered ALPH chunk."""
    cfg = ref.EncCfg(90.0, 4, alpha_filtering=2)
    data = ref.encode(ref.synth(1024, 1024, 77, alpha=True), cfg)
    for csp in (W.MODE_RGBA, W.MODE_rgbA):
        s_ref, want = ref.decode(data, csp, 0)
        st, out = W.WebPDecode(data, csp)
        assert st == s_ref == 0 and np.array_equal(out.reshape(-1), want.reshape(-1))


def test_incremental_api(W, port, manifest, amanifest):
    """dwebp -incremental's call sequence: WebPIDecode(config) + WebPIUpdate with a growing prefix, and
    WebPINewDecoder + WebPIAppend in chunks; SUSPENDED until the last byte, then the same pixels as WebPDecode."""
    L = W.lib()
    L.WebPINewDecoder.restype = C.c_void_p
    L.WebPINewDecoder.argtypes = [C.c_void_p]
    L.WebPIDecode.restype = C.c_void_p
    L.WebPIDecode.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p]
    L.WebPIAppend.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
    L.WebPIUpdate.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
    L.WebPIDelete.argtypes = [C.c_void_p]
    for e in (manifest[2], amanifest[0]):
        data = e["data"]
        _, want = W.WebPDecode(data, W.MODE_RGBA)
        w, h = e["features"]["width"], e["features"]["height"]
        # update mode with a caller config
        cfg = W._new_config(W.MODE_RGBA, 0, 0)
        idec = L.WebPIDecode(data, len(data), C.byref(cfg))
        assert idec
        for n in (30, len(data) // 3, len(data) - 1):
            assert L.WebPIUpdate(idec, data, n) == W.VP8_STATUS_SUSPENDED
        assert L.WebPIUpdate(idec, data, len(data)) == W.VP8_STATUS_OK
        got = np.ctypeslib.as_array(C.cast(cfg.output.u.RGBA.rgba, C.POINTER(C.c_uint8)), (h, cfg.output.u.RGBA.stride)).copy()
        L.WebPIDelete(idec)
        L.WebPFreeDecBuffer(C.byref(cfg.output))
        assert np.array_equal(got[:, :w * 4], want[:, :w * 4])
        # append mode into a caller buffer object
        buf = W.WebPDecBuffer()
        L.WebPInitDecBufferInternal(C.byref(buf), W.WEBP_DECODER_ABI_VERSION)
        buf.colorspace = W.MODE_RGBA
        idec = L.WebPINewDecoder(C.byref(buf))
        st, pos = W.VP8_STATUS_SUSPENDED, 0
        for chunk in (100, 1000, len(data)):
            nxt = min(len(data), pos + chunk)
            st = L.WebPIAppend(idec, data[pos:nxt], nxt - pos)
            pos = nxt
            assert st == (W.VP8_STATUS_OK if pos == len(data) else W.VP8_STATUS_SUSPENDED)
        got = np.ctypeslib.as_array(C.cast(buf.u.RGBA.rgba, C.POINTER(C.c_uint8)), (h, buf.u.RGBA.stride)).copy()
        L.WebPIDelete(idec)
        L.WebPFreeDecBuffer(C.byref(buf))
        assert np.array_equal(got[:, :w * 4], want[:, :w * 4])
        # WebPINewRGB with an external buffer + the getters: no rows before the last byte, the whole picture after
        L.WebPINewRGB.restype = C.c_void_p
        L.WebPINewRGB.argtypes = [C.c_int, C.c_void_p, C.c_size_t, C.c_int]
        L.WebPIDecGetRGB.restype = C.c_void_p
        L.WebPIDecGetRGB.argtypes = [C.c_void_p] + [C.POINTER(C.c_int)] * 4
        L.WebPIDecodedArea.restype = C.c_void_p
        L.WebPIDecodedArea.argtypes = [C.c_void_p] + [C.POINTER(C.c_int)] * 4
        ext = np.zeros((h, w * 4 + 16), np.uint8)
        idec = L.WebPINewRGB(W.MODE_RGBA, ext.ctypes.data, ext.size, w * 4 + 16)
        assert idec and not L.WebPINewRGB(W.MODE_YUV, None, 0, 0) and not L.WebPINewRGB(W.MODE_RGBA, ext.ctypes.data, 0, 0)
        ly, ww, hh, ss = C.c_int(-1), C.c_int(-1), C.c_int(-1), C.c_int(-1)
        assert L.WebPIAppend(idec, data[:20], 20) == W.VP8_STATUS_SUSPENDED
        assert not L.WebPIDecGetRGB(idec, C.byref(ly), C.byref(ww), C.byref(hh), C.byref(ss))        # headers still arriving: no area
        assert not L.WebPIDecodedArea(idec, None, None, C.byref(ww), C.byref(hh)) and ww.value == 0 and hh.value == 0
        assert L.WebPIAppend(idec, data[20:len(data) - 1], len(data) - 21) == W.VP8_STATUS_SUSPENDED
        # all but the last byte in: the caller's buffer with the picture's dimensions and no rows yet (tests/test_abi.py walks this
        # against the reference append by append)
        p = L.WebPIDecGetRGB(idec, C.byref(ly), C.byref(ww), C.byref(hh), C.byref(ss))
        assert p == ext.ctypes.data and (ly.value, ww.value, hh.value, ss.value) == (0, w, h, w * 4 + 16)
        assert L.WebPIDecodedArea(idec, None, None, C.byref(ww), C.byref(hh)) and (ww.value, hh.value) == (w, 0)
        assert not ext.any()
        assert L.WebPIAppend(idec, data[len(data) - 1:], 1) == W.VP8_STATUS_OK
        p = L.WebPIDecGetRGB(idec, C.byref(ly), C.byref(ww), C.byref(hh), C.byref(ss))
        assert p == ext.ctypes.data and (ly.value, ww.value, hh.value, ss.value) == (h, w, h, w * 4 + 16)
        assert L.WebPIDecodedArea(idec, None, None, C.byref(ww), C.byref(hh)) and (ww.value, hh.value) == (w, h)
        L.WebPIDelete(idec)
        assert np.array_equal(ext[:, :w * 4], want[:, :w * 4])
        # WebPINewYUVA with library-allocated planes
        L.WebPINewYUVA.restype = C.c_void_p
        L.WebPINewYUVA.argtypes = [C.c_void_p, C.c_size_t, C.c_int] * 4
        L.WebPIDecGetYUVA.restype = C.c_void_p
        L.WebPIDecGetYUVA.argtypes = [C.c_void_p, C.POINTER(C.c_int)] + [C.POINTER(C.c_void_p)] * 3 + [C.POINTER(C.c_int)] * 5
        idec = L.WebPINewYUVA(None, 0, 0, None, 0, 0, None, 0, 0, None, 0, 0)
        assert idec and L.WebPIAppend(idec, data, len(data)) == W.VP8_STATUS_OK
        up, vp, ap = C.c_void_p(), C.c_void_p(), C.c_void_p()
        st_, uvs, as_ = C.c_int(), C.c_int(), C.c_int()
        yp = L.WebPIDecGetYUVA(idec, C.byref(ly), C.byref(up), C.byref(vp), C.byref(ap), C.byref(ww), C.byref(hh), C.byref(st_), C.byref(uvs), C.byref(as_))
        assert yp and up.value and vp.value and (ly.value, ww.value, hh.value) == (h, w, h)
        _, want_yuv = W.WebPDecode(data, W.MODE_YUVA)
        got_y = np.ctypeslib.as_array(C.cast(yp, C.POINTER(C.c_uint8)), (h, st_.value))[:, :w]
        assert np.array_equal(got_y.reshape(-1), want_yuv[:w * h])
        assert not L.WebPIDecGetRGB(idec, None, None, None, None)   # wrong family
        L.WebPIDelete(idec)


def test_crop_and_flip(W, ref, manifest, amanifest):
    """options.use_cropping / options.flip through the C ABI against the reference (see tests/test_emu.py for the rules)."""
    rng = np.random.default_rng(21)
    for e in list(manifest) + list(amanifest):
        Wd, Hd = e["features"]["width"], e["features"]["height"]
        for it in range(3):
            crop = None
            if it > 0:
                cw, ch = int(rng.integers(1, Wd + 1)), int(rng.integers(1, Hd + 1))
                crop = (int(rng.integers(0, Wd - cw + 1)), int(rng.integers(0, Hd - ch + 1)), cw, ch)
            flip = bool(rng.integers(0, 2))
            for csp in (W.MODE_RGBA, W.MODE_rgbA, W.MODE_BGR, W.MODE_YUV, W.MODE_RGB_565, W.MODE_RGBA_4444, W.MODE_rgbA_4444,
                        W.MODE_YUVA, W.MODE_Argb, W.MODE_bgrA):
                s_ref, want = ref.decode_window(e["data"], csp, 8 if flip else 0, crop)
                st, got = W.WebPDecode(e["data"], csp, crop=crop, flip=flip)
                assert st == s_ref, (e["file"], crop, flip, csp, st, s_ref)
                if s_ref == 0:
                    assert np.array_equal(want, got.reshape(-1)), (e["file"], crop, flip, csp)
    data = manifest[1]["data"]
    for crop in ((300, 0, 40, 10), (0, 0, 0, 10), (0, 199, 10, 3)):
        st, _ = W.WebPDecode(data, W.MODE_RGBA, crop=crop)
        assert st == W.VP8_STATUS_INVALID_PARAM
    cut = data[: len(data) * 2 // 3]
    for crop in ((0, 0, 320, 32), (0, 120, 320, 80)):
        s_ref, want = ref.decode_window(cut, W.MODE_RGBA, 0, crop)
        st, got = W.WebPDecode(cut, W.MODE_RGBA, crop=crop)
        assert st == s_ref
        if s_ref == 0:
            assert np.array_equal(want, got.reshape(-1))


@pytest.mark.parametrize("mapping", ["f:0", "f:1", "f:0:band", "f:1:band", "f:0:ring", "f:1:ring", "warp", "k", "k:1", "k:2", "lanes"])
def test_every_token_mapping(mapping):
    """The default token parser (f: lockstep lanes with the fp32 boolean decoder and token-stream output) picks its run style
    by lanes per warp; here both are forced in turn (f:0 a branch per decode, f:1 straight-line groups), and so is each of
    the older mappings that write the dense level plane (one warp per partition, lockstep lanes, lane state machine) (WEBP_B200_TOKEN_MAP is read once per process, hence the subprocess) and
    must pass the manifest, mixed-batch and fresh-corpus parity tests above."""
    import os
    import subprocess
    import sys
    if os.environ.get("WEBP_B200_TOKEN_MAP_INNER"):
        pytest.skip("inner run")
    env = dict(os.environ, WEBP_B200_TOKEN_MAP=mapping.split(":")[0], WEBP_B200_TOKEN_MAP_INNER="1")
    if ":" in mapping:   # the lockstep parsers' ways of running their lanes (block ends on the spot, grouped event points, straight-line groups)
        env["WEBP_B200_TOKEN_GROUPED"] = mapping.split(":")[1]
    if mapping.endswith(":band"):   # the banded layout of the probability rows (launches with very many small images take it by themselves)
        env["WEBP_B200_TOKEN_BAND"] = "1"
    if mapping.endswith(":ring"):   # compressed bytes through shared-memory rings filled by cp.async.bulk instead of global loads
        env["WEBP_B200_TOKEN_RING"] = "1"
    select = "manifest or mixed_sizes or fresh_corpora or full_size or damaged or many_streams"
    if not mapping.startswith("f"):   # only the fp parser records the rows the both-chunks-damaged rule needs
        select = "(%s) and not both_damaged" % select
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-x", "-q", "-m", "gpu", "-k", select], env=env, capture_output=True, text=True,
                       cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("switch", ["WEBP_B200_MODES=lockstep", "WEBP_B200_MODES=lockstep WEBP_B200_MODES_LANES=3", "WEBP_B200_RECON_ROWS=0 WEBP_B200_RECON_WARPS=16",
                                    "WEBP_B200_RECON_ROWS=0 WEBP_B200_PIXEL_WARPS=8"])
def test_every_mode_and_reconstruction_mapping(switch):
    """The second instantiations of K1 and K3 kept behind switches: the intra-mode parse as lockstep lanes of a table-driven
    state machine (vp8_modes_lockstep.h; default: one image per warp) and the reconstruction as an anti-diagonal wavefront
    with a block-wide barrier, at 8 / 16 warps per image whatever the picture's size (default: one warp per macroblock row,
    8 warps, 4 for small pictures, in K3 and K4 alike). The switches are read once per process, hence the subprocess; each
    must pass the manifest, mixed-batch, fresh-corpus, damaged-file and stage-level tests."""
    import os
    import subprocess
    import sys
    if os.environ.get("WEBP_B200_TOKEN_MAP_INNER"):
        pytest.skip("inner run")
    env = dict(os.environ, WEBP_B200_TOKEN_MAP_INNER="1")
    for kv in switch.split():
        k, v = kv.split("=")
        env[k] = v
    select = "manifest or mixed_sizes or fresh_corpora or full_size or damaged or parse_stages"
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-x", "-q", "-m", "gpu", "-k", select], env=env, capture_output=True, text=True,
                       cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


def test_many_streams_take_the_lockstep_parser(W, ref):
    """Enough streams per partition count for the default choice to be the lockstep parser (>= two per SM sub-partition):
    1300 one-partition and 300 eight-partition images of mixed small sizes, every one compared with the reference."""
    d1 = [ref.encode(ref.synth(48 + 16 * (k % 3), 40 + 8 * (k % 5), 7000 + k), ref.cfg_simple_1part(40 + k)) for k in range(40)]
    d8 = [ref.encode(ref.synth(64 + 16 * (k % 4), 144 + 16 * (k % 2), 7100 + k), ref.cfg_normal_8part(50 + 2 * k)) for k in range(20)]
    distinct = d1 + d8
    order = [i % 40 for i in range(1300)] + [40 + i % 20 for i in range(300)]
    datas = [distinct[j] for j in order]
    sts, outs = W.decode_batch(datas, W.MODE_RGBA)
    wants = [ref.decode(d, ref.MODE_RGBA, 0) for d in distinct]
    for i, j in enumerate(order):
        st, want = wants[j]
        assert st == 0 and sts[i] == 0, i
        assert np.array_equal(outs[i], want), i


def test_row_bands(ref):
    """WEBP_B200_BANDS=4 (read once per process, hence the subprocess): a wave of single-partition images large enough for
    the lockstep parser is decoded in row bands -- token parse, reconstruction, loop filter, output and download of band
    k before band k+1, the parser's and the reconstruction's state handed from launch to launch. Images taller and
    shorter than a band, simple and normal loop filter, several 4-byte colourspaces; every image against the reference."""
    import os
    import subprocess
    import sys
    code = (
        "import numpy as np, sys\n"
        "sys.path.insert(0, %r)\n"
        "import libwebp_b200 as W\n"
        "from oracle import refwebp as ref\n"
        "cfg_n = ref.EncCfg(70, 4, partitions=0, low_memory=0, segments=4)\n"
        "distinct = []\n"
        "for k in range(24):\n"
        "    w, h = 48 + 16 * (k %% 3), (40, 130, 272, 300, 415, 512)[k %% 6]\n"
        "    distinct.append(ref.encode(ref.synth(w, h, 9000 + k), ref.cfg_simple_1part(50 + k) if k %% 2 else cfg_n))\n"
        "datas = [distinct[i %% 24] for i in range(1300)]\n"
        "for csp in (W.MODE_RGBA, W.MODE_BGRA, W.MODE_ARGB):\n"
        "    sts, outs = W.decode_batch(datas, csp)\n"
        "    wants = [ref.decode(d, csp, 0) for d in distinct]\n"
        "    for i in range(1300):\n"
        "        st, want = wants[i %% 24]\n"
        "        assert st == 0 and sts[i] == 0, (csp, i, sts[i])\n"
        "        assert np.array_equal(outs[i].reshape(-1), want.reshape(-1)), (csp, i)\n"
        "print('bands ok')\n"
    ) % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for overlap in ("0", "1"):
        r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, WEBP_B200_BANDS="4", WEBP_B200_BAND_OVERLAP=overlap, WEBP_B200_TOKEN_MAP="k"),   # row bands: the older lockstep parser
                           capture_output=True, text=True)
        assert r.returncode == 0 and "bands ok" in r.stdout, r.stdout[-1500:] + r.stderr[-1500:]


def test_scaling(W, ref, manifest, amanifest):
    """options.use_scaling through the C ABI: the rescaler runs on the device (one thread per output column), output
    identical to the reference's for up- and downscaling, ratio-preserving requests, crop + scale, flip, packed and
    planar colourspaces; impossible requests are refused with the reference's status; files with an ALPH chunk have their alpha rescaled too."""
    rng = np.random.default_rng(22)
    for e in manifest:
        Wd, Hd = e["features"]["width"], e["features"]["height"]
        for it in range(5):
            crop = None
            if it >= 3:
                cw, ch = int(rng.integers(1, Wd + 1)), int(rng.integers(1, Hd + 1))
                crop = (int(rng.integers(0, Wd - cw + 1)), int(rng.integers(0, Hd - ch + 1)), cw, ch)
            w, h = (crop[2], crop[3]) if crop else (Wd, Hd)
            req = [(max(1, w // 2), max(1, h // 3)), (w * 2 + 1, h + 7), (0, max(1, h // 2)),
                   (int(rng.integers(1, 2 * w + 2)), int(rng.integers(1, 2 * h + 2))), (max(1, w - 1), h * 2)][it]
            flip = bool(rng.integers(0, 2))
            for csp in (W.MODE_RGBA, W.MODE_RGB, W.MODE_YUV, W.MODE_RGB_565):
                s_ref, (sw, sh), want = ref.decode_scaled(e["data"], csp, 8 if flip else 0, crop, req)
                st, out = W.WebPDecode(e["data"], csp, crop=crop, flip=flip, scaled=req)
                assert st == s_ref == 0, (e["file"], req, crop, csp, st, s_ref, W.last_error())
                assert np.array_equal(out.reshape(-1)[:want.size], want), (e["file"], req, crop, flip, csp, (sw, sh))
    for e in manifest:   # beyond 16383 (round 1 refused these): the reference rescales to anything its allocator accepts
        if e["features"]["width"] * e["features"]["height"] > 4096:
            continue
        for req in ((20001, 5), (3, 16500), (40000, 0)):
            for csp in (W.MODE_RGBA, W.MODE_YUV):
                s_ref, (sw, sh), want = ref.decode_scaled(e["data"], csp, 0, None, req)
                st, out = W.WebPDecode(e["data"], csp, scaled=req)
                assert st == s_ref == 0, (e["file"], req, csp, st, s_ref, W.last_error())
                assert np.array_equal(out.reshape(-1)[:want.size], want), (e["file"], req, csp, (sw, sh))
    data = manifest[0]["data"]
    for req in ((0, 0), (-3, 10)):
        s_ref, _, _ = ref.decode_scaled(data, W.MODE_RGBA, 0, None, req)
        st, _ = W.WebPDecode(data, W.MODE_RGBA, scaled=req)
        assert st == s_ref != 0, (req, st, s_ref)
    for e in amanifest:   # scaled alpha: plain, premultiplied, 4-bit and planar
        w, h = e["features"]["width"], e["features"]["height"]
        for req in ((max(1, w // 2), max(1, h // 2)), (w + 9, 2 * h), (0, max(1, h - 3))):
            for csp in (W.MODE_RGBA, W.MODE_rgbA, W.MODE_YUVA, W.MODE_rgbA_4444, W.MODE_Argb):
                s_ref, (sw, sh), want = ref.decode_scaled(e["data"], csp, 0, None, req)
                st, out = W.WebPDecode(e["data"], csp, scaled=req)
                assert st == s_ref == 0, (e["file"], req, csp, st, s_ref, W.last_error())
                assert np.array_equal(out.reshape(-1)[:want.size], want), (e["file"], req, csp, (sw, sh))


@pytest.mark.gpu
def test_lossless_pictures(W, ref, manifest, amanifest):
    """Whole-picture VP8L on the device: every transform, colour cache, meta-Huffman groups and
    palette bundling (tests/test_emu.py:lossless_cases), every colourspace incl. MODE_YUV / MODE_YUVA, crop windows at odd offsets, flip;
    then one batch mixing lossless, lossy, alpha and damaged lossless files (per-item status as the reference's), decoded
    twice on the resident path; options.use_scaling on lossless pictures (four-channel rescaler between premultiply and un-premultiply)."""
    from test_emu import lossless_cases
    cases = lossless_cases(ref)
    rng = np.random.default_rng(6)
    for data in cases:
        _, f = ref.features(data)
        w, h = f["width"], f["height"]
        for it in range(3):
            crop = None
            if it > 0:
                cw, ch = int(rng.integers(1, w + 1)), int(rng.integers(1, h + 1))
                crop = (int(rng.integers(0, w - cw + 1)), int(rng.integers(0, h - ch + 1)), cw, ch)
            flip = bool(rng.integers(0, 2))
            for csp in (1, 7, 0, 2, 3, 8, 4, 9, 5, 6, 10, 11, 12):
                s_ref, want = ref.decode_window(data, csp, 8 if flip else 0, crop)
                st, out = W.WebPDecode(data, csp, crop=crop, flip=flip)
                assert st == s_ref == 0, (len(data), crop, flip, csp, st, W.last_error())
                assert np.array_equal(out.reshape(-1)[:want.size], want), (len(data), w, h, crop, flip, csp)
        for csp, req, crop in ((W.MODE_RGBA, (max(1, w // 2), max(1, h // 3)), None), (W.MODE_rgbA, (w * 2 + 1, h + 7), None),
                               (W.MODE_YUVA, (max(1, w - 1), 0), None), (W.MODE_BGR, (w + 3, max(1, h // 2)), (w // 4, h // 4, max(1, w // 2), max(1, h // 2)))):
            s_ref, (sw, sh), want = ref.decode_scaled(data, csp, 0, crop, req)
            st, out = W.WebPDecode(data, csp, crop=crop, scaled=req)
            assert st == s_ref == 0, (len(data), csp, req, crop, st, W.last_error())
            assert np.array_equal(out.reshape(-1)[:want.size], want), (len(data), w, h, csp, req, crop)
    datas = list(cases) + [e["data"] for e in manifest[:3]] + [e["data"] for e in amanifest[:3]]
    for data in cases[:8]:
        for k in range(4):
            b = bytearray(data)
            if k == 3 and len(b) > 40:
                b = b[: int(rng.integers(30, len(b)))]
            else:
                b[int(rng.integers(20, len(b)))] ^= int(rng.integers(1, 256))
            datas.append(bytes(b))
    order = rng.permutation(len(datas))
    datas = [datas[i] for i in order]
    for csp in (W.MODE_RGBA, W.MODE_bgrA):
        for rep in range(2):
            sts, outs = W.decode_batch(datas, csp, device=0)
            for d, st, out in zip(datas, sts, outs):
                s_ref, want = ref.decode(d, csp, 0)
                assert st == s_ref, (len(d), st, s_ref)
                if s_ref == 0:
                    assert np.array_equal(out.reshape(-1), want.reshape(-1)), len(d)
    # a 1080p lossless picture and the incremental shim
    big = np.zeros((1080, 1920, 4), np.uint8)
    big[..., :3] = ref.synth(1920, 1080, 4242)
    big[..., 3] = 255
    data = ref.encode(big, ref.EncCfg(50, 2, lossless=1))
    st, out = W.WebPDecode(data, W.MODE_RGBA)
    assert st == 0 and np.array_equal(out.reshape(1080, 1920, 4), big)


def test_concurrent_callers(W, ref, manifest, amanifest):
    """SURVEY.md 8(b) threading row: WebPDecode / WebPDecodeBatch are callable from many host threads at once (ctypes drops
    the GIL around the call; the library serialises batches per device). Eight threads, each decoding its own mix of
    files in its own colourspace, several rounds; every result equals the reference's."""
    import threading
    files = [e["data"] for e in list(manifest) + list(amanifest)]
    csps = (W.MODE_RGBA, W.MODE_BGR, W.MODE_YUV, W.MODE_rgbA, W.MODE_RGB_565, W.MODE_ARGB, W.MODE_YUVA, W.MODE_bgrA)
    want = {(i, csp): ref.decode(d, csp, 0) for i, d in enumerate(files) for csp in csps}
    errors = []

    def worker(t):
        try:
            csp = csps[t]
            for rnd in range(3):
                if (t + rnd) % 2 == 0:
                    for i in range(t % 3, len(files), 3):
                        st, out = W.WebPDecode(files[i], csp)
                        s_ref, w = want[(i, csp)]
                        if st != s_ref or (st == 0 and not np.array_equal(out.reshape(-1)[:w.size], w.reshape(-1))):
                            errors.append((t, rnd, i, st, s_ref))
                else:
                    sts, outs = W.decode_batch(files, csp, device=0)
                    for i, (st, out) in enumerate(zip(sts, outs)):
                        s_ref, w = want[(i, csp)]
                        if st != s_ref or (st == 0 and not np.array_equal(out.reshape(-1)[:w.size], w.reshape(-1))):
                            errors.append((t, rnd, i, st, s_ref))
        except Exception as e:   # noqa: BLE001 - report whatever a thread hit
            errors.append((t, repr(e)))

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(8)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors[:5]


def test_lossless_matches_manifest(W, lmanifest):
    for e in lmanifest:
        st, f = W.WebPGetFeatures(e["data"])
        assert st == 0 and f == e["features"]
        for key, want in e["sha256"].items():
            csp, fl = map(int, key.split(":"))
            st, out = W.WebPDecode(e["data"], csp)
            assert st == 0, (e["file"], key, st, W.last_error())
            assert sha(out) == want, (e["file"], key)


@pytest.mark.gpu
def test_lossless_palette_picture_losing_its_last_bits(W, lmanifest):
    """A whole VP8L picture runs the reference's 32-bit pixel loop even when it is a plain palette picture (VP8LDecodeImage,
    vp8l_dec.c:1761-1765): decoded to the last pixel with the reader past the end of the data is a BITSTREAM_ERROR there, unlike
    in an ALPH payload. The reference's statuses for these five damaged copies of a golden file were recorded by
    tools/fuzz_emu.py (all 3); the intact file and a copy damaged where it still decodes ride along in the same batch."""
    golden = next(e for e in lmanifest if e["file"] == "lossless_palette5_121x50.webp")
    datas, want = [golden["data"]], [0]
    for at, val in ((704, 0x97), (1476, 0xfd), (755, 0xd5), (1436, 0xd7), (235, 0xb0)):
        b = bytearray(golden["data"])
        b[at] = val
        datas.append(bytes(b))
        want.append(3)
    sts, outs = W.decode_batch(datas, W.MODE_RGBA)
    assert list(sts) == want, list(sts)
    assert sha(outs[0]) == golden["sha256"]["1:0"]
    for d, s_want in zip(datas, want):      # and one at a time through WebPDecode
        st, out = W.WebPDecode(d, W.MODE_RGBA)
        assert st == s_want, (st, s_want)


@pytest.mark.gpu
def test_extreme_dimensions(W, ref):
    """The format's largest dimensions (14 bits: 16383) on one axis at a time -- 1024 macroblocks per row, 1024 macroblock rows --
    lossy with 1 and 8 token partitions and both loop filters, with an ALPH chunk, and lossless; alone, mixed into one batch with
    small pictures (the wave's shared-memory layouts are sized by its widest picture), and as enough streams for the lockstep
    parser to be the default choice. Every picture against the compiled reference."""
    rng = np.random.default_rng(5)

    def pic(w, h, alpha=False):
        p = np.zeros((h, w, 4), np.uint8)
        p[..., :3] = rng.integers(0, 256, (h, w, 3), dtype=np.uint8) // 4 * 2 + (np.arange(w)[None, :, None] // 64 % 2 * 60).astype(np.uint8)
        p[..., 3] = (np.arange(w)[None, :] * 7 + np.arange(h)[:, None] * 3) % 256 if alpha else 255
        return p

    big = [ref.encode(pic(16383, 17), ref.EncCfg(60, 2, partitions=3, segments=4, filter_type=1)),
           ref.encode(pic(17, 16383), ref.EncCfg(60, 2, partitions=3, segments=4, filter_type=1)),
           ref.encode(pic(16383, 33), ref.EncCfg(75, 4, partitions=0, segments=1, filter_type=0)),
           ref.encode(pic(16383, 20, True), ref.EncCfg(80, 4, alpha_filtering=2)),
           ref.encode(pic(21, 16383, True), ref.EncCfg(80, 4, alpha_filtering=1, alpha_quality=50)),
           ref.encode(pic(16383, 9), ref.EncCfg(75, 2, lossless=1)),
           ref.encode(pic(9, 16383, True), ref.EncCfg(75, 2, lossless=1))]
    small = [ref.encode(ref.synth(48 + 16 * k, 40 + 8 * k, 7700 + k), ref.cfg_default(60 + 5 * k)) for k in range(6)]
    for csp in (W.MODE_RGBA, W.MODE_YUV, W.MODE_rgbA):
        wants = [ref.decode(d, csp, 0) for d in big + small]
        for d, (st_want, want) in zip(big, wants):       # one at a time
            st, out = W.WebPDecode(d, csp)
            assert st == st_want == 0, (csp, len(d), st)
            assert np.array_equal(out.reshape(-1), want.reshape(-1)), (csp, len(d))
        datas = [x for pair in zip(big, small) for x in pair] + big[len(small):]
        order = [(big + small).index(d) for d in datas]
        sts, outs = W.decode_batch(datas, csp)
        for i, j in enumerate(order):
            assert sts[i] == 0, (csp, i, sts[i])
            assert np.array_equal(outs[i].reshape(-1), wants[j][1].reshape(-1)), (csp, i)
    # many wide streams: the lockstep parser with 1024-macroblock context rows
    for d in (big[0], big[2]):
        st_want, want = ref.decode(d, W.MODE_RGBA, 0)
        sts, outs = W.decode_batch([d] * 300, W.MODE_RGBA)
        assert st_want == 0 and all(s == 0 for s in sts)
        for i in (0, 1, 149, 298, 299):
            assert np.array_equal(outs[i].reshape(-1), want.reshape(-1)), i


@pytest.mark.gpu
def test_damage_campaign_prefix(ref):
    """The first batches of tools/fuzz_gpu.py (seed 1: a prefix of the run logged in profiles/r01u_fuzz_gpu.log): mutated golden
    files, every eighth one intact, 2048 per WebPDecodeBatch call; per-item status and pixels against the compiled reference."""
    import os
    import subprocess
    import sys
    from conftest import ROOT
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fuzz_gpu.py"), "--seconds", "8", "--batch", "2048", "--seed", "1"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-1000:])


# ---------------------------------------------------------------------------------------------------------
# Round 2: BASELINE configs at their stated shapes, the asynchronous pair, in-library sharding, hygiene.

def _compare_all(W, ref, datas, csp, sts, outs, distinct=None):
    """Every image of a batch against the compiled reference; `distinct` = the batch tiles that many images."""
    cache = {}
    for i, data in enumerate(datas):
        key = i % distinct if distinct else i
        if key not in cache:
            cache[key] = ref.decode(data, csp, 0)
        s_ref, want = cache[key]
        assert sts[i] == s_ref, (i, sts[i], s_ref)
        if s_ref == 0:
            assert np.array_equal(outs[i].reshape(-1), want.reshape(-1)), (i, csp)


def test_config3_shape_8_partitions_normal_filter_yuv_and_rgba(W, ref):
    """BASELINE config 3: 1080p, 8 token partitions, 4 segments, normal loop filter, decoded to MODE_YUV and RGBA as one
    batch of 64 (16 distinct images tiled, every item compared)."""
    corpus = ref.encode_corpus(16, 1920, 1080, ref.cfg_normal_8part(), seed0=3100)
    datas = [corpus[i % 16] for i in range(64)]
    for csp in (W.MODE_YUV, W.MODE_RGBA):
        sts, outs = W.decode_batch(datas, csp)
        _compare_all(W, ref, datas, csp, sts, outs, distinct=16)


def test_config4_shape_4096_thumbnails_premultiplied(W, ref):
    """BASELINE config 4: 4096 thumbnails of 256x256 q80, fancy upsampling to premultiplied rgbA (64 distinct, all compared)."""
    corpus = ref.encode_corpus(64, 256, 256, ref.cfg_default(), seed0=4100)
    datas = [corpus[i % 64] for i in range(4096)]
    sts, outs = W.decode_batch(datas, W.MODE_rgbA)
    _compare_all(W, ref, datas, W.MODE_rgbA, sts, outs, distinct=64)


def test_many_small_images_take_the_banded_rows(W, ref):
    """More streams than the 4 KB-per-image layout of the probability rows can seat (16384 small pictures, one and four token
    partitions, twelve sizes): the launch picks the banded layout by itself (vp8_kernels.cu:launch_tokens_fp); every image compared."""
    cfgs = [ref.cfg_simple_1part(55), ref.cfg_default(), ref.EncCfg(70, 3, partitions=0, segments=2, filter_type=1),
            ref.EncCfg(45, 2, partitions=2, segments=4, filter_type=0)]   # 12288 one-partition streams (banded), 4096 x 4 partitions (not)
    corpus = [ref.encode(ref.synth(48 + 16 * (k % 3), 32 + 16 * (k % 4), 8200 + k), cfgs[k % 4]) for k in range(96)]
    datas = [corpus[i % 96] for i in range(16384)]
    sts, outs = W.decode_batch(datas, W.MODE_RGBA)
    _compare_all(W, ref, datas, W.MODE_RGBA, sts, outs, distinct=96)


def test_config5_shape_4096x4096_alpha(W, ref):
    """BASELINE config 5 at its stated shape: two 4096x4096 q90 images with a gradient-filtered ALPH chunk -> RGBA."""
    cfg = ref.cfg_alpha_q90()
    datas = ref.encode_corpus(2, 4096, 4096, cfg, seed0=5100, alpha=True)
    assert all(W.WebPGetFeatures(d)[1]["has_alpha"] for d in datas)
    sts, outs = W.decode_batch(datas, W.MODE_RGBA)
    _compare_all(W, ref, datas, W.MODE_RGBA, sts, outs)


def test_submit_wait_pipeline(W, ref, manifest):
    """WebPBatchSubmit / WebPBatchWait with two batches in flight on one device: same bytes as the blocking call, in any
    interleaving, also when one of the batches holds damaged files."""
    corpus = ref.encode_corpus(6, 640, 360, ref.cfg_simple_1part(), seed0=6100)
    d = manifest[0]["data"]
    a = [corpus[i % 6] for i in range(48)]
    b = [corpus[(i + 3) % 6] for i in range(40)] + [d[:len(d) // 2], b"junkjunkjunkjunk"]
    want = {id(x): W.decode_batch(x, W.MODE_RGBA) for x in (a, b)}
    slots = [W.Batch(a, W.MODE_RGBA, device=0), W.Batch(b, W.MODE_RGBA, device=0)]
    try:
        for rounds in range(3):
            for s in slots:
                s.submit()
            for s, datas in zip(slots, (a, b)):
                s.wait()
                sts, outs = want[id(datas)]
                assert s.statuses() == sts
                for i in range(s.n):
                    if sts[i] == 0:
                        assert np.array_equal(s.host_output(i), outs[i]), (rounds, i)
                s._out_arr[:] = 0
    finally:
        for s in slots:
            s.close()


def test_two_devices_in_one_process(W, ref):
    """WebPBatchOptions::devices: one call shards the items i % G over G devices (host thread per device, no collective);
    every image compared. Also plain per-device calls from one process (the per-device kernel attributes)."""
    if W.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    corpus = ref.encode_corpus(8, 1920, 1080, ref.cfg_simple_1part(), seed0=7100)
    datas = [corpus[i % 8] for i in range(1300)]   # enough streams per device for the lockstep parser (> 48 KB of shared memory)
    b = W.Batch(datas, W.MODE_RGBA, devices=[0, 1])
    try:
        assert b.decode_oneshot() == 0
        sts = b.statuses()
        outs = [b.host_output(i).copy() for i in range(b.n)]
    finally:
        b.close()
    _compare_all(W, ref, datas, W.MODE_RGBA, sts, outs, distinct=8)
    for dev in (1, 0):
        sts, outs = W.decode_batch(datas[:1200], W.MODE_RGBA, device=dev)
        _compare_all(W, ref, datas[:1200], W.MODE_RGBA, sts, outs, distinct=8)
    # resident form, sharded: device pointers of item i live on devices[i % G]
    r = W.Batch(datas[:64], W.MODE_RGBA, devices=[0, 1], output=W.WEBP_BATCH_DEVICE)
    try:
        assert r.create() == 0 and r.decode() == 0
        assert [r.device_output(i).device for i in range(4)] == [0, 1, 0, 1]
    finally:
        r.close()


def test_separately_allocated_buffers(W, ref, manifest):
    """One page-locked allocation per file and per output (merged copies used to span allocations), and plain
    pageable buffers."""
    datas = [e["data"] for e in manifest]
    want = [ref.decode(d, ref.MODE_RGBA, 0) for d in datas]
    L = W.lib()
    n = len(datas)
    items = (W.WebPBatchItem * n)()
    cfgs = (W.WebPDecoderConfig * n)()
    ins, outs = [], []
    try:
        for i, d in enumerate(datas):
            ib = W.HostBuffer(len(d)); ib.array[:len(d)] = np.frombuffer(d, np.uint8); ins.append(ib)
            _, f = W.WebPGetFeatures(d)
            ob = W.HostBuffer(f["width"] * f["height"] * 4); outs.append(ob)
            L.WebPInitDecoderConfigInternal(C.byref(cfgs[i]), W.WEBP_DECODER_ABI_VERSION)
            cfgs[i].output.colorspace = W.MODE_RGBA
            W._attach_external(cfgs[i], W.MODE_RGBA, f["width"], f["height"], ob.ptr, f["width"] * 4)
            items[i].data = ib.ptr; items[i].data_size = len(d); items[i].config = C.pointer(cfgs[i])
        assert L.WebPDecodeBatch(items, n, None) == 0, W.last_error()
        for i in range(n):
            _, f = W.WebPGetFeatures(datas[i])
            assert items[i].status == want[i][0] == 0
            assert np.array_equal(outs[i].array[:f["width"] * f["height"] * 4].reshape(f["height"], -1), want[i][1]), i
    finally:
        for b in ins + outs:
            b.free()
    sts, got = W.decode_batch(datas, W.MODE_RGBA, pinned=False)
    assert all(s == 0 and np.array_equal(g, w[1]) for s, g, w in zip(sts, got, want))


def test_header_only_giant_images_fail_alone(W, port, manifest):
    """Files of a few hundred bytes that declare 16383x16383 pixels: each gets the reference's status and the
    real images beside them decode."""
    good = [e["data"] for e in manifest[:6]]
    d = bytearray(next(m for m in manifest if m["file"] == "simple_1part_320x200.webp")["data"][:400])
    i = bytes(d).find(b"VP8 ")
    # 14-bit width / height fields of the frame header (vp8_dec.c:107-160): bytes 6..9 after the 3-byte frame tag
    d[i + 8 + 6] = 0xff; d[i + 8 + 7] = 0x3f; d[i + 8 + 8] = 0xff; d[i + 8 + 9] = 0x3f
    tag = int.from_bytes(d[i + 8:i + 11], "little")
    tag = (tag & 0x1f) | (100 << 5)   # a first partition that fits the 400 bytes, so that the file reaches the device planning
    d[i + 8:i + 11] = tag.to_bytes(3, "little")
    giant = bytes(d)
    st, f = W.WebPGetFeatures(giant)
    assert st == 0 and (f["width"], f["height"]) == (16383, 16383)
    datas = []
    for g in good:
        datas += [g, giant]
    # 90 giants declare 96 GB of RGBA between them: more than half the device, so the call runs in groups
    sts, outs = W.decode_batch(datas * 15, W.MODE_RGBA)
    for k, data in enumerate(datas * 15):
        if data is giant:
            assert sts[k] != 0
        else:
            want_st, want = port.decode(data, port.RGBA, 0)
            assert sts[k] == want_st == 0 and np.array_equal(outs[k], want), k


def test_cache_limit_and_trim(W, manifest):
    datas = [e["data"] for e in manifest]
    prev = W.set_cache_limit(1 << 30, 0)
    try:
        sts, _ = W.decode_batch(datas, W.MODE_RGBA, device=0)
        assert all(s == 0 for s in sts)
        assert W.trim_cache(0) > 0
        assert W.trim_cache(0) == 0
        sts, _ = W.decode_batch(datas, W.MODE_RGBA, device=0)   # works again after a trim
        assert all(s == 0 for s in sts)
    finally:
        W.set_cache_limit(prev, 0)


def test_callers_stream(W, ref, manifest):
    """WebPBatchOptions::stream: uploads and kernels are queued on the caller's stream -- behind work the caller queued
    before (here: a long fill kernel of torch's), the pixels complete when the call returns; device-resident output read
    back through the same stream. Every image compared."""
    torch = pytest.importorskip("torch")
    corpus = ref.encode_corpus(6, 640, 360, ref.cfg_simple_1part(), seed0=8100)
    datas = [corpus[i % 6] for i in range(60)] + [e["data"] for e in manifest[:4]]
    want = [ref.decode(d, ref.MODE_RGBA, 0) for d in datas]
    st = torch.cuda.Stream(device=0)
    with torch.cuda.stream(st):
        big = torch.empty(1 << 28, dtype=torch.uint8, device="cuda:0")
        for _ in range(8):
            big.fill_(7)          # something for the batch to queue behind
    b = W.Batch(datas, W.MODE_RGBA, device=0, stream=st.cuda_stream)
    try:
        assert b.decode_oneshot() == 0
        for i in range(b.n):
            assert b.statuses()[i] == want[i][0] == 0
            assert np.array_equal(b.host_output(i), want[i][1]), i
    finally:
        b.close()
    r = W.Batch(datas, W.MODE_RGBA, device=0, stream=st.cuda_stream, output=W.WEBP_BATCH_DEVICE)
    try:
        assert r.create() == 0 and r.decode() == 0
        for i in (0, 17, b.n - 1):
            p = r.device_output(i)
            w, h = r.dims[i]
            assert p.device == 0 and (p.width, p.height) == (w, h)
            t = torch.empty(h * p.stride, dtype=torch.uint8, device="cuda:0")
            # view the library's device memory as a tensor through the CUDA array interface
            class _V:
                __cuda_array_interface__ = {"shape": (h * p.stride,), "typestr": "|u1", "data": (p.y_or_rgba, False), "version": 3}
            with torch.cuda.stream(st):
                t.copy_(torch.as_tensor(_V(), device="cuda:0"))
            st.synchronize()
            assert np.array_equal(t.cpu().numpy().reshape(h, p.stride), want[i][1]), i
    finally:
        r.close()


def _damage_both_chunks(rng, d):
    """One byte of the ALPH payload changed and the VP8 payload damaged too (a changed byte, or the tail cut off)."""
    b = bytearray(d)
    i = d.find(b"ALPH")
    asz = int.from_bytes(d[i + 4:i + 8], "little")
    b[i + 8 + int(rng.integers(0, asz))] ^= int(rng.integers(1, 256))
    v = d.find(b"VP8 ", i + 8 + asz)
    vsz = int.from_bytes(d[v + 4:v + 8], "little")
    if rng.integers(0, 2):
        b[v + 8 + 10 + int(rng.integers(0, vsz - 10))] ^= int(rng.integers(1, 256))
    else:
        keep = int(rng.integers(vsz // 8, vsz))          # zero the tail: the chunk sizes stay, the partition runs dry
        b[v + 8 + keep:v + 8 + vsz] = bytes(vsz - keep)
    return bytes(b)


@pytest.mark.gpu
def test_alpha_and_vp8_both_damaged(W, ref, amanifest):
    """The reference decodes alpha rows as the macroblock rows above them finish and reports whichever failure its row loop
    meets first (frame_dec.c:440-460, alpha_dec.c:170-215); the product decodes the two chunks in separate kernels and
    rebuilds that order from the failing rows (vp8_dev.h:vp8b_vp8_failure_first)."""
    rng = np.random.default_rng(23)
    datas = []
    for e in amanifest:
        if len(e["data"]) < 400:
            continue
        for _ in range(400):
            datas.append(_damage_both_chunks(rng, e["data"]))
    sts, outs = W.decode_batch(datas, W.MODE_RGBA, device=0)
    seen = set()
    for d, st, out in zip(datas, sts, outs):
        s_ref, want = ref.decode(d, W.MODE_RGBA, 0)
        assert st == s_ref, (len(d), st, s_ref)
        seen.add(s_ref)
        if s_ref == 0 and not np.array_equal(out.reshape(-1), want.reshape(-1)):
            # out-of-range coefficients of a damaged stream: the reference's SSE2 transforms wrap at 16 bits where its C ones do
            # not (class 1); the C dsp path is the one to equal
            _, want = ref.decode(d, W.MODE_RGBA, 0, simd=False)
            assert np.array_equal(out.reshape(-1), want.reshape(-1))
    assert {3, 7} <= seen, seen      # both kinds of failure were met

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


@pytest.mark.gpu
def test_partition_starting_with_ff(W, ref, manifest, amanifest):
    """Files whose first or token partition starts with 0xFF go through k_parse_literal (the reference's reader taken literally)
    after the regular parse, in the same batch as intact files: status and pixels of the reference for all of them."""
    rng = np.random.default_rng(41)
    datas = []
    for e in list(manifest) + list(amanifest):
        data = e["data"]
        if data.find(b"VP8 ") < 0 or len(data) < 200:
            continue
        datas.append(data)
        datas += [b for b in _ff_mutants(rng, data, 24) if ref.features(b)[0] == 0]
    sts, outs = W.decode_batch(datas, W.MODE_RGBA, device=0)
    seen = {}
    for d, st, out in zip(datas, sts, outs):
        s_ref, want = ref.decode(d, W.MODE_RGBA, 0)
        assert st == s_ref, (len(d), st, s_ref)
        seen[s_ref] = seen.get(s_ref, 0) + 1
        if s_ref == 0 and not np.array_equal(out.reshape(-1), want.reshape(-1)):
            _, want = ref.decode(d, W.MODE_RGBA, 0, simd=False)
            assert np.array_equal(out.reshape(-1), want.reshape(-1))
    assert seen.get(0, 0) > 40 and seen.get(7, 0) > 20, seen


@pytest.mark.gpu
def test_crafted_vp8l_corners(W, ref, amanifest):
    """Hand-made VP8L streams (tests/vp8l_craft.py): group numbers beyond 1000 / beyond the pixel count, a colour-indexing transform
    that is not the first one; round 1 refused these (UNSUPPORTED_FEATURE), the reference decodes them."""
    import vp8l_craft
    host = [e["data"] for e in amanifest if e["file"] == "alpha_tiny_17x16.webp"][0]
    cases = vp8l_craft.crafted_cases(host)
    for csp in (W.MODE_RGBA, W.MODE_rgbA, W.MODE_YUV):
        sts, outs = W.decode_batch([d for _, d in cases], csp, device=0)
        for (name, data), st, out in zip(cases, sts, outs):
            s_ref, want = ref.decode(data, csp, 0)
            assert st == s_ref == 0, (name, csp, st, s_ref)
            assert np.array_equal(out.reshape(-1)[:want.size], want.reshape(-1)), (name, csp)


@pytest.mark.gpu
def test_unusable_slow_memory_buffer_is_reported_after_the_decode(W, ref, amanifest):
    """is_external_memory = 2 + premultiplied output + a file with alpha: the reference decodes into a buffer of its own and looks
    at the caller's only when it copies (webp_dec.c:769-786). So a caller's buffer that is too small ends in INVALID_PARAM when
    the file is sound, and in the file's own failure when it is not -- the product decodes (and copies nothing) to find out."""
    import ctypes as C
    L = W.lib()
    Q = ref.lib()
    for L_ in (L, Q):
        L_.WebPInitDecoderConfigInternal.argtypes = [C.POINTER(W.WebPDecoderConfig), C.c_int]
        L_.WebPDecode.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(W.WebPDecoderConfig)]
    e = [x for x in amanifest if x["file"] == "alpha_gradient_130x97.webp"][0]
    w, h = e["features"]["width"], e["features"]["height"]
    sound = e["data"]
    cut = sound[: len(sound) - 600]                       # the VP8 payload runs dry
    i = sound.find(b"ALPH")
    bad_alpha = bytearray(sound); bad_alpha[i + 9] ^= 0x55; bad_alpha = bytes(bad_alpha)
    arena = np.zeros(4 * w * h + 64, np.uint8)
    seen = set()
    for data in (sound, cut, bad_alpha):
        for stride, size in ((4 * w - 8, 4 * w * h), (4 * w, 4 * w * h - 5), (4 * w, 4 * w * h)):
            got = []
            for lib in (Q, L):
                cfg = W.WebPDecoderConfig()
                lib.WebPInitDecoderConfigInternal(C.byref(cfg), W.WEBP_DECODER_ABI_VERSION)
                cfg.output.colorspace = W.MODE_rgbA
                cfg.output.is_external_memory = 2
                cfg.output.u.RGBA.rgba = arena.ctypes.data
                cfg.output.u.RGBA.stride = stride
                cfg.output.u.RGBA.size = size
                arena[:] = 0
                st = lib.WebPDecode(data, len(data), C.byref(cfg))
                got.append((st, arena.copy() if st == 0 else None))
            assert got[0][0] == got[1][0], (len(data), stride, size, got[0][0], got[1][0])
            seen.add(got[0][0])
            if got[0][0] == 0:
                assert np.array_equal(got[0][1], got[1][1])
    assert {0, 2, 7} <= seen, seen


@pytest.mark.gpu
def test_parse_stages_match_the_host_build(W, manifest, amanifest):
    """Stage level, not pixels: what K1 and the token parser leave on the device (MbInfo: intra modes, non-zero codes, flags; the
    coefficient levels of every block) against the same two stages of the device code's host build (tests/emu, itself pinned to
    the oracle's stage dump by tests/test_emu.py). A regression in either parse shows here before it shows as wrong pixels."""
    import ctypes as C
    import subprocess
    from conftest import ROOT
    emu_dir = os.path.join(ROOT, "tests", "emu")
    subprocess.check_call(["make", "-s", "-C", emu_dir])
    E = C.CDLL(os.path.join(emu_dir, "libvp8_emu.so"))
    E.emu_decode.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p]
    E.emu_set_stage_dump.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    entries = [e for e in list(manifest) + list(amanifest) if e["data"].find(b"VP8 ") >= 0]
    datas = [e["data"] for e in entries]
    b = W.Batch(datas, W.MODE_RGBA, device=0, output=W.WEBP_BATCH_DEVICE)
    assert b.create() == 0 and b.decode() == 0, W.last_error()
    checked = 0
    for i, e in enumerate(entries):
        w, h = e["features"]["width"], e["features"]["height"]
        nmb = ((w + 15) // 16) * ((h + 15) // 16)
        mi_d = np.zeros((nmb, 4), np.uint32); lv_d = np.zeros((nmb, 400), np.int16)
        n = W.lib().WebPBatchDebugStages(b.handle, i, mi_d.ctypes.data, lv_d.ctypes.data, nmb)
        assert n == nmb, (e["file"], n)
        mi_e = np.zeros((nmb, 4), np.uint32); lv_e = np.zeros((nmb, 400), np.int16)
        E.emu_set_stage_dump(mi_e.ctypes.data, lv_e.ctypes.data, nmb)
        out = np.zeros((h, w * 4), np.uint8)
        st = E.emu_decode(e["data"], len(e["data"]), 1, 0, out.ctypes.data, out.size, w * 4, 64, None)
        E.emu_set_stage_dump(None, None, 0)
        assert st == 0, e["file"]
        assert np.array_equal(mi_d[:, :3], mi_e[:, :3]), e["file"]                       # modes, non-zero codes
        keep = np.uint32(~((1 << 23) | (1 << 24)) & 0xffffffff)                          # MBW_INNER / MBW_DITHER are K3's and the dither plan's
        assert np.array_equal(mi_d[:, 3] & keep, mi_e[:, 3] & keep), e["file"]
        assert np.array_equal(lv_d, lv_e), e["file"]
        checked += int(np.count_nonzero(lv_e))
    b.close()
    assert checked > 100000
