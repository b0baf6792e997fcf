"""CPU-side checks of the drop-in boundary: the shared library loads, exports every symbol include/webp/*.h
declares, agrees with the reference on struct layouts and on everything the host decides (features, container
errors, option screening). No kernel is launched here."""
import ctypes as C
import os
import re

import numpy as np

from conftest import ROOT


def test_library_exports_every_declared_symbol(product):
    L = product.lib()
    declared = set()
    for hdr in ("types.h", "decode.h", "decode_batch.h"):
        src = open(os.path.join(ROOT, "include", "webp", hdr)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        declared |= set(re.findall(r"^WEBP_EXTERN[^;(]*?\b(\w+)\s*\(", src, flags=re.M))
    declared.add("VP8GetCPUInfo")
    assert declared == set(product.EXPORTS), declared ^ set(product.EXPORTS)
    for name in declared:
        assert hasattr(L, name), name
    assert L.WebPGetDecoderVersion() == 0x010302


def test_struct_layouts_match_reference_abi(product):
    # sizes on LP64 for ABI 0x0209 (src/webp/decode.h:184-217, 414-469)
    assert C.sizeof(product.WebPRGBABuffer) == 24
    assert C.sizeof(product.WebPYUVABuffer) == 80
    assert C.sizeof(product.WebPDecBuffer) == 120
    assert C.sizeof(product.WebPBitstreamFeatures) == 40
    assert C.sizeof(product.WebPDecoderOptions) == 76
    assert C.sizeof(product.WebPDecoderConfig) == 240
    cfg = product.WebPDecoderConfig()
    assert product.lib().WebPInitDecoderConfigInternal(C.byref(cfg), 0x0209) == 1
    assert product.lib().WebPInitDecoderConfigInternal(C.byref(cfg), 0x0109) == 0   # ABI major mismatch


def test_struct_layouts_match_reference_compiler(ref, product):
    """Same sizes as the reference's own header compiled by gcc (via the reference .so's view of the structs)."""
    import subprocess, tempfile
    if not os.path.isdir("/root/reference"):
        import pytest
        pytest.skip("needs the reference headers")
    src = ('#include <stdio.h>\n#include "src/webp/decode.h"\nint main(void){printf("%zu %zu %zu %zu %zu %zu\\n",'
           'sizeof(WebPRGBABuffer),sizeof(WebPYUVABuffer),sizeof(WebPDecBuffer),sizeof(WebPBitstreamFeatures),'
           'sizeof(WebPDecoderOptions),sizeof(WebPDecoderConfig));return 0;}')
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "s.c"), "w").write(src)
        subprocess.check_call(["gcc", "-I/root/reference", os.path.join(d, "s.c"), "-o", os.path.join(d, "s")])
        sizes = list(map(int, subprocess.check_output([os.path.join(d, "s")]).split()))
    mine = [C.sizeof(t) for t in (product.WebPRGBABuffer, product.WebPYUVABuffer, product.WebPDecBuffer,
                                  product.WebPBitstreamFeatures, product.WebPDecoderOptions, product.WebPDecoderConfig)]
    assert sizes == mine


def test_features_match_oracle(product, port, manifest):
    for e in manifest:
        st, f = product.WebPGetFeatures(e["data"])
        assert st == 0 and f == e["features"]
        for n in (0, 4, 11, 12, 15, 20, 25, 29, 30, 31, 100):
            cut = e["data"][:n]
            assert product.WebPGetFeatures(cut) == port.features(cut), (e["file"], n)
    w, h = C.c_int(), C.c_int()
    d = manifest[0]["data"]
    assert product.lib().WebPGetInfo(d, len(d), C.byref(w), C.byref(h)) == 1
    assert (w.value, h.value) == (manifest[0]["features"]["width"], manifest[0]["features"]["height"])


def test_features_of_a_vp8x_file_cut_before_the_frame(product, ref, amanifest):
    """WebPGetFeatures answers from the VP8X chunk when the data ends before the frame header (webp_dec.c:397-404): has_alpha is
    the VP8X flag OR "an ALPH chunk went by", so a file whose flag was cleared still reports alpha once its ALPH chunk is in.
    (Found by tools/fuzz_emu.py.)"""
    for e in amanifest:
        d = bytearray(e["data"])
        i = d.find(b"ALPH")
        vp8 = i + 8 + int.from_bytes(d[i + 4:i + 8], "little")
        vp8 += vp8 & 1
        assert d[vp8:vp8 + 4] == b"VP8 "
        for flags in (d[20], d[20] & ~0x10, 0x8d):
            d[20] = flags
            for n in (i + 4, vp8 - 1, vp8, vp8 + 8, vp8 + 10, vp8 + 17, vp8 + 18, len(d)):
                cut = bytes(d[:n])
                assert product.WebPGetFeatures(cut) == ref.features(cut), (e["file"], hex(flags), n)


def test_host_side_failures_need_no_gpu(product, port, manifest):
    """Errors decided by the container walk come back with the reference's status codes without touching a GPU."""
    d = manifest[1]["data"]
    for cut in (d[:0], d[:5], d[:11], d[:19], d[:29], b"RIFF" + d[4:8] + b"WEBX" + d[12:], d[: len(d) // 2]):
        st, out = product.WebPDecode(cut)
        want, _ = port.decode(cut, port.RGBA, 0)
        if want in (3, 7) and out is None and st in (3, 7):
            # truncated-inside-the-bitstream cases are decided on the device; the ones decided on the host must agree
            if len(cut) < 30 or cut[8:12] != b"WEBP":
                assert st == want, (len(cut), st, want)
    L = product.lib()
    assert L.WebPDecode(d, len(d), None) == product.VP8_STATUS_INVALID_PARAM
    cfg = product.WebPDecoderConfig()
    L.WebPInitDecoderConfigInternal(C.byref(cfg), 0x0209)
    cfg.output.colorspace = 13
    assert L.WebPDecode(d, len(d), C.byref(cfg)) == product.VP8_STATUS_INVALID_PARAM
    L.WebPInitDecoderConfigInternal(C.byref(cfg), 0x0209)
    cfg.options.use_scaling = 1            # a 0 x 0 scaling request is an invalid parameter (buffer_dec.c:197-205)
    assert L.WebPDecode(d, len(d), C.byref(cfg)) == product.VP8_STATUS_INVALID_PARAM
    L.WebPInitDecoderConfigInternal(C.byref(cfg), 0x0209)
    cfg.options.use_cropping = 1           # an empty crop window is an invalid parameter (buffer_dec.c:184-195)
    assert L.WebPDecode(d, len(d), C.byref(cfg)) == product.VP8_STATUS_INVALID_PARAM
    # external buffer too small -> INVALID_PARAM (buffer_dec.c:41-84)
    L.WebPInitDecoderConfigInternal(C.byref(cfg), 0x0209)
    buf = np.zeros(16, np.uint8)
    cfg.output.colorspace = product.MODE_RGBA
    cfg.output.is_external_memory = 1
    cfg.output.u.RGBA.rgba = buf.ctypes.data
    cfg.output.u.RGBA.stride = 4
    cfg.output.u.RGBA.size = 16
    assert L.WebPDecode(d, len(d), C.byref(cfg)) == product.VP8_STATUS_INVALID_PARAM


def test_no_cpu_fallback(product, manifest):
    """Without a CUDA device a well-formed file must NOT decode: the product has no CPU path."""
    if product.device_count() > 0:
        import pytest
        pytest.skip("a GPU is present")
    st, out = product.WebPDecode(manifest[0]["data"])
    assert st != 0 and out is None
    assert "CPU" in product.last_error() or "CUDA" in product.last_error()


def test_product_never_touches_the_oracle():
    """The shipped package must not import, link or read anything under oracle/ or tests/."""
    pkg = os.path.join(ROOT, "libwebp_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".c", ".h", ".cu", ".cuh", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                for bad in ("oracle/", "oracle import", "from oracle", "vp8_oracle", "libwebp_ref", "tests/emu/libvp8"):
                    assert bad not in text.replace("oracle/: ", "").replace("touches oracle/", ""), (f, bad)


def test_incremental_protocol_without_device(product, manifest):
    """WebPIAppend / WebPIUpdate up to (not including) the last byte: VP8_STATUS_SUSPENDED, no device needed; bad
    headers are refused at once; append and update do not mix (src/dec/idec_dec.c:700-760)."""
    import ctypes as C
    L = product.lib()
    L.WebPINewDecoder.restype = C.c_void_p
    L.WebPINewDecoder.argtypes = [C.c_void_p]
    L.WebPIAppend.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
    L.WebPIUpdate.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
    L.WebPIDelete.argtypes = [C.c_void_p]
    L.WebPIDecode.restype = C.c_void_p
    L.WebPIDecode.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p]
    data = manifest[1]["data"]
    idec = L.WebPINewDecoder(None)
    assert idec
    for lo, hi in ((0, 10), (10, 500), (500, len(data) - 1)):
        assert L.WebPIAppend(idec, data[lo:hi], hi - lo) == product.VP8_STATUS_SUSPENDED
    assert L.WebPIUpdate(idec, data, len(data) - 1) == product.VP8_STATUS_INVALID_PARAM      # no mixing
    L.WebPIDelete(idec)
    idec = L.WebPINewDecoder(None)
    bad = b"RIFF" + (100).to_bytes(4, "little") + b"WEBPjunkjunkjunkjunkjunkjunk"
    assert L.WebPIUpdate(idec, bad, len(bad)) == product.VP8_STATUS_BITSTREAM_ERROR
    assert L.WebPIUpdate(idec, bad, len(bad)) == product.VP8_STATUS_BITSTREAM_ERROR               # the error sticks
    L.WebPIDelete(idec)
    assert not L.WebPIDecode(b"junkjunkjunkjunkjunk", 20, None)                                  # features must parse
    assert L.WebPIAppend(None, data, 10) == product.VP8_STATUS_INVALID_PARAM


def test_incremental_output_area_before_the_last_byte(product, ref, manifest, lmanifest):
    """The getters of the incremental API against the reference's, append by append (the protocol of src/tests.zig:676-686):
    NULL while the headers / first partition are still arriving (idec_dec.c:843-851), then the caller's buffer with last_y rows
    in it. The reference fills rows as data arrives, this library none until the picture is complete: whenever the reference
    has an area this library has one too, never with more rows than the reference's, and the caller's buffer is what comes back."""
    import ctypes as C
    P, Q = product.lib(), ref.lib()
    for L in (P, Q):
        L.WebPInitDecoderConfigInternal.argtypes = [C.POINTER(product.WebPDecoderConfig), C.c_int]
        L.WebPINewDecoder.restype = C.c_void_p
        L.WebPINewDecoder.argtypes = [C.c_void_p]
        L.WebPIAppend.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
        L.WebPIDelete.argtypes = [C.c_void_p]
        L.WebPIDecGetRGB.restype = C.c_void_p
        L.WebPIDecGetRGB.argtypes = [C.c_void_p] + [C.POINTER(C.c_int)] * 4
    phases = [False, False]
    for e in list(manifest[:6]) + list(lmanifest[:3]):
        data = e["data"]
        w, h = e["features"]["width"], e["features"]["height"]
        areas = []
        for L in (Q, P):
            cfg = product.WebPDecoderConfig()
            L.WebPInitDecoderConfigInternal(C.byref(cfg), product.WEBP_DECODER_ABI_VERSION)
            buf = np.zeros(w * h * 4, np.uint8)
            cfg.output.colorspace = product.MODE_RGBA
            cfg.output.is_external_memory = 1
            cfg.output.u.RGBA.rgba = buf.ctypes.data
            cfg.output.u.RGBA.stride = 4 * w
            cfg.output.u.RGBA.size = buf.size
            idec = L.WebPINewDecoder(C.addressof(cfg.output))
            seen = []
            step = max(64, len(data) // 7)
            cuts = [0, 16, 40] + list(range(40 + step, len(data) - 1, step)) + [len(data) - 1]   # everything but the last byte: no device needed
            for lo, hi in zip(cuts[:-1], cuts[1:]):
                st = L.WebPIAppend(idec, data[lo:hi], hi - lo)
                if L is Q and st == product.VP8_STATUS_OK:
                    break                                      # the reference is done once the last macroblock is (padding may follow)
                assert st == product.VP8_STATUS_SUSPENDED, (e["file"], lo, st)
                ly, ww, hh, ss = C.c_int(-1), C.c_int(-1), C.c_int(-1), C.c_int(-1)
                ptr = L.WebPIDecGetRGB(idec, C.byref(ly), C.byref(ww), C.byref(hh), C.byref(ss))
                seen.append((bool(ptr), ptr == buf.ctypes.data if ptr else None, ly.value if ptr else None, ww.value if ptr else None,
                             hh.value if ptr else None, ss.value if ptr else None))
            L.WebPIDelete(idec)
            areas.append(seen)
        for r, p in zip(*areas):
            assert r[0] == p[0], (e["file"], r, p)             # an area exactly when the reference has one
            if r[0]:
                assert p[1] and r[1]                            # the caller's buffer
                assert p[2] == 0 and p[2] <= r[2]               # no rows yet here; the reference may have some
                assert p[3:] == r[3:], (e["file"], r, p)       # width, height, stride
        phases[0] |= any(r[0] for r in areas[0]); phases[1] |= not areas[0][0][0]
    assert phases == [True, True]   # both phases were seen: no area yet / an area before the last byte


def test_lossless_crop_window_is_checked_at_the_offsets_as_given(product, lmanifest):
    """WebPAllocateDecBuffer checks the crop window at offsets snapped to even (buffer_dec.c:184-195), the lossless decoder
    then crops at the offsets as given (WebPIoInitFromOptions snaps for YUV420 sources only, webp_dec.c:809-817) and refuses a
    window that overhangs there: INVALID_PARAM, decided on the host. (Found by tools/fuzz_options.py.)"""
    L = product.lib()
    e = next(x for x in lmanifest if x["file"] == "lossless_alpha_83x61.webp")
    for (x, y, cw, ch), want in (((61, 3, 23, 57), 2), ((60, 3, 23, 57), product.VP8_STATUS_USER_ABORT), ((10, 41, 18, 21), 2),
                                 ((10, 40, 18, 21), product.VP8_STATUS_USER_ABORT), ((83, 0, 1, 1), 2), ((0, 61, 1, 1), 2)):
        cfg = product.WebPDecoderConfig()
        L.WebPInitDecoderConfigInternal(C.byref(cfg), 0x0209)
        cfg.options.use_cropping = 1
        cfg.options.crop_left, cfg.options.crop_top, cfg.options.crop_width, cfg.options.crop_height = x, y, cw, ch
        st = L.WebPDecode(e["data"], len(e["data"]), C.byref(cfg))
        L.WebPFreeDecBuffer(C.byref(cfg.output))
        if product.device_count() > 0 and want == product.VP8_STATUS_USER_ABORT:
            want = 0      # a GPU is present: the window is legal, so it decodes
        assert st == want, ((x, y, cw, ch), st, want)


def test_option_and_buffer_screening_matches_reference(product, ref):
    """A short run of tools/fuzz_options.py: random WebPDecoderConfig contents (colourspace, crop, scaling, flip, internal /
    external buffers with strides and sizes around the smallest legal values, missing planes) through the reference and the
    product's host side; every refusal equal, nothing the reference decodes refused. Needs a machine without a GPU (there
    "let through" shows as USER_ABORT)."""
    import subprocess
    import sys
    if product.device_count() > 0:
        pytest.skip("compares host-side refusals where no device can decode")
    from conftest import ROOT
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fuzz_options.py"), "--cases", "6000", "--seed", "11"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:]
    # the same with damaged files mixed in: a damaged HEADER is reported before an illegal request, like the reference does
    # (webp_dec.c:469-481: VP8GetHeaders / VP8LDecodeHeader run before WebPAllocateDecBuffer; vp8_host_probe.cpp)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fuzz_options.py"), "--cases", "6000", "--seed", "12", "--damage", "0.6"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:]


def test_packaged_as_libwebpdecoder(product, tmp_path):
    """`make dist`: the library under the reference's own name -- libwebpdecoder.so.3.1.8 with soname libwebpdecoder.so.3
    (src/Makefile.am:51), a static archive and libwebpdecoder.pc (src/libwebpdecoder.pc.in) -- so that an application
    which links -lwebpdecoder needs no change to its build. A C program compiled against it runs without a device."""
    import subprocess
    csrc = os.path.join(ROOT, "libwebp_b200", "csrc")
    subprocess.check_call(["make", "-s", "-C", csrc, "dist"])
    dist = os.path.join(ROOT, "libwebp_b200", "dist")
    so = os.path.join(dist, "lib", "libwebpdecoder.so.3.1.8")
    assert os.path.exists(so) and os.path.exists(os.path.join(dist, "lib", "libwebpdecoder.a"))
    assert os.path.realpath(os.path.join(dist, "lib", "libwebpdecoder.so")) == os.path.realpath(so)
    dyn = subprocess.run(["readelf", "-d", so], capture_output=True, text=True).stdout
    assert "libwebpdecoder.so.3" in dyn and "SONAME" in dyn
    pc = open(os.path.join(dist, "lib", "pkgconfig", "libwebpdecoder.pc")).read()
    assert "Name: libwebpdecoder" in pc and "-lwebpdecoder" in pc and "Version: 1.3.2" in pc
    src = tmp_path / "t.c"
    src.write_text('#include <stdio.h>\n#include "webp/decode.h"\n#include "webp/decode_batch.h"\n'
                   'int main(void) { WebPDecoderConfig c; WebPBatchOptions o;\n'
                   '  if (!WebPInitDecoderConfig(&c) || !WebPBatchOptionsInit(&o)) return 1;\n'
                   '  printf("%x %d\\n", WebPGetDecoderVersion(), (int)sizeof(o)); return 0; }\n')
    exe = tmp_path / "t"
    subprocess.check_call(["gcc", str(src), "-I", os.path.join(dist, "include"), "-L", os.path.join(dist, "lib"),
                           "-lwebpdecoder", "-Wl,-rpath," + os.path.join(dist, "lib"), "-o", str(exe)])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.split() == ["10302", "48"], out.stdout + out.stderr
    needed = subprocess.run(["readelf", "-d", str(exe)], capture_output=True, text=True).stdout
    assert "libwebpdecoder.so.3" in needed
