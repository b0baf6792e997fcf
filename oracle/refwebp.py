"""ctypes binding of oracle/_ref/libwebp_ref.so (the unmodified reference + oracle/reftool.c helpers).

TEST INFRASTRUCTURE ONLY: imported by tests/, bench.py (cpu_baseline and --impl reference legs, corpus
generation) and __graft_entry__.smoke(). The product package libwebp_b200 never imports this module.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libwebp_ref.so")

# WEBP_CSP_MODE (src/webp/decode.h:150-163)
MODE_RGB, MODE_RGBA, MODE_BGR, MODE_BGRA, MODE_ARGB = 0, 1, 2, 3, 4
MODE_rgbA, MODE_bgrA, MODE_Argb = 7, 8, 9
MODE_YUV, MODE_YUVA = 11, 12
BPP = {0: 3, 1: 4, 2: 3, 3: 4, 4: 4, 5: 2, 6: 2, 7: 4, 8: 4, 9: 4, 10: 2}
MODE_RGBA_4444, MODE_RGB_565, MODE_rgbA_4444 = 5, 6, 10

FLAG_BYPASS_FILTER, FLAG_NO_FANCY, FLAG_THREADS, FLAG_FLIP = 1, 2, 4, 8


class EncCfg(C.Structure):
    """Mirror of ReftEncCfg (oracle/reftool.c); -1 keeps the encoder default."""
    _fields_ = [("quality", C.c_float), ("method", C.c_int), ("segments", C.c_int), ("filter_type", C.c_int),
                ("filter_strength", C.c_int), ("filter_sharpness", C.c_int), ("partitions", C.c_int),
                ("low_memory", C.c_int), ("alpha_filtering", C.c_int), ("alpha_quality", C.c_int),
                ("sns_strength", C.c_int), ("lossless", C.c_int)]

    def __init__(self, quality=75.0, method=4, **kw):
        super().__init__()
        for name, _ in self._fields_:
            setattr(self, name, -1)
        self.quality, self.method = quality, method
        for k, v in kw.items():
            setattr(self, k, v)


# The corpus configurations of BASELINE.json / SURVEY.md 8(d).
def cfg_simple_1part(quality=75.0):      # config 2: 1 segment, simple filter, 1 token partition
    return EncCfg(quality, 4, segments=1, filter_type=0, partitions=0)


def cfg_normal_8part(quality=75.0):      # config 3: 4 segments, normal filter, 8 token partitions
    return EncCfg(quality, 4, segments=4, filter_type=1, partitions=3, low_memory=1)


def cfg_default(quality=80.0):           # config 4: cwebp defaults at q80 (4 segments, strong filter)
    return EncCfg(quality, 4)


def cfg_alpha_q90():                     # config 5: q90 with an ALPH chunk, gradient-filtered (alpha_filter best)
    return EncCfg(90.0, 4, alpha_filtering=2)


def cfg_lossless(quality=75.0):          # whole-picture VP8L (not a BASELINE config)
    return EncCfg(quality, 4, lossless=1)


_lib = None


def available():
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError(f"{LIB_PATH} missing: run `make -C oracle ref` where /root/reference exists")
        L = C.CDLL(LIB_PATH)
        L.reft_synth.argtypes = [C.c_int, C.c_int, C.c_uint32, C.c_int, C.c_void_p]
        L.reft_encode.restype = C.c_size_t
        L.reft_encode.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(EncCfg), C.POINTER(C.c_void_p)]
        L.reft_free.argtypes = [C.c_void_p]
        L.reft_encode_corpus.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint32, C.POINTER(EncCfg), C.c_int,
                                         C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
        L.reft_decode.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_int]
        L.reft_features.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.c_int)]
        if hasattr(L, "reft_decode_window"):
            L.reft_decode_window.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_int, C.POINTER(C.c_int), C.c_void_p, C.c_size_t]
        if hasattr(L, "reft_decode_dithered"):
            L.reft_decode_dithered.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_int, C.POINTER(C.c_int), C.c_int, C.c_int,
                                               C.c_void_p, C.c_size_t]
        if hasattr(L, "reft_decode_scaled"):
            L.reft_decode_scaled.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                             C.POINTER(C.c_int), C.c_void_p, C.c_size_t]
        L.reft_decode_bench.restype = C.c_double
        L.reft_decode_bench.argtypes = [C.POINTER(C.c_char_p), C.POINTER(C.c_size_t), C.c_int, C.c_int, C.c_int,
                                        C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int)]
        L.reft_set_simd.argtypes = [C.c_int]
        _lib = L
    return _lib


def synth(w, h, seed, alpha=False):
    bpp = 4 if alpha else 3
    out = np.empty((h, w, bpp), np.uint8)
    lib().reft_synth(w, h, seed, bpp, out.ctypes.data)
    return out


def encode(pix, cfg):
    h, w, bpp = pix.shape
    pix = np.ascontiguousarray(pix)
    p = C.c_void_p()
    n = lib().reft_encode(pix.ctypes.data, w, h, bpp, C.byref(cfg), C.byref(p))
    if n == 0:
        raise RuntimeError("reference encoder failed")
    data = C.string_at(p, n)
    lib().reft_free(p)
    return data


def encode_corpus(n, w, h, cfg, seed0=1, alpha=False, nthreads=None):
    """n synthetic images (seed0+k) -> list of .webp byte strings, encoded on nthreads host threads."""
    nthreads = nthreads or (os.cpu_count() or 1)
    outs = (C.c_void_p * n)()
    sizes = (C.c_size_t * n)()
    rc = lib().reft_encode_corpus(n, w, h, 4 if alpha else 3, seed0, C.byref(cfg), nthreads, outs, sizes)
    res = []
    for i in range(n):
        if outs[i]:
            res.append(C.string_at(outs[i], sizes[i]))
            lib().reft_free(outs[i])
    if rc != 0 or len(res) != n:
        raise RuntimeError("reference encoder failed on the corpus")
    return res


def features(data):
    f = (C.c_int * 5)()
    st = lib().reft_features(data, len(data), f)
    return st, dict(width=f[0], height=f[1], has_alpha=f[2], has_animation=f[3], format=f[4])


def decode(data, csp=MODE_RGBA, flags=0, simd=True, stride=None):
    """Reference WebPDecode. Returns (status, ndarray): (h, stride) bytes for RGB modes (visible part is
    [:, :w*bpp]); for MODE_YUV a flat y|u|v array with tight strides."""
    L = lib()
    L.reft_set_simd(1 if simd else 0)
    st, f = features(data)
    if st != 0:
        dummy = np.zeros(16, np.uint8)
        st = L.reft_decode(data, len(data), csp, flags, dummy.ctypes.data, 16, 16)
        L.reft_set_simd(1)
        return st, None
    w, h = f["width"], f["height"]
    if csp in (MODE_YUV, MODE_YUVA):
        n = w * h + 2 * ((w + 1) // 2) * ((h + 1) // 2) + (w * h if csp == MODE_YUVA else 0)
        out = np.zeros(n, np.uint8)
        st = L.reft_decode(data, len(data), csp, flags, out.ctypes.data, n, 0)
    else:
        stride = stride or w * BPP[csp]
        out = np.zeros((h, stride), np.uint8)
        st = L.reft_decode(data, len(data), csp, flags, out.ctypes.data, out.size, stride)
    L.reft_set_simd(1)
    return st, (out if st == 0 else None)


def decode_window(data, csp=MODE_RGBA, flags=0, crop=None):
    """Reference WebPDecode with options.use_cropping (crop = (left, top, width, height)) and/or FLAG_FLIP, into a
    tight buffer of the output size. Returns (status, flat ndarray or None)."""
    L = lib()
    st, f = features(data)
    w, h = (crop[2], crop[3]) if crop else (f["width"], f["height"])
    if st != 0 or w <= 0 or h <= 0:
        w = h = 4
    n = (w * h + 2 * ((w + 1) // 2) * ((h + 1) // 2) + (w * h if csp == MODE_YUVA else 0)) if csp in (MODE_YUV, MODE_YUVA) \
        else w * h * BPP[csp]
    out = np.zeros(max(n, 16), np.uint8)
    c4 = (C.c_int * 4)(*(crop if crop else (0, 0, 0, 0)))
    st = L.reft_decode_window(data, len(data), csp, flags, c4, out.ctypes.data, out.size)
    return st, (out[:n] if st == 0 else None)


def decode_dithered(data, csp=MODE_RGBA, flags=0, crop=None, strength=50, alpha_strength=0):
    """decode_window with options.dithering_strength / options.alpha_dithering_strength (dwebp's defaults are 50 / 100)."""
    L = lib()
    st, f = features(data)
    w, h = (crop[2], crop[3]) if crop else (f["width"], f["height"])
    if st != 0 or w <= 0 or h <= 0:
        w = h = 4
    n = (w * h + 2 * ((w + 1) // 2) * ((h + 1) // 2) + (w * h if csp == MODE_YUVA else 0)) if csp in (MODE_YUV, MODE_YUVA) \
        else w * h * BPP[csp]
    out = np.zeros(max(n, 16), np.uint8)
    c4 = (C.c_int * 4)(*(crop if crop else (0, 0, 0, 0)))
    st = L.reft_decode_dithered(data, len(data), csp, flags, c4, strength, alpha_strength, out.ctypes.data, out.size)
    return st, (out[:n] if st == 0 else None)


def decode_scaled(data, csp=MODE_RGBA, flags=0, crop=None, scaled=(0, 0)):
    """Reference WebPDecode with options.use_scaling (scaled = (width, height), 0 = keep the ratio), optionally on a crop
    window. Returns (status, (scaled_w, scaled_h), flat ndarray or None)."""
    L = lib()
    c4 = (C.c_int * 4)(*(crop if crop else (0, 0, 0, 0)))
    s2 = (C.c_int * 2)(*scaled)
    d2 = (C.c_int * 2)(0, 0)
    st0, f = features(data)
    w, h = (crop[2], crop[3]) if crop else (f["width"], f["height"])
    sw, sh = scaled
    if sw == 0 and h > 0:
        sw = (w * sh + h - 1) // h
    if sh == 0 and w > 0:
        sh = (h * sw + w - 1) // w
    sw, sh = max(sw, 1), max(sh, 1)
    n = (sw * sh + 2 * ((sw + 1) // 2) * ((sh + 1) // 2) + (sw * sh if csp == MODE_YUVA else 0)) if csp in (MODE_YUV, MODE_YUVA) \
        else sw * sh * BPP[csp]
    out = np.zeros(max(n, 16), np.uint8)
    st = L.reft_decode_scaled(data, len(data), csp, flags, c4, s2, d2, out.ctypes.data, out.size)
    return st, (d2[0], d2[1]), (out[:n] if st == 0 else None)


def decode_bench(datas, nthreads, csp=MODE_RGBA, passes=1, simd=True):
    """The CPU baseline of BASELINE.md section 3. Returns dict(seconds, mpix_per_pass, mpix_s, errors)."""
    L = lib()
    L.reft_set_simd(1 if simd else 0)
    n = len(datas)
    arr = (C.c_char_p * n)(*datas)
    sizes = (C.c_size_t * n)(*[len(d) for d in datas])
    mpix, errs = C.c_double(), C.c_int()
    sec = L.reft_decode_bench(arr, sizes, n, nthreads, csp, passes, C.byref(mpix), C.byref(errs))
    L.reft_set_simd(1)
    return dict(seconds=sec, mpix_per_pass=mpix.value, mpix_s=mpix.value * passes / max(sec, 1e-12),
                errors=errs.value, threads=nthreads)


# ---------------------------------------------------------------------------------------------------------
# Animated WebP (oracle/animtool.c): the reference's WebPAnimEncoder / WebPAnimDecoder, and the same WebPAnimDecoder objects
# relinked over the CUDA decoder (oracle/_ref/libanim_b200.so) -- tests only.
REF_DIR = os.path.dirname(LIB_PATH)
ANIM_B200_PATH = os.path.join(REF_DIR, "libanim_b200.so")
DWEBP_REF = os.path.join(REF_DIR, "dwebp_ref")
DWEBP_B200 = os.path.join(REF_DIR, "dwebp_b200")
_anim = {}


def _anim_lib(which):
    """which = 'reft' (everything is the reference) or 'b200' (reference demuxer + WebPAnimDecoder over libwebpdecoder_b200)."""
    if which not in _anim:
        L = lib() if which == "reft" else C.CDLL(ANIM_B200_PATH)
        for name, res, args in (("anim_decoder_library", C.c_char_p, []),
                                ("anim_info", C.c_int, [C.c_char_p, C.c_size_t, C.POINTER(C.c_int)]),
                                ("anim_decode_all", C.c_int, [C.c_char_p, C.c_size_t, C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_int)])):
            fn = getattr(L, f"{which}_{name}")
            fn.restype, fn.argtypes = res, args
        _anim[which] = L
    return _anim[which]


def anim_encode(frames, quality=75.0, method=4, kmin=3, kmax=5, minimize_size=0, lossless=0):
    """frames: (n, h, w, 4) uint8 RGBA -> animated WebP bytes (reference WebPAnimEncoder); lossless: 0 lossy frames,
    1 lossless frames, 2 the encoder picks per frame."""
    L = lib()
    L.reft_anim_encode.restype = C.c_size_t
    L.reft_anim_encode.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                   C.POINTER(C.c_void_p)]
    frames = np.ascontiguousarray(frames)
    n, h, w, _ = frames.shape
    p = C.c_void_p()
    size = L.reft_anim_encode(frames.ctypes.data, n, w, h, quality, method, kmin, kmax, minimize_size, lossless, C.byref(p))
    if size == 0:
        raise RuntimeError("WebPAnimEncoder failed")
    data = C.string_at(p, size)
    L.reft_free(p)
    return data


def anim_decoder_library(which="reft"):
    return getattr(_anim_lib(which), f"{which}_anim_decoder_library")().decode()


def anim_decode(data, csp=MODE_RGBA, which="reft"):
    """-> (frames delivered or -1 - delivered on failure, (n, h, w, 4) uint8 canvases, time stamps)."""
    L = _anim_lib(which)
    info = (C.c_int * 3)()
    if not getattr(L, f"{which}_anim_info")(data, len(data), info):
        return -1, None, None
    w, h, n = info[0], info[1], info[2]
    out = np.zeros((max(n, 1), h, w, 4), np.uint8)
    ts = (C.c_int * max(n, 1))()
    got = getattr(L, f"{which}_anim_decode_all")(data, len(data), csp, out.ctypes.data, out.size, ts)
    return got, out, list(ts)
