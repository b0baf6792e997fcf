"""ctypes binding of oracle/libvp8_oracle.so (the plain-C restatement, oracle/vp8_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, bench.py's cpu_baseline leg (when oracle/_ref is absent) and
__graft_entry__.smoke(). The product package libwebp_b200 never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvp8_oracle.so")

RGB, RGBA, BGR, BGRA, ARGB, rgbA, bgrA, Argb, YUV = 0, 1, 2, 3, 4, 7, 8, 9, 11
BPP = {0: 3, 1: 4, 2: 3, 3: 4, 4: 4, 7: 4, 8: 4, 9: 4}
FLAG_BYPASS_FILTER, FLAG_NO_FANCY = 1, 2


class Dump(C.Structure):
    _fields_ = [("status", C.c_int), ("width", C.c_int), ("height", C.c_int), ("mb_w", C.c_int), ("mb_h", C.c_int),
                ("filter_type", C.c_int), ("num_parts", C.c_int), ("dq", (C.c_int * 6) * 4),
                ("modes", C.POINTER(C.c_uint8)), ("coeffs", C.POINTER(C.c_int16)), ("nz", C.POINTER(C.c_uint32)),
                ("finfo", C.POINTER(C.c_uint8)), ("y_stride", C.c_int), ("uv_stride", C.c_int),
                ("unfiltered", C.POINTER(C.c_uint8)), ("filtered", C.POINTER(C.c_uint8))]


_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, "port"])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        L.vp8o_features.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.c_int)]
        L.vp8o_decode.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_int]
        L.vp8o_dump.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(Dump)]
        L.vp8o_dump_free.argtypes = [C.POINTER(Dump)]
        L.vp8o_count_decodes.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.c_uint64)]
        _lib = L
    return _lib


def features(data):
    f = (C.c_int * 5)()
    st = lib().vp8o_features(data, len(data), f)
    return st, dict(width=f[0], height=f[1], has_alpha=f[2], has_animation=f[3], format=f[4])


def decode(data, csp=RGBA, flags=0, stride=None):
    """Same contract as oracle.refwebp.decode."""
    st, f = features(data)
    if st != 0:
        return (3 if st == 7 else st), None
    w, h = f["width"], f["height"]
    if csp == YUV:
        n = w * h + 2 * ((w + 1) // 2) * ((h + 1) // 2)
        out = np.zeros(n, np.uint8)
        st = lib().vp8o_decode(data, len(data), csp, flags, out.ctypes.data, n, 0)
    else:
        stride = stride or w * BPP[csp]
        out = np.zeros((h, stride), np.uint8)
        st = lib().vp8o_decode(data, len(data), csp, flags, out.ctypes.data, out.size, stride)
    return st, (out if st == 0 else None)


def count_decodes(data):
    """-> (status, first-partition decodes, [decodes of each token partition]): the lengths of the frame's dependent chains."""
    out = (C.c_uint64 * 10)()
    st = lib().vp8o_count_decodes(data, len(data), out)
    return st, int(out[0]), [int(out[1 + p]) for p in range(int(out[9]))]


def dump(data):
    """Per-stage intermediates as numpy arrays (copied out of the C allocation)."""
    d = Dump()
    st = lib().vp8o_dump(data, len(data), C.byref(d))
    if st != 0:
        return st, None
    nmb = d.mb_w * d.mb_h
    ys, uvs = d.y_stride, d.uv_stride
    plane = ys * 16 * d.mb_h + 2 * uvs * 8 * d.mb_h
    res = dict(width=d.width, height=d.height, mb_w=d.mb_w, mb_h=d.mb_h, filter_type=d.filter_type,
               num_parts=d.num_parts, dq=np.array([[d.dq[s][k] for k in range(6)] for s in range(4)]),
               modes=np.ctypeslib.as_array(d.modes, (nmb, 20)).copy(),
               coeffs=np.ctypeslib.as_array(d.coeffs, (nmb, 384)).copy(),
               nz=np.ctypeslib.as_array(d.nz, (nmb, 2)).copy(),
               finfo=np.ctypeslib.as_array(d.finfo, (nmb, 4)).copy(),
               y_stride=ys, uv_stride=uvs,
               unfiltered=np.ctypeslib.as_array(d.unfiltered, (plane,)).copy(),
               filtered=np.ctypeslib.as_array(d.filtered, (plane,)).copy())
    lib().vp8o_dump_free(C.byref(d))
    return 0, res
