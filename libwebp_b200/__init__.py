"""libwebp_b200 -- Python view of the drop-in decode library (libwebpdecoder_b200.so).

The product is the C-ABI shared library built from libwebp_b200/csrc (host C + sm_100a CUDA, no torch). This
module is the host-side mirror of the reference's decode interface (src/webp/decode.h) used by the parity
tests and by bench.py: same names (WebPGetFeatures, WebPDecode, WebPDecoderConfig ...), same argument meaning,
same VP8StatusCode results, plus the batch entry points of include/webp/decode_batch.h.

Nothing here decodes on the CPU and nothing here touches oracle/: if the shared library is missing or no
CUDA device is usable, calls fail loudly.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libwebpdecoder_b200.so")

# WEBP_CSP_MODE (include/webp/decode.h)
MODE_RGB, MODE_RGBA, MODE_BGR, MODE_BGRA, MODE_ARGB, MODE_RGBA_4444, MODE_RGB_565 = 0, 1, 2, 3, 4, 5, 6
MODE_rgbA, MODE_bgrA, MODE_Argb, MODE_rgbA_4444, MODE_YUV, MODE_YUVA = 7, 8, 9, 10, 11, 12
BPP = {MODE_RGB: 3, MODE_RGBA: 4, MODE_BGR: 3, MODE_BGRA: 4, MODE_ARGB: 4, MODE_RGBA_4444: 2, MODE_RGB_565: 2,
       MODE_rgbA: 4, MODE_bgrA: 4, MODE_Argb: 4, MODE_rgbA_4444: 2}

# VP8StatusCode
(VP8_STATUS_OK, VP8_STATUS_OUT_OF_MEMORY, VP8_STATUS_INVALID_PARAM, VP8_STATUS_BITSTREAM_ERROR,
 VP8_STATUS_UNSUPPORTED_FEATURE, VP8_STATUS_SUSPENDED, VP8_STATUS_USER_ABORT, VP8_STATUS_NOT_ENOUGH_DATA) = range(8)

WEBP_DECODER_ABI_VERSION = 0x0209
WEBP_BATCH_ABI_VERSION = 0x0101
WEBP_BATCH_HOST, WEBP_BATCH_DEVICE = 0, 1


class WebPRGBABuffer(C.Structure):
    _fields_ = [("rgba", C.c_void_p), ("stride", C.c_int), ("size", C.c_size_t)]


class WebPYUVABuffer(C.Structure):
    _fields_ = [("y", C.c_void_p), ("u", C.c_void_p), ("v", C.c_void_p), ("a", C.c_void_p),
                ("y_stride", C.c_int), ("u_stride", C.c_int), ("v_stride", C.c_int), ("a_stride", C.c_int),
                ("y_size", C.c_size_t), ("u_size", C.c_size_t), ("v_size", C.c_size_t), ("a_size", C.c_size_t)]


class _BufUnion(C.Union):
    _fields_ = [("RGBA", WebPRGBABuffer), ("YUVA", WebPYUVABuffer)]


class WebPDecBuffer(C.Structure):
    _fields_ = [("colorspace", C.c_int), ("width", C.c_int), ("height", C.c_int), ("is_external_memory", C.c_int),
                ("u", _BufUnion), ("pad", C.c_uint32 * 4), ("private_memory", C.c_void_p)]


class WebPBitstreamFeatures(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("has_alpha", C.c_int), ("has_animation", C.c_int),
                ("format", C.c_int), ("pad", C.c_uint32 * 5)]


class WebPDecoderOptions(C.Structure):
    _fields_ = [("bypass_filtering", C.c_int), ("no_fancy_upsampling", C.c_int), ("use_cropping", C.c_int),
                ("crop_left", C.c_int), ("crop_top", C.c_int), ("crop_width", C.c_int), ("crop_height", C.c_int),
                ("use_scaling", C.c_int), ("scaled_width", C.c_int), ("scaled_height", C.c_int),
                ("use_threads", C.c_int), ("dithering_strength", C.c_int), ("flip", C.c_int),
                ("alpha_dithering_strength", C.c_int), ("pad", C.c_uint32 * 5)]


class WebPDecoderConfig(C.Structure):
    _fields_ = [("input", WebPBitstreamFeatures), ("output", WebPDecBuffer), ("options", WebPDecoderOptions)]


class WebPBatchItem(C.Structure):
    _fields_ = [("data", C.c_void_p), ("data_size", C.c_size_t), ("config", C.POINTER(WebPDecoderConfig)),
                ("status", C.c_int)]


class WebPBatchOptions(C.Structure):
    _fields_ = [("device", C.c_int), ("output", C.c_int), ("scratch_bytes", C.c_size_t),
                ("pipeline_waves", C.c_int), ("num_devices", C.c_int), ("devices", C.POINTER(C.c_int)),
                ("stream", C.c_void_p), ("pad", C.c_uint32 * 2)]


class WebPBatchPlane(C.Structure):
    _fields_ = [("y_or_rgba", C.c_void_p), ("u", C.c_void_p), ("v", C.c_void_p), ("stride", C.c_int),
                ("uv_stride", C.c_int), ("width", C.c_int), ("height", C.c_int), ("device", C.c_int)]


class WebPBatchTimings(C.Structure):
    _fields_ = [("total_ms", C.c_float), ("modes_ms", C.c_float), ("tokens_ms", C.c_float), ("recon_ms", C.c_float),
                ("filter_ms", C.c_float), ("emit_ms", C.c_float), ("launches", C.c_int), ("alpha_ms", C.c_float),
                ("pad", C.c_uint32 * 4)]


EXPORTS = [  # every symbol include/webp/*.h declares
    "WebPGetDecoderVersion", "WebPGetInfo", "WebPDecodeRGBA", "WebPDecodeARGB", "WebPDecodeBGRA", "WebPDecodeRGB",
    "WebPDecodeBGR", "WebPDecodeYUV", "WebPDecodeRGBAInto", "WebPDecodeARGBInto", "WebPDecodeBGRAInto",
    "WebPDecodeRGBInto", "WebPDecodeBGRInto", "WebPDecodeYUVInto", "WebPInitDecBufferInternal", "WebPFreeDecBuffer",
    "WebPINewDecoder", "WebPIDecode", "WebPIDelete", "WebPIAppend", "WebPIUpdate", "WebPINewRGB", "WebPINewYUVA", "WebPINewYUV",
    "WebPIDecGetRGB", "WebPIDecGetYUVA", "WebPIDecodedArea", "WebPGetFeaturesInternal",
    "WebPInitDecoderConfigInternal", "WebPDecode", "WebPMalloc", "WebPFree", "VP8GetCPUInfo",
    "WebPBatchOptionsInitInternal", "WebPDecodeBatch", "WebPBatchCreate", "WebPBatchDecode", "WebPBatchDownload",
    "WebPBatchDestroy", "WebPBatchOutput", "WebPBatchGetTimings", "WebPBatchHostAlloc", "WebPBatchHostFree",
    "WebPBatchDeviceCount", "WebPBatchLastError", "WebPBatchSubmit", "WebPBatchWait", "WebPBatchSetCacheLimit",
    "WebPBatchTrimCache", "WebPAnimBatchGetInfo", "WebPAnimDecodeBatch", "WebPBatchDebugStages",
]

_lib = None


def build():
    """Compile the shared library in-tree (nvcc cross-compiles sm_100a without a GPU)."""
    subprocess.check_call(["make", "-s", "-j8", "-C", os.path.join(_HERE, "csrc")])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} not built: run `make -C libwebp_b200/csrc` (or __graft_entry__.build())")
        L = C.CDLL(LIB_PATH)
        L.WebPGetFeaturesInternal.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(WebPBitstreamFeatures), C.c_int]
        L.WebPInitDecoderConfigInternal.argtypes = [C.POINTER(WebPDecoderConfig), C.c_int]
        L.WebPDecode.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(WebPDecoderConfig)]
        L.WebPFreeDecBuffer.argtypes = [C.POINTER(WebPDecBuffer)]
        L.WebPGetInfo.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.WebPDecodeRGBA.restype = C.c_void_p
        L.WebPDecodeRGBA.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.WebPFree.argtypes = [C.c_void_p]
        L.WebPBatchOptionsInitInternal.argtypes = [C.POINTER(WebPBatchOptions), C.c_int]
        L.WebPDecodeBatch.argtypes = [C.POINTER(WebPBatchItem), C.c_int, C.POINTER(WebPBatchOptions)]
        L.WebPBatchCreate.restype = C.c_void_p
        L.WebPBatchCreate.argtypes = [C.POINTER(WebPBatchItem), C.c_int, C.POINTER(WebPBatchOptions), C.POINTER(C.c_int)]
        L.WebPBatchDecode.argtypes = [C.c_void_p]
        L.WebPBatchDownload.argtypes = [C.c_void_p]
        L.WebPBatchDestroy.argtypes = [C.c_void_p]
        L.WebPBatchOutput.argtypes = [C.c_void_p, C.c_int, C.POINTER(WebPBatchPlane)]
        L.WebPBatchGetTimings.argtypes = [C.c_void_p, C.POINTER(WebPBatchTimings)]
        L.WebPBatchDebugStages.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t]
        L.WebPBatchHostAlloc.restype = C.c_void_p
        L.WebPBatchHostAlloc.argtypes = [C.c_size_t]
        L.WebPBatchHostFree.argtypes = [C.c_void_p]
        L.WebPBatchLastError.restype = C.c_char_p
        L.WebPBatchSubmit.restype = C.c_void_p
        L.WebPBatchSubmit.argtypes = [C.POINTER(WebPBatchItem), C.c_int, C.POINTER(WebPBatchOptions), C.POINTER(C.c_int)]
        L.WebPBatchWait.argtypes = [C.c_void_p]
        L.WebPBatchSetCacheLimit.restype = C.c_size_t
        L.WebPBatchSetCacheLimit.argtypes = [C.c_int, C.c_size_t]
        L.WebPBatchTrimCache.restype = C.c_size_t
        L.WebPBatchTrimCache.argtypes = [C.c_int]
        _lib = L
    return _lib


def device_count():
    return lib().WebPBatchDeviceCount()


def set_cache_limit(nbytes, device=-1):
    """WebPBatchSetCacheLimit: device memory the library may keep between batches; returns the previous limit."""
    return lib().WebPBatchSetCacheLimit(device, int(nbytes))


def trim_cache(device=-1):
    """WebPBatchTrimCache: give cached device blocks back to the driver; returns the bytes released."""
    return lib().WebPBatchTrimCache(device)


def last_error():
    return lib().WebPBatchLastError().decode()


def WebPGetFeatures(data):
    """-> (VP8StatusCode, dict) like the reference's WebPGetFeatures."""
    f = WebPBitstreamFeatures()
    st = lib().WebPGetFeaturesInternal(data, len(data), C.byref(f), WEBP_DECODER_ABI_VERSION)
    return st, dict(width=f.width, height=f.height, has_alpha=f.has_alpha, has_animation=f.has_animation, format=f.format)


def scaled_dims(w, h, scaled):
    """WebPRescalerGetScaledDimensions (src/utils/rescaler_utils.c:86-118): 0 = keep the ratio."""
    sw, sh = scaled
    if sw == 0 and h > 0:
        sw = (w * sh + h - 1) // h
    if sh == 0 and w > 0:
        sh = (h * sw + w - 1) // w
    return sw, sh


def _new_config(csp, bypass_filtering, no_fancy_upsampling, dithering_strength=0, crop=None, flip=False, scaled=None,
                alpha_dithering_strength=0):
    cfg = WebPDecoderConfig()
    if not lib().WebPInitDecoderConfigInternal(C.byref(cfg), WEBP_DECODER_ABI_VERSION):
        raise RuntimeError("WebPInitDecoderConfig failed (ABI mismatch)")
    cfg.output.colorspace = csp
    cfg.options.bypass_filtering = int(bool(bypass_filtering))
    cfg.options.no_fancy_upsampling = int(bool(no_fancy_upsampling))
    cfg.options.dithering_strength = dithering_strength
    cfg.options.alpha_dithering_strength = alpha_dithering_strength
    cfg.options.flip = int(bool(flip))
    if crop is not None:
        cfg.options.use_cropping = 1
        (cfg.options.crop_left, cfg.options.crop_top, cfg.options.crop_width, cfg.options.crop_height) = crop
    if scaled is not None:
        cfg.options.use_scaling = 1
        cfg.options.scaled_width, cfg.options.scaled_height = scaled
    return cfg


def _attach_external(cfg, csp, w, h, buf_addr, stride):
    """Point cfg.output at caller memory starting at buf_addr; returns the number of bytes it spans."""
    cfg.output.is_external_memory = 1
    if csp in (MODE_YUV, MODE_YUVA):
        uvw, uvh = (w + 1) // 2, (h + 1) // 2
        y = cfg.output.u.YUVA
        y.y, y.y_stride, y.y_size = buf_addr, w, w * h
        y.u, y.u_stride, y.u_size = buf_addr + w * h, uvw, uvw * uvh
        y.v, y.v_stride, y.v_size = buf_addr + w * h + uvw * uvh, uvw, uvw * uvh
        if csp == MODE_YUVA:
            y.a, y.a_stride, y.a_size = buf_addr + w * h + 2 * uvw * uvh, w, w * h
            return 2 * w * h + 2 * uvw * uvh
        return w * h + 2 * uvw * uvh
    r = cfg.output.u.RGBA
    r.rgba, r.stride, r.size = buf_addr, stride, stride * h
    return stride * h


def out_bytes(csp, w, h, stride=None):
    if csp in (MODE_YUV, MODE_YUVA):
        return w * h + 2 * ((w + 1) // 2) * ((h + 1) // 2) + (w * h if csp == MODE_YUVA else 0)
    return (stride or w * BPP[csp]) * h


def WebPDecode(data, csp=MODE_RGBA, bypass_filtering=False, no_fancy_upsampling=False, stride=None, external=True,
               dithering_strength=0, crop=None, flip=False, scaled=None, alpha_dithering_strength=0):
    """One image through the C-ABI WebPDecode (a GPU batch of one). Returns (status, ndarray or None):
    (h, stride) bytes for RGB-family modes, flat y|u|v for MODE_YUV. crop = (left, top, width, height);
    scaled = (width, height) turns options.use_scaling on (0 = keep the ratio)."""
    L = lib()
    cfg = _new_config(csp, bypass_filtering, no_fancy_upsampling, dithering_strength, crop, flip, scaled, alpha_dithering_strength)
    st, f = WebPGetFeatures(data)
    if st != VP8_STATUS_OK:
        return L.WebPDecode(data, len(data), C.byref(cfg)), None
    w, h = f["width"], f["height"]
    if crop is not None and crop[2] > 0 and crop[3] > 0:
        w, h = crop[2], crop[3]
    if scaled is not None:
        w, h = scaled_dims(w, h, scaled)
        if w <= 0 or h <= 0:
            return L.WebPDecode(data, len(data), C.byref(cfg)), None
    if external:
        if csp in (MODE_YUV, MODE_YUVA):
            out = np.zeros(out_bytes(csp, w, h), np.uint8)
        else:
            stride = stride or w * BPP.get(csp, 4)
            out = np.zeros((h, stride), np.uint8)
        _attach_external(cfg, csp, w, h, out.ctypes.data, stride)
        st = L.WebPDecode(data, len(data), C.byref(cfg))
        return st, (out if st == VP8_STATUS_OK else None)
    st = L.WebPDecode(data, len(data), C.byref(cfg))       # library-allocated output
    if st != VP8_STATUS_OK:
        return st, None
    if csp in (MODE_YUV, MODE_YUVA):
        out = np.ctypeslib.as_array(C.cast(cfg.output.u.YUVA.y, C.POINTER(C.c_uint8)), (out_bytes(csp, w, h),)).copy()
    else:
        s = cfg.output.u.RGBA.stride
        out = np.ctypeslib.as_array(C.cast(cfg.output.u.RGBA.rgba, C.POINTER(C.c_uint8)), (h, s)).copy()
    L.WebPFreeDecBuffer(C.byref(cfg.output))
    return st, out


class HostBuffer:
    """Page-locked host memory from WebPBatchHostAlloc, viewed as a numpy uint8 array."""

    def __init__(self, nbytes):
        self.nbytes = max(int(nbytes), 1)
        self.ptr = lib().WebPBatchHostAlloc(self.nbytes)
        if not self.ptr:
            raise MemoryError(f"WebPBatchHostAlloc({self.nbytes}) failed: {last_error()}")
        self.array = np.ctypeslib.as_array(C.cast(self.ptr, C.POINTER(C.c_uint8)), (self.nbytes,))

    def free(self):
        if self.ptr:
            self.array = None
            lib().WebPBatchHostFree(self.ptr)
            self.ptr = None


class HostArena:
    """One page-locked allocation handed out in pieces (page-locking tens of GB takes seconds: a caller that runs
    many batches does it once). take() never frees; reset() starts over."""

    def __init__(self, nbytes):
        self.buf = HostBuffer(nbytes)
        self.used = 0

    def take(self, nbytes):
        nbytes = max(int(nbytes), 1)
        start = (self.used + 4095) & ~4095
        if start + nbytes > self.buf.nbytes:
            raise MemoryError(f"HostArena: {nbytes} bytes wanted, {self.buf.nbytes - start} left")
        self.used = start + nbytes
        return self.buf.array[start:start + nbytes]

    def reset(self):
        self.used = 0

    def free(self):
        self.buf.free()


class Batch:
    """A batch of .webp files bound to one device: items, configs, packed input and output host buffers.

    inputs are packed back to back into one page-locked buffer (one H2D copy); outputs (WEBP_BATCH_HOST) land
    back to back in another (one D2H copy per 256 MiB run). `datas` may repeat the same bytes object."""

    def __init__(self, datas, csp=MODE_RGBA, bypass_filtering=False, no_fancy_upsampling=False, device=-1,
                 output=WEBP_BATCH_HOST, pinned=True, scratch_bytes=0, devices=None, stream=None, arena=None):
        L = lib()
        self.n = len(datas)
        self.csp = csp
        self.items = (WebPBatchItem * self.n)()
        self.configs = (WebPDecoderConfig * self.n)()
        sizes = [len(d) for d in datas]
        offs = np.concatenate([[0], np.cumsum([(s + 15) & ~15 for s in sizes])]).astype(np.int64)
        self.in_buf = HostBuffer(int(offs[-1]) + 64) if (pinned and arena is None) else None
        in_arr = (arena.take(int(offs[-1]) + 64) if arena is not None else
                  self.in_buf.array if pinned else np.zeros(int(offs[-1]) + 64, np.uint8))
        self._in_arr = in_arr
        self.dims = []
        out_off = [0]
        for i, d in enumerate(datas):
            in_arr[offs[i]:offs[i] + sizes[i]] = np.frombuffer(d, np.uint8)
            st, f = WebPGetFeatures(d)
            w, h = (f["width"], f["height"]) if st == VP8_STATUS_OK else (1, 1)
            self.dims.append((w, h))
            nb = out_bytes(csp, w, h) if csp in BPP or csp in (MODE_YUV, MODE_YUVA) else 4 * w * h
            out_off.append(out_off[-1] + ((nb + 255) & ~255))
        self.out_off = out_off
        self.output_mode = output
        self.out_buf = None
        if output == WEBP_BATCH_HOST:
            self.out_buf = HostBuffer(out_off[-1] + 64) if (pinned and arena is None) else None
            self._out_arr = (arena.take(out_off[-1] + 64) if arena is not None else
                             self.out_buf.array if pinned else np.zeros(out_off[-1] + 64, np.uint8))
        base_in = in_arr.ctypes.data
        for i in range(self.n):
            cfg = self.configs[i]
            L.WebPInitDecoderConfigInternal(C.byref(cfg), WEBP_DECODER_ABI_VERSION)
            cfg.output.colorspace = csp
            cfg.options.bypass_filtering = int(bool(bypass_filtering))
            cfg.options.no_fancy_upsampling = int(bool(no_fancy_upsampling))
            if output == WEBP_BATCH_HOST:
                w, h = self.dims[i]
                _attach_external(cfg, csp, w, h, self._out_arr.ctypes.data + out_off[i], w * BPP.get(csp, 4))
            self.items[i].data = base_in + int(offs[i])
            self.items[i].data_size = sizes[i]
            self.items[i].config = C.pointer(cfg)
        self.opt = WebPBatchOptions()
        L.WebPBatchOptionsInitInternal(C.byref(self.opt), WEBP_BATCH_ABI_VERSION)
        self.opt.device = device
        self.opt.output = output
        self.opt.scratch_bytes = scratch_bytes
        if devices:   # WebPBatchOptions::devices: item i on devices[i % len(devices)], sharded inside the library
            self._devices = (C.c_int * len(devices))(*devices)
            self.opt.devices = C.cast(self._devices, C.POINTER(C.c_int))
            self.opt.num_devices = len(devices)
        if stream:
            self.opt.stream = stream
        self.handle = None
        self.h2d_bytes = int(sum(sizes))
        self.d2h_bytes = int(sum(out_bytes(csp, w, h) for (w, h) in self.dims)) if output == WEBP_BATCH_HOST else 0
        self.mpix = sum(w * h for (w, h) in self.dims) * 1e-6

    # -- one-shot path: what a caller of the C API does (host buffers in, host buffers out)
    def decode_oneshot(self):
        return lib().WebPDecodeBatch(self.items, self.n, C.byref(self.opt))

    # -- asynchronous pair: everything queued by submit(), collected by wait() (two batches in flight hide the download)
    def submit(self):
        st = C.c_int()
        self.handle = lib().WebPBatchSubmit(self.items, self.n, C.byref(self.opt), C.byref(st))
        if not self.handle and st.value in (VP8_STATUS_USER_ABORT, VP8_STATUS_OUT_OF_MEMORY):
            raise RuntimeError(f"WebPBatchSubmit failed ({st.value}): {last_error()}")
        return st.value

    def wait(self):
        st = lib().WebPBatchWait(self.handle)
        self.destroy()
        return st

    # -- resident path
    def create(self):
        st = C.c_int()
        self.handle = lib().WebPBatchCreate(self.items, self.n, C.byref(self.opt), C.byref(st))
        if not self.handle and st.value in (VP8_STATUS_USER_ABORT, VP8_STATUS_OUT_OF_MEMORY):
            raise RuntimeError(f"WebPBatchCreate failed ({st.value}): {last_error()}")
        return st.value

    def decode(self):
        return lib().WebPBatchDecode(self.handle)

    def download(self):
        return lib().WebPBatchDownload(self.handle)

    def timings(self):
        t = WebPBatchTimings()
        lib().WebPBatchGetTimings(self.handle, C.byref(t))
        return {k: getattr(t, k) for k, _ in WebPBatchTimings._fields_ if k != "pad"}

    def device_output(self, i):
        p = WebPBatchPlane()
        if not lib().WebPBatchOutput(self.handle, i, C.byref(p)):
            return None
        return p

    def statuses(self):
        return [self.items[i].status for i in range(self.n)]

    def host_output(self, i):
        """numpy view of image i in the packed host output buffer."""
        w, h = self.dims[i]
        a = self._out_arr[self.out_off[i]:self.out_off[i] + out_bytes(self.csp, w, h)]
        return a if self.csp in (MODE_YUV, MODE_YUVA) else a.reshape(h, w * BPP[self.csp])

    def destroy(self):
        if self.handle:
            lib().WebPBatchDestroy(self.handle)
            self.handle = None

    def close(self):
        self.destroy()
        for b in (self.in_buf, self.out_buf):
            if b is not None:
                b.free()
        self.in_buf = self.out_buf = None


class WebPAnimBatchInfo(C.Structure):
    _fields_ = [("canvas_width", C.c_int), ("canvas_height", C.c_int), ("frame_count", C.c_int), ("loop_count", C.c_int),
                ("bgcolor", C.c_uint32), ("pad", C.c_uint32 * 3)]


def anim_decode_batch(data, csp=MODE_RGBA, device=-1):
    """WebPAnimDecodeBatch: all frames of an animated file in one batch -> (status, canvases [n, h, w, 4] or None, timestamps)."""
    L = lib()
    L.WebPAnimBatchGetInfo.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(WebPAnimBatchInfo)]
    L.WebPAnimDecodeBatch.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_int),
                                      C.POINTER(WebPBatchOptions)]
    info = WebPAnimBatchInfo()
    if not L.WebPAnimBatchGetInfo(data, len(data), C.byref(info)):
        return VP8_STATUS_BITSTREAM_ERROR, None, None
    n, w, h = info.frame_count, info.canvas_width, info.canvas_height
    out = np.zeros((n, h, w, 4), np.uint8)
    ts = (C.c_int * n)()
    opt = WebPBatchOptions()
    L.WebPBatchOptionsInitInternal(C.byref(opt), WEBP_BATCH_ABI_VERSION)
    opt.device = device
    st = L.WebPAnimDecodeBatch(data, len(data), csp, out.ctypes.data, out.size, ts, C.byref(opt))
    return st, (out if st == VP8_STATUS_OK else None), list(ts)


def shard_indices(n_items, rank, world):
    """Items of a batch that rank `rank` of `world` decodes: i % world == rank (images are independent, so the
    batch shards by index with no exchange). Every index belongs to exactly one rank."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return list(range(rank, n_items, world))


def decode_batch_sharded(datas, rank, world, csp=MODE_RGBA, device=-1, **kw):
    """This rank's share of `datas` through WebPDecodeBatch -> (indices, statuses, outputs)."""
    idx = shard_indices(len(datas), rank, world)
    sts, outs = decode_batch([datas[i] for i in idx], csp, device=device, **kw) if idx else ([], [])
    return idx, sts, outs


def decode_batch(datas, csp=MODE_RGBA, bypass_filtering=False, no_fancy_upsampling=False, device=-1, pinned=True):
    """WebPDecodeBatch over a list of files -> (statuses, [ndarray or None])."""
    b = Batch(datas, csp, bypass_filtering, no_fancy_upsampling, device, WEBP_BATCH_HOST, pinned)
    try:
        b.decode_oneshot()
        sts = b.statuses()
        outs = [b.host_output(i).copy() if sts[i] == VP8_STATUS_OK else None for i in range(b.n)]
    finally:
        b.close()
    return sts, outs
