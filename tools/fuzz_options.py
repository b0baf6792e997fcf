#!/usr/bin/env python
"""Option / output-buffer campaign without a GPU (test infrastructure): WebPDecode of intact golden files with random
WebPDecoderConfig contents -- colourspace in and out of range, crop windows and scaling requests that fit, overhang or are
negative, internal memory, external RGBA / YUVA buffers whose strides and sizes sit around the smallest legal values,
missing plane pointers -- through the compiled reference (oracle/_ref) and through the product's host side (plan_item,
prepare_host_buffer in vp8_batch.cu; buffer_dec.c:41-227 restated). Everything the reference refuses (INVALID_PARAM ...)
the product must refuse with the same status before it touches a device, and whatever the reference decodes the product
must let through to the device: in this container that shows as VP8_STATUS_USER_ABORT ("no CUDA device"), never as a
refusal. Exit code 1 on any difference.

    python tools/fuzz_options.py --cases 200000 [--seed 1]
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=100000)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--damage", type=float, default=0.0, help="fraction of cases that use a mutated file (tools/fuzz_emu.py)")
    a = ap.parse_args()
    import libwebp_b200 as W
    from oracle import refwebp as R
    P, Q = W.lib(), R.lib()
    for L in (P, Q):
        L.WebPInitDecoderConfigInternal.argtypes = [C.POINTER(W.WebPDecoderConfig), C.c_int]
        L.WebPDecode.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(W.WebPDecoderConfig)]
        L.WebPFreeDecBuffer.argtypes = [C.POINTER(W.WebPDecBuffer)]
    assert W.device_count() <= 0, "this campaign is for a machine without a GPU"
    g = os.path.join(ROOT, "tests", "golden")
    files = []
    for man in ("manifest.json", "manifest_alpha.json", "manifest_lossless.json"):
        for e in json.load(open(os.path.join(g, man))):
            files.append((e["file"], open(os.path.join(g, e["file"]), "rb").read(), e["features"]))
    rng = np.random.default_rng(a.seed)
    arena = np.zeros(1 << 24, np.uint8)      # external buffers live here; sizes are claimed, never more than the arena
    base = arena.ctypes.data
    hist, bad = {}, []
    seen_pairs = set()
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from fuzz_emu import mutate

    def near(v):
        r = rng.random()
        return int(v) if r < 0.4 else int(v) + int(rng.integers(-3, 4)) if r < 0.8 else int(rng.integers(-5, 2 * abs(int(v)) + 8))

    for case in range(a.cases):
        name, data, f = files[int(rng.integers(0, len(files)))]
        damaged = rng.random() < a.damage
        if damaged:
            data = mutate(rng, data)
            sf, f2 = R.features(data)
            if sf == 0 and f2["width"] * f2["height"] <= (1 << 18):
                f = f2
        w, h = f["width"], f["height"]
        csp = int(rng.integers(0, 13)) if rng.random() < 0.95 else int(rng.integers(-2, 16))
        opt = {}
        ow, oh = w, h
        if rng.random() < 0.4:
            cw, ch = near(rng.integers(1, w + 1)), near(rng.integers(1, h + 1))
            opt.update(use_cropping=1, crop_left=near(rng.integers(0, w)), crop_top=near(rng.integers(0, h)), crop_width=cw, crop_height=ch)
            ow, oh = cw, ch
        if rng.random() < 0.3:
            sw = 0 if rng.random() < 0.15 else near(rng.integers(1, 2 * max(ow, 1) + 2))
            sh = 0 if rng.random() < 0.15 else near(rng.integers(1, 2 * max(oh, 1) + 2))
            opt.update(use_scaling=1, scaled_width=sw, scaled_height=sh)
            if sw > 0 and sh > 0:
                ow, oh = sw, sh
            elif sw > 0 and ow > 0:
                ow, oh = sw, (oh * sw + ow - 1) // ow
            elif sh > 0 and oh > 0:
                ow, oh = (ow * sh + oh - 1) // oh, sh
        if rng.random() < 0.3:
            opt.update(flip=1)
        ow, oh = max(ow, 1), max(oh, 1)
        external = int(rng.choice([0, 1, 1, 1, 2]))
        cfgs = []
        buf = {}
        if external:
            if csp in (11, 12):
                uw, uh = (ow + 1) // 2, (oh + 1) // 2
                ys, us, vs, as_ = near(ow), near(uw), near(uw), near(ow)
                buf = dict(y_stride=ys, u_stride=us, v_stride=vs, a_stride=as_,
                           y_size=max(0, near(abs(ys) * (oh - 1) + ow)), u_size=max(0, near(abs(us) * (uh - 1) + uw)),
                           v_size=max(0, near(abs(vs) * (uh - 1) + uw)), a_size=max(0, near(abs(as_) * (oh - 1) + ow)),
                           y=base + (1 << 21), u=base + (3 << 21), v=base + (5 << 21), a=(base + (7 << 21)) if rng.random() < 0.7 else None)
                if rng.random() < 0.05:
                    buf[str(rng.choice(["y", "u", "v"]))] = None
            else:
                bpp = {0: 3, 2: 3, 5: 2, 6: 2, 10: 2}.get(csp, 4)
                st = near(ow * bpp)
                buf = dict(rgba=base + (1 << 21) if rng.random() < 0.97 else None, stride=st, size=max(0, near(abs(st) * (oh - 1) + ow * bpp)))
        for L in (Q, P):
            cfg = W.WebPDecoderConfig()
            L.WebPInitDecoderConfigInternal(C.byref(cfg), W.WEBP_DECODER_ABI_VERSION)
            cfg.output.colorspace = csp
            cfg.output.is_external_memory = external
            for k, v in opt.items():
                setattr(cfg.options, k, v)
            tgt = cfg.output.u.YUVA if csp in (11, 12) else cfg.output.u.RGBA
            for k, v in buf.items():
                setattr(tgt, k, v)
            cfgs.append(cfg)
        # never let the reference write outside the arena: every plane pointer sits in the middle of a 4 MB slot (a negative
        # stride walks downwards), requests whose extent would leave half a slot are skipped
        if external and any(buf.get(k, 0) and buf[k] > (1 << 21) for k in ("size", "y_size", "u_size", "v_size", "a_size")):
            continue
        if external and (ow * oh * 4 > (1 << 21) or max(abs(buf.get(k, 0)) for k in ("stride", "y_stride", "u_stride", "v_stride", "a_stride")) * oh > (1 << 21)):
            continue
        s_ref = Q.WebPDecode(data, len(data), C.byref(cfgs[0]))
        s_prod = P.WebPDecode(data, len(data), C.byref(cfgs[1]))
        if not external:
            Q.WebPFreeDecBuffer(C.byref(cfgs[0].output))
            P.WebPFreeDecBuffer(C.byref(cfgs[1].output))
        hist[s_ref] = hist.get(s_ref, 0) + 1
        want = W.VP8_STATUS_USER_ABORT if s_ref == 0 else s_ref
        if damaged and s_prod == W.VP8_STATUS_USER_ABORT:
            continue      # a damaged payload is judged on the device: nothing to compare here
        # is_external_memory >= 2 + premultiplied output + a file with alpha: the reference decodes into a buffer of its own and only
        # then looks at the caller's (webp_dec.c:769-786), so an unusable buffer is reported after the decode, and the product has to
        # decode as well before it can say INVALID_PARAM (ItemPlan::discard, vp8_batch.cu): judged on the device too
        if s_prod == W.VP8_STATUS_USER_ABORT and s_ref == 2 and external >= 2 and csp in (7, 8, 9, 10) and f.get("has_alpha"):
            hist["after_decode"] = hist.get("after_decode", 0) + 1
            continue
        if s_prod != want:
            bad.append((name, csp, external, json.dumps(opt), json.dumps({k: v for k, v in buf.items() if k not in ("rgba", "y", "u", "v", "a")}),
                        [k for k in ("rgba", "y", "u", "v", "a") if k in buf and buf[k] is None], (w, h), s_ref, s_prod))
    for t in bad[:40]:
        print("MISMATCH %s csp=%d external=%d opt=%s buf=%s null=%s size=%s ref=%d product=%d" % t)
    print(json.dumps({"cases": a.cases, "mismatches": len(bad), "reference_status_histogram": {str(k): v for k, v in sorted(hist.items(), key=lambda kv: str(kv[0]))},
                      "seed": a.seed}))
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
