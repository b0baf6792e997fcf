#!/usr/bin/env python
"""Damage campaign (CPU, test infrastructure): mutated copies of the golden files and of fresh encodes go through the
compiled reference (oracle/_ref) and through the host build of the device code (tests/emu); status must be equal
for every file and the pixels must be equal wherever the reference decodes. Mutations: byte flips anywhere (container
header included), bursts, truncation, zeroed ranges, chunk-size fields. Prints one line per mismatch and a summary;
exit code 1 if anything differed.

    python tools/fuzz_emu.py --seconds 300 --jobs 8 [--seed 1] [--kinds lossy,alpha,lossless]
"""
import argparse
import ctypes as C
import json
import multiprocessing as mp
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
EMU_DIR = os.path.join(ROOT, "tests", "emu")
MAX_PIXELS = 1 << 20


def load_emu():
    L = C.CDLL(os.path.join(EMU_DIR, "libvp8_emu.so"))
    L.emu_decode.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p]
    L.vp8b_get_features.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p]   # the product's WebPGetFeatures (vp8_container.c)
    return L


def seeds(kinds):
    from oracle import refwebp as R
    g = os.path.join(ROOT, "tests", "golden")
    out = []
    for kind, man in (("lossy", "manifest.json"), ("alpha", "manifest_alpha.json"), ("lossless", "manifest_lossless.json")):
        if kind not in kinds:
            continue
        for e in json.load(open(os.path.join(g, man))):
            out.append((kind, e["file"], open(os.path.join(g, e["file"]), "rb").read()))
    if "lossy" in kinds:   # fresh encodes: small pictures of every encoder configuration of the BASELINE shapes
        for k, cfg in enumerate((R.cfg_simple_1part(), R.cfg_normal_8part(), R.cfg_default(), R.cfg_simple_1part(20.0))):
            out.append(("lossy", "fresh%d" % k, R.encode(R.synth(96 + 16 * k, 64 + 9 * k, 700 + k), cfg)))
    return out


def mutate(rng, data):
    b = bytearray(data)
    kind = int(rng.integers(0, 9))
    n = len(b)
    if kind == 8:      # a partition that starts with 0xFF (what no encoder writes, vp8_literal.h), alone or with more damage
        i = data.find(b"VP8 ")
        if i < 0 or i + 8 + 11 > n:
            kind = 0
        else:
            fo = i + 8
            part0 = int.from_bytes(b[fo:fo + 3], "little") >> 5
            pick = int(rng.integers(0, 3))
            at = fo + 10 if pick == 0 else fo + 10 + part0 + 3 * ((1 << int(rng.integers(0, 4))) - 1)   # first partition / where a token partition may start
            b[min(at, n - 1)] = 0xFF
            for _ in range(int(rng.integers(0, 3))):
                b[int(rng.integers(fo + 10, n))] ^= int(rng.integers(1, 256))
            return bytes(b)
    if kind == 0:      # one byte anywhere
        b[int(rng.integers(0, n))] ^= int(rng.integers(1, 256))
    elif kind == 1:    # a few bytes in the payload
        for _ in range(int(rng.integers(2, 6))):
            b[int(rng.integers(min(30, n - 1), n))] ^= int(rng.integers(1, 256))
    elif kind == 2:    # truncation
        b = b[: int(rng.integers(1, n))]
    elif kind == 3:    # container / frame header region
        b[int(rng.integers(0, min(64, n)))] ^= int(rng.integers(1, 256))
    elif kind == 4:    # burst
        s = int(rng.integers(0, n)); e = min(n, s + int(rng.integers(2, 40)))
        b[s:e] = bytes(rng.integers(0, 256, e - s, dtype=np.uint8))
    elif kind == 5:    # zeroed range
        s = int(rng.integers(12, n)); e = min(n, s + int(rng.integers(1, 64)))
        b[s:e] = bytes(e - s)
    elif kind == 6:    # one bit
        b[int(rng.integers(0, n))] ^= 1 << int(rng.integers(0, 8))
    else:              # truncation + flip
        b = b[: int(rng.integers(max(1, n // 2), n))]
        b[int(rng.integers(0, len(b)))] ^= int(rng.integers(1, 256))
    return bytes(b)


class _Bool:
    """VP8GetBit at prob 1/2 (bit_reader_inl_utils.h:107-136), enough to walk a frame header up to the partition count."""
    def __init__(s, buf):
        s.buf, s.pos, s.range, s.value, s.bits, s.eof = buf, 0, 254, 0, -8, 0
    def bit(s):
        if s.bits < 0:
            if s.pos < len(s.buf):
                s.value = (s.value << 8) | s.buf[s.pos]; s.pos += 1; s.bits += 8
            elif not s.eof:
                s.value <<= 8; s.bits += 8; s.eof = 1
            else:
                s.bits = 0
        split = (s.range * 128) >> 8
        bit = (s.value >> s.bits) > split
        if bit:
            rng = s.range - split; s.value -= (split + 1) << s.bits
        else:
            rng = split + 1
        shift = 7 ^ (rng.bit_length() - 1)
        s.bits -= shift; s.range = (rng << shift) - 1
        return int(bit)
    def val(s, n):
        v = 0
        for _ in range(n):
            v = (v << 1) | s.bit()
        return v


def chunk_spans(b):
    """{fourcc: (payload offset, payload size)} of a RIFF file whose chunk headers are intact, else {}."""
    if b[:4] != b"RIFF" or b[8:12] != b"WEBP":
        return {}
    o, out = 12, {}
    while o + 8 <= len(b):
        sz = int.from_bytes(b[o + 4:o + 8], "little")
        out.setdefault(bytes(b[o:o + 4]), (o + 8, min(sz, len(b) - o - 8)))
        o += 8 + sz + (sz & 1)
    return out


def partition_starts_with_ff(b):
    """A partition whose first byte is 0xFF breaks the arithmetic decoder's invariant (value <= range) from the first bit on:
    no encoder emits it, and what the reference then decodes depends on its 64-bit window wrapping around. Known deviation."""
    sp = chunk_spans(b).get(b"VP8 ")
    if not sp or sp[1] < 11:
        return False
    fo, fs = sp
    part0 = int.from_bytes(b[fo:fo + 3], "little") >> 5
    if b[fo + 10] == 0xFF:
        return True
    br = _Bool(b[fo + 10:fo + 10 + part0])
    br.val(2)
    if br.bit():
        um = br.bit()
        if br.bit():
            br.bit()
            for n in (7, 7, 7, 7, 6, 6, 6, 6):
                if br.bit():
                    br.val(n + 1)
        if um:
            for _ in range(3):
                if br.bit():
                    br.val(8)
    br.val(10)
    if br.bit() and br.bit():
        for _ in range(8):
            if br.bit():
                br.val(7)
    nparts = 1 << br.val(2)
    o = fo + 10 + part0
    left = fo + fs - o - 3 * (nparts - 1)
    if left <= 0:
        return False
    start = o + 3 * (nparts - 1)
    for p in range(nparts):
        if start < len(b) and b[start] == 0xFF:
            return True
        if p < nparts - 1:
            psz = min(int.from_bytes(b[o + 3 * p:o + 3 * p + 3], "little"), left)
            start += psz; left -= psz
    return False


def damaged_in_both_chunks(R, b, data, s_ref, s_emu):
    """ALPH chunk and VP8 payload both damaged: the reference reports whichever failure its row loop meets first (the alpha
    rows are decoded as the macroblock rows above them finish, frame_dec.c:452-460). The default (fp) token parser keeps the
    failing rows and reports the same; the older parsers kept as debug paths report the VP8 status. Confirmed by damaging
    one chunk at a time."""
    sa, so = chunk_spans(b), chunk_spans(data)
    # (the VP8 chunk must start where it did; its size field may be part of the damage)
    if b"ALPH" not in sa or sa.get(b"ALPH") != so.get(b"ALPH") or b"VP8 " not in sa or sa[b"VP8 "][0] != so[b"VP8 "][0] or len(b) > len(data):
        return False
    ao, asz = sa[b"ALPH"]
    only_alpha = bytearray(data); only_alpha[ao:ao + asz] = b[ao:ao + asz]
    only_vp8 = bytearray(b); only_vp8[ao:ao + asz] = data[ao:ao + asz]
    a = R.decode(bytes(only_alpha), R.MODE_RGBA, 0)[0]
    v = R.decode(bytes(only_vp8), R.MODE_RGBA, 0)[0]
    return a != 0 and v != 0 and {s_ref, s_emu} <= {a, v}


def worker(args):
    wid, seed, seconds, kinds, use_port = args
    from oracle import refwebp as R
    L = load_emu()
    if use_port:      # the plain-C restatement (oracle/vp8_oracle.c) in the emulation's place: pins the oracle itself (lossy, opaque)
        from oracle import portwebp as PORT
    rng = np.random.default_rng(seed * 1000 + wid)
    S = seeds(kinds)
    t0 = time.time()
    n = ok = 0
    bad = []
    hist = {}
    while time.time() - t0 < seconds:
        kind, name, data = S[int(rng.integers(0, len(S)))]
        b = mutate(rng, data)
        sf, f = R.features(b)
        s_ref, want = R.decode(b, R.MODE_RGBA, 0)
        w, h = (f["width"], f["height"]) if sf == 0 else (1, 1)
        if w * h > MAX_PIXELS:
            continue
        variant = (0, 2, 8, 24, 56, 64, 320, 336, 192, 208, 320, 64)[int(rng.integers(0, 12))]   # +128: the banded probability rows, +256: the lockstep mode parser
        # the product's order (plan_item, vp8_batch.cu): feature probe first, its NOT_ENOUGH_DATA becomes BITSTREAM_ERROR
        # (webp_dec.c:761-767), any other failure is returned as it is; only then the decode proper
        if use_port:
            if sf == 0 and (f["format"] != 1 or f["has_alpha"]):
                continue      # the restatement covers opaque lossy pictures only
            n += 1
            hist[s_ref] = hist.get(s_ref, 0) + 1
            s_port, got = PORT.decode(b, PORT.RGBA, 0)
            same = s_port == s_ref and PORT.features(b) == (sf, f if sf == 0 else PORT.features(b)[1])
            if same and s_ref == 0 and not np.array_equal(want.reshape(-1), got.reshape(-1)):
                _, want = R.decode(b, R.MODE_RGBA, 0, simd=False)
                hist["simd_vs_c"] = hist.get("simd_vs_c", 0) + 1
                same = np.array_equal(want.reshape(-1), got.reshape(-1))
            if same:
                ok += 1
            else:
                bad.append(("port_%s_w%d_%d" % (name, wid, n), s_ref, s_port, -1))
                os.makedirs(os.path.join(ROOT, "gpurun_out", "fuzz"), exist_ok=True)
                open(os.path.join(ROOT, "gpurun_out", "fuzz", "port_%s_w%d_%d.webp" % (name, wid, n)), "wb").write(b)
            continue
        pf = (C.c_int * 10)()
        s_emu = L.vp8b_get_features(b, len(b), pf)
        feat_ok = s_emu == sf and (sf != 0 or list(pf[:5]) == [f["width"], f["height"], f["has_alpha"], f["has_animation"], f["format"]])
        if s_emu == 7:
            s_emu = 3
        out = np.zeros((max(h, 1), max(w, 1) * 4), np.uint8)
        if s_emu == 0:
            s_emu = L.emu_decode(b, len(b), 1, 0, out.ctypes.data, out.size, max(w, 1) * 4, variant, None)
        if s_ref == 0 and s_emu == 0 and not np.array_equal(want.reshape(-1), out.reshape(-1)):
            # out-of-range coefficients of a damaged stream: the reference's SSE2 transforms wrap at 16 bits where its C ones
            # do not, so the reference disagrees with itself there; the C dsp path (VP8GetCPUInfo = NULL) is the one to equal
            _, want = R.decode(b, R.MODE_RGBA, 0, simd=False)
            hist["simd_vs_c"] = hist.get("simd_vs_c", 0) + 1
        n += 1
        hist[s_ref] = hist.get(s_ref, 0) + 1
        differs = not feat_ok or s_emu != s_ref or (s_ref == 0 and not np.array_equal(want.reshape(-1), out.reshape(-1)))
        if differs and feat_ok and not (variant & 64) and partition_starts_with_ff(b):
            # the default parser hands such images to the literal reader (vp8_literal.h); only the older parsers keep this class
            hist["known_ff_first_byte"] = hist.get("known_ff_first_byte", 0) + 1
            ok += 1
        elif differs and feat_ok and s_emu != s_ref and not (variant & 64) and damaged_in_both_chunks(R, b, data, s_ref, s_emu):
            # only the older parsers (debug paths) keep this class: the fp parser records the failing row and the rule of
            # vp8_dev.h:vp8b_vp8_failure_first picks the status the reference reports
            hist["known_both_chunks_damaged"] = hist.get("known_both_chunks_damaged", 0) + 1
            ok += 1
        elif differs:
            tag = "%s_w%d_%d" % (name, wid, n)
            bad.append((tag, s_ref, s_emu, variant))
            os.makedirs(os.path.join(ROOT, "gpurun_out", "fuzz"), exist_ok=True)
            open(os.path.join(ROOT, "gpurun_out", "fuzz", tag + ".webp"), "wb").write(b)
        else:
            ok += 1
    return n, ok, bad, hist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=60)
    ap.add_argument("--jobs", type=int, default=max(1, (os.cpu_count() or 2) - 1))
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--kinds", default="lossy,alpha,lossless")
    ap.add_argument("--port", action="store_true", help="check the oracle's C restatement instead (use with --kinds lossy)")
    a = ap.parse_args()
    subprocess.check_call(["make", "-s", "-C", EMU_DIR])
    kinds = a.kinds.split(",")
    with mp.get_context("spawn").Pool(a.jobs) as p:
        res = p.map(worker, [(i, a.seed, a.seconds, kinds, a.port) for i in range(a.jobs)])
    n = sum(r[0] for r in res); ok = sum(r[1] for r in res)
    hist = {}
    for r in res:
        for k, v in r[3].items():
            hist[k] = hist.get(k, 0) + v
        for t in r[2]:
            print("MISMATCH file=%s ref=%d emu=%d variant=%d" % t)
    print(json.dumps({"cases": n, "equal": ok, "mismatches": n - ok, "reference_status_histogram": {str(k): v for k, v in sorted(hist.items(), key=lambda kv: str(kv[0]))},
                      "seed": a.seed, "jobs": a.jobs, "seconds": a.seconds, "kinds": kinds, "decoder": "port" if a.port else "emu"}))
    sys.exit(0 if n == ok else 1)


if __name__ == "__main__":
    main()
