"""The reference's own CALLERS of the decode path, unmodified, over the CUDA decoder (SURVEY.md 8(b) packaging row,
8(f) item 3): examples/dwebp.c + imageio relinked against libwebpdecoder_b200.so (BASELINE config 1 is literally
`dwebp examples/test.webp -ppm`), and src/demux's WebPAnimDecoder (anim_decode.c:376: one WebPDecode per frame into a
sub-rectangle of the canvas, blended on the host). oracle/Makefile `tools` builds them into oracle/_ref/ from the sources
under /root/reference; on the GPU box the prebuilt files travel with the snapshot. Test infrastructure only."""
import ctypes as C
import hashlib
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, REFERENCE_TREE, ROOT

# sha256 of `dwebp examples/test.webp -ppm` from the reference build (examples/test_ref.ppm itself
# is stale by <= 1 LSB)
PINNED_PPM = "db448ba15096dd0941cacb7e7cc8f0bf5461226c423c330fa73d0591bd8ac980"


@pytest.fixture(scope="module")
def callers(ref, product):
    have = all(os.path.exists(p) for p in (ref.DWEBP_REF, ref.DWEBP_B200, ref.ANIM_B200_PATH))
    if not have and os.path.isdir(REFERENCE_TREE):
        subprocess.check_call(["make", "-s", "-j8", "-C", os.path.join(ROOT, "oracle"), "tools"])
        have = True
    if not have:
        pytest.skip("oracle/_ref/dwebp_b200, libanim_b200.so not available")
    return ref


def animation(ref, n=8, w=96, h=80, seed=7, **kw):
    """A moving textured rectangle over a still background, every third frame with a translucent stripe: the encoder
    emits sub-rectangle frames, blended and not, some with ALPH chunks."""
    fr = np.zeros((n, h, w, 4), np.uint8)
    base = np.zeros((h, w, 4), np.uint8)
    base[..., :3] = ref.synth(w, h, seed)
    base[..., 3] = 255
    for i in range(n):
        f = base.copy()
        f[h // 8 + 4 * i:h // 8 + 4 * i + h // 4, w // 5 + 5 * i:w // 5 + 5 * i + w // 3, :3] = ref.synth(w // 3, h // 4, seed + 100 + i)
        if i % 3 == 2:
            f[h // 2:h // 2 + h // 4, w // 10:w // 10 + w // 3, 3] = np.linspace(0, 255, w // 3, dtype=np.uint8)[None, :]
        fr[i] = f
    return ref.anim_encode(fr, **kw)


# ------------------------------------------------------------------------------------------------- no GPU needed
def test_relinked_callers_bind_to_the_cuda_decoder(callers):
    """The relinked dwebp and the relinked WebPAnimDecoder resolve WebPDecode & co. to libwebpdecoder_b200.so; the
    reference builds of the same code resolve them to the reference."""
    assert "libwebpdecoder_b200" in callers.anim_decoder_library("b200")
    assert "libwebp_ref" in callers.anim_decoder_library("reft")
    env = dict(os.environ, LD_DEBUG="bindings")
    for exe, lib in ((callers.DWEBP_B200, "libwebpdecoder_b200.so"), (callers.DWEBP_REF, "libwebp_ref.so")):
        p = subprocess.run([exe, os.path.join(GOLDEN, "ref_examples_test.webp"), "-ppm", "-o", os.devnull], env=env,
                           capture_output=True, text=True)
        for sym in ("WebPDecode", "WebPGetFeaturesInternal", "WebPInitDecoderConfigInternal", "WebPFreeDecBuffer"):
            lines = [l for l in p.stderr.splitlines() if f"symbol `{sym}'" in l and "binding file " + exe in l]
            assert lines and all(lib in l.split(" to ")[1] for l in lines), (exe, sym, lines[:2])


def test_reference_animation_round_trip(callers):
    """The animation fixture generator does what the GPU test relies on: sub-rectangle frames, blending, ALPH chunks."""
    data = animation(callers)
    assert data.count(b"ANMF") == 8 and data.count(b"ALPH") >= 1 and data.count(b"VP8L") == 0
    n, canvases, ts = callers.anim_decode(data, callers.MODE_RGBA, "reft")
    assert n == 8 and ts == [40 * (i + 1) for i in range(8)]
    assert len({canvases[i].tobytes() for i in range(8)}) == 8


# ------------------------------------------------------------------------------------------------- on the GPU box
def run_dwebp(exe, src, args, out):
    p = subprocess.run([exe, src] + list(args) + ["-o", out], capture_output=True, text=True)
    data = None
    if p.returncode == 0 and os.path.exists(out):
        with open(out, "rb") as f:
            data = f.read()
        os.remove(out)
    return p.returncode, data, p.stderr


@pytest.mark.gpu
def test_dwebp_over_the_cuda_decoder(callers, manifest, amanifest, tmp_path):
    """BASELINE config 1: `dwebp examples/test.webp -ppm` through the relinked dwebp reproduces the pinned hash of the
    reference's output; then every output format dwebp can write here and its decoding flags (dwebp's defaults include
    -dither 50 and -alpha_dither off; -incremental goes through WebPIDecode/WebPIUpdate), file by file against the
    reference's dwebp."""
    src = os.path.join(GOLDEN, "ref_examples_test.webp")
    rc, out, err = run_dwebp(callers.DWEBP_B200, src, ["-ppm"], str(tmp_path / "t.ppm"))
    assert rc == 0, err
    assert hashlib.sha256(out).hexdigest() == PINNED_PPM
    variants = (["-ppm"], ["-pam"], ["-bmp"], ["-tiff"], ["-pgm"], ["-yuv"], ["-ppm", "-nofancy"], ["-pam", "-nofilter"],
                ["-pam", "-nodither"], ["-pam", "-dither", "100"], ["-pam", "-alpha_dither"], ["-pam", "-mt"],
                ["-pam", "-flip"], ["-ppm", "-crop", "3", "5", "40", "30"], ["-pam", "-resize", "50", "37"],
                ["-yuv", "-resize", "301", "0"], ["-pam", "-incremental"], ["-alpha", "-pgm"], ["-pam", "-noasm"],
                ["-pam", "-external_memory", "1"], ["-ppm", "-external_memory", "2"],
                ["-pam", "-crop", "2", "2", "60", "40", "-resize", "33", "90", "-flip", "-alpha_dither"])
    # every start of the relinked dwebp creates a CUDA context (~2 s): all variants on one file with quantised alpha, a few
    # on an opaque one and on a lossless one, the default PAM output on some other fixtures
    some = variants[:1] + variants[5:7] + variants[13:16]
    lossless = str(tmp_path / "lossless.webp")
    pix = np.zeros((90, 120, 4), np.uint8)
    pix[..., :3] = callers.synth(120, 90, 11)
    pix[..., 3] = np.linspace(0, 255, 120, dtype=np.uint8)[None, :]
    with open(lossless, "wb") as f:
        f.write(callers.encode(pix, callers.EncCfg(75, 4, lossless=1)))
    jobs = [(os.path.join(GOLDEN, "alpha_lowq_200x150.webp"), variants), (os.path.join(GOLDEN, "normal_8part_400x300.webp"), some),
            (lossless, variants[1:3] + variants[5:6] + variants[12:17])]
    jobs += [(os.path.join(GOLDEN, e["file"]), (["-pam"],)) for e in (manifest[0], manifest[3], amanifest[0], amanifest[2])]
    n = 0
    for src, todo in jobs:
        e = {"file": os.path.basename(src)}
        for args in todo:
            rc_ref, want, _ = run_dwebp(callers.DWEBP_REF, src, args, str(tmp_path / "r.out"))
            rc, got, err = run_dwebp(callers.DWEBP_B200, src, args, str(tmp_path / "b.out"))
            assert rc == rc_ref, (e["file"], args, rc_ref, rc, err)
            assert got == want, (e["file"], args)
            n += want is not None
    assert n >= 35


@pytest.mark.gpu
def test_anim_decoder_over_the_cuda_decoder(callers):
    """WebPAnimDecoder (reference objects) over the CUDA decoder: every reconstructed canvas of every frame equals the
    all-reference run, in the four colour modes anim_decode.c accepts (premultiplied ones blend differently)."""
    assert "libwebpdecoder_b200" in callers.anim_decoder_library("b200")
    files = [animation(callers), animation(callers, n=12, w=200, h=150, seed=21, quality=50.0, kmin=0, kmax=0),
             animation(callers, n=5, w=64, h=64, seed=33, quality=90.0, kmin=1, kmax=1),
             animation(callers, n=6, w=321, h=123, seed=5, minimize_size=1),
             animation(callers, n=6, w=96, h=80, seed=8, lossless=1), animation(callers, n=9, w=150, h=100, seed=9, lossless=2, quality=30.0)]
    for data in files:
        for csp in (callers.MODE_RGBA, callers.MODE_BGRA, callers.MODE_rgbA, callers.MODE_bgrA):
            n_ref, want, ts_ref = callers.anim_decode(data, csp, "reft")
            n, got, ts = callers.anim_decode(data, csp, "b200")
            assert n == n_ref > 0 and ts == ts_ref
            assert np.array_equal(got, want), (len(data), csp)


def test_anim_batch_info_matches_the_reference_demuxer(callers, product):
    """WebPAnimBatchGetInfo's container walk against the reference's demuxer (no GPU needed): canvas, frame count; still
    pictures and damaged files are refused."""
    for kw in (dict(), dict(n=12, w=200, h=150, seed=21, kmin=0, kmax=0), dict(n=6, w=321, h=123, seed=5, minimize_size=1),
               dict(n=6, w=96, h=80, seed=8, lossless=1)):
        data = animation(callers, **kw)
        n_ref, want, _ = callers.anim_decode(data, callers.MODE_RGBA, "reft")
        info = product.WebPAnimBatchInfo()
        L = product.lib()
        L.WebPAnimBatchGetInfo.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(product.WebPAnimBatchInfo)]
        assert L.WebPAnimBatchGetInfo(data, len(data), C.byref(info)) == 1
        assert (info.frame_count, info.canvas_height, info.canvas_width) == (n_ref, want.shape[1], want.shape[2])
        assert L.WebPAnimBatchGetInfo(data[:len(data) // 2], len(data) // 2, C.byref(info)) == 0
    still = open(os.path.join(ROOT, "tests", "golden", "simple_1part_320x200.webp"), "rb").read()
    assert L.WebPAnimBatchGetInfo(still, len(still), C.byref(info)) == 0


@pytest.mark.gpu
def test_anim_decode_batch(callers, product):
    """WebPAnimDecodeBatch: all frames of a file in ONE batch, canvases rebuilt on the host by the reference's rules --
    every canvas and time stamp equal to successive WebPAnimDecoderGetNext() calls of the all-reference build
    (sub-rectangle frames, blend / no-blend, dispose-to-background, ALPH, lossless and mixed frames, four colour modes)."""
    files = [animation(callers), animation(callers, n=12, w=200, h=150, seed=21, quality=50.0, kmin=0, kmax=0),
             animation(callers, n=5, w=64, h=64, seed=33, quality=90.0, kmin=1, kmax=1),
             animation(callers, n=6, w=321, h=123, seed=5, minimize_size=1),
             animation(callers, n=6, w=96, h=80, seed=8, lossless=1), animation(callers, n=9, w=150, h=100, seed=9, lossless=2, quality=30.0)]
    for data in files:
        for csp in (callers.MODE_RGBA, callers.MODE_BGRA, callers.MODE_rgbA, callers.MODE_bgrA):
            n_ref, want, ts_ref = callers.anim_decode(data, csp, "reft")
            st, got, ts = product.anim_decode_batch(data, csp)
            assert st == 0 and got.shape[0] == n_ref and ts == ts_ref, (st, csp)
            assert np.array_equal(got, want), (len(data), csp)
    assert product.anim_decode_batch(files[0], product.MODE_RGB)[0] == product.VP8_STATUS_INVALID_PARAM
