import hashlib
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
REFERENCE_TREE = "/root/reference"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="session")
def manifest():
    with open(os.path.join(GOLDEN, "manifest.json")) as f:
        m = json.load(f)
    for e in m:
        with open(os.path.join(GOLDEN, e["file"]), "rb") as f:
            e["data"] = f.read()
    return m


@pytest.fixture(scope="session")
def amanifest():
    """Fixtures with an ALPH chunk (reference-produced hashes, tests/golden/make_golden.py)."""
    with open(os.path.join(GOLDEN, "manifest_alpha.json")) as f:
        m = json.load(f)
    for e in m:
        with open(os.path.join(GOLDEN, e["file"]), "rb") as f:
            e["data"] = f.read()
    return m


@pytest.fixture(scope="session")
def lmanifest():
    """Lossless (VP8L) fixtures (reference-produced hashes, tests/golden/make_golden_lossless.py)."""
    with open(os.path.join(GOLDEN, "manifest_lossless.json")) as f:
        m = json.load(f)
    for e in m:
        with open(os.path.join(GOLDEN, e["file"]), "rb") as f:
            e["data"] = f.read()
    return m


@pytest.fixture(scope="session")
def port():
    """The plain-C oracle (oracle/vp8_oracle.c); compiled on demand with gcc."""
    from oracle import portwebp
    portwebp.build()
    portwebp.lib()
    return portwebp


@pytest.fixture(scope="session")
def ref():
    """The compiled, unmodified reference (oracle/_ref). Built here when /root/reference exists; on the GPU box
    the prebuilt .so travels with the snapshot. Tests that need it skip when neither is available."""
    from oracle import refwebp
    if not refwebp.available() and os.path.isdir(REFERENCE_TREE):
        subprocess.check_call(["make", "-s", "-j8", "-C", os.path.join(ROOT, "oracle"), "ref"])
    if not refwebp.available():
        pytest.skip("oracle/_ref/libwebp_ref.so not available")
    refwebp.lib()
    return refwebp


@pytest.fixture(scope="session")
def product():
    """The product library through its Python mirror; built in-tree on demand (nvcc needs no GPU)."""
    import libwebp_b200 as W
    if not os.path.exists(W.LIB_PATH):
        W.build()
    W.lib()
    return W


def truncated(data, n):
    """A RIFF file cut to n bytes with the RIFF size field left alone (what a short read looks like)."""
    return data[:n]


def smooth_image(w, h, seed):
    """A picture of slow sinusoids with one noisy patch: at fine quantisers most macroblocks end up without AC chroma
    coefficients, which is what the decoder's random dithering looks for (vp8_dec.c:603)."""
    import numpy as np
    y, x = np.mgrid[0:h, 0:w].astype(np.float64)
    rng = np.random.default_rng(seed)
    img = np.zeros((h, w, 3), np.uint8)
    for c in range(3):
        a, b, ph = rng.uniform(0.5, 2, 3)
        img[..., c] = np.clip(128 + 100 * np.sin(x * a * 0.01 + ph) * np.cos(y * b * 0.012), 0, 255)
    img[h // 3:h // 2, w // 4:w // 2] = rng.integers(0, 255, (h // 2 - h // 3, w // 2 - w // 4, 3), dtype=np.uint8)
    return img
