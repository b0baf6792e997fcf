"""The VP8 constant tables (RFC 6386): the oracle's copy and the product's copy must be the same numbers, and -- where the
reference tree is present (this container; not the GPU box) -- the numbers of src/dec/tree_dec.c and src/dec/quant_dec.c.
A fixed checksum of each table pins them on machines without the reference. (tools/gen_vp8_tables.py wrote both files.)"""
import hashlib
import os
import re
import sys

import pytest

from conftest import ROOT

TABLES = ("kVp8CoeffProba0", "kVp8CoeffUpdateProba", "kVp8BModeProba", "kVp8DcQ", "kVp8AcQ")
SIZES = {"kVp8CoeffProba0": 1056, "kVp8CoeffUpdateProba": 1056, "kVp8BModeProba": 900, "kVp8DcQ": 128, "kVp8AcQ": 128}
# sha256 of ",".join(str(v)) of each table, taken when the files were generated from the reference tree
PINNED = {
    "kVp8CoeffProba0": "81f03c3fd50de5c53581bda04eaf748d4320c1ab6dec6f76808abb8d34fe7fee",
    "kVp8CoeffUpdateProba": "323ed52946a8abac050982399676082fe99b1c7c379dc0b8618d54e163484f8c",
    "kVp8BModeProba": "6d2f62d27a4a27231169d24e9cf1687bbbf37c9cd8a3544df73c31dade3df607",
    "kVp8DcQ": "21418938f757858bba79dc08188003f6ea47d841c8e7856b4bb3efcdcd0d00c6",
    "kVp8AcQ": "0b83a83fd5b46b3e4b4dcf9ad59ed1a06e8ccc4165eff18707d93067427646ff",
}


def read_tables(path):
    src = open(path).read()
    out = {}
    for name in TABLES:
        m = re.search(re.escape(name) + r"\[(\d+)\]\s*=\s*\{([^}]*)\}", src)
        assert m, (path, name)
        vals = [int(x) for x in re.findall(r"\d+", m.group(2))]
        assert len(vals) == int(m.group(1)) == SIZES[name], (path, name, len(vals))
        out[name] = vals
    return out


def digest(vals):
    return hashlib.sha256(",".join(str(v) for v in vals).encode()).hexdigest()


def test_both_copies_hold_the_same_numbers():
    a = read_tables(os.path.join(ROOT, "oracle", "vp8_tables.h"))
    b = read_tables(os.path.join(ROOT, "libwebp_b200", "csrc", "vp8_tables.cuh"))
    for name in TABLES:
        assert a[name] == b[name], name


def test_tables_are_the_pinned_ones():
    a = read_tables(os.path.join(ROOT, "oracle", "vp8_tables.h"))
    for name in TABLES:
        assert digest(a[name]) == PINNED[name], name


def test_tables_match_the_reference_tree():
    ref = os.environ.get("WEBP_REF", "/root/reference")
    if not os.path.isdir(os.path.join(ref, "src", "dec")):
        pytest.skip("no reference tree here (the GPU box): the pinned checksums stand in")
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import gen_vp8_tables as G
    a = read_tables(os.path.join(ROOT, "oracle", "vp8_tables.h"))
    for name, (path, ref_name) in {"kVp8CoeffProba0": ("src/dec/tree_dec.c", "CoeffsProba0"),
                                   "kVp8CoeffUpdateProba": ("src/dec/tree_dec.c", "CoeffsUpdateProba"),
                                   "kVp8BModeProba": ("src/dec/tree_dec.c", "kBModesProba"),
                                   "kVp8DcQ": ("src/dec/quant_dec.c", "kDcTable"),
                                   "kVp8AcQ": ("src/dec/quant_dec.c", "kAcTable")}.items():
        assert a[name] == G.grab(path, ref_name), name
