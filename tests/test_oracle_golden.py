"""The CPU oracle (oracle/jinc_oracle.c) against golden digests of the UNMODIFIED reference's own tables and output
(tests/golden/reference_digests.json, produced by tests/golden/make_golden.py from oracle/_ref).  Bit-exact: the
oracle's LUT-as-float, meta[], factor[] and opt=0 output planes must hash to the reference's."""
import hashlib
import json
import os

import numpy as np
import pytest

from common import SMALL_CASES, make_planes, oracle_frame

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_digests.json")))


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("case", SMALL_CASES, ids=[c[0] for c in SMALL_CASES])
def test_oracle_reproduces_reference_digests(native_built, case):
    from oracle import cpu as oc

    name, fmt, w, h, tw, th, kw = case
    rec = GOLD["cases"][name]
    planes = make_planes(fmt, w, h, "noise")
    out, tabs = oracle_frame(fmt, w, h, tw, th, planes, **kw)
    assert digest(oc.make_lut(kw.get("tap", 3), kw.get("blur", 0.0)).astype(np.float32)) == rec["lut_f32"]
    assert len(tabs) == len(rec["tables"])
    for t, g in zip(tabs, rec["tables"]):
        assert (t.filter_size, t.coeff_stride) == (g["filter_size"], g["coeff_stride"])
        assert t.factor.size == g["n_floats"]
        assert digest(t.meta) == g["meta"]
        assert digest(t.factor) == g["factor"]
    for o, g in zip(out, rec["planes"]):
        assert list(o.shape) == g["shape"]
        assert digest(o) == g["sha256"]
