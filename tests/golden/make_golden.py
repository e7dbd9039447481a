#!/usr/bin/env python
"""Generates tests/golden/reference_digests.json by RUNNING THE UNMODIFIED REFERENCE (oracle/_ref, built from
/root/reference by `make ref`) under the mini-host, opt=0, on the seeded inputs of tests/common.py.

The reference ships no golden vectors of its own, so these are outputs of the reference itself.
For every case the file records SHA-256 digests of the reference's tables (meta, factor), LUT-as-float and output planes,
plus a few raw output samples -- small enough to commit, and enough for a box WITHOUT /root/reference (the GPU box) to
pin the oracle to the reference.   usage: python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "avisynth-jincresize_b200"))
sys.path.insert(0, os.path.join(REPO, "tests"))

from common import SMALL_CASES, make_planes  # noqa: E402
from minihost import avs_host as ah
from oracle import ref as oref  # noqa: E402
from jinc_b200 import paths  # noqa: E402


def digest(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    env = ah.Env()
    env.load_plugin(oref.REF_PLUGIN)
    rt = oref.RefTables()
    out = {"generator": "tests/golden/make_golden.py", "reference": "Asd-g/AviSynth-JincResize v2.1.4, opt=0, threads=1",
           "cases": {}}
    for name, fmt, w, h, tw, th, kw in SMALL_CASES:
        planes = make_planes(fmt, w, h, "noise")
        src = env.source(fmt, w, h, [planes])
        clip = env.invoke("JincResize", src, tw, th, opt=0, threads=1, **kw)
        frame, props = clip.get_frame(0)
        rec = {"lut_f32": digest(rt.lut(clip).astype(np.float32)), "tables": [], "planes": [], "props": props}
        sw, sh = fmt.subsampling
        for k in range(rt.count(clip)):
            dw, dh = (tw, th) if k == 0 else (tw >> sw, th >> sh)
            fs, cs, meta, factor = rt.table(clip, k, dw, dh)
            rec["tables"].append({"filter_size": fs, "coeff_stride": cs, "meta": digest(meta), "factor": digest(factor),
                                  "n_floats": int(factor.size)})
        for p in frame:
            ys = np.linspace(0, p.shape[0] - 1, 5).astype(int)
            xs = np.linspace(0, p.shape[1] - 1, 7).astype(int)
            rec["planes"].append({"sha256": digest(p), "shape": list(p.shape),
                                  "samples": [[float(p[y, x]) for x in xs] for y in ys]})
        out["cases"][name] = rec
        clip.release()
        src.release()
    with open(os.path.join(HERE, "reference_digests.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", len(out["cases"]), "cases")


if __name__ == "__main__":
    main()
