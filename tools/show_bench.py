#!/usr/bin/env python
import json, sys, glob
for f in sorted(sum([glob.glob(a) for a in sys.argv[1:]], [])):
    try:
        d = json.load(open(f))
    except Exception as e:
        print(f, "unreadable", e); continue
    r = d["roofline"]
    print(f"{f.split('/')[-1]:40s} value {d['value']:9.0f} ms/step {d['ms_per_step']:.3f} e2e {d['e2e']['value']:8.0f} frac {r['frac']:.3f} share {r['share_of_step']:.2f} launch_ms {r['launch_ms']:.3f} {r['kernel'][:14]}")
