#!/usr/bin/env python3
"""Per-launch device times of the step kernel (lbm_step + lbm_launch_times_ms): odd vs even iterations,
back to back in one stream.  Usage: python tools/step_times.py [variant] [dim] [precision] [n] [bx,by,bz] [stride]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from lbmcl_b200.capi import Simulation  # noqa: E402

variant = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dim = int(sys.argv[2]) if len(sys.argv) > 2 else 256
prec = sys.argv[3] if len(sys.argv) > 3 else "f32"
n = int(sys.argv[4]) if len(sys.argv) > 4 else 400
block = tuple(int(v) for v in sys.argv[5].split(",")) if len(sys.argv) > 5 else None
stride = int(sys.argv[6]) if len(sys.argv) > 6 else 32
kw = dict(block=block, exact_block=True) if block else {}
with Simulation(dim=dim, precision=prec, stride=stride, variant=variant, **kw) as s:
    s.init()
    for _ in range(n):
        s.step(False)
    t = s.launch_times_ms()[20:]
    bpc = 152 if prec == "f32" else 304
    wet = (dim - 2) ** 3
    odd, even = t[0::2], t[1::2]   # t[0] is iteration 21 (odd)
    print(f"variant {variant} {dim}^3 {prec} block {s.block_shape[0]} stride {stride}: mean {t.mean()*1e3:.1f} us  ({wet*bpc/t.mean()/1e6:.0f} GB/s);"
          f" odd its {np.median(odd)*1e3:.1f} us, even its {np.median(even)*1e3:.1f} us; min {t.min()*1e3:.1f} max {t.max()*1e3:.1f}")
