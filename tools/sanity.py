#!/usr/bin/env python3
"""Tiny end-to-end run for compute-sanitizer: all variants, both precisions, slab group."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from lbmcl_b200.capi import Group, Simulation  # noqa: E402

for prec in ("f32", "f64"):
    for variant in (1, 2, 4):
        for stride in (1, 8, 4096):
            with Simulation(dim=16, precision=prec, stride=stride, variant=variant) as s:
                s.run_snapshots(4, 2)
                s.read_f()
                s.read_map()
    with Group([0, 0, 0, 0], dim=16, precision=prec, stride=16) as g:
        g.run_snapshots(6, 3)
print("sanity ok")
