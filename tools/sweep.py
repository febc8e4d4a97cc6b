#!/usr/bin/env python3
"""Timing sweep over kernel variant / CSoA stride / block shape (the B200 counterpart of the
reference's benchmark.sh sweep over lws and stride).  Prints MLUPS and GB/s per configuration.
Usage: python tools/sweep.py [--dim 256] [--precision f32] [--its 200]"""
import argparse
import itertools
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from lbmcl_b200.capi import Simulation  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dim", type=int, default=256)
    ap.add_argument("--precision", default="f32")
    ap.add_argument("--its", type=int, default=200)
    ap.add_argument("--fast", type=int, nargs="*", default=[0, 1])
    ap.add_argument("--variants", type=int, nargs="*", default=[1, 2, 4])
    ap.add_argument("--strides", type=int, nargs="*", default=[32, 64, 128, 256, 65536, 16777216])
    ap.add_argument("--blocks", nargs="*", default=["256,1,1", "128,2,1", "64,4,1", "32,8,1", "256,4,1", "128,8,1"])
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    bpc = 152 if a.precision == "f32" else 304
    wet = (a.dim - 2) ** 3
    rows = []
    for fast, variant, stride, blk in itertools.product(a.fast, a.variants, a.strides, a.blocks):
        if stride > a.dim ** 3:
            continue
        block = tuple(int(v) for v in blk.split(","))
        try:
            with Simulation(dim=a.dim, precision=a.precision, stride=stride, block=block, variant=variant,
                            fast_math=bool(fast), exact_block=True) as s:
                s.init()
                s.run(10, 0)
                s.sync()
                s.init()
                s.run(a.its, 0)
                total, kernels = s.time_ms()
                eff_block, vec = s.block_shape
        except Exception as e:  # noqa: BLE001
            print("FAIL", fast, variant, stride, blk, e)
            continue
        mlups = wet * a.its / (kernels * 1e3)
        gbs = mlups * 1e6 * bpc / 1e9
        row = dict(fast=fast, vec=vec, stride=stride, req_block=blk, block=list(eff_block), ms_per_it=kernels / a.its,
                   mlups=mlups, gbs=gbs)
        rows.append(row)
        print(json.dumps(row), flush=True)
    rows.sort(key=lambda r: -r["mlups"])
    print("BEST:")
    for r in rows[:10]:
        print(json.dumps(r))
    if a.out:
        with open(a.out, "w") as fh:
            json.dump(rows, fh, indent=1)


if __name__ == "__main__":
    main()
