#!/usr/bin/env python3
"""Regenerate tests/golden/*.npz from the reference's golden fixtures.

Source: /root/reference/target_results/8/ldc.0.{00..10}.vti (11 files) and
/root/reference/target_results/32/ldc.0.{000,020,..,500}.vti (26 files): Sailfish fp32 output for
the lid-driven cavity (produced upstream by remoteVerify.sh:106-133).  They are the
only known-answer vectors the reference has for the collide-and-stream path.

The .vti files are not copied; their decoded Float32 payload is stored bit-for-bit:
    target8.npz   its[11],  rho[11, 6,6,6],    v[11, 6,6,6, 3]
    target32.npz  its[9],   rho[9, 30,30,30],  v[9, 30,30,30, 3]      (axes: z, y, x)
                  iterations 0,20,40,60,100,200,300,400,500 -- a subset, to keep the committed file
                  small; tests/test_oracle_golden.py checks all 26 in place when /root/reference exists
Target point (i,j,k) corresponds to LBMCL cell (x,y,z) = (i+1, j+1, k+1).

Run here (needs /root/reference); the GPU box only reads the committed .npz files.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from lbmcl_b200.vti import read_vti  # noqa: E402

REF = os.environ.get("LBMCL_REF", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def pack(sub: str, its, width: int, out: str) -> None:
    rhos, vs = [], []
    for it in its:
        d = read_vti(os.path.join(REF, "target_results", sub, f"ldc.0.{it:0{width}d}.vti"))
        nx, ny, nz = d["dims"]
        assert d["arrays"]["rho"].dtype == np.float32
        rhos.append(d["arrays"]["rho"].reshape(nz, ny, nx))
        vs.append(d["arrays"]["v"].reshape(nz, ny, nx, 3))
    np.savez_compressed(os.path.join(HERE, out), its=np.array(its, dtype=np.int32),
                        rho=np.stack(rhos), v=np.stack(vs))
    print(out, os.path.getsize(os.path.join(HERE, out)), "bytes")


if __name__ == "__main__":
    pack("8", list(range(0, 11)), 2, "target8.npz")
    pack("32", [0, 20, 40, 60, 100, 200, 300, 400, 500], 3, "target32.npz")
