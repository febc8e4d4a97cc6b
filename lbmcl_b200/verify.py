#!/usr/bin/env python3
"""verify.py of the reference, restated without pyvista/vtk/sklearn and with pass/fail thresholds.

Reference contract (reference verify.py:14-60): for it in range(0, iterations + 1, every) compare
target ``<t>/ldc.0.<it>.vti`` with prediction ``<p>/lbmcl.<it>.vti`` (zero padded to the number of
digits of ``iterations``) and print one row  ``it: MSE_RHO MSE_U MAE_RHO MAE_U``  where both fields
go through ``nan_to_num`` first, MSE is the mean squared error over all elements and "MAE" is the
MAXIMUM absolute error.  The reference never fails; it prints and exits 0.

Additions here
  * ``-t`` may also name one of the committed fixture archives (tests/golden/target8.npz /
    target32.npz, made from the reference's .vti files by tests/golden/make_golden.py); iterations
    the archive does not hold are reported as "no target";
  * ``--check`` turns the table into a test: NaN masks must be identical and
    max|d rho| <= --tol-rho (default 4e-6), max|d u| / U <= --tol-u (default 1.2e-5) -- the level at
    which the reference's own kernel agrees with these Sailfish fixtures;
    exit status 1 on failure.
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from lbmcl_b200.vti import read_vti  # noqa: E402


def load_fields(path: str):
    d = read_vti(path)
    return d["arrays"]["rho"], d["arrays"]["v"]


def metrics(t_rho, t_v, p_rho, p_v):
    """verify.py:47-56"""
    masks_equal = bool(np.array_equal(np.isnan(t_rho), np.isnan(p_rho)) and np.array_equal(np.isnan(t_v), np.isnan(p_v)))
    tr, pr = np.nan_to_num(t_rho).astype(np.float64), np.nan_to_num(p_rho).astype(np.float64)
    tv, pv = np.nan_to_num(t_v).astype(np.float64), np.nan_to_num(p_v).astype(np.float64)
    return {
        "mse_rho": float(np.mean((tr - pr) ** 2)),
        "mse_v": float(np.mean((tv - pv) ** 2)),
        "mae_rho": float(np.max(np.abs(tr - pr))),
        "mae_v": float(np.max(np.abs(tv - pv))),
        "masks_equal": masks_equal,
    }


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description="Calculate the Mean Squared Error (MSE) and the Max Absolute Error (MAE) "
                                             "of density (rho) and velocity (u) from two LBM simulation datasets "
                                             "(target and prediction).")
    ap.add_argument("-i", "--iterations", type=int, required=True, help="number of iterations in VTI dataset")
    ap.add_argument("-e", "--every", type=int, required=True, help="step number between two iterations")
    ap.add_argument("-t", "--target_path", type=str, required=True,
                    help="path of target VTI files (format ldc.0.*.vti) or a tests/golden/*.npz archive")
    ap.add_argument("-p", "--prediction_path", type=str, required=True,
                    help="path of prediction VTI files (format lbmcl.*.vti)")
    ap.add_argument("--check", action="store_true", help="fail (exit 1) when a threshold is exceeded")
    ap.add_argument("--tol-rho", type=float, default=4e-6)
    ap.add_argument("--tol-u", type=float, default=1.2e-5, help="relative to the lid speed")
    ap.add_argument("--velocity", type=float, default=0.05, help="lid speed used to normalise --tol-u")
    a = ap.parse_args(argv)

    width = int(np.log10(a.iterations)) + 1 if a.iterations > 0 else 1
    archive = np.load(a.target_path) if a.target_path.endswith(".npz") else None
    arch_its = {int(v): k for k, v in enumerate(archive["its"])} if archive is not None else {}

    print("{0:^{w}}  {1:^13}  {2:^13}  {3:^13}  {4:^13}".format("#it", "MSE_RHO", "MSE_U", "MAE_RHO", "MAE_U", w=width))
    ok = True
    compared = 0
    for it in range(0, a.iterations + 1, a.every):
        pred = "{p}/lbmcl.{i:0{w}}.vti".format(p=a.prediction_path, i=it, w=width)
        p_rho, p_v = load_fields(pred)
        if archive is not None:
            if it not in arch_its:
                print("{i:{w}}:  no target in the archive".format(i=it, w=width))
                continue
            k = arch_its[it]
            t_rho = archive["rho"][k].reshape(-1)
            t_v = archive["v"][k].reshape(-1, 3)
        else:
            t_rho, t_v = load_fields("{p}/ldc.0.{i:0{w}}.vti".format(p=a.target_path, i=it, w=width))
        m = metrics(t_rho, t_v, p_rho, p_v)
        compared += 1
        print("{i:{w}}:  {a:e}   {b:e}   {c:e}   {d:e}".format(i=it, w=width, a=m["mse_rho"], b=m["mse_v"],
                                                             c=m["mae_rho"], d=m["mae_v"]))
        if a.check:
            bad = []
            if not m["masks_equal"]:
                bad.append("NaN masks differ")
            if m["mae_rho"] > a.tol_rho:
                bad.append("max|d rho| %.3e > %.1e" % (m["mae_rho"], a.tol_rho))
            if m["mae_v"] > a.tol_u * abs(a.velocity):
                bad.append("max|d u|/U %.3e > %.1e" % (m["mae_v"] / abs(a.velocity), a.tol_u))
            if bad:
                ok = False
                print("     FAIL: " + "; ".join(bad))
    if a.check:
        if compared == 0:
            print("verify: nothing compared")
            return 1
        print("verify: %s (%d iterations compared)" % ("PASS" if ok else "FAIL", compared))
        return 0 if ok else 1
    return 0


if __name__ == "__main__":
    sys.exit(main())
