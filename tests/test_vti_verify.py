"""VTI reader/writer and the verify.py restatement (CPU only)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, wet
from lbmcl_b200 import verify
from lbmcl_b200.vti import read_vti, write_vti_ascii
from oracle import Oracle


@pytest.mark.parametrize("precision", ["f32", "f64"])
def test_ascii_vti_round_trip(tmp_path, precision):
    dim = 8
    run = Oracle(precision).run(dim, 8, 0.0089, 0.05, 3, 1)
    path = str(tmp_path / "lbmcl.3.vti")
    write_vti_ascii(path, run["rho"][3], run["u"][3], dim)
    d = read_vti(path)
    assert d["dims"] == (6, 6, 6) and d["extent"] == (0, 5, 0, 5, 0, 5)
    assert d["arrays"]["rho"].dtype == run["rho"].dtype
    # 17 significant digits round-trip both float32 and float64 exactly
    assert d["arrays"]["rho"].tobytes() == np.ascontiguousarray(wet(run["rho"][3], dim)).tobytes()
    v = np.moveaxis(wet(run["u"][3], dim), 0, -1).reshape(-1, 3)
    assert d["arrays"]["v"].tobytes() == np.ascontiguousarray(v).tobytes()


def _write_run(tmp_path, dim, stride, its, every, perturb=0.0):
    run = Oracle("f32").run(dim, stride, 0.0089, 0.05, its, every)
    width = len(str(its))
    for k in range(its // every + 1):
        rho = run["rho"][k].copy()
        if perturb and k > 0:
            rho[np.isfinite(rho)] += np.float32(perturb)
        write_vti_ascii(str(tmp_path / f"lbmcl.{k * every:0{width}d}.vti"), rho, run["u"][k], dim)


def test_verify_passes_on_oracle_output(tmp_path, capsys):
    _write_run(tmp_path, 8, 8, 10, 1)
    rc = verify.main(["-i", "10", "-e", "1", "-t", os.path.join(GOLDEN, "target8.npz"), "-p", str(tmp_path), "--check"])
    out = capsys.readouterr().out
    assert rc == 0 and "verify: PASS (11 iterations compared)" in out
    assert out.splitlines()[0].split() == ["#it", "MSE_RHO", "MSE_U", "MAE_RHO", "MAE_U"]   # verify.py:32-33
    assert out.splitlines()[1].startswith(" 0:  0.000000e+00   0.000000e+00   0.000000e+00   0.000000e+00")


def test_verify_fails_on_wrong_output(tmp_path, capsys):
    _write_run(tmp_path, 8, 8, 10, 1, perturb=1e-4)
    rc = verify.main(["-i", "10", "-e", "1", "-t", os.path.join(GOLDEN, "target8.npz"), "-p", str(tmp_path), "--check"])
    assert rc == 1 and "FAIL" in capsys.readouterr().out
    # without --check the reference's behaviour: print the table, exit 0
    assert verify.main(["-i", "10", "-e", "1", "-t", os.path.join(GOLDEN, "target8.npz"), "-p", str(tmp_path)]) == 0


@pytest.mark.skipif(not os.path.isdir("/root/reference/target_results/8"), reason="reference tree absent")
def test_verify_reads_the_reference_vti_directory(tmp_path, capsys):
    _write_run(tmp_path, 8, 8, 10, 1)
    rc = verify.main(["-i", "10", "-e", "1", "-t", "/root/reference/target_results/8", "-p", str(tmp_path), "--check"])
    assert rc == 0 and "PASS" in capsys.readouterr().out
