"""The oracle is pinned before it is trusted (CPU only).

(1) against the reference's golden fixtures: target_results/8 (11 files) and target_results/32
    (committed subset in tests/golden/, all 26 in place when /root/reference is present), with the
    comparator of verify.py:47-56 and the magnitudes recorded in SURVEY.md Appendix A;
(2) bit for bit -- rho, u, map and both complete f lattices -- against oracle/_ref, the reference's
    unmodified kernels.cl compiled as host C++.
"""
import glob
import os

import numpy as np
import pytest

from conftest import wet
from oracle import Oracle, RefKernel, ref_available

REF_ROOT = os.environ.get("LBMCL_REF", "/root/reference")
NU, U = 0.0089, 0.05


def _maxabs(a, b):
    return float(np.max(np.abs(np.nan_to_num(a).astype(np.float64) - np.nan_to_num(b).astype(np.float64))))


def test_oracle_vs_target8(golden8):
    run = Oracle("f32").run(8, 8, NU, U, 10, 1)
    for k, it in enumerate(golden8["its"]):
        p_rho = wet(run["rho"][it], 8)
        p_v = np.moveaxis(wet(run["u"][it], 8), 0, -1)
        assert np.array_equal(np.isnan(p_rho), np.isnan(golden8["rho"][k]))
        assert np.array_equal(np.isnan(p_v), np.isnan(golden8["v"][k]))
        # Appendix A: it 10 -> MAE_rho 5.4e-7, MAE_u 1.9e-7; it 0 exact
        assert _maxabs(p_rho, golden8["rho"][k]) <= (0.0 if it == 0 else 6e-7)
        assert _maxabs(p_v, golden8["v"][k]) <= (0.0 if it == 0 else 2e-7)
    assert np.count_nonzero(~np.isnan(golden8["rho"][0])) == 80  # 64 fluid + 16 lid cells


def test_oracle_vs_target32(golden32):
    run = Oracle("f32").run(32, 32, NU, U, 500, 20)
    for k, it in enumerate(golden32["its"]):
        p_rho = wet(run["rho"][it // 20], 32)
        p_v = np.moveaxis(wet(run["u"][it // 20], 32), 0, -1)
        assert np.array_equal(np.isnan(p_rho), np.isnan(golden32["rho"][k]))
        # Appendix A: it 500 -> MAE_rho 2.0e-6, MAE_u 5.2e-7 (= 1.04e-5 * U)
        assert _maxabs(p_rho, golden32["rho"][k]) <= 2.5e-6, it
        assert _maxabs(p_v, golden32["v"][k]) <= 1.2e-5 * U, it
    assert np.count_nonzero(~np.isnan(golden32["rho"][0])) == 22736  # 28^3 + 28^2


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF_ROOT, "target_results")), reason="reference tree absent")
def test_oracle_vs_every_reference_fixture_in_place():
    """All 37 .vti fixtures of the reference, read where they lie."""
    from lbmcl_b200.vti import read_vti
    for dim, stride, its, every, width in ((8, 8, 10, 1, 2), (32, 32, 500, 20, 3)):
        run = Oracle("f32").run(dim, stride, NU, U, its, every)
        files = sorted(glob.glob(os.path.join(REF_ROOT, "target_results", str(dim), "ldc.0.*.vti")))
        assert len(files) == its // every + 1
        for k, path in enumerate(files):
            it = int(os.path.basename(path).split(".")[2])
            assert it == k * every
            d = read_vti(path)
            n = dim - 2
            assert d["dims"] == (n, n, n)
            t_rho = d["arrays"]["rho"].reshape(n, n, n)
            t_v = d["arrays"]["v"].reshape(n, n, n, 3)
            p_rho = wet(run["rho"][k], dim)
            p_v = np.moveaxis(wet(run["u"][k], dim), 0, -1)
            assert np.array_equal(np.isnan(p_rho), np.isnan(t_rho))
            assert _maxabs(p_rho, t_rho) <= 2.5e-6
            assert _maxabs(p_v, t_v) <= 1.2e-5 * U


def test_golden_archives_match_reference_files_in_place(golden8, golden32):
    """The committed .npz archives are bit-for-bit the payload of the reference's .vti files."""
    if not os.path.isdir(os.path.join(REF_ROOT, "target_results")):
        pytest.skip("reference tree absent")
    from lbmcl_b200.vti import read_vti
    for arch, sub, width in ((golden8, "8", 2), (golden32, "32", 3)):
        for k, it in enumerate(arch["its"]):
            d = read_vti(os.path.join(REF_ROOT, "target_results", sub, f"ldc.0.{int(it):0{width}d}.vti"))
            assert d["arrays"]["rho"].tobytes() == arch["rho"][k].tobytes()
            assert d["arrays"]["v"].tobytes() == arch["v"][k].tobytes()


REF_CASES = [(8, 8, 10, 1), (8, 32, 10, 1), (16, 16, 12, 3), (32, 32, 40, 20), (32, 8, 10, 1), (64, 32, 4, 2)]


@pytest.mark.parametrize("precision", ["f32", "f64"])
@pytest.mark.parametrize("dim,stride,its,every", REF_CASES)
def test_oracle_bitwise_equals_reference_kernel(precision, dim, stride, its, every):
    if not ref_available(precision, dim, stride):
        pytest.skip("oracle/_ref not built for this configuration")
    ref = RefKernel(precision, dim, stride)
    a = Oracle(precision).run(dim, stride, ref.viscosity, ref.velocity, its, every, keep_state=True)
    b = ref.run(its, every, keep_state=True)
    assert np.array_equal(a["map"], b["map"])
    assert a["rho"].tobytes() == b["rho"].tobytes()
    assert a["u"].tobytes() == b["u"].tobytes()
    for k in ("f_stream", "f_collide"):
        assert a["state"][k].tobytes() == b["state"][k].tobytes(), k


def test_effective_parameters_follow_the_text_round_trip():
    """lbmcl.hpp:140-141 prints nu and U with 6 significant digits; the kernel compiler reads that."""
    p = Oracle("f32").params(0.0089, 0.05)
    assert p["viscosity"] == float(np.float32(0.0089)) and p["velocity"] == float(np.float32(0.05))
    assert p["inv_tau"] == float(np.float32(1.0) / (np.float32(3.0) * np.float32(0.0089) + np.float32(0.5)))
    p = Oracle("f64").params(0.0123456789, 0.05)
    assert p["viscosity"] == 0.0123457
    assert abs(Oracle("f64").params(0.0089, 0.05)["inv_tau"] - 1.89861401177140698415) < 1e-15  # kernels.cl:62


def test_cell_census_8():
    """FLUID 64, MOVING 16, CORNER 4, WALL 296 at 8^3."""
    m = Oracle("f32").cell_map(8)
    assert np.count_nonzero(m == 0x1) == 64
    assert np.count_nonzero((m & 0x2) != 0) == 16 and np.all(m[(m & 0x2) != 0] == 0x102)
    assert np.count_nonzero(m == 0x4) == 4
    assert np.count_nonzero(m == 0x8) == 296


def test_output_labels_lag_by_one_step():
    """file 1 holds the macros of the state read by iteration 1 = the initial state."""
    run = Oracle("f64").run(8, 8, NU, U, 2, 1)
    assert np.allclose(np.nan_to_num(run["rho"][0]), np.nan_to_num(run["rho"][1]), atol=1e-15)
    assert np.allclose(np.nan_to_num(run["u"][0]), np.nan_to_num(run["u"][1]), atol=1e-15)
    assert not np.array_equal(np.nan_to_num(run["rho"][1]), np.nan_to_num(run["rho"][2]))
