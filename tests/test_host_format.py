"""The VTI writer's number formatting (lbmcl_b200/host/fmt_e16.hpp): the hand-written "%.16e" must print the
same bytes as printf for every double -- the VTI files stay byte-identical to what the reference's
`std::scientific << std::setprecision(16)` writes (lbmcl.hpp:289-333)."""
import os
import subprocess

from conftest import ROOT


def test_fmt_e16_is_byte_identical_to_printf(tmp_path):
    exe = str(tmp_path / "format_check")
    subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++11", "-Wall", "-Wextra", "-o", exe,
                           os.path.join(ROOT, "tests", "cpp", "format_check.cpp")])
    r = subprocess.run([exe, "600000"], stdout=subprocess.PIPE, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:]
    assert " 0 mismatches" in r.stdout


def test_parallel_vti_writer_is_byte_identical_to_the_reference_writer(tmp_path):
    """lbmcl_b200/host/vti_writer.hpp (workers format planes and pwrite them at computed offsets) against a plain
    restatement of the reference's single-ofstream writer (lbmcl.hpp:261-334), for 1..32 threads, fp32 / fp64,
    NaN / zero / tiny / negative values."""
    exe = str(tmp_path / "vti_writer_check")
    subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++11", "-Wall", "-Wextra", "-pthread", "-o", exe,
                           os.path.join(ROOT, "tests", "cpp", "vti_writer_check.cpp")])
    r = subprocess.run([exe, str(tmp_path)], stdout=subprocess.PIPE, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:]
    assert "all files byte-identical" in r.stdout
