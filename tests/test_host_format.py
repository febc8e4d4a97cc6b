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
