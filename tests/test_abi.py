"""The drop-in boundary on a box without a GPU: the C-ABI library loads, exports every symbol that
include/lbm_b200.h declares, validates its arguments, and refuses to run without a CUDA device (no
CPU fallback).  No compute call is made here."""
import ctypes
import os
import re

import pytest

from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "lbm_b200.h")


def _declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lbm_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from lbmcl_b200 import capi
    lib = capi.load()
    names = _declared_functions()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"liblbm_b200.so does not export {n}"
    assert sorted(capi.EXPORTS) == names, "capi.EXPORTS and include/lbm_b200.h disagree"


def test_default_params_are_the_reference_defaults():
    from lbmcl_b200 import capi
    p = capi.LbmParams()
    capi.load().lbm_default_params(ctypes.byref(p))
    # lbm_options.hpp:32-50
    assert (p.dim, p.viscosity, p.velocity, p.stride) == (8, 0.0089, 0.05, 32)
    assert (p.block_x, p.block_y, p.block_z) == (8, 8, 8)
    assert p.precision == capi.F32 and p.fast_math == 0 and p.abi_version == capi.ABI_VERSION
    assert ctypes.sizeof(capi.LbmParams) == 104


@pytest.mark.parametrize("kw,fragment", [
    (dict(dim=12), "power of two"),
    (dict(dim=2), "power of two"),
    (dict(dim=8, stride=24), "stride"),
    (dict(dim=8, stride=1024), "stride"),          # > dim^3: the reference would overrun its buffers
    (dict(dim=8, viscosity=-1.0), "viscosity"),
    (dict(dim=8, z_range=(4, 2)), "z range"),
    (dict(dim=8, variant=3), "variant"),
])
def test_invalid_configurations_are_rejected(kw, fragment):
    from lbmcl_b200.capi import LbmError, Simulation
    with pytest.raises(LbmError) as e:
        Simulation(**kw)
    assert e.value.code == -1 and fragment in str(e.value)


def test_abi_version_is_checked():
    from lbmcl_b200 import capi
    p = capi.make_params(dim=8)
    p.abi_version = 99
    h = ctypes.c_void_p()
    assert capi.load().lbm_create(ctypes.byref(p), ctypes.byref(h)) == -1
    assert b"abi_version" in capi.load().lbm_last_error(None)


def test_no_cpu_fallback():
    """Without a CUDA device the product path must fail loudly, never compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from lbmcl_b200.capi import LbmError, Simulation
    with pytest.raises(LbmError) as e:
        Simulation(dim=8)
    assert e.value.code == -3 and "no CPU path" in str(e.value)


def test_product_path_does_not_touch_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may use oracle/."""
    pkg = os.path.join(ROOT, "lbmcl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".cpp", ".inl", ".h")) or f == "Makefile":
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in text and "from oracle import" not in text, f
                assert "liblbm_oracle" not in text and "oracle/_ref" not in text, f


def test_specialised_kernel_source_builds_without_a_gpu():
    """LBM_VARIANT_NVRTC's translation unit (the embedded kernel headers + -D options, the counterpart of the
    reference's run-time OpenCL build, lbmcl.hpp:131-156) compiles for sm_100a on a GPU-less box."""
    from lbmcl_b200 import capi
    for kw in (dict(dim=256, stride=32, precision="f32"), dict(dim=64, stride=4096, precision="f64", fast_math=True),
               dict(dim=16, stride=4096, precision="f32")):
        cubin = capi.spec_cubin(**kw)
        assert cubin[:4] == b"\x7fELF" and len(cubin) > 10000
        assert b"lbm_step_spec_m0" in cubin and b"lbm_step_spec_m1" in cubin


def test_compile_time_alternatives_of_the_kernels_still_build(tmp_path):
    """The measured-and-not-adopted forms stay in the sources behind compile-time knobs (profiles/r02_*.md): the TMA
    kernel with per-warp bulk-tensor stores, other register budgets.  They must keep compiling for sm_100a (nvcc
    cross-compiles without a GPU) so that the A/B builds of the profiles can be reproduced."""
    import shutil
    import subprocess
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    src = os.path.join(ROOT, "lbmcl_b200", "csrc")
    base = [nvcc, "-std=c++17", "-O1", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC", "-c"]
    for name, unit, flags in (
            ("tma_bulk_store", "lbm_launch_tma.cu", ["-DLBM_TMA_DIRECT_STORE=0"]),
            ("aa_shift_40_registers", "lbm_launch_aa.cu", ["-DLBM_MINB_AA_SHIFT_F32=6"])):
        r = subprocess.run(base + flags + ["-o", str(tmp_path / (name + ".o")), os.path.join(src, unit)],
                           stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
        assert r.returncode == 0, f"{name}: {r.stdout[-3000:]}"
    out = subprocess.run(["cuobjdump", "-sass", str(tmp_path / "tma_bulk_store.o")], stdout=subprocess.PIPE, text=True).stdout
    assert "UTMALDG" in out and "UTMASTG" in out       # bulk-tensor loads AND stores in that build
    lib = subprocess.run(["cuobjdump", "-sass", os.path.join(src, "liblbm_b200.so")], stdout=subprocess.PIPE, text=True).stdout
    assert "UTMALDG" in lib and "SYNCS" in lib         # the shipped TMA kernel: bulk-tensor loads + mbarriers
