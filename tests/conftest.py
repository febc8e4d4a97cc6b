"""pytest configuration: the `gpu` marker and shared fixtures.

`-m "not gpu"` runs here on a CPU-only box: the oracle against the reference's golden vectors and
against oracle/_ref, the host logic, and the C-ABI export checks.
`-m gpu` runs on a B200: the parity tests proper, all through the C ABI (liblbm_b200.so).
Nothing under -m gpu reads /root/reference.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Make sure the oracle library exists (cheap; the CUDA library is built by build())."""
    if not os.path.exists(os.path.join(ROOT, "oracle", "liblbm_oracle.so")):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "liblbm_oracle.so"])


@pytest.fixture(scope="session")
def golden8():
    return np.load(os.path.join(GOLDEN, "target8.npz"))


@pytest.fixture(scope="session")
def golden32():
    return np.load(os.path.join(GOLDEN, "target32.npz"))


def wet(field, dim):
    """[.., N] (N = dim^3, x fastest) -> [.., dim-2, dim-2, dim-2] (z, y, x) over the wet cube, the
    region the reference writes to VTI (lbmcl.hpp:289-317) and the Sailfish targets cover."""
    f = np.asarray(field).reshape(field.shape[:-1] + (dim, dim, dim))
    return f[..., 1:dim - 1, 1:dim - 1, 1:dim - 1]
