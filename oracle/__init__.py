"""TEST INFRASTRUCTURE ONLY: ctypes access to the CPU oracle.

Two checkers live here, neither of which the product path (lbmcl_b200/) may import:

* ``Oracle``    -- ``liblbm_oracle.so``, the CPU restatement (oracle/lbm_oracle.c), run-time
                   dim / stride / nu / U, fp32 and fp64.
* ``RefKernel`` -- ``_ref/libref_<p>_d<DIM>_s<STRIDE>.so``, the reference's own unmodified kernels.cl
                   compiled as host C++ (oracle/ref_shim.cpp); one library per configuration
                   because the reference takes DIM / stride / nu / U as -D macros
                   (reference lbmcl.hpp:131-156).

Both are built by ``make -C oracle`` (also run by ``__graft_entry__.build()``).
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_I64 = ctypes.c_int64


def _np_ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def host_cores() -> int:
    """Cores this process may run on (affinity-aware)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def n_snapshots(iterations: int, every: int) -> int:
    """Number of rho/u snapshots the reference host writes (lbmcl.hpp:502, 513-515)."""
    return 0 if every == 0 else 1 + iterations // every


class Oracle:
    """The CPU restatement.  ``precision`` is 'f32' or 'f64'."""

    _lib = None

    def __init__(self, precision: str = "f32"):
        assert precision in ("f32", "f64")
        self.precision = precision
        self.dtype = np.float32 if precision == "f32" else np.float64
        if Oracle._lib is None:
            path = os.path.join(HERE, "liblbm_oracle.so")
            if not os.path.exists(path):
                raise RuntimeError(f"{path} missing: run `make -C oracle` (or __graft_entry__.build())")
            Oracle._lib = ctypes.CDLL(path)
        lib = Oracle._lib
        s = "_" + precision
        self._params = getattr(lib, "lbm_oracle_params" + s)
        self._init = getattr(lib, "lbm_oracle_init" + s)
        self._step = getattr(lib, "lbm_oracle_step" + s)
        self._run = getattr(lib, "lbm_oracle_run" + s)
        vp = ctypes.c_void_p
        self._params.argtypes = [ctypes.c_double, ctypes.c_double, vp]
        self._params.restype = None
        self._init.argtypes = [ctypes.c_int, _I64, ctypes.c_double, ctypes.c_double, vp, vp, vp, vp, vp]
        self._init.restype = None
        self._step.argtypes = [ctypes.c_int, _I64, ctypes.c_double, ctypes.c_double, vp, vp, vp, vp, vp,
                               ctypes.c_int]
        self._step.restype = None
        self._run.argtypes = [ctypes.c_int, _I64, ctypes.c_double, ctypes.c_double, ctypes.c_int, ctypes.c_int,
                              vp, vp, vp, vp, vp, vp, vp]
        self._run.restype = ctypes.c_int
        lib.lbm_oracle_map.argtypes = [ctypes.c_int, vp]
        lib.lbm_oracle_map.restype = None
        lib.lbm_oracle_set_threads.argtypes = [ctypes.c_int]
        lib.lbm_oracle_set_threads.restype = ctypes.c_int

    def set_threads(self, n: int) -> int:
        """Ask for n OpenMP threads (overrides OMP_NUM_THREADS); returns the number in effect."""
        return Oracle._lib.lbm_oracle_set_threads(n)

    def params(self, nu: float, u_lid: float):
        out = np.zeros(3, dtype=np.float64)
        self._params(nu, u_lid, _np_ptr(out))
        return {"viscosity": out[0], "velocity": out[1], "inv_tau": out[2]}

    def cell_map(self, dim: int) -> np.ndarray:
        m = np.zeros(dim ** 3, dtype=np.int32)
        Oracle._lib.lbm_oracle_map(dim, _np_ptr(m))
        return m

    def alloc(self, dim: int):
        n = dim ** 3
        return {
            "f_stream": np.zeros(19 * n, dtype=self.dtype),
            "f_collide": np.zeros(19 * n, dtype=self.dtype),
            "rho": np.zeros(n, dtype=self.dtype),
            "u": np.zeros(3 * n, dtype=self.dtype),
            "map": np.zeros(n, dtype=np.int32),
        }

    def init(self, st, dim, stride, nu, u_lid):
        self._init(dim, stride, nu, u_lid, _np_ptr(st["f_stream"]), _np_ptr(st["f_collide"]),
                   _np_ptr(st["rho"]), _np_ptr(st["u"]), _np_ptr(st["map"]))

    def step(self, st, dim, stride, nu, u_lid, it: int, every: int):
        """Iteration ``it`` (1-based) with the reference's ping-pong and macro flag."""
        flag = 1 if (every != 0 and it % every == 0) else 0
        swap = it % 2 == 0
        dst = st["f_collide"] if swap else st["f_stream"]
        src = st["f_stream"] if swap else st["f_collide"]
        self._step(dim, stride, nu, u_lid, _np_ptr(dst), _np_ptr(src), _np_ptr(st["rho"]), _np_ptr(st["u"]),
                   _np_ptr(st["map"]), flag)

    def run(self, dim, stride, nu, u_lid, iterations, every, keep_state: bool = False):
        """Full reference schedule.  Returns dict with ``rho[k, N]``, ``u[k, 3, N]`` snapshots
        (k = 0 is the initial field, then every flagged iteration) and, if ``keep_state``, the
        final buffers."""
        n = dim ** 3
        st = self.alloc(dim)
        k = n_snapshots(iterations, every)
        snap_rho = np.zeros((max(k, 1), n), dtype=self.dtype)
        snap_u = np.zeros((max(k, 1), 3, n), dtype=self.dtype)
        got = self._run(dim, stride, nu, u_lid, iterations, every, _np_ptr(st["f_stream"]),
                        _np_ptr(st["f_collide"]), _np_ptr(st["rho"]), _np_ptr(st["u"]), _np_ptr(st["map"]),
                        _np_ptr(snap_rho), _np_ptr(snap_u))
        assert got == k, (got, k)
        out = {"rho": snap_rho[:k], "u": snap_u[:k], "map": st["map"]}
        if keep_state:
            out["state"] = st
        return out


def ref_lib_path(precision: str, dim: int, stride: int) -> str:
    return os.path.join(HERE, "_ref", f"libref_{precision}_d{dim}_s{stride}.so")


def ref_available(precision: str, dim: int, stride: int) -> bool:
    return os.path.exists(ref_lib_path(precision, dim, stride))


class RefKernel:
    """The reference's own kernels.cl (compiled as host C++) for one fixed configuration."""

    def __init__(self, precision: str, dim: int, stride: int):
        path = ref_lib_path(precision, dim, stride)
        if not os.path.exists(path):
            raise RuntimeError(f"{path} missing: run `make -C oracle ref` where /root/reference exists")
        self.lib = ctypes.CDLL(path)
        self.precision = precision
        self.dtype = np.float32 if precision == "f32" else np.float64
        self.dim = dim
        self.stride = stride
        assert self.lib.ref_dim() == dim and self.lib.ref_stride() == stride
        assert self.lib.ref_sizeof_real() == np.dtype(self.dtype).itemsize
        vp = ctypes.c_void_p
        self.lib.ref_viscosity.restype = ctypes.c_double
        self.lib.ref_velocity.restype = ctypes.c_double
        self.lib.ref_initialize.argtypes = [vp] * 5
        self.lib.ref_initialize.restype = None
        self.lib.ref_compute.argtypes = [vp] * 5 + [ctypes.c_int]
        self.lib.ref_compute.restype = None
        self.lib.ref_run.argtypes = [vp] * 5 + [ctypes.c_int, ctypes.c_int, vp, vp]
        self.lib.ref_run.restype = ctypes.c_int
        self.lib.ref_set_threads.argtypes = [ctypes.c_int]
        self.lib.ref_set_threads.restype = ctypes.c_int

    def set_threads(self, n: int) -> int:
        """Ask for n OpenMP threads (overrides OMP_NUM_THREADS); returns the number in effect."""
        return self.lib.ref_set_threads(n)

    @property
    def viscosity(self) -> float:
        return self.lib.ref_viscosity()

    @property
    def velocity(self) -> float:
        return self.lib.ref_velocity()

    def alloc(self):
        n = self.dim ** 3
        return {
            "f_stream": np.zeros(19 * n, dtype=self.dtype),
            "f_collide": np.zeros(19 * n, dtype=self.dtype),
            "rho": np.zeros(n, dtype=self.dtype),
            "u": np.zeros(3 * n, dtype=self.dtype),
            "map": np.zeros(n, dtype=np.int32),
        }

    def init(self, st):
        self.lib.ref_initialize(_np_ptr(st["f_stream"]), _np_ptr(st["f_collide"]), _np_ptr(st["rho"]),
                                _np_ptr(st["u"]), _np_ptr(st["map"]))

    def step(self, st, it: int, every: int):
        flag = 1 if (every != 0 and it % every == 0) else 0
        swap = it % 2 == 0
        dst = st["f_collide"] if swap else st["f_stream"]
        src = st["f_stream"] if swap else st["f_collide"]
        self.lib.ref_compute(_np_ptr(dst), _np_ptr(src), _np_ptr(st["rho"]), _np_ptr(st["u"]),
                             _np_ptr(st["map"]), flag)

    def run(self, iterations: int, every: int, keep_state: bool = False):
        n = self.dim ** 3
        st = self.alloc()
        k = n_snapshots(iterations, every)
        snap_rho = np.zeros((max(k, 1), n), dtype=self.dtype)
        snap_u = np.zeros((max(k, 1), 3, n), dtype=self.dtype)
        got = self.lib.ref_run(_np_ptr(st["f_stream"]), _np_ptr(st["f_collide"]), _np_ptr(st["rho"]),
                               _np_ptr(st["u"]), _np_ptr(st["map"]), iterations, every, _np_ptr(snap_rho),
                               _np_ptr(snap_u))
        assert got == k, (got, k)
        out = {"rho": snap_rho[:k], "u": snap_u[:k], "map": st["map"]}
        if keep_state:
            out["state"] = st
        return out
