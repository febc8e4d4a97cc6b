"""ctypes binding of the C ABI in include/lbm_b200.h (liblbm_b200.so).

This is the reference-side stub a Python host would use; the tests and bench.py
call the CUDA path through it.  There is no fallback: if the shared library is missing the import
of :class:`Library` fails loudly, and if no CUDA device is present ``lbm_create`` fails.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LBM_B200_LIB") or os.path.join(HERE, "csrc", "liblbm_b200.so")

ABI_VERSION = 2
F32, F64 = 0, 1
VARIANT_AUTO, VARIANT_SCALAR, VARIANT_VEC2, VARIANT_VEC4, VARIANT_AA, VARIANT_TMA, VARIANT_NVRTC = 0, 1, 2, 4, 8, 16, 32
FUSED_OFF, FUSED_FLAGS, FUSED_TOKEN = 0, 1, 2

# every symbol include/lbm_b200.h declares (tests check that the library exports all of them)
EXPORTS = [
    "lbm_default_params", "lbm_create", "lbm_destroy", "lbm_last_error", "lbm_init", "lbm_step", "lbm_run",
    "lbm_sync", "lbm_read_macros", "lbm_read_macros_slab", "lbm_host_alloc", "lbm_host_free", "lbm_read_macros_async",
    "lbm_read_wait", "lbm_mark_end", "lbm_read_map", "lbm_read_f", "lbm_time_ms", "lbm_launch_times_ms", "lbm_device_name",
    "lbm_effective_params", "lbm_block_shape", "lbm_device_bytes", "lbm_spec_cubin", "lbm_launch_count", "lbm_iteration",
    "lbm_set_stream", "lbm_step_planes", "lbm_advance", "lbm_z_range", "lbm_halo_elems", "lbm_halo_send_buffer",
    "lbm_halo_recv_buffer", "lbm_halo_pack", "lbm_halo_unpack", "lbm_comm_unique_id", "lbm_comm_init", "lbm_ipc_export",
    "lbm_ipc_attach", "lbm_peer_attach", "lbm_ipc_detach", "lbm_comm_fused", "lbm_group_create", "lbm_group_destroy",
    "lbm_group_last_error", "lbm_group_size", "lbm_group_ctx", "lbm_group_init", "lbm_group_run", "lbm_group_sync",
    "lbm_group_read_macros", "lbm_group_read_f", "lbm_group_time_ms",
]


class LbmParams(ctypes.Structure):
    _fields_ = [
        ("abi_version", ctypes.c_int32),
        ("dim", ctypes.c_int32),
        ("precision", ctypes.c_int32),
        ("fast_math", ctypes.c_int32),
        ("viscosity", ctypes.c_double),
        ("velocity", ctypes.c_double),
        ("stride", ctypes.c_int64),
        ("block_x", ctypes.c_int32),
        ("block_y", ctypes.c_int32),
        ("block_z", ctypes.c_int32),
        ("device", ctypes.c_int32),
        ("variant", ctypes.c_int32),
        ("z_begin", ctypes.c_int32),
        ("z_end", ctypes.c_int32),
        ("reserved", ctypes.c_int32 * 8),
    ]


class LbmError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"lbm_b200 error {code}: {message}")
        self.code = code


_lib = None


def load() -> ctypes.CDLL:
    """Load liblbm_b200.so (once) and declare the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} not found: build it with `make -C lbmcl_b200/csrc` "
                          "(or __graft_entry__.build()); there is no CPU fallback")
    lib = ctypes.CDLL(LIB_PATH)
    vp, ci, i64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64
    lib.lbm_default_params.argtypes = [ctypes.POINTER(LbmParams)]
    lib.lbm_default_params.restype = None
    lib.lbm_create.argtypes = [ctypes.POINTER(LbmParams), ctypes.POINTER(vp)]
    lib.lbm_create.restype = ci
    lib.lbm_destroy.argtypes = [vp]
    lib.lbm_destroy.restype = None
    lib.lbm_last_error.argtypes = [vp]
    lib.lbm_last_error.restype = ctypes.c_char_p
    lib.lbm_init.argtypes = [vp]
    lib.lbm_step.argtypes = [vp, ci]
    lib.lbm_run.argtypes = [vp, ci, ci]
    lib.lbm_sync.argtypes = [vp]
    lib.lbm_read_macros.argtypes = [vp, vp, vp]
    lib.lbm_read_macros_slab.argtypes = [vp, vp, vp]
    lib.lbm_host_alloc.argtypes = [ctypes.c_size_t, ctypes.POINTER(vp)]
    lib.lbm_host_free.argtypes = [vp]
    lib.lbm_host_free.restype = None
    lib.lbm_read_macros_async.argtypes = [vp, vp, vp]
    lib.lbm_read_wait.argtypes = [vp]
    lib.lbm_mark_end.argtypes = [vp]
    lib.lbm_spec_cubin.argtypes = [ctypes.POINTER(LbmParams), vp, ctypes.c_size_t, ctypes.POINTER(ctypes.c_size_t)]
    lib.lbm_read_map.argtypes = [vp, vp]
    lib.lbm_read_f.argtypes = [vp, vp]
    lib.lbm_time_ms.argtypes = [vp, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]
    lib.lbm_launch_times_ms.argtypes = [vp, vp, i64, ctypes.POINTER(i64)]
    lib.lbm_device_name.argtypes = [vp, ctypes.c_char_p, ctypes.c_size_t]
    lib.lbm_effective_params.argtypes = [vp, ctypes.POINTER(ctypes.c_double)]
    lib.lbm_block_shape.argtypes = [vp, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int32)]
    lib.lbm_device_bytes.argtypes = [vp]
    lib.lbm_device_bytes.restype = i64
    lib.lbm_launch_count.argtypes = [vp]
    lib.lbm_launch_count.restype = i64
    lib.lbm_iteration.argtypes = [vp]
    lib.lbm_iteration.restype = i64
    lib.lbm_set_stream.argtypes = [vp, vp]
    lib.lbm_step_planes.argtypes = [vp, ci, ci, ci]
    lib.lbm_advance.argtypes = [vp]
    lib.lbm_z_range.argtypes = [vp, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int32)]
    lib.lbm_halo_elems.argtypes = [vp]
    lib.lbm_halo_elems.restype = i64
    lib.lbm_halo_send_buffer.argtypes = [vp, ci]
    lib.lbm_halo_send_buffer.restype = vp
    lib.lbm_halo_recv_buffer.argtypes = [vp, ci]
    lib.lbm_halo_recv_buffer.restype = vp
    lib.lbm_halo_pack.argtypes = [vp]
    lib.lbm_halo_unpack.argtypes = [vp]
    lib.lbm_comm_unique_id.argtypes = [vp]
    lib.lbm_comm_init.argtypes = [vp, vp, ci, ci]
    lib.lbm_ipc_export.argtypes = [vp, vp]
    lib.lbm_ipc_attach.argtypes = [vp, ci, vp]
    lib.lbm_peer_attach.argtypes = [vp, ci, vp]
    lib.lbm_ipc_detach.argtypes = [vp]
    lib.lbm_comm_fused.argtypes = [vp, ci]
    lib.lbm_group_create.argtypes = [ctypes.POINTER(LbmParams), ctypes.POINTER(ctypes.c_int32), ci,
                                     ctypes.POINTER(vp)]
    lib.lbm_group_destroy.argtypes = [vp]
    lib.lbm_group_destroy.restype = None
    lib.lbm_group_last_error.argtypes = [vp]
    lib.lbm_group_last_error.restype = ctypes.c_char_p
    lib.lbm_group_size.argtypes = [vp]
    lib.lbm_group_ctx.argtypes = [vp, ci]
    lib.lbm_group_ctx.restype = vp
    lib.lbm_group_init.argtypes = [vp]
    lib.lbm_group_run.argtypes = [vp, ci, ci]
    lib.lbm_group_sync.argtypes = [vp]
    lib.lbm_group_read_macros.argtypes = [vp, vp, vp]
    lib.lbm_group_read_f.argtypes = [vp, vp]
    lib.lbm_group_time_ms.argtypes = [vp, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]
    _lib = lib
    return lib


def make_params(dim=8, precision="f32", viscosity=0.0089, velocity=0.05, stride=32, block=(8, 8, 8),
                fast_math=False, device=-1, variant=VARIANT_AUTO, z_range=None, exact_block=False,
                generic_addressing=False) -> LbmParams:
    lib = load()
    p = LbmParams()
    lib.lbm_default_params(ctypes.byref(p))
    p.dim = dim
    if precision in ("f32", "single", "float"):
        p.precision = F32
    elif precision in ("f64", "double"):
        p.precision = F64
    else:
        raise ValueError(f"precision {precision!r}: expected 'f32' or 'f64'")
    p.fast_math = 1 if fast_math else 0
    p.viscosity = viscosity
    p.velocity = velocity
    p.stride = stride
    p.block_x, p.block_y, p.block_z = block
    p.device = device
    p.variant = variant
    if z_range is not None:
        p.z_begin, p.z_end = z_range
    p.reserved[0] = 1 if generic_addressing else 0
    p.reserved[1] = 1 if exact_block else 0
    return p


def _ptr(a):
    return None if a is None else ctypes.c_void_p(a.ctypes.data)


class Simulation:
    """One context (= one device, one z-slab).  Mirrors the calls LBMCL<T> makes (lbmcl.hpp)."""

    def __init__(self, **kw):
        self.lib = load()
        self.params = make_params(**kw)
        self.dim = self.params.dim
        self.dtype = np.float32 if self.params.precision == F32 else np.float64
        h = ctypes.c_void_p()
        rc = self.lib.lbm_create(ctypes.byref(self.params), ctypes.byref(h))
        if rc != 0:
            raise LbmError(rc, self.lib.lbm_last_error(None).decode())
        self.h = h

    def _check(self, rc):
        if rc != 0:
            raise LbmError(rc, self.lib.lbm_last_error(self.h).decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.lbm_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def init(self):
        self._check(self.lib.lbm_init(self.h))

    def step(self, update_macro: bool):
        self._check(self.lib.lbm_step(self.h, 1 if update_macro else 0))

    def run(self, n_iterations: int, every: int = 0):
        self._check(self.lib.lbm_run(self.h, n_iterations, every))

    def sync(self):
        self._check(self.lib.lbm_sync(self.h))

    def set_stream(self, cuda_stream_ptr):
        self._check(self.lib.lbm_set_stream(self.h, ctypes.c_void_p(cuda_stream_ptr)))

    def read_macros(self, rho=None, u=None):
        """rho[N], u[3, N] in the reference's global layouts.  With both arguments omitted fresh arrays
        are allocated; if only one array is given, only that field is read."""
        n = self.dim ** 3
        if rho is None and u is None:
            rho = np.full(n, np.nan, dtype=self.dtype)
            u = np.full((3, n), np.nan, dtype=self.dtype)
        self._check(self.lib.lbm_read_macros(self.h, _ptr(rho), _ptr(u)))
        return rho, u

    def read_macros_slab(self, rho=None, u=None):
        """rho[P], u[3, P] over the owned planes only (P = planes * dim^2)."""
        z0, z1 = self.z_range
        p = (z1 - z0) * self.dim * self.dim
        if rho is None:
            rho = np.full(p, np.nan, dtype=self.dtype)
        if u is None:
            u = np.full((3, p), np.nan, dtype=self.dtype)
        self._check(self.lib.lbm_read_macros_slab(self.h, _ptr(rho), _ptr(u)))
        return rho, u

    def read_map(self):
        m = np.zeros(self.dim ** 3, dtype=np.int32)
        self._check(self.lib.lbm_read_map(self.h, _ptr(m)))
        return m

    def read_f(self):
        f = np.zeros(19 * self.dim ** 3, dtype=self.dtype)
        self._check(self.lib.lbm_read_f(self.h, _ptr(f)))
        return f

    def time_ms(self):
        t, k = ctypes.c_double(), ctypes.c_double()
        self._check(self.lib.lbm_time_ms(self.h, ctypes.byref(t), ctypes.byref(k)))
        return t.value, k.value

    def launch_times_ms(self):
        """Per-launch (lbm_step) / per-batch (lbm_run) durations in enqueue order."""
        n = ctypes.c_int64()
        self._check(self.lib.lbm_launch_times_ms(self.h, None, 0, ctypes.byref(n)))
        out = np.zeros(n.value, dtype=np.float64)
        self._check(self.lib.lbm_launch_times_ms(self.h, _ptr(out), n.value, ctypes.byref(n)))
        return out

    @property
    def device_name(self) -> str:
        buf = ctypes.create_string_buffer(256)
        self._check(self.lib.lbm_device_name(self.h, buf, 256))
        return buf.value.decode()

    @property
    def effective_params(self):
        out = (ctypes.c_double * 3)()
        self._check(self.lib.lbm_effective_params(self.h, out))
        return {"viscosity": out[0], "velocity": out[1], "inv_tau": out[2]}

    @property
    def block_shape(self):
        b = (ctypes.c_int32 * 3)()
        v = ctypes.c_int32()
        self._check(self.lib.lbm_block_shape(self.h, b, ctypes.byref(v)))
        return (b[0], b[1], b[2]), v.value

    @property
    def device_bytes(self) -> int:
        return self.lib.lbm_device_bytes(self.h)

    @property
    def launch_count(self) -> int:
        return self.lib.lbm_launch_count(self.h)

    @property
    def iteration(self) -> int:
        return self.lib.lbm_iteration(self.h)

    # -- split-phase iteration + dense halo transport (one process per device) --
    def step_planes(self, z_begin: int, z_end: int, update_macro: bool = False):
        self._check(self.lib.lbm_step_planes(self.h, z_begin, z_end, 1 if update_macro else 0))

    def advance(self):
        self._check(self.lib.lbm_advance(self.h))

    @property
    def z_range(self):
        a, b = ctypes.c_int32(), ctypes.c_int32()
        self._check(self.lib.lbm_z_range(self.h, ctypes.byref(a), ctypes.byref(b)))
        return a.value, b.value

    @property
    def halo_elems(self) -> int:
        return self.lib.lbm_halo_elems(self.h)

    def halo_send_ptr(self, face: int):
        return self.lib.lbm_halo_send_buffer(self.h, face)

    def halo_recv_ptr(self, face: int):
        return self.lib.lbm_halo_recv_buffer(self.h, face)

    def halo_pack(self):
        self._check(self.lib.lbm_halo_pack(self.h))

    def halo_unpack(self):
        self._check(self.lib.lbm_halo_unpack(self.h))

    # -- library-driven NCCL exchange (one process per device) --
    @staticmethod
    def comm_unique_id() -> bytes:
        lib = load()
        buf = (ctypes.c_uint8 * 128)()
        rc = lib.lbm_comm_unique_id(buf)
        if rc != 0:
            raise LbmError(rc, lib.lbm_last_error(None).decode())
        return bytes(buf)

    def comm_init(self, unique_id: bytes, rank: int, world: int):
        buf = (ctypes.c_uint8 * 128).from_buffer_copy(unique_id)
        self._check(self.lib.lbm_comm_init(self.h, buf, rank, world))

    IPC_BYTES = 256

    def ipc_export(self) -> bytes:
        buf = (ctypes.c_uint8 * self.IPC_BYTES)()
        self._check(self.lib.lbm_ipc_export(self.h, buf))
        return bytes(buf)

    def ipc_attach(self, face: int, blob: bytes):
        buf = (ctypes.c_uint8 * self.IPC_BYTES).from_buffer_copy(blob)
        self._check(self.lib.lbm_ipc_attach(self.h, face, buf))

    def peer_attach(self, face: int, neighbour: "Simulation"):
        """Same-process neighbour (any device with peer access): its halo plane becomes a store target."""
        self._check(self.lib.lbm_peer_attach(self.h, face, neighbour.h))

    def ipc_detach(self):
        self._check(self.lib.lbm_ipc_detach(self.h))

    def comm_fused(self, mode=FUSED_FLAGS):
        """FUSED_OFF / FUSED_FLAGS (in-kernel epoch flags, one launch per iteration) / FUSED_TOKEN (NCCL token)."""
        if mode is True:
            mode = FUSED_FLAGS
        elif mode is False:
            mode = FUSED_OFF
        self._check(self.lib.lbm_comm_fused(self.h, int(mode)))

    # -- asynchronous read-back into page-locked host memory --
    def read_macros_async(self, rho, u):
        self._check(self.lib.lbm_read_macros_async(self.h, _ptr(rho), _ptr(u)))

    def read_wait(self):
        self._check(self.lib.lbm_read_wait(self.h))

    def run_snapshots(self, iterations: int, every: int):
        """The schedule of lbmcl.hpp:490-521: snapshot of rho/u after init and after every flagged
        iteration.  Returns rho[k, N], u[k, 3, N]."""
        n = self.dim ** 3
        k = 0 if every == 0 else 1 + iterations // every
        rho = np.full((max(k, 1), n), np.nan, dtype=self.dtype)
        u = np.full((max(k, 1), 3, n), np.nan, dtype=self.dtype)
        self.init()
        s = 0
        if every != 0:
            self.read_macros(rho[s], u[s])
            s += 1
        done = 0
        while done < iterations:
            chunk = (every - done % every) if every != 0 else iterations - done
            chunk = min(chunk, iterations - done)
            self.run(chunk, every)
            done += chunk
            if every != 0 and done % every == 0:
                self.read_macros(rho[s], u[s])
                s += 1
        self.sync()
        return rho[:k], u[:k]


class Group:
    """Same-process z-slab group over several devices (or several slabs on one device)."""

    def __init__(self, devices, **kw):
        self.lib = load()
        self.params = make_params(**kw)
        self.dim = self.params.dim
        self.dtype = np.float32 if self.params.precision == F32 else np.float64
        dev = (ctypes.c_int32 * len(devices))(*devices)
        h = ctypes.c_void_p()
        rc = self.lib.lbm_group_create(ctypes.byref(self.params), dev, len(devices), ctypes.byref(h))
        if rc != 0:
            raise LbmError(rc, self.lib.lbm_group_last_error(None).decode())
        self.h = h

    def _check(self, rc):
        if rc != 0:
            raise LbmError(rc, self.lib.lbm_group_last_error(self.h).decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.lbm_group_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def init(self):
        self._check(self.lib.lbm_group_init(self.h))

    def run(self, n_iterations: int, every: int = 0):
        self._check(self.lib.lbm_group_run(self.h, n_iterations, every))

    def sync(self):
        self._check(self.lib.lbm_group_sync(self.h))

    def read_macros(self, rho=None, u=None):
        n = self.dim ** 3
        if rho is None and u is None:
            rho = np.full(n, np.nan, dtype=self.dtype)
            u = np.full((3, n), np.nan, dtype=self.dtype)
        self._check(self.lib.lbm_group_read_macros(self.h, _ptr(rho), _ptr(u)))
        return rho, u

    def read_f(self):
        f = np.zeros(19 * self.dim ** 3, dtype=self.dtype)
        self._check(self.lib.lbm_group_read_f(self.h, _ptr(f)))
        return f

    def time_ms(self):
        t, k = ctypes.c_double(), ctypes.c_double()
        self._check(self.lib.lbm_group_time_ms(self.h, ctypes.byref(t), ctypes.byref(k)))
        return t.value, k.value

    def run_snapshots(self, iterations: int, every: int):
        n = self.dim ** 3
        k = 0 if every == 0 else 1 + iterations // every
        rho = np.full((max(k, 1), n), np.nan, dtype=self.dtype)
        u = np.full((max(k, 1), 3, n), np.nan, dtype=self.dtype)
        self.init()
        s = 0
        if every != 0:
            self.read_macros(rho[s], u[s])
            s += 1
        done = 0
        while done < iterations:
            chunk = (every - done % every) if every != 0 else iterations - done
            chunk = min(chunk, iterations - done)
            self.run(chunk, every)
            done += chunk
            if every != 0 and done % every == 0:
                self.read_macros(rho[s], u[s])
                s += 1
        self.sync()
        return rho[:k], u[:k]


def pinned_array(shape, dtype):
    """numpy array over page-locked host memory from lbm_host_alloc (freed when the array is collected)."""
    lib = load()
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = ctypes.c_void_p()
    rc = lib.lbm_host_alloc(n, ctypes.byref(p))
    if rc != 0:
        raise LbmError(rc, lib.lbm_last_error(None).decode())
    buf = (ctypes.c_uint8 * max(n, 1)).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    import weakref
    weakref.finalize(buf, lib.lbm_host_free, p)
    return arr


def spec_cubin(**kw) -> bytes:
    """The run-time specialised (NVRTC) step kernel for a configuration, compiled without a device."""
    lib = load()
    p = make_params(**kw)
    n = ctypes.c_size_t()
    rc = lib.lbm_spec_cubin(ctypes.byref(p), None, 0, ctypes.byref(n))
    if rc != 0:
        raise LbmError(rc, lib.lbm_last_error(None).decode())
    buf = (ctypes.c_uint8 * n.value)()
    rc = lib.lbm_spec_cubin(ctypes.byref(p), buf, n.value, ctypes.byref(n))
    if rc != 0:
        raise LbmError(rc, lib.lbm_last_error(None).decode())
    return bytes(buf)
