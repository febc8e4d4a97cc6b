"""GPU parity: the CUDA path, called through the C ABI, against the CPU oracle on the same inputs.

The lid-driven cavity has no random inputs: a case is (dim, stride, nu, U, iterations,
every, precision).  The bar (BASELINE.json north_star): rho and u within 1e-5 (fp32) / 1e-12 (fp64)
of the field scale (rho: 1, u: lid speed), identical NaN masks.  The strict kernels reproduce the
reference's IEEE operation order, so the tests first ask for bit equality and only then fall back to
the tolerance (which is what the -o / fast variant is held to).
"""
import numpy as np
import pytest

from oracle import Oracle

pytestmark = pytest.mark.gpu

TOL = {"f32": 1e-5, "f64": 1e-12}


def _sim(**kw):
    from lbmcl_b200.capi import Simulation
    return Simulation(**kw)


def _compare(got_rho, got_u, exp_rho, exp_u, u_lid, tol, exact):
    assert got_rho.shape == exp_rho.shape and got_u.shape == exp_u.shape
    assert np.array_equal(np.isnan(got_rho), np.isnan(exp_rho)), "NaN mask of rho differs"
    assert np.array_equal(np.isnan(got_u), np.isnan(exp_u)), "NaN mask of u differs"
    if exact and got_rho.tobytes() == exp_rho.tobytes() and got_u.tobytes() == exp_u.tobytes():
        return 0.0, 0.0
    d_rho = np.nanmax(np.abs(got_rho.astype(np.float64) - exp_rho.astype(np.float64)))
    d_u = np.nanmax(np.abs(got_u.astype(np.float64) - exp_u.astype(np.float64))) / abs(u_lid)
    assert d_rho <= tol, f"max|d rho| = {d_rho:.3e} > {tol}"
    assert d_u <= tol, f"max|d u|/U = {d_u:.3e} > {tol}"
    return d_rho, d_u


CASES = [
    # dim, stride, nu,     U,    its, every      -- the reference's own test configurations first
    (8, 8, 0.0089, 0.05, 10, 1),       # make test8 (reference Makefile:73-78)
    (32, 32, 0.0089, 0.05, 10, 1),     # BASELINE config 2, 10 iterations
    (32, 32, 0.0089, 0.05, 60, 20),    # make test32 schedule (Makefile:81-86), shortened
    (8, 32, 0.0089, 0.05, 10, 1),      # default stride (lbm_options.hpp:43) on the default cube
    (16, 16, 0.02, 0.1, 12, 3),        # other physical parameters
    (16, 4096, 0.0089, 0.05, 7, 7),    # stride == dim^3: pure SoA
    (16, 1, 0.0089, 0.05, 5, 1),       # stride 1: AoS
    (16, 2, 0.0123456789, -0.03, 6, 2),  # 6-digit text round trip of nu; lid moving in -x
    (64, 64, 0.0089, 0.05, 9, 3),
    (4, 4, 0.0089, 0.05, 3, 1),        # smallest legal cube: no fluid cell at all
]


@pytest.mark.parametrize("precision", ["f32", "f64"])
@pytest.mark.parametrize("variant", [0, 1, 2, 4, 8, 16])
@pytest.mark.parametrize("case", CASES, ids=lambda c: "d%d_s%d_i%d_e%d" % (c[0], c[1], c[4], c[5]))
def test_strict_matches_oracle(case, variant, precision):
    dim, stride, nu, u_lid, its, every = case
    exp = Oracle(precision).run(dim, stride, nu, u_lid, its, every)
    with _sim(dim=dim, precision=precision, viscosity=nu, velocity=u_lid, stride=stride, variant=variant) as s:
        rho, u = s.run_snapshots(its, every)
        eff = s.effective_params
    o = Oracle(precision).params(nu, u_lid)
    assert eff["inv_tau"] == o["inv_tau"] and eff["velocity"] == o["velocity"]
    _compare(rho, u, exp["rho"], exp["u"], u_lid, TOL[precision], exact=True)
    # strict mode is expected to be bit-identical, not merely within tolerance
    assert rho.tobytes() == exp["rho"].tobytes()
    assert u.tobytes() == exp["u"].tobytes()


@pytest.mark.parametrize("precision", ["f32", "f64"])
@pytest.mark.parametrize("variant", [0, 4, 8, 16])
@pytest.mark.parametrize("case", CASES[:5], ids=lambda c: "d%d_s%d_i%d_e%d" % (c[0], c[1], c[4], c[5]))
def test_fast_within_tolerance(case, variant, precision):
    """-o (contracted arithmetic, approximate fp32 division) against the strict oracle."""
    dim, stride, nu, u_lid, its, every = case
    exp = Oracle(precision).run(dim, stride, nu, u_lid, its, every)
    with _sim(dim=dim, precision=precision, viscosity=nu, velocity=u_lid, stride=stride, fast_math=True,
              variant=variant) as s:
        rho, u = s.run_snapshots(its, every)
    _compare(rho, u, exp["rho"], exp["u"], u_lid, TOL[precision], exact=False)


@pytest.mark.parametrize("block", [(8, 8, 8), (32, 32, 1), (128, 1, 1), (4, 2, 16), (1, 1, 1), (64, 4, 1)])
def test_block_shapes(block):
    """-w is a hint: any requested work-group shape gives the same result."""
    dim, stride, nu, u_lid, its, every = 32, 32, 0.0089, 0.05, 6, 2
    exp = Oracle("f32").run(dim, stride, nu, u_lid, its, every)
    for variant in (1, 2, 4):
        with _sim(dim=dim, stride=stride, block=block, exact_block=True, variant=variant) as s:
            rho, u = s.run_snapshots(its, every)
        assert rho.tobytes() == exp["rho"].tobytes() and u.tobytes() == exp["u"].tobytes()
    with _sim(dim=dim, stride=stride, block=block) as s:   # default: the hint does not change the shape
        assert s.block_shape == ((32, 8, 1), 1)


@pytest.mark.parametrize("precision", ["f32", "f64"])
@pytest.mark.parametrize("variant", [1, 2, 4, 8])
@pytest.mark.parametrize("stride", [64, 1024, 16384])
def test_row_base_addressing_for_strides_between_dim_and_cells(stride, variant, precision):
    """DIM < stride < stored cells (LM_BLOCKROWS: 9 row bases per thread instead of a CSoA index per access):
    e.g. every `-s 64/128` of the reference's sweep at DIM <= 32..64, or -s DIM^2.  Bit-identical to the oracle,
    rho/u snapshots and the -f view, odd and even iteration counts (AA: LOCAL and SHIFT steps)."""
    dim, every = 32, 3
    for its in (6, 7):
        exp = Oracle(precision).run(dim, stride, 0.0089, 0.05, its, every, keep_state=True)
        with _sim(dim=dim, precision=precision, stride=stride, variant=variant) as s:
            rho, u = s.run_snapshots(its, every)
            got_f = s.read_f()
        assert rho.tobytes() == exp["rho"].tobytes() and u.tobytes() == exp["u"].tobytes(), its
        st = exp["state"]
        assert got_f.tobytes() == (st["f_stream"] if (its + 1) % 2 == 0 else st["f_collide"]).tobytes(), its


@pytest.mark.parametrize("precision", ["f32", "f64"])
@pytest.mark.parametrize("variant", [1, 2, 4, 8])
def test_generic_addressing_equals_fast_addressing(variant, precision):
    """LM_GENERIC (any stride) and LM_ROWS / LM_SOA / LM_BLOCKROWS (uniform offsets) are the same function."""
    dim, its, every = 32, 6, 3
    for stride in (32, 8, 32768, 1024):
        exp = Oracle(precision).run(dim, stride, 0.0089, 0.05, its, every)
        with _sim(dim=dim, precision=precision, stride=stride, variant=variant, generic_addressing=True) as s:
            rho, u = s.run_snapshots(its, every)
        assert rho.tobytes() == exp["rho"].tobytes() and u.tobytes() == exp["u"].tobytes()


@pytest.mark.parametrize("precision", ["f32", "f64"])
@pytest.mark.parametrize("dim,stride,block", [
    (64, 32, None),            # two warps per row
    (64, 32, (32, 8, 1)),      # one warp per block, two blocks per row
    (128, 32, (64, 4, 1)),     # warp edges and block edges inside a row
    (64, 64, None),            # stride == DIM
    (64, 512, None),           # LM_BLOCKROWS (a CSoA block holds 8 rows)
    (64, 262144, None),        # LM_SOA
    (32, 8, None),             # stride < warp: a warp's 32 cells span four CSoA runs
])
def test_in_place_variant_rows_wider_than_a_warp_or_a_block(dim, stride, block, precision):
    """The in-place variant on rows that span several warps / several blocks, for every addressing mode: rho/u
    snapshots and the complete -f view (which also shows the pushes into WALL cells) against the oracle, after
    an even and an odd number of iterations (LOCAL and SHIFT steps)."""
    every = 2
    kw = dict(dim=dim, precision=precision, stride=stride, variant=8)
    if block is not None:
        kw.update(block=block, exact_block=True)
    for its in (4, 5):
        exp = Oracle(precision).run(dim, stride, 0.0089, 0.05, its, every, keep_state=True)
        st = exp["state"]
        want_f = (st["f_stream"] if (its + 1) % 2 == 0 else st["f_collide"]).tobytes()
        with _sim(**kw) as s:
            rho, u = s.run_snapshots(its, every)
            got_f = s.read_f()
        assert rho.tobytes() == exp["rho"].tobytes() and u.tobytes() == exp["u"].tobytes(), its
        assert got_f.tobytes() == want_f, its


def test_step_api_equals_run_api():
    """lbm_step (one reference `compute` launch) and lbm_run (the loop) are the same path."""
    dim, stride, its, every = 16, 16, 9, 3
    with _sim(dim=dim, stride=stride) as a, _sim(dim=dim, stride=stride) as b:
        a.init()
        a.run(its, every)
        ra, ua = a.read_macros()
        b.init()
        for it in range(1, its + 1):
            b.step(every != 0 and it % every == 0)
        rb, ub = b.read_macros()
        assert a.iteration == b.iteration == its
        assert a.launch_count == b.launch_count == its
    assert ra.tobytes() == rb.tobytes() and ua.tobytes() == ub.tobytes()


def test_map_matches_reference_classification():
    for dim in (4, 8, 16, 32):
        with _sim(dim=dim, stride=4) as s:
            m = s.read_map()
        assert np.array_equal(m, Oracle("f32").cell_map(dim))


@pytest.mark.parametrize("precision", ["f32", "f64"])
@pytest.mark.parametrize("variant", [1, 8])
@pytest.mark.parametrize("its", [0, 1, 2, 5, 6])
def test_population_dump_view(its, variant, precision):
    """lbm_read_f presents the lattice as the reference stores it (pre-collision, CSoA order, pushed
    values in WALL cells, never-written slots at their initial value): the -f dump (lbmcl.hpp:206-258)."""
    dim, stride, nu, u_lid = 16, 16, 0.0089, 0.05
    o = Oracle(precision)
    st = o.alloc(dim)
    o.init(st, dim, stride, nu, u_lid)
    for it in range(1, its + 1):
        o.step(st, dim, stride, nu, u_lid, it, 0)
    # the buffer iteration its+1 would read (lbmcl.hpp:517-519)
    exp = st["f_stream"] if (its + 1) % 2 == 0 else st["f_collide"]
    with _sim(dim=dim, precision=precision, stride=stride, variant=variant) as s:
        s.init()
        s.run(its, 0)
        got = s.read_f()
    assert np.array_equal(np.isnan(got), np.isnan(exp))
    assert got.tobytes() == exp.tobytes()


def test_golden_8_through_cabi(golden8):
    """make test8: 8^3, 10 iterations, every 1, against the Sailfish fixtures (verify.py metrics).
    Expected magnitudes from SURVEY Appendix A: MAE_rho <~ 6e-7, MAE_u <~ 2e-7."""
    from conftest import wet
    with _sim(dim=8, stride=8) as s:
        rho, u = s.run_snapshots(10, 1)
    for k, it in enumerate(golden8["its"]):
        t_rho, t_v = golden8["rho"][k], golden8["v"][k]
        p_rho = wet(rho[it], 8)
        p_v = np.moveaxis(wet(u[it], 8), 0, -1)
        assert np.array_equal(np.isnan(p_rho), np.isnan(t_rho))
        assert np.nanmax(np.abs(p_rho - t_rho)) <= 1e-6
        assert np.nanmax(np.abs(p_v - t_v)) <= 5e-7


def test_golden_32_through_cabi(golden32):
    """make test32: 32^3, 500 iterations, every 20, stride 32, against the Sailfish fixtures.
    The reference kernel itself sits at MAE_rho ~2e-6, MAE_u ~5e-7 (1.0e-5 * U_lid) at step 500."""
    from conftest import wet
    with _sim(dim=32, stride=32, block=(32, 32, 1)) as s:
        rho, u = s.run_snapshots(500, 20)
    for k, it in enumerate(golden32["its"]):
        t_rho, t_v = golden32["rho"][k], golden32["v"][k]
        p_rho = wet(rho[it // 20], 32)
        p_v = np.moveaxis(wet(u[it // 20], 32), 0, -1)
        assert np.array_equal(np.isnan(p_rho), np.isnan(t_rho))
        assert np.nanmax(np.abs(p_rho - t_rho)) <= 4e-6, it
        assert np.nanmax(np.abs(p_v - t_v)) <= 1.2e-5 * 0.05, it


@pytest.mark.parametrize("precision,dim", [("f32", 256), ("f64", 128)])
def test_large_lattice_properties(precision, dim):
    """Full-size, size-independent checks where the oracle is too slow to be the comparator:
    (1) mass is conserved over the collision cells up to rounding drift,
    (2) the lid cells report u = (U, 0, 0) exactly and everything non-fluid is NaN,
    (3) the run is deterministic (two contexts give identical bits),
    (4) a centre sub-block agrees with the oracle run on the SAME lattice for a few iterations."""
    its = 6
    u_lid = 0.05
    with _sim(dim=dim, precision=precision, stride=32) as a, _sim(dim=dim, precision=precision, stride=32,
                                                                   variant=1) as b:
        a.init()
        a.run(its, its)
        ra, ua = a.read_macros()
        b.init()
        b.run(its, its)
        rb, ub = b.read_macros()
    assert ra.tobytes() == rb.tobytes() and ua.tobytes() == ub.tobytes()
    m = Oracle(precision).cell_map(dim)
    fluid = (m == 1) | ((m & 2) != 0)
    assert np.array_equal(~np.isnan(ra), fluid)
    lid = (m & 2) != 0
    assert np.all(ua[0][lid] == np.dtype(ra.dtype).type(u_lid)) and np.all(ua[1][lid] == 0) and np.all(ua[2][lid] == 0)
    exp = Oracle(precision).run(dim, 32, 0.0089, u_lid, its, its)
    assert ra.tobytes() == exp["rho"][1].tobytes()
    assert ua.tobytes() == exp["u"][1].tobytes()


def test_slab_group_on_one_device_matches_single():
    """z-slab decomposition logic (halo planes, crossing populations, event ordering) exercised with
    all slabs on device 0: 2, 4 and 8 slabs must reproduce the single-context bits."""
    from lbmcl_b200.capi import Group
    dim, stride, its, every = 32, 32, 12, 4
    exp = Oracle("f32").run(dim, stride, 0.0089, 0.05, its, every)
    for n in (2, 4, 8, 16):
        with Group([0] * n, dim=dim, stride=stride) as g:
            rho, u = g.run_snapshots(its, every)
        assert rho.tobytes() == exp["rho"].tobytes(), n
        assert u.tobytes() == exp["u"].tobytes(), n


@pytest.mark.parametrize("precision,stride", [("f32", 32), ("f64", 32), ("f32", 2048), ("f64", 1)])
def test_flag_transport_between_two_contexts_on_one_device(precision, stride, monkeypatch):
    """The one-launch-per-iteration transport (peer stores + in-kernel epoch flags, include/lbm_b200.h 2c) with
    both slabs in this process on device 0 (lbm_peer_attach): each context runs on its own stream, the kernels
    meet only through the flag words.  Re-initialisation is a phase of the protocol and needs no barrier."""
    from lbmcl_b200.capi import FUSED_FLAGS, Simulation
    monkeypatch.setenv("LBM_SYNC_TIMEOUT_S", "5")   # a broken protocol fails in seconds, not minutes
    dim, its, every = 32, 11, 4
    exp = Oracle(precision).run(dim, stride, 0.0089, 0.05, its, every)
    parts = [(0, 8), (8, 24), (24, 32)]
    sims = [Simulation(dim=dim, precision=precision, stride=stride, z_range=z) for z in parts]
    try:
        for i, s in enumerate(sims):
            if i > 0:
                s.peer_attach(0, sims[i - 1])
            if i + 1 < len(sims):
                s.peer_attach(1, sims[i + 1])
        for s in sims:
            s.comm_fused(FUSED_FLAGS)
        n = dim ** 3
        for rep in range(2):      # the second pass re-initialises without any host synchronisation in between
            for s in sims:
                s.init()
            k = 0
            done = 0
            rho = np.full((1 + its // every, n), np.nan, dtype=sims[0].dtype)
            u = np.full((1 + its // every, 3, n), np.nan, dtype=sims[0].dtype)
            for s in sims:
                s.read_macros(rho[k], u[k])
            k += 1
            while done < its:
                chunk = min(every - done % every, its - done)
                # iteration by iteration over the slabs: a kernel then only waits for kernels enqueued BEFORE it,
                # so the run cannot stall even if the device does not overlap the three streams (contexts of one
                # process on one device may share a hardware queue; separate processes / devices never do)
                for i in range(chunk):
                    flag = (done + i + 1) % every == 0
                    for s in (sims if rep == 0 else sims[::-1]):
                        s.step(flag)
                done += chunk
                if done % every == 0:
                    for s in sims:
                        s.read_macros(rho[k], u[k])
                    k += 1
            for s in sims:
                s.sync()
                assert s.launch_count == its          # ONE launch per iteration and slab
            assert rho.tobytes() == exp["rho"].tobytes(), rep
            assert u.tobytes() == exp["u"].tobytes(), rep
    finally:
        for s in sims:
            s.sync()
        for s in sims:
            s.close()


def test_neighbour_attachment_is_validated():
    """lbm_peer_attach / lbm_ipc_attach refuse a neighbour that is not adjacent or runs another configuration
    (a wrong blob would otherwise let a boundary kernel overwrite planes the other slab OWNS)."""
    from lbmcl_b200.capi import FUSED_FLAGS, LbmError, Simulation
    a = Simulation(dim=32, stride=32, z_range=(0, 8))
    b = Simulation(dim=32, stride=32, z_range=(8, 16))
    c = Simulation(dim=32, stride=32, z_range=(16, 32))
    d = Simulation(dim=32, stride=8, z_range=(8, 16))
    try:
        with pytest.raises(LbmError, match="not adjacent"):
            a.peer_attach(1, c)
        with pytest.raises(LbmError, match="not adjacent"):
            b.peer_attach(0, c)          # wrong face
        with pytest.raises(LbmError, match="different configuration"):
            a.peer_attach(1, d)
        with pytest.raises(LbmError, match="cube boundary"):
            a.peer_attach(0, b)
        with pytest.raises(LbmError, match="no attached neighbour"):
            b.comm_fused(FUSED_FLAGS)
        a.peer_attach(1, b)
        with pytest.raises(LbmError, match="already has a neighbour"):
            a.peer_attach(1, b)
        # the same checks through the IPC blob (exported and attached inside one process is refused by CUDA
        # only at the open; the geometry checks come first)
        blob = c.ipc_export()
        with pytest.raises(LbmError, match="not adjacent"):
            b.ipc_attach(0, blob)
        # a slab with neighbours but no transport is not drivable on its own
        a.init()
        with pytest.raises(LbmError, match="peer neighbours"):
            a.run(1, 0)
        with pytest.raises(LbmError, match="peer neighbours"):
            a.step(False)
        t = a.launch_times_ms()          # the refused calls left no half-recorded event pair behind
        assert len(t) == 0
        a.ipc_detach()
        a.run(2, 0)
        assert len(a.launch_times_ms()) == 1
    finally:
        for s in (a, b, c, d):
            s.close()


def test_dense_halo_transport_matches_single():
    """The one-process-per-device transport (pack -> copy -> unpack), driven from the host the way
    bench.py drives it under torchrun, here with both slabs on device 0 and cudaMemcpy as the wire."""
    import ctypes
    from lbmcl_b200.capi import Simulation
    dim, stride, its = 16, 16, 8
    exp = Oracle("f64").run(dim, stride, 0.0089, 0.05, its, its)
    cudart = ctypes.CDLL("libcudart.so")
    cudart.cudaMemcpy.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
    lo = Simulation(dim=dim, precision="f64", stride=stride, z_range=(0, 8))
    hi = Simulation(dim=dim, precision="f64", stride=stride, z_range=(8, 16))
    nbytes = lo.halo_elems * 8
    lo.init()
    hi.init()
    for it in range(1, its + 1):
        lo.run(1, its)
        hi.run(1, its)
        lo.halo_pack()
        hi.halo_pack()
        lo.sync()
        hi.sync()
        assert cudart.cudaMemcpy(hi.halo_recv_ptr(0), lo.halo_send_ptr(1), nbytes, 3) == 0
        assert cudart.cudaMemcpy(lo.halo_recv_ptr(1), hi.halo_send_ptr(0), nbytes, 3) == 0
        lo.halo_unpack()
        hi.halo_unpack()
    n = dim ** 3
    rho = np.full(n, np.nan)
    u = np.full((3, n), np.nan)
    lo.read_macros(rho, u)
    hi.read_macros(rho, u)
    # compact slab read-back = the owned planes of the global arrays
    r_hi, u_hi = hi.read_macros_slab()
    assert r_hi.tobytes() == rho[n // 2:].tobytes() and u_hi.tobytes() == np.ascontiguousarray(u[:, n // 2:]).tobytes()
    lo.close()
    hi.close()
    assert rho.tobytes() == exp["rho"][1].tobytes() and u.tobytes() == exp["u"][1].tobytes()


def test_slab_group_across_devices_matches_single():
    """Same-process group with one slab per physical GPU: the crossing populations travel as peer
    (NVLink) stores from inside the boundary-plane kernel.  Needs >= 2 GPUs (gpurun --gpus N)."""
    import torch
    from lbmcl_b200.capi import Group
    n_dev = torch.cuda.device_count()
    if n_dev < 2:
        pytest.skip("one GPU only")
    for precision, dim, stride, its, every in (("f32", 32, 32, 12, 4), ("f64", 64, 64, 8, 4)):
        exp = Oracle(precision).run(dim, stride, 0.0089, 0.05, its, every)
        for n in sorted({2, n_dev}):
            with Group(list(range(n)), dim=dim, precision=precision, stride=stride) as g:
                rho, u = g.run_snapshots(its, every)
            assert rho.tobytes() == exp["rho"].tobytes(), (precision, n)
            assert u.tobytes() == exp["u"].tobytes(), (precision, n)


def test_aa_variant_uses_half_the_lattice_memory_and_matches_at_size():
    """The in-place AA variant: one lattice instead of two, same bits as the two-lattice kernel at a
    size where the oracle is too slow to be run for long (128^3, 40 iterations, odd and even counts)."""
    with _sim(dim=128, stride=32, variant=1) as a, _sim(dim=128, stride=32, variant=8) as b:
        assert b.device_bytes < 0.6 * a.device_bytes
        for its in (40, 41):
            a.init(); a.run(its, its); ra, ua = a.read_macros()
            b.init(); b.run(its, its); rb, ub = b.read_macros()
            assert ra.tobytes() == rb.tobytes() and ua.tobytes() == ub.tobytes()
            assert a.read_f().tobytes() == b.read_f().tobytes()
    from lbmcl_b200.capi import LbmError
    with pytest.raises(LbmError):
        _sim(dim=32, variant=8, z_range=(0, 16))


@pytest.mark.parametrize("variant", [1, 8])
@pytest.mark.parametrize("every", [0, 7, 16, 33])
def test_graph_chunks_from_any_parity(every, variant):
    """Small lattices replay 16-iteration CUDA graphs inside lbm_run; chunks must start correctly from
    either lattice parity and step around the iterations that store rho/u."""
    dim, stride, nu, u_lid = 16, 16, 0.0089, 0.05
    o = Oracle("f32")
    st = o.alloc(dim)
    o.init(st, dim, stride, nu, u_lid)
    with _sim(dim=dim, stride=stride, variant=variant) as s:
        s.init()
        done = 0
        for n in (5, 40, 131, 1, 160, 35, 16):   # 131/160: graphs get captured (>= 128 left) at either parity, then reused
            s.run(n, every)
            for it in range(done + 1, done + n + 1):
                o.step(st, dim, stride, nu, u_lid, it, every)
            done += n
            rho, u = s.read_macros()
            assert rho.tobytes() == st["rho"].tobytes() and u.tobytes() == st["u"].reshape(3, -1).tobytes(), (done, every)
            exp_f = st["f_stream"] if (done + 1) % 2 == 0 else st["f_collide"]
            assert s.read_f().tobytes() == exp_f.tobytes(), done
        assert s.iteration == done and s.launch_count == done


@pytest.mark.parametrize("precision", ["f32", "f64"])
@pytest.mark.parametrize("tx", [None, "32", "64"])
def test_tma_variant_rows_segments_and_slabs(tx, precision, monkeypatch):
    """The TMA-fed kernels (bulk-tensor loads into an mbarrier ring, bulk-tensor stores): whole rows per
    tile, rows cut into several segments (edge threads fetch their neighbour themselves), different
    strides, macro stores, and as the interior kernel of a slab group."""
    from lbmcl_b200.capi import Group
    if tx is not None:
        monkeypatch.setenv("LBM_TMA_TX", tx)
    dim, its, every = 128, 7, 3
    for stride in (32, 128, 8):
        exp = Oracle(precision).run(dim, stride, 0.0089, 0.05, its, every)
        with _sim(dim=dim, precision=precision, stride=stride, variant=16) as s:
            rho, u = s.run_snapshots(its, every)
            got_f = s.read_f()
        assert rho.tobytes() == exp["rho"].tobytes() and u.tobytes() == exp["u"].tobytes(), stride
    o = Oracle(precision)
    st = o.alloc(dim)
    o.init(st, dim, 8, 0.0089, 0.05)
    for it in range(1, its + 1):
        o.step(st, dim, 8, 0.0089, 0.05, it, 0)
    assert got_f.tobytes() == (st["f_stream"] if (its + 1) % 2 == 0 else st["f_collide"]).tobytes()
    exp = Oracle(precision).run(64, 32, 0.0089, 0.05, 9, 3)
    with Group([0, 0, 0, 0], dim=64, precision=precision, stride=32, variant=16) as g:
        rho, u = g.run_snapshots(9, 3)
    assert rho.tobytes() == exp["rho"].tobytes() and u.tobytes() == exp["u"].tobytes()


def test_group_dump_view_matches_the_single_device_view():
    """lbm_group_read_f: the -f view gathered over z-slabs (the reference's storeF has no single-device
    restriction, lbmcl.hpp:206-258)."""
    from lbmcl_b200.capi import Group
    dim, stride = 16, 16
    for precision in ("f32", "f64"):
        o = Oracle(precision)
        st = o.alloc(dim)
        o.init(st, dim, stride, 0.0089, 0.05)
        with Group([0, 0, 0, 0], dim=dim, precision=precision, stride=stride) as g:
            g.init()
            for its in range(0, 5):
                exp = st["f_stream"] if (its + 1) % 2 == 0 else st["f_collide"]
                got = g.read_f()
                assert np.array_equal(np.isnan(got), np.isnan(exp)), its
                assert got.tobytes() == exp.tobytes(), its
                g.run(1, 0)
                o.step(st, dim, stride, 0.0089, 0.05, its + 1, 0)


@pytest.mark.parametrize("precision", ["f32", "f64"])
def test_nvrtc_specialised_variant_is_bit_identical(precision):
    """LBM_VARIANT_NVRTC: the step kernel compiled at run time with DIM / stride / offsets / INV_TAU / U as
    literals (the reference's -D kernel options, lbmcl.hpp:131-156)."""
    for dim, stride, nu, u_lid, its, every in ((32, 32, 0.0089, 0.05, 10, 1), (16, 2, 0.0123456789, -0.03, 6, 2),
                                               (32, 1024, 0.02, 0.1, 7, 7), (16, 4096, 0.0089, 0.05, 7, 7)):
        exp = Oracle(precision).run(dim, stride, nu, u_lid, its, every)
        with _sim(dim=dim, precision=precision, viscosity=nu, velocity=u_lid, stride=stride, variant=32) as s:
            rho, u = s.run_snapshots(its, every)
        assert rho.tobytes() == exp["rho"].tobytes() and u.tobytes() == exp["u"].tobytes(), (dim, stride)
    exp = Oracle(precision).run(32, 32, 0.0089, 0.05, 8, 4)
    with _sim(dim=32, precision=precision, stride=32, variant=32, fast_math=True) as s:
        rho, u = s.run_snapshots(8, 4)
    _compare(rho, u, exp["rho"], exp["u"], 0.05, TOL[precision], exact=False)


def test_tma_variant_inside_cuda_graphs():
    """TMA-fed kernels replayed from the 16-iteration CUDA graphs of lbm_run (small lattices, long runs): the
    shared-memory opt-in is set once per device at lbm_create, not inside the capture."""
    dim, stride, its = 32, 32, 16 * 9 + 5
    exp = Oracle("f32").run(dim, stride, 0.0089, 0.05, its, its)
    with _sim(dim=dim, stride=stride, variant=16) as s:
        s.init()
        s.run(its, its)
        rho, u = s.read_macros()
        assert s.launch_count == its
    assert rho.tobytes() == exp["rho"][1].tobytes() and u.tobytes() == exp["u"][1].tobytes()


def test_async_read_back_overlaps_and_is_ordered():
    """lbm_read_macros_async into page-locked memory: the copy sees exactly the state at the call, later
    iterations run alongside, and the next flagged iteration waits for the copy before it overwrites rho/u."""
    from lbmcl_b200.capi import pinned_array
    dim, stride, every, its = 64, 32, 5, 20
    exp = Oracle("f32").run(dim, stride, 0.0089, 0.05, its, every)
    n = dim ** 3
    bufs = [(pinned_array((n,), np.float32), pinned_array((3, n), np.float32)) for _ in range(2)]
    with _sim(dim=dim, stride=stride) as s:
        s.init()
        k = 0
        s.read_macros_async(*bufs[0])
        got = []
        for chunk in range(its // every):
            s.run(every, every)                       # enqueued while the previous copy may still be running
            s.read_wait()
            got.append((bufs[k % 2][0].copy(), bufs[k % 2][1].copy()))
            k += 1
            s.read_macros_async(*bufs[k % 2])
        s.read_wait()
        got.append((bufs[k % 2][0].copy(), bufs[k % 2][1].copy()))
        total, kernels = s.time_ms()
        assert total >= kernels > 0
    assert len(got) == 1 + its // every
    for i, (r, v) in enumerate(got):
        assert r.tobytes() == exp["rho"][i].tobytes(), i
        assert v.tobytes() == exp["u"][i].tobytes(), i


def test_per_launch_timings_like_the_reference_event_list():
    """lbm_launch_times_ms = the per-event list behind kernelsTimingsMS (lbmcl.hpp:580-593): one entry
    per lbm_step launch / per lbm_run batch, summing to kernels_ms."""
    with _sim(dim=64, stride=32) as s:
        s.init()
        for _ in range(7):
            s.step(False)
        s.run(5, 0)
        t = s.launch_times_ms()
        total, kernels = s.time_ms()
        assert len(t) == 8 and np.all(t > 0)
        assert abs(t.sum() - kernels) < 1e-6 * max(1.0, kernels)
        assert total >= kernels
        s.init()                       # a fresh profile, like a fresh LBMCL object
        assert len(s.launch_times_ms()) == 0
