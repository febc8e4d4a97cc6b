"""The one-process-per-GPU z-slab transports, launched the way the driver launches bench.py: under
torch.distributed.run, one rank per GPU.  tools/multi_gpu_check.py compares every transport (host-driven dense,
NCCL dense, NCCL token, in-kernel flags) bit for bit with the CPU oracle on rank 0.

On a one-GPU box the NCCL-free transport is still exercised across PROCESSES: two ranks share GPU 0 (gloo
bootstrap, CUDA IPC mappings, in-kernel epoch flags; the ranks' kernels are time-sliced)."""
import os
import socket
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

CHECK = os.path.join(ROOT, "tools", "multi_gpu_check.py")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _torchrun(nproc, script_args, timeout=900, env_extra=None):
    env = dict(os.environ)
    env.update(env_extra or {})
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port())] + script_args
    return subprocess.run(cmd, cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                          timeout=timeout)


def test_every_transport_matches_the_oracle_under_torchrun():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("one GPU only (the driver's scaling run covers N > 1; see test_flags_transport_across_processes_on_one_gpu)")
    for world in sorted({2, min(n, 8)}):
        r = _torchrun(world, [CHECK])
        assert r.returncode == 0, r.stdout[-4000:]
        assert "MISMATCH" not in r.stdout
        assert r.stdout.count("bit-identical") >= 10, r.stdout[-4000:]


def test_flags_transport_across_processes_on_one_gpu():
    """Two ranks on GPU 0: IPC peer stores + in-kernel epoch flags between two PROCESSES, no NCCL."""
    r = _torchrun(2, [CHECK, "--backend", "gloo", "--same-device", "--only", "flags"],
                  env_extra={"LBM_SYNC_TIMEOUT_S": "60"})
    if r.returncode != 0 and "timed out" in r.stdout and "MISMATCH" not in r.stdout:
        pytest.skip("the GPU did not time-slice the two ranks' kernels within the wait limit")
    assert r.returncode == 0, r.stdout[-4000:]
    assert "MISMATCH" not in r.stdout and r.stdout.count("bit-identical") >= 4, r.stdout[-4000:]


def test_bench_under_torchrun_carries_a_green_parity_gate():
    """bench.py at N = 2 on a small cube: the JSON line must carry parity_gate.ok and one launch per step and rank."""
    import json
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("one GPU only")
    r = _torchrun(2, [os.path.join(ROOT, "bench.py"), "--gpus", "2", "--steps", "6", "--warmup", "4", "--dim", "256",
                      "--no-extra"])
    assert r.returncode == 0, r.stdout[-4000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("{") and '"metric"' in l][-1]
    out = json.loads(line)
    assert out["parity_gate"]["ok"] is True and out["parity_gate"]["transport"] == "peer-stores+flags"
    assert all(c["vs_oracle"] and c["vs_single_gpu"] for c in out["parity_gate"]["cases"])
    assert out["n_gpus"] == 2 and out["gpu_launches"] == 2 * 6
    assert "peer-stores+flags" in out["config"]["workload"]
