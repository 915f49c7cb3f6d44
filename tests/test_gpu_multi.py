"""Sharded (multi-GPU) FastSLAM: every rank's shard must equal the corresponding slice of the single-process oracle,
bit for bit (global exact sums, global ancestry, cross-rank map exchange).  Needs >= 2 GPUs."""
import ctypes as C
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def n_gpus():
    import rust_robotics_b200 as rr
    c = C.c_int()
    rr.load_library().pfgpu_device_count(C.byref(c))
    return c.value


# mode 2: peer-memory step (fs_mg.cuh, the default when the shard size allows it); mode 1: NCCL collectives (fs_sharded.cuh)
@pytest.mark.parametrize("world,n,side,steps,mode,guests", [(2, 4096, 6, 16, 2, 0), (2, 1 << 16, 8, 6, 2, 0), (2, 4096, 6, 16, 1, 0),
                                                            (2, 4000, 6, 12, 1, 0), (2, 4096, 6, 40, 2, 40), (2, 4096, 6, 40, 1, 192),
                                                            (4, 8192, 6, 14, 2, 0), (8, 16384, 6, 14, 2, 0)])
def test_sharded_fastslam_matches_oracle(world, n, side, steps, mode, guests):
    if n_gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    env = dict(os.environ)
    env["PFGPU_SHARD_P2P"] = "1" if mode == 2 else "0"
    if guests:
        env["PFGPU_GUEST_COLS"] = str(guests)       # few guest columns: compaction (NCCL form) / eager rebuild (peer-memory form) has to run
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(ROOT, "tests", "mgpu_worker.py"), str(n), str(side), str(steps)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and "MGPU_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
    assert f"mode={mode}" in r.stdout, r.stdout[-2000:]
    if guests:
        assert "compactions=0 " not in r.stdout, r.stdout[-2000:]


@pytest.mark.parametrize("mode", [2, 1])
def test_sharded_fastslam_edge_cases(mode):
    """duplicate landmark ids, empty list, all-zero weights (every slot imports the last particle of the last rank), fresh landmarks"""
    if n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ)
    env["PFGPU_SHARD_P2P"] = "1" if mode == 2 else "0"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29543", os.path.join(ROOT, "tests", "mgpu_worker.py"), "edge"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and "MGPU_OK edge" in r.stdout and f"mode={mode}" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


@pytest.mark.parametrize("kind,n,steps", [("pf", 1 << 14, 40), ("mcl", 1 << 15, 8)])
def test_sharded_pf_matches_oracle(kind, n, steps):
    if n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29542", os.path.join(ROOT, "tests", "mgpu_worker.py"), kind, str(n), str(steps)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "MGPU_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
