"""Sharded (multi-GPU) FastSLAM / PF / MCL, one process per GPU (launched like the driver launches bench.py): every rank's
shard must equal the corresponding slice of the single-process oracle, bit for bit.  Needs >= 2 GPUs; the same engine with
all ranks in one process on ONE GPU is covered by tests/test_gpu_parity.py::test_fastslam_sharded_in_process_*."""
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


@pytest.mark.parametrize("world,n,side,steps", [(2, 4096, 6, 16), (2, 1 << 16, 8, 6), (2, 8192, 6, 40), (4, 8192, 6, 14), (8, 16384, 6, 14)])
def test_sharded_fastslam_matches_oracle(world, n, side, steps):
    if n_gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(ROOT, "tests", "mgpu_worker.py"), str(n), str(side), str(steps)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "MGPU_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
    assert "mode=2" in r.stdout, r.stdout[-2000:]


@pytest.mark.parametrize("world,n,side,steps", [(2, 8192, 6, 30), (8, 16384, 6, 14)])
def test_sharded_fastslam2_matches_oracle(world, n, side, steps):
    """FastSLAM 2.0 step (fastslam2.rs): the proposal kernel reads the first observation's landmark through remote rows"""
    if n_gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29544", os.path.join(ROOT, "tests", "mgpu_worker.py"), str(n), str(side), str(steps), "fs2"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "MGPU_OK" in r.stdout and "variant=2" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


def test_sharded_fastslam_edge_cases():
    """duplicate landmark ids, empty list, all-zero weights (every slot descends from the last particle of the last rank), fresh landmarks"""
    if n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29543", os.path.join(ROOT, "tests", "mgpu_worker.py"), "edge"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "MGPU_OK edge" in r.stdout and "mode=2" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


@pytest.mark.parametrize("kind,n,steps", [("pf", 1 << 14, 40), ("mcl", 1 << 15, 8)])
def test_sharded_pf_matches_oracle(kind, n, steps):
    if n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29542", os.path.join(ROOT, "tests", "mgpu_worker.py"), kind, str(n), str(steps)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "MGPU_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
