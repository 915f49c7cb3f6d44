"""world_size-2 tests (CPU) of the host-side multi-process plumbing, launched the way the driver launches bench.py
(torchrun sets RANK / WORLD_SIZE / MASTER_*): shard bounds, unique-id broadcast, barrier, max over ranks through the
torch-free TCP control plane (rust_robotics_b200/dist.py), and bench.py's reference arm (rank 0 prints, the others exit 0)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1])
from rust_robotics_b200 import dist as rdist
grp = rdist.TcpGroup()
rank, world = grp.rank, grp.world
assert world == 2
uid = rdist.broadcast_unique_id(grp, lambda: bytes(range(128)))
assert uid == bytes(range(128)), uid
lo, hi = rdist.shard_bounds(1 << 16, rank, world)
assert hi - lo == (1 << 16) // world and lo == rank * ((1 << 16) // world)
m = grp.max(1.0 + rank)
assert m == float(world), m
for _ in range(3): grp.barrier()
if rank == 0: print("DIST_OK")
grp.close()
'''


def _torchrun(args, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port)] + args
    return subprocess.run(cmd, capture_output=True, text=True, timeout=600)


def test_world2_plumbing(tmp_path):
    w = tmp_path / "worker.py"
    w.write_text(WORKER)
    r = _torchrun([str(w), ROOT], 29533)
    assert r.returncode == 0 and "DIST_OK" in r.stdout, r.stdout + r.stderr[-3000:]


def test_bench_reference_arm_under_torchrun():
    r = _torchrun([os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "4", "--warmup", "3"], 29534)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["n_gpus"] == 2 and j["e2e"]["h2d_bytes_per_step"] == 0 and j["cpu_baseline"]["kind"] == "port"
