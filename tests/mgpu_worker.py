"""Worker of tests/test_gpu_multi.py: one process per GPU (torchrun); the sharded engine against the full-size oracle."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import rust_robotics_b200 as rr  # noqa: E402
from rust_robotics_b200 import dist as rdist, scenarios  # noqa: E402
import _oracle  # noqa: E402


def run_pf(grp, kind, rank, world, local, n_global, steps, uid):
    """sharded ParticleFilterLocalizer / MonteCarloLocalizer vs the full-size oracle"""
    mcl = kind == "mcl"
    sc = scenarios.PfScenario("c2" if mcl else "c1", steps=steps)
    L = _oracle.load(libm=False)
    if mcl:
        g = rr.MonteCarloLocalizer(rr.MonteCarloLocalizationConfig(n_global, n_global, 0.05, 2.326, 0.25, 0.05, 0.02, 0.1), seed=5, device=local,
                                   shard=(uid, rank, world))
        o = _oracle.OraclePF(L, n_global, range_noise=0.25, velocity_noise=0.05, yaw_rate_noise=0.02, seed=5, mode=1, max_particles=n_global)
    else:
        g = rr.ParticleFilterLocalizer(rr.ParticleFilterConfig(n_global, 0.6, 0.25), seed=5, device=local, shard=(uid, rank, world))
        o = _oracle.OraclePF(L, n_global, threshold=0.6, range_noise=0.25, seed=5)
    L.orc_pf_set_fast_search(o.h, 1)
    rc = g.L.pfgpu_pf_init_state(g.h, rr.api._dp(np.asarray(sc.init, dtype=np.float64)))
    assert rc == 0
    o.init_state(sc.init)
    lo, hi = rdist.shard_bounds(n_global, rank, world)
    resamples = 0
    for t in range(steps):
        obs = sc.obs[t][:: max(1, sc.obs[t].shape[0] // 24)] if mcl else sc.obs[t]
        ge = g.try_step(sc.controls[t], obs)
        oe, did = o.step(sc.controls[t], obs)
        assert np.allclose(ge, oe, rtol=1e-6, atol=1e-9), f"rank {rank} step {t}: estimate {ge} vs {oe}"
        if did:
            resamples += 1
            assert np.array_equal(g.last_indices(), o.last_indices()[lo:hi]), f"rank {rank} step {t}: indices"
    assert np.array_equal(g.get_particles(), o.particles()[lo:hi]), f"rank {rank}: particles differ"
    assert resamples > 0
    grp.barrier()
    if rank == 0:
        print(f"MGPU_OK kind={kind} world={world} n={n_global} resamples={resamples}")


def run_edge(grp, rank, world, local, uid):
    """The single-GPU edge cases of tests/test_gpu_parity.py::test_fastslam_edge_cases, sharded."""
    n, m = 1024 * world, 4
    lm_xy = np.array([[5.0, 0.0], [0.0, 5.0], [5.0, 5.0], [-5.0, 2.0]])
    g = rr.FastSlam1(n, m, rr.FsConfig(nth=n / 1.5), seed=11, device=local, shard=(uid, rank, world))
    L = _oracle.load(libm=False)
    o = _oracle.OracleFS(L, n, m, seed=11, nth=n / 1.5)
    g.seed_map([0.0, 0.0, 0.0], lm_xy); o.seed_map([0.0, 0.0, 0.0], lm_xy)
    lo, hi = rdist.shard_bounds(n, rank, world)

    def same(tag):
        grp.barrier()                         # nobody steps on while a peer still reads through remote references
        gp, gl = g.state(); op, ol = o.state()
        grp.barrier()
        assert np.array_equal(gp, op[lo:hi]), f"rank {rank} {tag}: pose/weights differ"
        assert np.array_equal(gl, ol[lo:hi]), f"rank {rank} {tag}: landmarks differ"

    z = [(5.1, 0.02, 0), (5.0, 1.55, 1), (4.9, -0.01, 0)]          # duplicate lm_id
    assert g.fastslam_update([1.0, 0.1], z) == bool(o.step([1.0, 0.1], z)); same("duplicate ids")
    assert g.fastslam_update([1.0, 0.1], []) == bool(o.step([1.0, 0.1], [])); same("empty obs")
    p, l = g.state(); p[:, 0] = 0.0; g.set_state(p, l)              # all-zero weights: every slot clones particle n-1
    op, ol = o.state(); op[:, 0] = 0.0; o.set_state(op, ol)
    assert g.fastslam_update([1.0, 0.1], z[:2]) is True and o.step([1.0, 0.1], z[:2]) == 1
    assert np.all(g.last_indices() == n - 1); same("zero weights")
    p, l = g.state(); l[:, 2, :] = [0.0, 0.0, 1000.0, 0.0, 0.0, 1000.0]; g.set_state(p, l)      # fresh + initialised landmarks
    op, ol = o.state(); ol[:, 2, :] = [0.0, 0.0, 1000.0, 0.0, 0.0, 1000.0]; o.set_state(op, ol)
    z2 = [(7.0, 0.8, 2), (5.0, 0.1, 0)]
    for _ in range(6):
        assert g.fastslam_update([1.0, 0.0], z2) == bool(o.step([1.0, 0.0], z2))
    same("mixed fresh")
    grp.barrier()
    if rank == 0:
        print(f"MGPU_OK edge world={world} mode={g.shard_mode()}")


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    grp = rdist.TcpGroup()
    uid = rdist.broadcast_unique_id(grp, rdist.nccl_unique_id)
    if sys.argv[1] == "edge":
        run_edge(grp, rank, world, local, uid)
        grp.close()
        return
    if sys.argv[1] in ("pf", "mcl"):
        run_pf(grp, sys.argv[1], rank, world, local, int(sys.argv[2]), int(sys.argv[3]), uid)
        grp.close()
        return
    n_global, side, steps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    variant = 2 if len(sys.argv) > 4 and sys.argv[4] == "fs2" else 1      # fastslam2.rs step on the same engine
    sc = scenarios.FastSlamScenario(side, (10.0 * side / 2 - 5.0, 10.0 * side / 2 - 5.0, 0.0), (1.0, 0.025), steps)
    g = (rr.FastSlam2 if variant == 2 else rr.FastSlam1)(n_global, sc.m, rr.FsConfig(nth=n_global / 1.5), seed=9, device=local, shard=(uid, rank, world))
    L = _oracle.load(libm=False)
    o = _oracle.OracleFS(L, n_global, sc.m, seed=9, variant=variant, nth=n_global / 1.5)
    g.seed_map(sc.start, sc.landmarks)
    o.seed_map(sc.start, sc.landmarks)
    lo, hi = rdist.shard_bounds(n_global, rank, world)
    resamples = 0
    for t in range(steps):
        did = g.fastslam_update(sc.control, sc.obs[t])
        odid = bool(o.step(sc.control, sc.obs[t]))
        assert did == odid, f"rank {rank} step {t}: gate {did} vs oracle {odid}"
        assert abs(g.last_neff() - o.last_neff()) <= 1e-9 * abs(o.last_neff()), f"rank {rank} step {t}: neff"
        if did:
            resamples += 1
            assert np.array_equal(g.last_indices(), o.last_indices()[lo:hi]), f"rank {rank} step {t}: indices"
        bi, bp = g.get_best_particle()
        assert bi == o.best(), f"rank {rank} step {t}: best {bi} vs {o.best()}"
    grp.barrier()
    gp, gl = g.state()
    op, ol = o.state()
    assert np.array_equal(gp, op[lo:hi]), f"rank {rank}: pose/weights differ"
    assert np.array_equal(gl, ol[lo:hi]), f"rank {rank}: landmarks differ"
    assert resamples > 0
    grp.barrier()
    if rank == 0:
        st = g.stats()
        print(f"MGPU_OK world={world} n={n_global} variant={variant} resamples={resamples} mode={g.shard_mode()} serial_fallbacks={st.serial_fallbacks}")
    grp.close()


if __name__ == "__main__":
    main()
