"""FastSLAM 2.0 oracle (oracle/fs2_oracle.c) against the facts the reference pins (fs2.rs:443-545) and against an independent
numpy evaluation of the same formulas (tolerance level: numpy sums in a different order; bit-level order is the oracle's job)."""
import math

import numpy as np
import pytest

import _oracle


@pytest.fixture(scope="module")
def L():
    return _oracle.load(libm=False)


MOTION_COV = np.diag([0.1, 0.1, 0.01])          # fs2.rs:31
R = np.diag([0.5, 0.0305])                      # fs2.rs:28
DT = 0.1


def wrap(a):
    while a > math.pi:
        a -= 2 * math.pi
    while a < -math.pi:
        a += 2 * math.pi
    return a


def numpy_proposal(pose, u, z, lm):
    x, y, yaw = pose
    xp = np.array([x + u[0] * DT * math.cos(yaw), y + u[0] * DT * math.sin(yaw), wrap(yaw + u[1] * DT)])
    G = np.array([[1, 0, -u[0] * DT * math.sin(yaw)], [0, 1, u[0] * DT * math.cos(yaw)], [0, 0, 1.0]])
    Pp = G @ MOTION_COV @ G.T
    if not lm[2] < 100.0:
        return xp, Pp
    dx, dy = lm[0] - xp[0], lm[1] - xp[1]
    d2 = dx * dx + dy * dy
    d = math.sqrt(d2)
    Hp = np.array([[-dx / d, -dy / d, 0.0], [dy / d2, -dx / d2, -1.0]])
    Hl = np.array([[dx / d, dy / d], [-dy / d2, dx / d2]])
    P = np.array([[lm[2], lm[3]], [lm[4], lm[5]]])
    Q = Hl @ P @ Hl.T + R
    Qi = np.linalg.inv(Q)
    Ppost = np.linalg.inv(np.linalg.inv(Pp) + Hp.T @ Qi @ Hp)
    innov = np.array([z[0] - d, wrap(z[1] - wrap(math.atan2(dy, dx) - xp[2]))])
    return xp + Ppost @ Hp.T @ Qi @ innov, Ppost


def test_proposal_improves_with_observation(L):
    """fs2.rs:478-508: landmark (5,0) cov 0.5 I, u = (1,0), z = (5,0): det(cov) < det(motion-only cov), |mean - x_pred| < 1"""
    o = _oracle.OracleFS(L, 1, 1, variant=2)
    mean, cov = o.compute_proposal([0, 0, 0], [1.0, 0.0], (5.0, 0.0), [5.0, 0.0, 0.5, 0.0, 0.0, 0.5])
    G = np.array([[1, 0, 0.0], [0, 1, 0.1], [0, 0, 1.0]])
    motion_only = G @ MOTION_COV @ G.T
    assert np.linalg.det(cov) < np.linalg.det(motion_only)
    assert np.linalg.norm(mean - np.array([0.1, 0.0, 0.0])) < 1.0
    nm, nc = numpy_proposal([0, 0, 0], [1.0, 0.0], (5.0, 0.0), [5.0, 0.0, 0.5, 0.0, 0.0, 0.5])
    assert np.allclose(mean, nm, rtol=1e-11, atol=1e-13) and np.allclose(cov, nc, rtol=1e-10, atol=1e-14)


def test_proposal_matches_numpy_on_random_cases(L):
    o = _oracle.OracleFS(L, 1, 1, variant=2)
    rng = np.random.default_rng(5)
    for _ in range(300):
        pose = [rng.uniform(-20, 20), rng.uniform(-20, 20), rng.uniform(-3.1, 3.1)]
        u = [rng.uniform(0.1, 2.0), rng.uniform(-0.5, 0.5)]
        lm = [pose[0] + rng.uniform(3, 15) * rng.choice([-1, 1]), pose[1] + rng.uniform(3, 15) * rng.choice([-1, 1])]
        a, b = rng.uniform(0.05, 5.0, 2)
        c = rng.uniform(-0.5, 0.5) * math.sqrt(a * b)
        lm6 = lm + [a, c, c, b]
        z = (math.hypot(lm[0] - pose[0], lm[1] - pose[1]) + rng.normal(0, 0.5), rng.uniform(-3, 3))
        mean, cov = o.compute_proposal(pose, u, z, lm6)
        nm, nc = numpy_proposal(pose, u, z, lm6)
        assert np.allclose(mean, nm, rtol=1e-8, atol=1e-10)
        assert np.allclose(cov, nc, rtol=1e-8, atol=1e-12)


def test_proposal_uninitialised_landmark_is_motion_only(L):
    """fs2.rs:188-191"""
    o = _oracle.OracleFS(L, 1, 1, variant=2)
    mean, cov = o.compute_proposal([1, 2, 0.3], [1.0, 0.1], (5.0, 0.2), [0, 0, 1000.0, 0, 0, 1000.0])
    nm, nc = numpy_proposal([1, 2, 0.3], [1.0, 0.1], (5.0, 0.2), [0, 0, 1000.0, 0, 0, 1000.0])
    assert np.allclose(mean, nm, rtol=0, atol=1e-15) and np.allclose(cov, nc, rtol=1e-14, atol=1e-18)
    # the threshold is `cov00 < 100` (fs2.rs:50): exactly 100 counts as uninitialised, unlike FastSLAM 1.0's `> 100` (fs1.rs:144)
    m2, c2 = o.compute_proposal([1, 2, 0.3], [1.0, 0.1], (5.0, 0.2), [4, 4, 100.0, 0, 0, 100.0])
    assert np.array_equal(m2, mean) and np.array_equal(c2, cov)


def test_sample_pose_cholesky_and_fallback(L):
    """fs2.rs:219-239: mean + L n with the Cholesky factor; diagonal square roots when the matrix is not positive definite"""
    o = _oracle.OracleFS(L, 1, 1, variant=2)
    A = np.array([[4.0, 2.0, 0.6], [2.0, 5.0, 1.5], [0.6, 1.5, 3.0]])
    n = np.array([0.3, -1.2, 0.7])
    got = o.sample_pose([1.0, 2.0, 3.0], A, n)
    assert np.allclose(got, np.array([1.0, 2.0, 3.0]) + np.linalg.cholesky(A) @ n, rtol=1e-14)
    # hand case: diag(4, 9, 16) -> L = diag(2, 3, 4), exact
    assert np.array_equal(o.sample_pose([0, 0, 0], np.diag([4.0, 9.0, 16.0]), [1.0, 1.0, 1.0]), [2.0, 3.0, 4.0])
    # not positive definite (second pivot negative): fallback diag(sqrt(max(c_ii, 0)))
    B = np.array([[1.0, 2.0, 0.0], [2.0, 1.0, 0.0], [0.0, 0.0, -4.0]])
    assert np.array_equal(o.sample_pose([0, 0, 0], B, [1.0, 1.0, 1.0]), [1.0, 1.0, 0.0])
    # only the lower triangle is read (nalgebra's Cholesky works on the lower part)
    C2 = A.copy(); C2[0, 1] = 99.0; C2[0, 2] = -7.0; C2[1, 2] = 3.3
    assert np.array_equal(o.sample_pose([1.0, 2.0, 3.0], C2, n), got)


def test_update_does_not_panic_and_keeps_count(L):
    """fs2.rs:456-471: 20 particles, 3 landmarks, 5 steps"""
    o = _oracle.OracleFS(L, 20, 3, seed=7, variant=2)
    lms = [(10.0, 0.0), (0.0, 10.0), (10.0, 10.0)]
    for t in range(5):
        z = o.observations([0.0, 0.0, 0.0], lms, 7, t)
        o.step([1.0, 0.1], z)
    p, l = o.state()
    assert p.shape == (20, 4) and np.all(np.isfinite(p)) and np.all(np.isfinite(l))
    assert np.all(l[:, :, 2] < 100.0)                       # every landmark was initialised (cov = 10 I, then shrunk)


def test_landmark_convergence(L):
    """fs2.rs:511-544: 120 particles, one landmark at (5,5), 60 steps from (0,0,pi/4) with u = (0.5, 0): the weighted landmark
    estimate stays within 6 m.  The particles start with yaw 0 against a true yaw of pi/4, so the landmark is born ~5.4 m off
    and the bound is a "does not diverge" check; with OUR noise streams the error ranges 3.4 .. 6.8 m over seeds (the
    reference pins one ChaCha seed), so the check here is on the median over nine seeds."""
    errs = []
    for seed in range(11, 20):
        o = _oracle.OracleFS(L, 120, 1, seed=seed, variant=2)
        xt = np.array([0.0, 0.0, math.pi / 4])
        for t in range(60):
            xt = np.array([xt[0] + 0.5 * DT * math.cos(xt[2]), xt[1] + 0.5 * DT * math.sin(xt[2]), wrap(xt[2])])
            z = o.observations(xt, [(5.0, 5.0)], seed, t)
            o.step([0.5, 0.0], z)
        p, l = o.state()
        init = l[:, 0, 2] < 100.0
        assert init.any()
        w = p[init, 0]
        assert w.sum() > 0
        mx, my = np.sum(w * l[init, 0, 0]) / w.sum(), np.sum(w * l[init, 0, 1]) / w.sum()
        errs.append(math.hypot(mx - 5.0, my - 5.0))
    assert np.median(errs) < 6.0 and max(errs) < 8.0, errs


def test_first_step_initialises_with_cov_10_and_neutral_weight(L):
    """fs2.rs:250-256: a fresh landmark is placed from the SAMPLED pose, gets cov = 10 I and leaves the weight alone"""
    n = 8
    o = _oracle.OracleFS(L, n, 2, seed=3, variant=2)
    z0 = np.linspace(-1, 1, n); z1 = np.concatenate([np.linspace(1, -1, n), np.full(n, 0.5)])
    did = o.step([1.0, 0.0], [(5.0, 0.3, 1)], z0=z0, z1=z1, u01=0.5)
    p, l = o.state()
    assert did == 0 or did == 1
    # motion-only proposal (landmark uninitialised): cov = G M G^T with yaw = 0 -> L = chol([[.1,0,0],[0,.1+.01*.01,.001],[0,.001,.01]])
    G = np.array([[1, 0, 0.0], [0, 1, 0.1], [0, 0, 1.0]])
    Lc = np.linalg.cholesky(G @ MOTION_COV @ G.T)
    for i in range(n):
        pose = np.array([0.1, 0.0, 0.0]) + Lc @ np.array([z0[i], z1[i], z1[n + i]])
        assert np.allclose(p[i, 1:], pose, rtol=1e-13, atol=1e-15)
        assert np.allclose(l[i, 1, :2], [pose[0] + 5.0 * math.cos(pose[2] + 0.3), pose[1] + 5.0 * math.sin(pose[2] + 0.3)], rtol=1e-13)
        assert np.array_equal(l[i, 1, 2:], [10.0, 0.0, 0.0, 10.0])
        assert np.array_equal(l[i, 0, 2:], [1000.0, 0.0, 0.0, 1000.0])
    assert np.allclose(p[:, 0], 1.0 / n) or np.allclose(p[:, 0], 0.01 / (0.01 * n))
