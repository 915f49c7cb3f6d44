"""Pin the C oracle against the independent pure-Python restatement (tests/golden/make_golden.py) and the
hand-checkable known answers.  PARITY UNPINNED against the Rust reference itself: no
reference-generated vector can exist in this environment (see oracle/oracle.h)."""
import json
import os

import numpy as np
import pytest

from _oracle import OracleFS, OraclePF

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def unhex(v):
    if isinstance(v, list):
        return [unhex(a) for a in v]
    if isinstance(v, str):
        return float.fromhex(v)
    return v


def load(name):
    with open(os.path.join(G, name)) as f:
        return json.load(f)


def assert_close_ulp(a, b, rtol, what):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, what
    scale = np.maximum(np.abs(b), 1e-300)
    err = np.max(np.abs(a - b) / scale) if a.size else 0.0
    assert err <= rtol, f"{what}: max rel err {err:.3e} > {rtol:.1e}"


def run_pf_case(L, case, exact):
    n = case["n"]
    f = OraclePF(L, n, threshold=unhex(case["threshold"]), range_noise=unhex(case["sigma"]),
                 velocity_noise=unhex(case["sv"]), yaw_rate_noise=unhex(case["sw"]), dt=unhex(case["dt"]))
    f.set_particles(unhex(case["init"]))
    for t, st in enumerate(case["steps"]):
        f.predict(unhex(st["u"]), unhex(st["zv"]), unhex(st["zw"]))
        est_p, _ = f.estimate()
        f.update(np.array(unhex(st["obs"])).reshape(-1, 3))
        est_u, _ = f.estimate()
        w = f.particles()[:, 4]
        neff = f.neff()
        did = neff < n * unhex(case["threshold"])
        assert did == st["did_resample"], f"{case['name']} step {t}: resample gate"
        # slots 0,1 of step 1 carry rigged draws (r = 0 and r = 1-2^-53) that sit exactly on a cumsum edge: a
        # 1-ulp libm difference legitimately moves them, so the few-ulp (non-exact) comparison leaves them out
        keep = np.ones(n, dtype=bool)
        if not exact and t == 1:
            keep[:2] = False
        if did:
            f.resample(unhex(st["r"]))
            assert np.array_equal(f.last_indices()[keep], np.array(st["indices"])[keep]), f"{case['name']} step {t}: indices"
        est, cov = f.estimate()
        got = f.particles()
        want = np.array(unhex(st["particles"]))
        if not exact:
            got, want = got[keep], want[keep]
            if not keep.all():
                f.set_particles(np.array(unhex(st["particles"])))     # re-sync the trajectory after the rigged step
                est, cov = f.estimate()
        if exact:
            assert np.array_equal(got, want), f"{case['name']} step {t}: particles"
            assert np.array_equal(w, np.array(unhex(st["w_after_update"])))
            assert np.array_equal(est, np.array(unhex(st["est"])))
            assert np.array_equal(cov, np.array(unhex(st["cov"])))
            assert np.array_equal(est_p, np.array(unhex(st["est_after_predict"])))
            assert np.array_equal(est_u, np.array(unhex(st["est_after_update"])))
            assert neff == unhex(st["neff"])
        else:
            assert_close_ulp(got, want, 1e-9, f"{case['name']} step {t}: particles")
            assert_close_ulp(est, unhex(st["est"]), 1e-9, "est")
            np.testing.assert_allclose(cov, unhex(st["cov"]), rtol=1e-7, atol=1e-12)


@pytest.mark.parametrize("idx", range(4))
def test_pf_oracle_libm_bit_exact_vs_python(oracle_libm, idx):
    run_pf_case(oracle_libm, load("pf_golden.json")["cases"][idx], exact=True)


@pytest.mark.parametrize("idx", range(4))
def test_pf_oracle_contract_close_to_python(oracle, idx):
    run_pf_case(oracle, load("pf_golden.json")["cases"][idx], exact=False)


def run_mcl_case(L, case, exact):
    f = OraclePF(L, case["nmin"], range_noise=unhex(case["sigma"]), velocity_noise=unhex(case["sv"]),
                 yaw_rate_noise=unhex(case["sw"]), dt=unhex(case["dt"]), mode=1, max_particles=case["nmax"],
                 kld_epsilon=unhex(case["eps"]), kld_z=unhex(case["z"]))
    f.set_particles(unhex(case["init"]))
    for t, st in enumerate(case["steps"]):
        f.predict(unhex(st["u"]), unhex(st["zv"]), unhex(st["zw"]))
        f.update(np.array(unhex(st["obs"])).reshape(-1, 3))
        f.resample(unhex(st["r"]))
        assert f.count() == st["count"], f"{case['name']} step {t}: count"
        assert case["nmin"] <= f.count() <= case["nmax"]              # mcl.rs:573-574
        assert f.last_indices().tolist() == st["indices"], f"{case['name']} step {t}: indices"
        est, cov = f.estimate()
        got, want = f.particles(), np.array(unhex(st["particles"]))
        if exact:
            assert np.array_equal(got, want)
            assert np.array_equal(est, np.array(unhex(st["est"])))
            assert np.array_equal(cov, np.array(unhex(st["cov"])))
        else:
            assert_close_ulp(got, want, 1e-9, "particles")


@pytest.mark.parametrize("idx", range(2))
def test_mcl_oracle_libm_bit_exact_vs_python(oracle_libm, idx):
    run_mcl_case(oracle_libm, load("mcl_golden.json")["cases"][idx], exact=True)


@pytest.mark.parametrize("idx", range(2))
def test_mcl_oracle_contract_close_to_python(oracle, idx):
    run_mcl_case(oracle, load("mcl_golden.json")["cases"][idx], exact=False)


def run_fs_case(L, case, exact):
    n, m = case["n"], case["m"]
    cfg = {k: unhex(v) for k, v in case["cfg"].items()}
    f = OracleFS(L, n, m, **cfg)
    f.set_state(unhex(case["init_pose"]), unhex(case["init_lm"]))
    for t, st in enumerate(case["steps"]):
        if st["zero_weights"]:
            p, l = f.state()
            p[:, 0] = 0.0
            f.set_state(p, l)
        obs = [(unhex(o[0]), unhex(o[1]), o[2]) for o in st["obs"]]
        did = f.step(unhex(st["u"]), obs, unhex(st["z0"]), unhex(st["z1"]), unhex(st["u01"]))
        assert bool(did) == st["did_resample"], f"{case['name']} step {t}: gate"
        if did:
            assert f.last_indices().tolist() == st["indices"], f"{case['name']} step {t}: indices"
        assert f.best() == st["best"]
        p, l = f.state()
        wp, wl = np.array(unhex(st["pose"])), np.array(unhex(st["lm"]))
        if exact:
            assert np.array_equal(p, wp), f"{case['name']} step {t}: pose"
            assert np.array_equal(l, wl), f"{case['name']} step {t}: landmarks"
            assert f.last_neff() == unhex(st["neff"])
        else:
            np.testing.assert_allclose(p, wp, rtol=1e-9, atol=1e-12)
            np.testing.assert_allclose(l, wl, rtol=1e-8, atol=1e-11)


@pytest.mark.parametrize("idx", range(4))
def test_fs1_oracle_libm_bit_exact_vs_python(oracle_libm, idx):
    run_fs_case(oracle_libm, load("fs1_golden.json")["cases"][idx], exact=True)


@pytest.mark.parametrize("idx", range(4))
def test_fs1_oracle_contract_close_to_python(oracle, idx):
    run_fs_case(oracle, load("fs1_golden.json")["cases"][idx], exact=False)


def run_fs2_case(L, case, exact):
    """FastSLAM 2.0 (fs2.rs): the oracle's variant 2 against the independent Python restatement of tests/golden/make_golden.py"""
    n, m = case["n"], case["m"]
    cfg = {k: unhex(v) for k, v in case["cfg"].items()}
    f = OracleFS(L, n, m, variant=2, **cfg)
    f.set_state(unhex(case["init_pose"]), unhex(case["init_lm"]))
    for t, st in enumerate(case["steps"]):
        if st["zero_weights"]:
            p, l = f.state()
            p[:, 0] = 0.0
            f.set_state(p, l)
        obs = [(unhex(o[0]), unhex(o[1]), o[2]) for o in st["obs"]]
        z1 = np.concatenate([unhex(st["z1"]), unhex(st["z2"])])        # third draw of particle i rides behind z1 (oracle.h)
        did = f.step(unhex(st["u"]), obs, unhex(st["z0"]), z1, unhex(st["u01"]))
        assert bool(did) == st["did_resample"], f"{case['name']} step {t}: gate"
        if did:
            assert f.last_indices().tolist() == st["indices"], f"{case['name']} step {t}: indices"
        p, l = f.state()
        wp, wl = np.array(unhex(st["pose"])), np.array(unhex(st["lm"]))
        if exact:
            assert f.best() == st["best"]
            assert np.array_equal(p, wp), f"{case['name']} step {t}: pose"
            assert np.array_equal(l, wl), f"{case['name']} step {t}: landmarks"
            assert f.last_neff() == unhex(st["neff"])
        else:
            np.testing.assert_allclose(p, wp, rtol=1e-8, atol=1e-11)
            np.testing.assert_allclose(l, wl, rtol=1e-7, atol=1e-10)


@pytest.mark.parametrize("idx", range(4))
def test_fs2_oracle_libm_bit_exact_vs_python(oracle_libm, idx):
    run_fs2_case(oracle_libm, load("fs2_golden.json")["cases"][idx], exact=True)


@pytest.mark.parametrize("idx", range(4))
def test_fs2_oracle_contract_close_to_python(oracle, idx):
    run_fs2_case(oracle, load("fs2_golden.json")["cases"][idx], exact=False)


# ---------------- hand-checkable known answers ----------------
def test_kat_likelihood(oracle_libm):
    k = load("kat_golden.json")["likelihood_2x1"]
    f = OraclePF(oracle_libm, 2, range_noise=unhex(k["sigma"]))
    f.set_particles(unhex(k["particles"]))
    f.update([unhex(k["obs"])])
    assert f.particles()[:, 4].tolist() == unhex(k["normalised"])
    # closed form: raw_b = 1/sqrt(2 pi sigma^2), raw_a = raw_b * exp(-0.5)
    raw = unhex(k["raw"])
    assert raw[1] == pytest.approx(1.0 / np.sqrt(2 * np.pi * 0.25), rel=1e-15)
    assert raw[0] / raw[1] == pytest.approx(np.exp(-0.5), rel=1e-15)


def test_kat_index_rules(oracle_libm):
    k = load("kat_golden.json")["index_rules"]
    w, r = unhex(k["w"]), unhex(k["r"])
    part = np.zeros((4, 5))
    part[:, 0] = np.arange(4)
    part[:, 4] = w
    # hand-checked expectations: cum = [0.1, 0.30000000000000004, 0.6000000000000001, 1.0]
    assert k["pf"] == [0, 0, 1, 1, 2, 2, 3, 3, 0]      # r=1.5 above every cumsum -> PF falls back to index 0
    assert k["mcl"] == [0, 0, 1, 1, 2, 2, 3, 3, 3]     # ... MCL falls back to len-1
    got = []
    for rr in r:
        f = OraclePF(oracle_libm, 4, threshold=1.0)
        f.set_particles(part)
        f.resample([rr] * 4)
        got.append(int(f.last_indices()[0]))
    assert got == k["pf"]
    got = []
    for rr in r:
        f = OraclePF(oracle_libm, 4, mode=1, max_particles=4)
        f.set_particles(part)
        f.resample([rr] * 4)
        got.append(int(f.last_indices()[0]))
    assert got == k["mcl"]
    # FastSLAM systematic: r0 = 0.05 -> slots 0.05, 0.30, 0.55, 0.80 -> particles 0, 1, 2, 3
    assert k["fs"] == [0, 1, 2, 3]
    f = OracleFS(oracle_libm, 4, 1, nth=1e9)
    pose = np.zeros((4, 4))
    pose[:, 0] = w
    f.set_state(pose)
    assert f.step([0.0, 0.0], [], np.zeros(4), np.zeros(4), unhex(k["fs_u01"])) == 1
    assert f.last_indices().tolist() == k["fs"]


def test_kat_ekf_update(oracle_libm):
    k = load("kat_golden.json")["ekf_1"]
    f = OracleFS(oracle_libm, 1, 1, nth=0.0, q00=0.0, q11=0.0)
    f.set_state([[unhex(k["w0"]), 0.0, 0.0, 0.0]], [[[3.0, 4.0, 10.0, 0.0, 0.0, 10.0]]])
    z = unhex(k["z"])
    # u = 0, zero process noise -> the pose stays at the origin
    f.step([0.0, 0.0], [(z[0], z[1], 0)], [0.0], [0.0], 0.0)
    p, l = f.state()
    # the step normalises the single weight to 1; recover the raw product from the EKF output instead
    assert l[0, 0].tolist() == unhex(k["lm_after"])
    # independent numpy evaluation of the same update (textbook form)
    H = np.array([[3 / 5, 4 / 5], [-4 / 25, 3 / 25]])
    P = 10 * np.eye(2)
    S = H @ P @ H.T + np.diag([0.5, 0.0305])
    K = P @ H.T @ np.linalg.inv(S)
    y = np.array([z[0] - 5.0, 0.01])
    np.testing.assert_allclose(l[0, 0, :2], np.array([3.0, 4.0]) + K @ y, rtol=1e-12)
    np.testing.assert_allclose(l[0, 0, 2:].reshape(2, 2), (np.eye(2) - K @ H) @ P, rtol=1e-12, atol=1e-14)
    lik = np.exp(-0.5 * y @ np.linalg.inv(S) @ y) / (2 * np.pi * np.sqrt(np.linalg.det(S)))
    assert unhex(k["w_after"]) == pytest.approx(0.25 * lik, rel=1e-12)


# ---------------- value-level facts the reference's own tests assert ----------------
def test_reference_asserted_facts(oracle):
    # fs1.rs:386-398 initial constants
    f = OracleFS(oracle, 10, 4)
    p, l = f.state()
    assert np.all(p[:, 0] == 1.0 / 100.0) and np.all(p[:, 1:] == 0.0)
    assert np.all(l[:, :, 0:2] == 0.0) and np.all(l[:, :, 2] == 1000.0) and np.all(l[:, :, 5] == 1000.0)
    # pf.rs:621-622 sum w = 1 after update; pf.rs:582 count
    g = OraclePF(oracle, 100)
    g.init_state([0.0, 0.0, 0.0, 0.0])
    g.predict([1.0, 0.1])
    g.update([[5.0, 5.0, 0.0], [5.0, 0.0, 5.0]])
    assert abs(g.particles()[:, 4].sum() - 1.0) < 1e-3
    assert g.count() == 100
    # pf.rs:655-659 invalid config -> InvalidParameter
    with pytest.raises(ValueError):
        OraclePF(oracle, 0)
    with pytest.raises(ValueError):
        OraclePF(oracle, 10, range_noise=0.0)
    # proptest_filters.rs:79-88: empty observation list stays finite, weights uniform, no resample
    h = OraclePF(oracle, 50)
    for _ in range(20):
        est, did = h.step([1.0, 0.5], np.zeros((0, 3)))
        assert np.all(np.isfinite(est)) and did == 0
    # fs1.rs:372-378: 5 updates on 20 fresh particles keep len == 20 and (quirk B.3) never touch the weights
    k = OracleFS(oracle, 20, 3)
    lms = [(10.0, 0.0), (0.0, 10.0), (10.0, 10.0)]
    for t in range(5):
        k.step([1.0, 0.1], k.observations([0.0, 0.0, 0.0], lms, 7, t))
    p, l = k.state()
    assert p.shape[0] == 20 and np.all(l[:, :, 2] == 1000.0)
    assert np.all(p[:, 0] == 1.0 / 20.0) or np.all(p[:, 0] == p[0, 0])


def test_mcl_converges_like_reference_test(oracle):
    """mcl.rs:473-515: 60 steps, 4 landmarks, estimate within 1.0 m of truth."""
    f = OraclePF(oracle, 300, range_noise=0.25, velocity_noise=0.05, yaw_rate_noise=0.02, dt=0.1, mode=1,
                 max_particles=300)
    f.init_state([0.0, 0.0, 0.0, 1.0])
    lms = [(10.0, 0.0), (0.0, 10.0), (-10.0, 0.0), (0.0, -10.0)]
    x = np.zeros(3)
    for t in range(60):
        u = [1.0, 0.03]
        x[0] += u[0] * np.cos(x[2]) * 0.1
        x[1] += u[0] * np.sin(x[2]) * 0.1
        x[2] += u[1] * 0.1
        obs = [[np.hypot(x[0] - lx, x[1] - ly), lx, ly] for lx, ly in lms]
        est, _ = f.step(u, obs)
    assert np.hypot(est[0] - x[0], est[1] - x[1]) < 1.0
