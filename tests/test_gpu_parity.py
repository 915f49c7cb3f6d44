"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on identical seeded inputs.

Bar (BASELINE.json north_star): resample indices bit-exact; pose / weight / covariance within 1e-6 relative.
Because kernels and oracle share include/pf_contract_math.h and the device reproduces the reference's
sequential f64 sums exactly (xsum), particles, weights, landmarks and indices are in fact compared for
BIT EQUALITY; only estimate/covariance (tree-order sums on the device) use the 1e-6 tolerance.
"""
import ctypes as C
import os

import numpy as np
import pytest

import rust_robotics_b200 as rr
from rust_robotics_b200 import scenarios
from _oracle import OracleFS, OraclePF

pytestmark = pytest.mark.gpu

RTOL = 1e-6   # north_star tolerance for floating-point summaries


def assert_cov_close(got, want, what=""):
    got, want = np.asarray(got).reshape(4, 4), np.asarray(want).reshape(4, 4)
    # relative to sqrt(var_i var_j); a state component without spread (e.g. v when velocity_noise == 0) has
    # var ~ 1e-31 of pure rounding noise, so the scale is floored at 1e-9 of the largest variance
    d = np.maximum(np.abs(np.diag(want)), 1e-9 * np.max(np.abs(np.diag(want))) + 1e-300)
    scale = np.sqrt(np.outer(d, d))
    err = np.max(np.abs(got - want) / scale)
    assert err < RTOL, f"covariance {what}: {err}"


# ------------------------------------------------------------------------------------------------
# exact scan primitive
# ------------------------------------------------------------------------------------------------
def _xsum_cases():
    rng = np.random.default_rng(3)
    yield "uniform_65536", np.full(65536, 1.0 / 65536)
    yield "uniform_non_pow2", np.full(100003, 1.0 / 100003)
    w = rng.uniform(size=1 << 20); yield "random_1M", w / w.sum()
    w = np.exp(rng.normal(0, 12, 300000)); yield "lognormal_wide", w / w.sum()
    w = np.exp(rng.normal(0, 40, 300000)); yield "collapse", w / w.sum()
    w = np.full(50000, 2.0 ** -54); w[0] = 1.0; yield "ties", w
    w = np.full(50000, 2.0 ** -60); w[0] = 1.0 - 2.0 ** -40; yield "crawl_edge", w
    w = np.zeros(70000); w[35000:] = rng.uniform(size=35000); yield "leading_zeros", w
    yield "all_zero", np.zeros(5000)
    yield "subnormal", rng.uniform(size=30000) * 1e-310
    yield "comb", np.concatenate([[0.37 / 65536], np.full(65535, 1.0 / 65536)])
    yield "single", np.array([0.7])
    yield "tiny_n", rng.uniform(size=17)
    w = rng.uniform(size=4097); yield "tile_plus_one", w


@pytest.mark.parametrize("name,v", list(_xsum_cases()), ids=[c[0] for c in _xsum_cases()])
def test_xsum_device_bit_exact(name, v):
    L = rr.load_library()
    v = np.ascontiguousarray(v, dtype=np.float64)
    out = np.empty_like(v)
    tot = C.c_double()
    flags = (C.c_int * 4)()
    rc = L.pfgpu_test_xsum(v.ctypes.data_as(rr.api.c_dp), v.size, out.ctypes.data_as(rr.api.c_dp), C.byref(tot), flags, 0)
    assert rc == 0, L.pfgpu_last_error()
    ref = np.add.accumulate(v)          # strictly sequential left-to-right accumulation
    assert flags[0] == 0, "unexpected serial fallback"
    assert flags[2] == 0, "emit-time certificate failure"
    assert np.array_equal(out, ref), f"{name}: first diff at {np.flatnonzero(out != ref)[:5]}"
    assert tot.value == ref[-1]


def test_xsum_device_serial_fallback_on_bad_values():
    L = rr.load_library()
    v = np.random.default_rng(1).uniform(size=10000)
    v[1234] = np.nan
    out = np.empty_like(v)
    tot = C.c_double()
    flags = (C.c_int * 4)()
    assert L.pfgpu_test_xsum(v.ctypes.data_as(rr.api.c_dp), v.size, out.ctypes.data_as(rr.api.c_dp), C.byref(tot), flags, 0) == 0
    ref = np.add.accumulate(v)
    assert flags[0] == 1
    assert np.array_equal(out[:1234], ref[:1234]) and np.isnan(out[1234:]).all()


def test_device_division_is_ieee_exact():
    """PFC_DIV on the device (correctly rounded reciprocal + two fma corrections, guarded) must equal the IEEE quotient for
    every operand pair: 2e9 random pairs over exponents +-400 incl. all-ones / sparse mantissas and signed zeros."""
    L = rr.load_library()
    bad = C.c_ulonglong(12345)
    assert L.pfgpu_test_div(2_000_000_000, 99, C.byref(bad), 0) == 0, L.pfgpu_last_error()
    assert bad.value == 0


# ------------------------------------------------------------------------------------------------
# ParticleFilterLocalizer / MonteCarloLocalizer
# ------------------------------------------------------------------------------------------------
def _pf_pair(oracle, n, thr=0.5, sigma=0.25, sv=2.0, sw=np.deg2rad(40.0), mode=0, seed=42, init=(5.0, 5.0, 0.0, 0.0)):
    if mode == 0:
        g = rr.ParticleFilterLocalizer.try_with_initial_state(
            init, rr.ParticleFilterConfig(n, thr, sigma, sv, sw, 0.1), seed=seed)
    else:
        g = rr.MonteCarloLocalizer.try_with_initial_state(
            init, rr.MonteCarloLocalizationConfig(n, n, 0.05, 2.326, sigma, sv, sw, 0.1), seed=seed)
    o = OraclePF(oracle, n, threshold=thr, range_noise=sigma, velocity_noise=sv, yaw_rate_noise=sw, dt=0.1, seed=seed,
                 mode=mode, max_particles=n)
    o.L.orc_pf_set_fast_search(o.h, 1)
    o.init_state(init)
    return g, o


def _pf_compare(g, o, what):
    gp, op = g.get_particles(), o.particles()
    assert np.array_equal(gp, op), f"{what}: particles differ at rows {np.flatnonzero((gp != op).any(axis=1))[:5]}"
    ge, (oe, oc) = g.estimate(), o.estimate()
    np.testing.assert_allclose(ge, oe, rtol=RTOL, atol=1e-9, err_msg=what)
    assert_cov_close(g.calc_covariance().T.ravel(), oc, what)      # both column-major


@pytest.mark.parametrize("n,thr,steps", [(1000, 0.5, 300), (1000, 1.0, 40), (4099, 0.5, 60), (65536, 0.5, 25), (1 << 18, 1.0, 6)])
def test_pf_trajectory_bit_exact(oracle, n, thr, steps):
    """C1 scenario (render_gif_particle_filter.rs:21-79) at several particle counts."""
    sc = scenarios.PfScenario("c1", steps=steps)
    g, o = _pf_pair(oracle, n, thr=thr)
    _pf_compare(g, o, "init")
    resamples = 0
    for t in range(steps):
        ge = g.try_step(sc.controls[t], sc.obs[t])
        oe, did = o.step(sc.controls[t], sc.obs[t])
        np.testing.assert_allclose(ge, oe, rtol=RTOL, atol=1e-9)
        if did:
            resamples += 1
            assert np.array_equal(g.last_indices(), o.last_indices()), f"step {t}: resample indices"
        if t % 10 == 0 or t == steps - 1:
            _pf_compare(g, o, f"step {t}")
    assert g.stats().resamples == resamples
    assert resamples > 0
    assert g.stats().serial_fallbacks == 0
    # the reference's own assertions: finite estimate, sum w = 1 (pf.rs:621-635)
    assert np.all(np.isfinite(g.estimate())) and abs(g.get_particles()[:, 4].sum() - 1.0) < 1e-3


@pytest.mark.parametrize("fused,graph", [("0", "1"), ("0", "0"), ("1", "0")])
@pytest.mark.parametrize("kind", ["pf", "mcl"])
def test_pf_step_paths_agree_with_oracle(oracle, monkeypatch, kind, fused, graph):
    """The fused step has three host-side forms: predict + ONE cooperative tail launch (pf3.cuh, the default up to 2^18 particles),
    the ~22 separate launches replayed from a CUDA graph, and the same launches issued one by one.  The parametrised trajectory
    tests run the default; this one pins the other two (and the fused form without the graph) to the same oracle."""
    monkeypatch.setenv("PFGPU_PF_FUSED", fused)
    monkeypatch.setenv("PFGPU_PF_GRAPH", graph)
    steps = 40
    if kind == "pf":
        sc = scenarios.PfScenario("c1", steps=steps)
        g, o = _pf_pair(oracle, 3000, thr=0.6)
    else:
        sc = scenarios.PfScenario("c2", steps=steps)
        g, o = _pf_pair(oracle, 2048, sigma=0.25, sv=0.05, sw=0.02, mode=1, seed=5, init=tuple(sc.init))
    resamples = 0
    for t in range(steps):
        obs = sc.obs[t][:: max(1, sc.obs[t].shape[0] // 24)] if kind == "mcl" else sc.obs[t]     # <= 32 observations: rides in the launch parameters
        if t == 17:
            obs = obs[:-1]                                          # a different observation count re-captures the graph
        ge = g.try_step(sc.controls[t], obs)
        oe, did = o.step(sc.controls[t], obs)
        np.testing.assert_allclose(ge, oe, rtol=RTOL, atol=1e-9)
        if did:
            resamples += 1
            assert np.array_equal(g.last_indices(), o.last_indices()), f"step {t}: resample indices"
        if t % 8 == 0 or t == steps - 1:
            _pf_compare(g, o, f"step {t}")
    assert resamples > 0
    launches = g.stats().kernel_launches
    assert (launches < 4 * steps) == (fused == "1"), launches      # 2 launches per step when fused, ~22 otherwise


def test_pf_phase_api_matches_oracle(oracle):
    """predict / update / resample called one by one (the StateEstimator route pf.rs:552-573): caches after each phase."""
    sc = scenarios.PfScenario("c1", steps=12)
    g, o = _pf_pair(oracle, 2048, thr=0.9)
    for t in range(12):
        g.try_predict_with_control(sc.controls[t]); o.predict(sc.controls[t])
        _pf_compare(g, o, f"predict {t}")
        g.try_update_with_observations(sc.obs[t]); o.update(sc.obs[t])
        _pf_compare(g, o, f"update {t}")
        assert g.n_eff() == o.neff()
        assert g.resample() == bool(o.resample())
        _pf_compare(g, o, f"resample {t}")


def test_pf_edge_cases(oracle):
    g, o = _pf_pair(oracle, 500)
    # empty observation list (proptest_filters.rs:79-88): weight exactly 1 -> uniform -> no resample
    for _ in range(5):
        g.try_step([1.0, 0.5], np.zeros((0, 3))); o.step([1.0, 0.5], np.zeros((0, 3)))
    _pf_compare(g, o, "empty obs")
    assert g.stats().resamples == 0
    # every likelihood underflows -> sum_w == 0 -> uniform fallback (pf.rs:433-438)
    far = [[1.0e4, 0.0, 0.0], [1.0e4, 5.0, 5.0]]
    g.try_step([1.0, 0.0], far); o.step([1.0, 0.0], far)
    _pf_compare(g, o, "underflow")
    assert np.all(g.get_particles()[:, 4] == 1.0 / 500)
    # zero process noise (no draw at all when sigma == 0: pf.rs:259-276)
    g2, o2 = _pf_pair(oracle, 300, sv=0.0, sw=0.0)
    sc = scenarios.PfScenario("c1", steps=8)
    for t in range(8):
        g2.try_step(sc.controls[t], sc.obs[t]); o2.step(sc.controls[t], sc.obs[t])
    _pf_compare(g2, o2, "zero noise")
    # every likelihood tiny but not zero (co-located particles ~7 m off the measured range: exp(-392) ~ 1e-170 each): the
    # normalised weights are exactly uniform, so N_eff = n and nothing resamples — while the SQUARES of the raw weights all
    # underflow, which is what the tree-order N_eff shortcut would have looked at (it must hand over to the exact sum here)
    g3, o3 = _pf_pair(oracle, 400, thr=0.5, sigma=0.25, sv=0.0, sw=0.0)
    same = np.tile([5.0, 5.0, 0.0, 0.0, 1.0 / 400], (400, 1))
    g3.set_particles(same); o3.set_particles(same)
    tiny = [[10.0 + 6.99, 15.0, 5.0]]                             # landmark 10 m away, range reading 16.99
    ge = g3.try_step([0.0, 0.0], tiny); oe, did3 = o3.step([0.0, 0.0], tiny)
    assert not did3 and g3.stats().resamples == 0
    assert np.array_equal(g3.get_particles(), o3.particles()), "tiny uniform weights"     # (the covariance of 400 identical particles is
    np.testing.assert_allclose(ge, oe, rtol=RTOL, atol=1e-9)                                # rounding noise on both sides: not compared)
    assert np.allclose(g3.get_particles()[:, 4], 1.0 / 400, rtol=1e-12)      # (w_raw / S with a sequentially rounded S: uniform up to an ulp or two)
    # validation (pf.rs:515-549, 81-117)
    with pytest.raises(rr.InvalidParameter):
        g.try_step([np.nan, 0.0], far)
    with pytest.raises(rr.InvalidParameter):
        g.try_step([1.0, 0.0], [[-1.0, 0.0, 0.0]])
    with pytest.raises(rr.InvalidParameter):
        rr.ParticleFilterLocalizer(rr.ParticleFilterConfig(n_particles=0))
    with pytest.raises(rr.InvalidParameter):
        rr.ParticleFilterLocalizer(rr.ParticleFilterConfig(range_noise=0.0))
    with pytest.raises(rr.InvalidParameter):
        g.set_range_noise(-1.0)


@pytest.mark.parametrize("n,k", [(4096, 4), (1 << 16, 360)])
def test_mcl_fixed_n_bit_exact(oracle, n, k):
    """MonteCarloLocalizer with min == max (BASELINE config 2 shape): resamples every step, last cumsum forced to 1."""
    sc = scenarios.PfScenario("c2", steps=10)
    g, o = _pf_pair(oracle, n, sigma=0.25, sv=0.05, sw=0.02, mode=1, init=(0.0, 0.0, 0.0, 1.0))
    for t in range(10):
        obs = sc.obs[t][:: 360 // k][:k]
        ge = g.try_step(sc.controls[t], obs)
        oe, _ = o.step(sc.controls[t], obs)
        assert np.array_equal(g.last_indices(), o.last_indices()), f"step {t}"
        np.testing.assert_allclose(ge, oe, rtol=RTOL, atol=1e-9)
    _pf_compare(g, o, "mcl end")
    assert g.particle_count() == n and g.stats().resamples == 10


def test_mcl_converges_like_reference_test():
    """mcl.rs:473-515: 60 steps, estimate within 1.0 m of the truth."""
    g = rr.MonteCarloLocalizer.try_with_initial_state(
        [0.0, 0.0, 0.0, 1.0], rr.MonteCarloLocalizationConfig(300, 300, 0.05, 2.326, 0.25, 0.05, 0.02, 0.1))
    lms = [(10.0, 0.0), (0.0, 10.0), (-10.0, 0.0), (0.0, -10.0)]
    x = np.zeros(3)
    for _ in range(60):
        x[0] += np.cos(x[2]) * 0.1; x[1] += np.sin(x[2]) * 0.1; x[2] += 0.03 * 0.1
        est = g.try_step([1.0, 0.03], [[np.hypot(x[0] - lx, x[1] - ly), lx, ly] for lx, ly in lms])
    assert np.hypot(est[0] - x[0], est[1] - x[1]) < 1.0


@pytest.mark.parametrize("nmin,nmax,k,eps", [(100, 5000, 4, 0.05), (50, 400, 8, 0.05), (1000, 200000, 4, 0.0005), (64, 64 + 1, 4, 0.05)])
def test_mcl_kld_adaptive_bit_exact(oracle, nmin, nmax, k, eps):
    """resample_adaptive with min < max (mcl.rs:322-365; 100 / 5000 is the reference's default config): the particle count
    follows the KLD bound; count, ancestry and particles equal the oracle's at every step."""
    init = (0.0, 0.0, 0.0, 1.0)
    g = rr.MonteCarloLocalizer.try_with_initial_state(
        init, rr.MonteCarloLocalizationConfig(nmin, nmax, eps, 2.326, 0.25, 0.05, 0.02, 0.1), seed=3)
    o = OraclePF(oracle, nmin, range_noise=0.25, velocity_noise=0.05, yaw_rate_noise=0.02, dt=0.1, seed=3, mode=1, max_particles=nmax,
                 kld_epsilon=eps)
    o.L.orc_pf_set_fast_search(o.h, 1)
    o.init_state(init)
    sc = scenarios.PfScenario("c2", steps=25)
    counts = []
    for t in range(25):
        obs = sc.obs[t][:: 360 // k][:k]
        ge = g.try_step(sc.controls[t], obs)
        oe, _ = o.step(sc.controls[t], obs)
        assert g.particle_count() == o.count(), f"step {t}: count {g.particle_count()} vs oracle {o.count()} (history {counts})"
        assert np.array_equal(g.last_indices(), o.last_indices()), f"step {t}"
        np.testing.assert_allclose(ge, oe, rtol=RTOL, atol=1e-9)
        counts.append(o.count())
    _pf_compare(g, o, "kld end")
    assert min(counts) >= nmin and max(counts) <= nmax and (len(set(counts)) > 1 or nmax == nmin + 1), counts


# ------------------------------------------------------------------------------------------------
# FastSLAM 1.0
# ------------------------------------------------------------------------------------------------
def _fs_compare(g, o, what, landmarks=True):
    gp, gl = g.state(landmarks)
    op, ol = o.state()
    assert np.array_equal(gp, op), f"{what}: pose/weight rows {np.flatnonzero((gp != op).any(axis=1))[:5]}"
    if landmarks:
        assert np.array_equal(gl, ol), f"{what}: landmarks differ for particles {np.flatnonzero((gl != ol).any(axis=(1, 2)))[:5]}"


@pytest.mark.parametrize("n,side,steps", [(1024, 6, 90), (4096, 8, 60), (1000, 5, 70)])
def test_fastslam_generation_rows_bit_exact(oracle, n, side, steps):
    """long runs on small grids: landmarks leave and re-enter the view across many resamples, so several ancestry rows
    (one per inter-resample period with a not-since-observed landmark) are alive at once (fs3.cuh, lazy clone by generation)"""
    sc = scenarios.FastSlamScenario(side, (10.0 * side / 2 - 5.0, 10.0 * side / 2 - 15.0, 0.0), (8.0, 0.8), steps, max_range=12.0)
    g = rr.FastSlam1(n, sc.m, rr.FsConfig(nth=n / 1.5), seed=7)
    o = OracleFS(oracle, n, sc.m, seed=7, nth=n / 1.5)
    g.seed_map(sc.start, sc.landmarks); o.seed_map(sc.start, sc.landmarks)
    resamples = 0
    for t in range(steps):
        did = g.fastslam_update(sc.control, sc.obs[t])
        assert did == bool(o.step(sc.control, sc.obs[t])), f"step {t}"
        if did:
            resamples += 1
            assert np.array_equal(g.last_indices(), o.last_indices()), f"step {t}"
        if t % 7 == 0:
            _fs_compare(g, o, f"step {t}")
    _fs_compare(g, o, "end")
    assert resamples > 5
    assert g.stats().serial_fallbacks == 0


@pytest.mark.parametrize("n,side,steps", [(64, 4, 30), (1000, 6, 25), (4096, 8, 20), (1 << 16, 16, 4)])
def test_fastslam_trajectory_bit_exact(oracle, n, side, steps):
    """C3-shaped runs (initialised map, nth = n/1.5): pose, weights, every landmark EKF state and the resample
    ancestry equal the oracle's bit for bit."""
    sc = scenarios.FastSlamScenario(side, (10.0 * side / 2 - 5.0, 10.0 * side / 2 - 5.0, 0.0), (1.0, 0.025), steps)
    cfg = rr.FsConfig(nth=n / 1.5)
    g = rr.FastSlam1(n, sc.m, cfg, seed=7)
    o = OracleFS(oracle, n, sc.m, seed=7, nth=n / 1.5)
    g.seed_map(sc.start, sc.landmarks); o.seed_map(sc.start, sc.landmarks)
    _fs_compare(g, o, "seed")
    resamples = 0
    for t in range(steps):
        did = g.fastslam_update(sc.control, sc.obs[t])
        odid = o.step(sc.control, sc.obs[t])
        assert did == bool(odid), f"step {t}: gate (neff gpu {g.last_neff()} oracle {o.last_neff()})"
        # neff itself is reported from the tree-order sum unless it is within rounding of NTH (then the exact one decides)
        assert g.last_neff() == pytest.approx(o.last_neff(), rel=1e-9)
        if did:
            resamples += 1
            idx = g.last_indices()
            oi = o.last_indices()
            assert np.array_equal(idx, oi), f"step {t}: indices differ at {np.flatnonzero(idx != oi)[:8]} ({int((idx != oi).sum())} slots): {idx[idx != oi][:8]} vs {oi[idx != oi][:8]}"
            assert np.all(np.diff(idx.astype(np.int64)) >= 0)          # systematic resampling is monotone
        if n <= 4096 or t == steps - 1:
            _fs_compare(g, o, f"step {t}")
        bi, bp = g.get_best_particle()
        assert bi == o.best()
    assert resamples > 0, "scenario never resampled"
    assert g.stats().serial_fallbacks == 0


@pytest.mark.parametrize("tiles,n,nt", [(2, 4096, 256), (3, 5000, 256), (1, 2048, 512), (5, 1 << 14, 512)])
def test_fastslam_post_kernel_shapes_bit_exact(oracle, tiles, n, nt, monkeypatch):
    """the fused post kernel with several values per thread and few tiles (the shape large particle counts get), at sizes
    the oracle checks in seconds"""
    monkeypatch.setenv("PFGPU_POST_TILES", str(tiles))
    monkeypatch.setenv("PFGPU_POST_NT", str(nt))
    sc = scenarios.FastSlamScenario(6, (25.0, 25.0, 0.0), (1.0, 0.025), 16)
    g = rr.FastSlam1(n, sc.m, rr.FsConfig(nth=n / 1.5), seed=3)
    o = OracleFS(oracle, n, sc.m, seed=3, nth=n / 1.5)
    g.seed_map(sc.start, sc.landmarks); o.seed_map(sc.start, sc.landmarks)
    resamples = 0
    for t in range(16):
        did = g.fastslam_update(sc.control, sc.obs[t])
        assert did == bool(o.step(sc.control, sc.obs[t])), f"step {t}"
        if did:
            resamples += 1
            idx, oi = g.last_indices(), o.last_indices()
            assert np.array_equal(idx, oi), f"step {t}: {int((idx != oi).sum())} indices differ, first at {np.flatnonzero(idx != oi)[:4]}"
    _fs_compare(g, o, "end")
    assert resamples > 0 and g.stats().serial_fallbacks == 0


def test_fastslam_reference_constants_and_fresh_particles(oracle):
    """create_particles as shipped (fs1.rs:302-306, tests fs1.rs:362-398): cov stays 1000 I, weights untouched (App. B.3),
    NTH = 66.67 so 20 particles resample every step."""
    g = rr.FastSlam1(20, 3, seed=3)
    o = OracleFS(oracle, 20, 3, seed=3)
    p, l = g.state()
    assert np.all(p[:, 0] == 0.01) and np.all(p[:, 1:] == 0.0)
    assert np.all(l[:, :, 2] == 1000.0) and np.all(l[:, :, 5] == 1000.0) and np.all(l[:, :, :2] == 0.0)
    lms = [(10.0, 0.0), (0.0, 10.0), (10.0, 10.0)]
    rng = np.random.default_rng(0)
    for t in range(5):
        z = scenarios.get_observations([0.0, 0.0, 0.0], lms, rng)
        assert g.fastslam_update([1.0, 0.1], z) == bool(o.step([1.0, 0.1], z))
        _fs_compare(g, o, f"fresh {t}")
    p, l = g.state()
    assert p.shape[0] == 20 and np.all(l[:, :, 2] == 1000.0)


def test_fastslam_edge_cases(oracle):
    n, m = 256, 4
    lm_xy = np.array([[5.0, 0.0], [0.0, 5.0], [5.0, 5.0], [-5.0, 2.0]])
    g = rr.FastSlam1(n, m, rr.FsConfig(nth=n / 1.5), seed=11)
    o = OracleFS(oracle, n, m, seed=11, nth=n / 1.5)
    g.seed_map([0.0, 0.0, 0.0], lm_xy); o.seed_map([0.0, 0.0, 0.0], lm_xy)
    # duplicate lm_id in one observation list: sequential EKF updates of the same landmark (App. A8)
    z = [(5.1, 0.02, 0), (5.0, 1.55, 1), (4.9, -0.01, 0)]
    assert g.fastslam_update([1.0, 0.1], z) == bool(o.step([1.0, 0.1], z))
    _fs_compare(g, o, "duplicate ids")
    # empty observation list
    assert g.fastslam_update([1.0, 0.1], []) == bool(o.step([1.0, 0.1], []))
    _fs_compare(g, o, "empty obs")
    # all-zero weights: no normalisation, neff = 0 -> resample -> every slot clones particle n-1 (App. B.6)
    p, l = g.state()
    p[:, 0] = 0.0
    g.set_state(p, l); o.set_state(p, l)
    assert g.fastslam_update([1.0, 0.1], z[:2]) is True
    assert o.step([1.0, 0.1], z[:2]) == 1
    assert np.all(g.last_indices() == n - 1)
    _fs_compare(g, o, "zero weights")
    # tiny uniform weights (1e-170 each, e.g. after an outlier observation): normalised they are 1/n, N_eff = n, no resample —
    # but every raw square underflows, so the tree-order N_eff shortcut must hand over to the exact sum
    p, l = g.state()
    p[:, 0] = 1e-170
    g.set_state(p, l); o.set_state(p, l)
    assert g.fastslam_update([1.0, 0.1], []) is False and o.step([1.0, 0.1], []) == 0
    _fs_compare(g, o, "tiny uniform weights")
    assert g.last_neff() == pytest.approx(float(n), rel=1e-12)
    # mixed fresh (cov 1000) and initialised landmarks through upload
    p, l = g.state()
    l[:, 2, :] = [0.0, 0.0, 1000.0, 0.0, 0.0, 1000.0]
    g.set_state(p, l); o.set_state(p, l)
    z = [(7.0, 0.8, 2), (5.0, 0.1, 0)]
    assert g.fastslam_update([1.0, 0.0], z) == bool(o.step([1.0, 0.0], z))
    _fs_compare(g, o, "mixed fresh")
    # validation: lm_id out of range (the reference would panic on the Vec index), non-finite control
    with pytest.raises(rr.InvalidParameter):
        g.fastslam_update([1.0, 0.0], [(1.0, 0.0, 99)])
    with pytest.raises(rr.InvalidParameter):
        g.fastslam_update([np.inf, 0.0], [])


def test_fastslam_long_observation_list_uses_memcpy_path(oracle):
    n, side = 512, 8
    sc = scenarios.FastSlamScenario(side, (35.0, 35.0, 0.0), (1.0, 0.025), 3, max_range=60.0)   # sees all 64 landmarks
    assert len(sc.obs[0]) > 48
    g = rr.FastSlam1(n, sc.m, rr.FsConfig(nth=n / 1.5), seed=5)
    o = OracleFS(oracle, n, sc.m, seed=5, nth=n / 1.5)
    g.seed_map(sc.start, sc.landmarks); o.seed_map(sc.start, sc.landmarks)
    for t in range(3):
        assert g.fastslam_update(sc.control, sc.obs[t]) == bool(o.step(sc.control, sc.obs[t]))
    _fs_compare(g, o, "long list")


def test_fastslam_full_size_properties():
    """BASELINE config 3 at full size (65 536 x 256): size-independent properties over a short run."""
    n = 1 << 16
    sc = scenarios.c3_scenario(steps=12)
    g = rr.FastSlam1(n, sc.m, rr.FsConfig(nth=n / 1.5), seed=42)
    g.seed_map(sc.start, sc.landmarks)
    resampled = 0
    for t in range(12):
        before = g.state(landmarks=False)[0]
        did = g.fastslam_update(sc.control, sc.obs[t])
        p = g.state(landmarks=False)[0]
        assert np.all(np.isfinite(p)) and np.all(p[:, 0] >= 0.0)
        assert abs(p[:, 0].sum() - 1.0) < 1e-9                      # normalised
        assert np.all(np.abs(p[:, 3]) <= np.pi + 1e-12)             # yaw wrapped every step (fs1.rs:75)
        if did:
            resampled += 1
            idx = g.last_indices().astype(np.int64)
            assert np.all(np.diff(idx) >= 0) and idx.min() >= 0 and idx.max() < n
            assert np.all(p[:, 0] == 1.0 / n)                       # fs1.rs:228
            # gather check on a sample: the clone carries its ancestor's whole map
            for slot in (0, n // 3, n - 1):
                a = g.particle_landmarks(slot)
                assert np.all(np.isfinite(a))
        else:
            assert 1.0 / np.sum(p[:, 0] ** 2) >= n / 1.5 - 1e-6
    assert resampled > 0
    # determinism: same seed, same inputs -> identical state
    g2 = rr.FastSlam1(n, sc.m, rr.FsConfig(nth=n / 1.5), seed=42)
    g2.seed_map(sc.start, sc.landmarks)
    for t in range(12):
        g2.fastslam_update(sc.control, sc.obs[t], want_flag=False)
    assert np.array_equal(g.state(landmarks=False)[0], g2.state(landmarks=False)[0])
    assert np.array_equal(g.particle_landmarks(12345), g2.particle_landmarks(12345))


def test_fastslam_full_size_matches_oracle(oracle):
    """65 536 x 256, three steps, everything (805 MB of landmark state) compared bit for bit."""
    n = 1 << 16
    sc = scenarios.c3_scenario(steps=3)
    g = rr.FastSlam1(n, sc.m, rr.FsConfig(nth=n / 1.5), seed=42)
    o = OracleFS(oracle, n, sc.m, seed=42, nth=n / 1.5)
    o.L.orc_fs_set_threads(o.h, 8)
    g.seed_map(sc.start, sc.landmarks); o.seed_map(sc.start, sc.landmarks)
    for t in range(3):
        assert g.fastslam_update(sc.control, sc.obs[t]) == bool(o.step(sc.control, sc.obs[t]))
    _fs_compare(g, o, "full size")


# ---------------------------------------------------------------------------------------------------------------------
# the SHARDED engine on one GPU: all ranks inside this process, on the same device (pfgpu_fs_create_sharded_local).  Every
# cross-rank path runs — weights pushed into every rank's copy, arrive / done flags, ancestors and ancestry rows read from
# another rank's arena, lazy import of a remote map at the next EKF update — and every rank's shard must equal the oracle's
# slice bit for bit.  (tests/test_gpu_multi.py runs the same engine with one process per GPU when several GPUs exist.)
# ---------------------------------------------------------------------------------------------------------------------
def _shard_compare(ranks, o, what):
    op, ol = o.state()
    for r, g in enumerate(ranks):
        lo, hi = r * g.n_local, (r + 1) * g.n_local
        gp, gl = g.state()
        assert np.array_equal(gp, op[lo:hi]), f"{what}: rank {r} pose/weight rows {np.flatnonzero((gp != op[lo:hi]).any(axis=1))[:5]}"
        assert np.array_equal(gl, ol[lo:hi]), f"{what}: rank {r} landmarks differ for particles {np.flatnonzero((gl != ol[lo:hi]).any(axis=(1, 2)))[:5]}"


@pytest.mark.parametrize("world,n,side,steps,fast", [(2, 1024, 6, 60, True), (4, 4096, 6, 40, True), (2, 1 << 14, 8, 12, False), (8, 4096, 5, 30, True)])
def test_fastslam_sharded_in_process_bit_exact(oracle, world, n, side, steps, fast):
    if fast:   # landmarks leave and re-enter the view: remote references survive several resamples
        sc = scenarios.FastSlamScenario(side, (10.0 * side / 2 - 5.0, 10.0 * side / 2 - 15.0, 0.0), (8.0, 0.8), steps, max_range=12.0)
    else:
        sc = scenarios.FastSlamScenario(side, (10.0 * side / 2 - 5.0, 10.0 * side / 2 - 5.0, 0.0), (1.0, 0.025), steps)
    ranks = rr.FastSlam1.create_sharded_local(n, sc.m, [0] * world, rr.FsConfig(nth=n / 1.5), seed=9)
    o = OracleFS(oracle, n, sc.m, seed=9, nth=n / 1.5)
    for g in ranks:
        g.seed_map(sc.start, sc.landmarks)
    o.seed_map(sc.start, sc.landmarks)
    assert all(g.shard_mode() == 2 for g in ranks)
    resamples = 0
    for t in range(steps):
        did = rr.FastSlam1.step_all(ranks, sc.control, sc.obs[t])
        assert did == bool(o.step(sc.control, sc.obs[t])), f"step {t}: gate"
        assert all(g.last_gate() == did for g in ranks)
        if did:
            resamples += 1
            idx = np.concatenate([g.last_indices() for g in ranks])
            assert np.array_equal(idx, o.last_indices()), f"step {t}: indices"
        for g in ranks:
            assert g.get_best_particle()[0] == o.best(), f"step {t}: best particle"
        if t % 9 == 0:
            _shard_compare(ranks, o, f"step {t}")
    _shard_compare(ranks, o, "end")
    assert resamples > 2
    assert all(g.stats().serial_fallbacks == 0 for g in ranks)


def test_fastslam_sharded_in_process_edge_cases(oracle):
    """duplicate landmark ids, empty list, all-zero weights (every slot descends from the last particle of the last rank),
    fresh landmarks — sharded over 2 in-process ranks"""
    n, m = 2048, 4
    lm_xy = np.array([[5.0, 0.0], [0.0, 5.0], [5.0, 5.0], [-5.0, 2.0]])
    ranks = rr.FastSlam1.create_sharded_local(n, m, [0, 0], rr.FsConfig(nth=n / 1.5), seed=11)
    o = OracleFS(oracle, n, m, seed=11, nth=n / 1.5)
    for g in ranks:
        g.seed_map([0.0, 0.0, 0.0], lm_xy)
    o.seed_map([0.0, 0.0, 0.0], lm_xy)
    z = [(5.1, 0.02, 0), (5.0, 1.55, 1), (4.9, -0.01, 0)]
    assert rr.FastSlam1.step_all(ranks, [1.0, 0.1], z) == bool(o.step([1.0, 0.1], z)); _shard_compare(ranks, o, "duplicate ids")
    assert rr.FastSlam1.step_all(ranks, [1.0, 0.1], []) == bool(o.step([1.0, 0.1], [])); _shard_compare(ranks, o, "empty obs")
    for g in ranks:
        p, l = g.state(); p[:, 0] = 0.0; g.set_state(p, l)
    op, ol = o.state(); op[:, 0] = 0.0; o.set_state(op, ol)
    assert rr.FastSlam1.step_all(ranks, [1.0, 0.1], z[:2]) is True and o.step([1.0, 0.1], z[:2]) == 1
    assert all(np.all(g.last_indices() == n - 1) for g in ranks); _shard_compare(ranks, o, "zero weights")
    for g in ranks:
        p, l = g.state(); l[:, 2, :] = [0.0, 0.0, 1000.0, 0.0, 0.0, 1000.0]; g.set_state(p, l)
    op, ol = o.state(); ol[:, 2, :] = [0.0, 0.0, 1000.0, 0.0, 0.0, 1000.0]; o.set_state(op, ol)
    z2 = [(7.0, 0.8, 2), (5.0, 0.1, 0)]
    for _ in range(6):
        assert rr.FastSlam1.step_all(ranks, [1.0, 0.0], z2) == bool(o.step([1.0, 0.0], z2))
    _shard_compare(ranks, o, "mixed fresh")


# ------------------------------------------------------------------------------------------------
# FastSLAM 2.0 (fs2.rs = crates/rust_robotics_slam/src/fastslam2.rs): the same engine with pfgpu_fs_set_variant(2)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,side,steps,seeded", [(64, 4, 30, True), (1000, 6, 40, True), (1000, 5, 40, False), (4096, 8, 20, True), (1 << 16, 16, 4, True)])
def test_fastslam2_trajectory_bit_exact(oracle, n, side, steps, seeded):
    """fastslam2_update (fs2.rs:330-374): sampled poses (proposal of the first observation, Cholesky, three draws), landmark EKFs,
    weights and resample ancestry equal the oracle's bit for bit — from an initialised map and from fresh particles (where the
    first observations give birth to landmarks with cov = 10 I and the proposal falls back to the motion prior)."""
    sc = scenarios.FastSlamScenario(side, (10.0 * side / 2 - 5.0, 10.0 * side / 2 - 5.0, 0.0), (1.0, 0.025), steps)
    g = rr.FastSlam2(n, sc.m, rr.FsConfig(nth=n / 1.5), seed=7)
    o = OracleFS(oracle, n, sc.m, seed=7, variant=2, nth=n / 1.5)
    if seeded:
        g.seed_map(sc.start, sc.landmarks); o.seed_map(sc.start, sc.landmarks)
    else:
        p, l = o.state()
        p[:, 1:] = sc.start
        g.set_state(p, l); o.set_state(p, l)
    resamples = 0
    for t in range(steps):
        z = sc.obs[t]
        if t % 3 == 1 and len(z) > 1:
            z = z[1:] + z[:1]                                       # another landmark leads the list (it feeds the proposal)
        did = g.fastslam2_update(sc.control, z)
        assert did == bool(o.step(sc.control, z)), f"step {t}: gate (neff gpu {g.last_neff()} oracle {o.last_neff()})"
        if did:
            resamples += 1
            assert np.array_equal(g.last_indices(), o.last_indices()), f"step {t}: indices"
        if n <= 4096 or t == steps - 1:
            _fs_compare(g, o, f"step {t}")
        assert g.get_best_particle()[0] == o.best()
    assert resamples > 0, "scenario never resampled"
    assert g.stats().serial_fallbacks == 0


def test_fastslam2_edge_cases(oracle):
    n, m = 256, 4
    lm_xy = np.array([[5.0, 0.0], [0.0, 5.0], [5.0, 5.0], [-5.0, 2.0]])
    g = rr.FastSlam2(n, m, rr.FsConfig(nth=n / 1.5), seed=11)
    o = OracleFS(oracle, n, m, seed=11, variant=2, nth=n / 1.5)
    g.seed_map([0.0, 0.0, 0.0], lm_xy); o.seed_map([0.0, 0.0, 0.0], lm_xy)
    z = [(5.1, 0.02, 0), (5.0, 1.55, 1), (4.9, -0.01, 0)]          # duplicate lm_id: two EKF launches, ONE proposal
    assert g.fastslam2_update([1.0, 0.1], z) == bool(o.step([1.0, 0.1], z)); _fs_compare(g, o, "duplicate ids")
    assert g.fastslam2_update([1.0, 0.1], []) == bool(o.step([1.0, 0.1], [])); _fs_compare(g, o, "empty obs: motion model, two draws (fs2.rs:347-356)")
    p, l = g.state(); p[:, 0] = 0.0; g.set_state(p, l); o.set_state(p, l)
    assert g.fastslam2_update([1.0, 0.1], z[:2]) is True and o.step([1.0, 0.1], z[:2]) == 1
    assert np.all(g.last_indices() == n - 1); _fs_compare(g, o, "zero weights")
    # landmark 2 fresh (cov 1000), landmark 3 exactly on the threshold (cov00 = 100 counts as uninitialised, fs2.rs:50), observed
    # first (motion-prior proposal) and later
    p, l = g.state()
    l[:, 2, :] = [0.0, 0.0, 1000.0, 0.0, 0.0, 1000.0]
    l[:, 3, 2] = 100.0
    g.set_state(p, l); o.set_state(p, l)
    for z2 in ([(7.0, 0.8, 2), (5.0, 0.1, 0)], [(5.0, 0.1, 0), (5.4, 2.7, 3)], [(5.3, 2.6, 3), (7.1, 0.7, 2), (5.0, 0.2, 0)]):
        assert g.fastslam2_update([1.0, 0.0], z2) == bool(o.step([1.0, 0.0], z2))
        _fs_compare(g, o, f"fresh / threshold landmarks {z2[0][2]}")
    gp, gl = g.state()
    assert np.all(gl[:, 2, 2] < 10.0 + 1e-9) and np.all(gl[:, 3, 2] < 10.0 + 1e-9)       # born with cov = 10 I, then shrunk
    # a singular innovation covariance: det S <= 0 multiplies the weight by 1e-10 (fs2.rs:278)
    p, l = g.state()
    l[:, 1, 2:] = [-5.0, 0.0, 0.0, -5.0]
    g.set_state(p, l); o.set_state(p, l)
    assert g.fastslam2_update([1.0, 0.0], [(5.0, 0.1, 0), (5.0, 1.5, 1)]) == bool(o.step([1.0, 0.0], [(5.0, 0.1, 0), (5.0, 1.5, 1)]))
    _fs_compare(g, o, "negative-definite landmark covariance")
    # switching back to FastSLAM 1.0 between steps
    assert g.L.pfgpu_fs_set_variant(g.h, 1) == 0
    o.L.orc_fs_set_variant(o.h, 1)
    assert g.fastslam_update([1.0, 0.0], z[:2]) == bool(o.step([1.0, 0.0], z[:2])); _fs_compare(g, o, "variant 1 after variant 2")
    assert g.L.pfgpu_fs_set_variant(g.h, 3) != 0


@pytest.mark.parametrize("world,n,side,steps", [(2, 1024, 6, 40), (4, 4096, 6, 30)])
def test_fastslam2_sharded_in_process_bit_exact(oracle, world, n, side, steps):
    """the proposal reads the first observation's landmark through the lazy-clone rows, remote ancestors included"""
    sc = scenarios.FastSlamScenario(side, (10.0 * side / 2 - 5.0, 10.0 * side / 2 - 15.0, 0.0), (8.0, 0.8), steps, max_range=12.0)
    ranks = rr.FastSlam2.create_sharded_local(n, sc.m, [0] * world, rr.FsConfig(nth=n / 1.5), seed=9)
    o = OracleFS(oracle, n, sc.m, seed=9, variant=2, nth=n / 1.5)
    for g in ranks:
        g.seed_map(sc.start, sc.landmarks)
    o.seed_map(sc.start, sc.landmarks)
    resamples = 0
    for t in range(steps):
        did = rr.FastSlam2.step_all(ranks, sc.control, sc.obs[t])
        assert did == bool(o.step(sc.control, sc.obs[t])), f"step {t}: gate"
        if did:
            resamples += 1
            assert np.array_equal(np.concatenate([g.last_indices() for g in ranks]), o.last_indices()), f"step {t}: indices"
        if t % 9 == 0:
            _shard_compare(ranks, o, f"step {t}")
    _shard_compare(ranks, o, "end")
    assert resamples > 2


def test_fastslam_get_observations_on_device(oracle):
    """get_observations (fs1.rs:277-299) behind the ABI == the oracle's, bit for bit (same Philox stream), incl. the reference's own
    test geometry (fs1.rs:325-341: one landmark in range, one outside)"""
    g = rr.FastSlam1(64, 4, seed=42)
    o = OracleFS(oracle, 64, 4, seed=42)
    z = g.get_observations([0.0, 0.0, 0.0], [(5.0, 0.0), (100.0, 100.0)], 0)
    assert len(z) == 1 and z[0][2] == 0 and abs(z[0][0] - 5.0) < 5.0
    rng = np.random.default_rng(1)
    for call in range(6):
        lms = rng.uniform(-40, 40, size=(1500, 2))
        xt = [rng.uniform(-5, 5), rng.uniform(-5, 5), rng.uniform(-3, 3)]
        assert g.get_observations(xt, lms, call) == o.observations(xt, lms, 42, call)


def test_cpp_mirror_runs(tmp_path):
    """the C++ host mirror (rust_robotics_b200/host/*.hpp, what a C++ caller of the reference's API would use) driving the
    C ABI: same seeds and inputs as the Python mirror -> identical numbers"""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "mirror_run")
    pkg = os.path.join(root, "rust_robotics_b200")
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-O1", "-DMIRROR_MAIN", os.path.join(pkg, "host", "mirror_check.cpp"), "-I", os.path.join(root, "include"),
                    "-I", os.path.join(pkg, "host"), "-L", pkg, "-lpfgpu", f"-Wl,-rpath,{pkg}", "-o", exe], check=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    got = [float(x) for x in r.stdout.split()]
    pf = rr.ParticleFilterLocalizer.try_with_initial_state([5.0, 5.0, 0.0, 0.0], rr.ParticleFilterConfig(1000, 0.5, 0.25), seed=42)
    z = [(3.1, 2.0, 2.0), (5.0, 10.0, 2.0)]
    e1 = pf.try_step([1.1, 0.0], z)
    pf.predict([0.5, 0.63], 0.1); pf.update(z)                 # the StateEstimator route (pf.rs:552-573): update = update + resample
    e2 = pf.get_state()
    assert np.array_equal(pf.get_covariance(), pf.calc_covariance())
    fs = rr.FastSlam1(256, 4, seed=42)
    did = fs.fastslam_update([1.0, 0.1], [(5.0, 0.1, 0), (7.0, -0.4, 2)])
    bi, bp = fs.get_best_particle()
    lm = fs.particle_landmarks(bi)
    want = [e1[0], e1[1], e2[0], e2[3], 1000.0, bp[0], lm[0, 0] + lm[2, 1], 1.0 if did else 0.0]
    assert got == [float(w) for w in want], (got, want)
