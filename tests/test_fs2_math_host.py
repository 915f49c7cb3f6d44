"""include/fs2_math.h and the variant-2 update of include/fs_ekf_math.h on the CPU, against oracle/fs2_oracle.c: what the CUDA
kernels evaluate for FastSLAM 2.0 must equal the oracle bit for bit (same contract libm on both sides)."""
import ctypes as C
import math
import os
import subprocess

import numpy as np
import pytest

from _oracle import OracleFS

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
dp = C.POINTER(C.c_double)


@pytest.fixture(scope="module")
def probe():
    lib = os.path.join(ROOT, "tests", "host", "libfs2_math_test.so")
    subprocess.run(["/usr/bin/gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", lib,
                    os.path.join(ROOT, "tests", "host", "fs2_math_test.c"), "-lm"], check=True)
    L = C.CDLL(lib)
    L.fs2_probe_propose.argtypes = [dp, dp, dp, dp, dp, C.c_double, C.c_double, C.c_double, dp]
    L.fs2_probe_update.argtypes = [dp, dp, dp, C.c_double, C.c_double]
    L.fs2_probe_update.restype = C.c_double
    return L


def _p(a):
    return a.ctypes.data_as(dp)


def _cases(rng, count):
    for t in range(count):
        pose = np.array([rng.uniform(-30, 30), rng.uniform(-30, 30), rng.uniform(-math.pi, math.pi)])
        u = np.array([rng.uniform(-1, 3), rng.uniform(-1, 1)])
        lm = np.array([pose[0] + rng.uniform(-20, 20), pose[1] + rng.uniform(-20, 20), 0, 0, 0, 0], dtype=np.float64)
        a, b = 10 ** rng.uniform(-3, 1.5, 2)
        c = rng.uniform(-0.9, 0.9) * math.sqrt(a * b)
        lm[2:] = [a, c, c * rng.choice([1.0, 0.999]), b]                     # the reference never symmetrises the covariance
        if t % 11 == 0:
            lm[2:] = [1000.0, 0.0, 0.0, 1000.0]                              # uninitialised
        if t % 13 == 0:
            lm[2] = 100.0                                                    # exactly on the threshold: uninitialised under fs2.rs:50
        if t % 17 == 0:
            lm[:2] = pose[:2] + rng.uniform(-1e-3, 1e-3, 2)                  # landmark almost on top of the particle
        z = np.array([rng.uniform(0.1, 25), rng.uniform(-4, 4)])
        n3 = rng.normal(size=3)
        yield pose, u, z, lm, n3


def test_proposal_and_sample_equal_the_oracle(probe, oracle):
    rng = np.random.default_rng(77)
    o = OracleFS(oracle, 1, 1, variant=2)
    for pose, u, z, lm, n3 in _cases(rng, 3000):
        mean, cov = o.compute_proposal(pose, u, z, lm)
        want = o.sample_pose(mean, cov, n3)
        want[2] = _wrap(want[2])
        got = np.empty(3)
        probe.fs2_probe_propose(_p(pose), _p(u), _p(z), _p(lm), _p(n3), 0.1, 0.5, 0.0305, _p(got))
        assert np.array_equal(got, want) or (np.isnan(got) == np.isnan(want)).all() and np.array_equal(got[~np.isnan(got)], want[~np.isnan(want)]), (pose, u, z, lm, n3, got, want)


def _wrap(a):
    while a > math.pi:
        a -= 2.0 * math.pi
    while a < -math.pi:
        a += 2.0 * math.pi
    return a


def test_variant2_update_equals_the_oracle(probe, oracle):
    """one particle, one observation, no process noise is not available in FastSLAM 2.0 (the pose is always sampled), so the
    oracle runs a whole step with injected zero draws and an uninitialised proposal landmark ... simpler: a second landmark
    carries the case while the FIRST observation (the proposal's) is of a far, uninitialised one; zero draws keep the pose at
    the motion prediction, which the probe is given."""
    rng = np.random.default_rng(78)
    for t, (pose, u, z, lm, n3) in enumerate(_cases(rng, 600)):
        o = OracleFS(oracle, 1, 2, variant=2, nth=0.0)
        pw = np.array([[1.0, pose[0], pose[1], pose[2]]])
        lms = np.array([[[0.0, 0.0, 1000.0, 0.0, 0.0, 1000.0], lm]])
        o.set_state(pw, lms)
        o.step(u, [(5.0, 0.1, 0), (z[0], z[1], 1)], np.zeros(1), np.zeros(2), 0.0)
        p, l = o.state()
        newpose = np.ascontiguousarray(p[0, 1:])
        lm_in = lm.copy()
        f = probe.fs2_probe_update(_p(lm_in), _p(newpose), _p(z), 0.5, 0.0305)
        assert np.array_equal(lm_in, l[0, 1]) or np.isnan(lm_in).any(), (t, lm, lm_in, l[0, 1])
        # weight: 1.0 (fresh landmark 0) * factor, then normalised over one particle -> w / w = 1 unless the factor is 0
        # so compare through a two-particle run instead when the factor matters
        o2 = OracleFS(oracle, 2, 2, variant=2, nth=0.0)
        pw2 = np.array([[0.5, pose[0], pose[1], pose[2]], [0.5, pose[0], pose[1], pose[2]]])
        lms2 = np.array([[[0.0, 0.0, 1000.0, 0.0, 0.0, 1000.0], lm], [[0.0, 0.0, 1000.0, 0.0, 0.0, 1000.0], [0.0, 0.0, 1000.0, 0.0, 0.0, 1000.0]]])
        o2.set_state(pw2, lms2)
        o2.step(u, [(5.0, 0.1, 0), (z[0], z[1], 1)], np.zeros(2), np.zeros(4), 0.0)
        p2, _ = o2.state()
        w0, w1 = 0.5 * f, 0.5 * 1.0
        s = w0 + w1
        if s > 0 and np.isfinite(s):
            assert p2[0, 0] == w0 / s and p2[1, 0] == w1 / s, (t, f, p2[:, 0])
