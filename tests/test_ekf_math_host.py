"""include/fs_ekf_math.h on the CPU: the branch-free form of update_landmark (what the EKF kernel runs for almost every
(particle, observation) pair) must equal the contract form bit for bit wherever it declares itself applicable, must leave
the landmark untouched where it does not, and the contract form must equal the oracle's own update_landmark
(oracle/fs1_oracle.c, the restatement of fs1.rs:140-183).  W = 1 (C build) and W = 2 (C++ build, the device's width)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from _oracle import OracleFS

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "host", "ekf_math_test.c")
dp = C.POINTER(C.c_double)


def _build(cxx):
    lib = os.path.join(ROOT, "tests", "host", "libekf_math_test2.so" if cxx else "libekf_math_test.so")
    cmd = (["/usr/bin/g++", "-x", "c++"] if cxx else ["/usr/bin/gcc"]) + ["-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", lib, SRC, "-lm"]
    subprocess.run(cmd, check=True)
    L = C.CDLL(lib)
    L.ekf_math_compare.argtypes = [C.c_uint64, C.c_int, C.c_uint64, C.POINTER(C.c_uint64)]
    L.ekf_math_one.argtypes = [dp, dp, dp]
    return L


@pytest.fixture(scope="module", params=[False, True], ids=["W1", "W2"])
def lib(request):
    return _build(request.param)


# kind 2 = twelve adversarial shapes; almost all of them are outside the fast form's domain by construction (the one with cov00
# exactly 100 left it when the fresh-landmark test became `!(cov00 < 100)`, the union of fs1.rs:144 and fs2.rs:50)
@pytest.mark.parametrize("kind,min_cover", [(0, 0.999), (1, 0.3), (2, 0.001)])
def test_fast_form_is_bit_identical(lib, kind, min_cover):
    o = (C.c_uint64 * 4)()
    lib.ekf_math_compare(20260924 + kind, kind, 400000, o)
    cases, fast, bad, first = list(o)
    assert bad == 0, f"kind {kind}: {bad} mismatches, first at case {first - 1}"
    assert fast / cases >= min_cover, f"kind {kind}: only {fast}/{cases} pairs took the fast form"


def test_contract_form_equals_the_oracle(oracle):
    """fs_update_landmark of the header == update_landmark of oracle/fs1_oracle.c: two particles through one oracle step with
    zero process noise; particle 0 carries the case, particle 1 an uninitialised landmark (its weight is left alone,
    fs1.rs:144-149), so w0'/w1' is the likelihood factor"""
    L = _build(False)
    rng = np.random.default_rng(5)
    for trial in range(200):
        lm = np.array([rng.uniform(-20, 20), rng.uniform(-20, 20), 10 ** rng.uniform(-2, 1), rng.uniform(-0.01, 0.01),
                       rng.uniform(-0.01, 0.01), 10 ** rng.uniform(-2, 1)])
        if trial % 17 == 0:
            lm[2] = lm[5] = 1000.0                             # first observation: branch A
        pose = [rng.uniform(-5, 5), rng.uniform(-5, 5), rng.uniform(-3, 3)]
        z = (rng.uniform(1, 25), rng.uniform(-3, 3))
        o = OracleFS(oracle, 2, 1, seed=1, nth=0.0)            # nth = 0: never resamples
        pw = np.array([[0.5] + pose, [0.5] + pose])
        lms = np.array([[lm], [[0.0, 0.0, 1000.0, 0.0, 0.0, 1000.0]]])
        o.set_state(pw, lms)
        o.step([0.0, 0.0], [(z[0], z[1], 0)], np.zeros(2), np.zeros(2), 0.0)     # zero process noise: the poses stay put
        op, ol = o.state()
        inp = np.array(list(lm) + pose + [z[0], z[1], 0.5, 0.0305], dtype=np.float64)
        ref = np.zeros(7); fast = np.zeros(7)
        L.ekf_math_one(inp.ctypes.data_as(dp), ref.ctypes.data_as(dp), fast.ctypes.data_as(dp))
        assert np.array_equal(ref[:6], ol[0, 0]), f"trial {trial}: landmark {ref[:6]} vs {ol[0, 0]}"
        if op[1, 0] > 0:
            assert op[0, 0] / op[1, 0] == pytest.approx(ref[6], rel=1e-13), f"trial {trial}: likelihood"
