"""The numerical contract (include/pf_contract_math.h): Philox known answers and libm agreement."""
import ctypes as C

import numpy as np

from _oracle import c_dp, c_u32p


def _call1(fn, x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    o = np.empty_like(x)
    fn(x.ctypes.data_as(c_dp), o.ctypes.data_as(c_dp), x.size)
    return o


def _ulps(a, b):
    return np.abs(a - b) / np.spacing(np.abs(b))


def test_philox_known_answers(oracle):
    """Random123 kat_vectors for philox4x32-10."""
    out = (C.c_uint32 * 4)()
    oracle.orc_philox(0, 0, 0, 0, 0, 0, out)
    assert list(out) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    F = 0xffffffff
    oracle.orc_philox(F, F, F, F, F, F, out)
    assert list(out) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    oracle.orc_philox(0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344, 0xa4093822, 0x299f31d0, out)
    assert list(out) == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_contract_math_within_2ulp_of_glibc(oracle):
    rng = np.random.default_rng(7)
    x = np.concatenate([rng.uniform(-745, 709, 200000), rng.uniform(-60, 0, 200000), rng.uniform(-1e-3, 1e-3, 1000)])
    assert _ulps(_call1(oracle.orc_math_exp, x), np.exp(x)).max() <= 2
    x = np.concatenate([rng.uniform(0, 1, 200000), 10 ** rng.uniform(-300, 300, 50000)])
    assert _ulps(_call1(oracle.orc_math_log, x), np.log(x)).max() <= 2
    x = np.concatenate([rng.uniform(-7, 7, 200000), rng.uniform(-1e3, 1e3, 100000), rng.uniform(-1e6, 1e6, 100000)])
    assert _ulps(_call1(oracle.orc_math_sin, x), np.sin(x)).max() <= 2
    assert _ulps(_call1(oracle.orc_math_cos, x), np.cos(x)).max() <= 2
    y, xx = rng.normal(0, 10, 300000), rng.normal(0, 10, 300000)
    o = np.empty_like(y)
    oracle.orc_math_atan2(y.ctypes.data_as(c_dp), xx.ctypes.data_as(c_dp), o.ctypes.data_as(c_dp), y.size)
    assert _ulps(o, np.arctan2(y, xx)).max() <= 2


def test_contract_math_special_values(oracle):
    e = _call1(oracle.orc_math_exp, [0.0, -np.inf, np.inf, -800.0, 800.0, -744.0, 1.0])
    assert e[0] == 1.0 and e[1] == 0.0 and e[2] == np.inf and e[3] == 0.0 and e[4] == np.inf
    assert e[5] > 0 and _ulps(e[5:6], np.exp([-744.0]))[0] <= 1 and _ulps(e[6:7], np.array([np.e]))[0] <= 1
    y = np.array([0.0, 0.0, 1.0, -1.0, 0.0, -0.0, 3.0])
    x = np.array([1.0, -1.0, 0.0, 0.0, 0.0, -1.0, -4.0])
    o = np.empty_like(y)
    oracle.orc_math_atan2(y.ctypes.data_as(c_dp), x.ctypes.data_as(c_dp), o.ctypes.data_as(c_dp), y.size)
    np.testing.assert_array_equal(o, np.arctan2(y, x))
    assert np.isnan(_call1(oracle.orc_math_sin, [np.inf, np.nan])).all()


def test_normal_pair_moments(oracle):
    z = np.empty(2)
    zs = []
    for i in range(20000):
        oracle.orc_normal_pair(42, 0, 3, i, z.ctypes.data_as(c_dp))
        zs.append(z.copy())
    zs = np.array(zs)
    assert abs(zs.mean()) < 0.02 and abs(zs.std() - 1.0) < 0.02
    assert abs(np.corrcoef(zs[:, 0], zs[:, 1])[0, 1]) < 0.03
    u = np.array([oracle.orc_uniform53(42, 1, 0, i) for i in range(20000)])
    assert 0.0 <= u.min() and u.max() < 1.0 and abs(u.mean() - 0.5) < 0.01
