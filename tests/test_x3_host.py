"""Exact sequential-order sum of the fused FastSLAM post kernel (rust_robotics_b200/csrc/x3_core.h): host emulation of the
kernel's pass structure (tests/host/x3_emul.cpp) against a plain left-to-right loop, bit for bit, on adversarial inputs and
for several tile shapes, with and without the "derived" approximate offsets (sum of v partials divided by a scale, values
v_i / scale); and the closed form of the systematic comb for power-of-two particle counts (fs1.rs:219-230)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "host", "x3_emul.cpp")
LIB = os.path.join(ROOT, "tests", "host", "libx3_emul.so")
dp = C.POINTER(C.c_double)


@pytest.fixture(scope="module")
def emul():
    subprocess.run(["/usr/bin/g++", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", LIB, SRC], check=True)
    L = C.CDLL(LIB)
    L.x3_emul_scan.argtypes = [dp, C.c_size_t, C.c_size_t, C.c_size_t, C.c_double, dp, dp, C.POINTER(C.c_longlong)]
    L.x3_seq_scan.argtypes = [dp, C.c_size_t, C.c_double, dp]
    L.x3_emul_comb.argtypes = [C.c_double, C.c_int, C.c_size_t]
    L.x3_emul_comb.restype = C.c_longlong
    return L


def cases():
    rng = np.random.default_rng(3)
    n = 20000
    yield "uniform_pow2", np.full(16384, 1.0 / 16384)
    yield "uniform_non_pow2", np.full(n, 1.0 / n)
    w = rng.uniform(size=n); yield "random_normalised", w / w.sum()
    w = np.exp(rng.normal(0, 12, n)); yield "lognormal_wide", w / w.sum()
    w = np.exp(rng.normal(0, 40, n)); yield "collapse_one_dominant", w / w.sum()
    w = np.full(n, 1e-30); w[100] = 1.0; yield "dominant_then_negligible", w
    w = np.full(n, 2.0 ** -60); w[0] = 1.0 - 2.0 ** -40; yield "crawl_across_edge_1.0", w
    w = np.full(n, 2.0 ** -54); w[0] = 1.0; yield "ties_at_level_0", w          # every add is an exact tie
    w = np.full(n, 3 * 2.0 ** -55); w[0] = 1.0; yield "just_above_tie", w
    w = np.zeros(n); w[n // 2:] = rng.uniform(size=n - n // 2); yield "leading_zeros", w
    yield "all_zero", np.zeros(1000)
    w = np.full(n, 5e-324); yield "all_min_subnormal", w
    w = rng.uniform(size=n) * 1e-310; yield "subnormal_range", w
    w = np.concatenate([rng.uniform(size=500) * 1e-310, rng.uniform(size=500) * 1e-300, rng.uniform(size=500)]); yield "subnormal_to_normal", w
    w = 2.0 ** rng.integers(-60, 1, n).astype(np.float64); yield "powers_of_two", w
    w = np.sort(np.exp(rng.normal(0, 20, n))); yield "sorted_ascending", w
    w = np.sort(np.exp(rng.normal(0, 20, n)))[::-1].copy(); yield "sorted_descending", w
    w = rng.uniform(size=65536) ** 8; yield "weights_65536", w
    w = rng.uniform(size=1 << 18) ** 8; yield "weights_2^18", w


@pytest.mark.parametrize("name,v", list(cases()), ids=[c[0] for c in cases()])
@pytest.mark.parametrize("shape", [(256, 2), (512, 1), (512, 8)], ids=["256x2", "512x1", "512x8"])
def test_scan_is_the_sequential_sum(emul, name, v, shape):
    v = np.ascontiguousarray(v, dtype=np.float64)
    for scale in (0.0, float(v.sum())):
        out, ref = np.empty_like(v), np.empty_like(v)
        tot = C.c_double()
        st = (C.c_longlong * 3)()
        ok = emul.x3_emul_scan(v.ctypes.data_as(dp), v.size, shape[0], shape[1], scale, out.ctypes.data_as(dp), C.byref(tot), st)
        emul.x3_seq_scan(v.ctypes.data_as(dp), v.size, scale, ref.ctypes.data_as(dp))
        assert ok == 1 and st[1] == 0 and st[2] == 0, f"certificate failed: {list(st)}"
        assert np.array_equal(out.view(np.uint64), ref.view(np.uint64)), f"{name} scale {scale}: {int((out != ref).sum())} prefixes differ"
        assert tot.value == ref[-1]


@pytest.mark.parametrize("p", [0, 1, 4, 10, 16, 20])
def test_comb_closed_form(emul, p):
    rng = np.random.default_rng(p)
    for u in [0.0, 2.0 ** -52, 0.5, 1 - 2.0 ** -52] + list(rng.uniform(size=6)):
        u52 = float(np.floor(u * 2 ** 52) / 2 ** 52)
        assert emul.x3_emul_comb(u52, p, 7 if p >= 16 else 1) == 0, f"p={p} u={u52}"
