"""Exact sequential-order scan ("xsum"): host emulation of the device pipeline vs a plain sequential loop.

The element-level logic (rust_robotics_b200/csrc/xsum_core.h) is shared with the CUDA kernels; this test runs
it on the CPU in the same pass structure with deliberately different (pairwise) approximate prefix sums and
checks BIT equality with c_i = fl(c_{i-1} + v_i) on adversarial inputs.  The certificate-violation counter
must stay 0: a violation would mean the rigorous margin argument is wrong."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "host", "xsum_emul.cpp")
LIB = os.path.join(ROOT, "tests", "host", "libxsum_emul.so")
dp = C.POINTER(C.c_double)


@pytest.fixture(scope="module")
def emul():
    subprocess.run(["/usr/bin/g++", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", LIB, SRC], check=True)
    L = C.CDLL(LIB)
    L.xs_emul_scan.argtypes = [dp, C.c_size_t, C.c_size_t, dp, dp, C.POINTER(C.c_longlong)]
    L.xs_seq_scan.argtypes = [dp, C.c_size_t, dp]
    L.xs_emul_scan_scaled.argtypes = [dp, C.c_size_t, C.c_double, C.c_size_t, dp, dp, dp, C.POINTER(C.c_longlong)]
    L.xs_emul_scan_comb.argtypes = [C.c_double, C.c_double, C.c_size_t, C.c_size_t, dp, dp, dp, C.POINTER(C.c_longlong)]
    return L


def run(L, v, tile=256):
    v = np.ascontiguousarray(v, dtype=np.float64)
    out, ref = np.empty_like(v), np.empty_like(v)
    tot = C.c_double()
    st = (C.c_longlong * 3)()
    ok = L.xs_emul_scan(v.ctypes.data_as(dp), v.size, tile, out.ctypes.data_as(dp), C.byref(tot), st)
    L.xs_seq_scan(v.ctypes.data_as(dp), v.size, ref.ctypes.data_as(dp))
    return out, ref, tot.value, ok, list(st)


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
    yield "systematic_comb", np.concatenate([[0.37 / n], np.full(n - 1, 1.0 / n)])
    yield "single", np.array([0.7])
    w = rng.uniform(size=n) * 1e300; yield "huge_values", w / 1e4


@pytest.mark.parametrize("name,v", list(cases()), ids=[c[0] for c in cases()])
@pytest.mark.parametrize("tile", [64, 256, 2048])
def test_xsum_bit_exact(emul, name, v, tile):
    out, ref, tot, ok, st = run(emul, v, tile)
    assert ok == 1 and st[1] == 0, f"certificate violated ({name})"
    assert np.array_equal(out, ref), f"{name}: scan differs at {np.flatnonzero(out != ref)[:5]}"
    assert tot == ref[-1]


def test_xsum_dirty_counts_are_small(emul):
    rng = np.random.default_rng(5)
    n = 1 << 18
    w = rng.uniform(size=n); w /= w.sum()
    out, ref, tot, ok, st = run(emul, w, 2048)
    assert np.array_equal(out, ref)
    assert st[0] < 64, f"dirty elements {st[0]}"
    w = np.exp(rng.normal(0, 30, n)); w /= w.sum()          # weight collapse
    out, ref, tot, ok, st = run(emul, w, 2048)
    assert np.array_equal(out, ref)
    assert st[0] < 200, f"dirty elements {st[0]}"


def test_xsum_random_fuzz(emul):
    rng = np.random.default_rng(11)
    for trial in range(60):
        n = int(rng.integers(1, 5000))
        kind = trial % 4
        if kind == 0:
            w = rng.uniform(size=n)
        elif kind == 1:
            w = np.exp(rng.normal(0, rng.uniform(1, 60), n))
        elif kind == 2:
            w = 2.0 ** rng.integers(-80, 3, n) * rng.integers(1, 4, n)
        else:
            w = np.where(rng.uniform(size=n) < 0.5, 0.0, rng.uniform(size=n) * 10.0 ** rng.integers(-320, 0, n).astype(float))
        out, ref, tot, ok, st = run(emul, w, int(rng.choice([32, 128, 1024])))
        assert ok == 1 and np.array_equal(out, ref), f"trial {trial}"


# ---- derived tile prefixes (fs_post.cuh fx_classify_at): the approximate prefix that steers the classification comes from
# sums the kernel already holds, not from tile sums of the summed values themselves ----
@pytest.mark.parametrize("name,v", list(cases()), ids=[c[0] for c in cases()])
@pytest.mark.parametrize("tile", [64, 512])
def test_xsum_scaled_prefixes_bit_exact(emul, name, v, tile):
    """cumsum(w_i / S) with tile prefixes = (tile sums of w) / S, for S = the sequential sum of w (the resample's
    re-normalisation, fs1.rs:207) and for an S a few ulps off"""
    w = np.ascontiguousarray(v, dtype=np.float64)
    seq = np.empty_like(w)
    emul.xs_seq_scan(w.ctypes.data_as(dp), w.size, seq.ctypes.data_as(dp))
    for S in (float(seq[-1]), float(np.nextafter(seq[-1], np.inf)), 0.0):
        vv, out, ref = np.empty_like(w), np.empty_like(w), np.empty_like(w)
        tot = C.c_double(); st = (C.c_longlong * 3)()
        ok = emul.xs_emul_scan_scaled(w.ctypes.data_as(dp), w.size, S, tile, vv.ctypes.data_as(dp), out.ctypes.data_as(dp), C.byref(tot), st)
        if not np.all(np.isfinite(vv)):
            continue                                   # w / S overflowed: the kernels take the exact serial walk for such input
        emul.xs_seq_scan(vv.ctypes.data_as(dp), vv.size, ref.ctypes.data_as(dp))
        assert ok == 1 and st[1] == 0, f"certificate violated ({name}, S={S})"
        assert np.array_equal(out, ref), f"{name}, S={S}: scan differs at {np.flatnonzero(out != ref)[:5]}"


@pytest.mark.parametrize("n", [1, 7, 1000, 4096, 65536, 100003, 1 << 20])
@pytest.mark.parametrize("tile", [512])
def test_xsum_comb_closed_form_prefixes_bit_exact(emul, n, tile):
    """r_t = r0 + 1/n + 1/n + ... accumulated left to right (fs1.rs:219-230) with closed-form tile prefixes"""
    rng = np.random.default_rng(n)
    inv = 1.0 / n
    for r0 in (0.0, inv * rng.uniform(), float(np.nextafter(inv, 0.0)), inv * 2.0 ** -30):
        vv, out, ref = np.empty(n), np.empty(n), np.empty(n)
        tot = C.c_double(); st = (C.c_longlong * 3)()
        ok = emul.xs_emul_scan_comb(r0, inv, n, tile, vv.ctypes.data_as(dp), out.ctypes.data_as(dp), C.byref(tot), st)
        emul.xs_seq_scan(vv.ctypes.data_as(dp), n, ref.ctypes.data_as(dp))
        assert ok == 1 and st[1] == 0 and np.array_equal(out, ref), f"n={n} r0={r0}"

