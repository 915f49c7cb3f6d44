"""The N_eff gate of the fused post kernels (fs3_post_kernel / pf3_post_kernel) decides from a tree-order sum of the RAW squared
weights, Q ~ (sum w_raw^2 / S) / S, and hands over to the exact sequential sum of the normalised squares only when (a) N_eff is
within 16 (n + 64) 2^-52 of the threshold or (b) S lies outside [1e-120, 1e120].  This test restates that rule in numpy and checks,
on random and adversarial weight sets at every scale a double can hold, that whenever the shortcut is allowed to decide it decides
exactly what the reference's arithmetic decides (fs1.rs:186-193,262-263; pf.rs:337-345,416-423)."""
import math

import numpy as np
import pytest

EPS = 2.220446049250313e-16


def seq_sum(a):
    s = 0.0
    for x in a.tolist():
        s = s + x
    return s


def tree_sum(a, group=64):
    """64-value groups in pairwise order, then the groups pairwise: the shape of the kernels' tile sums (any tree is admissible)"""
    parts = [float(np.sum(a[i:i + group])) for i in range(0, len(a), group)]
    return float(np.sum(np.array(parts)))


def reference_decision(w_raw, thr, uniform_fallback):
    """normalise, then N_eff = 1 / sum w^2 (sequential), resample iff N_eff < thr"""
    S = seq_sum(w_raw)
    if S > 0.0:
        w = w_raw / S
    elif uniform_fallback:
        w = np.full_like(w_raw, 1.0 / len(w_raw))          # pf.rs:435-437
    else:
        w = w_raw                                           # fs1.rs:196-203: left alone
    Q = seq_sum(w * w)
    neff = 1.0 / Q if Q > 0.0 else 0.0
    return neff < thr, S


def shortcut(w_raw, S, thr, uniform_fallback):
    """(allowed, decision) exactly as the kernels evaluate it"""
    n = len(w_raw)
    qa = tree_sum(w_raw * w_raw)
    if S > 0.0:
        Q = (qa / S) / S
    else:
        Q = 1.0 / n if uniform_fallback else qa
    neff = 1.0 / Q if Q > 0.0 else 0.0
    slack = 16.0 * (n + 64) * EPS
    scale_ok = (not S > 0.0) or (1e-120 <= S <= 1e120)
    allowed = scale_ok and abs(neff - thr) > slack * max(abs(thr), abs(neff))
    return allowed, neff < thr, neff


def weight_sets(rng):
    for n in (64, 257, 1000, 4096):
        for scale_exp in (-300, -200, -160, -121, -119, -60, 0, 60, 119, 121, 150):
            scale = 10.0 ** scale_exp
            yield n, np.full(n, 1.0) * scale                                        # uniform
            yield n, rng.lognormal(0.0, 3.0, n) * scale                             # heavy-tailed
            yield n, rng.lognormal(0.0, 30.0, n) * scale                            # a few particles carry everything
            w = np.full(n, 1e-30 * scale); w[n // 3] = scale; yield n, w            # one-hot over a tiny floor
            yield n, rng.uniform(0.0, 1.0, n) * scale
    yield 512, np.zeros(512)
    yield 512, np.concatenate([np.zeros(511), [5e-324]])                            # one subnormal


@pytest.mark.parametrize("uniform_fallback", [False, True], ids=["fastslam", "pf"])
def test_shortcut_never_disagrees_with_the_reference(uniform_fallback):
    rng = np.random.default_rng(2026)
    decided = handed_over = 0
    with np.errstate(over="ignore", under="ignore", invalid="ignore", divide="ignore"):
        for n, w in weight_sets(rng):
            if not np.all(np.isfinite(w)):
                continue
            _, S = reference_decision(w, 1.0, uniform_fallback)
            _, _, neff_s = shortcut(w, S, 1.0, uniform_fallback)
            # thresholds: the configured ones, and adversarial ones hugging the shortcut's own N_eff from both sides
            thrs = [n / 1.5, 0.5 * n, float(n), 100.0 / 1.5]
            if neff_s > 0.0 and math.isfinite(neff_s):
                thrs += [neff_s * (1.0 + k * EPS * n) for k in (-64.0, -8.0, -1.0, 0.0, 1.0, 8.0, 64.0)]
            for thr in thrs:
                want, S = reference_decision(w, thr, uniform_fallback)
                allowed, got, _ = shortcut(w, S, thr, uniform_fallback)
                if allowed:
                    decided += 1
                    assert got == want, (n, thr, S, float(w.min()), float(w.max()))
                else:
                    handed_over += 1
    assert decided > 500 and handed_over > 50          # both branches exercised


def test_without_the_scale_guard_the_shortcut_is_wrong_for_tiny_weights():
    """the case that motivated the guard: 256 weights of 1e-170 (uniform after normalisation, N_eff = n, no resample)"""
    w = np.full(256, 1e-170)
    want, S = reference_decision(w, 256 / 1.5, False)
    assert want is False or want == False  # noqa: E712
    with np.errstate(under="ignore"):
        qa = tree_sum(w * w)
    assert qa == 0.0                                    # every raw square underflows ...
    allowed, _, _ = shortcut(w, S, 256 / 1.5, False)
    assert not allowed                                  # ... so the rule hands over to the exact sum
