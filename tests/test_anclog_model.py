"""Design model of the ANCESTRY LOG form of the lazy clone (csrc/fs_kernels.cuh, PFGPU_ANC_LOG=1).

The engine never copies maps at a resample.  Per landmark l it keeps (buf[l], ident[l], gen[l], anc[l][.]): the landmark of
the particle in slot i OF GENERATION gen[l] sits in column (i if ident[l] else anc[l][i]) of buffer buf[l].  A resample only
appends its index array to a ring log (generation G -> G+1); a reader at generation G first walks slot i back to generation
gen[l] through the logged maps.  An EKF update of landmark l writes column i of the other buffer and sets
(ident, gen) = (True, G); it may work in place only if the row is FRESH (ident and gen == G).  Every R resamples all rows are
recomposed to the current generation so that no chain is longer than R.

This test drives that bookkeeping (plain numpy, the same rules the kernels implement) against the obvious eager
implementation (copy every particle's whole map at every resample) on random schedules."""
import numpy as np
import pytest


class EagerMap:
    def __init__(self, n, m, rng):
        self.lm = rng.normal(size=(m, n))

    def update(self, l, delta):
        self.lm[l] += delta

    def resample(self, idx):
        self.lm = self.lm[:, idx]

    def dense(self):
        return self.lm.copy()


class LogMap:
    """the rules of the CUDA path, one scalar per (landmark, particle) standing for the six landmark fields"""

    def __init__(self, n, m, R, init):
        self.n, self.m, self.R = n, m, R
        self.buf = np.zeros((2, m, n))
        self.buf[0] = init
        self.cur = np.zeros(m, dtype=int)        # lmstate bit 0
        self.ident = np.ones(m, dtype=bool)      # lmstate bit 1
        self.gen = np.zeros(m, dtype=int)
        self.anc = np.zeros((2, m, n), dtype=int)
        self.anc_cur = 0
        self.G = 0                               # counters[0]
        self.log = np.zeros((R, n), dtype=int)   # idxlog

    def walk_back(self, l, i, G=None):
        G = self.G if G is None else G
        j = np.asarray(i).copy()
        for g in range(G, self.gen[l], -1):
            j = self.log[(g - 1) % self.R][j]
        return j

    def resolve(self, l, i):
        j = self.walk_back(l, i)
        return j if self.ident[l] else self.anc[self.anc_cur, l, j]

    def update(self, l, delta):                  # fs_ekf_kernel + the lmstate bookkeeping that follows it
        i = np.arange(self.n)
        fresh = self.ident[l] and self.gen[l] == self.G
        col = i if fresh else self.resolve(l, i)
        val = self.buf[self.cur[l], l, col] + delta
        if fresh:
            self.buf[self.cur[l], l, i] = val
        else:
            self.buf[self.cur[l] ^ 1, l, i] = val
            self.cur[l] ^= 1; self.ident[l] = True; self.gen[l] = self.G

    def resample(self, idx):                     # fs_search_pose_kernel + fs_compose_flip_kernel
        self.log[self.G % self.R] = idx
        if (self.G + 1) % self.R == 0:           # refresh: every row recomposed to generation G + 1
            t = np.arange(self.n)
            for l in range(self.m):
                j = self.walk_back(l, t, G=self.G + 1)
                self.anc[self.anc_cur ^ 1, l] = j if self.ident[l] else self.anc[self.anc_cur, l, j]
                self.ident[l] = False; self.gen[l] = self.G + 1
            self.anc_cur ^= 1
        self.G += 1

    def dense(self):                             # fs_unpack_lm_kernel (download)
        i = np.arange(self.n)
        return np.stack([self.buf[self.cur[l], l, self.resolve(l, i)] for l in range(self.m)])


def systematic_like(rng, n):
    """a monotone ancestry with duplicates and gaps, like systematic resampling produces"""
    w = rng.exponential(size=n) ** rng.uniform(0.5, 3.0)
    c = np.cumsum(w / w.sum())
    r = (rng.uniform() + np.arange(n)) / n
    return np.minimum(np.searchsorted(c, r), n - 1)


@pytest.mark.parametrize("R", [1, 2, 3, 8, 32])
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_ancestry_log_equals_eager_clone(R, seed):
    rng = np.random.default_rng(100 * R + seed)
    n, m = 97, 9
    e = EagerMap(n, m, rng)
    g = LogMap(n, m, R, e.dense())
    for step in range(400):
        k = rng.integers(0, 4)
        for l in rng.choice(m, size=k, replace=True):       # replace=True: the same landmark twice in a step (duplicate ids)
            d = rng.normal(size=n)
            e.update(l, d); g.update(l, d)
        if rng.uniform() < 0.6:
            idx = systematic_like(rng, n)
            e.resample(idx); g.resample(idx)
        if step % 37 == 0:
            assert np.array_equal(g.dense(), e.dense()), f"step {step}"
    assert np.array_equal(g.dense(), e.dense())
    assert g.G > 100 and np.all(g.G - g.gen <= R)            # no chain is ever longer than the ring


def test_ancestry_log_long_absence():
    """a landmark that is not observed for many resamples (the case the ring and the refresh exist for)"""
    rng = np.random.default_rng(7)
    n, m, R = 64, 3, 4
    e = EagerMap(n, m, rng)
    g = LogMap(n, m, R, e.dense())
    for step in range(50):
        d = rng.normal(size=n)
        e.update(0, d); g.update(0, d)                       # landmark 0 every step, 1 and 2 never
        idx = systematic_like(rng, n)
        e.resample(idx); g.resample(idx)
    d = rng.normal(size=n)
    e.update(2, d); g.update(2, d)
    assert np.array_equal(g.dense(), e.dense())
