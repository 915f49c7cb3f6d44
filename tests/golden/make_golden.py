#!/usr/bin/env python3
"""Golden-vector generator: an INDEPENDENT pure-Python restatement of the reference hot path.

Run (in the authoring container):  python tests/golden/make_golden.py
Writes tests/golden/pf_golden.json, mcl_golden.json, fs1_golden.json, kat_golden.json, fs2_golden.json.

Why Python: the reference (Rust) cannot be built here (no rustc/cargo, deps not vendored), so there is no
reference-generated vector ("parity unpinned", see oracle/oracle.h).  This script is a second, separately
written restatement of the same Rust source lines; Python floats are IEEE f64, CPython never fuses a*b+c,
and math.sin/cos/exp/atan2/sqrt call glibc — the same libm Rust's f64 methods call on Linux.  The C oracle
built with -DPF_ORACLE_LIBM must reproduce these files BIT FOR BIT (tests/test_oracle_golden.py); the
contract-math oracle and the CUDA path must match to a few ulp.

Random draws (N(0,1) noise, uniforms) are part of the fixture: they are generated here with numpy's
PCG64 and injected into the oracle through its *_with_noise entry points.

Reference citations: pf.rs = crates/rust_robotics_localization/src/particle_filter.rs,
mcl.rs = .../monte_carlo_localization.rs, fs1.rs = crates/rust_robotics_slam/src/fastslam1.rs,
fs2.rs = crates/rust_robotics_slam/src/fastslam2.rs.
"""
import json
import math
import os

import numpy as np

PI = math.pi
HERE = os.path.dirname(os.path.abspath(__file__))


def hx(v):
    """exact float serialisation"""
    if isinstance(v, (list, tuple)):
        return [hx(a) for a in v]
    return float(v).hex()


# ------------------------------------------------------------------------------------------------
# pf.rs / mcl.rs
# ------------------------------------------------------------------------------------------------
class P:  # pf.rs:26-32
    __slots__ = ("x", "y", "yaw", "v", "w")

    def __init__(self, x, y, yaw, v, w):
        self.x, self.y, self.yaw, self.v, self.w = x, y, yaw, v, w

    def clone(self):
        return P(self.x, self.y, self.yaw, self.v, self.w)

    def row(self):
        return [self.x, self.y, self.yaw, self.v, self.w]


def pf_predict(ps, u, zv, zw, sv, sw, dt):  # pf.rs:279-296
    for i, p in enumerate(ps):
        v_noise = (0.0 + sv * zv[i]) if sv > 0.0 else 0.0
        yaw_noise = (0.0 + sw * zw[i]) if sw > 0.0 else 0.0
        v_noisy = u[0] + v_noise
        yaw_rate_noisy = u[1] + yaw_noise
        c, s = math.cos(p.yaw), math.sin(p.yaw)
        p.x += v_noisy * c * dt
        p.y += v_noisy * s * dt
        p.yaw += yaw_rate_noisy * dt
        p.v = v_noisy


def gauss_likelihood(x, sigma):  # pf.rs:476-479
    coeff = 1.0 / math.sqrt(2.0 * PI * (sigma * sigma))
    return coeff * math.exp(-(x * x) / (2.0 * (sigma * sigma)))


def pf_normalize(ps):  # pf.rs:426-439
    s = 0.0
    for p in ps:
        s += p.w
    if s > 0.0:
        for p in ps:
            p.w /= s
    else:
        uw = 1.0 / float(len(ps))
        for p in ps:
            p.w = uw


def pf_update(ps, obs, sigma):  # pf.rs:316-331
    for p in ps:
        w = 1.0
        for (d_obs, lx, ly) in obs:
            dx = p.x - lx
            dy = p.y - ly
            d_pred = math.sqrt(dx * dx + dy * dy)
            diff = d_obs - d_pred
            w *= gauss_likelihood(diff, sigma)
        p.w = w
    pf_normalize(ps)


def pf_neff(ps):  # pf.rs:416-423
    s2 = 0.0
    for p in ps:
        s2 += p.w * p.w
    return 1.0 / s2 if s2 > 0.0 else 0.0


def pf_estimate(ps):  # pf.rs:382-396
    xe = ye = yawe = ve = 0.0
    for p in ps:
        xe += p.w * p.x
        ye += p.w * p.y
        yawe += p.w * p.yaw
        ve += p.w * p.v
    return [xe, ye, yawe, ve]


def pf_cov(ps, est):  # pf.rs:398-413 -> returned column-major like nalgebra storage
    cov = [[0.0] * 4 for _ in range(4)]
    for p in ps:
        dx = [p.x - est[0], p.y - est[1], p.yaw - est[2], p.v - est[3]]
        wdx = [p.w * d for d in dx]
        for a in range(4):
            for b in range(4):
                cov[a][b] += wdx[a] * dx[b]
    return [cov[i][j] for j in range(4) for i in range(4)]


def pf_resample_particles(ps, n, rs):  # pf.rs:442-473
    cum = []
    c = 0.0
    for p in ps:
        c += p.w
        cum.append(c)
    new, idxs = [], []
    for t in range(n):
        r = rs[t]
        index = 0
        for i, cw in enumerate(cum):
            if r <= cw:
                index = i
                break
        q = ps[index].clone()
        q.w = 1.0 / float(n)
        new.append(q)
        idxs.append(index)
    return new, idxs


def sat_i32(v):
    if v != v:
        return 0
    return int(max(-2147483648.0, min(2147483647.0, v)))


def kld_required(k_bins, nmin, nmax, eps, z):  # mcl.rs:367-378
    if k_bins <= 1:
        return nmin
    km1 = float(k_bins - 1)
    term = 1.0 - 2.0 / (9.0 * km1) + z * math.sqrt(2.0 / (9.0 * km1))
    n = (km1 / (2.0 * eps)) * (term * term * term)
    v = int(math.ceil(n)) if n > 0 else 0
    return max(nmin, min(nmax, v))


def mcl_resample_adaptive(ps, nmin, nmax, eps, z, rs):  # mcl.rs:322-365
    cum = []
    c = 0.0
    for p in ps:
        c += p.w
        cum.append(c)
    cum[-1] = 1.0
    bins = set()
    new, idxs = [], []
    required = nmin
    while len(new) < nmax:
        r = rs[len(new)]
        idx = len(cum) - 1
        for i, cw in enumerate(cum):
            if r <= cw:
                idx = i
                break
        s = ps[idx]
        bins.add((sat_i32(math.floor(s.x / 0.5)), sat_i32(math.floor(s.y / 0.5)),
                  sat_i32(math.floor(s.yaw / (15.0 * PI / 180.0)))))
        required = max(required, kld_required(len(bins), nmin, nmax, eps, z))
        new.append(s.clone())
        idxs.append(idx)
        if len(new) >= nmin and len(new) >= required:
            break
    uw = 1.0 / float(len(new))
    for p in new:
        p.w = uw
    return new, idxs


def run_pf_case(name, rng, n, T, K, thr, sv, sw, sigma, dt, zero_weight_step=None):
    ps = [P(5.0 + rng.uniform(-1, 1), 5.0 + rng.uniform(-1, 1), rng.uniform(-0.25, 0.25),
            rng.uniform(-0.5, 0.5), 1.0 / n) for _ in range(n)]
    case = {"name": name, "n": n, "threshold": hx(thr), "sv": hx(sv), "sw": hx(sw), "sigma": hx(sigma),
            "dt": hx(dt), "init": [hx(p.row()) for p in ps], "steps": []}
    lms = [(2.0, 2.0), (10.0, 2.0), (2.0, 8.0), (10.0, 8.0), (6.0, 5.0)][:K]
    truth = [5.0, 5.0, 0.0]
    for t in range(T):
        u = [1.1, 0.0] if (t // 3) % 2 == 0 else [0.5, 0.63]
        truth[0] += u[0] * math.cos(truth[2]) * dt
        truth[1] += u[0] * math.sin(truth[2]) * dt
        truth[2] += u[1] * dt
        obs = [[max(math.hypot(truth[0] - lx, truth[1] - ly) + rng.normal(0, 0.15), 0.0), lx, ly] for lx, ly in lms]
        if zero_weight_step == t:  # all likelihoods underflow -> sum_w == 0 -> uniform fallback (pf.rs:433-438)
            obs = [[1.0e4, lx, ly] for lx, ly in lms]
        zv = rng.normal(size=n).tolist()
        zw = rng.normal(size=n).tolist()
        rs = rng.uniform(size=n).tolist()
        if t == 1:
            rs[0] = 0.0          # edge: r == 0 -> index 0
            rs[1] = 1.0 - 2.0 ** -53  # edge: r above the last cumsum when it rounds below 1 -> fallback 0
        pf_predict(ps, u, zv, zw, sv, sw, dt)
        est_p = pf_estimate(ps)
        pf_update(ps, obs, sigma)
        est_u = pf_estimate(ps)
        w_after_update = [p.w for p in ps]
        neff = pf_neff(ps)
        did = neff < float(n) * thr
        idxs = []
        if did:
            ps, idxs = pf_resample_particles(ps, n, rs)
        est = pf_estimate(ps)
        cov = pf_cov(ps, est)
        case["steps"].append({"u": hx(u), "obs": [hx(o) for o in obs], "zv": hx(zv), "zw": hx(zw), "r": hx(rs),
                              "est_after_predict": hx(est_p), "est_after_update": hx(est_u),
                              "w_after_update": hx(w_after_update), "neff": hx(neff),
                              "did_resample": bool(did), "indices": idxs,
                              "particles": [hx(p.row()) for p in ps], "est": hx(est), "cov": hx(cov)})
    return case


def run_mcl_case(name, rng, nmin, nmax, T, K, sv, sw, sigma, dt, eps=0.05, z=2.326, spread=1.0):
    n = nmin
    ps = [P(spread * rng.uniform(-1, 1), spread * rng.uniform(-1, 1), rng.uniform(-0.25, 0.25),
            1.0 + rng.uniform(-0.5, 0.5), 1.0 / n) for _ in range(n)]
    case = {"name": name, "nmin": nmin, "nmax": nmax, "eps": hx(eps), "z": hx(z), "sv": hx(sv), "sw": hx(sw),
            "sigma": hx(sigma), "dt": hx(dt), "init": [hx(p.row()) for p in ps], "steps": []}
    lms = [(10.0, 0.0), (0.0, 10.0), (-10.0, 0.0), (0.0, -10.0)][:K]
    truth = [0.0, 0.0, 0.0]
    for t in range(T):
        u = [1.0, 0.03]
        truth[0] += u[0] * math.cos(truth[2]) * dt
        truth[1] += u[0] * math.sin(truth[2]) * dt
        truth[2] += u[1] * dt
        obs = [[max(math.hypot(truth[0] - lx, truth[1] - ly) + rng.normal(0, 0.1), 0.0), lx, ly] for lx, ly in lms]
        cur = len(ps)
        zv = rng.normal(size=cur).tolist()
        zw = rng.normal(size=cur).tolist()
        rs = rng.uniform(size=nmax).tolist()
        if t == 1:
            rs[0] = 1.0 - 2.0 ** -53   # edge: above every cumsum but the forced last 1.0 -> index len-1
        pf_predict(ps, u, zv, zw, sv, sw, dt)
        pf_update(ps, obs, sigma)
        w_after_update = [p.w for p in ps]
        ps, idxs = mcl_resample_adaptive(ps, nmin, nmax, eps, z, rs)
        est = pf_estimate(ps)
        cov = pf_cov(ps, est)
        case["steps"].append({"u": hx(u), "obs": [hx(o) for o in obs], "zv": hx(zv), "zw": hx(zw), "r": hx(rs),
                              "w_after_update": hx(w_after_update), "count": len(ps), "indices": idxs,
                              "particles": [hx(p.row()) for p in ps], "est": hx(est), "cov": hx(cov)})
    return case


# ------------------------------------------------------------------------------------------------
# fs1.rs
# ------------------------------------------------------------------------------------------------
def normalize_angle(a):  # fs1.rs:80-89
    while a > PI:
        a -= 2.0 * PI
    while a < -PI:
        a += 2.0 * PI
    return a


class FP:  # fs1.rs:44-51; landmark = [x, y, c00, c01, c10, c11]
    def __init__(self, w, x, y, yaw, lms):
        self.w, self.x, self.y, self.yaw, self.lms = w, x, y, yaw, lms

    def clone(self):
        return FP(self.w, self.x, self.y, self.yaw, [list(l) for l in self.lms])


def fs_predict(p, u, z0, z1, cfg):  # fs1.rs:123-137, 70-77
    u0 = u[0] + z0 * math.sqrt(cfg["q00"])
    u1 = u[1] + z1 * math.sqrt(cfg["q11"])
    yaw = p.yaw
    nx = p.x + u0 * cfg["dt"] * math.cos(yaw)
    ny = p.y + u0 * cfg["dt"] * math.sin(yaw)
    nyaw = normalize_angle(yaw + u1 * cfg["dt"])
    p.x, p.y, p.yaw = nx, ny, nyaw


def mm(a, b):  # nalgebra 2x2 product: entry = a_i0*b_0j + a_i1*b_1j
    return [[a[i][0] * b[0][j] + a[i][1] * b[1][j] for j in range(2)] for i in range(2)]


def tr(a):
    return [[a[0][0], a[1][0]], [a[0][1], a[1][1]]]


def fs_update_landmark(p, z0, z1, lm_id, cfg):  # fs1.rs:140-183
    L = p.lms[lm_id]
    if L[2] > 100.0:
        L[0] = p.x + z0 * math.cos(p.yaw + z1)
        L[1] = p.y + z0 * math.sin(p.yaw + z1)
        return
    dx, dy = L[0] - p.x, L[1] - p.y
    d = math.sqrt(dx * dx + dy * dy)
    zp1 = normalize_angle(math.atan2(dy, dx) - p.yaw)
    y = [z0 - d, normalize_angle(z1 - zp1)]
    d2 = dx * dx + dy * dy
    dd = math.sqrt(d2)
    H = [[dx / dd, dy / dd], [-dy / d2, dx / d2]]
    Pm = [[L[2], L[3]], [L[4], L[5]]]
    R = [[cfg["r00"], 0.0], [0.0, cfg["r11"]]]
    HPHt = mm(mm(H, Pm), tr(H))
    S = [[HPHt[i][j] + R[i][j] for j in range(2)] for i in range(2)]
    det = S[0][0] * S[1][1] - S[1][0] * S[0][1]
    if det == 0.0:
        Si = [[1.0, 0.0], [0.0, 1.0]]
    else:
        Si = [[S[1][1] / det, -S[0][1] / det], [-S[1][0] / det, S[0][0] / det]]
    K = mm(mm(Pm, tr(H)), Si)
    L[0] += K[0][0] * y[0] + K[0][1] * y[1]
    L[1] += K[1][0] * y[0] + K[1][1] * y[1]
    KH = mm(K, H)
    IKH = [[1.0 - KH[0][0], 0.0 - KH[0][1]], [0.0 - KH[1][0], 1.0 - KH[1][1]]]
    Pn = mm(IKH, Pm)
    L[2], L[3], L[4], L[5] = Pn[0][0], Pn[0][1], Pn[1][0], Pn[1][1]
    det_s = S[0][0] * S[1][1] - S[1][0] * S[0][1]
    if det_s > 0.0:
        t = [y[0] * Si[0][0] + y[1] * Si[1][0], y[0] * Si[0][1] + y[1] * Si[1][1]]
        mahal = t[0] * y[0] + t[1] * y[1]
        likelihood = math.exp(-0.5 * mahal) / (2.0 * PI * math.sqrt(det_s))
        p.w *= likelihood


def fs_normalize(ps):  # fs1.rs:196-203
    s = 0.0
    for p in ps:
        s += p.w
    if s > 0.0:
        for p in ps:
            p.w /= s


def fs_neff(ps):  # fs1.rs:186-193
    s2 = 0.0
    for p in ps:
        s2 += p.w * p.w
    return 1.0 / s2 if s2 > 0.0 else 0.0


def fs_resample(ps, u01):  # fs1.rs:206-234
    fs_normalize(ps)
    n = len(ps)
    cum = [0.0] * (n + 1)
    for i, p in enumerate(ps):
        cum[i + 1] = cum[i] + p.w
    r = u01 * (1.0 / float(n) - 0.0) + 0.0
    new, idxs = [], []
    j = 0
    for _ in range(n):
        while r > cum[j + 1] and j < n - 1:
            j += 1
        q = ps[j].clone()
        q.w = 1.0 / float(n)
        new.append(q)
        idxs.append(j)
        r += 1.0 / float(n)
    return new, idxs


def fs_state(ps):
    pose = [[p.w, p.x, p.y, p.yaw] for p in ps]
    lm = [[list(l) for l in p.lms] for p in ps]
    return pose, lm


def run_fs_case(name, rng, n, m, T, nth, init_cov, zero_weights_at=None):
    cfg = {"dt": 0.1, "max_range": 20.0, "nth": nth, "q00": 0.3, "q11": 0.0305, "r00": 0.5, "r11": 0.0305,
           "init_weight": 0.01}
    lm_true = [(10.0, -2.0), (15.0, 10.0), (3.0, 15.0), (-5.0, 20.0), (-5.0, 5.0), (25.0, 25.0)][:m]
    ps = []
    for _ in range(n):
        lms = []
        for l, (lx, ly) in enumerate(lm_true):
            if init_cov[l] > 100.0:
                lms.append([0.0, 0.0, 1000.0, 0.0, 0.0, 1000.0])       # fresh Landmark::new fs1.rs:34-40
            else:
                lms.append([lx + rng.normal(), ly + rng.normal(), init_cov[l], 0.0, 0.0, init_cov[l]])
        ps.append(FP(cfg["init_weight"], 0.0, 0.0, 0.0, lms))
    pose0, lm0 = fs_state(ps)
    case = {"name": name, "n": n, "m": m, "cfg": {k: hx(v) for k, v in cfg.items()},
            "init_pose": [hx(r) for r in pose0], "init_lm": [[hx(l) for l in row] for row in lm0], "steps": []}
    xt = [0.0, 0.0, 0.0]
    for t in range(T):
        u = [1.0, 0.1]
        xt = [xt[0] + u[0] * cfg["dt"] * math.cos(xt[2]), xt[1] + u[0] * cfg["dt"] * math.sin(xt[2]),
              normalize_angle(xt[2] + u[1] * cfg["dt"])]
        obs = []
        for l, (lx, ly) in enumerate(lm_true):                        # get_observations fs1.rs:277-299
            dx, dy = lx - xt[0], ly - xt[1]
            d = math.sqrt(dx * dx + dy * dy)
            if d <= cfg["max_range"]:
                ang = normalize_angle(math.atan2(dy, dx) - xt[2])
                obs.append([d + rng.normal() * math.sqrt(cfg["r00"]), ang + rng.normal() * math.sqrt(cfg["r11"]), l])
        if t == 2 and len(obs) > 1:
            obs.append(list(obs[0]))                                  # duplicate lm_id: sequential EKF updates
        z0 = rng.normal(size=n).tolist()
        z1 = rng.normal(size=n).tolist()
        u01 = float(rng.uniform())
        if zero_weights_at == t:                                      # caller zeroes the weights before the step
            for p in ps:
                p.w = 0.0
        for i, p in enumerate(ps):
            fs_predict(p, u, z0[i], z1[i], cfg)
        for (d, a, l) in obs:                                         # obs outer, particle inner fs1.rs:250-256
            for p in ps:
                fs_update_landmark(p, d, a, l, cfg)
        fs_normalize(ps)
        neff = fs_neff(ps)
        did = neff < cfg["nth"]
        idxs = []
        if did:
            ps, idxs = fs_resample(ps, u01)
        pose, lm = fs_state(ps)
        best = 0
        for i in range(1, n):                                         # max_by -> last max fs1.rs:269-274
            if ps[i].w >= ps[best].w:
                best = i
        case["steps"].append({"u": hx(u), "obs": [[hx(o[0]), hx(o[1]), int(o[2])] for o in obs], "z0": hx(z0),
                              "z1": hx(z1), "u01": hx(u01), "neff": hx(neff), "did_resample": bool(did),
                              "zero_weights": zero_weights_at == t, "indices": idxs, "best": best,
                              "pose": [hx(r) for r in pose], "lm": [[hx(l) for l in row] for row in lm]})
    return case


# ------------------------------------------------------------------------------------------------
# fs2.rs = crates/rust_robotics_slam/src/fastslam2.rs (FastSLAM 2.0).  Matrices are lists of rows; products follow nalgebra's
# static-size path (column by column, each entry accumulated left to right: ((a_i0*b_0j) + a_i1*b_1j) + a_i2*b_2j).
# ------------------------------------------------------------------------------------------------
FS2_MOTION_COV = [[0.1, 0.0, 0.0], [0.0, 0.1, 0.0], [0.0, 0.0, 0.01]]          # fs2.rs:31


def gmm(a, b):
    rows, inner, cols = len(a), len(b), len(b[0])
    out = [[0.0] * cols for _ in range(rows)]
    for j in range(cols):
        for i in range(rows):
            acc = a[i][0] * b[0][j]
            for k in range(1, inner):
                acc = a[i][k] * b[k][j] + acc
            out[i][j] = acc
    return out


def gtr(a):
    return [[a[i][j] for i in range(len(a))] for j in range(len(a[0]))]


def gadd(a, b):
    return [[a[i][j] + b[i][j] for j in range(len(a[0]))] for i in range(len(a))]


def inv2(m):  # nalgebra try_inverse, 2x2
    det = m[0][0] * m[1][1] - m[1][0] * m[0][1]
    if det == 0.0:
        return None
    return [[m[1][1] / det, -m[0][1] / det], [-m[1][0] / det, m[0][0] / det]]


def inv3(m):  # nalgebra try_inverse, 3x3
    (m11, m12, m13), (m21, m22, m23), (m31, m32, m33) = m
    minor_m12_m23 = m22 * m33 - m32 * m23
    minor_m11_m23 = m21 * m33 - m31 * m23
    minor_m11_m22 = m21 * m32 - m31 * m22
    det = m11 * minor_m12_m23 - m12 * minor_m11_m23 + m13 * minor_m11_m22
    if det == 0.0:
        return None
    return [[minor_m12_m23 / det, (m13 * m32 - m33 * m12) / det, (m12 * m23 - m22 * m13) / det],
            [-minor_m11_m23 / det, (m11 * m33 - m31 * m13) / det, (m13 * m21 - m23 * m11) / det],
            [minor_m11_m22 / det, (m12 * m31 - m32 * m11) / det, (m11 * m22 - m21 * m12) / det]]


def chol3_l(m):  # nalgebra Cholesky::new(...).l(): None when a pivot is zero, negative or NaN
    w = [list(r) for r in m]
    for j in range(3):
        for k in range(j):
            factor = -w[j][k]
            for i in range(j, 3):
                w[i][j] = factor * w[i][k] + w[i][j]
        diag = w[j][j]
        if diag == 0.0 or not diag >= 0.0:
            return None
        denom = math.sqrt(diag)
        w[j][j] = denom
        for i in range(j + 1, 3):
            w[i][j] = w[i][j] / denom
    return [[w[i][j] if j <= i else 0.0 for j in range(3)] for i in range(3)]


def fs2_motion_model(x, u, dt):  # fs2.rs:95-102
    yaw = x[2]
    return [x[0] + u[0] * dt * math.cos(yaw), x[1] + u[0] * dt * math.sin(yaw), normalize_angle(x[2] + u[1] * dt)]


def fs2_compute_proposal(p, u, z, lm_id, cfg):  # fs2.rs:173-216
    lm = p.lms[lm_id]
    pose = [p.x, p.y, p.yaw]
    x_pred = fs2_motion_model(pose, u, cfg["dt"])
    yaw, v = pose[2], u[0]
    g = [[1.0, 0.0, -v * cfg["dt"] * math.sin(yaw)], [0.0, 1.0, v * cfg["dt"] * math.cos(yaw)], [0.0, 0.0, 1.0]]
    p_pred = gmm(gmm(g, FS2_MOTION_COV), gtr(g))
    if not lm[2] < 100.0:
        return x_pred, p_pred
    dx, dy = lm[0] - x_pred[0], lm[1] - x_pred[1]
    d2 = dx * dx + dy * dy
    d = math.sqrt(d2)
    h_pose = [[-dx / d, -dy / d, 0.0], [dy / d2, -dx / d2, -1.0]]
    h_lm = [[dx / d, dy / d], [-dy / d2, dx / d2]]
    cov = [[lm[2], lm[3]], [lm[4], lm[5]]]
    r = [[cfg["r00"], 0.0], [0.0, cfg["r11"]]]
    q_obs = gadd(gmm(gmm(h_lm, cov), gtr(h_lm)), r)
    h_pose_t = gtr(h_pose)
    q_obs_inv = inv2(q_obs) or [[1.0, 0.0], [0.0, 1.0]]
    p_pred_inv = inv3(p_pred) or [[1.0 * 1e-6, 0.0 * 1e-6, 0.0 * 1e-6], [0.0 * 1e-6, 1.0 * 1e-6, 0.0 * 1e-6], [0.0 * 1e-6, 0.0 * 1e-6, 1.0 * 1e-6]]
    p_post_inv = gadd(p_pred_inv, gmm(gmm(h_pose_t, q_obs_inv), h_pose))
    p_post = inv3(p_post_inv) or p_pred
    ddx, ddy = lm[0] - x_pred[0], lm[1] - x_pred[1]                      # observation_model fs2.rs:122-128
    z_pred = [math.sqrt(ddx * ddx + ddy * ddy), normalize_angle(math.atan2(ddy, ddx) - x_pred[2])]
    innovation = [[z[0] - z_pred[0]], [normalize_angle(z[1] - z_pred[1])]]
    corr = gmm(gmm(gmm(p_post, h_pose_t), q_obs_inv), innovation)
    return [x_pred[i] + corr[i][0] for i in range(3)], p_post


def fs2_sample_pose(mean, cov, n3):  # fs2.rs:219-239
    l = chol3_l(cov)
    if l is None:
        l = [[math.sqrt(max(cov[i][i], 0.0)) if i == j else 0.0 for j in range(3)] for i in range(3)]
    ln = gmm(l, [[n3[0]], [n3[1]], [n3[2]]])
    return [mean[i] + ln[i][0] for i in range(3)]


def fs2_update_landmark_and_weight(p, z, lm_id, cfg):  # fs2.rs:242-280
    L = p.lms[lm_id]
    if not L[2] < 100.0:
        L[0] = p.x + z[0] * math.cos(p.yaw + z[1])
        L[1] = p.y + z[0] * math.sin(p.yaw + z[1])
        L[2], L[3], L[4], L[5] = 1.0 * 10.0, 0.0 * 10.0, 0.0 * 10.0, 1.0 * 10.0
        return 1.0
    dx, dy = L[0] - p.x, L[1] - p.y
    z_pred = [math.sqrt(dx * dx + dy * dy), normalize_angle(math.atan2(dy, dx) - p.yaw)]
    innovation = [[z[0] - z_pred[0]], [normalize_angle(z[1] - z_pred[1])]]
    d2 = dx * dx + dy * dy
    d = math.sqrt(d2)
    h = [[dx / d, dy / d], [-dy / d2, dx / d2]]
    cov = [[L[2], L[3]], [L[4], L[5]]]
    r = [[cfg["r00"], 0.0], [0.0, cfg["r11"]]]
    s = gadd(gmm(gmm(h, cov), gtr(h)), r)
    s_inv = inv2(s) or [[1.0, 0.0], [0.0, 1.0]]
    k = gmm(gmm(cov, gtr(h)), s_inv)
    delta = gmm(k, innovation)
    L[0] += delta[0][0]
    L[1] += delta[1][0]
    kh = gmm(k, h)
    ikh = [[1.0 - kh[0][0], 0.0 - kh[0][1]], [0.0 - kh[1][0], 1.0 - kh[1][1]]]
    pn = gmm(ikh, cov)
    L[2], L[3], L[4], L[5] = pn[0][0], pn[0][1], pn[1][0], pn[1][1]
    det_s = s[0][0] * s[1][1] - s[1][0] * s[0][1]
    if det_s > 0.0:
        mahal = gmm(gmm(gtr(innovation), s_inv), innovation)
        return math.exp(-0.5 * mahal[0][0]) / (2.0 * PI * math.sqrt(det_s))
    return 1e-10


def run_fs2_case(name, rng, n, m, T, nth, init_cov, no_obs_at=None, zero_weights_at=None):
    cfg = {"dt": 0.1, "max_range": 20.0, "nth": nth, "q00": 0.3, "q11": 0.0305, "r00": 0.5, "r11": 0.0305,
           "init_weight": 0.01}
    lm_true = [(10.0, -2.0), (15.0, 10.0), (3.0, 15.0), (-5.0, 20.0), (-5.0, 5.0), (25.0, 25.0)][:m]
    ps = []
    for _ in range(n):
        lms = []
        for l, (lx, ly) in enumerate(lm_true):
            if init_cov[l] >= 100.0:
                lms.append([0.0, 0.0, 1000.0, 0.0, 0.0, 1000.0])       # Landmark::new fs2.rs:41-47
            else:
                lms.append([lx + rng.normal(), ly + rng.normal(), init_cov[l], 0.0, 0.0, init_cov[l]])
        ps.append(FP(cfg["init_weight"], 0.0, 0.0, 0.0, lms))
    pose0, lm0 = fs_state(ps)
    case = {"name": name, "n": n, "m": m, "cfg": {k: hx(v) for k, v in cfg.items()},
            "init_pose": [hx(r) for r in pose0], "init_lm": [[hx(l) for l in row] for row in lm0], "steps": []}
    xt = [0.0, 0.0, 0.0]
    for t in range(T):
        u = [1.0, 0.1]
        xt = fs2_motion_model(xt, u, cfg["dt"])
        obs = []
        if no_obs_at != t:
            for l, (lx, ly) in enumerate(lm_true):                    # get_observations fs2.rs:392-416
                dx, dy = lx - xt[0], ly - xt[1]
                d = math.sqrt(dx * dx + dy * dy)
                if d <= cfg["max_range"]:
                    ang = normalize_angle(math.atan2(dy, dx) - xt[2])
                    obs.append([d + rng.normal() * math.sqrt(cfg["r00"]), ang + rng.normal() * math.sqrt(cfg["r11"]), l])
        if t == 2 and len(obs) > 1:
            obs.append(list(obs[0]))                                  # duplicate lm_id
        if t % 2 == 1 and len(obs) > 1:
            obs = obs[1:] + obs[:1]                                   # another landmark leads the list (it feeds the proposal)
        z0 = rng.normal(size=n).tolist()
        z1 = rng.normal(size=n).tolist()
        z2 = rng.normal(size=n).tolist()
        u01 = float(rng.uniform())
        if zero_weights_at == t:
            for p in ps:
                p.w = 0.0
        for i, p in enumerate(ps):                                    # fastslam2_update_with_rng fs2.rs:339-366
            if obs:
                mean, cov = fs2_compute_proposal(p, u, obs[0][:2], obs[0][2], cfg)
                sp = fs2_sample_pose(mean, cov, [z0[i], z1[i], z2[i]])
            else:
                un = [u[0] + z0[i] * math.sqrt(cfg["q00"]), u[1] + z1[i] * math.sqrt(cfg["q11"])]
                sp = fs2_motion_model([p.x, p.y, p.yaw], un, cfg["dt"])
            p.x, p.y, p.yaw = sp[0], sp[1], normalize_angle(sp[2])     # set_pose fs2.rs:77-81
            for (d, a, l) in obs:
                p.w *= fs2_update_landmark_and_weight(p, [d, a], l, cfg)
        fs_normalize(ps)                                              # fs2.rs:368-373 (same text as fs1)
        neff = fs_neff(ps)
        did = neff < cfg["nth"]
        idxs = []
        if did:
            ps, idxs = fs_resample(ps, u01)
        pose, lm = fs_state(ps)
        best = 0
        for i in range(1, n):
            if ps[i].w >= ps[best].w:
                best = i
        case["steps"].append({"u": hx(u), "obs": [[hx(o[0]), hx(o[1]), int(o[2])] for o in obs], "z0": hx(z0),
                              "z1": hx(z1), "z2": hx(z2), "u01": hx(u01), "neff": hx(neff), "did_resample": bool(did),
                              "zero_weights": zero_weights_at == t, "indices": idxs, "best": best,
                              "pose": [hx(r) for r in pose], "lm": [[hx(l) for l in row] for row in lm]})
    return case


# ------------------------------------------------------------------------------------------------
# hand-checkable known answers
# ------------------------------------------------------------------------------------------------
def kat():
    out = {}
    # two particles, one observation: A=(3,4) -> d_pred 5, diff 0.5; B=(0,5.5) -> diff 0; sigma = 0.5
    coeff = 1.0 / math.sqrt(2.0 * PI * 0.25)
    raw_a = 1.0 * (coeff * math.exp(-(0.5 * 0.5) / (2.0 * 0.25)))
    raw_b = 1.0 * (coeff * math.exp(-(0.0 * 0.0) / (2.0 * 0.25)))
    s = 0.0 + raw_a + raw_b
    out["likelihood_2x1"] = {"particles": [hx([3.0, 4.0, 0.0, 0.0, 0.5]), hx([0.0, 5.5, 0.0, 0.0, 0.5])],
                             "obs": hx([5.5, 0.0, 0.0]), "sigma": hx(0.5), "raw": hx([raw_a, raw_b]),
                             "normalised": hx([raw_a / s, raw_b / s])}
    # 4-particle cumsum, fixed r list, three index rules
    w = [0.1, 0.2, 0.3, 0.4]
    cum = []
    c = 0.0
    for v in w:
        c += v
        cum.append(c)
    rs = [0.05, 0.1, 0.10000000000000002, 0.3, 0.31, 0.6000000000000001, 0.99, 1.0, 1.5]

    def first_le(r, fb):
        for i, cw in enumerate(cum):
            if r <= cw:
                return i
        return fb
    cum_mcl = list(cum)
    cum_mcl[-1] = 1.0

    def first_le_mcl(r):
        for i, cw in enumerate(cum_mcl):
            if r <= cw:
                return i
        return len(cum_mcl) - 1
    out["index_rules"] = {"w": hx(w), "cum": hx(cum), "r": hx(rs),
                          "pf": [first_le(r, 0) for r in rs],          # pf.rs:459-465 default 0
                          "mcl": [first_le_mcl(r) for r in rs]}        # mcl.rs:334-336,387-392
    # FastSLAM systematic on the same weights: r0 = 0.2*(1/4)
    cum0 = [0.0] + cum
    r = 0.2 * (1.0 / 4.0 - 0.0) + 0.0
    j, idx = 0, []
    for _ in range(4):
        while r > cum0[j + 1] and j < 3:
            j += 1
        idx.append(j)
        r += 1.0 / 4.0
    out["index_rules"]["fs_u01"] = hx(0.2)
    out["index_rules"]["fs"] = idx
    # one EKF update: P = 10 I, landmark at (3,4) seen from the origin with yaw 0, z = (5.2, atan2(4,3)+0.01)
    p = FP(0.25, 0.0, 0.0, 0.0, [[3.0, 4.0, 10.0, 0.0, 0.0, 10.0]])
    cfg = {"r00": 0.5, "r11": 0.0305}
    z = [5.2, math.atan2(4.0, 3.0) + 0.01]
    fs_update_landmark(p, z[0], z[1], 0, cfg)
    out["ekf_1"] = {"z": hx(z), "w0": hx(0.25), "lm_after": hx(p.lms[0]), "w_after": hx(p.w)}
    return out


def main():
    rng = np.random.default_rng(20260924)
    pf = {"cases": [
        run_pf_case("thr1_resample_every_step", rng, n=16, T=6, K=3, thr=1.0, sv=2.0, sw=math.radians(40.0), sigma=0.25, dt=0.1),
        run_pf_case("thr0.5_default", rng, n=24, T=8, K=5, thr=0.5, sv=2.0, sw=math.radians(40.0), sigma=0.2, dt=0.1),
        run_pf_case("zero_noise_and_underflow", rng, n=8, T=4, K=2, thr=0.5, sv=0.0, sw=0.0, sigma=0.2, dt=0.1, zero_weight_step=2),
        run_pf_case("no_observations", rng, n=8, T=3, K=0, thr=0.5, sv=1.0, sw=0.5, sigma=0.2, dt=0.1),
    ]}
    mcl = {"cases": [
        run_mcl_case("fixed_n", rng, nmin=16, nmax=16, T=5, K=4, sv=0.05, sw=0.02, sigma=0.25, dt=0.1),
        run_mcl_case("adaptive_8_64", rng, nmin=8, nmax=64, T=6, K=4, sv=0.3, sw=0.2, sigma=0.5, dt=0.1, spread=3.0),
    ]}
    fs = {"cases": [
        run_fs_case("ekf_live", rng, n=8, m=4, T=6, nth=8 / 1.5, init_cov=[10.0, 10.0, 10.0, 10.0]),
        run_fs_case("mixed_init_and_fresh", rng, n=12, m=5, T=6, nth=12 / 1.5, init_cov=[10.0, 1000.0, 10.0, 1000.0, 10.0]),
        run_fs_case("fresh_reference_constants", rng, n=20, m=3, T=5, nth=100.0 / 1.5, init_cov=[1000.0, 1000.0, 1000.0]),
        run_fs_case("all_zero_weights", rng, n=6, m=3, T=4, nth=6 / 1.5, init_cov=[10.0, 10.0, 10.0], zero_weights_at=1),
    ]}
    # FastSLAM 2.0 cases draw AFTER everything above, so the earlier files do not change when this list does
    fs2 = {"cases": [
        run_fs2_case("ekf_live", rng, n=8, m=4, T=6, nth=8 / 1.5, init_cov=[10.0, 10.0, 10.0, 10.0]),
        run_fs2_case("mixed_init_and_fresh", rng, n=12, m=5, T=6, nth=12 / 1.5, init_cov=[10.0, 1000.0, 10.0, 1000.0, 10.0]),
        run_fs2_case("fresh_reference_constants", rng, n=20, m=3, T=5, nth=100.0 / 1.5, init_cov=[1000.0, 1000.0, 1000.0]),
        run_fs2_case("no_observation_step_and_zero_weights", rng, n=6, m=3, T=5, nth=6 / 1.5, init_cov=[10.0, 10.0, 10.0], no_obs_at=1, zero_weights_at=3),
    ]}
    for fn, obj in (("pf_golden.json", pf), ("mcl_golden.json", mcl), ("fs1_golden.json", fs), ("kat_golden.json", kat()), ("fs2_golden.json", fs2)):
        with open(os.path.join(HERE, fn), "w") as f:
            json.dump(obj, f, separators=(",", ":"))
        print("wrote", fn, os.path.getsize(os.path.join(HERE, fn)), "bytes")


if __name__ == "__main__":
    main()
