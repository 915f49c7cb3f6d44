"""ctypes access to the CPU oracle (oracle/liboracle*.so).  Test infrastructure only."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

c_dp = C.POINTER(C.c_double)
c_u32p = C.POINTER(C.c_uint32)


class PfConfig(C.Structure):
    _fields_ = [("n_particles", C.c_uint64), ("resample_threshold", C.c_double), ("range_noise", C.c_double),
                ("velocity_noise", C.c_double), ("yaw_rate_noise", C.c_double), ("dt", C.c_double),
                ("mode", C.c_int32), ("_pad", C.c_int32), ("max_particles", C.c_uint64),
                ("kld_epsilon", C.c_double), ("kld_z", C.c_double)]


class FsConfig(C.Structure):
    _fields_ = [("dt", C.c_double), ("max_range", C.c_double), ("nth", C.c_double), ("q00", C.c_double),
                ("q11", C.c_double), ("r00", C.c_double), ("r11", C.c_double), ("init_weight", C.c_double)]


class FsObs(C.Structure):
    _fields_ = [("d", C.c_double), ("angle", C.c_double), ("lm_id", C.c_uint64)]


def build_oracle():
    subprocess.run(["make", "-C", ORACLE_DIR, "-s"], check=True)


def _dp(a):
    return a.ctypes.data_as(c_dp)


def load(libm=False):
    name = "liboracle_libm.so" if libm else "liboracle.so"
    path = os.path.join(ORACLE_DIR, name)
    if not os.path.exists(path):
        build_oracle()
    L = C.CDLL(path)
    L.orc_pf_new.restype = C.c_void_p
    L.orc_pf_new.argtypes = [C.POINTER(PfConfig), C.c_uint64]
    L.orc_pf_free.argtypes = [C.c_void_p]
    L.orc_pf_init_state.argtypes = [C.c_void_p, c_dp]
    L.orc_pf_count.restype = C.c_size_t
    L.orc_pf_count.argtypes = [C.c_void_p]
    L.orc_pf_set_particles.argtypes = [C.c_void_p, c_dp, C.c_size_t]
    L.orc_pf_get_particles.argtypes = [C.c_void_p, c_dp]
    L.orc_pf_predict.argtypes = [C.c_void_p, c_dp]
    L.orc_pf_predict_with_noise.argtypes = [C.c_void_p, c_dp, c_dp, c_dp]
    L.orc_pf_update.argtypes = [C.c_void_p, c_dp, C.c_size_t]
    L.orc_pf_resample.argtypes = [C.c_void_p]
    L.orc_pf_resample_with_uniforms.argtypes = [C.c_void_p, c_dp, C.c_size_t]
    L.orc_pf_step.argtypes = [C.c_void_p, c_dp, c_dp, C.c_size_t, c_dp]
    L.orc_pf_estimate.argtypes = [C.c_void_p, c_dp, c_dp]
    L.orc_pf_neff.restype = C.c_double
    L.orc_pf_neff.argtypes = [C.c_void_p]
    L.orc_pf_set_range_noise.argtypes = [C.c_void_p, C.c_double]
    L.orc_pf_last_indices.restype = C.c_size_t
    L.orc_pf_last_indices.argtypes = [C.c_void_p, c_u32p, C.c_size_t]
    L.orc_pf_set_fast_search.argtypes = [C.c_void_p, C.c_int]
    L.orc_pf_set_threads.argtypes = [C.c_void_p, C.c_int]
    L.orc_fs_default_config.argtypes = [C.POINTER(FsConfig)]
    L.orc_fs_new.restype = C.c_void_p
    L.orc_fs_new.argtypes = [C.POINTER(FsConfig), C.c_size_t, C.c_size_t, C.c_uint64]
    L.orc_fs_free.argtypes = [C.c_void_p]
    L.orc_fs_set_state.argtypes = [C.c_void_p, c_dp, c_dp]
    L.orc_fs_get_state.argtypes = [C.c_void_p, c_dp, c_dp]
    L.orc_fs_seed_map.argtypes = [C.c_void_p, c_dp, c_dp, C.c_double, C.c_double]
    L.orc_fs_step.argtypes = [C.c_void_p, c_dp, C.POINTER(FsObs), C.c_size_t]
    L.orc_fs_step_with_noise.argtypes = [C.c_void_p, c_dp, C.POINTER(FsObs), C.c_size_t, c_dp, c_dp, C.c_double]
    L.orc_fs_best.restype = C.c_size_t
    L.orc_fs_best.argtypes = [C.c_void_p]
    L.orc_fs_last_indices.restype = C.c_size_t
    L.orc_fs_last_indices.argtypes = [C.c_void_p, c_u32p, C.c_size_t]
    L.orc_fs_last_neff.restype = C.c_double
    L.orc_fs_last_neff.argtypes = [C.c_void_p]
    L.orc_fs_get_observations.restype = C.c_size_t
    L.orc_fs_get_observations.argtypes = [C.POINTER(FsConfig), c_dp, c_dp, C.c_size_t, C.c_uint64, C.c_uint32,
                                          C.POINTER(FsObs)]
    L.orc_fs_set_threads.argtypes = [C.c_void_p, C.c_int]
    L.orc_fs_set_variant.argtypes = [C.c_void_p, C.c_int]
    L.orc_fs2_compute_proposal.argtypes = [C.POINTER(FsConfig), c_dp, c_dp, C.c_double, C.c_double, c_dp, c_dp, c_dp]
    L.orc_fs2_sample_pose.argtypes = [c_dp, c_dp, c_dp, c_dp]
    for fn in ("orc_math_exp", "orc_math_log", "orc_math_sin", "orc_math_cos"):
        getattr(L, fn).argtypes = [c_dp, c_dp, C.c_size_t]
    L.orc_math_atan2.argtypes = [c_dp, c_dp, c_dp, C.c_size_t]
    L.orc_philox.argtypes = [C.c_uint32] * 6 + [c_u32p]
    L.orc_normal_pair.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint64, c_dp]
    L.orc_uniform53.restype = C.c_double
    L.orc_uniform53.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint64]
    L.orc_uniform52.restype = C.c_double
    L.orc_uniform52.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint64]
    return L


def f64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


class OraclePF:
    """ParticleFilterLocalizer / MonteCarloLocalizer oracle (pf.rs / mcl.rs)."""

    def __init__(self, L, n, threshold=0.5, range_noise=0.2, velocity_noise=2.0, yaw_rate_noise=np.deg2rad(40.0),
                 dt=0.1, seed=42, mode=0, max_particles=None, kld_epsilon=0.05, kld_z=2.326):
        self.L = L
        self.cfg = PfConfig(n, threshold, range_noise, velocity_noise, yaw_rate_noise, dt, mode, 0,
                            max_particles if max_particles is not None else n, kld_epsilon, kld_z)
        self.h = L.orc_pf_new(C.byref(self.cfg), seed)
        if not self.h:
            raise ValueError("InvalidParameter")
        self.cap = int(self.cfg.max_particles) if mode == 1 else n

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_pf_free(self.h)
            self.h = None

    def init_state(self, s):
        s = f64(s)
        return self.L.orc_pf_init_state(self.h, _dp(s))

    def count(self):
        return int(self.L.orc_pf_count(self.h))

    def set_particles(self, aos5):
        a = f64(aos5)
        self.L.orc_pf_set_particles(self.h, _dp(a), a.shape[0])

    def particles(self):
        a = np.empty((self.count(), 5))
        self.L.orc_pf_get_particles(self.h, _dp(a))
        return a

    def predict(self, u, zv=None, zw=None):
        u = f64(u)
        if zv is None:
            return self.L.orc_pf_predict(self.h, _dp(u))
        zv, zw = f64(zv), f64(zw)
        return self.L.orc_pf_predict_with_noise(self.h, _dp(u), _dp(zv), _dp(zw))

    def update(self, obs):
        o = f64(obs).reshape(-1, 3)
        return self.L.orc_pf_update(self.h, _dp(o), o.shape[0])

    def resample(self, r=None):
        if r is None:
            return self.L.orc_pf_resample(self.h)
        r = f64(r)
        return self.L.orc_pf_resample_with_uniforms(self.h, _dp(r), r.size)

    def step(self, u, obs):
        u = f64(u)
        o = f64(obs).reshape(-1, 3)
        est = np.empty(4)
        did = self.L.orc_pf_step(self.h, _dp(u), _dp(o), o.shape[0], _dp(est))
        return est, did

    def estimate(self):
        est, cov = np.empty(4), np.empty(16)
        self.L.orc_pf_estimate(self.h, _dp(est), _dp(cov))
        return est, cov

    def neff(self):
        return float(self.L.orc_pf_neff(self.h))

    def last_indices(self):
        idx = np.empty(self.cap, dtype=np.uint32)
        n = self.L.orc_pf_last_indices(self.h, idx.ctypes.data_as(c_u32p), idx.size)
        return idx[:n].copy()


class OracleFS:
    """FastSLAM oracle: variant 1 = FastSLAM 1.0 (fs1.rs), variant 2 = FastSLAM 2.0 (fs2.rs)."""

    def __init__(self, L, n, m, seed=42, variant=1, **cfg):
        self.L = L
        self.cfg = FsConfig()
        L.orc_fs_default_config(C.byref(self.cfg))
        for k, v in cfg.items():
            setattr(self.cfg, k, v)
        self.n, self.m = n, m
        self.h = L.orc_fs_new(C.byref(self.cfg), n, m, seed)
        self.variant = variant
        if variant != 1:
            L.orc_fs_set_variant(self.h, variant)

    def compute_proposal(self, pose3, u, z, lm6):
        """compute_proposal fs2.rs:173-216 -> (mean[3], cov[3,3])"""
        mean, cov = np.empty(3), np.empty(9)
        p, uu, l = f64(pose3), f64(u), f64(lm6)
        self.L.orc_fs2_compute_proposal(C.byref(self.cfg), _dp(p), _dp(uu), float(z[0]), float(z[1]), _dp(l), _dp(mean), _dp(cov))
        return mean, cov.reshape(3, 3)

    def sample_pose(self, mean3, cov33, n3):
        out = np.empty(3)
        m, c, n = f64(mean3), f64(cov33).reshape(9), f64(n3)
        self.L.orc_fs2_sample_pose(_dp(m), _dp(c), _dp(n), _dp(out))
        return out

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_fs_free(self.h)
            self.h = None

    def set_state(self, pose_w, lm=None):
        p = f64(pose_w)
        if lm is None:
            self.L.orc_fs_set_state(self.h, _dp(p), None)
        else:
            l = f64(lm)
            self.L.orc_fs_set_state(self.h, _dp(p), _dp(l))

    def seed_map(self, pose3, landmarks_xy, sigma=1.0, cov0=10.0):
        p, l = f64(pose3), f64(landmarks_xy)
        self.L.orc_fs_seed_map(self.h, _dp(p), _dp(l), float(sigma), float(cov0))

    def state(self):
        p = np.empty((self.n, 4))
        l = np.empty((self.n, self.m, 6))
        self.L.orc_fs_get_state(self.h, _dp(p), _dp(l))
        return p, l

    @staticmethod
    def obs_array(obs):
        arr = (FsObs * max(len(obs), 1))()
        for i, (d, a, l) in enumerate(obs):
            arr[i].d, arr[i].angle, arr[i].lm_id = float(d), float(a), int(l)
        return arr

    def step(self, u, obs, z0=None, z1=None, u01=None):
        u = f64(u)
        arr = self.obs_array(obs)
        if z0 is None:
            return self.L.orc_fs_step(self.h, _dp(u), arr, len(obs))
        z0, z1 = f64(z0), f64(z1)
        return self.L.orc_fs_step_with_noise(self.h, _dp(u), arr, len(obs), _dp(z0), _dp(z1), float(u01))

    def best(self):
        return int(self.L.orc_fs_best(self.h))

    def last_indices(self):
        idx = np.empty(self.n, dtype=np.uint32)
        n = self.L.orc_fs_last_indices(self.h, idx.ctypes.data_as(c_u32p), idx.size)
        return idx[:n].copy()

    def last_neff(self):
        return float(self.L.orc_fs_last_neff(self.h))

    def observations(self, x_true, landmarks_xy, seed, call):
        xt = f64(x_true)
        lm = f64(landmarks_xy).reshape(-1, 2)
        out = (FsObs * max(lm.shape[0], 1))()
        k = self.L.orc_fs_get_observations(C.byref(self.cfg), _dp(xt), _dp(lm), lm.shape[0], seed, call, out)
        return [(out[i].d, out[i].angle, int(out[i].lm_id)) for i in range(k)]
