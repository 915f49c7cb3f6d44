"""Python mirror of the reference's public API for the hot path, over the C ABI (include/pfgpu.h).

Names and argument meanings follow the Rust reference so the parity tests read like its own tests:
  ParticleFilterConfig / ParticleFilterLocalizer   crates/rust_robotics_localization/src/particle_filter.rs:50-573
  MonteCarloLocalizationConfig / MonteCarloLocalizer   .../monte_carlo_localization.rs:50-471
  FastSlam1 (create_particles / fastslam_update / get_best_particle)   crates/rust_robotics_slam/src/fastslam1.rs:237-306
Errors: status < 0 -> InvalidParameter (RoboticsError::InvalidParameter, rust_robotics_core/src/error.rs:8-24);
status > 0 -> PfgpuError (CUDA/NCCL).  No CPU fallback exists.
"""
import ctypes as C
import math
import os

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
_LIB = None

c_dp = C.POINTER(C.c_double)
c_u32p = C.POINTER(C.c_uint32)


class InvalidParameter(ValueError):
    """RoboticsError::InvalidParameter"""


class PfgpuError(RuntimeError):
    """CUDA / NCCL failure, or no device"""


class _PfCfg(C.Structure):
    _fields_ = [("n_particles", C.c_uint64), ("resample_threshold", C.c_double), ("range_noise", C.c_double),
                ("velocity_noise", C.c_double), ("yaw_rate_noise", C.c_double), ("dt", C.c_double),
                ("mode", C.c_int32), ("_pad", C.c_int32), ("max_particles", C.c_uint64),
                ("kld_epsilon", C.c_double), ("kld_z", C.c_double)]


class _FsCfg(C.Structure):
    _fields_ = [("dt", C.c_double), ("max_range", C.c_double), ("nth", C.c_double), ("q00", C.c_double),
                ("q11", C.c_double), ("r00", C.c_double), ("r11", C.c_double), ("init_weight", C.c_double)]


class _FsObs(C.Structure):
    _fields_ = [("d", C.c_double), ("angle", C.c_double), ("lm_id", C.c_uint64)]


class Stats(C.Structure):
    _fields_ = [("kernel_launches", C.c_uint64), ("steps", C.c_uint64), ("resamples", C.c_uint64),
                ("serial_fallbacks", C.c_uint64), ("xsum_dirty_last", C.c_uint64),
                ("main_kernel_ms_sum", C.c_double), ("main_kernel_count", C.c_uint64), ("compactions", C.c_uint64),
                ("imported_particles", C.c_uint64)]


EXPORTS = [
    "pfgpu_strerror", "pfgpu_last_error", "pfgpu_device_count",
    "pfgpu_pf_default_config", "pfgpu_pf_config_validate", "pfgpu_pf_create", "pfgpu_pf_create_sharded",
    "pfgpu_pf_destroy", "pfgpu_pf_init_state", "pfgpu_pf_upload", "pfgpu_pf_download", "pfgpu_pf_count",
    "pfgpu_pf_predict", "pfgpu_pf_update", "pfgpu_pf_resample", "pfgpu_pf_step", "pfgpu_pf_estimate",
    "pfgpu_pf_neff", "pfgpu_pf_set_range_noise", "pfgpu_pf_last_indices", "pfgpu_pf_sync",
    "pfgpu_fs_default_config", "pfgpu_fs_create", "pfgpu_fs_create_sharded", "pfgpu_fs_create_sharded_local", "pfgpu_fs_destroy",
    "pfgpu_fs_upload", "pfgpu_fs_download", "pfgpu_fs_seed_map", "pfgpu_fs_step", "pfgpu_fs_best", "pfgpu_fs_particle_landmarks",
    "pfgpu_fs_get_observations", "pfgpu_fs_last_indices", "pfgpu_fs_last_neff", "pfgpu_fs_last_gate", "pfgpu_fs_set_variant", "pfgpu_fs_count", "pfgpu_fs_sync",
    "pfgpu_nccl_unique_id", "pfgpu_pf_stats", "pfgpu_fs_stats", "pfgpu_pf_time_main_kernel",
    "pfgpu_fs_time_main_kernel", "pfgpu_pf_mark", "pfgpu_pf_elapsed_ms", "pfgpu_fs_mark", "pfgpu_fs_elapsed_ms",
    "pfgpu_pf_flush_l2", "pfgpu_fs_flush_l2", "pfgpu_fs_post_trace", "pfgpu_fs_shard_mode",
]


def library_path():
    return os.path.join(_PKG, "libpfgpu.so")


def load_library():
    """Load libpfgpu.so (never builds; never falls back).  Raises PfgpuError if it is missing."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise PfgpuError(f"{path} not built: run `python -c 'import __graft_entry__ as g; g.build()'`")
    L = C.CDLL(path)
    vp = C.c_void_p
    L.pfgpu_strerror.restype = C.c_char_p
    L.pfgpu_last_error.restype = C.c_char_p
    L.pfgpu_device_count.argtypes = [C.POINTER(C.c_int)]
    L.pfgpu_pf_default_config.argtypes = [C.POINTER(_PfCfg), C.c_int]
    L.pfgpu_pf_config_validate.argtypes = [C.POINTER(_PfCfg)]
    L.pfgpu_pf_create.argtypes = [C.POINTER(_PfCfg), C.c_uint64, C.c_int, C.POINTER(vp)]
    L.pfgpu_pf_create_sharded.argtypes = [C.POINTER(_PfCfg), C.c_uint64, C.c_int, vp, C.c_int, C.c_int, C.POINTER(vp)]
    L.pfgpu_pf_destroy.argtypes = [vp]
    L.pfgpu_pf_destroy.restype = None
    L.pfgpu_pf_init_state.argtypes = [vp, c_dp]
    L.pfgpu_pf_upload.argtypes = [vp, c_dp, C.c_size_t]
    L.pfgpu_pf_download.argtypes = [vp, c_dp, C.c_size_t]
    L.pfgpu_pf_count.argtypes = [vp, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
    L.pfgpu_pf_predict.argtypes = [vp, c_dp]
    L.pfgpu_pf_update.argtypes = [vp, c_dp, C.c_size_t]
    L.pfgpu_pf_resample.argtypes = [vp, C.POINTER(C.c_int)]
    L.pfgpu_pf_step.argtypes = [vp, c_dp, c_dp, C.c_size_t, c_dp]
    L.pfgpu_pf_estimate.argtypes = [vp, c_dp, c_dp]
    L.pfgpu_pf_neff.argtypes = [vp, c_dp]
    L.pfgpu_pf_set_range_noise.argtypes = [vp, C.c_double]
    L.pfgpu_pf_last_indices.argtypes = [vp, c_u32p, C.c_size_t, C.POINTER(C.c_size_t)]
    L.pfgpu_pf_sync.argtypes = [vp]
    L.pfgpu_fs_default_config.argtypes = [C.POINTER(_FsCfg)]
    L.pfgpu_fs_create.argtypes = [C.POINTER(_FsCfg), C.c_size_t, C.c_size_t, C.c_uint64, C.c_int, C.POINTER(vp)]
    L.pfgpu_fs_create_sharded.argtypes = [C.POINTER(_FsCfg), C.c_size_t, C.c_size_t, C.c_uint64, C.c_int, vp, C.c_int,
                                          C.c_int, C.POINTER(vp)]
    L.pfgpu_fs_create_sharded_local.argtypes = [C.POINTER(_FsCfg), C.c_size_t, C.c_size_t, C.c_uint64, C.POINTER(C.c_int), C.c_int,
                                                C.POINTER(vp)]
    L.pfgpu_fs_destroy.argtypes = [vp]
    L.pfgpu_fs_destroy.restype = None
    L.pfgpu_fs_upload.argtypes = [vp, c_dp, c_dp, C.c_size_t]
    L.pfgpu_fs_download.argtypes = [vp, c_dp, c_dp, C.c_size_t]
    L.pfgpu_fs_seed_map.argtypes = [vp, c_dp, c_dp, C.c_size_t, C.c_double, C.c_double]
    L.pfgpu_fs_step.argtypes = [vp, c_dp, C.POINTER(_FsObs), C.c_size_t, C.POINTER(C.c_int)]
    L.pfgpu_fs_best.argtypes = [vp, C.POINTER(C.c_size_t), c_dp]
    L.pfgpu_fs_particle_landmarks.argtypes = [vp, C.c_size_t, c_dp]
    L.pfgpu_fs_last_indices.argtypes = [vp, c_u32p, C.c_size_t, C.POINTER(C.c_size_t)]
    L.pfgpu_fs_last_neff.argtypes = [vp, c_dp]
    L.pfgpu_fs_last_gate.argtypes = [vp, C.POINTER(C.c_int)]
    L.pfgpu_fs_set_variant.argtypes = [vp, C.c_int]
    L.pfgpu_fs_get_observations.argtypes = [vp, c_dp, c_dp, C.c_size_t, C.c_uint32, C.POINTER(_FsObs), C.POINTER(C.c_size_t)]
    L.pfgpu_fs_count.argtypes = [vp, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
    L.pfgpu_fs_sync.argtypes = [vp]
    L.pfgpu_nccl_unique_id.argtypes = [vp]
    L.pfgpu_pf_stats.argtypes = [vp, C.POINTER(Stats)]
    L.pfgpu_fs_stats.argtypes = [vp, C.POINTER(Stats)]
    L.pfgpu_pf_time_main_kernel.argtypes = [vp, C.c_int]
    L.pfgpu_fs_time_main_kernel.argtypes = [vp, C.c_int]
    for k in ("pf", "fs"):
        getattr(L, f"pfgpu_{k}_mark").argtypes = [vp, C.c_int]
        getattr(L, f"pfgpu_{k}_elapsed_ms").argtypes = [vp, C.c_int, C.c_int, c_dp]
        getattr(L, f"pfgpu_{k}_flush_l2").argtypes = [vp]
    L.pfgpu_fs_post_trace.argtypes = [vp, C.POINTER(C.c_ulonglong)]
    L.pfgpu_fs_shard_mode.argtypes = [vp, C.POINTER(C.c_int)]
    L.pfgpu_test_div.argtypes = [C.c_ulonglong, C.c_uint64, C.POINTER(C.c_ulonglong), C.c_int]
    L.pfgpu_test_xsum.argtypes = [c_dp, C.c_size_t, c_dp, c_dp, C.POINTER(C.c_int), C.c_int]
    _LIB = L
    return L


def _check(L, rc):
    if rc == 0:
        return
    msg = L.pfgpu_strerror(rc).decode()
    if rc < 0:
        raise InvalidParameter(msg)
    raise PfgpuError(f"{msg}: {L.pfgpu_last_error().decode()}")


def _f64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def _dp(a):
    return a.ctypes.data_as(c_dp)


# ------------------------------------------------------------------------------------------------
class ParticleFilterConfig:
    """pf.rs:50-78"""

    def __init__(self, n_particles=100, resample_threshold=0.5, range_noise=0.2, velocity_noise=2.0,
                 yaw_rate_noise=math.radians(40.0), dt=0.1):
        self.n_particles, self.resample_threshold, self.range_noise = n_particles, resample_threshold, range_noise
        self.velocity_noise, self.yaw_rate_noise, self.dt = velocity_noise, yaw_rate_noise, dt

    def _c(self):
        return _PfCfg(self.n_particles, self.resample_threshold, self.range_noise, self.velocity_noise,
                      self.yaw_rate_noise, self.dt, 0, 0, self.n_particles, 0.05, 2.326)

    def validate(self):
        L = load_library()
        _check(L, L.pfgpu_pf_config_validate(C.byref(self._c())))


class MonteCarloLocalizationConfig:
    """mcl.rs:50-74"""

    def __init__(self, min_particles=100, max_particles=5000, kld_epsilon=0.05, kld_z=2.326, range_noise=0.2,
                 velocity_noise=2.0, yaw_rate_noise=math.radians(40.0), dt=0.1):
        self.min_particles, self.max_particles, self.kld_epsilon, self.kld_z = min_particles, max_particles, kld_epsilon, kld_z
        self.range_noise, self.velocity_noise, self.yaw_rate_noise, self.dt = range_noise, velocity_noise, yaw_rate_noise, dt

    def _c(self):
        return _PfCfg(self.min_particles, 0.0, self.range_noise, self.velocity_noise, self.yaw_rate_noise, self.dt,
                      1, 0, self.max_particles, self.kld_epsilon, self.kld_z)

    def validate(self):
        L = load_library()
        _check(L, L.pfgpu_pf_config_validate(C.byref(self._c())))


class _PfBase:
    def __init__(self, ccfg, seed, device, shard=None):
        self.L = load_library()
        self.h = C.c_void_p()
        if shard is None:
            _check(self.L, self.L.pfgpu_pf_create(C.byref(ccfg), seed, device, C.byref(self.h)))
        else:
            uid, rank, world = shard
            buf = C.create_string_buffer(bytes(uid), 128)
            _check(self.L, self.L.pfgpu_pf_create_sharded(C.byref(ccfg), seed, device, buf, rank, world, C.byref(self.h)))

    def close(self):
        if getattr(self, "h", None) and self.h.value:
            self.L.pfgpu_pf_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- reference API --
    def try_predict_with_control(self, control):
        u = _f64(control)
        _check(self.L, self.L.pfgpu_pf_predict(self.h, _dp(u)))

    predict_with_control = try_predict_with_control

    def try_update_with_observations(self, observations):
        o = _f64(observations).reshape(-1, 3)
        _check(self.L, self.L.pfgpu_pf_update(self.h, _dp(o), o.shape[0]))

    update_with_observations = try_update_with_observations

    def resample(self):
        did = C.c_int()
        _check(self.L, self.L.pfgpu_pf_resample(self.h, C.byref(did)))
        return bool(did.value)

    def try_step(self, control, observations, want_estimate=True):
        u = control if isinstance(control, np.ndarray) and control.dtype == np.float64 else _f64(control)
        o = _f64(observations).reshape(-1, 3)
        est = np.empty(4)
        _check(self.L, self.L.pfgpu_pf_step(self.h, _dp(u), _dp(o), o.shape[0], _dp(est) if want_estimate else None))
        return est if want_estimate else None

    step = try_step

    # -- impl StateEstimator (traits.rs:31-52; pf.rs:552-573, mcl.rs:450-471) --
    def predict(self, control, dt=None):
        """`predict(&mut self, control, _dt)`: the dt argument is ignored, the filter steps with its configured dt (pf.rs:558-560)"""
        self.try_predict_with_control(control)

    def update(self, measurement):
        """`update(&mut self, measurement)` = update_with_observations + resample (pf.rs:562-565)"""
        self.try_update_with_observations(measurement)
        self.resample()

    def get_state(self):
        return self.estimate()

    def get_covariance(self):
        return self.calc_covariance()

    def estimate(self):
        est = np.empty(4)
        _check(self.L, self.L.pfgpu_pf_estimate(self.h, _dp(est), None))
        return est

    def calc_covariance(self):
        cov = np.empty(16)
        _check(self.L, self.L.pfgpu_pf_estimate(self.h, None, _dp(cov)))
        return cov.reshape(4, 4).T        # column-major -> [i, j]

    def state_2d(self):
        return tuple(self.estimate())

    def get_particles(self):
        n = self.particle_count(local=True)
        a = np.empty((n, 5))
        _check(self.L, self.L.pfgpu_pf_download(self.h, _dp(a), n))
        return a

    def set_particles(self, aos5):
        a = _f64(aos5)
        _check(self.L, self.L.pfgpu_pf_upload(self.h, _dp(a), a.shape[0]))

    def particle_count(self, local=False):
        nl, ng = C.c_size_t(), C.c_size_t()
        _check(self.L, self.L.pfgpu_pf_count(self.h, C.byref(nl), C.byref(ng)))
        return nl.value if local else ng.value

    def set_range_noise(self, s):
        _check(self.L, self.L.pfgpu_pf_set_range_noise(self.h, float(s)))

    # -- parity / bench hooks --
    def n_eff(self):
        v = C.c_double()
        _check(self.L, self.L.pfgpu_pf_neff(self.h, C.byref(v)))
        return v.value

    def last_indices(self):
        n = self.particle_count(local=True)
        idx = np.empty(n, dtype=np.uint32)
        cnt = C.c_size_t()
        _check(self.L, self.L.pfgpu_pf_last_indices(self.h, idx.ctypes.data_as(c_u32p), n, C.byref(cnt)))
        return idx[:cnt.value]

    def sync(self):
        _check(self.L, self.L.pfgpu_pf_sync(self.h))

    def stats(self):
        s = Stats()
        _check(self.L, self.L.pfgpu_pf_stats(self.h, C.byref(s)))
        return s

    def time_main_kernel(self, on=True):
        _check(self.L, self.L.pfgpu_pf_time_main_kernel(self.h, int(on)))

    def mark(self, slot):
        _check(self.L, self.L.pfgpu_pf_mark(self.h, slot))

    def elapsed_ms(self, a, b):
        v = C.c_double()
        _check(self.L, self.L.pfgpu_pf_elapsed_ms(self.h, a, b, C.byref(v)))
        return v.value

    def flush_l2(self):
        _check(self.L, self.L.pfgpu_pf_flush_l2(self.h))


class ParticleFilterLocalizer(_PfBase):
    """pf.rs:121-573"""

    def __init__(self, config=None, seed=42, device=0, shard=None):
        self.config = config or ParticleFilterConfig()
        super().__init__(self.config._c(), seed, device, shard)

    @classmethod
    def try_new(cls, config, **kw):
        return cls(config, **kw)

    new = try_new

    @classmethod
    def with_defaults(cls, **kw):
        return cls(ParticleFilterConfig(), **kw)

    @classmethod
    def try_with_initial_state(cls, initial_state, config, **kw):
        f = cls(config, **kw)
        s = _f64(initial_state)
        _check(f.L, f.L.pfgpu_pf_init_state(f.h, _dp(s)))
        return f

    with_initial_state = try_with_initial_state
    with_initial_state_2d = try_with_initial_state


class MonteCarloLocalizer(_PfBase):
    """mcl.rs:133-471"""

    def __init__(self, config=None, seed=42, device=0, shard=None):
        self.config = config or MonteCarloLocalizationConfig()
        super().__init__(self.config._c(), seed, device, shard)

    @classmethod
    def try_new(cls, config, **kw):
        return cls(config, **kw)

    new = try_new

    @classmethod
    def try_with_initial_state(cls, initial_state, config, **kw):
        f = cls(config, **kw)
        s = _f64(initial_state)
        _check(f.L, f.L.pfgpu_pf_init_state(f.h, _dp(s)))
        return f

    with_initial_state = try_with_initial_state


# ------------------------------------------------------------------------------------------------
class FsConfig:
    """fs1.rs:13-23 module constants as fields"""

    def __init__(self, dt=0.1, max_range=20.0, nth=100.0 / 1.5, q00=0.3, q11=0.0305, r00=0.5, r11=0.0305,
                 init_weight=1.0 / 100.0):
        self.dt, self.max_range, self.nth, self.q00, self.q11 = dt, max_range, nth, q00, q11
        self.r00, self.r11, self.init_weight = r00, r11, init_weight

    def _c(self):
        return _FsCfg(self.dt, self.max_range, self.nth, self.q00, self.q11, self.r00, self.r11, self.init_weight)


class FastSlam1:
    """Engine form of fs1.rs's free functions over a caller-owned Vec<Particle>:
    create_particles -> __init__; fastslam_update -> fastslam_update/step; get_best_particle -> get_best_particle."""

    def __init__(self, n_particles, n_landmarks, config=None, seed=42, device=0, shard=None):
        self.L = load_library()
        self.config = config or FsConfig()
        self.h = C.c_void_p()
        cc = self.config._c()
        if shard is None:
            _check(self.L, self.L.pfgpu_fs_create(C.byref(cc), n_particles, n_landmarks, seed, device, C.byref(self.h)))
        else:
            uid, rank, world = shard
            buf = C.create_string_buffer(bytes(uid), 128)
            _check(self.L, self.L.pfgpu_fs_create_sharded(C.byref(cc), n_particles, n_landmarks, seed, device, buf,
                                                          rank, world, C.byref(self.h)))
        nl, ng, m = C.c_size_t(), C.c_size_t(), C.c_size_t()
        _check(self.L, self.L.pfgpu_fs_count(self.h, C.byref(nl), C.byref(ng), C.byref(m)))
        self.n_local, self.n_global, self.m = nl.value, ng.value, m.value
        if self.VARIANT != 1:
            _check(self.L, self.L.pfgpu_fs_set_variant(self.h, self.VARIANT))

    VARIANT = 1                      # which reference module the step mirrors: fastslam1 (FastSlam2 below: fastslam2)
    create_particles = classmethod(lambda cls, n, m, **kw: cls(n, m, **kw))

    @classmethod
    def create_sharded_local(cls, n_particles_global, n_landmarks, devices, config=None, seed=42):
        """All ranks of the sharded engine inside this process (pfgpu_fs_create_sharded_local): one FastSlam1 per entry of
        `devices` (entries may repeat).  Drive them with step_all()."""
        L = load_library()
        config = config or FsConfig()
        cc = config._c()
        world = len(devices)
        devs = (C.c_int * world)(*devices)
        hs = (C.c_void_p * world)()
        _check(L, L.pfgpu_fs_create_sharded_local(C.byref(cc), n_particles_global, n_landmarks, seed, devs, world, hs))
        out = []
        for r in range(world):
            g = cls.__new__(cls)
            g.L, g.config, g.h = L, config, C.c_void_p(hs[r])
            nl, ng, m = C.c_size_t(), C.c_size_t(), C.c_size_t()
            _check(L, L.pfgpu_fs_count(g.h, C.byref(nl), C.byref(ng), C.byref(m)))
            g.n_local, g.n_global, g.m = nl.value, ng.value, m.value
            if cls.VARIANT != 1:
                _check(L, L.pfgpu_fs_set_variant(g.h, cls.VARIANT))
            out.append(g)
        return out

    @staticmethod
    def step_all(ranks, u, z):
        """one fastslam_update on every in-process rank: enqueue everywhere first, then synchronise; returns did_resample"""
        for g in ranks:
            g.fastslam_update(u, z, want_flag=False)
        for g in ranks:
            g.sync()
        return ranks[0].did_resample()

    def did_resample(self):
        """whether the last step resampled (synchronises)"""
        return self.last_gate()

    def last_gate(self):
        idx = C.c_size_t()
        _check(self.L, self.L.pfgpu_fs_best(self.h, C.byref(idx), None))      # synchronises the stream
        g = C.c_int()
        _check(self.L, self.L.pfgpu_fs_last_gate(self.h, C.byref(g)))
        return bool(g.value)

    def close(self):
        if getattr(self, "h", None) and self.h.value:
            self.L.pfgpu_fs_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def _obs(z):
        arr = (_FsObs * max(len(z), 1))()
        for i, (d, a, l) in enumerate(z):
            arr[i].d, arr[i].angle, arr[i].lm_id = float(d), float(a), int(l)
        return arr

    def fastslam_update(self, u, z, want_flag=True, obs_array=None):
        """fs1.rs:237-266.  Returns whether the step resampled (None when want_flag is False: no host sync)."""
        uu = _f64(u)
        arr = obs_array if obs_array is not None else self._obs(z)
        k = len(z)
        did = C.c_int()
        _check(self.L, self.L.pfgpu_fs_step(self.h, _dp(uu), arr, k, C.byref(did) if want_flag else None))
        return bool(did.value) if want_flag else None

    step = fastslam_update

    def get_observations(self, x_true, landmarks_xy, call):
        """fs1.rs:277-299 on the device (Philox stream OBS keyed by the handle's seed, `call` and the landmark id)"""
        xt, lm = _f64(x_true), _f64(landmarks_xy)
        n = lm.size // 2
        out = (_FsObs * max(n, 1))()
        k = C.c_size_t()
        _check(self.L, self.L.pfgpu_fs_get_observations(self.h, _dp(xt), _dp(lm), n, call, out, C.byref(k)))
        return [(out[i].d, out[i].angle, int(out[i].lm_id)) for i in range(k.value)]

    def get_best_particle(self):
        idx = C.c_size_t()
        pw = np.empty(4)
        _check(self.L, self.L.pfgpu_fs_best(self.h, C.byref(idx), _dp(pw)))
        return idx.value, pw

    def particle_landmarks(self, i):
        out = np.empty((self.m, 6))
        _check(self.L, self.L.pfgpu_fs_particle_landmarks(self.h, i, _dp(out)))
        return out

    def set_state(self, pose_w, lm=None):
        p = _f64(pose_w)
        l = _f64(lm) if lm is not None else None
        _check(self.L, self.L.pfgpu_fs_upload(self.h, _dp(p), _dp(l) if l is not None else None, p.shape[0]))

    def seed_map(self, pose3, landmarks_xy, sigma=1.0, cov0=10.0):
        """initialised map (EKF branch live from step 0); see include/pfgpu.h pfgpu_fs_seed_map"""
        p, l = _f64(pose3), _f64(landmarks_xy)
        _check(self.L, self.L.pfgpu_fs_seed_map(self.h, _dp(p), _dp(l), l.size // 2, float(sigma), float(cov0)))

    def state(self, landmarks=True):
        p = np.empty((self.n_local, 4))
        l = np.empty((self.n_local, self.m, 6)) if landmarks else None
        _check(self.L, self.L.pfgpu_fs_download(self.h, _dp(p), _dp(l) if landmarks else None, self.n_local))
        return p, l

    def last_indices(self):
        idx = np.empty(self.n_local, dtype=np.uint32)
        cnt = C.c_size_t()
        _check(self.L, self.L.pfgpu_fs_last_indices(self.h, idx.ctypes.data_as(c_u32p), self.n_local, C.byref(cnt)))
        return idx[:cnt.value]

    def last_neff(self):
        v = C.c_double()
        _check(self.L, self.L.pfgpu_fs_last_neff(self.h, C.byref(v)))
        return v.value

    def sync(self):
        _check(self.L, self.L.pfgpu_fs_sync(self.h))

    def stats(self):
        s = Stats()
        _check(self.L, self.L.pfgpu_fs_stats(self.h, C.byref(s)))
        return s

    def shard_mode(self):
        """0 = one GPU, 2 = sharded over peer memory (NVLink)."""
        m = C.c_int()
        _check(self.L, self.L.pfgpu_fs_shard_mode(self.h, C.byref(m)))
        return m.value

    def time_main_kernel(self, on=True):
        _check(self.L, self.L.pfgpu_fs_time_main_kernel(self.h, int(on)))

    def mark(self, slot):
        _check(self.L, self.L.pfgpu_fs_mark(self.h, slot))

    def elapsed_ms(self, a, b):
        v = C.c_double()
        _check(self.L, self.L.pfgpu_fs_elapsed_ms(self.h, a, b, C.byref(v)))
        return v.value

    def flush_l2(self):
        _check(self.L, self.L.pfgpu_fs_flush_l2(self.h))


class FastSlam2(FastSlam1):
    """rust_robotics_slam::fastslam2 (crates/rust_robotics_slam/src/fastslam2.rs): create_particles :418-422, fastslam2_update
    :376-383, get_best_particle :385-390, get_observations :418-421 over the same device-resident particle set as FastSlam1; the
    step samples every pose from the observation-informed proposal (:173-239) and runs update_landmark_and_weight (:242-280)."""
    VARIANT = 2

    def fastslam2_update(self, u, z, **kw):
        return self.fastslam_update(u, z, **kw)
