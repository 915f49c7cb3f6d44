"""rust_robotics_b200 — B200-native particle-filter / FastSLAM 1.0 engine behind the rust_robotics API.

The product is the CUDA library (csrc/ -> libpfgpu.so, C ABI in include/pfgpu.h).  `api` mirrors the
reference's public types (ParticleFilterLocalizer, MonteCarloLocalizer, fastslam1) over that ABI with
ctypes.  There is no CPU fallback: constructing any filter without a CUDA device raises.
"""
from .api import (FastSlam1, FastSlam2, FsConfig, InvalidParameter, MonteCarloLocalizationConfig, MonteCarloLocalizer,  # noqa: F401
                  ParticleFilterConfig, ParticleFilterLocalizer, PfgpuError, load_library)
