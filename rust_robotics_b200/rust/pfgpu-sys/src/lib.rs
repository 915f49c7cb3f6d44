//! pfgpu-sys — raw `extern "C"` bindings to libpfgpu.so (include/pfgpu.h).
//!
//! NOT COMPILED IN THIS REPOSITORY'S CI: the build image has no Rust toolchain.  The same ABI
//! is exercised by the C++ mirror (rust_robotics_b200/host/) and the Python mirror (rust_robotics_b200/api.py).
//! The reference's library crates are `#![forbid(unsafe_code)]` (crates/*/src/lib.rs:1), so the `extern` block lives
//! in this separate -sys crate; `rust_robotics_gpu` wraps it behind the reference's safe API.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)]
#[derive(Clone, Copy)]
pub struct pfgpu_pf_config {
    pub n_particles: u64,
    pub resample_threshold: f64,
    pub range_noise: f64,
    pub velocity_noise: f64,
    pub yaw_rate_noise: f64,
    pub dt: f64,
    pub mode: i32,
    pub _pad: i32,
    pub max_particles: u64,
    pub kld_epsilon: f64,
    pub kld_z: f64,
}
#[repr(C)]
#[derive(Clone, Copy)]
pub struct pfgpu_fs_config {
    pub dt: f64,
    pub max_range: f64,
    pub nth: f64,
    pub q00: f64,
    pub q11: f64,
    pub r00: f64,
    pub r11: f64,
    pub init_weight: f64,
}
#[repr(C)]
#[derive(Clone, Copy)]
pub struct pfgpu_fs_obs {
    pub d: f64,
    pub angle: f64,
    pub lm_id: u64,
}
pub enum pfgpu_pf {}
pub enum pfgpu_fs {}

#[link(name = "pfgpu")]
extern "C" {
    pub fn pfgpu_strerror(status: c_int) -> *const c_char;
    pub fn pfgpu_last_error() -> *const c_char;
    pub fn pfgpu_pf_config_validate(cfg: *const pfgpu_pf_config) -> c_int;
    pub fn pfgpu_pf_create(cfg: *const pfgpu_pf_config, seed: u64, device: c_int, out: *mut *mut pfgpu_pf) -> c_int;
    pub fn pfgpu_pf_create_sharded(cfg: *const pfgpu_pf_config, seed: u64, device: c_int, nccl_unique_id: *const c_void,
                                   rank: c_int, world: c_int, out: *mut *mut pfgpu_pf) -> c_int;
    pub fn pfgpu_pf_destroy(h: *mut pfgpu_pf);
    pub fn pfgpu_pf_init_state(h: *mut pfgpu_pf, init: *const f64) -> c_int;
    pub fn pfgpu_pf_upload(h: *mut pfgpu_pf, aos5: *const f64, n: usize) -> c_int;
    pub fn pfgpu_pf_download(h: *mut pfgpu_pf, aos5: *mut f64, n: usize) -> c_int;
    pub fn pfgpu_pf_count(h: *mut pfgpu_pf, n_local: *mut usize, n_global: *mut usize) -> c_int;
    pub fn pfgpu_pf_predict(h: *mut pfgpu_pf, u: *const f64) -> c_int;
    pub fn pfgpu_pf_update(h: *mut pfgpu_pf, obs3: *const f64, k: usize) -> c_int;
    pub fn pfgpu_pf_resample(h: *mut pfgpu_pf, did_resample: *mut c_int) -> c_int;
    pub fn pfgpu_pf_step(h: *mut pfgpu_pf, u: *const f64, obs3: *const f64, k: usize, est: *mut f64) -> c_int;
    pub fn pfgpu_pf_estimate(h: *mut pfgpu_pf, est: *mut f64, cov16_colmajor: *mut f64) -> c_int;
    pub fn pfgpu_pf_set_range_noise(h: *mut pfgpu_pf, range_noise: f64) -> c_int;
    pub fn pfgpu_fs_default_config(cfg: *mut pfgpu_fs_config);
    pub fn pfgpu_fs_create(cfg: *const pfgpu_fs_config, n_particles: usize, n_landmarks: usize, seed: u64, device: c_int,
                           out: *mut *mut pfgpu_fs) -> c_int;
    pub fn pfgpu_fs_destroy(h: *mut pfgpu_fs);
    pub fn pfgpu_fs_upload(h: *mut pfgpu_fs, pose_w: *const f64, lm: *const f64, n: usize) -> c_int;
    pub fn pfgpu_fs_download(h: *mut pfgpu_fs, pose_w: *mut f64, lm: *mut f64, n: usize) -> c_int;
    pub fn pfgpu_fs_step(h: *mut pfgpu_fs, u: *const f64, z: *const pfgpu_fs_obs, k: usize, did_resample: *mut c_int) -> c_int;
    pub fn pfgpu_fs_best(h: *mut pfgpu_fs, index_global: *mut usize, pose_w4: *mut f64) -> c_int;
    pub fn pfgpu_fs_get_observations(h: *mut pfgpu_fs, x_true: *const f64, landmarks_xy: *const f64, n_landmarks: usize, call: u32,
                                     out: *mut pfgpu_fs_obs, k: *mut usize) -> c_int;
    pub fn pfgpu_fs_last_gate(h: *mut pfgpu_fs, did_resample: *mut c_int) -> c_int;
    pub fn pfgpu_fs_set_variant(h: *mut pfgpu_fs, variant: c_int) -> c_int;
    pub fn pfgpu_fs_last_neff(h: *mut pfgpu_fs, neff: *mut f64) -> c_int;
    pub fn pfgpu_fs_particle_landmarks(h: *mut pfgpu_fs, index_local: usize, lm6: *mut f64) -> c_int;
    pub fn pfgpu_fs_count(h: *mut pfgpu_fs, n_local: *mut usize, n_global: *mut usize, n_landmarks: *mut usize) -> c_int;
    pub fn pfgpu_fs_sync(h: *mut pfgpu_fs) -> c_int;
    pub fn pfgpu_pf_sync(h: *mut pfgpu_pf) -> c_int;
    // multi-GPU: one process per GPU; rank 0 makes the id, the host program broadcasts its 128 bytes, every rank creates
    // its shard with the GLOBAL particle count (INTEGRATION.md "Multi-GPU")
    pub fn pfgpu_device_count(count: *mut c_int) -> c_int;
    pub fn pfgpu_nccl_unique_id(out128: *mut c_void) -> c_int;
    pub fn pfgpu_fs_create_sharded(cfg: *const pfgpu_fs_config, n_particles_global: usize, n_landmarks: usize, seed: u64,
                                   device: c_int, nccl_unique_id: *const c_void, rank: c_int, world: c_int,
                                   out: *mut *mut pfgpu_fs) -> c_int;
    /// all ranks inside one process (devices may repeat)
    pub fn pfgpu_fs_create_sharded_local(cfg: *const pfgpu_fs_config, n_particles_global: usize, n_landmarks: usize, seed: u64,
                                         devices: *const c_int, world: c_int, out: *mut *mut pfgpu_fs) -> c_int;
    /// 0 = one GPU, 2 = sharded over peer memory (NVLink)
    pub fn pfgpu_fs_shard_mode(h: *mut pfgpu_fs, mode: *mut c_int) -> c_int;
}
