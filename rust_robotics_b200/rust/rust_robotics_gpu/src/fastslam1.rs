//! FastSLAM 1.0 over the GPU engine — mirrors crates/rust_robotics_slam/src/fastslam1.rs.
//! The reference's functions mutate a caller-owned `Vec<Particle>`; here the particles live on the device inside
//! `FastSlam` and `download()` materialises the reference's `Vec<Particle>` when a caller wants it.
use nalgebra::{Matrix2, Vector2};
use pfgpu_sys as sys;

#[derive(Clone)]
pub struct Landmark { pub x: f64, pub y: f64, pub cov: Matrix2<f64> }                  // fs1.rs:27-31
#[derive(Clone)]
pub struct Particle { pub weight: f64, pub x: f64, pub y: f64, pub yaw: f64, pub landmarks: Vec<Landmark> }   // fs1.rs:45-51

pub struct FastSlam { h: *mut sys::pfgpu_fs, n: usize, m: usize }
unsafe impl Send for FastSlam {}

/// create_particles fs1.rs:302-306
pub fn create_particles(n_particles: usize, n_landmarks: usize) -> FastSlam {
    let mut cfg = std::mem::MaybeUninit::<sys::pfgpu_fs_config>::uninit();
    unsafe { sys::pfgpu_fs_default_config(cfg.as_mut_ptr()) };
    let mut h = std::ptr::null_mut();
    let rc = unsafe { sys::pfgpu_fs_create(cfg.as_ptr(), n_particles, n_landmarks, 42, 0, &mut h) };
    assert_eq!(rc, 0, "pfgpu_fs_create failed");
    FastSlam { h, n: n_particles, m: n_landmarks }
}
/// fastslam_update fs1.rs:237-266
pub fn fastslam_update(particles: &mut FastSlam, u: Vector2<f64>, z: &[(f64, f64, usize)]) {
    let obs: Vec<sys::pfgpu_fs_obs> = z.iter().map(|&(d, angle, id)| sys::pfgpu_fs_obs { d, angle, lm_id: id as u64 }).collect();
    let rc = unsafe { sys::pfgpu_fs_step(particles.h, u.as_ptr(), obs.as_ptr(), obs.len(), std::ptr::null_mut()) };
    assert_eq!(rc, 0, "pfgpu_fs_step failed");       // lm_id out of range panics in the reference too (Vec index, fs1.rs:141)
}
/// get_best_particle fs1.rs:269-274
pub fn get_best_particle(particles: &FastSlam) -> Particle {
    let (mut idx, mut pw) = (0usize, [0.0f64; 4]);
    unsafe { sys::pfgpu_fs_best(particles.h, &mut idx, pw.as_mut_ptr()) };
    let mut lm = vec![0.0f64; 6 * particles.m];
    unsafe { sys::pfgpu_fs_particle_landmarks(particles.h, idx, lm.as_mut_ptr()) };
    Particle { weight: pw[0], x: pw[1], y: pw[2], yaw: pw[3],
               landmarks: lm.chunks(6).map(|q| Landmark { x: q[0], y: q[1], cov: Matrix2::new(q[2], q[3], q[4], q[5]) }).collect() }
}
impl FastSlam {
    pub fn len(&self) -> usize { self.n }
    /// the reference's Vec<Particle>, materialised (checkpoint / API-compat)
    pub fn download(&self) -> Vec<Particle> {
        let (mut pw, mut lm) = (vec![0.0f64; 4 * self.n], vec![0.0f64; 6 * self.n * self.m]);
        unsafe { sys::pfgpu_fs_download(self.h, pw.as_mut_ptr(), lm.as_mut_ptr(), self.n) };
        (0..self.n).map(|i| Particle { weight: pw[4 * i], x: pw[4 * i + 1], y: pw[4 * i + 2], yaw: pw[4 * i + 3],
            landmarks: lm[6 * self.m * i..6 * self.m * (i + 1)].chunks(6)
                .map(|q| Landmark { x: q[0], y: q[1], cov: Matrix2::new(q[2], q[3], q[4], q[5]) }).collect() }).collect()
    }
}
impl Drop for FastSlam { fn drop(&mut self) { unsafe { sys::pfgpu_fs_destroy(self.h) } } }
