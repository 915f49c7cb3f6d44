//! MonteCarloLocalizer over the GPU engine — mirrors crates/rust_robotics_localization/src/monte_carlo_localization.rs.
//! The particle count follows the KLD bound between `min_particles` and `max_particles` (mcl.rs:322-378) exactly like the
//! reference; `particle_count()` therefore changes from step to step.
use nalgebra::{DMatrix, Vector2, Vector4};
use pfgpu_sys as sys;
use rust_robotics_core::{RoboticsError, RoboticsResult, State2D, StateEstimator};

use crate::particle_filter::{PFControl, PFMeasurement, PFState, Particle};

#[derive(Debug, Clone)]
pub struct MonteCarloLocalizationConfig {                                                   // mcl.rs:50-74
    pub min_particles: usize, pub max_particles: usize, pub kld_epsilon: f64, pub kld_z: f64,
    pub range_noise: f64, pub velocity_noise: f64, pub yaw_rate_noise: f64, pub dt: f64,
}
impl Default for MonteCarloLocalizationConfig {
    fn default() -> Self {
        Self { min_particles: 100, max_particles: 5000, kld_epsilon: 0.05, kld_z: 2.326, range_noise: 0.2,
               velocity_noise: 2.0, yaw_rate_noise: 40.0_f64.to_radians(), dt: 0.1 }
    }
}
impl MonteCarloLocalizationConfig {
    fn to_c(&self) -> sys::pfgpu_pf_config {
        sys::pfgpu_pf_config { n_particles: self.min_particles as u64, resample_threshold: 0.0, range_noise: self.range_noise,
            velocity_noise: self.velocity_noise, yaw_rate_noise: self.yaw_rate_noise, dt: self.dt, mode: 1, _pad: 0,
            max_particles: self.max_particles as u64, kld_epsilon: self.kld_epsilon, kld_z: self.kld_z }
    }
    pub fn validate(&self) -> RoboticsResult<()> { status(unsafe { sys::pfgpu_pf_config_validate(&self.to_c()) }) }   // mcl.rs:87-130
}

fn status(rc: i32) -> RoboticsResult<()> {
    if rc == 0 { return Ok(()); }
    let msg = unsafe { std::ffi::CStr::from_ptr(sys::pfgpu_strerror(rc)) }.to_string_lossy().into_owned();
    Err(RoboticsError::InvalidParameter(msg))
}

pub struct MonteCarloLocalizer {
    h: *mut sys::pfgpu_pf,
    state_estimate: PFState,
    covariance_dyn: DMatrix<f64>,
    particles: Vec<Particle>,
    dirty: bool,
}
unsafe impl Send for MonteCarloLocalizer {}

impl MonteCarloLocalizer {
    pub fn try_new(config: MonteCarloLocalizationConfig) -> RoboticsResult<Self> {                  // mcl.rs:150-164
        // the reference draws from rand::rng() (mcl.rs:187,216,338): entropy, not a constant, so that independent localizers
        // (Monte-Carlo trials) are not correlated; try_new_seeded pins the Philox seed for tests
        let t = std::time::SystemTime::now().duration_since(std::time::UNIX_EPOCH).map(|d| d.as_nanos() as u64).unwrap_or(0);
        Self::try_new_seeded(config, t.wrapping_mul(0x9E37_79B9_7F4A_7C15), 0)
    }
    pub fn try_new_seeded(config: MonteCarloLocalizationConfig, seed: u64, device: i32) -> RoboticsResult<Self> {
        let mut h = std::ptr::null_mut();
        status(unsafe { sys::pfgpu_pf_create(&config.to_c(), seed, device, &mut h) })?;
        let mut s = Self { h, state_estimate: PFState::zeros(), covariance_dyn: DMatrix::zeros(4, 4), particles: vec![], dirty: true };
        s.refresh_cache()?;
        Ok(s)
    }
    pub fn new(config: MonteCarloLocalizationConfig) -> Self { Self::try_new(config).expect("invalid MCL configuration") }
    pub fn try_with_initial_state(initial_state: PFState, config: MonteCarloLocalizationConfig) -> RoboticsResult<Self> {   // mcl.rs:176-206
        let mut s = Self::try_new(config)?;
        status(unsafe { sys::pfgpu_pf_init_state(s.h, initial_state.as_ptr()) })?;
        s.refresh_cache()?;
        Ok(s)
    }
    pub fn with_initial_state(initial_state: PFState, config: MonteCarloLocalizationConfig) -> Self {
        Self::try_with_initial_state(initial_state, config).expect("invalid MCL initial state")
    }
    pub fn try_predict_with_control(&mut self, control: &PFControl) -> RoboticsResult<()> {          // mcl.rs:209-257
        status(unsafe { sys::pfgpu_pf_predict(self.h, control.as_ptr()) })?;
        self.refresh_cache()
    }
    pub fn try_update_with_observations(&mut self, observations: &PFMeasurement) -> RoboticsResult<()> {   // mcl.rs:260-288
        let flat: Vec<f64> = observations.iter().flat_map(|&(d, x, y)| [d, x, y]).collect();
        status(unsafe { sys::pfgpu_pf_update(self.h, flat.as_ptr(), observations.len()) })?;
        self.refresh_cache()
    }
    pub fn try_step(&mut self, control: &PFControl, observations: &PFMeasurement) -> RoboticsResult<PFState> {   // mcl.rs:291-300
        let flat: Vec<f64> = observations.iter().flat_map(|&(d, x, y)| [d, x, y]).collect();
        let mut est = [0.0f64; 4];
        status(unsafe { sys::pfgpu_pf_step(self.h, control.as_ptr(), flat.as_ptr(), observations.len(), est.as_mut_ptr()) })?;
        self.refresh_cache()?;
        Ok(self.state_estimate)
    }
    pub fn estimate(&self) -> PFState { self.state_estimate }                                          // mcl.rs:302-304
    pub fn state_2d(&self) -> State2D { let e = self.state_estimate; State2D::new(e[0], e[1], e[2], e[3]) }
    pub fn particle_count(&self) -> usize {                                                            // mcl.rs:318-320
        let (mut nl, mut ng) = (0usize, 0usize);
        let _ = unsafe { sys::pfgpu_pf_count(self.h, &mut nl, &mut ng) };
        ng
    }
    pub fn get_particles(&mut self) -> &[Particle] {
        if self.dirty {
            let n = self.particle_count();
            let mut aos = vec![0.0f64; 5 * n];
            let _ = unsafe { sys::pfgpu_pf_download(self.h, aos.as_mut_ptr(), n) };
            self.particles = aos.chunks(5).map(|c| Particle { x: c[0], y: c[1], yaw: c[2], v: c[3], w: c[4] }).collect();
            self.dirty = false;
        }
        &self.particles
    }
    fn resample(&mut self) {                                                                           // resample_adaptive mcl.rs:322-365
        let mut did = 0;
        let _ = unsafe { sys::pfgpu_pf_resample(self.h, &mut did) };
        let _ = self.refresh_cache();
    }
    fn refresh_cache(&mut self) -> RoboticsResult<()> {                                                // mcl.rs:413-447
        let (mut est, mut cov) = ([0.0f64; 4], [0.0f64; 16]);
        status(unsafe { sys::pfgpu_pf_estimate(self.h, est.as_mut_ptr(), cov.as_mut_ptr()) })?;
        self.state_estimate = PFState::from_column_slice(&est);
        self.covariance_dyn = DMatrix::from_column_slice(4, 4, &cov);
        self.dirty = true;
        Ok(())
    }
}
impl Drop for MonteCarloLocalizer { fn drop(&mut self) { unsafe { sys::pfgpu_pf_destroy(self.h) } } }

impl StateEstimator for MonteCarloLocalizer {                                                          // mcl.rs:450-471 (errors are swallowed there too)
    type State = Vector4<f64>; type Measurement = PFMeasurement; type Control = Vector2<f64>;
    fn predict(&mut self, control: &Self::Control, _dt: f64) { let _ = self.try_predict_with_control(control); }
    fn update(&mut self, measurement: &Self::Measurement) { let _ = self.try_update_with_observations(measurement); self.resample(); }
    fn get_state(&self) -> &Self::State { &self.state_estimate }
    fn get_covariance(&self) -> Option<&DMatrix<f64>> { Some(&self.covariance_dyn) }
}
