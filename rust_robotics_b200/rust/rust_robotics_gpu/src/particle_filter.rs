//! ParticleFilterLocalizer over the GPU engine — mirrors crates/rust_robotics_localization/src/particle_filter.rs.
use nalgebra::{DMatrix, Matrix4, Vector2, Vector4};
use pfgpu_sys as sys;
use rust_robotics_core::{ControlInput, RoboticsError, RoboticsResult, State2D, StateEstimator};

pub type PFState = Vector4<f64>;
pub type PFControl = Vector2<f64>;
pub type PFMeasurement = Vec<(f64, f64, f64)>;

#[derive(Debug, Clone)]
pub struct Particle { pub x: f64, pub y: f64, pub yaw: f64, pub v: f64, pub w: f64 }   // pf.rs:26-32

#[derive(Debug, Clone)]
pub struct ParticleFilterConfig {                                                      // pf.rs:52-78
    pub n_particles: usize, pub resample_threshold: f64, pub range_noise: f64,
    pub velocity_noise: f64, pub yaw_rate_noise: f64, pub dt: f64,
}
impl Default for ParticleFilterConfig {
    fn default() -> Self {
        Self { n_particles: 100, resample_threshold: 0.5, range_noise: 0.2, velocity_noise: 2.0,
               yaw_rate_noise: 40.0_f64.to_radians(), dt: 0.1 }
    }
}
impl ParticleFilterConfig {
    fn to_c(&self) -> sys::pfgpu_pf_config {
        sys::pfgpu_pf_config { n_particles: self.n_particles as u64, resample_threshold: self.resample_threshold,
            range_noise: self.range_noise, velocity_noise: self.velocity_noise, yaw_rate_noise: self.yaw_rate_noise,
            dt: self.dt, mode: 0, _pad: 0, max_particles: self.n_particles as u64, kld_epsilon: 0.05, kld_z: 2.326 }
    }
    pub fn validate(&self) -> RoboticsResult<()> { status(unsafe { sys::pfgpu_pf_config_validate(&self.to_c()) }) }  // pf.rs:81-117
}

fn status(rc: i32) -> RoboticsResult<()> {
    if rc == 0 { return Ok(()); }
    let msg = unsafe { std::ffi::CStr::from_ptr(sys::pfgpu_strerror(rc)) }.to_string_lossy().into_owned();
    Err(RoboticsError::InvalidParameter(msg))            // rc > 0 (CUDA/NCCL) would map to an EstimationError variant
}

pub struct ParticleFilterLocalizer {
    h: *mut sys::pfgpu_pf,
    config: ParticleFilterConfig,
    state_estimate: PFState,                // refreshed after every phase, like pf.rs:499-503
    covariance_dyn: DMatrix<f64>,
    particles: Vec<Particle>,               // host mirror for get_particles() -> &[Particle] (pf.rs:244), refreshed lazily
    dirty: bool,
}
unsafe impl Send for ParticleFilterLocalizer {}

impl ParticleFilterLocalizer {
    pub fn try_new(config: ParticleFilterConfig) -> RoboticsResult<Self> {                       // pf.rs:139-156
        let mut h = std::ptr::null_mut();
        status(unsafe { sys::pfgpu_pf_create(&config.to_c(), 42, 0, &mut h) })?;
        let mut s = Self { h, config, state_estimate: PFState::zeros(), covariance_dyn: DMatrix::zeros(4, 4), particles: vec![], dirty: true };
        s.refresh_cache()?;
        Ok(s)
    }
    pub fn new(config: ParticleFilterConfig) -> Self { Self::try_new(config).expect("invalid particle filter configuration") }
    pub fn try_with_initial_state(initial_state: PFState, config: ParticleFilterConfig) -> RoboticsResult<Self> {   // pf.rs:170-199
        let mut s = Self::try_new(config)?;
        status(unsafe { sys::pfgpu_pf_init_state(s.h, initial_state.as_ptr()) })?;
        s.refresh_cache()?;
        Ok(s)
    }
    pub fn try_predict_with_control(&mut self, control: &PFControl) -> RoboticsResult<()> {      // pf.rs:255-301
        status(unsafe { sys::pfgpu_pf_predict(self.h, control.as_ptr()) })?;
        self.refresh_cache()
    }
    pub fn try_update_with_observations(&mut self, observations: &PFMeasurement) -> RoboticsResult<()> {   // pf.rs:310-334
        let flat: Vec<f64> = observations.iter().flat_map(|&(d, x, y)| [d, x, y]).collect();
        status(unsafe { sys::pfgpu_pf_update(self.h, flat.as_ptr(), observations.len()) })?;
        self.refresh_cache()
    }
    pub fn resample(&mut self) {                                                                  // pf.rs:337-345
        let mut did = 0;
        let _ = unsafe { sys::pfgpu_pf_resample(self.h, &mut did) };
        let _ = self.refresh_cache();
    }
    pub fn try_step(&mut self, control: &PFControl, observations: &PFMeasurement) -> RoboticsResult<PFState> {   // pf.rs:488-497
        let flat: Vec<f64> = observations.iter().flat_map(|&(d, x, y)| [d, x, y]).collect();
        let mut est = [0.0f64; 4];
        status(unsafe { sys::pfgpu_pf_step(self.h, control.as_ptr(), flat.as_ptr(), observations.len(), est.as_mut_ptr()) })?;
        self.refresh_cache()?;
        Ok(self.state_estimate)
    }
    pub fn step(&mut self, control: &PFControl, observations: &PFMeasurement) -> PFState {
        self.try_step(control, observations).expect("invalid particle filter step input")
    }
    pub fn try_step_state(&mut self, control: ControlInput, observations: &PFMeasurement) -> RoboticsResult<State2D> {   // pf.rs:373-380
        self.try_step(&control.to_vector(), observations)?;
        Ok(self.state_2d())
    }
    pub fn estimate(&self) -> PFState { self.state_estimate }                                      // pf.rs:348
    pub fn state_2d(&self) -> State2D { let e = self.state_estimate; State2D::new(e[0], e[1], e[2], e[3]) }
    pub fn calc_covariance(&self) -> Matrix4<f64> { Matrix4::from_fn(|i, j| self.covariance_dyn[(i, j)]) }
    pub fn set_range_noise(&mut self, s: f64) -> RoboticsResult<()> { status(unsafe { sys::pfgpu_pf_set_range_noise(self.h, s) })?; self.config.range_noise = s; Ok(()) }
    pub fn get_particles(&mut self) -> &[Particle] {                                               // pf.rs:244 (lazy D2H)
        if self.dirty {
            let (mut n, mut ng) = (0usize, 0usize);
            let _ = unsafe { sys::pfgpu_pf_count(self.h, &mut n, &mut ng) };
            let mut aos = vec![0.0f64; 5 * n];
            let _ = unsafe { sys::pfgpu_pf_download(self.h, aos.as_mut_ptr(), n) };
            self.particles = aos.chunks(5).map(|c| Particle { x: c[0], y: c[1], yaw: c[2], v: c[3], w: c[4] }).collect();
            self.dirty = false;
        }
        &self.particles
    }
    fn refresh_cache(&mut self) -> RoboticsResult<()> {
        let (mut est, mut cov) = ([0.0f64; 4], [0.0f64; 16]);
        status(unsafe { sys::pfgpu_pf_estimate(self.h, est.as_mut_ptr(), cov.as_mut_ptr()) })?;
        self.state_estimate = PFState::from_column_slice(&est);
        self.covariance_dyn = DMatrix::from_column_slice(4, 4, &cov);       // the ABI is column-major like nalgebra
        self.dirty = true;
        Ok(())
    }
}
impl Drop for ParticleFilterLocalizer { fn drop(&mut self) { unsafe { sys::pfgpu_pf_destroy(self.h) } } }

impl StateEstimator for ParticleFilterLocalizer {                                                  // pf.rs:552-573
    type State = PFState; type Measurement = PFMeasurement; type Control = PFControl;
    fn predict(&mut self, control: &Self::Control, _dt: f64) { self.try_predict_with_control(control).expect("invalid particle filter prediction input") }
    fn update(&mut self, measurement: &Self::Measurement) { self.try_update_with_observations(measurement).expect("invalid particle filter observations"); self.resample(); }
    fn get_state(&self) -> &Self::State { &self.state_estimate }
    fn get_covariance(&self) -> Option<&DMatrix<f64>> { Some(&self.covariance_dyn) }
}
