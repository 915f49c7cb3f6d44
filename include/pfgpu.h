/*
 * pfgpu.h — C ABI of the B200-native particle-filter / FastSLAM 1.0 engine (libpfgpu.so).
 *
 * This is the drop-in boundary.  The reference (rsasaki0109/rust_robotics) has no FFI of its own: its
 * boundary is the public Rust API of rust_robotics_localization::{ParticleFilterLocalizer,
 * MonteCarloLocalizer} and rust_robotics_slam::fastslam1.  A Rust shim (rust_robotics_b200/rust/, shown in
 * INTEGRATION.md) keeps those type and method names and forwards each method body to ONE entry point below;
 * the C++ mirror (rust_robotics_b200/host/ headers) and the Python mirror (rust_robotics_b200/api.py) do the same.
 * Each entry point cites the reference item it replaces ("pf.rs" = crates/rust_robotics_localization/src/
 * particle_filter.rs, "mcl.rs" = .../monte_carlo_localization.rs, "fs1.rs" = crates/rust_robotics_slam/src/
 * fastslam1.rs).
 *
 * Conventions
 *   - plain pointers and sizes only; all floating point is IEEE f64; matrices are column-major (nalgebra).
 *   - every call returns a status: 0 ok; <0 invalid parameter (maps to RoboticsError::InvalidParameter,
 *     crates/rust_robotics_core/src/error.rs:8-24); >0 CUDA / NCCL runtime failure.
 *   - one handle = one CUDA device + one stream; a handle is Send, not Sync (the reference API is &mut self).
 *   - there is NO CPU fallback: without a usable CUDA device create() fails with PFGPU_ERR_NO_DEVICE.
 *   - random draws follow the Philox contract of include/pf_contract_math.h (the reference is unseeded).
 */
#ifndef PFGPU_H
#define PFGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PFGPU_OK                 0
#define PFGPU_ERR_INVALID      (-1)   /* RoboticsError::InvalidParameter                       */
#define PFGPU_ERR_UNSUPPORTED  (-2)   /* valid in the reference, not built here (KLD-adaptive MCL on more than one GPU, > 1024 landmarks, ...) */
#define PFGPU_ERR_NO_DEVICE     1000  /* no CUDA device / extension cannot run: fail loudly      */
#define PFGPU_ERR_CUDA          1001
#define PFGPU_ERR_NCCL          1002

const char* pfgpu_strerror(int status);
/* last CUDA/NCCL error text of the calling thread (empty if none) */
const char* pfgpu_last_error(void);
int pfgpu_device_count(int* count);

/* ============================== ParticleFilterLocalizer / MonteCarloLocalizer ======================= */

/* ParticleFilterConfig (pf.rs:52-65) and MonteCarloLocalizationConfig (mcl.rs:50-59) */
typedef struct {
    uint64_t n_particles;        /* pf: n_particles (100); mcl: min_particles (100)        */
    double   resample_threshold; /* pf only (0.5)                                          */
    double   range_noise;        /* 0.2                                                    */
    double   velocity_noise;     /* 2.0                                                    */
    double   yaw_rate_noise;     /* 40 deg                                                 */
    double   dt;                 /* 0.1                                                    */
    int32_t  mode;               /* 0 = ParticleFilterLocalizer, 1 = MonteCarloLocalizer   */
    int32_t  _pad;
    uint64_t max_particles;      /* mcl only (5000); > n_particles: KLD-adaptive particle count (mcl.rs:322-365) */
    double   kld_epsilon;        /* mcl only (0.05)                                        */
    double   kld_z;              /* mcl only (2.326)                                       */
} pfgpu_pf_config;

typedef struct pfgpu_pf pfgpu_pf;

void pfgpu_pf_default_config(pfgpu_pf_config* cfg, int mode);      /* Default impls pf.rs:67-78, mcl.rs:61-74 */
int  pfgpu_pf_config_validate(const pfgpu_pf_config* cfg);         /* validate(): pf.rs:81-117, mcl.rs:87-130 */

/* try_new (pf.rs:139-156, mcl.rs:150-164): n particles at the origin, w = 1/n */
int  pfgpu_pf_create(const pfgpu_pf_config* cfg, uint64_t seed, int device, pfgpu_pf** out);
/* Sharded over `world` GPUs, one process per GPU.  `nccl_unique_id` = the 128 bytes of an
 * ncclUniqueId produced by pfgpu_nccl_unique_id() on rank 0 and broadcast by the host program.  cfg->n_particles
 * is the GLOBAL particle count and must divide evenly. */
int  pfgpu_pf_create_sharded(const pfgpu_pf_config* cfg, uint64_t seed, int device,
                             const void* nccl_unique_id, int rank, int world, pfgpu_pf** out);
void pfgpu_pf_destroy(pfgpu_pf*);
/* try_with_initial_state (pf.rs:170-199, mcl.rs:176-206): init + uniform jitter, on device */
int  pfgpu_pf_init_state(pfgpu_pf*, const double init[4]);
/* get_particles() (pf.rs:244-246) / checkpoint: AoS (x, y, yaw, v, w), local shard */
int  pfgpu_pf_upload(pfgpu_pf*, const double* aos5, size_t n);
int  pfgpu_pf_download(pfgpu_pf*, double* aos5, size_t n);
int  pfgpu_pf_count(pfgpu_pf*, size_t* n_local, size_t* n_global);          /* particle_count() mcl.rs:318-320 */
int  pfgpu_pf_predict(pfgpu_pf*, const double u[2]);                         /* try_predict_with_control pf.rs:255-301, mcl.rs:209-257 */
int  pfgpu_pf_update(pfgpu_pf*, const double* obs3, size_t k);               /* try_update_with_observations pf.rs:310-334, mcl.rs:260-288; obs3 = k x (d, lx, ly) */
int  pfgpu_pf_resample(pfgpu_pf*, int* did_resample);                        /* resample() pf.rs:337-345; resample_adaptive mcl.rs:322-365 */
/* try_step (pf.rs:488-497, mcl.rs:291-300).  est (nullable): if non-NULL the call synchronises and returns
 * estimate(); if NULL the step is only enqueued on the handle's stream. */
int  pfgpu_pf_step(pfgpu_pf*, const double u[2], const double* obs3, size_t k, double est[4]);
int  pfgpu_pf_estimate(pfgpu_pf*, double est[4], double cov16_colmajor[16]); /* estimate()/calc_covariance() pf.rs:348-365 */
int  pfgpu_pf_neff(pfgpu_pf*, double* neff);                                 /* calc_n_eff pf.rs:416-423 */
int  pfgpu_pf_set_range_noise(pfgpu_pf*, double range_noise);                /* pf.rs:228-236 */
/* parity hook: ancestry of the last step's resample (global indices); *n = 0 when the last step did not resample */
int  pfgpu_pf_last_indices(pfgpu_pf*, uint32_t* idx, size_t cap, size_t* n);
int  pfgpu_pf_sync(pfgpu_pf*);

/* ============================================ FastSLAM 1.0 ========================================== */

/* Module constants of fs1.rs:13-23 as fields; pfgpu_fs_default_config() fills in the reference values. */
typedef struct {
    double dt;           /* DT = 0.1                 */
    double max_range;    /* MAX_RANGE = 20 (only get_observations uses it) */
    double nth;          /* NTH = 100/1.5            */
    double q00, q11;     /* Q_SIM = diag(0.3, 0.0305) */
    double r00, r11;     /* R_SIM = diag(0.5, 0.0305) */
    double init_weight;  /* 1/N_PARTICLE = 0.01 (fs1.rs:56) */
} pfgpu_fs_config;

/* (distance, angle, landmark_id): the tuple of fs1.rs:240 */
typedef struct { double d, angle; uint64_t lm_id; } pfgpu_fs_obs;

typedef struct pfgpu_fs pfgpu_fs;

void pfgpu_fs_default_config(pfgpu_fs_config* cfg);
/* create_particles(n, m) fs1.rs:302-306 */
int  pfgpu_fs_create(const pfgpu_fs_config* cfg, size_t n_particles, size_t n_landmarks, uint64_t seed,
                     int device, pfgpu_fs** out);
/* Sharded over `world` (<= 8) GPUs of one NVLink domain, one process per GPU; collective over all ranks (same arguments
 * everywhere; n_particles_global / world must be a multiple of 64).  NCCL is used once, here, to exchange the cudaIpc handles of
 * the per-rank arenas.  The step itself runs over peer memory: every rank's EKF kernel pushes its 8 bytes per particle of
 * unnormalised weight into every rank's copy, every rank's post kernel evaluates the exact sums / CDF / resample indices
 * of ALL particles, and ancestors that live on another rank are read through NVLink when (and only when) one of their
 * landmarks is next observed — no NCCL call, no host synchronisation and no map copy per step.  Every rank must issue the
 * same sequence of pfgpu_fs_step calls; a rank that stops surfaces as PFGPU_ERR_CUDA at the others' next synchronising
 * call (spins time out; no hang). */
int  pfgpu_fs_create_sharded(const pfgpu_fs_config* cfg, size_t n_particles_global, size_t n_landmarks,
                             uint64_t seed, int device, const void* nccl_unique_id, int rank, int world,
                             pfgpu_fs** out);
/* The same sharded engine with all `world` ranks inside ONE process (no NCCL): out[r] runs on devices[r]; devices may
 * repeat (several ranks on one GPU — what the single-GPU parity tests use to exercise every cross-rank path).  One host
 * thread drives the ranks: issue each step to every handle with did_resample == NULL before synchronising any of them. */
int  pfgpu_fs_create_sharded_local(const pfgpu_fs_config* cfg, size_t n_particles_global, size_t n_landmarks,
                                   uint64_t seed, const int* devices, int world, pfgpu_fs** out);
void pfgpu_fs_destroy(pfgpu_fs*);
/* Vec<Particle> <-> device (fs1.rs:44-51).  pose_w: n x (weight, x, y, yaw); lm (nullable): n x m x
 * (x, y, c00, c01, c10, c11), particle-major AoS exactly like the reference's memory order. */
int  pfgpu_fs_upload(pfgpu_fs*, const double* pose_w, const double* lm, size_t n);
int  pfgpu_fs_download(pfgpu_fs*, double* pose_w, double* lm, size_t n);
/* Benchmark / test convenience (not in the reference): start from an INITIALISED map so that the EKF branch
 * fs1.rs:151-182 is live from step 0 (a fresh create_particles never reaches it: SURVEY.md App. B.3).
 * Every particle gets pose (x, y, yaw), weight 1/n, and for every landmark l: position = landmarks_xy[l] +
 * sigma * N(0,1)^2 (Philox stream INIT_A, index = particle*m + l), cov = cov0 * I (fs2.rs:254 convention). */
int  pfgpu_fs_seed_map(pfgpu_fs*, const double pose3[3], const double* landmarks_xy, size_t m, double sigma, double cov0);
/* fastslam_update fs1.rs:237-266.  did_resample (nullable): non-NULL synchronises. */
int  pfgpu_fs_step(pfgpu_fs*, const double u[2], const pfgpu_fs_obs* z, size_t k, int* did_resample);
/* get_observations fs1.rs:277-299, the simulator next to the filter, on the device: landmarks within cfg.max_range of x_true
 * (x, y, yaw), in landmark order; range / bearing noise N(0,1) * sqrt(R) from Philox stream PFC_STREAM_OBS keyed by (seed of the
 * handle, call, landmark id).  out has room for n_landmarks tuples; *k receives their number. */
int  pfgpu_fs_get_observations(pfgpu_fs*, const double x_true[3], const double* landmarks_xy, size_t n_landmarks, uint32_t call,
                               pfgpu_fs_obs* out, size_t* k);
/* get_best_particle fs1.rs:269-274 (last maximum wins); pose_w4 = (weight, x, y, yaw) */
int  pfgpu_fs_best(pfgpu_fs*, size_t* index_global, double pose_w4[4]);
/* landmarks of one particle (what render_gif_slam.rs:183-191 reads): lm6 = m x 6 */
int  pfgpu_fs_particle_landmarks(pfgpu_fs*, size_t index_local, double* lm6);
int  pfgpu_fs_last_indices(pfgpu_fs*, uint32_t* idx, size_t cap, size_t* n);  /* *n = 0 when the last step did not resample */
int  pfgpu_fs_last_neff(pfgpu_fs*, double* neff);
int  pfgpu_fs_last_gate(pfgpu_fs*, int* did_resample);     /* whether the last step resampled (fs1.rs:263); synchronises */
/* Which update pfgpu_fs_step runs: 1 = fastslam1::fastslam_update (fs1.rs:237-266, the default), 2 = fastslam2::fastslam2_update
 * (crates/rust_robotics_slam/src/fastslam2.rs:376-383 -> :330-374): the same particle set, normalisation, N_eff gate and
 * resampler; the pose of every particle is sampled from the proposal that fuses the motion prior (MOTION_COV, fs2.rs:31) with
 * the step's first observation (compute_proposal :173-216, sample_pose :219-239; three N(0,1) per particle), and
 * update_landmark_and_weight (:242-280) replaces update_landmark (landmark test `cov00 < 100`, birth with cov = 10 I, weight
 * factor 1e-10 when det S <= 0).  A step without observations is the motion model with two draws (:347-356).  May be changed
 * between steps; every rank of a sharded engine must make the same call. */
int  pfgpu_fs_set_variant(pfgpu_fs*, int variant);
int  pfgpu_fs_count(pfgpu_fs*, size_t* n_local, size_t* n_global, size_t* n_landmarks);
int  pfgpu_fs_sync(pfgpu_fs*);

/* ============================================ plumbing ============================================== */
int  pfgpu_nccl_unique_id(void* out128);     /* ncclGetUniqueId; 128 bytes */

/* Counters for benches / tests.  kernel_launches = kernels of this library launched by the handle;
 * serial_fallbacks = times an exact-sum pipeline fell back to its single-thread path (should be 0). */
typedef struct {
    uint64_t kernel_launches;
    uint64_t steps;
    uint64_t resamples;
    uint64_t serial_fallbacks;
    uint64_t xsum_dirty_last;    /* dirty elements in the most recent exact scan */
    double   main_kernel_ms_sum; /* sum of CUDA-event times of the dominant kernel when timing is on */
    uint64_t main_kernel_count;
    uint64_t compactions;        /* sharded FastSLAM: guest-column compactions so far */
    uint64_t imported_particles; /* sharded FastSLAM: particles whose map came from another rank so far */
} pfgpu_stats;
int  pfgpu_pf_stats(pfgpu_pf*, pfgpu_stats*);
int  pfgpu_fs_stats(pfgpu_fs*, pfgpu_stats*);
/* debug (PFGPU_POST_TRACE=1): accumulated per-phase times [ns] of the fused post-step kernel; out32[31] = launches */
int  pfgpu_fs_post_trace(pfgpu_fs*, unsigned long long* out32);
/* how the coupled part of the step runs: 0 = one GPU, 2 = sharded over peer memory (NVLink loads / stores inside the kernels;
   no NCCL call and no host sync per step) */
int  pfgpu_fs_shard_mode(pfgpu_fs*, int* mode);
int  pfgpu_pf_time_main_kernel(pfgpu_pf*, int on);
int  pfgpu_fs_time_main_kernel(pfgpu_fs*, int on);
/* CUDA events on the handle's own stream (bench.py times steps with these): mark(slot 0..16383) records an
 * event; elapsed(a, b) synchronises on slot b and returns the device time between the two marks. */
int  pfgpu_pf_mark(pfgpu_pf*, int slot);
int  pfgpu_pf_elapsed_ms(pfgpu_pf*, int slot_a, int slot_b, double* ms);
int  pfgpu_fs_mark(pfgpu_fs*, int slot);
int  pfgpu_fs_elapsed_ms(pfgpu_fs*, int slot_a, int slot_b, double* ms);
/* Evict the L2 cache between timed steps: overwrites a scratch buffer larger than L2 on the handle's stream. */
int  pfgpu_pf_flush_l2(pfgpu_pf*);
int  pfgpu_fs_flush_l2(pfgpu_fs*);

#ifdef __cplusplus
}
#endif
#endif /* PFGPU_H */
