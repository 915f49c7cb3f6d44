/*
 * oracle.h — CPU ORACLE for the particle-filter / FastSLAM 1.0 hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This is a statement-by-statement C restatement of the reference's Rust loops:
 *   crates/rust_robotics_localization/src/particle_filter.rs          (pf.rs)
 *   crates/rust_robotics_localization/src/monte_carlo_localization.rs (mcl.rs)
 *   crates/rust_robotics_slam/src/fastslam1.rs                        (fs1.rs)
 * in the reference's AoS layout and in the reference's exact operation order.
 *
 * PARITY UNPINNED: the reference cannot be compiled or run here (no rustc/cargo; nalgebra/rand/rand_distr
 * are not vendored) and none of its own tests pins a numeric value on this path, so no
 * reference-generated golden vector exists.  The oracle is instead pinned by (a) hand-computed known-answer
 * vectors, (b) an independent pure-Python restatement using glibc libm (tests/golden/make_golden.py) and
 * (c) the value-level facts the reference's tests do assert (sum w = 1, init constants, count bounds).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this
 * library.  The product (rust_robotics_b200/) never does.
 *
 * Two builds of the same source:
 *   liboracle.so       math = include/pf_contract_math.h   (bit-identical to the CUDA kernels)
 *   liboracle_libm.so  math = glibc sin/cos/exp/atan2      (what Rust's f64 methods call; -DPF_ORACLE_LIBM)
 * Random draws always come from the Philox contract (the reference is unseeded; see pf_contract_math.h).
 */
#ifndef PF_ORACLE_H
#define PF_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------- PF / MCL --------------------------------------------------------- */
typedef struct {            /* pf.rs:26-32 / mcl.rs:29-35 */
    double x, y, yaw, v, w;
} orc_particle;

typedef struct {            /* pf.rs:52-65 (+ mcl.rs:50-59 when mode == 1) */
    uint64_t n_particles;       /* PF: n_particles; MCL: min_particles */
    double resample_threshold;  /* PF only */
    double range_noise, velocity_noise, yaw_rate_noise, dt;
    int32_t mode;               /* 0 = ParticleFilterLocalizer, 1 = MonteCarloLocalizer */
    int32_t _pad;
    uint64_t max_particles;     /* MCL only */
    double kld_epsilon, kld_z;  /* MCL only */
} orc_pf_config;

typedef struct orc_pf orc_pf;

/* 0 ok, -1 InvalidParameter (pf.rs:81-117 / mcl.rs:87-130) */
int  orc_pf_config_validate(const orc_pf_config* cfg);
orc_pf* orc_pf_new(const orc_pf_config* cfg, uint64_t seed);                 /* try_new: pf.rs:139-156 */
int  orc_pf_init_state(orc_pf*, const double init[4]);                       /* try_with_initial_state: pf.rs:170-199, mcl.rs:176-206 */
void orc_pf_free(orc_pf*);
size_t orc_pf_count(const orc_pf*);                                          /* mcl.rs:318-320 */
void orc_pf_set_particles(orc_pf*, const double* aos5, size_t n);
void orc_pf_get_particles(const orc_pf*, double* aos5);                      /* pf.rs:244-246 */
int  orc_pf_predict(orc_pf*, const double u[2]);                             /* pf.rs:255-301 */
int  orc_pf_predict_with_noise(orc_pf*, const double u[2], const double* zv, const double* zw);
int  orc_pf_update(orc_pf*, const double* obs3, size_t k);                   /* pf.rs:310-334 */
int  orc_pf_resample(orc_pf*);                                               /* pf.rs:337-345 / mcl.rs:322-365; returns 1 if it resampled */
int  orc_pf_resample_with_uniforms(orc_pf*, const double* r, size_t nr);    /* injected draws, forced (no N_eff gate) */
int  orc_pf_step(orc_pf*, const double u[2], const double* obs3, size_t k, double est[4]); /* pf.rs:488-497 */
void orc_pf_estimate(const orc_pf*, double est[4], double cov16_colmajor[16]); /* pf.rs:348-365 */
double orc_pf_neff(const orc_pf*);                                           /* pf.rs:416-423 */
int  orc_pf_set_range_noise(orc_pf*, double);                                /* pf.rs:228-236 */
size_t orc_pf_last_indices(const orc_pf*, uint32_t* idx, size_t cap);        /* parity hook */
void orc_pf_set_fast_search(orc_pf*, int on);  /* 0: as-written linear scan (O(N^2)); 1: lower_bound, identical indices */
void orc_pf_set_threads(orc_pf*, int nthreads); /* OpenMP over particles for predict/update (baseline timing) */

/* ------------------------------- FastSLAM 1.0 ------------------------------------------------------ */
typedef struct {            /* fs1.rs:13-23: module constants become fields, reference values as defaults */
    double dt;              /* DT = 0.1 */
    double max_range;       /* MAX_RANGE = 20.0 (simulator only) */
    double nth;             /* NTH = 100/1.5 */
    double q00, q11;        /* Q_SIM diag = 0.3, 0.0305 */
    double r00, r11;        /* R_SIM diag = 0.5, 0.0305 */
    double init_weight;     /* 1/N_PARTICLE = 0.01 (fs1.rs:56) */
} orc_fs_config;

typedef struct { double d, angle; uint64_t lm_id; } orc_fs_obs;   /* fs1.rs:240 (f64, f64, usize) */

typedef struct orc_fs orc_fs;

void orc_fs_default_config(orc_fs_config*);
orc_fs* orc_fs_new(const orc_fs_config*, size_t n_particles, size_t n_landmarks, uint64_t seed); /* create_particles fs1.rs:302-306 */
void orc_fs_free(orc_fs*);
/* pose_w: n x 4 (weight, x, y, yaw) AoS; lm: n x m x 6 (x, y, c00, c01, c10, c11) AoS, particle-major */
void orc_fs_set_state(orc_fs*, const double* pose_w, const double* lm);
void orc_fs_get_state(const orc_fs*, double* pose_w, double* lm);
/* twin of pfgpu_fs_seed_map (include/pfgpu.h): initialised map so the EKF branch is live from step 0 */
void orc_fs_seed_map(orc_fs*, const double pose3[3], const double* landmarks_xy, double sigma, double cov0);
int  orc_fs_step(orc_fs*, const double u[2], const orc_fs_obs* z, size_t k);  /* fastslam_update fs1.rs:237-266; returns 1 if resampled */
int  orc_fs_step_with_noise(orc_fs*, const double u[2], const orc_fs_obs* z, size_t k,
                            const double* z0, const double* z1, double r_uniform01);
size_t orc_fs_best(const orc_fs*);                                              /* get_best_particle fs1.rs:269-274 */
size_t orc_fs_last_indices(const orc_fs*, uint32_t* idx, size_t cap);
double orc_fs_last_neff(const orc_fs*);
/* get_observations fs1.rs:277-299; noise from Philox stream OBS, call = step; returns count */
size_t orc_fs_get_observations(const orc_fs_config*, const double x_true[3], const double* landmarks_xy,
                               size_t n_landmarks, uint64_t seed, uint32_t call, orc_fs_obs* out);
void orc_fs_set_threads(orc_fs*, int nthreads);
/* 1 = FastSLAM 1.0 (default), 2 = FastSLAM 2.0 (crates/rust_robotics_slam/src/fastslam2.rs, "fs2.rs"): same particle set,
 * same normalise / N_eff / resample; the pose is sampled from the observation-informed proposal of the FIRST observation
 * (fs2.rs:173-239) and update_landmark_and_weight (fs2.rs:242-280) replaces update_landmark.  With injected noise
 * (orc_fs_step_with_noise) z1 must then hold 2n values: z1[i] and, as the third draw of particle i, z1[n + i]. */
void orc_fs_set_variant(orc_fs*, int variant);
/* compute_proposal fs2.rs:173-216 for one particle (test probe): pose3 = (x, y, yaw), lm6 = (x, y, c00, c01, c10, c11) */
void orc_fs2_compute_proposal(const orc_fs_config*, const double pose3[3], const double u[2], double z_d, double z_angle,
                              const double lm6[6], double mean3[3], double cov9_rowmajor[9]);
/* sample_pose fs2.rs:219-239 (test probe): mean + L * (n0, n1, n2), L = Cholesky factor or the diagonal fallback */
void orc_fs2_sample_pose(const double mean3[3], const double cov9_rowmajor[9], const double n3[3], double out3[3]);

/* ------------------------------- contract-math probes (for tests) ---------------------------------- */
void orc_math_exp(const double* in, double* out, size_t n);
void orc_math_log(const double* in, double* out, size_t n);
void orc_math_sin(const double* in, double* out, size_t n);
void orc_math_cos(const double* in, double* out, size_t n);
void orc_math_atan2(const double* y, const double* x, double* out, size_t n);
void orc_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4]);
void orc_normal_pair(uint64_t seed, uint32_t stream, uint32_t call, uint64_t index, double z[2]);
double orc_uniform53(uint64_t seed, uint32_t stream, uint32_t call, uint64_t index);
double orc_uniform52(uint64_t seed, uint32_t stream, uint32_t call, uint64_t index);
int  orc_math_mode(void);   /* 0 = contract, 1 = glibc */

#ifdef __cplusplus
}
#endif
#endif
