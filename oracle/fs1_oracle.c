/*
 * fs1_oracle.c — CPU oracle for FastSLAM 1.0 (crates/rust_robotics_slam/src/fastslam1.rs = "fs1.rs").
 * TEST INFRASTRUCTURE ONLY — see oracle.h.  PARITY UNPINNED (no reference-generated vectors exist).
 *
 * Layout mirrors the reference: Vec<Particle{weight,x,y,yaw,landmarks: Vec<Landmark{x,y,cov 2x2}>}>
 * (fs1.rs:26-51) stored as particle-major AoS.  nalgebra 0.33 semantics used:
 *   Matrix2 * Matrix2 : entry(i,j) = a_i0*b_0j + a_i1*b_1j  (two products, one add, no fma)
 *   try_inverse (2x2) : det = m11*m22 - m21*m12; None if det == 0; else [m22/det, -m12/det; -m21/det, m11/det]
 *   determinant (2x2) : m11*m22 - m21*m12
 */
#include "fs_state.h"

void orc_fs_default_config(orc_fs_config* c) {                  /* fs1.rs:13-23 */
    c->dt = 0.1; c->max_range = 20.0; c->nth = 100.0 / 1.5;
    c->q00 = 0.3; c->q11 = 0.0305; c->r00 = 0.5; c->r11 = 0.0305;
    c->init_weight = 1.0 / 100.0;
}

orc_fs* orc_fs_new(const orc_fs_config* cfg, size_t n, size_t m, uint64_t seed) {
    orc_fs* f = (orc_fs*)calloc(1, sizeof(orc_fs));
    f->cfg = *cfg; f->n = n; f->m = m; f->seed = seed; f->threads = 1; f->variant = 1;
    f->w = (double*)calloc(n, 8); f->x = (double*)calloc(n, 8); f->y = (double*)calloc(n, 8); f->yaw = (double*)calloc(n, 8);
    f->w2 = (double*)calloc(n, 8); f->x2 = (double*)calloc(n, 8); f->y2 = (double*)calloc(n, 8); f->yaw2 = (double*)calloc(n, 8);
    f->lm = (lm_t*)calloc(n * m + 1, sizeof(lm_t)); f->lm2 = (lm_t*)calloc(n * m + 1, sizeof(lm_t));
    f->last_idx = (uint32_t*)calloc(n, 4);
    /* Particle::new fs1.rs:54-62, Landmark::new fs1.rs:34-40 */
    for (size_t i = 0; i < n; ++i) {
        f->w[i] = cfg->init_weight;
        for (size_t l = 0; l < m; ++l) { lm_t* q = &f->lm[i * m + l]; q->c00 = 1000.0; q->c11 = 1000.0; }
    }
    return f;
}
void orc_fs_free(orc_fs* f) {
    if (!f) return;
    free(f->w); free(f->x); free(f->y); free(f->yaw); free(f->w2); free(f->x2); free(f->y2); free(f->yaw2);
    free(f->lm); free(f->lm2); free(f->last_idx); free(f);
}
void orc_fs_set_state(orc_fs* f, const double* pw, const double* lm) {
    for (size_t i = 0; i < f->n; ++i) { f->w[i] = pw[4 * i]; f->x[i] = pw[4 * i + 1]; f->y[i] = pw[4 * i + 2]; f->yaw[i] = pw[4 * i + 3]; }
    if (lm) memcpy(f->lm, lm, f->n * f->m * sizeof(lm_t));
}
void orc_fs_get_state(const orc_fs* f, double* pw, double* lm) {
    if (pw) for (size_t i = 0; i < f->n; ++i) { pw[4 * i] = f->w[i]; pw[4 * i + 1] = f->x[i]; pw[4 * i + 2] = f->y[i]; pw[4 * i + 3] = f->yaw[i]; }
    if (lm) memcpy(lm, f->lm, f->n * f->m * sizeof(lm_t));
}

void orc_fs_seed_map(orc_fs* f, const double pose[3], const double* lxy, double sigma, double cov0) {
    for (size_t i = 0; i < f->n; ++i) {
        f->x[i] = pose[0]; f->y[i] = pose[1]; f->yaw[i] = pose[2]; f->w[i] = 1.0 / (double)f->n;
        for (size_t l = 0; l < f->m; ++l) {
            double z0, z1;
            pfc_normal_pair(pfc_rng_block(f->seed, PFC_STREAM_INIT_A, 0, (uint64_t)(i * f->m + l)), &z0, &z1);
            lm_t* q = &f->lm[i * f->m + l];
            q->x = lxy[2 * l] + sigma * z0; q->y = lxy[2 * l + 1] + sigma * z1;
            q->c00 = cov0; q->c01 = 0.0; q->c10 = 0.0; q->c11 = cov0;
        }
    }
}

/* normalize_angle fs1.rs:80-89 */
#define normalize_angle orc_fs_normalize_angle

/* predict_particle fs1.rs:123-137 + motion_model fs1.rs:70-77 */
static inline void predict_particle(orc_fs* f, size_t i, const double u[2], double z0, double z1) {
    double u0 = u[0] + z0 * sqrt(f->cfg.q00);                  /* fs1.rs:129 */
    double u1 = u[1] + z1 * sqrt(f->cfg.q11);                  /* fs1.rs:130 */
    double yaw = f->yaw[i];
    double nx = f->x[i] + u0 * f->cfg.dt * M_COS(yaw);         /* fs1.rs:73 */
    double ny = f->y[i] + u0 * f->cfg.dt * M_SIN(yaw);         /* fs1.rs:74 */
    double nyaw = normalize_angle(yaw + u1 * f->cfg.dt);       /* fs1.rs:75 */
    f->x[i] = nx; f->y[i] = ny; f->yaw[i] = nyaw;
}

/* update_landmark fs1.rs:140-183 */
static inline void update_landmark(orc_fs* f, size_t i, double z0, double z1, size_t lm_id) {
    lm_t* L = &f->lm[i * f->m + lm_id];
    const double px = f->x[i], py = f->y[i], pyaw = f->yaw[i];
    if (L->c00 > 100.0) {                                       /* fs1.rs:144-149 (cov left untouched) */
        L->x = px + z0 * M_COS(pyaw + z1);
        L->y = py + z0 * M_SIN(pyaw + z1);
        return;
    }
    /* observation_model fs1.rs:92-99 */
    double dx = L->x - px, dy = L->y - py;
    double d = sqrt(dx * dx + dy * dy);
    double zp1 = normalize_angle(M_ATAN2(dy, dx) - pyaw);
    /* innovation fs1.rs:155 */
    double y0 = z0 - d, y1 = normalize_angle(z1 - zp1);
    /* compute_jacobian fs1.rs:102-110 */
    double d2 = dx * dx + dy * dy;
    double dd = sqrt(d2);
    double h00 = dx / dd, h01 = dy / dd, h10 = -dy / d2, h11 = dx / d2;
    double p00 = L->c00, p01 = L->c01, p10 = L->c10, p11 = L->c11;
    /* S = H P H^T + R  fs1.rs:161 */
    double a00 = h00 * p00 + h01 * p10, a01 = h00 * p01 + h01 * p11;
    double a10 = h10 * p00 + h11 * p10, a11 = h10 * p01 + h11 * p11;
    double s00 = (a00 * h00 + a01 * h01) + f->cfg.r00;
    double s01 = (a00 * h10 + a01 * h11) + 0.0;
    double s10 = (a10 * h00 + a11 * h01) + 0.0;
    double s11 = (a10 * h10 + a11 * h11) + f->cfg.r11;
    /* try_inverse().unwrap_or(identity)  fs1.rs:164 */
    double det = s00 * s11 - s10 * s01;
    double i00, i01, i10, i11;
    if (det == 0.0) { i00 = 1.0; i01 = 0.0; i10 = 0.0; i11 = 1.0; }
    else { i00 = s11 / det; i01 = -s01 / det; i10 = -s10 / det; i11 = s00 / det; }
    /* K = P H^T S^-1  fs1.rs:165 */
    double b00 = p00 * h00 + p01 * h01, b01 = p00 * h10 + p01 * h11;
    double b10 = p10 * h00 + p11 * h01, b11 = p10 * h10 + p11 * h11;
    double k00 = b00 * i00 + b01 * i10, k01 = b00 * i01 + b01 * i11;
    double k10 = b10 * i00 + b11 * i10, k11 = b10 * i01 + b11 * i11;
    /* landmark += K y  fs1.rs:168-170 */
    L->x += k00 * y0 + k01 * y1;
    L->y += k10 * y0 + k11 * y1;
    /* P = (I - K H) P  fs1.rs:173-174 */
    double m00 = 1.0 - (k00 * h00 + k01 * h10), m01 = 0.0 - (k00 * h01 + k01 * h11);
    double m10 = 0.0 - (k10 * h00 + k11 * h10), m11 = 1.0 - (k10 * h01 + k11 * h11);
    L->c00 = m00 * p00 + m01 * p10; L->c01 = m00 * p01 + m01 * p11;
    L->c10 = m10 * p00 + m11 * p10; L->c11 = m10 * p01 + m11 * p11;
    /* weight fs1.rs:177-182 */
    double det_s = s00 * s11 - s10 * s01;
    if (det_s > 0.0) {
        double t0 = y0 * i00 + y1 * i10, t1 = y0 * i01 + y1 * i11;
        double mahal = t0 * y0 + t1 * y1;
        double likelihood = M_EXP(-0.5 * mahal) / (2.0 * PFC_PI * sqrt(det_s));
        f->w[i] *= likelihood;
    }
}

/* normalize_weights fs1.rs:196-203 (no uniform fallback) */
void orc_fs_normalize_weights_(orc_fs* f) {
    double sum_w = 0.0;
    for (size_t i = 0; i < f->n; ++i) sum_w += f->w[i];
    if (sum_w > 0.0) for (size_t i = 0; i < f->n; ++i) f->w[i] /= sum_w;
}
/* compute_neff fs1.rs:186-193 */
double orc_fs_compute_neff_(const orc_fs* f) {
    double s2 = 0.0;
    for (size_t i = 0; i < f->n; ++i) s2 += f->w[i] * f->w[i];
    return s2 > 0.0 ? 1.0 / s2 : 0.0;
}
/* resample fs1.rs:206-234 */
void orc_fs_resample_(orc_fs* f, double u01) {
    orc_fs_normalize_weights_(f);
    size_t n = f->n, m = f->m;
    double* cum = (double*)malloc(sizeof(double) * (n + 1));
    cum[0] = 0.0;
    for (size_t i = 0; i < n; ++i) cum[i + 1] = cum[i] + f->w[i];
    /* Uniform::new(0, 1/n).sample = u01 * scale + low (rand 0.9 UniformFloat) */
    double r = u01 * (1.0 / (double)n - 0.0) + 0.0;
    size_t j = 0;
    for (size_t t = 0; t < n; ++t) {
        while (r > cum[j + 1] && j < n - 1) j++;
        f->last_idx[t] = (uint32_t)j;
        r += 1.0 / (double)n;
    }
    /* particles[j].clone(): the copies are independent, so the all-core baseline may do them in parallel */
    long nn = (long)n;
#pragma omp parallel for num_threads(f->threads) schedule(static) if (f->threads > 1)
    for (long t = 0; t < nn; ++t) {
        size_t jj = f->last_idx[t];
        f->w2[t] = 1.0 / (double)n; f->x2[t] = f->x[jj]; f->y2[t] = f->y[jj]; f->yaw2[t] = f->yaw[jj];
        memcpy(&f->lm2[(size_t)t * m], &f->lm[jj * m], m * sizeof(lm_t));
    }
    double* t; lm_t* tl;
    t = f->w; f->w = f->w2; f->w2 = t; t = f->x; f->x = f->x2; f->x2 = t;
    t = f->y; f->y = f->y2; f->y2 = t; t = f->yaw; f->yaw = f->yaw2; f->yaw2 = t;
    tl = f->lm; f->lm = f->lm2; f->lm2 = tl;
    f->last_idx_n = n;
    free(cum);
}

/* fastslam_update fs1.rs:237-266 */
static int step_impl(orc_fs* f, const double u[2], const orc_fs_obs* z, size_t k,
                     const double* nz0, const double* nz1, const double* r01) {
    long n = (long)f->n;
    const uint64_t seed = f->seed; const uint32_t call = f->n_step;
    if (f->variant == 2) {
        /* fastslam2_update_with_rng fs2.rs:330-374: particle-outer; per particle the draws are (n0, n1, n2) with observations
         * (sample_pose fs2.rs:234) or (n0, n1) without (fs2.rs:346-349).  n0, n1 come from the FS_PREDICT block of the
         * particle, n2 from the FS2_POSE3 block (injected arrays: nz0, nz1 and, behind them, nz1 + n as the third column). */
#pragma omp parallel for num_threads(f->threads) schedule(static) if (f->threads > 1)
        for (long i = 0; i < n; ++i) {
            double a0, a1, a2, dummy;
            if (nz0) { a0 = nz0[i]; a1 = nz1[i]; a2 = nz1[n + i]; }
            else {
                pfc_normal_pair(pfc_rng_block(seed, PFC_STREAM_FS_PREDICT, call, (uint64_t)i), &a0, &a1);
                pfc_normal_pair(pfc_rng_block(seed, PFC_STREAM_FS2_POSE3, call, (uint64_t)i), &a2, &dummy);
            }
            orc_fs2_particle_(f, (size_t)i, u, z, k, a0, a1, a2);
        }
    } else {
#pragma omp parallel for num_threads(f->threads) schedule(static) if (f->threads > 1)
    for (long i = 0; i < n; ++i) {                              /* fs1.rs:245-247 */
        double z0, z1;
        if (nz0) { z0 = nz0[i]; z1 = nz1[i]; }
        else pfc_normal_pair(pfc_rng_block(seed, PFC_STREAM_FS_PREDICT, call, (uint64_t)i), &z0, &z1);
        predict_particle(f, (size_t)i, u, z0, z1);
    }
    /* fs1.rs:250-256 is obs-outer / particle-inner; particles are independent, so iterating
     * particle-outer / obs-inner visits each (particle, obs) pair in the same per-particle order. */
#pragma omp parallel for num_threads(f->threads) schedule(static) if (f->threads > 1)
    for (long i = 0; i < n; ++i)
        for (size_t j = 0; j < k; ++j)
            if (z[j].lm_id < f->m) update_landmark(f, (size_t)i, z[j].d, z[j].angle, (size_t)z[j].lm_id);
    }
    orc_fs_normalize_weights_(f);                               /* fs1.rs:259 */
    double neff = orc_fs_compute_neff_(f);                      /* fs1.rs:262 */
    f->last_neff = neff;
    f->n_step++;
    if (neff < f->cfg.nth) {                                    /* fs1.rs:263-265 */
        double u01 = r01 ? *r01
                         : pfc_u01_52(pfc_blk_u64(pfc_rng_block(seed, PFC_STREAM_FS_RESAMPLE, f->n_resample, 0), 0));
        orc_fs_resample_(f, u01);
        f->n_resample++;
        return 1;
    }
    f->last_idx_n = 0;
    return 0;
}
int orc_fs_step(orc_fs* f, const double u[2], const orc_fs_obs* z, size_t k) {
    return step_impl(f, u, z, k, NULL, NULL, NULL);
}
int orc_fs_step_with_noise(orc_fs* f, const double u[2], const orc_fs_obs* z, size_t k,
                           const double* z0, const double* z1, double r01) {
    return step_impl(f, u, z, k, z0, z1, &r01);
}

/* get_best_particle fs1.rs:269-274: Iterator::max_by keeps the LAST maximum among equals */
size_t orc_fs_best(const orc_fs* f) {
    size_t best = 0;
    for (size_t i = 1; i < f->n; ++i) if (f->w[i] >= f->w[best]) best = i;
    return best;
}
size_t orc_fs_last_indices(const orc_fs* f, uint32_t* idx, size_t cap) {
    size_t n = f->last_idx_n < cap ? f->last_idx_n : cap;
    for (size_t i = 0; i < n; ++i) idx[i] = f->last_idx[i];
    return f->last_idx_n;
}
double orc_fs_last_neff(const orc_fs* f) { return f->last_neff; }
void orc_fs_set_threads(orc_fs* f, int t) { f->threads = t < 1 ? 1 : t; }
void orc_fs_set_variant(orc_fs* f, int v) { f->variant = v == 2 ? 2 : 1; }

/* get_observations fs1.rs:277-299 */
size_t orc_fs_get_observations(const orc_fs_config* c, const double xt[3], const double* lxy, size_t nl,
                               uint64_t seed, uint32_t call, orc_fs_obs* out) {
    size_t cnt = 0;
    for (size_t id = 0; id < nl; ++id) {
        double dx = lxy[2 * id] - xt[0], dy = lxy[2 * id + 1] - xt[1];
        double d = sqrt(dx * dx + dy * dy);
        if (d <= c->max_range) {
            double angle = normalize_angle(M_ATAN2(dy, dx) - xt[2]);
            double z0, z1;
            pfc_normal_pair(pfc_rng_block(seed, PFC_STREAM_OBS, call, id), &z0, &z1);
            out[cnt].d = d + z0 * sqrt(c->r00);                 /* fs1.rs:291 */
            out[cnt].angle = angle + z1 * sqrt(c->r11);         /* fs1.rs:292 */
            out[cnt].lm_id = id;
            cnt++;
        }
    }
    return cnt;
}
