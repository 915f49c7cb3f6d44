/*
 * pf_oracle.c — CPU oracle for ParticleFilterLocalizer (pf.rs) and MonteCarloLocalizer (mcl.rs).
 * TEST INFRASTRUCTURE ONLY — see oracle.h.  PARITY UNPINNED (no reference-generated vectors exist).
 *
 * Every function cites the reference lines it restates.  "pf.rs" = crates/rust_robotics_localization/
 * src/particle_filter.rs, "mcl.rs" = .../monte_carlo_localization.rs.
 */
#include "oracle.h"
#include "../include/pf_contract_math.h"
#include <stdlib.h>
#include <stdio.h>

#ifdef PF_ORACLE_LIBM
#define M_EXP(x) exp(x)
#define M_SIN(x) sin(x)
#define M_COS(x) cos(x)
#else
#define M_EXP(x) pfc_exp(x)
#define M_SIN(x) pfc_sin(x)
#define M_COS(x) pfc_cos(x)
#endif

struct orc_pf {
    orc_pf_config cfg;
    size_t n;                 /* current particle count (MCL: varies) */
    size_t cap;
    orc_particle* p;
    orc_particle* scratch;
    double est[4];            /* state_estimate  pf.rs:126 */
    double cov[16];           /* covariance_dyn  pf.rs:127, stored [i*4+j] */
    uint64_t seed;
    uint32_t n_predict, n_resample;
    uint32_t* last_idx; size_t last_idx_n;
    int fast_search;
    int threads;
};

/* ---- validation: pf.rs:81-117, mcl.rs:87-130 ---- */
static int finite_(double v) { return v == v && fabs(v) <= 1.7976931348623157e308; }

int orc_pf_config_validate(const orc_pf_config* c) {
    if (c->n_particles == 0) return -1;
    if (c->mode == 0) {
        if (!finite_(c->resample_threshold) || c->resample_threshold < 0.0 || c->resample_threshold > 1.0) return -1;
    } else {
        if (c->max_particles < c->n_particles) return -1;
        if (!finite_(c->kld_epsilon) || c->kld_epsilon <= 0.0) return -1;
        if (!finite_(c->kld_z) || c->kld_z <= 0.0) return -1;
    }
    if (!finite_(c->range_noise) || c->range_noise <= 0.0) return -1;
    if (!finite_(c->velocity_noise) || c->velocity_noise < 0.0) return -1;
    if (!finite_(c->yaw_rate_noise) || c->yaw_rate_noise < 0.0) return -1;
    if (!finite_(c->dt) || c->dt <= 0.0) return -1;
    return 0;
}

/* ---- compute_estimate pf.rs:382-396 / mcl.rs:413-427 ---- */
static void compute_estimate(const orc_pf* f, double est[4]) {
    double x_est = 0.0, y_est = 0.0, yaw_est = 0.0, v_est = 0.0;
    for (size_t i = 0; i < f->n; ++i) {
        const orc_particle* q = &f->p[i];
        x_est += q->w * q->x;
        y_est += q->w * q->y;
        yaw_est += q->w * q->yaw;
        v_est += q->w * q->v;
    }
    est[0] = x_est; est[1] = y_est; est[2] = yaw_est; est[3] = v_est;
}

/* ---- compute_covariance pf.rs:398-413: cov += (w * dx) * dx^T, all 16 entries ---- */
static void compute_covariance(const orc_pf* f, const double est[4], double cov[16]) {
    for (int k = 0; k < 16; ++k) cov[k] = 0.0;
    for (size_t i = 0; i < f->n; ++i) {
        const orc_particle* q = &f->p[i];
        double dx[4] = { q->x - est[0], q->y - est[1], q->yaw - est[2], q->v - est[3] };
        double wdx[4] = { q->w * dx[0], q->w * dx[1], q->w * dx[2], q->w * dx[3] };
        for (int a = 0; a < 4; ++a)
            for (int b = 0; b < 4; ++b)
                cov[a * 4 + b] += wdx[a] * dx[b];
    }
}

/* ---- refresh_cache pf.rs:499-503 ---- */
static void refresh_cache(orc_pf* f) {
    compute_estimate(f, f->est);
    compute_covariance(f, f->est, f->cov);
}

orc_pf* orc_pf_new(const orc_pf_config* cfg, uint64_t seed) {
    if (orc_pf_config_validate(cfg) != 0) return NULL;
    orc_pf* f = (orc_pf*)calloc(1, sizeof(orc_pf));
    f->cfg = *cfg;
    f->seed = seed;
    f->n = cfg->n_particles;
    f->cap = cfg->mode == 1 ? cfg->max_particles : cfg->n_particles;
    f->p = (orc_particle*)calloc(f->cap, sizeof(orc_particle));
    f->scratch = (orc_particle*)calloc(f->cap, sizeof(orc_particle));
    f->last_idx = (uint32_t*)calloc(f->cap, sizeof(uint32_t));
    f->threads = 1;
    /* pf.rs:142-144: Particle::new(0,0,0,0,n) -> w = 1/n */
    for (size_t i = 0; i < f->n; ++i) f->p[i].w = 1.0 / (double)f->n;
    refresh_cache(f);
    return f;
}

void orc_pf_free(orc_pf* f) {
    if (!f) return;
    free(f->p); free(f->scratch); free(f->last_idx); free(f);
}

/* try_with_initial_state pf.rs:170-199 (random::<f64>()*2-1 ...) and mcl.rs:176-206
 * (random_range(-1.0..1.0): rand 0.9 UniformFloat::sample_single = (value1_2 - 1) * scale + low, 52 bits). */
int orc_pf_init_state(orc_pf* f, const double s[4]) {
    for (int k = 0; k < 4; ++k) if (!finite_(s[k])) return -1;
    size_t n = f->cfg.n_particles;
    f->n = n;
    for (size_t i = 0; i < n; ++i) {
        pfc_u32x4 a = pfc_rng_block(f->seed, PFC_STREAM_INIT_A, 0, i);
        pfc_u32x4 b = pfc_rng_block(f->seed, PFC_STREAM_INIT_B, 0, i);
        orc_particle* q = &f->p[i];
        if (f->cfg.mode == 0) {
            q->x   = s[0] + pfc_u01_53(pfc_blk_u64(a, 0)) * 2.0 - 1.0;
            q->y   = s[1] + pfc_u01_53(pfc_blk_u64(a, 1)) * 2.0 - 1.0;
            q->yaw = s[2] + pfc_u01_53(pfc_blk_u64(b, 0)) * 0.5 - 0.25;
            q->v   = s[3] + pfc_u01_53(pfc_blk_u64(b, 1)) * 1.0 - 0.5;
        } else {
            q->x   = s[0] + (pfc_u01_52(pfc_blk_u64(a, 0)) * 2.0 + -1.0);
            q->y   = s[1] + (pfc_u01_52(pfc_blk_u64(a, 1)) * 2.0 + -1.0);
            q->yaw = s[2] + (pfc_u01_52(pfc_blk_u64(b, 0)) * 0.5 + -0.25);
            q->v   = s[3] + (pfc_u01_52(pfc_blk_u64(b, 1)) * 1.0 + -0.5);
        }
        q->w = 1.0 / (double)n;
    }
    refresh_cache(f);
    return 0;
}

size_t orc_pf_count(const orc_pf* f) { return f->n; }

void orc_pf_set_particles(orc_pf* f, const double* a, size_t n) {
    if (n > f->cap) n = f->cap;
    f->n = n;
    for (size_t i = 0; i < n; ++i) {
        f->p[i].x = a[5 * i]; f->p[i].y = a[5 * i + 1]; f->p[i].yaw = a[5 * i + 2];
        f->p[i].v = a[5 * i + 3]; f->p[i].w = a[5 * i + 4];
    }
    refresh_cache(f);
}
void orc_pf_get_particles(const orc_pf* f, double* a) {
    for (size_t i = 0; i < f->n; ++i) {
        a[5 * i] = f->p[i].x; a[5 * i + 1] = f->p[i].y; a[5 * i + 2] = f->p[i].yaw;
        a[5 * i + 3] = f->p[i].v; a[5 * i + 4] = f->p[i].w;
    }
}

/* ---- try_predict_with_control pf.rs:255-301 / mcl.rs:209-257 ---- */
static int predict_impl(orc_pf* f, const double u[2], const double* zv, const double* zw) {
    if (!finite_(u[0]) || !finite_(u[1])) return -1;          /* validate_control pf.rs:515-523 */
    const double sv = f->cfg.velocity_noise, sw = f->cfg.yaw_rate_noise, dt = f->cfg.dt;
    const uint64_t seed = f->seed; const uint32_t call = f->n_predict;
    long n = (long)f->n;
#pragma omp parallel for num_threads(f->threads) schedule(static) if (f->threads > 1)
    for (long i = 0; i < n; ++i) {
        orc_particle* q = &f->p[i];
        double z0, z1;
        if (zv) { z0 = zv[i]; z1 = zw[i]; }
        else pfc_normal_pair(pfc_rng_block(seed, PFC_STREAM_PF_PREDICT, call, (uint64_t)i), &z0, &z1);
        /* Normal::sample = mean + std_dev * z; no draw when sigma == 0 (pf.rs:259-287) */
        double v_noise   = sv > 0.0 ? 0.0 + sv * z0 : 0.0;
        double yaw_noise = sw > 0.0 ? 0.0 + sw * z1 : 0.0;
        double v_noisy = u[0] + v_noise;                       /* pf.rs:289 */
        double yaw_rate_noisy = u[1] + yaw_noise;              /* pf.rs:290 */
        double c = M_COS(q->yaw), s = M_SIN(q->yaw);
        q->x += v_noisy * c * dt;                              /* pf.rs:292 */
        q->y += v_noisy * s * dt;                              /* pf.rs:293 */
        q->yaw += yaw_rate_noisy * dt;                         /* pf.rs:294 (no wrap) */
        q->v = v_noisy;                                        /* pf.rs:295 */
    }
    f->n_predict++;
    refresh_cache(f);                                          /* pf.rs:299 */
    return 0;
}
int orc_pf_predict(orc_pf* f, const double u[2]) { return predict_impl(f, u, NULL, NULL); }
int orc_pf_predict_with_noise(orc_pf* f, const double u[2], const double* zv, const double* zw) {
    return predict_impl(f, u, zv, zw);
}

/* ---- gauss_likelihood pf.rs:476-479 ---- */
static inline double gauss_likelihood(double x, double sigma) {
    double coeff = 1.0 / sqrt(2.0 * PFC_PI * (sigma * sigma));
    return coeff * M_EXP(-(x * x) / (2.0 * (sigma * sigma)));
}

/* ---- normalize_weights pf.rs:426-439 / mcl.rs:394-406 ---- */
static void normalize_weights(orc_pf* f) {
    double sum_w = 0.0;
    for (size_t i = 0; i < f->n; ++i) sum_w += f->p[i].w;
    if (sum_w > 0.0) {
        for (size_t i = 0; i < f->n; ++i) f->p[i].w /= sum_w;
    } else {
        double uw = 1.0 / (double)f->n;
        for (size_t i = 0; i < f->n; ++i) f->p[i].w = uw;
    }
}

/* ---- try_update_with_observations pf.rs:310-334 / mcl.rs:260-288 ---- */
int orc_pf_update(orc_pf* f, const double* obs, size_t k) {
    for (size_t j = 0; j < k; ++j)                             /* validate_observations pf.rs:538-549 */
        if (!finite_(obs[3 * j]) || !finite_(obs[3 * j + 1]) || !finite_(obs[3 * j + 2]) || obs[3 * j] < 0.0)
            return -1;
    const double sigma = f->cfg.range_noise;
    long n = (long)f->n;
#pragma omp parallel for num_threads(f->threads) schedule(static) if (f->threads > 1)
    for (long i = 0; i < n; ++i) {
        orc_particle* q = &f->p[i];
        double w = 1.0;                                        /* pf.rs:317: previous weight discarded */
        for (size_t j = 0; j < k; ++j) {
            double dx = q->x - obs[3 * j + 1];
            double dy = q->y - obs[3 * j + 2];
            double d_pred = sqrt(dx * dx + dy * dy);
            double diff = obs[3 * j] - d_pred;
            w *= gauss_likelihood(diff, sigma);
        }
        q->w = w;
    }
    normalize_weights(f);                                      /* pf.rs:331 */
    refresh_cache(f);                                          /* pf.rs:332 */
    return 0;
}

/* ---- calc_n_eff pf.rs:416-423 ---- */
double orc_pf_neff(const orc_pf* f) {
    double s2 = 0.0;
    for (size_t i = 0; i < f->n; ++i) s2 += f->p[i].w * f->p[i].w;
    return s2 > 0.0 ? 1.0 / s2 : 0.0;
}

/* first i with r <= c_i; `fallback` if none.  Linear scan as written (pf.rs:459-465, mcl.rs:387-392) or an
 * equivalent lower_bound on the non-decreasing cumsum (identical result, used for large-N baselines). */
static size_t find_index(const double* cum, size_t n, double r, size_t fallback, int fast) {
    if (!fast) {
        for (size_t i = 0; i < n; ++i) if (r <= cum[i]) return i;
        return fallback;
    }
    size_t lo = 0, hi = n;
    while (lo < hi) { size_t mid = lo + (hi - lo) / 2; if (cum[mid] < r) lo = mid + 1; else hi = mid; }
    return lo < n ? lo : fallback;
}

/* ---- resample_particles pf.rs:442-473 ---- */
static void pf_resample_particles(orc_pf* f, const double* rin) {
    size_t n = f->cfg.n_particles;
    double* cum = (double*)malloc(sizeof(double) * f->n);
    double cum_sum = 0.0;
    for (size_t i = 0; i < f->n; ++i) { cum_sum += f->p[i].w; cum[i] = cum_sum; }
    for (size_t t = 0; t < n; ++t) {
        double r = rin ? rin[t]
                       : pfc_u01_53(pfc_blk_u64(pfc_rng_block(f->seed, PFC_STREAM_PF_RESAMPLE, f->n_resample, t), 0));
        size_t index = find_index(cum, f->n, r, 0, f->fast_search);   /* default 0: pf.rs:459 */
        f->scratch[t] = f->p[index];
        f->scratch[t].w = 1.0 / (double)n;
        f->last_idx[t] = (uint32_t)index;
    }
    orc_particle* tmp = f->p; f->p = f->scratch; f->scratch = tmp;
    f->n = n; f->last_idx_n = n;
    f->n_resample++;
    free(cum);
}

/* ---- MCL helpers ---- */
static int32_t sat_i32(double v) {                              /* Rust `as i32` saturates, NaN -> 0 */
    if (v != v) return 0;
    if (v >= 2147483647.0) return 2147483647;
    if (v <= -2147483648.0) return (int32_t)(-2147483647 - 1);
    return (int32_t)v;
}
static size_t sat_usize(double v) {                             /* Rust `as usize` */
    if (v != v || v <= 0.0) return 0;
    if (v >= 18446744073709551615.0) return (size_t)-1;
    return (size_t)v;
}
/* kld_required_particles mcl.rs:367-378 */
static size_t kld_required(const orc_pf_config* c, size_t k_bins) {
    if (k_bins <= 1) return c->n_particles;
    double km1 = (double)(k_bins - 1);
    double term = 1.0 - 2.0 / (9.0 * km1) + c->kld_z * sqrt(2.0 / (9.0 * km1));
    double nn = (km1 / (2.0 * c->kld_epsilon)) * (term * term * term);
    size_t v = sat_usize(ceil(nn));
    if (v < c->n_particles) v = c->n_particles;
    if (v > c->max_particles) v = c->max_particles;
    return v;
}
typedef struct { int32_t a, b, c; int used; } bin_t;
static int bins_insert(bin_t* tab, size_t cap, int32_t a, int32_t b, int32_t c) {   /* 1 if new */
    uint64_t h = (uint64_t)(uint32_t)a * 0x9E3779B97F4A7C15ull ^ (uint64_t)(uint32_t)b * 0xC2B2AE3D27D4EB4Full
               ^ (uint64_t)(uint32_t)c * 0x165667B19E3779F9ull;
    size_t i = (size_t)(h % cap);
    while (tab[i].used) {
        if (tab[i].a == a && tab[i].b == b && tab[i].c == c) return 0;
        i = (i + 1) % cap;
    }
    tab[i].a = a; tab[i].b = b; tab[i].c = c; tab[i].used = 1;
    return 1;
}

/* ---- resample_adaptive mcl.rs:322-365 ---- */
static void mcl_resample_adaptive(orc_pf* f, const double* rin, size_t nr) {
    size_t n_current = f->n;
    if (n_current == 0) return;
    const orc_pf_config* c = &f->cfg;
    double* cum = (double*)malloc(sizeof(double) * n_current);
    double cum_sum = 0.0;
    for (size_t i = 0; i < n_current; ++i) { cum_sum += f->p[i].w; cum[i] = cum_sum; }
    cum[n_current - 1] = 1.0;                                   /* mcl.rs:334-336 */
    size_t cap = 2 * c->max_particles + 16;
    bin_t* bins = (bin_t*)calloc(cap, sizeof(bin_t));
    size_t nbins = 0, required = c->n_particles, len = 0;
    const double X_BIN = 0.5, Y_BIN = 0.5, YAW_BIN = 15.0 * PFC_PI / 180.0;   /* mcl.rs:26-28 */
    while (len < c->max_particles) {
        double r;
        if (rin) { if (len >= nr) break; r = rin[len]; }
        else r = pfc_u01_53(pfc_blk_u64(pfc_rng_block(f->seed, PFC_STREAM_PF_RESAMPLE, f->n_resample, len), 0));
        size_t idx = find_index(cum, n_current, r, n_current - 1, f->fast_search);  /* mcl.rs:387-392 */
        const orc_particle* s = &f->p[idx];
        nbins += (size_t)bins_insert(bins, cap, sat_i32(floor(s->x / X_BIN)), sat_i32(floor(s->y / Y_BIN)),
                                     sat_i32(floor(s->yaw / YAW_BIN)));           /* mcl.rs:380-385 */
        size_t kr = kld_required(c, nbins);
        if (kr > required) required = kr;
        f->scratch[len] = *s;
        f->last_idx[len] = (uint32_t)idx;
        len++;
        if (len >= c->n_particles && len >= required) break;    /* mcl.rs:352-354 */
    }
    double uw = 1.0 / (double)len;
    for (size_t i = 0; i < len; ++i) f->scratch[i].w = uw;
    orc_particle* tmp = f->p; f->p = f->scratch; f->scratch = tmp;
    f->n = len; f->last_idx_n = len;
    f->n_resample++;
    free(cum); free(bins);
    refresh_cache(f);                                           /* mcl.rs:364 */
}

/* ---- resample pf.rs:337-345; MCL resamples every step (mcl.rs:298) ---- */
int orc_pf_resample(orc_pf* f) {
    if (f->cfg.mode == 1) { mcl_resample_adaptive(f, NULL, 0); return 1; }
    double n_eff = orc_pf_neff(f);
    double threshold = (double)f->cfg.n_particles * f->cfg.resample_threshold;
    if (n_eff < threshold) {
        pf_resample_particles(f, NULL);
        refresh_cache(f);
        return 1;
    }
    return 0;
}
int orc_pf_resample_with_uniforms(orc_pf* f, const double* r, size_t nr) {
    if (f->cfg.mode == 1) { mcl_resample_adaptive(f, r, nr); return 1; }
    if (nr < f->cfg.n_particles) return -1;
    pf_resample_particles(f, r);
    refresh_cache(f);
    return 1;
}

/* ---- try_step pf.rs:488-497 / mcl.rs:291-300 ---- */
int orc_pf_step(orc_pf* f, const double u[2], const double* obs, size_t k, double est[4]) {
    int rc = orc_pf_predict(f, u);
    if (rc) return rc;
    rc = orc_pf_update(f, obs, k);
    if (rc) return rc;
    int did = orc_pf_resample(f);
    if (est) for (int i = 0; i < 4; ++i) est[i] = f->est[i];
    return did;
}

void orc_pf_estimate(const orc_pf* f, double est[4], double cov_cm[16]) {
    if (est) for (int i = 0; i < 4; ++i) est[i] = f->est[i];
    if (cov_cm) for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) cov_cm[j * 4 + i] = f->cov[i * 4 + j];
}

int orc_pf_set_range_noise(orc_pf* f, double s) {              /* pf.rs:228-236 */
    if (!finite_(s) || s <= 0.0) return -1;
    f->cfg.range_noise = s; return 0;
}
size_t orc_pf_last_indices(const orc_pf* f, uint32_t* idx, size_t cap) {
    size_t n = f->last_idx_n < cap ? f->last_idx_n : cap;
    for (size_t i = 0; i < n; ++i) idx[i] = f->last_idx[i];
    return f->last_idx_n;
}
void orc_pf_set_fast_search(orc_pf* f, int on) { f->fast_search = on; }
void orc_pf_set_threads(orc_pf* f, int t) { f->threads = t < 1 ? 1 : t; }

/* ---- contract-math probes ---- */
void orc_math_exp(const double* in, double* out, size_t n) { for (size_t i = 0; i < n; ++i) out[i] = pfc_exp(in[i]); }
void orc_math_log(const double* in, double* out, size_t n) { for (size_t i = 0; i < n; ++i) out[i] = pfc_log(in[i]); }
void orc_math_sin(const double* in, double* out, size_t n) { for (size_t i = 0; i < n; ++i) out[i] = pfc_sin(in[i]); }
void orc_math_cos(const double* in, double* out, size_t n) { for (size_t i = 0; i < n; ++i) out[i] = pfc_cos(in[i]); }
void orc_math_atan2(const double* y, const double* x, double* out, size_t n) {
    for (size_t i = 0; i < n; ++i) out[i] = pfc_atan2(y[i], x[i]);
}
void orc_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4]) {
    pfc_u32x4 o = pfc_philox4x32_10(c0, c1, c2, c3, k0, k1);
    for (int i = 0; i < 4; ++i) out[i] = o.v[i];
}
void orc_normal_pair(uint64_t seed, uint32_t stream, uint32_t call, uint64_t index, double z[2]) {
    pfc_normal_pair(pfc_rng_block(seed, stream, call, index), &z[0], &z[1]);
}
double orc_uniform53(uint64_t seed, uint32_t stream, uint32_t call, uint64_t index) {
    return pfc_u01_53(pfc_blk_u64(pfc_rng_block(seed, stream, call, index), 0));
}
double orc_uniform52(uint64_t seed, uint32_t stream, uint32_t call, uint64_t index) {
    return pfc_u01_52(pfc_blk_u64(pfc_rng_block(seed, stream, call, index), 0));
}
int orc_math_mode(void) {
#ifdef PF_ORACLE_LIBM
    return 1;
#else
    return 0;
#endif
}
