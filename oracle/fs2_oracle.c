/*
 * fs2_oracle.c — CPU oracle for FastSLAM 2.0 (crates/rust_robotics_slam/src/fastslam2.rs = "fs2.rs").
 * TEST INFRASTRUCTURE ONLY — see oracle.h.  PARITY UNPINNED: the two seeded tests of the reference (fs2.rs:456-545) draw from
 * rand::StdRng (ChaCha12) through rand_distr's ziggurat, neither of which can be rebuilt offline, so no reference-generated
 * vector exists; noise is injected (Philox), as for fs1.  The properties those tests assert are mirrored in
 * tests/test_fs2_oracle.py.
 *
 * The particle set, normalise / N_eff / resample are fs1_oracle.c's (fs2.rs:282-323 is the same text as fs1.rs:186-234).
 * What differs (fs2.rs:330-374): the pose of a particle is SAMPLED from a Gaussian that fuses the motion prior with the first
 * observation of the step (compute_proposal fs2.rs:173-216, sample_pose fs2.rs:219-239), the landmark test is
 * `cov[(0,0)] < 100` (fs2.rs:49-51), a fresh landmark gets cov = 10·I (fs2.rs:254) and a non-positive det S multiplies the
 * weight by 1e-10 (fs2.rs:278).
 *
 * nalgebra 0.33 semantics restated from upstream (not vendored):
 *   A * B (static sizes)  column by column, entry (i,j) = ((a_i0*b_0j) + a_i1*b_1j) + a_i2*b_2j  (gemv / axcpy order)
 *   try_inverse 2x2       det = m11*m22 - m21*m12; None if det == 0; entries divided by det
 *   try_inverse 3x3       three minors of the first row, det = (m11*M1 - m12*M2) + m13*M3, None if det == 0,
 *                         every adjugate entry computed as a difference of two products and divided by det
 *   cholesky              column j: subtract l_jk * column k (k < j, in order) from rows j.. of column j, then
 *                         l_jj = sqrt(diag) (None if diag is zero, negative or NaN), rows below divided by l_jj
 */
#include "fs_state.h"

static const double MOTION_COV[3][3] = { { 0.1, 0.0, 0.0 }, { 0.0, 0.1, 0.0 }, { 0.0, 0.0, 0.01 } };   /* fs2.rs:31 */

typedef struct { double m[3][3]; } m3;

static m3 m3_mul(const m3* a, const m3* b) {
    m3 c;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double acc = a->m[i][0] * b->m[0][j];
            acc = a->m[i][1] * b->m[1][j] + acc;
            acc = a->m[i][2] * b->m[2][j] + acc;
            c.m[i][j] = acc;
        }
    return c;
}
static m3 m3_transpose(const m3* a) {
    m3 t;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) t.m[i][j] = a->m[j][i];
    return t;
}
static int m3_try_inverse(const m3* a, m3* out) {
    const double m11 = a->m[0][0], m12 = a->m[0][1], m13 = a->m[0][2];
    const double m21 = a->m[1][0], m22 = a->m[1][1], m23 = a->m[1][2];
    const double m31 = a->m[2][0], m32 = a->m[2][1], m33 = a->m[2][2];
    const double minor_m12_m23 = m22 * m33 - m32 * m23;
    const double minor_m11_m23 = m21 * m33 - m31 * m23;
    const double minor_m11_m22 = m21 * m32 - m31 * m22;
    const double det = m11 * minor_m12_m23 - m12 * minor_m11_m23 + m13 * minor_m11_m22;
    if (det == 0.0) return 0;
    out->m[0][0] = minor_m12_m23 / det;
    out->m[0][1] = (m13 * m32 - m33 * m12) / det;
    out->m[0][2] = (m12 * m23 - m22 * m13) / det;
    out->m[1][0] = -minor_m11_m23 / det;
    out->m[1][1] = (m11 * m33 - m31 * m13) / det;
    out->m[1][2] = (m13 * m21 - m23 * m11) / det;
    out->m[2][0] = minor_m11_m22 / det;
    out->m[2][1] = (m12 * m31 - m32 * m11) / det;
    out->m[2][2] = (m11 * m22 - m21 * m12) / det;
    return 1;
}
/* Cholesky::new + l(): lower factor, upper triangle zeroed; 0 when a pivot is zero, negative or NaN */
static int m3_cholesky_l(const m3* a, m3* l) {
    m3 w = *a;
    for (int j = 0; j < 3; ++j) {
        for (int k = 0; k < j; ++k) {
            const double factor = -w.m[j][k];
            for (int i = j; i < 3; ++i) w.m[i][j] = factor * w.m[i][k] + w.m[i][j];
        }
        const double diag = w.m[j][j];
        if (diag == 0.0 || !(diag >= 0.0)) return 0;
        const double denom = sqrt(diag);
        w.m[j][j] = denom;
        for (int i = j + 1; i < 3; ++i) w.m[i][j] = w.m[i][j] / denom;
    }
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) l->m[i][j] = j <= i ? w.m[i][j] : 0.0;
    return 1;
}

/* motion_model fs2.rs:95-102 */
static void motion_model(const double x[3], const double u[2], double dt, double out[3]) {
    const double yaw = x[2];
    out[0] = x[0] + u[0] * dt * M_COS(yaw);
    out[1] = x[1] + u[0] * dt * M_SIN(yaw);
    out[2] = orc_fs_normalize_angle(x[2] + u[1] * dt);
}

/* compute_proposal fs2.rs:173-216 */
static void compute_proposal(const orc_fs_config* c, const double pose[3], const double u[2], double z0, double z1,
                             const lm_t* lm, double mean[3], m3* cov) {
    double x_pred[3];
    motion_model(pose, u, c->dt, x_pred);                                     /* fs2.rs:183 */
    const double yaw = pose[2], v = u[0];
    m3 g = { { { 1.0, 0.0, -v * c->dt * M_SIN(yaw) }, { 0.0, 1.0, v * c->dt * M_COS(yaw) }, { 0.0, 0.0, 1.0 } } };   /* fs2.rs:105-120 */
    m3 mc;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) mc.m[i][j] = MOTION_COV[i][j];
    m3 gm = m3_mul(&g, &mc), gt = m3_transpose(&g);
    m3 p_pred = m3_mul(&gm, &gt);                                             /* fs2.rs:186 */
    if (!(lm->c00 < 100.0)) {                                                 /* fs2.rs:188-191 */
        mean[0] = x_pred[0]; mean[1] = x_pred[1]; mean[2] = x_pred[2];
        *cov = p_pred;
        return;
    }
    /* obs_jacobian_pose fs2.rs:141-148 at the predicted pose, obs_jacobian_landmark fs2.rs:132-138 */
    const double dx = lm->x - x_pred[0], dy = lm->y - x_pred[1];
    const double d2 = dx * dx + dy * dy;
    const double d = sqrt(d2);
    const double hp[2][3] = { { -dx / d, -dy / d, 0.0 }, { dy / d2, -dx / d2, -1.0 } };
    const double hl[2][2] = { { dx / d, dy / d }, { -dy / d2, dx / d2 } };
    /* q_obs = h_lm * lm.cov * h_lm^T + r  fs2.rs:198 */
    const double p00 = lm->c00, p01 = lm->c01, p10 = lm->c10, p11 = lm->c11;
    const double a00 = hl[0][0] * p00 + hl[0][1] * p10, a01 = hl[0][0] * p01 + hl[0][1] * p11;
    const double a10 = hl[1][0] * p00 + hl[1][1] * p10, a11 = hl[1][0] * p01 + hl[1][1] * p11;
    const double q00 = (a00 * hl[0][0] + a01 * hl[0][1]) + c->r00;
    const double q01 = (a00 * hl[1][0] + a01 * hl[1][1]) + 0.0;
    const double q10 = (a10 * hl[0][0] + a11 * hl[0][1]) + 0.0;
    const double q11 = (a10 * hl[1][0] + a11 * hl[1][1]) + c->r11;
    /* q_obs.try_inverse().unwrap_or(identity)  fs2.rs:203 */
    const double qdet = q00 * q11 - q10 * q01;
    double qi[2][2];
    if (qdet == 0.0) { qi[0][0] = 1.0; qi[0][1] = 0.0; qi[1][0] = 0.0; qi[1][1] = 1.0; }
    else { qi[0][0] = q11 / qdet; qi[0][1] = -q01 / qdet; qi[1][0] = -q10 / qdet; qi[1][1] = q00 / qdet; }
    /* p_pred.try_inverse().unwrap_or(identity * 1e-6)  fs2.rs:205 */
    m3 ppi;
    if (!m3_try_inverse(&p_pred, &ppi))
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) ppi.m[i][j] = (i == j ? 1.0 : 0.0) * 1e-6;
    /* h_pose_t * q_obs_inv * h_pose  fs2.rs:206: (3x2 . 2x2) . 2x3 */
    double hq[3][2];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 2; ++j) hq[i][j] = hp[0][i] * qi[0][j] + hp[1][i] * qi[1][j];
    m3 p_post_inv;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) p_post_inv.m[i][j] = ppi.m[i][j] + (hq[i][0] * hp[0][j] + hq[i][1] * hp[1][j]);
    m3 p_post;
    if (!m3_try_inverse(&p_post_inv, &p_post)) p_post = p_pred;                /* fs2.rs:207 */
    /* posterior mean fs2.rs:210-213 */
    const double zp0 = d;                                                      /* observation_model fs2.rs:122-128 at x_pred */
    const double zp1 = orc_fs_normalize_angle(M_ATAN2(dy, dx) - x_pred[2]);
    const double in0 = z0 - zp0, in1 = orc_fs_normalize_angle(z1 - zp1);
    /* ((p_post * h_pose_t) * q_obs_inv) * innovation */
    double ph[3][2];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 2; ++j) {
            double acc = p_post.m[i][0] * hp[j][0];
            acc = p_post.m[i][1] * hp[j][1] + acc;
            acc = p_post.m[i][2] * hp[j][2] + acc;
            ph[i][j] = acc;
        }
    for (int i = 0; i < 3; ++i) {
        const double k0 = ph[i][0] * qi[0][0] + ph[i][1] * qi[1][0], k1 = ph[i][0] * qi[0][1] + ph[i][1] * qi[1][1];
        mean[i] = x_pred[i] + (k0 * in0 + k1 * in1);
    }
    *cov = p_post;
}

/* sample_pose_with_rng fs2.rs:219-239 */
static void sample_pose(const double mean[3], const m3* cov, double n0, double n1, double n2, double out[3]) {
    m3 l;
    if (!m3_cholesky_l(cov, &l)) {                                            /* fs2.rs:227-233 */
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) l.m[i][j] = 0.0;
        for (int i = 0; i < 3; ++i) { const double c = cov->m[i][i]; l.m[i][i] = sqrt(c > 0.0 ? c : 0.0); }   /* f64::max(NaN, 0) = 0 */
    }
    const double nz[3] = { n0, n1, n2 };
    for (int i = 0; i < 3; ++i) {
        double acc = l.m[i][0] * nz[0];
        acc = l.m[i][1] * nz[1] + acc;
        acc = l.m[i][2] * nz[2] + acc;
        out[i] = mean[i] + acc;
    }
}

/* update_landmark_and_weight fs2.rs:242-280; returns the factor the weight is multiplied by */
static double update_landmark_and_weight(orc_fs* f, size_t i, double z0, double z1, size_t lm_id) {
    lm_t* L = &f->lm[i * f->m + lm_id];
    const double px = f->x[i], py = f->y[i], pyaw = f->yaw[i];
    if (!(L->c00 < 100.0)) {                                                  /* fs2.rs:250-256 */
        L->x = px + z0 * M_COS(pyaw + z1);
        L->y = py + z0 * M_SIN(pyaw + z1);
        L->c00 = 10.0; L->c01 = 0.0; L->c10 = 0.0; L->c11 = 10.0;
        return 1.0;
    }
    const double dx = L->x - px, dy = L->y - py;
    const double d = sqrt(dx * dx + dy * dy);
    const double zp1 = orc_fs_normalize_angle(M_ATAN2(dy, dx) - pyaw);
    const double y0 = z0 - d, y1 = orc_fs_normalize_angle(z1 - zp1);          /* fs2.rs:259 */
    const double d2 = dx * dx + dy * dy;
    const double dd = sqrt(d2);
    const double h00 = dx / dd, h01 = dy / dd, h10 = -dy / d2, h11 = dx / d2; /* fs2.rs:261 */
    const double p00 = L->c00, p01 = L->c01, p10 = L->c10, p11 = L->c11;
    const double a00 = h00 * p00 + h01 * p10, a01 = h00 * p01 + h01 * p11;
    const double a10 = h10 * p00 + h11 * p10, a11 = h10 * p01 + h11 * p11;
    const double s00 = (a00 * h00 + a01 * h01) + f->cfg.r00;                  /* fs2.rs:262 */
    const double s01 = (a00 * h10 + a01 * h11) + 0.0;
    const double s10 = (a10 * h00 + a11 * h01) + 0.0;
    const double s11 = (a10 * h10 + a11 * h11) + f->cfg.r11;
    const double det = s00 * s11 - s10 * s01;
    double i00, i01, i10, i11;
    if (det == 0.0) { i00 = 1.0; i01 = 0.0; i10 = 0.0; i11 = 1.0; }           /* fs2.rs:263 */
    else { i00 = s11 / det; i01 = -s01 / det; i10 = -s10 / det; i11 = s00 / det; }
    const double b00 = p00 * h00 + p01 * h01, b01 = p00 * h10 + p01 * h11;    /* fs2.rs:264 */
    const double b10 = p10 * h00 + p11 * h01, b11 = p10 * h10 + p11 * h11;
    const double k00 = b00 * i00 + b01 * i10, k01 = b00 * i01 + b01 * i11;
    const double k10 = b10 * i00 + b11 * i10, k11 = b10 * i01 + b11 * i11;
    L->x += k00 * y0 + k01 * y1;                                              /* fs2.rs:267-269 */
    L->y += k10 * y0 + k11 * y1;
    const double m00 = 1.0 - (k00 * h00 + k01 * h10), m01 = 0.0 - (k00 * h01 + k01 * h11);
    const double m10 = 0.0 - (k10 * h00 + k11 * h10), m11 = 1.0 - (k10 * h01 + k11 * h11);
    L->c00 = m00 * p00 + m01 * p10; L->c01 = m00 * p01 + m01 * p11;           /* fs2.rs:270 */
    L->c10 = m10 * p00 + m11 * p10; L->c11 = m10 * p01 + m11 * p11;
    const double det_s = s00 * s11 - s10 * s01;                               /* fs2.rs:273-279 */
    if (det_s > 0.0) {
        const double t0 = y0 * i00 + y1 * i10, t1 = y0 * i01 + y1 * i11;
        const double mahal = t0 * y0 + t1 * y1;
        return M_EXP(-0.5 * mahal) / (2.0 * PFC_PI * sqrt(det_s));
    }
    return 1e-10;
}

/* the loop body of fastslam2_update_with_rng fs2.rs:339-366 for particle i */
void orc_fs2_particle_(orc_fs* f, size_t i, const double u[2], const orc_fs_obs* z, size_t k, double n0, double n1, double n2) {
    double pose[3] = { f->x[i], f->y[i], f->yaw[i] }, np[3];
    if (k > 0) {                                                              /* fs2.rs:341-346 */
        double mean[3]; m3 cov;
        compute_proposal(&f->cfg, pose, u, z[0].d, z[0].angle, &f->lm[i * f->m + (size_t)z[0].lm_id], mean, &cov);
        sample_pose(mean, &cov, n0, n1, n2, np);
    } else {                                                                  /* fs2.rs:347-356 */
        const double un[2] = { u[0] + n0 * sqrt(f->cfg.q00), u[1] + n1 * sqrt(f->cfg.q11) };
        motion_model(pose, un, f->cfg.dt, np);
    }
    f->x[i] = np[0]; f->y[i] = np[1]; f->yaw[i] = orc_fs_normalize_angle(np[2]);   /* set_pose fs2.rs:77-81 */
    for (size_t j = 0; j < k; ++j)                                            /* fs2.rs:359-364 */
        if (z[j].lm_id < f->m) f->w[i] *= update_landmark_and_weight(f, i, z[j].d, z[j].angle, (size_t)z[j].lm_id);
}

void orc_fs2_compute_proposal(const orc_fs_config* c, const double pose3[3], const double u[2], double z_d, double z_angle,
                              const double lm6[6], double mean3[3], double cov9[9]) {
    lm_t L = { lm6[0], lm6[1], lm6[2], lm6[3], lm6[4], lm6[5] };
    m3 cov;
    compute_proposal(c, pose3, u, z_d, z_angle, &L, mean3, &cov);
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) cov9[3 * i + j] = cov.m[i][j];
}
void orc_fs2_sample_pose(const double mean3[3], const double cov9[9], const double n3[3], double out3[3]) {
    m3 cov;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) cov.m[i][j] = cov9[3 * i + j];
    sample_pose(mean3, &cov, n3[0], n3[1], n3[2], out3);
}
