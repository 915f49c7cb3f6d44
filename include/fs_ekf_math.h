/*
 * fs_ekf_math.h — update_landmark (fs1.rs:140-183) for one (particle, observation) pair, shared by the CUDA kernels and
 * by a host test (tests/host/ekf_math_test.c).  fs1.rs = crates/rust_robotics_slam/src/fastslam1.rs.
 *
 *   fs_update_landmark       the CONTRACT: every operation in the reference's order, IEEE f64, the libm of
 *                            pf_contract_math.h.  Bit-identical to oracle/fs1_oracle.c (tests compare them).
 *   fs_update_landmark_fast  the same function on the common domain, written for the FP64 pipe of sm_100a:
 *                            no data-dependent branch, no range guard inside the 13 divisions, comparisons done on the
 *                            integer unit.  It returns 0 ("not applicable") whenever an operand leaves the domain on which
 *                            the unguarded sequences are proven equal to the contract (zeros, |x| outside [2^-498, 2^498),
 *                            atan2 special cases, angles beyond 3 pi, exp argument beyond +-340, singular S, first
 *                            observation of the landmark); the caller then evaluates fs_update_landmark on the same
 *                            inputs.  Wherever it returns 1 the results are bit-identical to the contract
 *                            (tests/test_ekf_math_host.py: random + adversarial pairs on the CPU; GPU parity tests).
 */
#ifndef FS_EKF_MATH_H
#define FS_EKF_MATH_H

#include "pf_contract_math.h"

typedef struct { double x, y, c00, c01, c10, c11; } FsLm;

/* normalize_angle fs1.rs:80-89.  The reference loops without bound (and would spin forever on +-inf); the guard caps the
 * loop at 2^22 turns, i.e. |angle| up to ~2.6e7 rad behaves exactly like the reference. */
PFC_HD double fs_normalize_angle(double a) {
    int guard = 0;
    while (a > PFC_PI && guard < (1 << 22)) { a -= 2.0 * PFC_PI; ++guard; }
    while (a < -PFC_PI && guard < (1 << 23)) { a += 2.0 * PFC_PI; ++guard; }
    return a;
}

/* update_landmark fs1.rs:140-183; returns the likelihood factor (1.0 when the weight is left untouched).  L is updated
 * in place; *wrote_cov tells the caller whether the covariance changed (branch A leaves it alone, fs1.rs:144-149) — the
 * weight is multiplied only then (fs1.rs:181 sits inside the EKF branch). */
/* variant 1 = FastSLAM 1.0 (fs1.rs), variant 2 = FastSLAM 2.0's update_landmark_and_weight (fs2.rs:242-280, fs2.rs =
 * crates/rust_robotics_slam/src/fastslam2.rs): the same EKF text; a landmark is fresh when NOT `cov00 < 100` (fs2.rs:49-51)
 * rather than when `cov00 > 100` (fs1.rs:144), a fresh landmark also gets cov = 10 I (fs2.rs:254), and a non-positive det S
 * multiplies the weight by 1e-10 (fs2.rs:278) instead of leaving it alone (fs1.rs:178). */
PFC_HD double fs_update_landmark_v(FsLm* Lp, double px, double py, double pyaw, double z0, double z1,
                                   double r00, double r11, int* wrote_cov, int variant) {
    FsLm L = *Lp;
    if (variant == 2 ? !(L.c00 < 100.0) : (L.c00 > 100.0)) {   /* first observation of this landmark */
        double s, c;
        pfc_sincos(pyaw + z1, &s, &c);
        Lp->x = px + z0 * c;
        Lp->y = py + z0 * s;
        if (variant == 2) { Lp->c00 = 10.0; Lp->c01 = 0.0; Lp->c10 = 0.0; Lp->c11 = 10.0; }
        *wrote_cov = variant == 2 ? 1 : 0;
        return 1.0;
    }
    *wrote_cov = 1;
    /* observation_model fs1.rs:92-99 */
    double dx = L.x - px, dy = L.y - py;
    double d2 = dx * dx + dy * dy;
    double d = sqrt(d2);
    double zp1 = fs_normalize_angle(pfc_atan2(dy, dx) - pyaw);
    double y0 = z0 - d, y1 = fs_normalize_angle(z1 - zp1);     /* innovation fs1.rs:155 */
    /* compute_jacobian fs1.rs:102-110: four IEEE quotients over two denominators */
    const pfc_rcp_t rd = pfc_rcp_make(d), rd2 = pfc_rcp_make(d2);
    double h00 = pfc_div_by(dx, rd), h01 = pfc_div_by(dy, rd), h10 = pfc_div_by(-dy, rd2), h11 = pfc_div_by(dx, rd2);
    double p00 = L.c00, p01 = L.c01, p10 = L.c10, p11 = L.c11;
    /* S = H P H^T + R  fs1.rs:161 */
    double a00 = h00 * p00 + h01 * p10, a01 = h00 * p01 + h01 * p11;
    double a10 = h10 * p00 + h11 * p10, a11 = h10 * p01 + h11 * p11;
    double s00 = (a00 * h00 + a01 * h01) + r00;
    double s01 = (a00 * h10 + a01 * h11) + 0.0;
    double s10 = (a10 * h00 + a11 * h01) + 0.0;
    double s11 = (a10 * h10 + a11 * h11) + r11;
    /* try_inverse().unwrap_or(identity) fs1.rs:164 */
    double det = s00 * s11 - s10 * s01;
    double i00, i01, i10, i11;
    if (det == 0.0) { i00 = 1.0; i01 = 0.0; i10 = 0.0; i11 = 1.0; }
    else {
        const pfc_rcp_t rdet = pfc_rcp_make(det);
        i00 = pfc_div_by(s11, rdet); i01 = pfc_div_by(-s01, rdet); i10 = pfc_div_by(-s10, rdet); i11 = pfc_div_by(s00, rdet);
    }
    /* K = P H^T S^-1 fs1.rs:165 */
    double b00 = p00 * h00 + p01 * h01, b01 = p00 * h10 + p01 * h11;
    double b10 = p10 * h00 + p11 * h01, b11 = p10 * h10 + p11 * h11;
    double k00 = b00 * i00 + b01 * i10, k01 = b00 * i01 + b01 * i11;
    double k10 = b10 * i00 + b11 * i10, k11 = b10 * i01 + b11 * i11;
    L.x = L.x + (k00 * y0 + k01 * y1);                         /* fs1.rs:168-170 */
    L.y = L.y + (k10 * y0 + k11 * y1);
    /* P = (I - K H) P fs1.rs:173-174 (not symmetrised) */
    double m00 = 1.0 - (k00 * h00 + k01 * h10), m01 = 0.0 - (k00 * h01 + k01 * h11);
    double m10 = 0.0 - (k10 * h00 + k11 * h10), m11 = 1.0 - (k10 * h01 + k11 * h11);
    L.c00 = m00 * p00 + m01 * p10; L.c01 = m00 * p01 + m01 * p11;
    L.c10 = m10 * p00 + m11 * p10; L.c11 = m10 * p01 + m11 * p11;
    *Lp = L;
    /* likelihood fs1.rs:177-182 */
    double det_s = s00 * s11 - s10 * s01;
    if (det_s > 0.0) {
        double t0 = y0 * i00 + y1 * i10, t1 = y0 * i01 + y1 * i11;
        double mahal = t0 * y0 + t1 * y1;
        return PFC_DIV(pfc_exp(-0.5 * mahal), 2.0 * PFC_PI * sqrt(det_s));
    }
    return variant == 2 ? 1e-10 : 1.0;
}
PFC_HD double fs_update_landmark(FsLm* Lp, double px, double py, double pyaw, double z0, double z1,
                                 double r00, double r11, int* wrote_cov) {
    return fs_update_landmark_v(Lp, px, py, pyaw, z0, z1, r00, r11, wrote_cov, 1);
}

/* ------------------------------------------------------------------------------------------------------------------ */
/* fast form                                                                                                          */
/* ------------------------------------------------------------------------------------------------------------------ */
PFC_HD int fsm_hi(double x) {
#if defined(__CUDA_ARCH__)
    return __double2hiint(x);
#else
    return (int)(pfc_d2u(x) >> 32);
#endif
}
PFC_HD double fsm_flip_sign_if(double x, int neg) {      /* x or -x, on the integer unit */
#if defined(__CUDA_ARCH__)
    return __hiloint2double(__double2hiint(x) ^ (neg ? (int)0x80000000u : 0), __double2loint(x));
#else
    return pfc_u2d(pfc_d2u(x) ^ ((uint64_t)(neg ? 1 : 0) << 63));
#endif
}
/* 1 unless |x| lies in [2^-498, 2^498) (inside the (1e-150, 1e150) window of pfc_div_by); zero, inf and NaN are outside */
PFC_HD unsigned fsm_out(double x) { return ((unsigned)(fsm_hi(x) & 0x7ff00000) - 0x20D00000u) > 0x3E400000u ? 1u : 0u; }
/* correctly rounded reciprocal: the device instruction sequence / the IEEE quotient (same value) */
PFC_HD double fsm_rcp(double b) {
#if defined(__CUDA_ARCH__)
    return __drcp_rn(b);
#else
    return 1.0 / b;
#endif
}
/* RN(a/b) from y = RN(1/b), valid for operands inside the window (pf_contract_math.h, Division) */
PFC_HD double fsm_div(double a, double b, double y) {
    double q0 = a * y;
    double r0 = fma(-b, q0, a);
    double q1 = fma(r0, y, q0);
    double r1 = fma(-b, q1, a);
    return fma(r1, y, q1);
}
/* normalize_angle for |a| < 9 (< 3 pi): at most one turn each way, selected without a branch.  After a - 2pi with
 * a in (pi, 3pi] the result is exact and > -pi, so the second loop of the reference does not fire either. */
PFC_HD double fsm_wrap(double a, unsigned* bad) {
    *bad |= (unsigned)(fsm_hi(a) & 0x7fffffff) >= 0x40220000u ? 1u : 0u;       /* |a| >= 9.0 or NaN */
    double lo = a - 2.0 * PFC_PI, hi = a + 2.0 * PFC_PI;
    a = a > PFC_PI ? lo : a;
    return a < -PFC_PI ? hi : a;
}

/* ---- the stages below process W pairs in lockstep: every statement is issued for all W pairs before the next one, so that
 * the independent dependency chains of the pairs interleave in the instruction stream (the reciprocal / square-root
 * intrinsics end basic blocks; a pair-after-pair formulation would serialise the pairs) ---- */
#if defined(__CUDACC__)
#define FSM_VV _Pragma("unroll") for (int q = 0; q < W; ++q)
#else
#define FSM_VV for (int q = 0; q < W; ++q)
#endif

/* polynomial part of pfc_atan on the reduced argument t, recombined with (hi, lo) */
PFC_HD double fsm_atan_poly(double t, double hi, double lo) {
    const double z = t * t, w = z * z;
    const double s1 = z * fma(w, fma(w, fma(w, fma(w, fma(w, PFC_K(AT10), PFC_K(AT8)), PFC_K(AT6)), PFC_K(AT4)), PFC_K(AT2)), PFC_K(AT0));
    const double s2 = w * fma(w, fma(w, fma(w, fma(w, PFC_K(AT9), PFC_K(AT7)), PFC_K(AT5)), PFC_K(AT3)), PFC_K(AT1));
    return hi - ((t * (s1 + s2) - lo) - t);
}

/* pfc_exp(x) for |x| <= 340 (no special case can fire, and the result stays inside the division window) */
PFC_HD double fsm_exp(double x, unsigned* bad) {
    *bad |= (unsigned)(fsm_hi(x) & 0x7fffffff) > 0x40754000u ? 1u : 0u;       /* |x| > 340 or NaN */
    double kf = floor(fma(x, PFC_K(LOG2E), 0.5));
    double r = fma(-kf, PFC_K(LN2_HI), x);
    r = fma(-kf, PFC_K(LN2_LO), r);
    double q = PFC_K(E13);
    q = fma(q, r, PFC_K(E12)); q = fma(q, r, PFC_K(E11)); q = fma(q, r, PFC_K(E10)); q = fma(q, r, PFC_K(E9));
    q = fma(q, r, PFC_K(E8)); q = fma(q, r, PFC_K(E7)); q = fma(q, r, PFC_K(E6)); q = fma(q, r, PFC_K(E5));
    q = fma(q, r, PFC_K(E4)); q = fma(q, r, PFC_K(E3)); q = fma(q, r, 0.5);
    double t = fma(r * r, q, r);
    double y = 1.0 + t;
    const int k = (int)kf, k1 = k >> 1, k2 = k - k1;
    return (y * pfc_pow2i(k1)) * pfc_pow2i(k2);
}

/* W pairs at once.  ok[q] = 1: L[q], lik[q] hold the contract's results; ok[q] = 0: L[q] untouched, the caller runs
 * fs_update_landmark for that pair.  The atan2 follows pfc_atan2 / pfc_atan on their common domain: both operands inside
 * the window, exponents at most 60 apart, quotient >= 2^-27, reduced numerator inside the window. */
#if defined(__cplusplus)
template <int W>
#else
#define W 1
#endif
PFC_HD void fs_update_landmark_fastw(FsLm* L, const double* px, const double* py, const double* pyaw, double z0, double z1,
                                     double r00, double r11, double* lik, int* ok) {
    unsigned bad[W];
    double dx[W], dy[W], d2[W], d[W], ax[W], yax[W], yd[W], yd2[W];
    FSM_VV {
        bad[q] = !(L[q].c00 < 100.0) ? 1u : 0u;                 /* a fresh landmark under either variant's test (or NaN) */
        dx[q] = L[q].x - px[q]; dy[q] = L[q].y - py[q];
        d2[q] = dx[q] * dx[q] + dy[q] * dy[q];
        bad[q] |= fsm_out(dx[q]) | fsm_out(dy[q]) | fsm_out(d2[q]);
        ax[q] = fabs(dx[q]);
    }
    FSM_VV d[q] = sqrt(d2[q]);
    FSM_VV yax[q] = fsm_rcp(ax[q]);
    FSM_VV yd[q] = fsm_rcp(d[q]);
    FSM_VV yd2[q] = fsm_rcp(d2[q]);
    /* atan2: q = |dy| / |dx|, interval of pfc_atan (thresholds have zero low words: compare the high words).
     * (2q-1)/(2+q), (q-1)/(q+1), (q-1.5)/(1+1.5q) are (q-c)/(1+c*q) with c = 0.5, 1, 1.5 scaled by exact powers of two and
     * c = 0 gives q/1: identical quotients bit for bit.  The last interval is -1/q. */
    double num[W], den[W], ahi[W], alo[W], yden[W];
    FSM_VV {
        const int hy = fsm_hi(dy[q]), hx = fsm_hi(dx[q]);
        bad[q] |= (unsigned)(((hy >> 20) & 0x7ff) - ((hx >> 20) & 0x7ff) + 60) > 120u ? 1u : 0u;
        const double qq = fsm_div(fabs(dy[q]), ax[q], yax[q]);
        const int hq = fsm_hi(qq);
        bad[q] |= hq < 0x3E400000 ? 1u : 0u;                    /* q < 2^-27: pfc_atan returns its argument */
        const int id = (hq >= 0x3FDC0000) + (hq >= 0x3FE60000) + (hq >= 0x3FF30000) + (hq >= 0x40038000);
        const double c = 0.5 * (double)(id & 3);
        double n_ = qq - c, d_ = 1.0 + c * qq, h_ = 0.0, l_ = 0.0;
        if (id == 1) { h_ = 4.63647609000806093515e-01; l_ = 2.26987774529616870924e-17; }
        if (id == 2) { h_ = 7.85398163397448278999e-01; l_ = 3.06161699786838301793e-17; }
        if (id == 3) { h_ = 9.82793723247329054082e-01; l_ = 1.39033110312309984516e-17; }
        if (id == 4) { h_ = 1.57079632679489655800e+00; l_ = 6.12323399573676603587e-17; n_ = -1.0; d_ = qq; }
        bad[q] |= fsm_out(n_);
        num[q] = n_; den[q] = d_; ahi[q] = h_; alo[q] = l_;
    }
    FSM_VV yden[q] = fsm_rcp(den[q]);
    double y0[W], y1[W];
    FSM_VV {
        const double t = fsm_div(num[q], den[q], yden[q]);
        double r = fsm_atan_poly(t, ahi[q], alo[q]);
        /* quadrants of pfc_atan2: x > 0: +-r; x < 0: +-(PI - (r - PI_LO)) */
        const double PI = 3.14159265358979311600e+00, PI_LO = 1.2246467991473531772e-16;
        const double rr = PI - (r - PI_LO);
        r = fsm_hi(dx[q]) < 0 ? rr : r;
        r = fsm_flip_sign_if(r, fsm_hi(dy[q]) < 0);
        const double zp1 = fsm_wrap(r - pyaw[q], &bad[q]);
        y0[q] = z0 - d[q];
        y1[q] = fsm_wrap(z1 - zp1, &bad[q]);
    }
    double h00[W], h01[W], h10[W], h11[W], s00[W], s01[W], s10[W], s11[W], det[W], ydet[W];
    FSM_VV {
        h00[q] = fsm_div(dx[q], d[q], yd[q]); h01[q] = fsm_div(dy[q], d[q], yd[q]);
        h10[q] = fsm_div(-dy[q], d2[q], yd2[q]); h11[q] = fsm_div(dx[q], d2[q], yd2[q]);
        const double p00 = L[q].c00, p01 = L[q].c01, p10 = L[q].c10, p11 = L[q].c11;
        const double a00 = h00[q] * p00 + h01[q] * p10, a01 = h00[q] * p01 + h01[q] * p11;
        const double a10 = h10[q] * p00 + h11[q] * p10, a11 = h10[q] * p01 + h11[q] * p11;
        s00[q] = (a00 * h00[q] + a01 * h01[q]) + r00;
        s01[q] = (a00 * h10[q] + a01 * h11[q]) + 0.0;
        s10[q] = (a10 * h00[q] + a11 * h01[q]) + 0.0;
        s11[q] = (a10 * h10[q] + a11 * h11[q]) + r11;
        det[q] = s00[q] * s11[q] - s10[q] * s01[q];
        bad[q] |= fsm_out(det[q]) | fsm_out(s00[q]) | fsm_out(s01[q]) | fsm_out(s10[q]) | fsm_out(s11[q]);
        bad[q] |= fsm_hi(det[q]) < 0 ? 1u : 0u;                 /* det_s <= 0: the weight is left alone (fs1.rs:178) */
    }
    FSM_VV ydet[q] = fsm_rcp(det[q]);
    FsLm N[W];
    double e[W], sd[W], den2[W], yden2[W];
    FSM_VV {
        const double i00 = fsm_div(s11[q], det[q], ydet[q]), i01 = fsm_div(-s01[q], det[q], ydet[q]);
        const double i10 = fsm_div(-s10[q], det[q], ydet[q]), i11 = fsm_div(s00[q], det[q], ydet[q]);
        const double p00 = L[q].c00, p01 = L[q].c01, p10 = L[q].c10, p11 = L[q].c11;
        const double b00 = p00 * h00[q] + p01 * h01[q], b01 = p00 * h10[q] + p01 * h11[q];
        const double b10 = p10 * h00[q] + p11 * h01[q], b11 = p10 * h10[q] + p11 * h11[q];
        const double k00 = b00 * i00 + b01 * i10, k01 = b00 * i01 + b01 * i11;
        const double k10 = b10 * i00 + b11 * i10, k11 = b10 * i01 + b11 * i11;
        N[q].x = L[q].x + (k00 * y0[q] + k01 * y1[q]);
        N[q].y = L[q].y + (k10 * y0[q] + k11 * y1[q]);
        const double m00 = 1.0 - (k00 * h00[q] + k01 * h10[q]), m01 = 0.0 - (k00 * h01[q] + k01 * h11[q]);
        const double m10 = 0.0 - (k10 * h00[q] + k11 * h10[q]), m11 = 1.0 - (k10 * h01[q] + k11 * h11[q]);
        N[q].c00 = m00 * p00 + m01 * p10; N[q].c01 = m00 * p01 + m01 * p11;
        N[q].c10 = m10 * p00 + m11 * p10; N[q].c11 = m10 * p01 + m11 * p11;
        const double t0 = y0[q] * i00 + y1[q] * i10, t1 = y0[q] * i01 + y1[q] * i11;
        const double mahal = t0 * y0[q] + t1 * y1[q];
        e[q] = fsm_exp(-0.5 * mahal, &bad[q]);
    }
    FSM_VV sd[q] = sqrt(det[q]);
    FSM_VV { den2[q] = 2.0 * PFC_PI * sd[q]; }
    FSM_VV yden2[q] = fsm_rcp(den2[q]);
    FSM_VV {
        const double lk = fsm_div(e[q], den2[q], yden2[q]);
        ok[q] = bad[q] ? 0 : 1;
        if (!bad[q]) { L[q] = N[q]; lik[q] = lk; }
    }
}
#if !defined(__cplusplus)
#undef W
#endif

#endif /* FS_EKF_MATH_H */
