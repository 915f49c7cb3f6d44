/* Host check of include/fs_ekf_math.h: the branch-free form fs_update_landmark_fast must equal the contract form
 * fs_update_landmark bit for bit wherever it declares itself applicable.  Built with -ffp-contract=off; on the host the
 * correctly rounded reciprocal is 1.0/b and `/` is the IEEE quotient, so this also re-checks the reciprocal + two-fma
 * division on the operand ranges the EKF produces.  Driven by tests/test_ekf_math_host.py. */
#include <stdint.h>
#include <string.h>
#include <math.h>
#include "../../include/fs_ekf_math.h"

#ifdef __cplusplus
#define EXT extern "C"
/* one pair through the W-wide form: pair 0 is the case under test, the other lanes carry a second, unrelated case */
static int fast1(FsLm* B, const double* in, double* lb, const double* in2) {
    FsLm L[2] = { *B, { in2[0], in2[1], in2[2], in2[3], in2[4], in2[5] } };
    double px[2] = { in[6], in2[6] }, py[2] = { in[7], in2[7] }, pyaw[2] = { in[8], in2[8] }, lik[2] = { 0.0, 0.0 };
    int ok[2] = { 0, 0 };
    fs_update_landmark_fastw<2>(L, px, py, pyaw, in[9], in[10], in[11], in[12], lik, ok);
    *B = L[0]; *lb = lik[0];
    return ok[0];
}
#else
#define EXT
static int fast1(FsLm* B, const double* in, double* lb, const double* in2) {
    int ok = 0; (void)in2;
    fs_update_landmark_fastw(B, &in[6], &in[7], &in[8], in[9], in[10], in[11], in[12], lb, &ok);
    return ok;
}
#endif

static uint64_t sm64(uint64_t* s) { uint64_t z = (*s += 0x9E3779B97F4A7C15ull); z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }
static double u01(uint64_t* s) { return (double)(sm64(s) >> 11) * 1.1102230246251565e-16; }
static double sym(uint64_t* s, double a) { return (2.0 * u01(s) - 1.0) * a; }

/* kind 0: bench-like geometry (landmark 1..25 m away, cov ~ 0.01..10, small innovations)
 * kind 1: wide (distances 1e-6..1e6, covariances 1e-8..1e8, arbitrary angles)
 * kind 2: adversarial (axis-aligned, tiny / huge operands, zeros, non-PD covariances, first observations) */
static void make_case(uint64_t* s, int kind, double in[14]) {
    double px, py, pyaw, lx, ly, c00, c01, c10, c11, z0, z1, r00 = 0.5, r11 = 0.0305;
    px = sym(s, 150.0); py = sym(s, 150.0); pyaw = sym(s, 3.14159);
    if (kind == 0) {
        double dist = 1.0 + 24.0 * u01(s), ang = sym(s, 3.14159);
        lx = px + dist * cos(ang); ly = py + dist * sin(ang);
        double sc = pow(10.0, -2.0 + 3.0 * u01(s));
        c00 = sc * (0.5 + u01(s)); c11 = sc * (0.5 + u01(s)); c01 = sym(s, 0.3) * sc; c10 = c01 + sym(s, 1e-6) * sc;
        z0 = dist + sym(s, 1.0); z1 = ang - pyaw + sym(s, 0.3);
        while (z1 > 3.141592653589793) z1 -= 6.283185307179586;
        while (z1 < -3.141592653589793) z1 += 6.283185307179586;
    } else if (kind == 1) {
        double dist = pow(10.0, sym(s, 6.0)), ang = sym(s, 3.14159);
        lx = px + dist * cos(ang); ly = py + dist * sin(ang);
        double sc = pow(10.0, sym(s, 8.0));
        c00 = sc * (0.5 + u01(s)); c11 = sc * (0.5 + u01(s)); c01 = sym(s, 0.7) * sc; c10 = sym(s, 0.7) * sc;
        z0 = dist * (0.5 + u01(s)); z1 = sym(s, 3.14159);
        pyaw = sym(s, 8.0);
    } else {
        int v = (int)(sm64(s) % 12);
        double dist = 5.0;
        lx = px + dist; ly = py; c00 = 1.0; c11 = 1.0; c01 = 0.0; c10 = 0.0; z0 = 5.0; z1 = 0.0;
        if (v == 0) { ly = py; }                                     /* dy == 0 */
        if (v == 1) { lx = px; ly = py + 3.0; }                      /* dx == 0 */
        if (v == 2) { lx = px; ly = py; }                            /* on top of the landmark */
        if (v == 3) { ly = py + 1e-30; }                             /* huge exponent gap */
        if (v == 4) { c00 = 1000.0; c11 = 1000.0; }                  /* first observation: branch A */
        if (v == 5) { c00 = -3.0; c11 = 2.0; c01 = 5.0; c10 = -4.0; }/* not PD */
        if (v == 6) { c00 = 1e-300; c11 = 1e-300; r00 = 0.0; r11 = 0.0; }
        if (v == 7) { z0 = 1e4; }                                    /* huge innovation: exp underflows */
        if (v == 8) { pyaw = 100.0; }                                /* many turns */
        if (v == 9) { lx = px + 1e-170; ly = py + 1e-170; }          /* outside the division window */
        if (v == 10) { lx = px + 3.0; ly = py + 3.0 * 0.4375; }      /* atan interval edges */
        if (v == 11) { lx = px + 2.0; ly = py + 2.0 * 2.4375; c00 = 100.0; }
    }
    in[0] = lx; in[1] = ly; in[2] = c00; in[3] = c01; in[4] = c10; in[5] = c11;
    in[6] = px; in[7] = py; in[8] = pyaw; in[9] = z0; in[10] = z1; in[11] = r00; in[12] = r11; in[13] = 0.0;
}

/* out4: [0] cases, [1] cases on the fast path, [2] mismatches among those, [3] first mismatching case index + 1 */
EXT void ekf_math_compare(uint64_t seed, int kind, uint64_t n, uint64_t out4[4]) {
    uint64_t s = seed;
    out4[0] = n; out4[1] = out4[2] = out4[3] = 0;
    for (uint64_t i = 0; i < n; ++i) {
        double in[14], in2[14];
        make_case(&s, kind, in);
        { uint64_t s2 = s ^ 0x5555; make_case(&s2, (int)(i % 3), in2); }
        FsLm A = { in[0], in[1], in[2], in[3], in[4], in[5] }, B = A;
        int wrote = 0;
        double la = fs_update_landmark(&A, in[6], in[7], in[8], in[9], in[10], in[11], in[12], &wrote);
        double lb = 0.0;
        int ok = fast1(&B, in, &lb, in2);
        if (!ok) {                                  /* must leave the landmark untouched */
            FsLm C = { in[0], in[1], in[2], in[3], in[4], in[5] };
            if (memcmp(&B, &C, sizeof(B)) != 0) { out4[2]++; if (!out4[3]) out4[3] = i + 1; }
            continue;
        }
        out4[1]++;
        if (!wrote || memcmp(&A, &B, sizeof(A)) != 0 || memcmp(&la, &lb, 8) != 0) { out4[2]++; if (!out4[3]) out4[3] = i + 1; }
    }
}

/* one explicit pair through both forms (for the cross-check against oracle/fs1_oracle.c done in Python) */
EXT int ekf_math_one(const double in[13], double out_ref[7], double out_fast[7]) {
    FsLm A = { in[0], in[1], in[2], in[3], in[4], in[5] }, B = A;
    int wrote = 0;
    double la = fs_update_landmark(&A, in[6], in[7], in[8], in[9], in[10], in[11], in[12], &wrote);
    double lb = 1.0;
    int ok = fast1(&B, in, &lb, in);
    memcpy(out_ref, &A, 48); out_ref[6] = wrote ? la : 1.0;
    memcpy(out_fast, &B, 48); out_fast[6] = lb;
    return ok | (wrote << 1);
}
