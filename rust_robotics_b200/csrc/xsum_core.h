/*
 * xsum_core.h — element-level logic of the EXACT SEQUENTIAL-ORDER SUM / SCAN ("xsum").
 *
 * Problem.  The reference accumulates weights left to right in f64:
 *     sum_w:      pf.rs:427, mcl.rs:395, fs1.rs:197        (normalisation divisor)
 *     sum_w^2:    pf.rs:417, fs1.rs:187                     (N_eff gate)
 *     cum_sum:    pf.rs:448-453, mcl.rs:328-333, fs1.rs:213-216   (resampling CDF)
 *     r += 1/n:   fs1.rs:230                                (systematic comb)
 * FP addition is not associative, so a tree / look-back scan gives different low bits, and a resample
 * index flips whenever a draw lands inside the rounding band of a CDF edge.  To get BIT-EXACT indices the
 * device must reproduce c_i = fl(c_{i-1} + v_i) exactly, in parallel.
 *
 * Idea.  While the running sum c stays inside one "level" L (a binade [2^L, 2^(L+1)), or the union of the
 * subnormals and the first normal binade, which share ulp 2^-1074) its ulp u = 2^(L-52) is constant and
 * c = m*u with an integer mantissa m.  Then
 *       fl(c + v) = (m + q + round_bit) * u,   v = q*u + f, 0 <= f < u,
 *       round_bit = [f > u/2] + [f == u/2 and (m + q) odd]          (round-half-to-even)
 * Unless f is EXACTLY u/2 (a tie), adding v is the integer increment q + [f > u/2] of the mantissa, independent of the
 * state.  So any run of tie-free elements that provably stays inside one level is summed EXACTLY by an ordinary
 * parallel prefix sum over int64 increments, and is applied to an exact double by adding the increment to its raw
 * bit pattern.  (A tie would need the parity of m; ties are rare — the bits of v below u must be exactly 100...0 —
 * and are simply treated like the other "dirty" elements below.)
 *
 * Which level an element sees is decided from an APPROXIMATE prefix sum a_i (plain parallel scan) with a
 * rigorous margin: |c_i - a_i| <= rel * a_i for every summation order of n non-negative terms, rel =
 * 4(n+64)2^-53.  Elements whose approximate prefix (before or after the add) comes within the margin of a
 * level edge, and ties, are "dirty": they are applied with one genuine FP add, in order, by a short sequential
 * chain (there are ~log2(n) + a few of them).  Everything else is "clean" and goes through the scan.
 */
#ifndef XSUM_CORE_H
#define XSUM_CORE_H

#include "../../include/pf_contract_math.h"

#define XS_EMPTY  (-100000)   /* level of the identity transducer          */
#define XS_BAD    (-100001)   /* composition of two different levels: bug   */

typedef struct { long long inc; int lvl; } xs_t;    /* mantissa increment of a clean run at level lvl */

PFC_HD xs_t xs_identity(void) { xs_t t; t.inc = 0; t.lvl = XS_EMPTY; return t; }

/* level of a non-negative finite double: unbiased exponent, clamped so that subnormals and the first
 * normal binade (same ulp) form one level */
PFC_HD int xs_level(double s) {
    int e = (int)((pfc_d2u(s) >> 52) & 0x7FF) - 1023;
    return e < -1022 ? -1022 : e;
}
PFC_HD double xs_level_lo(int lvl) { return lvl <= -1022 ? 0.0 : pfc_pow2i(lvl); }
PFC_HD double xs_level_hi(int lvl) { return lvl >= 1023 ? pfc_u2d(0x7FF0000000000000ull) : pfc_pow2i(lvl + 1); }

/* increment of "add v" while the running sum is in level `lvl`.  ok=0 if v cannot be a clean element
 * (negative / non-finite / larger than the level / an exact tie). */
PFC_HD xs_t xs_elem(double v, int lvl, int* ok) {
    xs_t t; t.lvl = lvl; t.inc = 0;
    uint64_t bits = pfc_d2u(v);
    int be = (int)((bits >> 52) & 0x7FF);
    if ((bits >> 63) || be == 0x7FF) { *ok = 0; return t; }            /* negative / inf / nan */
    uint64_t frac = bits & 0x000FFFFFFFFFFFFFull;
    uint64_t mant = be ? (frac | 0x0010000000000000ull) : frac;        /* v = mant * 2^(ev-52) */
    int ev = be ? be - 1023 : -1022;
    int sh = lvl - ev;
    if (sh < 0) { *ok = 0; return t; }
    if (sh >= 64) return t;                                            /* v < u/2^11: no effect */
    if (sh == 0) { t.inc = (long long)mant; return t; }                /* exact multiple of u  */
    uint64_t q = mant >> sh;
    uint64_t rem = mant & ((1ull << sh) - 1ull);
    uint64_t half = 1ull << (sh - 1);
    if (rem == half) { *ok = 0; return t; }                            /* tie: needs the parity of the sum -> dirty */
    t.inc = (long long)q + (rem > half ? 1 : 0);
    return t;
}

/* A first, then B */
PFC_HD xs_t xs_compose(const xs_t a, const xs_t b) {
    xs_t r;
    const long long CAP = 1ll << 60;
    long long x = a.inc + b.inc;
    r.inc = x > CAP ? CAP : x;
    r.lvl = (a.lvl == XS_EMPTY) ? b.lvl : ((b.lvl == XS_EMPTY || b.lvl == a.lvl) ? a.lvl : XS_BAD);
    return r;
}

/* apply to an exact running sum.  *ok=0 if the certificate is violated (level mismatch / leaves level). */
PFC_HD double xs_apply(const xs_t t, double s, int* ok) {
    if (t.lvl == XS_EMPTY) return s;
    if (t.lvl == XS_BAD || xs_level(s) != t.lvl) { *ok = 0; return s; }
    uint64_t bits = pfc_d2u(s);
    uint64_t nb = bits + (uint64_t)t.inc;
    /* must stay in level: same exponent field, except the lowest level may move from exp 0 to exp 1 */
    int e0 = (int)(bits >> 52), e1 = (int)(nb >> 52);
    if (t.inc < 0 || t.inc >= (1ll << 53) || !(e1 == e0 || (t.lvl == -1022 && e1 <= 1))) { *ok = 0; return s; }
    return pfc_u2d(nb);
}

/* Classification of one element from its approximate prefix before (a_prev) and after (a_cur) the add.
 * Returns 1 and the transducer when the element is provably clean; 0 when it must take the FP add. */
PFC_HD int xs_classify(double v, double a_prev, double a_cur, double rel, xs_t* t) {
    if (!(v >= 0.0) || !(a_prev >= 0.0) || !(a_cur <= 1.7976931348623157e308)) return 0;
    /* Universal identity: the true running sum is >= c_lo, so its ulp is >= 2^(level(c_lo)-52); an addend
     * strictly below half of that never changes it, whatever level it is really in.  Keeps long tails of
     * negligible weights (weight collapse) out of the dirty chain even right at a level edge. */
    {
        double c_lo = a_prev - a_prev * rel;
        int l_lo = xs_level(c_lo > 0.0 ? c_lo : 0.0);
        double half_ulp = (l_lo - 53 >= -1022) ? pfc_pow2i(l_lo - 53) : 0.0;
        if (v == 0.0 || v < half_ulp) { *t = xs_identity(); return 1; }
    }
    int lvl = xs_level(a_prev);
    double lo = xs_level_lo(lvl), hi = xs_level_hi(lvl);
    double m_prev = a_prev * rel, m_cur = a_cur * rel;
    if (a_prev - m_prev < lo || a_cur - m_cur < lo) return 0;
    if (!(a_prev + m_prev < hi) || !(a_cur + m_cur < hi)) return 0;
    int ok = 1;
    *t = xs_elem(v, lvl, &ok);
    return ok;
}

PFC_HD double xs_margin(unsigned long long n) { return (double)(n + 64ull) * 4.440892098500626e-16; } /* 4(n+64)2^-53 */

#endif
