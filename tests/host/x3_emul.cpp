// Host emulation of the exact sequential-order sum of fs3_post.cuh (element logic: rust_robotics_b200/csrc/x3_core.h),
// in the same pass structure as the kernel: approximate prefixes assembled from 64-value partial sums (tree order), tile
// offsets, per-thread runs of K values; classification; a plain 64-bit prefix sum of the clean increments; the chain over
// the dirty values; emission.  Compared with c_i = fl(c_{i-1} + v_i) by tests/test_x3_host.py.
#include <cstdint>
#include <cstring>
#include <vector>
#include <cmath>
#include "../../rust_robotics_b200/csrc/x3_core.h"

static double tree_sum(const double* p, size_t n) {           // pairwise, like a warp shuffle tree
    if (n == 0) return 0.0;
    if (n == 1) return p[0];
    size_t h = n / 2;
    return tree_sum(p, h) + tree_sum(p + h, n - h);
}

// scale != 0: the values are v_i / scale (one rounding each) and the tile offsets are derived as (sum of v partials) / scale
// stats: [0] dirty values  [1] certificate failures  [2] emit failures
extern "C" int x3_emul_scan(const double* vin, size_t n, size_t NT, size_t K, double scale, double* out, double* total, long long* stats) {
    const size_t T = NT * K, nt = (n + T - 1) / T;
    const unsigned m32 = x3_margin32(n);
    std::vector<double> v(nt * T, 0.0);
    for (size_t i = 0; i < n; ++i) v[i] = scale != 0.0 ? vin[i] / scale : vin[i];
    // partial sums of the ORIGINAL values per 64 (what the EKF kernel's epilogue publishes)
    std::vector<double> raw(nt * T, 0.0);
    for (size_t i = 0; i < n; ++i) raw[i] = vin[i];
    std::vector<double> part(nt * T / 64);
    for (size_t p = 0; p < part.size(); ++p) part[p] = tree_sum(&raw[p * 64], 64);
    struct Ent { unsigned long long P; double v; int lvl; };
    std::vector<Ent> ent;
    std::vector<unsigned long long> Pel(nt * T);               // P before each element (global)
    std::vector<char> dirty(nt * T, 0);
    std::vector<unsigned long long> incs(nt * T, 0);
    std::vector<int> lvls(nt * T, 0);
    unsigned long long P = 0;
    stats[0] = stats[1] = stats[2] = 0;
    for (size_t b = 0; b < nt; ++b) {
        double toff = tree_sum(part.data(), b * T / 64);
        if (scale != 0.0) toff = toff / scale;
        // thread sums and their exclusive prefix (tree-ish: sequential over warps of 32, tree inside)
        std::vector<double> ts(NT, 0.0);
        for (size_t t = 0; t < NT; ++t) { double s = 0.0; for (size_t k = 0; k < K; ++k) s += v[b * T + t * K + k]; ts[t] = s; }
        for (size_t t = 0; t < NT; ++t) {
            size_t w = t / 32;
            double woff = 0.0;
            for (size_t ww = 0; ww < w; ++ww) woff += tree_sum(&ts[ww * 32], 32);
            double excl = woff + tree_sum(&ts[w * 32], t % 32);
            double a = toff + excl;
            const int e_run = x3_interior(a, a + ts[t], m32);       // the kernel's per-thread shortcut: the whole run inside one binade
            for (size_t k = 0; k < K; ++k) {
                size_t i = b * T + t * K + k;
                double a1 = a + v[i];
                unsigned long long inc; int lvl;
                int d;
                if (e_run >= 0) { d = x3_classify_at(v[i], e_run, &inc); lvl = e_run; }
                else d = x3_classify(v[i], a, a1, m32, &inc, &lvl);
                Pel[i] = P; dirty[i] = (char)d; incs[i] = inc; lvls[i] = lvl;
                if (d) { ent.push_back({P, v[i], lvl}); stats[0]++; }
                else P += inc;
                a = a1;
            }
        }
    }
    // chain
    std::vector<double> after(ent.size());
    double s = 0.0; unsigned long long prev = 0;
    int ok = 1;
    for (size_t k = 0; k < ent.size(); ++k) {
        s = x3_apply(s, ent[k].P - prev, ent[k].P - prev ? ent[k].lvl : -1, &ok);
        s = s + ent[k].v;
        after[k] = s; prev = ent[k].P;
    }
    double tot = x3_apply(s, P - prev, -1, &ok);
    if (!ok) stats[1]++;
    *total = tot;
    // emit
    size_t ko = 0; double base = 0.0; unsigned long long Pb = 0;
    for (size_t i = 0; i < nt * T; ++i) {
        double c;
        if (dirty[i]) { base = after[ko]; Pb = ent[ko].P; ko++; c = base; }
        else {
            int ok2 = 1;
            unsigned long long dp = Pel[i] + incs[i] - Pb;
            c = x3_apply(base, dp, incs[i] ? lvls[i] : -1, &ok2);
            if (!ok2) stats[2]++;
        }
        if (i < n) out[i] = c;
    }
    return ok;
}

extern "C" void x3_seq_scan(const double* v, size_t n, double scale, double* out) {
    double s = 0.0;
    for (size_t i = 0; i < n; ++i) { double x = scale != 0.0 ? v[i] / scale : v[i]; s = s + x; out[i] = s; }
}

// comb: closed form vs the loop; returns the number of mismatches over t in [0, n) (n = 2^p), sampled with stride
extern "C" long long x3_emul_comb(double u01, int p, size_t stride) {
    const double inv = std::ldexp(1.0, -p), ninv = std::ldexp(1.0, p);
    const size_t n = (size_t)1 << p;
    const double r0 = u01 * (inv - 0.0) + 0.0;
    long long bad = 0;
    double r = r0;
    x3_comb_table tb;
    x3_comb_build(&tb, r0, inv, ninv, n);
    for (size_t t = 0; t < n; ++t) {
        if (t % stride == 0 || t + 3 > n) {
            double c = x3_comb_pow2(r0, inv, ninv, t);
            if (std::memcmp(&c, &r, 8) != 0) bad++;
        }
        {
            double c2 = x3_comb_eval(&tb, inv, t);
            if (std::memcmp(&c2, &r, 8) != 0) bad++;
        }
        r = r + inv;
    }
    return bad;
}
