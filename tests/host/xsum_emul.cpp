// Host emulation of the device xsum pipeline (same pass structure, same element-level code from
// rust_robotics_b200/csrc/xsum_core.h).  Built and driven by tests/test_xsum_host.py.  Test code only.
#include "../../rust_robotics_b200/csrc/xsum_core.h"
#include <vector>
#include <cstdio>

struct Dirty { xs_t run; double v; };

static double pairwise(const double* v, size_t n) {
    if (n == 0) return 0.0;
    if (n == 1) return v[0];
    size_t h = n / 2;
    return pairwise(v + h, n - h) + pairwise(v, h);     // deliberately NOT the sequential order
}

extern "C" void xs_seq_scan(const double* v, size_t n, double* out) {
    double c = 0.0;
    for (size_t i = 0; i < n; ++i) { c = c + v[i]; out[i] = c; }
}

// stats: [0] dirty count, [1] certificate violations (must be 0), [2] universal-identity count
// toff_in: optional approximate tile prefixes supplied by the caller (the fused kernels derive them from sums they already
// hold instead of summing these very values first, see fs_post.cuh fx_classify_at); null = from pairwise tile sums of v
static int xs_emul_scan_impl(const double* v, size_t n, size_t tile, const double* toff_in, double* out, double* total, long long* stats) {
    const double rel = xs_margin(n);
    size_t nt = (n + tile - 1) / tile;
    std::vector<double> tsum(nt), toff(nt);
    for (size_t b = 0; b < nt; ++b) { size_t lo = b * tile, hi = lo + tile < n ? lo + tile : n; tsum[b] = pairwise(v + lo, hi - lo); }
    { // pass B: a *different* order again: offsets from a pairwise-ish running sum
        double acc = 0.0; for (size_t b = 0; b < nt; ++b) { toff[b] = acc; acc = tsum[b] + acc; }
    }
    if (toff_in) for (size_t b = 0; b < nt; ++b) toff[b] = toff_in[b];
    std::vector<xs_t> tail(nt); std::vector<int> anyd(nt); std::vector<std::vector<Dirty>> ent(nt);
    long long nd = 0, viol = 0, nid = 0;
    const size_t chunk = 8;   // "per-thread items"
    auto approx_prefix = [&](size_t b, size_t i_local, size_t lo, size_t hi) {
        // tile_off + (sum of whole chunks before, pairwise) + running inside the chunk: differs from sequential order
        size_t c0 = (i_local / chunk) * chunk;
        double pre = pairwise(v + lo, c0);
        double run = 0.0; for (size_t j = c0; j <= i_local; ++j) run += v[lo + j];
        (void)hi;
        return toff[b] + (pre + run);
    };
    for (size_t b = 0; b < nt; ++b) {
        size_t lo = b * tile, hi = lo + tile < n ? lo + tile : n;
        xs_t run = xs_identity(); anyd[b] = 0;
        for (size_t i = lo; i < hi; ++i) {
            double a_prev = (i == lo) ? toff[b] : approx_prefix(b, i - lo - 1, lo, hi);
            double a_cur = approx_prefix(b, i - lo, lo, hi);
            xs_t t;
            if (xs_classify(v[i], a_prev, a_cur, rel, &t)) { if (t.lvl == XS_EMPTY) nid++; run = xs_compose(run, t); }
            else { ent[b].push_back({run, v[i]}); run = xs_identity(); anyd[b] = 1; nd++; }
        }
        tail[b] = run;
    }
    // pass D: chain
    std::vector<xs_t> tin(nt); std::vector<double> sbase(nt);
    xs_t carry = xs_identity(); double s = 0.0; int ok = 1;
    for (size_t b = 0; b < nt; ++b) {
        tin[b] = carry; sbase[b] = s;
        if (anyd[b]) {
            s = xs_apply(carry, s, &ok);
            for (auto& e : ent[b]) { s = xs_apply(e.run, s, &ok); s = s + e.v; }
            carry = tail[b];
        } else carry = xs_compose(carry, tail[b]);
        if (carry.lvl == XS_BAD) ok = 0;
    }
    *total = xs_apply(carry, s, &ok);
    // pass E: output
    if (out) for (size_t b = 0; b < nt; ++b) {
        size_t lo = b * tile, hi = lo + tile < n ? lo + tile : n;
        xs_t run = tin[b]; double base = sbase[b];
        for (size_t i = lo; i < hi; ++i) {
            double a_prev = (i == lo) ? toff[b] : approx_prefix(b, i - lo - 1, lo, hi);
            double a_cur = approx_prefix(b, i - lo, lo, hi);
            xs_t t;
            if (xs_classify(v[i], a_prev, a_cur, rel, &t)) { run = xs_compose(run, t); out[i] = xs_apply(run, base, &ok); }
            else { base = xs_apply(run, base, &ok); base = base + v[i]; out[i] = base; run = xs_identity(); }
        }
    }
    if (!ok) viol++;
    stats[0] = nd; stats[1] = viol; stats[2] = nid;
    return ok;
}

extern "C" int xs_emul_scan(const double* v, size_t n, size_t tile, double* out, double* total, long long* stats) {
    return xs_emul_scan_impl(v, n, tile, nullptr, out, total, stats);
}
// scan of w_i / S with the tile prefixes taken from the tile sums of w, divided by S afterwards
// (fs_post.cuh: the CDF from the tile sums of w scaled by 1/S2; fs_mg.cuh: S2 from the tile sums of w_raw scaled by 1/S)
extern "C" int xs_emul_scan_scaled(const double* w, size_t n, double S, size_t tile, double* v_out, double* out, double* total, long long* stats) {
    size_t nt = (n + tile - 1) / tile;
    std::vector<double> v(n), toff(nt);
    for (size_t i = 0; i < n; ++i) v[i] = S > 0.0 ? w[i] / S : w[i];
    double acc = 0.0;
    for (size_t b = 0; b < nt; ++b) {
        size_t lo = b * tile, hi = lo + tile < n ? lo + tile : n;
        toff[b] = S > 0.0 ? acc / S : acc;
        acc = pairwise(w + lo, hi - lo) + acc;
    }
    for (size_t i = 0; i < n; ++i) v_out[i] = v[i];
    return xs_emul_scan_impl(v.data(), n, tile, toff.data(), out, total, stats);
}
// the systematic comb r0, 1/n, 1/n, ... (fs1.rs:219-230) with closed-form tile prefixes r0 + (b*tile - 1)/n
extern "C" int xs_emul_scan_comb(double r0, double inv, size_t n, size_t tile, double* v_out, double* out, double* total, long long* stats) {
    size_t nt = (n + tile - 1) / tile;
    std::vector<double> v(n), toff(nt);
    for (size_t i = 0; i < n; ++i) v[i] = i == 0 ? r0 : inv;
    for (size_t b = 0; b < nt; ++b) toff[b] = b == 0 ? 0.0 : r0 + ((double)(b * tile) - 1.0) * inv;
    for (size_t i = 0; i < n; ++i) v_out[i] = v[i];
    return xs_emul_scan_impl(v.data(), n, tile, toff.data(), out, total, stats);
}

