// fs_post.cuh — everything fastslam_update does AFTER the per-particle work (fs1.rs:258-265 + resample fs1.rs:206-231),
// fused into ONE cooperative kernel for particle counts that fit one co-resident wave of FX_TILE-particle CTAs:
//
//     S  = sum w_raw (exact, sequential order)            normalize_weights        fs1.rs:196-203
//     w  = w_raw / S                                                                fs1.rs:200
//     Q  = sum w^2  -> neff = 1/Q -> gate = neff < NTH     compute_neff + gate      fs1.rs:186-193, 262-263
//     S2 = sum w ;  w2 = w / S2                           resample re-normalises   fs1.rs:207
//     c  = cumsum(w2),  r_t = r0 + t/n accumulated        cum_sum + comb           fs1.rs:213-216, 219-230
//   (j_t = first j with c_j >= r_t and the pose clone run in fs_search_pose_kernel, at full occupancy)
//
// Each CTA owns one tile of FX_TILE = 512 particles and keeps its values in REGISTERS across all phases; the
// five exact sums run through the same tile-aggregate / chain scheme as xsum.cuh, but the chain (tens of entries) is
// evaluated redundantly by every CTA, so an exact total costs 2 grid-wide barriers and no extra launches.  Q and S2
// share their barriers, so do the two scans: 4 barriers for a step that does not resample, 8 for one that does
// (instead of ~25 launches).
#pragma once
#include <cooperative_groups.h>
#include "common.cuh"
#include "xsum.cuh"
#include "fs_kernels.cuh"

namespace cg = cooperative_groups;

#define FX_MAX_TILES 512
#define FX_CHUNK 256
#define FX_ITEMS 2                         // values per thread: small on purpose — the kernel is instruction-fetch bound, so
#define FX_TILE (XS_NT * FX_ITEMS)         // short unrolled bodies on many CTAs (128 for 65 536 particles) beat long ones on few
#define FX_SLOTS 5            // S, Q, S2, cum, comb

struct FxSlot { double* tsum; xs_t* ttail; int* tnd; XsEntry* ent; };
struct FxWork {
    unsigned long long* dbg;  // [32] phase timestamps of CTA 0 (ns, %globaltimer): [k] accumulates t_k - t_{k-1}; [31] = launches
    FxSlot slot[FX_SLOTS];
    int* flags;               // [0] bad value seen (-> exact serial walk)  [1] cumulative serial walks  [2] cumulative certificate failures
                              // [3] cumulative overflow tiles
};

template <int MAXT>
struct FxSharedT {
    double sm_d[XS_NT / 32];
    int sm_i[XS_NT / 32];
    XsSeg sm_s[XS_NT / 32];
    int sm_j[XS_NT / 32];
    XsSeg carry_seg;
    int carry_nd;
    int ndl;                           // number of tiles holding dirty values
    unsigned short dl_tile[MAXT];      // ... their ids, in order
    int dl_off[MAXT];                  // ... and the ordinal of their first dirty value
    xs_t tin[MAXT];
    int tdoff[MAXT];
    int tnd[MAXT];
    XsEntry ent[FX_CHUNK];
    double before[FX_CHUNK];           // exact running sum in front of / right after each staged dirty value
    double after[FX_CHUNK];
    double after_win[XS_MAXD + 2];     // exact prefixes right after the dirty values of MY tile (+ the one before it)
    double total;
    double s_run;
    int ok;
    int serial;
    double r0;
};
typedef FxSharedT<FX_MAX_TILES> FxShared;

// Where a tile's aggregates are published.  One GPU: plain stores into the local slot arrays, tile id = blockIdx.x.
// (fs_mg.cuh has the multi-GPU publisher: the same stores repeated into every peer's window over NVLink.)
struct FxPubLocal {
    __device__ __forceinline__ unsigned me() const { return blockIdx.x; }
    __device__ __forceinline__ void tsum(const FxSlot& s, int, unsigned b, double v) const { s.tsum[b] = v; }
    __device__ __forceinline__ void tail(const FxSlot& s, int, unsigned b, xs_t t, int nd) const { s.ttail[b] = t; s.tnd[b] = nd; }
    __device__ __forceinline__ void ent(const FxSlot& s, int, size_t k, const XsEntry& e) const { s.ent[k] = e; }
    __device__ __forceinline__ void bad(int* flags) const { flags[0] = 1; }
    __device__ __forceinline__ int is_bad(const int* flags) const { return flags[0]; }
};

struct FxTile {            // what a thread keeps between classify and emit
    double toff, excl;
    XsSeg carry;           // clean values of this tile before the thread since the last dirty one (flag: such a dirty value exists)
    int doff;              // dirty values of this tile before the thread
    int nd_tile;           // dirty values in the tile (-1: overflow)
};

// ---- phase 1: approximate tile sum -> global -------------------------------------------------------
template <class P, class SH>
__device__ __forceinline__ void fx_tile_sum(const double (&v)[FX_ITEMS], const FxSlot& s, int si, SH& sh, int* flags, const P& pub) {
    double t = 0.0; bool bad = false;
#pragma unroll
    for (int k = 0; k < FX_ITEMS; ++k) { t += v[k]; if (!(v[k] >= 0.0) || !(v[k] <= 1.7976931348623157e308)) bad = true; }
    if (bad) pub.bad(flags);
    double tot = block_sum<XS_NT>(t, sh.sm_d);
    if (threadIdx.x == 0) pub.tsum(s, si, pub.me(), tot);
}

// ---- phase 2 (after a grid barrier): classify the tile, publish its aggregate and dirty entries ----------------
// approximate sum of the tiles in front of tile b, from the published tile sums of slot s (same order in every CTA)
template <class SH>
__device__ __forceinline__ double fx_tile_offset(const FxSlot& s, unsigned b, SH& sh) {
    double part = 0.0;
    for (unsigned t = threadIdx.x; t < b; t += XS_NT) part += s.tsum[t];
    double toff_b = block_sum<XS_NT>(part, sh.sm_d);
    __syncthreads();
    if (threadIdx.x == 0) sh.total = toff_b;
    __syncthreads();
    return sh.total;
}
template <class P, class SH>
__device__ __forceinline__ FxTile fx_classify_impl(const double (&v)[FX_ITEMS], double toff, const FxSlot& s, int si, SH& sh, double rel, const P& pub);
// `toff`: approximate prefix in front of this tile.  It only steers the classification (any value within the margin `rel`
// of the exact prefix gives the same exact result), so callers may derive it from sums they already hold instead of
// publishing tile sums of these very values first — e.g. sum(w_i / S) from sum(w_i) / S (one rounding per term apart).
template <class P, class SH>
__device__ __forceinline__ FxTile fx_classify_at(const double (&v)[FX_ITEMS], double toff, const FxSlot& s, int si, SH& sh, double rel, const P& pub, int* flags) {
    bool bad = false;
#pragma unroll
    for (int k = 0; k < FX_ITEMS; ++k) if (!(v[k] >= 0.0) || !(v[k] <= 1.7976931348623157e308)) bad = true;
    if (bad) pub.bad(flags);
    return fx_classify_impl(v, toff, s, si, sh, rel, pub);
}
template <class P, class SH>
__device__ __forceinline__ FxTile fx_classify(const double (&v)[FX_ITEMS], const FxSlot& s, int si, SH& sh, double rel, const P& pub) {
    return fx_classify_impl(v, fx_tile_offset(s, pub.me(), sh), s, si, sh, rel, pub);
}
template <class P, class SH>
__device__ __forceinline__ FxTile fx_classify_impl(const double (&v)[FX_ITEMS], double toff, const FxSlot& s, int si, SH& sh, double rel, const P& pub) {
    const unsigned b = pub.me();
    FxTile c;
    c.toff = toff;
    double tsum = 0.0;
#pragma unroll
    for (int k = 0; k < FX_ITEMS; ++k) tsum += v[k];
    double btot;
    c.excl = block_excl_scan<XS_NT>(tsum, &btot, sh.sm_d);
    XsThreadScan ts = xs_thread_scan(v, c.toff, c.excl, rel);
    XsSeg stot;
    XsSeg carry = xs_block_seg_excl<XS_NT>(xs_seg_make(ts.tail, ts.nd > 0), &stot, sh.sm_s);
    int ndtot;
    c.doff = block_excl_scan_int<XS_NT>(ts.nd, &ndtot, sh.sm_i);
    c.carry = carry;
    const bool overflow = ndtot > XS_MAXD;
    c.nd_tile = overflow ? -1 : ndtot;
    if (ts.nd > 0 && !overflow) {
        xs_t run = carry.t;
        double running = 0.0, a_prev = c.toff + c.excl;
        int slot = c.doff;
#pragma unroll
        for (int k = 0; k < FX_ITEMS; ++k) {
            running += v[k];
            double a_cur = c.toff + (c.excl + running);
            xs_t t;
            if (xs_classify(v[k], a_prev, a_cur, rel, &t)) run = xs_compose(run, t);
            else {
                XsEntry e; e.inc = run.inc; e.lvl = run.lvl; e.pad = 0; e.v = v[k];
                pub.ent(s, si, (size_t)b * XS_MAXD + slot, e);
                slot++;
                run = xs_identity();
            }
            a_prev = a_cur;
        }
    }
    if (threadIdx.x == 0) {
        if (overflow) {
            XsEntry e; e.inc = 0; e.lvl = XS_EMPTY; e.pad = 1; e.v = 0.0;
            pub.ent(s, si, (size_t)b * XS_MAXD, e);
            pub.tail(s, si, b, xs_identity(), -1);
        } else {
            pub.tail(s, si, b, stot.t, ndtot);
        }
    }
    return c;
}

// ---- phase 3 (after a grid barrier): every CTA evaluates the whole chain; results land in shared memory ---------
// sh.total = exact total; sh.tin[blockIdx.x], sh.tdoff[blockIdx.x], sh.after_win[] serve fx_emit for this tile.
#define FXC_STAMP(k) do { if (dbg && me == 0 && tid == 0) { unsigned long long t__; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t__)); dbg[k] += t__ - tp; tp = t__; } } while (0)
template <class F, class P, class SH>
__device__ __noinline__ void fx_chain(const FxSlot& s, unsigned nt, F f, size_t n, SH& sh, int* flags, const P& pub, unsigned long long* dbg = nullptr) {
    const int tid = threadIdx.x;
    const unsigned me = pub.me();
    unsigned long long tp = 0;
    if (dbg) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(tp));
    __syncthreads();
    // ---- tile-level segmented scan over all nt tiles, by the whole CTA: thread t owns `per` consecutive tiles (one for
    // nt <= 256), warps scan by shuffles, the 8 warp totals are combined by every thread.  (One warp doing all of it took
    // 5 us per chain at 128 tiles and grew with the tile count, i.e. with the number of GPUs.)
    for (unsigned b = tid; b < nt; b += XS_NT) { sh.tnd[b] = s.tnd[b]; sh.tin[b] = s.ttail[b]; }
    if (tid == 0) { sh.s_run = 0.0; sh.ok = 1; sh.serial = pub.is_bad(flags); }
    __syncthreads();
    {
        const int lane = tid & 31, wid = tid >> 5;
        const unsigned per = (nt + XS_NT - 1) / XS_NT;
        const unsigned b0 = (unsigned)tid * per < nt ? (unsigned)tid * per : nt, b1 = (b0 + per < nt) ? b0 + per : nt;
        XsSeg run = xs_seg_make(xs_identity(), 0);
        int ndrun = 0, ntl = 0;                                // dirty values / dirty tiles in my tiles
        for (unsigned b = b0; b < b1; ++b) {                   // pass 1: my total
            int nd = sh.tnd[b]; int nde = nd < 0 ? 1 : nd;
            run = xs_seg_op(run, xs_seg_make(sh.tin[b], nde > 0));
            ndrun += nde; ntl += nde > 0 ? 1 : 0;
        }
        XsSeg inc = run; int ndinc = ndrun, ntinc = ntl;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            XsSeg y = xs_seg_shfl_up(inc, o);
            int yn = __shfl_up_sync(0xffffffffu, ndinc, o), yt = __shfl_up_sync(0xffffffffu, ntinc, o);
            if (lane >= o) { inc = xs_seg_op(y, inc); ndinc += yn; ntinc += yt; }
        }
        XsSeg ex = xs_seg_shfl_up(inc, 1);
        int ndex = __shfl_up_sync(0xffffffffu, ndinc, 1), ntex = __shfl_up_sync(0xffffffffu, ntinc, 1);
        if (lane == 0) { ex = xs_seg_make(xs_identity(), 0); ndex = 0; ntex = 0; }
        if (lane == 31) { sh.sm_s[wid] = inc; sh.sm_i[wid] = ndinc; sh.sm_j[wid] = ntinc; }
        __syncthreads();
        XsSeg wex = xs_seg_make(xs_identity(), 0);
        int wnd = 0, wnt = 0;
        for (int w = 0; w < wid; ++w) { wex = xs_seg_op(wex, sh.sm_s[w]); wnd += sh.sm_i[w]; wnt += sh.sm_j[w]; }
        ex = xs_seg_op(wex, ex); ndex += wnd; ntex += wnt;
        if (tid == XS_NT - 1) { sh.carry_seg = xs_seg_op(wex, inc); sh.carry_nd = wnd + ndinc; sh.ndl = wnt + ntinc; }
        __syncthreads();                                       // pass 2 overwrites sh.tin: every pass 1 is done
        for (unsigned b = b0; b < b1; ++b) {                   // pass 2: exclusive prefixes per tile + compact dirty-tile list
            int nd = sh.tnd[b]; int nde = nd < 0 ? 1 : nd;
            xs_t tl = sh.tin[b];
            sh.tin[b] = ex.t; sh.tdoff[b] = ndex;
            if (nde > 0) { sh.dl_tile[ntex] = (unsigned short)b; sh.dl_off[ntex] = ndex; ntex++; }
            ex = xs_seg_op(ex, xs_seg_make(tl, nde > 0));
            ndex += nde;
        }
    }
    __syncthreads();
    const int D = sh.carry_nd;
    if (tid == 0 && sh.carry_seg.t.lvl == XS_BAD) sh.serial = 1;
    __syncthreads();
    FXC_STAMP(0);
    const int my_o0 = sh.tdoff[me];                         // window of ordinals fx_emit needs: [my_o0 - 1, my_o0 + nd_me)
    if (!sh.serial) {
        for (int cbase = 0; cbase < D; cbase += FX_CHUNK) {
            const int cnt = D - cbase < FX_CHUNK ? D - cbase : FX_CHUNK;
            // stage exactly the dirty entries: thread o finds its (tile, slot) in the compact dirty-tile list
            for (int o = tid; o < cnt; o += XS_NT) {
                const int og = cbase + o;
                int lo = 0, hi = sh.ndl - 1;
                while (lo < hi) { int mid = (lo + hi + 1) >> 1; if (sh.dl_off[mid] <= og) lo = mid; else hi = mid - 1; }
                const unsigned b = sh.dl_tile[lo];
                const int e = og - sh.dl_off[lo];
                XsEntry en = s.ent[(size_t)b * XS_MAXD + e];
                if (e == 0) {
                    xs_t r; r.inc = en.inc; r.lvl = en.lvl;
                    r = xs_compose(sh.tin[b], r);
                    en.inc = r.inc; en.lvl = r.lvl;
                }
                if (en.pad == 1) en.v = (double)b;
                sh.ent[o] = en;
            }
            __syncthreads();
            FXC_STAMP(1);
            if (dbg && me == 0 && tid == 0) dbg[6] += (unsigned long long)D;
            // Serial part, kept minimal (a lone GPU thread retires ~1 dependent instruction per 6-8 cycles): per entry one
            // 64-bit integer add on the bit pattern and one FP add.  The certificate checks run afterwards, in parallel.
            if (tid == 0) {
                double sacc = sh.s_run;
                int o = 0;
                for (; o + 4 <= cnt; o += 4) {              // 4 entries per trip: their operands are fetched up front
                    const long long i0 = sh.ent[o].inc, i1 = sh.ent[o + 1].inc, i2 = sh.ent[o + 2].inc, i3 = sh.ent[o + 3].inc;
                    const double v0 = sh.ent[o].v, v1 = sh.ent[o + 1].v, v2 = sh.ent[o + 2].v, v3 = sh.ent[o + 3].v;
                    const int anypad = sh.ent[o].pad | sh.ent[o + 1].pad | sh.ent[o + 2].pad | sh.ent[o + 3].pad;
                    if (anypad) break;
                    sh.before[o] = sacc;     sacc = pfc_u2d(pfc_d2u(sacc) + (unsigned long long)i0) + v0; sh.after[o] = sacc;
                    sh.before[o + 1] = sacc; sacc = pfc_u2d(pfc_d2u(sacc) + (unsigned long long)i1) + v1; sh.after[o + 1] = sacc;
                    sh.before[o + 2] = sacc; sacc = pfc_u2d(pfc_d2u(sacc) + (unsigned long long)i2) + v2; sh.after[o + 2] = sacc;
                    sh.before[o + 3] = sacc; sacc = pfc_u2d(pfc_d2u(sacc) + (unsigned long long)i3) + v3; sh.after[o + 3] = sacc;
                }
                for (; o < cnt; ++o) {
                    sh.before[o] = sacc;
                    sacc = pfc_u2d(pfc_d2u(sacc) + (unsigned long long)sh.ent[o].inc);
                    if (sh.ent[o].pad == 1) {               // overflow tile: genuine FP adds over the whole tile
                        unsigned b = (unsigned)sh.ent[o].v;
                        if (b == me) sh.after_win[0] = sacc;
                        size_t lo = (size_t)b * FX_TILE, hi = lo + FX_TILE < n ? lo + FX_TILE : n;
                        for (size_t i = lo; i < hi; ++i) sacc = sacc + f(i);
                        if (me == 0) flags[3] += 1;
                    } else {
                        sacc = sacc + sh.ent[o].v;
                    }
                    sh.after[o] = sacc;
                }
                sh.s_run = sacc;
            }
            __syncthreads();
            FXC_STAMP(2);
            for (int o = tid; o < cnt; o += XS_NT) {
                // certificate of entry o: the run in front of it really was at the level it was classified for and stayed there
                xs_t r; r.inc = sh.ent[o].inc; r.lvl = sh.ent[o].lvl;
                int ok = 1;
                (void)xs_apply(r, sh.before[o], &ok);
                if (!ok) sh.ok = 0;
                const int wi = cbase + o - (my_o0 - 1);
                if (wi >= 0 && wi < XS_MAXD + 2 && !(sh.tnd[me] < 0 && wi == 0)) sh.after_win[wi] = sh.after[o];
            }
            __syncthreads();
        }
        if (tid == 0) {
            int ok = sh.ok;
            double tot = xs_apply(sh.carry_seg.t, sh.s_run, &ok);
            if (!ok) { sh.serial = 1; if (me == 0) flags[2] += 1; }
            else sh.total = tot;
        }
        __syncthreads();
        FXC_STAMP(3);
    }
    if (sh.serial) {      // exact by construction: one thread, left to right; also yields this tile's base
        if (tid == 0) {
            double sacc = 0.0;
            size_t mylo = (size_t)me * FX_TILE;
            for (size_t i = 0; i < n; ++i) { if (i == mylo) sh.after_win[0] = sacc; sacc = sacc + f(i); }
            sh.total = sacc;
            if (me == 0) flags[1] += 1;
        }
        __syncthreads();
    }
}

// ---- phase 4: exact inclusive prefix of this thread's 8 values ------------------------------------------------------
template <class F, class P, class SH>
__device__ __forceinline__ void fx_emit(const double (&v)[FX_ITEMS], const FxTile& c, SH& sh, F f, size_t n, double rel,
                                        double (&out)[FX_ITEMS], int* flags, double* gscratch /* [n], this tile's slice is ours */,
                                        const P& pub) {
    const unsigned me = pub.me();
    const size_t first = (size_t)me * FX_TILE + (size_t)threadIdx.x * FX_ITEMS;
    if (sh.serial || sh.tnd[me] < 0) {          // serial walk / overflow tile: thread 0 walks the tile from its exact base
        if (threadIdx.x == 0) {
            double sacc = sh.after_win[0];
            size_t lo = (size_t)me * FX_TILE, hi = lo + FX_TILE < n ? lo + FX_TILE : n;
            for (size_t i = lo; i < hi; ++i) { sacc = sacc + f(i); gscratch[i] = sacc; }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < FX_ITEMS; ++k) out[k] = first + k < n ? gscratch[first + k] : 0.0;
        __syncthreads();
        return;
    }
    xs_t run = xs_seg_op(xs_seg_make(sh.tin[me], 0), c.carry).t;   // the carry-in counts only if no dirty value precedes in the tile
    // after_win[0] = exact prefix right after the last dirty value BEFORE my tile (ordinal tdoff-1), [1+e] = after my e-th
    int w = c.doff;                                        // dirty values of my tile before this thread
    double base_s = (sh.tdoff[me] + w > 0) ? sh.after_win[w] : 0.0;
    double running = 0.0, a_prev = c.toff + c.excl;
    int ok = 1;
#pragma unroll
    for (int k = 0; k < FX_ITEMS; ++k) {
        running += v[k];
        double a_cur = c.toff + (c.excl + running);
        xs_t t;
        if (xs_classify(v[k], a_prev, a_cur, rel, &t)) { run = xs_compose(run, t); out[k] = xs_apply(run, base_s, &ok); }
        else { w++; base_s = sh.after_win[w]; out[k] = base_s; run = xs_identity(); }
        a_prev = a_cur;
    }
    (void)first;
    if (!ok) flags[2] += 1;
}

__device__ __forceinline__ unsigned long long fx_now() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define FX_STAMP(k) do { if (b == 0 && tid == 0 && fw.dbg) { unsigned long long t__ = fx_now(); fw.dbg[k] += t__ - t_prev; t_prev = t__; } } while (0)

struct FxValW2 {      // w / S2 recomputed on the fly for the rare global walks (overflow tiles / serial mode)
    const double* w; double S2;
    __device__ __forceinline__ double operator()(size_t i) const { double x = w[i]; return S2 > 0.0 ? x / S2 : x; }
};
struct FxValComb { double r0, inv; __device__ __forceinline__ double operator()(size_t i) const { return i == 0 ? r0 : inv; } };

// =====================================================================================================================
__global__ void __launch_bounds__(XS_NT, 2) fs_post_kernel(FsDev d, FxWork fw, double nth, uint64_t seed, unsigned nt, double rel,
                                                           const __grid_constant__ FsObsParam po, int k_obs) {
    cg::grid_group grid = cg::this_grid();
    __shared__ FxShared sh;
    const FxPubLocal pub;
    const unsigned b = blockIdx.x;
    const int tid = threadIdx.x;
    const size_t first = (size_t)b * FX_TILE + (size_t)tid * FX_ITEMS;
    const size_t n = d.n;
    unsigned long long t_prev = fx_now();
    if (b == 0 && tid == 0) { fw.flags[0] = 0; if (fw.dbg) fw.dbg[31] += 1; }
    if (b == 0 && tid < k_obs) fs_mark_updated(d, po.o[tid].lm_id);   // lazy-clone bookkeeping of the EKF launch that just ran
    double v[FX_ITEMS];
#pragma unroll
    for (int k = 0; k < FX_ITEMS; ++k) { size_t i = first + k; v[k] = i < n ? d.w_raw[i] : 0.0; }
    grid.sync();                                               // flags[0] reset is visible before anyone raises it
    FX_STAMP(0);
    // ---------------- S = sum w_raw; w = w_raw / S (fs1.rs:196-203) ----------------
    fx_tile_sum(v, fw.slot[0], 0, sh, fw.flags, pub);
    FX_STAMP(1);
    grid.sync();
    FX_STAMP(2);
    fx_classify(v, fw.slot[0], 0, sh, rel, pub);
    FX_STAMP(3);
    grid.sync();
    FX_STAMP(4);
    fx_chain(fw.slot[0], nt, XsValArray{d.w_raw}, n, sh, fw.flags, pub, fw.dbg ? fw.dbg + 16 : nullptr);
    FX_STAMP(5);
    const double S = sh.total;
    double q[FX_ITEMS];
#pragma unroll
    for (int k = 0; k < FX_ITEMS; ++k) {
        size_t i = first + k;
        if (S > 0.0) v[k] = v[k] / S;
        if (i < n) d.w[i] = v[k];
        q[k] = v[k] * v[k];
    }
    // ---------------- gate: neff = 1 / sum w^2 < NTH (fs1.rs:186-193, 262-263) ----------------
    // Only the DECISION feeds back into the state, so Q is first summed in tree order (one barrier).  The sequential sum
    // the reference computes differs from it by at most (n+64)*2^-52 relatively; only when neff lands that close to NTH is
    // the exact sequential sum evaluated (same decision as the reference in every case).
    fx_tile_sum(q, fw.slot[1], 1, sh, fw.flags, pub);                  // per-tile partial sums of w^2
    fx_tile_sum(v, fw.slot[2], 2, sh, fw.flags, pub);                  // per-tile sums of w (needed only if the gate opens)
    FX_STAMP(6);
    grid.sync();
    FX_STAMP(7);
    double qpart = 0.0;
    for (unsigned t = tid; t < nt; t += XS_NT) qpart += fw.slot[1].tsum[t];
    double qa = block_sum<XS_NT>(qpart, sh.sm_d);
    __syncthreads();
    if (tid == 0) sh.total = qa;
    __syncthreads();
    double Q = sh.total;                                       // bit-identical in every CTA (same order everywhere)
    double neff = Q > 0.0 ? 1.0 / Q : 0.0;
    const double slack = 8.0 * (double)(d.n_global + 64) * 2.220446049250313e-16;
    const bool border = !(fabs(neff - nth) > slack * fmax(fabs(nth), fabs(neff))) || (fw.flags[0] != 0);
    FX_STAMP(8);
    if (border) {                                              // rare; uniform over the grid
        fx_classify(q, fw.slot[1], 1, sh, rel, pub);
        grid.sync();
        fx_chain(fw.slot[1], nt, FsValWSq{d.w}, n, sh, fw.flags, pub);
        Q = sh.total;
        neff = Q > 0.0 ? 1.0 / Q : 0.0;                        // compute_neff fs1.rs:186-193
    }
    FX_STAMP(9);
    const int gate = neff < nth ? 1 : 0;                       // fs1.rs:263
    if (b == 0 && tid == 0) {
        d.scal[0] = S; d.scal[1] = Q; d.scal[3] = neff;
        *d.gate = gate;
    }
    if (!gate) return;                                         // the whole grid takes the same branch
    // ---------------- resample: S2 = sum w (fs1.rs:207) ----------------
    fx_classify(v, fw.slot[2], 2, sh, rel, pub);
    grid.sync();
    fx_chain(fw.slot[2], nt, XsValArray{d.w}, n, sh, fw.flags, pub);
    const double S2 = sh.total;
    if (b == 0 && tid == 0) d.scal[2] = S2;
    FX_STAMP(10);
    // ---------------- resample (fs1.rs:206-231) ----------------
    const double inv = 1.0 / (double)d.n_global;
    if (tid == 0) {
        double u01 = pfc_u01_52(pfc_blk_u64(pfc_rng_block(seed, PFC_STREAM_FS_RESAMPLE, d.counters[0], 0), 0));
        sh.r0 = u01 * (inv - 0.0) + 0.0;                       // Uniform::new(0, 1/n).sample
    }
    __syncthreads();
    const double r0 = sh.r0;
    double cv[FX_ITEMS];
#pragma unroll
    for (int k = 0; k < FX_ITEMS; ++k) {
        size_t i = first + k;
        if (S2 > 0.0) v[k] = v[k] / S2;                        // normalize_weights inside resample() fs1.rs:207
        cv[k] = i < n ? (i == 0 ? r0 : inv) : 0.0;
    }
    // approximate tile prefixes without another publish + barrier: the CDF's from the tile sums of w (slot 2) scaled by
    // 1/S2, the comb's in closed form
    FX_STAMP(11);
    double toff_c = fx_tile_offset(fw.slot[2], b, sh);
    if (S2 > 0.0) toff_c = toff_c / S2;
    const double toff_r = b == 0 ? 0.0 : r0 + ((double)((size_t)b * FX_TILE) - 1.0) * inv;
    FX_STAMP(12);
    FxTile tc = fx_classify_at(v, toff_c, fw.slot[3], 3, sh, rel, pub, fw.flags);
    FxTile tr = fx_classify_at(cv, toff_r, fw.slot[4], 4, sh, rel, pub, fw.flags);
    FX_STAMP(13);
    grid.sync();
    FX_STAMP(14);
    double c[FX_ITEMS], r[FX_ITEMS];
    fx_chain(fw.slot[3], nt, FxValW2{d.w, S2}, n, sh, fw.flags, pub);
    fx_emit(v, tc, sh, FxValW2{d.w, S2}, n, rel, c, fw.flags, d.cum, pub);
    fx_chain(fw.slot[4], nt, FxValComb{r0, inv}, n, sh, fw.flags, pub);
    fx_emit(cv, tr, sh, FxValComb{r0, inv}, n, rel, r, fw.flags, d.rcomb, pub);
    FX_STAMP(15);
#pragma unroll
    for (int k = 0; k < FX_ITEMS; ++k) { size_t i = first + k; if (i < n) { d.cum[i] = c[k]; d.rcomb[i] = r[k]; } }
    // the index walk (fs1.rs:224-226) and the pose clone (fs1.rs:227-229) need every slice of cum: they run in the
    // next, full-occupancy launch (fs_search_pose_kernel), gated on *d.gate
}
