// xsum.cuh — device-wide EXACT sequential-order sum / inclusive scan of non-negative f64 values.
//
// Reproduces, bit for bit and in parallel, the reference's left-to-right accumulations
//     sum_w (pf.rs:427, mcl.rs:395, fs1.rs:197), sum_w^2 (pf.rs:417, fs1.rs:187),
//     cum_sum (pf.rs:448-453, mcl.rs:328-333, fs1.rs:213-216) and the comb r += 1/n (fs1.rs:230).
// Theory and the element-level code: xsum_core.h (verified on the CPU by tests/test_xsum_host.py).
//
// Pipeline (tiles of XS_TILE = 256 threads x 8 consecutive values):
//   A  xs_tile_sums      approximate tile sums (tree order)                         read 8 B/value
//   B  xs_scan_tiles     exclusive scan of the tile sums (1 CTA)                    tiny
//   C  xs_classify_tiles per value: clean (an integer mantissa increment) or dirty (needs a real FP add);
//                        segmented transducer scan inside the tile -> tile aggregate + its dirty entries
//                                                                                   read 8 B/value
//   D  xs_chain          segmented scan of tile aggregates (parallel), then ONE thread applies the
//                        ~log2(n) dirty entries in order with genuine FP adds -> exact total,
//                        exact value after every dirty element, per-tile carry-in   tiny
//   E  xs_emit_tiles     (scan only) recompute C with the carry-ins and emit c_i     read 8 B, write sink
// If a certificate fails or a tile holds more than XS_MAXD dirty values, D falls back to a single-thread
// loop (flag[0]; counted in pfgpu_stats.serial_fallbacks) — slow but still exact.
#pragma once
#include "common.cuh"
#include "xsum_core.h"

#define XS_NT     256
#define XS_ITEMS  8
#define XS_TILE   (XS_NT * XS_ITEMS)
#define XS_MAXD   32
#define XS_CHAIN_NT 256
#define XS_CHUNK  768

struct __align__(8) XsEntry { long long inc; int lvl; int pad; double v; };
struct XsSeg { xs_t t; int flag; };

struct XsWork {
    size_t cap_n = 0;
    unsigned nt_cap = 0;
    double* tsum = nullptr;    // [nt] approximate tile sums
    double* toff = nullptr;    // [nt] approximate exclusive tile offsets
    xs_t*   ttail = nullptr;   // [nt] transducer of the clean values after the tile's last dirty value
    int*    tnd = nullptr;     // [nt] dirty values in the tile
    xs_t*   tin = nullptr;     // [nt] carry-in transducer (clean values since the last dirty before the tile)
    int*    tdoff = nullptr;   // [nt] dirty ordinal of the tile's first dirty value
    double* sbase = nullptr;   // [nt] serial mode: exact prefix at the tile start
    XsEntry* ent = nullptr;    // [nt * XS_MAXD]
    double* s_after = nullptr; // [nt * XS_MAXD] exact prefix right after each dirty value (by ordinal)
    int*    flags = nullptr;   // [0] serial mode  [1] dirty total  [2] emit-time certificate failure  [3] bad value seen
                               // [4] cumulative serial fallbacks  [5] dirty total of the last chain  [6] cumulative emit failures
                               // [7] cumulative overflow tiles (walked serially, still exact)
    double  approx_offset = 0.0;   // approximate sum of everything before this shard (multi-GPU)
    const double* approx_offset_ptr = nullptr;   // ... or the same on the device (sharded mode)
    const double* s_start_ptr = nullptr;         // exact prefix at the shard start, on the device (sharded mode)
    const int* gate = nullptr;     // device flag: when non-null and 0, every kernel of the pipeline returns at once
};
#define XS_GATE(w) do { if ((w).gate != nullptr && *(w).gate == 0) return; } while (0)

static int xs_work_alloc(XsWork& w, size_t n) {
    unsigned nt = cdiv_u(n, XS_TILE);
    if (nt == 0) nt = 1;
    w.cap_n = n; w.nt_cap = nt;
    PF_CUDA(cudaMalloc(&w.tsum, nt * sizeof(double)));
    PF_CUDA(cudaMalloc(&w.toff, nt * sizeof(double)));
    PF_CUDA(cudaMalloc(&w.ttail, nt * sizeof(xs_t)));
    PF_CUDA(cudaMalloc(&w.tnd, nt * sizeof(int)));
    PF_CUDA(cudaMalloc(&w.tin, nt * sizeof(xs_t)));
    PF_CUDA(cudaMalloc(&w.tdoff, nt * sizeof(int)));
    PF_CUDA(cudaMalloc(&w.sbase, nt * sizeof(double)));
    PF_CUDA(cudaMalloc(&w.ent, (size_t)nt * XS_MAXD * sizeof(XsEntry)));
    PF_CUDA(cudaMalloc(&w.s_after, (size_t)nt * XS_MAXD * sizeof(double)));
    PF_CUDA(cudaMalloc(&w.flags, 8 * sizeof(int)));
    PF_CUDA(cudaMemset(w.flags, 0, 8 * sizeof(int)));
    return 0;
}
static void xs_work_free(XsWork& w) {
    cudaFree(w.tsum); cudaFree(w.toff); cudaFree(w.ttail); cudaFree(w.tnd); cudaFree(w.tin); cudaFree(w.tdoff);
    cudaFree(w.sbase); cudaFree(w.ent); cudaFree(w.s_after); cudaFree(w.flags);
    w = XsWork();
}

// ---------------------------------------------------------------------------------------------------
// segmented transducer scan helpers
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ XsSeg xs_seg_make(xs_t t, int flag) { XsSeg s; s.t = t; s.flag = flag; return s; }
__device__ __forceinline__ XsSeg xs_seg_op(const XsSeg a, const XsSeg b) {   // a earlier, b later
    if (b.flag) return b;
    XsSeg r; r.t = xs_compose(a.t, b.t); r.flag = a.flag; return r;
}
__device__ __forceinline__ XsSeg xs_seg_shfl_up(const XsSeg s, int o) {
    XsSeg r;
    r.t.inc = __shfl_up_sync(0xffffffffu, s.t.inc, o);
    r.t.lvl = __shfl_up_sync(0xffffffffu, s.t.lvl, o);
    r.flag = __shfl_up_sync(0xffffffffu, s.flag, o);
    return r;
}
// exclusive segmented scan over the block's threads (thread order); *total = aggregate of the whole block
template <int NT>
__device__ __forceinline__ XsSeg xs_block_seg_excl(XsSeg x, XsSeg* total, XsSeg* smem /* NT/32 */) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    XsSeg inc = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        XsSeg y = xs_seg_shfl_up(inc, o);
        if (lane >= o) inc = xs_seg_op(y, inc);
    }
    XsSeg excl = xs_seg_shfl_up(inc, 1);
    if (lane == 0) excl = xs_seg_make(xs_identity(), 0);
    __syncthreads();
    if (lane == 31) smem[wid] = inc;
    __syncthreads();
    XsSeg carry = xs_seg_make(xs_identity(), 0), tot = xs_seg_make(xs_identity(), 0);
#pragma unroll
    for (int w = 0; w < NT / 32; ++w) {
        XsSeg t = smem[w];
        if (w < wid) carry = xs_seg_op(carry, t);
        tot = xs_seg_op(tot, t);
    }
    *total = tot;
    return xs_seg_op(carry, excl);
}

// ---------------------------------------------------------------------------------------------------
// A: approximate tile sums
// ---------------------------------------------------------------------------------------------------
template <class F>
__global__ void __launch_bounds__(XS_NT) xs_tile_sums(F f, size_t n, XsWork w) {
    __shared__ double sm[XS_NT / 32];
    XS_GATE(w);
    const size_t first = (size_t)blockIdx.x * XS_TILE + (size_t)threadIdx.x * XS_ITEMS;
    double s = 0.0;
    bool bad = false;
#pragma unroll
    for (int k = 0; k < XS_ITEMS; ++k) {
        size_t i = first + k;
        double v = i < n ? f(i) : 0.0;
        if (!(v >= 0.0) || !(v <= 1.7976931348623157e308)) bad = true;
        s += v;
    }
    if (bad) w.flags[3] = 1;          // xs_scan_tiles turns this into serial mode
    double t = block_sum<XS_NT>(s, sm);
    if (threadIdx.x == 0) w.tsum[blockIdx.x] = t;
}

// ---------------------------------------------------------------------------------------------------
// B: exclusive scan of tile sums (one CTA, any number of tiles)
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) xs_scan_tiles(unsigned nt, XsWork w) {
    __shared__ double sm[32];
    __shared__ double carry_s;
    XS_GATE(w);
    if (threadIdx.x == 0) { carry_s = w.approx_offset_ptr ? *w.approx_offset_ptr : w.approx_offset; w.flags[0] = w.flags[3]; w.flags[1] = 0; w.flags[2] = 0; }
    __syncthreads();
    for (unsigned base = 0; base < nt; base += 1024) {
        unsigned b = base + threadIdx.x;
        double x = b < nt ? w.tsum[b] : 0.0;
        double tot;
        double ex = block_excl_scan<1024>(x, &tot, sm);
        double c = carry_s;
        if (b < nt) w.toff[b] = c + ex;
        __syncthreads();
        if (threadIdx.x == 0) carry_s = c + tot;
        __syncthreads();
    }
}

// per-thread pass over its XS_ITEMS values: counts dirty values and builds the tail transducer.
// a0 = approximate prefix before the thread's first value.  Must be evaluated IDENTICALLY in C and E.
struct XsThreadScan { xs_t tail; int nd; };
template <int ITEMS>
__device__ __forceinline__ XsThreadScan xs_thread_scan(const double (&v)[ITEMS], double toff, double excl, double rel) {
    XsThreadScan r; r.tail = xs_identity(); r.nd = 0;
    double running = 0.0, a_prev = toff + excl;
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        running += v[k];
        double a_cur = toff + (excl + running);
        xs_t t;
        if (xs_classify(v[k], a_prev, a_cur, rel, &t)) r.tail = xs_compose(r.tail, t);
        else { r.tail = xs_identity(); r.nd++; }
        a_prev = a_cur;
    }
    return r;
}

// ---------------------------------------------------------------------------------------------------
// C: classify + tile aggregates + dirty entries
// ---------------------------------------------------------------------------------------------------
template <class F>
__global__ void __launch_bounds__(XS_NT) xs_classify_tiles(F f, size_t n, double rel, XsWork w) {
    __shared__ double sm_d[XS_NT / 32];
    __shared__ int sm_i[XS_NT / 32];
    __shared__ XsSeg sm_s[XS_NT / 32];
    XS_GATE(w);
    const unsigned b = blockIdx.x;
    const size_t first = (size_t)b * XS_TILE + (size_t)threadIdx.x * XS_ITEMS;
    double v[XS_ITEMS];
    double tsum = 0.0;
#pragma unroll
    for (int k = 0; k < XS_ITEMS; ++k) { size_t i = first + k; v[k] = i < n ? f(i) : 0.0; tsum += v[k]; }
    double btot;
    const double excl = block_excl_scan<XS_NT>(tsum, &btot, sm_d);
    const double toff = w.toff[b];
    XsThreadScan ts = xs_thread_scan(v, toff, excl, rel);
    XsSeg stot;
    XsSeg carry = xs_block_seg_excl<XS_NT>(xs_seg_make(ts.tail, ts.nd > 0), &stot, sm_s);
    int ndtot;
    int doff = block_excl_scan_int<XS_NT>(ts.nd, &ndtot, sm_i);
    // A tile with more than XS_MAXD dirty values (a long crawl along a level edge) is not itemised: it becomes
    // ONE pseudo entry (pad = 1) that tells the chain to walk the whole tile with genuine FP adds; its tail
    // transducer is the identity because the walk ends at the tile end.
    const bool overflow = ndtot > XS_MAXD;
    if (ts.nd > 0 && !overflow) {   // rare: redo the pass and write this thread's dirty entries
        xs_t run = carry.t;
        double running = 0.0, a_prev = toff + excl;
        int slot = doff;
#pragma unroll
        for (int k = 0; k < XS_ITEMS; ++k) {
            running += v[k];
            double a_cur = toff + (excl + running);
            xs_t t;
            if (xs_classify(v[k], a_prev, a_cur, rel, &t)) run = xs_compose(run, t);
            else {
                if (slot < XS_MAXD) {
                    XsEntry e; e.inc = run.inc; e.lvl = run.lvl; e.pad = 0; e.v = v[k];
                    w.ent[(size_t)b * XS_MAXD + slot] = e;
                }
                slot++;
                run = xs_identity();
            }
            a_prev = a_cur;
        }
    }
    if (threadIdx.x == 0) {
        if (overflow) {
            XsEntry e; e.inc = 0; e.lvl = XS_EMPTY; e.pad = 1; e.v = 0.0;
            w.ent[(size_t)b * XS_MAXD] = e;
            w.ttail[b] = xs_identity();
            w.tnd[b] = -1;
        } else {
            w.ttail[b] = stot.t;
            w.tnd[b] = ndtot;
        }
        if (stot.t.lvl == XS_BAD) w.flags[0] = 1;
    }
}

// ---------------------------------------------------------------------------------------------------
// D: chain
// ---------------------------------------------------------------------------------------------------
template <class F>
__global__ void __launch_bounds__(XS_CHAIN_NT) xs_chain(F f, size_t n, unsigned nt, XsWork w, double s_start,
                                                        double* total_out) {
    __shared__ XsSeg sm_s[XS_CHAIN_NT / 32];
    __shared__ int sm_i[XS_CHAIN_NT / 32];
    __shared__ XsSeg carry_seg;
    __shared__ int carry_nd;
    __shared__ XsEntry sm_ent[XS_CHUNK];
    __shared__ double sm_after[XS_CHUNK];
    __shared__ double sm_before[XS_CHUNK];
    __shared__ double s_run;
    __shared__ int ok_s;
    XS_GATE(w);
    const int tid = threadIdx.x;
    if (tid == 0) { carry_seg = xs_seg_make(xs_identity(), 0); carry_nd = 0; s_run = s_start; ok_s = 1; }
    __syncthreads();
    // phase 1: segmented exclusive scan over tiles
    for (unsigned base = 0; base < nt; base += XS_CHAIN_NT) {
        unsigned b = base + tid;
        int nd = b < nt ? w.tnd[b] : 0;
        if (nd < 0) nd = 1;                              // overflow tile: one pseudo entry
        xs_t tl = b < nt ? w.ttail[b] : xs_identity();
        XsSeg tot; int ndtot;
        XsSeg ex = xs_block_seg_excl<XS_CHAIN_NT>(xs_seg_make(tl, nd > 0), &tot, sm_s);
        int dex = block_excl_scan_int<XS_CHAIN_NT>(nd, &ndtot, sm_i);
        XsSeg cs = carry_seg; int cn = carry_nd;
        if (b < nt) { w.tin[b] = xs_seg_op(cs, ex).t; w.tdoff[b] = cn + dex; }
        __syncthreads();
        if (tid == 0) { carry_seg = xs_seg_op(cs, tot); carry_nd = cn + ndtot; }
        __syncthreads();
    }
    const int D = carry_nd;
    if (carry_seg.t.lvl == XS_BAD && tid == 0) w.flags[0] = 1;
    __syncthreads();
    // phase 2: apply the dirty entries in order (chunks staged through shared memory)
    if (!w.flags[0]) {
        for (int cbase = 0; cbase < D; cbase += XS_CHUNK) {
            // one (tile, entry) pair per thread and iteration (parallel fetch of a tile's entries)
            for (size_t idx = tid; idx < (size_t)nt * XS_MAXD; idx += XS_CHAIN_NT) {
                const unsigned b = (unsigned)(idx / XS_MAXD);
                const int e = (int)(idx % XS_MAXD);
                int nd = w.tnd[b];
                if (nd < 0) nd = 1;
                if (e >= nd) continue;
                const int o = w.tdoff[b] + e - cbase;
                if (o < 0 || o >= XS_CHUNK) continue;
                XsEntry en = w.ent[(size_t)b * XS_MAXD + e];
                if (e == 0) {
                    xs_t r; r.inc = en.inc; r.lvl = en.lvl;
                    r = xs_compose(w.tin[b], r);
                    en.inc = r.inc; en.lvl = r.lvl;
                }
                if (en.pad == 1) en.v = (double)b;
                sm_ent[o] = en;
            }
            __syncthreads();
            const int cnt = D - cbase < XS_CHUNK ? D - cbase : XS_CHUNK;
            // minimal serial part: per entry one integer add on the bit pattern and one FP add; certificates are checked
            // afterwards in parallel
            if (tid == 0) {
                double s = s_run;
                for (int o = 0; o < cnt; ++o) {
                    sm_before[o] = s;
                    s = pfc_u2d(pfc_d2u(s) + (unsigned long long)sm_ent[o].inc);
                    if (sm_ent[o].pad == 1) {               // overflow tile: walk it with genuine FP adds
                        unsigned b = (unsigned)sm_ent[o].v;
                        w.sbase[b] = s;
                        size_t lo = (size_t)b * XS_TILE, hi = lo + XS_TILE < n ? lo + XS_TILE : n;
                        for (size_t i = lo; i < hi; ++i) s = s + f(i);
                        w.flags[7] += 1;
                    } else {
                        s = s + sm_ent[o].v;
                    }
                    sm_after[o] = s;
                }
                s_run = s;
            }
            __syncthreads();
            for (int o = tid; o < cnt; o += XS_CHAIN_NT) {
                xs_t r; r.inc = sm_ent[o].inc; r.lvl = sm_ent[o].lvl;
                int ok = 1;
                (void)xs_apply(r, sm_before[o], &ok);
                if (!ok) ok_s = 0;
                w.s_after[cbase + o] = sm_after[o];
            }
            __syncthreads();
        }
        if (tid == 0) {
            int ok = ok_s;
            double tot = xs_apply(carry_seg.t, s_run, &ok);
            if (!ok) w.flags[0] = 1;
            else { *total_out = tot; w.flags[1] = D; w.flags[5] = D; }
        }
        __syncthreads();
    }
    // serial fallback: one thread, genuine left-to-right loop (exact by construction)
    if (w.flags[0]) {
        if (tid == 0) {
            double s = s_start;
            for (unsigned b = 0; b < nt; ++b) {
                w.sbase[b] = s;
                size_t lo = (size_t)b * XS_TILE, hi = lo + XS_TILE < n ? lo + XS_TILE : n;
                for (size_t i = lo; i < hi; ++i) s = s + f(i);
            }
            *total_out = s;
            w.flags[1] = -1;
            w.flags[4] += 1;
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// E: emit the exact inclusive prefix of every value
// ---------------------------------------------------------------------------------------------------
template <class F, class S>
__global__ void __launch_bounds__(XS_NT) xs_emit_tiles(F f, S sink, size_t n, double rel, XsWork w, double s_start) {
    __shared__ double sm_d[XS_NT / 32];
    __shared__ int sm_i[XS_NT / 32];
    __shared__ XsSeg sm_s[XS_NT / 32];
    XS_GATE(w);
    const unsigned b = blockIdx.x;
    if (w.flags[0] || w.tnd[b] < 0) {      // serial mode / overflow tile: walked by one thread from its exact base
        if (threadIdx.x == 0) {
            double s = w.sbase[b];
            size_t lo = (size_t)b * XS_TILE, hi = lo + XS_TILE < n ? lo + XS_TILE : n;
            for (size_t i = lo; i < hi; ++i) { s = s + f(i); sink(i, s); }
        }
        return;
    }
    const size_t first = (size_t)b * XS_TILE + (size_t)threadIdx.x * XS_ITEMS;
    double v[XS_ITEMS];
    double tsum = 0.0;
#pragma unroll
    for (int k = 0; k < XS_ITEMS; ++k) { size_t i = first + k; v[k] = i < n ? f(i) : 0.0; tsum += v[k]; }
    double btot;
    const double excl = block_excl_scan<XS_NT>(tsum, &btot, sm_d);
    const double toff = w.toff[b];
    XsThreadScan ts = xs_thread_scan(v, toff, excl, rel);
    XsSeg stot;
    XsSeg carry = xs_block_seg_excl<XS_NT>(xs_seg_make(ts.tail, ts.nd > 0), &stot, sm_s);
    carry = xs_seg_op(xs_seg_make(w.tin[b], 0), carry);
    int ndtot;
    int ord = w.tdoff[b] + block_excl_scan_int<XS_NT>(ts.nd, &ndtot, sm_i);
    double base_s = ord > 0 ? w.s_after[ord - 1] : (w.s_start_ptr ? *w.s_start_ptr : s_start);
    xs_t run = carry.t;
    double running = 0.0, a_prev = toff + excl;
    int ok = 1;
#pragma unroll
    for (int k = 0; k < XS_ITEMS; ++k) {
        running += v[k];
        double a_cur = toff + (excl + running);
        xs_t t;
        double c;
        if (xs_classify(v[k], a_prev, a_cur, rel, &t)) { run = xs_compose(run, t); c = xs_apply(run, base_s, &ok); }
        else { c = w.s_after[ord]; ord++; run = xs_identity(); base_s = c; }
        a_prev = a_cur;
        size_t i = first + k;
        if (i < n) sink(i, c);
    }
    if (!ok) { w.flags[2] = 1; w.flags[6] = 1; }
}

// ---------------------------------------------------------------------------------------------------
// host-side drivers.  n_margin = total number of terms of the (possibly multi-GPU) sum.
// ---------------------------------------------------------------------------------------------------
template <class F>
static int xs_total(Ctx& ctx, XsWork& w, F f, size_t n, size_t n_margin, double s_start, double* d_total) {
    unsigned nt = cdiv_u(n, XS_TILE);
    double rel = xs_margin(n_margin);
    PF_CUDA(cudaMemsetAsync(w.flags, 0, 4 * sizeof(int), ctx.stream));   // [0..3] are per-invocation
    PF_LAUNCH(ctx, xs_tile_sums<F>, nt, XS_NT, 0, f, n, w);
    PF_LAUNCH(ctx, xs_scan_tiles, 1, 1024, 0, nt, w);
    PF_LAUNCH(ctx, xs_classify_tiles<F>, nt, XS_NT, 0, f, n, rel, w);
    PF_LAUNCH(ctx, xs_chain<F>, 1, XS_CHAIN_NT, 0, f, n, nt, w, s_start, d_total);
    return 0;
}
template <class F, class S>
static int xs_scan(Ctx& ctx, XsWork& w, F f, S sink, size_t n, size_t n_margin, double s_start, double* d_total) {
    int rc = xs_total(ctx, w, f, n, n_margin, s_start, d_total);
    if (rc) return rc;
    unsigned nt = cdiv_u(n, XS_TILE);
    double rel = xs_margin(n_margin);
    PF_LAUNCH(ctx, (xs_emit_tiles<F, S>), nt, XS_NT, 0, f, sink, n, rel, w, s_start);
    return 0;
}

// ---- value functors / sinks ----
struct XsValArray { const double* p; __device__ __forceinline__ double operator()(size_t i) const { return p[i]; } };
struct XsSinkStore { double* c; __device__ __forceinline__ void operator()(size_t i, double x) const { c[i] = x; } };
