// xsum_sharded.cuh — the exact sequential-order sum / scan (xsum.cuh) over particles sharded across G GPUs (one process per
// GPU, NCCL): used by the sharded ParticleFilterLocalizer / MonteCarloLocalizer.
//
//     1. local approximate tile sums -> local total T_g            | ncclAllGather of G doubles   (approximate offsets)
//     2. classify with the GLOBAL approximate prefix, margin from the GLOBAL n; reduce the shard to its SUMMARY: the
//        ordered list of its dirty values with the integer increment of the clean run in front of each, plus the
//        increment of the trailing run                               | ncclAllGather of G x 3 KB     (summaries)
//     3. every rank walks the summaries of ranks 0..G-1 in order with genuine FP adds (tens of entries): this yields
//        the exact prefix at its own shard start, the exact value after each of its dirty values, and the exact
//        global total — without a rank-to-rank dependency chain.
#pragma once
#include <nccl.h>
#include "common.cuh"
#include "xsum.cuh"

#define PF_NCCL(call)                                                                                   \
    do {                                                                                                \
        ncclResult_t r__ = (call);                                                                      \
        if (r__ != ncclSuccess) {                                                                       \
            snprintf(g_pfgpu_err, sizeof(g_pfgpu_err), "%s:%d: %s -> %s", __FILE__, __LINE__, #call,    \
                     ncclGetErrorString(r__));                                                          \
            return PFGPU_ERR_NCCL;                                                                      \
        }                                                                                               \
    } while (0)

#define SH_SUM_CAP 126
#define SH_MAX_WORLD 16
struct __align__(8) ShardSummary {
    int count;                 // dirty values of the shard
    int bad;                   // 1: not summarisable (overflow tile, > SH_SUM_CAP dirty values, bad input value)
    long long tail_inc; int tail_lvl; int pad;
    XsEntry ent[SH_SUM_CAP];
};

struct FsShard {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
    double* t_loc = nullptr;        // [1]  local approximate total
    double* t_all = nullptr;        // [G]
    double* approx_off = nullptr;   // [1]  approximate sum of the shards before mine
    ShardSummary* sum_loc = nullptr;
    ShardSummary* sum_all = nullptr;   // [G]
    double* s_start = nullptr;      // [1]  exact prefix at my shard start
    int* err = nullptr;             // device: set when a shard was not summarisable
    double* cum_all = nullptr;      // [n_global]
    uint32_t* idx_all = nullptr;    // [n_global]
    double* pose_all = nullptr;     // [3][n_global]
    uint32_t* h_idx = nullptr;      // pinned [n_global]
    double* sendbuf = nullptr; size_t send_cap = 0;
    double* recvbuf = nullptr; size_t recv_cap = 0;
    size_t n_guest = 0, guest_used = 0;   // guest columns per landmark array / in use since the last compaction
    unsigned long long compactions = 0, imported = 0;
    double* best_loc = nullptr;     // [8]
    double* best_all = nullptr;     // [8*G]
};

// ---------------------------------------------------------------------------------------------------------------------
__global__ void sh_local_total_kernel(const double* tsum, unsigned nt, double* out) {
    __shared__ double sm[256 / 32];
    double a = 0.0;
    for (unsigned b = threadIdx.x; b < nt; b += 256) a += tsum[b];
    double t = block_sum<256>(a, sm);
    if (threadIdx.x == 0) *out = t;
}
__global__ void sh_offset_kernel(const double* t_all, int rank, double* approx_off) {
    double a = 0.0;
    for (int g = 0; g < rank; ++g) a += t_all[g];
    *approx_off = a;
}

// chain, part 1: segmented scan over my tiles + my shard summary (no walk yet)
__global__ void __launch_bounds__(XS_CHAIN_NT) sh_chain_summary_kernel(unsigned nt, XsWork w, ShardSummary* out) {
    __shared__ XsSeg sm_s[XS_CHAIN_NT / 32];
    __shared__ int sm_i[XS_CHAIN_NT / 32];
    __shared__ XsSeg carry_seg;
    __shared__ int carry_nd;
    __shared__ int bad_s;
    const int tid = threadIdx.x;
    if (tid == 0) { carry_seg = xs_seg_make(xs_identity(), 0); carry_nd = 0; bad_s = w.flags[0] ? 1 : 0; }
    __syncthreads();
    for (unsigned base = 0; base < nt; base += XS_CHAIN_NT) {
        unsigned b = base + tid;
        int nd = b < nt ? w.tnd[b] : 0;
        if (nd < 0) { nd = 1; bad_s = 1; }                    // overflow tile: not summarisable
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
    if (tid == 0) {
        if (D > SH_SUM_CAP || carry_seg.t.lvl == XS_BAD) bad_s = 1;
        out->count = D; out->tail_inc = carry_seg.t.inc; out->tail_lvl = carry_seg.t.lvl; out->pad = 0;
    }
    __syncthreads();
    if (!bad_s) {
        for (unsigned b = tid; b < nt; b += XS_CHAIN_NT) {
            int nd = w.tnd[b];
            if (nd <= 0) continue;
            int o0 = w.tdoff[b];
            xs_t tinb = w.tin[b];
            for (int e = 0; e < nd; ++e) {
                XsEntry en = w.ent[(size_t)b * XS_MAXD + e];
                if (e == 0) {
                    xs_t r; r.inc = en.inc; r.lvl = en.lvl;
                    r = xs_compose(tinb, r);
                    en.inc = r.inc; en.lvl = r.lvl;
                }
                out->ent[o0 + e] = en;
            }
        }
    }
    __syncthreads();
    if (tid == 0) out->bad = bad_s;
}

// chain, part 2: walk all shards' summaries in rank order (one thread; tens of entries)
__global__ void sh_chain_walk_kernel(const ShardSummary* all, int rank, int world, XsWork w, double* s_start_out, double* total_out, int* err) {
    if (threadIdx.x != 0) return;
    double s = 0.0; int ok = 1;
    for (int g = 0; g < world; ++g) {
        const ShardSummary* sm = &all[g];
        if (sm->bad) { ok = 0; break; }
        if (g == rank) *s_start_out = s;
        for (int e = 0; e < sm->count; ++e) {
            xs_t r; r.inc = sm->ent[e].inc; r.lvl = sm->ent[e].lvl;
            s = xs_apply(r, s, &ok);
            s = s + sm->ent[e].v;
            if (g == rank) w.s_after[e] = s;
        }
        xs_t tl; tl.inc = sm->tail_inc; tl.lvl = sm->tail_lvl;
        s = xs_apply(tl, s, &ok);
    }
    if (!ok) *err = 1;
    *total_out = s;
    w.flags[1] = all[rank].count;
}

// host drivers -----------------------------------------------------------------------------------------------------------
template <class F>
static int xs_total_sharded(Ctx& ctx, XsWork& w, FsShard& sh, F f, size_t n_local, size_t n_global, double* d_total) {
    unsigned nt = cdiv_u(n_local, XS_TILE);
    double rel = xs_margin(n_global);
    PF_CUDA(cudaMemsetAsync(w.flags, 0, 4 * sizeof(int), ctx.stream));
    PF_LAUNCH(ctx, xs_tile_sums<F>, nt, XS_NT, 0, f, n_local, w);
    PF_LAUNCH(ctx, sh_local_total_kernel, 1, 256, 0, w.tsum, nt, sh.t_loc);
    PF_NCCL(ncclAllGather(sh.t_loc, sh.t_all, 1, ncclDouble, sh.comm, ctx.stream));
    PF_LAUNCH(ctx, sh_offset_kernel, 1, 1, 0, sh.t_all, sh.rank, sh.approx_off);
    w.approx_offset_ptr = sh.approx_off;
    PF_LAUNCH(ctx, xs_scan_tiles, 1, 1024, 0, nt, w);
    PF_LAUNCH(ctx, xs_classify_tiles<F>, nt, XS_NT, 0, f, n_local, rel, w);
    PF_LAUNCH(ctx, sh_chain_summary_kernel, 1, XS_CHAIN_NT, 0, nt, w, sh.sum_loc);
    PF_NCCL(ncclAllGather(sh.sum_loc, sh.sum_all, sizeof(ShardSummary), ncclChar, sh.comm, ctx.stream));
    PF_LAUNCH(ctx, sh_chain_walk_kernel, 1, 32, 0, sh.sum_all, sh.rank, sh.world, w, sh.s_start, d_total, sh.err);
    w.approx_offset_ptr = nullptr;
    return 0;
}
template <class F, class S>
static int xs_scan_sharded(Ctx& ctx, XsWork& w, FsShard& sh, F f, S sink, size_t n_local, size_t n_global, double* d_total) {
    int rc = xs_total_sharded(ctx, w, sh, f, n_local, n_global, d_total);
    if (rc) return rc;
    unsigned nt = cdiv_u(n_local, XS_TILE);
    double rel = xs_margin(n_global);
    w.s_start_ptr = sh.s_start;
    PF_LAUNCH(ctx, (xs_emit_tiles<F, S>), nt, XS_NT, 0, f, sink, n_local, rel, w, 0.0);
    w.s_start_ptr = nullptr;
    return 0;
}

