// fs_mg.cuh — sharded FastSLAM 1.0 post-step over PEER MEMORY (NVLink / NVSwitch), no NCCL and no host round trip on
// the step path.  fs_sharded.cuh keeps the NCCL form (fallback when peer mapping is unavailable, PFGPU_SHARD_P2P=0).
//
// Every rank allocates one ARENA with an identical layout and maps the arenas of all peers (cudaIpc*).  The coupled part of
// fastslam_update (fs1.rs:258-265 + resample fs1.rs:206-231) then is the single-GPU fused kernel of fs_post.cuh with two
// changes:
//   * a tile's aggregates (approximate sum, trailing clean run, dirty entries) are PUSHED into every rank's arena at the
//     tile's GLOBAL id, so after a barrier each rank holds the aggregates of all G*ntl tiles locally and evaluates the
//     same exact chain as one GPU would — sums stay bit-identical to the sequential CPU order;
//   * the grid barrier becomes a cross-GPU barrier: every CTA adds 1 to an arrival counter in EVERY rank's arena
//     (red.sys over NVLink, after a system fence) and spins on its own rank's counter until G*ntl arrivals are in.
// The resample indices are searched in the global CDF (each rank pushes its exact slice to all peers); ancestors that
// live on another rank are PULLED: pose and map are read from the owner's arena through its lazy-clone ancestry and
// parked in guest columns (see fs_sharded.cuh).  Peers only read a rank's state between its post kernel and the
// "imports done" counter that its flip kernel waits for, so nothing they read is being rewritten.
#pragma once
#include "fs_post.cuh"
#include "fs_sharded.cuh"

#define MG_MAX_TILES 2048

struct MgDev {
    char* const* peer;          // device array [G]: base of every rank's arena (peer[rank] is my own)
    int G, rank;
    unsigned ntl, NT;           // tiles per rank / in total
    size_t o_ctr;               // 3 arrival counters, 128 B apart: post-kernel barriers | CDF slices published | imports done
    size_t o_bad;               // epoch-tagged "a non-finite / negative value was seen" (forces the exact serial walk everywhere)
    size_t o_tsum[FX_SLOTS], o_ttail[FX_SLOTS], o_tnd[FX_SLOTS], o_ent[FX_SLOTS];
    size_t o_cum_all;           // [n_global] exact CDF, every rank holds a full copy
    size_t o_w_raw, o_w;
    size_t o_px[2], o_py[2], o_pyaw[2], o_lm[2], o_anc[2], o_lmstate;
    unsigned* tgt;              // local [4]: arrivals already consumed on the three counters; [3] = launch epoch
    int* err;                   // local, sticky: 2 = a peer never arrived (timeout)
    unsigned* gcol;             // local [n]: guest column of each importing slot
    unsigned long long* plan;   // local [32]: 0 head-run length, 1 tail-run start, 2 imports of this resample, 3 guests in use,
                                //            4 imported (total), 5 eager rebuilds, 6 / 8 finished CTAs of the clone / search launch,
                                //            7 this resample rebuilds eagerly; trace (PFGPU_POST_TRACE):
                                //            13 on/off, 14 last stamp, 16+k accumulated ns up to stage k of a resample step
    size_t n_guest;
};

template <class T>
__device__ __forceinline__ T* mg_at(const MgDev& mg, int g, size_t off) { return reinterpret_cast<T*>(mg.peer[g] + off); }

// spin until *ctr has reached target (wrap-safe); gives up after about 15 s (a rank may legitimately be late by as
// much as its host is: first-launch module loading, a descheduled process) so that a missing peer is an error, not a hang
__device__ __forceinline__ bool mg_wait(const unsigned* ctr, unsigned target, int* err) {
    unsigned long long spins = 0;
    for (;;) {
        unsigned v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
        if ((int)(v - target) >= 0) return true;
        if (*(volatile int*)err) return false;
        if (++spins > (1ull << 24)) { *err = 2; return false; }      // ~1 us per probe: about 15 s
        __nanosleep(100);
    }
}
// trace: time since the previous stamp, accumulated per stage (one thread of the first CTA of each kernel calls it)
__device__ __forceinline__ void mg_stamp(const MgDev& mg, int k) {
    if (!mg.plan[13]) return;
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    if (k > 0) mg.plan[16 + k] += t - mg.plan[14];
    mg.plan[14] = t;
}
// one arrival per CTA on every rank's counter `which`
__device__ __forceinline__ void mg_arrive(const MgDev& mg, int which) {
    __syncthreads();
    if ((int)threadIdx.x < mg.G) {
        __threadfence_system();
        atomicAdd_system(mg_at<unsigned>(mg, threadIdx.x, mg.o_ctr) + 32 * which, 1u);
    }
}
__device__ __forceinline__ bool mg_barrier(const MgDev& mg, unsigned& target, int* s_ok) {
    mg_arrive(mg, 0);
    target += (unsigned)mg.G * mg.ntl;
    if (threadIdx.x == 0) *s_ok = mg_wait(mg_at<unsigned>(mg, mg.rank, mg.o_ctr), target, mg.err) ? 1 : 0;
    __syncthreads();
    return *s_ok != 0;
}

struct FxPubPeers {
    const MgDev& mg; unsigned me_; unsigned epoch;
    __device__ __forceinline__ unsigned me() const { return me_; }
    __device__ __forceinline__ void tsum(const FxSlot&, int si, unsigned b, double v) const {
        for (int g = 0; g < mg.G; ++g) mg_at<double>(mg, g, mg.o_tsum[si])[b] = v;
    }
    __device__ __forceinline__ void tail(const FxSlot&, int si, unsigned b, xs_t t, int nd) const {
        for (int g = 0; g < mg.G; ++g) { mg_at<xs_t>(mg, g, mg.o_ttail[si])[b] = t; mg_at<int>(mg, g, mg.o_tnd[si])[b] = nd; }
    }
    __device__ __forceinline__ void ent(const FxSlot&, int si, size_t k, const XsEntry& e) const {
        for (int g = 0; g < mg.G; ++g) mg_at<XsEntry>(mg, g, mg.o_ent[si])[k] = e;
    }
    __device__ __forceinline__ void bad(int*) const {
        for (int g = 0; g < mg.G; ++g) *mg_at<volatile unsigned>(mg, g, mg.o_bad) = epoch + 1u;
    }
    __device__ __forceinline__ int is_bad(const int*) const { return *mg_at<volatile unsigned>(mg, mg.rank, mg.o_bad) == epoch + 1u; }
};

// value at a GLOBAL index, fetched from its owner (only the rare serial walks / overflow tiles use these)
struct MgValRaw { const MgDev* mg; size_t nl;
    __device__ __forceinline__ double operator()(size_t i) const { size_t g = i / nl; return mg_at<double>(*mg, (int)g, mg->o_w_raw)[i - g * nl]; } };
struct MgValW { const MgDev* mg; size_t nl;
    __device__ __forceinline__ double operator()(size_t i) const { size_t g = i / nl; return mg_at<double>(*mg, (int)g, mg->o_w)[i - g * nl]; } };
struct MgValWSq { const MgDev* mg; size_t nl;
    __device__ __forceinline__ double operator()(size_t i) const { size_t g = i / nl; double x = mg_at<double>(*mg, (int)g, mg->o_w)[i - g * nl]; return x * x; } };
struct MgValW2 { const MgDev* mg; size_t nl; double S2;
    __device__ __forceinline__ double operator()(size_t i) const { size_t g = i / nl; double x = mg_at<double>(*mg, (int)g, mg->o_w)[i - g * nl]; return S2 > 0.0 ? x / S2 : x; } };

typedef FxSharedT<MG_MAX_TILES> MgShared;

// =====================================================================================================================
// Same phases as fs_post_kernel; comments there.  `fw.slot[]` point at MY arena's copies of the (global) tile arrays.
__global__ void __launch_bounds__(XS_NT, 2) fs_post_mg_kernel(FsDev d, FxWork fw, const __grid_constant__ MgDev mg, double nth, uint64_t seed,
                                                              double rel, const __grid_constant__ FsObsParam po, int k_obs) {
    extern __shared__ __align__(16) unsigned char mg_smem[];
    MgShared& sh = *reinterpret_cast<MgShared*>(mg_smem);
    __shared__ int s_ok;
    const unsigned bl = blockIdx.x;
    const unsigned b = (unsigned)mg.rank * mg.ntl + bl;          // global tile id
    const int tid = threadIdx.x;
    const size_t firstl = (size_t)bl * FX_TILE + (size_t)tid * FX_ITEMS;
    const size_t n = d.n, ng = d.n_global, nl = d.n;
    unsigned target = mg.tgt[0];
    const unsigned epoch = mg.tgt[3];
    const FxPubPeers pub{mg, b, epoch};
    const unsigned NT = mg.NT;
    unsigned long long t_prev = fx_now();
#define MG_FINISH() do { if (bl == 0 && tid == 0) { mg.tgt[0] = target; mg.tgt[3] = epoch + 1u; } } while (0)
#define MG_BARRIER() do { if (!mg_barrier(mg, target, &s_ok)) { MG_FINISH(); return; } } while (0)
    if (bl == 0 && tid == 0 && fw.dbg) fw.dbg[31] += 1;
    if (bl == 0 && tid == 0) mg_stamp(mg, 0);
    if (bl == 0 && tid < k_obs) fs_mark_updated(d, po.o[tid].lm_id);   // lazy-clone bookkeeping of the EKF launch that just ran
    double v[FX_ITEMS];
#pragma unroll
    for (int k = 0; k < FX_ITEMS; ++k) { size_t i = firstl + k; v[k] = i < n ? d.w_raw[i] : 0.0; }
    // ---------------- S = sum w_raw; w = w_raw / S (fs1.rs:196-203) ----------------
    fx_tile_sum(v, fw.slot[0], 0, sh, fw.flags, pub);
    MG_BARRIER();
    if (bl == 0 && tid == 0 && fw.dbg) { unsigned long long t = fx_now(); fw.dbg[0] += t - t_prev; t_prev = t; }
    fx_classify(v, fw.slot[0], 0, sh, rel, pub);
    MG_BARRIER();
    if (bl == 0 && tid == 0 && fw.dbg) { unsigned long long t = fx_now(); fw.dbg[1] += t - t_prev; t_prev = t; }
    fx_chain(fw.slot[0], NT, MgValRaw{&mg, nl}, ng, sh, fw.flags, pub);
    const double S = sh.total;
    double q[FX_ITEMS];
#pragma unroll
    for (int k = 0; k < FX_ITEMS; ++k) {
        size_t i = firstl + k;
        if (S > 0.0) v[k] = v[k] / S;
        if (i < n) d.w[i] = v[k];
        q[k] = v[k] * v[k];
    }
    // ---------------- gate: neff = 1 / sum w^2 < NTH (fs1.rs:186-193, 262-263) ----------------
    fx_tile_sum(q, fw.slot[1], 1, sh, fw.flags, pub);
    fx_tile_sum(v, fw.slot[2], 2, sh, fw.flags, pub);
    // A cross-GPU barrier costs more than a classification, so S2 = sum w (needed only if the gate opens) is classified
    // speculatively and published under the same barrier; its approximate tile prefix comes from the tile sums of w_raw / S.
    double toff_w = fx_tile_offset(fw.slot[0], b, sh);
    if (S > 0.0) toff_w = toff_w / S;
    fx_classify_at(v, toff_w, fw.slot[2], 2, sh, rel, pub, fw.flags);
    MG_BARRIER();
    if (bl == 0 && tid == 0 && fw.dbg) { unsigned long long t = fx_now(); fw.dbg[2] += t - t_prev; t_prev = t; }
    double qpart = 0.0;
    for (unsigned t = tid; t < NT; t += XS_NT) qpart += fw.slot[1].tsum[t];
    double qa = block_sum<XS_NT>(qpart, sh.sm_d);
    __syncthreads();
    if (tid == 0) sh.total = qa;
    __syncthreads();
    double Q = sh.total;                                       // bit-identical on every rank (same values, same order)
    double neff = Q > 0.0 ? 1.0 / Q : 0.0;
    const double slack = 8.0 * (double)(ng + 64) * 2.220446049250313e-16;
    const bool border = !(fabs(neff - nth) > slack * fmax(fabs(nth), fabs(neff))) || pub.is_bad(nullptr);
    if (border) {
        fx_classify(q, fw.slot[1], 1, sh, rel, pub);
        MG_BARRIER();
        fx_chain(fw.slot[1], NT, MgValWSq{&mg, nl}, ng, sh, fw.flags, pub);
        Q = sh.total;
        neff = Q > 0.0 ? 1.0 / Q : 0.0;
    }
    const int gate = neff < nth ? 1 : 0;
    if (bl == 0 && tid == 0) {
        d.scal[0] = S; d.scal[1] = Q; d.scal[3] = neff;
        *d.gate = gate;
    }
    if (!gate) { MG_FINISH(); return; }
    if (bl == 0 && tid == 0) { mg.plan[6] = 0; mg.plan[8] = 0; }     // completion counters of the two launches that follow
    // ---------------- resample: S2 = sum w (fs1.rs:207) ----------------
    fx_chain(fw.slot[2], NT, MgValW{&mg, nl}, ng, sh, fw.flags, pub);
    const double S2 = sh.total;
    if (bl == 0 && tid == 0) d.scal[2] = S2;
    const double inv = 1.0 / (double)ng;
    if (tid == 0) {
        double u01 = pfc_u01_52(pfc_blk_u64(pfc_rng_block(seed, PFC_STREAM_FS_RESAMPLE, d.counters[0], 0), 0));
        sh.r0 = u01 * (inv - 0.0) + 0.0;                       // Uniform::new(0, 1/n).sample
    }
    __syncthreads();
    const double r0 = sh.r0;
    double cv[FX_ITEMS];
#pragma unroll
    for (int k = 0; k < FX_ITEMS; ++k) {
        size_t i = firstl + k;
        if (S2 > 0.0) v[k] = v[k] / S2;
        cv[k] = i < n ? ((d.offset + i) == 0 ? r0 : inv) : 0.0;
    }
    // approximate tile prefixes without another publish + barrier (see fs_post_kernel)
    double toff_c = fx_tile_offset(fw.slot[2], b, sh);
    if (S2 > 0.0) toff_c = toff_c / S2;
    const double toff_r = b == 0 ? 0.0 : r0 + ((double)((size_t)b * FX_TILE) - 1.0) * inv;
    FxTile tc = fx_classify_at(v, toff_c, fw.slot[3], 3, sh, rel, pub, fw.flags);
    FxTile tr = fx_classify_at(cv, toff_r, fw.slot[4], 4, sh, rel, pub, fw.flags);
    MG_BARRIER();
    if (bl == 0 && tid == 0 && fw.dbg) { unsigned long long t = fx_now(); fw.dbg[3] += t - t_prev; t_prev = t; }
    double c[FX_ITEMS], r[FX_ITEMS];
    fx_chain(fw.slot[3], NT, MgValW2{&mg, nl, S2}, ng, sh, fw.flags, pub);
    fx_emit(v, tc, sh, MgValW2{&mg, nl, S2}, ng, rel, c, fw.flags, d.cum - d.offset, pub);
    fx_chain(fw.slot[4], NT, FxValComb{r0, inv}, ng, sh, fw.flags, pub);
    fx_emit(cv, tr, sh, FxValComb{r0, inv}, ng, rel, r, fw.flags, d.rcomb - d.offset, pub);
#pragma unroll
    for (int k = 0; k < FX_ITEMS; ++k) {
        size_t i = firstl + k;
        if (i < n) {
            d.cum[i] = c[k]; d.rcomb[i] = r[k];
            for (int g = 0; g < mg.G; ++g) mg_at<double>(mg, g, mg.o_cum_all)[d.offset + i] = c[k];   // my CDF slice -> everyone
        }
    }
    mg_arrive(mg, 1);                                          // fs_mg_search_pose_kernel waits for G*ntl of these
    if (bl == 0 && tid == 0 && fw.dbg) { unsigned long long t = fx_now(); fw.dbg[4] += t - t_prev; t_prev = t; }
    MG_FINISH();
#undef MG_BARRIER
#undef MG_FINISH
}

// ---- resample, part 2 -------------------------------------------------------------------------------------------------
// number of leading slots whose ancestor index is below `bound` (idx ascends): all NT threads of the CTA, two rounds
template <int NT>
__device__ __forceinline__ unsigned mg_count_below(const uint32_t* idx, size_t n, unsigned long long bound, unsigned* s_acc) {
    const size_t chunk = (n + NT - 1) / NT;
    const size_t t0 = (size_t)threadIdx.x * chunk;
    const int c1 = __syncthreads_count(t0 < n && (unsigned long long)__ldcg(idx + t0) < bound);   // chunk starts below: boundary is in chunk c1-1
    if (c1 == 0) return 0;
    const size_t base = (size_t)(c1 - 1) * chunk;
    if (threadIdx.x == 0) *s_acc = 0;
    __syncthreads();
    unsigned cnt = 0;
    for (size_t o = threadIdx.x; o < chunk; o += NT) { const size_t u = base + o; cnt += (u < n && (unsigned long long)__ldcg(idx + u) < bound) ? 1u : 0u; }
    if (cnt) atomicAdd(s_acc, cnt);
    __syncthreads();
    const unsigned r = (unsigned)base + *s_acc;
    __syncthreads();
    return r;
}
// Which of my slots have an ancestor on another rank, and which guest column each of them gets.  The ancestry is
// monotone, so these slots are a head run [0, nh) (ancestors below my block) and a tail run [t1, n) (above); slots that
// share an ancestor share a guest column (the first of them copies).  If the guests left do not suffice, this rank
// rebuilds its shard EAGERLY in this resample instead (plan[7] = 1): every landmark of every slot is copied (pulled, if
// remote) into its own column of the other buffer, which needs no guest column and frees all of them.
template <int NT>
__device__ __forceinline__ void mg_plan(const FsDev& d, const MgDev& mg, int* sm_i /* NT/32 */, unsigned* s_u /* 4 */) {
    const int tid = threadIdx.x;
    const size_t n = d.n;
    const unsigned nh = mg_count_below<NT>(d.idx, n, (unsigned long long)d.offset, &s_u[3]);
    const unsigned t1 = mg_count_below<NT>(d.idx, n, (unsigned long long)d.offset + n, &s_u[3]);
    if (tid == 0) { s_u[0] = nh; s_u[1] = t1; s_u[2] = 0; }
    __syncthreads();
    const unsigned base = (unsigned)mg.plan[3];
    for (int run = 0; run < 2; ++run) {
        const size_t start = run == 0 ? 0 : s_u[1], end = run == 0 ? s_u[0] : n;
        for (size_t c = start; c < end; c += NT) {
            const size_t t = c + tid;
            const int flag = (t < end && (t == start || __ldcg(d.idx + t) != __ldcg(d.idx + t - 1))) ? 1 : 0;
            int tot;
            const int ex = block_excl_scan_int<NT>(flag, &tot, sm_i);
            if (t < end) mg.gcol[t] = (unsigned)n + base + s_u[2] + (unsigned)(ex + flag - 1);
            __syncthreads();
            if (tid == 0) s_u[2] += (unsigned)tot;
            __syncthreads();
        }
    }
    if (tid == 0) {
        const unsigned need = s_u[2];
        mg.plan[0] = s_u[0]; mg.plan[1] = s_u[1]; mg.plan[2] = need;
        mg.plan[4] += need;
        if ((size_t)base + need > mg.n_guest) { mg.plan[7] = 1; mg.plan[5] += 1; }
        else { mg.plan[7] = 0; mg.plan[3] = base + need; }
    }
}

// index walk fs1.rs:224-226 in the global CDF + pose clone fs1.rs:227-229 (the ancestor's pose is pulled from its owner);
// the last CTA to finish plans the map import
__global__ void __launch_bounds__(256) fs_mg_search_plan_kernel(FsDev d, const __grid_constant__ MgDev mg) {
    if (!*d.gate) return;
    __shared__ int s_ok, s_last;
    __shared__ int sm_i[8];
    __shared__ unsigned s_u[4];
    if (blockIdx.x == 0 && threadIdx.x == 0) mg_stamp(mg, 1);
    if (threadIdx.x == 0)
        s_ok = mg_wait(mg_at<unsigned>(mg, mg.rank, mg.o_ctr) + 32, mg.tgt[1] + (unsigned)mg.G * mg.ntl, mg.err) ? 1 : 0;
    __syncthreads();
    const size_t t = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (s_ok && t < d.n) {
        const double* __restrict__ cum_all = mg_at<double>(mg, mg.rank, mg.o_cum_all);
        const double r = d.rcomb[t];
        size_t lo = 0, hi = d.n_global;
        while (lo < hi) {
            size_t mid = lo + ((hi - lo) >> 1);
            if (cum_all[mid] < r) lo = mid + 1; else hi = mid;
        }
        const size_t j = lo < d.n_global ? lo : d.n_global - 1;
        d.idx[t] = (uint32_t)j;
        const int cur = *d.cur;                                // the same on every rank: all ranks flip together
        const size_t nl = d.n, g = j / nl, jl = j - g * nl;
        fs_px(d, cur ^ 1)[t] = mg_at<double>(mg, (int)g, cur ? mg.o_px[1] : mg.o_px[0])[jl];
        fs_py(d, cur ^ 1)[t] = mg_at<double>(mg, (int)g, cur ? mg.o_py[1] : mg.o_py[0])[jl];
        fs_pyaw(d, cur ^ 1)[t] = mg_at<double>(mg, (int)g, cur ? mg.o_pyaw[1] : mg.o_pyaw[0])[jl];
        d.w[t] = 1.0 / (double)d.n_global;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(&mg.plan[8], 1ull) + 1 == gridDim.x) ? 1 : 0;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (threadIdx.x == 0) mg_stamp(mg, 2);
    if (*(volatile int*)mg.err) return;                        // a peer went missing: idx is not valid, nothing to plan
    mg_plan<256>(d, mg, sm_i, s_u);
}

// One launch finishes the resample (fs1.rs:227-229 for the maps, then the ping-pong flip):
//   CTAs with blockIdx.y <  MG_IMPORT_Y: PULL the maps of remote ancestors through their owner's lazy-clone ancestry into my
//                                        guest columns (NVLink reads; latency-bound, so they start first and overlap the rest)
//   CTAs with blockIdx.y >= MG_IMPORT_Y: lazy clone — anc'[l][t] = guest column for imported slots, composed local ancestry
//                                        otherwise; 4 slots per thread, one 16-byte store per landmark row
//                                        (eager rebuild, plan[7]: copy the landmarks themselves instead, see mg_plan)
//   the last CTA to finish tells every rank that my reads of their state are over, waits for the same from all of them,
//   and flips the ping-pong state (nobody may see the flip while still reading the old state).
#define MG_IMPORT_Y 4
__global__ void __launch_bounds__(256) fs_mg_clone_kernel(FsDev d, const __grid_constant__ MgDev mg) {
    if (!*d.gate) return;
    __shared__ int s_last, s_ok;
    const size_t n = d.n, nl = d.n;
    const int eager = (int)mg.plan[7];
    const int ac = *d.anc_cur;                                     // the same on every rank
    const bool dead = *(volatile int*)mg.err != 0;                 // a peer went missing: skip the data movement, keep the protocol
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) mg_stamp(mg, 3);
    if (blockIdx.y < MG_IMPORT_Y) {
        if (!eager && !dead) {
            const size_t vb = (size_t)blockIdx.y * gridDim.x + blockIdx.x, nvb = (size_t)MG_IMPORT_Y * gridDim.x;
            const size_t rows = 6 * d.m;
            const size_t nh = (size_t)mg.plan[0], t1 = (size_t)mg.plan[1];
            const size_t cnt = nh + (n - t1), total = cnt * rows;
            for (size_t e = vb * 256 + threadIdx.x; e < total; e += nvb * 256) {
                const size_t row = e / cnt, q = e - row * cnt;
                const size_t t = q < nh ? q : t1 + (q - nh);
                if (!(t == 0 || t == t1 || d.idx[t] != d.idx[t - 1])) continue;     // a neighbour copies this ancestor
                const size_t l = row / 6; const int f = (int)(row - l * 6);
                const size_t j = d.idx[t], g = j / nl, jl = j - g * nl;
                const int rst = mg_at<int>(mg, (int)g, mg.o_lmstate)[l];
                const size_t col = (rst & 2) ? jl : (size_t)mg_at<uint32_t>(mg, (int)g, ac ? mg.o_anc[1] : mg.o_anc[0])[l * nl + jl];
                const double val = mg_at<double>(mg, (int)g, (rst & 1) ? mg.o_lm[1] : mg.o_lm[0])[lm_index(d.ld, l, f, col)];
                const int st = d.lmstate[l];
                fs_lm(d, st & 1)[lm_index(d.ld, l, f, (size_t)mg.gcol[t])] = val;
            }
        }
    } else {
        const size_t t = ((size_t)blockIdx.x * 256 + threadIdx.x) * 4;
        const size_t l0 = (size_t)(blockIdx.y - MG_IMPORT_Y) * FS_COMPOSE_ROWS;
        if (t < n && !eager && !dead) {
            const uint32_t* __restrict__ src = fs_anc(d, ac);
            uint32_t* __restrict__ dst = fs_anc(d, ac ^ 1);
            const uint4 jj = *reinterpret_cast<const uint4*>(d.idx + t);
            const size_t lo = d.offset, hi = d.offset + nl;
            const bool r0 = jj.x < lo || jj.x >= hi, r1 = jj.y < lo || jj.y >= hi, r2 = jj.z < lo || jj.z >= hi, r3 = jj.w < lo || jj.w >= hi;
            uint4 gc = make_uint4(0, 0, 0, 0);
            if (r0 | r1 | r2 | r3) gc = *reinterpret_cast<const uint4*>(mg.gcol + t);
            const uint32_t a0 = jj.x - (uint32_t)lo, a1 = jj.y - (uint32_t)lo, a2 = jj.z - (uint32_t)lo, a3 = jj.w - (uint32_t)lo;
#pragma unroll
            for (int rr = 0; rr < FS_COMPOSE_ROWS; ++rr) {
                const size_t l = l0 + rr;
                if (l >= d.m) break;
                const bool ident = (d.lmstate[l] & 2) != 0;
                const uint32_t* __restrict__ row = src + l * nl;
                uint4 o;
                o.x = r0 ? gc.x : (ident ? a0 : row[a0]);
                o.y = r1 ? gc.y : (ident ? a1 : row[a1]);
                o.z = r2 ? gc.z : (ident ? a2 : row[a2]);
                o.w = r3 ? gc.w : (ident ? a3 : row[a3]);
                *reinterpret_cast<uint4*>(dst + l * nl + t) = o;
            }
        } else if (!dead) {
            // eager rebuild: lm[other][l][.][t] = the ancestor's landmark, local or remote, through its owner's ancestry.
            // This CTA owns slots [1024 bx, +1024) x landmarks [l0, l0 + ROWS): consecutive threads take consecutive slots
            // (coalesced columns), iterations are independent (several gathers in flight per thread).
            const size_t tb = (size_t)blockIdx.x * 1024;
#pragma unroll 4
            for (int it = 0; it < 4 * FS_COMPOSE_ROWS; ++it) {
                const int p = it * 256 + threadIdx.x;
                const size_t tk = tb + (size_t)(p & 1023), l = l0 + (size_t)(p >> 10);
                if (tk >= n || l >= d.m) continue;
                const size_t j = d.idx[tk], g = j / nl, jl = j - g * nl;
                const int rst = mg_at<int>(mg, (int)g, mg.o_lmstate)[l];
                const size_t col = (rst & 2) ? jl : (size_t)mg_at<uint32_t>(mg, (int)g, ac ? mg.o_anc[1] : mg.o_anc[0])[l * nl + jl];
                const double* __restrict__ sl = mg_at<double>(mg, (int)g, (rst & 1) ? mg.o_lm[1] : mg.o_lm[0]);
                double* __restrict__ o = fs_lm(d, (d.lmstate[l] & 1) ^ 1);
#pragma unroll
                for (int f = 0; f < 6; ++f) o[lm_index(d.ld, l, f, tk)] = sl[lm_index(d.ld, l, f, col)];
            }
        }
    }
    // ---- completion: the last CTA signals, waits for all ranks, flips ----
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(&mg.plan[6], 1ull) + 1 == (unsigned long long)gridDim.x * gridDim.y) ? 1 : 0;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (threadIdx.x == 0) {
        mg_stamp(mg, 5);
        for (int g = 0; g < mg.G; ++g) atomicAdd_system(mg_at<unsigned>(mg, g, mg.o_ctr) + 64, 1u);   // my reads of your state are over
        s_ok = mg_wait(mg_at<unsigned>(mg, mg.rank, mg.o_ctr) + 64, mg.tgt[2] + (unsigned)mg.G, mg.err) ? 1 : 0;
    }
    __syncthreads();
    for (size_t l = threadIdx.x; l < d.m; l += blockDim.x) {
        const int st = __ldcg(d.lmstate + l);
        d.lmstate[l] = eager ? (((st & 1) ^ 1) | 2) : (st & 1);   // eager: every landmark now sits in its own column of the other buffer
    }
    if (threadIdx.x == 0) {
        *d.cur ^= 1; *d.anc_cur ^= 1; d.counters[0] += 1;
        mg.tgt[1] += (unsigned)mg.G * mg.ntl; mg.tgt[2] += (unsigned)mg.G;
        if (eager) mg.plan[3] = 0;
        mg_stamp(mg, 6);
    }
}
