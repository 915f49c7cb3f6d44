// fs3_host.cuh — host side of the FastSLAM 1.0 engine (kernels: fs3.cuh): the pfgpu_fs_* entry points of include/pfgpu.h.
// Included by pfgpu.cu after the shared helpers (Ctx, Marks, KernelTimer, PF_NCCL, ...).
#pragma once
#include "fs3.cuh"

struct pfgpu_fs {
    Ctx ctx;
    pfgpu_fs_config cfg;
    uint64_t seed = 0;
    Fs3Dev d = {};
    uint32_t n_step = 0;
    uint64_t steps = 0;
    int world = 1, rank = 0;
    KernelTimer timer;
    Marks marks;
    char* arena = nullptr; size_t arena_bytes = 0;
    void* peer_ptr[FS3_MAXG] = {};
    ncclComm_t comm = nullptr;
    Fs3Rec* h_rec = nullptr;          // pinned + mapped
    double* stage = nullptr; size_t stage_bytes = 0;   // device staging buffer for upload / download / seed_map
    bool pdl = true;
    bool early = false;               // PFGPU_EARLY_LAUNCH=1: release the dependent kernel at the START of the previous grid (measured: 2 % slower;
                                      // releasing the next EKF launch when the post kernel's CTAs are through their phases: also 1.7 % slower)
    bool ekf_attr[2] = { false, false };
    int variant = 1;                  // 1 = FastSLAM 1.0 (fs1.rs), 2 = FastSLAM 2.0 (fs2.rs); pfgpu_fs_set_variant
    int ekf_helpers = 0;              // PFGPU_EKF_HELPERS: cap on the helper warps per CTA (0 = as many as fit, at most 3)
    int post_nt = 256; unsigned post_K = 1, post_tiles = 1, m32 = 0; int log2n = -1; size_t post_smem = 0;
};

extern "C" void pfgpu_fs_default_config(pfgpu_fs_config* c) {            // fs1.rs:13-23
    c->dt = 0.1; c->max_range = 20.0; c->nth = 100.0 / 1.5; c->q00 = 0.3; c->q11 = 0.0305; c->r00 = 0.5; c->r11 = 0.0305;
    c->init_weight = 1.0 / 100.0;
}

template <int NT>
static int fs3_post_prepare(pfgpu_fs* h) {
    if (cudaFuncSetAttribute(fs3_post_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->post_smem) != cudaSuccess) { cudaGetLastError(); return 1; }
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, fs3_post_kernel<NT>, NT, h->post_smem) != cudaSuccess) { cudaGetLastError(); return 1; }
    return (size_t)nb * (size_t)h->ctx.num_sms >= h->post_tiles ? 0 : 1;
}

static int fs_stage(pfgpu_fs* h, size_t bytes) {
    if (bytes <= h->stage_bytes) return 0;
    if (h->stage) { cudaFree(h->stage); h->stage = nullptr; h->stage_bytes = 0; }
    PF_CUDA(cudaMalloc(&h->stage, bytes));
    h->stage_bytes = bytes;
    return 0;
}

static int fs_create_impl(const pfgpu_fs_config* cfg, size_t n, size_t n_global, size_t offset, size_t m, uint64_t seed, int device,
                          const void* uid, int rank, int world, pfgpu_fs** out) {
    if (!out) return PFGPU_ERR_INVALID;
    *out = nullptr;
    if (!cfg || n == 0 || n_global >= (1ull << 28) * (size_t)world || n >= (1ull << 28) || n_global > 0xFFFFFFF0ull) return PFGPU_ERR_INVALID;
    if (m > 1024) { snprintf(g_pfgpu_err, sizeof(g_pfgpu_err), "more than 1024 landmarks per particle are not supported"); return PFGPU_ERR_UNSUPPORTED; }
    if (world > 1 && n % 64 != 0) { snprintf(g_pfgpu_err, sizeof(g_pfgpu_err), "sharded FastSLAM needs a multiple of 64 particles per GPU"); return PFGPU_ERR_UNSUPPORTED; }
    pfgpu_fs* h = new (std::nothrow) pfgpu_fs();
    if (!h) return PFGPU_ERR_CUDA;
    int rc = ctx_open(h->ctx, device);
    if (rc) { delete h; return rc; }
    h->cfg = *cfg; h->seed = seed; h->world = world; h->rank = rank;
    Fs3Dev& d = h->d;
    const size_t ld = (n + 63) / 64 * 64, mm = m ? m : 1;
    d.n = (unsigned)n; d.n_glob = (unsigned)n_global; d.off = (unsigned)offset; d.m = (unsigned)m; d.ld = (unsigned)ld;
    d.G = world; d.rank = rank; d.npart = (unsigned)(ld / 64); d.wait_inline = 1;
    auto fail = [&](int code) { pfgpu_fs_destroy(h); return code; };
#define FS_TRY(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) { snprintf(g_pfgpu_err, sizeof(g_pfgpu_err), "%s -> %s", #x, cudaGetErrorString(e__)); return fail(PFGPU_ERR_CUDA); } } while (0)
    // post kernel shape: <= 128 tiles (one CTA each, co-resident), NT threads x K values
    {
        // below 65 536 weights: up to 128 tiles of 256 threads (latency matters, not throughput); at 65 536: 128 tiles of 512 threads,
        // one value per thread (measured 1.3 % faster than 256 x 2; fewer, fatter tiles are slower: 64 x 512 x 2 -2 %, 32 x 512 x 4
        // -13 %); beyond that the per-value phases (classify, normalise, emit: ~100 instructions per value and sum) dominate, so
        // every SM gets a tile of 512 threads
        const char* env = getenv("PFGPU_POST_NT");
        const bool big = n_global > (size_t)128 * 256 * 2;
        h->post_nt = env ? (atoi(env) == 512 ? 512 : 256) : (n_global >= (size_t)128 * 512 ? 512 : 256);
        unsigned want = (unsigned)std::min<int>(big ? 148 : 128, h->ctx.num_sms);
        { const char* ew = getenv("PFGPU_POST_TILES"); if (ew && atoi(ew) >= 1 && atoi(ew) <= (int)want) want = (unsigned)atoi(ew); }   // tests: several values per thread at small n
        unsigned K = (unsigned)((n_global + (size_t)want * h->post_nt - 1) / ((size_t)want * h->post_nt));
        if (K == 0) K = 1;
        h->post_K = K;
        h->post_tiles = (unsigned)((n_global + (size_t)h->post_nt * K - 1) / ((size_t)h->post_nt * K));
        h->post_smem = (size_t)h->post_nt * K * sizeof(double);
        h->m32 = x3_margin32(n_global);
        h->log2n = -1;
        for (int p = 0; p < 32; ++p) if (((size_t)1 << p) == n_global) h->log2n = p;
        int bad = h->post_nt == 512 ? fs3_post_prepare<512>(h) : fs3_post_prepare<256>(h);
        if (bad || h->post_tiles > FS3_MAX_TILES || h->post_tiles > (unsigned)h->post_nt) {
            snprintf(g_pfgpu_err, sizeof(g_pfgpu_err), "%zu particles do not fit the fused post-step kernel (%u tiles x %zu B of shared memory)", n_global, h->post_tiles, h->post_smem);
            return fail(PFGPU_ERR_UNSUPPORTED);
        }
    }
    // ONE allocation for everything a peer may touch, same layout on every rank: one IPC mapping per peer exposes all of it
    {
        size_t off = 0;
        auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
        d.o_flags = take(2 * FS3_MAXG * 128);
        const size_t ngp = (size_t)world * ld + 64;
        for (int b = 0; b < 2; ++b) { d.o_wraw[b] = take(ngp * sizeof(double)); d.o_part[b] = take((size_t)world * d.npart * sizeof(double)); }
        for (int b = 0; b < 2; ++b) { d.o_px[b] = take(ld * sizeof(double)); d.o_py[b] = take(ld * sizeof(double)); d.o_pyaw[b] = take(ld * sizeof(double)); }
        for (int b = 0; b < 2; ++b) d.o_rows[b] = take(mm * ld * sizeof(unsigned));
        const size_t small = off;
        for (int b = 0; b < 2; ++b) d.o_lm[b] = take(mm * 6 * ld * sizeof(double));
        h->arena_bytes = off;
        FS_TRY(cudaMalloc(&h->arena, off));
        FS_TRY(cudaMemset(h->arena, 0, small));
        FS_TRY(cudaMemset(h->arena + d.o_lm[0], 0, mm * 6 * ld * sizeof(double)));
        FS_TRY(cudaMemset(h->arena + d.o_lm[1], 0, mm * 6 * ld * sizeof(double)));
        char* A = h->arena;
        for (int b = 0; b < 2; ++b) {
            d.px[b] = (double*)(A + d.o_px[b]); d.py[b] = (double*)(A + d.o_py[b]); d.pyaw[b] = (double*)(A + d.o_pyaw[b]);
            d.lm[b] = (double*)(A + d.o_lm[b]); d.rows[b] = (unsigned*)(A + d.o_rows[b]);
            d.wraw[b] = (double*)(A + d.o_wraw[b]); d.part[b] = (double*)(A + d.o_part[b]);
        }
        d.peer[rank] = A; h->peer_ptr[rank] = A;
    }
    {
        Fs3State* stp = nullptr;
        FS_TRY(cudaMalloc(&stp, sizeof(Fs3State))); FS_TRY(cudaMemset(stp, 0, sizeof(Fs3State)));
        d.st = stp;
        FS_TRY(cudaMalloc(&d.lmst, mm * sizeof(int))); FS_TRY(cudaMemset(d.lmst, 0, mm * sizeof(int)));
        FS_TRY(cudaMalloc(&d.w, ld * sizeof(double))); FS_TRY(cudaMemset(d.w, 0, ld * sizeof(double)));
        for (int b = 0; b < 2; ++b) { FS_TRY(cudaMalloc(&d.nz[b], ld * sizeof(double))); FS_TRY(cudaMemset(d.nz[b], 0, ld * sizeof(double))); }
        FS_TRY(cudaMalloc(&d.wn_all, (n_global + 64) * sizeof(double)));
        FS_TRY(cudaMalloc(&d.cum_all, (n_global + 64) * sizeof(double)));
        if (h->log2n < 0) FS_TRY(cudaMalloc(&d.rcomb_all, (n_global + 64) * sizeof(double)));
        FS_TRY(cudaMalloc(&d.idx, ld * sizeof(unsigned))); FS_TRY(cudaMemset(d.idx, 0, ld * sizeof(unsigned)));
        const size_t nsl = (size_t)FS3_SLOTS * FS3_MAX_TILES, nen = (size_t)FS3_SLOTS * FS3_ENT_CAP;
        FS_TRY(cudaMalloc(&d.tileP, nsl * sizeof(unsigned long long)));
        FS_TRY(cudaMalloc(&d.tileQ, FS3_MAX_TILES * sizeof(double)));
        FS_TRY(cudaMalloc(&d.entCnt, 8 * sizeof(unsigned))); FS_TRY(cudaMemset(d.entCnt, 0, 8 * sizeof(unsigned)));
        FS_TRY(cudaMalloc(&d.entKey, nen * sizeof(unsigned))); FS_TRY(cudaMalloc(&d.entTile, nen * sizeof(unsigned)));
        FS_TRY(cudaMalloc(&d.entP, nen * sizeof(unsigned long long))); FS_TRY(cudaMalloc(&d.entV, nen * sizeof(double)));
        FS_TRY(cudaMalloc(&d.entL, nen * sizeof(int)));
        FS_TRY(cudaMalloc(&d.bar, 8 * sizeof(unsigned))); FS_TRY(cudaMemset(d.bar, 0, 8 * sizeof(unsigned)));
        FS_TRY(cudaMalloc(&d.resflag, 8 * sizeof(unsigned))); FS_TRY(cudaMemset(d.resflag, 0, 8 * sizeof(unsigned)));
        FS_TRY(cudaMalloc(&d.res, 8 * sizeof(Fs3Res))); FS_TRY(cudaMemset(d.res, 0, 8 * sizeof(Fs3Res)));
        FS_TRY(cudaMalloc(&d.resTP, (size_t)8 * FS3_MAX_TILES * sizeof(unsigned long long)));
        FS_TRY(cudaMalloc(&d.resKey, (size_t)8 * FS3_ENT_CAP * sizeof(unsigned))); FS_TRY(cudaMalloc(&d.resP, (size_t)8 * FS3_ENT_CAP * sizeof(unsigned long long)));
        FS_TRY(cudaMalloc(&d.resAft, (size_t)8 * FS3_ENT_CAP * sizeof(double)));
        FS_TRY(cudaMalloc(&d.tileEnd, FS3_MAX_TILES * sizeof(double)));
        FS_TRY(cudaMalloc(&d.rowlist, 1024 * sizeof(unsigned short))); FS_TRY(cudaMalloc(&d.rowinfo, 2 * sizeof(int)));
        FS_TRY(cudaMemset(d.rowinfo, 0, 2 * sizeof(int)));
        FS_TRY(cudaMalloc(&d.tileBw, FS3_MAX_TILES * sizeof(double))); FS_TRY(cudaMalloc(&d.tileBi, FS3_MAX_TILES * sizeof(unsigned)));
        FS_TRY(cudaMalloc(&d.flagsg, 8 * sizeof(int))); FS_TRY(cudaMemset(d.flagsg, 0, 8 * sizeof(int)));
        FS_TRY(cudaHostAlloc(&h->h_rec, sizeof(Fs3Rec), cudaHostAllocMapped));
        memset(h->h_rec, 0, sizeof(Fs3Rec));
        FS_TRY(cudaHostGetDevicePointer((void**)&d.rec, h->h_rec, 0));
        if (getenv("PFGPU_POST_TRACE")) { FS_TRY(cudaMalloc(&d.trace, 48 * sizeof(unsigned long long))); FS_TRY(cudaMemset(d.trace, 0, 48 * sizeof(unsigned long long))); }
    }
    { const char* e5 = getenv("PFGPU_PDL"); h->pdl = !(e5 && e5[0] == '0'); }
    { const char* e7 = getenv("PFGPU_EARLY_LAUNCH"); if (e7 && atoi(e7) == 1) h->early = true; }
    { const char* e6 = getenv("PFGPU_EKF_HELPERS"); if (e6 && atoi(e6) >= 1 && atoi(e6) <= 3) h->ekf_helpers = atoi(e6); }
    if (world > 1 && uid) {
        ncclUniqueId id;
        memcpy(&id, uid, sizeof(id));
        ncclResult_t nr = ncclCommInitRank(&h->comm, world, id, rank);
        if (nr != ncclSuccess) { snprintf(g_pfgpu_err, sizeof(g_pfgpu_err), "ncclCommInitRank: %s", ncclGetErrorString(nr)); return fail(PFGPU_ERR_NCCL); }
        // map every peer's arena (the 64-byte IPC handles travel over the communicator); all ranks must agree on the outcome
        int ok = 1;
        cudaIpcMemHandle_t mine, all[FS3_MAXG];
        char* d_hand = nullptr;
        FS_TRY(cudaMalloc(&d_hand, (size_t)(world + 1) * sizeof(cudaIpcMemHandle_t)));
        if (cudaIpcGetMemHandle(&mine, h->arena) != cudaSuccess) { ok = 0; memset(&mine, 0, sizeof(mine)); cudaGetLastError(); }
        FS_TRY(cudaMemcpy(d_hand + (size_t)world * sizeof(mine), &mine, sizeof(mine), cudaMemcpyHostToDevice));
        PF_NCCL(ncclAllGather(d_hand + (size_t)world * sizeof(mine), d_hand, sizeof(mine), ncclChar, h->comm, h->ctx.stream));
        FS_TRY(cudaStreamSynchronize(h->ctx.stream));
        FS_TRY(cudaMemcpy(all, d_hand, (size_t)world * sizeof(mine), cudaMemcpyDeviceToHost));
        for (int g = 0; g < world && ok; ++g) {
            if (g == rank) continue;
            if (cudaIpcOpenMemHandle(&h->peer_ptr[g], all[g], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; h->peer_ptr[g] = nullptr; cudaGetLastError(); }
        }
        int* d_ok = reinterpret_cast<int*>(d_hand);
        FS_TRY(cudaMemcpy(d_ok + world, &ok, sizeof(int), cudaMemcpyHostToDevice));
        PF_NCCL(ncclAllGather(d_ok + world, d_ok, 1, ncclInt, h->comm, h->ctx.stream));     // also the "everybody has zeroed and mapped" fence
        FS_TRY(cudaStreamSynchronize(h->ctx.stream));
        int oks[FS3_MAXG];
        FS_TRY(cudaMemcpy(oks, d_ok, (size_t)world * sizeof(int), cudaMemcpyDeviceToHost));
        cudaFree(d_hand);
        for (int g = 0; g < world; ++g) ok = ok && oks[g];
        if (!ok) { snprintf(g_pfgpu_err, sizeof(g_pfgpu_err), "sharded FastSLAM needs peer access between all GPUs (cudaIpc mapping failed)"); return fail(PFGPU_ERR_UNSUPPORTED); }
        for (int g = 0; g < world; ++g) d.peer[g] = (char*)h->peer_ptr[g];
    }
#undef FS_TRY
    fs3_init_kernel<<<cdiv_u(n, 256), 256, 0, h->ctx.stream>>>(d, cfg->init_weight);
    h->ctx.launches += 1;
    if (cudaStreamSynchronize(h->ctx.stream) != cudaSuccess) return fail(PFGPU_ERR_CUDA);
    *out = h;
    return PFGPU_OK;
}
extern "C" int pfgpu_fs_create(const pfgpu_fs_config* cfg, size_t n, size_t m, uint64_t seed, int device, pfgpu_fs** out) {
    return fs_create_impl(cfg, n, n, 0, m, seed, device, nullptr, 0, 1, out);
}
extern "C" int pfgpu_fs_create_sharded(const pfgpu_fs_config* cfg, size_t n_global, size_t m, uint64_t seed, int device,
                                       const void* uid, int rank, int world, pfgpu_fs** out) {
    if (out) *out = nullptr;
    if (!uid || world < 1 || world > FS3_MAXG || rank < 0 || rank >= world || n_global == 0 || n_global % (size_t)world != 0)
        return PFGPU_ERR_INVALID;
    size_t nl = n_global / (size_t)world;
    return fs_create_impl(cfg, nl, n_global, (size_t)rank * nl, m, seed, device, uid, rank, world, out);
}
// All ranks in ONE process (one host thread drives them, no NCCL): peers are plain device pointers (same device) or
// peer-access pointers (different devices).
extern "C" int pfgpu_fs_create_sharded_local(const pfgpu_fs_config* cfg, size_t n_global, size_t m, uint64_t seed, const int* devices,
                                             int world, pfgpu_fs** out) {
    if (!out || !devices || world < 1 || world > FS3_MAXG || n_global == 0 || n_global % (size_t)world != 0) return PFGPU_ERR_INVALID;
    for (int r = 0; r < world; ++r) out[r] = nullptr;
    const size_t nl = n_global / (size_t)world;
    static const char dummy_uid = 0;
    for (int r = 0; r < world; ++r) {
        int rc = fs_create_impl(cfg, nl, n_global, (size_t)r * nl, m, seed, devices[r], world > 1 ? nullptr : &dummy_uid, r, world, &out[r]);
        if (rc) { for (int q = 0; q < r; ++q) { pfgpu_fs_destroy(out[q]); out[q] = nullptr; } return rc; }
    }
    for (int a = 0; a < world; ++a)
        for (int b = 0; b < world; ++b) {
            if (devices[a] != devices[b]) {
                cudaSetDevice(devices[a]);
                cudaError_t e = cudaDeviceEnablePeerAccess(devices[b], 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
                    snprintf(g_pfgpu_err, sizeof(g_pfgpu_err), "no peer access from device %d to device %d", devices[a], devices[b]);
                    for (int q = 0; q < world; ++q) { pfgpu_fs_destroy(out[q]); out[q] = nullptr; }
                    return PFGPU_ERR_UNSUPPORTED;
                }
                cudaGetLastError();
            }
            out[a]->d.peer[b] = out[b]->arena;
            if (a != b && devices[a] == devices[b]) out[a]->d.wait_inline = 0;     // ranks sharing a GPU must not hold SMs while they wait
        }
    return PFGPU_OK;
}
extern "C" void pfgpu_fs_destroy(pfgpu_fs* h) {
    if (!h) return;
    cudaSetDevice(h->ctx.device);
    if (h->ctx.stream) cudaStreamSynchronize(h->ctx.stream);
    Fs3Dev& d = h->d;
    for (int g = 0; g < h->world; ++g) if (g != h->rank && h->peer_ptr[g]) cudaIpcCloseMemHandle(h->peer_ptr[g]);   // (local mode: none were opened)
    cudaFree(h->arena);
    cudaFree(d.st); cudaFree(d.lmst); cudaFree(d.w); cudaFree(d.nz[0]); cudaFree(d.nz[1]); cudaFree(d.wn_all); cudaFree(d.cum_all); cudaFree(d.rcomb_all); cudaFree(d.idx);
    cudaFree(d.tileP); cudaFree(d.tileQ); cudaFree(d.entCnt); cudaFree(d.entKey); cudaFree(d.entTile); cudaFree(d.entP); cudaFree(d.entV); cudaFree(d.entL);
    cudaFree(d.bar); cudaFree(d.rowlist); cudaFree(d.rowinfo); cudaFree(d.resflag); cudaFree(d.res); cudaFree(d.resTP); cudaFree(d.resKey);
    cudaFree(d.resP); cudaFree(d.resAft); cudaFree(d.tileEnd);
    cudaFree(d.tileBw); cudaFree(d.tileBi); cudaFree(d.flagsg); cudaFree(d.trace); cudaFree(h->stage);
    if (h->h_rec) cudaFreeHost(h->h_rec);
    if (h->comm) ncclCommDestroy(h->comm);
    marks_free(h->marks);
    for (auto& p : h->timer.pending) { cudaEventDestroy(p.first); cudaEventDestroy(p.second); }
    if (h->ctx.stream) cudaStreamDestroy(h->ctx.stream);
    delete h;
}
static int fs_check_err(pfgpu_fs* h) {        // device-side time-outs are sticky and surface at the next synchronising call
    if (h->h_rec && h->h_rec->err) {
        snprintf(g_pfgpu_err, sizeof(g_pfgpu_err), "FastSLAM step: a barrier or a peer GPU timed out (sticky; destroy the handle)");
        return PFGPU_ERR_CUDA;
    }
    return 0;
}
extern "C" int pfgpu_fs_sync(pfgpu_fs* h) {
    if (!h) return PFGPU_ERR_INVALID;
    PF_CUDA(cudaSetDevice(h->ctx.device));
    PF_CUDA(cudaStreamSynchronize(h->ctx.stream));
    return fs_check_err(h);
}
extern "C" int pfgpu_fs_count(pfgpu_fs* h, size_t* nl, size_t* ng, size_t* m) {
    if (!h) return PFGPU_ERR_INVALID;
    if (nl) *nl = h->d.n;
    if (ng) *ng = h->d.n_glob;
    if (m) *m = h->d.m;
    return 0;
}
static const size_t FS_XFER_CHUNK_BYTES = (size_t)256 << 20;   // staging chunk for AoS<->SoA conversion
extern "C" int pfgpu_fs_upload(pfgpu_fs* h, const double* pose_w, const double* lm, size_t n) {
    if (!h || !pose_w || n != h->d.n) return PFGPU_ERR_INVALID;
    PF_CUDA(cudaSetDevice(h->ctx.device));
    Fs3Dev& d = h->d;
    int rc = fs_stage(h, n * 4 * sizeof(double));
    if (rc) return rc;
    PF_CUDA(cudaMemcpyAsync(h->stage, pose_w, n * 4 * sizeof(double), cudaMemcpyHostToDevice, h->ctx.stream));
    PF_LAUNCH(h->ctx, fs3_unpack_pose_kernel, cdiv_u(n, 256), 256, 0, d, h->stage);
    PF_CUDA(cudaStreamSynchronize(h->ctx.stream));
    if (lm && d.m) {
        const size_t per = (size_t)d.m * 6 * sizeof(double);
        size_t chunk = std::max<size_t>(1, FS_XFER_CHUNK_BYTES / per);
        if (chunk > n) chunk = n;
        rc = fs_stage(h, chunk * per);
        if (rc) return rc;
        for (size_t i0 = 0; i0 < n; i0 += chunk) {
            const size_t cnt = std::min(chunk, n - i0);
            PF_CUDA(cudaMemcpyAsync(h->stage, lm + i0 * d.m * 6, cnt * per, cudaMemcpyHostToDevice, h->ctx.stream));
            PF_LAUNCH(h->ctx, fs3_unpack_lm_kernel, cdiv_u(cnt * d.m * 6, 256), 256, 0, d, h->stage, i0, cnt);
            PF_CUDA(cudaStreamSynchronize(h->ctx.stream));
        }
        PF_LAUNCH(h->ctx, fs3_lmst_reset_kernel, 1, 256, 0, d);
    }
    return 0;
}
extern "C" int pfgpu_fs_download(pfgpu_fs* h, double* pose_w, double* lm, size_t n) {
    if (!h || n != h->d.n) return PFGPU_ERR_INVALID;
    PF_CUDA(cudaSetDevice(h->ctx.device));
    Fs3Dev& d = h->d;
    if (pose_w) {
        int rc = fs_stage(h, n * 4 * sizeof(double));
        if (rc) return rc;
        PF_LAUNCH(h->ctx, fs3_pack_pose_kernel, cdiv_u(n, 256), 256, 0, d, h->stage);
        PF_CUDA(cudaMemcpyAsync(pose_w, h->stage, n * 4 * sizeof(double), cudaMemcpyDeviceToHost, h->ctx.stream));
        PF_CUDA(cudaStreamSynchronize(h->ctx.stream));
    }
    if (lm && d.m) {
        const size_t per = (size_t)d.m * 6 * sizeof(double);
        size_t chunk = std::max<size_t>(1, FS_XFER_CHUNK_BYTES / per);
        if (chunk > n) chunk = n;
        int rc = fs_stage(h, chunk * per);
        if (rc) return rc;
        for (size_t i0 = 0; i0 < n; i0 += chunk) {
            const size_t cnt = std::min(chunk, n - i0);
            PF_LAUNCH(h->ctx, fs3_pack_lm_kernel, cdiv_u(cnt * d.m * 6, 256), 256, 0, d, h->stage, i0, cnt);
            PF_CUDA(cudaMemcpyAsync(lm + i0 * d.m * 6, h->stage, cnt * per, cudaMemcpyDeviceToHost, h->ctx.stream));
            PF_CUDA(cudaStreamSynchronize(h->ctx.stream));
        }
    }
    return fs_check_err(h);
}

extern "C" int pfgpu_fs_seed_map(pfgpu_fs* h, const double pose3[3], const double* lm_xy, size_t m, double sigma, double cov0) {
    if (!h || !pose3 || (m && !lm_xy) || m != h->d.m) return PFGPU_ERR_INVALID;
    PF_CUDA(cudaSetDevice(h->ctx.device));
    Fs3Dev& d = h->d;
    PF_LAUNCH(h->ctx, fs3_seed_pose_kernel, cdiv_u(d.n, 256), 256, 0, d, pose3[0], pose3[1], pose3[2]);
    if (m) {
        int rc = fs_stage(h, m * 2 * sizeof(double));
        if (rc) return rc;
        PF_CUDA(cudaMemcpyAsync(h->stage, lm_xy, m * 2 * sizeof(double), cudaMemcpyHostToDevice, h->ctx.stream));
        dim3 grid(cdiv_u(d.n, 256), (unsigned)m);
        PF_LAUNCH(h->ctx, fs3_seed_lm_kernel, grid, 256, 0, d, h->stage, sigma, cov0, h->seed);
        PF_LAUNCH(h->ctx, fs3_lmst_reset_kernel, 1, 256, 0, d);
        PF_CUDA(cudaStreamSynchronize(h->ctx.stream));
    }
    return 0;
}

template <int MAXT>
static int fs3_launch_ekf(pfgpu_fs* h, const Fs3ObsParam& po, const double u[2], int kk, int flags) {
    const Fs3Dev& d = h->d;
    // helper warps (predict ahead / weights behind): as many as fit the CTA beside the kk EKF warps, at most 3
    int nh = kk > 0 ? std::min(3, MAXT / 32 - kk) : 1;
    if (h->ekf_helpers > 0 && kk > 0) nh = std::min(nh, h->ekf_helpers);
    const unsigned threads = 32u * (unsigned)(kk + nh);
    auto smem_of = [](int k, int n) { return ((size_t)n * 192 + (size_t)n * k * 64 + (size_t)2 * k * 384) * sizeof(double); };
    const size_t smem = smem_of(kk, nh);
    const unsigned groups = d.ld / 64;
    // persistent: one CTA per SM walks the 64-particle groups; without observations the launch is predict-only and latency
    // bound, so every group gets its own (one-warp) CTA
    const unsigned grid = kk > 0 ? std::min<unsigned>(groups, (unsigned)h->ctx.num_sms) : groups;
    if (!h->ekf_attr[MAXT == 512 ? 0 : 1]) {
        size_t mx = 0;
        for (int k = 0; k < MAXT / 32; ++k) mx = std::max(mx, smem_of(k, std::min(3, MAXT / 32 - k)));
        PF_CUDA(cudaFuncSetAttribute(fs3_ekf_kernel<MAXT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mx));
        h->ekf_attr[MAXT == 512 ? 0 : 1] = true;
    }
    PF_LAUNCH_PDL(h->ctx, h->pdl, (fs3_ekf_kernel<MAXT>), grid, threads, smem, d, po, u[0], u[1], h->cfg.dt, sqrt(h->cfg.q00), sqrt(h->cfg.q11),
                  h->cfg.r00, h->cfg.r11, h->seed, (uint32_t)h->n_step, kk, flags, (unsigned)h->n_step);
    return 0;
}

extern "C" int pfgpu_fs_step(pfgpu_fs* h, const double u[2], const pfgpu_fs_obs* z, size_t k, int* did) {
    if (!h || !u || (k && !z)) return PFGPU_ERR_INVALID;
    if (!finite_d(u[0]) || !finite_d(u[1])) return PFGPU_ERR_INVALID;
    Fs3Dev& d = h->d;
    for (size_t j = 0; j < k; ++j) {
        if (!finite_d(z[j].d) || !finite_d(z[j].angle)) return PFGPU_ERR_INVALID;
        if (z[j].lm_id >= d.m) return PFGPU_ERR_INVALID;             // the reference would panic on the Vec index (fs1.rs:141)
    }
    PF_CUDA(cudaSetDevice(h->ctx.device));
    // One EKF launch runs one warp per observation and the lazy-clone bookkeeping is per launch, so a launch must not see the
    // same lm_id twice and holds at most FS3_MAX_OBS observations: the list is cut before every repeated id / every
    // FS3_MAX_OBS entries and the pieces run as consecutive launches (same per-particle order as fs1.rs:250-256).
    std::vector<size_t> cuts;
    cuts.push_back(0);
    {
        std::vector<uint64_t> seen;
        for (size_t j = 0; j < k; ++j) {
            bool dup = seen.size() >= FS3_MAX_OBS;
            for (uint64_t v : seen) if (v == z[j].lm_id) { dup = true; break; }
            if (dup) { cuts.push_back(j); seen.clear(); }
            seen.push_back(z[j].lm_id);
        }
    }
    cuts.push_back(k);
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (h->timer.on) { PF_CUDA(cudaEventCreate(&e0)); PF_CUDA(cudaEventCreate(&e1)); PF_CUDA(cudaEventRecord(e0, h->ctx.stream)); }
    const size_t nseg = cuts.size() - 1;
    Fs3ObsParam po;
    int k_last = 0;
    const bool host_waits = d.G > 1 && !d.wait_inline;
    if (host_waits) PF_LAUNCH(h->ctx, fs3_wait_kernel, 1, 32, 0, d, 1, (unsigned)h->n_step);            // peers' previous post kernels are over
    const bool proposed = h->variant == 2 && k > 0;
    if (proposed) {        // FastSLAM 2.0: sample every pose from the proposal of the first observation (fs2.rs:341-346)
        Fs3Obs ob0; ob0.d = z[0].d; ob0.angle = z[0].angle; ob0.lm_id = (int)z[0].lm_id;
        PF_LAUNCH_PDL(h->ctx, h->pdl, fs2_propose_kernel, cdiv_u(d.n, 128), 128, 0, d, ob0, u[0], u[1], h->cfg.dt, h->cfg.r00, h->cfg.r11,
                      h->seed, (uint32_t)h->n_step, (unsigned)h->n_step);
    }
    for (size_t seg = 0; seg < nseg; ++seg) {
        const size_t j0 = cuts[seg], kk = cuts[seg + 1] - cuts[seg];
        // (ranks that share a GPU take turns on its SMs through the wait / signal launches: parked early CTAs of one rank could keep
        // another rank's persistent CTAs from ever being scheduled, so nothing is released early there)
        const int flags = (seg == 0 ? 1 : 0) | (proposed ? 2 : 0) | (h->variant == 2 ? 4 : 0) | (h->early && h->pdl && !host_waits ? 8 : 0);
        memset(&po, 0, sizeof(po));
        for (size_t j = 0; j < kk; ++j) { po.o[j].d = z[j0 + j].d; po.o[j].angle = z[j0 + j].angle; po.o[j].lm_id = (int)z[j0 + j].lm_id; }
        int rc = kk <= 15 ? fs3_launch_ekf<512>(h, po, u, (int)kk, flags) : fs3_launch_ekf<1024>(h, po, u, (int)kk, flags);
        if (rc) return rc;
        if (seg + 1 < nseg) PF_LAUNCH_PDL(h->ctx, h->pdl, fs3_mark_kernel, 1, 64, 0, d, po, (int)kk);    // the last piece: the post kernel does it
        k_last = (int)kk;
    }
    if (h->timer.on) { PF_CUDA(cudaEventRecord(e1, h->ctx.stream)); h->timer.pending.push_back({e0, e1}); }
    if (host_waits) {      // the post kernel signals "my weights are pushed" itself; the wait for the others' is its own launch here
        PF_LAUNCH(h->ctx, fs3_signal_kernel, 1, 32, 0, d, 0, (unsigned)h->n_step + 1u);
        PF_LAUNCH(h->ctx, fs3_wait_kernel, 1, 32, 0, d, 0, (unsigned)h->n_step + 1u);
    }
    // normalise, N_eff gate and (when it opens) the whole resample: one launch
    if (h->post_nt == 512)
        PF_LAUNCH_PDL(h->ctx, h->pdl, fs3_post_kernel<512>, h->post_tiles, 512, h->post_smem, d, po, k_last, h->cfg.nth, h->seed, (unsigned)h->n_step, h->post_K, h->m32, h->log2n, h->early && h->pdl && !host_waits ? 1 : 0);
    else
        PF_LAUNCH_PDL(h->ctx, h->pdl, fs3_post_kernel<256>, h->post_tiles, 256, h->post_smem, d, po, k_last, h->cfg.nth, h->seed, (unsigned)h->n_step, h->post_K, h->m32, h->log2n, h->early && h->pdl && !host_waits ? 1 : 0);
    h->n_step++;
    h->steps++;
    if (did) {     // the gate lives on the device; only a caller who asks pays a sync
        PF_CUDA(cudaStreamSynchronize(h->ctx.stream));
        *did = h->h_rec->gate;
        return fs_check_err(h);
    }
    return 0;
}
extern "C" int pfgpu_fs_set_variant(pfgpu_fs* h, int variant) {
    if (!h || (variant != 1 && variant != 2)) return PFGPU_ERR_INVALID;
    h->variant = variant;
    return 0;
}
extern "C" int pfgpu_fs_best(pfgpu_fs* h, size_t* index, double pose_w4[4]) {
    if (!h) return PFGPU_ERR_INVALID;
    PF_CUDA(cudaSetDevice(h->ctx.device));
    PF_CUDA(cudaStreamSynchronize(h->ctx.stream));
    const Fs3Rec* r = h->h_rec;
    if (h->steps == 0) {        // no step yet: every weight equals init_weight, the last particle wins (fs1.rs:269-274)
        if (index) *index = (size_t)h->d.n_glob - 1;
        if (pose_w4) { pose_w4[0] = h->cfg.init_weight; pose_w4[1] = pose_w4[2] = pose_w4[3] = 0.0; }
        return 0;
    }
    if (index) *index = (size_t)r->best_idx;
    if (pose_w4) { pose_w4[0] = r->best_w; pose_w4[1] = r->bx; pose_w4[2] = r->by; pose_w4[3] = r->byaw; }
    return fs_check_err(h);
}
extern "C" int pfgpu_fs_particle_landmarks(pfgpu_fs* h, size_t il, double* lm6) {
    if (!h || !lm6 || il >= h->d.n) return PFGPU_ERR_INVALID;
    PF_CUDA(cudaSetDevice(h->ctx.device));
    Fs3Dev& d = h->d;
    if (!d.m) return 0;
    int rc = fs_stage(h, (size_t)d.m * 6 * sizeof(double));
    if (rc) return rc;
    PF_LAUNCH(h->ctx, fs3_pack_lm_kernel, cdiv_u((size_t)d.m * 6, 256), 256, 0, d, h->stage, il, (size_t)1);
    PF_CUDA(cudaMemcpyAsync(lm6, h->stage, (size_t)d.m * 6 * sizeof(double), cudaMemcpyDeviceToHost, h->ctx.stream));
    PF_CUDA(cudaStreamSynchronize(h->ctx.stream));
    return 0;
}
extern "C" int pfgpu_fs_last_indices(pfgpu_fs* h, uint32_t* idx, size_t cap, size_t* n) {
    if (!h || !idx) return PFGPU_ERR_INVALID;
    PF_CUDA(cudaSetDevice(h->ctx.device));
    PF_CUDA(cudaStreamSynchronize(h->ctx.stream));
    if (!h->h_rec->gate) { if (n) *n = 0; return 0; }         // the last step did not resample: no ancestry (as the oracle reports)
    size_t c = cap < h->d.n ? cap : h->d.n;
    PF_CUDA(cudaMemcpyAsync(idx, h->d.idx, c * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->ctx.stream));
    PF_CUDA(cudaStreamSynchronize(h->ctx.stream));
    if (n) *n = c;
    return 0;
}
extern "C" int pfgpu_fs_last_neff(pfgpu_fs* h, double* neff) {
    if (!h || !neff) return PFGPU_ERR_INVALID;
    PF_CUDA(cudaSetDevice(h->ctx.device));
    PF_CUDA(cudaStreamSynchronize(h->ctx.stream));
    *neff = h->h_rec->neff;
    return 0;
}
// get_observations fs1.rs:277-299 on the device: out (capacity n_landmarks) receives the observed (d, angle, lm_id) tuples
extern "C" int pfgpu_fs_get_observations(pfgpu_fs* h, const double x_true[3], const double* landmarks_xy, size_t n_landmarks, uint32_t call,
                                         pfgpu_fs_obs* out, size_t* k) {
    if (!h || !x_true || (n_landmarks && (!landmarks_xy || !out)) || !k) return PFGPU_ERR_INVALID;
    PF_CUDA(cudaSetDevice(h->ctx.device));
    *k = 0;
    if (!n_landmarks) return 0;
    const size_t bytes_xy = n_landmarks * 2 * sizeof(double), bytes_out = n_landmarks * sizeof(Fs3Obs);
    int rc = fs_stage(h, bytes_xy + bytes_out + 256);
    if (rc) return rc;
    char* base = reinterpret_cast<char*>(h->stage);
    double* dxy = reinterpret_cast<double*>(base);
    Fs3Obs* dob = reinterpret_cast<Fs3Obs*>(base + ((bytes_xy + 15) & ~(size_t)15));
    unsigned* dk = reinterpret_cast<unsigned*>(base + ((bytes_xy + 15) & ~(size_t)15) + bytes_out);
    PF_CUDA(cudaMemcpyAsync(dxy, landmarks_xy, bytes_xy, cudaMemcpyHostToDevice, h->ctx.stream));
    PF_LAUNCH(h->ctx, fs3_get_observations_kernel, 1, 1024, 0, x_true[0], x_true[1], x_true[2], dxy, (unsigned)n_landmarks, h->cfg.max_range,
              sqrt(h->cfg.r00), sqrt(h->cfg.r11), h->seed, call, dob, dk);
    std::vector<Fs3Obs> tmp(n_landmarks);
    unsigned kk = 0;
    PF_CUDA(cudaMemcpyAsync(&kk, dk, sizeof(unsigned), cudaMemcpyDeviceToHost, h->ctx.stream));
    PF_CUDA(cudaMemcpyAsync(tmp.data(), dob, bytes_out, cudaMemcpyDeviceToHost, h->ctx.stream));
    PF_CUDA(cudaStreamSynchronize(h->ctx.stream));
    for (unsigned j = 0; j < kk; ++j) { out[j].d = tmp[j].d; out[j].angle = tmp[j].angle; out[j].lm_id = (uint64_t)tmp[j].lm_id; }
    *k = kk;
    return 0;
}
extern "C" int pfgpu_fs_last_gate(pfgpu_fs* h, int* did) {
    if (!h || !did) return PFGPU_ERR_INVALID;
    PF_CUDA(cudaSetDevice(h->ctx.device));
    PF_CUDA(cudaStreamSynchronize(h->ctx.stream));
    *did = h->steps ? h->h_rec->gate : 0;
    return fs_check_err(h);
}
extern "C" int pfgpu_fs_stats(pfgpu_fs* h, pfgpu_stats* s) {
    if (!h || !s) return PFGPU_ERR_INVALID;
    PF_CUDA(cudaSetDevice(h->ctx.device));
    PF_CUDA(cudaStreamSynchronize(h->ctx.stream));
    memset(s, 0, sizeof(*s));
    timer_drain(h->timer);
    Fs3State st;
    PF_CUDA(cudaMemcpy(&st, h->d.st, sizeof(st), cudaMemcpyDeviceToHost));
    s->kernel_launches = h->ctx.launches; s->steps = h->steps; s->resamples = st.resamples;
    s->main_kernel_ms_sum = h->timer.ms_sum; s->main_kernel_count = h->timer.count;
    s->serial_fallbacks = (uint64_t)st.serial_walks + (uint64_t)st.cert_fail;
    s->xsum_dirty_last = (uint64_t)st.dirty_last;
    return 0;
}
extern "C" int pfgpu_fs_shard_mode(pfgpu_fs* h, int* mode) {
    if (!h || !mode) return PFGPU_ERR_INVALID;
    *mode = h->world <= 1 ? 0 : 2;
    return 0;
}
// debug: accumulated phase times of the post kernel (PFGPU_POST_TRACE=1)
extern "C" int pfgpu_fs_post_trace(pfgpu_fs* h, unsigned long long* out32) {
    if (!h || !out32) return PFGPU_ERR_INVALID;
    PF_CUDA(cudaSetDevice(h->ctx.device));
    PF_CUDA(cudaStreamSynchronize(h->ctx.stream));
    for (int k = 0; k < 32; ++k) out32[k] = 0;
    if (h->d.trace) PF_CUDA(cudaMemcpy(out32, h->d.trace, 32 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    out32[31] = h->steps;
    // step timeline [ns], folded into the free slot pairs: [7] idle before the EKF launch, [24..26] EKF launch, idle between the
    // launches, post launch
    if (h->d.trace) {
        unsigned long long t8[9] = {};
        PF_CUDA(cudaMemcpy(t8, h->d.trace + 32, 9 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
        out32[7] = t8[0]; out32[24] = t8[1]; out32[25] = t8[2]; out32[26] = t8[3]; out32[27] = t8[8];
    }
    return 0;
}
extern "C" int pfgpu_fs_time_main_kernel(pfgpu_fs* h, int on) {
    if (!h) return PFGPU_ERR_INVALID;
    PF_CUDA(cudaSetDevice(h->ctx.device));
    PF_CUDA(cudaStreamSynchronize(h->ctx.stream));
    timer_drain(h->timer);
    h->timer.on = on != 0; h->timer.ms_sum = 0.0; h->timer.count = 0;
    return 0;
}
