// pfgpu.cu — the C ABI of include/pfgpu.h: handles, step orchestration, upload/download.
// Single translation unit: nvcc -gencode arch=compute_100a,code=sm_100a --fmad=false (see __graft_entry__.build()).
#include "common.cuh"
#include "xsum.cuh"
#include "pf_kernels.cuh"
#include "fs_kernels.cuh"
#include "fs_post.cuh"
#include "pf_kld.cuh"
#include "fs_sharded.cuh"
#include "fs_mg.cuh"
#include <cstdlib>
#include <new>
#include <vector>
#include <cmath>
#include <algorithm>

thread_local char g_pfgpu_err[512] = {0};

extern "C" const char* pfgpu_last_error(void) { return g_pfgpu_err; }
extern "C" const char* pfgpu_strerror(int s) {
    switch (s) {
        case PFGPU_OK: return "ok";
        case PFGPU_ERR_INVALID: return "invalid parameter";
        case PFGPU_ERR_UNSUPPORTED: return "valid in the reference but not supported by this build";
        case PFGPU_ERR_NO_DEVICE: return "no usable CUDA device (this library has no CPU fallback)";
        case PFGPU_ERR_CUDA: return "CUDA runtime error (see pfgpu_last_error)";
        case PFGPU_ERR_NCCL: return "NCCL error (see pfgpu_last_error)";
        default: return "unknown status";
    }
}
extern "C" int pfgpu_device_count(int* count) {
    int c = 0;
    cudaError_t e = cudaGetDeviceCount(&c);
    if (e != cudaSuccess) { *count = 0; snprintf(g_pfgpu_err, sizeof(g_pfgpu_err), "cudaGetDeviceCount: %s", cudaGetErrorString(e)); return PFGPU_ERR_NO_DEVICE; }
    *count = c;
    return c > 0 ? PFGPU_OK : PFGPU_ERR_NO_DEVICE;
}

static bool finite_d(double v) { return std::isfinite(v); }

static int ctx_open(Ctx& ctx, int device) {
    int count = 0;
    int rc = pfgpu_device_count(&count);
    if (rc) return rc;
    if (device < 0 || device >= count) { snprintf(g_pfgpu_err, sizeof(g_pfgpu_err), "device %d out of range (%d devices)", device, count); return PFGPU_ERR_NO_DEVICE; }
    PF_CUDA(cudaSetDevice(device));
    ctx.device = device;
    PF_CUDA(cudaStreamCreateWithFlags(&ctx.stream, cudaStreamNonBlocking));
    int sms = 0;
    PF_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    ctx.num_sms = sms > 0 ? sms : PFGPU_NUM_SMS;
    return 0;
}

#define PF_MARK_SLOTS 16384
__global__ void pf_l2_read_kernel(const double4* __restrict__ p, size_t n4, double* sink) {
    double acc = 0.0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        double4 v = p[i];
        acc += v.x + v.y + v.z + v.w;
    }
    if (acc == 123.456) *sink = acc;      // never true: keeps the loads alive
}
struct Marks {
    std::vector<cudaEvent_t> ev = std::vector<cudaEvent_t>(PF_MARK_SLOTS, nullptr);
    void* l2buf = nullptr;
    void* l2buf_rd = nullptr;
    size_t l2bytes = (size_t)256 << 20;      // > 126 MB L2
};
static int marks_mark(Ctx& ctx, Marks& m, int slot) {
    if (slot < 0 || slot >= PF_MARK_SLOTS) return PFGPU_ERR_INVALID;
    if (!m.ev[slot]) PF_CUDA(cudaEventCreate(&m.ev[slot]));
    PF_CUDA(cudaEventRecord(m.ev[slot], ctx.stream));
    return 0;
}
static int marks_elapsed(Marks& m, int a, int b, double* ms) {
    if (a < 0 || a >= PF_MARK_SLOTS || b < 0 || b >= PF_MARK_SLOTS || !m.ev[a] || !m.ev[b] || !ms) return PFGPU_ERR_INVALID;
    PF_CUDA(cudaEventSynchronize(m.ev[b]));
    float f = 0.f;
    PF_CUDA(cudaEventElapsedTime(&f, m.ev[a], m.ev[b]));
    *ms = (double)f;
    return 0;
}
static int marks_flush(Ctx& ctx, Marks& m) {
    if (!m.l2buf) {
        PF_CUDA(cudaMalloc(&m.l2buf, m.l2bytes));
        PF_CUDA(cudaMalloc(&m.l2buf_rd, m.l2bytes));
        PF_CUDA(cudaMemsetAsync(m.l2buf_rd, 0, m.l2bytes, ctx.stream));
    }
    // write a buffer larger than L2 (evicts everything), then stream a second one through it so that the cache is left
    // full of CLEAN lines: the timed kernel then starts cold without inheriting the flush's own write-backs
    PF_CUDA(cudaMemsetAsync(m.l2buf, 0, m.l2bytes, ctx.stream));
    pf_l2_read_kernel<<<ctx.num_sms * 8, 256, 0, ctx.stream>>>((const double4*)m.l2buf_rd, m.l2bytes / 32, (double*)m.l2buf);
    return cudaGetLastError() == cudaSuccess ? 0 : PFGPU_ERR_CUDA;
}
static void marks_free(Marks& m) {
    for (auto e : m.ev) if (e) cudaEventDestroy(e);
    if (m.l2buf) cudaFree(m.l2buf);
    if (m.l2buf_rd) cudaFree(m.l2buf_rd);
}

struct KernelTimer {
    bool on = false;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending;
    double ms_sum = 0.0; uint64_t count = 0;
};

// ====================================================================================================
// ParticleFilterLocalizer / MonteCarloLocalizer
// ====================================================================================================
struct pfgpu_pf {
    Ctx ctx;
    pfgpu_pf_config cfg;
    uint64_t seed = 0;
    PfDev d;
    XsWork xs;
    size_t obs_cap = 0;
    double* mom15 = nullptr;
    int mom_blocks = 0;
    uint32_t n_predict = 0, n_resample = 0;
    uint64_t steps = 0, resamples_unknown = 0;
    int world = 1, rank = 0;
    KernelTimer timer;
    Marks marks;
    double* h_pin = nullptr;   // pinned scratch (>= 64 doubles)
    FsShard sh;                // multi-GPU state (world == 1: unused)
    int cur_host = 0;          // sharded mode: host mirror of *d.cur (the host knows every gate there)
    bool adaptive = false;     // MCL with min_particles < max_particles: the particle count changes per step (pf_kld.cuh)
    PfKld kld;
};

extern "C" void pfgpu_pf_default_config(pfgpu_pf_config* c, int mode) {
    memset(c, 0, sizeof(*c));
    c->n_particles = 100; c->resample_threshold = 0.5; c->range_noise = 0.2; c->velocity_noise = 2.0;
    c->yaw_rate_noise = 40.0 * PFC_PI / 180.0; c->dt = 0.1; c->mode = mode;
    c->max_particles = mode == 1 ? 5000 : 100; c->kld_epsilon = 0.05; c->kld_z = 2.326;
}
extern "C" int pfgpu_pf_config_validate(const pfgpu_pf_config* c) {
    if (!c) return PFGPU_ERR_INVALID;
    if (c->n_particles == 0) return PFGPU_ERR_INVALID;                                                   // pf.rs:82, mcl.rs:88
    if (c->mode == 0) {
        if (!finite_d(c->resample_threshold) || c->resample_threshold < 0.0 || c->resample_threshold > 1.0) return PFGPU_ERR_INVALID;  // pf.rs:87-94
    } else if (c->mode == 1) {
        if (c->max_particles < c->n_particles) return PFGPU_ERR_INVALID;                                 // mcl.rs:93-97
        if (!finite_d(c->kld_epsilon) || c->kld_epsilon <= 0.0) return PFGPU_ERR_INVALID;                // mcl.rs:98-102
        if (!finite_d(c->kld_z) || c->kld_z <= 0.0) return PFGPU_ERR_INVALID;                            // mcl.rs:103-107
    } else return PFGPU_ERR_INVALID;
    if (!finite_d(c->range_noise) || c->range_noise <= 0.0) return PFGPU_ERR_INVALID;                    // pf.rs:95-99
    if (!finite_d(c->velocity_noise) || c->velocity_noise < 0.0) return PFGPU_ERR_INVALID;               // pf.rs:100-104
    if (!finite_d(c->yaw_rate_noise) || c->yaw_rate_noise < 0.0) return PFGPU_ERR_INVALID;               // pf.rs:105-109
    if (!finite_d(c->dt) || c->dt <= 0.0) return PFGPU_ERR_INVALID;                                      // pf.rs:110-114
    return PFGPU_OK;
}

__global__ void pf_init_zero_kernel(PfDev d) {
    const size_t i = (size_t)blockIdx.x * PF_NT + threadIdx.x;
    if (i >= d.n) return;
    Pose4 z; z.x = 0.0; z.y = 0.0; z.yaw = 0.0; z.v = 0.0;
    pose_store(d.pose[0], i, z);
    d.w[i] = 1.0 / (double)d.n_global;                 // Particle::new pf.rs:35-43
    d.w_raw[i] = d.w[i];
}
// try_with_initial_state pf.rs:181-187 (random::<f64>()*2-1 ...) / mcl.rs:190-196 (random_range(-1.0..1.0) ...)
__global__ void pf_init_state_kernel(PfDev d, double s0, double s1, double s2, double s3, uint64_t seed, int mode) {
    const size_t i = (size_t)blockIdx.x * PF_NT + threadIdx.x;
    if (i >= d.n) return;
    pfc_u32x4 a = pfc_rng_block(seed, PFC_STREAM_INIT_A, 0, d.offset + i);
    pfc_u32x4 b = pfc_rng_block(seed, PFC_STREAM_INIT_B, 0, d.offset + i);
    Pose4 p;
    if (mode == 0) {
        p.x = s0 + pfc_u01_53(pfc_blk_u64(a, 0)) * 2.0 - 1.0;
        p.y = s1 + pfc_u01_53(pfc_blk_u64(a, 1)) * 2.0 - 1.0;
        p.yaw = s2 + pfc_u01_53(pfc_blk_u64(b, 0)) * 0.5 - 0.25;
        p.v = s3 + pfc_u01_53(pfc_blk_u64(b, 1)) * 1.0 - 0.5;
    } else {
        p.x = s0 + (pfc_u01_52(pfc_blk_u64(a, 0)) * 2.0 + -1.0);
        p.y = s1 + (pfc_u01_52(pfc_blk_u64(a, 1)) * 2.0 + -1.0);
        p.yaw = s2 + (pfc_u01_52(pfc_blk_u64(b, 0)) * 0.5 + -0.25);
        p.v = s3 + (pfc_u01_52(pfc_blk_u64(b, 1)) * 1.0 + -0.5);
    }
    pose_store(pf_pose(d, *d.cur), i, p);
    d.w[i] = 1.0 / (double)d.n_global;
    d.w_raw[i] = d.w[i];
}
__global__ void pf_unpack_kernel(PfDev d, const double* aos5) {
    const size_t i = (size_t)blockIdx.x * PF_NT + threadIdx.x;
    if (i >= d.n) return;
    Pose4 p; p.x = aos5[5 * i]; p.y = aos5[5 * i + 1]; p.yaw = aos5[5 * i + 2]; p.v = aos5[5 * i + 3];
    pose_store(pf_pose(d, *d.cur), i, p);
    d.w[i] = aos5[5 * i + 4];
    d.w_raw[i] = d.w[i];
}
__global__ void pf_pack_kernel(PfDev d, double* aos5) {
    const size_t i = (size_t)blockIdx.x * PF_NT + threadIdx.x;
    if (i >= d.n) return;
    Pose4 p;
    pose_load(pf_pose(d, *d.cur), i, p);
    aos5[5 * i] = p.x; aos5[5 * i + 1] = p.y; aos5[5 * i + 2] = p.yaw; aos5[5 * i + 3] = p.v; aos5[5 * i + 4] = d.w[i];
}

static int pf_refresh_cache(pfgpu_pf* h) {      // refresh_cache pf.rs:499-503
    PF_LAUNCH(h->ctx, pf_moments_kernel, h->mom_blocks, PF_NT, 0, h->d, h->mom_blocks);
    PF_LAUNCH(h->ctx, pf_moments_reduce_kernel, 1, PF_NT, 0, h->d.partial, h->mom_blocks, h->mom15);
    if (h->world > 1)   // shard moments about the common centre (the previous, replicated estimate) add up
        PF_NCCL(ncclAllReduce(h->mom15, h->mom15, PF_MOM, ncclDouble, ncclSum, h->sh.comm, h->ctx.stream));
    PF_LAUNCH(h->ctx, pf_moments_final_kernel, 1, 32, 0, h->d, h->mom15);
    return 0;
}

static int pf_alloc(pfgpu_pf* h, size_t cap) {
    PfDev& d = h->d;
    const size_t n = cap;       // every per-particle array is sized for the largest generation
    PF_CUDA(cudaMalloc(&d.pose[0], n * sizeof(Pose4)));
    PF_CUDA(cudaMalloc(&d.pose[1], n * sizeof(Pose4)));
    PF_CUDA(cudaMalloc(&d.cur, sizeof(int)));
    PF_CUDA(cudaMemset(d.cur, 0, sizeof(int)));
    PF_CUDA(cudaMalloc(&d.w_raw, n * sizeof(double)));
    PF_CUDA(cudaMalloc(&d.w, n * sizeof(double)));
    PF_CUDA(cudaMalloc(&d.cum, n * sizeof(double)));
    PF_CUDA(cudaMalloc(&d.idx, n * sizeof(uint32_t)));
    PF_CUDA(cudaMalloc(&d.scal, 32 * sizeof(double)));
    PF_CUDA(cudaMemset(d.scal, 0, 32 * sizeof(double)));
    PF_CUDA(cudaMalloc(&d.gate, sizeof(int)));
    PF_CUDA(cudaMemset(d.gate, 0, sizeof(int)));
    PF_CUDA(cudaMalloc(&d.counters, 4 * sizeof(unsigned int)));
    PF_CUDA(cudaMemset(d.counters, 0, 4 * sizeof(unsigned int)));
    h->mom_blocks = (int)std::min<size_t>((size_t)h->ctx.num_sms * 4, cdiv_u(n, PF_NT));
    if (h->mom_blocks < 1) h->mom_blocks = 1;
    PF_CUDA(cudaMalloc(&d.partial, (size_t)h->mom_blocks * PF_MOM * sizeof(double)));
    PF_CUDA(cudaMalloc(&h->mom15, PF_MOM * sizeof(double)));
    h->obs_cap = 1024;
    PF_CUDA(cudaMalloc(&d.obs, h->obs_cap * 3 * sizeof(double)));
    PF_CUDA(cudaMallocHost(&h->h_pin, 64 * sizeof(double)));
    if (h->adaptive) {
        PfKld& k = h->kld;
        k.cap = cap;
        unsigned tc = 64;
        while ((size_t)tc < 2 * cap + 16) tc <<= 1;
        k.tcap = tc;
        PF_CUDA(cudaMalloc(&k.keys, 3 * cap * sizeof(int)));
        PF_CUDA(cudaMalloc(&k.owner, (size_t)tc * sizeof(int)));
        PF_CUDA(cudaMalloc(&k.mint, (size_t)tc * sizeof(unsigned)));
        PF_CUDA(cudaMalloc(&k.slot, cap * sizeof(int)));
        PF_CUDA(cudaMalloc(&k.n_new, sizeof(unsigned)));
    }
    return xs_work_alloc(h->xs, n);
}

static int pf_create_impl(const pfgpu_pf_config* cfg, uint64_t seed, int device, const void* uid, int rank, int world, pfgpu_pf** out) {
    if (!out) return PFGPU_ERR_INVALID;
    *out = nullptr;
    int rc = pfgpu_pf_config_validate(cfg);
    if (rc) return rc;
    const bool adaptive = cfg->mode == 1 && cfg->max_particles != cfg->n_particles;     // KLD-adaptive particle count, mcl.rs:322-365
    if (adaptive && world > 1) return PFGPU_ERR_UNSUPPORTED;                            // a changing count is not sharded (yet)
    if (cfg->n_particles > 0xFFFFFFFFull || (adaptive && cfg->max_particles > 0x7FFFFFFFull)) return PFGPU_ERR_UNSUPPORTED;
    if (world > 1 && (cfg->n_particles % (uint64_t)world) != 0) return PFGPU_ERR_INVALID;
    pfgpu_pf* h = new (std::nothrow) pfgpu_pf();
    if (!h) return PFGPU_ERR_CUDA;
    rc = ctx_open(h->ctx, device);
    if (rc) { delete h; return rc; }
    h->cfg = *cfg; h->seed = seed; h->world = world; h->rank = rank;
    h->d.n_global = cfg->n_particles; h->d.n = cfg->n_particles / (uint64_t)world; h->d.offset = (size_t)rank * h->d.n;
    h->adaptive = adaptive;
    rc = pf_alloc(h, adaptive ? (size_t)cfg->max_particles : h->d.n);
    if (rc) { pfgpu_pf_destroy(h); return rc; }
    if (world > 1) {
        FsShard& sh = h->sh;
        sh.rank = rank; sh.world = world;
        ncclUniqueId id;
        memcpy(&id, uid, sizeof(id));
        ncclResult_t nr = ncclCommInitRank(&sh.comm, world, id, rank);
        if (nr != ncclSuccess) { snprintf(g_pfgpu_err, sizeof(g_pfgpu_err), "ncclCommInitRank: %s", ncclGetErrorString(nr)); pfgpu_pf_destroy(h); return PFGPU_ERR_NCCL; }
        const size_t ng = h->d.n_global;
        bool ok = cudaMalloc(&sh.t_loc, sizeof(double)) == cudaSuccess && cudaMalloc(&sh.t_all, world * sizeof(double)) == cudaSuccess &&
                  cudaMalloc(&sh.approx_off, sizeof(double)) == cudaSuccess && cudaMalloc(&sh.sum_loc, sizeof(ShardSummary)) == cudaSuccess &&
                  cudaMalloc(&sh.sum_all, world * sizeof(ShardSummary)) == cudaSuccess && cudaMalloc(&sh.s_start, sizeof(double)) == cudaSuccess &&
                  cudaMalloc(&sh.err, sizeof(int)) == cudaSuccess && cudaMemset(sh.err, 0, sizeof(int)) == cudaSuccess &&
                  cudaMalloc(&sh.cum_all, ng * sizeof(double)) == cudaSuccess && cudaMalloc(&sh.pose_all, 4 * ng * sizeof(double)) == cudaSuccess;
        if (!ok) { pfgpu_pf_destroy(h); return PFGPU_ERR_CUDA; }
    }
    PF_LAUNCH(h->ctx, pf_init_zero_kernel, cdiv_u(h->d.n, PF_NT), PF_NT, 0, h->d);
    rc = pf_refresh_cache(h);
    if (rc) { pfgpu_pf_destroy(h); return rc; }
    PF_CUDA(cudaStreamSynchronize(h->ctx.stream));
    *out = h;
    return PFGPU_OK;
}
extern "C" int pfgpu_pf_create(const pfgpu_pf_config* cfg, uint64_t seed, int device, pfgpu_pf** out) {
    return pf_create_impl(cfg, seed, device, nullptr, 0, 1, out);
}
extern "C" int pfgpu_pf_create_sharded(const pfgpu_pf_config* cfg, uint64_t seed, int device, const void* uid, int rank, int world, pfgpu_pf** out) {
    if (out) *out = nullptr;
    if (!uid || world < 1 || world > SH_MAX_WORLD || rank < 0 || rank >= world) return PFGPU_ERR_INVALID;
    return pf_create_impl(cfg, seed, device, uid, rank, world, out);
}
extern "C" void pfgpu_pf_destroy(pfgpu_pf* h) {
    if (!h) return;
    cudaSetDevice(h->ctx.device);
    if (h->ctx.stream) cudaStreamSynchronize(h->ctx.stream);
    PfDev& d = h->d;
    cudaFree(d.pose[0]); cudaFree(d.pose[1]); cudaFree(d.cur); cudaFree(d.w_raw); cudaFree(d.w); cudaFree(d.cum);
    cudaFree(d.idx); cudaFree(d.scal); cudaFree(d.gate); cudaFree(d.partial); cudaFree(d.obs); cudaFree(h->mom15); cudaFree(d.counters);
    if (h->h_pin) cudaFreeHost(h->h_pin);
    cudaFree(h->kld.keys); cudaFree(h->kld.owner); cudaFree(h->kld.mint); cudaFree(h->kld.slot); cudaFree(h->kld.n_new);
    {
        FsShard& sh = h->sh;
        cudaFree(sh.t_loc); cudaFree(sh.t_all); cudaFree(sh.approx_off); cudaFree(sh.sum_loc); cudaFree(sh.sum_all); cudaFree(sh.s_start);
        cudaFree(sh.err); cudaFree(sh.cum_all); cudaFree(sh.pose_all);
        if (sh.comm) ncclCommDestroy(sh.comm);
    }
    marks_free(h->marks);
    xs_work_free(h->xs);
    for (auto& p : h->timer.pending) { cudaEventDestroy(p.first); cudaEventDestroy(p.second); }
    if (h->ctx.stream) cudaStreamDestroy(h->ctx.stream);
    delete h;
}
extern "C" int pfgpu_pf_sync(pfgpu_pf* h) {
    if (!h) return PFGPU_ERR_INVALID;
    PF_CUDA(cudaSetDevice(h->ctx.device));
    PF_CUDA(cudaStreamSynchronize(h->ctx.stream));
    return 0;
}
extern "C" int pfgpu_pf_init_state(pfgpu_pf* h, const double s[4]) {
    if (!h || !s) return PFGPU_ERR_INVALID;
    for (int k = 0; k < 4; ++k) if (!finite_d(s[k])) return PFGPU_ERR_INVALID;     // validate_state pf.rs:505-513
    PF_CUDA(cudaSetDevice(h->ctx.device));
    PF_LAUNCH(h->ctx, pf_init_state_kernel, cdiv_u(h->d.n, PF_NT), PF_NT, 0, h->d, s[0], s[1], s[2], s[3], h->seed, h->cfg.mode);
    return pf_refresh_cache(h);
}
extern "C" int pfgpu_pf_upload(pfgpu_pf* h, const double* aos5, size_t n) {
    if (!h || !aos5 || n != h->d.n) return PFGPU_ERR_INVALID;
    PF_CUDA(cudaSetDevice(h->ctx.device));
    double* tmp = nullptr;
    PF_CUDA(cudaMalloc(&tmp, n * 5 * sizeof(double)));
    PF_CUDA(cudaMemcpyAsync(tmp, aos5, n * 5 * sizeof(double), cudaMemcpyHostToDevice, h->ctx.stream));
    PF_LAUNCH(h->ctx, pf_unpack_kernel, cdiv_u(n, PF_NT), PF_NT, 0, h->d, tmp);
    int rc = pf_refresh_cache(h);
    PF_CUDA(cudaStreamSynchronize(h->ctx.stream));
    cudaFree(tmp);
    return rc;
}
extern "C" int pfgpu_pf_download(pfgpu_pf* h, double* aos5, size_t n) {
    if (!h || !aos5 || n != h->d.n) return PFGPU_ERR_INVALID;
    PF_CUDA(cudaSetDevice(h->ctx.device));
    double* tmp = nullptr;
    PF_CUDA(cudaMalloc(&tmp, n * 5 * sizeof(double)));
    PF_LAUNCH(h->ctx, pf_pack_kernel, cdiv_u(n, PF_NT), PF_NT, 0, h->d, tmp);
    PF_CUDA(cudaMemcpyAsync(aos5, tmp, n * 5 * sizeof(double), cudaMemcpyDeviceToHost, h->ctx.stream));
    PF_CUDA(cudaStreamSynchronize(h->ctx.stream));
    cudaFree(tmp);
    return 0;
}
extern "C" int pfgpu_pf_count(pfgpu_pf* h, size_t* nl, size_t* ng) {
    if (!h) return PFGPU_ERR_INVALID;
    if (nl) *nl = h->d.n;
    if (ng) *ng = h->d.n_global;
    return 0;
}

static int pf_stage_obs(pfgpu_pf* h, const double* obs3, size_t k) {
    for (size_t j = 0; j < k; ++j)                                   // validate_observations pf.rs:538-549
        if (!finite_d(obs3[3 * j]) || !finite_d(obs3[3 * j + 1]) || !finite_d(obs3[3 * j + 2]) || obs3[3 * j] < 0.0)
            return PFGPU_ERR_INVALID;
    if (k <= PF_PARAM_OBS) return 0;                                 // short lists ride in the launch parameters
    if (k > h->obs_cap) {
        PF_CUDA(cudaStreamSynchronize(h->ctx.stream));
        cudaFree(h->d.obs);
        h->obs_cap = k * 2;
        PF_CUDA(cudaMalloc(&h->d.obs, h->obs_cap * 3 * sizeof(double)));
    }
    if (k > 0) PF_CUDA(cudaMemcpyAsync(h->d.obs, obs3, k * 3 * sizeof(double), cudaMemcpyHostToDevice, h->ctx.stream));
    return 0;
}
static size_t pf_obs_smem(size_t k) { return (k ? k : 1) * 3 * sizeof(double); }

template <bool P, bool W>
static int pf_launch_main(pfgpu_pf* h, const double u[2], const double* obs3, size_t k) {
    size_t smem = W ? pf_obs_smem(k) : 0;
    const bool param = !W || k <= PF_PARAM_OBS;
    PfObsParam po;
    if (W && param) for (size_t j = 0; j < 3 * k; ++j) po.o[j] = obs3[j];
    if (smem > 48 * 1024) {
        PF_CUDA(cudaFuncSetAttribute((pf_predict_weight_kernel<P, W, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (h->timer.on) { PF_CUDA(cudaEventCreate(&e0)); PF_CUDA(cudaEventCreate(&e1)); PF_CUDA(cudaEventRecord(e0, h->ctx.stream)); }
    if (param)
        PF_LAUNCH(h->ctx, (pf_predict_weight_kernel<P, W, true>), cdiv_u(h->d.n, PF_NT), PF_NT, smem, h->d, po, u ? u[0] : 0.0, u ? u[1] : 0.0,
                  h->cfg.velocity_noise, h->cfg.yaw_rate_noise, h->cfg.dt, h->seed, h->n_predict, (int)k, h->cfg.range_noise);
    else
        PF_LAUNCH(h->ctx, (pf_predict_weight_kernel<P, W, false>), cdiv_u(h->d.n, PF_NT), PF_NT, smem, h->d, po, u ? u[0] : 0.0, u ? u[1] : 0.0,
                  h->cfg.velocity_noise, h->cfg.yaw_rate_noise, h->cfg.dt, h->seed, h->n_predict, (int)k, h->cfg.range_noise);
    if (h->timer.on) { PF_CUDA(cudaEventRecord(e1, h->ctx.stream)); h->timer.pending.push_back({e0, e1}); }
    return 0;
}
// normalize_weights: exact sequential sum of the raw weights, then the division pass
// global index search + pose gather of the sharded mode (pose_all = ncclAllGather of the 32-byte records, rank order)
__global__ void __launch_bounds__(PF_NT) pf_search_sharded_kernel(PfDev d, const double* cum_all, uint64_t seed, int mode) {
    if (!*d.gate) return;
    const size_t t = (size_t)blockIdx.x * PF_NT + threadIdx.x;
    if (t >= d.n) return;
    double r = pfc_u01_53(pfc_blk_u64(pfc_rng_block(seed, PFC_STREAM_PF_RESAMPLE, d.counters[0], d.offset + t), 0));
    size_t lo = 0, hi = d.n_global;
    while (lo < hi) {
        size_t mid = lo + ((hi - lo) >> 1);
        if (cum_all[mid] < r) lo = mid + 1; else hi = mid;
    }
    d.idx[t] = (uint32_t)(lo < d.n_global ? lo : (mode == 1 ? d.n_global - 1 : 0));
}
__global__ void __launch_bounds__(PF_NT) pf_gather_sharded_kernel(PfDev d, const Pose4* pose_all) {
    if (!*d.gate) return;
    const size_t t = (size_t)blockIdx.x * PF_NT + threadIdx.x;
    if (t >= d.n) return;
    Pose4 p;
    pose_load(pose_all, d.idx[t], p);
    pose_store(pf_pose(d, *d.cur ^ 1), t, p);
    d.w[t] = 1.0 / (double)d.n_global;
}
template <class F>
static int pf_total(pfgpu_pf* h, F f, double* out) {
    if (h->world > 1) return xs_total_sharded(h->ctx, h->xs, h->sh, f, h->d.n, h->d.n_global, out);
    return xs_total(h->ctx, h->xs, f, h->d.n, h->d.n_global, 0.0, out);
}
static int pf_resample_sharded(pfgpu_pf* h) {
    PfDev& d = h->d; FsShard& sh = h->sh; Ctx& ctx = h->ctx;
    int rc = xs_total_sharded(ctx, h->xs, sh, PfValWSq{d.w}, d.n, d.n_global, d.scal + 1);          // calc_n_eff pf.rs:416-423
    if (rc) return rc;
    PF_LAUNCH(ctx, pf_gate_kernel, 1, 1, 0, d, h->cfg.resample_threshold, h->cfg.mode);
    int gate = 1;
    int* hp = reinterpret_cast<int*>(h->h_pin + 32);
    // the collectives below are host-enqueued: every rank must know the gate (MCL: always open, mcl.rs:298)
    PF_CUDA(cudaMemcpyAsync(hp, d.gate, sizeof(int), cudaMemcpyDeviceToHost, ctx.stream));
    PF_CUDA(cudaMemcpyAsync(hp + 1, sh.err, sizeof(int), cudaMemcpyDeviceToHost, ctx.stream));
    PF_CUDA(cudaStreamSynchronize(ctx.stream));
    if (hp[1]) { snprintf(g_pfgpu_err, sizeof(g_pfgpu_err), "sharded exact sum: a shard was not summarisable (degenerate weights)"); return PFGPU_ERR_UNSUPPORTED; }
    gate = *hp;
    if (!gate) return 0;
    rc = xs_scan_sharded(ctx, h->xs, sh, XsValArray{d.w}, XsSinkStore{d.cum}, d.n, d.n_global, d.scal + 2);   // pf.rs:448-453
    if (rc) return rc;
    if (h->cfg.mode == 1 && sh.rank == sh.world - 1) PF_LAUNCH(ctx, pf_force_last_kernel, 1, 1, 0, d);        // mcl.rs:334-336
    PF_NCCL(ncclAllGather(d.cum, sh.cum_all, d.n, ncclDouble, sh.comm, ctx.stream));
    PF_LAUNCH(ctx, pf_search_sharded_kernel, cdiv_u(d.n, PF_NT), PF_NT, 0, d, sh.cum_all, h->seed, h->cfg.mode);
    PF_NCCL(ncclAllGather(d.pose[h->cur_host], sh.pose_all, 4 * d.n, ncclDouble, sh.comm, ctx.stream));
    PF_LAUNCH(ctx, pf_gather_sharded_kernel, cdiv_u(d.n, PF_NT), PF_NT, 0, d, reinterpret_cast<const Pose4*>(sh.pose_all));
    PF_LAUNCH(ctx, pf_flip_kernel, 1, 1, 0, d);
    h->cur_host ^= 1;
    return 0;
}
static int pf_normalize(pfgpu_pf* h) {
    int rc = pf_total(h, XsValArray{h->d.w_raw}, h->d.scal + 0);
    if (rc) return rc;
    PF_LAUNCH(h->ctx, pf_normalize_kernel, cdiv_u(h->d.n, PF_NT), PF_NT, 0, h->d);
    return 0;
}
// resample_adaptive with a changing particle count (mcl.rs:322-365), see pf_kld.cuh
static int pf_resample_adaptive(pfgpu_pf* h) {
    PfDev& d = h->d; PfKld& k = h->kld; Ctx& ctx = h->ctx;
    PF_LAUNCH(ctx, pf_gate_kernel, 1, 1, 0, d, h->cfg.resample_threshold, h->cfg.mode);           // MCL resamples every step (mcl.rs:298)
    int rc = xs_scan(ctx, h->xs, XsValArray{d.w}, XsSinkStore{d.cum}, d.n, d.n_global, 0.0, d.scal + 2);   // mcl.rs:328-333
    if (rc) return rc;
    PF_LAUNCH(ctx, pf_force_last_kernel, 1, 1, 0, d);                                             // mcl.rs:334-336
    PF_CUDA(cudaMemsetAsync(k.owner, 0xFF, (size_t)k.tcap * sizeof(int), ctx.stream));
    PF_CUDA(cudaMemsetAsync(k.mint, 0xFF, (size_t)k.tcap * sizeof(unsigned), ctx.stream));
    PF_LAUNCH(ctx, pf_kld_draw_kernel, cdiv_u(k.cap, PF_NT), PF_NT, 0, d, h->seed, k);
    PF_LAUNCH(ctx, pf_kld_insert_kernel, cdiv_u(k.cap, PF_NT), PF_NT, 0, k);
    PF_LAUNCH(ctx, pf_kld_stop_kernel, 1, 1024, 0, k, (unsigned long long)h->cfg.n_particles, (unsigned long long)h->cfg.max_particles,
              h->cfg.kld_epsilon, h->cfg.kld_z);
    PF_LAUNCH(ctx, pf_kld_gather_kernel, cdiv_u(k.cap, PF_NT), PF_NT, 0, d, k);
    PF_LAUNCH(ctx, pf_flip_kernel, 1, 1, 0, d);
    unsigned* hp = reinterpret_cast<unsigned*>(h->h_pin + 34);
    PF_CUDA(cudaMemcpyAsync(hp, k.n_new, sizeof(unsigned), cudaMemcpyDeviceToHost, ctx.stream));
    PF_CUDA(cudaStreamSynchronize(ctx.stream));                    // the next launches are sized by the new count
    d.n = d.n_global = (size_t)*hp;
    return 0;
}
static int pf_resample_impl(pfgpu_pf* h) {
    PfDev& d = h->d;
    if (h->world > 1) return pf_resample_sharded(h);
    if (h->adaptive) return pf_resample_adaptive(h);
    int rc = xs_total(h->ctx, h->xs, PfValWSq{d.w}, d.n, d.n_global, 0.0, d.scal + 1);        // calc_n_eff pf.rs:416-423
    if (rc) return rc;
    PF_LAUNCH(h->ctx, pf_gate_kernel, 1, 1, 0, d, h->cfg.resample_threshold, h->cfg.mode);
    // cumulative weights (pf.rs:448-453), exact; the kernels below are no-ops when the gate is closed
    h->xs.gate = d.gate;
    rc = xs_scan(h->ctx, h->xs, XsValArray{d.w}, XsSinkStore{d.cum}, d.n, d.n_global, 0.0, d.scal + 2);
    h->xs.gate = nullptr;
    if (rc) return rc;
    if (h->cfg.mode == 1) PF_LAUNCH(h->ctx, pf_force_last_kernel, 1, 1, 0, d);
    PF_LAUNCH(h->ctx, pf_search_kernel, cdiv_u(d.n, PF_NT), PF_NT, 0, d, h->seed, h->cfg.mode);
    PF_LAUNCH(h->ctx, pf_gather_kernel, cdiv_u(d.n, PF_NT), PF_NT, 0, d);
    PF_LAUNCH(h->ctx, pf_flip_kernel, 1, 1, 0, d);
    return 0;
}
// The resample draw counter (Philox "call" index) advances only when a resample happened; it lives on the
// device (PfDev::counters[0]) next to the gate, so a step needs no host round trip.
static int pf_read_gate(pfgpu_pf* h, int* gate) {
    int* hp = reinterpret_cast<int*>(h->h_pin + 32);
    PF_CUDA(cudaMemcpyAsync(hp, h->d.gate, sizeof(int), cudaMemcpyDeviceToHost, h->ctx.stream));
    PF_CUDA(cudaStreamSynchronize(h->ctx.stream));
    *gate = *hp;
    return 0;
}

extern "C" int pfgpu_pf_predict(pfgpu_pf* h, const double u[2]) {
    if (!h || !u) return PFGPU_ERR_INVALID;
    if (!finite_d(u[0]) || !finite_d(u[1])) return PFGPU_ERR_INVALID;               // validate_control pf.rs:515-523
    PF_CUDA(cudaSetDevice(h->ctx.device));
    int rc = pf_launch_main<true, false>(h, u, nullptr, 0);
    if (rc) return rc;
    h->n_predict++;
    return pf_refresh_cache(h);                                                      // pf.rs:299
}
extern "C" int pfgpu_pf_update(pfgpu_pf* h, const double* obs3, size_t k) {
    if (!h || (k && !obs3)) return PFGPU_ERR_INVALID;
    PF_CUDA(cudaSetDevice(h->ctx.device));
    int rc = pf_stage_obs(h, obs3, k);
    if (rc) return rc;
    rc = pf_launch_main<false, true>(h, nullptr, obs3, k);
    if (rc) return rc;
    rc = pf_normalize(h);                                                            // pf.rs:331
    if (rc) return rc;
    return pf_refresh_cache(h);                                                      // pf.rs:332
}
extern "C" int pfgpu_pf_resample(pfgpu_pf* h, int* did) {
    if (!h) return PFGPU_ERR_INVALID;
    PF_CUDA(cudaSetDevice(h->ctx.device));
    int rc = pf_resample_impl(h);
    if (rc) return rc;
    rc = pf_refresh_cache(h);                                                        // pf.rs:343 (same values when the gate was closed)
    if (rc) return rc;
    if (did) { int g = 0; rc = pf_read_gate(h, &g); if (rc) return rc; *did = g; }
    return 0;
}
extern "C" int pfgpu_pf_step(pfgpu_pf* h, const double u[2], const double* obs3, size_t k, double est[4]) {
    if (!h || !u || (k && !obs3)) return PFGPU_ERR_INVALID;
    if (!finite_d(u[0]) || !finite_d(u[1])) return PFGPU_ERR_INVALID;
    PF_CUDA(cudaSetDevice(h->ctx.device));
    int rc = pf_stage_obs(h, obs3, k);
    if (rc) return rc;
    rc = pf_launch_main<true, true>(h, u, obs3, k);                                  // predict + likelihood, one pass
    if (rc) return rc;
    h->n_predict++;
    rc = pf_normalize(h);
    if (rc) return rc;
    rc = pf_resample_impl(h);
    if (rc) return rc;
    rc = pf_refresh_cache(h);
    if (rc) return rc;
    h->steps++;
    if (est) {
        PF_CUDA(cudaMemcpyAsync(h->h_pin, h->d.scal + 4, 4 * sizeof(double), cudaMemcpyDeviceToHost, h->ctx.stream));
        PF_CUDA(cudaStreamSynchronize(h->ctx.stream));
        for (int a = 0; a < 4; ++a) est[a] = h->h_pin[a];
    }
    return 0;
}
extern "C" int pfgpu_pf_estimate(pfgpu_pf* h, double est[4], double cov_cm[16]) {
    if (!h) return PFGPU_ERR_INVALID;
    PF_CUDA(cudaSetDevice(h->ctx.device));
    PF_CUDA(cudaMemcpyAsync(h->h_pin, h->d.scal + 4, 20 * sizeof(double), cudaMemcpyDeviceToHost, h->ctx.stream));
    PF_CUDA(cudaStreamSynchronize(h->ctx.stream));
    if (est) for (int a = 0; a < 4; ++a) est[a] = h->h_pin[a];
    if (cov_cm) for (int a = 0; a < 4; ++a) for (int b = 0; b < 4; ++b) cov_cm[b * 4 + a] = h->h_pin[4 + a * 4 + b];
    return 0;
}
extern "C" int pfgpu_pf_neff(pfgpu_pf* h, double* neff) {
    if (!h || !neff) return PFGPU_ERR_INVALID;
    PF_CUDA(cudaSetDevice(h->ctx.device));
    int rc = pf_total(h, PfValWSq{h->d.w}, h->d.scal + 1);
    if (rc) return rc;
    PF_CUDA(cudaMemcpyAsync(h->h_pin, h->d.scal + 1, sizeof(double), cudaMemcpyDeviceToHost, h->ctx.stream));
    PF_CUDA(cudaStreamSynchronize(h->ctx.stream));
    double Q = h->h_pin[0];
    *neff = Q > 0.0 ? 1.0 / Q : 0.0;
    return 0;
}
extern "C" int pfgpu_pf_set_range_noise(pfgpu_pf* h, double s) {
    if (!h || !finite_d(s) || s <= 0.0) return PFGPU_ERR_INVALID;                   // pf.rs:228-236
    h->cfg.range_noise = s;
    return 0;
}
extern "C" int pfgpu_pf_last_indices(pfgpu_pf* h, uint32_t* idx, size_t cap, size_t* n) {
    if (!h || !idx) return PFGPU_ERR_INVALID;
    PF_CUDA(cudaSetDevice(h->ctx.device));
    size_t c = cap < h->d.n ? cap : h->d.n;
    PF_CUDA(cudaMemcpyAsync(idx, h->d.idx, c * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->ctx.stream));
    PF_CUDA(cudaStreamSynchronize(h->ctx.stream));
    if (n) *n = h->d.n;
    return 0;
}

static void timer_drain(KernelTimer& t) {
    for (auto& p : t.pending) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, p.first, p.second) == cudaSuccess) { t.ms_sum += ms; t.count++; }
        cudaEventDestroy(p.first); cudaEventDestroy(p.second);
    }
    t.pending.clear();
}
static int read_fx_flags(FxWork& fx, bool on, pfgpu_stats* s) {
    if (!on) return 0;
    int f[4] = {0, 0, 0, 0};
    PF_CUDA(cudaMemcpy(f, fx.flags, 4 * sizeof(int), cudaMemcpyDeviceToHost));
    s->serial_fallbacks += (uint64_t)f[1] + (uint64_t)f[2];
    return 0;
}
static int read_xs_flags(Ctx& ctx, XsWork& xs, pfgpu_stats* s) {
    int f[4] = {0, 0, 0, 0};
    PF_CUDA(cudaMemcpy(f, xs.flags + 4, 4 * sizeof(int), cudaMemcpyDeviceToHost));
    s->serial_fallbacks = (uint64_t)f[0];
    s->xsum_dirty_last = (uint64_t)(f[1] < 0 ? 0 : f[1]);
    (void)ctx;
    return 0;
}
extern "C" int pfgpu_pf_stats(pfgpu_pf* h, pfgpu_stats* s) {
    if (!h || !s) return PFGPU_ERR_INVALID;
    PF_CUDA(cudaSetDevice(h->ctx.device));
    PF_CUDA(cudaStreamSynchronize(h->ctx.stream));
    memset(s, 0, sizeof(*s));
    timer_drain(h->timer);
    unsigned int cnt = 0;
    PF_CUDA(cudaMemcpy(&cnt, h->d.counters, sizeof(unsigned int), cudaMemcpyDeviceToHost));
    s->kernel_launches = h->ctx.launches; s->steps = h->steps; s->resamples = cnt;
    s->main_kernel_ms_sum = h->timer.ms_sum; s->main_kernel_count = h->timer.count;
    return read_xs_flags(h->ctx, h->xs, s);
}
extern "C" int pfgpu_pf_time_main_kernel(pfgpu_pf* h, int on) {
    if (!h) return PFGPU_ERR_INVALID;
    PF_CUDA(cudaSetDevice(h->ctx.device));
    PF_CUDA(cudaStreamSynchronize(h->ctx.stream));
    timer_drain(h->timer);
    h->timer.on = on != 0; h->timer.ms_sum = 0.0; h->timer.count = 0;
    return 0;
}

// ====================================================================================================
// FastSLAM 1.0
// ====================================================================================================
struct pfgpu_fs {
    Ctx ctx;
    pfgpu_fs_config cfg;
    uint64_t seed = 0;
    FsDev d;
    XsWork xs;
    size_t obs_cap = 0;
    int best_blocks = 0;
    uint32_t n_step = 0, n_resample = 0;
    uint64_t steps = 0;
    int world = 1, rank = 0;
    KernelTimer timer;
    Marks marks;
    double* h_pin = nullptr;
    FsObsDev* h_obs = nullptr;     // pinned staging for the observation list
    size_t lm_bytes = 0;
    FxWork fx;                     // workspace of the fused post-step kernel
    bool fused_post = false;
    unsigned fx_nt = 0;
    bool step_v2 = true;           // observation-parallel step kernel (PFGPU_STEP_V2=0 selects the one-thread-per-particle form)
    bool pdl = true;               // programmatic dependent launch for the kernels of a step (PFGPU_PDL=0 turns it off)
    bool compose_vec = true;       // 4 slots per thread in the ancestry composition (PFGPU_COMPOSE_VEC=0: one)
    int ekf_variant = 3;           // register budget of fs_ekf_kernel: 0 = 64 regs, 1 = 72 regs (2 CTAs/SM), 2 = up to 128 regs (1 CTA/SM)
    FsShard sh;                    // multi-GPU state (world == 1: unused)
    // peer-memory form of the sharded step (fs_mg.cuh): one arena per rank, mapped by every peer
    char* arena = nullptr; size_t arena_bytes = 0;
    void* peer_ptr[SH_MAX_WORLD] = {};
    char** d_peer = nullptr;
    MgDev mg = {};
    bool mg_on = false;
};
static int fs_mg_error(int code);

extern "C" void pfgpu_fs_default_config(pfgpu_fs_config* c) {            // fs1.rs:13-23
    c->dt = 0.1; c->max_range = 20.0; c->nth = 100.0 / 1.5; c->q00 = 0.3; c->q11 = 0.0305; c->r00 = 0.5; c->r11 = 0.0305;
    c->init_weight = 1.0 / 100.0;
}

static int fs_create_impl(const pfgpu_fs_config* cfg, size_t n, size_t n_global, size_t offset, size_t m, uint64_t seed, int device,
                          const void* uid, int rank, int world, pfgpu_fs** out) {
    if (!out) return PFGPU_ERR_INVALID;
    *out = nullptr;
    if (!cfg || n == 0 || n_global > 0xFFFFFFFFull) return PFGPU_ERR_INVALID;
    pfgpu_fs* h = new (std::nothrow) pfgpu_fs();
    if (!h) return PFGPU_ERR_CUDA;
    int rc = ctx_open(h->ctx, device);
    if (rc) { delete h; return rc; }
    h->cfg = *cfg; h->seed = seed; h->world = world; h->rank = rank;
    FsDev& d = h->d;
    d.n = n; d.n_global = n_global; d.offset = offset; d.m = m; d.eager = 0;
    h->sh.n_guest = world > 1 ? std::max<size_t>(2048, n / 4) : 0;
    if (world > 1) { const char* eg = getenv("PFGPU_GUEST_COLS"); if (eg && atoll(eg) > 0) h->sh.n_guest = (size_t)atoll(eg); }
    d.ld = n + h->sh.n_guest;
    h->lm_bytes = (m ? m : 1) * 6 * d.ld * sizeof(double);
    auto fail = [&](int code) { pfgpu_fs_destroy(h); return code; };
#define FS_TRY(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) { snprintf(g_pfgpu_err, sizeof(g_pfgpu_err), "%s -> %s", #x, cudaGetErrorString(e__)); return fail(PFGPU_ERR_CUDA); } } while (0)
    // Sharded mode keeps everything a peer may touch in ONE allocation with the same layout on every rank, so that a
    // single IPC mapping per peer gives access to all of it (fs_mg.cuh).
    bool want_mg = false;
    {
        const char* env = getenv("PFGPU_SHARD_P2P");
        want_mg = world > 1 && !(env && env[0] == '0') && n % FX_TILE == 0 && (size_t)world * (n / FX_TILE) <= MG_MAX_TILES;
    }
    if (want_mg) {
        MgDev& mg = h->mg;
        const size_t mm = m ? m : 1;
        const unsigned ntl = (unsigned)(n / FX_TILE), NT = ntl * (unsigned)world;
        size_t off = 0;
        auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
        mg.G = world; mg.rank = rank; mg.ntl = ntl; mg.NT = NT;
        mg.o_ctr = take(512); mg.o_bad = mg.o_ctr + 384;
        for (int sl = 0; sl < FX_SLOTS; ++sl) {
            mg.o_tsum[sl] = take(NT * sizeof(double)); mg.o_ttail[sl] = take(NT * sizeof(xs_t));
            mg.o_tnd[sl] = take(NT * sizeof(int)); mg.o_ent[sl] = take((size_t)NT * XS_MAXD * sizeof(XsEntry));
        }
        mg.o_cum_all = take(n_global * sizeof(double));
        mg.o_w_raw = take(n * sizeof(double)); mg.o_w = take(n * sizeof(double));
        for (int b = 0; b < 2; ++b) { mg.o_px[b] = take(n * sizeof(double)); mg.o_py[b] = take(n * sizeof(double)); mg.o_pyaw[b] = take(n * sizeof(double)); }
        mg.o_lmstate = take(mm * sizeof(int));
        for (int b = 0; b < 2; ++b) mg.o_anc[b] = take(mm * n * sizeof(uint32_t));
        for (int b = 0; b < 2; ++b) mg.o_lm[b] = take(h->lm_bytes);
        h->arena_bytes = off;
        FS_TRY(cudaMalloc(&h->arena, off));
        FS_TRY(cudaMemset(h->arena, 0, mg.o_cum_all));
        char* A = h->arena;
        for (int b = 0; b < 2; ++b) {
            d.px[b] = (double*)(A + mg.o_px[b]); d.py[b] = (double*)(A + mg.o_py[b]); d.pyaw[b] = (double*)(A + mg.o_pyaw[b]);
            d.lm[b] = (double*)(A + mg.o_lm[b]); d.anc[b] = (uint32_t*)(A + mg.o_anc[b]);
        }
        d.w = (double*)(A + mg.o_w); d.w_raw = (double*)(A + mg.o_w_raw); d.lmstate = (int*)(A + mg.o_lmstate);
        for (int sl = 0; sl < FX_SLOTS; ++sl) {
            h->fx.slot[sl].tsum = (double*)(A + mg.o_tsum[sl]); h->fx.slot[sl].ttail = (xs_t*)(A + mg.o_ttail[sl]);
            h->fx.slot[sl].tnd = (int*)(A + mg.o_tnd[sl]); h->fx.slot[sl].ent = (XsEntry*)(A + mg.o_ent[sl]);
        }
    } else {
    for (int b = 0; b < 2; ++b) {
        FS_TRY(cudaMalloc(&d.px[b], n * sizeof(double))); FS_TRY(cudaMalloc(&d.py[b], n * sizeof(double)));
        FS_TRY(cudaMalloc(&d.pyaw[b], n * sizeof(double))); FS_TRY(cudaMalloc(&d.lm[b], h->lm_bytes));
    }
    FS_TRY(cudaMalloc(&d.w, n * sizeof(double))); FS_TRY(cudaMalloc(&d.w_raw, n * sizeof(double)));
    }
    FS_TRY(cudaMalloc(&d.cur, sizeof(int))); FS_TRY(cudaMemset(d.cur, 0, sizeof(int)));
    FS_TRY(cudaMalloc(&d.cum, n * sizeof(double))); FS_TRY(cudaMalloc(&d.rcomb, n * sizeof(double)));
    FS_TRY(cudaMalloc(&d.idx, n * sizeof(uint32_t)));
    FS_TRY(cudaMalloc(&d.scal, 16 * sizeof(double))); FS_TRY(cudaMemset(d.scal, 0, 16 * sizeof(double)));
    FS_TRY(cudaMalloc(&d.gate, sizeof(int))); FS_TRY(cudaMemset(d.gate, 0, sizeof(int)));
    d.anc16 = (n <= 65536 && world == 1) ? 1 : 0;
    if (!h->arena) for (int b = 0; b < 2; ++b) FS_TRY(cudaMalloc(&d.anc[b], (m ? m : 1) * n * sizeof(uint32_t)));
    FS_TRY(cudaMalloc(&d.anc_cur, sizeof(int))); FS_TRY(cudaMemset(d.anc_cur, 0, sizeof(int)));
    if (!h->arena) FS_TRY(cudaMalloc(&d.lmstate, (m ? m : 1) * sizeof(int)));
    FS_TRY(cudaMalloc(&d.counters, 4 * sizeof(unsigned int))); FS_TRY(cudaMemset(d.counters, 0, 4 * sizeof(unsigned int)));
    h->obs_cap = FS_MAX_OBS;
    FS_TRY(cudaMalloc(&d.obs, h->obs_cap * sizeof(FsObsDev)));
    h->best_blocks = (int)std::min<size_t>((size_t)h->ctx.num_sms * 2, cdiv_u(n, 256));
    FS_TRY(cudaMalloc(&d.best_w, h->best_blocks * sizeof(double)));
    FS_TRY(cudaMalloc(&d.best_i, h->best_blocks * sizeof(unsigned long long)));
    FS_TRY(cudaMallocHost(&h->h_pin, (64 + 2 * (size_t)h->best_blocks) * sizeof(double)));
    FS_TRY(cudaMallocHost(&h->h_obs, h->obs_cap * sizeof(FsObsDev)));
#undef FS_TRY
    rc = xs_work_alloc(h->xs, n);
    if (rc) return fail(rc);
    { const char* e2 = getenv("PFGPU_STEP_V2"); h->step_v2 = !(e2 && e2[0] == '0'); }
    { const char* e5 = getenv("PFGPU_PDL"); h->pdl = (e5 && e5[0] == '2') || (!(e5 && e5[0] == '0') && world == 1); }   // sharded steps: not measured with it yet (PFGPU_PDL=2 forces it)
    { const char* e4 = getenv("PFGPU_COMPOSE_VEC"); h->compose_vec = !(e4 && e4[0] == '0'); }
    { const char* e3 = getenv("PFGPU_EKF_VARIANT"); if (e3 && e3[0] >= '0' && e3[0] <= '4') h->ekf_variant = e3[0] - '0'; }
    {   // fused post-step kernel: usable when one co-resident wave covers all tiles
        unsigned nt = cdiv_u(n, FX_TILE);
        int nb = 0;
        const char* env = getenv("PFGPU_FUSED_POST");
        bool want = !(env && env[0] == '0') && world == 1;
        if (want && nt <= FX_MAX_TILES &&
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, fs_post_kernel, XS_NT, 0) == cudaSuccess &&
            (size_t)nb * (size_t)h->ctx.num_sms >= nt) {
            int coop = 0;
            cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device);
            if (coop) {
                bool okalloc = true;
                for (int sl = 0; sl < FX_SLOTS && okalloc; ++sl) {
                    okalloc = cudaMalloc(&h->fx.slot[sl].tsum, nt * sizeof(double)) == cudaSuccess &&
                              cudaMalloc(&h->fx.slot[sl].ttail, nt * sizeof(xs_t)) == cudaSuccess &&
                              cudaMalloc(&h->fx.slot[sl].tnd, nt * sizeof(int)) == cudaSuccess &&
                              cudaMalloc(&h->fx.slot[sl].ent, (size_t)nt * XS_MAXD * sizeof(XsEntry)) == cudaSuccess;
                }
                okalloc = okalloc && cudaMalloc(&h->fx.flags, 8 * sizeof(int)) == cudaSuccess &&
                          cudaMemset(h->fx.flags, 0, 8 * sizeof(int)) == cudaSuccess;
                h->fx.dbg = nullptr;
                if (getenv("PFGPU_POST_TRACE")) {
                    okalloc = okalloc && cudaMalloc(&h->fx.dbg, 32 * sizeof(unsigned long long)) == cudaSuccess &&
                              cudaMemset(h->fx.dbg, 0, 32 * sizeof(unsigned long long)) == cudaSuccess;
                }
                if (!okalloc) return fail(PFGPU_ERR_CUDA);
                h->fused_post = true; h->fx_nt = nt;
            }
        }
    }
    {   // ancestry log (PFGPU_ANC_LOG=1; fs_kernels.cuh FsDev::alog): resamples stop composing the ancestry rows
        const char* ea = getenv("PFGPU_ANC_LOG");
        if (ea && ea[0] == '1' && world == 1 && h->fused_post && n % 4 == 0 && h->compose_vec && m > 0) {
            int R = 32;
            const char* er = getenv("PFGPU_ANC_LOG_R");
            if (er && atoi(er) >= 1 && atoi(er) <= 4096) R = atoi(er);
            if (cudaMalloc(&d.idxlog, (size_t)R * n * sizeof(uint32_t)) != cudaSuccess || cudaMalloc(&d.gen, m * sizeof(int)) != cudaSuccess ||
                cudaMemset(d.gen, 0, m * sizeof(int)) != cudaSuccess) return fail(PFGPU_ERR_CUDA);
            d.alog = R;
        }
    }
    if (world > 1) {
        FsShard& sh = h->sh;
        sh.rank = rank; sh.world = world;
        ncclUniqueId id;
        memcpy(&id, uid, sizeof(id));
        ncclResult_t nr = ncclCommInitRank(&sh.comm, world, id, rank);
        if (nr != ncclSuccess) { snprintf(g_pfgpu_err, sizeof(g_pfgpu_err), "ncclCommInitRank: %s", ncclGetErrorString(nr)); return fail(PFGPU_ERR_NCCL); }
#define SH_TRY(x) do { if ((x) != cudaSuccess) { snprintf(g_pfgpu_err, sizeof(g_pfgpu_err), "%s failed", #x); return fail(PFGPU_ERR_CUDA); } } while (0)
        SH_TRY(cudaMalloc(&sh.t_loc, sizeof(double))); SH_TRY(cudaMalloc(&sh.t_all, world * sizeof(double)));
        SH_TRY(cudaMalloc(&sh.approx_off, sizeof(double))); SH_TRY(cudaMalloc(&sh.sum_loc, sizeof(ShardSummary)));
        SH_TRY(cudaMalloc(&sh.sum_all, world * sizeof(ShardSummary))); SH_TRY(cudaMalloc(&sh.s_start, sizeof(double)));
        SH_TRY(cudaMalloc(&sh.err, sizeof(int))); SH_TRY(cudaMemset(sh.err, 0, sizeof(int)));
        SH_TRY(cudaMalloc(&sh.cum_all, n_global * sizeof(double))); SH_TRY(cudaMalloc(&sh.idx_all, n_global * sizeof(uint32_t)));
        SH_TRY(cudaMalloc(&sh.pose_all, 3 * n_global * sizeof(double))); SH_TRY(cudaMallocHost(&sh.h_idx, n_global * sizeof(uint32_t)));
        SH_TRY(cudaMalloc(&sh.best_loc, 8 * sizeof(double))); SH_TRY(cudaMalloc(&sh.best_all, 8 * world * sizeof(double)));
        {   // NCCL sets up peer connections lazily, on the first send/recv or collective of each kind (>100 ms once): do it here,
            // not inside somebody's timed step
            PF_NCCL(ncclAllGather(sh.t_loc, sh.t_all, 1, ncclDouble, sh.comm, h->ctx.stream));
            PF_NCCL(ncclAllGather(sh.sum_loc, sh.sum_all, sizeof(ShardSummary), ncclChar, sh.comm, h->ctx.stream));
            PF_NCCL(ncclAllGather(d.idx, sh.idx_all, n, ncclUint32, sh.comm, h->ctx.stream));
            PF_NCCL(ncclAllGather(d.cum, sh.cum_all, n, ncclDouble, sh.comm, h->ctx.stream));
            PF_NCCL(ncclGroupStart());
            for (int g = 0; g < world; ++g) {
                if (g == rank) continue;
                PF_NCCL(ncclSend(sh.best_loc, 8, ncclDouble, g, sh.comm, h->ctx.stream));
                PF_NCCL(ncclRecv(sh.best_all + 8 * g, 8, ncclDouble, g, sh.comm, h->ctx.stream));
            }
            PF_NCCL(ncclGroupEnd());
            SH_TRY(cudaStreamSynchronize(h->ctx.stream));
        }
        if (h->arena) {
            // map every peer's arena; all ranks must agree on the outcome, so the verdict is gathered too
            int ok = 1;
            cudaIpcMemHandle_t mine, all[SH_MAX_WORLD];
            char* d_hand = nullptr;
            SH_TRY(cudaMalloc(&d_hand, (size_t)(world + 1) * sizeof(cudaIpcMemHandle_t)));
            if (cudaIpcGetMemHandle(&mine, h->arena) != cudaSuccess) { ok = 0; memset(&mine, 0, sizeof(mine)); cudaGetLastError(); }
            SH_TRY(cudaMemcpy(d_hand + (size_t)world * sizeof(mine), &mine, sizeof(mine), cudaMemcpyHostToDevice));
            PF_NCCL(ncclAllGather(d_hand + (size_t)world * sizeof(mine), d_hand, sizeof(mine), ncclChar, sh.comm, h->ctx.stream));
            SH_TRY(cudaStreamSynchronize(h->ctx.stream));
            SH_TRY(cudaMemcpy(all, d_hand, (size_t)world * sizeof(mine), cudaMemcpyDeviceToHost));
            for (int g = 0; g < world && ok; ++g) {
                if (g == rank) { h->peer_ptr[g] = h->arena; continue; }
                if (cudaIpcOpenMemHandle(&h->peer_ptr[g], all[g], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                    ok = 0; h->peer_ptr[g] = nullptr; cudaGetLastError();
                }
            }
            int nb = 0, coop = 0;
            cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device);
            if (cudaFuncSetAttribute(fs_post_mg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(MgShared)) != cudaSuccess ||
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, fs_post_mg_kernel, XS_NT, sizeof(MgShared)) != cudaSuccess ||
                (size_t)nb * (size_t)h->ctx.num_sms < h->mg.ntl || !coop) { ok = 0; cudaGetLastError(); }
            int* d_ok = reinterpret_cast<int*>(d_hand);
            SH_TRY(cudaMemcpy(d_ok + world, &ok, sizeof(int), cudaMemcpyHostToDevice));
            PF_NCCL(ncclAllGather(d_ok + world, d_ok, 1, ncclInt, sh.comm, h->ctx.stream));
            SH_TRY(cudaStreamSynchronize(h->ctx.stream));
            int oks[SH_MAX_WORLD];
            SH_TRY(cudaMemcpy(oks, d_ok, (size_t)world * sizeof(int), cudaMemcpyDeviceToHost));
            cudaFree(d_hand);
            for (int g = 0; g < world; ++g) ok = ok && oks[g];
            if (ok) {
                MgDev& mg = h->mg;
                SH_TRY(cudaMalloc(&h->d_peer, world * sizeof(char*)));
                SH_TRY(cudaMemcpy(h->d_peer, h->peer_ptr, world * sizeof(char*), cudaMemcpyHostToDevice));
                mg.peer = h->d_peer;
                SH_TRY(cudaMalloc(&mg.tgt, 4 * sizeof(unsigned))); SH_TRY(cudaMemset(mg.tgt, 0, 4 * sizeof(unsigned)));
                SH_TRY(cudaMalloc(&mg.err, sizeof(int))); SH_TRY(cudaMemset(mg.err, 0, sizeof(int)));
                SH_TRY(cudaMalloc(&mg.gcol, n * sizeof(unsigned)));
                SH_TRY(cudaMalloc(&mg.plan, 32 * sizeof(unsigned long long))); SH_TRY(cudaMemset(mg.plan, 0, 32 * sizeof(unsigned long long)));
                mg.n_guest = sh.n_guest;
                SH_TRY(cudaMalloc(&h->fx.flags, 8 * sizeof(int))); SH_TRY(cudaMemset(h->fx.flags, 0, 8 * sizeof(int)));
                h->fx.dbg = nullptr;
                if (getenv("PFGPU_POST_TRACE")) {
                    SH_TRY(cudaMalloc(&h->fx.dbg, 32 * sizeof(unsigned long long)));
                    SH_TRY(cudaMemset(h->fx.dbg, 0, 32 * sizeof(unsigned long long)));
                    unsigned long long one = 1;
                    SH_TRY(cudaMemcpy(mg.plan + 13, &one, sizeof(one), cudaMemcpyHostToDevice));
                }
                h->mg_on = true;
                // nobody may start pushing into an arena before its owner has zeroed it (cudaMemset above, synchronous) and
                // everybody has mapped it: one more collective as the fence
                PF_NCCL(ncclAllGather(sh.t_loc, sh.t_all, 1, ncclDouble, sh.comm, h->ctx.stream));
                SH_TRY(cudaStreamSynchronize(h->ctx.stream));
            }
        }
        {   // exchange buffers sized for the worst admissible resample up front: no cudaMalloc on the step path
            size_t cap = std::max<size_t>(3 * n, sh.n_guest * 6 * (m ? m : 1)) + 1024;
            SH_TRY(cudaMalloc(&sh.sendbuf, cap * sizeof(double))); sh.send_cap = cap;
            SH_TRY(cudaMalloc(&sh.recvbuf, cap * sizeof(double))); sh.recv_cap = cap;
        }
#undef SH_TRY
    }
    fs_init_kernel<<<cdiv_u(n, 256), 256, 0, h->ctx.stream>>>(d, cfg->init_weight);
    fs_lmstate_reset_kernel<<<1, 256, 0, h->ctx.stream>>>(d);
    h->ctx.launches += 2;
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
    if (!uid || world < 1 || world > SH_MAX_WORLD || rank < 0 || rank >= world || n_global == 0 || n_global % (size_t)world != 0)
        return PFGPU_ERR_INVALID;
    size_t nl = n_global / (size_t)world;
    return fs_create_impl(cfg, nl, n_global, (size_t)rank * nl, m, seed, device, uid, rank, world, out);
}
extern "C" void pfgpu_fs_destroy(pfgpu_fs* h) {
    if (!h) return;
    cudaSetDevice(h->ctx.device);
    if (h->ctx.stream) cudaStreamSynchronize(h->ctx.stream);
    FsDev& d = h->d;
    if (h->arena) {      // these live inside the arena
        for (int b = 0; b < 2; ++b) { d.px[b] = d.py[b] = d.pyaw[b] = d.lm[b] = nullptr; d.anc[b] = nullptr; }
        d.w = d.w_raw = nullptr; d.lmstate = nullptr;
        for (int sl = 0; sl < FX_SLOTS; ++sl) { h->fx.slot[sl].tsum = nullptr; h->fx.slot[sl].ttail = nullptr; h->fx.slot[sl].tnd = nullptr; h->fx.slot[sl].ent = nullptr; }
        for (int g = 0; g < h->world; ++g) if (g != h->rank && h->peer_ptr[g]) cudaIpcCloseMemHandle(h->peer_ptr[g]);
        cudaFree(h->d_peer); cudaFree(h->mg.tgt); cudaFree(h->mg.err); cudaFree(h->mg.gcol); cudaFree(h->mg.plan);
        cudaFree(h->arena);
    }
    for (int b = 0; b < 2; ++b) { cudaFree(d.px[b]); cudaFree(d.py[b]); cudaFree(d.pyaw[b]); cudaFree(d.lm[b]); }
    cudaFree(d.cur); cudaFree(d.w); cudaFree(d.w_raw); cudaFree(d.cum); cudaFree(d.rcomb); cudaFree(d.idx); cudaFree(d.scal);
    cudaFree(d.gate); cudaFree(d.obs); cudaFree(d.best_w); cudaFree(d.best_i); cudaFree(d.counters);
    cudaFree(d.anc[0]); cudaFree(d.anc[1]); cudaFree(d.anc_cur); cudaFree(d.lmstate); cudaFree(d.idxlog); cudaFree(d.gen);
    for (int sl = 0; sl < FX_SLOTS; ++sl) { cudaFree(h->fx.slot[sl].tsum); cudaFree(h->fx.slot[sl].ttail); cudaFree(h->fx.slot[sl].tnd); cudaFree(h->fx.slot[sl].ent); }
    cudaFree(h->fx.flags); cudaFree(h->fx.dbg);
    {
        FsShard& sh = h->sh;
        cudaFree(sh.t_loc); cudaFree(sh.t_all); cudaFree(sh.approx_off); cudaFree(sh.sum_loc); cudaFree(sh.sum_all); cudaFree(sh.s_start);
        cudaFree(sh.err); cudaFree(sh.cum_all); cudaFree(sh.idx_all); cudaFree(sh.pose_all); cudaFree(sh.sendbuf); cudaFree(sh.recvbuf);
        cudaFree(sh.best_loc); cudaFree(sh.best_all);
        if (sh.h_idx) cudaFreeHost(sh.h_idx);
        if (sh.comm) ncclCommDestroy(sh.comm);
    }
    if (h->h_pin) cudaFreeHost(h->h_pin);
    if (h->h_obs) cudaFreeHost(h->h_obs);
    marks_free(h->marks);
    xs_work_free(h->xs);
    for (auto& p : h->timer.pending) { cudaEventDestroy(p.first); cudaEventDestroy(p.second); }
    if (h->ctx.stream) cudaStreamDestroy(h->ctx.stream);
    delete h;
}
extern "C" int pfgpu_fs_sync(pfgpu_fs* h) {
    if (!h) return PFGPU_ERR_INVALID;
    PF_CUDA(cudaSetDevice(h->ctx.device));
    PF_CUDA(cudaStreamSynchronize(h->ctx.stream));
    if (h->mg_on) {        // errors of the device-side sharded step are sticky and surface at the next sync
        int e = 0;
        PF_CUDA(cudaMemcpy(&e, h->mg.err, sizeof(int), cudaMemcpyDeviceToHost));
        if (e) return fs_mg_error(e);
    }
    return 0;
}
extern "C" int pfgpu_fs_count(pfgpu_fs* h, size_t* nl, size_t* ng, size_t* m) {
    if (!h) return PFGPU_ERR_INVALID;
    if (nl) *nl = h->d.n;
    if (ng) *ng = h->d.n_global;
    if (m) *m = h->d.m;
    return 0;
}
static const size_t FS_XFER_CHUNK_BYTES = (size_t)256 << 20;   // staging chunk for AoS<->SoA conversion
extern "C" int pfgpu_fs_upload(pfgpu_fs* h, const double* pose_w, const double* lm, size_t n) {
    if (!h || !pose_w || n != h->d.n) return PFGPU_ERR_INVALID;
    PF_CUDA(cudaSetDevice(h->ctx.device));
    FsDev& d = h->d;
    double* tmp = nullptr;
    PF_CUDA(cudaMalloc(&tmp, n * 4 * sizeof(double)));
    PF_CUDA(cudaMemcpyAsync(tmp, pose_w, n * 4 * sizeof(double), cudaMemcpyHostToDevice, h->ctx.stream));
    PF_LAUNCH(h->ctx, fs_unpack_pose_kernel, cdiv_u(n, 256), 256, 0, d, tmp);
    PF_CUDA(cudaStreamSynchronize(h->ctx.stream));
    cudaFree(tmp);
    if (lm && d.m) {
        size_t per = d.m * 6 * sizeof(double);
        size_t chunk = std::max<size_t>(1, FS_XFER_CHUNK_BYTES / per);
        if (chunk > n) chunk = n;
        PF_CUDA(cudaMalloc(&tmp, chunk * per));
        for (size_t i0 = 0; i0 < n; i0 += chunk) {
            size_t cnt = std::min(chunk, n - i0);
            PF_CUDA(cudaMemcpyAsync(tmp, lm + i0 * d.m * 6, cnt * per, cudaMemcpyHostToDevice, h->ctx.stream));
            PF_LAUNCH(h->ctx, fs_unpack_lm_kernel, cdiv_u(cnt * d.m * 6, 256), 256, 0, d, tmp, i0, cnt);
            PF_CUDA(cudaStreamSynchronize(h->ctx.stream));
        }
        PF_LAUNCH(h->ctx, fs_lmstate_reset_kernel, 1, 256, 0, d);
        cudaFree(tmp);
    }
    return 0;
}
extern "C" int pfgpu_fs_download(pfgpu_fs* h, double* pose_w, double* lm, size_t n) {
    if (!h || n != h->d.n) return PFGPU_ERR_INVALID;
    PF_CUDA(cudaSetDevice(h->ctx.device));
    FsDev& d = h->d;
    double* tmp = nullptr;
    if (pose_w) {
        PF_CUDA(cudaMalloc(&tmp, n * 4 * sizeof(double)));
        PF_LAUNCH(h->ctx, fs_pack_pose_kernel, cdiv_u(n, 256), 256, 0, d, tmp);
        PF_CUDA(cudaMemcpyAsync(pose_w, tmp, n * 4 * sizeof(double), cudaMemcpyDeviceToHost, h->ctx.stream));
        PF_CUDA(cudaStreamSynchronize(h->ctx.stream));
        cudaFree(tmp);
    }
    if (lm && d.m) {
        size_t per = d.m * 6 * sizeof(double);
        size_t chunk = std::max<size_t>(1, FS_XFER_CHUNK_BYTES / per);
        if (chunk > n) chunk = n;
        PF_CUDA(cudaMalloc(&tmp, chunk * per));
        for (size_t i0 = 0; i0 < n; i0 += chunk) {
            size_t cnt = std::min(chunk, n - i0);
            PF_LAUNCH(h->ctx, fs_pack_lm_kernel, cdiv_u(cnt * d.m * 6, 256), 256, 0, d, tmp, i0, cnt);
            PF_CUDA(cudaMemcpyAsync(lm + i0 * d.m * 6, tmp, cnt * per, cudaMemcpyDeviceToHost, h->ctx.stream));
            PF_CUDA(cudaStreamSynchronize(h->ctx.stream));
        }
        cudaFree(tmp);
    }
    return 0;
}

extern "C" int pfgpu_fs_seed_map(pfgpu_fs* h, const double pose3[3], const double* lm_xy, size_t m, double sigma, double cov0) {
    if (!h || !pose3 || (m && !lm_xy) || m != h->d.m) return PFGPU_ERR_INVALID;
    PF_CUDA(cudaSetDevice(h->ctx.device));
    FsDev& d = h->d;
    PF_LAUNCH(h->ctx, fs_seed_pose_kernel, cdiv_u(d.n, 256), 256, 0, d, pose3[0], pose3[1], pose3[2]);
    if (m) {
        double* dxy = nullptr;
        PF_CUDA(cudaMalloc(&dxy, m * 2 * sizeof(double)));
        PF_CUDA(cudaMemcpyAsync(dxy, lm_xy, m * 2 * sizeof(double), cudaMemcpyHostToDevice, h->ctx.stream));
        dim3 grid(cdiv_u(d.n, 256), (unsigned)m);
        PF_LAUNCH(h->ctx, fs_seed_lm_kernel, grid, 256, 0, d, dxy, sigma, cov0, h->seed);
        PF_LAUNCH(h->ctx, fs_lmstate_reset_kernel, 1, 256, 0, d);
        PF_CUDA(cudaStreamSynchronize(h->ctx.stream));
        cudaFree(dxy);
    }
    return 0;
}
// Post-step pipeline of the sharded mode (see fs_sharded.cuh).  One host sync per step (the gate decides which
// collectives follow; every rank computes the same gate from the same exact global sums).
static int sh_grow(double** buf, size_t* cap, size_t need) {
    if (need <= *cap) return 0;
    if (*buf) cudaFree(*buf);
    *cap = need + need / 4 + 1024;
    PF_CUDA(cudaMalloc(buf, *cap * sizeof(double)));
    return 0;
}
static int fs_mg_error(int code) {
    if (code == 3) { snprintf(g_pfgpu_err, sizeof(g_pfgpu_err), "sharded resample: more imported particles than guest columns (shard weights too unbalanced)"); return PFGPU_ERR_UNSUPPORTED; }
    snprintf(g_pfgpu_err, sizeof(g_pfgpu_err), "sharded step: a peer GPU never arrived at a barrier (code %d)", code);
    return PFGPU_ERR_CUDA;
}
static int fs_post_sharded(pfgpu_fs* h) {
    FsDev& d = h->d; FsShard& sh = h->sh; Ctx& ctx = h->ctx;
    const size_t nl = d.n, ng = d.n_global, rows = 6 * d.m;
    int rc = xs_total_sharded(ctx, h->xs, sh, XsValArray{d.w_raw}, nl, ng, d.scal + 0);              // fs1.rs:259
    if (rc) return rc;
    PF_LAUNCH(ctx, fs_normalize_kernel, cdiv_u(nl, 256), 256, 0, d);
    // gate (fs1.rs:262-263): tree-order sum of w^2 first (one small allgather); the exact sequential sum only when neff is
    // within rounding of NTH — same decision as the reference in every case (see fs_post.cuh)
    int* hp = reinterpret_cast<int*>(h->h_pin + 32);
    double* hq = h->h_pin + 40;
    {
        unsigned nt = cdiv_u(nl, XS_TILE);
        PF_CUDA(cudaMemsetAsync(h->xs.flags, 0, 4 * sizeof(int), ctx.stream));
        PF_LAUNCH(ctx, xs_tile_sums<FsValWSq>, nt, XS_NT, 0, FsValWSq{d.w}, nl, h->xs);
        PF_LAUNCH(ctx, sh_local_total_kernel, 1, 256, 0, h->xs.tsum, nt, sh.t_loc);
        PF_NCCL(ncclAllGather(sh.t_loc, sh.t_all, 1, ncclDouble, sh.comm, ctx.stream));
        PF_CUDA(cudaMemcpyAsync(hq, sh.t_all, sh.world * sizeof(double), cudaMemcpyDeviceToHost, ctx.stream));
        PF_CUDA(cudaMemcpyAsync(hp + 1, sh.err, sizeof(int), cudaMemcpyDeviceToHost, ctx.stream));
        PF_CUDA(cudaMemcpyAsync(hp + 2, h->xs.flags + 3, sizeof(int), cudaMemcpyDeviceToHost, ctx.stream));
        PF_CUDA(cudaStreamSynchronize(ctx.stream));
    }
    if (hp[1]) { snprintf(g_pfgpu_err, sizeof(g_pfgpu_err), "sharded exact sum: a shard was not summarisable (degenerate weights)"); return PFGPU_ERR_UNSUPPORTED; }
    double Qa = 0.0;
    for (int g = 0; g < sh.world; ++g) Qa += hq[g];                       // same order on every rank: identical value
    double neff = Qa > 0.0 ? 1.0 / Qa : 0.0;
    const double slack = 8.0 * (double)(ng + 64) * 2.220446049250313e-16;
    const bool border = !(std::fabs(neff - h->cfg.nth) > slack * std::fmax(std::fabs(h->cfg.nth), std::fabs(neff))) || hp[2] != 0;
    if (border) {
        rc = xs_total_sharded(ctx, h->xs, sh, FsValWSq{d.w}, nl, ng, d.scal + 1);                    // fs1.rs:262, exact
        if (rc) return rc;
        PF_LAUNCH(ctx, fs_gate_kernel, 1, 1, 0, d, h->cfg.nth);
        PF_CUDA(cudaMemcpyAsync(hp, d.gate, sizeof(int), cudaMemcpyDeviceToHost, ctx.stream));
        PF_CUDA(cudaStreamSynchronize(ctx.stream));
    } else {
        *hp = neff < h->cfg.nth ? 1 : 0;
        double* hs = h->h_pin + 60;
        hs[0] = Qa; hs[1] = neff;
        PF_CUDA(cudaMemcpyAsync(d.scal + 1, hs, sizeof(double), cudaMemcpyHostToDevice, ctx.stream));
        PF_CUDA(cudaMemcpyAsync(d.scal + 3, hs + 1, sizeof(double), cudaMemcpyHostToDevice, ctx.stream));
        PF_CUDA(cudaMemcpyAsync(d.gate, hp, sizeof(int), cudaMemcpyHostToDevice, ctx.stream));
    }
    if (!*hp) return 0;
    // ---------------- resample fs1.rs:206-234 ----------------
    rc = xs_total_sharded(ctx, h->xs, sh, XsValArray{d.w}, nl, ng, d.scal + 2);                      // fs1.rs:207
    if (!rc) rc = xs_scan_sharded(ctx, h->xs, sh, FsValWNorm2{d.w, d.scal, d.gate}, XsSinkStore{d.cum}, nl, ng, d.scal + 4);
    if (rc) return rc;
    PF_LAUNCH(ctx, fs_comb_kernel, 1, 1, 0, d, h->seed);
    rc = xs_scan_sharded(ctx, h->xs, sh, FsValCombG{d.scal, 1.0 / (double)ng, d.offset}, XsSinkStore{d.rcomb}, nl, ng, d.scal + 5);
    if (rc) return rc;
    PF_NCCL(ncclAllGather(d.cum, sh.cum_all, nl, ncclDouble, sh.comm, ctx.stream));
    PF_LAUNCH(ctx, sh_search_kernel, cdiv_u(nl, 256), 256, 0, d, sh.cum_all);
    PF_NCCL(ncclAllGather(d.idx, sh.idx_all, nl, ncclUint32, sh.comm, ctx.stream));
    // poses: everyone needs any ancestor's pose; 24 B per particle
    rc = sh_grow(&sh.sendbuf, &sh.send_cap, 3 * nl);
    if (rc) return rc;
    PF_LAUNCH(ctx, sh_pack_pose_kernel, cdiv_u(nl, 256), 256, 0, d, sh.sendbuf);
    PF_NCCL(ncclAllGather(sh.sendbuf, sh.pose_all, 3 * nl, ncclDouble, sh.comm, ctx.stream));
    PF_LAUNCH(ctx, sh_gather_pose_kernel, cdiv_u(nl, 256), 256, 0, d, sh.pose_all);
    // maps: contiguous runs of slots per (source, destination) pair
    PF_CUDA(cudaMemcpyAsync(sh.h_idx, sh.idx_all, ng * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx.stream));
    PF_CUDA(cudaMemcpyAsync(hp + 1, sh.err, sizeof(int), cudaMemcpyDeviceToHost, ctx.stream));
    PF_CUDA(cudaStreamSynchronize(ctx.stream));
    if (hp[1]) { snprintf(g_pfgpu_err, sizeof(g_pfgpu_err), "sharded exact scan: a shard was not summarisable (degenerate weights)"); return PFGPU_ERR_UNSUPPORTED; }
    const int G = sh.world, me = sh.rank;
    // slots [lo, hi) of destination block dst whose ancestor lives on rank src
    auto run_of = [&](int dst, int src, size_t* lo, size_t* hi) {
        const uint32_t* a = sh.h_idx + (size_t)dst * nl;
        size_t b0 = std::lower_bound(a, a + nl, (uint32_t)((size_t)src * nl)) - a;
        size_t b1 = std::lower_bound(a, a + nl, (uint32_t)((size_t)(src + 1) * nl)) - a;
        if ((size_t)(src + 1) * nl > 0xFFFFFFFFull) b1 = nl;
        *lo = (size_t)dst * nl + b0; *hi = (size_t)dst * nl + b1;
    };
    size_t send_tot = 0, recv_tot = 0, recv_cnt = 0;
    size_t s_lo[SH_MAX_WORLD], s_hi[SH_MAX_WORLD], r_lo[SH_MAX_WORLD], r_hi[SH_MAX_WORLD];
    for (int g = 0; g < G; ++g) {
        s_lo[g] = s_hi[g] = r_lo[g] = r_hi[g] = 0;
        if (g == me) continue;
        run_of(g, me, &s_lo[g], &s_hi[g]);           // what I send to g
        run_of(me, g, &r_lo[g], &r_hi[g]);           // what I receive from g
        send_tot += (s_hi[g] - s_lo[g]) * rows; recv_tot += (r_hi[g] - r_lo[g]) * rows; recv_cnt += r_hi[g] - r_lo[g];
    }
    if (recv_cnt > sh.n_guest) { snprintf(g_pfgpu_err, sizeof(g_pfgpu_err), "resample imports %zu particles, guest capacity %zu", recv_cnt, sh.n_guest); return PFGPU_ERR_UNSUPPORTED; }
    if (rows && sh.guest_used + recv_cnt > sh.n_guest) {   // guests exhausted: compact (everything identity-mapped, all guests free)
        dim3 cg(cdiv_u(nl, 256), cdiv_u(d.m, 4));
        PF_LAUNCH(ctx, sh_compact_kernel, cg, 256, 0, d);
        PF_LAUNCH(ctx, sh_compact_finish_kernel, 1, 256, 0, d);
        sh.guest_used = 0; sh.compactions++;
    }
    rc = sh_grow(&sh.sendbuf, &sh.send_cap, send_tot > 3 * nl ? send_tot : 3 * nl);
    if (!rc) rc = sh_grow(&sh.recvbuf, &sh.recv_cap, recv_tot + 1);
    if (rc) return rc;
    ShRecvTable tab;
    size_t soff[SH_MAX_WORLD], so = 0, ro = 0, gc = nl + sh.guest_used;
    for (int g = 0; g < G; ++g) {
        tab.t0[g] = r_lo[g]; tab.base[g] = ro; tab.cnt[g] = r_hi[g] - r_lo[g]; tab.gcol[g] = gc;
        ro += (r_hi[g] - r_lo[g]) * rows; gc += r_hi[g] - r_lo[g];
        soff[g] = so; so += (s_hi[g] - s_lo[g]) * rows;
        size_t cnt = s_hi[g] - s_lo[g];
        if (cnt && rows) PF_LAUNCH(ctx, sh_pack_map_kernel, cdiv_u(cnt * rows, 256), 256, 0, d, sh.idx_all, s_lo[g], cnt, me, sh.sendbuf + soff[g]);
    }
    if (rows) {
        PF_NCCL(ncclGroupStart());
        for (int g = 0; g < G; ++g) {
            if (g == me) continue;
            size_t sc = (s_hi[g] - s_lo[g]) * rows, rcn = (r_hi[g] - r_lo[g]) * rows;
            if (sc) PF_NCCL(ncclSend(sh.sendbuf + soff[g], sc, ncclDouble, g, sh.comm, ctx.stream));
            if (rcn) PF_NCCL(ncclRecv(sh.recvbuf + tab.base[g], rcn, ncclDouble, g, sh.comm, ctx.stream));
        }
        PF_NCCL(ncclGroupEnd());
        for (int g = 0; g < G; ++g) {
            size_t cnt = r_hi[g] - r_lo[g];
            if (g != me && cnt) PF_LAUNCH(ctx, sh_unpack_guest_kernel, cdiv_u(cnt * rows, 256), 256, 0, d, sh.recvbuf, tab, g);
        }
        sh.guest_used += recv_cnt; sh.imported += recv_cnt;
        dim3 grid(cdiv_u(nl, 256), cdiv_u(d.m, FS_COMPOSE_ROWS));
        PF_LAUNCH(ctx, sh_compose_anc_kernel, grid, 256, 0, d, tab, me);
    }
    PF_LAUNCH(ctx, sh_flip_kernel, 1, 256, 0, d);
    return 0;
}
extern "C" int pfgpu_fs_step(pfgpu_fs* h, const double u[2], const pfgpu_fs_obs* z, size_t k, int* did) {
    if (!h || !u || (k && !z)) return PFGPU_ERR_INVALID;
    if (!finite_d(u[0]) || !finite_d(u[1])) return PFGPU_ERR_INVALID;
    if (k > h->obs_cap) return PFGPU_ERR_UNSUPPORTED;
    FsDev& d = h->d;
    for (size_t j = 0; j < k; ++j) {
        if (!finite_d(z[j].d) || !finite_d(z[j].angle)) return PFGPU_ERR_INVALID;
        if (z[j].lm_id >= d.m) return PFGPU_ERR_INVALID;             // the reference would panic on the Vec index (fs1.rs:141)
    }
    PF_CUDA(cudaSetDevice(h->ctx.device));
    // The lazy-clone bookkeeping (lmstate) is per launch, so one launch must not see the same lm_id twice: the list is
    // cut before every repeated id and the pieces run as consecutive launches (same per-particle order as fs1.rs:250-256).
    std::vector<size_t> cuts;
    cuts.push_back(0);
    {
        std::vector<uint64_t> seen;
        for (size_t j = 0; j < k; ++j) {
            bool dup = false;
            for (uint64_t v : seen) if (v == z[j].lm_id) { dup = true; break; }
            if (dup) { cuts.push_back(j); seen.clear(); }
            seen.push_back(z[j].lm_id);
        }
    }
    cuts.push_back(k);
    FsObsParam last_po; int last_k = 0;
    memset(&last_po, 0, sizeof(last_po));
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (h->timer.on) { PF_CUDA(cudaEventCreate(&e0)); PF_CUDA(cudaEventCreate(&e1)); PF_CUDA(cudaEventRecord(e0, h->ctx.stream)); }
    for (size_t seg = 0; seg + 1 < cuts.size(); ++seg) {
        const size_t j0 = cuts[seg], kk = cuts[seg + 1] - cuts[seg];
        const int do_predict = seg == 0 ? 1 : 0;
        const bool param = kk <= FS_PARAM_OBS;
        FsObsParam po;
        if (param) {
            for (size_t j = 0; j < kk; ++j) { po.o[j].d = z[j0 + j].d; po.o[j].angle = z[j0 + j].angle; po.o[j].lm_id = (int)z[j0 + j].lm_id; po.o[j].pad = 0; }
            if (h->step_v2 && kk <= FS2_MAX_OBS) {
                // observation-parallel form: predict at full occupancy, then 32 particles x kk warps per CTA
                if (do_predict)
                    PF_LAUNCH_PDL(h->ctx, h->pdl, fs_predict_kernel, cdiv_u(d.n, 256), 256, 0, d, u[0], u[1], h->cfg.dt, sqrt(h->cfg.q00),
                                  sqrt(h->cfg.q11), h->seed, (uint32_t)h->n_step);
                if (kk) {
                    size_t smem = kk * 32 * sizeof(double) + kk * sizeof(unsigned) + 8;
                    const int var = kk <= 14 ? h->ekf_variant : 0;
                    if (d.alog) {      // ancestry log: the read path walks the ring (two instantiations: 3 CTAs/SM up to 14 observations, else 1)
                        if (kk <= 14) PF_LAUNCH_PDL(h->ctx, h->pdl, (fs_ekf_kernel<true, 448, 3, true>), cdiv_u(d.n, 32), (unsigned)(32 * kk), smem, d, po, h->cfg.r00, h->cfg.r11, (int)kk);
                        else          PF_LAUNCH_PDL(h->ctx, h->pdl, (fs_ekf_kernel<true, 1024, 1, true>), cdiv_u(d.n, 32), (unsigned)(32 * kk), smem, d, po, h->cfg.r00, h->cfg.r11, (int)kk);
                    } else
                    if (var == 1)      PF_LAUNCH(h->ctx, (fs_ekf_kernel<true, 448, 2>), cdiv_u(d.n, 32), (unsigned)(32 * kk), smem, d, po, h->cfg.r00, h->cfg.r11, (int)kk);
                    else if (var == 2) PF_LAUNCH(h->ctx, (fs_ekf_kernel<true, 448, 1>), cdiv_u(d.n, 32), (unsigned)(32 * kk), smem, d, po, h->cfg.r00, h->cfg.r11, (int)kk);
                    else if (var == 3) PF_LAUNCH_PDL(h->ctx, h->pdl, (fs_ekf_kernel<true, 448, 3>), cdiv_u(d.n, 32), (unsigned)(32 * kk), smem, d, po, h->cfg.r00, h->cfg.r11, (int)kk);
                    else if (var == 4) PF_LAUNCH(h->ctx, (fs_ekf_kernel<true, 448, 4>), cdiv_u(d.n, 32), (unsigned)(32 * kk), smem, d, po, h->cfg.r00, h->cfg.r11, (int)kk);
                    else               PF_LAUNCH(h->ctx, (fs_ekf_kernel<true, 1024, 1>), cdiv_u(d.n, 32), (unsigned)(32 * kk), smem, d, po, h->cfg.r00, h->cfg.r11, (int)kk);
                }
                if ((h->fused_post || h->mg_on) && seg + 2 == cuts.size() && kk <= 256) {   // last piece: the fused post kernel does the bookkeeping
                    last_po = po; last_k = (int)kk;
                    continue;
                }
            } else
            PF_LAUNCH(h->ctx, fs_step_kernel<true>, cdiv_u(d.n, FS_NT), FS_NT, (kk ? kk : 1) * sizeof(FsObsDev), d, po, u[0], u[1], h->cfg.dt,
                      sqrt(h->cfg.q00), sqrt(h->cfg.q11), h->cfg.r00, h->cfg.r11, h->seed, h->n_step, (int)kk, do_predict);
            if (kk) PF_LAUNCH(h->ctx, fs_lmstate_after_step_kernel<true>, 1, 64, 0, d, po, (int)kk);
        } else {
            // long lists go through one pinned staging slot: wait until the previous copy has left it
            PF_CUDA(cudaStreamSynchronize(h->ctx.stream));
            for (size_t j = 0; j < kk; ++j) { h->h_obs[j].d = z[j0 + j].d; h->h_obs[j].angle = z[j0 + j].angle; h->h_obs[j].lm_id = (int)z[j0 + j].lm_id; h->h_obs[j].pad = 0; }
            PF_CUDA(cudaMemcpyAsync(d.obs, h->h_obs, kk * sizeof(FsObsDev), cudaMemcpyHostToDevice, h->ctx.stream));
            PF_LAUNCH(h->ctx, fs_step_kernel<false>, cdiv_u(d.n, FS_NT), FS_NT, kk * sizeof(FsObsDev), d, po, u[0], u[1], h->cfg.dt,
                      sqrt(h->cfg.q00), sqrt(h->cfg.q11), h->cfg.r00, h->cfg.r11, h->seed, h->n_step, (int)kk, do_predict);
            PF_LAUNCH(h->ctx, fs_lmstate_after_step_kernel<false>, 1, 64, 0, d, po, (int)kk);
        }
    }
    if (h->timer.on) { PF_CUDA(cudaEventRecord(e1, h->ctx.stream)); h->timer.pending.push_back({e0, e1}); }
    h->n_step++;
    if (h->mg_on) {
        // peer-memory form (fs_mg.cuh): everything stays on the device, no NCCL call and no host sync on this path
        double nth = h->cfg.nth; uint64_t seed = h->seed; double rel = xs_margin(d.n_global);
        void* args[] = { (void*)&d, (void*)&h->fx, (void*)&h->mg, (void*)&nth, (void*)&seed, (void*)&rel, (void*)&last_po, (void*)&last_k };
        PF_CUDA(cudaLaunchCooperativeKernel((void*)fs_post_mg_kernel, dim3(h->mg.ntl), dim3(XS_NT), args, sizeof(MgShared), h->ctx.stream));
        h->ctx.launches++;
        PF_LAUNCH(h->ctx, fs_mg_search_plan_kernel, cdiv_u(d.n, 256), 256, 0, d, h->mg);
        {
            dim3 grid(cdiv_u(d.n / 4, 256), MG_IMPORT_Y + cdiv_u(d.m, FS_COMPOSE_ROWS));
            PF_LAUNCH(h->ctx, fs_mg_clone_kernel, grid, 256, 0, d, h->mg);
        }
        h->steps++;
        if (did) {
            int* hp = reinterpret_cast<int*>(h->h_pin + 32);
            PF_CUDA(cudaMemcpyAsync(hp, d.gate, sizeof(int), cudaMemcpyDeviceToHost, h->ctx.stream));
            PF_CUDA(cudaMemcpyAsync(hp + 1, h->mg.err, sizeof(int), cudaMemcpyDeviceToHost, h->ctx.stream));
            PF_CUDA(cudaStreamSynchronize(h->ctx.stream));
            *did = *hp;
            if (hp[1]) return fs_mg_error(hp[1]);
        }
        return 0;
    }
    if (h->world > 1) {
        int rcs = fs_post_sharded(h);
        if (rcs) return rcs;
        h->steps++;
        if (did) {
            int* hp = reinterpret_cast<int*>(h->h_pin + 32);
            *did = *hp;                                   // fs_post_sharded already brought the gate to the host
        }
        return 0;
    }
    if (h->fused_post) {
        // normalise, N_eff gate and (when it opens) the whole index computation + pose clone: one cooperative launch
        double nth = h->cfg.nth; uint64_t seed = h->seed; unsigned nt = h->fx_nt; double rel = xs_margin(d.n_global);
        void* args[] = { (void*)&d, (void*)&h->fx, (void*)&nth, (void*)&seed, (void*)&nt, (void*)&rel, (void*)&last_po, (void*)&last_k };
        PF_CUDA(cudaLaunchCooperativeKernel((void*)fs_post_kernel, dim3(nt), dim3(XS_NT), args, 0, h->ctx.stream));
        h->ctx.launches++;
        PF_LAUNCH_PDL(h->ctx, h->pdl, fs_search_pose_kernel, cdiv_u(d.n, 256), 256, 0, d);
        if (d.m && d.n % 4 == 0 && h->compose_vec) {
            dim3 grid(cdiv_u(d.n / 4, 256), cdiv_u(d.m, FS_COMPOSE_ROWS));
            if (d.alog) {
                if (d.anc16) PF_LAUNCH_PDL(h->ctx, h->pdl, fs_compose_flip_alog_kernel<unsigned short>, grid, 256, 0, d);
                else         PF_LAUNCH_PDL(h->ctx, h->pdl, fs_compose_flip_alog_kernel<uint32_t>, grid, 256, 0, d);
            } else {
                if (d.anc16) PF_LAUNCH_PDL(h->ctx, h->pdl, fs_compose_flip_kernel<unsigned short>, grid, 256, 0, d);
                else         PF_LAUNCH_PDL(h->ctx, h->pdl, fs_compose_flip_kernel<uint32_t>, grid, 256, 0, d);
            }
            h->steps++;
            if (did) {
                int* hp = reinterpret_cast<int*>(h->h_pin + 32);
                PF_CUDA(cudaMemcpyAsync(hp, d.gate, sizeof(int), cudaMemcpyDeviceToHost, h->ctx.stream));
                PF_CUDA(cudaStreamSynchronize(h->ctx.stream));
                *did = *hp;
            }
            return 0;
        }
        if (d.m) {      // particle counts that are not a multiple of 4 (or PFGPU_COMPOSE_VEC=0): one slot per thread, separate flip
            dim3 grid(cdiv_u(d.n, 256), cdiv_u(d.m, FS_COMPOSE_ROWS));
            if (d.anc16) PF_LAUNCH(h->ctx, fs_compose_anc_kernel<unsigned short>, grid, 256, 0, d);
            else         PF_LAUNCH(h->ctx, fs_compose_anc_kernel<uint32_t>, grid, 256, 0, d);
        }
        PF_LAUNCH(h->ctx, fs_flip_kernel, 1, 256, 0, d);
        h->steps++;
        if (did) {
            int* hp = reinterpret_cast<int*>(h->h_pin + 32);
            PF_CUDA(cudaMemcpyAsync(hp, d.gate, sizeof(int), cudaMemcpyDeviceToHost, h->ctx.stream));
            PF_CUDA(cudaStreamSynchronize(h->ctx.stream));
            *did = *hp;
        }
        return 0;
    } else {
    // normalize_weights fs1.rs:259
    int rc = xs_total(h->ctx, h->xs, XsValArray{d.w_raw}, d.n, d.n_global, 0.0, d.scal + 0);
    if (rc) return rc;
    PF_LAUNCH(h->ctx, fs_normalize_kernel, cdiv_u(d.n, 256), 256, 0, d);
    // compute_neff fs1.rs:262 and the gate fs1.rs:263
    rc = xs_total(h->ctx, h->xs, FsValWSq{d.w}, d.n, d.n_global, 0.0, d.scal + 1);
    if (rc) return rc;
    PF_LAUNCH(h->ctx, fs_gate_kernel, 1, 1, 0, d, h->cfg.nth);
    // resample fs1.rs:206-234 (every kernel below returns at once when the gate is closed)
    h->xs.gate = d.gate;
    rc = xs_total(h->ctx, h->xs, XsValArray{d.w}, d.n, d.n_global, 0.0, d.scal + 2);                 // fs1.rs:207 re-normalise
    if (!rc) rc = xs_scan(h->ctx, h->xs, FsValWNorm2{d.w, d.scal, d.gate}, XsSinkStore{d.cum}, d.n, d.n_global, 0.0, d.scal + 4);
    if (!rc) {
        PF_LAUNCH(h->ctx, fs_comb_kernel, 1, 1, 0, d, h->seed);      // the single uniform draw fs1.rs:219-220
        rc = xs_scan(h->ctx, h->xs, FsValComb{d.scal, 1.0 / (double)d.n_global}, XsSinkStore{d.rcomb}, d.n, d.n_global, 0.0, d.scal + 5);
    }
    h->xs.gate = nullptr;
    if (rc) return rc;
    PF_LAUNCH(h->ctx, fs_search_kernel, cdiv_u(d.n, 256), 256, 0, d);
    PF_LAUNCH(h->ctx, fs_gather_pose_kernel, cdiv_u(d.n, 256), 256, 0, d);
    }
    if (d.m) {
        dim3 grid(cdiv_u(d.n, 256), cdiv_u(d.m, FS_COMPOSE_ROWS));
        if (d.anc16) PF_LAUNCH(h->ctx, fs_compose_anc_kernel<unsigned short>, grid, 256, 0, d);      // lazy clone: ancestry columns instead of the map
        else         PF_LAUNCH(h->ctx, fs_compose_anc_kernel<uint32_t>, grid, 256, 0, d);
    }
    PF_LAUNCH(h->ctx, fs_flip_kernel, 1, 256, 0, d);
    h->steps++;
    if (did) {     // the gate and the resample draw counter live on the device; only a caller who asks pays a sync
        int* hp = reinterpret_cast<int*>(h->h_pin + 32);
        PF_CUDA(cudaMemcpyAsync(hp, d.gate, sizeof(int), cudaMemcpyDeviceToHost, h->ctx.stream));
        PF_CUDA(cudaStreamSynchronize(h->ctx.stream));
        *did = *hp;
    }
    return 0;
}
extern "C" int pfgpu_fs_best(pfgpu_fs* h, size_t* index, double pose_w4[4]) {
    if (!h) return PFGPU_ERR_INVALID;
    PF_CUDA(cudaSetDevice(h->ctx.device));
    FsDev& d = h->d;
    PF_LAUNCH(h->ctx, fs_best_kernel, h->best_blocks, 256, 0, d, h->best_blocks);
    double* hw = h->h_pin + 64;
    unsigned long long* hi = reinterpret_cast<unsigned long long*>(h->h_pin + 64 + h->best_blocks);
    PF_CUDA(cudaMemcpyAsync(hw, d.best_w, h->best_blocks * sizeof(double), cudaMemcpyDeviceToHost, h->ctx.stream));
    PF_CUDA(cudaMemcpyAsync(hi, d.best_i, h->best_blocks * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->ctx.stream));
    PF_CUDA(cudaStreamSynchronize(h->ctx.stream));
    size_t bi = 0; double bw = -1.0; bool have = false;
    for (int b = 0; b < h->best_blocks; ++b)
        if (!have || hw[b] > bw || (hw[b] == bw && hi[b] > bi)) { bw = hw[b]; bi = (size_t)hi[b]; have = true; }
    double pw[4] = {bw, 0.0, 0.0, 0.0};
    if (pose_w4 || h->world > 1) {
        int cur = 0;
        PF_CUDA(cudaMemcpy(&cur, d.cur, sizeof(int), cudaMemcpyDeviceToHost));
        PF_CUDA(cudaMemcpy(&pw[1], d.px[cur] + bi, sizeof(double), cudaMemcpyDeviceToHost));
        PF_CUDA(cudaMemcpy(&pw[2], d.py[cur] + bi, sizeof(double), cudaMemcpyDeviceToHost));
        PF_CUDA(cudaMemcpy(&pw[3], d.pyaw[cur] + bi, sizeof(double), cudaMemcpyDeviceToHost));
    }
    size_t gi = d.offset + bi;
    if (h->world > 1) {     // every rank contributes its shard's best; the LAST maximum in global order wins (fs1.rs:269-274)
        FsShard& sh = h->sh;
        double loc[8] = {pw[0], (double)gi, pw[1], pw[2], pw[3], 0.0, 0.0, 0.0};
        PF_CUDA(cudaMemcpyAsync(sh.best_loc, loc, sizeof(loc), cudaMemcpyHostToDevice, h->ctx.stream));
        PF_NCCL(ncclAllGather(sh.best_loc, sh.best_all, 8, ncclDouble, sh.comm, h->ctx.stream));
        std::vector<double> all(8 * (size_t)sh.world);
        PF_CUDA(cudaMemcpyAsync(all.data(), sh.best_all, all.size() * sizeof(double), cudaMemcpyDeviceToHost, h->ctx.stream));
        PF_CUDA(cudaStreamSynchronize(h->ctx.stream));
        int bg = 0;
        for (int g = 1; g < sh.world; ++g) if (all[8 * g] >= all[8 * bg]) bg = g;
        gi = (size_t)all[8 * bg + 1];
        pw[0] = all[8 * bg]; pw[1] = all[8 * bg + 2]; pw[2] = all[8 * bg + 3]; pw[3] = all[8 * bg + 4];
    }
    if (index) *index = gi;
    if (pose_w4) for (int a = 0; a < 4; ++a) pose_w4[a] = pw[a];
    return 0;
}
extern "C" int pfgpu_fs_particle_landmarks(pfgpu_fs* h, size_t il, double* lm6) {
    if (!h || !lm6 || il >= h->d.n) return PFGPU_ERR_INVALID;
    PF_CUDA(cudaSetDevice(h->ctx.device));
    FsDev& d = h->d;
    if (!d.m) return 0;
    double* tmp = nullptr;
    PF_CUDA(cudaMalloc(&tmp, d.m * 6 * sizeof(double)));
    PF_LAUNCH(h->ctx, fs_pack_lm_kernel, cdiv_u(d.m * 6, 256), 256, 0, d, tmp, il, (size_t)1);
    PF_CUDA(cudaMemcpyAsync(lm6, tmp, d.m * 6 * sizeof(double), cudaMemcpyDeviceToHost, h->ctx.stream));
    PF_CUDA(cudaStreamSynchronize(h->ctx.stream));
    cudaFree(tmp);
    return 0;
}
extern "C" int pfgpu_fs_last_indices(pfgpu_fs* h, uint32_t* idx, size_t cap, size_t* n) {
    if (!h || !idx) return PFGPU_ERR_INVALID;
    PF_CUDA(cudaSetDevice(h->ctx.device));
    size_t c = cap < h->d.n ? cap : h->d.n;
    PF_CUDA(cudaMemcpyAsync(idx, h->d.idx, c * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->ctx.stream));
    PF_CUDA(cudaStreamSynchronize(h->ctx.stream));
    if (n) *n = h->d.n;
    return 0;
}
extern "C" int pfgpu_fs_last_neff(pfgpu_fs* h, double* neff) {
    if (!h || !neff) return PFGPU_ERR_INVALID;
    PF_CUDA(cudaSetDevice(h->ctx.device));
    PF_CUDA(cudaMemcpyAsync(h->h_pin, h->d.scal + 3, sizeof(double), cudaMemcpyDeviceToHost, h->ctx.stream));
    PF_CUDA(cudaStreamSynchronize(h->ctx.stream));
    *neff = h->h_pin[0];
    return 0;
}
extern "C" int pfgpu_fs_stats(pfgpu_fs* h, pfgpu_stats* s) {
    if (!h || !s) return PFGPU_ERR_INVALID;
    PF_CUDA(cudaSetDevice(h->ctx.device));
    PF_CUDA(cudaStreamSynchronize(h->ctx.stream));
    memset(s, 0, sizeof(*s));
    timer_drain(h->timer);
    unsigned int cnt = 0;
    PF_CUDA(cudaMemcpy(&cnt, h->d.counters, sizeof(unsigned int), cudaMemcpyDeviceToHost));
    s->kernel_launches = h->ctx.launches; s->steps = h->steps; s->resamples = cnt;
    s->main_kernel_ms_sum = h->timer.ms_sum; s->main_kernel_count = h->timer.count;
    s->compactions = h->sh.compactions; s->imported_particles = h->sh.imported;
    if (h->mg_on) {
        unsigned long long pl[8];
        PF_CUDA(cudaMemcpy(pl, h->mg.plan, sizeof(pl), cudaMemcpyDeviceToHost));
        s->imported_particles = pl[4]; s->compactions = pl[5];
    }
    int rcx = read_xs_flags(h->ctx, h->xs, s);
    if (rcx) return rcx;
    return read_fx_flags(h->fx, h->fused_post || h->mg_on, s);
}
extern "C" int pfgpu_fs_shard_mode(pfgpu_fs* h, int* mode) {
    if (!h || !mode) return PFGPU_ERR_INVALID;
    *mode = h->world <= 1 ? 0 : (h->mg_on ? 2 : 1);
    return 0;
}
// debug: accumulated phase times of the fused post kernel (PFGPU_POST_TRACE=1); out32[31] = launches
extern "C" int pfgpu_fs_post_trace(pfgpu_fs* h, unsigned long long* out32) {
    if (!h || !out32) return PFGPU_ERR_INVALID;
    PF_CUDA(cudaSetDevice(h->ctx.device));
    PF_CUDA(cudaStreamSynchronize(h->ctx.stream));
    for (int k = 0; k < 32; ++k) out32[k] = 0;
    if ((h->fused_post || h->mg_on) && h->fx.dbg) PF_CUDA(cudaMemcpy(out32, h->fx.dbg, 32 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    if (h->mg_on && h->fx.dbg) PF_CUDA(cudaMemcpy(out32 + 8, h->mg.plan + 17, 6 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));   // [8..13]: stages of a resample step
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

extern "C" int pfgpu_pf_mark(pfgpu_pf* h, int slot) { if (!h) return PFGPU_ERR_INVALID; PF_CUDA(cudaSetDevice(h->ctx.device)); return marks_mark(h->ctx, h->marks, slot); }
extern "C" int pfgpu_pf_elapsed_ms(pfgpu_pf* h, int a, int b, double* ms) { if (!h) return PFGPU_ERR_INVALID; PF_CUDA(cudaSetDevice(h->ctx.device)); return marks_elapsed(h->marks, a, b, ms); }
extern "C" int pfgpu_fs_mark(pfgpu_fs* h, int slot) { if (!h) return PFGPU_ERR_INVALID; PF_CUDA(cudaSetDevice(h->ctx.device)); return marks_mark(h->ctx, h->marks, slot); }
extern "C" int pfgpu_fs_elapsed_ms(pfgpu_fs* h, int a, int b, double* ms) { if (!h) return PFGPU_ERR_INVALID; PF_CUDA(cudaSetDevice(h->ctx.device)); return marks_elapsed(h->marks, a, b, ms); }
extern "C" int pfgpu_pf_flush_l2(pfgpu_pf* h) { if (!h) return PFGPU_ERR_INVALID; PF_CUDA(cudaSetDevice(h->ctx.device)); return marks_flush(h->ctx, h->marks); }
extern "C" int pfgpu_fs_flush_l2(pfgpu_fs* h) { if (!h) return PFGPU_ERR_INVALID; PF_CUDA(cudaSetDevice(h->ctx.device)); return marks_flush(h->ctx, h->marks); }

extern "C" int pfgpu_nccl_unique_id(void* out128) {
    if (!out128) return PFGPU_ERR_INVALID;
    ncclUniqueId id;
    PF_NCCL(ncclGetUniqueId(&id));
    static_assert(sizeof(id) == 128, "ncclUniqueId is 128 bytes");
    memcpy(out128, &id, sizeof(id));
    return 0;
}

// ====================================================================================================
// test hook: the device's reciprocal-based division (PFC_DIV) against the IEEE `/` on n random + adversarial pairs
__global__ void pf_test_div_kernel(unsigned long long n, uint64_t seed, unsigned long long* mismatches) {
    unsigned long long bad = 0;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x) {
        pfc_u32x4 r = pfc_rng_block(seed, 7, 0, i), r2 = pfc_rng_block(seed, 7, 1, i);
        uint64_t x = pfc_blk_u64(r, 0), z = pfc_blk_u64(r, 1), m = pfc_blk_u64(r2, 0);
        int ea = (int)(m % 801) - 400, eb = (int)((m >> 10) % 801) - 400;
        int kind = (int)((m >> 20) & 7);
        if (kind == 0) z |= 0x000FFFFFFFFFF000ull; else if (kind == 1) z &= 0xFFF0000000000FFFull;
        else if (kind == 2) x |= 0x000FFFFFFFFFFF00ull; else if (kind == 3) { x &= 0xFFF00000000000FFull; z &= 0xFFF00000000000FFull; }
        double a = pfc_u2d((x & 0x800FFFFFFFFFFFFFull) | ((uint64_t)(ea + 1023) << 52));
        double b = pfc_u2d((z & 0x800FFFFFFFFFFFFFull) | ((uint64_t)(eb + 1023) << 52));
        if (kind == 7) a = (m & (1ull << 40)) ? 0.0 : -0.0;
        double q = PFC_DIV(a, b), t = a / b;
        if (pfc_d2u(q) != pfc_d2u(t)) bad++;
    }
    if (bad) atomicAdd(mismatches, bad);
}
extern "C" int pfgpu_test_div(unsigned long long n, uint64_t seed, unsigned long long* mismatches, int device) {
    PF_CUDA(cudaSetDevice(device));
    unsigned long long* d = nullptr;
    PF_CUDA(cudaMalloc(&d, sizeof(*d)));
    PF_CUDA(cudaMemset(d, 0, sizeof(*d)));
    pf_test_div_kernel<<<148 * 8, 256>>>(n, seed, d);
    PF_CUDA(cudaDeviceSynchronize());
    PF_CUDA(cudaMemcpy(mismatches, d, sizeof(*d), cudaMemcpyDeviceToHost));
    cudaFree(d);
    return 0;
}

// ====================================================================================================
// test hook: the exact scan on an arbitrary host array (used by tests/test_gpu_xsum.py)
// ====================================================================================================
extern "C" int pfgpu_test_xsum(const double* host_v, size_t n, double* host_scan, double* host_total, int* flags4, int device) {
    Ctx ctx;
    int rc = ctx_open(ctx, device);
    if (rc) return rc;
    XsWork xs;
    rc = xs_work_alloc(xs, n);
    if (rc) return rc;
    double *dv = nullptr, *dc = nullptr, *dt = nullptr;
    PF_CUDA(cudaMalloc(&dv, n * sizeof(double))); PF_CUDA(cudaMalloc(&dc, n * sizeof(double))); PF_CUDA(cudaMalloc(&dt, sizeof(double)));
    PF_CUDA(cudaMemcpy(dv, host_v, n * sizeof(double), cudaMemcpyHostToDevice));
    rc = xs_scan(ctx, xs, XsValArray{dv}, XsSinkStore{dc}, n, n, 0.0, dt);
    if (rc) return rc;
    PF_CUDA(cudaStreamSynchronize(ctx.stream));
    if (host_scan) PF_CUDA(cudaMemcpy(host_scan, dc, n * sizeof(double), cudaMemcpyDeviceToHost));
    if (host_total) PF_CUDA(cudaMemcpy(host_total, dt, sizeof(double), cudaMemcpyDeviceToHost));
    if (flags4) PF_CUDA(cudaMemcpy(flags4, xs.flags, 4 * sizeof(int), cudaMemcpyDeviceToHost));
    cudaFree(dv); cudaFree(dc); cudaFree(dt);
    xs_work_free(xs);
    cudaStreamDestroy(ctx.stream);
    return 0;
}
