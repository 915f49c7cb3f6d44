// pfgpu.cu — the C ABI of include/pfgpu.h: handles, step orchestration, upload/download.
// Single translation unit: nvcc -gencode arch=compute_100a,code=sm_100a --fmad=false (see __graft_entry__.build()).
#include "common.cuh"
#include "xsum.cuh"
#include "pf_kernels.cuh"
#include "pf3.cuh"
#include "pf_kld.cuh"
#include "xsum_sharded.cuh"
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
    // The fused step (pfgpu_pf_step) of a fixed-size, single-GPU filter is the same ~22 launches every time: predict + weight,
    // the exact-sum pipelines, gate, search, gather, flip, moments — a launch-bound sequence below ~10^5 particles.  It is
    // captured ONCE into a CUDA graph and replayed; only the first kernel's arguments (control, observations, draw counter,
    // range noise) change from step to step and are patched into the instantiated graph before each replay.
    // fused tail of the step (pf3.cuh): everything after the predict + likelihood kernel in one cooperative launch
    struct Fused {
        bool on = false;
        Fs3Dev x = {};             // workspace of the exact-sum routine (fs3_xsum)
        Pf3Arg arg = {};
        unsigned tiles = 0;
        size_t smem = 0;
    } fu;
    struct StepGraph {
        cudaGraphExec_t exec = nullptr;
        cudaGraph_t graph = nullptr;
        cudaGraphNode_t main_node = nullptr;
        cudaKernelNodeParams main_params = {};
        size_t k = ~(size_t)0;
        uint64_t launches = 0;
        int captures = 0;          // an observation count that keeps changing would re-capture every step: give up after a few
        bool off = false;          // PFGPU_PF_GRAPH=0, capture failed, or too many re-captures: plain launches from then on
    } sg;
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

// workspace + shape of the fused tail (pf3.cuh); leaves h->fu.on false when the configuration keeps the multi-kernel path
static int pf3_setup(pfgpu_pf* h) {
    const char* e = getenv("PFGPU_PF_FUSED");
    if (e && e[0] == '0') return 0;
    const size_t n = h->d.n;
    // measured (profiles/r02_sweep): 2.0x at 2^10, 1.4x at 2^14, 1.15x at 2^16, even at 2^18, slower at 2^20 (there the separate
    // kernels fill the GPU and their launch latency is hidden behind the graph replay)
    if (h->world != 1 || h->adaptive || n < 1 || n > ((size_t)1 << 18)) return 0;
    const unsigned NT = 256;
    unsigned tiles = (unsigned)std::min<size_t>((size_t)std::min(h->ctx.num_sms, FS3_MAX_TILES), (n + NT - 1) / NT);
    unsigned K = (unsigned)((n + (size_t)tiles * NT - 1) / ((size_t)tiles * NT));
    tiles = (unsigned)((n + (size_t)NT * K - 1) / ((size_t)NT * K));                       // no empty tile: the last one holds index n - 1
    const size_t smem = (size_t)2 * K * NT * sizeof(double);
    if (cudaFuncSetAttribute(pf3_post_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { cudaGetLastError(); return 0; }
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, pf3_post_kernel<256>, (int)NT, smem) != cudaSuccess || (size_t)nb * (size_t)h->ctx.num_sms < tiles) { cudaGetLastError(); return 0; }
    Fs3Dev& x = h->fu.x;
    x.n = (unsigned)n; x.n_glob = n; x.off = 0; x.G = 1; x.rank = 0; x.wait_inline = 1;
    x.wraw[0] = h->d.w_raw; x.wraw[1] = h->d.w_raw; x.wn_all = h->d.w;
    const size_t nsl = (size_t)FS3_SLOTS * FS3_MAX_TILES, nen = (size_t)FS3_SLOTS * FS3_ENT_CAP;
    Fs3State* stp = nullptr;
    PF_CUDA(cudaMalloc(&stp, sizeof(Fs3State))); PF_CUDA(cudaMemset(stp, 0, sizeof(Fs3State)));
    x.st = stp;
    PF_CUDA(cudaMalloc(&x.tileP, nsl * sizeof(unsigned long long)));
    PF_CUDA(cudaMalloc(&x.tileQ, FS3_MAX_TILES * sizeof(double)));
    PF_CUDA(cudaMalloc(&x.entCnt, 8 * sizeof(unsigned))); PF_CUDA(cudaMemset(x.entCnt, 0, 8 * sizeof(unsigned)));
    PF_CUDA(cudaMalloc(&x.entKey, nen * sizeof(unsigned))); PF_CUDA(cudaMalloc(&x.entTile, nen * sizeof(unsigned)));
    PF_CUDA(cudaMalloc(&x.entP, nen * sizeof(unsigned long long))); PF_CUDA(cudaMalloc(&x.entV, nen * sizeof(double)));
    PF_CUDA(cudaMalloc(&x.entL, nen * sizeof(int)));
    PF_CUDA(cudaMalloc(&x.bar, 8 * sizeof(unsigned))); PF_CUDA(cudaMemset(x.bar, 0, 8 * sizeof(unsigned)));
    PF_CUDA(cudaMalloc(&x.resflag, 8 * sizeof(unsigned))); PF_CUDA(cudaMemset(x.resflag, 0, 8 * sizeof(unsigned)));
    PF_CUDA(cudaMalloc(&x.res, 8 * sizeof(Fs3Res))); PF_CUDA(cudaMemset(x.res, 0, 8 * sizeof(Fs3Res)));
    PF_CUDA(cudaMalloc(&x.resTP, (size_t)8 * FS3_MAX_TILES * sizeof(unsigned long long)));
    PF_CUDA(cudaMalloc(&x.resKey, (size_t)8 * FS3_ENT_CAP * sizeof(unsigned))); PF_CUDA(cudaMalloc(&x.resP, (size_t)8 * FS3_ENT_CAP * sizeof(unsigned long long)));
    PF_CUDA(cudaMalloc(&x.resAft, (size_t)8 * FS3_ENT_CAP * sizeof(double)));
    PF_CUDA(cudaMalloc(&x.tileEnd, FS3_MAX_TILES * sizeof(double)));
    PF_CUDA(cudaMalloc(&x.flagsg, 8 * sizeof(int))); PF_CUDA(cudaMemset(x.flagsg, 0, 8 * sizeof(int)));
    Pf3Arg& a = h->fu.arg;
    a.pd = h->d; a.threshold = h->cfg.resample_threshold; a.mode = h->cfg.mode; a.seed = h->seed; a.K = K; a.m32 = x3_margin32(n);
    PF_CUDA(cudaMalloc(&a.tsum, FS3_MAX_TILES * sizeof(double))); PF_CUDA(cudaMalloc(&a.tsq, FS3_MAX_TILES * sizeof(double)));
    PF_CUDA(cudaMalloc(&a.mom, (size_t)FS3_MAX_TILES * PF_MOM * sizeof(double)));
    h->fu.tiles = tiles; h->fu.smem = smem;
    h->fu.on = true;
    return 0;
}
static void pf3_free(pfgpu_pf* h) {
    Fs3Dev& x = h->fu.x;
    cudaFree(x.st); cudaFree(x.tileP); cudaFree(x.tileQ); cudaFree(x.entCnt); cudaFree(x.entKey); cudaFree(x.entTile); cudaFree(x.entP);
    cudaFree(x.entV); cudaFree(x.entL); cudaFree(x.bar); cudaFree(x.resflag); cudaFree(x.res); cudaFree(x.resTP); cudaFree(x.resKey);
    cudaFree(x.resP); cudaFree(x.resAft); cudaFree(x.tileEnd); cudaFree(x.flagsg);
    cudaFree(h->fu.arg.tsum); cudaFree(h->fu.arg.tsq); cudaFree(h->fu.arg.mom);
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
    { const char* e = getenv("PFGPU_PF_GRAPH"); if (e && e[0] == '0') h->sg.off = true; }
    h->fu.on = false;
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
    rc = pf3_setup(h);
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
    if (h->sg.exec) cudaGraphExecDestroy(h->sg.exec);
    if (h->sg.graph) cudaGraphDestroy(h->sg.graph);
    pf3_free(h);
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
// device temporary of the bulk transfers: released on every exit path (PF_CUDA / PF_LAUNCH return early on errors)
struct PfScopedDev {
    double* p = nullptr;
    ~PfScopedDev() { if (p) cudaFree(p); }
};
extern "C" int pfgpu_pf_upload(pfgpu_pf* h, const double* aos5, size_t n) {
    if (!h || !aos5 || n != h->d.n) return PFGPU_ERR_INVALID;
    PF_CUDA(cudaSetDevice(h->ctx.device));
    PfScopedDev t;
    PF_CUDA(cudaMalloc(&t.p, n * 5 * sizeof(double)));
    double* tmp = t.p;
    PF_CUDA(cudaMemcpyAsync(tmp, aos5, n * 5 * sizeof(double), cudaMemcpyHostToDevice, h->ctx.stream));
    PF_LAUNCH(h->ctx, pf_unpack_kernel, cdiv_u(n, PF_NT), PF_NT, 0, h->d, tmp);
    int rc = pf_refresh_cache(h);
    PF_CUDA(cudaStreamSynchronize(h->ctx.stream));
    return rc;
}
extern "C" int pfgpu_pf_download(pfgpu_pf* h, double* aos5, size_t n) {
    if (!h || !aos5 || n != h->d.n) return PFGPU_ERR_INVALID;
    PF_CUDA(cudaSetDevice(h->ctx.device));
    PfScopedDev t;
    PF_CUDA(cudaMalloc(&t.p, n * 5 * sizeof(double)));
    double* tmp = t.p;
    PF_LAUNCH(h->ctx, pf_pack_kernel, cdiv_u(n, PF_NT), PF_NT, 0, h->d, tmp);
    PF_CUDA(cudaMemcpyAsync(aos5, tmp, n * 5 * sizeof(double), cudaMemcpyDeviceToHost, h->ctx.stream));
    PF_CUDA(cudaStreamSynchronize(h->ctx.stream));
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
// the launches of one fused step, in stream order (what the graph captures)
static int pf_step_launches(pfgpu_pf* h, const double u[2], const double* obs3, size_t k) {
    int rc = pf_launch_main<true, true>(h, u, obs3, k);                              // predict + likelihood, one pass
    if (rc) return rc;
    if (h->fu.on) {                                                                  // normalise .. refresh_cache: one launch (pf3.cuh)
        h->fu.arg.pd = h->d;
        h->fu.arg.threshold = h->cfg.resample_threshold;
        PF_LAUNCH(h->ctx, pf3_post_kernel<256>, h->fu.tiles, 256, h->fu.smem, h->fu.x, h->fu.arg);
        return 0;
    }
    rc = pf_normalize(h);
    if (rc) return rc;
    rc = pf_resample_impl(h);
    if (rc) return rc;
    return pf_refresh_cache(h);
}
static void pf_graph_drop(pfgpu_pf* h) {
    if (h->sg.exec) cudaGraphExecDestroy(h->sg.exec);
    if (h->sg.graph) cudaGraphDestroy(h->sg.graph);
    h->sg.exec = nullptr; h->sg.graph = nullptr; h->sg.main_node = nullptr; h->sg.k = ~(size_t)0;
}
// capture the step at observation count k (no work is executed by the capture itself)
static int pf_graph_capture(pfgpu_pf* h, const double u[2], const double* obs3, size_t k) {
    pf_graph_drop(h);
    const uint64_t l0 = h->ctx.launches;
    if (cudaStreamBeginCapture(h->ctx.stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { cudaGetLastError(); return 1; }
    const int rc = pf_step_launches(h, u, obs3, k);
    cudaGraph_t g = nullptr;
    const cudaError_t e = cudaStreamEndCapture(h->ctx.stream, &g);
    h->sg.launches = h->ctx.launches - l0;
    h->ctx.launches = l0;
    if (rc || e != cudaSuccess || !g) { cudaGetLastError(); if (g) cudaGraphDestroy(g); return 1; }
    h->sg.graph = g;
    size_t nn = 0;
    if (cudaGraphGetNodes(g, nullptr, &nn) != cudaSuccess || nn == 0) { cudaGetLastError(); pf_graph_drop(h); return 1; }
    std::vector<cudaGraphNode_t> nodes(nn);
    if (cudaGraphGetNodes(g, nodes.data(), &nn) != cudaSuccess) { cudaGetLastError(); pf_graph_drop(h); return 1; }
    const void* want = (const void*)pf_predict_weight_kernel<true, true, true>;
    for (cudaGraphNode_t nd : nodes) {
        cudaGraphNodeType ty;
        if (cudaGraphNodeGetType(nd, &ty) != cudaSuccess || ty != cudaGraphNodeTypeKernel) continue;
        cudaKernelNodeParams kp = {};
        if (cudaGraphKernelNodeGetParams(nd, &kp) != cudaSuccess) continue;
        if (kp.func == want) { h->sg.main_node = nd; h->sg.main_params = kp; break; }
    }
    if (!h->sg.main_node || cudaGraphInstantiate(&h->sg.exec, g, 0) != cudaSuccess) { cudaGetLastError(); pf_graph_drop(h); return 1; }
    h->sg.k = k;
    return 0;
}
// replay with this step's arguments patched into the first kernel
static int pf_graph_replay(pfgpu_pf* h, const double u[2], const double* obs3, size_t k) {
    PfObsParam po;
    for (size_t j = 0; j < 3 * k; ++j) po.o[j] = obs3[j];
    double u0 = u[0], u1 = u[1], sv = h->cfg.velocity_noise, sw = h->cfg.yaw_rate_noise, dt = h->cfg.dt, sigma = h->cfg.range_noise;
    uint64_t seed = h->seed; uint32_t call = h->n_predict; int kk = (int)k;
    void* args[] = { &h->d, &po, &u0, &u1, &sv, &sw, &dt, &seed, &call, &kk, &sigma };
    cudaKernelNodeParams kp = h->sg.main_params;
    kp.kernelParams = args; kp.extra = nullptr;
    PF_CUDA(cudaGraphExecKernelNodeSetParams(h->sg.exec, h->sg.main_node, &kp));
    PF_CUDA(cudaGraphLaunch(h->sg.exec, h->ctx.stream));
    h->ctx.launches += h->sg.launches;
    return 0;
}

extern "C" int pfgpu_pf_step(pfgpu_pf* h, const double u[2], const double* obs3, size_t k, double est[4]) {
    if (!h || !u || (k && !obs3)) return PFGPU_ERR_INVALID;
    if (!finite_d(u[0]) || !finite_d(u[1])) return PFGPU_ERR_INVALID;
    PF_CUDA(cudaSetDevice(h->ctx.device));
    int rc = pf_stage_obs(h, obs3, k);
    if (rc) return rc;
    // graph replay: fixed particle count, one GPU, observations short enough to ride in the launch parameters, no per-kernel
    // timing events; the first step of a handle runs plainly (lazy set-up such as function attributes happens there)
    const bool graphable = !h->sg.off && h->world == 1 && !h->adaptive && k <= PF_PARAM_OBS && !h->timer.on && h->steps > 0;
    bool done = false;
    if (graphable) {
        if (h->sg.exec && h->sg.k == k) done = true;
        else if (++h->sg.captures <= 16 && pf_graph_capture(h, u, obs3, k) == 0) done = true;
        else { h->sg.off = true; pf_graph_drop(h); }
        if (done) { rc = pf_graph_replay(h, u, obs3, k); if (rc) return rc; }
    }
    if (!done) { rc = pf_step_launches(h, u, obs3, k); if (rc) return rc; }
    h->n_predict++;
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
    int gate = 0;
    { int rc = pf_read_gate(h, &gate); if (rc) return rc; }
    if (!gate) { if (n) *n = 0; return 0; }                   // the last step did not resample: no ancestry (as the oracle reports)
    size_t c = cap < h->d.n ? cap : h->d.n;
    PF_CUDA(cudaMemcpyAsync(idx, h->d.idx, c * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->ctx.stream));
    PF_CUDA(cudaStreamSynchronize(h->ctx.stream));
    if (n) *n = c;
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
// FastSLAM 1.0: fs3_host.cuh (entry points) + fs3.cuh (kernels)
// ====================================================================================================
#include "fs3_host.cuh"

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
