// fs_kernels.cuh — FastSLAM 1.0 device kernels.  Reference: fs1.rs = crates/rust_robotics_slam/src/fastslam1.rs.
//
// HBM layout (per shard of n particles, m landmarks), structure of arrays:
//   px[2][n], py[2][n], pyaw[2][n]     pose columns (ping-pong; device int `cur` selects the live set)
//   w[n]                               persistent particle weight (fs1.rs:46)
//   w_raw[n]                           weight after this step's likelihood products, before normalisation
//   lm[2][m][6][n]                     landmark EKF state: field-major columns x, y, c00, c01, c10, c11 of
//                                      landmark l are 6 contiguous n-vectors, so one observed landmark is 6
//                                      fully coalesced column segments for a warp of consecutive particles
//   anc[2][m][n] (u32), lmstate[m]     LAZY CLONE.  The reference's resample deep-copies every particle's whole map
//                                      (particles[j].clone(), fs1.rs:227: 48*m bytes per particle, 1.6 GB of traffic
//                                      at 65 536 x 256).  Here a resample only composes, per landmark, an ancestry
//                                      column: "particle i's copy of landmark l lives in column anc[l][i] of buffer
//                                      lmstate[l]&1" (8 B per particle and landmark instead of 96).  The actual copy is
//                                      folded into the next EKF update of that landmark: it reads through anc (systematic
//                                      ancestries are monotone, so the gather stays nearly coalesced), writes column i of
//                                      the other buffer and marks the landmark "identity" (lmstate[l]&2) again.
//   cum[n], rcomb[n], idx[n]           exact cumulative weights, exact comb positions, ancestry
#pragma once
#include "common.cuh"
#include "xsum.cuh"

#define FS_NT 128
#define FS_MAX_OBS 1024

struct FsObsDev { double d, angle; int lm_id; int pad; };

struct FsDev {
    size_t n = 0, n_global = 0, offset = 0, m = 0;
    double* px[2] = {nullptr, nullptr};
    double* py[2] = {nullptr, nullptr};
    double* pyaw[2] = {nullptr, nullptr};
    double* lm[2] = {nullptr, nullptr};
    int* cur = nullptr;
    double* w = nullptr;
    double* w_raw = nullptr;
    double* cum = nullptr;
    double* rcomb = nullptr;
    uint32_t* idx = nullptr;
    double* scal = nullptr;      // [0] S  [1] Q  [2] S2 (resample re-normalisation)  [3] neff  [4] cum total  [5] comb total
    int* gate = nullptr;
    FsObsDev* obs = nullptr;     // device copy of this step's observation list
    double* best_w = nullptr; unsigned long long* best_i = nullptr;   // per-block argmax partials
    unsigned int* counters = nullptr;   // device: [0] resamples done so far (= Philox call index of the next comb draw)
    uint32_t* anc[2] = {nullptr, nullptr};   // [m][n] per-landmark ancestry columns (ping-pong across resamples)
    int* anc_cur = nullptr;                  // device: live ancestry buffer
    int* lmstate = nullptr;                  // device [m]: bit0 = buffer holding landmark l, bit1 = ancestry is the identity
    int eager = 0;                           // 1 (sharded mode): maps are cloned eagerly; every landmark lives in buffer *cur
    int anc16 = 0;                           // 1: ancestry columns hold u16 (n <= 65 536)
    size_t ld = 0;                           // column stride of the landmark arrays: n, plus the GUEST columns of sharded mode
                                             // (maps imported from other ranks live in columns [n, ld) until the next compaction)
    // ANCESTRY LOG (PFGPU_ANC_LOG=1, one GPU; model + rules in tests/test_anclog_model.py): a resample does not compose the
    // ancestry rows at all, it appends its index array to a ring of `alog` generations; row l (explicit or identity) is valid
    // for generation gen[l], and a reader at generation G = counters[0] first walks its slot back to gen[l] through the logged
    // maps.  Every `alog` resamples all rows are recomposed to the current generation, so no walk is longer than the ring.
    int alog = 0;                            // 0: off (every resample composes every row); else the ring size R
    uint32_t* idxlog = nullptr;              // [R][n] index arrays of the last R resamples; slot g % R maps generation g+1 -> g
    int* gen = nullptr;                      // [m] generation for which row l / its identity flag is valid
};

// Observation lists of up to FS_PARAM_OBS entries travel inside the kernel's launch parameters (no H2D copy).
#define FS_PARAM_OBS 48
struct FsObsParam { FsObsDev o[FS_PARAM_OBS]; };

// ping-pong buffer selection by ternary (a runtime index into the by-value parameter struct would force a stack copy)
__device__ __forceinline__ double* fs_px(const FsDev& d, int c) { return c ? d.px[1] : d.px[0]; }
__device__ __forceinline__ double* fs_py(const FsDev& d, int c) { return c ? d.py[1] : d.py[0]; }
__device__ __forceinline__ double* fs_pyaw(const FsDev& d, int c) { return c ? d.pyaw[1] : d.pyaw[0]; }
__device__ __forceinline__ double* fs_lm(const FsDev& d, int c) { return c ? d.lm[1] : d.lm[0]; }
__device__ __forceinline__ uint32_t* fs_anc(const FsDev& d, int c) { return c ? d.anc[1] : d.anc[0]; }
// ancestry columns are stored as u16 when every index fits (n <= 65 536: half the traffic of a resample), else u32
__device__ __forceinline__ size_t fs_anc_load(const FsDev& d, int c, size_t k) {
    const uint32_t* p = fs_anc(d, c);
    return d.anc16 ? (size_t)reinterpret_cast<const unsigned short*>(p)[k] : (size_t)p[k];
}
__device__ __forceinline__ size_t lm_index(size_t n, size_t l, int f, size_t i) { return (l * 6 + (size_t)f) * n + i; }
// ancestry log: slot i of generation G -> the slot it descends from in generation gen[l]
template <bool ALOG>
__device__ __forceinline__ size_t fs_walk_back(const FsDev& d, size_t l, size_t i, unsigned G) {
    if (!ALOG) return i;
    size_t j = i;
    const unsigned R = (unsigned)d.alog, gl = (unsigned)d.gen[l];
    for (unsigned g = G; g != gl; --g) j = d.idxlog[(size_t)((g - 1u) % R) * d.n + j];
    return j;
}
// landmark l may be updated in place: its row is the identity AND belongs to the current generation
template <bool ALOG>
__device__ __forceinline__ bool fs_row_fresh(const FsDev& d, size_t l, int st, unsigned G) {
    return (st & 2) != 0 && (!ALOG || (unsigned)d.gen[l] == G);
}
// column that holds landmark l of the particle in slot i (current generation)
template <bool ALOG>
__device__ __forceinline__ size_t fs_lm_col(const FsDev& d, size_t l, size_t i, int st, unsigned G) {
    const size_t j = fs_walk_back<ALOG>(d, l, i, G);
    return (st & 2) ? j : fs_anc_load(d, *d.anc_cur, l * d.n + j);
}
// bookkeeping after an EKF launch updated landmark l: it now lives in the other buffer, own columns, current generation
__device__ __forceinline__ void fs_mark_updated(const FsDev& d, int l) {
    const int st = d.lmstate[l];
    const unsigned G = d.counters[0];
    const bool fresh = d.alog ? fs_row_fresh<true>(d, (size_t)l, st, G) : fs_row_fresh<false>(d, (size_t)l, st, G);
    if (!fresh) { d.lmstate[l] = ((st & 1) ^ 1) | 2; if (d.alog) d.gen[l] = (int)G; }
}

// normalize_angle fs1.rs:80-89.  The reference loops without bound (and would spin forever on +-inf); the
// device caps the loop at 2^22 turns, i.e. |angle| up to ~2.6e7 rad behaves exactly like the reference.
__device__ __forceinline__ double fs_normalize_angle(double a) {
    int guard = 0;
    while (a > PFC_PI && guard < (1 << 22)) { a -= 2.0 * PFC_PI; ++guard; }
    while (a < -PFC_PI && guard < (1 << 23)) { a += 2.0 * PFC_PI; ++guard; }
    return a;
}

struct FsLm { double x, y, c00, c01, c10, c11; };

// update_landmark fs1.rs:140-183 for one (particle, observation) pair; returns the likelihood factor
// (1.0 when the weight is left untouched).  L is updated in registers; *wrote_cov tells the caller whether
// the covariance changed (branch A leaves it alone, fs1.rs:144-149).
__device__ __forceinline__ double fs_update_landmark(FsLm& L, double px, double py, double pyaw, double z0, double z1,
                                                     double r00, double r11, bool* wrote_cov) {
    if (L.c00 > 100.0) {                                       // first observation of this landmark
        double s, c;
        pfc_sincos(pyaw + z1, &s, &c);
        L.x = px + z0 * c;
        L.y = py + z0 * s;
        *wrote_cov = false;
        return 1.0;
    }
    *wrote_cov = true;
    // observation_model fs1.rs:92-99
    double dx = L.x - px, dy = L.y - py;
    double d2 = dx * dx + dy * dy;
    double d = sqrt(d2);
    double zp1 = fs_normalize_angle(pfc_atan2(dy, dx) - pyaw);
    double y0 = z0 - d, y1 = fs_normalize_angle(z1 - zp1);     // innovation fs1.rs:155
    // compute_jacobian fs1.rs:102-110
    // four IEEE quotients over two denominators: one correctly rounded reciprocal each (pf_contract_math.h, PFC_DIV)
    const pfc_rcp_t rd = pfc_rcp_make(d), rd2 = pfc_rcp_make(d2);
    double h00 = pfc_div_by(dx, rd), h01 = pfc_div_by(dy, rd), h10 = pfc_div_by(-dy, rd2), h11 = pfc_div_by(dx, rd2);
    double p00 = L.c00, p01 = L.c01, p10 = L.c10, p11 = L.c11;
    // S = H P H^T + R  fs1.rs:161
    double a00 = h00 * p00 + h01 * p10, a01 = h00 * p01 + h01 * p11;
    double a10 = h10 * p00 + h11 * p10, a11 = h10 * p01 + h11 * p11;
    double s00 = (a00 * h00 + a01 * h01) + r00;
    double s01 = (a00 * h10 + a01 * h11) + 0.0;
    double s10 = (a10 * h00 + a11 * h01) + 0.0;
    double s11 = (a10 * h10 + a11 * h11) + r11;
    // try_inverse().unwrap_or(identity) fs1.rs:164
    double det = s00 * s11 - s10 * s01;
    double i00, i01, i10, i11;
    if (det == 0.0) { i00 = 1.0; i01 = 0.0; i10 = 0.0; i11 = 1.0; }
    else {
        const pfc_rcp_t rdet = pfc_rcp_make(det);
        i00 = pfc_div_by(s11, rdet); i01 = pfc_div_by(-s01, rdet); i10 = pfc_div_by(-s10, rdet); i11 = pfc_div_by(s00, rdet);
    }
    // K = P H^T S^-1 fs1.rs:165
    double b00 = p00 * h00 + p01 * h01, b01 = p00 * h10 + p01 * h11;
    double b10 = p10 * h00 + p11 * h01, b11 = p10 * h10 + p11 * h11;
    double k00 = b00 * i00 + b01 * i10, k01 = b00 * i01 + b01 * i11;
    double k10 = b10 * i00 + b11 * i10, k11 = b10 * i01 + b11 * i11;
    L.x = L.x + (k00 * y0 + k01 * y1);                         // fs1.rs:168-170
    L.y = L.y + (k10 * y0 + k11 * y1);
    // P = (I - K H) P fs1.rs:173-174 (not symmetrised)
    double m00 = 1.0 - (k00 * h00 + k01 * h10), m01 = 0.0 - (k00 * h01 + k01 * h11);
    double m10 = 0.0 - (k10 * h00 + k11 * h10), m11 = 1.0 - (k10 * h01 + k11 * h11);
    L.c00 = m00 * p00 + m01 * p10; L.c01 = m00 * p01 + m01 * p11;
    L.c10 = m10 * p00 + m11 * p10; L.c11 = m10 * p01 + m11 * p11;
    // likelihood fs1.rs:177-182
    double det_s = s00 * s11 - s10 * s01;
    if (det_s > 0.0) {
        double t0 = y0 * i00 + y1 * i10, t1 = y0 * i01 + y1 * i11;
        double mahal = t0 * y0 + t1 * y1;
        return PFC_DIV(pfc_exp(-0.5 * mahal), 2.0 * PFC_PI * sqrt(det_s));
    }
    return 1.0;
}

// fastslam_update fs1.rs:245-256: predict_particle (fs1.rs:123-137) then, per observation in list order,
// update_landmark.  One thread = one particle; the observation list is staged in shared memory; each
// observed landmark is 6 coalesced column loads and 2 or 6 column stores.
template <bool PARAM_OBS>
__global__ void __launch_bounds__(FS_NT) fs_step_kernel(FsDev d, const __grid_constant__ FsObsParam po, double u0, double u1,
                                                        double dt, double sq0, double sq1, double r00, double r11,
                                                        uint64_t seed, uint32_t call, int k_obs, int do_predict) {
    extern __shared__ FsObsDev s_obs_fs[];
    for (int j = threadIdx.x; j < k_obs; j += FS_NT) s_obs_fs[j] = PARAM_OBS ? po.o[j] : d.obs[j];
    __syncthreads();
    const size_t i = (size_t)blockIdx.x * FS_NT + threadIdx.x;
    if (i >= d.n) return;
    const int cur = *d.cur;
    double px = fs_px(d, cur)[i], py = fs_py(d, cur)[i], pyaw = fs_pyaw(d, cur)[i];
    double w;
    if (do_predict) {   // predict_particle + motion_model fs1.rs:70-77,128-136
        double z0, z1;
        pfc_normal_pair(pfc_rng_block(seed, PFC_STREAM_FS_PREDICT, call, d.offset + i), &z0, &z1);
        double un0 = u0 + z0 * sq0;
        double un1 = u1 + z1 * sq1;
        double s, c;
        pfc_sincos(pyaw, &s, &c);
        double nx = px + un0 * dt * c;
        double ny = py + un0 * dt * s;
        double nyaw = fs_normalize_angle(pyaw + un1 * dt);
        px = nx; py = ny; pyaw = nyaw;
        w = d.w[i];
    } else {
        w = d.w_raw[i];       // continuation launch of the same step (observation list split at a repeated lm_id)
    }
    const size_t n = d.n;
    const int anc_c = *d.anc_cur;
    for (int j = 0; j < k_obs; ++j) {
        const size_t l = (size_t)s_obs_fs[j].lm_id;
        const double zd = s_obs_fs[j].d, za = s_obs_fs[j].angle;
        const int st = d.lmstate[l];
        const unsigned G = d.counters[0];
        const bool ident = d.alog ? fs_row_fresh<true>(d, l, st, G) : (st & 2) != 0;     // in place only if the row is fresh
        const double* __restrict__ src = fs_lm(d, st & 1);
        double* __restrict__ dst = fs_lm(d, ident ? (st & 1) : ((st & 1) ^ 1));
        const size_t col = ident ? i : (d.alog ? fs_lm_col<true>(d, l, i, st, G) : fs_anc_load(d, anc_c, l * n + i));   // lazy clone: read the ancestor's copy
        FsLm L;
        L.x = src[lm_index(d.ld, l, 0, col)]; L.y = src[lm_index(d.ld, l, 1, col)];
        L.c00 = src[lm_index(d.ld, l, 2, col)]; L.c01 = src[lm_index(d.ld, l, 3, col)];
        L.c10 = src[lm_index(d.ld, l, 4, col)]; L.c11 = src[lm_index(d.ld, l, 5, col)];
        bool wrote_cov;
        double lik = fs_update_landmark(L, px, py, pyaw, zd, za, r00, r11, &wrote_cov);
        dst[lm_index(d.ld, l, 0, i)] = L.x; dst[lm_index(d.ld, l, 1, i)] = L.y;
        if (wrote_cov || !ident) {                             // a materialising write must carry the covariance too
            dst[lm_index(d.ld, l, 2, i)] = L.c00; dst[lm_index(d.ld, l, 3, i)] = L.c01;
            dst[lm_index(d.ld, l, 4, i)] = L.c10; dst[lm_index(d.ld, l, 5, i)] = L.c11;
        }
        if (wrote_cov) w = w * lik;                            // fs1.rs:181 (only when det_s > 0: lik == 1.0 otherwise)
    }
    fs_px(d, cur)[i] = px; fs_py(d, cur)[i] = py; fs_pyaw(d, cur)[i] = pyaw;
    d.w_raw[i] = w;
}
// ---------------------------------------------------------------------------------------------------------------------
// Observation-parallel form of the step: fs_predict_kernel + fs_ekf_kernel.  Same arithmetic, decomposed for latency:
// the one-thread-per-particle kernel has only n/32 = 2048 warps for the whole chip (14 per SM), each carrying ~7 700
// dependent instructions, so it idles on FP64 div/sqrt and load latency (profiles/r01a_summary.md).
//   fs_predict_kernel : predict_particle (fs1.rs:123-137), one thread per particle, in place; also w_raw = w.
//   fs_ekf_kernel     : a CTA owns 32 particles and runs one warp per observation (lanes = particles, so the six
//                       landmark columns stay coalesced); every warp does one update_landmark (fs1.rs:140-183); then
//                       warp 0 multiplies the likelihood factors into the weight IN OBSERVATION ORDER,
//                       w = (((w*l_0)*l_1)...), exactly like the sequential loop fs1.rs:250-256.
// Updates of different landmarks commute (a launch never holds the same lm_id twice; the host splits such lists).
// ---------------------------------------------------------------------------------------------------------------------
#define FS2_MAX_OBS 32
__global__ void __launch_bounds__(256) fs_predict_kernel(FsDev d, double u0, double u1, double dt, double sq0, double sq1,
                                                         uint64_t seed, uint32_t call) {
    pf_grid_dep_sync();

    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= d.n) return;
    const int cur = *d.cur;
    double px = fs_px(d, cur)[i], py = fs_py(d, cur)[i], pyaw = fs_pyaw(d, cur)[i];
    double z0, z1;
    pfc_normal_pair(pfc_rng_block(seed, PFC_STREAM_FS_PREDICT, call, d.offset + i), &z0, &z1);
    double un0 = u0 + z0 * sq0;                                // fs1.rs:129
    double un1 = u1 + z1 * sq1;                                // fs1.rs:130
    double s, c;
    pfc_sincos(pyaw, &s, &c);
    fs_px(d, cur)[i] = px + un0 * dt * c;                      // motion_model fs1.rs:73-75
    fs_py(d, cur)[i] = py + un0 * dt * s;
    fs_pyaw(d, cur)[i] = fs_normalize_angle(pyaw + un1 * dt);
    d.w_raw[i] = d.w[i];
}

// MAXT/MINB: launch bounds = register budget.  (1024,1): 64 registers, any k_obs <= 32; (448,2): 72 registers, two CTAs of
// <= 14 warps per SM; (448,1): up to 128 registers, one CTA per SM.  Chosen at run time by the host (PFGPU_EKF_VARIANT).
template <bool PARAM_OBS, int MAXT, int MINB, bool ALOG = false>
__global__ void __launch_bounds__(MAXT, MINB) fs_ekf_kernel(FsDev d, const __grid_constant__ FsObsParam po, double r00, double r11, int k_obs) {
    pf_grid_dep_sync();

    extern __shared__ double s_v2[];
    double* s_lik = s_v2;                                                          // [k_obs][32]
    unsigned* s_mask = reinterpret_cast<unsigned*>(s_lik + (size_t)k_obs * 32);    // [k_obs]
    const int lane = threadIdx.x & 31, wj = threadIdx.x >> 5;
    const size_t i = (size_t)blockIdx.x * 32 + lane;
    const bool valid = i < d.n;
    const size_t n = d.n;
    const int cur = *d.cur;
    const FsObsDev ob = PARAM_OBS ? po.o[wj] : d.obs[wj];
    const size_t l = (size_t)ob.lm_id;
    const int st = d.lmstate[l];
    const unsigned G = ALOG ? d.counters[0] : 0u;                                  // ancestry log: current generation
    const bool ident = fs_row_fresh<ALOG>(d, l, st, G);                            // update in place (own column, same buffer)
    const size_t ld = d.ld;
    const double* __restrict__ src = fs_lm(d, st & 1) + l * 6 * ld;
    double* __restrict__ dst = fs_lm(d, ident ? (st & 1) : ((st & 1) ^ 1)) + l * 6 * ld + i;
    bool wrote_cov = false;
    double lik = 1.0;
    if (valid) {
        const size_t col = ident ? i : (ALOG ? fs_lm_col<ALOG>(d, l, i, st, G) : fs_anc_load(d, *d.anc_cur, l * n + i));     // lazy clone: the ancestor's copy
        const double* __restrict__ sp = src + col;
        FsLm L;
        L.x = sp[0]; L.y = sp[ld]; L.c00 = sp[2 * ld]; L.c01 = sp[3 * ld]; L.c10 = sp[4 * ld]; L.c11 = sp[5 * ld];
        const double px = fs_px(d, cur)[i], py = fs_py(d, cur)[i], pyaw = fs_pyaw(d, cur)[i];
        lik = fs_update_landmark(L, px, py, pyaw, ob.d, ob.angle, r00, r11, &wrote_cov);
        dst[0] = L.x; dst[ld] = L.y;
        if (wrote_cov || !ident) { dst[2 * ld] = L.c00; dst[3 * ld] = L.c01; dst[4 * ld] = L.c10; dst[5 * ld] = L.c11; }
    }
    s_lik[wj * 32 + lane] = lik;
    const unsigned mask = __ballot_sync(0xffffffffu, wrote_cov);
    if (lane == 0) s_mask[wj] = mask;
    __syncthreads();
    if (wj == 0 && valid) {
        double w = d.w_raw[i];
        for (int j = 0; j < k_obs; ++j)
            if ((s_mask[j] >> lane) & 1u) w = w * s_lik[j * 32 + lane];              // fs1.rs:181, in observation order
        d.w_raw[i] = w;
    }
}

// after a step launch: every landmark it updated through its ancestry now lives in the other buffer, own columns
template <bool PARAM_OBS>
__global__ void fs_lmstate_after_step_kernel(FsDev d, const __grid_constant__ FsObsParam po, int k_obs) {
    for (int j = threadIdx.x; j < k_obs; j += blockDim.x) {
        int l = PARAM_OBS ? po.o[j].lm_id : d.obs[j].lm_id;
        fs_mark_updated(d, l);
    }
}

// normalize_weights fs1.rs:196-203 (no uniform fallback)
__global__ void __launch_bounds__(256) fs_normalize_kernel(FsDev d) {
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= d.n) return;
    const double S = d.scal[0];
    double x = d.w_raw[i];
    d.w[i] = S > 0.0 ? x / S : x;
}
struct FsValWSq { const double* w; __device__ __forceinline__ double operator()(size_t i) const { double x = w[i]; return x * x; } };

// compute_neff + gate fs1.rs:262-265
__global__ void fs_gate_kernel(FsDev d, double nth) {
    double Q = d.scal[1];
    double neff = Q > 0.0 ? 1.0 / Q : 0.0;
    d.scal[3] = neff;
    *d.gate = neff < nth ? 1 : 0;
}

// resample() re-normalises first (fs1.rs:207): values w_i / S2 (or w_i when S2 == 0) are recomputed on the fly
struct FsValWNorm2 {
    const double* w; const double* scal; const int* gate;
    __device__ __forceinline__ double operator()(size_t i) const { double S2 = scal[2]; double x = w[i]; return S2 > 0.0 ? x / S2 : x; }
};
// the comb r, r + 1/n, ... accumulated sequentially (fs1.rs:219-230)
struct FsValComb { const double* scal; double inv; __device__ __forceinline__ double operator()(size_t i) const { return i == 0 ? scal[6] : inv; } };
// let r = Uniform::new(0, 1/n).sample(rng)  (fs1.rs:219-220; rand 0.9: u01 * scale + low), drawn once per resample
__global__ void fs_comb_kernel(FsDev d, uint64_t seed) {
    if (!*d.gate) return;
    double inv = 1.0 / (double)d.n_global;
    double u01 = pfc_u01_52(pfc_blk_u64(pfc_rng_block(seed, PFC_STREAM_FS_RESAMPLE, d.counters[0], 0), 0));
    d.scal[6] = u01 * (inv - 0.0) + 0.0;
}

// while r > cum_sum[j+1] && j < n-1 { j += 1 }  (fs1.rs:224-226): r and j are both non-decreasing, so slot t's
// j is the first j with cum_incl[j] >= r_t, clamped to n-1.
__global__ void __launch_bounds__(256) fs_search_kernel(FsDev d) {
    if (!*d.gate) return;
    const size_t t = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (t >= d.n) return;
    const double r = d.rcomb[t];
    const double* __restrict__ c = d.cum;
    size_t lo = 0, hi = d.n;
    while (lo < hi) {
        size_t mid = lo + ((hi - lo) >> 1);
        if (c[mid] < r) lo = mid + 1; else hi = mid;
    }
    d.idx[t] = (uint32_t)(lo < d.n ? lo : d.n - 1);
}

// index walk + pose clone in one pass (used after the fused post kernel)
__global__ void __launch_bounds__(256) fs_search_pose_kernel(FsDev d) {
    pf_grid_dep_sync();

    if (!*d.gate) return;
    const size_t t = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (t >= d.n) return;
    const double r = d.rcomb[t];
    const double* __restrict__ c = d.cum;
    size_t lo = 0, hi = d.n;
    while (lo < hi) {
        size_t mid = lo + ((hi - lo) >> 1);
        if (c[mid] < r) lo = mid + 1; else hi = mid;
    }
    const size_t j = lo < d.n ? lo : d.n - 1;
    d.idx[t] = (uint32_t)j;
    if (d.alog) d.idxlog[(size_t)(d.counters[0] % (unsigned)d.alog) * d.n + t] = (uint32_t)j;   // ancestry log: generation G+1 -> G
    const int cur = *d.cur;
    fs_px(d, cur ^ 1)[t] = fs_px(d, cur)[j];
    fs_py(d, cur ^ 1)[t] = fs_py(d, cur)[j];
    fs_pyaw(d, cur ^ 1)[t] = fs_pyaw(d, cur)[j];
    d.w[t] = 1.0 / (double)d.n_global;                      // fs1.rs:228
}
// particles[j].clone(): the pose columns and weight (fs1.rs:227-229)
__global__ void __launch_bounds__(256) fs_gather_pose_kernel(FsDev d) {
    if (!*d.gate) return;
    const size_t t = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (t >= d.n) return;
    const int cur = *d.cur;
    const size_t j = d.idx[t];
    fs_px(d, cur ^ 1)[t] = fs_px(d, cur)[j];
    fs_py(d, cur ^ 1)[t] = fs_py(d, cur)[j];
    fs_pyaw(d, cur ^ 1)[t] = fs_pyaw(d, cur)[j];
    d.w[t] = 1.0 / (double)d.n_global;
}
// ... and the map: instead of copying 48*m bytes per particle, compose every landmark's ancestry column with this
// resample's (monotone) ancestry idx: anc'[l][t] = anc[l][idx[t]]  (idx[t] itself where the landmark is identity-mapped).
// grid.x = particle chunks, grid.y = groups of FS_COMPOSE_ROWS landmarks.
#define FS_COMPOSE_ROWS 16
template <class AncT>
__global__ void __launch_bounds__(256) fs_compose_anc_kernel(FsDev d) {
    pf_grid_dep_sync();

    if (!*d.gate) return;
    const size_t t = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (t >= d.n) return;
    const int ac = *d.anc_cur;
    const AncT* __restrict__ src = reinterpret_cast<const AncT*>(fs_anc(d, ac));
    AncT* __restrict__ dst = reinterpret_cast<AncT*>(fs_anc(d, ac ^ 1));
    const AncT j = (AncT)d.idx[t];
    const size_t l0 = (size_t)blockIdx.y * FS_COMPOSE_ROWS;
#pragma unroll
    for (int rr = 0; rr < FS_COMPOSE_ROWS; ++rr) {
        size_t l = l0 + rr;
        if (l < d.m) dst[l * d.n + t] = (d.lmstate[l] & 2) ? j : src[l * d.n + j];
    }
}
// Everything a resample does once the exact CDF and comb are known, one pass per particle slot t (gated):
//   index walk (fs1.rs:224-226) as a lower bound, pose clone + weight (fs1.rs:227-229), and the lazy map clone: compose
// Same composition, 4 consecutive slots per thread: one 8 / 16-byte store per landmark row instead of four 2 / 4-byte ones
// (the kernel is bound by load/store instruction issue, not by bytes; the four gathers of a row mostly share a sector because
// systematic ancestries ascend by about one per slot).  Needs n % 4 == 0.
template <class AncT> struct AncVec4;
template <> struct AncVec4<unsigned short> { typedef ushort4 type; };
template <> struct AncVec4<uint32_t> { typedef uint4 type; };
// ... and the ping-pong flip (fs_flip_kernel) is done by the last CTA to finish: one launch fewer per step
template <class AncT>
__global__ void __launch_bounds__(256) fs_compose_flip_kernel(FsDev d) {
    pf_grid_dep_sync();

    if (!*d.gate) return;
    typedef typename AncVec4<AncT>::type V;
    __shared__ int s_last;
    const size_t t = ((size_t)blockIdx.x * 256 + threadIdx.x) * 4;
    if (t < d.n) {
        const int ac = *d.anc_cur;
        const AncT* __restrict__ src = reinterpret_cast<const AncT*>(fs_anc(d, ac));
        AncT* __restrict__ dst = reinterpret_cast<AncT*>(fs_anc(d, ac ^ 1));
        const uint4 jj = *reinterpret_cast<const uint4*>(d.idx + t);
        const size_t l0 = (size_t)blockIdx.y * FS_COMPOSE_ROWS;
#pragma unroll
        for (int rr = 0; rr < FS_COMPOSE_ROWS; ++rr) {
            const size_t l = l0 + rr;
            if (l >= d.m) break;
            V o;
            if (d.lmstate[l] & 2) { o.x = (AncT)jj.x; o.y = (AncT)jj.y; o.z = (AncT)jj.z; o.w = (AncT)jj.w; }
            else {
                const AncT* __restrict__ row = src + l * d.n;
                o.x = row[jj.x]; o.y = row[jj.y]; o.z = row[jj.z]; o.w = row[jj.w];
            }
            *reinterpret_cast<V*>(dst + l * d.n + t) = o;
        }
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(&d.counters[2], 1u) + 1u == gridDim.x * gridDim.y) ? 1 : 0;
    __syncthreads();
    if (!s_last) return;                                           // every other CTA has read lmstate / anc_cur by now
    for (size_t l = threadIdx.x; l < d.m; l += blockDim.x) d.lmstate[l] &= 1;
    if (threadIdx.x == 0) { *d.cur ^= 1; *d.anc_cur ^= 1; d.counters[0] += 1; d.counters[2] = 0; }
}
// the same kernel for the ANCESTRY LOG form (FsDev::alog != 0): most resamples only flip; see the comment inside
template <class AncT>
__global__ void __launch_bounds__(256) fs_compose_flip_alog_kernel(FsDev d) {
    pf_grid_dep_sync();
    if (!*d.gate) return;
    typedef typename AncVec4<AncT>::type V;
    __shared__ int s_last;
    const unsigned G = d.counters[0];
    // ancestry log: rows are left stale (the resample's index array is in the ring); only every alog-th resample recomposes
    // all of them to the new generation G + 1, walking each slot back through the ring to the row's own generation
    const bool alog = true;
    const bool compose = (G + 1u) % (unsigned)d.alog == 0u;
    const size_t t = ((size_t)blockIdx.x * 256 + threadIdx.x) * 4;
    if (compose && t < d.n) {
        const int ac = *d.anc_cur;
        const AncT* __restrict__ src = reinterpret_cast<const AncT*>(fs_anc(d, ac));
        AncT* __restrict__ dst = reinterpret_cast<AncT*>(fs_anc(d, ac ^ 1));
        const uint4 jj = *reinterpret_cast<const uint4*>(d.idx + t);
        const size_t l0 = (size_t)blockIdx.y * FS_COMPOSE_ROWS;
#pragma unroll
        for (int rr = 0; rr < FS_COMPOSE_ROWS; ++rr) {
            const size_t l = l0 + rr;
            if (l >= d.m) break;
            uint4 j4 = jj;                                         // slots of generation G
            if (alog) {
                j4.x = (unsigned)fs_walk_back<true>(d, l, jj.x, G); j4.y = (unsigned)fs_walk_back<true>(d, l, jj.y, G);
                j4.z = (unsigned)fs_walk_back<true>(d, l, jj.z, G); j4.w = (unsigned)fs_walk_back<true>(d, l, jj.w, G);
            }
            V o;
            if (d.lmstate[l] & 2) { o.x = (AncT)j4.x; o.y = (AncT)j4.y; o.z = (AncT)j4.z; o.w = (AncT)j4.w; }
            else {
                const AncT* __restrict__ row = src + l * d.n;
                o.x = row[j4.x]; o.y = row[j4.y]; o.z = row[j4.z]; o.w = row[j4.w];
            }
            *reinterpret_cast<V*>(dst + l * d.n + t) = o;
        }
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(&d.counters[2], 1u) + 1u == gridDim.x * gridDim.y) ? 1 : 0;
    __syncthreads();
    if (!s_last) return;                                           // every other CTA has read lmstate / anc_cur / gen by now
    if (compose) for (size_t l = threadIdx.x; l < d.m; l += blockDim.x) { d.lmstate[l] &= 1; if (alog) d.gen[l] = (int)(G + 1u); }
    if (threadIdx.x == 0) { *d.cur ^= 1; if (compose) *d.anc_cur ^= 1; d.counters[0] = G + 1u; d.counters[2] = 0; }
}
__global__ void fs_flip_kernel(FsDev d) {
    pf_grid_dep_sync();

    if (!*d.gate) return;
    for (size_t l = threadIdx.x; l < d.m; l += blockDim.x) d.lmstate[l] &= 1;     // no landmark is identity-mapped any more
    if (threadIdx.x == 0) { *d.cur ^= 1; *d.anc_cur ^= 1; d.counters[0] += 1; }
}

// get_best_particle fs1.rs:269-274: max_by keeps the LAST maximum
__global__ void __launch_bounds__(256) fs_best_kernel(FsDev d, int nblocks) {
    __shared__ double sw[256];
    __shared__ unsigned long long si[256];
    double bw = -1.0; unsigned long long bi = 0; bool have = false;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < d.n; i += (size_t)nblocks * 256) {
        double x = d.w[i];
        if (!have || x >= bw) { bw = x; bi = i; have = true; }
    }
    sw[threadIdx.x] = have ? bw : -1.0; si[threadIdx.x] = have ? bi : 0ull;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            double ow = sw[threadIdx.x + o]; unsigned long long oi = si[threadIdx.x + o];
            if (ow > sw[threadIdx.x] || (ow == sw[threadIdx.x] && oi > si[threadIdx.x])) { sw[threadIdx.x] = ow; si[threadIdx.x] = oi; }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { d.best_w[blockIdx.x] = sw[0]; d.best_i[blockIdx.x] = si[0]; }
}

// AoS <-> SoA converters for upload/download (pose_w: n x 4, lm: n x m x 6 particle-major like Vec<Particle>)
__global__ void __launch_bounds__(256) fs_unpack_pose_kernel(FsDev d, const double* pose_w) {
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= d.n) return;
    const int cur = *d.cur;
    d.w[i] = pose_w[4 * i]; fs_px(d, cur)[i] = pose_w[4 * i + 1]; fs_py(d, cur)[i] = pose_w[4 * i + 2]; fs_pyaw(d, cur)[i] = pose_w[4 * i + 3];
}
__global__ void __launch_bounds__(256) fs_pack_pose_kernel(FsDev d, double* pose_w) {
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= d.n) return;
    const int cur = *d.cur;
    pose_w[4 * i] = d.w[i]; pose_w[4 * i + 1] = fs_px(d, cur)[i]; pose_w[4 * i + 2] = fs_py(d, cur)[i]; pose_w[4 * i + 3] = fs_pyaw(d, cur)[i];
}
// chunk of particles [i0, i0+cnt): aos = cnt x m x 6
__global__ void __launch_bounds__(256) fs_unpack_lm_kernel(FsDev d, const double* aos, size_t i0, size_t cnt) {
    const size_t e = (size_t)blockIdx.x * 256 + threadIdx.x;      // element index in the AoS chunk
    const size_t tot = cnt * d.m * 6;
    if (e >= tot) return;
    size_t ip = e / (d.m * 6), rem = e % (d.m * 6);
    size_t l = rem / 6; int f = (int)(rem % 6);
    fs_lm(d, d.eager ? *d.cur : 0)[lm_index(d.ld, l, f, i0 + ip)] = aos[e];
}
__global__ void __launch_bounds__(256) fs_pack_lm_kernel(FsDev d, double* aos, size_t i0, size_t cnt) {
    const size_t e = (size_t)blockIdx.x * 256 + threadIdx.x;
    const size_t tot = cnt * d.m * 6;
    if (e >= tot) return;
    size_t ip = e / (d.m * 6), rem = e % (d.m * 6);
    size_t l = rem / 6; int f = (int)(rem % 6);
    const int st = d.lmstate[l];
    const size_t col = d.alog ? fs_lm_col<true>(d, l, i0 + ip, st, d.counters[0])
                              : ((st & 2) ? (i0 + ip) : fs_anc_load(d, *d.anc_cur, l * d.n + i0 + ip));   // materialise through the ancestry
    aos[e] = fs_lm(d, st & 1)[lm_index(d.ld, l, f, col)];
}
__global__ void fs_lmstate_reset_kernel(FsDev d) {     // every landmark identity-mapped in buffer 0 (eager mode: *cur) after init/upload/seed
    const int buf = d.eager ? *d.cur : 0;
    for (size_t l = (size_t)blockIdx.x * blockDim.x + threadIdx.x; l < d.m; l += (size_t)gridDim.x * blockDim.x) {
        d.lmstate[l] = buf | 2;
        if (d.alog) d.gen[l] = (int)d.counters[0];             // identity of the CURRENT generation
    }
}
// create_particles fs1.rs:302-306: Particle::new (fs1.rs:54-62) with Landmark::new (fs1.rs:34-40)
__global__ void __launch_bounds__(256) fs_init_kernel(FsDev d, double init_weight) {
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= d.n) return;
    d.w[i] = init_weight; d.w_raw[i] = init_weight;
    d.px[0][i] = 0.0; d.py[0][i] = 0.0; d.pyaw[0][i] = 0.0;
    for (size_t l = 0; l < d.m; ++l) {
        d.lm[0][lm_index(d.ld, l, 0, i)] = 0.0; d.lm[0][lm_index(d.ld, l, 1, i)] = 0.0;
        d.lm[0][lm_index(d.ld, l, 2, i)] = 1000.0; d.lm[0][lm_index(d.ld, l, 3, i)] = 0.0;
        d.lm[0][lm_index(d.ld, l, 4, i)] = 0.0; d.lm[0][lm_index(d.ld, l, 5, i)] = 1000.0;
    }
}

// pfgpu_fs_seed_map: initialised map for benchmarks/tests (see include/pfgpu.h)
__global__ void __launch_bounds__(256) fs_seed_pose_kernel(FsDev d, double x, double y, double yaw) {
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= d.n) return;
    const int cur = *d.cur;
    fs_px(d, cur)[i] = x; fs_py(d, cur)[i] = y; fs_pyaw(d, cur)[i] = yaw;
    d.w[i] = 1.0 / (double)d.n_global; d.w_raw[i] = d.w[i];
}
__global__ void __launch_bounds__(256) fs_seed_lm_kernel(FsDev d, const double* lm_xy, double sigma, double cov0, uint64_t seed) {
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    const size_t l = blockIdx.y;
    if (i >= d.n) return;
    double* lm = fs_lm(d, d.eager ? *d.cur : 0);
    double z0, z1;
    pfc_normal_pair(pfc_rng_block(seed, PFC_STREAM_INIT_A, 0, (d.offset + i) * d.m + l), &z0, &z1);
    lm[lm_index(d.ld, l, 0, i)] = lm_xy[2 * l] + sigma * z0;
    lm[lm_index(d.ld, l, 1, i)] = lm_xy[2 * l + 1] + sigma * z1;
    lm[lm_index(d.ld, l, 2, i)] = cov0; lm[lm_index(d.ld, l, 3, i)] = 0.0;
    lm[lm_index(d.ld, l, 4, i)] = 0.0; lm[lm_index(d.ld, l, 5, i)] = cov0;
}
