// pf_kernels.cuh — ParticleFilterLocalizer / MonteCarloLocalizer device kernels.
// Reference: pf.rs = crates/rust_robotics_localization/src/particle_filter.rs, mcl.rs = .../monte_carlo_localization.rs.
//
// HBM layout (per shard of n particles):
//   pose[2][n]   32-byte records {x, y, yaw, v} (ping-pong; `cur` on the device selects the live one)
//                — one record = one 32 B DRAM sector, so the multinomial gather (random indices, pf.rs:455-470)
//                costs one sector per particle instead of four with four separate columns.
//   w_raw[n]     likelihood products before normalisation (pf.rs:317-328)
//   w[n]         normalised weights (pf.rs:426-439)
//   cum[n]       exact sequential cumulative weights (pf.rs:448-453)
//   idx[n]       resample ancestry (u32)
#pragma once
#include "common.cuh"
#include "xsum.cuh"

struct __align__(32) Pose4 { double x, y, yaw, v; };

#define PF_NT 256
#define PF_MOM 15              // sum w, 4 first moments, 10 second moments (upper triangle)

struct PfDev {
    size_t n = 0, n_global = 0, offset = 0;
    Pose4* pose[2] = {nullptr, nullptr};
    int* cur = nullptr;            // device: index of the live pose buffer
    double* w_raw = nullptr;
    double* w = nullptr;
    double* cum = nullptr;
    uint32_t* idx = nullptr;
    double* scal = nullptr;        // [0] S=sum w_raw  [1] Q=sum w^2  [2] cum total  [3] neff  [4..7] est  [8..23] cov (row-major)
    int* gate = nullptr;           // device: 1 if this step resamples
    double* partial = nullptr;     // [blocks][PF_MOM]
    double* obs = nullptr;         // device copy of the observation list (k x 3) when it does not fit the launch parameters
    unsigned int* counters = nullptr;   // device: [0] resamples done so far (= Philox call index of the next resample)
};

// Observation lists of up to PF_PARAM_OBS entries travel inside the kernel's launch parameters (no H2D copy).
#define PF_PARAM_OBS 32
struct PfObsParam { double o[3 * PF_PARAM_OBS]; };

// select by ternary: indexing the by-value parameter struct with a runtime index would force a stack copy
__device__ __forceinline__ Pose4* pf_pose(const PfDev& d, int cur) { return cur ? d.pose[1] : d.pose[0]; }
__device__ __forceinline__ void pose_load(const Pose4* p, size_t i, Pose4& o) {
    const double2* q = reinterpret_cast<const double2*>(p + i);
    double2 a = q[0], b = q[1];
    o.x = a.x; o.y = a.y; o.yaw = b.x; o.v = b.y;
}
__device__ __forceinline__ void pose_store(Pose4* p, size_t i, const Pose4& o) {
    double2* q = reinterpret_cast<double2*>(p + i);
    q[0] = make_double2(o.x, o.y); q[1] = make_double2(o.yaw, o.v);
}

// try_predict_with_control (pf.rs:279-296, mcl.rs:236-253) and/or the likelihood loop of
// try_update_with_observations (pf.rs:316-329, mcl.rs:273-283), fused in one pass over the pose records.
template <bool DO_PREDICT, bool DO_WEIGHT, bool PARAM_OBS>
__global__ void __launch_bounds__(PF_NT) pf_predict_weight_kernel(PfDev d, const __grid_constant__ PfObsParam po,
                                                                  double u0, double u1, double sv, double sw,
                                                                  double dt, uint64_t seed, uint32_t call,
                                                                  int k_obs, double sigma) {
    extern __shared__ double s_obs_pf[];     // k_obs x (d, lx, ly): the observation vector staged once per CTA
    if (DO_WEIGHT) {
        for (int j = threadIdx.x; j < 3 * k_obs; j += PF_NT) s_obs_pf[j] = PARAM_OBS ? po.o[j] : d.obs[j];
        __syncthreads();
    }
    const size_t i = (size_t)blockIdx.x * PF_NT + threadIdx.x;
    if (i >= d.n) return;
    Pose4* pose = pf_pose(d, *d.cur);
    Pose4 p;
    pose_load(pose, i, p);
    if (DO_PREDICT) {
        double z0, z1;
        pfc_normal_pair(pfc_rng_block(seed, PFC_STREAM_PF_PREDICT, call, d.offset + i), &z0, &z1);
        double v_noise = sv > 0.0 ? 0.0 + sv * z0 : 0.0;            // Normal::sample = mean + std*z; no draw if sigma == 0
        double yaw_noise = sw > 0.0 ? 0.0 + sw * z1 : 0.0;
        double v_noisy = u0 + v_noise;                              // pf.rs:289
        double yaw_rate_noisy = u1 + yaw_noise;                     // pf.rs:290
        double s, c;
        pfc_sincos(p.yaw, &s, &c);
        p.x = p.x + v_noisy * c * dt;                               // pf.rs:292
        p.y = p.y + v_noisy * s * dt;                               // pf.rs:293
        p.yaw = p.yaw + yaw_rate_noisy * dt;                        // pf.rs:294 (yaw is not wrapped)
        p.v = v_noisy;                                              // pf.rs:295
        pose_store(pose, i, p);
    }
    if (DO_WEIGHT) {
        const double coeff = 1.0 / sqrt(2.0 * PFC_PI * (sigma * sigma));   // gauss_likelihood pf.rs:476-479
        const double denom = 2.0 * (sigma * sigma);
        const pfc_rcp_t rdenom = pfc_rcp_make(denom);               // one reciprocal for all k_obs IEEE quotients
        double w = 1.0;                                             // pf.rs:317: the previous weight is discarded
        for (int j = 0; j < k_obs; ++j) {
            double dx = p.x - s_obs_pf[3 * j + 1];
            double dy = p.y - s_obs_pf[3 * j + 2];
            double d_pred = sqrt(dx * dx + dy * dy);
            double diff = s_obs_pf[3 * j] - d_pred;
            w = w * (coeff * pfc_exp(pfc_div_by(-(diff * diff), rdenom)));
        }
        d.w_raw[i] = w;
    }
}

// normalize_weights pf.rs:426-439 / mcl.rs:394-406
__global__ void __launch_bounds__(PF_NT) pf_normalize_kernel(PfDev d) {
    const size_t i = (size_t)blockIdx.x * PF_NT + threadIdx.x;
    if (i >= d.n) return;
    const double S = d.scal[0];
    d.w[i] = S > 0.0 ? d.w_raw[i] / S : 1.0 / (double)d.n_global;
}

struct PfValWSq { const double* w; __device__ __forceinline__ double operator()(size_t i) const { double x = w[i]; return x * x; } };

// calc_n_eff + gate: pf.rs:337-345, 416-423.  MCL resamples every step (mcl.rs:298).
__global__ void pf_gate_kernel(PfDev d, double threshold, int mode) {
    double Q = d.scal[1];
    double neff = Q > 0.0 ? 1.0 / Q : 0.0;
    d.scal[3] = neff;
    *d.gate = (mode == 1) ? 1 : (neff < (double)d.n_global * threshold ? 1 : 0);
}

// MCL: "if let Some(last) = cumulative_weights.last_mut() { *last = 1.0 }"  mcl.rs:334-336
__global__ void pf_force_last_kernel(PfDev d) {
    if (!*d.gate) return;
    d.cum[d.n - 1] = 1.0;
}

// index search: first i with r <= c_i (pf.rs:459-465, fallback 0; mcl.rs:387-392, fallback len-1).
// The cumulative weights are non-decreasing, so the linear scan equals a lower_bound.
__global__ void __launch_bounds__(PF_NT) pf_search_kernel(PfDev d, uint64_t seed, int mode) {
    if (!*d.gate) return;
    const size_t t = (size_t)blockIdx.x * PF_NT + threadIdx.x;
    if (t >= d.n) return;
    const uint32_t call = d.counters[0];
    double r = pfc_u01_53(pfc_blk_u64(pfc_rng_block(seed, PFC_STREAM_PF_RESAMPLE, call, d.offset + t), 0));
    const double* __restrict__ c = d.cum;
    size_t lo = 0, hi = d.n;
    while (lo < hi) {
        size_t mid = lo + ((hi - lo) >> 1);
        if (c[mid] < r) lo = mid + 1; else hi = mid;
    }
    size_t index = lo < d.n ? lo : (mode == 1 ? d.n - 1 : 0);
    d.idx[t] = (uint32_t)index;
}

// new_particles.push(particles[index].clone()); w = 1/n   (pf.rs:467-469, mcl.rs:351,357-361)
__global__ void __launch_bounds__(PF_NT) pf_gather_kernel(PfDev d) {
    if (!*d.gate) return;
    const size_t t = (size_t)blockIdx.x * PF_NT + threadIdx.x;
    if (t >= d.n) return;
    const int cur = *d.cur;
    Pose4 p;
    pose_load(pf_pose(d, cur), d.idx[t], p);
    pose_store(pf_pose(d, cur ^ 1), t, p);
    d.w[t] = 1.0 / (double)d.n_global;
}
__global__ void pf_flip_kernel(PfDev d) { if (*d.gate) { *d.cur ^= 1; d.counters[0] += 1; } }

__device__ __forceinline__ double pf_finite_or_zero(double c) { return (c - c == 0.0) ? c : 0.0; }
// compute_estimate + compute_covariance (pf.rs:382-413) in one pass with a shifted centre c (the previous
// estimate): sum w, sum w(p-c), sum w(p-c)(p-c)^T; finalised by pf_moments_final_kernel.  Tolerance-level
// quantity (1e-6): summed in tree order, not in the reference's sequential order.
__global__ void __launch_bounds__(PF_NT) pf_moments_kernel(PfDev d, int nblocks) {
    __shared__ double sm[PF_NT / 32];
    const Pose4* pose = pf_pose(d, *d.cur);
    // centre = previous estimate, or 0 when that is not finite (an overflowed pose must not poison every later estimate:
    // the reference's refresh_cache recomputes from scratch, pf.rs:382-413).  pf_moments_final_kernel applies the same rule.
    const double c0 = pf_finite_or_zero(d.scal[4]), c1 = pf_finite_or_zero(d.scal[5]), c2 = pf_finite_or_zero(d.scal[6]), c3 = pf_finite_or_zero(d.scal[7]);
    double acc[PF_MOM];
#pragma unroll
    for (int k = 0; k < PF_MOM; ++k) acc[k] = 0.0;
    for (size_t i = (size_t)blockIdx.x * PF_NT + threadIdx.x; i < d.n; i += (size_t)nblocks * PF_NT) {
        Pose4 p;
        pose_load(pose, i, p);
        double w = d.w[i];
        double e0 = p.x - c0, e1 = p.y - c1, e2 = p.yaw - c2, e3 = p.v - c3;
        double w0 = w * e0, w1 = w * e1, w2 = w * e2, w3 = w * e3;
        acc[0] += w;
        acc[1] += w0; acc[2] += w1; acc[3] += w2; acc[4] += w3;
        acc[5] += w0 * e0; acc[6] += w0 * e1; acc[7] += w0 * e2; acc[8] += w0 * e3;
        acc[9] += w1 * e1; acc[10] += w1 * e2; acc[11] += w1 * e3;
        acc[12] += w2 * e2; acc[13] += w2 * e3;
        acc[14] += w3 * e3;
    }
#pragma unroll
    for (int k = 0; k < PF_MOM; ++k) {
        double t = block_sum<PF_NT>(acc[k], sm);
        if (threadIdx.x == 0) d.partial[(size_t)blockIdx.x * PF_MOM + k] = t;
    }
}
// one CTA: reduce the per-block partials; out = moments about the centre (15 doubles)
__global__ void __launch_bounds__(PF_NT) pf_moments_reduce_kernel(const double* partial, int nblocks, double* out15) {
    __shared__ double sm[PF_NT / 32];
    for (int k = 0; k < PF_MOM; ++k) {
        double a = 0.0;
        for (int b = threadIdx.x; b < nblocks; b += PF_NT) a += partial[(size_t)b * PF_MOM + k];
        double t = block_sum<PF_NT>(a, sm);
        if (threadIdx.x == 0) out15[k] = t;
    }
}
// est = c + M1 (weights are normalised: sum w = 1 up to rounding; the reference does not divide either),
// cov_ij = M2_ij - M1_i*m_j - m_i*M1_j + W*m_i*m_j  with m = est - c, which equals sum w (p-est)(p-est)^T.
__global__ void pf_moments_final_kernel(PfDev d, const double* mom15) {
    if (threadIdx.x != 0) return;
    const double W = mom15[0];
    double c[4] = { pf_finite_or_zero(d.scal[4]), pf_finite_or_zero(d.scal[5]), pf_finite_or_zero(d.scal[6]), pf_finite_or_zero(d.scal[7]) };
    double M1[4] = { mom15[1], mom15[2], mom15[3], mom15[4] };
    double M2[4][4];
    int q = 5;
    for (int a = 0; a < 4; ++a) for (int b = a; b < 4; ++b) { M2[a][b] = mom15[q]; M2[b][a] = mom15[q]; q++; }
    double m[4];
    for (int a = 0; a < 4; ++a) m[a] = M1[a];            // est - c = sum w (p - c)   (exactly what sum w*p - c*W gives for W=1)
    for (int a = 0; a < 4; ++a) d.scal[4 + a] = c[a] * W + M1[a];   // = sum w*p
    for (int a = 0; a < 4; ++a)
        for (int b = 0; b < 4; ++b) {
            // deviations about est: (p - est) = (p - c) - (est - c), est - c = c*(W-1) + M1
            double ea = c[a] * (W - 1.0) + m[a], eb = c[b] * (W - 1.0) + m[b];
            d.scal[8 + a * 4 + b] = M2[a][b] - M1[a] * eb - ea * M1[b] + W * ea * eb;
        }
}
