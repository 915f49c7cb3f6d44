// fastslam1.hpp — C++ host-side mirror of crates/rust_robotics_slam/src/fastslam1.rs (fs1.rs) over include/pfgpu.h.
// The reference is a set of free functions over a caller-owned Vec<Particle>; the GPU form keeps the particles on the
// device inside an engine object and offers the same four entry points under the same names.
#pragma once
#include <array>
#include <tuple>
#include <vector>
#include "pfgpu.h"
#include "particle_filter.hpp"

namespace rust_robotics_b200 { namespace fastslam1 {

struct Landmark { double x, y; std::array<double, 4> cov; };              // fs1.rs:27-31 (cov row-major c00 c01 c10 c11)
struct Particle { double weight, x, y, yaw; std::vector<Landmark> landmarks; };   // fs1.rs:45-51
using Observation = std::tuple<double, double, size_t>;                   // (distance, angle, landmark_id) fs1.rs:240

class FastSlam {
    pfgpu_fs* h_ = nullptr;
    size_t n_ = 0, m_ = 0;
public:
    FastSlam(size_t n_particles, size_t n_landmarks, uint64_t seed = 42, int device = 0, const pfgpu_fs_config* cfg = nullptr) {
        pfgpu_fs_config c; pfgpu_fs_default_config(&c); if (cfg) c = *cfg;
        check(pfgpu_fs_create(&c, n_particles, n_landmarks, seed, device, &h_), "create_particles");
        n_ = n_particles; m_ = n_landmarks;
    }
    FastSlam(const FastSlam&) = delete;
    FastSlam& operator=(const FastSlam&) = delete;
    ~FastSlam() { pfgpu_fs_destroy(h_); }
    // fastslam_update fs1.rs:237-266
    bool step(const std::array<double, 2>& u, const std::vector<Observation>& z) {
        std::vector<pfgpu_fs_obs> o(z.size());
        for (size_t i = 0; i < z.size(); ++i) { o[i].d = std::get<0>(z[i]); o[i].angle = std::get<1>(z[i]); o[i].lm_id = std::get<2>(z[i]); }
        int did = 0;
        check(pfgpu_fs_step(h_, u.data(), o.data(), o.size(), &did), "fastslam_update");
        return did != 0;
    }
    // get_best_particle fs1.rs:269-274 (pose + weight + that particle's landmarks, what render_gif_slam.rs:183-191 reads)
    Particle best() const {
        size_t idx = 0; double pw[4];
        check(pfgpu_fs_best(h_, &idx, pw), "get_best_particle");
        Particle p{pw[0], pw[1], pw[2], pw[3], {}};
        std::vector<double> lm(6 * m_);
        check(pfgpu_fs_particle_landmarks(h_, idx, lm.data()), "landmarks");
        for (size_t l = 0; l < m_; ++l) p.landmarks.push_back({lm[6 * l], lm[6 * l + 1], {lm[6 * l + 2], lm[6 * l + 3], lm[6 * l + 4], lm[6 * l + 5]}});
        return p;
    }
    // Vec<Particle> view (checkpoint / API-compat tests)
    std::vector<Particle> download() const {
        std::vector<double> pw(4 * n_), lm(6 * n_ * m_);
        check(pfgpu_fs_download(h_, pw.data(), lm.data(), n_), "download");
        std::vector<Particle> out(n_);
        for (size_t i = 0; i < n_; ++i) {
            out[i] = {pw[4 * i], pw[4 * i + 1], pw[4 * i + 2], pw[4 * i + 3], {}};
            for (size_t l = 0; l < m_; ++l) { const double* q = &lm[(i * m_ + l) * 6]; out[i].landmarks.push_back({q[0], q[1], {q[2], q[3], q[4], q[5]}}); }
        }
        return out;
    }
    size_t len() const { return n_; }
    // 1 = fastslam1::fastslam_update, 2 = fastslam2::fastslam2_update (crates/rust_robotics_slam/src/fastslam2.rs:376-383)
    void set_variant(int variant) { check(pfgpu_fs_set_variant(h_, variant), "set_variant"); }
};

// the reference's free-function names
inline FastSlam create_particles(size_t n_particles, size_t n_landmarks) = delete;   // engines are not copyable: construct FastSlam directly
inline bool fastslam_update(FastSlam& particles, const std::array<double, 2>& u, const std::vector<Observation>& z) { return particles.step(u, z); }
inline Particle get_best_particle(const FastSlam& particles) { return particles.best(); }

}  // namespace fastslam1

// crates/rust_robotics_slam/src/fastslam2.rs: the same engine object with the proposal-sampling step
namespace fastslam2 {
using fastslam1::Landmark; using fastslam1::Particle; using fastslam1::Observation; using fastslam1::get_best_particle;
struct FastSlam : fastslam1::FastSlam {
    FastSlam(size_t n_particles, size_t n_landmarks, uint64_t seed = 42, int device = 0, const pfgpu_fs_config* cfg = nullptr)
        : fastslam1::FastSlam(n_particles, n_landmarks, seed, device, cfg) { set_variant(2); }
};
inline bool fastslam2_update(FastSlam& particles, const std::array<double, 2>& u, const std::vector<Observation>& z) { return particles.step(u, z); }
}  // namespace fastslam2
}  // namespace rust_robotics_b200
