// particle_filter.hpp — C++ host-side mirror of the reference's ParticleFilterLocalizer / MonteCarloLocalizer over
// the C ABI (include/pfgpu.h).  The reference is compiled code (Rust) and this image has no Rust toolchain, so the
// host side above the ABI is written in C++ with the SAME type names, method names, argument meaning and error
// behaviour as crates/rust_robotics_localization/src/particle_filter.rs (pf.rs) and monte_carlo_localization.rs
// (mcl.rs).  Header-only; link against libpfgpu.so.  No CPU fallback: constructors throw when no CUDA device exists.
#pragma once
#include <array>
#include <stdexcept>
#include <string>
#include <tuple>
#include <vector>
#include "pfgpu.h"

namespace rust_robotics_b200 {

// crates/rust_robotics_core/src/error.rs:8-24 — this path only ever produces InvalidParameter
struct RoboticsError : std::runtime_error {
    enum Kind { InvalidParameter, Runtime } kind;
    RoboticsError(Kind k, const std::string& m) : std::runtime_error(m), kind(k) {}
};
inline void check(int status, const char* what) {
    if (status == PFGPU_OK) return;
    std::string msg = std::string(what) + ": " + pfgpu_strerror(status);
    if (status > 0) msg += std::string(" (") + pfgpu_last_error() + ")";
    throw RoboticsError(status < 0 ? RoboticsError::InvalidParameter : RoboticsError::Runtime, msg);
}

using PFState = std::array<double, 4>;                                   // Vector4<f64> pf.rs:16
using PFControl = std::array<double, 2>;                                 // Vector2<f64> pf.rs:19
using PFMeasurement = std::vector<std::tuple<double, double, double>>;   // (distance, landmark_x, landmark_y) pf.rs:22
struct Particle { double x, y, yaw, v, w; };                             // pf.rs:26-32

struct ParticleFilterConfig {                                            // pf.rs:52-78
    size_t n_particles = 100;
    double resample_threshold = 0.5, range_noise = 0.2, velocity_noise = 2.0;
    double yaw_rate_noise = 40.0 * 3.14159265358979323846 / 180.0, dt = 0.1;
    pfgpu_pf_config to_c() const { pfgpu_pf_config c{}; c.n_particles = n_particles; c.resample_threshold = resample_threshold;
        c.range_noise = range_noise; c.velocity_noise = velocity_noise; c.yaw_rate_noise = yaw_rate_noise; c.dt = dt; c.mode = 0;
        c.max_particles = n_particles; c.kld_epsilon = 0.05; c.kld_z = 2.326; return c; }
    void validate() const { auto c = to_c(); check(pfgpu_pf_config_validate(&c), "particle filter configuration"); }   // pf.rs:81-117
};

struct MonteCarloLocalizationConfig {                                    // mcl.rs:50-74
    size_t min_particles = 100, max_particles = 5000;
    double kld_epsilon = 0.05, kld_z = 2.326, range_noise = 0.2, velocity_noise = 2.0;
    double yaw_rate_noise = 40.0 * 3.14159265358979323846 / 180.0, dt = 0.1;
    pfgpu_pf_config to_c() const { pfgpu_pf_config c{}; c.n_particles = min_particles; c.range_noise = range_noise;
        c.velocity_noise = velocity_noise; c.yaw_rate_noise = yaw_rate_noise; c.dt = dt; c.mode = 1; c.max_particles = max_particles;
        c.kld_epsilon = kld_epsilon; c.kld_z = kld_z; return c; }
    void validate() const { auto c = to_c(); check(pfgpu_pf_config_validate(&c), "MCL configuration"); }               // mcl.rs:87-130
};

namespace detail {
class PfHandle {
protected:
    pfgpu_pf* h_ = nullptr;
    mutable std::vector<Particle> mirror_;     // get_particles() returns a reference in the reference API: lazily refreshed host mirror
    mutable bool dirty_ = true;
    PfHandle(const pfgpu_pf_config& c, uint64_t seed, int device) { check(pfgpu_pf_create(&c, seed, device, &h_), "create"); }
public:
    PfHandle(const PfHandle&) = delete;
    PfHandle& operator=(const PfHandle&) = delete;
    ~PfHandle() { pfgpu_pf_destroy(h_); }
    static std::vector<double> flat(const PFMeasurement& z) {
        std::vector<double> o; o.reserve(3 * z.size());
        for (auto& t : z) { o.push_back(std::get<0>(t)); o.push_back(std::get<1>(t)); o.push_back(std::get<2>(t)); }
        return o;
    }
    void try_predict_with_control(const PFControl& u) { check(pfgpu_pf_predict(h_, u.data()), "predict"); dirty_ = true; }        // pf.rs:255
    void try_update_with_observations(const PFMeasurement& z) { auto f = flat(z); check(pfgpu_pf_update(h_, f.data(), z.size()), "update"); dirty_ = true; }  // pf.rs:310
    void resample() { int did = 0; check(pfgpu_pf_resample(h_, &did), "resample"); dirty_ = true; }                               // pf.rs:337
    PFState try_step(const PFControl& u, const PFMeasurement& z) {                                                               // pf.rs:488
        auto f = flat(z); PFState est{}; check(pfgpu_pf_step(h_, u.data(), f.data(), z.size(), est.data()), "step"); dirty_ = true; return est; }
    PFState step(const PFControl& u, const PFMeasurement& z) { return try_step(u, z); }
    PFState estimate() const { PFState e{}; check(pfgpu_pf_estimate(h_, e.data(), nullptr), "estimate"); return e; }            // pf.rs:348
    std::array<double, 16> calc_covariance() const { std::array<double, 16> c{}; check(pfgpu_pf_estimate(h_, nullptr, c.data()), "covariance"); return c; }  // column-major, pf.rs:363
    size_t particle_count() const { size_t nl = 0, ng = 0; check(pfgpu_pf_count(h_, &nl, &ng), "count"); return ng; }            // mcl.rs:318
    const std::vector<Particle>& get_particles() const {                                                                         // pf.rs:244
        if (dirty_) { size_t nl = 0, ng = 0; check(pfgpu_pf_count(h_, &nl, &ng), "count"); mirror_.resize(nl);
            check(pfgpu_pf_download(h_, reinterpret_cast<double*>(mirror_.data()), nl), "download"); dirty_ = false; }
        return mirror_;
    }
    void init_state(const PFState& s) { check(pfgpu_pf_init_state(h_, s.data()), "initial state"); dirty_ = true; }
};
}  // namespace detail

class ParticleFilterLocalizer : public detail::PfHandle {                // pf.rs:121-573
public:
    explicit ParticleFilterLocalizer(const ParticleFilterConfig& c = {}, uint64_t seed = 42, int device = 0) : PfHandle(c.to_c(), seed, device) {}
    static ParticleFilterLocalizer try_with_initial_state(const PFState& s, const ParticleFilterConfig& c, uint64_t seed = 42, int device = 0) = delete;
    void with_initial_state(const PFState& s) { init_state(s); }                                                               // pf.rs:164-199
    void set_range_noise(double s) { check(pfgpu_pf_set_range_noise(h_, s), "range_noise"); }                                   // pf.rs:228
    // StateEstimator (traits.rs:31-52; pf.rs:552-573): update = update_with_observations + resample
    void predict(const PFControl& u, double /*dt ignored, pf.rs:557*/) { try_predict_with_control(u); }
    void update(const PFMeasurement& z) { try_update_with_observations(z); resample(); }
    PFState get_state() const { return estimate(); }
};

class MonteCarloLocalizer : public detail::PfHandle {                    // mcl.rs:133-471
public:
    explicit MonteCarloLocalizer(const MonteCarloLocalizationConfig& c = {}, uint64_t seed = 42, int device = 0) : PfHandle(c.to_c(), seed, device) {}
    void with_initial_state(const PFState& s) { init_state(s); }                                                               // mcl.rs:167-206
    void predict(const PFControl& u, double) { try_predict_with_control(u); }                                                   // mcl.rs:455
    void update(const PFMeasurement& z) { try_update_with_observations(z); resample(); }                                        // mcl.rs:459-462
};

}  // namespace rust_robotics_b200
