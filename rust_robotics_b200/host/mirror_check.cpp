// Compile check of the C++ host mirror (no device needed): g++ -std=c++17 -c mirror_check.cpp -I include
#include "particle_filter.hpp"
#include "fastslam1.hpp"

using namespace rust_robotics_b200;

double mirror_smoke(int device) {
    ParticleFilterConfig cfg; cfg.n_particles = 1000; cfg.range_noise = 0.25;
    cfg.validate();
    ParticleFilterLocalizer pf(cfg, 42, device);
    pf.with_initial_state({5.0, 5.0, 0.0, 0.0});
    PFMeasurement z = {{3.1, 2.0, 2.0}, {5.0, 10.0, 2.0}};
    PFState est = pf.try_step({1.1, 0.0}, z);
    fastslam1::FastSlam fs(256, 4);
    fastslam1::fastslam_update(fs, {1.0, 0.1}, {{5.0, 0.1, 0}});
    return est[0] + fastslam1::get_best_particle(fs).weight + pf.get_particles().size();
}
