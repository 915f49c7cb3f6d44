// The C++ host mirror, compiled (no device needed: g++ -std=c++17 -c, __graft_entry__.build()) and RUN (tests/test_gpu_parity.py::
// test_cpp_mirror_runs builds it with -DMIRROR_MAIN, links libpfgpu.so and compares what it prints with the Python mirror
// driving the same ABI with the same seeds and inputs: identical numbers, bit for bit).
#include <cstdio>
#include "particle_filter.hpp"
#include "fastslam1.hpp"

using namespace rust_robotics_b200;

// one ParticleFilterLocalizer step (pf.rs:488-497) + StateEstimator-style update, one fastslam_update (fs1.rs:237-266)
int mirror_run(int device, double out[8]) {
    ParticleFilterConfig cfg; cfg.n_particles = 1000; cfg.range_noise = 0.25;
    cfg.validate();
    ParticleFilterLocalizer pf(cfg, 42, device);
    pf.with_initial_state({5.0, 5.0, 0.0, 0.0});
    PFMeasurement z = {{3.1, 2.0, 2.0}, {5.0, 10.0, 2.0}};
    PFState est = pf.try_step({1.1, 0.0}, z);
    pf.predict({0.5, 0.63}, 0.1);
    pf.update(z);                                             // update + resample (pf.rs:561-564)
    PFState est2 = pf.get_state();
    fastslam1::FastSlam fs(256, 4, 42, device);
    bool did = fastslam1::fastslam_update(fs, {1.0, 0.1}, {{5.0, 0.1, 0}, {7.0, -0.4, 2}});
    fastslam1::Particle best = fastslam1::get_best_particle(fs);
    fastslam2::FastSlam fs2(64, 4, 42, device);                // FastSLAM 2.0 (fs2.rs:376-383): exercised, not printed
    (void)fastslam2::fastslam2_update(fs2, {1.0, 0.1}, {{5.0, 0.1, 0}, {7.0, -0.4, 2}});
    out[0] = est[0]; out[1] = est[1]; out[2] = est2[0]; out[3] = est2[3];
    out[4] = (double)pf.get_particles().size(); out[5] = best.weight; out[6] = best.landmarks[0].x + best.landmarks[2].y; out[7] = did ? 1.0 : 0.0;
    return 0;
}
#ifdef MIRROR_MAIN
int main() {
    double o[8];
    try { mirror_run(0, o); } catch (const std::exception& e) { std::fprintf(stderr, "mirror: %s\n", e.what()); return 1; }
    for (int i = 0; i < 8; ++i) std::printf("%.17g\n", o[i]);
    return 0;
}
#endif
