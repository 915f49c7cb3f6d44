/* host probe of include/fs2_math.h (the pose proposal of FastSLAM 2.0 as the CUDA kernel evaluates it) and of the variant-2
 * landmark update of include/fs_ekf_math.h; tests/test_fs2_math_host.py compares both with oracle/fs2_oracle.c. */
#include "../../include/fs2_math.h"

/* in: pose3, u2, z2, lm6, n3  -> out3 */
void fs2_probe_propose(const double* pose, const double* u, const double* z, const double* lm6, const double* n3, double dt,
                       double r00, double r11, double* out3) {
    static const double mc[9] = { 0.1, 0.0, 0.0, 0.0, 0.1, 0.0, 0.0, 0.0, 0.01 };
    FsLm L = { lm6[0], lm6[1], lm6[2], lm6[3], lm6[4], lm6[5] };
    double x = pose[0], y = pose[1], a = pose[2];
    fs2_propose_pose(&x, &y, &a, &L, u[0], u[1], dt, z[0], z[1], r00, r11, mc, n3[0], n3[1], n3[2]);
    out3[0] = x; out3[1] = y; out3[2] = a;
}
/* update_landmark_and_weight through the contract form: lm6 in/out, returns the weight factor */
double fs2_probe_update(double* lm6, const double* pose, const double* z, double r00, double r11) {
    FsLm L = { lm6[0], lm6[1], lm6[2], lm6[3], lm6[4], lm6[5] };
    int wrote;
    const double f = fs_update_landmark_v(&L, pose[0], pose[1], pose[2], z[0], z[1], r00, r11, &wrote, 2);
    lm6[0] = L.x; lm6[1] = L.y; lm6[2] = L.c00; lm6[3] = L.c01; lm6[4] = L.c10; lm6[5] = L.c11;
    return f;
}
