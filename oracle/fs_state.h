/*
 * fs_state.h — the particle set shared by fs1_oracle.c (FastSLAM 1.0) and fs2_oracle.c (FastSLAM 2.0).
 * TEST INFRASTRUCTURE ONLY — see oracle.h.
 */
#ifndef ORC_FS_STATE_H
#define ORC_FS_STATE_H
#include "oracle.h"
#include "../include/pf_contract_math.h"
#include <stdlib.h>

#ifdef PF_ORACLE_LIBM
#define M_EXP(x) exp(x)
#define M_SIN(x) sin(x)
#define M_COS(x) cos(x)
#define M_ATAN2(y, x) atan2(y, x)
#else
#define M_EXP(x) pfc_exp(x)
#define M_SIN(x) pfc_sin(x)
#define M_COS(x) pfc_cos(x)
#define M_ATAN2(y, x) pfc_atan2(y, x)
#endif

typedef struct { double x, y, c00, c01, c10, c11; } lm_t;       /* fs1.rs:27-31 */

struct orc_fs {
    orc_fs_config cfg;
    size_t n, m;
    double *w, *x, *y, *yaw;      /* fs1.rs:45-50 */
    lm_t* lm;                     /* [particle][landmark] */
    double *w2, *x2, *y2, *yaw2; lm_t* lm2;   /* resample target */
    uint64_t seed; uint32_t n_step, n_resample;
    uint32_t* last_idx; size_t last_idx_n;
    double last_neff;
    int threads;
    int variant;                  /* 1 = FastSLAM 1.0 (fs1.rs), 2 = FastSLAM 2.0 (fs2.rs) */
};

static inline double orc_fs_normalize_angle(double a) {         /* fs1.rs:80-89 = fs2.rs:84-93 */
    while (a > PFC_PI) a -= 2.0 * PFC_PI;
    while (a < -PFC_PI) a += 2.0 * PFC_PI;
    return a;
}
/* shared by the two variants (fs1.rs:186-234 and fs2.rs:282-323 are the same text) */
void   orc_fs_normalize_weights_(orc_fs* f);
double orc_fs_compute_neff_(const orc_fs* f);
void   orc_fs_resample_(orc_fs* f, double u01);
/* fs2_oracle.c */
void   orc_fs2_particle_(orc_fs* f, size_t i, const double u[2], const orc_fs_obs* z, size_t k, double n0, double n1, double n2);


#endif
