/* nbody_oracle.h -- CPU restatement of the reference hot path.  TEST INFRASTRUCTURE ONLY:
 * see the header of nbody_oracle.c for who may use it and how it is pinned. */
#ifndef NBODY_ORACLE_H_
#define NBODY_ORACLE_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

int oracle_disk_galaxy(uint64_t n, float *x, float *y, float *z, float *vx, float *vy, float *vz);

/* method: 0 = BRANCH, 1 = PREDICATED (as shipped) */
int oracle_accel(uint64_t n, const float *x, const float *y, const float *z, float eps, int method,
                 uint64_t i_begin, uint64_t i_end, float *ax, float *ay, float *az);
int oracle_accel_mass(uint64_t n, const float *x, const float *y, const float *z, const float *m, float eps,
                      uint64_t i_begin, uint64_t i_end, float *ax, float *ay, float *az);
int oracle_accel_f64(uint64_t n, const float *x, const float *y, const float *z, float eps,
                     uint64_t i_begin, uint64_t i_end, double *ax, double *ay, double *az);
int oracle_step(uint64_t n, float *x, float *y, float *z, float *vx, float *vy, float *vz, float G,
                float dt, float damping, float eps, int method, int iters);

double   oracle_time_accel(uint64_t n, const float *x, const float *y, const float *z, float eps,
                           uint64_t i_begin, uint64_t i_count, int reps);
int      oracle_num_threads(void);
void     oracle_set_num_threads(int n);
uint64_t oracle_fnv1a64(uint64_t n, int k, const float *const *arr);
uint64_t oracle_fnv1a64_state(uint64_t n, const float *x, const float *y, const float *z,
                              const float *vx, const float *vy, const float *vz);

#ifdef __cplusplus
}
#endif
#endif
