/*
 * nbody_oracle.c -- CPU restatement of the reference's hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this file's library.  The product (cuda-to-sycl-nbody_b200/) never links,
 * imports or calls it; it has no CPU path at all.
 *
 * What is restated (reference = codeplaysoftware/cuda-to-sycl-nbody, paths below /root/reference):
 *   oracle_disk_galaxy      src/simulator.cu:131-158 (+ cross/length/normalize :165-181)
 *   oracle_accel            src/simulator.cu:196-211 (the j-loop of particle_interaction)
 *   oracle_step             src/simulator.cu:186-229 (whole kernel) x src/simulator.cu:57-66 (launch loop)
 *   the CPU-baseline role   src_sycl/simulator.dp.cpp:315-360 (same kernel as a SYCL parallel_for)
 *
 * Parity pinning: the reference ships no tests or golden vectors, so this
 * restatement is pinned against outputs of the reference's own code run on a B200
 * (oracle/_ref, built from the unmodified sources by oracle/Makefile) and committed as
 * tests/golden/ by tests/golden/make_golden.py:
 *   - the generator is pinned BIT-EXACTLY (host code, same libstdc++/glibc);
 *   - force sums and stepped states are pinned to a stated tolerance only, because the
 *     reference kernel is built with -use_fast_math and evaluates rsqrt with the GPU's
 *     MUFU.RSQ approximation, which no CPU reproduces bit-for-bit.  The bit-exact oracle of
 *     the CUDA path is therefore oracle/_ref itself (tests/test_parity_gpu.py).
 *
 * Arithmetic follows the op order of the reference kernel's sm_100a SASS (nvcc 12.9,
 * -O3 -use_fast_math): r = (-p_i) + p_j; t = ry*ry; t = fma(rx,rx,t); t = fma(rz,rz,t);
 * d = t + eps; c = d*(d*d); inv = rsqrt(c); a = fma(r, inv, a) with one accumulator per
 * component and j ascending; flush-to-zero on.  Build with -ffp-contract=off.
 */
#include "nbody_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#if defined(__x86_64__)
#include <xmmintrin.h>
#endif
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------------
 * std::mt19937 (default seed 5489) + libstdc++ std::uniform_real_distribution<double>(0,1),
 * i.e. std::generate_canonical<double,53>: two 32-bit draws, (lo + hi * 2^32) / 2^64.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  uint32_t s[624];
  int idx;
} mt19937_t;

static void mt_seed(mt19937_t *g, uint32_t seed) {
  g->s[0] = seed;
  for (int i = 1; i < 624; i++)
    g->s[i] = 1812433253u * (g->s[i - 1] ^ (g->s[i - 1] >> 30)) + (uint32_t)i;
  g->idx = 624;
}

static uint32_t mt_next(mt19937_t *g) {
  if (g->idx >= 624) {
    for (int k = 0; k < 624; k++) {
      uint32_t y = (g->s[k] & 0x80000000u) | (g->s[(k + 1) % 624] & 0x7fffffffu);
      g->s[k] = g->s[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    g->idx = 0;
  }
  uint32_t y = g->s[g->idx++];
  y ^= y >> 11;
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= y >> 18;
  return y;
}

static double mt_canonical(mt19937_t *g) {
  double lo = (double)mt_next(g);
  double hi = (double)mt_next(g);
  double r = (lo + hi * 4294967296.0) / 18446744073709551616.0;
  if (r >= 1.0) r = nextafter(1.0, 0.0);
  return r;
}

/* src/simulator.cu:131-158.  The reference file is compiled by nvcc, whose headers resolve
 * cos(float)/sin(float) to the single-precision functions; length() goes through
 * std::pow(float,int) -> double and std::sqrt(double). */
int oracle_disk_galaxy(uint64_t n, float *x, float *y, float *z, float *vx, float *vy, float *vz) {
  mt19937_t g;
  mt_seed(&g, 5489u);
  const float pi_f = 3.14159265358979323846f; /* src/simulator.cuh:35 */
  for (uint64_t i = 0; i < n; i++) {
    float t = (float)(mt_canonical(&g) * 2 * (double)pi_f);
    float s = (float)(mt_canonical(&g) * 100);
    x[i] = cosf(t) * s;
    y[i] = sinf(t) * s;
  }
  for (uint64_t i = 0; i < n; i++) z[i] = (float)(4.0 * mt_canonical(&g));

  for (uint64_t i = 0; i < n; i++) {
    /* cross((x,y,z), (0,0,1)), src/simulator.cu:165-168 */
    const float ux = 0.0f, uy = 0.0f, uz = 1.0f;
    float cx = y[i] * uz - z[i] * uy;
    float cy = z[i] * ux - x[i] * uz;
    float cz = x[i] * uy - y[i] * ux;
    /* length(): double pow/sqrt narrowed to coords_t, src/simulator.cu:170-172 */
    float len = (float)sqrt(pow((double)cx, 2) + pow((double)cy, 2) + pow((double)cz, 2));
    float orbital = (float)sqrt(2.0 * (double)len); /* src/simulator.cu:152 */
    vx[i] = (cx / len) * orbital;
    vy[i] = (cy / len) * orbital;
    vz[i] = (cz / len) * orbital;
  }
  return 0;
}

/* ------------------------------------------------------------------------------------------ */

/* FTZ | DAZ while the restated kernel runs (the reference kernel is all .ftz); the caller's
 * floating-point environment is restored afterwards (the master thread is the Python thread) */
static unsigned set_ftz(void) {
#if defined(__x86_64__)
  unsigned old = _mm_getcsr();
  _mm_setcsr(old | 0x8040u);
  return old;
#else
  return 0;
#endif
}
static void restore_csr(unsigned old) {
#if defined(__x86_64__)
  _mm_setcsr(old);
#else
  (void)old;
#endif
}

#define LANES 16

/* force sums for bodies [i0, i0+cnt) against j in [0, n) -- src/simulator.cu:196-211.
 * `method` is a compile-time constant in both instantiations below so the lane loop vectorises. */
static inline __attribute__((always_inline)) void
accel_block_impl(uint64_t n, const float *x, const float *y, const float *z, float eps,
                 const int method, uint64_t i0, int cnt, float *ax, float *ay, float *az) {
  float px[LANES], py[LANES], pz[LANES], fx[LANES], fy[LANES], fz[LANES];
  int32_t id[LANES];
  for (int l = 0; l < LANES; l++) {
    uint64_t i = i0 + (uint64_t)(l < cnt ? l : 0);
    px[l] = -x[i]; py[l] = -y[i]; pz[l] = -z[i];
    fx[l] = fy[l] = fz[l] = 0.0f;
    id[l] = (int32_t)i;
  }
  for (uint64_t j = 0; j < n; j++) {
    const float jx = x[j], jy = y[j], jz = z[j];
    const int32_t jj = (int32_t)j;
#pragma omp simd
    for (int l = 0; l < LANES; l++) {
      float rx = jx + px[l], ry = jy + py[l], rz = jz + pz[l];
      float t = ry * ry;
      t = fmaf(rx, rx, t);
      t = fmaf(rz, rz, t);
      float d = t + eps;
      float c = d * d;
      c = d * c;
      float inv = 1.0f / sqrtf(c); /* stands in for MUFU.RSQ */
      if (method == 0) {           /* BRANCH: `if (i == id) continue;` :206 */
        /* select, not multiply: the skipped term must not touch the accumulator */
        float nx = fmaf(rx, inv, fx[l]), ny = fmaf(ry, inv, fy[l]), nz = fmaf(rz, inv, fz[l]);
        fx[l] = (jj != id[l]) ? nx : fx[l];
        fy[l] = (jj != id[l]) ? ny : fy[l];
        fz[l] = (jj != id[l]) ? nz : fz[l];
      } else {                     /* PREDICATED as shipped: `* (i == id)` :209 */
        float sel = (jj == id[l]) ? 1.0f : 0.0f;
        fx[l] = fmaf(rx * inv, sel, fx[l]);
        fy[l] = fmaf(ry * inv, sel, fy[l]);
        fz[l] = fmaf(rz * inv, sel, fz[l]);
      }
    }
  }
  for (int l = 0; l < cnt; l++) { ax[l] = fx[l]; ay[l] = fy[l]; az[l] = fz[l]; }
}

#if defined(__x86_64__) && defined(__GNUC__) && !defined(__clang__)
#define ORACLE_CLONES __attribute__((target_clones("avx512f", "avx2,fma", "default")))
#else
#define ORACLE_CLONES
#endif

ORACLE_CLONES static void accel_block_branch(uint64_t n, const float *x, const float *y,
                                             const float *z, float eps, uint64_t i0, int cnt,
                                             float *ax, float *ay, float *az) {
  accel_block_impl(n, x, y, z, eps, 0, i0, cnt, ax, ay, az);
}
ORACLE_CLONES static void accel_block_pred(uint64_t n, const float *x, const float *y,
                                           const float *z, float eps, uint64_t i0, int cnt,
                                           float *ax, float *ay, float *az) {
  accel_block_impl(n, x, y, z, eps, 1, i0, cnt, ax, ay, az);
}

int oracle_accel(uint64_t n, const float *x, const float *y, const float *z, float eps, int method,
                 uint64_t i_begin, uint64_t i_end, float *ax, float *ay, float *az) {
  if (i_end > n || i_begin > i_end || n > 0x7fffffffull) return 1;
  int64_t nblk = (int64_t)((i_end - i_begin + LANES - 1) / LANES);
#pragma omp parallel
  {
    unsigned csr = set_ftz();
#pragma omp for schedule(dynamic, 4)
    for (int64_t b = 0; b < nblk; b++) {
      uint64_t i0 = i_begin + (uint64_t)b * LANES;
      int cnt = (int)((i_end - i0) < LANES ? (i_end - i0) : LANES);
      if (method == 0)
        accel_block_branch(n, x, y, z, eps, i0, cnt, ax + (i0 - i_begin), ay + (i0 - i_begin),
                           az + (i0 - i_begin));
      else
        accel_block_pred(n, x, y, z, eps, i0, cnt, ax + (i0 - i_begin), ay + (i0 - i_begin),
                         az + (i0 - i_begin));
    }
    restore_csr(csr);
  }
  return 0;
}

/* Extension (no reference counterpart: the reference is unit-mass, src/simulator.cu:204):
 * per-body masses.  Each term's weight is w*m_j, one FP32 multiply after the rsqrt, then the same
 * fma(r, w*m_j, a) accumulate -- the op order of the CUDA kernels' MASS variants.  BRANCH semantics. */
int oracle_accel_mass(uint64_t n, const float *x, const float *y, const float *z, const float *m, float eps,
                      uint64_t i_begin, uint64_t i_end, float *ax, float *ay, float *az) {
  if (i_end > n || i_begin > i_end) return 1;
#pragma omp parallel
  {
    unsigned csr = set_ftz();
#pragma omp for schedule(dynamic, 16)
    for (int64_t i = (int64_t)i_begin; i < (int64_t)i_end; i++) {
      float fx = 0.0f, fy = 0.0f, fz = 0.0f;
      const float px = -x[i], py = -y[i], pz = -z[i];
      for (uint64_t j = 0; j < n; j++) {
        if ((int64_t)j == i) continue;
        float rx = x[j] + px, ry = y[j] + py, rz = z[j] + pz;
        float t = ry * ry;
        t = fmaf(rx, rx, t);
        t = fmaf(rz, rz, t);
        float d = t + eps;
        float c = d * d;
        c = d * c;
        float w = (1.0f / sqrtf(c)) * m[j];
        fx = fmaf(rx, w, fx);
        fy = fmaf(ry, w, fy);
        fz = fmaf(rz, w, fz);
      }
      ax[i - (int64_t)i_begin] = fx; ay[i - (int64_t)i_begin] = fy; az[i - (int64_t)i_begin] = fz;
    }
    restore_csr(csr);
  }
  return 0;
}

/* FP64 truth for the same sum (self term skipped), for error-budget reporting only */
int oracle_accel_f64(uint64_t n, const float *x, const float *y, const float *z, float eps,
                     uint64_t i_begin, uint64_t i_end, double *ax, double *ay, double *az) {
  if (i_end > n || i_begin > i_end) return 1;
#pragma omp parallel for schedule(dynamic, 8)
  for (int64_t i = (int64_t)i_begin; i < (int64_t)i_end; i++) {
    double fx = 0, fy = 0, fz = 0;
    for (uint64_t j = 0; j < n; j++) {
      if ((int64_t)j == i) continue;
      double rx = (double)x[j] - x[i], ry = (double)y[j] - y[i], rz = (double)z[j] - z[i];
      double d = rx * rx + ry * ry + rz * rz + (double)eps;
      double inv = 1.0 / sqrt(d * d * d);
      fx += rx * inv; fy += ry * inv; fz += rz * inv;
    }
    ax[i - (int64_t)i_begin] = fx; ay[i - (int64_t)i_begin] = fy; az[i - (int64_t)i_begin] = fz;
  }
  return 0;
}

/* velocity/position update of src/simulator.cu:213-228 in the SASS op order:
 * t = F*dt; vd = v*damping; v' = fma(t, G, vd); x' = fma(v', dt, x). */
static inline void integrate1(float f, float *v, float *p, float dt, float G, float damping) {
  float t = f * dt;
  float vd = *v * damping;
  float vn = fmaf(t, G, vd);
  *v = vn;
  *p = fmaf(vn, dt, *p);
}

int oracle_step(uint64_t n, float *x, float *y, float *z, float *vx, float *vy, float *vz, float G,
                float dt, float damping, float eps, int method, int iters) {
  float *ax = (float *)malloc(3 * n * sizeof(float));
  if (!ax) return 2;
  float *ay = ax + n, *az = ay + n;
  for (int it = 0; it < iters; it++) {
    int rc = oracle_accel(n, x, y, z, eps, method, 0, n, ax, ay, az);
    if (rc) { free(ax); return rc; }
    unsigned csr = set_ftz();
    for (uint64_t i = 0; i < n; i++) {
      integrate1(ax[i], &vx[i], &x[i], dt, G, damping);
      integrate1(ay[i], &vy[i], &y[i], dt, G, damping);
      integrate1(az[i], &vz[i], &z[i], dt, G, damping);
    }
    restore_csr(csr);
  }
  free(ax);
  return 0;
}

/* CPU baseline timing (reported, never the target): forces of `i_count` bodies starting at
 * i_begin against all n, repeated `reps` times; returns seconds of the fastest repetition
 * measured the way src_sycl/simulator.dp.cpp:68-108 does (host steady clock around the work). */
double oracle_time_accel(uint64_t n, const float *x, const float *y, const float *z, float eps,
                         uint64_t i_begin, uint64_t i_count, int reps) {
  float *a = (float *)malloc(3 * i_count * sizeof(float));
  if (!a) return -1.0;
  double best = 1e300;
  for (int r = 0; r < reps; r++) {
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    oracle_accel(n, x, y, z, eps, 0, i_begin, i_begin + i_count, a, a + i_count, a + 2 * i_count);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    double s = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
    if (s < best) best = s;
  }
  free(a);
  return best;
}

/* launchers such as torchrun export OMP_NUM_THREADS=1; the CPU-baseline legs ask for all cores */
void oracle_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* FNV-1a-64 over the float bit patterns of k arrays, interleaved per body (a0[i], a1[i], ...) */
uint64_t oracle_fnv1a64(uint64_t n, int k, const float *const *arr) {
  uint64_t h = 1469598103934665603ull;
  for (uint64_t i = 0; i < n; i++)
    for (int a = 0; a < k; a++) {
      uint32_t b;
      memcpy(&b, &arr[a][i], 4);
      for (int s = 0; s < 4; s++) {
        h ^= (b >> (8 * s)) & 0xffu;
        h *= 1099511628211ull;
      }
    }
  return h;
}

uint64_t oracle_fnv1a64_state(uint64_t n, const float *x, const float *y, const float *z,
                              const float *vx, const float *vy, const float *vz) {
  const float *arr[6] = {x, y, z, vx, vy, vz};
  return oracle_fnv1a64(n, 6, arr);
}
