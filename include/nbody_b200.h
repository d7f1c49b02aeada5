/*
 * nbody_b200.h -- C ABI of the B200-native all-pairs N-body step.
 *
 * This is the drop-in boundary for the ONE hot path of codeplaysoftware/cuda-to-sycl-nbody:
 * the O(N^2) softened-gravity force accumulation fused with the damped semi-implicit Euler
 * update (reference: src/simulator.cu:186-229, driven by DiskGalaxySimulator::stepSim,
 * src/simulator.cu:47-75).  Every entry point names the reference interface it replaces.
 * The C++ mirror of the reference's class (simulation::DiskGalaxySimulator,
 * src/simulator.cuh:129-160) in cuda-to-sycl-nbody_b200/cxx/ is a thin wrapper over this ABI.
 *
 * Conventions
 *   - plain pointers and sizes only; no C++/torch types cross this boundary;
 *   - every function returns 0 on success, non-zero on failure (CUDA / NCCL error code or
 *     NBODY_E_*); nbody_last_error() returns a static description of the last failure of
 *     the calling thread;
 *   - a handle is NOT re-entrant: one host thread at a time (the reference is single
 *     threaded, src/nbody.cpp:31-141); different handles may be driven from different threads;
 *   - handles made by nbody_create_rank (one process per GPU) have COLLECTIVE calls, which every
 *     rank must make in the same order: nbody_create_rank, nbody_step, nbody_read_vel[_f4],
 *     nbody_read_state, nbody_compute_accel, nbody_save_state, nbody_destroy.  nbody_set_state,
 *     nbody_set_mass, nbody_read_pos[_f4] and nbody_read_local are local to the calling rank;
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef NBODY_B200_H_
#define NBODY_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NBODY_B200_ABI_VERSION 2

/* error codes beyond CUDA's (which are passed through unchanged, all < 10000) */
#define NBODY_E_INVALID 10001 /* bad argument                                  */
#define NBODY_E_NOGPU   10002 /* no CUDA device / driver                       */
#define NBODY_E_NCCL    10003 /* NCCL missing or NCCL call failed              */
#define NBODY_E_STATE   10004 /* call not valid in the handle's current state  */

/* calculation method: reference enum CalculationMethod, src/sim_param.hpp:8-11 */
#define NBODY_CALC_BRANCH     0
#define NBODY_CALC_PREDICATED 1 /* reproduces the shipped (i == id) behaviour, src/simulator.cu:208-209 */
/* opt-in extension: PREDICATED as the README describes it, force += r * inv * (i != id)
 * (README.md:229-231,247-250); bit-exact against the reference built with that one character changed
 * (oracle/Makefile: ref_fixed).  The C++ class selects it for calcMethod == PREDICATED when the
 * environment has NBODY_PREDICATED_FIXED=1. */
#define NBODY_CALC_PREDICATED_FIXED 2

/* kernel selection (new knob; the reference has one kernel) */
#define NBODY_KERNEL_AUTO    0 /* packed f32x2 kernel when bit-exactness allows, else generic */
#define NBODY_KERNEL_GENERIC 1 /* scalar, predicated self-term; always valid                  */
/* comparison variants, present only in lib/libnbody_b200_variants.so (make VARIANTS=1); the product
 * library rejects them with NBODY_E_INVALID */
#define NBODY_KERNEL_PACKED  2 /* CTA-tiled packed kernel (fails if eps makes it inexact)     */
#define NBODY_KERNEL_SCALAR  3 /* CTA-tiled register-blocked scalar FFMA kernel               */

/*
 * Mirror of the fields of the reference's SimParam that the hot path reads
 * (src/sim_param.hpp:30-39; numFrames is a main-loop field and is not part of the step).
 */
typedef struct nbody_params {
  float    G;                 /* SimParam::G                      */
  float    dt;                /* SimParam::dt                     */
  uint64_t num_particles;     /* SimParam::numParticles           */
  int32_t  iters_per_frame;   /* SimParam::simIterationsPerFrame  */
  float    damping;           /* SimParam::damping                */
  float    dist_eps;          /* SimParam::distEps (added to r^2) */
  int32_t  gw_size;           /* SimParam::gwSize  (advisory)     */
  int32_t  calc_method;       /* SimParam::calcMethod             */
} nbody_params;

typedef struct nbody_handle nbody_handle;

/* ---- library-level ------------------------------------------------------------------ */

int         nbody_abi_version(void);
const char *nbody_last_error(void);
/* number of visible CUDA devices (0 when there is no driver/GPU); never fails */
int         nbody_device_count(void);

/* reference defaults, src/sim_param.cpp:12-22 */
void        nbody_default_params(nbody_params *out);

/*
 * Host-only disk-galaxy generator: replaces DiskGalaxySimulator::randomParticlePos +
 * initialParticleVel (src/simulator.cu:131-158, helpers :165-181).  Bit-identical to the
 * reference's nvcc-compiled host code.  Needs no GPU.  All six arrays have n floats.
 */
int nbody_generate_disk_galaxy(uint64_t n, float *x, float *y, float *z,
                               float *vx, float *vy, float *vz);

/*
 * Host-only: the contiguous i-range [*begin, *begin + *count) that rank `rank` of `world` owns
 * for n bodies (shards are multiples of 128 bodies except the last).  New functionality: the
 * reference is single-GPU (device 0 hard-wired, src/simulator.cu:40).
 */
int nbody_plan_shard(uint64_t n, int world, int rank, uint64_t *begin, uint64_t *count);

/* ---- simulator object ---------------------------------------------------------------- */

/*
 * Replaces DiskGalaxySimulator::DiskGalaxySimulator(SimParam) (src/simulator.cu:24-34):
 * allocates device state on `n_gpus` devices of this process (devices 0..n_gpus-1, bodies
 * sharded by contiguous i-range, NCCL all-gather of positions per iteration), generates the
 * reference's default-seeded disk galaxy and uploads it.  n_gpus <= 0 means "read
 * NBODY_GPUS from the environment, default 1".
 */
int nbody_create(const nbody_params *p, int n_gpus, nbody_handle **out);

/*
 * One-process-per-GPU form of the same constructor (torchrun / MPI style launch): this
 * process owns rank `rank` of `world` and drives CUDA device `device`.  `nccl_unique_id`
 * is the 128-byte id from nbody_nccl_unique_id() of rank 0, distributed by the caller
 * (ignored when world == 1).  Collective over all ranks.
 */
int nbody_create_rank(const nbody_params *p, int device, int rank, int world,
                      const void *nccl_unique_id, nbody_handle **out);
int nbody_nccl_unique_id(void *out128);

/* frees device, pinned and NCCL resources (the reference never frees: no dtor) */
int nbody_destroy(nbody_handle *h);

/* kernel selection; must be called before the next nbody_step */
int nbody_set_kernel(nbody_handle *h, int kernel);
/* human-readable description of the kernel configuration the next step will use */
const char *nbody_kernel_name(nbody_handle *h);
/*
 * Host-only: the configuration AUTO picks for a shard of `shard_bodies` i-bodies on a device with `sms` SMs
 * (`has_mass`: per-body masses set).  Same naming as nbody_kernel_name() without the post-link suffix.  Needs no
 * GPU: lets a caller (and the CPU test-suite) see the kernel switch points.  The reference has one kernel and no
 * such choice (src/simulator.cu:60-66 launches particle_interaction<> with gwSize threads whatever N is).
 */
int nbody_describe_auto(const nbody_params *p, uint64_t shard_bodies, int sms, int has_mass, char *buf, size_t len);

/*
 * Replaces DiskGalaxySimulator::sendToDevice (src/simulator.cu:79-103) for caller-provided
 * state: six host arrays of num_particles floats (full N on every rank).
 */
int nbody_set_state(nbody_handle *h, const float *x, const float *y, const float *z,
                    const float *vx, const float *vy, const float *vz);

/* optional per-body masses (float4.w of the position array).  The reference is unit-mass
 * (src/simulator.cu:204); passing NULL restores unit masses and the exact reference path. */
int nbody_set_mass(nbody_handle *h, const float *m);

/*
 * Replaces the device part of DiskGalaxySimulator::stepSim (src/simulator.cu:47-72):
 * iters_per_frame fused force+integrate iterations, then blocks until the device is idle.
 * Afterwards nbody_last_step_ms() == the reference's lastStepTime (host steady_clock from
 * before the first launch to after the synchronize) and nbody_last_step_device_ms() is the
 * same span measured with CUDA events on the compute stream (max over local devices).
 */
int   nbody_step(nbody_handle *h);
float nbody_last_step_ms(nbody_handle *h);
float nbody_last_step_device_ms(nbody_handle *h);
/* kernels launched by this library on behalf of the handle since creation */
uint64_t nbody_launch_count(nbody_handle *h);

/*
 * Replaces DiskGalaxySimulator::recvFromDevice + getParticlePos/getParticleVel
 * (src/simulator.cu:106-129, :160-162): SoA read-back into caller memory, n floats each.
 * In rank mode every rank receives all N positions; velocities of other ranks' bodies are
 * gathered on demand.
 */
int nbody_read_pos(nbody_handle *h, float *x, float *y, float *z);
int nbody_read_vel(nbody_handle *h, float *vx, float *vy, float *vz);
/* AoS float4 (x,y,z,w) read-back: the layout RendererGL::setParticleData builds by hand
 * (src/renderer_gl.cpp:156-172).  w = mass for positions (1.0f, as the renderer writes),
 * 0 for velocities. */
int nbody_read_pos_f4(nbody_handle *h, float *xyzw);
int nbody_read_vel_f4(nbody_handle *h, float *xyzw);

/* positions and velocities in one call (one de-interleave pass per array, the device->host copies on
 * a copy stream overlapping the next de-interleave): what stepSim() hands to the host every frame,
 * src/simulator.cu:74 */
int nbody_read_state(nbody_handle *h, float *x, float *y, float *z, float *vx, float *vy, float *vz);
/* the bodies [begin, begin + count) owned by the devices of this handle (all of them for nbody_create;
 * one rank's shard for nbody_create_rank) and their positions + velocities, count floats per array:
 * the read-back a rank needs when every process keeps only its own bodies.  Not collective. */
int nbody_local_range(nbody_handle *h, uint64_t *begin, uint64_t *count);
int nbody_read_local(nbody_handle *h, float *x, float *y, float *z, float *vx, float *vy, float *vz);
/* page-lock caller memory (cudaHostRegister) so nbody_set_state / nbody_read_* move it by DMA instead
 * of through the driver's pageable staging; the C++ class registers its ParticleData vectors once
 * (they never reallocate, src/simulator.cuh:75-84).  Unregister before freeing the memory. */
int nbody_host_register(void *ptr, size_t bytes);
int nbody_host_unregister(void *ptr);

/*
 * Checkpoint / resume (the reference has none).  Little-endian binary file:
 *   char magic[8] = "NBB200\0\1"; uint64 n; float G, dt, damping, dist_eps; int32 iters_per_frame,
 *   calc_method, has_mass, reserved;  then float32[n] x, y, z, vx, vy, vz (and m when has_mass).
 * Loading restores positions, velocities and masses into a handle created for the same n (the
 * handle's own SimParam stays in force); stepping then continues bit-identically.
 */
int nbody_save_state(nbody_handle *h, const char *path);
int nbody_load_state(nbody_handle *h, const char *path);

/* replaces DiskGalaxySimulator::getDeviceName (src/simulator.cu:36-45) */
const char *nbody_device_name(nbody_handle *h);
uint64_t    nbody_num_particles(nbody_handle *h);
int         nbody_num_gpus(nbody_handle *h); /* local devices driven by this handle */
int         nbody_world_size(nbody_handle *h);

/*
 * Test hook: raw force sums sum_j r_ij * rsqrt((r.r + eps)^3) of the CURRENT state, in the
 * reference's accumulation order, without touching the state.  (The reference exposes the
 * same numbers only through the damping=0, dt=1, G=1 trick.)  n floats each, full N.
 */
int nbody_compute_accel(nbody_handle *h, float *ax, float *ay, float *az);

/* ---- stateless device-pointer entry (for callers that own device memory) ----------- */

/*
 * One fused force+integrate launch on caller-owned DEVICE memory, asynchronous on
 * `cuda_stream` (a cudaStream_t, 0 = default stream) of the current device.
 *   pos4      : n_bodies float4 (x,y,z,mass) -- j-bodies and the i-bodies' old positions
 *   vel4      : float4 velocity of bodies [i_begin, i_begin+i_count), indexed from 0
 *   pos4_next : n_bodies float4, entries [i_begin, i_begin+i_count) are written
 * Bodies j in [0, n_bodies) are accumulated in ascending order with one FP32 accumulator per
 * component, exactly as src/simulator.cu:196-211 does.  flags: 0, or NBODY_DEVSTEP_MASS when pos4[].w
 * holds per-body masses that are not all 1 (without it w is ignored: the reference is unit-mass).
 */
#define NBODY_DEVSTEP_MASS 1
int nbody_launch_step_device(const nbody_params *p, const void *pos4, void *vel4,
                             void *pos4_next, uint64_t i_begin, uint64_t i_count,
                             int kernel, int flags, void *cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* NBODY_B200_H_ */
