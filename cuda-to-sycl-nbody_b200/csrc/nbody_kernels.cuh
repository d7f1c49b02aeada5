// nbody_kernels.cuh -- launch interface of the sm_100a force+integrate kernels.
//
// One launch = for every i-body of a shard, accumulate the softened-gravity force of the
// j-bodies [j_begin, j_end) in ascending j with ONE FP32 accumulator per component
// (reference loop: src/simulator.cu:196-211), optionally carrying the accumulators in/out
// through `acc` so that a step can be cut into j-chunks without changing a single bit, and
// optionally finishing with the reference's velocity/position update (src/simulator.cu:213-228).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace nbody {

enum StepFlags : int {
  kFirstChunk = 1,  // accumulators start at +0 (else they are loaded from `acc`)
  kLastChunk = 2,   // integrate and write vel / pos_next (else accumulators go to `acc`)
  kAccelOut = 4,    // with kLastChunk: write raw force sums to `acc` instead of integrating
};

constexpr int kMaxPeers = 15;  // up to 16 GPUs per exchange group

// Host-side state of the j-segmented launches of one device (see force_wseg_kernel).
struct SegSync {
  unsigned int *words = nullptr;  // device: [0] = ticket counter, [1 + g] = hand-off word of body group g
  unsigned int *error = nullptr;  // device-visible flag (mapped pinned host word): a hand-off wait timed out
  uint32_t n_groups = 0;          // hand-off words available
  unsigned int epoch = 0;         // running segment counter: hand-off words are compared against epoch + seg
  unsigned int ticket_base = 0;   // value of the ticket counter when the next launch starts
  unsigned long long timeout_ns = 20000000000ull;  // hand-off wait time-out (0 = wait for ever)
};

struct StepArgs {
  const float4 *pos;  // n float4 (x,y,z,mass): j-bodies and the i-bodies' old positions
  float4 *pos_next;   // n float4; [i_begin, i_begin+i_count) written when integrating
  float4 *vel;        // i_count float4, shard-local index
  float4 *acc;        // i_count float4, shard-local index (carried sums / accel output)
  uint32_t n;
  uint32_t i_begin, i_count;
  uint32_t j_begin, j_end;
  float eps, dt, G, damping;
  int flags;
  // peer-push exchange: next-position replicas of the OTHER GPUs (NVLink peer / IPC mappings);
  // the integrate epilogue stores each new position to all of them
  int n_peers;
  float4 *peer_next[kMaxPeers];
  // j-segmented launches (host-side launch state, ignored by the kernels); nullptr: one segment
  SegSync *sync;
};

// self-term handling of the scalar kernels
enum SelfMode : int {
  kSelfNone = 0,             // no predicate: valid when rsqrt((0+eps)^3) is finite (self term adds +0)
  kSelfBranch = 1,           // skip j == i             (src/simulator.cu:206)
  kSelfPredicated = 2,       // multiply by (j == i)    (src/simulator.cu:209, as shipped)
  kSelfPredicatedFixed = 3,  // multiply by (j != i)    (README.md:229-231,247-250: what :209 was meant to be)
};

enum Family : int {
  kFamGeneric = 0,     // scalar, one body per lane, any self-term mode (always valid)
  kFamPackedCta = 1,   // comparison: CTA-tiled packed f32x2            (VARIANTS build only)
  kFamScalarCta = 2,   // comparison: CTA-tiled scalar, register blocked (VARIANTS build only)
  kFamUnsegmented = 3, // production kernel launched with a single j-segment
  kFamSegmented = 4,   // production: warp-streaming packed f32x2 with j-segmented hand-off
  kFamTma = 5,         // comparison: TMA-staged production kernel       (VARIANTS build only)
  kFamSmall = 6,       // small shards: scalar, one body per lane, no predicate
  kFamRelay = 7,       // smallest shards: block/32 warps relay the sums of the same 32 bodies over j-tiles of r bodies
};

struct KernelConfig {
  int family;
  int r;       // i-bodies per thread
  int block;   // threads per CTA
  int self_mode;
  int sms;     // SM count of the device the launch goes to (segment planning)
  int mass;    // 1: multiply each term by the j-body's mass (float4.w); extension, SURVEY 8(f)-3
};

// true when the unpredicated kernels reproduce the BRANCH result bit-for-bit for this eps:
// c = eps*(eps*eps) evaluated in FP32 with flush-to-zero must be a positive normal number.
bool eps_allows_unpredicated(float eps);

// picks the configuration for a shard of i_count bodies on a device with `sms` SMs
KernelConfig choose_config(int requested_kernel, int calc_method, float eps, uint32_t i_count,
                           int sms, bool has_mass);
const char *config_name(const KernelConfig &c, char *buf, size_t len);
// true when this build of the library contains the comparison kernels (make VARIANTS=1)
bool variants_built();

// asynchronous launch on `stream`; returns the CUDA error of the launch
cudaError_t launch_step(const KernelConfig &c, const StepArgs &a, cudaStream_t stream);
#ifdef NBODY_VARIANTS
cudaError_t launch_variant(const KernelConfig &c, const StepArgs &a, cudaStream_t stream);
#endif
// segment planning shared with the TMA variant
uint32_t plan_segments(uint32_t groups, uint32_t nj, int sms, int k_max);

// overwrite float4.w (the mass slot) of `count` bodies from m[] (device) or with the constant w
cudaError_t launch_set_w(float4 *pos, const float *m, float w, uint32_t count, cudaStream_t stream);
// float4 AoS -> three SoA arrays (read-back for the reference's ParticleData host layout)
cudaError_t launch_deinterleave(const float4 *src, float *x, float *y, float *z, uint32_t count,
                                cudaStream_t stream);
// three SoA arrays (+ optional mass, else w) -> float4 AoS
cudaError_t launch_interleave(const float *x, const float *y, const float *z, const float *m,
                              float w, float4 *dst, uint32_t count, cudaStream_t stream);

}  // namespace nbody
