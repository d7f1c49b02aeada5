// nbody_kernels.cu -- the production sm_100a kernels of the all-pairs force + integrate step and
// their host-side dispatch.  The arithmetic lives in nbody_body.cuh (one definition of the packed
// 12-op body, one of the scalar body, shared by every kernel); this file decides how it is fed:
//
//   * force_wseg_kernel<R, MINB, MASS>  (AUTO: R = 6 / 4 / 2 by shard size, see choose_config)
//     one warp per CTA, R i-bodies per lane in f32x2 pairs, warp-private 32-body j-tiles (LDG.128 ->
//     STS.128 -> __syncwarp -> broadcast LDS.128, double buffered, no CTA barrier), and the j-sweep
//     of a body group cut into consecutive work units of ONE grid that hand the accumulators on
//     through L2 in order -- still one FP32 chain per body over ascending j (bit-exact), but short
//     units, which removes the low-occupancy tail of the launch;
//   * force_wscalar_kernel<R, SELF, MASS>
//     one warp per CTA, scalar ops, one (or two) bodies per lane: small shards (lanes, not issue
//     slots, are scarce there) and the generic/faithful path with the reference's self-term variants
//     (BRANCH predicate for any eps, PREDICATED as shipped, PREDICATED as the README intends);
//   * force_wrelay_kernel<W, TJ, MINB, MASS>  (AUTO below 128 bodies per SM)
//     W warps per CTA serve the SAME 32 i-bodies and take the j-tiles round-robin: a warp computes the differences
//     and weights of its tile into registers (10 of the 13 operations, no dependence on the running sums), then
//     receives the sums from the warp of the previous tile (shared memory + mbarrier), runs the accumulate FMAs in
//     ascending j and passes them on -- more warps out of the same bodies without touching the summation order;
//   * no warp shuffles, no atomics, no j-split reduction in the accumulate: a per-thread FMA chain.
// Comparison kernels (CTA-tiled, TMA-staged) live in nbody_variants.cu and are only built with
// `make VARIANTS=1`.  Measurements behind every choice: profiles/, DESIGN.md section 5.
#include "nbody_kernels.cuh"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <mutex>

#include "nbody_body.cuh"

// Resident warps per SM promised to ptxas (= register budget) per register-blocking factor.  The unrolled tile
// body is regenerated after linking (tools/sass_gen.py) with its own register allocation: double-buffered
// differences and weights need ~44 registers next to the negated positions and accumulators, so every
// instantiation gets a 128-register budget (16 resident warps per SM; the generated single-warp schedule keeps a
// sub-partition's issue port busy with 3-4 warps, profiles/r02_sched_sweep.txt).
#ifndef NB_MINB2
#define NB_MINB2 16
#endif
#ifndef NB_MINB4
#define NB_MINB4 16
#endif
#ifndef NB_MINB6
#define NB_MINB6 14
#endif
// the per-body-mass instantiations (one more FMUL2 per pair-unit) are regenerated the same way
#ifndef NB_MINB2M
#define NB_MINB2M NB_MINB2
#endif
#ifndef NB_MINB4M
#define NB_MINB4M NB_MINB4
#endif
#ifndef NB_MINB6M
#define NB_MINB6M NB_MINB6
#endif
// AUTO switch points in bodies per SM (see choose_config)
#ifndef NB_SW6
#define NB_SW6 3100u
#endif
#ifndef NB_SW4
#define NB_SW4 1650u
#endif
// accumulator relay (smallest shards): up to 3 / 4 resident CTAs of 32 bodies per SM
#ifndef NB_RELAY32
#define NB_RELAY32 96u
#endif
#ifndef NB_RELAY16
#define NB_RELAY16 128u
#endif

namespace nbody {

// j-segmented launch.  A body group's sweep over j is cut into `segs` consecutive segments that are
// separate CTAs of the SAME grid: unit (seg, g) handles group g over segment seg, takes the
// accumulators from `acc` and hands them on through `acc`, in order -- so the per-body sum is still
// one FP32 chain over ascending j (bit-exact), but the schedulable unit is `segs` times shorter.
// Why: with equal units the launch ends with every SM sub-partition holding ~W/2 units of
// leftovers that finish one by one at falling occupancy; that tail costs ~0.46 unit-times
// (3.3 % at N = 1M, 6.6 % at 512K bodies, measured) and shrinks in proportion to the unit.
// (The register budget is fixed through MINB because ptxas' schedule is sensitive to it:
// +-6 % between neighbouring budgets, profiles/r01_tuning_log.txt.)
template <int R, int MINB, bool MASS>
__global__ void __launch_bounds__(32, MINB)
    force_wseg_kernel(const StepArgs a, const uint32_t groups, const uint32_t segs, const uint32_t seg_len,
                      unsigned int *words, unsigned int *error, const unsigned int epoch,
                      const unsigned int ticket_base, const unsigned long long timeout_ns) {
  __shared__ __align__(16) float4 s_tile[2][32];
  const int lane = threadIdx.x & 31;
#ifdef NB_SMEM_PAD  // tuning builds only (tools/sass_lab.py): extra static shared memory caps the resident warps per SM
  __shared__ float s_pad[NB_SMEM_PAD / 4];
  if (a.n == 0xffffffffu) ((volatile float *)s_pad)[lane] = 0.0f;
#endif
  uint32_t unit = blockIdx.x;
  if (segs > 1) {  // take a ticket: units are numbered in the order CTAs actually start
    unsigned int t = 0;
    if (lane == 0) t = atomicAdd(words, 1u) - ticket_base;
    unit = __shfl_sync(0xffffffffu, t, 0);
  }
  const uint32_t seg = unit / groups;
  const uint32_t g = unit - seg * groups;
  const uint32_t warp_i = g * (uint32_t)(32 * R);
  const uint32_t j_begin = a.j_begin + seg * seg_len;
  const uint32_t j_end = min(a.j_end, j_begin + seg_len);
  const int flags = (seg == 0 ? (a.flags & kFirstChunk) : 0) | (seg == segs - 1 ? (a.flags & (kLastChunk | kAccelOut)) : 0);
  if (seg > 0) {
    if (lane == 0) wait_for_segment(words + 1 + g, epoch + seg, error, timeout_ns);
    __syncwarp();
  }
  warp_sweep_packed<R, MASS>(a, j_begin, j_end, flags, warp_i, s_tile, lane);
  if (seg + 1 < segs) {
    __threadfence();  // every lane publishes its accumulator stores ...
    __syncwarp();
    if (lane == 0) atomicExch(words + 1 + g, epoch + seg + 1);  // ... before the group is handed on
  }
}

// scalar warp-streaming kernel: small shards (SELF = none) and the generic path (any self-term mode)
template <int R, int SELF, bool MASS>
__global__ void __launch_bounds__(32) force_wscalar_kernel(const StepArgs a) {
  __shared__ __align__(16) float4 s_tile[2][32];
  const int lane = threadIdx.x & 31;
  const uint32_t warp_i = blockIdx.x * (uint32_t)(32 * R);
  if (warp_i >= a.i_count) return;
  warp_sweep_scalar<R, SELF, MASS>(a, a.j_begin, a.j_end, a.flags, warp_i, s_tile, lane);
}

// accumulator relay: W warps per CTA serve the same 32 i-bodies, j-tiles of TJ bodies round-robin (cta_relay_scalar)
template <int W, int TJ, int MINB, bool MASS>
__global__ void __launch_bounds__(32 * W, MINB) force_wrelay_kernel(const StepArgs a) {
  __shared__ __align__(16) float4 s_tile[W][TJ];
  __shared__ __align__(16) float4 s_tok[32];
  __shared__ __align__(8) u64 s_bar[W];
  if (blockIdx.x * 32u >= a.i_count) return;
  cta_relay_scalar<W, TJ, MASS>(a, s_tile, s_tok, s_bar);
}

// ---- layout helpers -------------------------------------------------------------------------
__global__ void deinterleave_kernel(const float4 *__restrict__ src, float *__restrict__ x,
                                    float *__restrict__ y, float *__restrict__ z, uint32_t count) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  float4 v = src[i];
  x[i] = v.x;
  y[i] = v.y;
  z[i] = v.z;
}
__global__ void interleave_kernel(const float *__restrict__ x, const float *__restrict__ y,
                                  const float *__restrict__ z, const float *__restrict__ m, float w,
                                  float4 *__restrict__ dst, uint32_t count) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  dst[i] = make_float4(x[i], y[i], z[i], m ? m[i] : w);
}

__global__ void set_w_kernel(float4 *__restrict__ pos, const float *__restrict__ m, float w, uint32_t count) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  pos[i].w = m ? m[i] : w;
}

cudaError_t launch_set_w(float4 *pos, const float *m, float w, uint32_t count, cudaStream_t stream) {
  if (count == 0) return cudaSuccess;
  set_w_kernel<<<(count + 255) / 256, 256, 0, stream>>>(pos, m, w, count);
  return cudaGetLastError();
}

cudaError_t launch_deinterleave(const float4 *src, float *x, float *y, float *z, uint32_t count,
                                cudaStream_t stream) {
  if (count == 0) return cudaSuccess;
  deinterleave_kernel<<<(count + 255) / 256, 256, 0, stream>>>(src, x, y, z, count);
  return cudaGetLastError();
}
cudaError_t launch_interleave(const float *x, const float *y, const float *z, const float *m,
                              float w, float4 *dst, uint32_t count, cudaStream_t stream) {
  if (count == 0) return cudaSuccess;
  interleave_kernel<<<(count + 255) / 256, 256, 0, stream>>>(x, y, z, m, w, dst, count);
  return cudaGetLastError();
}

// ---- host side: configuration and dispatch ---------------------------------------------------

// FP32 flush-to-zero product as the GPU evaluates c = d*(d*d) for r = 0, d = 0 + eps
static float ftz(float v) { return (fpclassify(v) == FP_SUBNORMAL) ? copysignf(0.0f, v) : v; }

bool eps_allows_unpredicated(float eps) {
  volatile float d = ftz(0.0f + ftz(eps));
  volatile float dd = ftz(d * d);
  volatile float c = ftz(d * dd);
  // rsqrt(c) must be finite and the self term fma(0, w, a) == a: c positive, normal, finite
  return isfinite(c) && c > 0.0f && fpclassify(c) == FP_NORMAL;
}

bool variants_built() {
#ifdef NBODY_VARIANTS
  return true;
#else
  return false;
#endif
}

KernelConfig choose_config(int requested_kernel, int calc_method, float eps, uint32_t i_count,
                           int sms, bool has_mass) {
  KernelConfig c;
  const bool exact_unpred = eps_allows_unpredicated(eps);
  if (calc_method == 1) return KernelConfig{kFamGeneric, 1, 32, kSelfPredicated, sms, has_mass ? 1 : 0};       // as shipped
  if (calc_method == 2) return KernelConfig{kFamGeneric, 1, 32, kSelfPredicatedFixed, sms, has_mass ? 1 : 0};  // README intent
  if (requested_kernel == 1 /*GENERIC*/ || !exact_unpred) return KernelConfig{kFamGeneric, 1, 32, kSelfBranch, sms, has_mass ? 1 : 0};
  int family = requested_kernel == 3 ? kFamScalarCta : (requested_kernel == 2 ? kFamPackedCta : kFamSegmented);
  int r = 4, block = 128;
  if (family == kFamSegmented) {
    // one warp per CTA.  What decides is how many warps the shard yields: with the generated tile body a warp's
    // cost per pair-interaction hardly depends on R (26.1 cycles at R = 6, 26.2 at R = 4, 27.0 at R = 2: the
    // LDS.128 and the tile bookkeeping are shared by R/2 pairs), but every one of the 4 x SMs sub-partitions
    // needs about three warps before the hand-off segments balance the machine.  Measured on 148 SMs
    // (profiles/r02_sched_sweep.txt, section 3):
    //   >= NB_SW6 bodies per SM (459K) : R = 6     76.7 % of the FP32 roofline at N = 1M
    //   >= NB_SW4 (244K)               : R = 4     74.8 % at N = 262 144, 76.3 % at 400 003
    //   >= 768 (114K)                  : R = 2     71.7 % at N = 131 072
    //   below: the sub-partitions hold 0-2 warps each and the quantisation decides -- pick the cheapest of
    //   {scalar one-body-per-lane, R = 2, R = 4} by  ceil(warps / sub-partitions) x bodies per warp / efficiency
    //   of a sub-partition holding that many warps of that kind (table below, from the same sweep).
    block = 32;
    r = 6;
    if ((uint64_t)i_count < (uint64_t)sms * NB_SW6) r = 4;
    if ((uint64_t)i_count < (uint64_t)sms * NB_SW4) r = 2;
    if ((uint64_t)i_count <= (uint64_t)sms * NB_RELAY16) {
      // smallest shards (at most 4 groups of 32 bodies per SM): even the scalar kernel leaves each sub-partition
      // with one warp at 0.6 instructions per cycle.  The accumulator relay turns one group into 4 warps without
      // changing the summation order (cta_relay_scalar): measured 1.25x at N = 12 800, 1.6-1.8x at 2K-6.4K bodies,
      // 1.13x at 18 944 (profiles/r02_relay.txt).  32-body tiles need 164 registers = 3 CTAs per SM, 16-body tiles 108 = 4.
      family = kFamRelay;
      block = 128;
      r = (uint64_t)i_count <= (uint64_t)sms * NB_RELAY32 ? 32 : 16;
    } else if ((uint64_t)i_count < (uint64_t)sms * 768u) {
      static const double eff[3][4] = {{0.376, 0.47, 0.53, 0.57},   // scalar, 1 / 2 / 3 / >= 4 warps per sub-partition
                                       {0.58, 0.66, 0.72, 0.74},    // R = 2
                                       {0.68, 0.74, 0.755, 0.763}}; // R = 4
      static const int bodies[3] = {32, 64, 128};
      const uint64_t smsp = 4ull * (uint64_t)sms;
      double best = 0.0;
      int pick = 0;
      for (int k = 0; k < 3; k++) {
        const uint64_t warps = ((uint64_t)i_count + bodies[k] - 1) / bodies[k];
        const uint64_t w = (warps + smsp - 1) / smsp;
        const double cost = (double)w * bodies[k] / eff[k][w >= 4 ? 3 : (w < 1 ? 0 : w - 1)];
        if (k == 0 || cost < best * 0.999) {
          best = cost;
          pick = k;
        }
      }
      if (pick == 0) {
        family = kFamSmall;
        r = 1;
      } else {
        r = pick == 1 ? 2 : 4;
      }
    }
  } else {
    // CTA-tiled comparison kernels: i-bodies per CTA = block*r; keep ~4 CTA-tiles per SM
    if ((uint64_t)i_count < (uint64_t)sms * 4u * 512u) r = 2;
    if ((uint64_t)i_count < (uint64_t)sms * 4u * 256u) block = 64;
  }
  // tuning override for sweeps: NBODY_KERNEL_CONFIG="r,block[,family]" (exact unpredicated families)
  if (const char *e = getenv("NBODY_KERNEL_CONFIG")) {
    int er = 0, eb = 0, ef = 0;
    int got = sscanf(e, "%d,%d,%d", &er, &eb, &ef);
    if (got >= 2 && er > 0 && eb > 0) {
      r = er;
      block = eb;
      if (got == 3 && ef >= 1 && ef <= 7) family = ef;
    }
  }
  c = {family, r, block, kSelfNone, sms, has_mass ? 1 : 0};
  return c;
}

const char *config_name(const KernelConfig &c, char *buf, size_t len) {
  static const char *fam[] = {"generic_scalar", "cta_packed_f32x2", "cta_scalar", "wstream_f32x2", "wseg_f32x2", "wseg_tma_f32x2", "wsmall_scalar", "wrelay_scalar"};
  static const char *self[] = {"nopred", "branch", "predicated", "predicated_fixed"};
  snprintf(buf, len, "%s_r%d_b%d_%s%s", fam[c.family >= 0 && c.family <= 7 ? c.family : 0], c.r, c.block,
           self[c.self_mode >= 0 && c.self_mode <= 3 ? c.self_mode : 0], c.mass ? "_mass" : "");
  return buf;
}

// resident CTAs per SM of a kernel on the current device, cached per (kernel, device); the cache is
// shared by every handle of the process, hence the lock
static cudaError_t resident_ctas(const void *kern, int threads, int *out) {
  struct Slot { const void *kern; int dev; int k; };
  static Slot slots[128];
  static int n_slots = 0;
  static std::mutex mu;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  std::lock_guard<std::mutex> lock(mu);
  for (int i = 0; i < n_slots; i++)
    if (slots[i].kern == kern && slots[i].dev == dev) {
      *out = slots[i].k;
      return cudaSuccess;
    }
  int k = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&k, kern, threads, 0);
  if (e != cudaSuccess) return e;
  if (k < 1) k = 1;
  if (n_slots < 128) slots[n_slots++] = Slot{kern, dev, k};
  *out = k;
  return cudaSuccess;
}

// segment planning: aim at ~16 schedulable units per resident warp slot
uint32_t plan_segments(uint32_t groups, uint32_t nj, int sms, int k_max) {
  if (const char *env = getenv("NBODY_SEGS")) {  // tuning override
    int v = atoi(env);
    if (v >= 1 && v <= 1024) return (uint32_t)v;
  }
  const double gens = (double)groups / ((double)sms * k_max);
  if (gens < 0.4) return 1;  // far fewer warps than slots: later segments would only spin (measured, N < ~100K at R = 2)
  int segs = (int)ceil(16.0 / (gens > 0.05 ? gens : 0.05));
  if (segs > 64) segs = 64;
  while (segs > 1 && nj / segs < 4096) segs--;  // keep segments long: hand-off cost stays invisible
  return (uint32_t)(segs < 1 ? 1 : segs);
}

template <int R, int MINB, bool MASS>
static cudaError_t launch_wseg(const StepArgs &a, int sms, bool segmented, cudaStream_t s) {
  const uint32_t groups = (a.i_count + 32 * R - 1) / (32 * R);
  const uint32_t nj = a.j_end - a.j_begin;
  SegSync *sy = a.sync;
  uint32_t segs = 1, seg_len = (nj + 31u) / 32u * 32u;
  cudaError_t e;
  if (segmented && sy && sy->words && groups <= sy->n_groups && nj > 0) {
    int k_max = 1;
    if ((e = resident_ctas((const void *)force_wseg_kernel<R, MINB, MASS>, 32, &k_max)) != cudaSuccess) return e;
    segs = plan_segments(groups, nj, sms, k_max);
    seg_len = ((nj + segs - 1) / segs + 31u) / 32u * 32u;
    segs = (nj + seg_len - 1) / seg_len;
  }
  unsigned int ep = 0, tb = 0;
  if (segs > 1) {
    if (sy->epoch > 0xf0000000u) {
      // hand-off words are compared for equality with epoch + seg: restart the numbering long before
      // it wraps (stream order makes the reset safe: every earlier launch has drained)
      if ((e = cudaMemsetAsync(sy->words + 1, 0, (size_t)sy->n_groups * sizeof(unsigned int), s)) != cudaSuccess) return e;
      sy->epoch = 0;
    }
    ep = sy->epoch;
    tb = sy->ticket_base;
    sy->epoch += segs;
    sy->ticket_base += groups * segs;  // wraps with the device counter (unsigned arithmetic on both sides)
  }
  force_wseg_kernel<R, MINB, MASS><<<groups * segs, 32, 0, s>>>(a, groups, segs, seg_len, sy ? sy->words : nullptr,
                                                                sy ? sy->error : nullptr, ep, tb, sy ? sy->timeout_ns : 0ull);
  e = cudaGetLastError();
  if (e != cudaSuccess && segs > 1) {  // nothing ran: keep host and device numbering in step
    sy->epoch -= segs;
    sy->ticket_base -= groups * segs;
  }
  return e;
}

template <int R, int SELF, bool MASS>
static cudaError_t launch_wscalar(const StepArgs &a, cudaStream_t s) {
  const uint32_t per = 32u * (uint32_t)R;
  force_wscalar_kernel<R, SELF, MASS><<<(a.i_count + per - 1) / per, 32, 0, s>>>(a);
  return cudaGetLastError();
}

template <int W, int TJ, int MINB>
static cudaError_t launch_wrelay(const StepArgs &a, bool mass, cudaStream_t s) {
  const uint32_t ctas = (a.i_count + 31u) / 32u;
  if (mass)
    force_wrelay_kernel<W, TJ, MINB, true><<<ctas, 32 * W, 0, s>>>(a);
  else
    force_wrelay_kernel<W, TJ, MINB, false><<<ctas, 32 * W, 0, s>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_step(const KernelConfig &c, const StepArgs &a, cudaStream_t s) {
  if (a.i_count == 0) return cudaSuccess;
  const bool m = c.mass != 0;
  if (c.family == kFamGeneric) {
    switch (c.self_mode) {
      case kSelfPredicated: return m ? launch_wscalar<1, kSelfPredicated, true>(a, s) : launch_wscalar<1, kSelfPredicated, false>(a, s);
      case kSelfPredicatedFixed: return m ? launch_wscalar<1, kSelfPredicatedFixed, true>(a, s) : launch_wscalar<1, kSelfPredicatedFixed, false>(a, s);
      default: return m ? launch_wscalar<1, kSelfBranch, true>(a, s) : launch_wscalar<1, kSelfBranch, false>(a, s);
    }
  }
  if (c.family == kFamSmall) {
    if (c.r == 1) return m ? launch_wscalar<1, kSelfNone, true>(a, s) : launch_wscalar<1, kSelfNone, false>(a, s);
    if (c.r == 2) return m ? launch_wscalar<2, kSelfNone, true>(a, s) : launch_wscalar<2, kSelfNone, false>(a, s);
    return cudaErrorInvalidConfiguration;
  }
  if (c.family == kFamRelay) {  // (warps per CTA, j-bodies per tile): register budget = resident CTAs promised to ptxas
    if (c.block == 128 && c.r == 16) return launch_wrelay<4, 16, 4>(a, m, s);
    if (c.block == 128 && c.r == 32) return launch_wrelay<4, 32, 3>(a, m, s);
#ifdef NBODY_VARIANTS  // shapes AUTO never picks (measured slower, profiles/r02_relay.txt): comparison build only
    if (c.block == 64 && c.r == 16) return launch_wrelay<2, 16, 8>(a, m, s);
    if (c.block == 64 && c.r == 32) return launch_wrelay<2, 32, 6>(a, m, s);
    if (c.block == 256 && c.r == 16) return launch_wrelay<8, 16, 2>(a, m, s);
#endif
    return cudaErrorInvalidConfiguration;
  }
  if (c.family == kFamSegmented || c.family == kFamUnsegmented) {
    const bool seg = c.family == kFamSegmented;
    // (R, resident warps per SM promised to ptxas): the tuned points, profiles/r01_tuning_log.txt and r02_sched_sweep.txt
    if (c.r == 2) return m ? launch_wseg<2, NB_MINB2M, true>(a, c.sms, seg, s) : launch_wseg<2, NB_MINB2, false>(a, c.sms, seg, s);
    if (c.r == 4) return m ? launch_wseg<4, NB_MINB4M, true>(a, c.sms, seg, s) : launch_wseg<4, NB_MINB4, false>(a, c.sms, seg, s);
    if (c.r == 6) return m ? launch_wseg<6, NB_MINB6M, true>(a, c.sms, seg, s) : launch_wseg<6, NB_MINB6, false>(a, c.sms, seg, s);
#ifdef NB_MINB8
    if (c.r == 8 && !m) return launch_wseg<8, NB_MINB8, false>(a, c.sms, seg, s);  // tuning builds only
#endif
    return cudaErrorInvalidConfiguration;
  }
#ifdef NBODY_VARIANTS
  return launch_variant(c, a, s);
#else
  return cudaErrorInvalidConfiguration;  // comparison kernels: make VARIANTS=1
#endif
}

}  // namespace nbody
