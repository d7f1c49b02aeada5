// nbody_variants.cu -- COMPARISON kernels, built only with `make VARIANTS=1` (-DNBODY_VARIANTS) into
// lib/libnbody_b200_variants.so.  They share the arithmetic of nbody_body.cuh with the production
// kernels (so they are bit-identical) and differ only in how the j-bodies are fed; they are what
// the production design was measured against (profiles/r01_tuning_log.txt):
//   family 1  force_packed_kernel<R, BLOCK>     CTA-wide 256-body tiles, one __syncthreads per tile, f32x2
//   family 2  force_scalar_kernel<R, BLOCK>     the same with scalar FFMA, register blocked
//   family 5  force_wseg_tma_kernel<R, MINB>    production kernel with cp.async.bulk + mbarrier tile staging
// Select with nbody_set_kernel(NBODY_KERNEL_PACKED / _SCALAR) or NBODY_KERNEL_CONFIG="r,block,family".
#ifdef NBODY_VARIANTS
#include "nbody_body.cuh"
#include "nbody_kernels.cuh"

namespace nbody {

// CTA-tiled j-loop: every thread fetches one j-body per tile; `body(q, gj)` for all j ascending
template <int TJ, class F>
__device__ __forceinline__ void sweep_cta_tiles(const StepArgs &a, float4 (*s_p)[TJ], const int tid, F body) {
  const uint32_t nj = a.j_end - a.j_begin;
  const uint32_t ntiles = (nj + TJ - 1) / TJ;
  auto fetch = [&](uint32_t t) -> float4 {
    uint32_t j = a.j_begin + t * TJ + tid;
    return a.pos[j < a.j_end ? j : a.j_end - 1];
  };
  if (ntiles > 0) s_p[0][tid] = fetch(0);
  __syncthreads();
  for (uint32_t t = 0; t < ntiles; t++) {
    const int buf = t & 1;
    float4 nxt;
    const bool more = t + 1 < ntiles;
    if (more) nxt = fetch(t + 1);
    const uint32_t cnt = min((uint32_t)TJ, nj - t * TJ);
    const uint32_t gj0 = a.j_begin + t * TJ;
    if (cnt == TJ) {
#pragma unroll 32
      for (int j = 0; j < TJ; j++) body(s_p[buf][j], gj0 + j);
    } else {
      for (uint32_t j = 0; j < cnt; j++) body(s_p[buf][j], gj0 + j);
    }
    if (more) s_p[buf ^ 1][tid] = nxt;
    __syncthreads();
  }
}

template <int R, int BLOCK>
__global__ void __launch_bounds__(BLOCK) force_packed_kernel(const StepArgs a) {
  static_assert(R % 2 == 0, "packed kernel pairs i-bodies");
  constexpr int NP = R / 2;
  __shared__ __align__(16) float4 s_p[2][BLOCK];
  const int tid = threadIdx.x;
  const uint32_t tile_i = blockIdx.x * (uint32_t)(BLOCK * R);
  u64 nx[NP], ny[NP], nz[NP], ax[NP], ay[NP], az[NP];
  float4 own[R];
#pragma unroll
  for (int k = 0; k < R; k++) {
    uint32_t li = tile_i + k * BLOCK + tid;
    uint32_t lc = li < a.i_count ? li : a.i_count - 1;
    own[k] = a.pos[a.i_begin + lc];
  }
#pragma unroll
  for (int p = 0; p < NP; p++) {
    nx[p] = pack2(-own[2 * p].x, -own[2 * p + 1].x);
    ny[p] = pack2(-own[2 * p].y, -own[2 * p + 1].y);
    nz[p] = pack2(-own[2 * p].z, -own[2 * p + 1].z);
    if (a.flags & kFirstChunk) {
      ax[p] = ay[p] = az[p] = 0ull;
    } else {
      float4 c[2];
#pragma unroll
      for (int h = 0; h < 2; h++) {
        uint32_t li = tile_i + (2 * p + h) * BLOCK + tid;
        c[h] = __ldcg(&a.acc[li < a.i_count ? li : a.i_count - 1]);
      }
      ax[p] = pack2(c[0].x, c[1].x);
      ay[p] = pack2(c[0].y, c[1].y);
      az[p] = pack2(c[0].z, c[1].z);
    }
  }
  const u64 eps2 = pack2(a.eps, a.eps);
  sweep_cta_tiles<BLOCK>(a, s_p, tid, [&](const float4 q, uint32_t) { interact_packed<NP, false>(q, eps2, nx, ny, nz, ax, ay, az); });
#pragma unroll
  for (int k = 0; k < R; k++) {
    const uint32_t li = tile_i + k * BLOCK + tid;
    if (li >= a.i_count) continue;
    float fx0, fx1, fy0, fy1, fz0, fz1;
    unpack2(ax[k / 2], fx0, fx1);
    unpack2(ay[k / 2], fy0, fy1);
    unpack2(az[k / 2], fz0, fz1);
    finish_body(a, a.flags, li, (k & 1) ? fx1 : fx0, (k & 1) ? fy1 : fy0, (k & 1) ? fz1 : fz0, own[k]);
  }
}

template <int R, int BLOCK>
__global__ void __launch_bounds__(BLOCK) force_scalar_kernel(const StepArgs a) {
  __shared__ __align__(16) float4 s_p[2][BLOCK];
  const int tid = threadIdx.x;
  const uint32_t tile_i = blockIdx.x * (uint32_t)(BLOCK * R);
  float nx[R], ny[R], nz[R], ax[R], ay[R], az[R];
  float4 own[R];
  uint32_t gi[R];
#pragma unroll
  for (int k = 0; k < R; k++) {
    uint32_t li = tile_i + k * BLOCK + tid;
    uint32_t lc = li < a.i_count ? li : a.i_count - 1;
    gi[k] = a.i_begin + lc;
    own[k] = a.pos[gi[k]];
    nx[k] = -own[k].x;
    ny[k] = -own[k].y;
    nz[k] = -own[k].z;
    if (a.flags & kFirstChunk) {
      ax[k] = ay[k] = az[k] = 0.0f;
    } else {
      float4 c = __ldcg(&a.acc[lc]);
      ax[k] = c.x;
      ay[k] = c.y;
      az[k] = c.z;
    }
  }
  const float eps = a.eps;
  sweep_cta_tiles<BLOCK>(a, s_p, tid, [&](const float4 q, uint32_t gj) {
    interact_scalar<R, kSelfNone, false>(q, gj, eps, nx, ny, nz, gi, ax, ay, az);
  });
#pragma unroll
  for (int k = 0; k < R; k++) {
    const uint32_t li = tile_i + k * BLOCK + tid;
    if (li >= a.i_count) continue;
    finish_body(a, a.flags, li, ax[k], ay[k], az[k], own[k]);
  }
}

// TMA-staged variant of the production kernel.  Same arithmetic and ticketed j-segment hand-off; the
// warp's 32-body tiles are fetched by one lane with cp.async.bulk (SASS: UBLKCP) into a 4-stage
// shared-memory ring, completion tracked by one mbarrier per stage.  north_star: "TMA bulk copies
// where ncu shows they help" -- measured, they do not: the loop is FMA-pipe / register-file bound
// with long_scoreboard ~ 0 (1.5-2 % slower than LDG->STS staging, profiles/r01_tuning_log.txt section 7).
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int R, int MINB>
__global__ void __launch_bounds__(32, MINB)
    force_wseg_tma_kernel(const StepArgs a, const uint32_t groups, const uint32_t segs, const uint32_t seg_len,
                          unsigned int *words, unsigned int *error, const unsigned int epoch,
                          const unsigned int ticket_base, const unsigned long long timeout_ns) {
  static_assert(R % 2 == 0, "packed kernel pairs i-bodies");
  constexpr int NP = R / 2;
  constexpr int TJ = 32;
  constexpr int STAGES = 4;
  __shared__ __align__(128) float4 s_tile[STAGES][TJ];
  __shared__ __align__(8) unsigned long long s_bar[STAGES];
  const int lane = threadIdx.x & 31;
  uint32_t unit = blockIdx.x;
  if (segs > 1) {
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
  const uint32_t nj = j_end - j_begin;
  const uint32_t ntiles = (nj + TJ - 1) / TJ;

  auto issue = [&](uint32_t t) {  // lane 0 only: arm the stage's barrier and start the bulk copy of tile t
    const int st = t % STAGES;
    const uint32_t bytes = min((uint32_t)TJ, nj - t * TJ) * 16u;
    const uint32_t bar = smem_u32(&s_bar[st]);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(&s_tile[st][0])),
                 "l"(a.pos + j_begin + t * TJ), "r"(bytes), "r"(bar)
                 : "memory");
  };
  if (lane == 0) {
#pragma unroll
    for (int st = 0; st < STAGES; st++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_bar[st])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (uint32_t t = 0; t < STAGES && t < ntiles; t++) issue(t);
  }
  __syncwarp();

  u64 nx[NP], ny[NP], nz[NP], ax[NP], ay[NP], az[NP];
  float4 own[R];
#pragma unroll
  for (int k = 0; k < R; k++) {
    uint32_t li = warp_i + k * 32 + lane;
    own[k] = a.pos[a.i_begin + (li < a.i_count ? li : a.i_count - 1)];
  }
#pragma unroll
  for (int p = 0; p < NP; p++) {
    nx[p] = pack2(-own[2 * p].x, -own[2 * p + 1].x);
    ny[p] = pack2(-own[2 * p].y, -own[2 * p + 1].y);
    nz[p] = pack2(-own[2 * p].z, -own[2 * p + 1].z);
  }
  if (seg > 0) {
    if (lane == 0) wait_for_segment(words + 1 + g, epoch + seg, error, timeout_ns);
    __syncwarp();
  }
#pragma unroll
  for (int p = 0; p < NP; p++) {
    if (flags & kFirstChunk) {
      ax[p] = ay[p] = az[p] = 0ull;
    } else {
      float4 c[2];
#pragma unroll
      for (int h = 0; h < 2; h++) {
        uint32_t li = warp_i + (2 * p + h) * 32 + lane;
        c[h] = __ldcg(&a.acc[li < a.i_count ? li : a.i_count - 1]);
      }
      ax[p] = pack2(c[0].x, c[1].x);
      ay[p] = pack2(c[0].y, c[1].y);
      az[p] = pack2(c[0].z, c[1].z);
    }
  }
  const u64 eps2 = pack2(a.eps, a.eps);
  for (uint32_t t = 0; t < ntiles; t++) {
    const int st = t % STAGES;
    const uint32_t parity = (t / STAGES) & 1u;
    const uint32_t bar = smem_u32(&s_bar[st]);
    uint32_t done = 0;
    while (!done)
      asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                   : "=r"(done)
                   : "r"(bar), "r"(parity)
                   : "memory");
    const uint32_t cnt = min((uint32_t)TJ, nj - t * TJ);
    if (cnt == TJ) {
#pragma unroll
      for (int j = 0; j < TJ; j++) interact_packed<NP, false>(s_tile[st][j], eps2, nx, ny, nz, ax, ay, az);
    } else {
      for (uint32_t j = 0; j < cnt; j++) interact_packed<NP, false>(s_tile[st][j], eps2, nx, ny, nz, ax, ay, az);
    }
    __syncwarp();  // every lane is done reading the stage before it is refilled
    if (lane == 0 && t + STAGES < ntiles) issue(t + STAGES);
  }
#pragma unroll
  for (int k = 0; k < R; k++) {
    const uint32_t li = warp_i + k * 32 + lane;
    if (li >= a.i_count) continue;
    float fx0, fx1, fy0, fy1, fz0, fz1;
    unpack2(ax[k / 2], fx0, fx1);
    unpack2(ay[k / 2], fy0, fy1);
    unpack2(az[k / 2], fz0, fz1);
    finish_body(a, flags, li, (k & 1) ? fx1 : fx0, (k & 1) ? fy1 : fy0, (k & 1) ? fz1 : fz0, own[k]);
  }
  if (seg + 1 < segs) {
    __threadfence();
    __syncwarp();
    if (lane == 0) atomicExch(words + 1 + g, epoch + seg + 1);
  }
}

template <int R, int MINB>
static cudaError_t launch_wseg_tma(const StepArgs &a, int sms, cudaStream_t s) {
  const uint32_t groups = (a.i_count + 32 * R - 1) / (32 * R);
  const uint32_t nj = a.j_end - a.j_begin;
  SegSync *sy = a.sync;
  uint32_t segs = 1, seg_len = (nj + 31u) / 32u * 32u;
  if (sy && sy->words && groups <= sy->n_groups && nj > 0) {
    segs = plan_segments(groups, nj, sms, MINB);
    seg_len = ((nj + segs - 1) / segs + 31u) / 32u * 32u;
    segs = (nj + seg_len - 1) / seg_len;
  }
  unsigned int ep = 0, tb = 0;
  if (segs > 1) {
    if (sy->epoch > 0xf0000000u) {
      cudaError_t e = cudaMemsetAsync(sy->words + 1, 0, (size_t)sy->n_groups * sizeof(unsigned int), s);
      if (e != cudaSuccess) return e;
      sy->epoch = 0;
    }
    ep = sy->epoch;
    tb = sy->ticket_base;
    sy->epoch += segs;
    sy->ticket_base += groups * segs;
  }
  force_wseg_tma_kernel<R, MINB><<<groups * segs, 32, 0, s>>>(a, groups, segs, seg_len, sy ? sy->words : nullptr,
                                                              sy ? sy->error : nullptr, ep, tb, sy ? sy->timeout_ns : 0ull);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess && segs > 1) {
    sy->epoch -= segs;
    sy->ticket_base -= groups * segs;
  }
  return e;
}

template <int R, int BLOCK>
static cudaError_t launch_packed(const StepArgs &a, cudaStream_t s) {
  force_packed_kernel<R, BLOCK><<<(a.i_count + BLOCK * R - 1) / (BLOCK * R), BLOCK, 0, s>>>(a);
  return cudaGetLastError();
}
template <int R, int BLOCK>
static cudaError_t launch_scalar(const StepArgs &a, cudaStream_t s) {
  force_scalar_kernel<R, BLOCK><<<(a.i_count + BLOCK * R - 1) / (BLOCK * R), BLOCK, 0, s>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_variant(const KernelConfig &c, const StepArgs &a, cudaStream_t s) {
  if (c.mass || c.self_mode != kSelfNone) return cudaErrorInvalidConfiguration;  // masses / predicates: production kernels only
  if (c.family == kFamTma) {
    if (c.r == 6) return launch_wseg_tma<6, 14>(a, c.sms, s);
    if (c.r == 4) return launch_wseg_tma<4, 20>(a, c.sms, s);
    return cudaErrorInvalidConfiguration;
  }
#define NB_PACKED(RR, BB) \
  if (c.family == kFamPackedCta && c.r == RR && c.block == BB) return launch_packed<RR, BB>(a, s);
#define NB_SCALAR(RR, BB) \
  if (c.family == kFamScalarCta && c.r == RR && c.block == BB) return launch_scalar<RR, BB>(a, s);
  NB_PACKED(2, 64)
  NB_PACKED(2, 128)
  NB_PACKED(4, 128)
  NB_PACKED(4, 256)
  NB_SCALAR(2, 64)
  NB_SCALAR(2, 128)
  NB_SCALAR(4, 128)
  NB_SCALAR(4, 256)
#undef NB_PACKED
#undef NB_SCALAR
  return cudaErrorInvalidConfiguration;
}

}  // namespace nbody
#endif  // NBODY_VARIANTS
