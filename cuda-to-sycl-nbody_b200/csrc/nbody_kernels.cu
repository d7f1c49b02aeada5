// nbody_kernels.cu -- hand-written sm_100a kernels of the all-pairs force + integrate step.
//
// What is computed is fixed by the reference kernel particle_interaction<BRANCH>
// (src/simulator.cu:186-229) compiled with -use_fast_math; its sm_100a SASS does, per pair,
//     r  = p_j + (-p_i)                      3 FADD.FTZ
//     t  = ry*ry ; t = fma(rx,rx,t) ; t = fma(rz,rz,t)     FMUL + 2 FFMA
//     d  = t + distEps                       FADD   (softening is added to r^2, :201)
//     c  = d * (d*d)                         2 FMUL
//     w  = MUFU.RSQ(c)
//     a  = fma(r, w, a)                      3 FFMA, skipped when j == i
// with one accumulator per component and j ascending.  FP32 addition is not associative and
// the sums cancel heavily, so any other order differs from the reference by ~1e-5 relative at
// N = 262144.  Every kernel here therefore keeps that exact op sequence and
// order -- spelled in PTX with explicit .rn.ftz so ptxas cannot re-contract it -- and is
// bit-identical to the reference; the speed comes from how the sequence is fed and issued:
//
//   * the packed kernels pair two i-bodies in the two lanes of Blackwell's f32x2 instructions
//     (FADD2 / FMUL2 / FFMA2): 12 FP32 instructions serve TWO interactions; the j-body is a
//     scalar operand broadcast to both lanes by the instruction itself (SASS operand form
//     `R.F32`), so the tile stays in its HBM float4 layout and costs one LDS.128 per j;
//   * each thread register-blocks R i-bodies, so one LDS feeds R interactions;
//   * production kernel (force_wseg_kernel): one warp per CTA, warp-private 32-body j-tiles
//     (LDG.128 -> STS.128 -> __syncwarp -> broadcast LDS.128, double buffered, no CTA barrier),
//     and the j-sweep of a body group cut into consecutive CTAs of one grid that hand the
//     accumulators on through L2 in order -- still one FP32 chain per body, but short units,
//     which removes the low-occupancy tail of the launch;
//   * comparison kernels kept selectable: unsegmented warp-streaming, CTA-tiled packed (256-body
//     tiles, one __syncthreads per tile), TMA (cp.async.bulk + mbarrier) staged, scalar FFMA;
//   * no warp shuffles, no atomics, no j-split reduction: the accumulate is a per-thread FMA chain.
// Measurements behind every choice: profiles/r01_tuning_log.txt, DESIGN.md section 5.
#include "nbody_kernels.cuh"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>

namespace nbody {

typedef unsigned long long u64;

// ---- FP32 primitives with the reference's rounding/flush behaviour --------------------------
__device__ __forceinline__ float fadd(float a, float b) {
  float d;
  asm("add.rn.ftz.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b));
  return d;
}
__device__ __forceinline__ float fmul(float a, float b) {
  float d;
  asm("mul.rn.ftz.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b));
  return d;
}
__device__ __forceinline__ float ffma(float a, float b, float c) {
  float d;
  asm("fma.rn.ftz.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ float frsq(float a) {  // MUFU.RSQ, what rsqrt() is under -use_fast_math
  float d;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(d) : "f"(a));
  return d;
}
// packed pairs: two independent IEEE lanes per instruction (sm_100+)
__device__ __forceinline__ u64 pack2(float lo, float hi) {
  u64 d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi));
  return d;
}
__device__ __forceinline__ void unpack2(u64 v, float &lo, float &hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ u64 fadd2(u64 a, u64 b) {
  u64 d;
  asm("add.rn.ftz.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ u64 fmul2(u64 a, u64 b) {
  u64 d;
  asm("mul.rn.ftz.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) {
  u64 d;
  asm("fma.rn.ftz.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}

// velocity / position update, src/simulator.cu:213-228 in the reference SASS op order
__device__ __forceinline__ void integrate_component(float f, float &v, float &p, float dt, float G,
                                                    float damping) {
  float t = fmul(f, dt);
  float vd = fmul(v, damping);
  v = ffma(t, G, vd);
  p = ffma(v, dt, p);
}

// epilogue of one i-body: carry / dump the force sum, or integrate.  When integrating, the new
// position goes to this GPU's next-position replica AND, in peer-push mode, straight into every
// peer GPU's replica with plain stores over NVLink (the position "all-gather" is fused into the
// kernel: by the time the last warp retires, every GPU already holds this shard).
__device__ __forceinline__ void finish_body(const StepArgs &a, const int flags, uint32_t li, float fx,
                                            float fy, float fz, float4 p) {
  if (!(flags & kLastChunk) || (flags & kAccelOut)) {
    __stcg(&a.acc[li], make_float4(fx, fy, fz, 0.0f));  // L2: the next j-segment may run on another SM
    return;
  }
  float4 v = a.vel[li];
  integrate_component(fx, v.x, p.x, a.dt, a.G, a.damping);
  integrate_component(fy, v.y, p.y, a.dt, a.G, a.damping);
  integrate_component(fz, v.z, p.z, a.dt, a.G, a.damping);
  a.vel[li] = v;
  const uint32_t gi = a.i_begin + li;
  a.pos_next[gi] = p;
  for (int k = 0; k < a.n_peers; k++) a.peer_next[k][gi] = p;
}

// =============================================================================================
// packed kernel: R (even) i-bodies per thread, pairs (2p, 2p+1) share one f32x2 lane pair
// =============================================================================================
template <int R, int BLOCK>
__global__ void __launch_bounds__(BLOCK) force_packed_kernel(const StepArgs a) {
  static_assert(R % 2 == 0, "packed kernel pairs i-bodies");
  constexpr int NP = R / 2;
  constexpr int TJ = BLOCK;  // one j-body per thread per tile fill
  __shared__ __align__(16) float4 s_p[2][TJ];  // tile of j-bodies as they lie in HBM

  const int tid = threadIdx.x;
  const uint32_t tile_i = blockIdx.x * (uint32_t)(BLOCK * R);

  u64 nx[NP], ny[NP], nz[NP];  // negated i positions, packed
  u64 ax[NP], ay[NP], az[NP];  // accumulators, packed
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
  }
  if (a.flags & kFirstChunk) {
#pragma unroll
    for (int p = 0; p < NP; p++) ax[p] = ay[p] = az[p] = 0ull;
  } else {
#pragma unroll
    for (int p = 0; p < NP; p++) {
      float4 c[2];
#pragma unroll
      for (int h = 0; h < 2; h++) {
        uint32_t li = tile_i + (2 * p + h) * BLOCK + tid;
        uint32_t lc = li < a.i_count ? li : a.i_count - 1;
        c[h] = __ldcg(&a.acc[lc]);
      }
      ax[p] = pack2(c[0].x, c[1].x);
      ay[p] = pack2(c[0].y, c[1].y);
      az[p] = pack2(c[0].z, c[1].z);
    }
  }
  const u64 eps2 = pack2(a.eps, a.eps);

  const uint32_t nj = a.j_end - a.j_begin;
  const uint32_t ntiles = (nj + TJ - 1) / TJ;

  auto fill = [&](int buf, float4 v) { s_p[buf][tid] = v; };
  auto fetch = [&](uint32_t t) -> float4 {
    uint32_t j = a.j_begin + t * TJ + tid;
    return a.pos[j < a.j_end ? j : a.j_end - 1];
  };
  auto interact = [&](int buf, int j) {
    // one broadcast LDS.128 per j; pack2(q.x, q.x) costs nothing: ptxas encodes it as the
    // scalar-broadcast operand form of FADD2 (`R.F32`), see profiles/sass_*.txt
    const float4 q = s_p[buf][j];
    const u64 qx = pack2(q.x, q.x), qy = pack2(q.y, q.y), qz = pack2(q.z, q.z);
#pragma unroll
    for (int p = 0; p < NP; p++) {
      u64 rx = fadd2(qx, nx[p]);
      u64 ry = fadd2(qy, ny[p]);
      u64 rz = fadd2(qz, nz[p]);
      u64 t = fmul2(ry, ry);
      t = ffma2(rx, rx, t);
      t = ffma2(rz, rz, t);
      u64 d = fadd2(t, eps2);
      u64 c = fmul2(d, d);
      c = fmul2(d, c);
      float c0, c1;
      unpack2(c, c0, c1);
      u64 w = pack2(frsq(c0), frsq(c1));
      ax[p] = ffma2(rx, w, ax[p]);
      ay[p] = ffma2(ry, w, ay[p]);
      az[p] = ffma2(rz, w, az[p]);
    }
  };

  if (ntiles > 0) fill(0, fetch(0));
  __syncthreads();
  for (uint32_t t = 0; t < ntiles; t++) {
    const int buf = t & 1;
    float4 nxt;
    const bool more = t + 1 < ntiles;
    if (more) nxt = fetch(t + 1);
    const uint32_t cnt = min((uint32_t)TJ, nj - t * TJ);
    if (cnt == TJ) {
#pragma unroll 32
      for (int j = 0; j < TJ; j++) interact(buf, j);
    } else {
      for (uint32_t j = 0; j < cnt; j++) interact(buf, (int)j);
    }
    if (more) fill(buf ^ 1, nxt);
    __syncthreads();
  }

  // epilogue: carry, dump, or integrate
#pragma unroll
  for (int k = 0; k < R; k++) {
    const uint32_t li = tile_i + k * BLOCK + tid;
    if (li >= a.i_count) continue;
    float fx0, fx1, fy0, fy1, fz0, fz1;
    unpack2(ax[k / 2], fx0, fx1);
    unpack2(ay[k / 2], fy0, fy1);
    unpack2(az[k / 2], fz0, fz1);
    const float fx = (k & 1) ? fx1 : fx0, fy = (k & 1) ? fy1 : fy0, fz = (k & 1) ? fz1 : fz0;
    finish_body(a, a.flags, li, fx, fy, fz, own[k]);
  }
}

// =============================================================================================
// warp-streaming packed kernel (the production kernel)
//
// Same arithmetic as force_packed_kernel, different feeding: every WARP stages its own 32-body
// j-tiles (one coalesced LDG.128 per lane -> STS.128 -> __syncwarp -> 32 broadcast LDS.128),
// double buffered, so there is no CTA-wide barrier and warps never wait for each other.  A CTA
// is WARPS independent warps; with WARPS = 1 the hardware block scheduler balances the grid at
// warp granularity.  The host caps the number of resident CTAs per SM through the dynamic
// shared-memory size so that the grid runs as an integer number of equally full "generations"
// (see plan_wstream): all SM sub-partitions then keep >= 5-7 warps until the very end, which
// is what the FMA pipe needs to stay saturated (tools/ubench_fma2.cu, profiles/).
// =============================================================================================
// one warp's work: R*32 i-bodies starting at shard-local index warp_i, all j of the launch
template <int R, bool MASS>
__device__ __forceinline__ void wstream_body(const StepArgs &a, const uint32_t j_begin, const uint32_t j_end,
                                             const int flags, const uint32_t warp_i, float4 (*tile)[32],
                                             const int lane) {
  static_assert(R % 2 == 0, "packed kernel pairs i-bodies");
  constexpr int NP = R / 2;
  constexpr int TJ = 32;

  u64 nx[NP], ny[NP], nz[NP];
  u64 ax[NP], ay[NP], az[NP];
  float4 own[R];
#pragma unroll
  for (int k = 0; k < R; k++) {
    uint32_t li = warp_i + k * 32 + lane;
    uint32_t lc = li < a.i_count ? li : a.i_count - 1;
    own[k] = a.pos[a.i_begin + lc];
  }
#pragma unroll
  for (int p = 0; p < NP; p++) {
    nx[p] = pack2(-own[2 * p].x, -own[2 * p + 1].x);
    ny[p] = pack2(-own[2 * p].y, -own[2 * p + 1].y);
    nz[p] = pack2(-own[2 * p].z, -own[2 * p + 1].z);
  }
  if (flags & kFirstChunk) {
#pragma unroll
    for (int p = 0; p < NP; p++) ax[p] = ay[p] = az[p] = 0ull;
  } else {
#pragma unroll
    for (int p = 0; p < NP; p++) {
      float4 c[2];
#pragma unroll
      for (int h = 0; h < 2; h++) {
        uint32_t li = warp_i + (2 * p + h) * 32 + lane;
        uint32_t lc = li < a.i_count ? li : a.i_count - 1;
        c[h] = __ldcg(&a.acc[lc]);
      }
      ax[p] = pack2(c[0].x, c[1].x);
      ay[p] = pack2(c[0].y, c[1].y);
      az[p] = pack2(c[0].z, c[1].z);
    }
  }
  const u64 eps2 = pack2(a.eps, a.eps);
  const uint32_t nj = j_end - j_begin;
  const uint32_t ntiles = (nj + TJ - 1) / TJ;

  auto fetch = [&](uint32_t t) -> float4 {
    uint32_t j = j_begin + t * TJ + lane;
    return a.pos[j < j_end ? j : j_end - 1];
  };
  auto interact = [&](int buf, int j) {
    const float4 q = tile[buf][j];
    const u64 qx = pack2(q.x, q.x), qy = pack2(q.y, q.y), qz = pack2(q.z, q.z);
#pragma unroll
    for (int p = 0; p < NP; p++) {
      u64 rx = fadd2(qx, nx[p]);
      u64 ry = fadd2(qy, ny[p]);
      u64 rz = fadd2(qz, nz[p]);
      u64 t = fmul2(ry, ry);
      t = ffma2(rx, rx, t);
      t = ffma2(rz, rz, t);
      u64 d = fadd2(t, eps2);
      u64 c = fmul2(d, d);
      c = fmul2(d, c);
      float c0, c1;
      unpack2(c, c0, c1);
      u64 w = pack2(frsq(c0), frsq(c1));
      if (MASS) w = fmul2(w, pack2(q.w, q.w));  // extension: per-body mass m_j (float4.w); m = 1 changes no bit
      ax[p] = ffma2(rx, w, ax[p]);
      ay[p] = ffma2(ry, w, ay[p]);
      az[p] = ffma2(rz, w, az[p]);
    }
  };

  if (ntiles > 0) tile[0][lane] = fetch(0);
  __syncwarp();
  for (uint32_t t = 0; t < ntiles; t++) {
    const int buf = t & 1;
    float4 nxt;
    const bool more = t + 1 < ntiles;
    if (more) nxt = fetch(t + 1);
    const uint32_t cnt = min((uint32_t)TJ, nj - t * TJ);
    if (cnt == TJ) {
#pragma unroll
      for (int j = 0; j < TJ; j++) interact(buf, j);
    } else {
      for (uint32_t j = 0; j < cnt; j++) interact(buf, (int)j);
    }
    if (more) tile[buf ^ 1][lane] = nxt;
    __syncwarp();
  }

#pragma unroll
  for (int k = 0; k < R; k++) {
    const uint32_t li = warp_i + k * 32 + lane;
    if (li >= a.i_count) continue;
    float fx0, fx1, fy0, fy1, fz0, fz1;
    unpack2(ax[k / 2], fx0, fx1);
    unpack2(ay[k / 2], fy0, fy1);
    unpack2(az[k / 2], fz0, fz1);
    const float fx = (k & 1) ? fx1 : fx0, fy = (k & 1) ? fy1 : fy0, fz = (k & 1) ? fz1 : fz0;
    finish_body(a, flags, li, fx, fy, fz, own[k]);
  }
}

template <int R, int WARPS, bool MASS>
__global__ void __launch_bounds__(32 * WARPS, 28 / WARPS) force_wstream_kernel(const StepArgs a) {
  __shared__ __align__(16) float4 s_tile[WARPS][2][32];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const uint32_t warp_i = (blockIdx.x * (uint32_t)WARPS + warp) * (uint32_t)(32 * R);
  if (warp_i >= a.i_count) return;  // no CTA-wide synchronisation anywhere
  wstream_body<R, MASS>(a, a.j_begin, a.j_end, a.flags, warp_i, s_tile[warp], lane);
}

// Spin (with back-off) until the predecessor segment of this body group has published its
// accumulators.  The predecessor CTA has a lower blockIdx and is therefore already resident or done;
// legitimate waits are at most one unit long (<= a few seconds at the largest sizes).  Should the
// dispatch-order assumption ever be violated, trap after ~20 s instead of hanging the device.
__device__ __forceinline__ void wait_for_segment(unsigned int *word, unsigned int target) {
  volatile unsigned int *p = word;
  unsigned int spins = 0;
  while (*p != target) {
    __nanosleep(128);
    if (++spins > (1u << 27)) __trap();
  }
  __threadfence();
}

// (28 resident warps per SM = a 72-register budget: the schedule ptxas finds there is the fastest
// measured -- 60 or 79 registers lose 6-11 %, profiles/r01_tuning_log.txt)
// j-segmented launch.  A body-group's sweep over j is cut into `segs` consecutive segments that
// are separate CTAs of the SAME grid: CTA (seg, g) handles group g over segment seg, takes the
// accumulators from `acc` and hands them on through `acc`, in order -- so the per-body sum is still
// one FP32 chain over ascending j (bit-exact), but the schedulable unit is `segs` times shorter.
// Why: with equal units the launch ends with every SM sub-partition holding ~W/2 units of
// leftovers that finish one by one at falling occupancy; that tail costs ~0.46 unit-times
// (3.3 % at N = 1M, 6.6 % at 512K bodies, measured) and shrinks in proportion to the unit.
// Hand-off: CTA (seg, g) spins on progress[g] until CTA (seg-1, g) has published seg.  The
// predecessor has a lower blockIdx, so it was dispatched earlier and never waits on a later
// CTA: forward progress is guaranteed (same argument as decoupled look-back scans).
template <int R, int MINB, bool MASS>
__global__ void __launch_bounds__(32, MINB) force_wseg_kernel(const StepArgs a, const uint32_t groups,
                                                        const uint32_t segs, const uint32_t seg_len,
                                                        unsigned int *progress, const unsigned int epoch) {
  __shared__ __align__(16) float4 s_tile[2][32];
  const int lane = threadIdx.x & 31;
  const uint32_t seg = blockIdx.x / groups;
  const uint32_t g = blockIdx.x - seg * groups;
  const uint32_t warp_i = g * (uint32_t)(32 * R);
  if (warp_i >= a.i_count) return;
  const uint32_t j_begin = a.j_begin + seg * seg_len;
  const uint32_t j_end = min(a.j_end, j_begin + seg_len);
  const int flags = (seg == 0 ? (a.flags & kFirstChunk) : 0) | (seg == segs - 1 ? (a.flags & (kLastChunk | kAccelOut)) : 0);
  if (seg > 0) {
    if (lane == 0) wait_for_segment(progress + g, epoch + seg);
    __syncwarp();
  }
  wstream_body<R, MASS>(a, j_begin, j_end, flags, warp_i, s_tile, lane);
  if (seg + 1 < segs) {
    __threadfence();  // every lane publishes its accumulator stores ...
    __syncwarp();
    if (lane == 0) atomicExch(progress + g, epoch + seg + 1);  // ... before the group is handed on
  }
}

// =============================================================================================
// small-N kernel: one warp per CTA, warp-private tiles like the production kernel, but SCALAR
// FP32 ops and R i-bodies per lane (R = 1 or 2).  With N of a few ten thousand bodies there are
// fewer warps than SM sub-partitions can hold, so lanes, not issue slots, are scarce: one body per
// lane doubles the number of warps relative to the packed R = 2 kernel, and a scalar op issues in
// one cycle where a packed one holds the pipe for two.  Same op sequence, bit-exact.
// =============================================================================================
template <int R>
__global__ void __launch_bounds__(32) force_wsmall_kernel(const StepArgs a) {
  constexpr int TJ = 32;
  __shared__ __align__(16) float4 tile[2][TJ];
  const int lane = threadIdx.x & 31;
  const uint32_t warp_i = blockIdx.x * (uint32_t)(32 * R);
  if (warp_i >= a.i_count) return;
  float nx[R], ny[R], nz[R], ax[R], ay[R], az[R];
  float4 own[R];
#pragma unroll
  for (int k = 0; k < R; k++) {
    uint32_t li = warp_i + k * 32 + lane;
    uint32_t lc = li < a.i_count ? li : a.i_count - 1;
    own[k] = a.pos[a.i_begin + lc];
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
  const uint32_t nj = a.j_end - a.j_begin;
  const uint32_t ntiles = (nj + TJ - 1) / TJ;
  auto fetch = [&](uint32_t t) -> float4 {
    uint32_t j = a.j_begin + t * TJ + lane;
    return a.pos[j < a.j_end ? j : a.j_end - 1];
  };
  auto interact = [&](int buf, int j) {
    const float4 q = tile[buf][j];
#pragma unroll
    for (int k = 0; k < R; k++) {
      float rx = fadd(q.x, nx[k]);
      float ry = fadd(q.y, ny[k]);
      float rz = fadd(q.z, nz[k]);
      float t = fmul(ry, ry);
      t = ffma(rx, rx, t);
      t = ffma(rz, rz, t);
      float d = fadd(t, eps);
      float c = fmul(d, d);
      c = fmul(d, c);
      float w = frsq(c);
      ax[k] = ffma(rx, w, ax[k]);
      ay[k] = ffma(ry, w, ay[k]);
      az[k] = ffma(rz, w, az[k]);
    }
  };
  if (ntiles > 0) tile[0][lane] = fetch(0);
  __syncwarp();
  for (uint32_t t = 0; t < ntiles; t++) {
    const int buf = t & 1;
    float4 nxt;
    const bool more = t + 1 < ntiles;
    if (more) nxt = fetch(t + 1);
    const uint32_t cnt = min((uint32_t)TJ, nj - t * TJ);
    if (cnt == TJ) {
#pragma unroll
      for (int j = 0; j < TJ; j++) interact(buf, j);
    } else {
      for (uint32_t j = 0; j < cnt; j++) interact(buf, (int)j);
    }
    if (more) tile[buf ^ 1][lane] = nxt;
    __syncwarp();
  }
#pragma unroll
  for (int k = 0; k < R; k++) {
    const uint32_t li = warp_i + k * 32 + lane;
    if (li >= a.i_count) continue;
    finish_body(a, a.flags, li, ax[k], ay[k], az[k], own[k]);
  }
}

// =============================================================================================
// TMA-staged variant of the production kernel (comparison only, NBODY_KERNEL_CONFIG="6,32,5").
// Same arithmetic and j-segmented hand-off; the warp's 32-body tiles are fetched by one lane with
// cp.async.bulk (SASS: UBLKCP) into a 4-stage shared-memory ring, completion tracked by one
// mbarrier per stage.  north_star: "TMA bulk copies where ncu shows they help" -- measured, they
// do not: the loop is FMA-pipe bound with long_scoreboard ~ 0, see profiles/r01_tuning_log.txt.
// =============================================================================================
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int R, int MINB>
__global__ void __launch_bounds__(32, MINB) force_wseg_tma_kernel(const StepArgs a, const uint32_t groups,
                                                            const uint32_t segs, const uint32_t seg_len,
                                                            unsigned int *progress, const unsigned int epoch) {
  static_assert(R % 2 == 0, "packed kernel pairs i-bodies");
  constexpr int NP = R / 2;
  constexpr int TJ = 32;
  constexpr int STAGES = 4;
  __shared__ __align__(128) float4 s_tile[STAGES][TJ];
  __shared__ __align__(8) unsigned long long s_bar[STAGES];
  const int lane = threadIdx.x & 31;
  const uint32_t seg = blockIdx.x / groups;
  const uint32_t g = blockIdx.x - seg * groups;
  const uint32_t warp_i = g * (uint32_t)(32 * R);
  if (warp_i >= a.i_count) return;
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
    uint32_t lc = li < a.i_count ? li : a.i_count - 1;
    own[k] = a.pos[a.i_begin + lc];
  }
#pragma unroll
  for (int p = 0; p < NP; p++) {
    nx[p] = pack2(-own[2 * p].x, -own[2 * p + 1].x);
    ny[p] = pack2(-own[2 * p].y, -own[2 * p + 1].y);
    nz[p] = pack2(-own[2 * p].z, -own[2 * p + 1].z);
  }
  if (seg > 0) {
    if (lane == 0) wait_for_segment(progress + g, epoch + seg);
    __syncwarp();
  }
  if (flags & kFirstChunk) {
#pragma unroll
    for (int p = 0; p < NP; p++) ax[p] = ay[p] = az[p] = 0ull;
  } else {
#pragma unroll
    for (int p = 0; p < NP; p++) {
      float4 c[2];
#pragma unroll
      for (int h = 0; h < 2; h++) {
        uint32_t li = warp_i + (2 * p + h) * 32 + lane;
        uint32_t lc = li < a.i_count ? li : a.i_count - 1;
        c[h] = __ldcg(&a.acc[lc]);
      }
      ax[p] = pack2(c[0].x, c[1].x);
      ay[p] = pack2(c[0].y, c[1].y);
      az[p] = pack2(c[0].z, c[1].z);
    }
  }
  const u64 eps2 = pack2(a.eps, a.eps);
  auto interact = [&](int st, int j) {
    const float4 q = s_tile[st][j];
    const u64 qx = pack2(q.x, q.x), qy = pack2(q.y, q.y), qz = pack2(q.z, q.z);
#pragma unroll
    for (int p = 0; p < NP; p++) {
      u64 rx = fadd2(qx, nx[p]);
      u64 ry = fadd2(qy, ny[p]);
      u64 rz = fadd2(qz, nz[p]);
      u64 t = fmul2(ry, ry);
      t = ffma2(rx, rx, t);
      t = ffma2(rz, rz, t);
      u64 d = fadd2(t, eps2);
      u64 c = fmul2(d, d);
      c = fmul2(d, c);
      float c0, c1;
      unpack2(c, c0, c1);
      u64 w = pack2(frsq(c0), frsq(c1));
      ax[p] = ffma2(rx, w, ax[p]);
      ay[p] = ffma2(ry, w, ay[p]);
      az[p] = ffma2(rz, w, az[p]);
    }
  };
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
      for (int j = 0; j < TJ; j++) interact(st, j);
    } else {
      for (uint32_t j = 0; j < cnt; j++) interact(st, (int)j);
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
    const float fx = (k & 1) ? fx1 : fx0, fy = (k & 1) ? fy1 : fy0, fz = (k & 1) ? fz1 : fz0;
    finish_body(a, flags, li, fx, fy, fz, own[k]);
  }
  if (seg + 1 < segs) {
    __threadfence();
    __syncwarp();
    if (lane == 0) atomicExch(progress + g, epoch + seg + 1);
  }
}

// =============================================================================================
// scalar kernel: R i-bodies per thread, scalar FADD/FMUL/FFMA; also the generic/faithful path
// (BRANCH predicate for any eps, PREDICATED as shipped)
// =============================================================================================
template <int R, int BLOCK, int SELF, bool MASS>
__global__ void __launch_bounds__(BLOCK) force_scalar_kernel(const StepArgs a) {
  constexpr int TJ = BLOCK;
  __shared__ __align__(16) float4 s_p[2][TJ];

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
  const uint32_t nj = a.j_end - a.j_begin;
  const uint32_t ntiles = (nj + TJ - 1) / TJ;

  auto fetch = [&](uint32_t t) -> float4 {
    uint32_t j = a.j_begin + t * TJ + tid;
    return a.pos[j < a.j_end ? j : a.j_end - 1];
  };
  auto interact = [&](int buf, int j, uint32_t gj) {
    const float4 q = s_p[buf][j];
#pragma unroll
    for (int k = 0; k < R; k++) {
      float rx = fadd(q.x, nx[k]);
      float ry = fadd(q.y, ny[k]);
      float rz = fadd(q.z, nz[k]);
      float t = fmul(ry, ry);
      t = ffma(rx, rx, t);
      t = ffma(rz, rz, t);
      float d = fadd(t, eps);
      float c = fmul(d, d);
      c = fmul(d, c);
      float w = frsq(c);
      if (MASS) w = fmul(w, q.w);  // extension: per-body mass m_j
      if (SELF == kSelfNone) {
        ax[k] = ffma(rx, w, ax[k]);
        ay[k] = ffma(ry, w, ay[k]);
        az[k] = ffma(rz, w, az[k]);
      } else if (SELF == kSelfBranch) {
        if (gj != gi[k]) {
          ax[k] = ffma(rx, w, ax[k]);
          ay[k] = ffma(ry, w, ay[k]);
          az[k] = ffma(rz, w, az[k]);
        }
      } else {  // as shipped: force += r * inv * (i == id)
        const float sel = (gj == gi[k]) ? 1.0f : 0.0f;
        ax[k] = ffma(fmul(rx, w), sel, ax[k]);
        ay[k] = ffma(fmul(ry, w), sel, ay[k]);
        az[k] = ffma(fmul(rz, w), sel, az[k]);
      }
    }
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
#pragma unroll 8
      for (int j = 0; j < TJ; j++) interact(buf, j, gj0 + j);
    } else {
      for (uint32_t j = 0; j < cnt; j++) interact(buf, (int)j, gj0 + j);
    }
    if (more) s_p[buf ^ 1][tid] = nxt;
    __syncthreads();
  }

#pragma unroll
  for (int k = 0; k < R; k++) {
    const uint32_t li = tile_i + k * BLOCK + tid;
    if (li >= a.i_count) continue;
    finish_body(a, a.flags, li, ax[k], ay[k], az[k], own[k]);
  }
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

KernelConfig choose_config(int requested_kernel, int calc_method, float eps, uint32_t i_count,
                           int sms, bool has_mass) {
  KernelConfig c;
  const bool exact_unpred = eps_allows_unpredicated(eps);
  if (calc_method != 0) {  // PREDICATED, as shipped
    c = {0, 1, 128, kSelfPredicated, sms, has_mass ? 1 : 0};
    return c;
  }
  if (requested_kernel == 1 /*GENERIC*/ || !exact_unpred) {
    c = {0, 1, 128, kSelfBranch, sms, has_mass ? 1 : 0};
    return c;
  }
  // family: 4 = j-segmented warp-streaming packed (AUTO), 6 = small-N scalar warp-streaming (AUTO, tiny shards),
  // 3 = unsegmented, 5 = TMA-staged, 1 = CTA-tiled packed, 2 = CTA-tiled scalar
  int family = requested_kernel == 3 ? 2 : (requested_kernel == 2 ? 1 : 4);
  int r = 4, block = 128;
  if (family == 4) {
    // one warp per CTA.  Wider register blocking (fewer LDS per interaction, more ILP, fewer
    // resident warps) as long as the shard still supplies ~2 warps per resident slot:
    // R = 6 at 14 warps/SM, R = 4 at 20, R = 2 at 28 (measured: profiles/r01_tuning_log.txt)
    block = 32;
    r = 6;
    if ((uint64_t)i_count < (uint64_t)sms * 2700u) r = 4;  // < ~400K bodies on 148 SMs
    if ((uint64_t)i_count < (uint64_t)sms * 1350u) r = 2;  // < ~200K bodies
    if ((uint64_t)i_count <= (uint64_t)sms * 128u && !has_mass) {
      // at most one warp per SM sub-partition even at one body per lane (the reference's interactive
      // sizes, N <= ~19K): scalar ops, R = 1 -- 1.5x the packed kernel there (tools/small_n.py)
      family = 6;
      r = 1;
    }
  } else if (family == 3) {
    block = 32;
    if ((uint64_t)i_count < (uint64_t)sms * 20u * 128u) r = 2;
  } else {
    // i-bodies per CTA = block*r.  Keep at least ~4 CTA-tiles per SM so the tail of the grid is
    // short; with fewer bodies fall back to narrower register blocking.
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
      if (got == 3 && ef >= 1 && ef <= 6) family = ef;
    }
  }
  c = {family, r, block, kSelfNone, sms, has_mass ? 1 : 0};
  return c;
}

const char *config_name(const KernelConfig &c, char *buf, size_t len) {
  const char *fam = c.family == 1 ? "packed_f32x2" : (c.family == 2 ? "scalar_blocked" : (c.family == 3 ? "wstream_f32x2" : (c.family == 4 ? "wseg_f32x2" : (c.family == 5 ? "wseg_tma_f32x2" : (c.family == 6 ? "wsmall_scalar" : "generic")))));
  const char *self = c.self_mode == kSelfNone ? "nopred"
                                              : (c.self_mode == kSelfBranch ? "branch" : "predicated");
  snprintf(buf, len, "%s_r%d_b%d_%s%s", fam, c.r, c.block, self, c.mass ? "_mass" : "");
  return buf;
}

template <int R, int BLOCK>
static cudaError_t launch_packed(const StepArgs &a, cudaStream_t s) {
  uint32_t grid = (a.i_count + BLOCK * R - 1) / (BLOCK * R);
  force_packed_kernel<R, BLOCK><<<grid, BLOCK, 0, s>>>(a);
  return cudaGetLastError();
}
template <int R, int BLOCK, int SELF, bool MASS = false>
static cudaError_t launch_scalar(const StepArgs &a, cudaStream_t s) {
  uint32_t grid = (a.i_count + BLOCK * R - 1) / (BLOCK * R);
  force_scalar_kernel<R, BLOCK, SELF, MASS><<<grid, BLOCK, 0, s>>>(a);
  return cudaGetLastError();
}

// ---- warp-streaming kernel: residency planning ------------------------------------------------
// The grid is `ctas` identical CTAs.  If an SM can hold k_max of them, the grid runs in
// gens = ceil(ctas_per_sm / k_max) generations; capping residency at k = ceil(ctas_per_sm / gens)
// makes every generation equally full instead of leaving a thin last one (e.g. 55.4 CTAs/SM
// -> 28 + 27.4 rather than 32 + 23.4).  The cap is enforced with dynamic shared memory.
struct WstreamPlan {
  int k_cap = 0;        // resident CTAs per SM to aim for
  size_t dyn_smem = 0;  // dynamic shared memory per CTA that enforces it
};

// per (kernel, device): attributes set once, last plan cached (a process may drive several GPUs)
struct PlanSlot {
  const void *kern = nullptr;
  int dev = -1;
  uint32_t ctas = 0;
  int sms = 0;
  WstreamPlan plan;
};
static PlanSlot g_plan_slots[256];
static int g_plan_slot_count = 0;

static cudaError_t plan_resident(const void *kern, int threads, uint32_t ctas, int sms, WstreamPlan *out) {
  cudaError_t e;
  int dev = 0;
  if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
  PlanSlot *slot = nullptr;
  for (int i = 0; i < g_plan_slot_count; i++)
    if (g_plan_slots[i].kern == kern && g_plan_slots[i].dev == dev) slot = &g_plan_slots[i];
  if (!slot) {
    if (g_plan_slot_count >= 256) return cudaErrorMemoryAllocation;
    if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)) != cudaSuccess) return e;
    slot = &g_plan_slots[g_plan_slot_count++];
    slot->kern = kern;
    slot->dev = dev;
  }
  if (slot->ctas == ctas && slot->sms == sms && slot->plan.k_cap > 0) {
    *out = slot->plan;
    return cudaSuccess;
  }
  int k_max = 0;
  if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&k_max, kern, threads, 0)) != cudaSuccess) return e;
  if (k_max < 1) return cudaErrorInvalidConfiguration;
  int k = k_max;  // ctas == 0: no cap (grids whose CTAs differ in size drain as one queue)
  if (ctas > 0) {
    const double per_sm = (double)ctas / sms;
    int gens = (int)ceil(per_sm / k_max);
    if (gens < 1) gens = 1;
    k = (int)ceil(per_sm / gens);
    if (k > k_max) k = k_max;
    if (k < 1) k = 1;
  }
  if (const char *env = getenv("NBODY_RESIDENT_CTAS")) {  // tuning override
    int v = atoi(env);
    if (v >= 1 && v <= k_max) k = v;
  }
  size_t dyn = 0;
  if (k < k_max) {
    // largest dynamic size that still lets k CTAs fit, found with the occupancy API itself
    size_t lo = 0, hi = 200 * 1024;
    while (hi - lo > 256) {
      size_t mid = (lo + hi) / 2;
      int occ = 0;
      if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, mid)) != cudaSuccess) return e;
      if (occ >= k) lo = mid; else hi = mid;
    }
    dyn = lo;
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, dyn);
    if (occ != k) dyn = 0;  // cannot hit k exactly: leave the hardware limit
  }
  slot->ctas = ctas;
  slot->sms = sms;
  slot->plan.k_cap = k;
  slot->plan.dyn_smem = dyn;
  *out = slot->plan;
  return cudaSuccess;
}

template <int R, int WARPS, bool MASS>
static cudaError_t launch_wstream(const StepArgs &a, int sms, cudaStream_t s) {
  const uint32_t warps = (a.i_count + 32 * R - 1) / (32 * R);
  const uint32_t ctas = (warps + WARPS - 1) / WARPS;
  WstreamPlan plan;
  cudaError_t e = plan_resident((const void *)force_wstream_kernel<R, WARPS, MASS>, 32 * WARPS, ctas, sms, &plan);
  if (e != cudaSuccess) return e;
  force_wstream_kernel<R, WARPS, MASS><<<ctas, 32 * WARPS, plan.dyn_smem, s>>>(a);
  return cudaGetLastError();
}

// segment planning: aim at ~16 schedulable units per resident warp slot
static uint32_t plan_segments(uint32_t groups, uint32_t nj, int sms, int k_max) {
  if (const char *env = getenv("NBODY_SEGS")) {  // tuning override
    int v = atoi(env);
    if (v >= 1 && v <= 1024) return (uint32_t)v;
  }
  const double gens = (double)groups / ((double)sms * k_max);
  if (gens < 0.25) return 1;  // far fewer warps than slots: later segments would only spin (measured, N < 64K)
  int segs = (int)ceil(16.0 / (gens > 0.05 ? gens : 0.05));
  if (segs > 64) segs = 64;
  while (segs > 1 && nj / segs < 4096) segs--;  // keep segments long: hand-off cost stays invisible
  return (uint32_t)(segs < 1 ? 1 : segs);
}

template <int R, int MINB, bool MASS>
static cudaError_t launch_wseg_mb(const StepArgs &a, int sms, unsigned int *progress, unsigned int *epoch,
                                  cudaStream_t s) {
  const uint32_t groups = (a.i_count + 32 * R - 1) / (32 * R);
  static int k_max_cache[64] = {0};  // per device: resident CTAs per SM of this instantiation
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
  if (k_max_cache[dev] == 0) {
    int k = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&k, (const void *)force_wseg_kernel<R, MINB, MASS>, 32, 0);
    if (e != cudaSuccess) return e;
    k_max_cache[dev] = k > 0 ? k : 1;
  }
  const int k_max = k_max_cache[dev];
  const uint32_t nj = a.j_end - a.j_begin;
  uint32_t segs = plan_segments(groups, nj, sms, k_max);
  uint32_t seg_len = ((nj + segs - 1) / segs + 31u) / 32u * 32u;
  segs = (nj + seg_len - 1) / seg_len;
  if (*epoch > 0xf0000000u) return cudaErrorInvalidValue;  // 4e9 segment-launches: recreate the handle
  const unsigned int ep = *epoch;
  *epoch += segs;
  force_wseg_kernel<R, MINB, MASS><<<groups * segs, 32, 0, s>>>(a, groups, segs, seg_len, progress, ep);
  return cudaGetLastError();
}

template <int R, int MINB>
static cudaError_t launch_wseg_tma(const StepArgs &a, int sms, unsigned int *progress, unsigned int *epoch,
                                   cudaStream_t s) {
  const uint32_t groups = (a.i_count + 32 * R - 1) / (32 * R);
  const uint32_t nj = a.j_end - a.j_begin;
  uint32_t segs = plan_segments(groups, nj, sms, MINB);
  uint32_t seg_len = ((nj + segs - 1) / segs + 31u) / 32u * 32u;
  segs = (nj + seg_len - 1) / seg_len;
  if (*epoch > 0xf0000000u) return cudaErrorInvalidValue;
  const unsigned int ep = *epoch;
  *epoch += segs;
  force_wseg_tma_kernel<R, MINB><<<groups * segs, 32, 0, s>>>(a, groups, segs, seg_len, progress, ep);
  return cudaGetLastError();
}

cudaError_t launch_step(const KernelConfig &c, const StepArgs &a, cudaStream_t s) {
  if (a.i_count == 0) return cudaSuccess;
  if (c.family == 0) {
    if (c.self_mode == kSelfPredicated)
      return c.mass ? launch_scalar<1, 128, kSelfPredicated, true>(a, s) : launch_scalar<1, 128, kSelfPredicated>(a, s);
    return c.mass ? launch_scalar<1, 128, kSelfBranch, true>(a, s) : launch_scalar<1, 128, kSelfBranch>(a, s);
  }
  if (c.family == 6 && !c.mass) {  // small-N scalar warp-streaming kernel
    const uint32_t per = 32u * (uint32_t)c.r;
    const uint32_t ctas = (a.i_count + per - 1) / per;
    if (c.r == 1) force_wsmall_kernel<1><<<ctas, 32, 0, s>>>(a);
    else if (c.r == 2) force_wsmall_kernel<2><<<ctas, 32, 0, s>>>(a);
    else return cudaErrorInvalidConfiguration;
    return cudaGetLastError();
  }
  if (c.family == 5 && a.acc && a.progress && a.epoch && !c.mass) {  // TMA-staged comparison variant
    if (c.r == 6) return launch_wseg_tma<6, 14>(a, c.sms, a.progress, a.epoch, s);
    if (c.r == 4) return launch_wseg_tma<4, 20>(a, c.sms, a.progress, a.epoch, s);
    return cudaErrorInvalidConfiguration;
  }
  if (c.family == 4 && a.acc && a.progress && a.epoch) {  // j-segmented warp-streaming launch
    // (R, resident warps per SM promised to ptxas): the three tuned points, profiles/r01_tuning_log.txt
    if (c.r == 2) return c.mass ? launch_wseg_mb<2, 28, true>(a, c.sms, a.progress, a.epoch, s) : launch_wseg_mb<2, 28, false>(a, c.sms, a.progress, a.epoch, s);
    if (c.r == 4) return c.mass ? launch_wseg_mb<4, 20, true>(a, c.sms, a.progress, a.epoch, s) : launch_wseg_mb<4, 20, false>(a, c.sms, a.progress, a.epoch, s);
    if (c.r == 6) return c.mass ? launch_wseg_mb<6, 14, true>(a, c.sms, a.progress, a.epoch, s) : launch_wseg_mb<6, 14, false>(a, c.sms, a.progress, a.epoch, s);
  }
  if (c.family == 3 || c.family == 4) {  // unsegmented (also: caller-owned memory entry, no hand-off buffers)
#define NB_WSTREAM(RR, WW)                                   \
  if (c.r == RR && (c.block == 32 * WW || (c.family == 4 && WW == 1))) \
    return c.mass ? launch_wstream<RR, WW, true>(a, c.sms, s) : launch_wstream<RR, WW, false>(a, c.sms, s);
    NB_WSTREAM(2, 1)
    NB_WSTREAM(4, 1)
    NB_WSTREAM(6, 1)
    NB_WSTREAM(8, 1)
    NB_WSTREAM(4, 2)
    NB_WSTREAM(4, 4)
#undef NB_WSTREAM
  }
  if (c.mass) return cudaErrorInvalidConfiguration;  // masses: AUTO / GENERIC kernels only
#define NB_PACKED(RR, BB) \
  if (c.family == 1 && c.r == RR && c.block == BB) return launch_packed<RR, BB>(a, s);
#define NB_SCALAR(RR, BB) \
  if (c.family == 2 && c.r == RR && c.block == BB) return launch_scalar<RR, BB, kSelfNone>(a, s);
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
