// nbody_body.cuh -- the arithmetic of one body-body interaction and of the integration tail,
// written ONCE and used by every kernel of the library (production and comparison variants).
//
// What is computed is fixed by the reference kernel particle_interaction<> (src/simulator.cu:186-229)
// compiled with -use_fast_math; its sm_100a SASS does, per pair,
//     r  = p_j + (-p_i)                      3 FADD.FTZ
//     t  = ry*ry ; t = fma(rx,rx,t) ; t = fma(rz,rz,t)     FMUL + 2 FFMA   (y first)
//     d  = t + distEps                       FADD   (softening is added to r^2, :201)
//     c  = d * (d*d)                         2 FMUL
//     w  = MUFU.RSQ(c)
//     a  = fma(r, w, a)                      3 FFMA, skipped when j == i          (BRANCH, :206-207)
//     a  = fma(r*w, (j == i), a)             FMUL + FFMA                          (PREDICATED as shipped, :209)
//     a  = fma(r*w, (j != i), a)             FMUL + FFMA                          (README-intended PREDICATED)
// with one accumulator per component and j ascending.  FP32 addition is not associative and the
// sums cancel heavily, so any other order differs from the reference by ~1e-5 relative at N = 262144
// (SURVEY.md section 0.3).  Everything below therefore keeps that exact op sequence -- spelled in PTX
// with explicit .rn.ftz so ptxas cannot re-contract it -- and is bit-identical to the reference.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "nbody_kernels.cuh"

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
// packed pairs: two independent IEEE lanes per instruction (sm_100+: FADD2 / FMUL2 / FFMA2)
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

// ---- one j-body against NP packed pairs of i-bodies (no self-term predicate) ------------------
// Two i-bodies share the two lanes of each f32x2 instruction: 12 instructions serve TWO interactions.
// The j-body is the scalar-broadcast operand of the instruction (SASS operand form `R.F32`), so
// pack2(q.x, q.x) costs nothing and the tile stays in its HBM float4 layout.
template <int NP, bool MASS>
__device__ __forceinline__ void interact_packed(const float4 q, const u64 eps2, const u64 (&nx)[NP],
                                                const u64 (&ny)[NP], const u64 (&nz)[NP], u64 (&ax)[NP],
                                                u64 (&ay)[NP], u64 (&az)[NP]) {
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
}

// ---- one j-body against R scalar i-bodies, with the reference's self-term variants ------------
template <int R, int SELF, bool MASS>
__device__ __forceinline__ void interact_scalar(const float4 q, const uint32_t gj, const float eps,
                                                const float (&nx)[R], const float (&ny)[R], const float (&nz)[R],
                                                const uint32_t (&gi)[R], float (&ax)[R], float (&ay)[R],
                                                float (&az)[R]) {
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
    } else {
      // as shipped: force += r * inv * (i == id);  README-intended: ... * (i != id)
      const bool same = gj == gi[k];
      const float sel = (SELF == kSelfPredicated ? same : !same) ? 1.0f : 0.0f;
      ax[k] = ffma(fmul(rx, w), sel, ax[k]);
      ay[k] = ffma(fmul(ry, w), sel, ay[k]);
      az[k] = ffma(fmul(rz, w), sel, az[k]);
    }
  }
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

// ---- j-segment hand-off ------------------------------------------------------------------------
// Wait (with back-off) until the predecessor unit of this body group has published its
// accumulators.  Work units are numbered by an atomic TICKET taken when the CTA starts running
// (not by blockIdx), so the predecessor -- which holds a smaller ticket -- is by construction
// already running or finished, whatever order the hardware dispatches CTAs in: forward progress
// does not depend on dispatch order (the decoupled look-back construction).  A legitimate wait is
// at most one unit long.  If a wait nevertheless exceeds the time-out (20 s by default; debugger,
// time-slicing, a sanitizer slowing the kernel 1000x: NBODY_HANDOFF_TIMEOUT_S, 0 = never) the kernel
// raises the host-visible error word and carries on -- nbody_step() then reports NBODY_E_STATE
// instead of the context being torn down by a trap.
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ void wait_for_segment(unsigned int *word, unsigned int target, unsigned int *error,
                                                 const unsigned long long timeout_ns) {
  volatile unsigned int *p = word;
  if (*p != target) {
    const unsigned long long t0 = global_ns();
    unsigned int spins = 0;
    while (*p != target) {
      __nanosleep(128);
      if ((++spins & 0x3fffu) == 0 && timeout_ns && global_ns() - t0 > timeout_ns) {
        *(volatile unsigned int *)error = 1u;
        __threadfence_system();
        break;
      }
    }
  }
  __threadfence();
}

// ---- warp-private j-tiles ----------------------------------------------------------------------
// Every WARP stages its own 32-body j-tiles: one coalesced LDG.128 per lane -> STS.128 -> __syncwarp
// -> 32 broadcast LDS.128, double buffered; no CTA-wide barrier, warps never wait for each other.
// `body(q, gj)` is called for every j of [j_begin, j_end) in ascending order.
template <class F>
__device__ __forceinline__ void sweep_warp_tiles(const float4 *__restrict__ pos, const uint32_t j_begin,
                                                 const uint32_t j_end, float4 (*tile)[32], const int lane,
                                                 F body) {
  constexpr int TJ = 32;
  const uint32_t nj = j_end - j_begin;
  const uint32_t ntiles = (nj + TJ - 1) / TJ;
  if (ntiles == 0) return;
#ifndef NB_LOOP_BRANCHY
  // The per-tile bookkeeping competes with the arithmetic for issue slots (every instruction costs the
  // sub-partition at least one issue cycle, tools/sass_lab.py), so it is kept branch-free: the next tile is
  // always fetched (index clamped to the last j-body, a harmless re-read behind the last tile) and always
  // stored (behind the last tile nobody reads it).
  const uint32_t last = j_end - 1;
  uint32_t jn = j_begin + lane;
  tile[0][lane] = pos[min(jn, last)];
  __syncwarp();
  for (uint32_t t = 0; t < ntiles; t++) {
    const int buf = t & 1;
    jn += TJ;
    const float4 nxt = pos[min(jn, last)];
    const uint32_t cnt = min((uint32_t)TJ, nj - t * TJ);
    const uint32_t gj0 = j_begin + t * TJ;
    if (cnt == TJ) {
#pragma unroll
      for (int j = 0; j < TJ; j++) body(tile[buf][j], gj0 + j);
    } else {
      for (uint32_t j = 0; j < cnt; j++) body(tile[buf][j], gj0 + j);
    }
    tile[buf ^ 1][lane] = nxt;
    __syncwarp();
  }
#else
  auto fetch = [&](uint32_t t) -> float4 {
    uint32_t j = j_begin + t * TJ + lane;
    return pos[j < j_end ? j : j_end - 1];
  };
  tile[0][lane] = fetch(0);
  __syncwarp();
  for (uint32_t t = 0; t < ntiles; t++) {
    const int buf = t & 1;
    float4 nxt;
    const bool more = t + 1 < ntiles;
    if (more) nxt = fetch(t + 1);
    const uint32_t cnt = min((uint32_t)TJ, nj - t * TJ);
    const uint32_t gj0 = j_begin + t * TJ;
    if (cnt == TJ) {
#pragma unroll
      for (int j = 0; j < TJ; j++) body(tile[buf][j], gj0 + j);
    } else {
      for (uint32_t j = 0; j < cnt; j++) body(tile[buf][j], gj0 + j);
    }
    if (more) tile[buf ^ 1][lane] = nxt;
    __syncwarp();
  }
#endif
}

// one warp's packed sweep: R*32 i-bodies starting at shard-local index warp_i against j in [j_begin, j_end)
template <int R, bool MASS>
__device__ __forceinline__ void warp_sweep_packed(const StepArgs &a, const uint32_t j_begin, const uint32_t j_end,
                                                  const int flags, const uint32_t warp_i, float4 (*tile)[32],
                                                  const int lane) {
  static_assert(R % 2 == 0, "packed kernel pairs i-bodies");
  constexpr int NP = R / 2;
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
  sweep_warp_tiles(a.pos, j_begin, j_end, tile, lane,
                   [&](const float4 q, uint32_t) { interact_packed<NP, MASS>(q, eps2, nx, ny, nz, ax, ay, az); });
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

// one warp's scalar sweep (small shards; the generic/faithful path with the self-term variants)
template <int R, int SELF, bool MASS>
__device__ __forceinline__ void warp_sweep_scalar(const StepArgs &a, const uint32_t j_begin, const uint32_t j_end,
                                                  const int flags, const uint32_t warp_i, float4 (*tile)[32],
                                                  const int lane) {
  float nx[R], ny[R], nz[R], ax[R], ay[R], az[R];
  float4 own[R];
  uint32_t gi[R];
#pragma unroll
  for (int k = 0; k < R; k++) {
    uint32_t li = warp_i + k * 32 + lane;
    uint32_t lc = li < a.i_count ? li : a.i_count - 1;
    gi[k] = a.i_begin + lc;
    own[k] = a.pos[gi[k]];
    nx[k] = -own[k].x;
    ny[k] = -own[k].y;
    nz[k] = -own[k].z;
    if (flags & kFirstChunk) {
      ax[k] = ay[k] = az[k] = 0.0f;
    } else {
      float4 c = __ldcg(&a.acc[lc]);
      ax[k] = c.x;
      ay[k] = c.y;
      az[k] = c.z;
    }
  }
  const float eps = a.eps;
  sweep_warp_tiles(a.pos, j_begin, j_end, tile, lane, [&](const float4 q, uint32_t gj) {
    interact_scalar<R, SELF, MASS>(q, gj, eps, nx, ny, nz, gi, ax, ay, az);
  });
#pragma unroll
  for (int k = 0; k < R; k++) {
    const uint32_t li = warp_i + k * 32 + lane;
    if (li >= a.i_count) continue;
    finish_body(a, flags, li, ax[k], ay[k], az[k], own[k]);
  }
}

// ---- accumulator relay (small shards) --------------------------------------------------------------
// With fewer i-bodies than the machine has lanes, one warp per 32 bodies leaves every SM sub-partition with at
// most one warp, and a single warp issues only ~0.6 instructions per cycle (profiles/r02_ncu_wscalar_*).  More
// warps cannot come from the bodies -- but they can come from the j-loop WITHOUT touching the summation order:
// W warps of one CTA serve the SAME 32 i-bodies and take the j-tiles round-robin.  For its tile a warp first
// computes everything that does not depend on the running sums -- the differences r and the weights
// w = rsqrt((r^2+eps)^3), 10 of the 13 operations -- into registers; then it takes the accumulators from the warp
// that handled the previous tile (through shared memory), runs the 3 x TJ accumulate FMAs in ascending j, and
// passes them on.  Every lane still owns one FP32 chain per component over ascending j: bit-identical.
// The serial part is the FMA chain alone (4 cycles per j) instead of the whole 13-op body.
// The hand-off uses one shared-memory mbarrier per warp ("your turn"): the warp that finishes tile t stores its
// lanes' sums and arrives (release) on the barrier of the warp that holds tile t + 1; that warp sleeps in
// mbarrier.try_wait (acquire) instead of spinning -- a spin loop would take issue slots from the warps of other
// CTAs that share the sub-partition (measured: profiles/r02_relay.txt).  Every lane arrives, so each lane's own
// store is ordered before the wake-up of the same lane of the next warp.
__device__ __forceinline__ void mbar_init(u64 *bar, unsigned int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned int)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(u64 *bar) {
  asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.release.cta.shared::cta.b64 st, [%0];\n}" ::"r"((unsigned int)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(u64 *bar, unsigned int parity) {
  asm volatile(
      "{\n .reg .pred p;\n"
      "RELAY_WAIT:\n"
      " mbarrier.try_wait.parity.acquire.cta.shared::cta.b64 p, [%0], %1;\n"
      " @p bra RELAY_DONE;\n"
      " bra RELAY_WAIT;\n"
      "RELAY_DONE:\n}" ::"r"((unsigned int)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
}

template <int W, int TJ, bool MASS>
__device__ __forceinline__ void cta_relay_scalar(const StepArgs &a, float4 (*tile)[TJ], float4 *tok, u64 *bar) {
  static_assert(TJ == 16 || TJ == 32, "one coalesced load per tile");
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const uint32_t li = blockIdx.x * 32u + lane;
  const uint32_t lc = li < a.i_count ? li : a.i_count - 1;
  const float4 own = a.pos[a.i_begin + lc];
  const float nx = -own.x, ny = -own.y, nz = -own.z;
  const float eps = a.eps;
  const uint32_t nj = a.j_end - a.j_begin;
  const uint32_t ntiles = (nj + TJ - 1) / TJ;
  float ax = 0.0f, ay = 0.0f, az = 0.0f;
  if (w == 0) {  // warp 0 takes tile 0 and with it the incoming sums
    if (!(a.flags & kFirstChunk)) {
      const float4 c = __ldcg(&a.acc[lc]);
      ax = c.x;
      ay = c.y;
      az = c.z;
    }
    if (ntiles == 0) {
      if (li < a.i_count) finish_body(a, a.flags, li, ax, ay, az, own);
      return;
    }
  }
  if (ntiles == 0) return;
  if (threadIdx.x < W) mbar_init(&bar[threadIdx.x], 32u);  // bar[w]: the 32 lanes holding the previous tile have arrived
  __syncthreads();
  unsigned int turns = 0;  // completed waits of this warp = phase parity of its barrier
  const uint32_t last = a.j_end - 1;
  const int jl = lane & (TJ - 1);
  float4 nxt = a.pos[min(a.j_begin + (uint32_t)w * TJ + jl, last)];
  for (uint32_t t = w; t < ntiles; t += W) {
    __syncwarp();  // every lane is done reading the previous tile
    if (TJ == 32 || lane < TJ) tile[w][jl] = nxt;
    nxt = a.pos[min(a.j_begin + (t + W) * TJ + jl, last)];  // always fetched (clamped): branch-free bookkeeping
    __syncwarp();
    float rx[TJ], ry[TJ], rz[TJ], wt[TJ];
#pragma unroll
    for (int j = 0; j < TJ; j++) {  // a ragged last tile computes its unused entries from the clamped re-read
      const float4 q = tile[w][j];
      rx[j] = fadd(q.x, nx);
      ry[j] = fadd(q.y, ny);
      rz[j] = fadd(q.z, nz);
      float s = fmul(ry[j], ry[j]);
      s = ffma(rx[j], rx[j], s);
      s = ffma(rz[j], rz[j], s);
      const float d = fadd(s, eps);
      float c = fmul(d, d);
      c = fmul(d, c);
      wt[j] = frsq(c);
      if (MASS) wt[j] = fmul(wt[j], q.w);
    }
    if (t > 0) {  // the sums arrive from the warp that handled tile t - 1
      mbar_wait(&bar[w], turns & 1u);
      turns++;
      const float4 k = tok[lane];
      ax = k.x;
      ay = k.y;
      az = k.z;
    }
    const uint32_t cnt = min((uint32_t)TJ, nj - t * TJ);
    if (cnt == TJ) {
#pragma unroll
      for (int j = 0; j < TJ; j++) {
        ax = ffma(rx[j], wt[j], ax);
        ay = ffma(ry[j], wt[j], ay);
        az = ffma(rz[j], wt[j], az);
      }
    } else {
#pragma unroll
      for (int j = 0; j < TJ; j++) {
        if (j < cnt) {
          ax = ffma(rx[j], wt[j], ax);
          ay = ffma(ry[j], wt[j], ay);
          az = ffma(rz[j], wt[j], az);
        }
      }
    }
    if (t + 1 < ntiles) {  // hand the sums to the warp that holds tile t + 1
      tok[lane] = make_float4(ax, ay, az, 0.0f);
      mbar_arrive(&bar[(w + 1) % W]);
    } else if (li < a.i_count) {
      finish_body(a, a.flags, li, ax, ay, az, own);
    }
  }
}

}  // namespace nbody
