// ubench_fma2.cu -- micro-benchmarks that pin the B200 ceilings the force kernel is judged
// against: issue rate of FFMA / FFMA2 / FADD2(.F32 broadcast) / MUFU.RSQ per SMSP, and the
// arithmetic-only rate of the 12-op interaction body (no memory) vs resident warps per SM.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o bin/ubench_fma2 tools/ubench_fma2.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

typedef unsigned long long u64;
__device__ __forceinline__ u64 pack2(float lo, float hi) { u64 d; asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi)); return d; }
__device__ __forceinline__ void unpack2(u64 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 fadd2(u64 a, u64 b) { u64 d; asm("add.rn.ftz.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 fmul2(u64 a, u64 b) { u64 d; asm("mul.rn.ftz.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.ftz.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ float ffma(float a, float b, float c) { float d; asm("fma.rn.ftz.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ float fadd(float a, float b) { float d; asm("add.rn.ftz.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d; }
__device__ __forceinline__ float fmul(float a, float b) { float d; asm("mul.rn.ftz.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d; }
__device__ __forceinline__ float frsq(float a) { float d; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(d) : "f"(a)); return d; }

#define CHAINS 8

// mode 0: FFMA2 x CHAINS ; 1: FFMA x 2*CHAINS ; 2: FADD2 with .F32 broadcast ; 3: MUFU only ;
// 4: 6 FFMA2 : 1 MUFU mix (kernel ratio 12:2) ; 5: FMUL2
template <int MODE>
__global__ void k_pipe(float *out, int iters, float s0, float s1, long long *cyc) {
  u64 acc[CHAINS];
  float f[2 * CHAINS];
  const u64 a = pack2(s0, s1), b = pack2(s1, s0);
#pragma unroll
  for (int c = 0; c < CHAINS; c++) { acc[c] = pack2(threadIdx.x + c, c); f[2 * c] = threadIdx.x + c; f[2 * c + 1] = c + 0.5f; }
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int c = 0; c < CHAINS; c++) {
      if (MODE == 0) acc[c] = ffma2(a, b, acc[c]);
      if (MODE == 1) { f[2 * c] = ffma(s0, s1, f[2 * c]); f[2 * c + 1] = ffma(s1, s0, f[2 * c + 1]); }
      if (MODE == 2) acc[c] = fadd2(pack2(s0, s0), acc[c]);
      if (MODE == 3) { f[c] = frsq(f[c]); }
      if (MODE == 4) {
        acc[c] = ffma2(a, b, acc[c]); acc[c] = ffma2(b, a, acc[c]); acc[c] = ffma2(a, a, acc[c]);
        acc[c] = ffma2(b, b, acc[c]); acc[c] = ffma2(a, b, acc[c]); acc[c] = ffma2(b, a, acc[c]);
        f[c] = frsq(f[c]);
      }
      if (MODE == 5) acc[c] = fmul2(a, acc[c]);
      if (MODE == 6) acc[c] = fadd2(pack2(f[c], f[c]), acc[c]);  // R.F32 broadcast from a VECTOR register
      if (MODE == 7) acc[c] = ffma2(pack2(f[c], f[c]), a, acc[c]);
    }
  }
  long long t1 = clock64();
  float r = 0;
#pragma unroll
  for (int c = 0; c < CHAINS; c++) { float lo, hi; unpack2(acc[c], lo, hi); r += lo + hi + f[2 * c] + f[2 * c + 1]; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// the interaction body with the j-body coming from registers (no LDS): arithmetic-only ceiling.
// NP pairs per thread (R = 2*NP i-bodies); PACKED selects f32x2 vs scalar ops.
template <int NP, bool PACKED, bool WITH_MUFU>
__global__ void k_body(float *out, int iters, float qx, float qy, float qz, float eps, long long *cyc) {
  u64 nx[NP], ny[NP], nz[NP], ax[NP], ay[NP], az[NP];
#pragma unroll
  for (int p = 0; p < NP; p++) {
    nx[p] = pack2(threadIdx.x + p, threadIdx.x - p); ny[p] = pack2(p + 1.f, p + 2.f); nz[p] = pack2(0.5f * p, 3.f);
    ax[p] = ay[p] = az[p] = 0ull;
  }
  const u64 eps2 = pack2(eps, eps);
  long long t0 = clock64();
#pragma unroll 8
  for (int it = 0; it < iters; it++) {
    float jx = qx + it, jy = qy, jz = qz;  // varies so nothing is hoisted
    if (PACKED) {
      const u64 X = pack2(jx, jx), Y = pack2(jy, jy), Z = pack2(jz, jz);
#pragma unroll
      for (int p = 0; p < NP; p++) {
        u64 rx = fadd2(X, nx[p]), ry = fadd2(Y, ny[p]), rz = fadd2(Z, nz[p]);
        u64 t = fmul2(ry, ry); t = ffma2(rx, rx, t); t = ffma2(rz, rz, t);
        u64 d = fadd2(t, eps2);
        u64 c = fmul2(d, d); c = fmul2(d, c);
        float c0, c1; unpack2(c, c0, c1);
        u64 w = WITH_MUFU ? pack2(frsq(c0), frsq(c1)) : c;
        ax[p] = ffma2(rx, w, ax[p]); ay[p] = ffma2(ry, w, ay[p]); az[p] = ffma2(rz, w, az[p]);
      }
    } else {
#pragma unroll
      for (int p = 0; p < NP; p++) {
#pragma unroll
        for (int h = 0; h < 2; h++) {
          float nxx, nxy, nyx, nyy, nzx, nzy, axx, axy, ayx, ayy, azx, azy;
          unpack2(nx[p], nxx, nxy); unpack2(ny[p], nyx, nyy); unpack2(nz[p], nzx, nzy);
          unpack2(ax[p], axx, axy); unpack2(ay[p], ayx, ayy); unpack2(az[p], azx, azy);
          float rx = fadd(jx, h ? nxy : nxx), ry = fadd(jy, h ? nyy : nyx), rz = fadd(jz, h ? nzy : nzx);
          float t = fmul(ry, ry); t = ffma(rx, rx, t); t = ffma(rz, rz, t);
          float d = fadd(t, eps);
          float c = fmul(d, d); c = fmul(d, c);
          float w = WITH_MUFU ? frsq(c) : c;
          if (h) { axy = ffma(rx, w, axy); ayy = ffma(ry, w, ayy); azy = ffma(rz, w, azy); }
          else { axx = ffma(rx, w, axx); ayx = ffma(ry, w, ayx); azx = ffma(rz, w, azx); }
          ax[p] = pack2(axx, axy); ay[p] = pack2(ayx, ayy); az[p] = pack2(azx, azy);
        }
      }
    }
  }
  long long t1 = clock64();
  float r = 0;
#pragma unroll
  for (int p = 0; p < NP; p++) { float lo, hi; unpack2(ax[p], lo, hi); r += lo + hi; unpack2(ay[p], lo, hi); r += lo + hi; unpack2(az[p], lo, hi); r += lo + hi; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// same body, j-body read from a 32-entry shared-memory tile with a warp-uniform (broadcast) address
// SRC: 0 = LDS.128 broadcast, 1 = per-lane register tile + 3 SHFL, 2 = LDS.128 + tile refilled from
// global every 32 j (the production warp-streaming loop)
template <int NP, int SRC, bool WITH_MUFU = true, int PF = 0, int U = 8, int VAR = 0>
__global__ void k_body_mem(float *out, const float4 *gpos, int iters, float eps, long long *cyc) {
  __shared__ __align__(16) float4 tile[16][2][32];
  __shared__ __align__(16) ulonglong2 dxy[16][32];
  __shared__ __align__(8) u64 dz[16][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  u64 nx[NP], ny[NP], nz[NP], ax[NP], ay[NP], az[NP];
#pragma unroll
  for (int p = 0; p < NP; p++) {
    nx[p] = pack2(threadIdx.x + p, threadIdx.x - p); ny[p] = pack2(p + 1.f, p + 2.f); nz[p] = pack2(0.5f * p, 3.f);
    ax[p] = ay[p] = az[p] = 0ull;
  }
  const u64 eps2 = pack2(eps, eps);
  tile[warp][0][lane] = gpos[lane];
  tile[warp][1][lane] = gpos[32 + lane];
  float4 mine = gpos[lane];
  dxy[warp][lane] = make_ulonglong2(pack2(mine.x, mine.x), pack2(mine.y, mine.y));
  dz[warp][lane] = pack2(mine.z, mine.z);
  __syncwarp();
  for (int t = 0; t < iters / 32; t++) {
    const int buf = t & 1;
    float4 nxt;
    if (SRC == 2) nxt = gpos[((t + 1) * 32 + lane) & 0xffff];
    float4 qq[PF + 1];
#pragma unroll
    for (int f = 0; f < PF; f++) qq[f] = tile[warp][buf][f];
#pragma unroll U
    for (int j = 0; j < 32; j++) {
      float4 q;
      if (PF > 0) {
        // explicit register pipeline: the LDS for j+PF is issued before the math of j
        q = qq[0];
#pragma unroll
        for (int f = 0; f + 1 < PF; f++) qq[f] = qq[f + 1];
        qq[PF - 1] = tile[warp][buf][(j + PF) & 31];
      } else if (SRC == 3) {
        q = make_float4(0, 0, 0, 0);
      } else if (SRC == 1) {
        q.x = __shfl_sync(0xffffffffu, mine.x, j); q.y = __shfl_sync(0xffffffffu, mine.y, j); q.z = __shfl_sync(0xffffffffu, mine.z, j);
      } else {
        q = tile[warp][buf][j];
      }
      u64 X = pack2(q.x, q.x), Y = pack2(q.y, q.y), Z = pack2(q.z, q.z);
      if (SRC == 3) { ulonglong2 xy = dxy[warp][j]; X = xy.x; Y = xy.y; Z = dz[warp][j]; }
#pragma unroll
      for (int p = 0; p < NP; p++) {
        u64 rx = fadd2(X, nx[p]), ry = fadd2(Y, ny[p]), rz = fadd2(Z, nz[p]);
        u64 tt = fmul2(ry, ry); tt = ffma2(rx, rx, tt); tt = ffma2(rz, rz, tt);
        u64 d = fadd2(tt, eps2);
        u64 c = fmul2(d, d); c = fmul2(d, c);
        float c0, c1; unpack2(c, c0, c1);
        u64 w = WITH_MUFU ? pack2(frsq(c0), frsq(c1)) : c;
        if (VAR == 0) { ax[p] = ffma2(rx, w, ax[p]); ay[p] = ffma2(ry, w, ay[p]); az[p] = ffma2(rz, w, az[p]); }
        if (VAR == 1) { ax[p] = ffma2(rx, rx, ax[p]); ay[p] = ffma2(ry, ry, ay[p]); az[p] = ffma2(w, w, az[p]); }
        if (VAR == 2) { ax[p] = fadd2(ax[p], w); ay[p] = fadd2(ay[p], rx); az[p] = fadd2(az[p], fadd2(ry, rz)); }
        if (VAR == 3) { ax[p] = ffma2(w, rx, ax[p]); ay[p] = ffma2(w, ry, ay[p]); az[p] = ffma2(w, rz, az[p]); }
      }
    }
    if (SRC == 2) { tile[warp][buf ^ 1][lane] = nxt; __syncwarp(); }
  }
  float r = 0;
#pragma unroll
  for (int p = 0; p < NP; p++) { float lo, hi; unpack2(ax[p], lo, hi); r += lo + hi; unpack2(ay[p], lo, hi); r += lo + hi; unpack2(az[p], lo, hi); r += lo + hi; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

static float *d_out;
static long long *d_cyc;

// returns SM cycles of the whole launch (CUDA-event time x 1.965 GHz: clocks are at max under
// this load, see profiles/), so that rates are whole-GPU averages, ramp and tail included
template <typename F>
static double run(F launch, int grid, int block) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  launch();  // warm
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 3; r++) {
    cudaEventRecord(e0);
    launch();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return (double)best * 1e-3 * 1.965e9;
}

int main() {
  cudaMalloc(&d_out, 148 * 64 * 1024 * sizeof(float));
  cudaMalloc(&d_cyc, 148 * 64 * sizeof(long long));
  const int iters = 1 << 17;
  { k_pipe<0><<<148 * 16, 128>>>(d_out, 1 << 20, 1.0001f, 0.9999f, d_cyc); cudaDeviceSynchronize(); }  // clock warm-up
  printf("# pipe tests: per-SMSP instruction rate (warp-instr / cycle), CHAINS=%d independent chains per thread\n", CHAINS);
  const char *names[] = {"FFMA2", "FFMA(x2)", "FADD2.URbcast", "MUFU.RSQ", "6xFFMA2+1MUFU", "FMUL2", "FADD2.Rbcast", "FFMA2.Rbcast"};
  for (int wps = 1; wps <= 16; wps *= 2) {  // warps per SMSP
    int block = 128, grid = 148 * wps;     // each block = 4 warps = 1 per SMSP
    double c[8];
    c[0] = run([&] { k_pipe<0><<<grid, block>>>(d_out, iters, 1.0001f, 0.9999f, d_cyc); }, grid, block);
    c[1] = run([&] { k_pipe<1><<<grid, block>>>(d_out, iters, 1.0001f, 0.9999f, d_cyc); }, grid, block);
    c[2] = run([&] { k_pipe<2><<<grid, block>>>(d_out, iters, 1.0001f, 0.9999f, d_cyc); }, grid, block);
    c[3] = run([&] { k_pipe<3><<<grid, block>>>(d_out, iters, 1.0001f, 0.9999f, d_cyc); }, grid, block);
    c[4] = run([&] { k_pipe<4><<<grid, block>>>(d_out, iters, 1.0001f, 0.9999f, d_cyc); }, grid, block);
    c[5] = run([&] { k_pipe<5><<<grid, block>>>(d_out, iters, 1.0001f, 0.9999f, d_cyc); }, grid, block);
    c[6] = run([&] { k_pipe<6><<<grid, block>>>(d_out, iters, 1.0001f, 0.9999f, d_cyc); }, grid, block);
    c[7] = run([&] { k_pipe<7><<<grid, block>>>(d_out, iters, 1.0001f, 0.9999f, d_cyc); }, grid, block);
    double n_inst[8] = {CHAINS, 2.0 * CHAINS, CHAINS, CHAINS, 7.0 * CHAINS, CHAINS, CHAINS, CHAINS};
    for (int m = 0; m < 8; m++)
      printf("warps/SMSP=%2d %-14s cycles/iter=%8.2f  warp-inst/clk/SMSP=%.3f\n", wps, names[m], c[m] / iters,
             n_inst[m] * iters * wps / c[m]);
  }
  printf("# interaction body, j from registers (no LDS): cycles per warp-level j-step and interactions/clk/SM\n");
  for (int wps = 1; wps <= 8; wps++) {
    int block = 128, grid = 148 * wps;
    double c;
#define BODY(NP, PK, MU, label)                                                                              \
  c = run([&] { k_body<NP, PK, MU><<<grid, block>>>(d_out, iters, 1.f, 2.f, 3.f, 1e-7f, d_cyc); }, grid, block); \
  printf("warps/SMSP=%d %-22s cycles/j=%7.2f  inter/clk/SM=%6.3f  (%%of 10.667 pipe bound: %5.1f)\n", wps, label, \
         c / iters, 4.0 * wps * 32 * 2 * NP * iters / c, 100.0 * (4.0 * wps * 32 * 2 * NP * iters / c) / 10.6667);
    BODY(1, true, true, "packed R=2")
    BODY(2, true, true, "packed R=4")
    BODY(3, true, true, "packed R=6")
    BODY(4, true, true, "packed R=8")
    BODY(2, true, false, "packed R=4 noMUFU")
    BODY(2, false, true, "scalar R=4")
    BODY(2, false, false, "scalar R=4 noMUFU")
  }
  printf("# interaction body fed from memory (packed): 0=LDS.128 bcast static tile, 1=register tile+SHFL, 2=LDS.128 + LDG/STS refill\n");
  float4 *d_pos;
  cudaMalloc(&d_pos, 65536 * sizeof(float4));
  cudaMemset(d_pos, 0, 65536 * sizeof(float4));
  for (int wps = 4; wps <= 12; wps += 2) {
    int block = 128, grid = 148 * wps;
    double c;
#define BODYM(NP, SRC, label) BODYMX(NP, SRC, true, 0, 8, label)
#define BODYMX(NP, SRC, MU, PF, UU, label) BODYMV(NP, SRC, MU, PF, UU, 0, label)
#define BODYMV(NP, SRC, MU, PF, UU, VV, label)                                                                              \
  c = run([&] { k_body_mem<NP, SRC, MU, PF, UU, VV><<<grid, block>>>(d_out, d_pos, iters, 1e-7f, d_cyc); }, grid, block); \
  printf("warps/SMSP=%d %-22s cycles/j=%7.2f  inter/clk/SM=%6.3f  (%%of 10.667 pipe bound: %5.1f)\n", wps, label, \
         c / iters, 4.0 * wps * 32 * 2 * NP * iters / c, 100.0 * (4.0 * wps * 32 * 2 * NP * iters / c) / 10.6667);
    BODYM(1, 0, "R=2 LDS")
    BODYM(2, 0, "R=4 LDS")
    BODYM(4, 0, "R=8 LDS")
    BODYM(2, 1, "R=4 SHFL")
    BODYMX(2, 0, false, 0, 8, "R=4 LDS noMUFU")
    BODYMX(2, 0, true, 2, 8, "R=4 LDS prefetch2")
    BODYMX(2, 0, true, 4, 8, "R=4 LDS prefetch4")
    BODYMX(1, 2, true, 0, 1, "R=2 refill U1")
    BODYMX(1, 2, true, 0, 2, "R=2 refill U2")
    BODYMX(1, 2, true, 0, 4, "R=2 refill U4")
    BODYMX(2, 2, true, 0, 1, "R=4 refill U1")
    BODYMX(2, 2, true, 0, 2, "R=4 refill U2")
    BODYMX(2, 2, true, 0, 4, "R=4 refill U4")
    BODYMX(2, 2, true, 0, 16, "R=4 refill U16")
    BODYMX(2, 2, true, 0, 32, "R=4 refill U32")
    BODYMV(2, 0, true, 0, 32, 0, "V0 R=4 U32 baseline")
    BODYMV(2, 0, true, 0, 32, 1, "V1 acc 2-operand")
    BODYMV(2, 0, true, 0, 32, 2, "V2 acc as FADD2")
    BODYMV(2, 0, true, 0, 32, 3, "V3 acc w in slot A")
    BODYMV(2, 0, false, 0, 32, 1, "V1 noMUFU")
    BODYM(1, 3, "R=2 dupLDS")
    BODYM(2, 3, "R=4 dupLDS")
    BODYM(3, 3, "R=6 dupLDS")
    BODYM(4, 3, "R=8 dupLDS")
    BODYM(1, 2, "R=2 LDS+refill")
    BODYM(2, 2, "R=4 LDS+refill")
    BODYM(3, 2, "R=6 LDS+refill")
    BODYM(4, 2, "R=8 LDS+refill")
  }
  return 0;
}
