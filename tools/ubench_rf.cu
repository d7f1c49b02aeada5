// ubench_rf.cu -- how the register-file operand stage and the FMA pipe compose on B200.
// Question behind it (round 2): the production loop costs ~28.2 cycles per pair-interaction while
// its FMA-pipe time is 24.  The RF-banking rule (one register per even/odd bank per cycle; an FFMA2
// with three distinct 64-bit operands needs 3 cycles) explains the gap only if RF time and pipe
// time do NOT overlap between neighbouring instructions.  The modes below measure exactly that:
// do light instructions (1 RF cycle, 2 pipe cycles) absorb the extra RF cycle of heavy ones, is a
// MUFU free in the shadow of a packed op, and does `.reuse` make the 2nd/3rd accumulate of a
// triplet a 2-cycle instruction.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o bin/ubench_rf tools/ubench_rf.cu
#include <cuda_runtime.h>
#include <stdio.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pack2(float lo, float hi) { u64 d; asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi)); return d; }
__device__ __forceinline__ void unpack2(u64 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 fmul2(u64 a, u64 b) { u64 d; asm volatile("mul.rn.ftz.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.ftz.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ float frsq(float a) { float d; asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(d) : "f"(a)); return d; }

#define C 6
template <int MODE>
__global__ void __launch_bounds__(128) k(float *out, const float *in, int iters) {
  u64 x[C], y[C], w[C], a0[C], a1[C], a2[C], z[C];
  float f[C];
#pragma unroll
  for (int c = 0; c < C; c++) {
    const float *p = in + threadIdx.x + 8 * c;
    x[c] = pack2(p[0], p[1]); y[c] = pack2(p[2], p[3]); w[c] = pack2(p[4], p[5]);
    a0[c] = pack2(p[6], p[7]); a1[c] = pack2(p[8], p[9]); a2[c] = pack2(p[10], p[11]);
    z[c] = pack2(p[12], p[13]); f[c] = p[14];
  }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 2; u++) {
#pragma unroll
      for (int c = 0; c < C; c++) {
        if (MODE == 0) { a0[c] = ffma2(x[c], w[c], a0[c]); }                                   // heavy: 3 RF / 2 pipe
        if (MODE == 1) { z[c] = fmul2(z[c], z[c]); }                                           // light: 1 RF / 2 pipe
        if (MODE == 2) { a0[c] = ffma2(x[c], w[c], a0[c]); z[c] = fmul2(z[c], z[c]); }          // heavy + light
        if (MODE == 3) { a0[c] = ffma2(x[c], w[c], a0[c]); z[c] = fmul2(z[c], z[c]); y[c] = fmul2(y[c], y[c]); }  // heavy + 2 light
        if (MODE == 4) { a0[c] = ffma2(x[c], x[c], a0[c]); }                                   // medium: 2 RF / 2 pipe
        if (MODE == 5) { a0[c] = ffma2(x[c], w[c], a0[c]); a1[c] = ffma2(y[c], w[c], a1[c]); a2[c] = ffma2(z[c], w[c], a2[c]); }  // accumulate triplet, shared w
        if (MODE == 6) { a0[c] = ffma2(x[c], w[c], a0[c]); x[c] = fmul2(x[c], x[c]); a1[c] = ffma2(y[c], w[c], a1[c]); y[c] = fmul2(y[c], y[c]); a2[c] = ffma2(z[c], w[c], a2[c]); z[c] = fmul2(z[c], z[c]); }  // triplet broken by light ops
      }
      // 6 FP2 : 1 MUFU groups (the kernel's ratio); MUFU chains are independent of the FP2 chains
      if (MODE == 7) { _Pragma("unroll") for (int c = 0; c < C; c++) z[c] = fmul2(z[c], z[c]); f[u] = frsq(f[u]); }
      if (MODE == 8) { _Pragma("unroll") for (int c = 0; c < C; c++) a0[c] = ffma2(x[c], x[c], a0[c]); f[u] = frsq(f[u]); }
      if (MODE == 9) { _Pragma("unroll") for (int c = 0; c < C; c++) a0[c] = ffma2(x[c], w[c], a0[c]); f[u] = frsq(f[u]); }
      // kernel-like mix per "pair": 3 FADD2-like (2 RF), 1 light, 2 medium, 1 medium, 1 light, 1 medium, 2 MUFU, 3 heavy
      if (MODE == 10 || MODE == 11) {
#pragma unroll
        for (int c = 0; c < C; c += 2) {
          x[c] = ffma2(x[c], x[c], y[c]); y[c] = ffma2(y[c], y[c], z[c]); z[c] = ffma2(z[c], z[c], x[c]);   // 3 medium
          z[c + 1] = fmul2(z[c + 1], z[c + 1]);                                                               // light
          x[c + 1] = ffma2(x[c + 1], x[c + 1], y[c + 1]); y[c + 1] = ffma2(y[c + 1], y[c + 1], z[c + 1]);    // 2 medium
          w[c + 1] = fmul2(w[c + 1], w[c + 1]);                                                               // light
          if (MODE == 11) { f[c] = frsq(f[c]); f[c + 1] = frsq(f[c + 1]); }
          a0[c] = ffma2(x[c], w[c], a0[c]); a1[c] = ffma2(y[c], w[c], a1[c]); a2[c] = ffma2(z[c], w[c], a2[c]);  // triplet
          x[c] = fmul2(x[c], y[c]); y[c] = fmul2(y[c], z[c]);                                                  // 2 medium (FMUL2 2 regs)
        }
      }
    }
  }
  float r = 0;
#pragma unroll
  for (int c = 0; c < C; c++) {
    float lo, hi;
    unpack2(x[c], lo, hi); r += lo + hi; unpack2(y[c], lo, hi); r += lo + hi; unpack2(w[c], lo, hi); r += lo + hi;
    unpack2(a0[c], lo, hi); r += lo + hi; unpack2(a1[c], lo, hi); r += lo + hi; unpack2(a2[c], lo, hi); r += lo + hi;
    unpack2(z[c], lo, hi); r += lo + hi; r += f[c];
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <typename F> static double run(F launch) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  launch(); cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 3; r++) { cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
  return (double)best * 1e-3 * 1.965e9;
}

int main() {
  float *d_out, *d_in;
  cudaMalloc(&d_out, 148 * 16 * 128 * sizeof(float));
  cudaMalloc(&d_in, 4096 * sizeof(float));
  cudaMemset(d_in, 0, 4096 * sizeof(float));
  const int iters = 1 << 14;
  k<1><<<148 * 8, 128>>>(d_out, d_in, 1 << 17); cudaDeviceSynchronize();  // warm the clocks
  struct M { const char *name; double fp2_per_body; double expect_inelastic, expect_elastic; };
  // per inner body (one c, or one group): number of FP2 instructions, and the two model predictions in cycles per body
  const M m[12] = {
      {"heavy  FFMA2 x*w+a (3 RF)", 1, 3, 3},
      {"light  FMUL2 z*z   (1 RF)", 1, 2, 2},
      {"heavy + light", 2, 5, 4},
      {"heavy + 2 light", 3, 7, 6},
      {"medium FFMA2 x*x+a (2 RF)", 1, 2, 2},
      {"triplet shared w, adjacent", 3, 7, 7},
      {"triplet broken by light ops", 6, 15, 12},
      {"6 light  + 1 MUFU", 6, 13, 12},
      {"6 medium + 1 MUFU", 6, 13, 12},
      {"6 heavy  + 1 MUFU", 6, 19, 18},
      {"kernel-like pair mix, no MUFU", 12, 25, 24},
      {"kernel-like pair mix, 2 MUFU", 12, 27, 24},
  };
  for (int wps = 2; wps <= 8; wps *= 2) {
    int grid = 148 * wps, block = 128;
    double cyc[12];
    cyc[0] = run([&] { k<0><<<grid, block>>>(d_out, d_in, iters); });
    cyc[1] = run([&] { k<1><<<grid, block>>>(d_out, d_in, iters); });
    cyc[2] = run([&] { k<2><<<grid, block>>>(d_out, d_in, iters); });
    cyc[3] = run([&] { k<3><<<grid, block>>>(d_out, d_in, iters); });
    cyc[4] = run([&] { k<4><<<grid, block>>>(d_out, d_in, iters); });
    cyc[5] = run([&] { k<5><<<grid, block>>>(d_out, d_in, iters); });
    cyc[6] = run([&] { k<6><<<grid, block>>>(d_out, d_in, iters); });
    cyc[7] = run([&] { k<7><<<grid, block>>>(d_out, d_in, iters); });
    cyc[8] = run([&] { k<8><<<grid, block>>>(d_out, d_in, iters); });
    cyc[9] = run([&] { k<9><<<grid, block>>>(d_out, d_in, iters); });
    cyc[10] = run([&] { k<10><<<grid, block>>>(d_out, d_in, iters); });
    cyc[11] = run([&] { k<11><<<grid, block>>>(d_out, d_in, iters); });
    for (int i = 0; i < 12; i++) {
      // bodies per thread-iteration: modes 0-6: 2*C ; 7-9: 2 groups ; 10-11: 2 * C/2 pairs
      double bodies = (i <= 6 ? 2.0 * C : (i <= 9 ? 2.0 : 2.0 * (C / 2))) * iters * wps;
      printf("warps/SMSP=%d %-32s cycles/body = %6.3f   (model: inelastic %g, elastic %g; %g FP2 instr)\n", wps, m[i].name,
             cyc[i] / bodies, m[i].expect_inelastic, m[i].expect_elastic, m[i].fp2_per_body);
    }
  }
  return 0;
}
