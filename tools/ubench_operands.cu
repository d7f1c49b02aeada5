// ubench_operands.cu -- register-operand patterns of the FP32 pipe on B200: does an f32x2
// instruction with three distinct 64-bit VECTOR operands still issue every 2 cycles?
#include <cuda_runtime.h>
#include <stdio.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pack2(float lo, float hi) { u64 d; asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi)); return d; }
__device__ __forceinline__ void unpack2(u64 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 fadd2(u64 a, u64 b) { u64 d; asm("add.rn.ftz.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 fmul2(u64 a, u64 b) { u64 d; asm("mul.rn.ftz.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.ftz.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ float ffma(float a, float b, float c) { float d; asm("fma.rn.ftz.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }

#define C 6
// MODE 0: FFMA2 acc = x*y+acc, x,y,acc all distinct per chain (3 vector pairs, nothing reusable)
// MODE 1: FFMA2 acc = x*x+acc (2 distinct)        MODE 2: FMUL2 acc = x*acc (2 distinct)
// MODE 3: FFMA  3 distinct scalars (2*C chains)   MODE 4: FADD2 acc = bcast(s)+acc, s vector scalar
// MODE 5: FFMA2 acc = x*y+acc with x,y shared by all chains (reuse-friendly)
// MODE 6: FFMA2 acc = x*y+acc, x distinct per chain, y shared (2 fresh pairs)
template <int MODE>
__global__ void k(float *out, const float *in, int iters) {
  u64 x[C], y[C], acc[C];
  float fx[2 * C], fy[2 * C], fa[2 * C];
#pragma unroll
  for (int c = 0; c < C; c++) {
    x[c] = pack2(in[threadIdx.x + c], in[threadIdx.x + c + 7]);
    y[c] = pack2(in[threadIdx.x + c + 13], in[threadIdx.x + c + 19]);
    acc[c] = pack2(in[threadIdx.x + c + 23], in[threadIdx.x + c + 29]);
    fx[2 * c] = in[threadIdx.x + 2 * c]; fx[2 * c + 1] = in[threadIdx.x + 2 * c + 1];
    fy[2 * c] = in[threadIdx.x + 2 * c + 40]; fy[2 * c + 1] = in[threadIdx.x + 2 * c + 41];
    fa[2 * c] = in[threadIdx.x + 2 * c + 80]; fa[2 * c + 1] = in[threadIdx.x + 2 * c + 81];
  }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 4; u++) {
#pragma unroll
      for (int c = 0; c < C; c++) {
        if (MODE == 0) acc[c] = ffma2(x[c], y[c], acc[c]);
        if (MODE == 1) acc[c] = ffma2(x[c], x[c], acc[c]);
        if (MODE == 2) acc[c] = fmul2(x[c], acc[c]);
        if (MODE == 3) { fa[2 * c] = ffma(fx[2 * c], fy[2 * c], fa[2 * c]); fa[2 * c + 1] = ffma(fx[2 * c + 1], fy[2 * c + 1], fa[2 * c + 1]); }
        if (MODE == 4) acc[c] = fadd2(pack2(fx[c], fx[c]), acc[c]);
        if (MODE == 5) acc[c] = ffma2(x[0], y[0], acc[c]);
        if (MODE == 6) acc[c] = ffma2(x[c], y[0], acc[c]);
      }
    }
  }
  float r = 0;
#pragma unroll
  for (int c = 0; c < C; c++) { float lo, hi; unpack2(acc[c], lo, hi); r += lo + hi + fa[2 * c] + fa[2 * c + 1]; unpack2(x[c], lo, hi); r += lo + hi; unpack2(y[c], lo, hi); r += lo + hi; }
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
  const int iters = 1 << 15;
  k<5><<<148 * 16, 128>>>(d_out, d_in, 1 << 18); cudaDeviceSynchronize();
  const char *names[] = {"FFMA2 3 distinct pairs", "FFMA2 x*x+acc", "FMUL2 x*acc", "FFMA 3 distinct (x2)", "FADD2 Rbcast+acc", "FFMA2 shared x,y", "FFMA2 x[c]*y0+acc"};
  for (int wps = 2; wps <= 8; wps *= 2) {
    int grid = 148 * wps, block = 128;
    double cyc[7];
    cyc[0] = run([&] { k<0><<<grid, block>>>(d_out, d_in, iters); });
    cyc[1] = run([&] { k<1><<<grid, block>>>(d_out, d_in, iters); });
    cyc[2] = run([&] { k<2><<<grid, block>>>(d_out, d_in, iters); });
    cyc[3] = run([&] { k<3><<<grid, block>>>(d_out, d_in, iters); });
    cyc[4] = run([&] { k<4><<<grid, block>>>(d_out, d_in, iters); });
    cyc[5] = run([&] { k<5><<<grid, block>>>(d_out, d_in, iters); });
    cyc[6] = run([&] { k<6><<<grid, block>>>(d_out, d_in, iters); });
    for (int m = 0; m < 7; m++) {
      double ninst = (m == 3 ? 2.0 : 1.0) * C * 4.0 * iters * wps;
      printf("warps/SMSP=%d %-24s cycles per warp-inst per SMSP = %.3f\n", wps, names[m], cyc[m] / ninst);
    }
  }
  return 0;
}
