// ubench_const.cu -- can the j-bodies be fed through the constant bank / uniform datapath?
#include <cuda_runtime.h>
#include <stdio.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pack2(float lo, float hi) { u64 d; asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi)); return d; }
__device__ __forceinline__ void unpack2(u64 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 fadd2(u64 a, u64 b) { u64 d; asm("add.rn.ftz.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 fmul2(u64 a, u64 b) { u64 d; asm("mul.rn.ftz.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.ftz.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ float frsq(float a) { float d; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(d) : "f"(a)); return d; }

__constant__ float4 c_pos[4096];

template <int NP>
__global__ void k_const(float *out, const float *in, int sweeps, float eps) {
  u64 nx[NP], ny[NP], nz[NP], ax[NP], ay[NP], az[NP];
#pragma unroll
  for (int p = 0; p < NP; p++) {
    nx[p] = pack2(in[threadIdx.x + p], in[threadIdx.x + p + 3]); ny[p] = pack2(in[threadIdx.x + p + 5], in[threadIdx.x + p + 9]);
    nz[p] = pack2(in[threadIdx.x + p + 11], in[threadIdx.x + p + 17]);
    ax[p] = ay[p] = az[p] = 0ull;
  }
  const u64 eps2 = pack2(eps, eps);
  for (int s = 0; s < sweeps; s++) {
#pragma unroll 8
    for (int j = 0; j < 4096; j++) {
      const float4 q = c_pos[j];
      const u64 X = pack2(q.x, q.x), Y = pack2(q.y, q.y), Z = pack2(q.z, q.z);
#pragma unroll
      for (int p = 0; p < NP; p++) {
        u64 rx = fadd2(X, nx[p]), ry = fadd2(Y, ny[p]), rz = fadd2(Z, nz[p]);
        u64 t = fmul2(ry, ry); t = ffma2(rx, rx, t); t = ffma2(rz, rz, t);
        u64 d = fadd2(t, eps2);
        u64 c = fmul2(d, d); c = fmul2(d, c);
        float c0, c1; unpack2(c, c0, c1);
        u64 w = pack2(frsq(c0), frsq(c1));
        ax[p] = ffma2(rx, w, ax[p]); ay[p] = ffma2(ry, w, ay[p]); az[p] = ffma2(rz, w, az[p]);
      }
    }
  }
  float r = 0;
#pragma unroll
  for (int p = 0; p < NP; p++) { float lo, hi; unpack2(ax[p], lo, hi); r += lo + hi; unpack2(ay[p], lo, hi); r += lo + hi; unpack2(az[p], lo, hi); r += lo + hi; }
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
  static float4 h[4096];
  for (int i = 0; i < 4096; i++) h[i] = make_float4(i * 0.37f, i * 0.11f, i * 0.05f, 1.f);
  cudaMemcpyToSymbol(c_pos, h, sizeof h);
  k_const<2><<<148 * 8, 128>>>(d_out, d_in, 64, 1e-7f); cudaDeviceSynchronize();
  const int sweeps = 16;
  for (int wps = 1; wps <= 8; wps++) {
    int grid = 148 * wps, block = 128;
    double c;
#define T(NP, label) c = run([&] { k_const<NP><<<grid, block>>>(d_out, d_in, sweeps, 1e-7f); }); \
    printf("warps/SMSP=%d %-12s cycles/j=%7.2f inter/clk/SM=%6.3f (%%of 10.667: %5.1f)\n", wps, label, c / (sweeps * 4096.0), \
           4.0 * wps * 32 * 2 * NP * sweeps * 4096.0 / c, 100.0 * (4.0 * wps * 32 * 2 * NP * sweeps * 4096.0 / c) / 10.6667);
    T(1, "const R=2") T(2, "const R=4") T(3, "const R=6") T(4, "const R=8")
  }
  return 0;
}
