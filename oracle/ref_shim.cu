/*
 * ref_shim.cu -- C ABI around the UNMODIFIED reference simulator.  TEST INFRASTRUCTURE ONLY.
 *
 * This translation unit #includes the reference's own src/simulator.cu (and through it
 * src/simulator.cuh, src/sim_param.hpp) from where they lie under /root/reference; no
 * reference source is copied into this repository.  oracle/Makefile compiles it together
 * with the reference's src/sim_param.cpp, with the reference's own flags
 * (-O3 -use_fast_math, src/CMakeLists.txt:51) plus -gencode arch=compute_100a,code=sm_100a,
 * into oracle/_ref/libnbody_ref.so (git-ignored, shipped to the GPU box by gpurun).
 *
 * It is the bit-exact oracle of the CUDA path: tests/ drive it through ctypes on the GPU box,
 * tests/golden/make_golden.py uses it to write the golden fixtures, bench.py times its kernel
 * as "the kernel to beat".  The product never loads it.
 *
 * `#define private public` only widens access so the shim can upload caller-provided state
 * and launch the reference kernel directly; it changes no layout and no code.
 */
#include <cuda.h>
#include <cuda_runtime_api.h>
#include <stdio.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <iostream>
#include <map>
#include <random>
#include <string>
#include <tuple>
#include <vector>

#define private public
#include "simulator.cuh"
#undef private
#include "simulator.cu" /* resolved by -I/root/reference/src */

using simulation::DiskGalaxySimulator;

extern "C" {

int ref_device_count() {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

/* constructs the reference simulator: runs ITS generator and ITS sendToDevice */
void *ref_create(float G, float dt, unsigned long long n, int iters, float damping, float eps,
                 int gw, int calc) {
  SimParam p;
  p.G = G;
  p.dt = dt;
  p.numParticles = n;
  p.simIterationsPerFrame = iters;
  p.damping = damping;
  p.distEps = eps;
  p.gwSize = gw;
  p.calcMethod = calc == 0 ? CalculationMethod::BRANCH : CalculationMethod::PREDICATED;
  return new DiskGalaxySimulator(p);
}

void ref_destroy(void *h) {
  auto *s = static_cast<DiskGalaxySimulator *>(h);
  /* the reference has no destructor; release its device buffers here */
  for (auto *d : {&s->pos_d, &s->pos_next_d, &s->vel_d}) {
    cudaFree(d->x);
    cudaFree(d->y);
    cudaFree(d->z);
  }
  delete s;
}

void ref_set_state(void *h, const float *x, const float *y, const float *z, const float *vx,
                   const float *vy, const float *vz) {
  auto *s = static_cast<DiskGalaxySimulator *>(h);
  size_t n = s->getNumParticles();
  std::copy(x, x + n, s->pos.x.begin());
  std::copy(y, y + n, s->pos.y.begin());
  std::copy(z, z + n, s->pos.z.begin());
  std::copy(vx, vx + n, s->vel.x.begin());
  std::copy(vy, vy + n, s->vel.y.begin());
  std::copy(vz, vz + n, s->vel.z.begin());
  s->sendToDevice();
}

/* host copy of the state the reference exposes through getParticlePos/getParticleVel */
void ref_get_state(void *h, float *x, float *y, float *z, float *vx, float *vy, float *vz) {
  auto *s = static_cast<DiskGalaxySimulator *>(h);
  const simulation::ParticleData &p = s->getParticlePos();
  const simulation::ParticleData &v = s->getParticleVel();
  std::copy(p.x.begin(), p.x.end(), x);
  std::copy(p.y.begin(), p.y.end(), y);
  std::copy(p.z.begin(), p.z.end(), z);
  std::copy(v.x.begin(), v.x.end(), vx);
  std::copy(v.y.begin(), v.y.end(), vy);
  std::copy(v.z.begin(), v.z.end(), vz);
}

void ref_step(void *h) { static_cast<DiskGalaxySimulator *>(h)->stepSim(); }
float ref_last_step_ms(void *h) { return static_cast<DiskGalaxySimulator *>(h)->getLastStepTime(); }
const char *ref_device_name(void *h) {
  return static_cast<DiskGalaxySimulator *>(h)->getDeviceName()->c_str();
}

/* CUDA-event time (ms) of `launches` back-to-back launches of the reference BRANCH kernel with
 * work-group size gw, state left advanced (the launch loop of src/simulator.cu:57-66) */
float ref_time_kernel(void *h, int gw, int launches) {
  auto *s = static_cast<DiskGalaxySimulator *>(h);
  int nblocks = ((s->getNumParticles() - 1) / gw) + 1;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0, 0);
  for (int i = 0; i < launches; i++) {
    simulation::particle_interaction<CalculationMethod::BRANCH>
        <<<nblocks, gw>>>(s->pos_d, s->pos_next_d, s->vel_d, s->params);
    std::swap(s->pos_d, s->pos_next_d);
  }
  cudaEventRecord(e1, 0);
  cudaEventSynchronize(e1);
  float ms = -1.0f;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (cudaGetLastError() != cudaSuccess) return -1.0f;
  return ms;
}

} /* extern "C" */
