// galaxy.cpp -- host-side initial-condition generator of the drop-in simulator.
//
// Replaces DiskGalaxySimulator::randomParticlePos / initialParticleVel
// (reference src/simulator.cu:131-158 with helpers cross/length/normalize :165-181) and must
// reproduce them bit-for-bit, because parity of every later step is judged on "identical
// generated galaxies".  What makes the bits:
//   * one default-seeded std::mt19937 feeding std::uniform_real_distribution<double>(0,1);
//     per body an angle draw then a radius draw, and only afterwards all the z draws;
//   * angle = u*2*PI with PI a *float* constant, narrowed to float before the trig call;
//     radius = u*100 narrowed to float;
//   * single-precision cosf/sinf (the reference file is compiled by nvcc, whose headers resolve
//     cos(float) to cosf; plain <cmath> would promote to double and differ in the last ulp);
//   * tangential velocity through DOUBLE pow/sqrt narrowed to float: |c| = (float)sqrt(cx^2+cy^2+cz^2),
//     speed = (float)sqrt(2.0*|c|), v = (c/|c|)*speed with float division and multiply.
// Compile with -ffp-contract=off so that no FMA is formed on hosts that have one.
#include <cmath>
#include <cstdint>
#include <random>

#include "../../include/nbody_b200.h"

namespace {
constexpr float kPi = 3.14159265358979323846;  // float, as simulation::PI (src/simulator.cuh:35)
}

extern "C" int nbody_generate_disk_galaxy(uint64_t n, float *x, float *y, float *z, float *vx,
                                          float *vy, float *vz) {
  if (n && (!x || !y || !z || !vx || !vy || !vz)) return NBODY_E_INVALID;
  std::mt19937 engine;  // default seed 5489
  std::uniform_real_distribution<double> unit(0.0, 1.0);

  for (uint64_t i = 0; i < n; ++i) {
    const float angle = unit(engine) * 2 * kPi;
    const float radius = unit(engine) * 100;
    x[i] = cosf(angle) * radius;
    y[i] = sinf(angle) * radius;
  }
  for (uint64_t i = 0; i < n; ++i) z[i] = 4.0 * unit(engine);

  for (uint64_t i = 0; i < n; ++i) {
    // c = p x (0,0,1), written out so that signed zeros come out as in the reference
    const float ex = 0.0f, ey = 0.0f, ez = 1.0f;
    const float cx = y[i] * ez - z[i] * ey;
    const float cy = z[i] * ex - x[i] * ez;
    const float cz = x[i] * ey - y[i] * ex;
    const float norm = std::sqrt(std::pow((double)cx, 2) + std::pow((double)cy, 2) + std::pow((double)cz, 2));
    const float speed = std::sqrt(2.0 * norm);
    vx[i] = cx / norm * speed;
    vy[i] = cy / norm * speed;
    vz[i] = cz / norm * speed;
  }
  return 0;
}
