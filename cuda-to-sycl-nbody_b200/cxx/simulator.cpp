// simulator.cpp -- DiskGalaxySimulator as a thin C++ wrapper over the C ABI
// (include/nbody_b200.h).  Mirrors the behaviour of reference src/simulator.cu:24-75,160-162;
// all GPU work lives behind the ABI in libnbody_b200.so.
#include "simulator.cuh"

#include <cstdio>
#include <cstdlib>

#include "../../include/nbody_b200.h"

namespace simulation {

// same fail-fast convention as the reference's gpuErrchk (src/simulator.cuh:22-31):
// print "GPUassert: <msg> <file> <line>" to stderr and exit with the error code
void DiskGalaxySimulator::check(int rc, const char *what, int line) {
  if (rc == 0) return;
  fprintf(stderr, "GPUassert: %s (%s) %s %d\n", nbody_last_error(), what, __FILE__, line);
  exit(rc);
}
#define NB_CHECK(call) check((call), #call, __LINE__)

DiskGalaxySimulator::DiskGalaxySimulator(SimParam params_)
    : params(params_), pos(params_.numParticles), vel(params_.numParticles) {
  nbody_params p;
  p.G = params.G;
  p.dt = params.dt;
  p.num_particles = params.numParticles;
  p.iters_per_frame = params.simIterationsPerFrame;
  p.damping = params.damping;
  p.dist_eps = params.distEps;
  p.gw_size = params.gwSize;
  p.calc_method = params.calcMethod == CalculationMethod::BRANCH ? NBODY_CALC_BRANCH : NBODY_CALC_PREDICATED;
  // generates the reference's default-seeded disk galaxy and uploads it (ctor, src/simulator.cu:31-33)
  NB_CHECK(nbody_create(&p, /*n_gpus: NBODY_GPUS or 1*/ 0, &impl));
  // the reference's host vectors hold the initial state right after construction
  // (renderer.updateParticles() is called before the first step, src/nbody.cpp:78)
  refreshHost();
}

DiskGalaxySimulator::~DiskGalaxySimulator() { nbody_destroy(impl); }

void DiskGalaxySimulator::stepSim() {
  NB_CHECK(nbody_step(impl));
  lastStepTime = nbody_last_step_ms(impl);
  lastStepDeviceTime = nbody_last_step_device_ms(impl);
  hostFresh = false;
}

void DiskGalaxySimulator::refreshHost() {
  if (hostFresh) return;
  NB_CHECK(nbody_read_pos(impl, pos.x.data(), pos.y.data(), pos.z.data()));
  NB_CHECK(nbody_read_vel(impl, vel.x.data(), vel.y.data(), vel.z.data()));
  hostFresh = true;
}

const ParticleData &DiskGalaxySimulator::getParticlePos() {
  refreshHost();
  return pos;
}

const ParticleData &DiskGalaxySimulator::getParticleVel() {
  refreshHost();
  return vel;
}

const std::string *DiskGalaxySimulator::getDeviceName() {
  if (devName.empty()) {
    const char *n = nbody_device_name(impl);
    devName = (n && *n) ? n : "Unknown Device";
  }
  return &devName;
}

}  // namespace simulation
