// simulator.cpp -- DiskGalaxySimulator as a thin C++ wrapper over the C ABI
// (include/nbody_b200.h).  Mirrors the behaviour of reference src/simulator.cu:24-75,160-162;
// all GPU work lives behind the ABI in libnbody_b200.so.
#include "simulator.cuh"

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

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
  // opt-in: PREDICATED as the README describes it, (i != id), instead of the shipped (i == id)
  // (src/simulator.cu:209 vs README.md:229-231); the constructor signature stays the reference's
  const char *fixed = getenv("NBODY_PREDICATED_FIXED");
  if (p.calc_method == NBODY_CALC_PREDICATED && fixed && atoi(fixed) != 0) p.calc_method = NBODY_CALC_PREDICATED_FIXED;
  // generates the reference's default-seeded disk galaxy and uploads it (ctor, src/simulator.cu:31-33)
  NB_CHECK(nbody_create(&p, /*n_gpus: NBODY_GPUS or 1*/ 0, &impl));
  // the host vectors are sized here and never reallocate (src/simulator.cuh:83): page-lock them once so
  // every read-back is a DMA transfer straight into them (best effort: pageable memory still works)
  for (std::vector<coords_t> *v : {&pos.x, &pos.y, &pos.z, &vel.x, &vel.y, &vel.z})
    if (!v->empty() && nbody_host_register(v->data(), v->size() * sizeof(coords_t)) == 0) registered.push_back(v->data());
  // the reference's host vectors hold the initial state right after construction
  // (renderer.updateParticles() is called before the first step, src/nbody.cpp:78)
  refreshHost();
}

DiskGalaxySimulator::~DiskGalaxySimulator() {
  // NBODY_DUMP_HASH=1: FNV-1a-64 of the final host state (x,y,z,vx,vy,vz bit patterns per body) on
  // stderr, so a run of ANY main -- including the reference's own src/nbody.cpp built against this
  // header -- can be compared bit-for-bit with the reference simulator
  const char *dump = getenv("NBODY_DUMP_HASH");
  if (impl && dump && atoi(dump) != 0) {
    refreshHost();
    uint64_t h = 1469598103934665603ull;
    const std::vector<coords_t> *arr[6] = {&pos.x, &pos.y, &pos.z, &vel.x, &vel.y, &vel.z};
    for (size_t i = 0; i < params.numParticles; i++)
      for (const auto *a : arr) {
        uint32_t bits;
        memcpy(&bits, &(*a)[i], sizeof bits);
        for (int b = 0; b < 4; b++) h = (h ^ ((bits >> (8 * b)) & 0xffu)) * 1099511628211ull;
      }
    fprintf(stderr, "nbody-b200 final state fnv1a64 %016llx\n", (unsigned long long)h);
  }
  for (void *ptr : registered) nbody_host_unregister(ptr);
  nbody_destroy(impl);
}

void DiskGalaxySimulator::stepSim() {
  NB_CHECK(nbody_step(impl));
  lastStepTime = nbody_last_step_ms(impl);
  lastStepDeviceTime = nbody_last_step_device_ms(impl);
  hostFresh = false;
}

void DiskGalaxySimulator::refreshHost() {
  if (hostFresh) return;
  NB_CHECK(nbody_read_state(impl, pos.x.data(), pos.y.data(), pos.z.data(), vel.x.data(), vel.y.data(), vel.z.data()));
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
