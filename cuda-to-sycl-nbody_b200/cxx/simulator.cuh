// simulator.cuh -- drop-in mirror of the reference's simulator header
// (reference src/simulator.cuh:107-160), written from scratch for nbody-b200.
//
// Same namespace, class, method and struct names and signatures as the reference, so
// src/nbody.cpp and src/renderer_gl.cpp compile against it unchanged.  Unlike the reference
// header it needs no CUDA headers and holds no device pointers: DiskGalaxySimulator is a thin
// wrapper over the opaque C handle of include/nbody_b200.h (libnbody_b200.so).
#pragma once

#include <cstddef>
#include <string>
#include <vector>

#include "sim_param.hpp"

struct nbody_handle;  // include/nbody_b200.h

namespace simulation {

const float PI = 3.14159265358979323846;  // src/simulator.cuh:35

typedef float coords_t;

// host-side SoA the renderer reads (src/simulator.cuh:75-84, src/renderer_gl.cpp:156-172)
struct ParticleData {
  std::vector<coords_t> x;
  std::vector<coords_t> y;
  std::vector<coords_t> z;

  ParticleData(std::vector<coords_t> x_, std::vector<coords_t> y_, std::vector<coords_t> z_)
      : x(std::move(x_)), y(std::move(y_)), z(std::move(z_)) {}
  explicit ParticleData(size_t n) : x(n, 0.0f), y(n, 0.0f), z(n, 0.0f) {}
};

// interface the renderer holds a pointer to (src/simulator.cuh:107-116)
class Simulator {
 public:
  virtual ~Simulator() = default;
  virtual void stepSim() = 0;
  virtual size_t getNumParticles() = 0;
  virtual const ParticleData &getParticlePos() = 0;
  virtual const ParticleData &getParticleVel() = 0;
  virtual float getLastStepTime() = 0;
  virtual const std::string *getDeviceName() = 0;
  virtual int getGwSize() = 0;
};

// src/simulator.cuh:129-160.  Post-conditions are the reference's: after stepSim() returns the
// state is advanced by simIterationsPerFrame iterations, getLastStepTime() is the host-clock
// time in ms from before the first launch to after the device synchronise (read-back
// excluded), and getParticlePos()/getParticleVel() return host copies valid until the next
// stepSim().  The read-back is lazy: the device->host copy happens on the first getter call
// after a step (the headless main loop never asks, src/nbody.cpp:95-123).
// Extra knobs come from the environment so that the constructor signature stays the
// reference's: NBODY_GPUS=<n> shards the bodies over n GPUs of this process; NBODY_PREDICATED_FIXED=1
// gives calcMethod PREDICATED the README's (i != id) meaning; NBODY_DUMP_HASH=1 prints the FNV-1a-64
// hash of the final host state on stderr when the simulator is destroyed.
class DiskGalaxySimulator : public Simulator {
 public:
  explicit DiskGalaxySimulator(SimParam params_);
  ~DiskGalaxySimulator() override;
  DiskGalaxySimulator(const DiskGalaxySimulator &) = delete;
  DiskGalaxySimulator &operator=(const DiskGalaxySimulator &) = delete;

  void stepSim() override;
  float getLastStepTime() override { return lastStepTime; }
  size_t getNumParticles() override { return params.numParticles; }
  const ParticleData &getParticlePos() override;
  const ParticleData &getParticleVel() override;
  const std::string *getDeviceName() override;
  int getGwSize() override { return params.gwSize; }
  CalculationMethod getCM() { return params.calcMethod; }

  // additions (not in the reference): device-event time of the last step, C handle access
  float getLastStepDeviceTime() const { return lastStepDeviceTime; }
  nbody_handle *handle() { return impl; }

 private:
  void check(int rc, const char *what, int line);
  void refreshHost();

  SimParam params;
  std::string devName;
  float lastStepTime{0.0f};
  float lastStepDeviceTime{0.0f};
  ParticleData pos;
  ParticleData vel;
  bool hostFresh{false};
  nbody_handle *impl{nullptr};
  std::vector<void *> registered;  // host vectors page-locked through nbody_host_register
};

}  // namespace simulation
