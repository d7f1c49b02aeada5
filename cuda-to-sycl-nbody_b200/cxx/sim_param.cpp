// sim_param.cpp -- defaults and argv parsing of SimParam, behaviour-compatible with the
// reference (src/sim_param.cpp:12-67), written from scratch for nbody-b200.
#include "sim_param.hpp"

#include <cstdint>
#include <stdexcept>
#include <string>

SimParam::SimParam()
    : G(2.0f),
      dt(0.005f),
      numParticles(50 * 256),
      numFrames(SIZE_MAX),
      simIterationsPerFrame(4),
      damping(0.999998f),
      distEps(1.0e-7f),
      gwSize(64),
      calcMethod(CalculationMethod::BRANCH) {}

namespace {
CalculationMethod method_from_name(const std::string &name) {
  if (name == "BRANCH") return CalculationMethod::BRANCH;
  if (name == "PREDICATED") return CalculationMethod::PREDICATED;
  // same exception type and message as the reference (src/sim_param.cpp:36)
  throw std::invalid_argument("Valid calculation methods are BRANCH or PREDICATED");
}
}  // namespace

void SimParam::parseArgs(int argc, char **argv) {
  // each argument is optional; the unit of argv[1] is 256 bodies
  switch (argc > 10 ? 10 : argc) {
    case 10: calcMethod = method_from_name(argv[9]);  // fallthrough
    case 9: gwSize = atoi(argv[8]);                   // fallthrough
    case 8: numFrames = atoi(argv[7]);                // fallthrough
    case 7: G = atof(argv[6]);                        // fallthrough
    case 6: distEps = atof(argv[5]);                  // fallthrough
    case 5: dt = atof(argv[4]);                       // fallthrough
    case 4: damping = atof(argv[3]);                  // fallthrough
    case 3: simIterationsPerFrame = atoi(argv[2]);    // fallthrough
    case 2: numParticles = 256 * atoi(argv[1]);       // fallthrough
    default: break;
  }
}
