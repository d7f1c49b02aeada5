// sim_param.hpp -- drop-in mirror of the reference's SimParam (reference src/sim_param.hpp:16-40).
// The nine public fields keep the reference's names, types and ORDER (the struct is part of
// the DiskGalaxySimulator constructor signature and callers poke the fields directly,
// e.g. params.numFrames in src/nbody.cpp:92).  Written from scratch for nbody-b200.
#pragma once

#include <cstddef>
#include <cstdlib>

enum class CalculationMethod { BRANCH, PREDICATED };

class SimParam {
 public:
  SimParam();  // reference defaults, src/sim_param.cpp:12-22

  // positional argv: [1] numParticles/256 [2] simIterationsPerFrame [3] damping [4] dt
  // [5] distEps [6] G [7] numFrames [8] gwSize [9] BRANCH|PREDICATED  (src/sim_param.cpp:40-67)
  void parseArgs(int argc, char **argv);

  float G;
  float dt;
  size_t numParticles;
  size_t numFrames;
  int simIterationsPerFrame;
  float damping;
  float distEps;  // added to r^2 (src/simulator.cu:201)
  int gwSize;     // advisory here: the B200 kernels pick their own CTA shape
  CalculationMethod calcMethod;
};
