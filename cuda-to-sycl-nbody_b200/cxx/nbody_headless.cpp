// nbody_headless.cpp -- headless driver with the reference's argv and stdout contract
// (reference src/nbody.cpp:31-37,83-123 with DISABLE_GL), written from scratch.
// Same positional arguments (src/sim_param.cpp:40-67), same per-step line
//   "At step S kernel time is T and mean is M and stddev is: D"
// after two warm-up frames; mean/stddev are kept as running sums instead of being recomputed
// over the whole history every frame (the reference is O(steps^2)).
#include <cmath>
#include <iostream>

#include "sim_param.hpp"
#include "simulator.cuh"

int main(int argc, char **argv) {
  SimParam params;
  params.parseArgs(argc, argv);
  simulation::DiskGalaxySimulator sim(params);

  const int warm_steps = 2;
  double sum = 0.0, sum_sq = 0.0;
  size_t samples = 0;
  for (size_t step = 1; step <= params.numFrames; ++step) {
    sim.stepSim();
    if (step <= (size_t)warm_steps) continue;
    const float t = sim.getLastStepTime();
    sum += t;
    sum_sq += (double)t * t;
    ++samples;
    const float mean = (float)(sum / samples);
    const double var = sum_sq / samples - (sum / samples) * (sum / samples);
    const float stddev = (float)std::sqrt(var > 0.0 ? var : 0.0);
    std::cout << "At step " << step << " kernel time is " << t << " and mean is " << mean
              << " and stddev is: " << stddev << "\n";
  }
  return 0;
}
