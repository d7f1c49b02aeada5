"""The C++ drop-in header must serve the reference's callers unchanged.  This compiles (CPU only, no
GPU needed) a translation unit that uses the simulator exactly the way the reference's renderer and
main loop do (src/renderer_gl.cpp:39,111-114,144,156-172; src/nbody.cpp:33-36,92-110) against
cuda-to-sycl-nbody_b200/cxx/ and links it with the library."""
from __future__ import annotations

import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "cuda-to-sycl-nbody_b200")

CALLER = r'''
#include <cstdio>
#include <string>
#include <vector>
#include "sim_param.hpp"
#include "simulator.cuh"      // plain C++ TU, as renderer.hpp:12 includes it
using namespace simulation;

// what RendererGL::setParticleData does with the host SoA (src/renderer_gl.cpp:156-172)
static void fill_vec4(std::vector<float> &dst, const ParticleData &data, size_t n) {
  dst.resize(4 * n);
  for (size_t i = 0; i < n; i++) { dst[4*i] = data.x[i]; dst[4*i+1] = data.y[i]; dst[4*i+2] = data.z[i]; dst[4*i+3] = 1.0f; }
}

int main(int argc, char **argv) {
  SimParam params;                       // src/nbody.cpp:33-34
  params.parseArgs(argc, argv);
  static_assert(sizeof(params.numParticles) == sizeof(size_t), "field types are the reference's");
  if (argc > 10) return 3;
  if (params.numFrames == 0) {           // argv[7] = 0: only exercise the parser, never touch the GPU
    std::printf("%zu %d %g %g %g %g %d %d\n", params.numParticles, params.simIterationsPerFrame, params.damping,
                params.dt, params.distEps, params.G, params.gwSize, (int)params.calcMethod);
    return 0;
  }
  DiskGalaxySimulator nbodySim(params);  // src/nbody.cpp:36
  Simulator *sim = &nbodySim;            // the renderer holds a Simulator*  (src/renderer_gl.hpp)
  std::vector<float> vbo, ssbo;
  for (size_t step = 0; step < params.numFrames; step++) {
    sim->stepSim();
    fill_vec4(vbo, sim->getParticlePos(), sim->getNumParticles());
    fill_vec4(ssbo, sim->getParticleVel(), sim->getNumParticles());
    float t = sim->getLastStepTime();
    const std::string *name = sim->getDeviceName();
    std::printf("%s %f %d %d\n", name->c_str(), t, sim->getGwSize(), (int)nbodySim.getCM());
  }
  return 0;
}
'''


def _build(tmp_path):
    src = tmp_path / "caller.cpp"
    src.write_text(CALLER)
    exe = tmp_path / "caller"
    lib = os.path.join(PKG, "lib")
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-I", os.path.join(PKG, "cxx"), str(src),
           os.path.join(PKG, "cxx", "simulator.cpp"), os.path.join(PKG, "cxx", "sim_param.cpp"),
           "-L", lib, "-lnbody_b200", f"-Wl,-rpath,{lib}", "-o", str(exe)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return str(exe)


def test_reference_style_caller_compiles_and_parses_args(nb, tmp_path):
    exe = _build(tmp_path)
    # positional argv of src/sim_param.cpp:40-67; numFrames = 0 -> parser only
    r = subprocess.run([exe, "100", "10", "0.999", "0.001", "1.0e-3", "2.0", "0", "128", "PREDICATED"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert r.stdout.split() == ["25600", "10", "0.999", "0.001", "0.001", "2", "128", "1"]
    # defaults of src/sim_param.cpp:12-22 (numFrames given so that the GPU is not touched)
    r = subprocess.run([exe, "50", "4", "0.999998", "0.005", "1.0e-7", "2.0", "0"], capture_output=True, text=True)
    assert r.stdout.split() == ["12800", "4", "0.999998", "0.005", "1e-07", "2", "64", "0"]
    # bad calculation method: std::invalid_argument as in the reference (src/sim_param.cpp:36) -> abort
    r = subprocess.run([exe, "1", "1", "1", "1", "1", "1", "0", "64", "NOPE"], capture_output=True, text=True)
    assert r.returncode != 0 and "BRANCH or PREDICATED" in r.stderr


def test_reference_style_caller_fails_fast_without_gpu(nb, tmp_path):
    import pytest
    if nb.device_count() > 0:
        pytest.skip("GPU box: covered by tests/test_parity_gpu.py::test_cxx_dropin_binaries")
    exe = _build(tmp_path)
    r = subprocess.run([exe, "4", "1", "0.999", "0.001", "1e-3", "2.0", "2"], capture_output=True, text=True)
    assert r.returncode != 0 and r.stderr.startswith("GPUassert:")
