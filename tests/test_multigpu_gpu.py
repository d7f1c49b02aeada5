"""Multi-GPU sharding (new functionality; the reference is single-GPU, src/simulator.cu:40).
Oracle = the single-GPU result at the same N, which test_parity_gpu.py pins to the reference:
sharding by i-range with the j-loop cut into rank-ordered chunks must not change a single bit."""
from __future__ import annotations

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _state(sim):
    p, v = sim.getParticlePos(), sim.getParticleVel()
    return [p.x.copy(), p.y.copy(), p.z.copy(), v.x.copy(), v.y.copy(), v.z.copy()]


def _need(nb, k):
    if nb.device_count() < k:
        pytest.skip(f"needs {k} GPUs")


@pytest.mark.parametrize("n,gpus", [(25600, 2), (1000, 2), (262144, 2), (4099, 3), (65536, 4), (262144, 8)])
def test_sharded_step_bit_equal_to_single_gpu(nb, n, gpus):
    _need(nb, gpus)
    one = nb.DiskGalaxySimulator(nb.SimParam(numParticles=n, simIterationsPerFrame=5), n_gpus=1)
    many = nb.DiskGalaxySimulator(nb.SimParam(numParticles=n, simIterationsPerFrame=5), n_gpus=gpus)
    for frame in range(2):
        one.stepSim()
        many.stepSim()
        for k, (a, b) in enumerate(zip(_state(many), _state(one))):
            assert np.array_equal(a, b), f"frame {frame} component {k}: {(a != b).sum()} bodies differ"
    fa, fb = many.computeAccel(), one.computeAccel()
    for a, b in zip(fa, fb):
        assert np.array_equal(a, b)
    p4 = many.readPosF4()
    assert np.array_equal(p4[:, 0], _state(one)[0])
    one.close()
    many.close()


def test_sharded_generic_kernel_and_set_state(nb):
    _need(nb, 2)
    n = 5000
    rng = np.random.default_rng(3)
    st = [rng.uniform(-30, 30, n).astype(np.float32) for _ in range(3)] + \
         [rng.uniform(-1, 1, n).astype(np.float32) for _ in range(3)]
    out = []
    for gpus in (1, 2):
        sim = nb.DiskGalaxySimulator(nb.SimParam(numParticles=n, simIterationsPerFrame=3, distEps=0.0), n_gpus=gpus)
        assert "generic" in sim.kernelName()
        sim.setState(*st)
        sim.stepSim()
        out.append(_state(sim))
        sim.close()
    for a, b in zip(*out):
        assert np.array_equal(a, b)
