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


@pytest.mark.parametrize("exchange", ["p2p", "nccl"])
@pytest.mark.parametrize("n,gpus", [(25600, 2), (1000, 2), (262144, 2), (4099, 3), (65536, 4), (262144, 8)])
def test_sharded_step_bit_equal_to_single_gpu(nb, n, gpus, exchange, monkeypatch):
    """both exchange modes: peer push from the kernel epilogue (NVLink stores) and rank-ordered
    NCCL broadcasts overlapped with j-chunk kernels"""
    _need(nb, gpus)
    monkeypatch.setenv("NBODY_EXCHANGE", exchange)
    one = nb.DiskGalaxySimulator(nb.SimParam(numParticles=n, simIterationsPerFrame=5), n_gpus=1)
    many = nb.DiskGalaxySimulator(nb.SimParam(numParticles=n, simIterationsPerFrame=5), n_gpus=gpus)
    assert ("peer-push" if exchange == "p2p" else "nccl-bcast") in many.kernelName()
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


RANK_WORKER = r"""
import os, sys, json
import numpy as np
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "cuda-to-sycl-nbody_b200"))
import bench, nbody_b200 as nb
d = bench.Dist()
uid = d.broadcast_bytes(nb.nccl_unique_id() if d.rank == 0 else None, 128)
n = {n}
sim = nb.DiskGalaxySimulator(nb.SimParam(numParticles=n, simIterationsPerFrame=4), rank=d.rank, world=d.world,
                             device=d.local_rank, unique_id=uid)
sim.stepSim(); sim.stepSim()
p, v = sim.getParticlePos(), sim.getParticleVel()
if d.rank == 0:
    np.savez({out!r}, x=p.x, y=p.y, z=p.z, vx=v.x, vy=v.y, vz=v.z, name=sim.kernelName())
sim.close(); d.close()
"""


@pytest.mark.parametrize("exchange", ["p2p", "nccl"])
def test_one_process_per_gpu_matches_single_gpu(nb, tmp_path, exchange):
    """torchrun-style launch (one rank per GPU, NCCL id broadcast by torch.distributed; CUDA IPC
    peer mappings in p2p mode) against the single-GPU result"""
    import os
    import socket
    import subprocess
    import sys
    _need(nb, 2)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    n = 40000
    out = str(tmp_path / "rank0.npz")
    script = tmp_path / "w.py"
    script.write_text(RANK_WORKER.format(root=root, n=n, out=out))
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), NBODY_EXCHANGE=exchange)
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.PIPE, text=True))
    for p in procs:
        o, e = p.communicate(timeout=300)
        assert p.returncode == 0, e[-3000:]
    g = np.load(out)
    assert ("peer-push" if exchange == "p2p" else "nccl-bcast") in str(g["name"])
    one = nb.DiskGalaxySimulator(nb.SimParam(numParticles=n, simIterationsPerFrame=4), n_gpus=1)
    one.stepSim(); one.stepSim()
    for a, k in zip(_state(one), ("x", "y", "z", "vx", "vy", "vz")):
        assert np.array_equal(a, g[k]), k
    one.close()
