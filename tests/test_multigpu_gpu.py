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


@pytest.mark.parametrize("gpus", [2, 4, 8])
@pytest.mark.parametrize("exchange", ["p2p", "nccl"])
def test_sharded_forces_and_step_vs_reference_golden(nb, golden_dir, gpus, exchange, monkeypatch):
    """pinned to the REFERENCE, not to this repo's single-GPU kernel: forces and one default step of the
    sharded handle at N = 262144 (BASELINE configs[1]) against the SHA-256 of what the unmodified
    reference kernel produced on a B200 (tests/golden/make_golden.py --big-only)"""
    import hashlib
    import json
    import os
    _need(nb, gpus)
    monkeypatch.setenv("NBODY_EXCHANGE", exchange)
    meta = json.load(open(os.path.join(golden_dir, "golden_meta.json")))
    n = 262144

    def sha(arrays):
        return hashlib.sha256(np.stack(arrays, axis=1).reshape(-1).tobytes()).hexdigest()

    many = nb.DiskGalaxySimulator(nb.SimParam(numParticles=n, simIterationsPerFrame=1), n_gpus=gpus)
    assert sha(many.computeAccel()) == meta["force_sha256"][str(n)]
    many.stepSim()
    assert sha(_state(many)) == meta["step1_sha256"][str(n)]
    b, c = many.localRange()
    assert (b, c) == (0, n)
    loc = [np.empty(n, np.float32) for _ in range(6)]
    many.readLocalInto(*loc)
    assert sha(loc) == meta["step1_sha256"][str(n)]
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
full = [a.copy() for a in (p.x, p.y, p.z, v.x, v.y, v.z)]
# a process-per-GPU application loop: upload state, step, read back only the bodies this rank owns.
# No rank waits for the others between the calls (nbody_step itself is the only collective).
b, c = sim.localRange()
assert (b, c) == nb.plan_shard(n, d.world, d.rank)
loc = [np.empty(c, np.float32) for _ in range(6)]
cur = [a.copy() for a in full]
for frame in range(3):
    sim.setState(*cur)
    sim.stepSim()
    sim.readLocalInto(*loc)
    # the full state for the next upload comes from the collective read
    sim._host_fresh = False
    p, v = sim.getParticlePos(), sim.getParticleVel()
    cur = [a.copy() for a in (p.x, p.y, p.z, v.x, v.y, v.z)]
    for k in range(6):
        assert np.array_equal(loc[k], cur[k][b:b + c]), (frame, k)
if d.rank == 0:
    np.savez({out!r}, x=full[0], y=full[1], z=full[2], vx=full[3], vy=full[4], vz=full[5], name=sim.kernelName(),
             x2=cur[0], vz2=cur[5])
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
    for _ in range(3):  # the upload / step / local read-back loop of the workers, on one GPU
        one.setState(*_state(one))
        one.stepSim()
    st = _state(one)
    assert np.array_equal(st[0], g["x2"]) and np.array_equal(st[5], g["vz2"])
    one.close()
