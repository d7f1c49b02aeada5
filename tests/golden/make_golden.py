"""Generates the golden fixtures in this directory from the UNMODIFIED reference simulator.

The reference ships no tests or golden vectors, so parity is pinned by
running the reference's own code (oracle/_ref/libnbody_ref.so = /root/reference/src/simulator.cu
+ sim_param.cpp compiled for sm_100a with the reference's flags, see oracle/Makefile) on a B200:

    make -C oracle ref && gpurun -- python tests/golden/make_golden.py

and committing what it writes (gpurun_out/golden/* copied to tests/golden/).  Nothing from this
repository's product or CPU oracle takes part in producing these files.

Fixtures
  golden_meta.json          FNV-1a-64 hashes of generator output / forces / stepped states at
                            several N, the first bodies' bit patterns, GPU + toolchain provenance
  force_n2048.npz           raw force sums (damping=0, dt=1, G=1 trick) of the N=2048 galaxy
  step10_n2048.npz          pos+vel after 10 iterations with SimParam defaults, N=2048
  cloud_n1000.npz           a seeded uniform cloud (inputs included) and its reference forces, eps=1e-3
  predicated_n1024.npz      state after 1 step of the shipped PREDICATED kernel (zero force)
  predicated_fixed_n1024.npz  state after 1 step of the README-intended PREDICATED kernel, i.e. the reference
                            with src/simulator.cu:209 changed to (i != id) (oracle/Makefile: ref_fixed)
  golden_meta.json["force_sha256"], ["step1_sha256"], ["force_fixed_sha256"]
                            SHA-256 of the float32 byte stream (fx[i],fy[i],fz[i] interleaved per body; for
                            step1: x,y,z,vx,vy,vz) at the BASELINE sizes 262144 / 1M / 4M / 16M -- what bench.py,
                            smoke() and the full-size parity tests compare the CUDA path with

    python tests/golden/make_golden.py            # everything (the 16M force pass alone is ~4 min of B200)
    python tests/golden/make_golden.py --big-only # only the *_sha256 entries, merged into the committed meta
"""
from __future__ import annotations

import hashlib
import json
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import refsim  # noqa: E402

OUT = os.path.join(ROOT, "gpurun_out", "golden")


def fnv1a64(arrays) -> str:
    """FNV-1a-64 over the float32 bit patterns, interleaved per body (a0[i], a1[i], ...)."""
    inter = np.stack([np.ascontiguousarray(a, np.float32) for a in arrays], axis=1).reshape(-1)
    data = inter.view(np.uint8)
    h = 1469598103934665603
    # chunked pure-python would be slow at 262144*6*4 bytes; use the same recurrence vectorised per byte
    for b in data.tobytes():
        h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return f"{h:016x}"


def sha256_f32(arrays) -> str:
    """SHA-256 over the float32 bit patterns, interleaved per body (a0[i], a1[i], ...)."""
    inter = np.stack([np.ascontiguousarray(a, np.float32) for a in arrays], axis=1).reshape(-1)
    return hashlib.sha256(inter.tobytes()).hexdigest()


BIG_FORCE = (262144, 400003, 1048576, 4194304, 16777216)
BIG_STEP1 = (262144, 1048576)


def big(meta):
    """full-size pins: one launch of the reference kernel per size (1M: 0.9 s, 4M: 14 s, 16M: ~220 s on B200)"""
    meta.setdefault("force_sha256", {})
    meta.setdefault("step1_sha256", {})
    meta.setdefault("force_fixed_sha256", {})
    sizes = BIG_FORCE if "--no-16m" not in sys.argv else BIG_FORCE[:-1]
    for n in sizes:
        fx, fy, fz, _ = refsim.reference_forces(n)
        meta["force_sha256"][str(n)] = sha256_f32([fx, fy, fz])
        print("force_sha256", n, meta["force_sha256"][str(n)], flush=True)
        with open(os.path.join(OUT, "golden_meta.json"), "w") as f:  # keep what is done if the lease ends early
            json.dump(meta, f, indent=1)
    for n in BIG_STEP1:
        sim = refsim.RefSimulator(n, iters=1)
        sim.step()
        meta["step1_sha256"][str(n)] = sha256_f32(sim.state())
        sim.close()
        print("step1_sha256", n, meta["step1_sha256"][str(n)], flush=True)
    if refsim.available("fixed"):
        for n in (25600, 262144):
            fx, fy, fz, _ = refsim.reference_forces(n, calc=1, lib="fixed")
            meta["force_fixed_sha256"][str(n)] = sha256_f32([fx, fy, fz])
            print("force_fixed_sha256", n, meta["force_fixed_sha256"][str(n)], flush=True)
        n = 1024
        sim = refsim.RefSimulator(n, iters=1, calc=1, lib="fixed")
        sim.step()
        s = sim.state()
        sim.close()
        np.savez(os.path.join(OUT, "predicated_fixed_n1024.npz"), x=s[0], y=s[1], z=s[2], vx=s[3], vy=s[4], vz=s[5])
        fx, fy, fz, _ = refsim.reference_forces(2048, calc=1, lib="fixed")
        np.savez(os.path.join(OUT, "force_fixed_n2048.npz"), fx=fx, fy=fy, fz=fz)


def main():
    os.makedirs(OUT, exist_ok=True)
    if "--big-only" in sys.argv:
        with open(os.path.join(HERE, "golden_meta.json")) as f:
            meta = json.load(f)
        big(meta)
        with open(os.path.join(OUT, "golden_meta.json"), "w") as f:
            json.dump(meta, f, indent=1)
        print("wrote", OUT)
        return
    meta = {"provenance": {}, "init": {}, "force": {}, "step10": {}}
    try:
        meta["provenance"]["gpu"] = subprocess.run(
            ["nvidia-smi", "--query-gpu=name,driver_version", "--format=csv,noheader"],
            capture_output=True, text=True).stdout.strip()
        meta["provenance"]["nvcc"] = subprocess.run(["nvcc", "--version"], capture_output=True,
                                                    text=True).stdout.strip().splitlines()[-2]
    except Exception as e:  # noqa: BLE001
        meta["provenance"]["error"] = str(e)
    meta["provenance"]["reference_flags"] = "-O3 -use_fast_math -DDISABLE_GL -gencode arch=compute_100a,code=sm_100a"

    # generator + forces
    for n in (256, 2048, 12800, 25600, 262144):
        fx, fy, fz, init = refsim.reference_forces(n)
        meta["init"][str(n)] = {
            "fnv1a64": fnv1a64(init),
            "first4_bits": [[f"{int(v):08x}" for v in np.asarray(a[:4]).view(np.uint32)] for a in init],
        }
        meta["force"][str(n)] = {"fnv1a64": fnv1a64([fx, fy, fz]), "eps": 1.0e-7}
        if n == 2048:
            np.savez(os.path.join(OUT, "force_n2048.npz"), fx=fx, fy=fy, fz=fz)
        print("force", n, meta["force"][str(n)], flush=True)

    # 10 iterations, SimParam defaults
    for n in (2048, 25600):
        sim = refsim.RefSimulator(n, iters=10)
        sim.step()
        s = sim.state()
        sim.close()
        meta["step10"][str(n)] = {"fnv1a64": fnv1a64(s)}
        if n == 2048:
            np.savez(os.path.join(OUT, "step10_n2048.npz"), x=s[0], y=s[1], z=s[2], vx=s[3], vy=s[4], vz=s[5])
        print("step10", n, meta["step10"][str(n)], flush=True)

    # seeded uniform cloud through set_state, larger softening, ragged N
    rng = np.random.default_rng(20261017)
    n = 1000
    cloud = [rng.uniform(-50, 50, n).astype(np.float32) for _ in range(3)] + \
            [rng.uniform(-1, 1, n).astype(np.float32) for _ in range(3)]
    cloud[0][10] = cloud[0][11]; cloud[1][10] = cloud[1][11]; cloud[2][10] = cloud[2][11]  # coincident pair
    fx, fy, fz, _ = refsim.reference_forces(n, eps=1.0e-3, state=cloud)
    np.savez(os.path.join(OUT, "cloud_n1000.npz"), x=cloud[0], y=cloud[1], z=cloud[2], vx=cloud[3],
             vy=cloud[4], vz=cloud[5], fx=fx, fy=fy, fz=fz, eps=np.float32(1.0e-3))

    # shipped PREDICATED kernel: one default step
    n = 1024
    sim = refsim.RefSimulator(n, iters=1, calc=1)
    sim.step()
    s = sim.state()
    sim.close()
    np.savez(os.path.join(OUT, "predicated_n1024.npz"), x=s[0], y=s[1], z=s[2], vx=s[3], vy=s[4], vz=s[5])

    big(meta)
    with open(os.path.join(OUT, "golden_meta.json"), "w") as f:
        json.dump(meta, f, indent=1)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
