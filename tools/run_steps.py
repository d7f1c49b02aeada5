"""Runs a few stepSim() calls of one kernel configuration (profiling / clock-sampling target).

    python tools/run_steps.py --n 1048576 --kernel packed --cfg 4,256 --steps 5 [--iters 1]
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cuda-to-sycl-nbody_b200"))
import nbody_b200 as nb  # noqa: E402

KERNELS = {"auto": nb.KERNEL_AUTO, "generic": nb.KERNEL_GENERIC, "packed": nb.KERNEL_PACKED, "scalar": nb.KERNEL_SCALAR}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1048576)
    ap.add_argument("--kernel", default="auto", choices=sorted(KERNELS))
    ap.add_argument("--cfg", default="")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--iters", type=int, default=1)
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--segs", type=int, default=0, help="NBODY_SEGS override")
    ap.add_argument("--variants", action="store_true", help="load libnbody_b200_variants.so (comparison kernels)")
    ap.add_argument("--mass", action="store_true", help="per-body masses (uniform in [0.5, 1.5)): times the MASS instantiations")
    a = ap.parse_args()
    if a.cfg:
        os.environ["NBODY_KERNEL_CONFIG"] = a.cfg
    if a.segs:
        os.environ["NBODY_SEGS"] = str(a.segs)
    lib = nb.load_library(nb.VARIANTS_LIB_PATH) if a.variants or a.kernel in ("packed", "scalar") else None
    sim = nb.DiskGalaxySimulator(nb.SimParam(numParticles=a.n, simIterationsPerFrame=a.iters), n_gpus=a.gpus, lib=lib)
    sim.setKernel(KERNELS[a.kernel])
    if a.mass:
        import numpy as np
        sim.setMass(np.random.default_rng(1).uniform(0.5, 1.5, a.n).astype(np.float32))
    for s in range(a.steps):
        sim.stepSim()
        ms = sim.getLastStepDeviceTime() / a.iters
        print(json.dumps({"step": s, "kernel": sim.kernelName(), "n": a.n, "gpus": a.gpus, "ms_per_iter": ms,
                          "host_ms": sim.getLastStepTime() / a.iters,
                          "ginter_s": a.n * a.n / ms / 1e6}), flush=True)
    sim.close()


if __name__ == "__main__":
    main()
