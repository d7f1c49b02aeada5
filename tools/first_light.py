"""GPU bring-up: parity against the reference kernel + kernel-variant sweep.  Run on a B200:

    gpurun -- python tools/first_light.py [--quick]

Writes gpurun_out/first_light.json.  Not part of the product or of the test-suite; the parity
tests proper live in tests/.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cuda-to-sycl-nbody_b200"))
sys.path.insert(0, os.path.join(ROOT, "tools"))

import nbody_b200 as nb  # noqa: E402
import refsim  # noqa: E402


def bits_equal(a, b):
    return bool(np.array_equal(a, b))


def relerr(ax, ay, az, bx, by, bz):
    num = np.sqrt((ax - bx).astype(np.float64) ** 2 + (ay - by).astype(np.float64) ** 2 + (az - bz).astype(np.float64) ** 2)
    den = np.sqrt(bx.astype(np.float64) ** 2 + by.astype(np.float64) ** 2 + bz.astype(np.float64) ** 2)
    return num / np.maximum(den, 1e-30)


def time_variant(n, kernel, cfg, iters, reps=3):
    """Returns best device-ms per iteration for a kernel variant."""
    if cfg:
        os.environ["NBODY_KERNEL_CONFIG"] = cfg
    else:
        os.environ.pop("NBODY_KERNEL_CONFIG", None)
    sim = nb.DiskGalaxySimulator(nb.SimParam(numParticles=n, simIterationsPerFrame=iters))
    sim.setKernel(kernel)
    name = sim.kernelName()
    sim.stepSim()  # warm-up
    best = 1e30
    for _ in range(reps):
        sim.stepSim()
        best = min(best, sim.getLastStepDeviceTime() / iters)
    sim.close()
    os.environ.pop("NBODY_KERNEL_CONFIG", None)
    return name, best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--sizes", default="")
    args = ap.parse_args()
    out = {"parity": [], "sweep": [], "reference_kernel": []}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    try:
        print(subprocess.run(["nvidia-smi", "--query-gpu=name,clocks.sm,clocks.max.sm,power.draw",
                              "--format=csv"], capture_output=True, text=True).stdout)
    except Exception as e:  # noqa: BLE001
        print("nvidia-smi failed", e)

    # ---- 1. generator + force parity against the reference kernel ----
    for n in ([12800, 25600] if args.quick else [12800, 25600, 262144]):
        fx, fy, fz, init = refsim.reference_forces(n)
        mine = nb.generate_disk_galaxy(n)
        gen_ok = all(bits_equal(a, b) for a, b in zip(mine, init))
        rec = {"n": n, "generator_bit_equal": gen_ok}
        for kernel, kname in ((nb.KERNEL_AUTO, "auto"), (nb.KERNEL_SCALAR, "scalar"), (nb.KERNEL_GENERIC, "generic")):
            sim = nb.DiskGalaxySimulator(nb.SimParam(numParticles=n))
            sim.setKernel(kernel)
            ax, ay, az = sim.computeAccel()
            e = relerr(ax, ay, az, fx, fy, fz)
            rec[kname] = {"kernel": sim.kernelName(),
                          "bit_equal": bits_equal(ax, fx) and bits_equal(ay, fy) and bits_equal(az, fz),
                          "max_rel": float(e.max()), "median_rel": float(np.median(e))}
            sim.close()
        print(json.dumps(rec), flush=True)
        out["parity"].append(rec)

    # ---- 2. 10 iterations, default params, vs reference ----
    for n in [25600]:
        ref = refsim.RefSimulator(n, iters=10)
        ref.step()
        rs = ref.state()
        ref.close()
        sim = nb.DiskGalaxySimulator(nb.SimParam(numParticles=n, simIterationsPerFrame=10))
        sim.stepSim()
        p, v = sim.getParticlePos(), sim.getParticleVel()
        ms = [p.x, p.y, p.z, v.x, v.y, v.z]
        rec = {"n": n, "steps": 10, "bit_equal": all(bits_equal(a, b) for a, b in zip(ms, rs)),
               "max_abs_pos_diff": float(max(np.abs(a - b).max() for a, b in zip(ms[:3], rs[:3])))}
        sim.close()
        print(json.dumps(rec), flush=True)
        out["parity"].append(rec)

    # ---- 3. reference kernel throughput (the kernel to beat) ----
    for n in ([] if args.quick else [1048576]):
        ref = refsim.RefSimulator(n, iters=1)
        for gw in (64, 128, 256):
            ref.time_kernel(gw, 1)
            ms = ref.time_kernel(gw, 2) / 2
            rec = {"n": n, "gw": gw, "ms": ms, "ginter_s": n * n / ms / 1e6}
            print("reference", json.dumps(rec), flush=True)
            out["reference_kernel"].append(rec)
        ref.close()

    # ---- 4. variant sweep ----
    sizes = [262144] if args.quick else [262144, 524288, 1048576]
    if args.sizes:
        sizes = [int(v) for v in args.sizes.split(",")]
    variants = [(nb.KERNEL_AUTO, "")]
    variants += [(nb.KERNEL_AUTO, c) for c in ("2,32,3", "4,32,3", "6,32,3", "8,32,3", "2,64,3", "4,64,3", "4,128,3")]
    variants += [(nb.KERNEL_PACKED, c) for c in ("2,128,1", "4,256,1")]
    variants += [(nb.KERNEL_SCALAR, c) for c in ("4,256,2",)]
    for n in sizes:
        for kernel, cfg in variants:
            iters = 4 if n <= 262144 else 1
            t0 = time.time()
            name, ms = time_variant(n, kernel, cfg, iters)
            g = n * n / ms / 1e6
            rec = {"n": n, "kernel": name, "ms_per_iter": ms, "ginter_s": g, "pct_roofline_3722": 100 * g / 3722.0,
                   "wall_s": time.time() - t0}
            print(json.dumps(rec), flush=True)
            out["sweep"].append(rec)

    with open(os.path.join(ROOT, "gpurun_out", "first_light.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
