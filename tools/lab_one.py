"""lab_one.py <lib.so> <bodies> [steps] -- times stepSim() of one (possibly SASS-patched) library; prints one line.
Used by tools/sass_lab.py run and the sweep scripts.  NBODY_LAB_PARITY=1: the force hash is computed and compared with
the reference golden where one exists (synthetic lab blocks compute garbage by construction -- leave it unset there).
The kernel is chosen by the library (AUTO) or forced through NBODY_KERNEL_CONFIG="r,32,4"."""
import hashlib, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cuda-to-sycl-nbody_b200"))
import nbody_b200 as nb
path, n = sys.argv[1], int(sys.argv[2])
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
lib = nb.load_library(os.path.abspath(path))
sim = nb.DiskGalaxySimulator(nb.SimParam(numParticles=n, simIterationsPerFrame=1), lib=lib)
par = ""
if os.environ.get("NBODY_LAB_PARITY") == "1":
    meta = json.load(open(os.path.join(ROOT, "tests", "golden", "golden_meta.json")))
    want = meta["force_sha256"].get(str(n))
    got = hashlib.sha256(np.stack(sim.computeAccel(), axis=1).reshape(-1).tobytes()).hexdigest()
    par = f" sha={got[:12]} parity={'n/a' if want is None else got == want}"
sim.stepSim()
ms = []
for _ in range(steps):
    sim.stepSim()
    ms.append(sim.getLastStepDeviceTime())
name = sim.kernelName()
sim.close()
best = min(ms)
cfg = os.environ.get("NBODY_KERNEL_CONFIG", "")
r = int(cfg.split(",")[0]) if cfg else 6
groups = (n + 32 * r - 1) // (32 * r)
blocks_per_smsp = groups * (n / 32.0) / (148 * 4)
cyc = best * 1e-3 * 1.965e9 / blocks_per_smsp
g = float(n) * n / best / 1e6
print(f"ms={best:9.3f} cycles/pair={cyc / (16 * r):6.2f} G/s={g:7.1f} ({g / 37.225:5.2f}%){par} {name}", flush=True)
