"""lab_one.py <lib.so> <bodies> [steps] -- times stepSim() of one (possibly SASS-patched) library; prints one line.
Used by tools/sass_lab.py run and tools/sweep_libs.py; checks the force hash against the reference golden when
NBODY_LAB_PARITY=1 (real schedules) -- synthetic lab blocks compute garbage by construction."""
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
    par = f" parity={'n/a' if want is None else got == want}"
sim.stepSim()
ms = []
for _ in range(steps):
    sim.stepSim()
    ms.append(sim.getLastStepDeviceTime())
name = sim.kernelName()
sim.close()
best = min(ms)
r = 6
groups = (n + 32 * r - 1) // (32 * r)
blocks_per_smsp = groups * (n / 32.0) / (148 * 4)
cyc = best * 1e-3 * 1.965e9 / blocks_per_smsp
print(f"ms={best:9.3f} cycles/block/SMSP={cyc:8.1f} cycles/pair={cyc / 96:6.2f} G/s={float(n) * n / best / 1e6:7.1f}{par} {name}")
