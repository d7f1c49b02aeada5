"""try_libs.py -- load each given build of the library, check the forces of the R = 6 production kernel
against the reference golden (N = 400003 and 1M) and time a few steps at N = 1M.
    python tools/try_libs.py lib_a.so lib_b.so ..."""
import hashlib, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cuda-to-sycl-nbody_b200"))
import nbody_b200 as nb
meta = json.load(open(os.path.join(ROOT, "tests", "golden", "golden_meta.json")))
sha = lambda a: hashlib.sha256(np.stack(a, axis=1).reshape(-1).tobytes()).hexdigest()
sizes = [int(x) for x in os.environ.get("TRY_SIZES", "400003,1048576").split(",")]
for path in [os.path.abspath(p) for p in sys.argv[1:]]:
    lib = nb.load_library(path)
    res = {}
    for n in sizes:
        sim = nb.DiskGalaxySimulator(nb.SimParam(numParticles=n, simIterationsPerFrame=1), lib=lib)
        ok = [sha(sim.computeAccel()) == meta["force_sha256"][str(n)] for _ in range(2)]
        name = sim.kernelName()
        ms = []
        if n == sizes[-1]:
            sim.stepSim()
            for _ in range(3):
                sim.stepSim(); ms.append(sim.getLastStepDeviceTime())
        sim.close()
        res[n] = (ok, name, [round(float(n) * n / m / 1e6, 1) for m in ms])
    print(os.path.basename(path), res, flush=True)
