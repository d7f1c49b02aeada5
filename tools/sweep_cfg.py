"""sweep_cfg.py -- G inter/s of the production kernel over (N, R, segments); checks the force hash where a golden exists.
    python tools/sweep_cfg.py 262144 "2,4,6" "0,1,8,24"      (segments 0 = the library's own plan)"""
import hashlib, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cuda-to-sycl-nbody_b200"))
import nbody_b200 as nb
meta = json.load(open(os.path.join(ROOT, "tests", "golden", "golden_meta.json")))
sha = lambda a: hashlib.sha256(np.stack(a, axis=1).reshape(-1).tobytes()).hexdigest()
n = int(sys.argv[1])
rs = [int(x) for x in sys.argv[2].split(",")]
segs = [int(x) for x in sys.argv[3].split(",")]
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
for r in rs:
    for sg in segs:
        fam = int(os.environ.get("SWEEP_FAMILY", "4"))
        os.environ["NBODY_KERNEL_CONFIG"] = f"{r},32,{fam}"
        if sg: os.environ["NBODY_SEGS"] = str(sg)
        else: os.environ.pop("NBODY_SEGS", None)
        sim = nb.DiskGalaxySimulator(nb.SimParam(numParticles=n, simIterationsPerFrame=1))
        ok = None
        if str(n) in meta["force_sha256"]:
            ok = sha(sim.computeAccel()) == meta["force_sha256"][str(n)]
        sim.stepSim(); sim.stepSim()
        ms = []
        for _ in range(steps):
            sim.stepSim(); ms.append(sim.getLastStepDeviceTime())
        name = sim.kernelName()
        sim.close()
        best, med = min(ms), sorted(ms)[len(ms) // 2]
        print(f"N={n} R={r} segs={sg or 'auto'}: best {float(n)*n/best/1e6:7.1f}  median {float(n)*n/med/1e6:7.1f} G inter/s  ({100*float(n)*n/med/1e6/3722.5:.1f}%)  parity={ok}  {name}", flush=True)
