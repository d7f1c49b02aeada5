"""small_n.py -- ms per iteration at the reference's interactive sizes and at BASELINE configs[1]: this library (AUTO)
against the UNMODIFIED reference kernel at its best gwSize of 64/128/256, same GPU, same galaxy.

    python tools/small_n.py [N ...]        (under gpurun; prints one table row per size)
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cuda-to-sycl-nbody_b200"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import nbody_b200 as nb
import refsim
sizes = [int(a) for a in sys.argv[1:]] or [12800, 25600, 51200, 64000, 102400, 131072, 262144]
ITERS = 20
print(f"{'N':>8s} {'kernel':44s} {'ours ms/iter':>12s} {'G inter/s':>10s} {'%roof':>6s} | {'reference ms/iter (gwSize)':>27s} {'G inter/s':>10s} {'speed-up':>8s}")
for n in sizes:
    sim = nb.DiskGalaxySimulator(nb.SimParam(numParticles=n, simIterationsPerFrame=ITERS))
    sim.stepSim()
    ms = []
    for _ in range(5):
        sim.stepSim()
        ms.append(sim.getLastStepDeviceTime() / ITERS)
    name = sim.kernelName()
    sim.close()
    ours = min(ms)
    r = refsim.RefSimulator(n, iters=1)
    per = {}
    for gw in (64, 128, 256):
        r.time_kernel(gw, 2)
        per[gw] = r.time_kernel(gw, 10) / 10
    r.close()
    gw = min(per, key=per.get)
    g = n * float(n) / ours / 1e6
    gr = n * float(n) / per[gw] / 1e6
    print(f"{n:8d} {name:44s} {ours:12.4f} {g:10.1f} {g / 37.225:6.1f} | {per[gw]:20.4f} ({gw:3d}) {gr:10.1f} {per[gw] / ours:8.2f}", flush=True)
