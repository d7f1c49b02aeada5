"""small-N comparison: our AUTO kernel vs the unmodified reference kernel (oracle/_ref), ms per iteration"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cuda-to-sycl-nbody_b200")); sys.path.insert(0, os.path.join(ROOT, "tools"))
import nbody_b200 as nb, refsim
for n in [int(v) for v in (sys.argv[1].split(',') if len(sys.argv) > 1 else '12800,25600,64000,131072'.split(','))]:
    iters = 64
    ref = refsim.RefSimulator(n, iters=1)
    best_ref = 1e9
    for gw in (64, 128, 256):
        ref.time_kernel(gw, 4)
        best_ref = min(best_ref, ref.time_kernel(gw, iters) / iters)
    ref.close()
    out = {"n": n, "reference_ms": best_ref}
    for name, kern, cfg in (("auto", nb.KERNEL_AUTO, ""), ("generic", nb.KERNEL_GENERIC, ""), ("scalar_r2_b64", nb.KERNEL_SCALAR, "2,64,2"),
                            ("wseg_r2_seg1", nb.KERNEL_AUTO, "2,32,3"), ("wsmall_r1", nb.KERNEL_AUTO, "1,32,6"),
                            ("wsmall_r2", nb.KERNEL_AUTO, "2,32,6")):
        if cfg: os.environ["NBODY_KERNEL_CONFIG"] = cfg
        else: os.environ.pop("NBODY_KERNEL_CONFIG", None)
        sim = nb.DiskGalaxySimulator(nb.SimParam(numParticles=n, simIterationsPerFrame=iters))
        sim.setKernel(kern)
        sim.stepSim(); sim.stepSim()
        out[name] = sim.getLastStepDeviceTime() / iters
        out[name + "_kernel"] = sim.kernelName()
        sim.close()
    print(json.dumps(out))
