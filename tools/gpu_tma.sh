#!/bin/bash
# TMA-staged variant vs production kernel: parity (hash vs the reference's golden) and rate
python - <<'PY'
import os, sys, json
sys.path.insert(0, "cuda-to-sycl-nbody_b200"); sys.path.insert(0, "tests")
import nbody_b200 as nb, oracle_lib
o = oracle_lib.Oracle()
meta = json.load(open("tests/golden/golden_meta.json"))
for cfg in ("6,32,5", "4,32,5"):
    os.environ["NBODY_KERNEL_CONFIG"] = cfg
    sim = nb.DiskGalaxySimulator(nb.SimParam(numParticles=262144))
    ok = o.fnv1a64(sim.computeAccel()) == meta["force"]["262144"]["fnv1a64"]
    print(cfg, sim.kernelName(), "forces bit-equal to reference golden:", ok)
    sim.close()
    sim = nb.DiskGalaxySimulator(nb.SimParam(numParticles=25600, simIterationsPerFrame=10)); sim.stepSim()
    p, v = sim.getParticlePos(), sim.getParticleVel()
    print("   step10 bit-equal:", o.fnv1a64([p.x, p.y, p.z, v.x, v.y, v.z]) == meta["step10"]["25600"]["fnv1a64"])
    sim.close()
PY
for n in 1048576 524288; do
python tools/run_steps.py --n $n --kernel auto --steps 3 | tail -1 | cut -c12-200
python tools/run_steps.py --n $n --kernel auto --cfg 6,32,5 --steps 3 | tail -1 | cut -c12-200
python tools/run_steps.py --n $n --kernel auto --cfg 4,32,5 --steps 3 | tail -1 | cut -c12-200
done
