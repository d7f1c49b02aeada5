"""lab_relay.py [lib.so] -- one process on one B200: the accumulator-relay kernel (force_wrelay_kernel<W, TJ>, family 7)
against the scalar one-body-per-lane kernel (family 6, validated bit-exact against the reference kernel) and the packed
R = 2 / R = 4 kernels: SHA-256 of the forces, of the state after 2 x 10 iterations, with per-body masses, and the
best-of-5 device time per iteration.  Kernels are forced through NBODY_KERNEL_CONFIG="r,block,family".
Writes gpurun_out/lab_relay.txt / .json as it goes."""
import hashlib, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cuda-to-sycl-nbody_b200"))
import nbody_b200 as nb

OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)
logf = open(os.path.join(OUT, "lab_relay.txt"), "w")
T0 = time.time()
BUDGET = float(os.environ.get("LAB_BUDGET_S", "120"))


def log(*a):
    s = " ".join(str(x) for x in a)
    print(s, flush=True)
    logf.write(s + "\n")
    logf.flush()


meta = json.load(open(os.path.join(ROOT, "tests", "golden", "golden_meta.json")))
sha = lambda arrs: hashlib.sha256(np.stack(arrs, axis=1).reshape(-1).tobytes()).hexdigest()
# the 2- and 8-warp shapes only exist in the VARIANTS build of the library
lib = nb.load_library(os.path.abspath(sys.argv[1]) if len(sys.argv) > 1 else nb.VARIANTS_LIB_PATH)
res = {}
CFGS = {"scalar": "1,32,6", "relay4x16": "16,128,7", "relay4x32": "32,128,7", "relay2x16": "16,64,7", "relay2x32": "32,64,7",
        "relay8x16": "16,256,7", "r2": "2,32,4", "r4": "4,32,4"}


def run(n, cfg, steps=5, iters=1, forces=True, mass=False):
    if cfg:
        os.environ["NBODY_KERNEL_CONFIG"] = cfg
    else:
        os.environ.pop("NBODY_KERNEL_CONFIG", None)
    sim = nb.DiskGalaxySimulator(nb.SimParam(numParticles=n, simIterationsPerFrame=iters), lib=lib)
    if mass:
        sim.setMass(np.random.default_rng(1).uniform(0.5, 1.5, n).astype(np.float32))
    h = sha(sim.computeAccel()) if forces else None
    sim.stepSim()
    ms = []
    for _ in range(steps):
        sim.stepSim()
        ms.append(sim.getLastStepDeviceTime() / iters)
    name = sim.kernelName()
    p, v = sim.getParticlePos(), sim.getParticleVel()
    hs = sha([p.x, p.y, p.z, v.x, v.y, v.z])
    sim.close()
    return dict(ms=min(ms), force_sha=h, state_sha=hs, kernel=name)


def save():
    json.dump(res, open(os.path.join(OUT, "lab_relay.json"), "w"), indent=1)


allok = True
# ---- phase 1: parity (forces + one-iteration states) and time, all relay shapes -----------------------------------
for n in (12800, 1, 31, 33, 1000, 2048, 4097, 6400, 12801, 18944, 25600, 37888, 51200, 65536):
    if time.time() - T0 > BUDGET:
        log("phase 1 cut short (time budget)")
        break
    row = {}
    for k, cfg in CFGS.items():
        if k in ("r2", "r4") and n < 64:
            continue
        row[k] = run(n, cfg)
    base = row["scalar"]
    gold = meta["force_sha256"].get(str(n))
    same = {k: (r["force_sha"] == base["force_sha"] and r["state_sha"] == base["state_sha"]) for k, r in row.items()}
    allok = allok and all(same.values()) and (gold is None or gold == base["force_sha"])
    res[f"p1:{n}"] = row
    log(f"P1 N={n:6d} " + " ".join(f"{k}={r['ms']:.4f}{'' if same[k] else '(MISMATCH)'}" for k, r in row.items())
        + f" golden={'n/a' if gold is None else gold == base['force_sha']}")
    save()
log("kernel names:", ", ".join(sorted({r["kernel"] for k in res for r in res[k].values()})))
# ---- phase 2: 2 x 10 iterations (the hand-off of the integrate epilogue, ragged sizes), and per-body masses -----------
for n in (1, 17, 4097, 12800, 20011):
    for mass in (False, True):
        a = run(n, CFGS["scalar"], steps=1, iters=10, forces=True, mass=mass)
        for k in ("relay4x16", "relay4x32", "relay2x16", "relay2x32", "relay8x16"):
            b = run(n, CFGS[k], steps=1, iters=10, forces=True, mass=mass)
            same = a["state_sha"] == b["state_sha"] and a["force_sha"] == b["force_sha"]
            allok = allok and same
            res[f"p2:{n}:{int(mass)}:{k}"] = same
            log(f"P2 N={n:6d} mass={int(mass)} {k}: forces + state after 20 iterations == scalar kernel's: {same}  ({b['kernel']})")
    save()
log("ALL BIT-EXACT" if allok else "MISMATCH SOMEWHERE")
res["all_bit_exact"] = allok
save()
log(f"done in {time.time() - T0:.1f} s")
