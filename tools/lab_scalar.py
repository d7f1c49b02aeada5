"""lab_scalar.py <V0.so> <variant.so> ... -- one process: builds of the library whose scalar small-shard kernel
(force_wscalar_kernel<1, none, unit mass>) carries a generated tile body, against the build that keeps ptxas' code (V0).

 phase 1  every library, kernel forced to the scalar one (NBODY_KERNEL_CONFIG=1,32,6): SHA-256 of the forces must equal
          V0's at every size (and the reference golden where one is committed), best-of-5 stepSim() device time;
 phase 2  the fastest bit-exact variant: 10-iteration states against V0 (integrate epilogue, ragged sizes);
 phase 3  the numbers AUTO's cost model needs: the variant's scalar kernel, and R = 2 / R = 4 of V0, at 1..4 warps per
          sub-partition and at the reference's interactive sizes.
Writes gpurun_out/lab_scalar.txt (lines, flushed as they come), gpurun_out/lab_scalar.json and the name of the chosen
library into gpurun_out/lab_scalar_best.txt ("none" when no variant is both bit-exact and faster)."""
import hashlib, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cuda-to-sycl-nbody_b200"))
import nbody_b200 as nb

OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)
logf = open(os.path.join(OUT, "lab_scalar.txt"), "w")
T0 = time.time()
BUDGET = float(os.environ.get("LAB_BUDGET_S", "150"))


def log(*a):
    s = " ".join(str(x) for x in a)
    print(s, flush=True)
    logf.write(s + "\n")
    logf.flush()


meta = json.load(open(os.path.join(ROOT, "tests", "golden", "golden_meta.json")))
sha = lambda arrs: hashlib.sha256(np.stack(arrs, axis=1).reshape(-1).tobytes()).hexdigest()
paths = [os.path.abspath(p) for p in sys.argv[1:]]
names = [os.path.basename(p)[:-3] for p in paths]
libs = {n: nb.load_library(p) for n, p in zip(names, paths)}
res = {"phase1": {}, "phase2": {}, "phase3": {}}


def run(lib, n, cfg, steps=5, iters=1, forces=True):
    os.environ["NBODY_KERNEL_CONFIG"] = cfg
    sim = nb.DiskGalaxySimulator(nb.SimParam(numParticles=n, simIterationsPerFrame=iters), lib=lib)
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


# ---- phase 1 --------------------------------------------------------------------------------------------------------
sizes1 = [12800, 25600, 51200, 12801, 2048, 33]
ok = {n: True for n in names}
tot = {n: 0.0 for n in names}
for n in sizes1:
    base = None
    for k in names:
        r = run(libs[k], n, "1,32,6")
        if k == names[0]:
            base = r
        same = r["force_sha"] == base["force_sha"] and r["state_sha"] == base["state_sha"]
        gold = meta["force_sha256"].get(str(n))
        if gold is not None:
            same = same and r["force_sha"] == gold
        ok[k] = ok[k] and same
        if n >= 12800:
            tot[k] += r["ms"] / base["ms"]
        res["phase1"][f"{k}:{n}"] = r
        log(f"P1 {k:4s} N={n:6d} ms={r['ms']:8.4f} cyc/j={r['ms'] * 1e-3 * 1.965e9 / n:6.2f} vsV0={r['ms'] / base['ms']:.3f} "
            f"bit-exact={same} golden={'n/a' if gold is None else r['force_sha'] == gold} {r['kernel']}")
cands = [k for k in names[1:] if ok[k] and "+sass-gen" in res["phase1"][f"{k}:12800"]["kernel"]]
best = min(cands, key=lambda k: tot[k]) if cands else None
if best is not None and tot[best] >= tot[names[0]] * 0.985:
    log(f"best variant {best} is not faster than ptxas' code ({tot[best]:.3f} vs {tot[names[0]]:.3f})")
    best = None
log("ranking:", ", ".join(f"{k}={tot[k] / 4:.3f}{'' if ok[k] else '(MISMATCH)'}" for k in sorted(names, key=lambda k: tot[k])))
log("BEST", best)
open(os.path.join(OUT, "lab_scalar_best.txt"), "w").write(best or "none")
json.dump(res, open(os.path.join(OUT, "lab_scalar.json"), "w"), indent=1)
if best is None:
    sys.exit(0)

# ---- phase 2: states after 10 iterations, ragged sizes ----------------------------------------------------------------
for n in (1, 31, 1000, 4097, 12800, 40003):
    a = run(libs[names[0]], n, "1,32,6", steps=1, iters=10, forces=False)
    b = run(libs[best], n, "1,32,6", steps=1, iters=10, forces=False)
    same = a["state_sha"] == b["state_sha"]
    res["phase2"][str(n)] = same
    log(f"P2 N={n:6d} 20 iterations, state of {best} == state of V0: {same}")
    if not same:
        open(os.path.join(OUT, "lab_scalar_best.txt"), "w").write("none")
        log("BEST none (state mismatch)")
        json.dump(res, open(os.path.join(OUT, "lab_scalar.json"), "w"), indent=1)
        sys.exit(0)

# ---- phase 3: cost-model inputs ---------------------------------------------------------------------------------------
smsp = 148 * 4
sizes3 = [smsp * 32 * k for k in (1, 2, 3, 4, 5, 6)] + [6400, 20000, 32768, 40000, 57720, 64000, 65536, 80000, 90000, 102400, 113664, 131072]
for n in sorted(set(sizes3)):
    if time.time() - T0 > BUDGET:
        log("phase 3 cut short (time budget)")
        break
    row = {}
    for label, lib, cfg in (("scalar_gen", libs[best], "1,32,6"), ("scalar_ptxas", libs[names[0]], "1,32,6"),
                            ("r2", libs[names[0]], "2,32,4"), ("r4", libs[names[0]], "4,32,4")):
        r = run(lib, n, cfg, steps=3, forces=True)
        row[label] = r
    same = len({row[k]["force_sha"] for k in row}) == 1
    res["phase3"][str(n)] = row
    log(f"P3 N={n:6d} " + " ".join(f"{k}={row[k]['ms']:.4f}ms/{float(n) * n / row[k]['ms'] / 1e6 / 37.225:5.1f}%" for k in row) + f" all-equal={same}")
    json.dump(res, open(os.path.join(OUT, "lab_scalar.json"), "w"), indent=1)
log(f"done in {time.time() - T0:.1f} s")
