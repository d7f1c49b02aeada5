#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native all-pairs N-body step.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--bodies N | --weak-base B]

Metric (BASELINE.json): G body-interactions/s = steps * N^2 / seconds / 1e9 (self pair
included), and its fraction of the FP32 roofline at 20 flop/interaction.

Workload: the reference's default-seeded disk galaxy with SimParam defaults (G=2, dt=0.005,
damping=0.999998, distEps=1e-7, BRANCH), one force+integrate iteration per "step".
  --gpus 1 : N = 1,048,576  (BASELINE configs[2]: 1xB200, 1M bodies)
  --gpus>1 : N = 4,194,304  (BASELINE configs[3]: bodies sharded by i-range, position exchange
             per step), strong scaling.  Launched by torchrun, one rank per GPU;
             without torchrun env one process drives all N GPUs (the drop-in class's mode).
  --bodies N         any other size (configs[4]: 16,777,216 strong scaling at 1/2/4/8 GPUs)
  --weak-base B      weak scaling in WORK: N = B*sqrt(gpus) rounded to 256, so N^2/gpus is constant
                     (B = 5931642 ends at 16,777,216 bodies on 8 GPUs)

Parity: before anything is timed, one force pass of the (sharded) handle is hashed (SHA-256 of all 3N
float32 force components) and compared with what the UNMODIFIED reference kernel produced for the same
galaxy on a B200 (tests/golden/golden_meta.json: 262144, 400003, 1M, 4M, 16M).  A mismatch aborts the
run; sizes without a committed golden report "matches_reference_golden": null.

Timing: W >= 3 untimed warm-up steps, then K steps, each timed ON THE DEVICE by CUDA events
on the library's compute stream (nbody_last_step_device_ms: first launch -> last kernel and
position exchange of that step), L2 flushed between steps, barrier + device synchronise on
both sides of the timed region, MAX over ranks.  `value` has the state resident in HBM;
`e2e` is the same step driven through the C ABI with HOST buffers (nbody_set_state from
pinned memory + nbody_step + read-back into pinned memory) timed by the host clock; with one rank
per GPU every rank uploads all N positions (each GPU keeps a full replica) and reads back the bodies
it owns (nbody_read_local), as a process-per-GPU application would.

--impl reference: the reference's CPU implementation of this path cannot be built here (SYCL /
OpenCL-CPU toolchain absent), so the arm times the oracle's C/OpenMP port of
src_sycl/simulator.dp.cpp:315-360 on all host threads, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "cuda-to-sycl-nbody_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))

METRIC = "G body-interactions/s"
FLOP_PER_INTERACTION = 20.0
SMS, FP32_LANES = 148, 128


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


def fp32_peak_tflops(peaks):
    """FP32 FMA peak = SMs x 128 lanes x 2 flop x max SM clock (MEASURED_PEAKS.json: sm_max_mhz)."""
    return SMS * FP32_LANES * 2 * peaks.get("sm_max_mhz", 1965.0) * 1e6 / 1e12


# ---- torch.distributed plumbing (rendezvous, barrier, max over ranks) -------------------------
class Dist:
    """One process per GPU under torchrun; a no-op single rank otherwise."""

    def __init__(self, backend: str | None = None):
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.active = self.world > 1
        self.backend = backend
        if self.active:
            import torch
            import torch.distributed as dist
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29577")
            if backend is None:
                backend = "nccl" if torch.cuda.is_available() else "gloo"
            self.backend = backend
            if backend == "nccl":
                torch.cuda.set_device(self.local_rank)
            dist.init_process_group(backend=backend, rank=self.rank, world_size=self.world)
            self.dist = dist
            self.torch = torch

    def _dev(self):
        return self.torch.device("cuda", self.local_rank) if self.backend == "nccl" else self.torch.device("cpu")

    def barrier(self):
        if self.active:
            self.dist.barrier()

    def broadcast_bytes(self, payload: bytes | None, n: int) -> bytes:
        """rank 0's `payload` (n bytes) to every rank"""
        if not self.active:
            return payload
        t = self.torch.zeros(n, dtype=self.torch.uint8, device=self._dev())
        if self.rank == 0:
            t.copy_(self.torch.frombuffer(bytearray(payload), dtype=self.torch.uint8))
        self.dist.broadcast(t, src=0)
        return bytes(t.cpu().numpy().tobytes())

    def max(self, v: float) -> float:
        if not self.active:
            return v
        t = self.torch.tensor([v], dtype=self.torch.float64, device=self._dev())
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum(self, v: float) -> float:
        if not self.active:
            return v
        t = self.torch.tensor([v], dtype=self.torch.float64, device=self._dev())
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def close(self):
        if self.active:
            self.dist.destroy_process_group()


# ---- clocks during the timed region -----------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                c = float(r[1])
                if c < 500:  # idle sample before/after the load
                    continue
                sm.append(c)
                mx.append(float(r[2]))
                pw.append(float(r[3]))
                for k, nm in enumerate(names):
                    if r[4 + k].lower().startswith("active"):
                        reasons.add(nm)
            except Exception:  # noqa: BLE001
                continue
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_median": statistics.median(pw) if pw else None, "samples_under_load": len(sm),
                "reasons": sorted(reasons)}


# ---- CPU baseline (oracle port), bounded sample -----------------------------------------------
def cpu_baseline(n_bodies: int, target_s: float = 12.0):
    """Times the oracle's OpenMP port on forces of the first i_sample bodies against all N."""
    import oracle_lib
    if not os.path.exists(oracle_lib.ORACLE_LIB):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "oracle"], check=True, capture_output=True)
    o = oracle_lib.Oracle()
    o.set_num_threads(len(os.sched_getaffinity(0)))
    st = o.disk_galaxy(n_bodies)
    # two-stage calibration (thread start-up dominates a tiny probe), then ~target_s of work
    probe = 16 * o.num_threads()
    o.time_accel(st[0], st[1], st[2], 1.0e-7, 0, probe, 1)
    probe = max(probe, 4096)
    t = o.time_accel(st[0], st[1], st[2], 1.0e-7, 0, probe, 1)
    rate = probe * n_bodies / t
    i_sample = int(min(n_bodies, max(probe, (rate * target_s / n_bodies) // 16 * 16)))
    t = o.time_accel(st[0], st[1], st[2], 1.0e-7, 0, i_sample, 1)
    g = i_sample * n_bodies / t / 1e9
    return {"value": g, "unit": METRIC, "cores": o.num_threads(), "kind": "port",
            "sample": f"forces of the first {i_sample} bodies against all {n_bodies} (1 pass, {t:.2f} s), "
                      f"C/OpenMP restatement of src_sycl/simulator.dp.cpp:315-360, not the SYCL binary",
            "seconds": t, "host_cpus": os.cpu_count(), "cpu_model": cpu_model(),
            "config0": cpu_config0(o)}


def cpu_model() -> str:
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def cpu_config0(o, frames: int = 3):
    """BASELINE configs[0] as named: 25600 particles (100 x 256), 10 steps per frame, the whole frame
    (force + integrate, all 10 iterations) timed by the host clock as src_sycl/simulator.dp.cpp:59-112
    does; invocation README.md:71-75 (`run_nbody.sh -b dpcpp 100 10`).  Full frames, no sampling."""
    n, iters = 25600, 10
    st = o.disk_galaxy(n)
    st = o.step(st, iters=1)  # warm-up (thread pool, page faults)
    times = []
    for _ in range(frames):
        t0 = time.perf_counter()
        st = o.step(st, iters=iters)
        times.append(time.perf_counter() - t0)
    best = min(times)
    return {"workload": "N=25600 (100x256), simIterationsPerFrame=10, SimParam defaults, full frames",
            "ms_per_frame": best * 1e3, "ms_per_frame_all": [t * 1e3 for t in times],
            "value": iters * float(n) * n / best / 1e9, "unit": METRIC, "cores": o.num_threads(),
            "kind": "port", "timing": "host clock around the whole frame (10 fused force+integrate iterations)"}


def run_reference_arm(args, dist, emit):
    """--impl reference: the CPU port on all host threads, rank 0 only."""
    if dist.rank != 0:
        return
    n = workload_size(args)
    import oracle_lib
    if not os.path.exists(oracle_lib.ORACLE_LIB):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "oracle"], check=True, capture_output=True)
    o = oracle_lib.Oracle()
    o.set_num_threads(len(os.sched_getaffinity(0)))  # torchrun exports OMP_NUM_THREADS=1: use every core we may
    st = o.disk_galaxy(n)
    # bounded sample per step: ~3 s of CPU work (two-stage calibration: thread start-up dominates a tiny probe)
    probe = 16 * o.num_threads()
    o.time_accel(st[0], st[1], st[2], 1.0e-7, 0, probe, 1)
    probe = min(n, max(probe, 2048))
    t = o.time_accel(st[0], st[1], st[2], 1.0e-7, 0, probe, 1)
    i_sample = int(min(n, max(16, (probe / t * 3.0) // 16 * 16)))
    for _ in range(args.warmup):
        o.time_accel(st[0], st[1], st[2], 1.0e-7, 0, i_sample, 1)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        o.time_accel(st[0], st[1], st[2], 1.0e-7, 0, i_sample, 1)
    dt = time.perf_counter() - t0
    g = args.steps * i_sample * n / dt / 1e9
    sample = (f"each step = forces of {i_sample} of the {n} bodies against all {n}, i.e. {i_sample / n:.4f} of a full step, "
              f"rate-normalised to the same metric (a full CPU step of this N would take {n / i_sample * dt / args.steps:.0f} s); "
              f"C/OpenMP restatement of the reference's SYCL/OpenCL-CPU kernel; nbody_dpcpp itself cannot be built here")
    line = {"impl": "reference", "metric": METRIC, "value": g, "unit": "G inter/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": {"workload": f"disk galaxy N={n}, SimParam defaults, bounded i-sample", "n_bodies": n,
                                            "sample_fraction_of_a_step": i_sample / n, "same_config_as_gpu_arm": False},
            "cpu_baseline": {"value": g, "unit": "G inter/s", "cores": o.num_threads(), "kind": "port", "sample": sample,
                             "cpu_model": cpu_model(), "config0": cpu_config0(o)},
            "e2e": {"value": g, "unit": "G inter/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


def workload_size(args) -> int:
    if args.weak_base:
        return max(256, int(round(args.weak_base * math.sqrt(args.gpus) / 256.0)) * 256)
    return args.bodies or (1048576 if args.gpus == 1 else 4194304)


def golden_force_hash(n: int):
    try:
        meta = json.load(open(os.path.join(ROOT, "tests", "golden", "golden_meta.json")))
        return meta.get("force_sha256", {}).get(str(n))
    except Exception:  # noqa: BLE001
        return None


def parity_check(sim, n, np):
    """One force pass of the (sharded) handle against the reference kernel's committed golden."""
    a = sim.computeAccel()
    got = hashlib.sha256(np.stack(a, axis=1).reshape(-1).tobytes()).hexdigest()
    want = golden_force_hash(n)
    c64 = [c.astype(np.float64) for c in a]
    third_law = max(abs(c.sum()) / max(np.abs(c).sum(), 1e-300) for c in c64)
    return {"force_sha256": got, "reference_golden_sha256": want,
            "matches_reference_golden": (got == want) if want else None,
            "oracle": "unmodified reference kernel (/root/reference/src/simulator.cu:186-229) on B200, "
                      "tests/golden/make_golden.py --big-only" if want else "no committed golden for this N",
            "sum_F_over_sum_absF": third_law}


def traffic_for(n_gpus: int, n: int):
    """dram bytes per launch of the dominant kernel from the committed ncu capture OF THIS CONFIG, else None."""
    p = os.path.join(ROOT, "profiles", "r02_traffic.json")
    try:
        t = json.load(open(p))
        e = t.get(f"{n_gpus}x{n}")
        return (e or {}).get("dram_bytes_per_launch")
    except Exception:  # noqa: BLE001
        return None


def run_b200_arm(args, dist, emit):
    import numpy as np
    import torch

    import nbody_b200 as nb

    nb.load_library()  # raises if the CUDA library is not built -- no fallback
    if nb.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device visible; the product has no CPU path")
    n = workload_size(args)
    params = nb.SimParam(numParticles=n, simIterationsPerFrame=1)
    multi_proc = dist.active
    if multi_proc:
        uid = dist.broadcast_bytes(nb.nccl_unique_id() if dist.rank == 0 else None, 128)
        sim = nb.DiskGalaxySimulator(params, rank=dist.rank, world=dist.world, device=dist.local_rank, unique_id=uid)
        torch.cuda.set_device(dist.local_rank)
    else:
        sim = nb.DiskGalaxySimulator(params, n_gpus=args.gpus)
    dev = torch.device("cuda", dist.local_rank if multi_proc else 0)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def l2_flush():
        flush.zero_()
        torch.cuda.synchronize(dev)

    # ---- parity first: a fast wrong kernel is not measured --------------------------------------
    parity = parity_check(sim, n, np)  # collective in rank mode
    if parity["matches_reference_golden"] is False:
        raise SystemExit(f"bench.py: forces at N={n} on {args.gpus} GPU(s) differ from the reference golden "
                         f"({parity['force_sha256']} != {parity['reference_golden_sha256']})")

    # --quick (the 16M-body sweeps, >= 13 s per step even on 8 GPUs): one warm-up step -- the parity pass
    # above already ran the same kernel once -- and one e2e step; reported as such in the line
    warmup = max(1, args.warmup) if args.quick else max(3, args.warmup)
    for _ in range(warmup):
        sim.stepSim()

    # ---- device-resident timing ---------------------------------------------------------------
    sampler = ClockSampler(dist.local_rank if multi_proc else 0)
    if dist.rank == 0:
        sampler.start()
    launches0 = sim.launchCount()
    dist.barrier()
    torch.cuda.synchronize(dev)
    wall0 = time.perf_counter()
    dev_ms = 0.0
    per_step = []
    for _ in range(args.steps):
        l2_flush()
        sim.stepSim()
        per_step.append(sim.getLastStepDeviceTime())
        dev_ms += per_step[-1]
    torch.cuda.synchronize(dev)
    dist.barrier()
    wall = time.perf_counter() - wall0
    launches = sim.launchCount() - launches0
    clocks = sampler.stop() if dist.rank == 0 else None
    dev_ms = dist.max(dev_ms)
    launches = int(dist.sum(launches))
    value = args.steps * float(n) * n / (dev_ms * 1e-3) / 1e9

    # ---- end to end through the C ABI with pinned host buffers -----------------------------------
    b0, cnt = sim.localRange()
    host = [torch.empty(n, dtype=torch.float32).pin_memory() for _ in range(6)]
    hv = [t.numpy() for t in host]
    sim.readInto(*hv)  # full state once: every rank uploads all N positions below
    if multi_proc:
        loc = [torch.empty(cnt, dtype=torch.float32).pin_memory() for _ in range(6)]
        lv = [t.numpy() for t in loc]

    def read_back():
        if multi_proc:
            sim.readLocalInto(*lv)  # this rank's bodies: positions + velocities
        else:
            sim.readInto(*hv)       # D2H: positions + velocities, SoA, as recvFromDevice does

    e2e_steps = 1 if args.quick else max(2, min(args.steps, 5))
    if not args.quick:
        sim.setState(*hv)
        sim.stepSim()
        read_back()  # warm
    dist.barrier()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        sim.setState(*hv)          # H2D: positions (all N) + velocities (owned shard)
        sim.stepSim()              # one iteration
        read_back()
    torch.cuda.synchronize(dev)
    dist.barrier()
    e2e_s = dist.max(time.perf_counter() - t0)
    e2e_value = e2e_steps * float(n) * n / e2e_s / 1e9
    if multi_proc:
        h2d = sum(12 * n + 12 * nb.plan_shard(n, dist.world, r)[1] for r in range(dist.world))
        d2h = 24 * n  # every body's position + velocity is read back exactly once, by its owner
    else:
        h2d = 12 * n * args.gpus + 12 * n
        d2h = 24 * n

    kname = sim.kernelName()
    sim.close()

    # ---- same-N single-GPU point for the scaling curve (rank 0, outside every timed region) ------
    same_n = None
    if args.gpus > 1 and dist.rank == 0 and not args.no_same_n and float(n) * n <= 1.8e13:
        try:
            one = nb.DiskGalaxySimulator(params, n_gpus=1) if not multi_proc else \
                nb.DiskGalaxySimulator(params, rank=0, world=1, device=dist.local_rank)
            one.stepSim()
            ms = []
            for _ in range(2):
                l2_flush()
                one.stepSim()
                ms.append(one.getLastStepDeviceTime())
            one.close()
            v1 = float(n) * n / (sum(ms) / len(ms) * 1e-3) / 1e9
            same_n = {"value_same_n_1gpu": v1, "ms_per_step_1gpu": sum(ms) / len(ms), "steps": len(ms),
                      "efficiency_same_n": value / (args.gpus * v1)}
        except Exception as e:  # noqa: BLE001
            same_n = {"error": str(e)}
    dist.barrier()
    if dist.rank != 0:
        return

    peaks, peak_src = measured_peaks()
    peak_tf = fp32_peak_tflops(peaks)
    n_gpus = args.gpus
    achieved_tf = value * 1e9 * FLOP_PER_INTERACTION / 1e12 / n_gpus  # per GPU
    # HBM side of the roofline, to show the kernel is not memory bound: algorithmic bytes per step
    # per GPU = N*16 (positions read) + (N/P)*(16 vel r/w *2 + 16 pos write)
    alg_bytes = n * 16 + (n / n_gpus) * 48
    hbm_gbs = alg_bytes / (dev_ms / args.steps * 1e-3) / 1e9
    roofline = {"bound": "fp32", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                "frac": achieved_tf / peak_tf, "traffic": traffic_for(n_gpus, n),
                "peak_source": f"148 SMs x 128 FP32 lanes x 2 flop x sm_max_mhz of MEASURED_PEAKS.json ({peak_src})",
                "convention": "20 flop/interaction (north_star); the exact 12-op recipe + 1 MUFU per interaction is issue bound at "
                              "~76.9% of this (24 + 2 issue cycles per 64 interactions per SM sub-partition, profiles/r02_sass_lab.txt)",
                "per_gpu": True,
                # the practical ceiling of the exact recipe on this chip: per 64 interactions a sub-partition's issue port is
                # held 24 cycles by the 12 packed ops and 2 by the MUFUs (profiles/r02_sass_lab.txt) vs 20 at the 20-flop roofline
                "issue_bound": {"frac_of_roofline": 20.0 / 26.0, "achieved_frac_of_issue_bound": (achieved_tf / peak_tf) / (20.0 / 26.0)},
                "hbm": {"achieved_gbs": hbm_gbs, "peak_gbs": peaks.get("hbm_gbs"), "frac": hbm_gbs / peaks.get("hbm_gbs", 6448.4),
                        "algorithmic_bytes_per_step_per_gpu": alg_bytes},
                "note": "path is bound by the FP32 issue port, neither HBM nor tensor: see DESIGN.md section 5; "
                        "traffic = dram bytes per launch from the ncu capture of THIS config (profiles/r02_traffic.json) or null"}

    if args.weak_base:
        cfg_name = f"weak scaling in work: N = {args.weak_base}*sqrt(gpus) (BASELINE.json configs[4])"
    elif n == 1048576 and n_gpus == 1:
        cfg_name = "BASELINE.json configs[2]"
    elif n == 4194304:
        cfg_name = "BASELINE.json configs[3]"
    elif n == 16777216:
        cfg_name = "BASELINE.json configs[4] (strong)"
    elif n == 262144:
        cfg_name = "BASELINE.json configs[1]"
    else:
        cfg_name = "custom size"
    line = {"metric": METRIC, "value": value, "unit": "G inter/s", "n_gpus": n_gpus, "steps": args.steps,
            "warmup": warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "weak" if args.weak_base else "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"reference disk galaxy (mt19937 seed 5489), N={n}, SimParam defaults, "
                                   f"1 force+integrate iteration per step",
                       "n_bodies": n, "kernel": kname, "sharding": f"i-range x{n_gpus}" if n_gpus > 1 else "none",
                       "process_model": "torchrun, one rank per GPU" if multi_proc else "single process",
                       "l2": "flushed between timed steps (256 MiB memset)",
                       "baseline_config": cfg_name},
            "pct_fp32_roofline": 100.0 * achieved_tf / peak_tf,
            "roofline": roofline,
            "parity": parity,
            "e2e": {"value": e2e_value, "unit": "G inter/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "steps": e2e_steps, "ms_per_step": e2e_s / e2e_steps * 1e3,
                    "path": "nbody_set_state(pinned host SoA) + nbody_step + " +
                            ("nbody_read_local(pinned host SoA, the rank's own bodies)" if multi_proc
                             else "nbody_read_state(pinned host SoA)")},
            "quick": bool(args.quick),
            "gpu_launches": launches, "clocks": clocks, "wall_s_timed_region": wall,
            "ms_per_step_each": per_step}
    if same_n is not None:
        line["same_n_scaling"] = same_n
    if n_gpus == 1 and not args.no_cpu_baseline:
        try:
            line["cpu_baseline"] = cpu_baseline(n)
        except Exception as e:  # noqa: BLE001
            line["cpu_baseline"] = {"error": str(e)}
    if n_gpus == 1 and not args.no_ref_kernel and float(n) * n <= 2e12:
        try:
            import refsim
            if refsim.available():
                r = refsim.RefSimulator(n, iters=1)
                per_gw = {}
                for gw in (64, 128, 256):  # the reference's default (sim_param.cpp:20) and the README's table sizes
                    r.time_kernel(gw, 1)
                    per_gw[str(gw)] = r.time_kernel(gw, 3) / 3
                r.close()
                gw_best = min(per_gw, key=per_gw.get)
                ms = per_gw[gw_best]
                line["reference_cuda_kernel"] = {"value": float(n) * n / ms / 1e6, "unit": "G inter/s", "ms_per_step": ms,
                                                 "gw_size_best": int(gw_best), "ms_per_step_by_gw_size": per_gw,
                                                 "launches_timed_per_gw_size": 3,
                                                 "what": "unmodified particle_interaction<BRANCH> (oracle/_ref), best of gwSize 64/128/256, "
                                                         "same GPU, same N, CUDA events, after this arm's timed region",
                                                 "speedup_of_this_repo": value / (float(n) * n / ms / 1e6)}
        except Exception as e:  # noqa: BLE001
            line["reference_cuda_kernel"] = {"error": str(e)}
    emit(line)


def main():
    # stdout carries exactly ONE JSON line: everything else any library prints there (NCCL's version
    # banner, torchrun notices) is sent to stderr by pointing fd 1 at fd 2 for the duration of the run
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(line: dict):
        os.write(real_stdout, (json.dumps(line) + "\n").encode())

    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--bodies", type=int, default=0)
    ap.add_argument("--weak-base", type=int, default=0, help="weak scaling in work: N = base*sqrt(gpus)")
    ap.add_argument("--no-same-n", action="store_true", help="skip the same-N single-GPU point of multi-GPU runs")
    ap.add_argument("--quick", action="store_true", help="1 warm-up step and 1 e2e step (for the 16M-body sweeps)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ref-kernel", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        dist = Dist(backend="gloo")
        try:
            run_reference_arm(args, dist, emit)
        finally:
            dist.close()
        return
    dist = Dist()
    try:
        run_b200_arm(args, dist, emit)
    finally:
        dist.close()


if __name__ == "__main__":
    main()
