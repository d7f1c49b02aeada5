"""Parity of the CUDA path (through the C ABI) against
  (1) the UNMODIFIED reference simulator running on the same GPU (oracle/_ref)  -> bit-exact,
  (2) the committed golden fixtures the reference produced on a B200             -> bit-exact,
  (3) the CPU oracle (oracle/nbody_oracle.c)                                    -> stated tolerance
      (the CPU cannot reproduce MUFU.RSQ bit-for-bit),
and size-independent properties at BASELINE.json's full sizes.

north_star contract: per-body single-step accelerations within 1e-5 relative, positions within
1e-4 relative after 10 steps.  The kernels keep the reference's op order, so the observed
difference is 0 ulp; both the contract tolerance and bit equality are asserted.
"""
from __future__ import annotations

import hashlib
import json
import os

import numpy as np
import pytest

from oracle_lib import rel_err

pytestmark = pytest.mark.gpu

ACCEL_TOL = 1e-5   # north_star: per-body single-step accelerations, relative
POS_TOL = 1e-4     # north_star: positions after 10 steps, relative


def _mk(nb, n, lib=None, **kw):
    return nb.DiskGalaxySimulator(nb.SimParam(numParticles=n, **kw), lib=lib)


def sha256_f32(arrays) -> str:
    """SHA-256 of the float32 byte stream interleaved per body (tests/golden/make_golden.py)."""
    inter = np.stack([np.ascontiguousarray(a, np.float32) for a in arrays], axis=1).reshape(-1)
    return hashlib.sha256(inter.tobytes()).hexdigest()


@pytest.fixture(scope="module")
def vlib(nb):
    """libnbody_b200_variants.so: the product library plus the comparison kernels (make VARIANTS=1)."""
    import subprocess
    if not os.path.exists(nb.VARIANTS_LIB_PATH):
        subprocess.run(["make", "-C", os.path.dirname(os.path.dirname(nb.VARIANTS_LIB_PATH)), "variants"],
                       check=True, capture_output=True)
    return nb.load_library(nb.VARIANTS_LIB_PATH)


def _state(sim):
    p, v = sim.getParticlePos(), sim.getParticleVel()
    return [p.x.copy(), p.y.copy(), p.z.copy(), v.x.copy(), v.y.copy(), v.z.copy()]


def _assert_bits(got, want, what):
    for k, (a, b) in enumerate(zip(got, want)):
        assert np.array_equal(a, b), f"{what}: component {k} differs in {(a != b).sum()} of {a.size} bodies"


KERNELS = ["auto", "packed", "scalar", "generic"]


def _kernel_id(nb, name):
    return {"auto": nb.KERNEL_AUTO, "packed": nb.KERNEL_PACKED, "scalar": nb.KERNEL_SCALAR,
            "generic": nb.KERNEL_GENERIC}[name]


# ---------------------------------------------------------------------------------------------
# (1) live reference on the same GPU
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [256, 12800, 25600])
def test_generator_matches_reference_ctor(nb, ref, n):
    sim = ref.RefSimulator(n)
    want = sim.state()
    sim.close()
    _assert_bits(nb.generate_disk_galaxy(n), want, f"galaxy N={n}")
    ours = _mk(nb, n)
    _assert_bits(_state(ours), want, f"constructed state N={n}")
    ours.close()


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("n", [25600, 262144])
def test_single_step_accelerations_vs_reference(nb, vlib, ref, n, kernel):
    """BASELINE config 2: 262144 bodies, single-step force accuracy vs the reference kernel.
    "packed" / "scalar" are the CTA-tiled comparison kernels of the VARIANTS library."""
    fx, fy, fz, _ = ref.reference_forces(n)
    sim = _mk(nb, n, lib=vlib if kernel in ("packed", "scalar") else None)
    sim.setKernel(_kernel_id(nb, kernel))
    a = sim.computeAccel()
    sim.close()
    e = rel_err(a, [fx, fy, fz])
    assert e.max() <= ACCEL_TOL, f"max rel err {e.max():.3g}"
    _assert_bits(a, [fx, fy, fz], f"forces N={n} {kernel}")


@pytest.mark.parametrize("kernel", ["auto", "generic"])
def test_ten_steps_vs_reference(nb, ref, kernel):
    """positions after 10 steps (SimParam defaults) -- contract 1e-4 relative, observed 0 ulp."""
    n = 25600
    r = ref.RefSimulator(n, iters=10)
    r.step()
    want = r.state()
    r.close()
    sim = _mk(nb, n, simIterationsPerFrame=10)
    sim.setKernel(_kernel_id(nb, kernel))
    sim.stepSim()
    got = _state(sim)
    sim.close()
    e = rel_err(got[:3], want[:3])
    assert e.max() <= POS_TOL
    _assert_bits(got, want, "state after 10 steps")


def test_frames_accumulate_like_reference(nb, ref):
    """3 frames of 4 iterations == what the reference holds after 3 stepSim() calls."""
    n = 12800
    r = ref.RefSimulator(n, iters=4)
    sim = _mk(nb, n, simIterationsPerFrame=4)
    for _ in range(3):
        r.step()
        sim.stepSim()
        _assert_bits(_state(sim), r.state(), "frame state")
    r.close()
    sim.close()


@pytest.mark.parametrize("n", [1, 2, 31, 255, 257, 1000, 4097])
def test_ragged_sizes_vs_reference(nb, ref, n):
    """numParticles need not be a multiple of any tile size (SimParam can be set directly)."""
    rng = np.random.default_rng(n)
    st = [rng.uniform(-20, 20, n).astype(np.float32) for _ in range(3)] + \
         [rng.uniform(-1, 1, n).astype(np.float32) for _ in range(3)]
    r = ref.RefSimulator(n, iters=3, eps=1e-3)
    r.set_state(*st)
    r.step()
    want = r.state()
    r.close()
    for kernel in ("auto", "generic"):
        sim = _mk(nb, n, simIterationsPerFrame=3, distEps=1e-3)
        sim.setKernel(_kernel_id(nb, kernel))
        sim.setState(*st)
        sim.stepSim()
        _assert_bits(_state(sim), want, f"N={n} {kernel}")
        sim.close()


def test_zero_softening_uses_predicated_self_term(nb, ref):
    """distEps = 0: rsqrt(0) = inf for the self pair, so the unpredicated kernels would produce
    NaN; AUTO must fall back to the predicated generic kernel and still match bit-for-bit."""
    n = 2048
    r = ref.RefSimulator(n, iters=2, eps=0.0)
    r.step()
    want = r.state()
    r.close()
    sim = _mk(nb, n, simIterationsPerFrame=2, distEps=0.0)
    assert "generic" in sim.kernelName()
    with pytest.raises(nb.NBodyError, match="bit-exact"):
        sim.setKernel(nb.KERNEL_PACKED)
    sim.stepSim()
    got = _state(sim)
    sim.close()
    assert all(np.isfinite(a).all() for a in got)
    _assert_bits(got, want, "eps=0")


def test_predicated_method_reproduces_shipped_behaviour(nb, ref):
    """calcMethod=PREDICATED multiplies by (i == id) in the shipped source (src/simulator.cu:209):
    zero force.  Drop-in means reproducing that, not the README's intent."""
    n = 4096
    r = ref.RefSimulator(n, iters=2, calc=1)
    r.step()
    want = r.state()
    r.close()
    sim = _mk(nb, n, simIterationsPerFrame=2, calcMethod=nb.CALC_PREDICATED)
    sim.stepSim()
    _assert_bits(_state(sim), want, "PREDICATED")
    sim.close()


def test_coincident_bodies_and_large_coordinates(nb, ref):
    n = 3000
    rng = np.random.default_rng(7)
    st = [rng.uniform(-1e4, 1e4, n).astype(np.float32) for _ in range(3)] + \
         [np.zeros(n, np.float32) for _ in range(3)]
    for k in range(3):
        st[k][100:110] = st[k][100]  # ten bodies at one point: r = 0 for j != i
    fx, fy, fz, _ = ref.reference_forces(n, eps=1e-2, state=st)
    sim = _mk(nb, n, distEps=1e-2)
    sim.setState(*st)
    _assert_bits(sim.computeAccel(), [fx, fy, fz], "coincident bodies")
    sim.close()


# ---------------------------------------------------------------------------------------------
# (2) committed golden fixtures (made by tests/golden/make_golden.py from the reference on a B200)
# ---------------------------------------------------------------------------------------------
def test_golden_forces_n2048(nb, vlib, golden_dir):
    g = np.load(os.path.join(golden_dir, "force_n2048.npz"))
    for kernel in KERNELS:
        sim = _mk(nb, 2048, lib=vlib if kernel in ("packed", "scalar") else None)
        sim.setKernel(_kernel_id(nb, kernel))
        _assert_bits(sim.computeAccel(), [g["fx"], g["fy"], g["fz"]], f"golden forces {kernel}")
        sim.close()


def test_golden_step10_n2048(nb, golden_dir):
    g = np.load(os.path.join(golden_dir, "step10_n2048.npz"))
    sim = _mk(nb, 2048, simIterationsPerFrame=10)
    sim.stepSim()
    _assert_bits(_state(sim), [g[k] for k in ("x", "y", "z", "vx", "vy", "vz")], "golden step10")
    sim.close()


def test_golden_cloud_n1000(nb, golden_dir):
    g = np.load(os.path.join(golden_dir, "cloud_n1000.npz"))
    sim = _mk(nb, 1000, distEps=float(g["eps"]))
    sim.setState(*[g[k] for k in ("x", "y", "z", "vx", "vy", "vz")])
    _assert_bits(sim.computeAccel(), [g["fx"], g["fy"], g["fz"]], "golden cloud")
    sim.close()


def test_golden_predicated_n1024(nb, golden_dir):
    g = np.load(os.path.join(golden_dir, "predicated_n1024.npz"))
    sim = _mk(nb, 1024, simIterationsPerFrame=1, calcMethod=nb.CALC_PREDICATED)
    sim.stepSim()
    _assert_bits(_state(sim), [g[k] for k in ("x", "y", "z", "vx", "vy", "vz")], "golden predicated")
    sim.close()


def test_golden_predicated_fixed(nb, golden_dir):
    """README-intended PREDICATED, force += r*inv*(i != id) (README.md:229-231): fixtures from the
    reference built with src/simulator.cu:209 patched to (i != id) (oracle/Makefile: ref_fixed)."""
    g = np.load(os.path.join(golden_dir, "predicated_fixed_n1024.npz"))
    sim = _mk(nb, 1024, simIterationsPerFrame=1, calcMethod=nb.CALC_PREDICATED_FIXED)
    assert "predicated_fixed" in sim.kernelName()
    sim.stepSim()
    _assert_bits(_state(sim), [g[k] for k in ("x", "y", "z", "vx", "vy", "vz")], "golden predicated (fixed)")
    sim.close()
    g = np.load(os.path.join(golden_dir, "force_fixed_n2048.npz"))
    sim = _mk(nb, 2048, calcMethod=nb.CALC_PREDICATED_FIXED)
    _assert_bits(sim.computeAccel(), [g["fx"], g["fy"], g["fz"]], "golden forces, predicated (fixed)")
    sim.close()
    meta = json.load(open(os.path.join(golden_dir, "golden_meta.json")))
    for n in (25600, 262144):
        sim = _mk(nb, n, calcMethod=nb.CALC_PREDICATED_FIXED)
        assert sha256_f32(sim.computeAccel()) == meta["force_fixed_sha256"][str(n)], n
        sim.close()


@pytest.mark.parametrize("n", [25600, 4097])
def test_predicated_fixed_vs_live_patched_reference(nb, ref, n):
    """the same against the patched reference running on this GPU (forces and one default step);
    for the default eps it must also equal BRANCH bit-for-bit only where r*inv*1 + a == fma(r, inv, a)
    -- it does not in general (FMUL then FFMA rounds twice), which is why it has its own oracle"""
    if not ref.available("fixed"):
        pytest.skip("oracle/_ref/libnbody_ref_fixed.so not built")
    fx, fy, fz, _ = ref.reference_forces(n, calc=1, lib="fixed")
    sim = _mk(nb, n, calcMethod=nb.CALC_PREDICATED_FIXED)
    _assert_bits(sim.computeAccel(), [fx, fy, fz], f"forces N={n} predicated-fixed")
    sim.close()
    r = ref.RefSimulator(n, iters=3, calc=1, lib="fixed")
    r.step()
    want = r.state()
    r.close()
    sim = _mk(nb, n, simIterationsPerFrame=3, calcMethod=nb.CALC_PREDICATED_FIXED)
    sim.stepSim()
    _assert_bits(_state(sim), want, f"3 steps N={n} predicated-fixed")
    sim.close()


def test_golden_hashes_large(nb, oracle, golden_dir):
    """forces at N=262144 and 10 steps at N=25600 against the hashes the reference produced"""
    meta = json.load(open(os.path.join(golden_dir, "golden_meta.json")))
    sim = _mk(nb, 262144)
    assert oracle.fnv1a64(sim.computeAccel()) == meta["force"]["262144"]["fnv1a64"]
    sim.close()
    sim = _mk(nb, 25600, simIterationsPerFrame=10)
    sim.stepSim()
    assert oracle.fnv1a64(_state(sim)) == meta["step10"]["25600"]["fnv1a64"]
    sim.close()


# ---------------------------------------------------------------------------------------------
# (3) CPU oracle, stated tolerance
# ---------------------------------------------------------------------------------------------
def test_cuda_vs_cpu_oracle_forces(nb, oracle):
    """The CPU restatement uses 1/sqrtf where the GPU uses MUFU.RSQ (~1 ulp apart per term);
    the sums then differ at the level of the reference's own rounding noise.  Tolerance:
    median <= 1e-6, max <= 2e-4 relative at N=4096 (the reference itself is ~2e-6 median /
    4e-5 max from FP64 truth at N=25600)."""
    n = 4096
    sim = _mk(nb, n)
    got = sim.computeAccel()
    st = _state(sim)
    sim.close()
    want = oracle.accel(st[0], st[1], st[2], 1.0e-7)
    e = rel_err(got, want)
    assert np.median(e) <= 1e-6 and e.max() <= 2e-4, (np.median(e), e.max())


def test_cuda_vs_fp64_truth_no_worse_than_contract(nb, oracle):
    """512-body subsample against FP64: the exact-order sum is as far from truth as the
    reference is (same bits), reported for the error budget."""
    n = 25600
    sim = _mk(nb, n)
    got = sim.computeAccel()
    st = _state(sim)
    sim.close()
    truth = oracle.accel_f64(st[0], st[1], st[2], 1.0e-7, 0, 512)
    e = rel_err([g[:512] for g in got], truth)
    assert np.median(e) < 1e-4 and e.max() < 1e-2


# ---------------------------------------------------------------------------------------------
# full-size properties (BASELINE configs 2/3: N = 1,048,576)
# ---------------------------------------------------------------------------------------------
def test_full_size_forces_and_step_vs_reference_golden(nb, golden_dir):
    """BASELINE configs[2], N = 1,048,576: the production kernel (R = 6, j-segmented) against what
    the UNMODIFIED reference kernel produced for the same galaxy on a B200 -- SHA-256 of all 3M force
    components and of the full state after one default step (tests/golden/make_golden.py --big-only).
    Plus Newton's third law as a size-independent sanity property."""
    meta = json.load(open(os.path.join(golden_dir, "golden_meta.json")))
    n = 1048576
    sim = _mk(nb, n, simIterationsPerFrame=1)
    assert "wseg_f32x2_r6" in sim.kernelName() and "+sass-gen" in sim.kernelName(), sim.kernelName()
    a = sim.computeAccel()
    assert sha256_f32(a) == meta["force_sha256"][str(n)]
    for c in a:
        c64 = c.astype(np.float64)
        assert abs(c64.sum()) <= 1e-3 * np.abs(c64).sum()
    sim.stepSim()
    assert sha256_f32(_state(sim)) == meta["step1_sha256"][str(n)]
    sim.close()


@pytest.mark.parametrize("n", [262144, 400003])
def test_mid_size_forces_vs_reference_golden(nb, golden_dir, n):
    """configs[1] (262144, R = 4 kernel) and a ragged size above the R = 6 switch point"""
    meta = json.load(open(os.path.join(golden_dir, "golden_meta.json")))
    sim = _mk(nb, n, simIterationsPerFrame=1)
    assert sha256_f32(sim.computeAccel()) == meta["force_sha256"][str(n)]
    if str(n) in meta["step1_sha256"]:
        sim.stepSim()
        assert sha256_f32(_state(sim)) == meta["step1_sha256"][str(n)]
    sim.close()


def test_generic_kernel_agrees_at_full_size(nb, golden_dir):
    """the independently written predicated scalar kernel against the same 1M golden"""
    meta = json.load(open(os.path.join(golden_dir, "golden_meta.json")))
    n = 1048576
    sim = _mk(nb, n)
    sim.setKernel(nb.KERNEL_GENERIC)
    assert sha256_f32(sim.computeAccel()) == meta["force_sha256"][str(n)]
    sim.close()


def test_device_pointer_entry_matches_handle_path(nb):
    """nbody_launch_step_device on caller-owned device memory (torch tensors) == nbody_step."""
    import ctypes

    import torch
    n = 8192
    st = nb.generate_disk_galaxy(n)
    sim = _mk(nb, n, simIterationsPerFrame=1)
    sim.stepSim()
    want = _state(sim)
    sim.close()
    dev = torch.device("cuda:0")
    pos4 = torch.tensor(np.stack([st[0], st[1], st[2], np.ones(n, np.float32)], 1), device=dev).contiguous()
    vel4 = torch.tensor(np.stack([st[3], st[4], st[5], np.zeros(n, np.float32)], 1), device=dev).contiguous()
    nxt = torch.empty_like(pos4)
    p = nb.SimParam(numParticles=n, simIterationsPerFrame=1).to_c()
    lib = nb.load_library()
    stream = torch.cuda.current_stream().cuda_stream
    rc = lib.nbody_launch_step_device(ctypes.byref(p), pos4.data_ptr(), vel4.data_ptr(), nxt.data_ptr(),
                                      0, n, nb.KERNEL_AUTO, 0, stream)
    assert rc == 0, lib.nbody_last_error()
    torch.cuda.synchronize()
    got_p, got_v = nxt.cpu().numpy(), vel4.cpu().numpy()
    _assert_bits([got_p[:, 0], got_p[:, 1], got_p[:, 2], got_v[:, 0], got_v[:, 1], got_v[:, 2]], want,
                 "device-pointer entry")
    assert np.all(got_p[:, 3] == 1.0)


def test_f4_readback_layout(nb):
    """AoS float4(x,y,z,1) is what RendererGL::setParticleData builds (src/renderer_gl.cpp:156-172)."""
    n = 5000
    sim = _mk(nb, n, simIterationsPerFrame=2)
    sim.stepSim()
    st = _state(sim)
    p4, v4 = sim.readPosF4(), sim.readVelF4()
    sim.close()
    _assert_bits([p4[:, 0], p4[:, 1], p4[:, 2], v4[:, 0], v4[:, 1], v4[:, 2]], st, "f4 read-back")
    assert np.all(p4[:, 3] == 1.0)


def test_step_timing_and_launch_count(nb):
    n = 25600
    sim = _mk(nb, n, simIterationsPerFrame=10)
    c0 = sim.launchCount()
    sim.stepSim()
    assert sim.launchCount() - c0 == 10  # one fused force+integrate launch per iteration
    assert sim.getLastStepTime() > 0 and sim.getLastStepDeviceTime() > 0
    assert "B200" in sim.getDeviceName() or len(sim.getDeviceName()) > 0
    sim.close()


def test_cxx_dropin_binaries(nb):
    """The reference's own main (src/nbody.cpp, compiled against OUR simulator.cuh) and our
    headless driver: same argv contract (src/sim_param.cpp:40-67) and the same stdout line
    (src/nbody.cpp:120-122), printed after the two warm-up frames."""
    import os
    import re
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    line = re.compile(r"^At step (\d+) kernel time is \S+ and mean is \S+ and stddev is: \S+$")
    ran = 0
    for exe in ("nbody_b200", "nbody_refmain"):
        path = os.path.join(root, "cuda-to-sycl-nbody_b200", "bin", exe)
        if not os.path.exists(path):
            continue
        r = subprocess.run([path, "8", "3", "0.999", "0.001", "1.0e-3", "2.0", "6", "128", "BRANCH"],
                           capture_output=True, text=True, timeout=120)
        assert r.returncode == 0, r.stderr
        rows = [l for l in r.stdout.splitlines() if l.strip()]
        assert [int(line.match(l).group(1)) for l in rows] == [3, 4, 5, 6], rows
        ran += 1
    assert ran >= 1, "no drop-in binary built"
    # the STATE behind those lines: NBODY_DUMP_HASH=1 makes the C++ class print the FNV-1a-64 of its host
    # vectors when it is destroyed; it must equal the reference simulator's after the same frames
    # (covers cxx/simulator.cpp: constructor upload, lazy refreshHost, registered host vectors)
    import oracle_lib
    import refsim
    if refsim.available():
        r = refsim.RefSimulator(8 * 256, G=2.0, dt=0.001, iters=3, damping=0.999, eps=1.0e-3, gw=128)
        for _ in range(6):
            r.step()
        want = oracle_lib.Oracle().fnv1a64(r.state())
        r.close()
        for exe in ("nbody_b200", "nbody_refmain"):
            path = os.path.join(root, "cuda-to-sycl-nbody_b200", "bin", exe)
            if not os.path.exists(path):
                continue
            env = dict(os.environ, NBODY_DUMP_HASH="1")
            out = subprocess.run([path, "8", "3", "0.999", "0.001", "1.0e-3", "2.0", "6", "128", "BRANCH"],
                                 capture_output=True, text=True, timeout=120, env=env)
            m = re.search(r"final state fnv1a64 ([0-9a-f]{16})", out.stderr)
            assert m and m.group(1) == want, (exe, out.stderr, want)
    bad = subprocess.run([os.path.join(root, "cuda-to-sycl-nbody_b200", "bin", "nbody_b200"), "8", "1", "1", "1", "1",
                          "1", "1", "64", "NOPE"], capture_output=True, text=True)
    assert bad.returncode != 0  # std::invalid_argument, as the reference (src/sim_param.cpp:36)


# ---------------------------------------------------------------------------------------------
# extension: per-body masses in float4.w (the reference is unit-mass)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,eps", [(20000, 1e-7), (300000, 1e-7), (3000, 0.0)])
def test_masses_exact_properties_and_oracle(nb, oracle, n, eps):
    """No reference exists for masses, so: (a) m = 1 is bit-identical to the reference path,
    (b) m = 2 doubles every force bit-exactly (power-of-two scaling commutes with every rounding),
    (c) massless bodies add exactly nothing, (d) random masses agree with the CPU oracle's massful
    sum to the MUFU-vs-sqrt tolerance, (e) masses survive set_state and stepping."""
    sim = _mk(nb, n, distEps=eps, simIterationsPerFrame=2)
    base = sim.computeAccel()
    st = _state(sim)
    sim.setMass(np.ones(n, np.float32))
    assert "mass" in sim.kernelName()
    _assert_bits(sim.computeAccel(), base, "m=1")
    sim.setMass(np.full(n, 2.0, np.float32))
    _assert_bits(sim.computeAccel(), [2.0 * b for b in base], "m=2")
    rng = np.random.default_rng(n)
    m = rng.uniform(0.1, 3.0, n).astype(np.float32)
    m[::3] = 0.0
    sim.setMass(m)
    got = sim.computeAccel()
    sub = min(n, 2000)
    want = oracle.accel_mass(st[0], st[1], st[2], m, eps, 0, sub)
    e = rel_err([g[:sub] for g in got], want)
    assert np.median(e) <= 2e-6 and e.max() <= 5e-4, (np.median(e), e.max())
    # massless bodies contribute exactly nothing: drop them from the j set by moving them far away
    sim.setState(*st)                       # masses stay with their bodies across set_state
    assert np.array_equal(sim.readPosF4()[:, 3], m)
    _assert_bits(sim.computeAccel(), got, "after set_state")
    sim.stepSim()
    assert np.array_equal(sim.readPosF4()[:, 3], m)      # and across steps
    sim.setMass(None)
    assert "mass" not in sim.kernelName()
    sim.setState(*st)
    _assert_bits(sim.computeAccel(), base, "unit masses restored")
    sim.close()


def test_checkpoint_resume_is_bit_identical(nb, tmp_path):
    """state dump / restore: 3+3 iterations through a file == 6 iterations straight"""
    n = 30000
    a = _mk(nb, n, simIterationsPerFrame=3)
    m = np.linspace(0.5, 1.5, n).astype(np.float32)
    a.setMass(m)
    a.stepSim()
    path = str(tmp_path / "ckpt.nbb")
    a.saveState(path)
    a.stepSim()
    want = _state(a)
    a.close()
    raw = open(path, "rb").read()
    assert raw[:6] == b"NBB200" and len(raw) == 48 + 7 * 4 * n
    assert np.frombuffer(raw, np.uint64, 1, 8)[0] == n
    b = _mk(nb, n, simIterationsPerFrame=3)
    b.loadState(path)
    assert "mass" in b.kernelName()
    b.stepSim()
    _assert_bits(_state(b), want, "resumed run")
    c = _mk(nb, n + 256)
    with pytest.raises(nb.NBodyError, match="bodies"):
        c.loadState(path)
    b.close()
    c.close()


@pytest.mark.parametrize("n,cfg", [(32003, None), (40003, None), (57001, None), (65003, None), (113003, None),
                                   (114003, None), (200001, None), (244003, None), (244403, None), (400003, None),
                                   (458003, None), (459003, None),
                                   (58003, "2,32,4"), (200001, "4,32,4"), (200001, "6,32,4"), (400003, "2,32,4"),
                                   (400003, "4,32,4"), (131072, "6,32,4")])
def test_register_blocking_switch_points_ragged(nb, ref, n, cfg, monkeypatch):
    """ragged sizes (not a multiple of anything) around the scalar -> R = 2 -> R = 4 -> R = 6 switch points of AUTO,
    and every register-blocking factor forced at sizes AUTO would not use it for: exercises the ragged last group
    and the generated tile body (tools/sass_gen.py) of every instantiation against the reference kernel"""
    if cfg:
        monkeypatch.setenv("NBODY_KERNEL_CONFIG", cfg)
    fx, fy, fz, _ = ref.reference_forces(n)
    sim = _mk(nb, n)
    name = sim.kernelName()
    if cfg:
        assert f"wseg_f32x2_r{cfg[0]}" in name, name
    else:
        assert "wseg_f32x2_r" in name or "wsmall_scalar_r1" in name, name
    _assert_bits(sim.computeAccel(), [fx, fy, fz], f"forces N={n} ({name})")
    sim.close()


def test_post_link_step_is_reported(nb):
    """the library says what the post-link step did to the kernel it runs: the production kernels, unit-mass and
    per-body-mass alike, carry a generated tile body (+sass-gen)"""
    sim = _mk(nb, 400003)
    assert sim.kernelName().endswith("+sass-gen"), sim.kernelName()
    sim.setMass(np.full(400003, 1.0, np.float32))
    assert "mass" in sim.kernelName() and sim.kernelName().endswith("+sass-gen"), sim.kernelName()
    sim.close()


@pytest.mark.parametrize("segs", ["1", "3", "7", "64"])
def test_segment_count_does_not_change_a_bit(nb, golden_dir, segs, monkeypatch):
    """the j-segmented hand-off keeps one FP32 chain per body whatever the number of segments"""
    monkeypatch.setenv("NBODY_SEGS", segs)
    meta = json.load(open(os.path.join(golden_dir, "golden_meta.json")))
    import oracle_lib
    o = oracle_lib.Oracle()
    sim = _mk(nb, 262144)
    assert o.fnv1a64(sim.computeAccel()) == meta["force"]["262144"]["fnv1a64"]
    sim.close()
    sim = _mk(nb, 25600, simIterationsPerFrame=10)
    sim.stepSim()
    assert o.fnv1a64(_state(sim)) == meta["step10"]["25600"]["fnv1a64"]
    sim.close()


def test_two_handles_interleaved(nb):
    """two simulators alive in one process, stepped alternately (per-device launch-plan caches,
    per-handle hand-off epochs)"""
    a = _mk(nb, 30000, simIterationsPerFrame=2)
    b = _mk(nb, 50000, simIterationsPerFrame=3)
    ra = _mk(nb, 30000, simIterationsPerFrame=4)
    for _ in range(2):
        a.stepSim()
        b.stepSim()
    ra.stepSim()
    _assert_bits(_state(a), _state(ra), "interleaved handle")
    for s in (a, b, ra):
        s.close()


def test_bad_gpu_count_is_rejected(nb):
    with pytest.raises(nb.NBodyError, match="not present"):
        nb.DiskGalaxySimulator(nb.SimParam(numParticles=1024), n_gpus=nb.device_count() + 1)


@pytest.mark.parametrize("cfg", ["6,32,5", "4,32,5", "4,32,3", "4,256,1", "4,128,2"])
def test_comparison_variants_stay_bit_exact(nb, vlib, golden_dir, cfg, monkeypatch):
    """the comparison kernels of the VARIANTS library (TMA/cp.async.bulk-staged, unsegmented
    warp-streaming, CTA-tiled packed and scalar) compute the same bits as the production kernel and
    the reference; the product library refuses the ones it does not contain"""
    monkeypatch.setenv("NBODY_KERNEL_CONFIG", cfg)
    meta = json.load(open(os.path.join(golden_dir, "golden_meta.json")))
    sim = _mk(nb, 262144, lib=vlib)
    if cfg.endswith(",1"):
        sim.setKernel(nb.KERNEL_PACKED)
    if cfg.endswith(",2"):
        sim.setKernel(nb.KERNEL_SCALAR)
    assert sha256_f32(sim.computeAccel()) == meta["force_sha256"]["262144"], sim.kernelName()
    sim.stepSim()
    sim.close()
    if cfg.endswith(",1"):
        prod = _mk(nb, 4096)
        with pytest.raises(nb.NBodyError, match="VARIANTS"):
            prod.setKernel(nb.KERNEL_PACKED)
        prod.close()


def test_many_launches_keep_ticket_and_epoch_numbering(nb, golden_dir, monkeypatch):
    """hand-off words and the ticket counter run on across launches (and across kernels of different
    group counts on one handle): 40 segmented launches, then the 10-step golden must still hold"""
    monkeypatch.setenv("NBODY_SEGS", "5")
    g = np.load(os.path.join(golden_dir, "step10_n2048.npz"))
    sim = _mk(nb, 2048, simIterationsPerFrame=10)
    init = _state(sim)
    for _ in range(4):
        sim.stepSim()
    a1 = sim.computeAccel()
    a2 = sim.computeAccel()
    _assert_bits(a1, a2, "repeated accel pass")
    sim.setState(*init)
    sim.stepSim()
    _assert_bits(_state(sim), [g[k] for k in ("x", "y", "z", "vx", "vy", "vz")], "golden after 40 earlier launches")
    sim.close()


def test_read_state_read_local_and_registered_host_memory(nb):
    """nbody_read_state == read_pos + read_vel; nbody_read_local == the owned range of it; page-locked
    (nbody_host_register) destination buffers give the same bits"""
    import ctypes
    n = 70000
    sim = _mk(nb, n, simIterationsPerFrame=2)
    sim.stepSim()
    lib = nb.load_library()
    sep = [np.empty(n, np.float32) for _ in range(6)]
    nb._check(lib, lib.nbody_read_pos(sim._h, *[nb._ptr(a) for a in sep[:3]]), "read_pos")
    nb._check(lib, lib.nbody_read_vel(sim._h, *[nb._ptr(a) for a in sep[3:]]), "read_vel")
    one = [np.empty(n, np.float32) for _ in range(6)]
    for a in one:
        assert lib.nbody_host_register(a.ctypes.data_as(ctypes.c_void_p), a.nbytes) == 0, lib.nbody_last_error()
    sim.readInto(*one)
    _assert_bits(one, sep, "read_state into registered memory")
    assert sim.localRange() == (0, n)
    loc = [np.zeros(n, np.float32) for _ in range(6)]
    sim.readLocalInto(*loc)
    _assert_bits(loc, sep, "read_local")
    for a in one:
        assert lib.nbody_host_unregister(a.ctypes.data_as(ctypes.c_void_p)) == 0
    sim.close()


# ---------------------------------------------------------------------------------------------
# accumulator relay (AUTO's choice for the smallest shards): W warps of a CTA serve the same 32 bodies
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,cfg", [(2048, None), (12800, None), (14208, None), (14209, None), (18944, None),
                                   (4099, "16,64,7"), (4099, "32,64,7"), (4099, "16,256,7"), (20011, "32,128,7"),
                                   (20011, "16,128,7"), (33, "32,128,7"), (17, "16,64,7")])
def test_relay_kernel_forces_and_steps_vs_reference(nb, vlib, ref, n, cfg, monkeypatch):
    """every shape of force_wrelay_kernel (warps per CTA x bodies per tile) at ragged sizes: forces and the state after
    2 x 5 iterations are bit-equal to the unmodified reference kernel's (the sums travel between warps through shared
    memory, so the integrate epilogue runs in whichever warp took the last tile)"""
    if cfg:
        monkeypatch.setenv("NBODY_KERNEL_CONFIG", cfg)
    fx, fy, fz, _ = ref.reference_forces(n)
    # the 2- and 8-warp shapes are comparison kernels of the VARIANTS library; AUTO's two shapes are in the product
    sim = _mk(nb, n, lib=vlib if cfg and "128" not in cfg else None, simIterationsPerFrame=5)
    name = sim.kernelName()
    assert "wrelay_scalar" in name, name
    if cfg:
        r, b, _f = cfg.split(",")
        assert f"_r{r}_b{b}_" in name, name
    _assert_bits(sim.computeAccel(), [fx, fy, fz], f"forces N={n} ({name})")
    rs = ref.RefSimulator(n, iters=5)
    for _ in range(2):
        rs.step()
        sim.stepSim()
    _assert_bits(_state(sim), rs.state(), f"state after 10 iterations N={n} ({name})")
    rs.close()
    sim.close()


@pytest.mark.parametrize("n", [1000, 12800, 18944])
def test_relay_kernel_with_masses_equals_scalar_kernel(nb, n, monkeypatch):
    """per-body masses through the relay kernel: bit-equal to the one-body-per-lane scalar kernel (which
    test_masses_exact_properties_and_oracle pins), forces and a 6-iteration state"""
    m = np.random.default_rng(n).uniform(0.25, 4.0, n).astype(np.float32)
    out = {}
    for label, cfg in (("relay", None), ("scalar", "1,32,6")):
        if cfg:
            monkeypatch.setenv("NBODY_KERNEL_CONFIG", cfg)
        sim = _mk(nb, n, simIterationsPerFrame=3)
        sim.setMass(m)
        name = sim.kernelName()
        assert ("wrelay_scalar" if label == "relay" else "wsmall_scalar_r1") in name and "mass" in name, name
        f = sim.computeAccel()
        sim.stepSim()
        sim.stepSim()
        out[label] = (f, _state(sim))
        sim.close()
    _assert_bits(out["relay"][0], out["scalar"][0], f"forces with masses N={n}")
    _assert_bits(out["relay"][1], out["scalar"][1], f"state with masses N={n}")
