"""CPU-side checks of the C ABI: the library loads, exports every symbol include/nbody_b200.h
declares, host-only entry points work, and compute entry points FAIL LOUDLY without a GPU
(no CPU fallback exists)."""
from __future__ import annotations

import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "nbody_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nbody_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_all_exported(nb):
    lib = nb.load_library()
    declared = _declared_symbols()
    assert len(declared) >= 25
    for s in declared:
        assert hasattr(lib, s), f"{s} declared in include/nbody_b200.h but not exported"
    assert sorted(nb.ABI_SYMBOLS) == declared, "python binding list out of sync with the header"
    assert lib.nbody_abi_version() == nb.ABI_VERSION == 2


def test_library_has_sm100a_code_only(nb):
    out = subprocess.run(["cuobjdump", "-lelf", nb.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_no_oracle_or_reference_in_product(nb):
    """the product must not link or reference the checkers"""
    out = subprocess.run(["ldd", nb.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out and "nbody_ref" not in out
    for dirpath, _, files in os.walk(os.path.join(ROOT, "cuda-to-sycl-nbody_b200")):
        if os.sep + "lib" in dirpath or os.sep + "bin" in dirpath or "__pycache__" in dirpath:
            continue
        for f in files:
            if f.endswith((".cu", ".cuh", ".cpp", ".hpp", ".py")) or f == "Makefile":
                t = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle_" not in t and "libnbody_oracle" not in t and "libnbody_ref" not in t, f


def test_default_params_match_reference_defaults(nb):
    lib = nb.load_library()
    p = nb.Params()
    lib.nbody_default_params(ctypes.byref(p))
    # reference src/sim_param.cpp:12-22
    assert (p.G, p.dt, p.num_particles, p.iters_per_frame) == (2.0, np.float32(0.005), 12800, 4)
    assert (p.damping, p.dist_eps, p.gw_size, p.calc_method) == (np.float32(0.999998), np.float32(1e-7), 64, 0)
    d = nb.SimParam()
    assert (d.G, d.numParticles, d.simIterationsPerFrame, d.gwSize) == (2.0, 12800, 4, 64)


def test_generator_matches_oracle_and_golden(nb, oracle, golden_dir):
    import json
    meta = json.load(open(os.path.join(golden_dir, "golden_meta.json")))
    for n in (256, 2048, 12800, 25600):
        mine = nb.generate_disk_galaxy(n)
        port = oracle.disk_galaxy(n)
        for a, b in zip(mine, port):
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
        assert oracle.fnv1a64(mine) == meta["init"][str(n)]["fnv1a64"]


def test_generator_edge_sizes(nb):
    assert all(len(a) == 0 for a in nb.generate_disk_galaxy(0))
    one = nb.generate_disk_galaxy(1)
    big = nb.generate_disk_galaxy(1000)
    # body 0's x,y do not depend on N; z does (z draws come after all angle/radius draws)
    assert one[0][0] == big[0][0] and one[1][0] == big[1][0]
    r = np.hypot(big[0], big[1])
    assert r.max() < 100.0 and 0.0 <= big[2].min() and big[2].max() < 4.0
    assert np.all(big[5] == 0.0)
    # tangential: v . r == 0 up to rounding, |v| = sqrt(2 r)
    dot = big[3] * big[0] + big[4] * big[1]
    assert np.abs(dot).max() < 1e-3
    assert np.allclose(np.hypot(big[3], big[4]), np.sqrt(2 * r), rtol=1e-5)


@pytest.mark.parametrize("n,world", [(4194304, 8), (1048576, 2), (1000, 3), (100, 8), (25600, 4), (1, 2)])
def test_shard_plan_partitions_bodies(nb, n, world):
    shards = [nb.plan_shard(n, world, r) for r in range(world)]
    pos = 0
    for b, c in shards:
        assert b == pos
        pos += c
    assert pos == n
    for b, c in shards[:-1]:
        if c:
            assert b % 128 == 0
    sizes = [c for _, c in shards if c]
    assert max(sizes) - min(sizes) <= max(128, sizes[0])  # balanced up to the ragged tail


def test_compute_fails_loudly_without_gpu(nb):
    if nb.device_count() > 0:
        pytest.skip("a GPU is visible: this test covers the CPU-only box")
    with pytest.raises(nb.NBodyError, match="no CUDA device"):
        nb.DiskGalaxySimulator(nb.SimParam(numParticles=256))
    lib = nb.load_library()
    h = ctypes.c_void_p()
    p = nb.SimParam(numParticles=256).to_c()
    assert lib.nbody_create(ctypes.byref(p), 1, ctypes.byref(h)) == 10002  # NBODY_E_NOGPU
    assert not h.value


def test_bad_arguments_are_rejected(nb):
    lib = nb.load_library()
    h = ctypes.c_void_p()
    p = nb.SimParam(numParticles=0).to_c()
    assert lib.nbody_create(ctypes.byref(p), 1, ctypes.byref(h)) != 0
    assert lib.nbody_step(None) != 0
    assert b"null" in lib.nbody_last_error()
    b, c = ctypes.c_uint64(), ctypes.c_uint64()
    assert lib.nbody_plan_shard(100, 2, 5, ctypes.byref(b), ctypes.byref(c)) != 0


def test_headless_driver_and_reference_main_fail_fast_without_gpu(nb):
    """the C++ drop-in keeps the reference's print-and-exit convention (src/simulator.cuh:22-31)"""
    if nb.device_count() > 0:
        pytest.skip("GPU box")
    exe = os.path.join(ROOT, "cuda-to-sycl-nbody_b200", "bin", "nbody_b200")
    if not os.path.exists(exe):
        pytest.skip("driver not built")
    r = subprocess.run([exe, "4", "1", "0.999", "0.001", "1e-3", "2.0", "3"], capture_output=True, text=True)
    assert r.returncode != 0 and "GPUassert:" in r.stderr


def test_shard_plan_properties_hypothesis(nb):
    """property test: shards tile [0, n) in rank order, 128-aligned except the ragged tail"""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=200, deadline=None)
    @given(n=st.integers(min_value=1, max_value=(1 << 31) - 1), world=st.integers(min_value=1, max_value=16))
    def check(n, world):
        pos = 0
        for r in range(world):
            b, c = nb.plan_shard(n, world, r)
            assert b == pos and (c == 0 or b % 128 == 0)
            pos += c
        assert pos == n

    check()


def test_auto_switch_points(nb):
    """AUTO's kernel choice per shard size on a 148-SM device (host-only nbody_describe_auto): accumulator relay for the
    smallest shards (<= 3 / 4 groups of 32 bodies per SM), the quantisation cost model up to 768 bodies per SM, then
    R = 2 / 4 / 6 of the production kernel; PREDICATED and eps = 0 always take the generic scalar kernel."""
    d = lambda n, **k: nb.describe_auto(n, 148, **k)
    assert d(1) == d(2048) == d(12800) == d(148 * 96) == "wrelay_scalar_r32_b128_nopred"
    assert d(148 * 96 + 1) == d(18944) == "wrelay_scalar_r16_b128_nopred"
    assert d(12800, has_mass=True) == "wrelay_scalar_r32_b128_nopred_mass"
    for n in (18945, 25600, 40000, 57720, 64000, 100000):
        assert d(n) in ("wsmall_scalar_r1_b32_nopred", "wseg_f32x2_r2_b32_nopred", "wseg_f32x2_r4_b32_nopred"), (n, d(n))
    assert d(131072) == "wseg_f32x2_r2_b32_nopred"
    assert d(262144) == "wseg_f32x2_r4_b32_nopred"
    assert d(1048576) == d(4194304 // 8) == "wseg_f32x2_r6_b32_nopred"
    assert d(1048576, has_mass=True) == "wseg_f32x2_r6_b32_nopred_mass"
    # a shard of an 8-GPU run is chosen by ITS size, not by N
    assert d(102400 // 8) == "wrelay_scalar_r32_b128_nopred"
    # the faithful paths never leave the generic kernel
    assert d(12800, params=nb.SimParam(calcMethod=nb.CALC_PREDICATED)) == "generic_scalar_r1_b32_predicated"
    assert d(12800, params=nb.SimParam(distEps=0.0)) == "generic_scalar_r1_b32_branch"
    # other machine sizes scale the switch points with the SM count
    assert nb.describe_auto(132 * 96, 132) == "wrelay_scalar_r32_b128_nopred"
    assert nb.describe_auto(132 * 96 + 1, 132) == "wrelay_scalar_r16_b128_nopred"
    with pytest.raises(nb.NBodyError):
        nb.describe_auto(0, 148)


def test_auto_choice_properties_hypothesis(nb):
    """AUTO over arbitrary shard sizes and SM counts: always one of the shipped kernels; the relay exactly up to 128
    bodies per SM; above 768 bodies per SM the register-blocking factor never shrinks as the shard grows"""
    from hypothesis import given, settings, strategies as st

    names = {"wrelay_scalar_r32_b128_nopred": 0, "wrelay_scalar_r16_b128_nopred": 0, "wsmall_scalar_r1_b32_nopred": 1,
             "wseg_f32x2_r2_b32_nopred": 2, "wseg_f32x2_r4_b32_nopred": 4, "wseg_f32x2_r6_b32_nopred": 6}

    @settings(max_examples=300, deadline=None)
    @given(st.integers(1, 1 << 24), st.integers(1, 1 << 24), st.integers(64, 192))
    def prop(n1, n2, sms):
        a, b = sorted((n1, n2))
        ka, kb = nb.describe_auto(a, sms), nb.describe_auto(b, sms)
        assert ka in names and kb in names
        assert ("wrelay" in ka) == (a <= 128 * sms)
        assert ("_r32_b128" in ka) == (a <= 96 * sms)
        if a >= 768 * sms:
            assert names[ka] >= 2 and names[kb] >= names[ka]
        assert nb.describe_auto(a, sms, has_mass=True) == ka + "_mass"

    prop()


def test_relay_turn_protocol_model():
    """model of the accumulator relay's hand-off (csrc/nbody_body.cuh, cta_relay_scalar): warp w takes tiles w, w + W, ...;
    before tile t > 0 it waits on ITS barrier with parity (number of its earlier waits) & 1; after tile t it arrives on
    the barrier of warp (w + 1) % W unless t is the last tile.  Under any interleaving of the warps the tiles are
    accumulated in ascending order, every wait is eventually satisfied, and a warp never finds more than one completed
    phase it has not waited for (the parity bit would alias)."""
    import random

    for W in (2, 4, 8):
        for ntiles in list(range(0, 20)) + [37, 64, 101]:
            rng = random.Random(W * 1000 + ntiles)
            phase = [0] * W            # completed phases per barrier
            nxt = [w for w in range(W)]  # next tile of each warp
            waits = [0] * W
            order = []
            stuck = 0
            while any(t < ntiles for t in nxt):
                w = rng.randrange(W)
                t = nxt[w]
                if t >= ntiles:
                    continue
                if t > 0:
                    if phase[w] <= waits[w]:      # try_wait(parity = waits & 1) not yet satisfied
                        stuck += 1
                        assert stuck < 100000, "deadlock"
                        continue
                    assert phase[w] == waits[w] + 1, "a phase was skipped"
                    waits[w] += 1
                stuck = 0
                order.append(t)
                if t + 1 < ntiles:
                    d = (w + 1) % W
                    phase[d] += 1                 # all 32 lanes arrive: the phase completes
                nxt[w] = t + W
            assert order == list(range(ntiles))
