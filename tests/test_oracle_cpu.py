"""The CPU oracle (oracle/nbody_oracle.c) against the golden fixtures the UNMODIFIED reference
produced on a B200 (tests/golden/make_golden.py).  Generator: bit-exact.  Forces / stepped
states: stated tolerance, because the reference kernel evaluates rsqrt with MUFU.RSQ."""
from __future__ import annotations

import json
import os

import numpy as np

from oracle_lib import rel_err


def _meta(golden_dir):
    return json.load(open(os.path.join(golden_dir, "golden_meta.json")))


def test_generator_bit_exact_vs_reference(oracle, golden_dir):
    meta = _meta(golden_dir)
    for n_s, rec in meta["init"].items():
        st = oracle.disk_galaxy(int(n_s))
        assert oracle.fnv1a64(st) == rec["fnv1a64"], f"N={n_s}"
        for arr, bits in zip(st, rec["first4_bits"]):
            assert [f"{int(v):08x}" for v in arr[:4].view(np.uint32)] == bits


def test_python_and_c_hash_agree(oracle):
    """make_golden.py hashes in pure python, the tests in C: same function"""
    st = oracle.disk_galaxy(64)
    inter = np.stack(st, axis=1).reshape(-1).view(np.uint8)
    h = 1469598103934665603
    for b in inter.tobytes():
        h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    assert oracle.fnv1a64(st) == f"{h:016x}"


def test_forces_vs_reference_golden_n2048(oracle, golden_dir):
    """tolerance: per-body relative error median <= 1e-6, max <= 1e-4 (CPU 1/sqrtf vs MUFU.RSQ)"""
    g = np.load(os.path.join(golden_dir, "force_n2048.npz"))
    st = oracle.disk_galaxy(2048)
    a = oracle.accel(st[0], st[1], st[2], 1.0e-7)
    e = rel_err(a, [g["fx"], g["fy"], g["fz"]])
    assert np.median(e) <= 1e-6 and e.max() <= 1e-4, (np.median(e), e.max())


def test_step10_vs_reference_golden_n2048(oracle, golden_dir):
    """positions after 10 steps: north_star tolerance 1e-4 relative"""
    g = np.load(os.path.join(golden_dir, "step10_n2048.npz"))
    st = oracle.step(oracle.disk_galaxy(2048), iters=10)
    e = rel_err(st[:3], [g["x"], g["y"], g["z"]])
    assert e.max() <= 1e-4, e.max()
    ev = rel_err(st[3:], [g["vx"], g["vy"], g["vz"]])
    assert np.median(ev) <= 1e-5


def test_cloud_with_coincident_pair_vs_reference_golden(oracle, golden_dir):
    g = np.load(os.path.join(golden_dir, "cloud_n1000.npz"))
    a = oracle.accel(g["x"], g["y"], g["z"], float(g["eps"]))
    e = rel_err(a, [g["fx"], g["fy"], g["fz"]])
    assert np.median(e) <= 1e-6 and e.max() <= 1e-4, (np.median(e), e.max())
    assert np.isfinite(np.stack(a)).all()


def test_predicated_as_shipped_is_force_free(oracle, golden_dir):
    """PREDICATED multiplies by (i == id): zero force (src/simulator.cu:209).  With F = 0 the update
    is exact arithmetic (v*damping, fma(v', dt, x)), so the CPU matches the GPU golden bit-for-bit."""
    g = np.load(os.path.join(golden_dir, "predicated_n1024.npz"))
    init = oracle.disk_galaxy(1024)
    a = oracle.accel(init[0], init[1], init[2], 1.0e-7, method=1)
    assert all(np.all(c == 0.0) for c in a)
    st = oracle.step(init, method=1, iters=1)
    for got, name in zip(st, ("x", "y", "z", "vx", "vy", "vz")):
        assert np.array_equal(got, g[name]), name


def test_oracle_edge_cases(oracle):
    # single body: no force, pure drift
    one = [np.array([v], np.float32) for v in (1.0, 2.0, 3.0, 0.5, 0.0, -0.5)]
    a = oracle.accel(one[0], one[1], one[2], 1e-7)
    assert all(c[0] == 0.0 for c in a)
    # two bodies: equal and opposite, along the separation
    x = np.array([0.0, 3.0], np.float32); y = np.array([0.0, 4.0], np.float32); z = np.zeros(2, np.float32)
    ax, ay, az = oracle.accel(x, y, z, 0.0)
    assert ax[0] == -ax[1] and ay[0] == -ay[1] and az[0] == 0.0
    assert np.isclose(np.hypot(ax[0], ay[0]), 1.0 / 25.0, rtol=1e-6)
    # zero softening with the BRANCH skip stays finite
    st = oracle.disk_galaxy(300)
    a = oracle.accel(st[0], st[1], st[2], 0.0)
    assert np.isfinite(np.stack(a)).all()
    # sub-range evaluation equals the slice of the full evaluation (ragged, not lane-aligned)
    full = oracle.accel(st[0], st[1], st[2], 1e-7)
    part = oracle.accel(st[0], st[1], st[2], 1e-7, i_begin=37, i_end=250)
    for f, p in zip(full, part):
        assert np.array_equal(f[37:250], p)


def test_oracle_fp32_vs_fp64_truth(oracle):
    """error budget: the reference-order FP32 sum vs FP64 (~2e-6 median at 25600)"""
    st = oracle.disk_galaxy(4096)
    a32 = oracle.accel(st[0], st[1], st[2], 1e-7, i_begin=0, i_end=256)
    a64 = oracle.accel_f64(st[0], st[1], st[2], 1e-7, 0, 256)
    e = rel_err(a32, a64)
    assert np.median(e) < 1e-5
