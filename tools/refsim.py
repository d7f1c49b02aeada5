"""ctypes binding of oracle/_ref/libnbody_ref.so -- the UNMODIFIED reference CUDA simulator.

Test infrastructure (see oracle/ref_shim.cu).  Used by tests/, tests/golden/make_golden.py,
tools/ and bench.py's comparison legs; never by the product.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_LIB = os.path.join(ROOT, "oracle", "_ref", "libnbody_ref.so")
# the same reference with src/simulator.cu:209 changed to (i != id): oracle of PREDICATED_FIXED
REF_FIXED_LIB = os.path.join(ROOT, "oracle", "_ref", "libnbody_ref_fixed.so")
_fp = ctypes.POINTER(ctypes.c_float)
_libs = {}


def available(which: str = "ref") -> bool:
    return os.path.exists(REF_FIXED_LIB if which == "fixed" else REF_LIB)


def load(which: str = "ref"):
    if which not in _libs:
        lib = ctypes.CDLL(REF_FIXED_LIB if which == "fixed" else REF_LIB)
        lib.ref_create.restype = ctypes.c_void_p
        lib.ref_create.argtypes = [ctypes.c_float, ctypes.c_float, ctypes.c_ulonglong, ctypes.c_int,
                                   ctypes.c_float, ctypes.c_float, ctypes.c_int, ctypes.c_int]
        lib.ref_destroy.argtypes = [ctypes.c_void_p]
        lib.ref_set_state.argtypes = [ctypes.c_void_p] + [_fp] * 6
        lib.ref_get_state.argtypes = [ctypes.c_void_p] + [_fp] * 6
        lib.ref_step.argtypes = [ctypes.c_void_p]
        lib.ref_last_step_ms.argtypes = [ctypes.c_void_p]
        lib.ref_last_step_ms.restype = ctypes.c_float
        lib.ref_device_name.argtypes = [ctypes.c_void_p]
        lib.ref_device_name.restype = ctypes.c_char_p
        lib.ref_time_kernel.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
        lib.ref_time_kernel.restype = ctypes.c_float
        _libs[which] = lib
    return _libs[which]


def _p(a):
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_fp)


class RefSimulator:
    """The reference's DiskGalaxySimulator (src/simulator.cuh:129-160), driven through the shim."""

    def __init__(self, n, G=2.0, dt=0.005, iters=4, damping=0.999998, eps=1.0e-7, gw=64, calc=0, lib="ref"):
        self.lib = load(lib)
        self.n = n
        self.h = self.lib.ref_create(G, dt, n, iters, damping, eps, gw, calc)

    def state(self):
        a = [np.empty(self.n, np.float32) for _ in range(6)]
        self.lib.ref_get_state(self.h, *[_p(v) for v in a])
        return a

    def set_state(self, *arrs):
        arrs = [np.ascontiguousarray(a, np.float32) for a in arrs]
        self.lib.ref_set_state(self.h, *[_p(v) for v in arrs])

    def step(self):
        self.lib.ref_step(self.h)

    def last_step_ms(self):
        return float(self.lib.ref_last_step_ms(self.h))

    def time_kernel(self, gw, launches):
        return float(self.lib.ref_time_kernel(self.h, gw, launches))

    def device_name(self):
        return self.lib.ref_device_name(self.h).decode()

    def close(self):
        if self.h:
            self.lib.ref_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def reference_forces(n, eps=1.0e-7, state=None, calc=0, lib="ref"):
    """Raw force sums of the reference kernel via the damping=0, dt=1, G=1 trick:
    v' = fma(F*1, 1, v*0) = F exactly.  Returns (fx, fy, fz, initial_state)."""
    sim = RefSimulator(n, G=1.0, dt=1.0, iters=1, damping=0.0, eps=eps, calc=calc, lib=lib)
    if state is not None:
        sim.set_state(*state)
    init = sim.state()
    sim.step()
    s = sim.state()
    sim.close()
    return s[3], s[4], s[5], init
