"""ctypes binding of the CPU oracle (oracle/_build/libnbody_oracle.so).  Test infrastructure:
imported only by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs."""
from __future__ import annotations

import ctypes
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_LIB = os.path.join(ROOT, "oracle", "_build", "libnbody_oracle.so")
_fp = ctypes.POINTER(ctypes.c_float)
_dp = ctypes.POINTER(ctypes.c_double)


def _p(a):
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_fp)


class Oracle:
    def __init__(self, path: str = ORACLE_LIB):
        L = ctypes.CDLL(path)
        u64, f32, i32 = ctypes.c_uint64, ctypes.c_float, ctypes.c_int
        L.oracle_disk_galaxy.argtypes = [u64] + [_fp] * 6
        L.oracle_accel.argtypes = [u64, _fp, _fp, _fp, f32, i32, u64, u64, _fp, _fp, _fp]
        L.oracle_accel_mass.argtypes = [u64, _fp, _fp, _fp, _fp, f32, u64, u64, _fp, _fp, _fp]
        L.oracle_accel_f64.argtypes = [u64, _fp, _fp, _fp, f32, u64, u64, _dp, _dp, _dp]
        L.oracle_step.argtypes = [u64] + [_fp] * 6 + [f32, f32, f32, f32, i32, i32]
        L.oracle_time_accel.argtypes = [u64, _fp, _fp, _fp, f32, u64, u64, i32]
        L.oracle_time_accel.restype = ctypes.c_double
        L.oracle_num_threads.restype = i32
        L.oracle_fnv1a64.argtypes = [u64, i32, ctypes.POINTER(_fp)]
        L.oracle_fnv1a64.restype = u64
        self.L = L

    def disk_galaxy(self, n):
        a = [np.empty(n, np.float32) for _ in range(6)]
        assert self.L.oracle_disk_galaxy(n, *[_p(v) for v in a]) == 0
        return a

    def accel(self, x, y, z, eps, method=0, i_begin=0, i_end=None):
        n = len(x)
        i_end = n if i_end is None else i_end
        out = [np.empty(i_end - i_begin, np.float32) for _ in range(3)]
        rc = self.L.oracle_accel(n, _p(x), _p(y), _p(z), eps, method, i_begin, i_end, *[_p(v) for v in out])
        assert rc == 0
        return out

    def accel_mass(self, x, y, z, m, eps, i_begin=0, i_end=None):
        n = len(x)
        i_end = n if i_end is None else i_end
        out = [np.empty(i_end - i_begin, np.float32) for _ in range(3)]
        rc = self.L.oracle_accel_mass(n, _p(x), _p(y), _p(z), _p(np.ascontiguousarray(m, np.float32)), eps, i_begin, i_end,
                                      *[_p(v) for v in out])
        assert rc == 0
        return out

    def accel_f64(self, x, y, z, eps, i_begin=0, i_end=None):
        n = len(x)
        i_end = n if i_end is None else i_end
        out = [np.empty(i_end - i_begin, np.float64) for _ in range(3)]
        rc = self.L.oracle_accel_f64(n, _p(x), _p(y), _p(z), eps, i_begin, i_end,
                                     *[v.ctypes.data_as(_dp) for v in out])
        assert rc == 0
        return out

    def step(self, state, G=2.0, dt=0.005, damping=0.999998, eps=1.0e-7, method=0, iters=1):
        s = [np.array(a, np.float32, copy=True) for a in state]
        rc = self.L.oracle_step(len(s[0]), *[_p(v) for v in s], G, dt, damping, eps, method, iters)
        assert rc == 0
        return s

    def time_accel(self, x, y, z, eps, i_begin, i_count, reps=1):
        return float(self.L.oracle_time_accel(len(x), _p(x), _p(y), _p(z), eps, i_begin, i_count, reps))

    def set_num_threads(self, n: int):
        self.L.oracle_set_num_threads(int(n))

    def num_threads(self):
        return int(self.L.oracle_num_threads())

    def fnv1a64(self, arrays) -> str:
        arrays = [np.ascontiguousarray(a, np.float32) for a in arrays]
        ptrs = (_fp * len(arrays))(*[_p(a) for a in arrays])
        return f"{self.L.oracle_fnv1a64(len(arrays[0]), len(arrays), ptrs):016x}"


def rel_err(a, b):
    """per-body |a-b| / |b| over 3-vectors given as (x,y,z) lists"""
    num = np.sqrt(sum((np.asarray(p, np.float64) - np.asarray(q, np.float64)) ** 2 for p, q in zip(a, b)))
    den = np.sqrt(sum(np.asarray(q, np.float64) ** 2 for q in b))
    return num / np.maximum(den, 1e-300)
