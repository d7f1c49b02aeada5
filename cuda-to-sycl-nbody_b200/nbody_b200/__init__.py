"""ctypes binding of the C ABI in include/nbody_b200.h (lib/libnbody_b200.so).

Test/bench harness only: the product is the shared library and the C++ mirror of the
reference's ``simulation::DiskGalaxySimulator`` (cuda-to-sycl-nbody_b200/cxx/).  The class below
mirrors that interface (reference src/simulator.cuh:129-160) method for method so the parity
tests read like calls into the reference:

    stepSim / getLastStepTime / getNumParticles / getParticlePos / getParticleVel /
    getDeviceName / getGwSize / getCM

There is no CPU fallback: if the library is missing or no GPU is visible, construction raises.
"""
from __future__ import annotations

import ctypes
import os
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.normpath(os.path.join(_HERE, "..", "lib", "libnbody_b200.so"))
# the same library plus the comparison kernels (make VARIANTS=1); tests of those kernels load it explicitly
VARIANTS_LIB_PATH = os.path.normpath(os.path.join(_HERE, "..", "lib", "libnbody_b200_variants.so"))

CALC_BRANCH, CALC_PREDICATED, CALC_PREDICATED_FIXED = 0, 1, 2
DEVSTEP_MASS = 1
ABI_VERSION = 2
KERNEL_AUTO, KERNEL_GENERIC, KERNEL_PACKED, KERNEL_SCALAR = 0, 1, 2, 3

# every symbol include/nbody_b200.h declares (checked by tests/test_capi_cpu.py)
ABI_SYMBOLS = [
    "nbody_abi_version", "nbody_last_error", "nbody_device_count", "nbody_default_params",
    "nbody_generate_disk_galaxy", "nbody_plan_shard", "nbody_create", "nbody_create_rank", "nbody_nccl_unique_id",
    "nbody_destroy", "nbody_set_kernel", "nbody_kernel_name", "nbody_describe_auto", "nbody_set_state", "nbody_set_mass",
    "nbody_step", "nbody_last_step_ms", "nbody_last_step_device_ms", "nbody_launch_count",
    "nbody_read_pos", "nbody_read_vel", "nbody_read_pos_f4", "nbody_read_vel_f4",
    "nbody_read_state", "nbody_local_range", "nbody_read_local", "nbody_host_register", "nbody_host_unregister",
    "nbody_save_state", "nbody_load_state",
    "nbody_device_name", "nbody_num_particles", "nbody_num_gpus", "nbody_world_size",
    "nbody_compute_accel", "nbody_launch_step_device",
]


class NBodyError(RuntimeError):
    pass


class Params(ctypes.Structure):
    """struct nbody_params == the hot-path fields of SimParam (reference src/sim_param.hpp:30-39)."""
    _fields_ = [
        ("G", ctypes.c_float),
        ("dt", ctypes.c_float),
        ("num_particles", ctypes.c_uint64),
        ("iters_per_frame", ctypes.c_int32),
        ("damping", ctypes.c_float),
        ("dist_eps", ctypes.c_float),
        ("gw_size", ctypes.c_int32),
        ("calc_method", ctypes.c_int32),
    ]


@dataclass
class SimParam:
    """Python mirror of the reference's SimParam with its defaults (src/sim_param.cpp:12-22)."""
    G: float = 2.0
    dt: float = 0.005
    numParticles: int = 50 * 256
    numFrames: int = 2**64 - 1
    simIterationsPerFrame: int = 4
    damping: float = 0.999998
    distEps: float = 1.0e-7
    gwSize: int = 64
    calcMethod: int = CALC_BRANCH

    def to_c(self) -> Params:
        return Params(self.G, self.dt, self.numParticles, self.simIterationsPerFrame, self.damping,
                      self.distEps, self.gwSize, self.calcMethod)


_lib = None
_fp = ctypes.POINTER(ctypes.c_float)


def load_library(path: str | None = None) -> ctypes.CDLL:
    """Loads libnbody_b200.so; raises if it is not built (no fallback of any kind)."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise NBodyError(f"{p} is not built: run `make -C cuda-to-sycl-nbody_b200` "
                         "(or __graft_entry__.build()); there is no CPU fallback")
    lib = ctypes.CDLL(p)
    H = ctypes.c_void_p
    lib.nbody_abi_version.restype = ctypes.c_int
    lib.nbody_last_error.restype = ctypes.c_char_p
    lib.nbody_device_count.restype = ctypes.c_int
    lib.nbody_default_params.argtypes = [ctypes.POINTER(Params)]
    lib.nbody_default_params.restype = None
    lib.nbody_generate_disk_galaxy.argtypes = [ctypes.c_uint64] + [_fp] * 6
    lib.nbody_plan_shard.argtypes = [ctypes.c_uint64, ctypes.c_int, ctypes.c_int,
                                     ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_uint64)]
    lib.nbody_create.argtypes = [ctypes.POINTER(Params), ctypes.c_int, ctypes.POINTER(H)]
    lib.nbody_create_rank.argtypes = [ctypes.POINTER(Params), ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                      ctypes.c_void_p, ctypes.POINTER(H)]
    lib.nbody_nccl_unique_id.argtypes = [ctypes.c_void_p]
    lib.nbody_destroy.argtypes = [H]
    lib.nbody_set_kernel.argtypes = [H, ctypes.c_int]
    lib.nbody_describe_auto.argtypes = [ctypes.POINTER(Params), ctypes.c_uint64, ctypes.c_int, ctypes.c_int,
                                        ctypes.c_char_p, ctypes.c_size_t]
    lib.nbody_kernel_name.argtypes = [H]
    lib.nbody_kernel_name.restype = ctypes.c_char_p
    lib.nbody_set_state.argtypes = [H] + [_fp] * 6
    lib.nbody_set_mass.argtypes = [H, _fp]
    lib.nbody_step.argtypes = [H]
    lib.nbody_last_step_ms.argtypes = [H]
    lib.nbody_last_step_ms.restype = ctypes.c_float
    lib.nbody_last_step_device_ms.argtypes = [H]
    lib.nbody_last_step_device_ms.restype = ctypes.c_float
    lib.nbody_launch_count.argtypes = [H]
    lib.nbody_launch_count.restype = ctypes.c_uint64
    lib.nbody_read_pos.argtypes = [H] + [_fp] * 3
    lib.nbody_read_vel.argtypes = [H] + [_fp] * 3
    lib.nbody_read_pos_f4.argtypes = [H, _fp]
    lib.nbody_read_vel_f4.argtypes = [H, _fp]
    lib.nbody_read_state.argtypes = [H] + [_fp] * 6
    lib.nbody_local_range.argtypes = [H, ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_uint64)]
    lib.nbody_read_local.argtypes = [H] + [_fp] * 6
    lib.nbody_host_register.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
    lib.nbody_host_unregister.argtypes = [ctypes.c_void_p]
    lib.nbody_save_state.argtypes = [H, ctypes.c_char_p]
    lib.nbody_load_state.argtypes = [H, ctypes.c_char_p]
    lib.nbody_device_name.argtypes = [H]
    lib.nbody_device_name.restype = ctypes.c_char_p
    lib.nbody_num_particles.argtypes = [H]
    lib.nbody_num_particles.restype = ctypes.c_uint64
    lib.nbody_num_gpus.argtypes = [H]
    lib.nbody_world_size.argtypes = [H]
    lib.nbody_compute_accel.argtypes = [H] + [_fp] * 3
    lib.nbody_launch_step_device.argtypes = [ctypes.POINTER(Params), ctypes.c_void_p, ctypes.c_void_p,
                                             ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint64,
                                             ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    if path is None:
        _lib = lib
    return lib


def _ptr(a: np.ndarray):
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_fp)


def _check(lib, rc: int, what: str):
    if rc != 0:
        raise NBodyError(f"{what} failed with code {rc}: {lib.nbody_last_error().decode()}")


def device_count() -> int:
    return load_library().nbody_device_count()


def generate_disk_galaxy(n: int):
    """Host-only generator == randomParticlePos + initialParticleVel (src/simulator.cu:131-158)."""
    lib = load_library()
    arrs = [np.empty(n, np.float32) for _ in range(6)]
    _check(lib, lib.nbody_generate_disk_galaxy(n, *[_ptr(a) for a in arrs]), "nbody_generate_disk_galaxy")
    return arrs


def plan_shard(n: int, world: int, rank: int) -> tuple[int, int]:
    """(begin, count) of the i-range rank `rank` of `world` owns."""
    lib = load_library()
    b, c = ctypes.c_uint64(), ctypes.c_uint64()
    _check(lib, lib.nbody_plan_shard(n, world, rank, ctypes.byref(b), ctypes.byref(c)), "nbody_plan_shard")
    return int(b.value), int(c.value)


def describe_auto(n: int, sms: int = 148, has_mass: bool = False, params: "SimParam | None" = None, lib=None) -> str:
    """the kernel configuration AUTO picks for a shard of n bodies on a device with `sms` SMs (host-only)"""
    lib = lib or load_library()
    buf = ctypes.create_string_buffer(96)
    c = (params or SimParam()).to_c()
    _check(lib, lib.nbody_describe_auto(ctypes.byref(c), n, sms, int(has_mass), buf, len(buf)), "nbody_describe_auto")
    return buf.value.decode()


def nccl_unique_id() -> bytes:
    lib = load_library()
    buf = ctypes.create_string_buffer(128)
    _check(lib, lib.nbody_nccl_unique_id(buf), "nbody_nccl_unique_id")
    return buf.raw


@dataclass
class ParticleData:
    """Host SoA, as the reference's ParticleData (src/simulator.cuh:75-84)."""
    x: np.ndarray
    y: np.ndarray
    z: np.ndarray


class DiskGalaxySimulator:
    """Mirror of simulation::DiskGalaxySimulator (reference src/simulator.cuh:129-160)."""

    def __init__(self, params: SimParam, n_gpus: int = 1, *, rank: int | None = None,
                 world: int | None = None, device: int = 0, unique_id: bytes | None = None, lib=None):
        self._lib = lib if lib is not None else load_library()
        self.params = params
        self._h = ctypes.c_void_p()
        cp = params.to_c()
        if rank is None:
            rc = self._lib.nbody_create(ctypes.byref(cp), n_gpus, ctypes.byref(self._h))
            what = "nbody_create"
        else:
            uid = ctypes.create_string_buffer(unique_id, 128) if unique_id else None
            rc = self._lib.nbody_create_rank(ctypes.byref(cp), device, rank, world, uid, ctypes.byref(self._h))
            what = "nbody_create_rank"
        _check(self._lib, rc, what)
        n = params.numParticles
        self._pos = ParticleData(*[np.zeros(n, np.float32) for _ in range(3)])
        self._vel = ParticleData(*[np.zeros(n, np.float32) for _ in range(3)])
        self._host_fresh = False

    # -- reference interface -------------------------------------------------------------------
    def stepSim(self):
        _check(self._lib, self._lib.nbody_step(self._h), "nbody_step")
        self._host_fresh = False

    def getLastStepTime(self) -> float:
        return float(self._lib.nbody_last_step_ms(self._h))

    def getNumParticles(self) -> int:
        return int(self._lib.nbody_num_particles(self._h))

    def getParticlePos(self) -> ParticleData:
        self._refresh()
        return self._pos

    def getParticleVel(self) -> ParticleData:
        self._refresh()
        return self._vel

    def getDeviceName(self) -> str:
        return self._lib.nbody_device_name(self._h).decode()

    def getGwSize(self) -> int:
        return self.params.gwSize

    def getCM(self) -> int:
        return self.params.calcMethod

    # -- extensions of the C ABI ---------------------------------------------------------------
    def getLastStepDeviceTime(self) -> float:
        return float(self._lib.nbody_last_step_device_ms(self._h))

    def setState(self, x, y, z, vx, vy, vz):
        arrs = [np.ascontiguousarray(a, np.float32) for a in (x, y, z, vx, vy, vz)]
        assert all(a.shape == (self.params.numParticles,) for a in arrs)
        _check(self._lib, self._lib.nbody_set_state(self._h, *[_ptr(a) for a in arrs]), "nbody_set_state")
        self._host_fresh = False

    def setMass(self, m):
        """per-body masses (float4.w); None restores the reference's unit masses"""
        if m is None:
            _check(self._lib, self._lib.nbody_set_mass(self._h, None), "nbody_set_mass")
        else:
            a = np.ascontiguousarray(m, np.float32)
            assert a.shape == (self.params.numParticles,)
            _check(self._lib, self._lib.nbody_set_mass(self._h, _ptr(a)), "nbody_set_mass")
        self._host_fresh = False

    def saveState(self, path: str):
        _check(self._lib, self._lib.nbody_save_state(self._h, path.encode()), "nbody_save_state")

    def loadState(self, path: str):
        _check(self._lib, self._lib.nbody_load_state(self._h, path.encode()), "nbody_load_state")
        self._host_fresh = False

    def setKernel(self, kernel: int):
        _check(self._lib, self._lib.nbody_set_kernel(self._h, kernel), "nbody_set_kernel")

    def kernelName(self) -> str:
        return self._lib.nbody_kernel_name(self._h).decode()

    def launchCount(self) -> int:
        return int(self._lib.nbody_launch_count(self._h))

    def computeAccel(self):
        n = self.params.numParticles
        a = [np.empty(n, np.float32) for _ in range(3)]
        _check(self._lib, self._lib.nbody_compute_accel(self._h, *[_ptr(v) for v in a]), "nbody_compute_accel")
        return a

    def readPosF4(self) -> np.ndarray:
        out = np.empty((self.params.numParticles, 4), np.float32)
        _check(self._lib, self._lib.nbody_read_pos_f4(self._h, _ptr(out.reshape(-1))), "nbody_read_pos_f4")
        return out

    def readVelF4(self) -> np.ndarray:
        out = np.empty((self.params.numParticles, 4), np.float32)
        _check(self._lib, self._lib.nbody_read_vel_f4(self._h, _ptr(out.reshape(-1))), "nbody_read_vel_f4")
        return out

    def readInto(self, x, y, z, vx, vy, vz):
        """Read-back into caller buffers (e.g. pinned memory) -- recvFromDevice, src/simulator.cu:106-129."""
        _check(self._lib, self._lib.nbody_read_state(self._h, *[_ptr(a) for a in (x, y, z, vx, vy, vz)]),
               "nbody_read_state")

    def localRange(self) -> tuple[int, int]:
        b, c = ctypes.c_uint64(), ctypes.c_uint64()
        _check(self._lib, self._lib.nbody_local_range(self._h, ctypes.byref(b), ctypes.byref(c)), "nbody_local_range")
        return int(b.value), int(c.value)

    def readLocalInto(self, x, y, z, vx, vy, vz):
        """positions + velocities of the bodies this handle's devices own (count floats per array)"""
        _check(self._lib, self._lib.nbody_read_local(self._h, *[_ptr(a) for a in (x, y, z, vx, vy, vz)]),
               "nbody_read_local")

    def close(self):
        if self._h:
            self._lib.nbody_destroy(self._h)
            self._h = ctypes.c_void_p()

    def _refresh(self):
        if not self._host_fresh:
            self.readInto(self._pos.x, self._pos.y, self._pos.z, self._vel.x, self._vel.y, self._vel.z)
            self._host_fresh = True

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
