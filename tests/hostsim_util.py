"""TEST INFRASTRUCTURE: runs the env logic + frame composition headers of the CUDA kernels on the CPU
(tests/hostsim/mg_hostsim.cpp) behind the same Python classes, so parity with the oracle can be checked
in the CPU-only tier. The product never uses this: ``memory_gym_b200`` has no CPU path."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "endless-memory-gym_b200", "csrc")
SRC = os.path.join(ROOT, "tests", "hostsim", "mg_hostsim.cpp")
OUT_DIR = os.path.join(ROOT, "tests", "hostsim", "_build")
OUT = os.path.join(OUT_DIR, "libmg_hostsim.so")


def build():
    deps = [SRC] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cpp"))]
    deps.append(os.path.join(ROOT, "include", "memgym_b200.h"))
    if os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in deps):
        return OUT
    os.makedirs(OUT_DIR, exist_ok=True)
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-Wall", "-o", OUT, SRC,
           os.path.join(CSRC, "mg_params_host.cpp")]
    subprocess.run(cmd, check=True)
    return OUT


class HostSimBackend:
    """Same interface as memory_gym_b200.vec_env._CudaBackend, on host memory."""

    name = "hostsim"

    def __init__(self):
        from memory_gym_b200 import _lib as L
        from memory_gym_b200.vec_env import _CudaBackend
        self.L = L
        self.sim = C.CDLL(build())
        self._host = _CudaBackend.__new__(_CudaBackend)  # host-side helpers (assets, seeding, validation)
        self._host.lib = L.lib()
        vp = C.c_void_p
        for fam, P in (("mm", L.MMParams), ("mp", L.MPParams), ("ss", L.SSParams), ("ep", L.EPParams)):
            getattr(self.sim, f"mgsim_{fam}").argtypes = [C.POINTER(P), vp, vp, vp, C.POINTER(L.StepIO), C.c_int, C.c_int]

    def check_device(self, device):
        assert device.type == "cpu"

    def stream(self, device):
        return None

    def build_assets(self, family, params):
        return self._host.build_assets(family, params)

    def seed(self, seeds):
        return self._host.seed(seeds)

    def check_params(self, family, params):
        return self._host.check_params(family, params)

    def reset(self, family, params, assets, state, rng4, io, n, stream):
        rc = getattr(self.sim, f"mgsim_{family}")(C.byref(params), assets, state, rng4, C.byref(io), n, 0)
        assert rc == 0, rc

    def step(self, family, params, assets, state, io, n, stream):
        rc = getattr(self.sim, f"mgsim_{family}")(C.byref(params), assets, state, None, C.byref(io), n, 1)
        assert rc == 0, rc

    def step_many(self, family, params, assets, state, ios, n_steps, n, stream):
        for t in range(n_steps):
            self.step(family, params, assets, state, ios[t], n, stream)

    def rollout(self, family, params, assets, state, rio, n, stream):
        """mg_*_rollout on the host: env-major, all T steps of an env back to back (like the kernel)."""
        fn = getattr(self.sim, f"mgsim_{family}_rollout")
        fn.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        rc = fn(C.byref(params), assets, state, C.byref(rio), n)
        assert rc == 0, rc
