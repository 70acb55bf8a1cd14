"""ctypes front end of oracle/libgddp_oracle.so (generic unconstrained DDP oracle, model (B)).  TEST INFRASTRUCTURE."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_dp, _ip, _lp = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_int64)


class _Problem(C.Structure):
    _fields_ = [("model", C.c_int), ("nx", C.c_int), ("nu", C.c_int), ("B", C.c_int), ("N", C.c_int), ("iter_max", C.c_int),
                ("dt", C.c_double), ("tol", C.c_double), ("x0", _dp), ("xg", _dp), ("u_init", _dp),
                ("q", C.c_double * 12), ("qf", C.c_double * 12), ("r", C.c_double * 4), ("uh", C.c_double * 4)]


class _Result(C.Structure):
    _fields_ = [("rtn", _ip), ("iters", _ip), ("cost", _dp), ("x", _dp), ("u", _dp), ("stats", _lp)]


def build():
    so = os.path.join(_HERE, "libgddp_oracle.so")
    src = [os.path.join(_HERE, f) for f in ("gddp_oracle.c", "gddp_impl.h", "gddp_oracle.h")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src):
        subprocess.check_call(["gcc", "-O2", "-march=x86-64-v2", "-fPIC", "-fopenmp", "-Wall", "-std=c11", "-ffp-contract=off",
                               "-shared", "-o", so, src[0], "-lm"])
    return so


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
    return _lib


class Result:
    def __init__(self, B, N, nx, nu):
        self.rtn = np.zeros(B, np.int32); self.iters = np.zeros(B, np.int32); self.cost = np.zeros(B)
        self.x = np.zeros((B, N + 1, nx)); self.u = np.zeros((B, N, nu)); self.stats = np.zeros((B, 2), np.int64)


def solve_batch(gp, fp32=False, nthreads=1) -> Result:
    """gp: direct_b200.gddp.GddpProblem"""
    p = _Problem(gp.model, gp.nx, gp.nu, gp.B, gp.N, gp.iter_max, gp.dt, gp.tol, gp.x0.ctypes.data_as(_dp), gp.xg.ctypes.data_as(_dp),
                 None if gp.u_init is None else gp.u_init.ctypes.data_as(_dp))
    for k in range(gp.nx):
        p.q[k], p.qf[k] = gp.q[k], gp.qf[k]
    for k in range(gp.nu):
        p.r[k], p.uh[k] = gp.r[k], gp.uh[k]
    out = Result(gp.B, gp.N, gp.nx, gp.nu)
    o = _Result(out.rtn.ctypes.data_as(_ip), out.iters.ctypes.data_as(_ip), out.cost.ctypes.data_as(_dp), out.x.ctypes.data_as(_dp),
                out.u.ctypes.data_as(_dp), out.stats.ctypes.data_as(_lp))
    st = lib().gddp_oracle_solve_batch(C.byref(p), C.byref(o), int(fp32), int(nthreads))
    if st:
        raise RuntimeError(f"gddp oracle failed with status {st}")
    return out


def model(model_id, x, u):
    nx, nu = (12, 4) if model_id == 1 else (6, 3)
    x = np.ascontiguousarray(x, dtype=np.float64); u = np.ascontiguousarray(u, dtype=np.float64)
    f, F, G = np.zeros(nx), np.zeros((nx, nx)), np.zeros((nx, nu))
    lib().gddp_oracle_model(int(model_id), x.ctypes.data_as(_dp), u.ctypes.data_as(_dp), f.ctypes.data_as(_dp), F.ctypes.data_as(_dp),
                            G.ctypes.data_as(_dp))
    return f, F, G
