"""ctypes front end of tools/libemu.so (the CPU lane-by-lane emulation of the CUDA solver). DEBUG TOOL."""
import ctypes as C
import os
import subprocess

import numpy as np

from oracle import oracle_py as O

_HERE = os.path.dirname(os.path.abspath(__file__))


def build():
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-Wno-unknown-pragmas",
                           "-o", os.path.join(_HERE, "libemu.so"), os.path.join(_HERE, "emulate.cpp")])


_lib = None


def lib():
    global _lib
    if _lib is None:
        p = os.path.join(_HERE, "libemu.so")
        src = [os.path.join(_HERE, "emulate.cpp")] + [os.path.join(_HERE, "..", "direct_b200", "csrc", f) for f in ("ipddp_solver.h", "simt.h")]
        if not os.path.exists(p) or any(os.path.getmtime(s) > os.path.getmtime(p) for s in src):
            build()
        _lib = C.CDLL(p)
    return _lib


def solve_batch(pb, fp32=False, trace_cap=0, **kw):
    keep = []
    b = O._batch_struct(pb, keep, **kw)
    out = O.Result(pb.B, pb.N)
    o = out.c_struct()
    tr = np.zeros((max(trace_cap, 1), 12))
    n = C.c_int(0)
    st = lib().emu_solve_batch(C.byref(b), C.byref(o), int(fp32), tr.ctypes.data_as(C.POINTER(C.c_double)) if trace_cap else None,
                               trace_cap, C.byref(n))
    assert st == 0
    if trace_cap:
        return out, tr[:n.value]
    return out


def two_stage_batch(pb, fp32=False, stage0=None, stage1=None, time_power=2):
    from direct_b200.problems import STAGE0, STAGE1
    s0 = dict(STAGE0 if stage0 is None else stage0)
    s1 = dict(STAGE1 if stage1 is None else stage1)
    keep = []
    b = O._batch_struct(pb, keep)
    opts = O._TwoStage(s0["w_snap"], s0["w_terminal"], s0["w_time"], s0["iter_max"], s1["w_snap"], s1["w_terminal"],
                       s1["w_time"], s1["iter_max"], time_power)
    r0, r1 = O.Result(pb.B, pb.N), O.Result(pb.B, pb.N)
    o0, o1 = r0.c_struct(), r1.c_struct()
    st = lib().emu_two_stage_batch(C.byref(b), C.byref(opts), C.byref(o0), C.byref(o1), int(fp32))
    assert st == 0
    return r0, r1
