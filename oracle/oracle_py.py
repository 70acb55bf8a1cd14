"""ctypes front end of the CPU parity oracle (oracle/libipddp_oracle.so) and, when built, of the
reference's own translation unit (oracle/_ref/libddp_ref.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Nothing under direct_b200/ imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_lp = C.POINTER(C.c_long)


class _Batch(C.Structure):
    _fields_ = [("B", C.c_int), ("N", C.c_int), ("P_max", C.c_int), ("planes", _dp), ("nplanes", _ip),
                ("durations", _dp), ("seeds", _dp), ("x0", _dp), ("xd", _dp), ("init_bez", _dp),
                ("max_vel", C.c_double), ("max_acc", C.c_double), ("w_snap", C.c_double),
                ("w_terminal", C.c_double), ("w_time", C.c_double), ("iter_max", C.c_int),
                ("time_power", C.c_int), ("zero_init", C.c_int), ("line_init", C.c_int), ("minvo", C.c_int),
                ("infeas", _ip), ("infeas_all", C.c_int)]


class _Out(C.Structure):
    _fields_ = [("rtn", _ip), ("infeas_out", _ip), ("line_failed_out", _ip), ("iters", _ip), ("cost", _dp),
                ("x_final", _dp), ("poly_coeff", _dp), ("bez_coeff", _dp), ("poly_time", _dp), ("jerk", _dp),
                ("stats", _lp)]


class _TwoStage(C.Structure):
    _fields_ = [("w_snap0", C.c_double), ("w_terminal0", C.c_double), ("w_time0", C.c_double),
                ("iter_max0", C.c_int), ("w_snap", C.c_double), ("w_terminal", C.c_double),
                ("w_time", C.c_double), ("iter_max", C.c_int), ("time_power", C.c_int)]


class _Trace(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("cost", "costq", "logcost", "err", "mu", "reg", "stepsize", "opterr")] + \
               [(n, C.c_int) for n in ("step", "fp_failed", "n_bwd")]


def build(ref: bool = True) -> None:
    """Compile the oracle (and oracle/_ref when /root/reference is mounted)."""
    subprocess.check_call(["make", "-s", "-C", _HERE, "all"])
    if ref and os.path.exists("/root/reference/global_planner/src/ddp_optimizer.cpp") \
            and os.path.exists(os.path.join(_HERE, "ref_driver.cpp")):
        subprocess.check_call(["make", "-s", "-C", _HERE, "ref"])
        if os.path.exists(os.path.join(_HERE, "bezier_ref_driver.cpp")):
            subprocess.check_call(["make", "-s", "-C", _HERE, "bezier"])
        if os.path.exists(os.path.join(_HERE, "..", "direct_b200", "libdirect_ddp_b200.so")):
            subprocess.check_call(["make", "-s", "-C", _HERE, "dropin"])


_lib = None
_ref = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "libipddp_oracle.so")
        if not os.path.exists(path):
            build(ref=False)
        _lib = C.CDLL(path)
        _lib.ipddp_oracle_solve_batch.argtypes = [C.POINTER(_Batch), C.POINTER(_Out), C.c_int]
        _lib.ipddp_oracle_two_stage_batch.argtypes = [C.POINTER(_Batch), C.POINTER(_TwoStage), C.POINTER(_Out),
                                                      C.POINTER(_Out), C.c_int]
        _lib.ipddp_oracle_solve_traced.argtypes = [C.POINTER(_Batch), C.c_int, C.POINTER(_Out),
                                                   C.POINTER(_Trace), C.c_int, _ip]
        _lib.ipddp_oracle_time_allocation.argtypes = [C.c_int, _dp, _dp, _dp, C.c_double, C.c_double, _dp]
        _lib.ipddp_oracle_time_allocation.restype = None
    return _lib


def ref_available() -> bool:
    return os.path.exists(os.path.join(_HERE, "_ref", "libddp_ref.so"))


def ref_lib():
    global _ref
    if _ref is None:
        _ref = C.CDLL(os.path.join(_HERE, "_ref", "libddp_ref.so"))
        _ref.ddp_ref_solve_batch.argtypes = [C.POINTER(_Batch), C.POINTER(_Out), C.c_int]
    return _ref


_dropin = None


def dropin_available() -> bool:
    return os.path.exists(os.path.join(_HERE, "_ref", "libddp_dropin.so"))


def dropin_lib():
    """ref_driver.cpp linked with direct_b200/host/ddp_optimizer_b200.cpp (the B200 drop-in TU) instead of the
    reference's ddp_optimizer.cpp: exercises the product through the reference's C++ class API."""
    global _dropin
    if _dropin is None:
        _dropin = C.CDLL(os.path.join(_HERE, "_ref", "libddp_dropin.so"))
        _dropin.ddp_ref_solve_batch.argtypes = [C.POINTER(_Batch), C.POINTER(_Out), C.c_int]
    return _dropin


def _p(a, t=_dp):
    return None if a is None else a.ctypes.data_as(t)


class Result:
    """Per-trajectory outputs of polyCurveGeneration (host numpy arrays)."""

    def __init__(self, B, N):
        self.rtn = np.zeros(B, np.int32)
        self.infeas_out = np.zeros(B, np.int32)
        self.line_failed_out = np.zeros(B, np.int32)
        self.iters = np.zeros(B, np.int32)
        self.cost = np.zeros(B)
        self.x_final = np.zeros((B, 9))
        self.poly_coeff = np.zeros((B, N, 18))
        self.bez_coeff = np.zeros((B, N, 18))
        self.poly_time = np.zeros((B, N))
        self.jerk = np.zeros((B, N))
        self.stats = np.zeros((B, 8), np.int64)

    def c_struct(self):
        return _Out(_p(self.rtn, _ip), _p(self.infeas_out, _ip), _p(self.line_failed_out, _ip), _p(self.iters, _ip),
                    _p(self.cost), _p(self.x_final), _p(self.poly_coeff), _p(self.bez_coeff), _p(self.poly_time),
                    _p(self.jerk), _p(self.stats, _lp))


def _batch_struct(pb, keep, *, init_bez=None, durations=None, infeas=1, zero_init=1, line_init=0, minvo=0,
                  w_snap=1.0, w_terminal=1.0, w_time=1.0, iter_max=50, time_power=2):
    dur = np.ascontiguousarray(pb.durations if durations is None else durations, dtype=np.float64)
    ib = None if init_bez is None else np.ascontiguousarray(init_bez, dtype=np.float64)
    inf_arr = None
    inf_all = 0
    if np.ndim(infeas) == 0:
        inf_all = int(infeas)
    else:
        inf_arr = np.ascontiguousarray(infeas, dtype=np.int32)
    keep.extend([dur, ib, inf_arr])
    return _Batch(pb.B, pb.N, pb.P_max, _p(pb.planes), _p(pb.nplanes, _ip), _p(dur), _p(pb.seeds), _p(pb.x0),
                  _p(pb.xd), _p(ib), pb.max_vel, pb.max_acc, w_snap, w_terminal, w_time, iter_max, time_power,
                  int(zero_init), int(line_init), int(minvo), _p(inf_arr, _ip), inf_all)


def solve_batch(pb, nthreads=1, use_ref=False, use_dropin=False, **kw) -> Result:
    """One polyCurveGeneration call per trajectory of ProblemBatch ``pb``."""
    keep = []
    b = _batch_struct(pb, keep, **kw)
    out = Result(pb.B, pb.N)
    o = out.c_struct()
    if use_dropin:
        fn = dropin_lib().ddp_ref_solve_batch
    else:
        fn = ref_lib().ddp_ref_solve_batch if use_ref else lib().ipddp_oracle_solve_batch
    st = fn(C.byref(b), C.byref(o), int(nthreads))
    if st:
        raise RuntimeError(f"oracle solve failed with status {st}")
    return out


def two_stage_batch(pb, nthreads=1, stage0=None, stage1=None, time_power=2):
    """teach_repeat_planner.cpp:853-951 protocol; returns (stage-0 Result, stage-1 Result)."""
    from direct_b200.problems import STAGE0, STAGE1
    s0 = dict(STAGE0 if stage0 is None else stage0)
    s1 = dict(STAGE1 if stage1 is None else stage1)
    keep = []
    b = _batch_struct(pb, keep)
    opts = _TwoStage(s0["w_snap"], s0["w_terminal"], s0["w_time"], s0["iter_max"], s1["w_snap"],
                     s1["w_terminal"], s1["w_time"], s1["iter_max"], time_power)
    r0, r1 = Result(pb.B, pb.N), Result(pb.B, pb.N)
    o0, o1 = r0.c_struct(), r1.c_struct()
    st = lib().ipddp_oracle_two_stage_batch(C.byref(b), C.byref(opts), C.byref(o0), C.byref(o1), int(nthreads))
    if st:
        raise RuntimeError(f"oracle two-stage failed with status {st}")
    return r0, r1


def solve_traced(pb, i=0, cap=512, **kw):
    """Solve trajectory ``i`` and return (Result for it, list of per-iteration trace dicts)."""
    keep = []
    b = _batch_struct(pb, keep, **kw)
    out = Result(1, pb.N)
    o = out.c_struct()
    tr = (_Trace * cap)()
    n = C.c_int(0)
    st = lib().ipddp_oracle_solve_traced(C.byref(b), int(i), C.byref(o), tr, cap, C.byref(n))
    if st:
        raise RuntimeError(f"oracle solve failed with status {st}")
    rows = [{f: getattr(tr[k], f) for f, _ in _Trace._fields_} for k in range(n.value)]
    return out, rows


def max_threads() -> int:
    lib().ipddp_oracle_max_threads.restype = C.c_int
    return int(lib().ipddp_oracle_max_threads())


def bezier_ref_available() -> bool:
    return os.path.exists(os.path.join(_HERE, "_ref", "libbezier_ref.so"))


def bezier_sample_ref(bez_coeff, poly_time, S):
    """The REFERENCE's own Bernstein::getPos / getVel / getAcc (utils/bezier_base.h:77-115, compiled unmodified into
    oracle/_ref/libbezier_ref.so by `make -C oracle bezier`), scaled like its callers do.  Same shapes as bezier_sample."""
    lib_ = C.CDLL(os.path.join(_HERE, "_ref", "libbezier_ref.so"))
    lib_.bezier_ref_sample.argtypes = [C.c_int, C.c_int, C.c_int, _dp, _dp, _dp, _dp, _dp]
    bez = np.ascontiguousarray(bez_coeff, dtype=np.float64)
    T = np.ascontiguousarray(poly_time, dtype=np.float64)
    B, N = T.shape
    pos = np.zeros((B, N, S, 3)); vel = np.zeros_like(pos); acc = np.zeros_like(pos)
    st = lib_.bezier_ref_sample(B, N, int(S), _p(bez), _p(T), _p(pos), _p(vel), _p(acc))
    if st:
        raise RuntimeError(f"bezier_ref_sample failed with status {st}")
    return pos, vel, acc


# ---- §8(f) #3: trajectory sampling, restated from the reference's Bernstein evaluators ---------------------------
def bezier_sample(bez_coeff, poly_time, S):
    """Bernstein::getPos / getVel / getAcc of global_planner/include/global_planner/utils/bezier_base.h:77-115 at the
    S parameters s_k = k / (S - 1) of every segment, scaled like the callers do (teach_repeat_planner.cpp:1557-1560:
    position = time * getPosFromBezier; :681-682: velocity = getVel, acceleration = getAcc / time).
    bez_coeff [B][N][18] rows [x*6, y*6, z*6] (scaled by 1/T), poly_time [B][N].  Returns pos, vel, acc [B][N][S][3].
    Plain numpy with pow() exactly as the reference writes it; TEST INFRASTRUCTURE ONLY."""
    from math import comb
    bez = np.asarray(bez_coeff, dtype=np.float64)
    T = np.asarray(poly_time, dtype=np.float64)
    B, N = T.shape
    n = 5
    cp = bez.reshape(B, N, 3, n + 1)
    s = (np.arange(S, dtype=np.float64) / (S - 1)) if S > 1 else np.zeros(1)
    pos = np.zeros((B, N, S, 3)); vel = np.zeros_like(pos); acc = np.zeros_like(pos)
    for j in range(n + 1):
        w = comb(n, j) * np.power(s, j) * np.power(1 - s, n - j)
        pos += cp[:, :, None, :, j] * w[None, None, :, None]
    for j in range(n):
        w = comb(n - 1, j) * n * np.power(s, j) * np.power(1 - s, n - j - 1)
        vel += (cp[:, :, None, :, j + 1] - cp[:, :, None, :, j]) * w[None, None, :, None]
    for j in range(n - 1):
        w = comb(n - 2, j) * n * (n - 1) * np.power(s, j) * np.power(1 - s, n - j - 2)
        acc += (cp[:, :, None, :, j + 2] - 2 * cp[:, :, None, :, j + 1] + cp[:, :, None, :, j]) * w[None, None, :, None]
    return pos * T[:, :, None, None], vel, acc / T[:, :, None, None]
