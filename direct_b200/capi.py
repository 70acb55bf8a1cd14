"""ctypes binding of the C-ABI in include/direct_ddp.h (libdirect_ddp_b200.so).

This is the host-side mirror used by the tests and bench.py; the drop-in C++ translation unit for the
reference node is direct_b200/host/ddp_optimizer_b200.cpp.  There is no CPU fallback here either: if the
library is missing it is built with nvcc, and if no B200 is visible every solve raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build
from .problems import STAGE0, STAGE1, TIME_POWER, ProblemBatch

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_lp = C.POINTER(C.c_int64)


class Opts(C.Structure):
    _fields_ = [("device", C.c_int), ("precision", C.c_int), ("warps_per_block", C.c_int),
                ("blocks_per_sm", C.c_int), ("trace", C.c_int), ("ndevices", C.c_int), ("devices", C.POINTER(C.c_int))]


class Batch(C.Structure):
    _fields_ = [("B", C.c_int), ("N", C.c_int), ("P_max", C.c_int), ("planes", C.c_void_p), ("nplanes", C.c_void_p),
                ("durations", C.c_void_p), ("seeds", C.c_void_p), ("x0", C.c_void_p), ("xd", C.c_void_p),
                ("init_bez", C.c_void_p), ("infeas", C.c_void_p), ("infeas_all", C.c_int),
                ("max_vel", C.c_double), ("max_acc", C.c_double), ("w_snap", C.c_double),
                ("w_terminal", C.c_double), ("w_time", C.c_double), ("iter_max", C.c_int), ("time_power", C.c_int),
                ("zero_init", C.c_int), ("line_init", C.c_int), ("minvo", C.c_int), ("nknots", C.c_void_p)]


class ResultC(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("rtn", "infeas_out", "line_failed_out", "iters", "cost", "x_final",
                                           "poly_coeff", "bez_coeff", "poly_time", "jerk", "stats")]


class TwoStage(C.Structure):
    _fields_ = [("w_snap0", C.c_double), ("w_terminal0", C.c_double), ("w_time0", C.c_double), ("iter_max0", C.c_int),
                ("w_snap", C.c_double), ("w_terminal", C.c_double), ("w_time", C.c_double), ("iter_max", C.c_int),
                ("time_power", C.c_int)]


class Stats(C.Structure):
    _fields_ = [("kernel_ms", C.c_double), ("h2d_ms", C.c_double), ("d2h_ms", C.c_double),
                ("bwd_sweeps", C.c_int64), ("bwd_knots", C.c_int64), ("fwd_trials", C.c_int64), ("fwd_knots", C.c_int64),
                ("kernel_launches", C.c_int64), ("grid_blocks", C.c_int), ("block_threads", C.c_int),
                ("smem_bytes_per_block", C.c_int), ("workspace_slots", C.c_int), ("h2d_bytes", C.c_int64),
                ("d2h_bytes", C.c_int64), ("coop_jobs", C.c_int64), ("helper_units", C.c_int64),
                ("spec_searches", C.c_int64), ("spec_trials", C.c_int64), ("spec_sweeps", C.c_int64), ("spec_sweeps_used", C.c_int64)]


class CorridorC(C.Structure):
    """direct_ddp_corridor: msgs/msg/corridor.msg flattened (include/direct_ddp.h)."""
    _fields_ = [("path_id", C.c_int), ("N", C.c_int), ("P_max", C.c_int), ("planes", C.POINTER(C.c_double)),
                ("nplanes", C.POINTER(C.c_int32)), ("center", C.POINTER(C.c_double)), ("seed", C.POINTER(C.c_double))]


class TraceRow(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("cost", "costq", "logcost", "err", "mu", "reg", "stepsize", "opterr")] + \
               [(n, C.c_int32) for n in ("step", "fp_failed", "n_bwd", "t_us")]


EXPORTS = ["direct_ddp_version", "direct_ddp_create", "direct_ddp_destroy", "direct_ddp_last_error",
           "direct_ddp_solve_batch", "direct_ddp_solve_batch_device", "direct_ddp_solve_two_stage",
           "direct_ddp_solve_two_stage_device", "direct_ddp_time_allocation_device", "direct_ddp_last_stats",
           "direct_ddp_last_trace", "direct_ddp_measure_fma_peak", "direct_ddp_sample", "direct_ddp_sample_device",
           "direct_ddp_corridor_read", "direct_ddp_corridor_write", "direct_ddp_corridor_free", "direct_ddp_replay",
           "direct_ddp_replay_write", "direct_ddp_sm_clock_hz", "direct_ddp_device_count"]

_lib = None


def load_library(build_if_missing: bool = True):
    """dlopen libdirect_ddp_b200.so (building it with nvcc when absent or stale and nvcc is available)."""
    global _lib
    if _lib is None:
        path = os.environ.get("DIRECT_DDP_LIB")   # tuning experiments: another build of the same sources
        if not path:
            path = _build.LIB
            if build_if_missing and _build.is_stale():
                _build.build()
        if not os.path.exists(path):
            raise RuntimeError(f"{path} is missing: run `python -m direct_b200.build` (needs nvcc)")
        lib = C.CDLL(path)
        lib.direct_ddp_last_error.restype = C.c_char_p
        lib.direct_ddp_last_error.argtypes = [C.c_void_p]
        lib.direct_ddp_create.argtypes = [C.POINTER(Opts), C.POINTER(C.c_void_p)]
        lib.direct_ddp_destroy.argtypes = [C.c_void_p]
        lib.direct_ddp_solve_batch.argtypes = [C.c_void_p, C.POINTER(Batch), C.POINTER(ResultC)]
        lib.direct_ddp_solve_batch_device.argtypes = [C.c_void_p, C.POINTER(Batch), C.POINTER(ResultC), C.c_void_p]
        lib.direct_ddp_solve_two_stage.argtypes = [C.c_void_p, C.POINTER(Batch), C.POINTER(TwoStage), C.POINTER(ResultC),
                                                   C.POINTER(ResultC)]
        lib.direct_ddp_solve_two_stage_device.argtypes = [C.c_void_p, C.POINTER(Batch), C.POINTER(TwoStage),
                                                          C.POINTER(ResultC), C.POINTER(ResultC), C.c_void_p]
        lib.direct_ddp_time_allocation_device.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                                          C.c_void_p, C.c_double, C.c_double, C.c_void_p, C.c_void_p]
        lib.direct_ddp_last_stats.argtypes = [C.c_void_p, C.POINTER(Stats)]
        lib.direct_ddp_last_trace.argtypes = [C.c_void_p, C.POINTER(TraceRow), C.c_int, C.POINTER(C.c_int)]
        lib.direct_ddp_measure_fma_peak.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double)]
        lib.direct_ddp_sample.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 5
        lib.direct_ddp_sample_device.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 6
        lib.direct_ddp_corridor_read.argtypes = [C.c_char_p, C.POINTER(C.POINTER(CorridorC))]
        lib.direct_ddp_corridor_write.argtypes = [C.c_char_p, C.POINTER(CorridorC)]
        lib.direct_ddp_corridor_free.argtypes = [C.POINTER(CorridorC)]
        lib.direct_ddp_corridor_free.restype = None
        lib.direct_ddp_replay.argtypes = [C.c_void_p, C.POINTER(CorridorC), C.c_int, C.c_int, C.POINTER(TwoStage), C.c_double,
                                          C.c_double, C.c_void_p]
        lib.direct_ddp_replay_write.argtypes = [C.c_char_p, C.c_void_p, C.c_int]
        lib.direct_ddp_sm_clock_hz.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        if hasattr(lib, "direct_ddp_device_count"):   # absent from older builds loaded through DIRECT_DDP_LIB (tuning experiments)
            lib.direct_ddp_device_count.argtypes = [C.c_void_p]
        _lib = lib
    return _lib


class DirectDdpError(RuntimeError):
    pass


class HostResult:
    """Host (numpy) outputs of one polyCurveGeneration per trajectory."""

    def __init__(self, B: int, N: int):
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

    def c_struct(self) -> ResultC:
        return ResultC(*[getattr(self, n).ctypes.data for n, _ in ResultC._fields_])


def _ptr(a):
    return None if a is None else a.ctypes.data


def two_stage_opts(stage0=None, stage1=None, time_power=TIME_POWER) -> TwoStage:
    s0 = dict(STAGE0 if stage0 is None else stage0)
    s1 = dict(STAGE1 if stage1 is None else stage1)
    return TwoStage(s0["w_snap"], s0["w_terminal"], s0["w_time"], s0["iter_max"], s1["w_snap"], s1["w_terminal"],
                    s1["w_time"], s1["iter_max"], time_power)


class Solver:
    """One libdirect_ddp_b200 handle bound to one GPU, or (devices=[...]) sharding host-buffer batches over several."""

    def __init__(self, device: int = 0, precision: str = "fp64", warps_per_block: int = 0, blocks_per_sm: int = 0,
                 trace: bool = False, devices=None):
        self.lib = load_library()
        self.h = C.c_void_p()
        devs = None if devices is None else (C.c_int * len(devices))(*devices)
        opts = Opts(device if devices is None else devices[0], {"fp64": 0, "fp32": 1}[precision], warps_per_block, blocks_per_sm,
                    int(trace), 0 if devices is None else len(devices), devs)
        st = self.lib.direct_ddp_create(C.byref(opts), C.byref(self.h))
        if st != 0:
            msg = self.lib.direct_ddp_last_error(self.h).decode() if self.h else "create failed"
            if self.h:
                self.lib.direct_ddp_destroy(self.h)
                self.h = C.c_void_p()
            raise DirectDdpError(f"direct_ddp_create failed ({st}): {msg}")

    def close(self):
        if self.h:
            self.lib.direct_ddp_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, st):
        if st != 0:
            raise DirectDdpError(f"direct_ddp error {st}: {self.lib.direct_ddp_last_error(self.h).decode()}")

    @staticmethod
    def batch_struct(pb: ProblemBatch, keep: list, *, init_bez=None, durations=None, infeas=1, zero_init=1, line_init=0,
                     minvo=0, w_snap=1.0, w_terminal=1.0, w_time=1.0, iter_max=50, time_power=2, nknots=None) -> Batch:
        dur = np.ascontiguousarray(pb.durations if durations is None else durations, dtype=np.float64)
        ib = None if init_bez is None else np.ascontiguousarray(init_bez, dtype=np.float64)
        inf_arr, inf_all = None, 0
        if np.ndim(infeas) == 0:
            inf_all = int(infeas)
        else:
            inf_arr = np.ascontiguousarray(infeas, dtype=np.int32)
        nk = None if nknots is None else np.ascontiguousarray(nknots, dtype=np.int32)
        keep.extend([dur, ib, inf_arr, nk])
        return Batch(pb.B, pb.N, pb.P_max, _ptr(pb.planes), _ptr(pb.nplanes), _ptr(dur), _ptr(pb.seeds), _ptr(pb.x0),
                     _ptr(pb.xd), _ptr(ib), _ptr(inf_arr), inf_all, pb.max_vel, pb.max_acc, w_snap, w_terminal, w_time,
                     iter_max, time_power, int(zero_init), int(line_init), int(minvo), _ptr(nk))

    # ---- host-buffer entry points (the drop-in path: H2D + solve + D2H inside the call) ----------------
    def solve_batch(self, pb: ProblemBatch, **kw) -> HostResult:
        keep = []
        b = self.batch_struct(pb, keep, **kw)
        out = HostResult(pb.B, pb.N)
        o = out.c_struct()
        self._check(self.lib.direct_ddp_solve_batch(self.h, C.byref(b), C.byref(o)))
        return out

    def solve_two_stage(self, pb: ProblemBatch, stage0=None, stage1=None, time_power=TIME_POWER, want_stage0=True,
                        out0: HostResult | None = None, out1: HostResult | None = None, nknots=None):
        keep = []
        b = self.batch_struct(pb, keep, nknots=nknots)
        ts = two_stage_opts(stage0, stage1, time_power)
        r0 = out0 if out0 is not None else (HostResult(pb.B, pb.N) if want_stage0 else None)
        r1 = out1 if out1 is not None else HostResult(pb.B, pb.N)
        o0 = r0.c_struct() if r0 is not None else None
        o1 = r1.c_struct()
        self._check(self.lib.direct_ddp_solve_two_stage(self.h, C.byref(b), C.byref(ts),
                                                        C.byref(o0) if o0 is not None else None, C.byref(o1)))
        return r0, r1

    # ---- device-buffer entry points (raw pointers, e.g. torch tensors' data_ptr()) ----------------------
    def solve_two_stage_device(self, batch: Batch, ts: TwoStage, out0: ResultC | None, out1: ResultC, stream: int = 0):
        self._check(self.lib.direct_ddp_solve_two_stage_device(self.h, C.byref(batch), C.byref(ts),
                                                               C.byref(out0) if out0 is not None else None,
                                                               C.byref(out1), C.c_void_p(stream)))

    def solve_batch_device(self, batch: Batch, out: ResultC, stream: int = 0):
        self._check(self.lib.direct_ddp_solve_batch_device(self.h, C.byref(batch), C.byref(out), C.c_void_p(stream)))

    def time_allocation_device(self, B, N, start, end, seeds, max_vel, max_acc, durations, stream: int = 0):
        self._check(self.lib.direct_ddp_time_allocation_device(self.h, B, N, start, end, seeds, max_vel, max_acc,
                                                               durations, C.c_void_p(stream)))

    def sample(self, bez_coeff: np.ndarray, poly_time: np.ndarray, S: int):
        """Bernstein::getPos/getVel/getAcc of every segment at s_k = k/(S-1) (host buffers): pos, vel, acc [B][N][S][3]."""
        bez = np.ascontiguousarray(bez_coeff, dtype=np.float64)
        tim = np.ascontiguousarray(poly_time, dtype=np.float64)
        B, N = tim.shape
        out = [np.zeros((B, N, S, 3)) for _ in range(3)]
        self._check(self.lib.direct_ddp_sample(self.h, B, N, S, _ptr(bez), _ptr(tim), _ptr(out[0]), _ptr(out[1]), _ptr(out[2])))
        return tuple(out)

    def sample_device(self, B, N, S, bez, tim, pos, vel, acc, stream: int = 0):
        self._check(self.lib.direct_ddp_sample_device(self.h, B, N, S, bez, tim, pos, vel, acc, C.c_void_p(stream)))

    def replay(self, corridor: "Corridor", n_min: int = 2, n_max: int | None = None, stage0=None, stage1=None,
               time_power=TIME_POWER, max_vel: float = 2.0, max_acc: float = 2.0) -> np.ndarray:
        """corridorRecCallBack's comparison loop (teach_repeat_planner.cpp:309-350) as one ragged batch -> rows [n][12]."""
        n_max = corridor.N if n_max is None else n_max
        rows = np.zeros((n_max - n_min + 1, 12))
        ts = two_stage_opts(stage0, stage1, time_power)
        cs = corridor.c_struct()
        self._check(self.lib.direct_ddp_replay(self.h, C.byref(cs), n_min, n_max, C.byref(ts), max_vel, max_acc, _ptr(rows)))
        return rows

    def stats(self) -> Stats:
        s = Stats()
        self._check(self.lib.direct_ddp_last_stats(self.h, C.byref(s)))
        return s

    def fma_peak_tflops(self, precision: str) -> float:
        v = C.c_double(0.0)
        self._check(self.lib.direct_ddp_measure_fma_peak(self.h, {"fp64": 0, "fp32": 1}[precision], C.byref(v)))
        return v.value

    def trace(self, cap: int = 512):
        rows = (TraceRow * cap)()
        n = C.c_int(0)
        self._check(self.lib.direct_ddp_last_trace(self.h, rows, cap, C.byref(n)))
        return [{f: getattr(rows[k], f) for f, _ in TraceRow._fields_} for k in range(n.value)]


class Corridor:
    """A recorded corridor (msgs/msg/corridor.msg as readCorridorMsg flattens it, teach_repeat_planner.cpp:385-410)."""

    def __init__(self, path_id, planes, nplanes, center, seed):
        self.path_id = int(path_id)
        self.planes = np.ascontiguousarray(planes, dtype=np.float64)     # (N, P_max, 4)
        self.nplanes = np.ascontiguousarray(nplanes, dtype=np.int32)     # (N,)
        self.center = np.ascontiguousarray(center, dtype=np.float64)     # (N, 3)
        self.seed = np.ascontiguousarray(seed, dtype=np.float64)         # (N, 3)
        self.N, self.P_max = self.planes.shape[0], self.planes.shape[1]

    def c_struct(self) -> CorridorC:
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)
        return CorridorC(self.path_id, self.N, self.P_max, self.planes.ctypes.data_as(dp), self.nplanes.ctypes.data_as(ip),
                         self.center.ctypes.data_as(dp), self.seed.ctypes.data_as(dp))

    def write(self, path: str):
        cs = self.c_struct()
        st = load_library().direct_ddp_corridor_write(path.encode(), C.byref(cs))
        if st:
            raise DirectDdpError(f"direct_ddp_corridor_write failed with status {st}")

    @staticmethod
    def read(path: str) -> "Corridor":
        lib = load_library()
        p = C.POINTER(CorridorC)()
        st = lib.direct_ddp_corridor_read(path.encode(), C.byref(p))
        if st:
            raise DirectDdpError(f"direct_ddp_corridor_read failed with status {st}")
        c = p.contents
        N, PM = c.N, c.P_max
        out = Corridor(c.path_id, np.ctypeslib.as_array(c.planes, (N, PM, 4)).copy(), np.ctypeslib.as_array(c.nplanes, (N,)).copy(),
                       np.ctypeslib.as_array(c.center, (N, 3)).copy(), np.ctypeslib.as_array(c.seed, (N, 3)).copy())
        lib.direct_ddp_corridor_free(p)
        return out


def write_replay_rows(path: str, rows: np.ndarray):
    """The reference's result file, one "%d %f x11" line per prefix (teach_repeat_planner.cpp:347)."""
    rows = np.ascontiguousarray(rows, dtype=np.float64)
    st = load_library().direct_ddp_replay_write(path.encode(), _ptr(rows), rows.shape[0])
    if st:
        raise DirectDdpError(f"direct_ddp_replay_write failed with status {st}")
