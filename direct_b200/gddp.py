"""Generic unconstrained DDP (model (B) of SURVEY.md section 8(d)): problem generator and ctypes binding of the
direct_gddp_* entry points (include/direct_gddp.h).  BASELINE.json's literal "12-state / 4-input quadrotor" and "6-state
double integrator" - models the reference's DDP does not have, so there is no reference parity for
them; the checker is oracle/gddp_oracle.c ("parity unpinned")."""
from __future__ import annotations

import ctypes as C
import dataclasses

import numpy as np

from .problems import _Stream

MODEL_DINT6, MODEL_QUAD12 = 0, 1
QUAD_MASS, QUAD_G = 0.98, 9.81   # simulation/so3_quadrotor_simulator/src/dynamics/Quadrotor.cpp:15-20


@dataclasses.dataclass
class GddpProblem:
    model: int
    nx: int
    nu: int
    B: int
    N: int
    dt: float
    x0: np.ndarray          # (B, nx)
    xg: np.ndarray          # (B, nx)
    q: np.ndarray
    qf: np.ndarray
    r: np.ndarray
    uh: np.ndarray
    iter_max: int = 50
    tol: float = 1e-9
    u_init: np.ndarray | None = None

    def slice(self, lo, hi):
        return dataclasses.replace(self, B=hi - lo, x0=np.ascontiguousarray(self.x0[lo:hi]), xg=np.ascontiguousarray(self.xg[lo:hi]),
                                   u_init=None if self.u_init is None else np.ascontiguousarray(self.u_init[lo:hi]))


def make_quad_batch(B: int, N: int = 100, first: int = 0, dt: float = 0.05, workload: str = "hop") -> GddpProblem:
    """workload = "hop": hover-to-hover transfers: random start pose (small attitude, rates and velocity), goal 1.5-3.5 m away
    at rest (every one of the first 16384 problems converges in <= 16 iterations in the fp64 oracle).
    workload = "stated": the start / goal distribution BASELINE.md section 3 and SURVEY.md 8(d) state for the benchmark: start at
    rest in [-10, 10]^2 x [0.5, 2.5], goal at rest at radius U(15, 25) m, heading U(0, 2 pi), z ~ U(0.5, 1.8) (the reference's own
    generator, teach_repeat_planner.cpp:196-200, :215-217).  A 5 s horizon for a 15-25 m transfer: ~98 % of the problems converge,
    in ~17 iterations (fp64 oracle)."""
    if workload == "stated":
        rs = _Stream(np.arange(first, first + B, dtype=np.int64) + (3 << 40))
        x0 = np.zeros((B, 12)); xg = np.zeros((B, 12))
        x0[:, 0:3] = np.concatenate([rs.uniform(2, -10.0, 10.0), rs.uniform(1, 0.5, 2.5)], axis=1)
        th = rs.uniform(1, 0.0, 2.0 * np.pi)[:, 0]
        rad = rs.uniform(1, 15.0, 25.0)[:, 0]
        xg[:, 0] = x0[:, 0] + rad * np.cos(th); xg[:, 1] = x0[:, 1] + rad * np.sin(th); xg[:, 2] = rs.uniform(1, 0.5, 1.8)[:, 0]
        q = np.array([2.0] * 3 + [0.5] * 3 + [10.0] * 3 + [0.2] * 3)
        qf = np.array([100.0] * 3 + [10.0] * 3 + [50.0] * 3 + [2.0] * 3)
        r = np.array([1.0, 200.0, 200.0, 200.0])
        uh = np.array([QUAD_MASS * QUAD_G, 0.0, 0.0, 0.0])
        return GddpProblem(MODEL_QUAD12, 12, 4, B, N, dt, x0, xg, q, qf, r, uh)
    if workload != "hop":
        raise ValueError("workload must be 'hop' or 'stated'")
    rs = _Stream(np.arange(first, first + B, dtype=np.int64) + (1 << 40))
    x0 = np.zeros((B, 12)); xg = np.zeros((B, 12))
    x0[:, 0:3] = np.concatenate([rs.uniform(2, -3.0, 3.0), rs.uniform(1, 1.0, 2.0)], axis=1)
    x0[:, 3:6] = rs.uniform(3, -0.3, 0.3)
    x0[:, 6:9] = rs.uniform(3, -0.1, 0.1)
    x0[:, 9:12] = rs.uniform(3, -0.1, 0.1)
    d = rs.uniform(3, -1.0, 1.0)
    d /= np.linalg.norm(d, axis=1, keepdims=True) + 1e-9
    xg[:, 0:3] = x0[:, 0:3] + d * rs.uniform(1, 1.5, 3.5)
    xg[:, 2] = np.clip(xg[:, 2], 0.5, 3.0)
    q = np.array([2.0] * 3 + [0.5] * 3 + [10.0] * 3 + [0.2] * 3)
    qf = np.array([100.0] * 3 + [10.0] * 3 + [50.0] * 3 + [2.0] * 3)
    r = np.array([1.0, 200.0, 200.0, 200.0])
    uh = np.array([QUAD_MASS * QUAD_G, 0.0, 0.0, 0.0])
    return GddpProblem(MODEL_QUAD12, 12, 4, B, N, dt, x0, xg, q, qf, r, uh)


def make_dint_batch(B: int, N: int = 50, first: int = 0, dt: float = 0.1) -> GddpProblem:
    """BASELINE.json configs[0]: 3D double integrator (6-state point mass), quadratic cost, no constraints."""
    rs = _Stream(np.arange(first, first + B, dtype=np.int64) + (2 << 40))
    x0 = np.zeros((B, 6)); xg = np.zeros((B, 6))
    x0[:, 0:3] = rs.uniform(3, -10.0, 10.0)
    x0[:, 3:6] = rs.uniform(3, -1.0, 1.0)
    xg[:, 0:3] = rs.uniform(3, -10.0, 10.0)
    return GddpProblem(MODEL_DINT6, 6, 3, B, N, dt, x0, xg, np.array([1.0] * 3 + [0.1] * 3), np.array([100.0] * 3 + [10.0] * 3),
                       np.array([0.1] * 3), np.zeros(3))


# ---- ctypes binding of include/direct_gddp.h ------------------------------------------------------------------------------
GDDP_EXPORTS = ["direct_gddp_solve", "direct_gddp_solve_device"]


class ProblemC(C.Structure):
    _fields_ = [("model", C.c_int), ("B", C.c_int), ("N", C.c_int), ("iter_max", C.c_int), ("dt", C.c_double), ("tol", C.c_double),
                ("x0", C.c_void_p), ("xg", C.c_void_p), ("u_init", C.c_void_p),
                ("q", C.c_double * 12), ("qf", C.c_double * 12), ("r", C.c_double * 4), ("uh", C.c_double * 4)]


class ResultC(C.Structure):
    _fields_ = [("rtn", C.c_void_p), ("iters", C.c_void_p), ("cost", C.c_void_p), ("x", C.c_void_p), ("u", C.c_void_p),
                ("stats", C.c_void_p)]


class GddpResult:
    def __init__(self, B, N, nx, nu):
        self.rtn = np.zeros(B, np.int32); self.iters = np.zeros(B, np.int32); self.cost = np.zeros(B)
        self.x = np.zeros((B, N + 1, nx)); self.u = np.zeros((B, N, nu)); self.stats = np.zeros((B, 4), np.int64)

    def c_struct(self) -> ResultC:
        return ResultC(*[getattr(self, n).ctypes.data for n in ("rtn", "iters", "cost", "x", "u", "stats")])


def problem_struct(gp: GddpProblem, x0_ptr=None, xg_ptr=None, u_init_ptr=None) -> ProblemC:
    p = ProblemC(gp.model, gp.B, gp.N, gp.iter_max, gp.dt, gp.tol,
                 gp.x0.ctypes.data if x0_ptr is None else x0_ptr, gp.xg.ctypes.data if xg_ptr is None else xg_ptr,
                 (None if gp.u_init is None else gp.u_init.ctypes.data) if u_init_ptr is None else u_init_ptr)
    for k in range(gp.nx):
        p.q[k], p.qf[k] = gp.q[k], gp.qf[k]
    for k in range(gp.nu):
        p.r[k], p.uh[k] = gp.r[k], gp.uh[k]
    return p


def solve(solver, gp: GddpProblem, out: GddpResult | None = None) -> GddpResult:
    """Host-buffer call through the C-ABI on `solver` (a direct_b200.capi.Solver: device, precision, stream).  `out`: a result whose
    arrays the caller owns (e.g. views of pinned memory); a fresh pageable one is allocated otherwise."""
    lib = solver.lib
    lib.direct_gddp_solve.argtypes = [C.c_void_p, C.POINTER(ProblemC), C.POINTER(ResultC)]
    if out is None:
        out = GddpResult(gp.B, gp.N, gp.nx, gp.nu)
    p, o = problem_struct(gp), out.c_struct()
    solver._check(lib.direct_gddp_solve(solver.h, C.byref(p), C.byref(o)))
    return out


def solve_device(solver, p: ProblemC, o: ResultC, stream: int = 0):
    lib = solver.lib
    lib.direct_gddp_solve_device.argtypes = [C.c_void_p, C.POINTER(ProblemC), C.POINTER(ResultC), C.c_void_p]
    solver._check(lib.direct_gddp_solve_device(solver.h, C.byref(p), C.byref(o), stream))


def bwd_flops_per_knot(nx: int, nu: int) -> float:
    """Algorithmic (dense) flops of one knot of the backward sweep, SURVEY.md section 8(d) with m_c = 0:
    (12, 4) -> 13 397, (6, 3) -> 2 295."""
    n, m = nx, nu
    return 4 * n**3 + 8 * n * n * m + 4 * n * m * m + 2 * n * n + 8 * n * m + m**3 / 3.0 + 2 * m * m * (n + 1)
