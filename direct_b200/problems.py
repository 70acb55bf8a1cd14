"""Deterministic synthetic corridor problems for the batched IPDDP solver.

Every trajectory ``b`` draws from a counter-based splitmix64 stream seeded with
``0xD12EC7 ^ b``, so any subset of a batch can be regenerated anywhere
(CPU oracle, any GPU rank) byte-identically without communication.

Geometry mirrors the reference's own random tour (teach_repeat_planner.cpp:182-251): waypoints
on a ring of radius U(15,25) m at increasing heading, z ~ U(0.5,1.8); the corridor is a chain of
N overlapping convex cells along that polyline:

  * ``kind="box"``   axis-aligned boxes, P = 6 planes per cell
  * ``kind="poly"``  box + U{0..8} random cutting planes, P <= 14 (ragged; padded with the
                     always-inactive plane (0,0,0,-1), the trick teach_repeat_planner.cpp:867-879 uses)
  * ``kind="poly40"`` box + U{0..34} cutting planes, P <= 40: polytopes with more planes than a warp has lanes
                     (cdd H-representations have no bound, poly_utils.cpp:127-166)

Planes are unit outward normals with ``n.x + d <= 0`` inside (poly_utils.cpp:156-166).  Initial
segment times come from the reference's trapezoidal rule (teach_repeat_planner.cpp:583-639).
"""
from __future__ import annotations

import dataclasses

import numpy as np

SEED_BASE = 0xD12EC7
_GOLD = np.uint64(0x9E3779B97F4A7C15)
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)


def _splitmix(state: np.ndarray) -> np.ndarray:
    z = state.copy()
    z ^= z >> np.uint64(30)
    z *= _M1
    z ^= z >> np.uint64(27)
    z *= _M2
    z ^= z >> np.uint64(31)
    return z


class _Stream:
    """Per-trajectory counter-based uniform stream (vectorised over the batch)."""

    def __init__(self, ids: np.ndarray):
        self.seed = (np.uint64(SEED_BASE) ^ ids.astype(np.uint64)).astype(np.uint64)
        self.k = 0

    def uniform(self, n: int, lo: float = 0.0, hi: float = 1.0) -> np.ndarray:
        """(B, n) doubles in [lo, hi)."""
        ctr = np.arange(self.k + 1, self.k + n + 1, dtype=np.uint64)
        self.k += n
        with np.errstate(over="ignore"):
            st = self.seed[:, None] + ctr[None, :] * _GOLD
            z = _splitmix(st)
        u = (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
        return lo + (hi - lo) * u


@dataclasses.dataclass
class ProblemBatch:
    """Flat host arrays in the layout of the C-ABI ``direct_ddp_batch`` (include/direct_ddp.h)."""

    B: int
    N: int
    P_max: int
    planes: np.ndarray      # (B, N, P_max, 4) f64
    nplanes: np.ndarray     # (B, N) i32
    durations: np.ndarray   # (B, N) f64
    seeds: np.ndarray       # (B, N, 3) f64
    x0: np.ndarray          # (B, 9) f64  [pos, vel, acc]
    xd: np.ndarray          # (B, 9) f64
    max_vel: float = 2.0
    max_acc: float = 2.0

    def slice(self, lo: int, hi: int) -> "ProblemBatch":
        return ProblemBatch(hi - lo, self.N, self.P_max, self.planes[lo:hi], self.nplanes[lo:hi],
                            self.durations[lo:hi], self.seeds[lo:hi], self.x0[lo:hi], self.xd[lo:hi],
                            self.max_vel, self.max_acc)


def time_allocation(points: np.ndarray, max_vel: float, max_acc: float) -> np.ndarray:
    """teach_repeat_planner.cpp:583-639 with v0 = 0; points (B, N+1, 3) -> durations (B, N)."""
    d = np.linalg.norm(points[:, 1:] - points[:, :-1], axis=-1)
    acct = max_vel / max_acc
    accd = max_acc * acct * acct / 2
    dcct = max_vel / max_acc
    dccd = max_acc * dcct * dcct / 2
    t2 = np.sqrt(max_acc * d) / max_acc
    short = t2 + (max_acc * t2) / max_acc
    long_ = acct + (d - accd - dccd) / max_vel + dcct
    return np.where(d < accd + dccd, short, long_)


def make_batch(B: int, N: int, kind: str = "box", first: int = 0, max_vel: float = 2.0,
               max_acc: float = 2.0) -> ProblemBatch:
    """Problems ``first .. first+B-1`` of the infinite deterministic family."""
    if kind not in ("box", "poly", "poly40"):
        raise ValueError("kind must be 'box', 'poly' or 'poly40'")
    max_cuts = {"box": 0, "poly": 8, "poly40": 34}[kind]
    ids = np.arange(first, first + B, dtype=np.int64)
    rs = _Stream(ids)
    # --- tour waypoints (ring goals) --------------------------------------------------------
    n_leg = int(np.ceil(N * 1.05 / 7.0)) + 2
    start = np.concatenate([rs.uniform(2, -10.0, 10.0), rs.uniform(1, 0.5, 2.5)], axis=1)  # (B,3)
    rad = rs.uniform(n_leg, 15.0, 25.0)
    dth = rs.uniform(n_leg, 0.5, 2.5)
    zz = rs.uniform(n_leg, 0.5, 1.8)
    th = 1.25 * np.pi + np.cumsum(dth, axis=1)
    goals = np.stack([rad * np.cos(th), rad * np.sin(th), zz], axis=-1)  # (B,n_leg,3)
    way = np.concatenate([start[:, None, :], goals], axis=1)            # (B,n_leg+1,3)
    leg_vec = way[:, 1:] - way[:, :-1]
    leg_len = np.linalg.norm(leg_vec, axis=-1)
    cum = np.concatenate([np.zeros((B, 1)), np.cumsum(leg_len, axis=1)], axis=1)  # (B,n_leg+1)
    # --- cells ---------------------------------------------------------------------------------
    w = rs.uniform(N + 1, 0.6, 1.5)                     # half-widths (w[N] only spaces the end point)
    frac = rs.uniform(N, 0.4, 0.7)
    step = frac * np.minimum(w[:, :-1], w[:, 1:])       # spacing path point i -> i+1
    arc = np.concatenate([np.zeros((B, 1)), np.cumsum(step, axis=1)], axis=1)  # (B,N+1)
    leg = (arc[:, :, None] >= cum[:, None, 1:]).sum(-1)  # (B,N+1) leg index
    leg = np.minimum(leg, n_leg - 1)
    bi = np.arange(B)[:, None]
    t_in = (arc - cum[bi, leg]) / leg_len[bi, leg]
    pts = way[bi, leg] + t_in[..., None] * leg_vec[bi, leg]  # (B,N+1,3) path points
    seeds = pts[:, :N].copy()
    jit = rs.uniform(3 * N, -0.2, 0.2).reshape(B, N, 3)
    centre = seeds + jit * w[:, :N, None]
    P_max = 6 + max_cuts
    planes = np.zeros((B, N, P_max, 4))
    planes[..., 3] = -1.0  # inactive padding: c = -1 always
    for a in range(3):
        planes[:, :, 2 * a, a] = 1.0
        planes[:, :, 2 * a, 3] = -(centre[:, :, a] + w[:, :N])
        planes[:, :, 2 * a + 1, a] = -1.0
        planes[:, :, 2 * a + 1, 3] = centre[:, :, a] - w[:, :N]
    nplanes = np.full((B, N), 6, dtype=np.int32)
    if max_cuts:
        mc = max_cuts
        ncut = np.floor(rs.uniform(N, 0.0, mc + 1.0)).astype(np.int32).clip(0, mc)
        dirs = rs.uniform(N * mc * 3, -1.0, 1.0).reshape(B, N, mc, 3)
        rho = rs.uniform(N * mc, 0.8, 1.1).reshape(B, N, mc)
        nrm = np.linalg.norm(dirs, axis=-1, keepdims=True)
        nrm = np.where(nrm < 1e-3, 1.0, nrm)
        dirs = dirs / nrm
        for k in range(mc):
            on = ncut > k
            n = dirs[:, :, k]
            # plane through seed + rho*w*n with outward normal n
            d = -(np.einsum("bna,bna->bn", n, seeds) + rho[:, :, k] * w[:, :N])
            cut = np.concatenate([n, d[..., None]], axis=-1)
            planes[:, :, 6 + k] = np.where(on[..., None], cut, planes[:, :, 6 + k])
        nplanes = (6 + ncut).astype(np.int32)
    durations = time_allocation(pts, max_vel, max_acc)
    x0 = np.zeros((B, 9))
    xd = np.zeros((B, 9))
    x0[:, :3] = pts[:, 0]
    xd[:, :3] = pts[:, N]
    return ProblemBatch(B, N, P_max, np.ascontiguousarray(planes), np.ascontiguousarray(nplanes),
                        np.ascontiguousarray(durations), np.ascontiguousarray(seeds), x0, xd, max_vel, max_acc)


# Solver weights of the reference node, global_planner/launch/global_planner.launch:61-70.
STAGE0 = dict(w_snap=1.0, w_terminal=1.0, w_time=1.0, iter_max=50)
STAGE1 = dict(w_snap=1.0, w_terminal=100.0, w_time=20.0, iter_max=100)
TIME_POWER = 2
