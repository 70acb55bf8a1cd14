"""Generic unconstrained DDP, model (B) of SURVEY.md section 8(d): BASELINE.json's 12-state/4-input quadrotor and 6-state
double integrator.  The reference has neither model, so parity is UNPINNED: the CPU tests check the oracle on its own
(Jacobians vs central differences, LQ convergence, Riccati optimality, monotone cost), the GPU tests check the CUDA path
against that oracle through the C-ABI (1e-5 in fp64, 1e-3 in fp32 - the tolerances of BASELINE.json's north_star)."""
import ctypes as C
import dataclasses
import os
import re

import numpy as np
import pytest

from conftest import ROOT
from direct_b200 import gddp
from direct_b200.gddp import make_dint_batch, make_quad_batch


@pytest.fixture(scope="module")
def G():
    from oracle import gddp_py
    gddp_py.build()
    return gddp_py


def test_gddp_header_symbols_exported_and_struct_sizes(tmp_path):
    from direct_b200 import capi
    lib = capi.load_library()
    hdr = open(os.path.join(ROOT, "include", "direct_gddp.h")).read()
    declared = sorted(set(re.findall(r"\b(direct_gddp_[a-z_0-9]+)\s*\(", hdr)))
    assert set(declared) == set(gddp.GDDP_EXPORTS)
    for name in declared:
        assert hasattr(lib, name)
    import subprocess
    src = tmp_path / "s.c"
    src.write_text('#include <stdio.h>\n#include "direct_gddp.h"\nint main(void){printf("%zu %zu\\n", sizeof(direct_gddp_problem), sizeof(direct_gddp_result));return 0;}\n')
    exe = str(tmp_path / "s")
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", exe, str(src)])
    a, b = map(int, subprocess.check_output([exe]).split())
    assert a == C.sizeof(gddp.ProblemC) and b == C.sizeof(gddp.ResultC)


@pytest.mark.parametrize("model,nx,nu", [(1, 12, 4), (0, 6, 3)])
def test_oracle_jacobians_match_central_differences(G, model, nx, nu):
    rng = np.random.default_rng(3)
    for _ in range(5):
        x = rng.uniform(-0.6, 0.6, nx); u = rng.uniform(-0.5, 0.5, nu)
        if model == 1:
            u[0] += 9.6; u[1:] *= 0.05
        f, F, Gm = G.model(model, x, u)
        eps = 1e-6
        for j in range(nx):
            d = np.zeros(nx); d[j] = eps
            assert np.allclose((G.model(model, x + d, u)[0] - G.model(model, x - d, u)[0]) / (2 * eps), F[:, j], atol=2e-8)
        for j in range(nu):
            d = np.zeros(nu); d[j] = eps
            assert np.allclose((G.model(model, x, u + d)[0] - G.model(model, x, u - d)[0]) / (2 * eps), Gm[:, j], atol=2e-8)


def test_oracle_double_integrator_is_solved_in_one_iteration_and_is_optimal(G):
    """BASELINE.json configs[0]: single trajectories, 50 knots, 3D double integrator.  Linear-quadratic, so one Newton
    step is exact; the result must satisfy the optimality of the equivalent least-squares problem."""
    gp = make_dint_batch(4, 50)
    r = G.solve_batch(gp)
    assert (r.rtn == 1).all() and (r.iters <= 2).all()
    # brute force: J is quadratic in U = (u_0 .. u_{N-1}); build x_i = Phi_i x0 + Gam_i U and solve the normal equations
    N, dt = gp.N, gp.dt
    A = np.eye(6); A[:3, 3:] = dt * np.eye(3)
    Bm = np.zeros((6, 3)); Bm[3:, :] = dt * np.eye(3)
    for b in range(gp.B):
        Phi = [np.eye(6)]; Gam = [np.zeros((6, 3 * N))]
        for i in range(N):
            Gn = A @ Gam[-1]; Gn[:, 3 * i:3 * i + 3] += Bm
            Phi.append(A @ Phi[-1]); Gam.append(Gn)
        H = np.zeros((3 * N, 3 * N)); g = np.zeros(3 * N)
        for i in range(N + 1):
            W = np.diag(gp.qf) if i == N else dt * np.diag(gp.q)
            e0 = Phi[i] @ gp.x0[b] - gp.xg[b]
            H += Gam[i].T @ W @ Gam[i]; g += Gam[i].T @ W @ e0
        H += dt * np.kron(np.eye(N), np.diag(gp.r))
        U = np.linalg.solve(H, -g)
        assert np.allclose(r.u[b].ravel(), U, rtol=1e-7, atol=1e-8)


def test_oracle_quadrotor_converges_and_fp32_agrees(G):
    gp = make_quad_batch(96)
    r = G.solve_batch(gp, nthreads=4)
    assert (r.rtn == 1).mean() > 0.97
    ok = (r.rtn == 1) & (np.abs(r.x[:, :, 6:9]).max((1, 2)) < 1.2)
    assert ok.mean() > 0.95
    assert np.abs(r.x[ok, -1, :3] - gp.xg[ok, :3]).max() < 0.2          # reaches the goal
    open_loop = G.solve_batch(dataclasses.replace(gp, iter_max=0))
    assert (r.cost[ok] < open_loop.cost[ok]).all()                      # and pays less than hovering in place
    g32 = dataclasses.replace(gp, tol=1e-5)
    a, b = G.solve_batch(g32, nthreads=4), G.solve_batch(g32, fp32=True, nthreads=4)
    both = ok & (a.rtn == 1) & (b.rtn == 1)
    assert both.mean() > 0.9
    assert np.max(np.abs(a.cost[both] - b.cost[both]) / a.cost[both]) < 1e-3


def _screen(r, tol_rpy=1.2):
    """Trajectories the comparison is meaningful on: converged, away from the Euler-angle singularity."""
    return (r.rtn == 1) & (np.abs(r.x[:, :, 6:9]).max((1, 2)) < tol_rpy)


@pytest.mark.gpu
def test_gpu_quadrotor_fp64_matches_oracle(G):
    from direct_b200.capi import Solver
    gp = make_quad_batch(512)
    a = G.solve_batch(gp, nthreads=G_threads())
    s = Solver(0, "fp64")
    g = gddp.solve(s, gp)
    st = s.stats()
    s.close()
    assert st.kernel_launches == 1 and st.kernel_ms > 0
    ok = _screen(a)
    assert ok.mean() > 0.97
    same = ok & (g.rtn == a.rtn) & (g.iters == a.iters)
    assert same.mean() > 0.95                      # identical decisions (iteration counts) on nearly all of them
    rel = np.abs(g.cost[same] - a.cost[same]) / np.abs(a.cost[same])
    assert rel.max() < 1e-5                        # north_star: <= 1e-5 relative in fp64
    assert np.abs(g.x[same] - a.x[same]).max() < 1e-5 * max(1.0, np.abs(a.x[same]).max())
    assert np.abs(g.u[same] - a.u[same]).max() < 1e-5 * max(1.0, np.abs(a.u[same]).max())
    assert (g.stats[same, 0] == a.stats[same, 0]).all() and (g.stats[same, 1] == a.stats[same, 1]).all()
    assert np.isfinite(g.cost).all()


@pytest.mark.gpu
def test_gpu_quadrotor_fp32_within_stated_tolerance(G):
    from direct_b200.capi import Solver
    gp = dataclasses.replace(make_quad_batch(512), tol=1e-5)
    a = G.solve_batch(gp, nthreads=G_threads())      # fp64 oracle at the same stopping tolerance
    s = Solver(0, "fp32")
    g = gddp.solve(s, gp)
    s.close()
    ok = _screen(a) & (g.rtn == 1)
    assert ok.mean() > 0.95
    rel = np.abs(g.cost[ok] - a.cost[ok]) / np.abs(a.cost[ok])
    assert rel.max() < 1e-3                        # north_star: <= 1e-3 in fp32
    assert np.quantile(np.abs(g.x[ok] - a.x[ok]).max((1, 2)), 0.99) < 5e-2


@pytest.mark.gpu
def test_gpu_double_integrator_config0(G):
    from direct_b200.capi import Solver
    gp = make_dint_batch(64, 50)
    a = G.solve_batch(gp)
    for prec, tol in (("fp64", 1e-9), ("fp32", 2e-4)):
        s = Solver(0, prec)
        g = gddp.solve(s, gp if prec == "fp64" else dataclasses.replace(gp, tol=1e-5))
        s.close()
        assert (g.rtn == 1).all()
        assert np.max(np.abs(g.cost - a.cost) / a.cost) < max(tol, 1e-9)
        assert np.abs(g.u - a.u).max() < (1e-7 if prec == "fp64" else 5e-3) * max(1.0, np.abs(a.u).max())


def G_threads():
    return max(1, len(os.sched_getaffinity(0)))


@pytest.mark.gpu
def test_gpu_pair_kernel_ragged_batches_and_single_kernel_agree(G, monkeypatch):
    """Two trajectories share a warp (gddp_pair.cuh): odd batches leave a half-warp without a trajectory, a batch of one leaves a whole
    half idle, and a pair's halves finish at different iterations.  Every batch size must give, trajectory for trajectory, the
    decisions of the one-trajectory-per-warp kernel (DIRECT_GDDP_PAIR=0; same arithmetic except that the gains are multiplied by the
    pivot's rsqrt instead of divided by its square root) and its numbers to 1e-9, results that do not depend on the partner (bit for
    bit), and what the oracle gives within 1e-5."""
    from direct_b200.capi import Solver
    full = make_quad_batch(67, 60)
    u0 = np.tile(np.array([0.98 * 9.81, 0.0, 0.0, 0.0]), (67, 60, 1)) + 0.01 * np.sin(np.arange(67 * 60 * 4)).reshape(67, 60, 4)
    full = dataclasses.replace(full, u_init=np.ascontiguousarray(u0))     # warm-start controls: the u_init path of both kernels
    a = G.solve_batch(full, nthreads=G_threads())
    assert len(set(a.iters.tolist())) > 2                                  # partners do finish at different iterations
    s = Solver(0, "fp64")
    res = {}
    for B in (1, 2, 3, 66, 67):
        gp = full.slice(0, B)
        monkeypatch.setenv("DIRECT_GDDP_PAIR", "1")
        p = gddp.solve(s, gp)
        monkeypatch.setenv("DIRECT_GDDP_PAIR", "0")
        q = gddp.solve(s, gp)
        assert np.array_equal(p.rtn, q.rtn) and np.array_equal(p.iters, q.iters)
        assert np.array_equal(p.stats[:, :3], q.stats[:, :3])             # sweeps, rollouts, backward knots
        for f in ("cost", "x", "u"):
            assert np.abs(getattr(p, f) - getattr(q, f)).max() <= 1e-9 * max(1.0, np.abs(getattr(q, f)).max()), (B, f)
        res[B] = p
    monkeypatch.delenv("DIRECT_GDDP_PAIR")
    s.close()
    g = res[67]
    for B in (1, 2, 3, 66):                                                # a trajectory's result does not depend on its partner
        for f in ("rtn", "iters", "cost", "x", "u"):
            assert np.array_equal(getattr(res[B], f), getattr(g, f)[:B]), (B, f)
    ok = _screen(a)
    same = ok & (g.rtn == a.rtn) & (g.iters == a.iters)
    assert same.mean() > 0.9
    assert (np.abs(g.cost[same] - a.cost[same]) / np.abs(a.cost[same])).max() < 1e-5
