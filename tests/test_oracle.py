"""CPU tests of the parity oracle (oracle/ipddp_oracle.c): pinned against the fixtures produced by the
reference's own translation unit, against that translation unit live when it is built, and against
algorithm invariants.  No GPU needed."""
import numpy as np
import pytest

from conftest import GOLDEN_TWO_STAGE, OUT_FIELDS, load_golden, rel_err
from direct_b200.problems import STAGE0, STAGE1, make_batch


def _stage_kwargs(d):
    ov = dict(minvo=int(d["minvo"]), time_power=int(d["time_power"]))
    return dict(STAGE0, **ov), dict(STAGE1, **ov)


@pytest.mark.parametrize("name", GOLDEN_TWO_STAGE)
def test_oracle_matches_reference_fixtures(oracle, name):
    pb, d = load_golden(name)
    s0, s1 = _stage_kwargs(d)
    r0 = oracle.solve_batch(pb, nthreads=2, infeas=1, zero_init=1, **s0)
    assert (r0.rtn == d["s0_rtn"]).all() and (r0.iters == d["s0_iters"]).all()
    assert (r0.infeas_out == d["s0_infeas_out"]).all()
    for f in OUT_FIELDS:
        assert rel_err(getattr(r0, f), d["s0_" + f]) < 1e-9, f
    assert rel_err(r0.jerk.sum(1), d["s0_jerk_sum"]) < 1e-9
    assert rel_err(((r0.x_final - pb.xd) ** 2).sum(1), d["s0_terminal_norm"]) < 1e-9
    # stage 1 from the REFERENCE's stage-0 output (the fixture), so both sides start identically
    r1 = oracle.solve_batch(pb, nthreads=2, infeas=d["s0_infeas_out"], zero_init=0, init_bez=d["s0_bez_coeff"],
                            durations=d["dur1"], **s1)
    assert (r1.rtn == d["s1_rtn"]).all() and (r1.iters == d["s1_iters"]).all()
    for f in OUT_FIELDS:
        assert rel_err(getattr(r1, f), d["s1_" + f]) < 1e-9, f
    assert rel_err(r1.jerk.sum(1), d["s1_jerk_sum"]) < 1e-9


def test_oracle_line_init_fixture(oracle):
    pb, d = load_golden("box_n6_lineinit")
    r = oracle.solve_batch(pb, nthreads=1, infeas=1, zero_init=0, line_init=1, w_snap=1.0, w_terminal=100.0,
                           w_time=50.0, iter_max=60)
    assert (r.rtn == d["s1_rtn"]).all() and (r.iters == d["s1_iters"]).all()
    assert (r.line_failed_out == d["s1_line_failed_out"]).all()
    for f in OUT_FIELDS:
        assert rel_err(getattr(r, f), d["s1_" + f]) < 1e-9, f


def test_oracle_two_stage_equals_two_calls(oracle):
    pb = make_batch(5, 9, "poly", first=77)
    a0, a1 = oracle.two_stage_batch(pb, nthreads=2)
    b0 = oracle.solve_batch(pb, infeas=1, zero_init=1, **STAGE0)
    dur = np.where((b0.rtn == 2)[:, None], b0.poly_time, pb.durations)
    b1 = oracle.solve_batch(pb, infeas=b0.infeas_out, zero_init=0, init_bez=b0.bez_coeff, durations=dur, **STAGE1)
    assert (a0.rtn == b0.rtn).all() and (a1.rtn == b1.rtn).all()
    assert np.array_equal(a1.poly_coeff, b1.poly_coeff) and np.array_equal(a1.cost, b1.cost)


def test_oracle_vs_reference_live(oracle):
    """Only where oracle/_ref was built (the dev container or a box it travelled to)."""
    if not oracle.ref_available():
        pytest.skip("oracle/_ref/libddp_ref.so not built")
    for kind, N, first in (("box", 7, 11), ("poly", 16, 23)):
        pb = make_batch(4, N, kind, first=first)
        a = oracle.solve_batch(pb, infeas=1, zero_init=1, **STAGE0)
        r = oracle.solve_batch(pb, use_ref=True, infeas=1, zero_init=1, **STAGE0)
        assert (a.rtn == r.rtn).all() and (a.iters == r.iters).all()
        for f in OUT_FIELDS:
            assert rel_err(getattr(a, f), getattr(r, f)) < 1e-9


def test_converged_trajectory_is_feasible_and_connected(oracle):
    """Invariants of an accepted solution: every control point inside its polytope, velocity/acceleration
    control points within limits, positive segment times >= 0.3, C2 continuity between segments."""
    pb = make_batch(6, 14, "poly", first=5)
    _, r = oracle.two_stage_batch(pb, nthreads=2)
    assert np.isin(r.rtn, (0, 1)).all() and (r.rtn == 1).sum() >= 4  # interior-point iterates stay feasible
    assert (r.poly_time > 0.3).all()
    bez = r.bez_coeff.reshape(pb.B, pb.N, 3, 6) * r.poly_time[:, :, None, None]  # control points (positions)
    for b in range(pb.B):
        for i in range(pb.N):
            P = pb.nplanes[b, i]
            pl = pb.planes[b, i, :P]
            val = pl[:, :3] @ bez[b, i] + pl[:, 3:4]
            assert (val < 1e-9).all()
    # continuity: x_{i+1} = end of segment i  <=>  poly_coeff[i+1][0:9] equals the derivatives of segment i at T
    pc = r.poly_coeff.reshape(pb.B, pb.N, 6, 3)
    T = r.poly_time
    for k, fac in ((0, [1, 1, 1, 1, 1, 1]), (1, [0, 1, 2, 3, 4, 5]), (2, [0, 0, 1, 3, 6, 10])):
        end = sum(fac[l] * pc[:, :-1, l] * T[:, :-1, None] ** (l - k) for l in range(k, 6))
        assert rel_err(end, pc[:, 1:, k]) < 1e-9


def test_bezier_monomial_round_trip(oracle):
    """poly2bez(bez2poly(B)) == B: feeding a solution's Bezier back in as warm start with iter_max = 0
    reproduces the same high-order coefficients (ddp_optimizer.cpp:782-812)."""
    pb = make_batch(3, 6, "box", first=900)
    _, r = oracle.two_stage_batch(pb)
    again = oracle.solve_batch(pb, infeas=0, zero_init=0, init_bez=r.bez_coeff, durations=r.poly_time,
                               **dict(STAGE1, iter_max=0))
    assert rel_err(again.poly_coeff, r.poly_coeff) < 1e-9
    assert rel_err(again.bez_coeff, r.bez_coeff) < 1e-9
    assert (again.iters == 0).all()


def test_time_allocation_matches_numpy(oracle):
    import ctypes as C
    from direct_b200.problems import time_allocation
    rng = np.random.default_rng(3)
    N = 9
    pts = np.cumsum(rng.uniform(0.1, 3.0, size=(1, N + 1, 3)), axis=1)
    want = time_allocation(pts, 2.0, 2.0)[0]
    got = np.zeros(N)
    dp = C.POINTER(C.c_double)
    start, end, seeds = pts[0, 0].copy(), pts[0, N].copy(), np.ascontiguousarray(pts[0, :N])
    oracle.lib().ipddp_oracle_time_allocation(N, start.ctypes.data_as(dp), end.ctypes.data_as(dp),
                                              seeds.ctypes.data_as(dp), 2.0, 2.0, got.ctypes.data_as(dp))
    assert rel_err(got, want) < 1e-14


def test_rejects_undefined_time_power(oracle):
    pb = make_batch(1, 3, "box")
    with pytest.raises(RuntimeError):
        oracle.solve_batch(pb, infeas=1, zero_init=1, **dict(STAGE0, time_power=3))


def test_edge_cases_single_knot_and_unreachable_goal(oracle):
    # N = 1: one segment
    pb = make_batch(2, 1, "box", first=4)
    r0, r1 = oracle.two_stage_batch(pb)
    assert np.isfinite(r1.cost).all() and (r1.iters <= 100).all()
    # goal far outside the last polytope: stage 1 cannot satisfy everything, must still terminate cleanly
    pb = make_batch(2, 4, "box", first=8)
    pb.xd[:, :3] += 50.0
    r0, r1 = oracle.two_stage_batch(pb)
    assert set(np.unique(r1.rtn)).issubset({0, 1, 2, -3, -4})
    assert (r1.iters <= 100).all()


def test_bezier_sampling_oracle_endpoints_and_derivatives(oracle):
    """The numpy restatement of Bernstein::getPos/getVel/getAcc (bezier_base.h:77-115): end points are control points
    0 and 5, and velocity / acceleration are the finite-difference derivatives of position in time."""
    rng = np.random.default_rng(3)
    bez = rng.normal(size=(2, 3, 18)); T = rng.uniform(0.5, 2.0, size=(2, 3))
    S = 2001
    pos, vel, acc = oracle.bezier_sample(bez, T, S)
    cp = bez.reshape(2, 3, 3, 6) * T[:, :, None, None]
    assert np.allclose(pos[:, :, 0], cp[..., 0]) and np.allclose(pos[:, :, -1], cp[..., 5])
    dt = T[:, :, None, None] / (S - 1)
    inner = slice(1, -1)   # central differences; the one-sided end points are only first-order accurate
    assert np.allclose((np.gradient(pos, axis=2) / dt)[:, :, inner], vel[:, :, inner], rtol=0, atol=1e-4 * np.abs(vel).max())
    assert np.allclose((np.gradient(vel, axis=2) / dt)[:, :, inner], acc[:, :, inner], rtol=0, atol=1e-4 * np.abs(acc).max())


def test_bezier_restatement_matches_the_references_own_evaluators(oracle):
    """oracle_py.bezier_sample (numpy) against tests/golden/bezier_ref.npz, the output of the reference's own
    Bernstein::getPos / getVel / getAcc (utils/bezier_base.h:77-115, compiled unmodified: `make -C oracle bezier`,
    tests/golden/make_bezier_golden.py), and against that library itself where it is built."""
    import os
    from conftest import GOLDEN
    d = np.load(os.path.join(GOLDEN, "bezier_ref.npz"))
    S = int(d["S"])
    got = oracle.bezier_sample(d["bez_coeff"], d["poly_time"], S)
    for g, name in zip(got, ("pos", "vel", "acc")):
        assert rel_err(g, d[name]) < 1e-14, name
    if oracle.bezier_ref_available():
        live = oracle.bezier_sample_ref(d["bez_coeff"], d["poly_time"], S)
        for g, name in zip(live, ("pos", "vel", "acc")):
            assert np.array_equal(g, d[name]), name
