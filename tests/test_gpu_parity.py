"""Parity of the CUDA path (through the C-ABI, host buffers) against the oracle, the reference fixtures and
size-independent properties at BASELINE.json's full size.  Needs a B200: run with -m gpu."""
import numpy as np
import pytest

from conftest import (GOLDEN_TWO_STAGE, OUT_FIELDS, SCREEN_FLOOR, SCREEN_FLOOR_FULL, assert_screened_out_bounded, assert_valid_result,
                      conditioned_mask_two_stage, load_golden, rel_err, results_differ)
from direct_b200.problems import STAGE0, STAGE1, make_batch

pytestmark = pytest.mark.gpu

TOL64 = 1e-5  # BASELINE.json north_star: "<= 1e-5 relative in fp64"
TOL32 = 1e-3  # "... and <= 1e-3 in fp32"


@pytest.fixture(scope="module")
def solver():
    from direct_b200.capi import Solver
    s = Solver(0, "fp64")
    yield s
    s.close()


@pytest.mark.parametrize("name", GOLDEN_TWO_STAGE)
def test_gpu_matches_reference_fixtures(solver, name):
    pb, d = load_golden(name)
    ov = dict(minvo=int(d["minvo"]), time_power=int(d["time_power"]))
    r0 = solver.solve_batch(pb, infeas=1, zero_init=1, **dict(STAGE0, **ov))
    assert (r0.rtn == d["s0_rtn"]).all() and (r0.iters == d["s0_iters"]).all()
    assert (r0.infeas_out == d["s0_infeas_out"]).all()
    for f in OUT_FIELDS:
        assert rel_err(getattr(r0, f), d["s0_" + f]) < TOL64, f
    r1 = solver.solve_batch(pb, infeas=d["s0_infeas_out"], zero_init=0, init_bez=d["s0_bez_coeff"],
                            durations=d["dur1"], **dict(STAGE1, **ov))
    assert (r1.rtn == d["s1_rtn"]).all() and (r1.iters == d["s1_iters"]).all()
    for f in OUT_FIELDS:
        assert rel_err(getattr(r1, f), d["s1_" + f]) < TOL64, f
    assert rel_err(r1.jerk.sum(1), d["s1_jerk_sum"]) < TOL64
    assert rel_err(((r1.x_final - pb.xd) ** 2).sum(1), d["s1_terminal_norm"]) < TOL64


# horizons around the 32-knot blocks of the knot-parallel phases and around the depth of the bulk-copy rings (2 and 3 knots) included,
# and the longest horizon of BASELINE.json's sweep (400 knots)
@pytest.mark.parametrize("kind,N,B", [("box", 1, 5), ("box", 2, 6), ("box", 3, 6), ("poly", 31, 12), ("box", 32, 12), ("box", 33, 64),
                                      ("poly", 50, 48), ("box", 64, 10), ("poly", 65, 10), ("box", 96, 12), ("poly", 97, 16), ("poly", 100, 32), ("box", 200, 8), ("poly", 400, 6)])
def test_gpu_two_stage_matches_oracle(solver, oracle, kind, N, B):
    pb = make_batch(B, N, kind, first=2000 + N)
    pert = []
    ok, (a0, a1) = conditioned_mask_two_stage(oracle, pb, keep=pert)      # see conftest.py "Conditioning screen"
    assert ok.mean() >= SCREEN_FLOOR
    g0, g1 = solver.solve_two_stage(pb)
    for a, g in ((a0, g0), (a1, g1)):
        bad = results_differ(g, a, TOL64, OUT_FIELDS + ("jerk", "x_final")) & ok
        assert not bad.any(), np.nonzero(bad)[0]
        assert np.array_equal(a.stats[ok, :4], g.stats[ok, :4])   # same sweeps / rollouts, knot for knot
        assert_valid_result(pb, g)
    assert_screened_out_bounded(oracle, pb, ok, (a0, a1), (g0, g1), perturbed=pert)   # set aside is not exempt


def test_gpu_fused_two_stage_equals_two_single_calls(solver):
    pb = make_batch(40, 25, "poly", first=31)
    g0, g1 = solver.solve_two_stage(pb)
    s0 = solver.solve_batch(pb, infeas=1, zero_init=1, **STAGE0)
    dur = np.where((s0.rtn == 2)[:, None], s0.poly_time, pb.durations)
    s1 = solver.solve_batch(pb, infeas=s0.infeas_out, zero_init=0, init_bez=s0.bez_coeff, durations=dur, **STAGE1)
    assert np.array_equal(g0.poly_coeff, s0.poly_coeff) and np.array_equal(g0.rtn, s0.rtn)
    assert np.array_equal(g1.rtn, s1.rtn) and np.array_equal(g1.iters, s1.iters)
    assert np.array_equal(g1.poly_coeff, s1.poly_coeff) and np.array_equal(g1.cost, s1.cost)


def test_gpu_batch_of_one_reproduces_batch_element(solver):
    pb = make_batch(12, 20, "box", first=600)
    _, g = solver.solve_two_stage(pb)
    for i in (0, 7, 11):
        _, one = solver.solve_two_stage(pb.slice(i, i + 1))
        assert np.array_equal(one.poly_coeff[0], g.poly_coeff[i]) and one.cost[0] == g.cost[i]


def test_gpu_full_size_properties(solver):
    """BASELINE configs[1] size (4096 x 100 knots): determinism, feasibility of every returned trajectory,
    C2 continuity, Bezier <-> monomial consistency.  The oracle is too slow for all of it; a sample is compared."""
    B, N = 4096, 100
    pb = make_batch(B, N, "box")
    _, g = solver.solve_two_stage(pb, want_stage0=False)
    _, g2 = solver.solve_two_stage(pb, want_stage0=False)
    assert np.array_equal(g.poly_coeff, g2.poly_coeff) and np.array_equal(g.rtn, g2.rtn)  # run-to-run identical
    assert np.isin(g.rtn, (0, 1)).all() and (g.rtn == 1).mean() > 0.95
    assert (g.poly_time > 0.3 - 2.0e-4).all()                          # time row -T + 0.3 - 2e-4 < 0 (ddp_optimizer.cpp:1279-1283)
    cp = g.bez_coeff.reshape(B, N, 3, 6) * g.poly_time[:, :, None, None]
    val = np.einsum("bnpa,bnaj->bnpj", pb.planes[..., :3], cp) + pb.planes[..., 3:4]
    # every control point of a converged solve (rtn 1) is inside its polytope relaxed by the reference's own
    # 2e-4 margin (c = n.p + d - 2e-4 < 0, ddp_optimizer.cpp:1281-1283); solves that run into iter_max (rtn 0)
    # carry no such guarantee in the reference either (ddp_optimizer.cpp:346-378)
    assert (val[g.rtn == 1] < 2.0e-4).all()
    pc = g.poly_coeff.reshape(B, N, 6, 3)
    T = g.poly_time
    for k, fac in ((0, [1, 1, 1, 1, 1, 1]), (1, [0, 1, 2, 3, 4, 5]), (2, [0, 0, 1, 3, 6, 10])):
        end = sum(fac[l] * pc[:, :-1, l] * T[:, :-1, None] ** (l - k) for l in range(k, 6))
        assert rel_err(end, pc[:, 1:, k]) < 1e-9                      # x_{i+1} = f(x_i, u_i)
    # Bezier control point 0 is the segment start, control point 5 its end
    assert rel_err(cp[:, :, :, 0], pc[:, :, 0]) < 1e-9


def test_gpu_tail_balancing_leaves_every_bit_alone(solver, monkeypatch):
    """Cooperation inside a CTA, the speculative line search and the speculative backward sweep only change
    WHO computes a phase: a batch far below the grid (every mechanism active from the start) returns the same bits and the same
    sweep / trial / knot counters with each of them switched off (tuning knobs DIRECT_DDP_SPEC / _GSPEC / _COOP)."""
    pb = make_batch(320, 60, "box")
    ref = None
    used = {}
    for knobs in ({}, {"DIRECT_DDP_SPEC": "0"}, {"DIRECT_DDP_GSPEC": "0"}, {"DIRECT_DDP_COOP": "0"}):
        for k in ("DIRECT_DDP_SPEC", "DIRECT_DDP_GSPEC", "DIRECT_DDP_COOP"):
            monkeypatch.delenv(k, raising=False)
        for k, v in knobs.items():
            monkeypatch.setenv(k, v)
        g0, g1 = solver.solve_two_stage(pb, want_stage0=True)
        st = solver.stats()
        used[tuple(knobs)] = (st.helper_units, st.spec_trials, st.spec_sweeps_used)
        if ref is None:
            ref = (g0, g1)
            continue
        for a, b in zip(ref, (g0, g1)):
            for f in ("rtn", "iters", "infeas_out") + OUT_FIELDS:
                assert np.array_equal(getattr(a, f), getattr(b, f)), (knobs, f)
            assert np.array_equal(a.stats[:, :4], b.stats[:, :4]), knobs
    assert min(used[()]) > 0, used                                    # all three mechanisms did run in the default configuration
    assert used[("DIRECT_DDP_SPEC",)][2] == 0 and used[("DIRECT_DDP_GSPEC",)][1] == 0 and used[("DIRECT_DDP_COOP",)] == (0, 0, 0)


def test_gpu_full_size_sample_against_oracle(solver, oracle):
    pb = make_batch(4096, 100, "box")
    g0, g = solver.solve_two_stage(pb, want_stage0=True)
    assert_valid_result(pb, g)
    idx = np.arange(0, 4096, 32)          # 128 of the 4096 trajectories
    sub = pb.slice(0, 4096)
    for f in ("planes", "nplanes", "durations", "seeds", "x0", "xd"):
        setattr(sub, f, np.ascontiguousarray(getattr(pb, f)[idx]))
    sub.B = len(idx)
    pert = []
    ok, (a0, a) = conditioned_mask_two_stage(oracle, sub, keep=pert)
    assert ok.mean() >= SCREEN_FLOOR_FULL
    gs = type("R", (), {f: getattr(g, f)[idx] for f in ("rtn", "iters", "infeas_out") + OUT_FIELDS})
    bad = results_differ(gs, a, TOL64) & ok
    assert not bad.any(), idx[bad]
    gs0 = type("R", (), {f: getattr(g0, f)[idx] for f in ("rtn", "iters", "infeas_out") + OUT_FIELDS})
    assert_screened_out_bounded(oracle, sub, ok, (a0, a), (gs0, gs), perturbed=pert)


def test_gpu_fp32_within_stated_tolerance(oracle):
    """fp32 arithmetic: the fraction of trajectories within 1e-3 of the fp64 oracle is reported by bench.py;
    here the easy regime (short horizons) must agree outright."""
    from direct_b200.capi import Solver
    s = Solver(0, "fp32")
    pb = make_batch(64, 10, "box", first=123)
    a0, a1 = oracle.two_stage_batch(pb, nthreads=oracle.max_threads())
    g0, g1 = s.solve_two_stage(pb)
    ok = (g1.rtn == a1.rtn) & (np.abs(g1.cost - a1.cost) <= TOL32 * np.abs(a1.cost))
    assert ok.mean() >= 0.9
    s.close()


def test_gpu_rejects_bad_arguments(solver):
    from direct_b200.capi import DirectDdpError
    pb = make_batch(2, 3, "box")
    with pytest.raises(DirectDdpError, match="time_power"):
        solver.solve_batch(pb, infeas=1, zero_init=1, **dict(STAGE0, time_power=3))
    import dataclasses
    with pytest.raises(DirectDdpError, match="line_init"):   # line_init reads the seeds (ddp_optimizer.cpp:195-247)
        solver.solve_batch(dataclasses.replace(pb, seeds=None), infeas=1, zero_init=0, line_init=1, **STAGE1)


@pytest.mark.parametrize("name", ["box_n5", "poly_n12", "box_n50_single"])
def test_gpu_dropin_translation_unit_matches_reference_fixtures(oracle, name):
    """direct_b200/host/ddp_optimizer_b200.cpp driven through the reference's own C++ class API
    (ddpTrajOptimizer::polyCurveGeneration + the inline getters of ddp_optimizer.h:299-340) by the same driver
    that produced the fixtures from the reference's translation unit (oracle/ref_driver.cpp)."""
    if not oracle.dropin_available():
        pytest.skip("oracle/_ref/libddp_dropin.so not built (needs the reference's headers: `make -C oracle dropin`)")
    pb, d = load_golden(name)
    ov = dict(minvo=int(d["minvo"]), time_power=int(d["time_power"]))
    r0 = oracle.solve_batch(pb, use_dropin=True, infeas=1, zero_init=1, **dict(STAGE0, **ov))
    assert (r0.rtn == d["s0_rtn"]).all() and (r0.iters == d["s0_iters"]).all() and (r0.infeas_out == d["s0_infeas_out"]).all()
    for f in OUT_FIELDS:
        assert rel_err(getattr(r0, f), d["s0_" + f]) < TOL64, f
    r1 = oracle.solve_batch(pb, use_dropin=True, infeas=d["s0_infeas_out"], zero_init=0, init_bez=d["s0_bez_coeff"],
                            durations=d["dur1"], **dict(STAGE1, **ov))
    assert (r1.rtn == d["s1_rtn"]).all() and (r1.iters == d["s1_iters"]).all()
    for f in OUT_FIELDS:
        assert rel_err(getattr(r1, f), d["s1_" + f]) < TOL64, f
    assert rel_err(r1.jerk[:, 0], d["s1_jerk_sum"]) < TOL64               # getJerkCost()
    assert rel_err(r1.x_final[:, 0], d["s1_terminal_norm"]) < TOL64       # getTerminalNorm()


def test_gpu_bezier_sampling_matches_the_references_own_evaluators(solver):
    """bezier_sample_kernel against the fixture produced by the reference's own Bernstein::getPos /
    getVel / getAcc (utils/bezier_base.h:77-115 compiled unmodified, tests/golden/make_bezier_golden.py)."""
    import os
    from conftest import GOLDEN
    d = np.load(os.path.join(GOLDEN, "bezier_ref.npz"))
    pos, vel, acc = solver.sample(d["bez_coeff"], d["poly_time"], int(d["S"]))
    for g, name in zip((pos, vel, acc), ("pos", "vel", "acc")):
        assert rel_err(g, d[name]) < 1e-12, name


def test_gpu_bezier_sampling_matches_reference_formulas(solver, oracle):
    """batched Bernstein evaluation of solved trajectories against the numpy restatement of
    bezier_base.h:77-115 (itself pinned by tests/test_oracle.py against the reference's evaluators); plus the properties the
    node relies on (segment ends = control points 0 / 5, C0-C2 joins)."""
    pb = make_batch(32, 40, "poly", first=77)
    _, g = solver.solve_two_stage(pb, want_stage0=False)
    S = 9
    pos, vel, acc = solver.sample(g.bez_coeff, g.poly_time, S)
    rp, rv, ra = oracle.bezier_sample(g.bez_coeff, g.poly_time, S)
    for a, b in ((pos, rp), (vel, rv), (acc, ra)):
        assert rel_err(a, b) < 1e-12
    cp = g.bez_coeff.reshape(32, 40, 3, 6) * g.poly_time[:, :, None, None]
    assert rel_err(pos[:, :, 0], cp[..., 0]) < 1e-12 and rel_err(pos[:, :, -1], cp[..., 5]) < 1e-12
    ok = g.rtn == 1
    for arr in (pos, vel, acc):                       # consecutive segments join in position, velocity, acceleration
        assert rel_err(arr[ok][:, :-1, -1], arr[ok][:, 1:, 0]) < 1e-6


def test_gpu_line_init_fixture(solver):
    """line_init_flag = true (ddp_optimizer.cpp:195-247 straight-line initialisation with the time-doubling loop,
    :255-269 feasibility check, :381-388 exit) on the B200 against the output of the reference's own translation unit."""
    pb, d = load_golden("box_n6_lineinit")
    r = solver.solve_batch(pb, infeas=1, zero_init=0, line_init=1, w_snap=1.0, w_terminal=100.0, w_time=50.0, iter_max=60)
    assert (r.rtn == d["s1_rtn"]).all() and (r.iters == d["s1_iters"]).all()
    assert (r.line_failed_out == d["s1_line_failed_out"]).all() and (r.infeas_out == d["s1_infeas_out"]).all()
    for f in OUT_FIELDS:
        assert rel_err(getattr(r, f), d["s1_" + f]) < TOL64, f


def test_gpu_time_allocation_device_matches_reference_rule(solver, oracle):
    """direct_ddp_time_allocation_device = initTimeAllocation (teach_repeat_planner.cpp:583-639, v0 = 0) against the C
    restatement of the oracle (segment by segment) and the numpy statement the problem generator uses; covers the
    short-segment branch (D < accd + dccd), the cruise branch and a different (max_vel, max_acc) pair."""
    import ctypes as C
    import torch
    from direct_b200.problems import time_allocation
    rng = np.random.default_rng(11)
    B, N = 37, 23
    step = rng.uniform(0.05, 6.0, size=(B, N + 1, 3)) * rng.choice([-1.0, 1.0], size=(B, N + 1, 3))
    pts = np.cumsum(step, axis=1)
    dev = torch.device("cuda", 0)
    for mv, ma in ((2.0, 2.0), (3.0, 2.5)):
        start, end = np.ascontiguousarray(pts[:, 0]), np.ascontiguousarray(pts[:, N])
        seeds = np.ascontiguousarray(pts[:, :N])
        t = [torch.from_numpy(a).to(dev) for a in (start, end, seeds)]
        out = torch.zeros(B, N, dtype=torch.float64, device=dev)
        solver.time_allocation_device(B, N, t[0].data_ptr(), t[1].data_ptr(), t[2].data_ptr(), mv, ma, out.data_ptr(),
                                      torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        got = out.cpu().numpy()
        want = np.zeros((B, N))
        dp = C.POINTER(C.c_double)
        for b in range(B):
            oracle.lib().ipddp_oracle_time_allocation(N, start[b].ctypes.data_as(dp), end[b].ctypes.data_as(dp),
                                                      seeds[b].ctypes.data_as(dp), mv, ma, want[b].ctypes.data_as(dp))
        assert rel_err(got, want) < 1e-14
        assert rel_err(got, time_allocation(pts, mv, ma)) < 1e-14
        d = np.linalg.norm(pts[:, 1:] - pts[:, :-1], axis=-1)
        assert ((d < mv * mv / ma).any() and (d > mv * mv / ma).any())   # both branches of the rule were exercised


def test_gpu_polytopes_with_more_than_32_planes_match_oracle(solver, oracle):
    """The reference bounds the planes of a polytope nowhere (m_c = 6 P + 55, ddp_optimizer.cpp:145, :1181-1187; cdd
    H-representations, poly_utils.cpp:127-166): P up to 40 against the oracle (fixture poly40_n10 holds the reference's
    own output for three of these)."""
    pb = make_batch(24, 30, "poly40", first=4100)
    assert pb.nplanes.max() > 32
    ok, (a0, a1) = conditioned_mask_two_stage(oracle, pb)
    assert ok.mean() >= SCREEN_FLOOR
    g0, g1 = solver.solve_two_stage(pb)
    for a, g in ((a0, g0), (a1, g1)):
        bad = results_differ(g, a, TOL64, OUT_FIELDS + ("jerk", "x_final")) & ok
        assert not bad.any(), np.nonzero(bad)[0]
        assert np.array_equal(a.stats[ok, :4], g.stats[ok, :4])
        assert_valid_result(pb, g)
    assert_screened_out_bounded(oracle, pb, ok, (a0, a1), (g0, g1))


def test_gpu_multi_device_handle_shards_host_batches(solver):
    """direct_ddp_opts.devices[]: a handle over several devices shards every host-buffer batch into contiguous
    ranges, one host thread + stream per device, results D2H straight into the caller's arrays -- same bits as one device.
    Runs with the devices that are there: two handles on device 0 when the box has a single GPU."""
    import torch
    from direct_b200.capi import Solver
    ndev = torch.cuda.device_count()
    devices = [0, 1] if ndev >= 2 else [0, 0]
    pb = make_batch(37, 20, "poly", first=900)          # 37 = 18 + 19: uneven shards
    a0, a1 = solver.solve_two_stage(pb)
    m = Solver(0, "fp64", devices=devices)
    assert m.lib.direct_ddp_device_count(m.h) == 2
    g0, g1 = m.solve_two_stage(pb)
    st = m.stats()
    for a, g in ((a0, g0), (a1, g1)):
        for f in ("rtn", "iters", "infeas_out", "cost", "poly_coeff", "bez_coeff", "poly_time", "jerk", "x_final"):
            assert np.array_equal(getattr(a, f), getattr(g, f)), f
        assert np.array_equal(a.stats[:, :4], g.stats[:, :4])
    assert st.bwd_knots == int(a0.stats[:, 1].sum() + a1.stats[:, 1].sum()) and st.kernel_launches == 2
    r = m.solve_batch(pb.slice(0, 1), infeas=1, zero_init=1, **STAGE0)      # B < ndevices: runs on devices[0]
    assert r.rtn[0] == a0.rtn[0] and np.array_equal(r.poly_coeff[0], a0.poly_coeff[0])
    m.close()
