"""The CUDA solver source (direct_b200/csrc/ipddp_solver.h) compiled lane-by-lane for the CPU
(tools/emulate.cpp) against the oracle and the reference fixtures.  This exercises the device LOGIC in the
GPU-less container; the shipped path is tested by test_gpu_parity.py through the C-ABI on a B200."""
import numpy as np
import pytest

from conftest import (GOLDEN_TWO_STAGE, OUT_FIELDS, SCREEN_FLOOR, assert_screened_out_bounded, assert_valid_result,
                      conditioned_mask_two_stage, load_golden, rel_err, results_differ)
from direct_b200.problems import STAGE0, STAGE1, make_batch

TOL64 = 1e-5   # north_star: <= 1e-5 relative in fp64
TOL32 = 2e-2   # fp32 emulation is only smoke-checked here; the GPU test states the real fp32 bound


@pytest.fixture(scope="module")
def emu():
    from tools import emu_py
    emu_py.lib()
    return emu_py


@pytest.mark.parametrize("name", GOLDEN_TWO_STAGE)
def test_emulated_kernel_matches_reference_fixtures(emu, name):
    pb, d = load_golden(name)
    ov = dict(minvo=int(d["minvo"]), time_power=int(d["time_power"]))
    r0 = emu.solve_batch(pb, infeas=1, zero_init=1, **dict(STAGE0, **ov))
    assert (r0.rtn == d["s0_rtn"]).all() and (r0.iters == d["s0_iters"]).all()
    for f in OUT_FIELDS:
        assert rel_err(getattr(r0, f), d["s0_" + f]) < TOL64, f
    r1 = emu.solve_batch(pb, infeas=d["s0_infeas_out"], zero_init=0, init_bez=d["s0_bez_coeff"], durations=d["dur1"],
                         **dict(STAGE1, **ov))
    assert (r1.rtn == d["s1_rtn"]).all() and (r1.iters == d["s1_iters"]).all()
    for f in OUT_FIELDS:
        assert rel_err(getattr(r1, f), d["s1_" + f]) < TOL64, f
    assert rel_err(r1.jerk.sum(1), d["s1_jerk_sum"]) < TOL64


def test_emulated_kernel_line_init_fixture(emu):
    """line_init_flag = true (ddp_optimizer.cpp:195-247, :255-269, :381-388) against the reference's own output."""
    pb, d = load_golden("box_n6_lineinit")
    r = emu.solve_batch(pb, infeas=1, zero_init=0, line_init=1, w_snap=1.0, w_terminal=100.0, w_time=50.0, iter_max=60)
    assert (r.rtn == d["s1_rtn"]).all() and (r.iters == d["s1_iters"]).all()
    assert (r.line_failed_out == d["s1_line_failed_out"]).all()
    for f in OUT_FIELDS:
        assert rel_err(getattr(r, f), d["s1_" + f]) < TOL64, f


def test_emulated_two_stage_matches_oracle_ragged_planes(emu, oracle):
    pb = make_batch(6, 17, "poly", first=321)
    a0, a1 = oracle.two_stage_batch(pb, nthreads=2)
    e0, e1 = emu.two_stage_batch(pb)
    assert (a0.rtn == e0.rtn).all() and (a1.rtn == e1.rtn).all()
    assert (a0.iters == e0.iters).all() and (a1.iters == e1.iters).all()
    assert np.array_equal(a1.stats[:, :4], e1.stats[:, :4])  # same sweeps / rollouts, knot for knot
    for f in OUT_FIELDS + ("jerk", "x_final"):
        assert rel_err(getattr(e1, f), getattr(a1, f)) < TOL64, f


def test_emulated_kernel_parity_on_conditioned_trajectories(emu, oracle):
    """The batch of the GPU parity test that contains a chaotic trajectory (index 44 runs 72 iterations in the
    oracle; its iterates separate from ANY differently-rounded evaluation around iteration 33): every
    trajectory that passes the conditioning screen matches to 1e-5, the rest are still valid results."""
    pb = make_batch(64, 33, "box", first=2033)
    pert = []
    ok, (a0, a1) = conditioned_mask_two_stage(oracle, pb, nthreads=4, keep=pert)
    assert SCREEN_FLOOR <= ok.mean() < 1.0 and not ok[44]
    e0, e1 = emu.two_stage_batch(pb)
    for a, e in ((a0, e0), (a1, e1)):
        bad = results_differ(e, a, TOL64, OUT_FIELDS + ("jerk", "x_final")) & ok
        assert not bad.any(), np.nonzero(bad)[0]
        assert_valid_result(pb, e)
    assert_screened_out_bounded(oracle, pb, ok, (a0, a1), (e0, e1), perturbed=pert)


def test_emulated_fp32_smoke(emu, oracle):
    pb = make_batch(4, 10, "box", first=50)
    a0, a1 = oracle.two_stage_batch(pb)
    f0, f1 = emu.two_stage_batch(pb, fp32=True)
    assert (f0.rtn == a0.rtn).all()
    assert rel_err(f0.cost, a0.cost) < TOL32 and rel_err(f1.cost, a1.cost) < TOL32
