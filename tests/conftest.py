import os
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(1e-300, np.max(np.abs(b))))


def load_golden(name):
    """Fixture written by tests/golden/make_golden.py from the reference's own translation unit."""
    from direct_b200.problems import ProblemBatch
    d = np.load(os.path.join(GOLDEN, name + ".npz"))
    pb = ProblemBatch(int(d["B"]), int(d["N"]), int(d["P_max"]), np.ascontiguousarray(d["planes"]),
                      np.ascontiguousarray(d["nplanes"]), np.ascontiguousarray(d["durations"]),
                      np.ascontiguousarray(d["seeds"]), np.ascontiguousarray(d["x0"]), np.ascontiguousarray(d["xd"]),
                      float(d["max_vel"]), float(d["max_acc"]))
    return pb, d


GOLDEN_TWO_STAGE = ["box_n5", "poly_n12", "box_n50_single", "poly_n30_minvo", "box_n8_timepower1", "box_n100", "poly40_n10", "poly_n200"]
OUT_FIELDS = ("cost", "poly_coeff", "bez_coeff", "poly_time")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle_py as O
    O.build(ref=False)
    return O


# ---------------------------------------------------------------------------------------------------------
# Conditioning screen.  The reference algorithm takes discrete decisions (line-search step index, filter
# acceptance, Cholesky failure -> regularisation, early exits) inside a nonlinear iteration that amplifies
# perturbations by ~10x per iteration while it struggles at high regularisation (DESIGN.md "Parity").  On a
# few per cent of the synthetic corridors a 1-ulp change of the INPUTS therefore moves the reference's own
# outputs by far more than 1e-5 -- the oracle built with -ffp-contract=fast disagrees with itself on those,
# and so would the reference rebuilt with other compiler flags.  No implementation with a different rounding
# sequence can promise 1e-5 there.  The screen below finds them with the oracle alone: a trajectory is
# "conditioned" when four 2^-48-relative input perturbations leave every oracle output within 1e-6 relative
# (a tenth of the parity tolerance) and every discrete outcome unchanged.  Parity (<= 1e-5, identical rtn /
# iteration counts) is asserted on ALL conditioned trajectories; the others are bounded by
# assert_screened_out_bounded (same return codes as the oracle's own runs, valid, cost inside the oracle's own
# spread).  Measured pass fractions of the test batches (oracle alone, this container): 64 x 33 box 0.969,
# 48 x 50 poly 0.917 (4 trajectories on which the oracle disagrees with itself at 1e-5), 32 x 100 poly 0.969,
# 8 x 200 box 1.0, 24 x 30 poly40 0.917, the 128-trajectory sample of the 4096 x 100 bench batch 0.977.
# ---------------------------------------------------------------------------------------------------------
PERTURB_EPS = 2.0 ** -48
SCREEN_TOL = 1e-6


def _row_rel(a, b):
    a = np.asarray(a, dtype=np.float64).reshape(len(a), -1)
    b = np.asarray(b, dtype=np.float64).reshape(len(b), -1)
    return np.max(np.abs(a - b), axis=1) / np.maximum(1e-300, np.max(np.abs(b), axis=1))


def results_differ(a, b, tol, fields=OUT_FIELDS):
    """Per-trajectory bool: discrete outcome or any output field of Result `a` differs from `b` by > tol."""
    d = (a.rtn != b.rtn) | (a.iters != b.iters) | (a.infeas_out != b.infeas_out)
    for f in fields:
        d |= _row_rel(getattr(a, f), getattr(b, f)) > tol
    return d


def perturbed_batches(pb):
    from direct_b200.problems import ProblemBatch
    e = PERTURB_EPS
    mk = lambda **kw: ProblemBatch(**{**{f: getattr(pb, f) for f in ("B", "N", "P_max", "planes", "nplanes", "durations",
                                                                       "seeds", "x0", "xd", "max_vel", "max_acc")}, **kw})
    return [mk(durations=pb.durations * (1 + e)), mk(durations=pb.durations * (1 - e)),
            mk(x0=pb.x0 * (1 + e), xd=pb.xd * (1 - e)), mk(planes=np.ascontiguousarray(pb.planes * (1 + e)))]


SCREEN_FLOOR = 0.90        # at least this fraction of every small test batch must pass the screen (measured 0.917 - 1.0)
SCREEN_FLOOR_FULL = 0.95   # ... and of the sample of the full-size bench batch (measured 0.977)


def conditioned_mask_two_stage(oracle, pb, base=None, nthreads=None, keep=None):
    """(mask, (a0, a1)): trajectories on which the oracle's two-stage result is insensitive to 1-ulp inputs.
    `keep` (a list) receives the oracle's perturbed runs [(p0, p1), ...] for assert_screened_out_bounded."""
    nt = nthreads or oracle.max_threads()
    a0, a1 = base if base is not None else oracle.two_stage_batch(pb, nthreads=nt)
    ok = np.ones(pb.B, dtype=bool)
    for q in perturbed_batches(pb):
        p0, p1 = oracle.two_stage_batch(q, nthreads=nt)
        ok &= ~results_differ(p0, a0, SCREEN_TOL) & ~results_differ(p1, a1, SCREEN_TOL)
        if keep is not None:
            keep.append((p0, p1))
    return ok, (a0, a1)


def assert_screened_out_bounded(oracle, pb, ok, base, got, perturbed=None, slack=0.10):
    """The trajectories the screen sets aside are not exempt: on each of them the device result must (i) end with a
    return code that the oracle itself produces under its four 2^-48 input perturbations (or the unperturbed run),
    (ii) be a valid result (assert_valid_result: finite, inside the corridor when converged), and (iii) have a final
    cost inside the oracle's own min..max over those five runs widened by `slack` (10 %) -- i.e. the device lands no
    further from the oracle than the oracle lands from itself."""
    if ok.all():
        return
    if perturbed is None:
        perturbed = []
        conditioned_mask_two_stage(oracle, pb, base=base, keep=perturbed)
    idx = np.nonzero(~ok)[0]
    for st in (0, 1):
        runs = [base[st]] + [p[st] for p in perturbed]
        rt = np.stack([r.rtn[idx] for r in runs])            # (5, k)
        cs = np.stack([r.cost[idx] for r in runs])
        g = got[st]
        assert np.isfinite(g.cost[idx]).all()
        for j, i in enumerate(idx):
            assert g.rtn[i] in set(rt[:, j].tolist()), (st, int(i), int(g.rtn[i]), rt[:, j])
            lo, hi = cs[:, j].min(), cs[:, j].max()
            assert lo - slack * abs(lo) <= g.cost[i] <= hi + slack * abs(hi), (st, int(i), g.cost[i], lo, hi)


def assert_valid_result(pb, r):
    """What must hold for every returned trajectory, conditioned or not."""
    assert set(np.unique(r.rtn)).issubset({0, 1, 2, -3, -4})
    assert np.isfinite(r.cost).all() and np.isfinite(r.poly_coeff).all() and np.isfinite(r.bez_coeff).all()
    done = np.isin(r.rtn, (1, 2))
    if done.any():   # exits 1 and 2 need every c < 2e-4 where c already carries the -2e-4 margin (ddp_optimizer.cpp:346-378,
        # :1281-1283; hazard H8): control points are at most 4e-4 outside their polytope
        cp = r.bez_coeff.reshape(pb.B, pb.N, 3, 6) * r.poly_time[:, :, None, None]
        val = np.einsum("bnpa,bnaj->bnpj", pb.planes[..., :3], cp) + pb.planes[..., 3:4]
        assert (val[done] < 4e-4 + 1e-9).all()
        assert (r.poly_time[done] > 0.3).all()


def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
