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


GOLDEN_TWO_STAGE = ["box_n5", "poly_n12", "box_n50_single", "poly_n30_minvo", "box_n8_timepower1", "box_n100"]
OUT_FIELDS = ("cost", "poly_coeff", "bez_coeff", "poly_time")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle_py as O
    O.build(ref=False)
    return O


def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
