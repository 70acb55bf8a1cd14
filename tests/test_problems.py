import numpy as np

from direct_b200.problems import make_batch


def test_generator_is_deterministic_and_sliceable():
    a = make_batch(8, 10, "poly", first=3)
    b = make_batch(3, 10, "poly", first=5)
    assert np.array_equal(a.planes[2:5], b.planes) and np.array_equal(a.durations[2:5], b.durations)
    assert np.array_equal(a.x0[2:5], b.x0) and np.array_equal(a.nplanes[2:5], b.nplanes)


def test_corridor_geometry():
    for kind, pmax in (("box", 6), ("poly", 14)):
        pb = make_batch(16, 20, kind)
        assert pb.P_max == pmax and pb.nplanes.min() >= 6 and pb.nplanes.max() <= pmax
        nrm = np.linalg.norm(pb.planes[..., :3], axis=-1)
        active = np.arange(pmax)[None, None, :] < pb.nplanes[..., None]
        assert np.allclose(nrm[active], 1.0)
        assert (pb.planes[~active] == np.array([0, 0, 0, -1.0])).all()   # inactive padding
        # seed of cell i and of its neighbours lie strictly inside cell i (overlap)
        def inside(pts, cells):
            v = np.einsum("bnpa,bna->bnp", cells[..., :3], pts) + cells[..., 3]
            return (v < 0).all()
        assert inside(pb.seeds, pb.planes)
        assert inside(pb.seeds[:, 1:], pb.planes[:, :-1]) and inside(pb.seeds[:, :-1], pb.planes[:, 1:])
        assert inside(pb.xd[:, None, :3], pb.planes[:, -1:]) and inside(pb.x0[:, None, :3], pb.planes[:, :1])
        assert (pb.durations > 0).all() and np.isfinite(pb.durations).all()
