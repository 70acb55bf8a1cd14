"""Voxel-map kernels: paraConvexTest + paraResultCheck, paraCubeInflation and the two host loops
around them (polyhedron_generator/src/cluster_engine.cu, cluster_server.cu).

CPU suite: oracle/voxel_oracle.c against fixtures produced by the reference's own kernels on a B200
(tests/golden/voxel_ref_*.npz, tests/golden/make_voxel_golden.py), against a separately written pure-Python ray walk, and the
loop restatements against their step-by-step composition and their invariants.
GPU suite (-m gpu): direct_b200's kernels through the C-ABI == oracle == the reference's kernels run live
(oracle/_ref/libvoxel_ref.so), byte for byte."""
import glob
import os
import re
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT

sys.path.insert(0, GOLDEN)

from direct_b200 import voxel as X      # noqa: E402
from oracle import voxel_py as V        # noqa: E402  (checker)
import make_voxel_golden as MG          # noqa: E402


# ---- a separately written ray walk (Python floats are IEEE doubles: same operations, same comparisons) ------------------------
def py_ray(occ, inside, a, b):
    x, y, z = (int(t) for t in a)
    ex, ey, ez = (int(t) for t in b)
    d = [ex - x, ey - y, ez - z]
    step = [(t > 0) - (t < 0) for t in d]
    tmax = [99999.0 if t == 0 else 0.5 / abs(t) for t in d]
    with np.errstate(all="ignore"):
        delta = [float(np.float64(s) / np.float64(t)) for s, t in zip(step, d)]
    p = [x, y, z]
    ok = True
    while p != [ex, ey, ez]:
        if tmax[0] < tmax[1]:
            k = 0 if tmax[0] < tmax[2] else 2
        else:
            k = 1 if tmax[1] < tmax[2] else 2
        p[k] += step[k]; tmax[k] += delta[k]
        if inside[p[0], p[1], p[2]] > 0:
            return ok
        if p == [ex, ey, ez]:
            break
        if occ[p[0], p[1], p[2]] > 0:
            ok = False
    return ok


def py_convex_test(occ, inside, cand, clu, fill=2):
    Cn = len(cand)
    cc = np.full(Cn * (Cn + 1) // 2, fill, np.uint8); cl = np.zeros(Cn, np.uint8)
    for t in range(Cn):
        for i in range(t):
            cc[t * (t + 1) // 2 + i] = py_ray(occ, inside, cand[t], cand[i])
        cl[t] = all(py_ray(occ, inside, cand[t], c) for c in clu)
    return cc, cl


def small_case(shape=(24, 20, 10), pillars=8, seed=2, cell=(12, 10, 4), inflate=3):
    occ = X.make_map(shape, pillars, seed, clear=(*cell, 2))
    v, _ = V.inflate_box(occ, X.box_vertices(*cell, *cell), inflate)
    inside, use, shell = X.cube_shell(shape, v)
    return occ, v, inside, use, shell, MG.first_candidates(occ, inside, use, shell)


GOLDEN_VOXEL = sorted(glob.glob(os.path.join(GOLDEN, "voxel_ref_*.npz")))
# The reference's own kernels (built where /root/reference is mounted, shipped to the GPU box as a git-ignored .so).  When the file did not
# travel, the live comparison is skipped; the committed fixtures above, which are outputs of the same kernels, still pin both sides.
HAVE_REF = os.path.exists(V.REF_SO)


# ---- CPU suite --------------------------------------------------------------------------------------------------------------------
def test_library_exports_every_voxel_symbol():
    from direct_b200 import capi
    lib = capi.load_library()
    hdr = open(os.path.join(ROOT, "include", "direct_voxel.h")).read()
    declared = sorted(set(re.findall(r"\b(direct_voxel_[a-z_0-9]+)\s*\(", hdr)))
    assert set(declared) == set(X.VOXEL_EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name


def test_fixtures_of_the_reference_kernels_are_committed():
    assert len(GOLDEN_VOXEL) >= 3, "tests/golden/voxel_ref_*.npz (outputs of the reference's own kernels) missing"


@pytest.mark.parametrize("path", GOLDEN_VOXEL, ids=[os.path.basename(p) for p in GOLDEN_VOXEL])
def test_oracle_matches_reference_kernel_fixtures(path):
    d = np.load(path)
    cc, cl = V.convex_test(d["occ"], d["inside"], d["cand"], d["cluster"])
    assert np.array_equal(cc, d["can_can"]) and np.array_equal(cl, d["can_clu"])
    assert (cc == 0).any() and (cc == 1).any() and (cl == 0).any() and (cl == 1).any()   # the fixture exercises both answers
    mine = [V.cube_inflation(d["occ"], d["vertex_idx"], k, 1, 128 * 128) for k in range(6)]
    assert mine == d["inflation"][0].tolist()


def test_oracle_matches_python_ray_walk():
    occ, v, inside, use, shell, cand = small_case()
    assert 20 < len(cand) < 400
    cc, cl = V.convex_test(occ, inside, cand, shell)
    pc, pl = py_convex_test(occ, inside, cand, shell)
    assert np.array_equal(cc, pc) and np.array_equal(cl, pl)
    assert (cc == 0).any() and (cl == 0).any()


def test_oracle_empty_map_every_ray_is_free_and_edge_counts():
    occ = np.zeros((10, 9, 8), np.uint8); inside = np.zeros_like(occ)
    cand = np.array([[1, 1, 1], [8, 7, 6], [1, 7, 1], [5, 5, 5]], np.int32)
    cc, cl = V.convex_test(occ, inside, cand, np.array([[0, 0, 0], [9, 8, 7]], np.int32))
    tri = [t * (t + 1) // 2 + i for t in range(4) for i in range(t)]
    assert (cc[tri] == 1).all() and (np.delete(cc, tri) == 2).all() and (cl == 1).all()
    cc, cl = V.convex_test(occ, inside, np.zeros((0, 3), np.int32), cand)            # no candidate
    assert cc.size == 0 and cl.size == 0
    cc, cl = V.convex_test(occ, inside, cand[:1], np.zeros((0, 3), np.int32))         # one candidate, empty cluster
    assert cc.tolist() == [2] and cl.tolist() == [1]
    occ[4, :, :] = 1                                                                  # a wall between x < 4 and x > 4
    cc, cl = V.convex_test(occ, inside, cand, np.array([[0, 0, 0]], np.int32))
    assert cl.tolist() == [1, 0, 1, 0] and cc[1 * 2 // 2 + 0] == 0 and cc[2 * 3 // 2 + 0] == 1
    inside[5, :, :] = 1                                                               # rays stop at an inside voxel before the wall
    cc, cl = V.convex_test(occ, inside, cand, np.array([[0, 0, 0]], np.int32))
    assert cl.tolist() == [1, 1, 1, 1]


def test_oracle_cube_inflation_each_direction():
    occ = np.zeros((12, 12, 12), np.uint8)
    v = X.box_vertices(4, 4, 4, 7, 7, 7)
    assert [V.cube_inflation(occ, v, k) for k in range(6)] == [1] * 6
    for k, cell in enumerate([(5, 3, 5), (5, 8, 5), (3, 5, 5), (8, 5, 5), (5, 5, 3), (5, 5, 8)]):   # Y-, Y+, X-, X+, Z-, Z+
        o = occ.copy(); o[cell] = 1
        assert [V.cube_inflation(o, v, j) for j in range(6)] == [0 if j == k else 1 for j in range(6)]
    o = occ.copy(); o[3, 3, 5] = 1   # diagonal to the box: no face sees it
    assert [V.cube_inflation(o, v, j) for j in range(6)] == [1] * 6


def test_oracle_inflate_box_is_free_and_maximal():
    occ = X.make_map((40, 36, 14), 16, 9, clear=(20, 18, 6, 2))
    v, it = V.inflate_box(occ, X.box_vertices(20, 18, 6, 20, 18, 6), 1000)
    x0, y0, z0, x1, y1, z1 = X.box_bounds(v)
    assert it < 1000 and occ[x0:x1 + 1, y0:y1 + 1, z0:z1 + 1].sum() == 0
    at_wall = [y0 == 0, y1 == 35, x0 == 0, x1 == 39, z0 == 0, z1 == 13]
    assert all(w or V.cube_inflation(occ, v, k) == 0 for k, w in enumerate(at_wall))
    v2, it2 = V.inflate_box(occ, X.box_vertices(20, 18, 6, 20, 18, 6), 2)   # the iteration limit is honoured
    assert it2 == 2 and (X.box_bounds(v2)[3] - X.box_bounds(v2)[0]) <= 4


def py_cluster(occ, inside, use, invalid, shell, itr_max):
    """polytopeCluster_gpu composed step by step from the kernel restatement (cluster_server.cu:556-767)."""
    use = use.copy(); invalid = invalid.copy()
    cluster = [tuple(c) for c in shell]; active = list(cluster); itr = 0
    while itr < itr_max:
        cand = MG.first_candidates(np.where((occ == 1) | (invalid == 1), 1, 0).astype(np.uint8), inside, use, np.array(active, np.int32))
        for a in active:
            use[a] = 1
        for c in cand:
            use[tuple(c)] = 1
        if len(cand) == 0:
            break
        cc, cl = V.convex_test(occ, inside, cand, np.array(cluster, np.int32))
        acc = np.zeros(len(cand), bool); active = []
        for i in range(len(cand)):
            ok = bool(cl[i]) and not any(cc[i * (i + 1) // 2 + j] == 0 and acc[j] for j in range(i))
            if ok:
                acc[i] = True; cluster.append(tuple(cand[i])); active.append(tuple(cand[i]))
            else:
                invalid[tuple(cand[i])] = 1
        if not active:
            break
        itr += 1
    return np.array(cluster, np.int32), use, invalid, itr


def test_oracle_cluster_equals_stepwise_composition():
    occ, v, inside, use, shell, _ = small_case((28, 24, 12), 10, 4, (14, 12, 5), 3)
    inv0 = np.zeros_like(occ)
    for itr_max in (1, 3, 20):
        a = V.cluster(occ, inside, use, inv0, shell, 20000, 4000, itr_max)
        b = py_cluster(occ, inside, use, inv0, shell, itr_max)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2]) and a[3] == b[3]
    assert len(a[0]) > len(shell) and a[2].sum() > 0
    assert occ[a[0][:, 0], a[0][:, 1], a[0][:, 2]].sum() == 0            # the cluster never enters an obstacle
    assert len(np.unique(a[0], axis=0)) == len(a[0])                       # no voxel twice
    with pytest.raises(RuntimeError):
        V.cluster(occ, inside, use, inv0, shell, 20000, 8, 20)             # candidate capacity


# ---- GPU suite ----------------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def solver():
    from direct_b200 import capi
    s = capi.Solver(0, "fp64")   # raises without the CUDA library or a GPU: no CPU fallback
    yield s
    s.close()


def _cases_gpu():
    yield "small", small_case()
    yield "poly", small_case((28, 24, 12), 10, 4, (14, 12, 5), 3)
    for name, occ, v, inside, shell, cand in MG.cases():
        yield name, (occ, v, inside, None, shell, cand)


@pytest.mark.gpu
def test_gpu_convex_test_matches_oracle_and_reference_kernels(solver):
    n = 0
    for name, (occ, v, inside, _, shell, cand) in _cases_gpu():
        gc, gl = X.convex_test(solver, occ, inside, cand, shell)
        oc, ol = V.convex_test(occ, inside, cand, shell)
        assert np.array_equal(gc, oc) and np.array_equal(gl, ol), name
        if HAVE_REF:
            rc, rl, _ = V.ref_convex_test(occ, inside, cand, shell)
            assert np.array_equal(gc, rc) and np.array_equal(gl, rl), name
        n += len(cand)
    assert n > 2000
    # ragged / empty edge cases
    occ, v, inside, _, shell, cand = small_case()
    for c, k in ((cand[:0], shell), (cand[:1], shell[:0]), (cand[:33], shell[:1]), (cand[:64], shell[:0])):
        gc, gl = X.convex_test(solver, occ, inside, c, k)
        oc, ol = V.convex_test(occ, inside, c, k)
        assert np.array_equal(gc, oc) and np.array_equal(gl, ol)


@pytest.mark.gpu
def test_gpu_fixtures_reproduced_through_the_c_abi(solver):
    assert GOLDEN_VOXEL
    for path in GOLDEN_VOXEL:
        d = np.load(path)
        gc, gl = X.convex_test(solver, d["occ"], d["inside"], d["cand"], d["cluster"])
        assert np.array_equal(gc, d["can_can"]) and np.array_equal(gl, d["can_clu"])
        assert [X.cube_inflation(solver, d["occ"], d["vertex_idx"], k) for k in range(6)] == d["inflation"][0].tolist()


@pytest.mark.gpu
def test_gpu_cube_inflation_matches_oracle_and_reference_kernel(solver):
    occ = np.zeros((12, 12, 12), np.uint8)
    v = X.box_vertices(4, 4, 4, 7, 7, 7)
    maps = [occ]
    for cell in [(5, 3, 5), (5, 8, 5), (3, 5, 5), (8, 5, 5), (5, 5, 3), (5, 5, 8), (3, 3, 5)]:
        o = occ.copy(); o[cell] = 1; maps.append(o)
    for o in maps:
        for k in range(6):
            want = V.cube_inflation(o, v, k)
            assert X.cube_inflation(solver, o, v, k) == want
            assert not HAVE_REF or V.ref_cube_inflation(o, v, k) == want
    big = X.make_map((200, 200, 40), 150, 6, clear=(100, 100, 15, 5))   # a face of 8000 cells: more than one CTA's worth
    vb, _ = V.inflate_box(big, X.box_vertices(100, 100, 15, 100, 100, 15), 1000)
    for k in range(6):
        if [vb[8] == 0, vb[9] == 199, vb[3] == 0, vb[0] == 199, vb[20] == 0, vb[16] == 39][k]:
            continue
        assert X.cube_inflation(solver, big, vb, k) == V.cube_inflation(big, vb, k) == 0


@pytest.mark.gpu
def test_gpu_inflate_box_matches_oracle_and_reference_loop(solver):
    for shape, pillars, seed, cell in (((40, 36, 14), 16, 9, (20, 18, 6)), ((120, 120, 30), 60, 6, (60, 60, 12)), ((64, 64, 64), 0, 1, (5, 60, 30))):
        occ = X.make_map(shape, pillars, seed, clear=(*cell, 2))
        v0 = X.box_vertices(*cell, *cell)
        for itr_max in (0, 1, 3, 20, 1000):
            gv, gi = X.inflate_box(solver, occ, v0, itr_max)
            ov, oi = V.inflate_box(occ, v0, itr_max)
            assert np.array_equal(gv, ov) and gi == oi, (shape, itr_max)
        if HAVE_REF:
            rv, ri, _ = V.ref_inflate_box(occ, v0, 1000)
            assert np.array_equal(gv, rv) and gi == ri


@pytest.mark.gpu
def test_gpu_cluster_loop_matches_oracle(solver):
    for (shape, pillars, seed, cell, inflate) in (((24, 20, 10), 8, 2, (12, 10, 4), 3), ((28, 24, 12), 10, 4, (14, 12, 5), 3),
                                                  ((60, 60, 20), 30, 6, (30, 30, 8), 20)):
        occ, v, inside, use, shell, _ = small_case(shape, pillars, seed, cell, inflate)
        inv0 = np.zeros_like(occ)
        for itr_max in (0, 1, 2, 20):
            g = X.cluster(solver, occ, inside, use, inv0, shell, 40000, 8000, itr_max)
            o = V.cluster(occ, inside, use, inv0, shell, 40000, 8000, itr_max)
            assert np.array_equal(g[0], o[0]), (shape, itr_max, len(g[0]), len(o[0]))
            assert np.array_equal(g[1], o[1]) and np.array_equal(g[2], o[2]) and g[3] == o[3]
        assert len(g[0]) > len(shell)
    from direct_b200 import capi
    with pytest.raises(capi.DirectDdpError):
        X.cluster(solver, occ, inside, use, inv0, shell, 40000, 8, 20)       # candidate capacity exceeded: a loud error
    g2 = X.cluster(solver, occ, inside, use, inv0, shell, 40000, 8000, 20)   # and the handle is still good afterwards
    assert np.array_equal(g2[0], g[0])


@pytest.mark.gpu
def test_gpu_cluster_full_size_properties(solver):
    """The node's map size (global_planner.launch: 0.15 m voxels) is beyond what the sequential oracle finishes in seconds:
    size-independent properties instead."""
    occ = X.make_map((200, 200, 40), 160, 6, clear=(100, 100, 15, 6))
    v, _ = X.inflate_box(solver, occ, X.box_vertices(100, 100, 15, 100, 100, 15), 20)
    inside, use, shell = X.cube_shell(occ.shape, v)
    inv0 = np.zeros_like(occ)
    clu, use1, inv1, it = X.cluster(solver, occ, inside, use, inv0, shell, 50000, 10000, 8)
    assert np.array_equal(clu[:len(shell)], shell) and len(clu) > len(shell) and 1 <= it <= 8
    assert occ[clu[:, 0], clu[:, 1], clu[:, 2]].sum() == 0 and len(np.unique(clu, axis=0)) == len(clu)
    new = clu[len(shell):]
    assert inside[new[:, 0], new[:, 1], new[:, 2]].sum() == 0 and inv1[new[:, 0], new[:, 1], new[:, 2]].sum() == 0
    assert (use1[clu[:, 0], clu[:, 1], clu[:, 2]] == 1).all() and (use1[inv1 == 1] == 1).all()
    # every accepted voxel sees every voxel of the initial cluster (the defining property of the convex test)
    cc, cl = X.convex_test(solver, occ, inside, new[:512], shell)
    assert (cl == 1).all()
    # deterministic: a second run gives the same cluster in the same order
    clu2, _, _, it2 = X.cluster(solver, occ, inside, use, inv0, shell, 50000, 10000, 8)
    assert np.array_equal(clu, clu2) and it == it2


# ---- polygonGeneration as a whole (cluster_server.cu:769-966) ---------------------------------------------------------------------------
POLY_CASES = [  # shape, pillars, map seed, seed voxel, itr_inflate_max, itr_cluster_max
    ((40, 36, 14), 16, 9, (20, 18, 6), 20, 3),
    ((24, 20, 10), 8, 2, (12, 10, 4), 3, 20),
    ((30, 30, 30), 10, 5, (0, 15, 29), 20, 4),      # seed in a corner region: the box runs into the map boundary
    ((60, 60, 20), 30, 6, (30, 30, 8), 20, 6),
    ((24, 20, 10), 8, 2, (12, 10, 4), 0, 5),        # no inflation: a one-voxel box is degenerate, the seed is the polytope
]


def _slab_map():
    occ = np.ones((20, 20, 8), np.uint8)
    occ[2:18, 2:18, 3] = 0                           # a corridor one voxel high: the box is one voxel thick (degenerate, :911-920)
    return occ


def test_oracle_polytope_equals_its_parts():
    for shape, pillars, mseed, cell, inf, clu in POLY_CASES[:3]:
        occ = X.make_map(shape, pillars, mseed, clear=(*cell, 2))
        r = V.polytope(occ, cell, inf, clu, 40000, 8000)
        v, it = V.inflate_box(occ, X.box_vertices(*cell, *cell), inf)
        inside, use, shell = X.cube_shell(shape, v)               # numpy statement of :834-895
        c = V.cluster(occ, inside, use, np.zeros_like(occ), shell, 40000, 8000, clu)
        assert np.array_equal(r["vertex_idx"], v) and r["iters"] == [it, c[3]]
        assert np.array_equal(r["cluster"], c[0]) and np.array_equal(r["use"], c[1]) and np.array_equal(r["invalid"], c[2])
        assert np.array_equal(r["inside"], inside)
    r = V.polytope(_slab_map(), (10, 10, 3), 20, 5, 4000, 800)
    assert len(r["cluster"]) == 16 * 16 and r["iters"][1] == 0 and r["invalid"].sum() == 0


@pytest.mark.gpu
def test_gpu_polytope_matches_oracle(solver):
    cases = [(X.make_map(shape, pillars, mseed, clear=(*cell, 2)), cell, inf, clu) for shape, pillars, mseed, cell, inf, clu in POLY_CASES]
    cases.append((_slab_map(), (10, 10, 3), 20, 5))
    cases.append((np.zeros((12, 9, 7), np.uint8), (5, 4, 3), 50, 5))          # empty map: the box is the whole map, nothing to cluster
    for occ, cell, inf, clu in cases:
        g = X.polytope(solver, occ, cell, inf, clu, 40000, 8000)
        o = V.polytope(occ, cell, inf, clu, 40000, 8000)
        assert np.array_equal(g["vertex_idx"], o["vertex_idx"]) and g["iters"] == o["iters"], (occ.shape, cell)
        assert np.array_equal(g["cluster"], o["cluster"]), (occ.shape, cell, len(g["cluster"]), len(o["cluster"]))
        for f in ("inside", "use", "invalid"):
            assert np.array_equal(g[f], o[f]), (occ.shape, cell, f)
    from direct_b200 import capi
    with pytest.raises(capi.DirectDdpError):
        X.polytope(solver, cases[0][0], (99, 0, 0), 5, 5, 1000, 1000)          # seed outside the map
    with pytest.raises(capi.DirectDdpError):
        X.polytope(solver, cases[0][0], cases[0][1], 20, 3, 100, 8000)         # cluster capacity below the box's boundary


@pytest.mark.gpu
@pytest.mark.skipif(not V.ref_server_available(), reason="oracle/_ref/libvoxel_server_ref.so (the reference's host loop) did not travel")
def test_gpu_polytope_against_the_references_own_host_loop(solver):
    """cudaPolytopeGeneration::polygonGeneration itself (cluster_server.cu:769-966, compiled unmodified with its kernels for sm_100a, ROS
    reduced to a stopwatch) against the oracle and direct_voxel_polytope.  The reference reads the can_can row of the LAST candidate of
    every iteration from host bytes its download never wrote; with that stale read reproduced (oracle, host_can_can) the oracle returns
    the reference's cluster voxel for voxel, call after call on one generator object; without it the oracle is direct_voxel_polytope
    (test_gpu_polytope_matches_oracle).  So the product and the reference differ by that read and nothing else."""
    grew = differ = 0
    cases = [(X.make_map(shape, pillars, mseed, clear=(*cell, 2)), cell, inf, clu) for shape, pillars, mseed, cell, inf, clu in POLY_CASES]
    cases.append((_slab_map(), (10, 10, 3), 20, 5))
    for occ, cell, inf, clu in cases:
        refs, _ = V.ref_server_polytope(occ, cell, inf, clu, reps=3)
        host = V.reference_host_buffer(10000)
        for k, ref in enumerate(refs):   # the buffer lives as long as the generator: later calls see what earlier ones left
            o = V.polytope(occ, cell, inf, clu, 50000, 10000, host_can_can=host)
            assert np.array_equal(o["cluster"], ref), (occ.shape, cell, k, len(o["cluster"]), len(ref))
        g = X.polytope(solver, occ, cell, inf, clu, 50000, 10000)
        assert np.array_equal(g["cluster"], V.polytope(occ, cell, inf, clu, 50000, 10000)["cluster"])
        grew += g["iters"][1] > 0
        differ += not np.array_equal(g["cluster"], refs[0])
        if np.array_equal(g["vertex_idx"], o["vertex_idx"]):   # same box, same initial cluster
            n0 = len(X.cube_shell(occ.shape, g["vertex_idx"])[2])
            assert np.array_equal(g["cluster"][:n0], refs[0][:n0])
    assert grew >= 3       # the cases do run clustering iterations, not only the box
    assert differ >= 1     # and the stale read is exercised (the corner case loses its last candidate in the reference)


def test_closed_form_boundary_rank_of_cube_shell_kernel():
    """cube_shell_kernel (direct_b200/csrc/voxel.cuh) places a boundary voxel of the inflated box at its position in the reference's
    x, y, z scan (cluster_server.cu:848-886) by a closed form instead of a compaction pass.  The same formula, stated here in Python,
    against the scan itself for every box shape up to 6 x 6 x 6 (thin and one-voxel boxes included); the GPU suite checks the kernel."""
    import itertools

    def before_slab(a, ny, nz, nxb):
        full = ny * nz
        ring = full - (ny - 2 if ny > 2 else 0) * (nz - 2 if nz > 2 else 0)
        n = 0
        if a > 0:
            n += full
        if a > 1:
            n += (a - 1 if a - 1 < nxb - 1 else nxb - 2) * ring
        if a > nxb - 1 and nxb > 1:
            n += full
        return n

    def rank(a, b, k, bx, by, bz):
        ex, ey, ez = a in (0, bx - 1), b in (0, by - 1), k in (0, bz - 1)
        if not (ex or ey or ez):
            return None
        r = before_slab(a, by, bz, bx)
        if ex:
            return r + b * bz + k
        if b > 0:
            r += bz
        if b > 1:
            r += (b - 1) * (2 if bz > 1 else 1)
        return r + (k if ey else (0 if k == 0 else 1))

    for bx, by, bz in itertools.product(range(1, 7), repeat=3):
        scan = [(a, b, k) for a in range(bx) for b in range(by) for k in range(bz)
                if a in (0, bx - 1) or b in (0, by - 1) or k in (0, bz - 1)]
        got = {}
        for a, b, k in itertools.product(range(bx), range(by), range(bz)):
            r = rank(a, b, k, bx, by, bz)
            if r is not None:
                assert r not in got
                got[r] = (a, b, k)
        assert [got[i] for i in range(len(scan))] == scan and before_slab(bx, by, bz, bx) == len(scan), (bx, by, bz)


def test_group_of_32_acceptance_scan_equals_the_sequential_scan():
    """cluster_loop_kernel's phase 3 resolves the reference's order-dependent acceptance scan (cluster_server.cu:693-737) 32 candidates at
    a time: conflicts with groups already resolved come as one bit per candidate (prehit + the previous group's word), the dependence
    inside a group is settled by rounds of 'rejected if it conflicts with an accepted lane, accepted if it conflicts with no accepted
    and no still-undecided lower lane'.  The same procedure in Python against the plain sequential scan on random conflict sets."""
    rng = np.random.default_rng(7)
    for trial in range(60):
        C = int(rng.integers(1, 150))
        can_clu = rng.random(C) < rng.choice([0.3, 0.7, 1.0])
        dens = rng.choice([0.0, 0.02, 0.2, 0.6])
        conflict = np.tril(rng.random((C, C)) < dens, k=-1)           # conflict[i][j], j < i: the ray between i and j is blocked
        conflict &= can_clu[None, :] & can_clu[:, None]               # phase 2b only traces rays between surviving candidates
        seq = np.zeros(C, bool)
        for i in range(C):
            seq[i] = can_clu[i] and not (conflict[i, :i] & seq[:i]).any()
        acc = np.zeros(C, bool)
        for b in range(0, C, 32):
            lanes = range(b, min(b + 32, C))
            alive = {i: bool(can_clu[i]) and not (conflict[i, :b] & acc[:b]).any() for i in lanes}
            und = {i for i in lanes if alive[i]}
            accepted = set()
            rounds = 0
            while und:
                rounds += 1
                rej = {i for i in und if any(conflict[i, j] for j in accepted)}
                now = {i for i in und - rej if not any(conflict[i, j] for j in und if j < i)}
                assert rej or now                                     # the lowest undecided lane always decides
                accepted |= now
                und -= rej | now
            assert rounds <= 32
            for i in accepted:
                acc[i] = True
        assert np.array_equal(acc, seq), trial


def test_ray_walk_reaches_its_target_after_exactly_l1_steps_and_lookahead_reads_nothing_else():
    """voxel.cuh's ray walk replaces the reference's 'position == end' tests by a step counter (|dx| + |dy| + |dz| steps) and examines the
    voxels in groups of four fetched ahead.  Checked here on the reference's own fp64 walk (py_ray's arithmetic): the walk is at its
    target after exactly that many steps and not before, for axis-aligned, diagonal, tie-prone and random directions up to the node's
    map size - so the look-ahead never steps past the target and the counter is the reference's test."""
    rng = np.random.default_rng(3)
    dirs = [(d, 0, 0) for d in (1, -7, 333)] + [(5, 5, 5), (-9, 9, 0), (1, 3, 9), (3, 1, -9), (2, 4, 8), (6, -3, 2), (333, 333, 33), (1, 1, 332)]
    dirs += [tuple(int(v) for v in rng.integers(-333, 334, 3)) for _ in range(1500)]
    dirs += [tuple(int(v) for v in rng.integers(-6, 7, 3)) for _ in range(1500)]
    for d in dirs:
        if d == (0, 0, 0):
            continue
        step = [(t > 0) - (t < 0) for t in d]
        tmax = [99999.0 if t == 0 else 0.5 / abs(t) for t in d]
        delta = [0.0 if t == 0 else 1.0 / abs(t) for t in d]          # the kernel's table: 1 / |d|; tMax = 0.5 * (1 / |d|) exactly
        assert all(t == 0 or 0.5 / abs(t) == 0.5 * (1.0 / abs(t)) for t in d)
        p = [0, 0, 0]
        n = sum(abs(t) for t in d)
        for s in range(n):
            assert p != list(d), (d, s)
            if tmax[0] < tmax[1]:
                k = 0 if tmax[0] < tmax[2] else 2
            else:
                k = 1 if tmax[1] < tmax[2] else 2
            assert d[k] != 0                                          # an axis with d = 0 is never stepped
            p[k] += step[k]; tmax[k] += delta[k]
        assert p == list(d), d
