"""Voxel-map kernels: ctypes binding of include/direct_voxel.h and a seeded synthetic map generator.

The entry points rebuild the reference's only CUDA code (polyhedron_generator/src/cluster_engine.cu, cluster_server.cu) for
sm_100a; host-side helpers here mirror the reference's host code around the kernels (vertex layout of setVertexInitIndex,
getVoxelsInCube + shell extraction of polygonGeneration, cluster_server.cu:788-890) so that tests and tools can drive the kernels
the way the node does.  No CPU fallback: every call goes through libdirect_ddp_b200.so."""
from __future__ import annotations

import ctypes as C

import numpy as np

VOXEL_EXPORTS = ["direct_voxel_convex_test", "direct_voxel_convex_test_device", "direct_voxel_cube_inflation",
                 "direct_voxel_cube_inflation_device", "direct_voxel_inflate_box", "direct_voxel_inflate_box_device",
                 "direct_voxel_cluster", "direct_voxel_cluster_device", "direct_voxel_cluster_phases", "direct_voxel_polytope"]


class MapC(C.Structure):
    _fields_ = [("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int), ("occupied", C.c_void_p), ("inside", C.c_void_p)]


def _u8(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return a


def _map_struct(occ: np.ndarray, inside: np.ndarray | None, keep: list) -> MapC:
    occ = _u8(occ); keep.append(occ)
    if inside is not None:
        inside = _u8(inside); keep.append(inside)
        assert inside.shape == occ.shape
    nx, ny, nz = occ.shape
    return MapC(nx, ny, nz, occ.ctypes.data, None if inside is None else inside.ctypes.data)


def convex_test(solver, occ, inside, cand, clu, can_can_fill: int = 2):
    """paraConvexTest + paraResultCheck.  Returns (can_can [C (C + 1) / 2] uint8, can_clu [C] uint8); can_can entries the
    reference never writes keep `can_can_fill`."""
    keep = []
    m = _map_struct(occ, inside, keep)
    cand = np.ascontiguousarray(cand, dtype=np.int32).reshape(-1, 3); clu = np.ascontiguousarray(clu, dtype=np.int32).reshape(-1, 3)
    Cn, K = len(cand), len(clu)
    cc = np.full(Cn * (Cn + 1) // 2, can_can_fill, np.uint8); cl = np.zeros(Cn, np.uint8)
    solver._check(solver.lib.direct_voxel_convex_test(solver.h, C.byref(m), C.c_void_p(cand.ctypes.data), Cn, C.c_void_p(clu.ctypes.data), K,
                                                      C.c_void_p(cc.ctypes.data), C.c_void_p(cl.ctypes.data)))
    return cc, cl


def cube_inflation(solver, occ, vertex_idx, direction: int, inf_step: int = 1) -> int:
    keep = []
    m = _map_struct(occ, None, keep)
    v = np.ascontiguousarray(vertex_idx, dtype=np.int32); assert v.size == 24
    r = C.c_int32(-1)
    solver._check(solver.lib.direct_voxel_cube_inflation(solver.h, C.byref(m), C.c_void_p(v.ctypes.data), direction, inf_step, C.byref(r)))
    return r.value


def inflate_box(solver, occ, vertex_idx, itr_inflate_max: int, inf_step: int = 1):
    """cubeInflation_gpu.  Returns (vertex_idx [24] after inflation, outer iterations run)."""
    keep = []
    m = _map_struct(occ, None, keep)
    v = np.array(vertex_idx, dtype=np.int32); assert v.size == 24
    it = C.c_int32(0)
    solver._check(solver.lib.direct_voxel_inflate_box(solver.h, C.byref(m), C.c_void_p(v.ctypes.data), inf_step, itr_inflate_max, C.byref(it)))
    return v, it.value


def cluster(solver, occ, inside, use, invalid, cluster_xyz, cap: int, cand_cap: int, itr_cluster_max: int):
    """polytopeCluster_gpu.  Returns (cluster_xyz [n][3], use, invalid, iterations)."""
    keep = []
    m = _map_struct(occ, inside, keep)
    use = np.array(use, dtype=np.uint8); invalid = np.array(invalid, dtype=np.uint8)
    init = np.ascontiguousarray(cluster_xyz, dtype=np.int32).reshape(-1, 3)
    buf = np.zeros((cap, 3), np.int32); buf[:len(init)] = init
    n, it = C.c_int32(len(init)), C.c_int32(0)
    solver._check(solver.lib.direct_voxel_cluster(solver.h, C.byref(m), C.c_void_p(use.ctypes.data), C.c_void_p(invalid.ctypes.data),
                                                  C.c_void_p(buf.ctypes.data), C.byref(n), cap, cand_cap, itr_cluster_max, C.byref(it)))
    return buf[:n.value].copy(), use, invalid, it.value


def polytope(solver, occ, seed, itr_inflate_max: int, itr_cluster_max: int, cap: int, cand_cap: int, flags: bool = True) -> dict:
    """cudaPolytopeGeneration::polygonGeneration for a one-voxel seed, four launches, no host round trip in between.
    Returns dict(cluster [n][3], vertex_idx [24], iters [inflation, clustering], inside, use, invalid (None unless flags))."""
    keep = []
    m = _map_struct(occ, None, keep)
    seed = np.ascontiguousarray(seed, dtype=np.int32); assert seed.size == 3
    buf = np.zeros((cap, 3), np.int32); n = C.c_int32(0); it = (C.c_int32 * 2)(); v = np.zeros(24, np.int32)
    fl = [np.zeros(np.shape(occ), np.uint8) if flags else None for _ in range(3)]
    solver._check(solver.lib.direct_voxel_polytope(solver.h, C.byref(m), C.c_void_p(seed.ctypes.data), itr_inflate_max, itr_cluster_max, cap, cand_cap,
                                                   C.c_void_p(buf.ctypes.data), C.byref(n), it, C.c_void_p(v.ctypes.data),
                                                   *[C.c_void_p(f.ctypes.data) if f is not None else None for f in fl]))
    return dict(cluster=buf[:n.value].copy(), vertex_idx=v, iters=[it[0], it[1]], inside=fl[0], use=fl[1], invalid=fl[2])


def cluster_phases(solver) -> dict:
    ms = (C.c_double * 5)()
    solver._check(solver.lib.direct_voxel_cluster_phases(solver.h, ms))
    return dict(zip(("claims", "compaction", "cluster_rays", "candidate_rays", "acceptance"), [float(v) for v in ms]))


# ---- host-side mirrors of the reference's host code around the kernels ---------------------------------------------------------
def box_vertices(xmin, ymin, zmin, xmax, ymax, zmax) -> np.ndarray:
    """vertex_idx [24] = x of p1..p8, y of p1..p8, z of p1..p8 (setVertexInitIndex, cluster_server.cu:205-228: p1 = (xmax, ymin,
    zmax), p2 = (xmax, ymax, zmax), p3 = (xmin, ymax, zmax), p4 = (xmin, ymin, zmax), p5..p8 the same at zmin)."""
    x = [xmax, xmax, xmin, xmin] * 2
    y = [ymin, ymax, ymax, ymin] * 2
    z = [zmax] * 4 + [zmin] * 4
    return np.array(x + y + z, dtype=np.int32)


def box_bounds(v):
    v = np.asarray(v)
    return int(v[3]), int(v[8]), int(v[20]), int(v[0]), int(v[9]), int(v[16])   # xmin, ymin, zmin, xmax, ymax, zmax


def cube_shell(shape, v):
    """polygonGeneration between its two GPU stages (cluster_server.cu:822-895): the voxels of the inflated box become `inside`
    and `use`; those with a neighbour outside the box (or the map) are the initial cluster and lose their `inside` flag.
    Returns (inside, use, shell_xyz [n][3] in the reference's x, y, z scan order)."""
    nx, ny, nz = shape
    x0, y0, z0, x1, y1, z1 = box_bounds(v)
    inside = np.zeros(shape, np.uint8); use = np.zeros(shape, np.uint8)
    inside[x0:x1 + 1, y0:y1 + 1, z0:z1 + 1] = 1
    xs, ys, zs = np.meshgrid(np.arange(x0, x1 + 1), np.arange(y0, y1 + 1), np.arange(z0, z1 + 1), indexing="ij")
    cells = np.stack([xs.ravel(), ys.ravel(), zs.ravel()], axis=1).astype(np.int32)
    if len(cells) == 1:
        shell = cells
    else:
        use[x0:x1 + 1, y0:y1 + 1, z0:z1 + 1] = 1
        pad = np.zeros((nx + 2, ny + 2, nz + 2), np.uint8); pad[1:-1, 1:-1, 1:-1] = inside
        full = np.ones(len(cells), bool)
        for dx in (-1, 0, 1):
            for dy in (-1, 0, 1):
                for dz in (-1, 0, 1):
                    full &= pad[cells[:, 0] + 1 + dx, cells[:, 1] + 1 + dy, cells[:, 2] + 1 + dz] == 1
        shell = cells[~full]
    inside[shell[:, 0], shell[:, 1], shell[:, 2]] = 0
    return inside, use, np.ascontiguousarray(shell)


def make_map(shape=(120, 120, 30), n_pillars: int = 60, seed: int = 6, clear=None) -> np.ndarray:
    """Seeded pillar / slab obstacle map in the spirit of the reference's random_complex_generator (pillars of random footprint and
    height, rng seed 6 in map_generator.launch:24); `clear` = (x, y, z, r) keeps a ball of radius r free."""
    rng = np.random.default_rng(seed)
    nx, ny, nz = shape
    occ = np.zeros(shape, np.uint8)
    for _ in range(n_pillars):
        w = int(rng.integers(2, 7)); d = int(rng.integers(2, 7)); hgt = int(rng.integers(nz // 3, nz + 1))
        x = int(rng.integers(0, nx - w)); y = int(rng.integers(0, ny - d))
        if rng.random() < 0.25:   # a floating slab instead of a pillar
            z = int(rng.integers(nz // 3, nz - 2)); occ[x:x + 3 * w, y:y + d, z:z + 2] = 1
        else:
            occ[x:x + w, y:y + d, 0:hgt] = 1
    if clear is not None:
        cx, cy, cz, r = clear
        xs, ys, zs = np.ogrid[:nx, :ny, :nz]
        occ[(xs - cx) ** 2 + (ys - cy) ** 2 + (zs - cz) ** 2 <= r * r] = 0
    return occ
