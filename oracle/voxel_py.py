"""ctypes front end of oracle/libvoxel_oracle.so (CPU restatement of the reference's voxel kernels) and, on a GPU box, of
oracle/_ref/libvoxel_ref.so (the reference's own CUDA kernels compiled unmodified for sm_100a).  TEST INFRASTRUCTURE."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_PG = "/root/reference/polyhedron_generator"
REF_SO = os.path.join(_HERE, "_ref", "libvoxel_ref.so")
REF_SERVER_SO = os.path.join(_HERE, "_ref", "libvoxel_server_ref.so")


def build(ref: bool = False):
    so = os.path.join(_HERE, "libvoxel_oracle.so")
    src = [os.path.join(_HERE, f) for f in ("voxel_oracle.c", "voxel_oracle.h")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src):
        subprocess.check_call(["gcc", "-O2", "-march=x86-64-v2", "-fPIC", "-Wall", "-std=c11", "-ffp-contract=off", "-shared", "-o", so,
                               src[0], "-lm"])
    if ref and os.path.isdir(REF_PG):   # the reference's kernels, where the reference tree is mounted (this container only)
        drv = os.path.join(_HERE, "voxel_ref_driver.cu")
        eng = os.path.join(REF_PG, "src", "cluster_engine.cu")
        if not os.path.exists(REF_SO) or os.path.getmtime(drv) > os.path.getmtime(REF_SO):
            os.makedirs(os.path.dirname(REF_SO), exist_ok=True)
            # -O3 -use_fast_math are the reference's own flags (polyhedron_generator/CMakeLists.txt:19-31); only the arch differs
            subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-use_fast_math", "-shared", "-Xcompiler", "-fPIC", "-w",
                                   "-I", os.path.join(REF_PG, "include"), drv, eng, "-o", REF_SO])
        # the reference's host loop around those kernels (cluster_server.cu, unmodified; ROS only as the shim's stopwatch)
        sdrv = os.path.join(_HERE, "voxel_server_ref_driver.cu")
        if not os.path.exists(REF_SERVER_SO) or os.path.getmtime(sdrv) > os.path.getmtime(REF_SERVER_SO):
            src = [os.path.join(REF_PG, "src", f) for f in ("cluster_server.cu", "cluster_engine.cu")]
            obj = os.path.join(_HERE, "_ref", "cluster_engine_cpu.o")   # gcc 13 no longer pulls uint8_t in through <iostream>
            subprocess.check_call(["g++", "-O3", "-fPIC", "-w", "-include", "cstdint", "-I", os.path.join(REF_PG, "include"), "-c",
                                   os.path.join(REF_PG, "src", "cluster_engine_cpu.cpp"), "-o", obj])
            subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-use_fast_math", "-shared", "-Xcompiler", "-fPIC", "-w",
                                   "-I", os.path.join(_HERE, "shim"), "-I", os.path.join(REF_PG, "include"), sdrv, *src, obj, "-o", REF_SERVER_SO])
            os.remove(obj)
    return so


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.voxel_oracle_cube_inflation.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_longlong]
    return _lib


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def convex_test(occ, inside, cand, clu, can_can_fill=2):
    occ = np.ascontiguousarray(occ, np.uint8); inside = np.ascontiguousarray(inside, np.uint8)
    cand = _i32(cand).reshape(-1, 3); clu = _i32(clu).reshape(-1, 3)
    Cn, K = len(cand), len(clu)
    cc = np.full(Cn * (Cn + 1) // 2, can_can_fill, np.uint8); cl = np.zeros(Cn, np.uint8)
    lib().voxel_oracle_convex_test(C.c_void_p(occ.ctypes.data), C.c_void_p(inside.ctypes.data), occ.shape[1], occ.shape[2],
                                   C.c_void_p(cand.ctypes.data), Cn, C.c_void_p(clu.ctypes.data), K, C.c_void_p(cc.ctypes.data),
                                   C.c_void_p(cl.ctypes.data))
    return cc, cl


def cube_inflation(occ, v, direction, inf_step=1, max_threads=0):
    occ = np.ascontiguousarray(occ, np.uint8); v = _i32(v)
    return lib().voxel_oracle_cube_inflation(C.c_void_p(occ.ctypes.data), occ.shape[1], occ.shape[2], C.c_void_p(v.ctypes.data), direction,
                                             inf_step, max_threads)


def inflate_box(occ, v, itr_inflate_max, inf_step=1):
    occ = np.ascontiguousarray(occ, np.uint8); v = np.array(v, dtype=np.int32)
    it = lib().voxel_oracle_inflate_box(C.c_void_p(occ.ctypes.data), *occ.shape, C.c_void_p(v.ctypes.data), inf_step, itr_inflate_max)
    return v, it


def cluster(occ, inside, use, invalid, cluster_xyz, cap, cand_cap, itr_cluster_max):
    occ = np.ascontiguousarray(occ, np.uint8); inside = np.ascontiguousarray(inside, np.uint8)
    use = np.array(use, dtype=np.uint8); invalid = np.array(invalid, dtype=np.uint8)
    init = _i32(cluster_xyz).reshape(-1, 3)
    buf = np.zeros((cap, 3), np.int32); buf[:len(init)] = init
    it = C.c_int(0)
    n = lib().voxel_oracle_cluster(C.c_void_p(occ.ctypes.data), C.c_void_p(inside.ctypes.data), C.c_void_p(use.ctypes.data),
                                   C.c_void_p(invalid.ctypes.data), *occ.shape, C.c_void_p(buf.ctypes.data), len(init), cap, cand_cap,
                                   itr_cluster_max, C.byref(it))
    if n < 0:
        raise RuntimeError("cluster or candidate capacity exceeded")
    return buf[:n].copy(), use, invalid, it.value


def polytope(occ, seed, itr_inflate_max, itr_cluster_max, cap, cand_cap, host_can_can=None):
    """Returns dict(cluster, vertex_idx, iters, inside, use, invalid).  host_can_can (uint8 [cand_cap (cand_cap + 1) / 2], see
    reference_host_buffer): reproduce the reference host loop's stale read of the last candidate row (voxel_oracle.h)."""
    occ = np.ascontiguousarray(occ, np.uint8)
    seed = _i32(seed); buf = np.zeros((cap, 3), np.int32); v = np.zeros(24, np.int32); it = (C.c_int * 2)()
    fl = [np.zeros(occ.shape, np.uint8) for _ in range(3)]
    if host_can_can is not None:
        assert host_can_can.dtype == np.uint8 and host_can_can.size >= cand_cap * (cand_cap + 1) // 2 and host_can_can.flags.c_contiguous
    n = lib().voxel_oracle_polytope_hostbuf(C.c_void_p(occ.ctypes.data), *occ.shape, C.c_void_p(seed.ctypes.data), itr_inflate_max,
                                            itr_cluster_max, cap, cand_cap, C.c_void_p(buf.ctypes.data), C.c_void_p(v.ctypes.data), it,
                                            *[C.c_void_p(f.ctypes.data) for f in fl],
                                            C.c_void_p(host_can_can.ctypes.data) if host_can_can is not None else None)
    if n < 0:
        raise RuntimeError("cluster or candidate capacity exceeded")
    return dict(cluster=buf[:n].copy(), vertex_idx=v, iters=[it[0], it[1]], inside=fl[0], use=fl[1], invalid=fl[2])


def reference_host_buffer(cand_cap=10000):
    """A fresh h_can_can_result as the reference's generator object allocates it (cluster_server.cu:145), zero-filled."""
    return np.zeros(cand_cap * (cand_cap + 1) // 2, np.uint8)


# ---- the reference's own kernels (GPU box only) ---------------------------------------------------------------------------------
_ref = None


def ref_lib():
    global _ref
    if _ref is None:
        if not os.path.exists(REF_SO):
            raise FileNotFoundError(REF_SO + " (built by oracle.voxel_py.build(ref=True) where /root/reference is mounted)")
        _ref = C.CDLL(REF_SO)
    return _ref


def ref_convex_test(occ, inside, cand, clu, can_can_fill=2):
    occ = np.ascontiguousarray(occ, np.uint8); inside = np.ascontiguousarray(inside, np.uint8)
    cand = _i32(cand).reshape(-1, 3); clu = _i32(clu).reshape(-1, 3)
    Cn, K = len(cand), len(clu)
    cc = np.full(Cn * (Cn + 1) // 2, can_can_fill, np.uint8); cl = np.zeros(Cn, np.uint8)
    ms = C.c_float(0)
    st = ref_lib().voxel_ref_convex_test(C.c_void_p(occ.ctypes.data), C.c_void_p(inside.ctypes.data), *occ.shape, C.c_void_p(cand.ctypes.data),
                                         Cn, C.c_void_p(clu.ctypes.data), K, C.c_void_p(cc.ctypes.data), C.c_void_p(cl.ctypes.data),
                                         C.byref(ms))
    if st:
        raise RuntimeError(f"reference kernels: cudaError {st}")
    return cc, cl, ms.value


def ref_cube_inflation(occ, v, direction, inf_step=1):
    occ = np.ascontiguousarray(occ, np.uint8); v = _i32(v)
    r = C.c_int(-1)
    st = ref_lib().voxel_ref_cube_inflation(C.c_void_p(occ.ctypes.data), *occ.shape, C.c_void_p(v.ctypes.data), direction, inf_step, C.byref(r))
    if st:
        raise RuntimeError(f"reference kernels: cudaError {st}")
    return r.value


def ref_inflate_box(occ, v, itr_inflate_max, inf_step=1):
    occ = np.ascontiguousarray(occ, np.uint8); v = np.array(v, dtype=np.int32)
    it, sec = C.c_int(0), C.c_double(0)
    st = ref_lib().voxel_ref_inflate_box(C.c_void_p(occ.ctypes.data), *occ.shape, C.c_void_p(v.ctypes.data), inf_step, itr_inflate_max,
                                         C.byref(it), C.byref(sec))
    if st:
        raise RuntimeError(f"reference kernels: cudaError {st}")
    return v, it.value, sec.value


# ---- the reference's own host loop, cudaPolytopeGeneration::polygonGeneration (GPU box only) -------------------------------------
_ref_server = None


def ref_server_available() -> bool:
    return os.path.exists(REF_SERVER_SO)


def ref_server_polytope(occ, seed, itr_inflate_max, itr_cluster_max, resolution=0.1, reps=1, cap=50000):
    """The reference's polygonGeneration (cluster_server.cu:769-966) from a one-voxel seed on map `occ`, called `reps` times on ONE
    generator object.  Returns (list of the clusters [n][3] of every call, in the reference's order; best wall seconds)."""
    global _ref_server
    if _ref_server is None:
        if not ref_server_available():
            raise FileNotFoundError(REF_SERVER_SO + " (built by oracle.voxel_py.build(ref=True) where /root/reference is mounted)")
        _ref_server = C.CDLL(REF_SERVER_SO)
    occ = np.ascontiguousarray(occ, np.uint8); seed = _i32(seed)
    st = _ref_server.voxel_server_ref_setup(C.c_void_p(occ.ctypes.data), *occ.shape, C.c_double(resolution), int(itr_inflate_max),
                                            int(itr_cluster_max))
    if st:
        raise RuntimeError(f"reference host loop: cudaError {st}")
    buf = np.zeros((cap, 3), np.int32); best = float("inf"); out = []
    for _ in range(reps):
        sec = C.c_double(0)
        n = _ref_server.voxel_server_ref_polytope(C.c_void_p(seed.ctypes.data), C.c_void_p(buf.ctypes.data), cap, C.byref(sec))
        if n < 0:
            raise RuntimeError(f"reference host loop failed ({n})")
        best = min(best, sec.value)
        out.append(buf[:n].copy())
    _ref_server.voxel_server_ref_release()
    return out, best
