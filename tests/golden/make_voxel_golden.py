"""Writes tests/golden/voxel_ref_*.npz: inputs and outputs of the REFERENCE's own CUDA kernels (paraConvexTest + paraResultCheck,
paraCubeInflation; polyhedron_generator/src/cluster_engine.cu compiled unmodified for sm_100a into oracle/_ref/libvoxel_ref.so).

Needs a GPU, so it runs on the GPU box:  gpurun -- 'python tests/golden/make_voxel_golden.py gpurun_out/golden'  and the files
are then copied into tests/golden/.  The CPU suite checks oracle/voxel_oracle.c against them (tests/test_voxel.py)."""
import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)

from direct_b200 import voxel as X   # noqa: E402  (map generator and host-side mirrors only)
from oracle import voxel_py as V     # noqa: E402


def first_candidates(occ, inside, use, shell):
    """The candidate list of the first clustering iteration (cluster_server.cu:573-626)."""
    nx, ny, nz = occ.shape
    use = use.copy(); out = []
    for cx, cy, cz in shell:
        use[cx, cy, cz] = 1
        for dx in (-1, 0, 1):
            for dy in (-1, 0, 1):
                for dz in (-1, 0, 1):
                    x, y, z = cx + dx, cy + dy, cz + dz
                    if (dx, dy, dz) == (0, 0, 0) or not (0 <= x < nx and 0 <= y < ny and 0 <= z < nz):
                        continue
                    if occ[x, y, z] == 1 or use[x, y, z] == 1 or inside[x, y, z] == 1:
                        continue
                    out.append((x, y, z)); use[x, y, z] = 1
    return np.array(out, np.int32).reshape(-1, 3)


def cases():
    for name, shape, pillars, seed, seed_cell in (("a", (40, 40, 12), 14, 3, (20, 20, 5)), ("b", (48, 36, 10), 20, 11, (10, 25, 4)),
                                                  ("c", (30, 30, 30), 10, 5, (15, 15, 15))):
        occ = X.make_map(shape, pillars, seed, clear=(*seed_cell, 3))
        v, _ = V.inflate_box(occ, X.box_vertices(*seed_cell, *seed_cell), 6)
        inside, use, shell = X.cube_shell(shape, v)
        cand = first_candidates(occ, inside, use, shell)
        yield name, occ, v, inside, shell, cand


def main(out_dir):
    os.makedirs(out_dir, exist_ok=True)
    for name, occ, v, inside, shell, cand in cases():
        cc, cl, _ = V.ref_convex_test(occ, inside, cand, shell)
        infl = np.array([[V.ref_cube_inflation(occ, v, d) for d in range(6)]], np.int32)
        np.savez_compressed(os.path.join(out_dir, f"voxel_ref_{name}.npz"), occ=occ, inside=inside, vertex_idx=v, cluster=shell, cand=cand,
                            can_can=cc, can_clu=cl, inflation=infl)
        print(name, occ.shape, "cand", len(cand), "cluster", len(shell), "can_clu true", int(cl.sum()), "can_can true",
              int((cc == 1).sum()), "inflation", infl.tolist())


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden"))
