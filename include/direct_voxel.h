/*
 * direct_voxel.h -- C-ABI of the voxel-map kernels of libdirect_ddp_b200.so.
 *
 * The reference's only CUDA code, polyhedron_generator/src/cluster_engine.cu (written for sm_75, synchronous, switched off
 * in its build), rebuilt for sm_100a:
 *   direct_voxel_convex_test      = paraConvexTest (cluster_engine.cu:70-180) + paraResultCheck (:37-67) fused:
 *       for every candidate voxel, a 3-D DDA ray (Amanatides-Woo, fp64 like the reference) to every EARLIER candidate and
 *       to every voxel of the cluster; a ray fails when it crosses an occupied voxel before it reaches its target or an
 *       `inside` voxel.  Outputs exactly what the reference's pair of kernels leaves in d_can_can_result / d_can_clu_result:
 *       can_can[t (t + 1) / 2 + i] for i < t  (the reference's own packing, cluster_engine.cu:46-52: bias n (n - 1) / 2
 *       with n = t + 1), can_clu[t] = every cluster ray of candidate t is free.  The cand x (cand + clu) intermediate
 *       array of the reference never exists.
 *   direct_voxel_cube_inflation   = paraCubeInflation (cluster_engine.cu:185-349, launched cluster_server.cu:387): can
 *       face `dir` (0 Y-, 1 Y+, 2 X-, 3 X+, 4 Z-, 5 Z+) of the box given by its 8 vertex indices (vertex_idx[0..7] = x,
 *       [8..15] = y, [16..23] = z of the vertices) move out by inf_step voxels without touching an occupied voxel.
 *   direct_voxel_inflate_box      = cubeInflation_gpu (cluster_server.cu:343-440): the whole six-direction loop around
 *       paraCubeInflation, until the box stops growing or itr_inflate_max outer iterations have run, as ONE kernel launch (the
 *       reference uploads 96 bytes, launches, synchronises and downloads one byte per direction per iteration).
 *   direct_voxel_cluster          = polytopeCluster_gpu (cluster_server.cu:556-767): the whole convex-clustering loop (candidate
 *       generation, ray tests, acceptance scan, cluster update; up to itr_cluster_max iterations) as ONE cooperative kernel
 *       launch; nothing crosses PCIe inside the loop.  Same cluster, voxel for voxel and in the same order, as the reference's
 *       loop with every can_can entry read from the kernel's output (the reference downloads C (C - 1) / 2 entries and indexes
 *       up to C (C + 1) / 2 - 2, so its last candidate row is stale host memory; oracle/voxel_oracle.c does what this does).
 * Map layout as in the reference: uint8 occupancy, index x * (ny * nz) + y * nz + z.
 * Host-buffer entry points copy inputs and outputs; *_device variants take device pointers and a cudaStream_t.
 * Handles and status codes: direct_ddp.h.  No CPU fallback.
 */
#ifndef DIRECT_VOXEL_H_
#define DIRECT_VOXEL_H_

#include "direct_ddp.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct direct_voxel_map {
    int nx, ny, nz;
    const uint8_t *occupied; /* [nx * ny * nz] map_data    */
    const uint8_t *inside;   /* [nx * ny * nz] inside_data */
} direct_voxel_map;

/* can_can: [C (C + 1) / 2] bytes (entries the reference never writes are left untouched), can_clu: [C] bytes. */
int direct_voxel_convex_test(direct_ddp_handle h, const direct_voxel_map *map, const int32_t *candidate_xyz, int C,
                             const int32_t *cluster_xyz, int K, uint8_t *can_can, uint8_t *can_clu);
int direct_voxel_convex_test_device(direct_ddp_handle h, const direct_voxel_map *map, const int32_t *candidate_xyz, int C,
                                    const int32_t *cluster_xyz, int K, uint8_t *can_can, uint8_t *can_clu, void *stream);
/* *result = 1 when the face can be inflated, 0 otherwise. */
int direct_voxel_cube_inflation(direct_ddp_handle h, const direct_voxel_map *map, const int32_t *vertex_idx /*[24]*/, int dir,
                                int inf_step, int32_t *result);
int direct_voxel_cube_inflation_device(direct_ddp_handle h, const direct_voxel_map *map, const int32_t *vertex_idx, int dir,
                                       int inf_step, int32_t *result, void *stream);

/* vertex_idx [24] in/out; *iters (optional) = outer iterations run.  inf_step must be 1 (the reference's constant,
 * cluster_server.cu:173: its boundary tests are only right for 1). */
int direct_voxel_inflate_box(direct_ddp_handle h, const direct_voxel_map *map, int32_t *vertex_idx /*[24]*/, int inf_step,
                             int itr_inflate_max, int32_t *iters);
int direct_voxel_inflate_box_device(direct_ddp_handle h, const direct_voxel_map *map, int32_t *vertex_idx, int inf_step,
                                    int itr_inflate_max, int32_t *iters, void *stream);

/* use / invalid: [nx * ny * nz] flag arrays of the reference (use_data, invalid_data), in/out.  cluster_xyz: [cap][3], the first
 * *cluster_num voxels are the initial cluster (all active, cluster_server.cu:897-909); on return *cluster_num is the new size and the
 * accepted voxels follow in acceptance order.  cand_cap <= 32768 bounds the candidates of one iteration.  *iters (optional) =
 * iterations completed.  Returns DIRECT_DDP_ERR_ARG when cap or cand_cap would be exceeded (outputs then undefined). */
int direct_voxel_cluster(direct_ddp_handle h, const direct_voxel_map *map, uint8_t *use, uint8_t *invalid, int32_t *cluster_xyz,
                         int32_t *cluster_num, int cap, int cand_cap, int itr_cluster_max, int32_t *iters);
/* Device pointers; ctl: int32 [8] on the device, ctl[0] = cluster_num in/out, ctl[1] = iterations out, ctl[2] = 0 or -1 (overflow)
 * out, the rest scratch.  Scratch buffers belong to the handle.  Asynchronous on `stream`. */
int direct_voxel_cluster_device(direct_ddp_handle h, const direct_voxel_map *map, uint8_t *use, uint8_t *invalid, int32_t *cluster_xyz,
                                int32_t *ctl, int cap, int cand_cap, int itr_cluster_max, void *stream);

/* cudaPolytopeGeneration::polygonGeneration (cluster_server.cu:769-966) for a one-voxel seed, the case the node uses per path point:
 * flagClear, box inflation from the seed (<= itr_inflate_max outer iterations), the box's voxels -> inside / use flags and its boundary
 * voxels -> initial cluster, then (unless the box is one voxel thick, :911-920) the clustering loop (<= itr_cluster_max iterations).
 * Four launches on one stream, nothing but the seed goes up and nothing but the result comes down.  cluster_xyz [cap][3] receives the
 * *cluster_num voxels of the polytope in the reference's order; inside / use / invalid ([nx ny nz], optional, may be NULL) receive the
 * reference's flag arrays; vertex_idx [24] (optional) the inflated box.  (Seeds of several voxels are not accepted: the reference's
 * bounding box of such a seed is -100000 .. 100000, cluster_server.cu:804-816.) */
int direct_voxel_polytope(direct_ddp_handle h, const direct_voxel_map *map /* inside ignored */, const int32_t seed_xyz[3],
                          int itr_inflate_max, int itr_cluster_max, int cap, int cand_cap, int32_t *cluster_xyz, int32_t *cluster_num,
                          int32_t *iters /* optional [2]: inflation, clustering */, int32_t *vertex_idx, uint8_t *inside, uint8_t *use,
                          uint8_t *invalid);

/* Device time (ms) of the five phases of the last clustering launch on this handle, summed over its iterations: neighbour claims,
 * ordered candidate compaction, candidate -> cluster rays, candidate -> candidate rays, acceptance scan. */
int direct_voxel_cluster_phases(direct_ddp_handle h, double ms[5]);

#ifdef __cplusplus
}
#endif
#endif
