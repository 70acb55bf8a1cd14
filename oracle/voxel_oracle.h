/* voxel_oracle.h -- CPU restatement of the reference's voxel-map CUDA kernels and of the host loops around them
 * (polyhedron_generator/src/cluster_engine.cu, cluster_server.cu).  TEST INFRASTRUCTURE ONLY (see voxel_oracle.c). */
#ifndef VOXEL_ORACLE_H_
#define VOXEL_ORACLE_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* paraConvexTest + paraResultCheck (cluster_engine.cu:70-180, :37-67).  can_can: [C (C + 1) / 2] (only the entries the
 * reference writes are touched), can_clu: [C]. */
void voxel_oracle_convex_test(const uint8_t *occ, const uint8_t *inside, int ny, int nz, const int32_t *cand, int C,
                              const int32_t *clu, int K, uint8_t *can_can, uint8_t *can_clu);

/* paraCubeInflation (cluster_engine.cu:185-349).  max_threads > 0 restates the reference's fixed <<<128, 128>>> launch
 * (cluster_server.cu:170-171: only the first max_threads cells of the face are looked at); 0 = every cell. */
int voxel_oracle_cube_inflation(const uint8_t *occ, int ny, int nz, const int32_t *vertex_idx, int dir, int inf_step,
                                long long max_threads);

/* cubeInflation_gpu (cluster_server.cu:343-440): the six-direction loop until the box stops growing. vertex_idx[24] in/out.
 * Returns the number of outer iterations run. */
int voxel_oracle_inflate_box(const uint8_t *occ, int nx, int ny, int nz, int32_t *vertex_idx, int inf_step, int itr_inflate_max);

/* polytopeCluster_gpu (cluster_server.cu:556-767) with every can_can entry read from the kernel's own output (the reference
 * downloads C (C - 1) / 2 entries and indexes up to C (C + 1) / 2 - 2: its last candidate row is stale host memory).
 * cluster_xyz [cap][3] holds cluster_num voxels on entry (all of them active), more on return; inside / use / invalid are
 * the reference's per-voxel flag arrays (inside is read only).  Returns the new cluster size, or -1 when cap or cand_cap
 * would be exceeded.  iters_out (optional): iterations run. */
int voxel_oracle_cluster(const uint8_t *occ, const uint8_t *inside, uint8_t *use, uint8_t *invalid, int nx, int ny, int nz,
                         int32_t *cluster_xyz, int cluster_num, int cap, int cand_cap, int itr_cluster_max, int *iters_out);

/* polygonGeneration (cluster_server.cu:769-966) for a one-voxel seed: flagClear, box inflation, getVoxelsInCube + boundary extraction
 * (:834-895, restated with the reference's 26-neighbour product), degenerate test (:911-920), clustering.  inside / use / invalid
 * [nx ny nz] are outputs; vertex_idx [24] out; iters [2] out (inflation, clustering).  Returns the cluster size or -1. */
int voxel_oracle_polytope(const uint8_t *occ, int nx, int ny, int nz, const int32_t seed[3], int itr_inflate_max, int itr_cluster_max,
                          int cap, int cand_cap, int32_t *cluster_xyz, int32_t *vertex_idx, int *iters, uint8_t *inside, uint8_t *use,
                          uint8_t *invalid);

/* The same two functions with the reference's stale read reproduced, for the comparison with its own host loop
 * (oracle/_ref/libvoxel_server_ref.so): host_can_can [cand_cap (cand_cap + 1) / 2] plays the reference's pinned h_can_can_result, which
 * lives as long as the generator object: every iteration copies the first C (C - 1) / 2 entries of the kernels' output into it
 * (cluster_server.cu:677-682) and the acceptance scan reads entry i (i + 1) / 2 + j of it (:700-707), so the row of the LAST candidate
 * is whatever an earlier iteration or call left there (zeros in a fresh buffer).  NULL = the functions above. */
int voxel_oracle_cluster_hostbuf(const uint8_t *occ, const uint8_t *inside, uint8_t *use, uint8_t *invalid, int nx, int ny, int nz,
                                 int32_t *cluster_xyz, int cluster_num, int cap, int cand_cap, int itr_cluster_max, int *iters_out,
                                 uint8_t *host_can_can);
int voxel_oracle_polytope_hostbuf(const uint8_t *occ, int nx, int ny, int nz, const int32_t seed[3], int itr_inflate_max,
                                  int itr_cluster_max, int cap, int cand_cap, int32_t *cluster_xyz, int32_t *vertex_idx, int *iters,
                                  uint8_t *inside, uint8_t *use, uint8_t *invalid, uint8_t *host_can_can);

#ifdef __cplusplus
}
#endif
#endif
