/* voxel_oracle.c -- CPU oracle of the voxel-map kernels.
 *
 * TEST INFRASTRUCTURE ONLY: the checker of direct_b200/csrc/voxel.cuh in tests/ and tools/; nothing under direct_b200/
 * may include, link or call it.
 *
 * Restates, in plain sequential C, what the reference's three CUDA kernels compute
 * (polyhedron_generator/src/cluster_engine.cu) and the two host loops that drive them (cluster_server.cu).  PINNED on the
 * GPU box against the reference's own kernels: oracle/voxel_ref_driver.cu + the reference's cluster_engine.cu compile
 * (nvcc, sm_100a, sources where they lie) into oracle/_ref/libvoxel_ref.so, and tests/test_voxel.py compares oracle,
 * reference kernels and direct_b200's kernels byte for byte on the same maps; tests/golden/voxel_ref_*.npz are outputs of
 * the reference kernels brought back from a B200 (tests/golden/make_voxel_golden.py), checked by the CPU suite.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "voxel_oracle.h"

/* cluster_engine.cu:12-16 */
static double mod_ref(double value, double modulus) { return fmod(fmod(value, modulus) + modulus, modulus); }

/* cluster_engine.cu:19-35 */
static double intbound_ref(double s, int ds) {
    if (ds == 0) return 99999.0;
    if (ds < 0) return intbound_ref(-s, -ds);
    s = mod_ref(s, 1.0f);
    return (1 - s) / ds;
}

/* cluster_engine.cu:6-9 */
static double signum_ref(int x) { return x == 0 ? 0 : x < 0 ? -1 : 1; }

/* one thread of paraConvexTest past its target lookup, cluster_engine.cu:104-176; returns d_result[tid] */
static int ray_ref(const uint8_t *occ, const uint8_t *inside, int yz, int nz, int x, int y, int z, int endX, int endY, int endZ) {
    int result = 1;
    int dx = endX - x, dy = endY - y, dz = endZ - z;
    int stepX = (int)signum_ref(dx), stepY = (int)signum_ref(dy), stepZ = (int)signum_ref(dz);
    double tMaxX = intbound_ref(0.5, dx), tMaxY = intbound_ref(0.5, dy), tMaxZ = intbound_ref(0.5, dz);
    double tDeltaX = ((double)stepX) / dx, tDeltaY = ((double)stepY) / dy, tDeltaZ = ((double)stepZ) / dz;
    for (;;) {
        if (x == endX && y == endY && z == endZ) break;
        if (tMaxX < tMaxY) {
            if (tMaxX < tMaxZ) { x += stepX; tMaxX += tDeltaX; }
            else { z += stepZ; tMaxZ += tDeltaZ; }
        } else {
            if (tMaxY < tMaxZ) { y += stepY; tMaxY += tDeltaY; }
            else { z += stepZ; tMaxZ += tDeltaZ; }
        }
        int idx = x * yz + y * nz + z;
        if (inside[idx] > 0) return result;
        if (x == endX && y == endY && z == endZ) break;
        if (occ[idx] > 0) result = 0;
    }
    return result;
}

void voxel_oracle_convex_test(const uint8_t *occ, const uint8_t *inside, int ny, int nz, const int32_t *cand, int C,
                              const int32_t *clu, int K, uint8_t *can_can, uint8_t *can_clu) {
    const int yz = ny * nz;
    for (int t = 0; t < C; t++) {
        const int x = cand[3 * t], y = cand[3 * t + 1], z = cand[3 * t + 2];
        const long long bias = (long long)(t + 1) * t / 2;   /* cluster_engine.cu:46-47 */
        for (int i = 0; i < t; i++)
            can_can[bias + i] = (uint8_t)ray_ref(occ, inside, yz, nz, x, y, z, cand[3 * i], cand[3 * i + 1], cand[3 * i + 2]);
        int all = 1;
        for (int i = 0; i < K; i++)
            if (!ray_ref(occ, inside, yz, nz, x, y, z, clu[3 * i], clu[3 * i + 1], clu[3 * i + 2])) { all = 0; break; }
        can_clu[t] = (uint8_t)all;
    }
}

int voxel_oracle_cube_inflation(const uint8_t *occ, int ny, int nz, const int32_t *v, int dir, int inf_step, long long max_threads) {
    const int yz = ny * nz;
    int na, nb, a0, b0, c0, cs;
    switch (dir) {   /* (extent a, extent b, origin a, origin b, fixed coordinate, its direction), cluster_engine.cu:196-343 */
        case 0: na = v[0] - v[3] + 1;   nb = v[16] - v[20] + 1; a0 = v[3];  b0 = v[20]; c0 = v[8];  cs = -1; break;
        case 1: na = v[1] - v[2] + 1;   nb = v[17] - v[21] + 1; a0 = v[2];  b0 = v[21]; c0 = v[9];  cs = 1;  break;
        case 2: na = v[10] - v[11] + 1; nb = v[19] - v[23] + 1; a0 = v[11]; b0 = v[23]; c0 = v[3];  cs = -1; break;
        case 3: na = v[9] - v[8] + 1;   nb = v[16] - v[20] + 1; a0 = v[8];  b0 = v[20]; c0 = v[0];  cs = 1;  break;
        case 4: na = v[13] - v[12] + 1; nb = v[4] - v[7] + 1;   a0 = v[12]; b0 = v[7];  c0 = v[20]; cs = -1; break;
        case 5: na = v[9] - v[8] + 1;   nb = v[0] - v[3] + 1;   a0 = v[8];  b0 = v[3];  c0 = v[16]; cs = 1;  break;
        default: return 1;
    }
    long long n = (long long)na * nb;
    if (max_threads > 0 && n > max_threads) n = max_threads;
    int result = 1;
    for (long long tid = 0; tid < n; tid++) {
        int ia = (int)(tid / nb) + a0, ib = (int)(tid - (tid / nb) * nb) + b0;
        for (int i = 1; i <= inf_step; i++) {
            int c = c0 + cs * i, x, y, z;
            if (dir < 2) { x = ia; y = c; z = ib; }
            else if (dir < 4) { x = c; y = ia; z = ib; }
            else { x = ib; y = ia; z = c; }
            if (occ[x * yz + y * nz + z] > 0) result = 0;
        }
    }
    return result;
}

int voxel_oracle_inflate_box(const uint8_t *occ, int nx, int ny, int nz, int32_t *v, int inf_step, int itr_inflate_max) {
    int32_t last[24];
    memcpy(last, v, sizeof last);
    int iter = 0;
    while (iter < itr_inflate_max) {   /* cluster_server.cu:347-440 */
        for (int dir = 0; dir < 6; dir++) {
            int at_max = 0;
            switch (dir) {
                case 0: at_max = v[8] == 0; break;
                case 1: at_max = v[9] == ny - 1; break;
                case 2: at_max = v[3] == 0; break;
                case 3: at_max = v[0] == nx - 1; break;
                case 4: at_max = v[20] == 0; break;
                case 5: at_max = v[16] == nz - 1; break;
            }
            if (at_max) continue;
            if (!voxel_oracle_cube_inflation(occ, ny, nz, v, dir, inf_step, 0)) continue;
            switch (dir) {
                case 0: v[8] -= inf_step;  v[11] -= inf_step; v[12] -= inf_step; v[15] -= inf_step; break;
                case 1: v[9] += inf_step;  v[10] += inf_step; v[13] += inf_step; v[14] += inf_step; break;
                case 2: v[2] -= inf_step;  v[3] -= inf_step;  v[6] -= inf_step;  v[7] -= inf_step;  break;
                case 3: v[0] += inf_step;  v[1] += inf_step;  v[4] += inf_step;  v[5] += inf_step;  break;
                case 4: v[20] -= inf_step; v[21] -= inf_step; v[22] -= inf_step; v[23] -= inf_step; break;
                case 5: v[16] += inf_step; v[17] += inf_step; v[18] += inf_step; v[19] += inf_step; break;
            }
        }
        if (memcmp(last, v, sizeof last) == 0) break;
        memcpy(last, v, sizeof last);
        iter++;
    }
    return iter;
}

int voxel_oracle_cluster(const uint8_t *occ, const uint8_t *inside, uint8_t *use, uint8_t *invalid, int nx, int ny, int nz,
                         int32_t *cluster_xyz, int cluster_num, int cap, int cand_cap, int itr_cluster_max, int *iters_out) {
    return voxel_oracle_cluster_hostbuf(occ, inside, use, invalid, nx, ny, nz, cluster_xyz, cluster_num, cap, cand_cap, itr_cluster_max,
                                        iters_out, NULL);
}

int voxel_oracle_cluster_hostbuf(const uint8_t *occ, const uint8_t *inside, uint8_t *use, uint8_t *invalid, int nx, int ny, int nz,
                                 int32_t *cluster_xyz, int cluster_num, int cap, int cand_cap, int itr_cluster_max, int *iters_out,
                                 uint8_t *host_can_can) {
    const int yz = ny * nz;
    int32_t *active = (int32_t *)malloc(sizeof(int32_t) * 3 * (size_t)(cand_cap > cluster_num ? cand_cap : cluster_num));
    int32_t *cand = (int32_t *)malloc(sizeof(int32_t) * 3 * (size_t)cand_cap);
    uint8_t *can_can = (uint8_t *)malloc((size_t)cand_cap * ((size_t)cand_cap + 1) / 2 + 1);
    uint8_t *can_clu = (uint8_t *)malloc((size_t)cand_cap + 1), *accepted = (uint8_t *)malloc((size_t)cand_cap + 1);
    int active_num = cluster_num, itr = 0, status = 0;
    memcpy(active, cluster_xyz, sizeof(int32_t) * 3 * (size_t)cluster_num);
    while (itr < itr_cluster_max) {   /* cluster_server.cu:568-767 */
        int C = 0;
        for (int i = 0; i < active_num && !status; i++) {   /* :573-626 */
            int cx = active[3 * i], cy = active[3 * i + 1], cz = active[3 * i + 2];
            use[cx * yz + cy * nz + cz] = 1;
            for (int dx = -1; dx < 2; dx++) for (int dy = -1; dy < 2; dy++) for (int dz = -1; dz < 2; dz++) {
                if (dx == 0 && dy == 0 && dz == 0) continue;
                int x = cx + dx, y = cy + dy, z = cz + dz;
                if (x < 0 || x > nx - 1 || y < 0 || y > ny - 1 || z < 0 || z > nz - 1) continue;
                int idx = x * yz + y * nz + z;
                if (occ[idx] == 1 || use[idx] == 1 || invalid[idx] == 1 || inside[idx] == 1) continue;
                if (C >= cand_cap) { status = -1; continue; }
                cand[3 * C] = x; cand[3 * C + 1] = y; cand[3 * C + 2] = z;
                C++;
                use[idx] = 1;
            }
        }
        if (status) break;
        if (C == 0) break;   /* :652 */
        voxel_oracle_convex_test(occ, inside, ny, nz, cand, C, cluster_xyz, cluster_num, can_can, can_clu);
        const uint8_t *cc = can_can;
        if (host_can_can) {   /* the reference's download, :677-682: C (C - 1) / 2 entries, one candidate row short */
            memcpy(host_can_can, can_can, (size_t)C * ((size_t)C - 1) / 2);
            cc = host_can_can;
        }
        memset(accepted, 0, (size_t)C);
        active_num = 0;
        for (int i = 0; i < C; i++) {   /* :693-737 */
            int convex = 1;
            if (!can_clu[i]) convex = 0;
            else {
                long long bias = (long long)(i + 1) * i / 2;
                for (int j = 0; j < i; j++) if (!cc[bias + j] && accepted[j]) { convex = 0; break; }
            }
            int x = cand[3 * i], y = cand[3 * i + 1], z = cand[3 * i + 2];
            if (convex) {
                if (cluster_num >= cap) { status = -1; break; }
                accepted[i] = 1;
                cluster_xyz[3 * cluster_num] = x; cluster_xyz[3 * cluster_num + 1] = y; cluster_xyz[3 * cluster_num + 2] = z;
                active[3 * active_num] = x; active[3 * active_num + 1] = y; active[3 * active_num + 2] = z;
                cluster_num++; active_num++;
            } else invalid[x * yz + y * nz + z] = 1;
        }
        if (status) break;
        if (active_num == 0) break;   /* :739 */
        itr++;
    }
    free(active); free(cand); free(can_can); free(can_clu); free(accepted);
    if (iters_out) *iters_out = itr;
    return status ? status : cluster_num;
}

int voxel_oracle_polytope(const uint8_t *occ, int nx, int ny, int nz, const int32_t seed[3], int itr_inflate_max, int itr_cluster_max,
                          int cap, int cand_cap, int32_t *cluster_xyz, int32_t *v, int *iters, uint8_t *inside, uint8_t *use,
                          uint8_t *invalid) {
    return voxel_oracle_polytope_hostbuf(occ, nx, ny, nz, seed, itr_inflate_max, itr_cluster_max, cap, cand_cap, cluster_xyz, v, iters,
                                         inside, use, invalid, NULL);
}

int voxel_oracle_polytope_hostbuf(const uint8_t *occ, int nx, int ny, int nz, const int32_t seed[3], int itr_inflate_max,
                                  int itr_cluster_max, int cap, int cand_cap, int32_t *cluster_xyz, int32_t *v, int *iters,
                                  uint8_t *inside, uint8_t *use, uint8_t *invalid, uint8_t *host_can_can) {
    const int yz = ny * nz;
    const size_t cells = (size_t)nx * ny * nz;
    memset(use, 0, cells); memset(invalid, 0, cells); memset(inside, 0, cells);   /* flagClear, :39-45 */
    for (int k = 0; k < 8; k++) { v[k] = seed[0]; v[8 + k] = seed[1]; v[16 + k] = seed[2]; }   /* :793-800 */
    iters[0] = voxel_oracle_inflate_box(occ, nx, ny, nz, v, 1, itr_inflate_max);   /* cubeInflation_*: inf_step = 1 */
    iters[1] = 0;
    /* getVoxelsInCube, :79-98 */
    long long ncube = 0;
    for (int x = v[7]; x <= v[1]; x++) for (int y = v[8 + 7]; y <= v[8 + 1]; y++) for (int z = v[16 + 7]; z <= v[16 + 1]; z++) {
        inside[x * yz + y * nz + z] = 1;
        ncube++;
    }
    int n = 0;
    if (ncube == 1) {   /* :839-845 */
        if (cap < 1) return -1;
        cluster_xyz[0] = v[7]; cluster_xyz[1] = v[8 + 7]; cluster_xyz[2] = v[16 + 7];
        n = 1;
    } else {            /* :846-886 */
        for (int x = v[7]; x <= v[1]; x++) for (int y = v[8 + 7]; y <= v[8 + 1]; y++) for (int z = v[16 + 7]; z <= v[16 + 1]; z++) {
            use[x * yz + y * nz + z] = 1;
            int is_inside = 1;
            for (int dx = -1; dx < 2; dx++) for (int dy = -1; dy < 2; dy++) for (int dz = -1; dz < 2; dz++) {
                if (dx == 0 && dy == 0 && dz == 0) continue;
                int tx = x + dx, ty = y + dy, tz = z + dz;
                if (tx >= 0 && tx < nx && ty >= 0 && ty < ny && tz >= 0 && tz < nz) is_inside *= inside[tx * yz + ty * nz + tz];
                else is_inside = 0;
            }
            if (is_inside < 1) {
                if (n >= cap) return -1;
                cluster_xyz[3 * n] = x; cluster_xyz[3 * n + 1] = y; cluster_xyz[3 * n + 2] = z;
                n++;
            }
        }
    }
    for (int i = 0; i < n; i++) inside[cluster_xyz[3 * i] * yz + cluster_xyz[3 * i + 1] * nz + cluster_xyz[3 * i + 2]] = 0;   /* :889-892 */
    if (abs(v[7] - v[1]) == 0 || abs(v[8 + 7] - v[8 + 1]) == 0 || abs(v[16 + 7] - v[16 + 1]) == 0) return n;   /* :911-920 */
    return voxel_oracle_cluster_hostbuf(occ, inside, use, invalid, nx, ny, nz, cluster_xyz, n, cap, cand_cap, itr_cluster_max, &iters[1],
                                        host_can_can);
}
