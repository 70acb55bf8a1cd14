// voxel_ref_driver.cu -- C entry points around the REFERENCE's own CUDA kernels (test infrastructure).
//
// Compiled together with /root/reference/polyhedron_generator/src/cluster_engine.cu (unmodified, where it lies) into
// oracle/_ref/libvoxel_ref.so by `make -C oracle voxelref`; the .so travels to the GPU box, the reference sources do not.
// Launch shapes are the reference's: paraConvexTest / paraResultCheck as in cluster_server.cu:653-672, paraCubeInflation
// <<<128, 128>>> as in cluster_server.cu:170-171, :387.  Used to pin oracle/voxel_oracle.c and direct_b200's kernels
// (tests/test_voxel.py) and as the GPU baseline of tools/voxel_report.py.  Returns 0 or a cudaError_t.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cuda_runtime.h>
#include "polyhedron_generator/cluster_engine.cuh"

#define RCK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return (int)e_; } while (0)

// Device pointers in, device pointers out; *ms (optional) = kernel time of the two launches.
extern "C" int voxel_ref_convex_test_device(const uint8_t *d_map, const uint8_t *d_inside, int yz, int nz, const int *d_cand, int C,
                                            const int *d_clu, int K, bool *d_scratch /* [C (C + K)] */, bool *d_can_can, bool *d_can_clu,
                                            float *ms) {
    cudaEvent_t e0, e1;
    RCK(cudaEventCreate(&e0)); RCK(cudaEventCreate(&e1));
    RCK(cudaEventRecord(e0));
    int para_comp_num = C * (K + C);
    dim3 threads_cvx, blocks_cvx, threads_res_chk, blocks_res_chk;
    threads_cvx.x = std::min(1024, C);
    blocks_cvx.x = ceil(para_comp_num / threads_cvx.x) + 1;
    paraConvexTest<<<blocks_cvx, threads_cvx>>>(d_map, d_inside, d_cand, d_clu, d_scratch, yz, nz, C, K);
    RCK(cudaDeviceSynchronize());
    threads_res_chk.x = std::min(1024, C);
    blocks_res_chk.x = ceil(C / threads_res_chk.x) + 1;
    paraResultCheck<<<blocks_res_chk, threads_res_chk>>>(d_scratch, d_can_can, d_can_clu, C, K);
    RCK(cudaEventRecord(e1));
    RCK(cudaDeviceSynchronize());
    if (ms) RCK(cudaEventElapsedTime(ms, e0, e1));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return 0;
}

// Host buffers in and out.  can_can [C (C + 1) / 2] is uploaded first so that entries the kernels never write keep the caller's bytes.
extern "C" int voxel_ref_convex_test(const uint8_t *map, const uint8_t *inside, int nx, int ny, int nz, const int *cand, int C,
                                     const int *clu, int K, uint8_t *can_can, uint8_t *can_clu, float *ms) {
    if (C <= 0) return 0;
    static_assert(sizeof(bool) == 1, "bool");
    const size_t cells = (size_t)nx * ny * nz, ncc = (size_t)C * ((size_t)C + 1) / 2;
    uint8_t *d_map, *d_inside; int *d_cand, *d_clu; bool *d_scratch, *d_cc, *d_cl;
    RCK(cudaMalloc(&d_map, cells)); RCK(cudaMalloc(&d_inside, cells));
    RCK(cudaMalloc(&d_cand, sizeof(int) * 3 * (size_t)C)); RCK(cudaMalloc(&d_clu, sizeof(int) * 3 * (size_t)std::max(K, 1)));
    RCK(cudaMalloc(&d_scratch, (size_t)C * ((size_t)C + K))); RCK(cudaMalloc(&d_cc, ncc)); RCK(cudaMalloc(&d_cl, (size_t)C));
    RCK(cudaMemcpy(d_map, map, cells, cudaMemcpyHostToDevice)); RCK(cudaMemcpy(d_inside, inside, cells, cudaMemcpyHostToDevice));
    RCK(cudaMemcpy(d_cand, cand, sizeof(int) * 3 * (size_t)C, cudaMemcpyHostToDevice));
    if (K > 0) RCK(cudaMemcpy(d_clu, clu, sizeof(int) * 3 * (size_t)K, cudaMemcpyHostToDevice));
    RCK(cudaMemcpy(d_cc, can_can, ncc, cudaMemcpyHostToDevice));
    int st = voxel_ref_convex_test_device(d_map, d_inside, ny * nz, nz, d_cand, C, d_clu, K, d_scratch, d_cc, d_cl, ms);
    if (st) return st;
    RCK(cudaMemcpy(can_can, d_cc, ncc, cudaMemcpyDeviceToHost)); RCK(cudaMemcpy(can_clu, d_cl, (size_t)C, cudaMemcpyDeviceToHost));
    cudaFree(d_map); cudaFree(d_inside); cudaFree(d_cand); cudaFree(d_clu); cudaFree(d_scratch); cudaFree(d_cc); cudaFree(d_cl);
    return 0;
}

// One step of cubeInflation_gpu (cluster_server.cu:383-398): upload the 24 vertex indices, launch, synchronise, download the flag.
extern "C" int voxel_ref_cube_inflation(const uint8_t *map, int nx, int ny, int nz, const int *vertex_idx, int dir, int inf_step, int *result) {
    const size_t cells = (size_t)nx * ny * nz;
    uint8_t *d_map; int *d_v; bool *d_r; bool h = false;
    RCK(cudaMalloc(&d_map, cells)); RCK(cudaMalloc(&d_v, sizeof(int) * 24)); RCK(cudaMalloc(&d_r, sizeof(bool)));
    RCK(cudaMemcpy(d_map, map, cells, cudaMemcpyHostToDevice));
    RCK(cudaMemcpy(d_v, vertex_idx, sizeof(int) * 24, cudaMemcpyHostToDevice));
    paraCubeInflation<<<128, 128>>>(dir, inf_step, d_map, ny * nz, nz, d_v, d_r);
    RCK(cudaDeviceSynchronize());
    RCK(cudaMemcpy(&h, d_r, sizeof(bool), cudaMemcpyDeviceToHost));
    *result = h ? 1 : 0;
    cudaFree(d_map); cudaFree(d_v); cudaFree(d_r);
    return 0;
}

// The reference's stepwise loop (cluster_server.cu:343-440) with the map resident: per direction a 96-byte H2D, a launch, a
// device synchronise and a 1-byte D2H.  The baseline the fused loop of direct_voxel_inflate_box is timed against.
extern "C" int voxel_ref_inflate_box(const uint8_t *map, int nx, int ny, int nz, int *vertex_idx, int inf_step, int itr_inflate_max,
                                     int *iters, double *seconds) {
    const size_t cells = (size_t)nx * ny * nz;
    uint8_t *d_map; int *d_v; bool *d_r; bool h = false;
    RCK(cudaMalloc(&d_map, cells)); RCK(cudaMalloc(&d_v, sizeof(int) * 24)); RCK(cudaMalloc(&d_r, sizeof(bool)));
    RCK(cudaMemcpy(d_map, map, cells, cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1;
    RCK(cudaEventCreate(&e0)); RCK(cudaEventCreate(&e1));
    RCK(cudaEventRecord(e0));
    int last[24];
    std::copy(vertex_idx, vertex_idx + 24, last);
    int *v = vertex_idx, iter = 0;
    while (iter < itr_inflate_max) {
        for (int dir = 0; dir < 6; dir++) {
            bool at_max = false;
            switch (dir) {
                case 0: at_max = v[8] == 0; break;
                case 1: at_max = v[9] == ny - 1; break;
                case 2: at_max = v[3] == 0; break;
                case 3: at_max = v[0] == nx - 1; break;
                case 4: at_max = v[20] == 0; break;
                case 5: at_max = v[16] == nz - 1; break;
            }
            if (at_max) continue;
            RCK(cudaMemcpy(d_v, v, sizeof(int) * 24, cudaMemcpyHostToDevice));
            paraCubeInflation<<<128, 128>>>(dir, inf_step, d_map, ny * nz, nz, d_v, d_r);
            RCK(cudaDeviceSynchronize());
            RCK(cudaMemcpy(&h, d_r, sizeof(bool), cudaMemcpyDeviceToHost));
            if (!h) continue;
            static const int idx[6][4] = {{8, 11, 12, 15}, {9, 10, 13, 14}, {2, 3, 6, 7}, {0, 1, 4, 5}, {20, 21, 22, 23}, {16, 17, 18, 19}};
            for (int k = 0; k < 4; k++) v[idx[dir][k]] += (dir & 1) ? inf_step : -inf_step;
        }
        if (std::equal(last, last + 24, v)) break;
        std::copy(v, v + 24, last);
        iter++;
    }
    RCK(cudaEventRecord(e1));
    RCK(cudaDeviceSynchronize());
    float ms = 0;
    RCK(cudaEventElapsedTime(&ms, e0, e1));
    if (iters) *iters = iter;
    if (seconds) *seconds = ms * 1e-3;
    cudaFree(d_map); cudaFree(d_v); cudaFree(d_r);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return 0;
}
