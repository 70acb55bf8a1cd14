// voxel_server_ref_driver.cu -- one C entry point around the REFERENCE's own host loop (test infrastructure).
//
// Compiled together with /root/reference/polyhedron_generator/src/{cluster_server.cu, cluster_engine.cu, cluster_engine_cpu.cpp}
// (unmodified, where they lie; `ros/ros.h` is oracle/shim's stopwatch stand-in, the only thing the file takes from ROS:
// cluster_server.cu:382-399, :560-762, :777-932) into oracle/_ref/libvoxel_server_ref.so by oracle/voxel_py.build(ref=True).
// The .so travels to the GPU box, the reference sources do not.
//
// What it runs is what the reference's node runs per polytope: cudaPolytopeGeneration::paramSet once, the occupancy map through
// setObs + mapUpload, then polygonGeneration (cluster_server.cu:769-966) from a one-voxel seed - box inflation on the host (the
// `#if _is_gpu_on_stage_1` at :822 is a preprocessor test of a non-macro, so cubeInflation_cpu always runs), clustering through
// polytopeCluster_gpu (:556-767): per iteration candidate upload, paraConvexTest + paraResultCheck, two downloads, the acceptance scan
// on the host, cluster upload.  Used to pin direct_voxel_polytope voxel for voxel and as its timing baseline (tools/voxel_report.py).
#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <vector>
#define private public   // this driver zero-fills the generator's pinned h_can_can_result (below); the class itself is untouched
#include "polyhedron_generator/cluster_server.cuh"
#undef private

namespace {
cudaPolytopeGeneration *g_gen = nullptr;
int g_dims[3] = {0, 0, 0};
}  // namespace

// (Re)creates the generator for a map of nx x ny x nz voxels and uploads `occ` (1 = occupied).  `resolution` only selects the
// reference's buffer sizes (cluster_server.cu:124-136: < 0.15 -> 50000 cluster voxels / 10000 candidates).  Returns 0 or a cudaError_t.
extern "C" int voxel_server_ref_setup(const uint8_t *occ, int nx, int ny, int nz, double resolution, int itr_inflate_max,
                                      int itr_cluster_max) {
    delete g_gen;
    g_gen = new cudaPolytopeGeneration();
    g_gen->paramSet(false, true, true, nx, ny, nz, resolution, (double)itr_inflate_max, (double)itr_cluster_max);
    g_dims[0] = nx; g_dims[1] = ny; g_dims[2] = nz;
    for (int x = 0; x < nx; x++)
        for (int y = 0; y < ny; y++)
            for (int z = 0; z < nz; z++)
                if (occ[((size_t)x * ny + y) * nz + z]) g_gen->setObs(x, y, z);
    g_gen->mapUpload();
    // polytopeCluster_gpu reads the last candidate row of every iteration from bytes of this buffer that its download never wrote
    // (:677-682 against :700-707); cudaMallocHost does not promise their content, so give every generator the same start: zeros.
    memset(g_gen->h_can_can_result, 0, sizeof(bool) * (size_t)g_gen->_cluster_buffer_size_square);
    return (int)cudaDeviceSynchronize();
}

// polygonGeneration from the one-voxel seed; out_xyz [cap][3] receives the cluster in the reference's order.  *seconds = wall time
// of the call alone (steady_clock around it, device idle before and after).  Returns the number of cluster voxels, -1 if `cap` is
// too small, -2 without setup, or -(1000 + cudaError_t).
extern "C" int voxel_server_ref_polytope(const int *seed, int *out_xyz, int cap, double *seconds) {
    if (!g_gen) return -2;
    std::vector<int> cx(1, seed[0]), cy(1, seed[1]), cz(1, seed[2]);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) return -(1000 + (int)e);
    const auto t0 = std::chrono::steady_clock::now();
    g_gen->polygonGeneration(cx, cy, cz);
    e = cudaDeviceSynchronize();
    const auto t1 = std::chrono::steady_clock::now();
    if (e != cudaSuccess) return -(1000 + (int)e);
    if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
    const int n = (int)cx.size();
    if (n > cap) return -1;
    for (int i = 0; i < n; i++) { out_xyz[3 * i] = cx[i]; out_xyz[3 * i + 1] = cy[i]; out_xyz[3 * i + 2] = cz[i]; }
    return n;
}

extern "C" void voxel_server_ref_release() {
    delete g_gen;
    g_gen = nullptr;
}
