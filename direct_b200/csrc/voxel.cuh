// voxel.cuh -- voxel-map kernels for sm_100a (include/direct_voxel.h).
//
// Integer / byte work bound by the latency of uint8 gathers into a map that lives in L2 (a 200 x 200 x 40 map is 1.6 MB):
// nothing here is GEMM-shaped.  convex_test: one WARP per candidate voxel, its lanes stride over the targets (earlier
// candidates, then the cluster), every lane walks its own DDA ray; candidate-candidate results go straight into the
// reference's packed triangular array, the candidate-cluster results are AND-reduced with a warp vote - the reference's
// cand x (cand + clu) intermediate array and its second kernel (paraResultCheck) do not exist.  The DDA repeats the
// reference's fp64 operations one for one (cluster_engine.cu:6-35, :110-176), so every comparison of two ray parameters
// gives the same answer and the results are bit-identical.
#ifndef DIRECT_B200_VOXEL_CUH_
#define DIRECT_B200_VOXEL_CUH_

#include <cooperative_groups.h>

namespace voxel {
namespace cg = cooperative_groups;

// 1 / k for k = 1 .. RCP_N - 1, correctly rounded fp64 quotients (IEEE division gives the same bits on the host and on the device);
// filled once per handle.  The reference's per-ray quotients are tDelta = step / d = 1 / |d| and tMax = (1 - 0.5) / |d| (cluster_engine.cu
// :19-35 with s = mod(+-0.5, 1) = 0.5), and 0.5 / |d| is exactly half of 1 / |d|: three table reads replace six fp64 divisions, same bits.
constexpr int RCP_N = 4096;
__global__ void rcp_table_kernel(double *rcp) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < RCP_N) rcp[k] = k == 0 ? 0.0 : 1.0 / (double)k;
}
__device__ __forceinline__ double rcp_abs(const double *__restrict__ rcp, int d) {
    const int a = d < 0 ? -d : d;
    return a < RCP_N ? __ldg(rcp + a) : 1.0 / (double)a;
}

// Merged map: bit 0 = occupied (map_data > 0), bit 1 = inside (inside_data > 0).  One byte gather per ray step instead of two.
__global__ void merge_map_kernel(const uint8_t *__restrict__ occ, const uint8_t *__restrict__ inside, uint8_t *__restrict__ merged, size_t cells,
                                 uint8_t *ones, int n_ones) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < cells; i += (size_t)gridDim.x * blockDim.x) {
        merged[i] = (uint8_t)((occ[i] > 0 ? 1 : 0) | (inside[i] > 0 ? 2 : 0));
        if (i < (size_t)n_ones) ones[i] = 1;   // can_clu of the convex test starts true
    }
}

// true = the ray from (x, y, z) to (ex, ey, ez) is free (paraConvexTest's d_result[tid], cluster_engine.cu:104-176).
// The walk takes the reference's decisions in the reference's order (same fp64 tMax additions and comparisons).  Two things are
// arranged differently.  (1) The walk reaches its end voxel after exactly |dx| + |dy| + |dz| steps (axis a crosses its last
// boundary at t = 1 - 0.5 / |d_a| < 1 and its next one only at t > 1; the margin 0.5 / |d| is 13 orders of magnitude above the
// rounding of the additions), so "position == end" is a step counter.  (2) The reference's loop waits for each voxel before it
// takes the next step (the exit tests depend on the load); here the next LOOK steps are taken first, their voxels are fetched
// together and then examined in order - steps past the end are not taken, so nothing outside the ray is ever read.
constexpr int LOOK = 4;
__device__ __forceinline__ bool ray_free(const uint8_t *__restrict__ m, const double *__restrict__ rcp, int yz, int nz, int x, int y, int z,
                                         int ex, int ey, int ez) {
    const int dx = ex - x, dy = ey - y, dz = ez - z;
    const int sx = dx == 0 ? 0 : (dx < 0 ? -1 : 1), sy = dy == 0 ? 0 : (dy < 0 ? -1 : 1), sz = dz == 0 ? 0 : (dz < 0 ? -1 : 1);
    // an axis with d = 0 is never stepped (its tMax is the reference's 99999); its delta is never added
    const double ddx = rcp_abs(rcp, dx), ddy = rcp_abs(rcp, dy), ddz = rcp_abs(rcp, dz);
    double tx = dx == 0 ? 99999.0 : 0.5 * ddx, ty = dy == 0 ? 99999.0 : 0.5 * ddy, tz = dz == 0 ? 99999.0 : 0.5 * ddz;
    int rem = (dx < 0 ? -dx : dx) + (dy < 0 ? -dy : dy) + (dz < 0 ? -dz : dz);
    int idx = x * yz + y * nz + z;
    const int ix = sx * yz, iy = sy * nz, iz = sz;
    bool free_ray = true;
    while (rem > 0) {
        uint8_t v[LOOK];
        int left[LOOK];
#pragma unroll
        for (int u = 0; u < LOOK; u++) {
            if (rem > 0) {
                if (tx < ty) {
                    if (tx < tz) { idx += ix; tx += ddx; }
                    else { idx += iz; tz += ddz; }
                } else {
                    if (ty < tz) { idx += iy; ty += ddy; }
                    else { idx += iz; tz += ddz; }
                }
                rem--;
            }
            left[u] = rem;
            v[u] = m[idx];
        }
#pragma unroll
        for (int u = 0; u < LOOK; u++) {
            if (v[u] & 2) return free_ray;        // an inside voxel of the cluster: stop tracing
            if (left[u] == 0) return free_ray;    // the target itself is not examined
            if (v[u] & 1) free_ray = false;
        }
    }
    return free_ray;
}

// A task = one candidate x 32 consecutive targets (earlier candidates first, then the cluster voxels), chunk-major; can_clu is preset
// to 1 (merge_map_kernel) and cleared by any task that finds a blocked cluster ray; cluster tasks of a candidate already cleared are
// skipped (their rays decide nothing).
__global__ void __launch_bounds__(256) convex_test_kernel(const uint8_t *__restrict__ merged, const double *__restrict__ rcp, int yz, int nz,
                                                          const int *__restrict__ cand, int C, const int *__restrict__ clu, int K,
                                                          uint8_t *__restrict__ can_can, uint8_t *can_clu) {
    const int lane = threadIdx.x & 31;
    const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
    const int chunks = (C - 1 + K + 31) >> 5;
    // task = chunk * C + r, walked with stride `warps`: (chunk, r) advance by (warps / C, warps % C) with a carry - no division per task
    const long long first = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int dq = (int)(warps / C), dr = (int)(warps % C);
    int chunk = (int)(first / C), r = (int)(first % C);
    for (; chunk < chunks; chunk += dq, r += dr) {
        if (r >= C) { r -= C; chunk++; if (chunk >= chunks) break; }
        const int t = C - 1 - r, jb = chunk * 32, j = jb + lane;
        if (jb >= t + K) continue;                                        // warp-uniform
        if (jb >= t && !((volatile uint8_t *)can_clu)[t]) continue;       // warp-uniform
        bool clu_ok = true;
        if (j < t + K) {
            const int *e = j < t ? cand + 3 * j : clu + 3 * (j - t);
            const bool ok = ray_free(merged, rcp, yz, nz, cand[3 * t], cand[3 * t + 1], cand[3 * t + 2], e[0], e[1], e[2]);
            if (j < t) can_can[(long long)t * (t + 1) / 2 + j] = ok ? 1 : 0;   // the reference's packing: n (n - 1) / 2 with n = t + 1
            else clu_ok = ok;
        }
        if (!__all_sync(0xffffffffu, clu_ok) && lane == 0) can_clu[t] = 0;
    }
}

// Face `dir` of the box with vertex indices v (cluster_engine.cu:196-343): rectangle [a0, a0 + na) x [b0, b0 + nb), the fixed
// coordinate starts at c0 and moves by cs.
struct Face { int na, nb, a0, b0, c0, cs; };
__device__ __forceinline__ Face face_of(const int *v, int dir) {
    Face f;
    switch (dir) {
        case 0: f.na = v[0] - v[3] + 1; f.nb = v[16] - v[20] + 1; f.a0 = v[3]; f.b0 = v[20]; f.c0 = v[8]; f.cs = -1; break;       // Y-: (x, z), y = v[8] - i
        case 1: f.na = v[1] - v[2] + 1; f.nb = v[17] - v[21] + 1; f.a0 = v[2]; f.b0 = v[21]; f.c0 = v[9]; f.cs = 1; break;        // Y+
        case 2: f.na = v[10] - v[11] + 1; f.nb = v[19] - v[23] + 1; f.a0 = v[11]; f.b0 = v[23]; f.c0 = v[3]; f.cs = -1; break;    // X-: (y, z), x = v[3] - i
        case 3: f.na = v[9] - v[8] + 1; f.nb = v[16] - v[20] + 1; f.a0 = v[8]; f.b0 = v[20]; f.c0 = v[0]; f.cs = 1; break;        // X+
        case 4: f.na = v[13] - v[12] + 1; f.nb = v[4] - v[7] + 1; f.a0 = v[12]; f.b0 = v[7]; f.c0 = v[20]; f.cs = -1; break;      // Z-: (y, x), z = v[20] - i
        default: f.na = v[9] - v[8] + 1; f.nb = v[0] - v[3] + 1; f.a0 = v[8]; f.b0 = v[3]; f.c0 = v[16]; f.cs = 1; break;         // Z+
    }
    return f;
}

// Does any of the cells [first, first + stride, ...) of the face hit an occupied voxel within inf_step layers?
__device__ __forceinline__ bool face_hit(const uint8_t *__restrict__ occ, int yz, int nz, const Face &f, int dir, int inf_step,
                                         long long first, long long stride) {
    const long long n = (long long)f.na * f.nb;
    bool hit = false;
    for (long long tid = first; tid < n; tid += stride) {
        const int ia = (int)(tid / f.nb) + f.a0, ib = (int)(tid % f.nb) + f.b0;
        for (int i = 1; i <= inf_step; i++) {
            const int c = f.c0 + f.cs * i;
            int x, y, z;
            if (dir < 2) { x = ia; y = c; z = ib; }
            else if (dir < 4) { x = c; y = ia; z = ib; }
            else { x = ib; y = ia; z = c; }
            if (occ[x * yz + y * nz + z] > 0) hit = true;
        }
    }
    return hit;
}

// One thread per cell of the face; *result (preset to 1) becomes 0 when any of the inf_step layers is occupied.
__global__ void cube_inflation_kernel(const uint8_t *__restrict__ occ, int yz, int nz, const int *__restrict__ v, int dir, int inf_step,
                                      int *result) {
    if (dir < 0 || dir > 5) return;
    const Face f = face_of(v, dir);
    const bool hit = face_hit(occ, yz, nz, f, dir, inf_step, (long long)blockIdx.x * blockDim.x + threadIdx.x, (long long)gridDim.x * blockDim.x);
    if (__syncthreads_or(hit) && threadIdx.x == 0) *result = 0;
}

// cubeInflation_gpu (cluster_server.cu:343-440) as ONE launch of one CTA: the six-direction loop runs on the device with the
// vertex indices in shared memory, so the reference's per-step 96-byte upload, launch, device synchronise and 1-byte download
// (up to 6 x itr_inflate_max of each) do not exist.  A step is a scan of one face out of L2 plus two CTA barriers.
__global__ void __launch_bounds__(1024) inflate_box_kernel(const uint8_t *__restrict__ occ, int nx, int ny, int nz, int *v_io, int inf_step,
                                                           int itr_inflate_max, int *iters_out) {
    __shared__ int v[24], last[24];
    __shared__ int changed;
    const int yz = ny * nz;
    if (threadIdx.x < 24) { v[threadIdx.x] = v_io[threadIdx.x]; last[threadIdx.x] = v[threadIdx.x]; }
    __syncthreads();
    int iter = 0;
    while (iter < itr_inflate_max) {
        for (int dir = 0; dir < 6; dir++) {
            bool at_max;
            switch (dir) {
                case 0: at_max = v[8] == 0; break;
                case 1: at_max = v[9] == ny - 1; break;
                case 2: at_max = v[3] == 0; break;
                case 3: at_max = v[0] == nx - 1; break;
                case 4: at_max = v[20] == 0; break;
                default: at_max = v[16] == nz - 1; break;
            }
            if (at_max) continue;   // uniform: v is shared
            const Face f = face_of(v, dir);
            const bool hit = face_hit(occ, yz, nz, f, dir, inf_step, threadIdx.x, blockDim.x);
            const int any = __syncthreads_or(hit);   // every thread has read v before anyone moves it
            if (!any && threadIdx.x < 4) {
                const int base = dir == 0 ? 8 : dir == 1 ? 9 : dir == 2 ? 2 : dir == 3 ? 0 : dir == 4 ? 20 : 16;
                // moved vertices: Y- {8,11,12,15}, Y+ {9,10,13,14}, X- {2,3,6,7}, X+ {0,1,4,5}, Z- {20..23}, Z+ {16..19}
                const unsigned packed = dir == 0 ? 0x7430u : dir < 4 ? 0x5410u : 0x3210u;   // four vertex offsets, one per nibble
                const int idx = base + (int)((packed >> (4 * threadIdx.x)) & 15u);
                v[idx] += (dir & 1) ? inf_step : -inf_step;
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) changed = 0;
        __syncthreads();
        if (threadIdx.x < 24 && v[threadIdx.x] != last[threadIdx.x]) changed = 1;
        __syncthreads();
        if (!changed) break;
        if (threadIdx.x < 24) last[threadIdx.x] = v[threadIdx.x];
        __syncthreads();
        iter++;
    }
    if (threadIdx.x < 24) v_io[threadIdx.x] = v[threadIdx.x];
    if (threadIdx.x == 0 && iters_out) *iters_out = iter;
}

// ---- polytopeCluster_gpu (cluster_server.cu:556-767) as ONE cooperative launch ------------------------------------------------
// The reference runs, per clustering iteration: candidate generation on the host, an upload, two kernels with a device
// synchronise between them, two downloads (the can_can triangle: up to 50 MB), a sequential acceptance scan on the host and another
// upload.  Here the whole loop stays on the device; phases are separated by grid barriers:
//   1a  every (active voxel, neighbour) pair claims its free neighbour cell with atomicMin(key = 26 i + k): the smallest key is
//       the pair that reaches the cell first in the reference's nested loops (cluster_server.cu:573-626);
//   1b  the winning pairs are compacted in key order (per-segment counts, barrier, bases) -> the reference's candidate list,
//       element for element;
//   2a  one warp per (candidate, 32 cluster voxels): rays to the cluster voxels, candidates already found blocked are skipped (can_clu);
//   2b  one warp per (SURVIVING candidate, 32 earlier surviving candidates); the 32 results of a warp step are one ballot
//       word of that candidate's conflict bit-row (a candidate that failed 2a is rejected whatever its rays say and is never in
//       the accepted set, so its rays decide nothing: cluster_server.cu:696-711);
//   3   CTA 0 walks the candidates in order, 32 at a time: accepted <=> can_clu and (conflict row & accepted bit-set) == 0; accepted
//       voxels are appended to the cluster (they are the next iteration's active set), the others marked invalid.
// The reference reads the last candidate's can_can row from stale host memory (it downloads C (C - 1) / 2 entries and indexes up
// to C (C + 1) / 2 - 2); here every row is the kernel's own result, as in oracle/voxel_oracle.c.
struct ClusterCtl {
    int cluster_num;    // in / out
    int iters;          // out: iterations completed (the reference's itr_cluster_cnt)
    int status;         // out: 0, or -1 when cap / cand_cap would be exceeded
    int active_begin;   // scratch: the active voxels are cluster_xyz[active_begin .. cluster_num)
    int cand_num;       // scratch
    int accepted;       // scratch
    int skip;           // in: non-zero = do nothing (the degenerate box of polygonGeneration, cluster_server.cu:911-920)
    int pad;
};

constexpr int CLAIM_EMPTY = 0x7f7f7f7f;   // cudaMemset(0x7f)
constexpr int ACC_WORDS = 1024;           // accepted bit-set in shared memory: cand_cap <= 32768

// neighbour k = 0..25 in the reference's dx, dy, dz nesting order (the centre is skipped)
__device__ __forceinline__ bool neighbour_cell(const int *__restrict__ xyz, int i, int k, int nx, int ny, int nz, int &x, int &y, int &z) {
    const int kk = k < 13 ? k : k + 1;
    x = xyz[3 * i] + kk / 9 - 1; y = xyz[3 * i + 1] + (kk / 3) % 3 - 1; z = xyz[3 * i + 2] + kk % 3 - 1;
    return !(x < 0 || x > nx - 1 || y < 0 || y > ny - 1 || z < 0 || z > nz - 1);
}

__global__ void __launch_bounds__(256) cluster_loop_kernel(const uint8_t *__restrict__ occ, const uint8_t *__restrict__ inside, const uint8_t *__restrict__ merged, const double *__restrict__ rcp, uint8_t *use,
                                                           uint8_t *invalid, int *claim, int nx, int ny, int nz, int *cluster_xyz, int cap,
                                                           int *cand, int cand_cap, unsigned *conflict, uint8_t *can_clu, int *seg_count, int itr_cluster_max,
                                                           ClusterCtl *ctl, unsigned long long *phase_ns) {
    cg::grid_group grid = cg::this_grid();
    __shared__ unsigned acc[ACC_WORDS];
    const int yz = ny * nz, lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x, gthreads = (long long)gridDim.x * blockDim.x;
    const int gwarp = (int)(gtid >> 5), gwarps = (int)(gthreads >> 5);

    if (ctl->skip) return;   // uniform: nobody has reached a grid barrier
    if (gtid == 0) { ctl->active_begin = 0; ctl->iters = 0; ctl->status = 0; }
    for (long long i = gtid; i < ctl->cluster_num; i += gthreads)
        use[cluster_xyz[3 * i] * yz + cluster_xyz[3 * i + 1] * nz + cluster_xyz[3 * i + 2]] = 1;   // cluster_server.cu:583
    grid.sync();
    // per-phase device time (globaltimer, thread 0): the loop is one launch, so a profiler cannot split it
    unsigned long long t_prev = 0;
    auto mark = [&](int k) {
        if (gtid == 0 && phase_ns) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            if (k >= 0) phase_ns[k] += t - t_prev;
            t_prev = t;
        }
    };
    mark(-1);

    for (int itr = 0; itr < itr_cluster_max;) {
        const int K = ctl->cluster_num, a0 = ctl->active_begin, A = K - a0;
        const int *act = cluster_xyz + 3 * a0;
        // 1a: claims
        for (long long p = gtid; p < 26LL * A; p += gthreads) {
            int x, y, z;
            if (!neighbour_cell(act, (int)(p / 26), (int)(p % 26), nx, ny, nz, x, y, z)) continue;
            const int idx = x * yz + y * nz + z;
            if (occ[idx] == 1 || use[idx] == 1 || invalid[idx] == 1 || inside[idx] == 1) continue;
            atomicMin(&claim[idx], (int)p);
        }
        grid.sync();
        mark(0);
        // 1b: ordered compaction, grid-wide.  A warp owns segments of 128 consecutive keys (four per lane): count the winning keys of
        // each segment, barrier, then every segment finds its base (sum of the counts before it) and writes its winners in key order.
        const long long P = 26LL * A;
        const int nseg = (int)((P + 127) >> 7);
        for (int sg = gwarp; sg < nseg; sg += gwarps) {
            int mine = 0;
            for (int q = 0; q < 4; q++) {
                const long long p = ((long long)sg << 7) + 4 * lane + q;
                int x, y, z;
                if (p < P && neighbour_cell(act, (int)(p / 26), (int)(p % 26), nx, ny, nz, x, y, z)) mine += claim[x * yz + y * nz + z] == (int)p;
            }
            for (int d = 16; d > 0; d >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, d);
            if (lane == 0) seg_count[sg] = mine;
        }
        if (blockIdx.x == 0) for (int w = threadIdx.x; w < ACC_WORDS; w += blockDim.x) acc[w] = 0;
        grid.sync();
        for (int sg = gwarp; sg < nseg; sg += gwarps) {
            int base = 0;
            for (int q = lane; q < sg; q += 32) base += seg_count[q];
            for (int d = 16; d > 0; d >>= 1) base += __shfl_xor_sync(0xffffffffu, base, d);
            int cell[4], mine = 0;
            for (int q = 0; q < 4; q++) {
                const long long p = ((long long)sg << 7) + 4 * lane + q;
                int x, y, z;
                cell[q] = -1;
                if (p < P && neighbour_cell(act, (int)(p / 26), (int)(p % 26), nx, ny, nz, x, y, z) && claim[x * yz + y * nz + z] == (int)p) {
                    cell[q] = x * yz + y * nz + z; mine++;
                }
            }
            int incl = mine;
            for (int d = 1; d < 32; d <<= 1) { const int o = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += o; }
            int at = base + incl - mine;
            for (int q = 0; q < 4; q++) {
                if (cell[q] < 0) continue;
                claim[cell[q]] = CLAIM_EMPTY;
                if (at < cand_cap) {
                    const int x = cell[q] / yz, r = cell[q] - x * yz;
                    cand[3 * at] = x; cand[3 * at + 1] = r / nz; cand[3 * at + 2] = r % nz;
                    use[cell[q]] = 1; can_clu[at] = 1;
                }
                at++;
            }
            if (sg == nseg - 1 && lane == 31) {   // incl of the last lane of the last segment closes the count
                const int total = base + incl;
                ctl->cand_num = total <= cand_cap ? total : 0;
                if (total > cand_cap) ctl->status = -1;
            }
        }
        if (nseg == 0 && gtid == 0) ctl->cand_num = 0;
        grid.sync();
        mark(1);
        const int C = ctl->cand_num;
        if (C == 0) break;   // cluster_server.cu:652 (or overflow)
        const int W = (C + 31) >> 5;
        // 2a: candidate -> cluster rays.  A task = one candidate x 32 cluster voxels, chunk-major so that a candidate found blocked in
        // an early chunk is skipped by the later ones (the stores of 0 and the skip test race benignly: a late skip only costs rays).
        {
            const int chunks = (K + 31) >> 5;
            const int dq = gwarps / C, dr = gwarps % C;
            int chunk = gwarp / C, r = gwarp % C;
            for (; chunk < chunks; chunk += dq, r += dr) {   // task = chunk * C + r, stride gwarps, no division per task
                if (r >= C) { r -= C; chunk++; if (chunk >= chunks) break; }
                const int t = r, j = chunk * 32 + lane;
                if (!((volatile uint8_t *)can_clu)[t]) continue;   // warp-uniform: one address
                bool ok = true;
                if (j < K) ok = ray_free(merged, rcp, yz, nz, cand[3 * t], cand[3 * t + 1], cand[3 * t + 2], cluster_xyz[3 * j], cluster_xyz[3 * j + 1], cluster_xyz[3 * j + 2]);
                if (!__all_sync(0xffffffffu, ok) && lane == 0) can_clu[t] = 0;
            }
        }
        grid.sync();
        mark(2);
        // 2b: surviving candidate -> earlier surviving candidates.  A task = one candidate x 32 earlier candidates = one conflict word.
        {
            const int chunks = (C + 31) >> 5;
            const int dq = gwarps / C, dr = gwarps % C;
            int chunk = gwarp / C, r = gwarp % C;
            for (; chunk < chunks; chunk += dq, r += dr) {
                if (r >= C) { r -= C; chunk++; if (chunk >= chunks) break; }
                const int t = C - 1 - r, jb = chunk * 32, j = jb + lane;
                if (jb >= t || !can_clu[t]) continue;   // warp-uniform
                bool blocked = false;
                if (j < t && can_clu[j]) blocked = !ray_free(merged, rcp, yz, nz, cand[3 * t], cand[3 * t + 1], cand[3 * t + 2], cand[3 * j], cand[3 * j + 1], cand[3 * j + 2]);
                const unsigned word = __ballot_sync(0xffffffffu, blocked);
                if (lane == 0) conflict[(size_t)t * W + (jb >> 5)] = word;
            }
        }
        grid.sync();
        mark(3);
        // 3: acceptance scan (cluster_server.cu:693-737) by CTA 0, 32 candidates (one group) at a time.  Warp 0 resolves group g: the
        // order dependence inside the group from the group's diagonal conflict word, in registers (a few ballot rounds, no memory on
        // the dependent chain).  Meanwhile the other warps reduce, for group g + 1, the conflicts with everything accepted in groups < g
        // (rows read in parallel); the one word they cannot know yet - group g itself - is a single AND for warp 0 in the next round.
        if (blockIdx.x == 0) {
            __shared__ unsigned prehit[2];
            int n = K, overflow = 0;   // maintained by warp 0 (uniform across its lanes)
            if (threadIdx.x < 2) prehit[threadIdx.x] = 0;
            unsigned pf_cc = 0, pf_prev = 0, pf_diag = 0;
            int pf_x = 0, pf_y = 0, pf_z = 0;
            if (wib == 0 && lane < C) {   // group 0
                pf_cc = can_clu[lane]; pf_diag = lane > 0 ? conflict[(size_t)lane * W] : 0u;
                pf_x = cand[3 * lane]; pf_y = cand[3 * lane + 1]; pf_z = cand[3 * lane + 2];
            }
            __syncthreads();
            const int groups = (C + 31) >> 5;
            for (int g = 0; g < groups; g++) {
                if (wib == 0) {
                    const int i = g * 32 + lane;
                    const bool valid = i < C;
                    // this group's words were fetched during the previous round (pf_*); fetch the next group's now, off the chain
                    const unsigned cc = pf_cc, prev = pf_prev, dg = pf_diag;
                    const int x = pf_x, y = pf_y, z = pf_z;
                    {
                        const int i2 = i + 32;
                        if (i2 < C) {
                            pf_cc = can_clu[i2]; pf_prev = conflict[(size_t)i2 * W + g]; pf_diag = lane > 0 ? conflict[(size_t)i2 * W + g + 1] : 0u;
                            pf_x = cand[3 * i2]; pf_y = cand[3 * i2 + 1]; pf_z = cand[3 * i2 + 2];
                        } else pf_cc = 0;
                    }
                    bool alive = valid && cc && !((prehit[g & 1] >> lane) & 1u);
                    __syncwarp();
                    if (lane == 0) prehit[g & 1] = 0;   // free for group g + 2 (written after the next barrier)
                    if (alive && g > 0) alive = (prev & acc[g - 1]) == 0;
                    const unsigned diag = alive ? dg : 0u;   // bits j - 32 g < lane
                    // greedy in lane order without a 32-step chain: per round a lane is rejected when it conflicts with an accepted
                    // lane, accepted when it conflicts with no accepted and no still-undecided lower lane; the lowest undecided lane
                    // always decides, so the loop ends, usually after two or three rounds
                    unsigned accmask = 0, und = __ballot_sync(0xffffffffu, alive);
                    const unsigned lower = (1u << lane) - 1u;
                    while (und) {
                        const bool me = (und >> lane) & 1u;
                        const bool rej = me && (diag & accmask) != 0;
                        const bool acc_now = me && !rej && (diag & und & lower) == 0;
                        const unsigned a = __ballot_sync(0xffffffffu, acc_now), r = __ballot_sync(0xffffffffu, rej);
                        accmask |= a; und &= ~(a | r);
                    }
                    const bool mine = (accmask >> lane) & 1u;
                    const int at = n + __popc(accmask & ((1u << lane) - 1u));
                    const bool fits = at < cap;
                    if (mine && fits) { cluster_xyz[3 * at] = x; cluster_xyz[3 * at + 1] = y; cluster_xyz[3 * at + 2] = z; }
                    else if (valid) invalid[x * yz + y * nz + z] = 1;
                    const unsigned kept = __ballot_sync(0xffffffffu, mine && fits);
                    if (lane == 0) acc[g] = kept;
                    n += __popc(kept);
                    if (kept != accmask) overflow = 1;
                } else if (g >= 1 && g + 1 < groups) {
                    // 32 rows x g words of group g + 1, flat over the helper threads, eight independent reads in flight per thread
                    // (rows of candidates that failed 2a hold stale words: warp 0 ignores their bits)
                    const int items = 32 * g, th = (int)threadIdx.x - 32, nh = (int)blockDim.x - 32;
                    for (int base = th; base < items; base += nh * 8) {
                        unsigned v[8];
#pragma unroll
                        for (int u = 0; u < 8; u++) {
                            const int idx = base + u * nh;
                            v[u] = 0;
                            if (idx < items) {
                                const int q = idx / g, w = idx - q * g, i = (g + 1) * 32 + q;
                                if (i < C) v[u] = conflict[(size_t)i * W + w] & acc[w];
                            }
                        }
#pragma unroll
                        for (int u = 0; u < 8; u++)
                            if (v[u]) atomicOr(&prehit[(g + 1) & 1], 1u << ((base + u * nh) / g));
                    }
                }
                __syncthreads();
            }
            if (threadIdx.x == 0) {
                ctl->accepted = n - K; ctl->active_begin = K; ctl->cluster_num = n;
                if (overflow) ctl->status = -1;
            }
        }
        grid.sync();
        mark(4);
        if (ctl->status != 0 || ctl->accepted == 0) break;   // cluster_server.cu:739
        itr++;
        if (gtid == 0) ctl->iters = itr;
    }
}

// ---- polygonGeneration between its two stages (cluster_server.cu:822-920), one launch ----------------------------------------------
// The voxels of the inflated box become `inside` and `use`; those with a neighbour outside the box are the initial cluster and lose
// their `inside` flag.  After flagClear the `inside` flags are exactly the box, so "every one of the 26 neighbours is inside" is "the
// voxel is not on the box's boundary", and the position of a boundary voxel in the reference's x, y, z scan is a closed form: no
// compaction pass.  ctl->cluster_num = number of boundary voxels (a one-voxel box gets no use flag, :839-845), ctl->skip = the degenerate test of :911 (a box one voxel thick: the boundary is the result).  Flags must be
// zero on entry (flagClear, :39-45).
__device__ __forceinline__ long long shell_before_slab(long long a, long long ny, long long nz, long long nxb) {
    // boundary voxels in x-slabs 0 .. a-1 of an nxb x ny x nz box
    const long long full = ny * nz, ring = full - (ny > 2 ? ny - 2 : 0) * (nz > 2 ? nz - 2 : 0);
    long long n = 0;
    if (a > 0) n += full;                                              // slab 0
    if (a > 1) n += (a - 1 < nxb - 1 ? a - 1 : nxb - 2) * ring;        // slabs 1 .. min(a, nxb - 1) - 1
    if (a > nxb - 1 && nxb > 1) n += full;                             // slab nxb - 1
    return n;
}
__global__ void cube_shell_kernel(const int *__restrict__ v, int ny_map, int nz_map, uint8_t *inside, uint8_t *use, int *cluster_xyz, int cap,
                                  ClusterCtl *ctl) {
    const int x0 = v[7], x1 = v[1], y0 = v[8 + 7], y1 = v[8 + 1], z0 = v[16 + 7], z1 = v[16 + 1];
    const long long bx = x1 - x0 + 1, by = y1 - y0 + 1, bz = z1 - z0 + 1, cells = bx * by * bz;
    const int yz = ny_map * nz_map;
    for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < cells; c += (long long)gridDim.x * blockDim.x) {
        const long long a = c / (by * bz), rem = c - a * by * bz, b = rem / bz, k = rem - b * bz;
        const int x = x0 + (int)a, y = y0 + (int)b, z = z0 + (int)k, idx = x * yz + y * nz_map + z;
        const bool edge_x = a == 0 || a == bx - 1, edge_y = b == 0 || b == by - 1, edge_z = k == 0 || k == bz - 1;
        const bool boundary = edge_x || edge_y || edge_z;
        inside[idx] = boundary ? 0 : 1;   // (:889-892 clears the flag of every boundary voxel, the one-voxel box included)
        if (cells > 1) use[idx] = 1;
        if (!boundary) continue;
        long long rank = shell_before_slab(a, by, bz, bx);
        if (edge_x) rank += b * bz + k;                                // a full slab
        else {                                                         // an inner slab: full rows y0 / y1, two ends otherwise
            if (b > 0) rank += bz;                                                       // row 0
            if (b > 1) rank += (b - 1 < by - 1 ? b - 1 : by - 2) * (bz > 1 ? 2 : 1);   // rows 1 .. b - 1: their two end voxels
            rank += edge_y ? k : (k == 0 ? 0 : 1);
        }
        if (rank < cap) { cluster_xyz[3 * rank] = x; cluster_xyz[3 * rank + 1] = y; cluster_xyz[3 * rank + 2] = z; }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const long long n = cells == 1 ? 1 : shell_before_slab(bx, by, bz, bx);
        ctl->cluster_num = n <= cap ? (int)n : 0;
        ctl->iters = 0;
        ctl->status = n <= cap ? 0 : -1;
        ctl->skip = (bx == 1 || by == 1 || bz == 1 || n > cap) ? 1 : 0;
    }
}

}  // namespace voxel
#endif
