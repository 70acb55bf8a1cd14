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

// Smallest positive t with s + t ds integer, for s = 0.5 (cluster_engine.cu:19-35: mod(+-0.5, 1) = 0.5).
__device__ __forceinline__ double intbound_half(int ds) {
    if (ds == 0) return 99999.0;
    const int a = ds < 0 ? -ds : ds;
    return (1 - 0.5) / a;
}

// true = the ray from (x, y, z) to (ex, ey, ez) is free (paraConvexTest's d_result[tid], cluster_engine.cu:104-176).
__device__ __forceinline__ bool ray_free(const uint8_t *__restrict__ occ, const uint8_t *__restrict__ inside, int yz, int nz,
                                         int x, int y, int z, int ex, int ey, int ez) {
    const int dx = ex - x, dy = ey - y, dz = ez - z;
    const int sx = dx == 0 ? 0 : (dx < 0 ? -1 : 1), sy = dy == 0 ? 0 : (dy < 0 ? -1 : 1), sz = dz == 0 ? 0 : (dz < 0 ? -1 : 1);
    double tx = intbound_half(dx), ty = intbound_half(dy), tz = intbound_half(dz);
    const double ddx = ((double)sx) / dx, ddy = ((double)sy) / dy, ddz = ((double)sz) / dz;
    bool free_ray = true;
    while (true) {
        if (x == ex && y == ey && z == ez) break;
        if (tx < ty) {
            if (tx < tz) { x += sx; tx += ddx; }
            else { z += sz; tz += ddz; }
        } else {
            if (ty < tz) { y += sy; ty += ddy; }
            else { z += sz; tz += ddz; }
        }
        const int idx = x * yz + y * nz + z;
        if (inside[idx] > 0) return free_ray;
        if (x == ex && y == ey && z == ez) break;
        if (occ[idx] > 0) free_ray = false;
    }
    return free_ray;
}

__global__ void __launch_bounds__(256) convex_test_kernel(const uint8_t *__restrict__ occ, const uint8_t *__restrict__ inside, int yz, int nz,
                                                          const int *__restrict__ cand, int C, const int *__restrict__ clu, int K,
                                                          uint8_t *__restrict__ can_can, uint8_t *__restrict__ can_clu) {
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    // heaviest candidates (most earlier candidates to test against) first
    for (int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < C; w += warps) {
        const int t = C - 1 - w;
        const int x = cand[3 * t], y = cand[3 * t + 1], z = cand[3 * t + 2];
        const long long bias = (long long)t * (t + 1) / 2;   // the reference's packing: n (n - 1) / 2 with n = t + 1
        bool all_clu = true;
        for (int j = lane; j < t + K; j += 32) {
            const int *e = j < t ? cand + 3 * j : clu + 3 * (j - t);
            const bool ok = ray_free(occ, inside, yz, nz, x, y, z, e[0], e[1], e[2]);
            if (j < t) can_can[bias + j] = ok ? 1 : 0;
            else all_clu = all_clu && ok;
        }
        all_clu = __all_sync(0xffffffffu, all_clu);
        if (lane == 0) can_clu[t] = all_clu ? 1 : 0;
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
//   1b  CTA 0 compacts the winning pairs in key order -> the reference's candidate list, element for element;
//   2a  one warp per candidate: rays to the cluster voxels, stop at the first blocked one (can_clu);
//   2b  one warp per SURVIVING candidate: rays to the earlier surviving candidates; the 32 results of a warp step are one ballot
//       word of that candidate's conflict bit-row (a candidate that failed 2a is rejected whatever its rays say and is never in
//       the accepted set, so its rays decide nothing: cluster_server.cu:696-711);
//   3   warp 0 of CTA 0 walks the candidates in order: accepted <=> can_clu and (conflict row & accepted bit-set) == 0; accepted
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
};

constexpr int CLAIM_EMPTY = 0x7f7f7f7f;   // cudaMemset(0x7f)
constexpr int ACC_WORDS = 1024;           // accepted bit-set in shared memory: cand_cap <= 32768

// neighbour k = 0..25 in the reference's dx, dy, dz nesting order (the centre is skipped)
__device__ __forceinline__ bool neighbour_cell(const int *__restrict__ xyz, int i, int k, int nx, int ny, int nz, int &x, int &y, int &z) {
    const int kk = k < 13 ? k : k + 1;
    x = xyz[3 * i] + kk / 9 - 1; y = xyz[3 * i + 1] + (kk / 3) % 3 - 1; z = xyz[3 * i + 2] + kk % 3 - 1;
    return !(x < 0 || x > nx - 1 || y < 0 || y > ny - 1 || z < 0 || z > nz - 1);
}

__global__ void __launch_bounds__(256) cluster_loop_kernel(const uint8_t *__restrict__ occ, const uint8_t *__restrict__ inside, uint8_t *use,
                                                           uint8_t *invalid, int *claim, int nx, int ny, int nz, int *cluster_xyz, int cap,
                                                           int *cand, int cand_cap, unsigned *conflict, uint8_t *can_clu, int itr_cluster_max,
                                                           ClusterCtl *ctl, unsigned long long *phase_ns) {
    cg::grid_group grid = cg::this_grid();
    __shared__ unsigned acc[ACC_WORDS];
    __shared__ int warp_sum[8];
    const int yz = ny * nz, lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x, gthreads = (long long)gridDim.x * blockDim.x;
    const int gwarp = (int)(gtid >> 5), gwarps = (int)(gthreads >> 5);

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
        // 1b: ordered compaction by CTA 0 (each thread owns a contiguous range of keys)
        if (blockIdx.x == 0) {
            const long long P = 26LL * A, chunk = (P + blockDim.x - 1) / blockDim.x;
            const long long lo = chunk * threadIdx.x, hi = lo + chunk < P ? lo + chunk : P;
            int mine = 0;
            for (long long p = lo; p < hi; p++) {
                int x, y, z;
                if (!neighbour_cell(act, (int)(p / 26), (int)(p % 26), nx, ny, nz, x, y, z)) continue;
                mine += claim[x * yz + y * nz + z] == (int)p;
            }
            int incl = mine;
            for (int d = 1; d < 32; d <<= 1) { const int o = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += o; }
            if (lane == 31) warp_sum[wib] = incl;
            __syncthreads();
            int base = 0, total = 0;
            for (int w = 0; w < (int)(blockDim.x >> 5); w++) { if (w < wib) base += warp_sum[w]; total += warp_sum[w]; }
            int at = base + incl - mine;
            const bool fits = total <= cand_cap;
            for (long long p = lo; p < hi; p++) {
                int x, y, z;
                if (!neighbour_cell(act, (int)(p / 26), (int)(p % 26), nx, ny, nz, x, y, z)) continue;
                const int idx = x * yz + y * nz + z;
                if (claim[idx] != (int)p) continue;
                claim[idx] = CLAIM_EMPTY;
                if (fits) { cand[3 * at] = x; cand[3 * at + 1] = y; cand[3 * at + 2] = z; use[idx] = 1; at++; }
            }
            if (threadIdx.x == 0) { ctl->cand_num = fits ? total : 0; if (!fits) ctl->status = -1; }
            for (int w = threadIdx.x; w < ACC_WORDS; w += blockDim.x) acc[w] = 0;
            __syncthreads();
        }
        grid.sync();
        mark(1);
        const int C = ctl->cand_num;
        if (C == 0) break;   // cluster_server.cu:652 (or overflow)
        const int W = (C + 31) >> 5;
        // 2a: candidate -> cluster rays
        for (int t = gwarp; t < C; t += gwarps) {
            const int x = cand[3 * t], y = cand[3 * t + 1], z = cand[3 * t + 2];
            bool ok = true;
            for (int jb = 0; jb < K && ok; jb += 32) {
                const int j = jb + lane;
                const bool mine_ok = j < K ? ray_free(occ, inside, yz, nz, x, y, z, cluster_xyz[3 * j], cluster_xyz[3 * j + 1], cluster_xyz[3 * j + 2]) : true;
                ok = __all_sync(0xffffffffu, mine_ok);
            }
            if (lane == 0) can_clu[t] = ok ? 1 : 0;
        }
        grid.sync();
        mark(2);
        // 2b: surviving candidate -> earlier surviving candidates, heaviest rows first
        for (int w = gwarp; w < C; w += gwarps) {
            const int t = C - 1 - w;
            if (!can_clu[t]) continue;
            const int x = cand[3 * t], y = cand[3 * t + 1], z = cand[3 * t + 2];
            for (int jb = 0; jb < t; jb += 32) {
                const int j = jb + lane;
                bool blocked = false;
                if (j < t && can_clu[j]) blocked = !ray_free(occ, inside, yz, nz, x, y, z, cand[3 * j], cand[3 * j + 1], cand[3 * j + 2]);
                const unsigned word = __ballot_sync(0xffffffffu, blocked);
                if (lane == 0) conflict[(size_t)t * W + (jb >> 5)] = word;
            }
        }
        grid.sync();
        mark(3);
        // 3: acceptance scan (cluster_server.cu:693-737) by CTA 0, 32 candidates at a time: the conflicts with candidates accepted in
        // EARLIER groups are reduced by the whole CTA (rows read in parallel), the order dependence inside a group is resolved by
        // warp 0 from the group's diagonal conflict word in registers (32 shuffle steps, no memory on the dependent chain).
        if (blockIdx.x == 0) {
            __shared__ unsigned prehit;
            int n = K, overflow = 0;   // maintained by warp 0 (uniform across its lanes)
            for (int b = 0; b < C; b += 32) {
                if (threadIdx.x == 0) prehit = 0;
                __syncthreads();
                const int gw = b >> 5;   // words of earlier groups
                for (int q = wib; q < 32; q += (int)(blockDim.x >> 5)) {
                    const int i = b + q;
                    if (i >= C || !can_clu[i]) continue;   // warp-uniform
                    bool hit = false;
                    for (int w = lane; w < gw; w += 32) hit |= (conflict[(size_t)i * W + w] & acc[w]) != 0;
                    if (__any_sync(0xffffffffu, hit) && lane == 0) atomicOr(&prehit, 1u << q);
                }
                __syncthreads();
                if (wib == 0) {
                    const int i = b + lane;
                    const bool valid = i < C;
                    const bool alive = valid && can_clu[i] && !((prehit >> lane) & 1u);
                    const unsigned diag = (alive && lane > 0) ? conflict[(size_t)i * W + gw] : 0u;   // bits j - b < lane
                    int x = 0, y = 0, z = 0;
                    if (valid) { x = cand[3 * i]; y = cand[3 * i + 1]; z = cand[3 * i + 2]; }
                    unsigned accmask = 0;
                    for (int k = 0; k < 32; k++) {
                        const bool ok = alive && (diag & accmask) == 0;
                        if (__shfl_sync(0xffffffffu, (int)ok, k)) accmask |= 1u << k;
                    }
                    const bool mine = (accmask >> lane) & 1u;
                    const int at = n + __popc(accmask & ((1u << lane) - 1u));
                    const bool fits = at < cap;
                    if (mine && fits) { cluster_xyz[3 * at] = x; cluster_xyz[3 * at + 1] = y; cluster_xyz[3 * at + 2] = z; }
                    else if (valid) invalid[x * yz + y * nz + z] = 1;
                    const unsigned kept = __ballot_sync(0xffffffffu, mine && fits);
                    if (lane == 0) acc[gw] = kept;
                    n += __popc(kept);
                    if (kept != accmask) overflow = 1;
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

}  // namespace voxel
#endif
