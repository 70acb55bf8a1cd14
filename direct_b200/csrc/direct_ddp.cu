// direct_ddp.cu -- kernels + C-ABI of libdirect_ddp_b200.so (include/direct_ddp.h).
//
// One persistent kernel per call: every warp pulls trajectory indices from an atomic work queue and
// runs the complete polyCurveGeneration-equivalent solve (ipddp_solver.h) for each, both stages of the
// node's two-stage protocol back to back when asked.  No host round trips, no CPU fallback.
// Built for sm_100a only (see direct_b200/build.py).
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/direct_ddp.h"
#include "../../include/direct_gddp.h"
#include "../../include/direct_voxel.h"
#include "ipddp_solver.h"
#include "gddp.cuh"
#include "gddp_pair.cuh"
#include "voxel.cuh"

#ifndef DDP_MAX_THREADS
#define DDP_MAX_THREADS 128  // four trajectories (warps) per CTA: up to three helpers for the last solve of a CTA
#endif
#ifndef DDP_MIN_BLOCKS
#define DDP_MIN_BLOCKS 2     // => <= 255 registers/thread available, 8 resident trajectories per SM
#endif

namespace {

using ddp::SolveArgs;

template <class R>
__global__ void __launch_bounds__(DDP_MAX_THREADS, DDP_MIN_BLOCKS) ipddp_solve_kernel(SolveArgs A, const R *__restrict__ tabs_g) {
    extern __shared__ __align__(16) unsigned char smraw[];
    R *sm_all = reinterpret_cast<R *>(smraw);
    R *tabs = sm_all;  // 360 table entries shared by the block
    for (int i = threadIdx.x; i < 360; i += blockDim.x) tabs[i] = tabs_g[i];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    const int per_warp = ddp::smem_elems_per_warp(A.PM);
    R *sm = sm_all + 360 + warp * per_warp;
    // job boards (one per warp) + CTA control block behind the per-warp scratch areas, 16-byte aligned
    ddp::JobBoard<R> *boards = reinterpret_cast<ddp::JobBoard<R> *>(
        smraw + (((size_t)(360 + wpb * per_warp) * sizeof(R) + 15) & ~(size_t)15));
    ddp::BlockCtl *ctl = reinterpret_cast<ddp::BlockCtl *>(boards + wpb);
    if (lane == 0) { boards[warp].word = 0ull; boards[warp].done = 0; boards[warp].owner_seq = 0; }
    ddp::ring_init<R>(sm, lane);   // mbarriers of the line search's slack-row ring (ipddp_solver.h "Slack-row ring")
    if (threadIdx.x == 0) { ctl->active_owners = wpb; ctl->unit_running = 0; ctl->jobs_ctr = A.counter + 1; }
    __syncthreads();
    const long long slot = (long long)blockIdx.x * wpb + warp;
    R *ws = reinterpret_cast<R *>(A.ws) + slot * A.ws_stride;
    while (true) {
        unsigned int b = 0;
        if (lane == 0) b = atomicAdd(A.counter, 1u);
        b = __shfl_sync(0xffffffffu, b, 0);
        if (b >= (unsigned int)A.B) break;
        if (A.two_stage) ddp::solve_one<R>(A, 0, (int)b, sm, tabs, ws, lane, boards + warp, A.coop ? ctl : nullptr, wpb);
        ddp::solve_one<R>(A, A.two_stage ? 1 : 0, (int)b, sm, tabs, ws, lane, boards + warp, A.coop ? ctl : nullptr, wpb);
        __threadfence();
        __syncwarp();
        if (lane == 0) atomicAdd(A.counter + 3, 1u);   // trajectories finished (the warps of idle CTAs leave when all are)
    }
    // queue empty: help the warps of this CTA that still own a trajectory (ipddp_solver.h "Intra-CTA cooperation")
    if (lane == 0) atomicSub(&ctl->active_owners, 1);
    __syncwarp();
    if (wpb > 1 && A.coop) ddp::helper_loop<R>(boards, ctl, wpb, warp, sm, lane, A.counter + 2);
    // no warp of this CTA owns a trajectory any more: run whole line-search trials of the remaining solves of other CTAs
    // ("Speculative line search").  Measured: letting idle warps of CTAs that still own a trajectory take remote trials as
    // well is slower (204 vs 196 ms at B = 4096): their owner loses its row helpers while the SM is still contended.
    if (A.gspec) ddp::gspec_helper_loop<R>(A, sm, tabs, ws, (int)slot, lane, boards, (wpb > 1 && A.coop) ? ctl : nullptr, wpb, warp);
}

// initTimeAllocation, teach_repeat_planner.cpp:583-639 (v0 = 0): one thread per segment.
__global__ void time_allocation_kernel(int B, int N, const double *start, const double *end, const double *seeds,
                                       double vel, double accl, double *durations) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)B * N) return;
    const int b = (int)(idx / N), k = (int)(idx % N);
    double p0[3], p1[3];
    for (int a = 0; a < 3; a++) {
        p0[a] = (k == 0) ? start[(long long)b * 3 + a] : seeds[((long long)b * N + k) * 3 + a];
        p1[a] = (k == N - 1) ? end[(long long)b * 3 + a] : seeds[((long long)b * N + k + 1) * 3 + a];
    }
    const double d0 = p1[0] - p0[0], d1 = p1[1] - p0[1], d2 = p1[2] - p0[2];
    const double D = sqrt(d0 * d0 + d1 * d1 + d2 * d2);
    const double V0 = 0.0 * (d0 / D) + 0.0 * (d1 / D) + 0.0 * (d2 / D);
    const double aV0 = fabs(V0);
    const double acct = (vel - V0) / accl * ((vel > V0) ? 1 : -1);
    const double accd = V0 * acct + (accl * acct * acct / 2) * ((vel > V0) ? 1 : -1);
    const double dcct = vel / accl, dccd = accl * dcct * dcct / 2;
    double dt;
    if (D < aV0 * aV0 / (2 * accl)) {
        dt = ((V0 < 0) ? 2.0 * aV0 / accl : 0.0) + aV0 / accl;
    } else if (D < accd + dccd) {
        const double t1 = (V0 < 0) ? 2.0 * aV0 / accl : 0.0;
        const double t2 = (-aV0 + sqrt(aV0 * aV0 + accl * D - aV0 * aV0 / 2)) / accl;
        dt = t1 + t2 + (aV0 + accl * t2) / accl;
    } else {
        dt = acct + (D - accd - dccd) / vel + dcct;
    }
    durations[idx] = dt;
}

// Bernstein evaluation of solved trajectories (utils/bezier_base.h:77-115), one thread per (segment, sample).
// HBM-bound: 19 doubles read per segment (shared by its S threads through L1), 9 doubles written per sample.
__global__ void bezier_sample_kernel(long long nseg, int S, const double *__restrict__ bez, const double *__restrict__ times,
                                     double *__restrict__ pos, double *__restrict__ vel, double *__restrict__ acc) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nseg * S) return;
    const long long seg = idx / S;
    const int k = (int)(idx % S);
    const double s = S > 1 ? (double)k / (double)(S - 1) : 0.0, q = 1.0 - s, T = times[seg];
    double sp[6], qp[6];
    sp[0] = 1.0; qp[0] = 1.0;
    for (int j = 1; j < 6; j++) { sp[j] = sp[j - 1] * s; qp[j] = qp[j - 1] * q; }
    const double c5[6] = {1, 5, 10, 10, 5, 1}, c4[5] = {1, 4, 6, 4, 1}, c3[4] = {1, 3, 3, 1};
    const double *cp = bez + seg * 18;
    for (int a = 0; a < 3; a++) {
        double c[6];
        for (int j = 0; j < 6; j++) c[j] = cp[a * 6 + j];
        double p = 0.0, v = 0.0, w = 0.0;
        for (int j = 0; j < 6; j++) p += c5[j] * c[j] * sp[j] * qp[5 - j];
        for (int j = 0; j < 5; j++) v += c4[j] * 5.0 * (c[j + 1] - c[j]) * sp[j] * qp[4 - j];
        for (int j = 0; j < 4; j++) w += c3[j] * 20.0 * (c[j + 2] - 2.0 * c[j + 1] + c[j]) * sp[j] * qp[3 - j];
        if (pos) pos[idx * 3 + a] = T * p;
        if (vel) vel[idx * 3 + a] = v;
        if (acc) acc[idx * 3 + a] = w / T;
    }
}

// Register-resident FMA throughput probe: the measured denominator of the FMA roofline (MEASURED_PEAKS.json
// only carries HBM and bf16 tensor numbers).  8 independent chains per thread, 2 flops per FMA.
template <class R> __global__ void fma_peak_kernel(R *out, int iters, R a, R b) {
    R x0 = a + threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; i++) {
        x0 = x0 * a + b; x1 = x1 * a + b; x2 = x2 * a + b; x3 = x3 * a + b;
        x4 = x4 * a + b; x5 = x5 * a + b; x6 = x6 * a + b; x7 = x7 * a + b;
    }
    if (x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7 == R(12345.678)) out[0] = x0;
}

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
};

}  // namespace

struct direct_ddp_handle_s {
    direct_ddp_opts opts;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    std::string err;
    direct_ddp_stats stats;
    // grow-only device buffers
    DevBuf planes, nplanes, durations, seeds, x0, xd, init_bez, infeas, nknots;
    DevBuf o_int[2], o_cost[2], o_xf[2], o_pc[2], o_bz[2], o_pt[2], o_jk[2], o_st[2];
    DevBuf ws, counter, tabs, bez_tmp, time_tmp, trace, trace_len, scratch_i, gboards, gwords;
    DevBuf gd[8];   // generic DDP (gddp.cuh): x0, xg, u_init, ints, cost, x, u, stats
    DevBuf vx[19];  // voxel kernels (voxel.cuh): occupied, inside, candidates, cluster, can_can, can_clu, vertices, result,
                    // claim, loop candidates, conflict rows, loop can_clu, ctl, use, invalid, phase timers, merged map, segment counts, reciprocal table
    int vx_coop_blocks = 0;   // co-resident CTAs of cluster_loop_kernel
    std::vector<direct_ddp_handle_s *> peers;   // multi-GPU: handles of devices[1..] (owned); this handle is devices[0]
    bool last_multi = false;                    // the last host-buffer call was sharded over the peers
    const long long *last_stats_dev = nullptr;  // device [B][4] of the last solve (stage 1 / single)
    const long long *last_stats_dev0 = nullptr; // stage 0 when two-stage
    int last_B = 0;
    bool stats_valid = false;
};

namespace {

typedef direct_ddp_handle_s H;

bool cuda_ok(H *h, cudaError_t e, const char *what) {
    if (e == cudaSuccess) return true;
    h->err = std::string(what) + ": " + cudaGetErrorString(e);
    return false;
}
#define CK(call)                                                 \
    do {                                                         \
        if (!cuda_ok(h, (call), #call)) return DIRECT_DDP_ERR_CUDA; \
    } while (0)

int ensure(H *h, DevBuf &b, size_t bytes) {
    if (bytes == 0) bytes = 16;
    if (b.cap >= bytes) return 0;
    if (b.p) cudaFree(b.p);
    b.p = nullptr; b.cap = 0;
    if (!cuda_ok(h, cudaMalloc(&b.p, bytes), "cudaMalloc")) return DIRECT_DDP_ERR_NOMEM;
    b.cap = bytes;
    return 0;
}

int validate(H *h, const direct_ddp_batch *in) {
    if (!in) { h->err = "batch is NULL"; return DIRECT_DDP_ERR_ARG; }
    if (in->B <= 0 || in->N <= 0) { h->err = "B and N must be positive"; return DIRECT_DDP_ERR_ARG; }
    if (in->P_max < 0 || in->P_max > DIRECT_DDP_MAX_PLANES) { h->err = "P_max must be in [0, DIRECT_DDP_MAX_PLANES]"; return DIRECT_DDP_ERR_ARG; }
    if (!in->planes || !in->nplanes || !in->durations || !in->x0 || !in->xd) { h->err = "missing input pointer"; return DIRECT_DDP_ERR_ARG; }
    return 0;
}
int validate_cfg(H *h, int time_power, int line_init, int iter_max) {
    if (time_power != 1 && time_power != 2) { h->err = "time_power must be 1 or 2"; return DIRECT_DDP_ERR_ARG; }
    if (iter_max < 0) { h->err = "iter_max must be >= 0"; return DIRECT_DDP_ERR_ARG; }
    (void)line_init;
    return 0;
}

template <class R> int upload_tables(H *h) {
    std::vector<R> t(360);
    const ddp::BasisTables &bt = ddp::basis_tables();
    for (int m = 0; m < 2; m++)
        for (int e = 0; e < 90; e++) { t[m * 180 + e] = (R)bt.val[m][e]; t[m * 180 + 90 + e] = (R)bt.dt[e]; }
    int st = ensure(h, h->tabs, 360 * sizeof(R));
    if (st) return st;
    CK(cudaMemcpyAsync(h->tabs.p, t.data(), 360 * sizeof(R), cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

// Launch the persistent solve kernel on device-resident arguments.
template <class R> int launch(H *h, SolveArgs &A, cudaStream_t s) {
    const int wpb = h->opts.warps_per_block > 0 ? h->opts.warps_per_block : DDP_MAX_THREADS / 32;
    const int threads = wpb * 32;
    if (threads > DDP_MAX_THREADS) { h->err = "warps_per_block exceeds the kernel's launch bound"; return DIRECT_DDP_ERR_ARG; }
    const size_t smem = (((size_t)(360 + wpb * ddp::smem_elems_per_warp(A.PM)) * sizeof(R) + 15) & ~(size_t)15) +
                        (size_t)ddp::coop_smem_bytes<R>(wpb);
    auto kern = ipddp_solve_kernel<R>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (const char *e = getenv("DIRECT_DDP_CARVEOUT"))   // tuning knob: shared-memory share of the unified L1 (percent)
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(e)));
    int per_sm = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem));
    if (per_sm < 1) { h->err = "kernel does not fit on an SM (shared memory)"; return DIRECT_DDP_ERR_CUDA; }
    if (h->opts.blocks_per_sm > 0 && per_sm > h->opts.blocks_per_sm) per_sm = h->opts.blocks_per_sm;
    long long grid = (long long)per_sm * h->sm_count;
    const long long need = (A.B + wpb - 1) / wpb;
    if (grid > need) {
        // A small batch (the node's one corridor at a time) still gets idle CTAs: their warps run its line-search trials and
        // backward sweeps speculatively (ipddp_solver.h "Speculative line search").  Tuning knob: DIRECT_DDP_MIN_GRID (CTAs).
        // Measured (profiles/r2w_min_grid.md): one CTA per SM is the best floor - a single hard corridor of 100 knots 113 -> 47 ms.
        long long min_grid = h->sm_count;
        if (const char *e = getenv("DIRECT_DDP_MIN_GRID")) min_grid = atoll(e);
        if (min_grid > grid) min_grid = grid;
        grid = need > min_grid ? need : min_grid;
    }
    int max_iter = A.cfg[0].iter_max;
    if (A.two_stage && A.cfg[1].iter_max > max_iter) max_iter = A.cfg[1].iter_max;
    A.fcap = max_iter + 2;
    const ddp::WsLay wl = ddp::ws_layout(A.N, A.PM, A.fcap);
    A.ws_stride = wl.total;
    const size_t ws_bytes = (size_t)grid * wpb * wl.total * sizeof(R) + 65536;   // + slack: the row loops prefetch a few rows past the last array
    if (h->ws.cap < ws_bytes) {
        int st = ensure(h, h->ws, ws_bytes);
        if (st) return st;
        CK(cudaMemsetAsync(h->ws.p, 0, ws_bytes, s));
    }
    A.ws = h->ws.p;
    int st = ensure(h, h->counter, 64);
    if (st) return st;
    CK(cudaMemsetAsync(h->counter.p, 0, 64, s));
    A.counter = (unsigned int *)h->counter.p;
    {
        const char *e = getenv("DIRECT_DDP_COOP");   // tuning knob: 0 disables the tail balancing
        A.coop = (e && atoi(e) == 0) ? 0 : 1;
        e = getenv("DIRECT_DDP_GSPEC");              // tuning knob: 0 disables the speculative line search of the tail
        A.gspec = (e && atoi(e) == 0) ? 0 : A.coop;
        e = getenv("DIRECT_DDP_SPEC");               // tuning knob: 0 disables the speculative backward sweep of the tail
        A.spec = (e && atoi(e) == 0) ? 0 : A.coop;
    }
    A.gboards = nullptr; A.gwords = nullptr;
    if (A.gspec) {
        const size_t nb = (size_t)ddp::GSPEC_BOARDS * grid * wpb;
        if ((st = ensure(h, h->gboards, nb * sizeof(ddp::GBoard<R>)))) return st;
        const size_t wbytes = nb * sizeof(unsigned long long) + ((nb + 31) / 32 + 1) * sizeof(unsigned int);   // words + bitmap
        if ((st = ensure(h, h->gwords, wbytes))) return st;
        CK(cudaMemsetAsync(h->gboards.p, 0, nb * sizeof(ddp::GBoard<R>), s));
        CK(cudaMemsetAsync(h->gwords.p, 0, wbytes, s));
        A.gboards = h->gboards.p; A.gwords = (unsigned long long *)h->gwords.p;
    }
    if (h->opts.trace) {
        if ((st = ensure(h, h->trace, 512 * 12 * sizeof(double)))) return st;
        if ((st = ensure(h, h->trace_len, 16))) return st;
        CK(cudaMemsetAsync(h->trace_len.p, 0, 16, s));
        A.trace = (double *)h->trace.p; A.trace_cap = 512; A.trace_len = (int *)h->trace_len.p;
    }
    CK(cudaEventRecord(h->ev[2], s));
    kern<<<(unsigned)grid, threads, smem, s>>>(A, (const R *)h->tabs.p);
    CK(cudaGetLastError());
    CK(cudaEventRecord(h->ev[3], s));
    h->stats.kernel_launches = 1;
    h->stats.grid_blocks = (int)grid; h->stats.block_threads = threads;
    h->stats.smem_bytes_per_block = (int)smem; h->stats.workspace_slots = (int)(grid * wpb);
    return 0;
}

void fill_cfg(ddp::StageCfg &c, double ws, double wt, double wtime, int iters, int tp, int zero, int line, int minvo, int inf) {
    c.w_snap = ws; c.w_terminal = wt; c.w_time = wtime; c.iter_max = iters; c.time_power = tp;
    c.zero_init = zero; c.line_init = line; c.minvo = minvo; c.infeas_all = inf;
}

void fill_out(ddp::OutPtrs &O, const direct_ddp_result *r) {
    memset(&O, 0, sizeof O);
    if (!r) return;
    O.rtn = r->rtn; O.infeas_out = r->infeas_out; O.line_failed_out = r->line_failed_out; O.iters = r->iters;
    O.cost = r->cost; O.x_final = r->x_final; O.poly_coeff = r->poly_coeff; O.bez_coeff = r->bez_coeff;
    O.poly_time = r->poly_time; O.jerk = r->jerk; O.stats = (long long *)r->stats;
}

// Core: everything device-resident.  `ts` NULL = single stage using the batch's own scalars.
int solve_device(H *h, const direct_ddp_batch *in, const direct_ddp_two_stage *ts, const direct_ddp_result *out0,
                 const direct_ddp_result *out1, cudaStream_t s) {
    int st = validate(h, in);
    if (st) return st;
    if (!out1) { h->err = "result is NULL"; return DIRECT_DDP_ERR_ARG; }
    if (((size_t)in->planes & 15) != 0) { h->err = "planes must be 16-byte aligned (128-bit loads)"; return DIRECT_DDP_ERR_ARG; }
    SolveArgs A;
    memset(&A, 0, sizeof A);
    A.B = in->B; A.N = in->N; A.PM = in->P_max;
    A.planes = in->planes; A.nplanes = in->nplanes; A.durations = in->durations; A.seeds = in->seeds;
    A.x0 = in->x0; A.xd = in->xd; A.max_vel = in->max_vel; A.max_acc = in->max_acc;
    A.nknots = in->nknots;
    if (ts) {
        if ((st = validate_cfg(h, ts->time_power, 0, ts->iter_max0))) return st;
        if ((st = validate_cfg(h, ts->time_power, 0, ts->iter_max))) return st;
        A.two_stage = 1;
        // teach_repeat_planner.cpp:895-897 and :918-921
        fill_cfg(A.cfg[0], ts->w_snap0, ts->w_terminal0, ts->w_time0, ts->iter_max0, ts->time_power, 1, 0, 0, 1);
        fill_cfg(A.cfg[1], ts->w_snap, ts->w_terminal, ts->w_time, ts->iter_max, ts->time_power, 0, 0, 0, 0);
        fill_out(A.out[0], out0);
        fill_out(A.out[1], out1);
        if ((st = ensure(h, h->bez_tmp, (size_t)in->B * in->N * 18 * sizeof(double)))) return st;
        if ((st = ensure(h, h->time_tmp, (size_t)in->B * in->N * sizeof(double)))) return st;
        A.bez_tmp = (double *)h->bez_tmp.p; A.time_tmp = (double *)h->time_tmp.p;
        if (!A.out[0].rtn || !A.out[0].infeas_out) {  // stage 1 needs stage 0's rtn and infeas
            if ((st = ensure(h, h->scratch_i, (size_t)in->B * 2 * sizeof(int32_t)))) return st;
            if (!A.out[0].rtn) A.out[0].rtn = (int32_t *)h->scratch_i.p;
            if (!A.out[0].infeas_out) A.out[0].infeas_out = (int32_t *)h->scratch_i.p + in->B;
        }
    } else {
        if ((st = validate_cfg(h, in->time_power, in->line_init, in->iter_max))) return st;
        if (in->line_init && !in->seeds) { h->err = "line_init needs seeds"; return DIRECT_DDP_ERR_ARG; }   // ddp_optimizer.cpp:195-247
        A.two_stage = 0;
        A.init_bez = in->init_bez; A.infeas = in->infeas;
        fill_cfg(A.cfg[0], in->w_snap, in->w_terminal, in->w_time, in->iter_max, in->time_power, in->zero_init,
                 in->line_init, in->minvo, in->infeas_all);
        fill_out(A.out[1], out1);
    }
    // statistics are always collected (roofline accounting): use internal buffers when the caller has none
    for (int k = ts ? 0 : 1; k < 2; k++) {
        if (!A.out[k].stats) {
            if ((st = ensure(h, h->o_st[k], (size_t)in->B * 8 * sizeof(long long)))) return st;
            A.out[k].stats = (long long *)h->o_st[k].p;
        }
    }
    h->last_stats_dev = A.out[1].stats;
    h->last_stats_dev0 = ts ? A.out[0].stats : nullptr;
    h->last_B = in->B;
    h->stats_valid = false;
    if (h->opts.precision == DIRECT_DDP_FP32) return launch<float>(h, A, s);
    return launch<double>(h, A, s);
}

// Device mirror of a host result: allocate what the caller asked for.
int mirror_result(H *h, int k, const direct_ddp_result *host, int B, int N, direct_ddp_result *dev) {
    memset(dev, 0, sizeof *dev);
    if (!host) return 0;
    int st;
    if ((st = ensure(h, h->o_int[k], (size_t)B * 4 * sizeof(int32_t)))) return st;
    int32_t *ip = (int32_t *)h->o_int[k].p;
    dev->rtn = ip; dev->infeas_out = ip + B; dev->line_failed_out = ip + 2 * B; dev->iters = ip + 3 * B;
    if ((st = ensure(h, h->o_cost[k], (size_t)B * 8))) return st;
    dev->cost = (double *)h->o_cost[k].p;
    if ((st = ensure(h, h->o_xf[k], (size_t)B * 9 * 8))) return st;
    dev->x_final = (double *)h->o_xf[k].p;
    if (host->poly_coeff) { if ((st = ensure(h, h->o_pc[k], (size_t)B * N * 18 * 8))) return st; dev->poly_coeff = (double *)h->o_pc[k].p; }
    if (host->bez_coeff) { if ((st = ensure(h, h->o_bz[k], (size_t)B * N * 18 * 8))) return st; dev->bez_coeff = (double *)h->o_bz[k].p; }
    if (host->poly_time) { if ((st = ensure(h, h->o_pt[k], (size_t)B * N * 8))) return st; dev->poly_time = (double *)h->o_pt[k].p; }
    if (host->jerk) { if ((st = ensure(h, h->o_jk[k], (size_t)B * N * 8))) return st; dev->jerk = (double *)h->o_jk[k].p; }
    if ((st = ensure(h, h->o_st[k], (size_t)B * 8 * 8))) return st;
    dev->stats = (int64_t *)h->o_st[k].p;
    return 0;
}

int download_result(H *h, const direct_ddp_result *host, const direct_ddp_result *dev, int B, int N, cudaStream_t s,
                    int64_t *bytes) {
    if (!host) return 0;
#define DL(field, n)                                                                                         \
    if (host->field && dev->field) {                                                                         \
        CK(cudaMemcpyAsync(host->field, dev->field, (size_t)(n), cudaMemcpyDeviceToHost, s));                \
        *bytes += (int64_t)(n);                                                                              \
    }
    DL(rtn, (size_t)B * 4) DL(infeas_out, (size_t)B * 4) DL(line_failed_out, (size_t)B * 4) DL(iters, (size_t)B * 4)
    DL(cost, (size_t)B * 8) DL(x_final, (size_t)B * 72) DL(poly_coeff, (size_t)B * N * 144)
    DL(bez_coeff, (size_t)B * N * 144) DL(poly_time, (size_t)B * N * 8) DL(jerk, (size_t)B * N * 8)
    DL(stats, (size_t)B * 64)
#undef DL
    return 0;
}

int solve_host_one(H *h, const direct_ddp_batch *in, const direct_ddp_two_stage *ts, direct_ddp_result *out0,
                   direct_ddp_result *out1) {
    int st = validate(h, in);
    if (st) return st;
    if (!out1) { h->err = "result is NULL"; return DIRECT_DDP_ERR_ARG; }
    CK(cudaSetDevice(h->opts.device));
    cudaStream_t s = h->stream;
    const int B = in->B, N = in->N, PM = in->P_max;
    direct_ddp_batch d = *in;
    int64_t h2d = 0, d2h = 0;
    CK(cudaEventRecord(h->ev[0], s));
#define UP(buf, field, bytes, type)                                                                          \
    if (in->field) {                                                                                         \
        if ((st = ensure(h, h->buf, (size_t)(bytes)))) return st;                                            \
        CK(cudaMemcpyAsync(h->buf.p, in->field, (size_t)(bytes), cudaMemcpyHostToDevice, s));                \
        d.field = (const type *)h->buf.p;                                                                    \
        h2d += (int64_t)(bytes);                                                                             \
    }
    for (size_t e = 0; e < (size_t)B * N; e++)   // a count past P_max would read past the polytope's rows on the device
        if (in->nplanes[e] < 0 || in->nplanes[e] > PM) { h->err = "nplanes must be in [0, P_max]"; return DIRECT_DDP_ERR_ARG; }
    UP(planes, planes, (size_t)B * N * PM * 32, double)
    UP(nplanes, nplanes, (size_t)B * N * 4, int32_t)
    UP(durations, durations, (size_t)B * N * 8, double)
    UP(x0, x0, (size_t)B * 72, double)
    UP(xd, xd, (size_t)B * 72, double)
    if (in->nknots) {
        for (int b = 0; b < B; b++)
            if (in->nknots[b] < 1 || in->nknots[b] > N) { h->err = "nknots must be in [1, N]"; return DIRECT_DDP_ERR_ARG; }
    }
    UP(nknots, nknots, (size_t)B * 4, int32_t)
    if (!ts) {
        UP(init_bez, init_bez, (size_t)B * N * 144, double)
        UP(infeas, infeas, (size_t)B * 4, int32_t)
    }
    if (!ts && in->line_init) {  // only line_init reads the seeds (ddp_optimizer.cpp:195-247)
        if (!in->seeds) { h->err = "line_init needs seeds"; return DIRECT_DDP_ERR_ARG; }
        UP(seeds, seeds, (size_t)B * N * 24, double)
    } else d.seeds = nullptr;
#undef UP
    CK(cudaEventRecord(h->ev[1], s));
    direct_ddp_result dev0, dev1;
    if ((st = mirror_result(h, 0, ts ? out0 : nullptr, B, N, &dev0))) return st;
    if ((st = mirror_result(h, 1, out1, B, N, &dev1))) return st;
    if (in->nknots) {   // ragged batch: entries past a trajectory's own knot count come back as zeros
        for (direct_ddp_result *dv : {&dev0, &dev1}) {
            if (dv->poly_coeff) CK(cudaMemsetAsync(dv->poly_coeff, 0, (size_t)B * N * 144, s));
            if (dv->bez_coeff) CK(cudaMemsetAsync(dv->bez_coeff, 0, (size_t)B * N * 144, s));
            if (dv->poly_time) CK(cudaMemsetAsync(dv->poly_time, 0, (size_t)B * N * 8, s));
            if (dv->jerk) CK(cudaMemsetAsync(dv->jerk, 0, (size_t)B * N * 8, s));
        }
    }
    if ((st = solve_device(h, &d, ts, (ts && out0) ? &dev0 : nullptr, &dev1, s))) return st;
    CK(cudaEventRecord(h->ev[4], s));
    if (ts && out0 && (st = download_result(h, out0, &dev0, B, N, s, &d2h))) return st;
    if ((st = download_result(h, out1, &dev1, B, N, s, &d2h))) return st;
    CK(cudaEventRecord(h->ev[5], s));
    CK(cudaStreamSynchronize(s));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, h->ev[0], h->ev[1])); h->stats.h2d_ms = ms;
    CK(cudaEventElapsedTime(&ms, h->ev[2], h->ev[3])); h->stats.kernel_ms = ms;
    CK(cudaEventElapsedTime(&ms, h->ev[4], h->ev[5])); h->stats.d2h_ms = ms;
    h->stats.h2d_bytes = h2d; h->stats.d2h_bytes = d2h;
    return 0;
}

// Shard [lo, hi) of a host batch / result: the same struct with every per-trajectory pointer advanced.
direct_ddp_batch shard_batch(const direct_ddp_batch &in, int lo, int hi) {
    direct_ddp_batch d = in;
    const size_t N = (size_t)in.N, PM = (size_t)in.P_max, o = (size_t)lo;
    d.B = hi - lo;
    d.planes = in.planes + o * N * PM * 4; d.nplanes = in.nplanes + o * N; d.durations = in.durations + o * N;
    if (in.seeds) d.seeds = in.seeds + o * N * 3;
    d.x0 = in.x0 + o * 9; d.xd = in.xd + o * 9;
    if (in.init_bez) d.init_bez = in.init_bez + o * N * 18;
    if (in.infeas) d.infeas = in.infeas + o;
    if (in.nknots) d.nknots = in.nknots + o;
    return d;
}
direct_ddp_result shard_result(const direct_ddp_result &r, int lo, int N) {
    direct_ddp_result d = r;
    const size_t o = (size_t)lo, n = (size_t)N;
#define ADV(f, stride) if (r.f) d.f = r.f + o * (stride)
    ADV(rtn, 1); ADV(infeas_out, 1); ADV(line_failed_out, 1); ADV(iters, 1); ADV(cost, 1); ADV(x_final, 9);
    ADV(poly_coeff, n * 18); ADV(bez_coeff, n * 18); ADV(poly_time, n); ADV(jerk, n); ADV(stats, 8);
#undef ADV
    return d;
}

// Host-buffer solve: one device, or the batch sharded over the handle's devices with one host thread per device
// (contiguous ranges, no collective -- every shard's results go D2H straight into the caller's arrays).
int solve_host(H *h, const direct_ddp_batch *in, const direct_ddp_two_stage *ts, direct_ddp_result *out0,
               direct_ddp_result *out1) {
    h->last_multi = false;
    const int ndev = 1 + (int)h->peers.size();
    if (ndev == 1 || !in || in->B < ndev) return solve_host_one(h, in, ts, out0, out1);
    int st = validate(h, in);
    if (st) return st;
    if (!out1) { h->err = "result is NULL"; return DIRECT_DDP_ERR_ARG; }
    std::vector<direct_ddp_batch> sb(ndev);
    std::vector<direct_ddp_result> r0(ndev), r1(ndev);
    std::vector<int> status(ndev, 0);
    std::vector<std::thread> th;
    auto run = [&](int k) {
        H *hk = k == 0 ? h : h->peers[k - 1];
        status[k] = solve_host_one(hk, &sb[k], ts, (ts && out0) ? &r0[k] : nullptr, &r1[k]);
    };
    for (int k = 0; k < ndev; k++) {
        const int lo = (int)((long long)in->B * k / ndev), hi = (int)((long long)in->B * (k + 1) / ndev);
        sb[k] = shard_batch(*in, lo, hi);
        if (ts && out0) r0[k] = shard_result(*out0, lo, in->N);
        r1[k] = shard_result(*out1, lo, in->N);
    }
    for (int k = 1; k < ndev; k++) th.emplace_back(run, k);
    run(0);
    for (auto &t : th) t.join();
    for (int k = 0; k < ndev; k++)
        if (status[k]) {
            if (k > 0) h->err = "device " + std::to_string(h->peers[k - 1]->opts.device) + ": " + h->peers[k - 1]->err;
            return status[k];
        }
    h->last_multi = true;
    return 0;
}

}  // namespace

extern "C" {

int direct_ddp_version(void) { return DIRECT_DDP_VERSION; }

int direct_ddp_create(const direct_ddp_opts *opts, direct_ddp_handle *out) {
    if (!out) return DIRECT_DDP_ERR_ARG;
    *out = nullptr;
    H *h = new H();
    memset(&h->opts, 0, sizeof h->opts);
    if (opts) h->opts = *opts;
    if (h->opts.ndevices < 0 || h->opts.ndevices > 64) { h->err = "ndevices must be in [0, 64]"; *out = h; return DIRECT_DDP_ERR_ARG; }
    if (h->opts.ndevices > 0 && h->opts.devices) h->opts.device = h->opts.devices[0];
    memset(&h->stats, 0, sizeof h->stats);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev <= 0 || h->opts.device < 0 || h->opts.device >= ndev) {
        // No CPU fallback: the handle is returned so the caller can read the error text, every solve fails.
        h->err = std::string("no usable CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "bad ordinal");
        h->sm_count = 0;
        *out = h;
        return DIRECT_DDP_ERR_CUDA;
    }
    cudaDeviceProp prop;
    if (cudaSetDevice(h->opts.device) != cudaSuccess || cudaGetDeviceProperties(&prop, h->opts.device) != cudaSuccess) {
        h->err = "cudaSetDevice/cudaGetDeviceProperties failed";
        *out = h;
        return DIRECT_DDP_ERR_CUDA;
    }
    h->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) { h->err = "stream"; *out = h; return DIRECT_DDP_ERR_CUDA; }
    for (int i = 0; i < 6; i++) cudaEventCreate(&h->ev[i]);
    int st = h->opts.precision == DIRECT_DDP_FP32 ? upload_tables<float>(h) : upload_tables<double>(h);
    *out = h;
    if (st) return st;
    // multi-GPU: one ordinary single-device handle per further device
    const int nd = h->opts.ndevices;
    const int *devs = h->opts.devices;
    const int first = h->opts.device;
    h->opts.ndevices = 0; h->opts.devices = nullptr;   // the caller's array is not kept
    for (int k = 1; k < nd; k++) {
        direct_ddp_opts po = h->opts;
        po.device = devs ? devs[k] : first + k;
        po.trace = 0;
        direct_ddp_handle peer = nullptr;
        const int ps = direct_ddp_create(&po, &peer);
        if (ps != DIRECT_DDP_OK) {
            h->err = "device " + std::to_string(po.device) + ": " + (peer ? peer->err : std::string("create failed"));
            if (peer) direct_ddp_destroy(peer);
            h->sm_count = 0;   // the handle stays readable for the error text; every solve fails
            return ps;
        }
        h->peers.push_back(peer);
    }
    return 0;
}

int direct_ddp_device_count(direct_ddp_handle h) { return h ? 1 + (int)h->peers.size() : 0; }

void direct_ddp_destroy(direct_ddp_handle h) {
    if (!h) return;
    for (direct_ddp_handle_s *p : h->peers) direct_ddp_destroy(p);
    h->peers.clear();
    if (h->sm_count > 0) {
        cudaSetDevice(h->opts.device);
        DevBuf *bufs[] = {&h->planes, &h->nplanes, &h->durations, &h->seeds, &h->x0, &h->xd, &h->init_bez, &h->infeas, &h->nknots,
                          &h->ws, &h->counter, &h->tabs, &h->bez_tmp, &h->time_tmp, &h->trace, &h->trace_len, &h->scratch_i,
                          &h->gboards, &h->gwords};
        for (DevBuf *b : bufs) if (b->p) cudaFree(b->p);
        for (DevBuf &b : h->gd) if (b.p) cudaFree(b.p);
        for (DevBuf &b : h->vx) if (b.p) cudaFree(b.p);
        for (int k = 0; k < 2; k++) {
            DevBuf *ob[] = {&h->o_int[k], &h->o_cost[k], &h->o_xf[k], &h->o_pc[k], &h->o_bz[k], &h->o_pt[k], &h->o_jk[k], &h->o_st[k]};
            for (DevBuf *b : ob) if (b->p) cudaFree(b->p);
        }
        for (int i = 0; i < 6; i++) if (h->ev[i]) cudaEventDestroy(h->ev[i]);
        if (h->stream) cudaStreamDestroy(h->stream);
    }
    delete h;
}

const char *direct_ddp_last_error(direct_ddp_handle h) { return h ? h->err.c_str() : "null handle"; }

#define REQUIRE_DEVICE(h)                                                      \
    if (!(h)) return DIRECT_DDP_ERR_ARG;                                       \
    if ((h)->sm_count <= 0) { if ((h)->err.empty()) (h)->err = "no CUDA device"; return DIRECT_DDP_ERR_CUDA; }

int direct_ddp_solve_batch(direct_ddp_handle h, const direct_ddp_batch *in, direct_ddp_result *out) {
    REQUIRE_DEVICE(h)
    return solve_host(h, in, nullptr, nullptr, out);
}

int direct_ddp_solve_two_stage(direct_ddp_handle h, const direct_ddp_batch *in, const direct_ddp_two_stage *ts,
                               direct_ddp_result *out0, direct_ddp_result *out1) {
    REQUIRE_DEVICE(h)
    if (!ts) { h->err = "two_stage options are NULL"; return DIRECT_DDP_ERR_ARG; }
    return solve_host(h, in, ts, out0, out1);
}

int direct_ddp_solve_batch_device(direct_ddp_handle h, const direct_ddp_batch *in, direct_ddp_result *out, void *stream) {
    REQUIRE_DEVICE(h)
    CK(cudaSetDevice(h->opts.device));
    return solve_device(h, in, nullptr, nullptr, out, (cudaStream_t)stream);
}

int direct_ddp_solve_two_stage_device(direct_ddp_handle h, const direct_ddp_batch *in, const direct_ddp_two_stage *ts,
                                      direct_ddp_result *out0, direct_ddp_result *out1, void *stream) {
    REQUIRE_DEVICE(h)
    if (!ts) { h->err = "two_stage options are NULL"; return DIRECT_DDP_ERR_ARG; }
    CK(cudaSetDevice(h->opts.device));
    return solve_device(h, in, ts, out0, out1, (cudaStream_t)stream);
}

int direct_ddp_time_allocation_device(direct_ddp_handle h, int B, int N, const double *start, const double *end,
                                      const double *seeds, double max_vel, double max_acc, double *durations, void *stream) {
    REQUIRE_DEVICE(h)
    if (B <= 0 || N <= 0 || !start || !end || !seeds || !durations) { h->err = "bad argument"; return DIRECT_DDP_ERR_ARG; }
    CK(cudaSetDevice(h->opts.device));
    const long long n = (long long)B * N;
    time_allocation_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(B, N, start, end, seeds, max_vel, max_acc, durations);
    CK(cudaGetLastError());
    return 0;
}

int direct_ddp_sample_device(direct_ddp_handle h, int B, int N, int S, const double *bez_coeff, const double *poly_time,
                             double *pos, double *vel, double *acc, void *stream) {
    REQUIRE_DEVICE(h)
    if (B <= 0 || N <= 0 || S <= 0 || !bez_coeff || !poly_time) { h->err = "bad argument"; return DIRECT_DDP_ERR_ARG; }
    CK(cudaSetDevice(h->opts.device));
    const long long n = (long long)B * N * S;
    bezier_sample_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>((long long)B * N, S, bez_coeff, poly_time, pos, vel, acc);
    CK(cudaGetLastError());
    return 0;
}

int direct_ddp_sample(direct_ddp_handle h, int B, int N, int S, const double *bez_coeff, const double *poly_time,
                      double *pos, double *vel, double *acc) {
    REQUIRE_DEVICE(h)
    if (B <= 0 || N <= 0 || S <= 0 || !bez_coeff || !poly_time) { h->err = "bad argument"; return DIRECT_DDP_ERR_ARG; }
    CK(cudaSetDevice(h->opts.device));
    const size_t nb = (size_t)B * N * 18 * 8, nt = (size_t)B * N * 8, no = (size_t)B * N * S * 3 * 8;
    int st;
    if ((st = ensure(h, h->o_bz[0], nb))) return st;
    if ((st = ensure(h, h->o_pt[0], nt))) return st;
    if ((st = ensure(h, h->o_pc[0], no * 3))) return st;
    double *d_bez = (double *)h->o_bz[0].p, *d_t = (double *)h->o_pt[0].p, *d_o = (double *)h->o_pc[0].p;
    CK(cudaMemcpyAsync(d_bez, bez_coeff, nb, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(d_t, poly_time, nt, cudaMemcpyHostToDevice, h->stream));
    const size_t ne = (size_t)B * N * S * 3;
    if ((st = direct_ddp_sample_device(h, B, N, S, d_bez, d_t, pos ? d_o : nullptr, vel ? d_o + ne : nullptr, acc ? d_o + 2 * ne : nullptr, h->stream))) return st;
    if (pos) CK(cudaMemcpyAsync(pos, d_o, no, cudaMemcpyDeviceToHost, h->stream));
    if (vel) CK(cudaMemcpyAsync(vel, d_o + ne, no, cudaMemcpyDeviceToHost, h->stream));
    if (acc) CK(cudaMemcpyAsync(acc, d_o + 2 * ne, no, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

int direct_ddp_measure_fma_peak(direct_ddp_handle h, int precision, double *tflops) {
    REQUIRE_DEVICE(h)
    if (!tflops) return DIRECT_DDP_ERR_ARG;
    CK(cudaSetDevice(h->opts.device));
    int st = ensure(h, h->scratch_i, 64);
    if (st) return st;
    const int iters = 4096, blocks = h->sm_count * 8, threads = 512;
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        CK(cudaEventRecord(h->ev[0], h->stream));
        if (precision == DIRECT_DDP_FP32) fma_peak_kernel<float><<<blocks, threads, 0, h->stream>>>((float *)h->scratch_i.p, iters, 0.999f, 0.001f);
        else fma_peak_kernel<double><<<blocks, threads, 0, h->stream>>>((double *)h->scratch_i.p, iters, 0.999, 0.001);
        CK(cudaGetLastError());
        CK(cudaEventRecord(h->ev[1], h->stream));
        CK(cudaStreamSynchronize(h->stream));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]));
        if (ms < best) best = ms;
    }
    *tflops = 2.0 * 8.0 * (double)iters * blocks * threads / (best * 1e-3) / 1e12;
    return 0;
}

int direct_ddp_sm_clock_hz(direct_ddp_handle h, double *hz) {
    REQUIRE_DEVICE(h)
    if (!hz) return DIRECT_DDP_ERR_ARG;
    int khz = 0;
    CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, h->opts.device));
    *hz = 1000.0 * khz;
    return DIRECT_DDP_OK;
}

static int last_stats_one(direct_ddp_handle h, direct_ddp_stats *out);
int direct_ddp_last_stats(direct_ddp_handle h, direct_ddp_stats *out) {
    REQUIRE_DEVICE(h)
    if (!out) return DIRECT_DDP_ERR_ARG;
    int st = last_stats_one(h, out);
    if (st || !h->last_multi) return st;
    for (direct_ddp_handle_s *p : h->peers) {
        direct_ddp_stats q;
        if ((st = last_stats_one(p, &q))) { h->err = p->err; return st; }
        out->kernel_ms = q.kernel_ms > out->kernel_ms ? q.kernel_ms : out->kernel_ms;
        out->h2d_ms = q.h2d_ms > out->h2d_ms ? q.h2d_ms : out->h2d_ms;
        out->d2h_ms = q.d2h_ms > out->d2h_ms ? q.d2h_ms : out->d2h_ms;
        out->bwd_sweeps += q.bwd_sweeps; out->bwd_knots += q.bwd_knots; out->fwd_trials += q.fwd_trials; out->fwd_knots += q.fwd_knots;
        out->kernel_launches += q.kernel_launches; out->grid_blocks += q.grid_blocks; out->workspace_slots += q.workspace_slots;
        out->h2d_bytes += q.h2d_bytes; out->d2h_bytes += q.d2h_bytes; out->coop_jobs += q.coop_jobs; out->helper_units += q.helper_units;
        out->spec_searches += q.spec_searches; out->spec_trials += q.spec_trials;
        out->spec_sweeps += q.spec_sweeps; out->spec_sweeps_used += q.spec_sweeps_used;
    }
    return 0;
}
static int last_stats_one(direct_ddp_handle h, direct_ddp_stats *out) {
    REQUIRE_DEVICE(h)
    if (!out) return DIRECT_DDP_ERR_ARG;
    if (!h->stats_valid && h->last_stats_dev && h->last_B > 0) {
        CK(cudaSetDevice(h->opts.device));
        CK(cudaDeviceSynchronize());
        float ms = 0;
        if (cudaEventElapsedTime(&ms, h->ev[2], h->ev[3]) == cudaSuccess) h->stats.kernel_ms = ms;
        std::vector<long long> tmp((size_t)h->last_B * 8);
        long long tot[4] = {0, 0, 0, 0};
        for (int pass = 0; pass < 2; pass++) {
            const long long *src = pass == 0 ? h->last_stats_dev : h->last_stats_dev0;
            if (!src) continue;
            CK(cudaMemcpy(tmp.data(), src, tmp.size() * sizeof(long long), cudaMemcpyDeviceToHost));
            for (int i = 0; i < h->last_B; i++) for (int k = 0; k < 4; k++) tot[k] += tmp[(size_t)i * 8 + k];
        }
        h->stats.bwd_sweeps = tot[0]; h->stats.bwd_knots = tot[1]; h->stats.fwd_trials = tot[2]; h->stats.fwd_knots = tot[3];
        unsigned int cnt[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        CK(cudaMemcpy(cnt, h->counter.p, sizeof cnt, cudaMemcpyDeviceToHost));
        h->stats.coop_jobs = cnt[1]; h->stats.helper_units = cnt[2];
        h->stats.spec_searches = cnt[5]; h->stats.spec_trials = cnt[6];
        h->stats.spec_sweeps = cnt[7]; h->stats.spec_sweeps_used = cnt[8];
        h->stats_valid = true;
    }
    *out = h->stats;
    return 0;
}

int direct_ddp_last_trace(direct_ddp_handle h, direct_ddp_trace_row *rows, int cap, int *len) {
    REQUIRE_DEVICE(h)
    if (!rows || !len) return DIRECT_DDP_ERR_ARG;
    *len = 0;
    if (!h->opts.trace || !h->trace.p) return 0;
    CK(cudaSetDevice(h->opts.device));
    CK(cudaDeviceSynchronize());
    int n = 0;
    CK(cudaMemcpy(&n, h->trace_len.p, sizeof(int), cudaMemcpyDeviceToHost));
    if (n > cap) n = cap;
    std::vector<double> t((size_t)n * 12 + 1);
    if (n > 0) CK(cudaMemcpy(t.data(), h->trace.p, (size_t)n * 12 * sizeof(double), cudaMemcpyDeviceToHost));
    int khz = 0;
    CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, h->opts.device));
    const double cyc_per_us = khz > 0 ? khz / 1000.0 : 1965.0;
    for (int i = 0; i < n; i++) {
        const double *r = &t[(size_t)i * 12];
        rows[i].cost = r[0]; rows[i].costq = r[1]; rows[i].logcost = r[2]; rows[i].err = r[3]; rows[i].mu = r[4];
        rows[i].reg = r[5]; rows[i].stepsize = r[6]; rows[i].opterr = r[7];
        rows[i].step = (int)r[8]; rows[i].fp_failed = (int)r[9]; rows[i].n_bwd = (int)r[10]; rows[i].t_us = (int)(r[11] / cyc_per_us);
    }
    *len = n;
    return 0;
}

}  // extern "C"


// ---- generic unconstrained DDP (include/direct_gddp.h, gddp.cuh) ---------------------------------------------------------
namespace {

template <int MODEL, class R> int gddp_launch(H *h, const direct_gddp_problem *in, const direct_gddp_result *out, cudaStream_t s) {
    using D = gddp::Dim<MODEL>;
    gddp::Args<R> A;
    memset(&A, 0, sizeof A);
    A.B = in->B; A.N = in->N; A.iter_max = in->iter_max; A.dt = (R)in->dt; A.tol = (R)in->tol;
    A.x0 = in->x0; A.xg = in->xg; A.u_init = in->u_init;
    for (int a = 0; a < D::NX; a++) { A.q[a] = (R)in->q[a]; A.qf[a] = (R)in->qf[a]; }
    for (int m = 0; m < D::NU; m++) { A.r[m] = (R)in->r[m]; A.uh[m] = (R)in->uh[m]; }
    A.rtn = out->rtn; A.iters = out->iters; A.cost = out->cost; A.x = out->x; A.u = out->u; A.stats = (long long *)out->stats;
    // two trajectories per warp (gddp_pair.cuh) unless DIRECT_GDDP_PAIR=0 asks for the one-per-warp kernel (tuning / comparison)
    const char *pe = getenv("DIRECT_GDDP_PAIR");
    const int tpw = (pe && atoi(pe) == 0) ? 1 : 2;
    const int threads = 128, wpb = threads / 32;
    const size_t smem = (size_t)wpb * (tpw == 2 ? 2 * gddp::SmemP<MODEL>::TOTAL : gddp::Smem<MODEL>::TOTAL) * sizeof(R);
    auto kern = tpw == 2 ? gddp::gddp_pair_kernel<MODEL, R> : gddp::gddp_kernel<MODEL, R>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem));
    if (per_sm < 1) { h->err = "gddp kernel does not fit on an SM"; return DIRECT_DDP_ERR_CUDA; }
    long long grid = (long long)per_sm * h->sm_count;
    const long long need = ((long long)in->B + wpb * tpw - 1) / (wpb * tpw);
    if (grid > need) grid = need;
    const gddp::Ws<MODEL, R> wl(in->N);
    A.ws_stride = wl.total;
    int st = ensure(h, h->ws, (size_t)grid * wpb * tpw * wl.total * sizeof(R));
    if (st) return st;
    A.ws = (R *)h->ws.p;
    if ((st = ensure(h, h->counter, 64))) return st;
    CK(cudaMemsetAsync(h->counter.p, 0, 64, s));
    A.counter = (unsigned int *)h->counter.p;
    CK(cudaEventRecord(h->ev[2], s));
    kern<<<(unsigned)grid, threads, smem, s>>>(A);
    CK(cudaGetLastError());
    CK(cudaEventRecord(h->ev[3], s));
    h->stats.kernel_launches = 1;
    h->stats.grid_blocks = (int)grid; h->stats.block_threads = threads;
    h->stats.smem_bytes_per_block = (int)smem; h->stats.workspace_slots = (int)(grid * wpb * tpw);
    h->last_stats_dev = nullptr; h->last_stats_dev0 = nullptr; h->last_B = 0; h->stats_valid = false;
    return 0;
}

int gddp_validate(H *h, const direct_gddp_problem *in, const direct_gddp_result *out) {
    if (!in || !out) { h->err = "problem or result is NULL"; return DIRECT_DDP_ERR_ARG; }
    if (in->model != DIRECT_GDDP_DINT6 && in->model != DIRECT_GDDP_QUAD12) { h->err = "unknown model"; return DIRECT_DDP_ERR_ARG; }
    if (in->B <= 0 || in->N <= 0 || in->iter_max < 0 || !(in->dt > 0.0)) { h->err = "B, N, dt must be positive"; return DIRECT_DDP_ERR_ARG; }
    if (!in->x0 || !in->xg || !out->rtn || !out->iters || !out->cost || !out->x || !out->u) { h->err = "missing pointer"; return DIRECT_DDP_ERR_ARG; }
    return 0;
}

int gddp_dispatch(H *h, const direct_gddp_problem *in, const direct_gddp_result *out, cudaStream_t s) {
    const bool f32 = h->opts.precision == DIRECT_DDP_FP32;
    if (in->model == DIRECT_GDDP_QUAD12) return f32 ? gddp_launch<1, float>(h, in, out, s) : gddp_launch<1, double>(h, in, out, s);
    return f32 ? gddp_launch<0, float>(h, in, out, s) : gddp_launch<0, double>(h, in, out, s);
}

}  // namespace

extern "C" int direct_gddp_solve_device(direct_ddp_handle h, const direct_gddp_problem *in, direct_gddp_result *out, void *stream) {
    REQUIRE_DEVICE(h)
    int st = gddp_validate(h, in, out);
    if (st) return st;
    CK(cudaSetDevice(h->opts.device));
    return gddp_dispatch(h, in, out, (cudaStream_t)stream);   // NULL = the default stream, as in direct_ddp_solve_batch_device
}

extern "C" int direct_gddp_solve(direct_ddp_handle h, const direct_gddp_problem *in, direct_gddp_result *out) {
    REQUIRE_DEVICE(h)
    int st = gddp_validate(h, in, out);
    if (st) return st;
    CK(cudaSetDevice(h->opts.device));
    cudaStream_t s = h->stream;
    const int nx = in->model == DIRECT_GDDP_QUAD12 ? 12 : 6, nu = in->model == DIRECT_GDDP_QUAD12 ? 4 : 3;
    const size_t B = (size_t)in->B, N = (size_t)in->N;
    direct_gddp_problem d = *in;
    direct_gddp_result r;
    memset(&r, 0, sizeof r);
    int64_t h2d = 0, d2h = 0;
    CK(cudaEventRecord(h->ev[0], s));
    const size_t nb_x0 = B * nx * 8, nb_u = B * N * nu * 8, nb_x = B * (N + 1) * nx * 8;
    if ((st = ensure(h, h->gd[0], nb_x0)) || (st = ensure(h, h->gd[1], nb_x0))) return st;
    CK(cudaMemcpyAsync(h->gd[0].p, in->x0, nb_x0, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(h->gd[1].p, in->xg, nb_x0, cudaMemcpyHostToDevice, s));
    d.x0 = (const double *)h->gd[0].p; d.xg = (const double *)h->gd[1].p;
    h2d += 2 * (int64_t)nb_x0;
    if (in->u_init) {
        if ((st = ensure(h, h->gd[2], nb_u))) return st;
        CK(cudaMemcpyAsync(h->gd[2].p, in->u_init, nb_u, cudaMemcpyHostToDevice, s));
        d.u_init = (const double *)h->gd[2].p;
        h2d += (int64_t)nb_u;
    }
    if ((st = ensure(h, h->gd[3], B * 8)) || (st = ensure(h, h->gd[4], B * 8)) || (st = ensure(h, h->gd[5], nb_x)) ||
        (st = ensure(h, h->gd[6], nb_u)) || (st = ensure(h, h->gd[7], B * 32))) return st;
    r.rtn = (int32_t *)h->gd[3].p; r.iters = (int32_t *)h->gd[3].p + B; r.cost = (double *)h->gd[4].p;
    r.x = (double *)h->gd[5].p; r.u = (double *)h->gd[6].p; r.stats = (int64_t *)h->gd[7].p;
    CK(cudaEventRecord(h->ev[1], s));
    if ((st = gddp_dispatch(h, &d, &r, s))) return st;
    CK(cudaEventRecord(h->ev[4], s));
    CK(cudaMemcpyAsync(out->rtn, r.rtn, B * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(out->iters, r.iters, B * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(out->cost, r.cost, B * 8, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(out->x, r.x, nb_x, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(out->u, r.u, nb_u, cudaMemcpyDeviceToHost, s));
    d2h += (int64_t)(B * 16 + nb_x + nb_u);
    if (out->stats) { CK(cudaMemcpyAsync(out->stats, r.stats, B * 32, cudaMemcpyDeviceToHost, s)); d2h += (int64_t)B * 32; }
    CK(cudaEventRecord(h->ev[5], s));
    CK(cudaStreamSynchronize(s));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, h->ev[0], h->ev[1])); h->stats.h2d_ms = ms;
    CK(cudaEventElapsedTime(&ms, h->ev[2], h->ev[3])); h->stats.kernel_ms = ms;
    CK(cudaEventElapsedTime(&ms, h->ev[4], h->ev[5])); h->stats.d2h_ms = ms;
    h->stats.h2d_bytes = h2d; h->stats.d2h_bytes = d2h;
    return DIRECT_DDP_OK;
}


// ---- voxel-map kernels (include/direct_voxel.h, voxel.cuh) ----------------------------------------------------------------
namespace {
int voxel_rcp_table(H *h, cudaStream_t s) {
    if (h->vx[18].p) return 0;
    int st = ensure(h, h->vx[18], sizeof(double) * voxel::RCP_N);
    if (st) return st;
    voxel::rcp_table_kernel<<<voxel::RCP_N / 256, 256, 0, s>>>((double *)h->vx[18].p);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(s));   // later launches may come on other streams
    return 0;
}

int voxel_validate(H *h, const direct_voxel_map *m) {
    if (!m || !m->occupied || m->nx <= 0 || m->ny <= 0 || m->nz <= 0) { h->err = "bad voxel map"; return DIRECT_DDP_ERR_ARG; }
    if ((long long)m->nx * m->ny * m->nz > 0x7fffffffLL) { h->err = "voxel map too large for 32-bit cell indices"; return DIRECT_DDP_ERR_ARG; }
    return 0;
}
}  // namespace

extern "C" int direct_voxel_convex_test_device(direct_ddp_handle h, const direct_voxel_map *map, const int32_t *cand, int C,
                                               const int32_t *clu, int K, uint8_t *can_can, uint8_t *can_clu, void *stream) {
    REQUIRE_DEVICE(h)
    int st = voxel_validate(h, map);
    if (st) return st;
    if (!map->inside || C < 0 || K < 0 || (C > 0 && (!cand || !can_can || !can_clu)) || (K > 0 && !clu)) { h->err = "missing pointer"; return DIRECT_DDP_ERR_ARG; }
    CK(cudaSetDevice(h->opts.device));
    cudaStream_t s = (cudaStream_t)stream;
    CK(cudaEventRecord(h->ev[2], s));
    if (C > 0) {
        const int threads = 256;
        const long long tasks = (long long)C * ((C - 1 + K + 31) / 32);
        long long grid = (tasks + 7) / 8;
        const long long cap = (long long)h->sm_count * 8;   // 2048 threads per SM: a persistent grid, warps stride over the tasks
        if (grid > cap) grid = cap;
        if (grid < 1) grid = 1;
        const size_t cells = (size_t)map->nx * map->ny * map->nz;
        if ((st = ensure(h, h->vx[16], cells)) || (st = voxel_rcp_table(h, s))) return st;   // merged map (bit 0 occupied, bit 1 inside)
        voxel::merge_map_kernel<<<h->sm_count * 4, 256, 0, s>>>(map->occupied, map->inside, (uint8_t *)h->vx[16].p, cells, can_clu, C);
        voxel::convex_test_kernel<<<(unsigned)grid, threads, 0, s>>>((const uint8_t *)h->vx[16].p, (const double *)h->vx[18].p, map->ny * map->nz, map->nz, cand, C, clu, K,
                                                                     can_can, can_clu);
        CK(cudaGetLastError());
    }
    CK(cudaEventRecord(h->ev[3], s));
    h->stats.kernel_launches = C > 0 ? 2 : 0;
    h->last_stats_dev = nullptr; h->last_stats_dev0 = nullptr; h->last_B = 0; h->stats_valid = false;
    return DIRECT_DDP_OK;
}

extern "C" int direct_voxel_cube_inflation_device(direct_ddp_handle h, const direct_voxel_map *map, const int32_t *vertex_idx, int dir,
                                                  int inf_step, int32_t *result, void *stream) {
    REQUIRE_DEVICE(h)
    int st = voxel_validate(h, map);
    if (st) return st;
    if (!vertex_idx || !result || dir < 0 || dir > 5 || inf_step < 0) { h->err = "bad inflation argument"; return DIRECT_DDP_ERR_ARG; }
    CK(cudaSetDevice(h->opts.device));
    cudaStream_t s = (cudaStream_t)stream;
    const int32_t one = 1;
    CK(cudaMemcpyAsync(result, &one, sizeof one, cudaMemcpyHostToDevice, s));
    CK(cudaEventRecord(h->ev[2], s));
    // the face of a box inside the map has at most max(nx ny, nx nz, ny nz) cells; blocks stride over it
    voxel::cube_inflation_kernel<<<h->sm_count, 256, 0, s>>>(map->occupied, map->ny * map->nz, map->nz, vertex_idx, dir, inf_step, result);
    CK(cudaGetLastError());
    CK(cudaEventRecord(h->ev[3], s));
    h->stats.kernel_launches = 1;
    h->last_stats_dev = nullptr; h->last_stats_dev0 = nullptr; h->last_B = 0; h->stats_valid = false;
    return DIRECT_DDP_OK;
}

namespace {
int voxel_upload_map(H *h, const direct_voxel_map *map, direct_voxel_map *d, cudaStream_t s, bool need_inside) {
    const size_t cells = (size_t)map->nx * map->ny * map->nz;
    int st;
    if ((st = ensure(h, h->vx[0], cells))) return st;
    CK(cudaMemcpyAsync(h->vx[0].p, map->occupied, cells, cudaMemcpyHostToDevice, s));
    *d = *map;
    d->occupied = (const uint8_t *)h->vx[0].p;
    d->inside = nullptr;
    if (need_inside) {
        if ((st = ensure(h, h->vx[1], cells))) return st;
        CK(cudaMemcpyAsync(h->vx[1].p, map->inside, cells, cudaMemcpyHostToDevice, s));
        d->inside = (const uint8_t *)h->vx[1].p;
    }
    return 0;
}
}  // namespace

extern "C" int direct_voxel_convex_test(direct_ddp_handle h, const direct_voxel_map *map, const int32_t *cand, int C, const int32_t *clu,
                                        int K, uint8_t *can_can, uint8_t *can_clu) {
    REQUIRE_DEVICE(h)
    int st = voxel_validate(h, map);
    if (st) return st;
    if (!map->inside || C < 0 || K < 0 || (C > 0 && (!cand || !can_can || !can_clu)) || (K > 0 && !clu)) { h->err = "missing pointer"; return DIRECT_DDP_ERR_ARG; }
    CK(cudaSetDevice(h->opts.device));
    cudaStream_t s = h->stream;
    direct_voxel_map d;
    CK(cudaEventRecord(h->ev[0], s));
    if ((st = voxel_upload_map(h, map, &d, s, true))) return st;
    const size_t ncc = (size_t)C * ((size_t)C + 1) / 2;
    if ((st = ensure(h, h->vx[2], (size_t)C * 12)) || (st = ensure(h, h->vx[3], (size_t)K * 12)) || (st = ensure(h, h->vx[4], ncc)) ||
        (st = ensure(h, h->vx[5], (size_t)C))) return st;
    if (C > 0) CK(cudaMemcpyAsync(h->vx[2].p, cand, (size_t)C * 12, cudaMemcpyHostToDevice, s));
    if (K > 0) CK(cudaMemcpyAsync(h->vx[3].p, clu, (size_t)K * 12, cudaMemcpyHostToDevice, s));
    if (ncc > 0) CK(cudaMemcpyAsync(h->vx[4].p, can_can, ncc, cudaMemcpyHostToDevice, s));   // entries the kernel never writes keep the caller's bytes
    CK(cudaEventRecord(h->ev[1], s));
    if ((st = direct_voxel_convex_test_device(h, &d, (const int32_t *)h->vx[2].p, C, (const int32_t *)h->vx[3].p, K, (uint8_t *)h->vx[4].p,
                                              (uint8_t *)h->vx[5].p, s))) return st;
    CK(cudaEventRecord(h->ev[4], s));
    if (ncc > 0) CK(cudaMemcpyAsync(can_can, h->vx[4].p, ncc, cudaMemcpyDeviceToHost, s));
    if (C > 0) CK(cudaMemcpyAsync(can_clu, h->vx[5].p, (size_t)C, cudaMemcpyDeviceToHost, s));
    CK(cudaEventRecord(h->ev[5], s));
    CK(cudaStreamSynchronize(s));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, h->ev[0], h->ev[1])); h->stats.h2d_ms = ms;
    CK(cudaEventElapsedTime(&ms, h->ev[2], h->ev[3])); h->stats.kernel_ms = ms;
    CK(cudaEventElapsedTime(&ms, h->ev[4], h->ev[5])); h->stats.d2h_ms = ms;
    return DIRECT_DDP_OK;
}

extern "C" int direct_voxel_cube_inflation(direct_ddp_handle h, const direct_voxel_map *map, const int32_t *vertex_idx, int dir, int inf_step,
                                           int32_t *result) {
    REQUIRE_DEVICE(h)
    int st = voxel_validate(h, map);
    if (st) return st;
    if (!vertex_idx || !result) { h->err = "missing pointer"; return DIRECT_DDP_ERR_ARG; }
    CK(cudaSetDevice(h->opts.device));
    cudaStream_t s = h->stream;
    direct_voxel_map d;
    if ((st = voxel_upload_map(h, map, &d, s, false))) return st;
    if ((st = ensure(h, h->vx[6], 96)) || (st = ensure(h, h->vx[7], 16))) return st;
    CK(cudaMemcpyAsync(h->vx[6].p, vertex_idx, 96, cudaMemcpyHostToDevice, s));
    if ((st = direct_voxel_cube_inflation_device(h, &d, (const int32_t *)h->vx[6].p, dir, inf_step, (int32_t *)h->vx[7].p, s))) return st;
    CK(cudaMemcpyAsync(result, h->vx[7].p, 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, h->ev[2], h->ev[3])); h->stats.kernel_ms = ms;
    return DIRECT_DDP_OK;
}

extern "C" int direct_voxel_inflate_box_device(direct_ddp_handle h, const direct_voxel_map *map, int32_t *vertex_idx, int inf_step,
                                               int itr_inflate_max, int32_t *iters, void *stream) {
    REQUIRE_DEVICE(h)
    int st = voxel_validate(h, map);
    if (st) return st;
    if (!vertex_idx || inf_step != 1 || itr_inflate_max < 0) { h->err = "bad inflation argument (inf_step must be 1)"; return DIRECT_DDP_ERR_ARG; }
    CK(cudaSetDevice(h->opts.device));
    cudaStream_t s = (cudaStream_t)stream;
    CK(cudaEventRecord(h->ev[2], s));
    voxel::inflate_box_kernel<<<1, 1024, 0, s>>>(map->occupied, map->nx, map->ny, map->nz, vertex_idx, inf_step, itr_inflate_max, iters);
    CK(cudaGetLastError());
    CK(cudaEventRecord(h->ev[3], s));
    h->stats.kernel_launches = 1;
    h->last_stats_dev = nullptr; h->last_stats_dev0 = nullptr; h->last_B = 0; h->stats_valid = false;
    return DIRECT_DDP_OK;
}

extern "C" int direct_voxel_inflate_box(direct_ddp_handle h, const direct_voxel_map *map, int32_t *vertex_idx, int inf_step,
                                        int itr_inflate_max, int32_t *iters) {
    REQUIRE_DEVICE(h)
    int st = voxel_validate(h, map);
    if (st) return st;
    if (!vertex_idx) { h->err = "missing pointer"; return DIRECT_DDP_ERR_ARG; }
    CK(cudaSetDevice(h->opts.device));
    cudaStream_t s = h->stream;
    direct_voxel_map d;
    if ((st = voxel_upload_map(h, map, &d, s, false))) return st;
    if ((st = ensure(h, h->vx[6], 96)) || (st = ensure(h, h->vx[7], 16))) return st;
    CK(cudaMemcpyAsync(h->vx[6].p, vertex_idx, 96, cudaMemcpyHostToDevice, s));
    if ((st = direct_voxel_inflate_box_device(h, &d, (int32_t *)h->vx[6].p, inf_step, itr_inflate_max, (int32_t *)h->vx[7].p, s))) return st;
    int32_t it = 0;
    CK(cudaMemcpyAsync(vertex_idx, h->vx[6].p, 96, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(&it, h->vx[7].p, 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (iters) *iters = it;
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, h->ev[2], h->ev[3])); h->stats.kernel_ms = ms;
    return DIRECT_DDP_OK;
}

extern "C" int direct_voxel_cluster_device(direct_ddp_handle h, const direct_voxel_map *map, uint8_t *use, uint8_t *invalid,
                                           int32_t *cluster_xyz, int32_t *ctl, int cap, int cand_cap, int itr_cluster_max, void *stream) {
    REQUIRE_DEVICE(h)
    int st = voxel_validate(h, map);
    if (st) return st;
    if (!map->inside || !use || !invalid || !cluster_xyz || !ctl || cap <= 0 || cand_cap <= 0 || cand_cap > 32 * voxel::ACC_WORDS ||
        itr_cluster_max < 0) { h->err = "bad clustering argument (cand_cap <= 32768)"; return DIRECT_DDP_ERR_ARG; }
    CK(cudaSetDevice(h->opts.device));
    cudaStream_t s = (cudaStream_t)stream;
    const size_t cells = (size_t)map->nx * map->ny * map->nz;
    const bool fresh_claims = h->vx[8].cap < cells * 4;
    const size_t conflict_bytes = (size_t)cand_cap * ((cand_cap + 31) / 32) * 4;
    const bool fresh_conflict = h->vx[10].cap < conflict_bytes;
    if ((st = ensure(h, h->vx[8], cells * 4)) || (st = ensure(h, h->vx[9], (size_t)cand_cap * 12)) ||
        (st = ensure(h, h->vx[10], (size_t)cand_cap * ((cand_cap + 31) / 32) * 4)) || (st = ensure(h, h->vx[11], (size_t)cand_cap))) return st;
    // the kernel hands every claim word back empty; only a new (or regrown) array needs the fill
    if (fresh_claims) CK(cudaMemsetAsync(h->vx[8].p, 0x7f, h->vx[8].cap, s));
    if (fresh_conflict) CK(cudaMemsetAsync(h->vx[10].p, 0, h->vx[10].cap, s));   // rows of rejected candidates are read (and ignored)
    if (h->vx_coop_blocks == 0) {
        int per_sm = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, voxel::cluster_loop_kernel, 256, 0));
        if (per_sm < 1) { h->err = "cluster_loop_kernel does not fit an SM"; return DIRECT_DDP_ERR_CUDA; }
        h->vx_coop_blocks = per_sm * h->sm_count;
    }
    const uint8_t *occ = map->occupied, *inside = map->inside;
    int *claim = (int *)h->vx[8].p, *cand = (int *)h->vx[9].p;
    unsigned *conflict = (unsigned *)h->vx[10].p;
    uint8_t *can_clu = (uint8_t *)h->vx[11].p;
    int nx = map->nx, ny = map->ny, nz = map->nz;
    voxel::ClusterCtl *c = (voxel::ClusterCtl *)ctl;
    if ((st = ensure(h, h->vx[15], 64))) return st;
    unsigned long long *phase_ns = (unsigned long long *)h->vx[15].p;
    CK(cudaMemsetAsync(phase_ns, 0, 64, s));
    if ((st = ensure(h, h->vx[16], cells)) || (st = ensure(h, h->vx[17], ((size_t)26 * (size_t)(cap > cand_cap ? cap : cand_cap) / 128 + 2) * 4))) return st;
    if ((st = voxel_rcp_table(h, s))) return st;
    int *seg_count = (int *)h->vx[17].p;
    const double *rcp = (const double *)h->vx[18].p;
    const uint8_t *merged = (const uint8_t *)h->vx[16].p;
    CK(cudaEventRecord(h->ev[2], s));
    voxel::merge_map_kernel<<<h->sm_count * 4, 256, 0, s>>>(occ, inside, (uint8_t *)h->vx[16].p, cells, nullptr, 0);
    void *args[] = {&occ, &inside, &merged, &rcp, &use, &invalid, &claim, &nx, &ny, &nz, &cluster_xyz, &cap, &cand, &cand_cap, &conflict, &can_clu,
                    &seg_count, &itr_cluster_max, &c, &phase_ns};
    CK(cudaLaunchCooperativeKernel((const void *)voxel::cluster_loop_kernel, dim3(h->vx_coop_blocks), dim3(256), args, 0, s));
    CK(cudaEventRecord(h->ev[3], s));
    h->stats.kernel_launches = 2;
    h->last_stats_dev = nullptr; h->last_stats_dev0 = nullptr; h->last_B = 0; h->stats_valid = false;
    return DIRECT_DDP_OK;
}

extern "C" int direct_voxel_cluster(direct_ddp_handle h, const direct_voxel_map *map, uint8_t *use, uint8_t *invalid, int32_t *cluster_xyz,
                                    int32_t *cluster_num, int cap, int cand_cap, int itr_cluster_max, int32_t *iters) {
    REQUIRE_DEVICE(h)
    int st = voxel_validate(h, map);
    if (st) return st;
    if (!map->inside || !use || !invalid || !cluster_xyz || !cluster_num || *cluster_num < 0 || *cluster_num > cap) {
        h->err = "bad clustering argument"; return DIRECT_DDP_ERR_ARG;
    }
    CK(cudaSetDevice(h->opts.device));
    cudaStream_t s = h->stream;
    const size_t cells = (size_t)map->nx * map->ny * map->nz;
    direct_voxel_map d;
    CK(cudaEventRecord(h->ev[0], s));
    if ((st = voxel_upload_map(h, map, &d, s, true))) return st;
    if ((st = ensure(h, h->vx[12], 32)) || (st = ensure(h, h->vx[13], cells)) || (st = ensure(h, h->vx[14], cells)) ||
        (st = ensure(h, h->vx[3], (size_t)cap * 12))) return st;
    int32_t ctl[8] = {*cluster_num, 0, 0, 0, 0, 0, 0, 0};
    CK(cudaMemcpyAsync(h->vx[12].p, ctl, 32, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(h->vx[13].p, use, cells, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(h->vx[14].p, invalid, cells, cudaMemcpyHostToDevice, s));
    if (*cluster_num > 0) CK(cudaMemcpyAsync(h->vx[3].p, cluster_xyz, (size_t)*cluster_num * 12, cudaMemcpyHostToDevice, s));
    CK(cudaEventRecord(h->ev[1], s));
    if ((st = direct_voxel_cluster_device(h, &d, (uint8_t *)h->vx[13].p, (uint8_t *)h->vx[14].p, (int32_t *)h->vx[3].p, (int32_t *)h->vx[12].p,
                                          cap, cand_cap, itr_cluster_max, s))) return st;
    CK(cudaEventRecord(h->ev[4], s));
    CK(cudaMemcpyAsync(ctl, h->vx[12].p, 32, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (ctl[2] != 0) { h->err = "cluster or candidate capacity exceeded"; return DIRECT_DDP_ERR_ARG; }
    CK(cudaMemcpyAsync(use, h->vx[13].p, cells, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(invalid, h->vx[14].p, cells, cudaMemcpyDeviceToHost, s));
    if (ctl[0] > 0) CK(cudaMemcpyAsync(cluster_xyz, h->vx[3].p, (size_t)ctl[0] * 12, cudaMemcpyDeviceToHost, s));
    CK(cudaEventRecord(h->ev[5], s));
    CK(cudaStreamSynchronize(s));
    *cluster_num = ctl[0];
    if (iters) *iters = ctl[1];
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, h->ev[0], h->ev[1])); h->stats.h2d_ms = ms;
    CK(cudaEventElapsedTime(&ms, h->ev[2], h->ev[3])); h->stats.kernel_ms = ms;
    CK(cudaEventElapsedTime(&ms, h->ev[4], h->ev[5])); h->stats.d2h_ms = ms;
    return DIRECT_DDP_OK;
}

// Device time of the phases of the last direct_voxel_cluster[_device] launch on this handle, summed over its iterations (ms):
// claims, ordered compaction, cluster rays, candidate rays, acceptance scan.  Synchronises the handle's device.
extern "C" int direct_voxel_cluster_phases(direct_ddp_handle h, double ms[5]) {
    REQUIRE_DEVICE(h)
    if (!ms || !h->vx[15].p) { h->err = "no clustering launch on this handle yet"; return DIRECT_DDP_ERR_ARG; }
    CK(cudaSetDevice(h->opts.device));
    unsigned long long ns[8];
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(ns, h->vx[15].p, 64, cudaMemcpyDeviceToHost));
    for (int k = 0; k < 5; k++) ms[k] = (double)ns[k] * 1e-6;
    return DIRECT_DDP_OK;
}

extern "C" int direct_voxel_polytope(direct_ddp_handle h, const direct_voxel_map *map, const int32_t seed_xyz[3], int itr_inflate_max,
                                     int itr_cluster_max, int cap, int cand_cap, int32_t *cluster_xyz, int32_t *cluster_num, int32_t *iters,
                                     int32_t *vertex_idx, uint8_t *inside, uint8_t *use, uint8_t *invalid) {
    REQUIRE_DEVICE(h)
    int st = voxel_validate(h, map);
    if (st) return st;
    if (!seed_xyz || !cluster_xyz || !cluster_num || cap <= 0 || seed_xyz[0] < 0 || seed_xyz[0] >= map->nx || seed_xyz[1] < 0 ||
        seed_xyz[1] >= map->ny || seed_xyz[2] < 0 || seed_xyz[2] >= map->nz) { h->err = "bad polytope argument (seed outside the map?)"; return DIRECT_DDP_ERR_ARG; }
    CK(cudaSetDevice(h->opts.device));
    cudaStream_t s = h->stream;
    const size_t cells = (size_t)map->nx * map->ny * map->nz;
    direct_voxel_map d;
    CK(cudaEventRecord(h->ev[0], s));
    if ((st = voxel_upload_map(h, map, &d, s, false))) return st;
    if ((st = ensure(h, h->vx[1], cells)) || (st = ensure(h, h->vx[13], cells)) || (st = ensure(h, h->vx[14], cells)) ||
        (st = ensure(h, h->vx[3], (size_t)cap * 12)) || (st = ensure(h, h->vx[6], 96)) || (st = ensure(h, h->vx[7], 16)) ||
        (st = ensure(h, h->vx[12], 32))) return st;
    d.inside = (const uint8_t *)h->vx[1].p;
    int32_t v[24], ctl[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int k = 0; k < 8; k++) { v[k] = seed_xyz[0]; v[8 + k] = seed_xyz[1]; v[16 + k] = seed_xyz[2]; }   // cluster_server.cu:793-800
    CK(cudaMemcpyAsync(h->vx[6].p, v, 96, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(h->vx[12].p, ctl, 32, cudaMemcpyHostToDevice, s));
    CK(cudaEventRecord(h->ev[1], s));
    CK(cudaMemsetAsync(h->vx[1].p, 0, cells, s));    // flagClear
    CK(cudaMemsetAsync(h->vx[13].p, 0, cells, s));
    CK(cudaMemsetAsync(h->vx[14].p, 0, cells, s));
    if ((st = direct_voxel_inflate_box_device(h, &d, (int32_t *)h->vx[6].p, 1, itr_inflate_max, (int32_t *)h->vx[7].p, s))) return st;
    voxel::cube_shell_kernel<<<h->sm_count * 2, 256, 0, s>>>((const int *)h->vx[6].p, map->ny, map->nz, (uint8_t *)h->vx[1].p, (uint8_t *)h->vx[13].p,
                                                            (int *)h->vx[3].p, cap, (voxel::ClusterCtl *)h->vx[12].p);
    CK(cudaGetLastError());
    if (cand_cap > 0 && itr_cluster_max >= 0) {
        if ((st = direct_voxel_cluster_device(h, &d, (uint8_t *)h->vx[13].p, (uint8_t *)h->vx[14].p, (int32_t *)h->vx[3].p, (int32_t *)h->vx[12].p,
                                              cap, cand_cap, itr_cluster_max, s))) return st;
    }
    CK(cudaMemcpyAsync(ctl, h->vx[12].p, 32, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(v, h->vx[6].p, 96, cudaMemcpyDeviceToHost, s));
    int32_t it_inf = 0;
    CK(cudaMemcpyAsync(&it_inf, h->vx[7].p, 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (ctl[2] != 0) { h->err = "cluster or candidate capacity exceeded"; return DIRECT_DDP_ERR_ARG; }
    CK(cudaMemcpyAsync(cluster_xyz, h->vx[3].p, (size_t)ctl[0] * 12, cudaMemcpyDeviceToHost, s));
    if (inside) CK(cudaMemcpyAsync(inside, h->vx[1].p, cells, cudaMemcpyDeviceToHost, s));
    if (use) CK(cudaMemcpyAsync(use, h->vx[13].p, cells, cudaMemcpyDeviceToHost, s));
    if (invalid) CK(cudaMemcpyAsync(invalid, h->vx[14].p, cells, cudaMemcpyDeviceToHost, s));
    CK(cudaEventRecord(h->ev[5], s));
    CK(cudaStreamSynchronize(s));
    *cluster_num = ctl[0];
    if (iters) { iters[0] = it_inf; iters[1] = ctl[1]; }
    if (vertex_idx) memcpy(vertex_idx, v, 96);
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, h->ev[0], h->ev[1])); h->stats.h2d_ms = ms;
    CK(cudaEventElapsedTime(&ms, h->ev[1], h->ev[3])); h->stats.kernel_ms = ms;   // memsets + inflation + shell + clustering
    h->stats.kernel_launches = 4;
    return DIRECT_DDP_OK;
}
