/*
 * direct_ddp.h -- C-ABI of libdirect_ddp_b200.so, the B200-native batched IPDDP trajectory optimiser.
 *
 * This is the drop-in boundary for ONE path of ntu-caokun/DIRECT: the call
 *
 *     int ddpTrajOptimizer::polyCurveGeneration(corridor, MQM_u, MQM_l, pos, vel, acc, jer, minimize_order,
 *             max_vel, max_acc, max_jer, initbezCoeff, w_snap, w_terminal, w_time, iter_max,
 *             bool& infeas, zero_init_flag, line_init_flag, bool& line_failed, time_power, minvo_flag)
 *
 * (global_planner/include/global_planner/ddp_optimizer.h:267-289, body ddp_optimizer.cpp:5-438) and the
 * getters next to it (ddp_optimizer.h:299-340).  direct_b200/host/ddp_optimizer_b200.cpp is a
 * replacement translation unit for the reference's src/ddp_optimizer.cpp that keeps that header unchanged
 * and forwards to the entry points below with a batch of one; INTEGRATION.md shows the CMake change.
 *
 * Conventions: plain C, caller owns every buffer, all floating-point buffers are IEEE double (the
 * reference's type, ddp_optimizer.h:15-16) whatever arithmetic the device path is asked to use.
 * Return value = library status (0 ok, <0 error; text via direct_ddp_last_error); the reference's own
 * per-trajectory return code lives in direct_ddp_result::rtn and is never mixed with it.
 * There is no CPU fallback: every solve entry point fails with DIRECT_DDP_ERR_CUDA when no usable
 * sm_100 device is present.
 *
 * Environment variables read at every solve (tuning and A/B measurements only; none of them changes a result bit):
 *   DIRECT_DDP_COOP=0   no cooperation between the warps of a CTA (and none of the two mechanisms below)
 *   DIRECT_DDP_GSPEC=0  idle CTAs do not run line-search trials of the remaining solves
 *   DIRECT_DDP_SPEC=0   idle CTAs do not run backward sweeps speculatively
 *   DIRECT_DDP_MIN_GRID=n   CTAs launched for a small batch (default: one per SM, so that a single corridor has idle CTAs)
 *   DIRECT_DDP_CARVEOUT=p   shared-memory carve-out hint (percent)
 */
#ifndef DIRECT_DDP_H_
#define DIRECT_DDP_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DIRECT_DDP_VERSION 200 /* 0.2.0: opts.ndevices / devices[], stats.spec_sweeps*, direct_ddp_device_count (round 2) */

enum {
    DIRECT_DDP_OK = 0,
    DIRECT_DDP_ERR_ARG = -1,     /* bad argument (NULL, N<=0, nplanes > P_max, time_power not in {1,2}, ...) */
    DIRECT_DDP_ERR_CUDA = -2,    /* CUDA runtime error / no device                               */
    DIRECT_DDP_ERR_NOMEM = -3,
    DIRECT_DDP_ERR_UNSUPPORTED = -4
};

enum { DIRECT_DDP_FP64 = 0, DIRECT_DDP_FP32 = 1 };

/* Planes per polytope: no algorithmic limit (the reference has none, m_c = 6 P + 55 rows per knot,
 * ddp_optimizer.cpp:145, :1181-1187); the row loops of the kernel walk a polytope's planes one after the other and the
 * per-slot workspace grows linearly with P_max.  The constant is a sanity bound on the argument only. */
#define DIRECT_DDP_MAX_PLANES 4096

typedef struct direct_ddp_opts {
    int device;          /* CUDA device ordinal                                                    */
    int precision;       /* DIRECT_DDP_FP64 (parity path) or DIRECT_DDP_FP32                        */
    int warps_per_block; /* 0 = default (4)                                                        */
    int blocks_per_sm;   /* 0 = as many as shared memory / registers allow                         */
    int trace;           /* !=0: keep a per-iteration trace of trajectory 0 (debug)                */
    /* Multi-GPU, single host process: with ndevices > 1 every HOST-buffer entry point
     * (direct_ddp_solve_batch, direct_ddp_solve_two_stage, direct_ddp_replay) shards its batch into ndevices
     * contiguous ranges [B k / n, B (k + 1) / n), one per device, each driven by its own host thread and stream:
     * H2D of the shard, solve, D2H of the shard straight into the caller's arrays.  Trajectories are independent, so
     * there is no collective on the data path (the host is the consumer of the results; nothing is gathered on a GPU).
     * The *_device entry points keep running on devices[0] (`device` when devices is NULL).  0 or 1 = single device. */
    int ndevices;
    const int *devices;  /* [ndevices] CUDA ordinals, or NULL = device, device + 1, ...             */
} direct_ddp_opts;

/* One batch of B independent problems, all with N polytopes (= knots = polynomial segments).
 * Mirrors decomp_cvx_space::FlightCorridor (utils/data_type.h:190-245) flattened:
 *   polyhedrons[i].planes[k] (Vector4d, a x + b y + c z + d <= 0 inside) -> planes[b][i][k][0..3]
 *   durations[i]                                                            -> durations[b][i]
 *   polyhedrons[i].seed_coord (only read when line_init != 0)               -> seeds[b][i][0..2]
 * and the scalar arguments of polyCurveGeneration.  Rows k >= nplanes[b][i] of planes are ignored.
 * With host entry points the pointers are host pointers; with the *_device entry points they are
 * device pointers on opts.device. */
typedef struct direct_ddp_batch {
    int B, N, P_max;
    const double *planes;    /* [B][N][P_max][4]; device entry points: 16-byte aligned (cudaMalloc is)  */
    const int32_t *nplanes;  /* [B][N], each in [0, P_max] (checked by the host entry points)          */
    const double *durations; /* [B][N]                                                              */
    const double *seeds;     /* [B][N][3] or NULL                                                   */
    const double *x0;        /* [B][9] = [pos.row(0), vel.row(0), acc.row(0)] (ddp_optimizer.cpp:115-121) */
    const double *xd;        /* [B][9] = [pos.row(1), vel.row(1), acc.row(1)] (ddp_optimizer.cpp:104-111) */
    const double *init_bez;  /* [B][N][18] initbezCoeff rows [x*6,y*6,z*6], or NULL (= zeros)       */
    const int32_t *infeas;   /* [B] value of `bool& infeas` on entry, or NULL (= infeas_all)        */
    int infeas_all;
    double max_vel, max_acc;
    double w_snap, w_terminal, w_time;
    int iter_max;
    int time_power;          /* 1 or 2 (anything else is undefined behaviour in the reference,
                                ddp_optimizer.cpp:1294-1305; rejected here)                         */
    int zero_init, line_init, minvo;
    const int32_t *nknots;   /* [B] knots of each trajectory, each in [1, N], or NULL (= N for all): ragged batches such
                                as the prefixes 2..64 of a recorded corridor (teach_repeat_planner.cpp:316-350).  All
                                per-knot arrays keep the stride N; entries past a trajectory's own count are ignored on
                                input and left untouched on output.                                              */
} direct_ddp_batch;

/* Outputs, one entry per trajectory; any pointer may be NULL to skip that output.
 *   rtn            polyCurveGeneration's return value: 0, 1, 2, -3, -4 (ddp_optimizer.cpp:335-396)
 *   infeas_out     `bool& infeas` after the call; line_failed_out likewise
 *   iters          getIterUsed();  cost = getDDPObjective()
 *   poly_coeff     getPolyCoeff()  [N][18] = [Ek_inv*x_i, u_i[0:9]]
 *   bez_coeff      getBezCoeff()   [N][18] rows [x*6,y*6,z*6]
 *   poly_time      getPolyTime()   [N]
 *   jerk           per-segment jerk cost; getJerkCost() is its sum
 *   x_final        fp.x.back(); getTerminalNorm() = |x_final - xd|^2
 *   stats          [B][8] backward sweeps, backward knots, line-search rollouts, rollout knots,
 *                  SM cycles in backward passes, in line searches, in the whole solve; [7] = kilo-cycles of the
 *                  Riccati recursion (low 32 bits) and of the sequential state rollout (high 32 bits)
 */
typedef struct direct_ddp_result {
    int32_t *rtn, *infeas_out, *line_failed_out, *iters;
    double *cost;
    double *x_final;    /* [B][9]      */
    double *poly_coeff; /* [B][N][18]  */
    double *bez_coeff;  /* [B][N][18]  */
    double *poly_time;  /* [B][N]      */
    double *jerk;       /* [B][N]      */
    int64_t *stats;     /* [B][8]      */
} direct_ddp_result;

/* The node's two-stage protocol, teach_repeat_planner.cpp:853-951 (fastTrajPlanning): stage 0 =
 * zero-init infeasible IPDDP with (w_*0, iter_max0); if it returns 2 its segment times replace the
 * durations; stage 1 = warm start from the stage-0 Bezier with (w_*, iter_max). */
typedef struct direct_ddp_two_stage {
    double w_snap0, w_terminal0, w_time0; int iter_max0;
    double w_snap, w_terminal, w_time;    int iter_max;
    int time_power;
} direct_ddp_two_stage;

typedef struct direct_ddp_stats {
    double kernel_ms;       /* device time of the solve kernel(s) of the last call (CUDA events)   */
    double h2d_ms, d2h_ms;  /* host entry points only                                              */
    int64_t bwd_sweeps, bwd_knots, fwd_trials, fwd_knots; /* summed over the batch               */
    int64_t kernel_launches;
    int grid_blocks, block_threads, smem_bytes_per_block, workspace_slots;
    int64_t h2d_bytes, d2h_bytes;
    int64_t coop_jobs, helper_units; /* jobs posted to idle warps of the CTA / units those warps ran (tail balancing) */
    int64_t spec_searches, spec_trials; /* line searches posted to the warps of idle CTAs / trials those warps ran */
    int64_t spec_sweeps, spec_sweeps_used; /* speculative backward sweeps run by warps of idle CTAs / results taken by their owners */
} direct_ddp_stats;

typedef struct direct_ddp_trace_row {
    double cost, costq, logcost, err, mu, reg, stepsize, opterr;
    int32_t step, fp_failed, n_bwd;
    int32_t t_us;   /* device time at the end of the iteration, microseconds since the solve of this trajectory began */
} direct_ddp_trace_row;

typedef struct direct_ddp_handle_s *direct_ddp_handle;

int direct_ddp_version(void);
/* Creates a solver bound to opts->device, or to the opts->ndevices devices of opts->devices.  *out is NULL on failure.
 * A handle supports ONE call in flight at a time: the asynchronous *_device entry points share the handle's scratch
 * (workspace, work-queue counter, stage-0 -> stage-1 carry buffers), so a second call must not be enqueued on another
 * stream before the first has finished; use one handle per concurrent stream. */
int direct_ddp_create(const direct_ddp_opts *opts, direct_ddp_handle *out);
void direct_ddp_destroy(direct_ddp_handle h);
const char *direct_ddp_last_error(direct_ddp_handle h);

/* One polyCurveGeneration per trajectory; host buffers (H2D, solve, D2H inside the call). */
int direct_ddp_solve_batch(direct_ddp_handle h, const direct_ddp_batch *in, direct_ddp_result *out);
/* Same with device-resident buffers; asynchronous on `stream` (a cudaStream_t, may be NULL). */
int direct_ddp_solve_batch_device(direct_ddp_handle h, const direct_ddp_batch *in, direct_ddp_result *out,
                                  void *stream);
/* fastTrajPlanning's two calls fused on the device (no host round trip between the stages).
 * out0 (stage 0) may be NULL; in->init_bez/infeas/weights/flags are ignored. */
int direct_ddp_solve_two_stage(direct_ddp_handle h, const direct_ddp_batch *in, const direct_ddp_two_stage *ts,
                               direct_ddp_result *out0, direct_ddp_result *out1);
int direct_ddp_solve_two_stage_device(direct_ddp_handle h, const direct_ddp_batch *in,
                                      const direct_ddp_two_stage *ts, direct_ddp_result *out0,
                                      direct_ddp_result *out1, void *stream);

/* initTimeAllocation, teach_repeat_planner.cpp:583-639, batched on the device:
 * points = [start_b, seeds[b][1..N-1], end_b]; writes durations[B][N].  Device pointers. */
int direct_ddp_time_allocation_device(direct_ddp_handle h, int B, int N, const double *start /*[B][3]*/,
                                      const double *end /*[B][3]*/, const double *seeds /*[B][N][3]*/,
                                      double max_vel, double max_acc, double *durations, void *stream);

/* Batched trajectory sampling: Bernstein::getPos / getVel / getAcc (utils/bezier_base.h:77-115) of every segment at the
 * S parameters s_k = k / (S - 1), scaled as the node does (teach_repeat_planner.cpp:1557-1560 position = time * getPos,
 * :681-682 velocity = getVel, acceleration = getAcc / time).  bez_coeff [B][N][18] and poly_time [B][N] as returned by
 * the solve; pos / vel / acc [B][N][S][3], any of them may be NULL.  Device pointers; asynchronous on `stream`. */
int direct_ddp_sample_device(direct_ddp_handle h, int B, int N, int S, const double *bez_coeff, const double *poly_time,
                             double *pos, double *vel, double *acc, void *stream);
/* The same with host buffers (H2D, kernel, D2H inside the call). */
int direct_ddp_sample(direct_ddp_handle h, int B, int N, int S, const double *bez_coeff, const double *poly_time,
                      double *pos, double *vel, double *acc);

/* ---- recorded corridors and the authors' comparison loop (direct_b200/host/corridor_replay.cpp) ----------------------
 * direct_ddp_corridor = msgs/msg/corridor.msg flattened as readCorridorMsg does (teach_repeat_planner.cpp:385-410);
 * the on-disk form is a plain-text dump of the message:
 *     corridor <path_id> <N>
 *     polyhedron <cx> <cy> <cz> <sx> <sy> <sz> <P>       N times, each followed by P lines "<a> <b> <c> <d>"
 * direct_ddp_replay = corridorRecCallBack's loop with alg 0 (teach_repeat_planner.cpp:309-350): the prefixes
 * n = n_min .. n_max of the corridor are planned by fastTrajPlanning's protocol (:792-951) as ONE ragged batch;
 * rows[(n - n_min)][12] = {n, compTime, sum allocTime, sum initAllocTime, rtn0, iter_used0, jerkCost0, rtn, iter_used,
 * jerkCost, terminalNorm, 0 | -1}, the columns of teach_repeat_planner.cpp:347; direct_ddp_replay_write prints them in
 * the reference's "%d %f ... %f" format. */
typedef struct direct_ddp_corridor {
    int path_id, N, P_max;
    double *planes;   /* [N][P_max][4], rows past nplanes[i] hold the always-inactive plane (0,0,0,-1) */
    int32_t *nplanes; /* [N] */
    double *center;   /* [N][3] polyhedron.center */
    double *seed;     /* [N][3] polyhedron.seed_coord */
} direct_ddp_corridor;
int direct_ddp_corridor_read(const char *path, direct_ddp_corridor **out);
int direct_ddp_corridor_write(const char *path, const direct_ddp_corridor *c);
void direct_ddp_corridor_free(direct_ddp_corridor *c);
int direct_ddp_replay(direct_ddp_handle h, const direct_ddp_corridor *c, int n_min, int n_max, const direct_ddp_two_stage *ts,
                      double max_vel, double max_acc, double *rows);
int direct_ddp_replay_write(const char *path, const double *rows, int nrows);
/* SM clock of the handle's device in Hz (converts the cycle counts of direct_ddp_result::stats into seconds). */
int direct_ddp_sm_clock_hz(direct_ddp_handle h, double *hz);

/* Statistics of the last call.  After a call that was sharded over several devices: kernel_ms, h2d_ms, d2h_ms are the
 * maxima over the devices, counters and byte counts the sums, grid_blocks / workspace_slots the sums, block_threads /
 * smem_bytes_per_block those of devices[0]. */
int direct_ddp_last_stats(direct_ddp_handle h, direct_ddp_stats *out);
/* Number of devices the handle shards host-buffer batches over (1 for a single-device handle). */
int direct_ddp_device_count(direct_ddp_handle h);
/* Register-resident FMA throughput of the device (TFLOP/s, 2 flops per FMA) for DIRECT_DDP_FP64 or
 * DIRECT_DDP_FP32: the measured denominator of the FMA roofline bench.py reports. */
int direct_ddp_measure_fma_peak(direct_ddp_handle h, int precision, double *tflops);
/* Per-iteration trace of trajectory 0 of the last (stage-1 or single) solve when opts.trace != 0. */
int direct_ddp_last_trace(direct_ddp_handle h, direct_ddp_trace_row *rows, int cap, int *len);

#ifdef __cplusplus
}
#endif
#endif
