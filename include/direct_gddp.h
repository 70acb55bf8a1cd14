/*
 * direct_gddp.h -- C-ABI of the generic unconstrained batched DDP solver in libdirect_ddp_b200.so.
 *
 * This is model (B) of SURVEY.md section 8(d): BASELINE.json's literal "12-state / 4-input quadrotor" (configs[1]) and
 * "6-state double integrator" (configs[0]).  The reference's DDP (global_planner/src/ddp_optimizer.cpp) has neither
 * model - its only model is the flat 9-state / 10-input polynomial-segment one served by direct_ddp.h - so these entry
 * points replace no reference interface and have NO reference parity; they share the reference's solver skeleton
 * (backward Riccati sweep with a regularised Cholesky of Quu, ddp_optimizer.cpp:507-638 without the interior-point terms;
 * closed-loop rollout with step halving 2^0 .. 2^-10, ddp_optimizer.cpp:669-697) and are checked against
 * oracle/gddp_oracle.c ("parity unpinned").  Quadrotor constants: simulation/so3_quadrotor_simulator/src/dynamics/
 * Quadrotor.cpp:15-20.
 *
 * Specification:
 *   x+ = x + dt f(x, u)                          explicit Euler; A = I + dt df/dx, B = dt df/du (analytic)
 *   J  = sum_i dt/2 [(x_i-xg)' diag(q) (x_i-xg) + (u_i-uh)' diag(r) (u_i-uh)] + 1/2 (x_N-xg)' diag(qf) (x_N-xg)
 *   quad12: x = [p, v, (roll, pitch, yaw), body rates], u = [thrust, tau_x, tau_y, tau_z];  dint6: x = [p, v], u = a
 *   iteration, regularisation schedule, stopping rules: oracle/gddp_oracle.c header.
 * Handles, status codes and the precision option are those of direct_ddp.h.  No CPU fallback.
 * Kernel: two trajectories per warp (direct_b200/csrc/gddp_pair.cuh); the environment variable DIRECT_GDDP_PAIR=0 selects the
 * one-trajectory-per-warp kernel (gddp.cuh) for comparison - same decisions, numbers equal to 1e-9.
 */
#ifndef DIRECT_GDDP_H_
#define DIRECT_GDDP_H_

#include "direct_ddp.h"

#ifdef __cplusplus
extern "C" {
#endif

enum { DIRECT_GDDP_DINT6 = 0, DIRECT_GDDP_QUAD12 = 1 };

typedef struct direct_gddp_problem {
    int model, B, N, iter_max;
    double dt, tol;
    const double *x0;      /* [B][nx] */
    const double *xg;      /* [B][nx] */
    const double *u_init;  /* [B][N][nu] or NULL (= uh at every knot) */
    double q[12], qf[12], r[4], uh[4];
} direct_gddp_problem;

typedef struct direct_gddp_result {
    int32_t *rtn;    /* [B] 1 converged, 0 iter_max reached, -4 regularisation exhausted */
    int32_t *iters;  /* [B] */
    double *cost;    /* [B] */
    double *x;       /* [B][N+1][nx] */
    double *u;       /* [B][N][nu]   */
    int64_t *stats;  /* [B][4] backward sweeps, rollouts, backward knots visited, SM cycles; or NULL */
} direct_gddp_result;

/* host buffers (H2D, solve, D2H inside the call) */
int direct_gddp_solve(direct_ddp_handle h, const direct_gddp_problem *in, direct_gddp_result *out);
/* device buffers, asynchronous on `stream` (a cudaStream_t, may be NULL) */
int direct_gddp_solve_device(direct_ddp_handle h, const direct_gddp_problem *in, direct_gddp_result *out, void *stream);

#ifdef __cplusplus
}
#endif
#endif
