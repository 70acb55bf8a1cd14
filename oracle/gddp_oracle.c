/* gddp_oracle.c -- CPU oracle of the generic unconstrained DDP (model (B) of SURVEY.md section 8(d)).
 *
 * TEST INFRASTRUCTURE ONLY: the checker of direct_b200/csrc/gddp.cu in tests/, smoke() and bench.py's cpu_baseline.
 * Nothing under direct_b200/ may include, link or call it.
 *
 * PARITY UNPINNED.  BASELINE.json names a "12-state / 4-input quadrotor" and a "6-state double integrator"; the
 * reference's DDP has neither (its only model is the flat 9/10 polynomial-segment model, which is
 * what oracle/ipddp_oracle.c restates and pins against the reference's own translation unit).  There is therefore no
 * reference implementation, golden vector or test for this file to be pinned against.  It restates, in plain dense C,
 * the algorithm skeleton the reference's solver shares with textbook DDP - backward Riccati sweep with a regularised
 * Cholesky of Quu (ddp_optimizer.cpp:507-638 without the interior-point terms), closed-loop rollout with step halving
 * 2^0 .. 2^-10 (ddp_optimizer.cpp:669-697) - for the two models, with the quadrotor constants of
 * simulation/so3_quadrotor_simulator/src/dynamics/Quadrotor.cpp:15-20.  Its own checks (tests/test_gddp.py): Jacobians
 * against central differences, monotone cost, one-iteration convergence on the linear-quadratic double integrator, the
 * Riccati value function against a brute-force least-squares solve.
 *
 * Specification (shared with the CUDA implementation, include/direct_gddp.h):
 *   x+ = x + dt f(x, u)                         (explicit Euler; A = I + dt df/dx, B = dt df/du)
 *   J  = sum_i dt/2 [ (x_i-xg)' diag(q) (x_i-xg) + (u_i-uh)' diag(r) (u_i-uh) ] + 1/2 (x_N-xg)' diag(qf) (x_N-xg)
 *   sweep(rho): Vx = qf (x_N-xg), Vxx = diag(qf); per knot Qx, Qu, Qxx, Qux, Quu + rho I, LL' = Quu, [k|K] = -Quu^-1 [Qu|Qux],
 *               Vx = Qx + Qux' k, Vxx = sym(Qxx + Qux' K); fails on a non-positive pivot
 *   iterate: sweep (rho <- max(4 rho, 1e-6) until it succeeds, give up above 1e10: rtn -4); stop with rtn 1 when the expected
 *            first-order decrease -sum_i k_i' Qu_i is <= tol (1 + |J|); line search over alpha = 2^-s,
 *            s = 0..10, accept the first J_new < J; none: rho up, next iteration; accepted: rho <- rho/4 (0 below 1e-9),
 *            stop with rtn 1 when J - J_new <= tol (1 + |J_new|); rtn 0 after iter_max iterations.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "gddp_oracle.h"

#define REAL double
#define SUF _d
#define SIN sin
#define COS cos
#define SQRT sqrt
#define FABS fabs
#include "gddp_impl.h"
#undef REAL
#undef SUF
#undef SIN
#undef COS
#undef SQRT
#undef FABS

#define REAL float
#define SUF _f
#define SIN sinf
#define COS cosf
#define SQRT sqrtf
#define FABS fabsf
#include "gddp_impl.h"

int gddp_oracle_solve_batch(const gddp_problem *P, gddp_result *O, int fp32, int nthreads) {
    if (!P || !O || P->nx > GDDP_MAX_NX || P->nu > GDDP_MAX_NU || P->B <= 0 || P->N <= 0) return -1;
    if ((P->model == GDDP_MODEL_QUAD12 && (P->nx != 12 || P->nu != 4)) || (P->model == GDDP_MODEL_DINT6 && (P->nx != 6 || P->nu != 3))) return -1;
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
    for (int b = 0; b < P->B; b++) {
        if (fp32) solve_one_f(P, b, O);
        else solve_one_d(P, b, O);
    }
    return 0;
}
void gddp_oracle_model(int model, const double *x, const double *u, double *f, double *F, double *G) { model_f_d(model, x, u, f, F, G); }
