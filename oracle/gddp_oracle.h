/* gddp_oracle.h -- CPU oracle of the generic unconstrained DDP of SURVEY.md section 8(d), model (B).  TEST INFRASTRUCTURE. */
#ifndef GDDP_ORACLE_H_
#define GDDP_ORACLE_H_
#include <stdint.h>
#define GDDP_MODEL_DINT6 0   /* 3D double integrator, nx = 6, nu = 3  (BASELINE.json configs[0]) */
#define GDDP_MODEL_QUAD12 1  /* rigid-body quadrotor, nx = 12, nu = 4 (BASELINE.json configs[1]) */
#define GDDP_MAX_NX 12
#define GDDP_MAX_NU 4
typedef struct gddp_problem {
    int model, nx, nu, B, N, iter_max;
    double dt, tol;
    const double *x0, *xg;        /* [B][nx] */
    const double *u_init;         /* [B][N][nu] or NULL (= uh everywhere) */
    double q[GDDP_MAX_NX], qf[GDDP_MAX_NX], r[GDDP_MAX_NU], uh[GDDP_MAX_NU];
} gddp_problem;
typedef struct gddp_result {
    int32_t *rtn, *iters;         /* [B]: 1 converged, 0 iter_max, -4 regularisation exhausted */
    double *cost;                 /* [B] */
    double *x, *u;                /* [B][N+1][nx], [B][N][nu] */
    int64_t *stats;               /* [B][2] backward sweeps, rollouts; or NULL */
} gddp_result;
int gddp_oracle_solve_batch(const gddp_problem *P, gddp_result *O, int fp32, int nthreads);
/* continuous dynamics and Jacobians of a model at one point (fp64): f[nx], F[nx*nx], G[nx*nu] */
void gddp_oracle_model(int model, const double *x, const double *u, double *f, double *F, double *G);
#endif
