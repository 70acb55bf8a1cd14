/*
 * ipddp_batch.c -- flat-array batch front end of the oracle (TEST INFRASTRUCTURE ONLY).
 * Same array layout as the product's C-ABI (include/direct_ddp.h) so tests can feed both sides
 * the very same buffers.  OpenMP runs one trajectory per thread: this is the CPU baseline.
 */
#include "ipddp_oracle.h"
#include <string.h>
#include <stdlib.h>

typedef struct {
    int B, N, P_max;
    const double *planes;    /* [B][N][P_max][4] */
    const int *nplanes;      /* [B][N] */
    const double *durations; /* [B][N] */
    const double *seeds;     /* [B][N][3] or NULL */
    const double *x0, *xd;   /* [B][9] */
    const double *init_bez;  /* [B][N][18] or NULL */
    double max_vel, max_acc, w_snap, w_terminal, w_time;
    int iter_max, time_power, zero_init, line_init, minvo;
    const int *infeas;       /* [B] or NULL (-> infeas_all) */
    int infeas_all;
} oracle_batch;

typedef struct {
    int *rtn, *infeas_out, *line_failed_out, *iters; /* [B] */
    double *cost, *x_final;                          /* [B], [B][9] */
    double *poly_coeff, *bez_coeff, *poly_time, *jerk; /* [B][N][18] x2, [B][N] x2 */
    long *stats;                                     /* [B][8]: bwd sweeps, bwd knots, fwd trials, fwd knots, 4 x reserved (0) */
} oracle_out;

static void fill_problem(const oracle_batch *b, int i, ipddp_problem *p) {
    memset(p, 0, sizeof *p);
    p->N = b->N; p->P_max = b->P_max;
    p->planes = b->planes + (size_t)i * b->N * b->P_max * 4;
    p->nplanes = b->nplanes + (size_t)i * b->N;
    p->durations = b->durations + (size_t)i * b->N;
    p->seeds = b->seeds ? b->seeds + (size_t)i * b->N * 3 : NULL;
    memcpy(p->x0, b->x0 + (size_t)i * 9, 72);
    memcpy(p->xd, b->xd + (size_t)i * 9, 72);
    p->init_bez = b->init_bez ? b->init_bez + (size_t)i * b->N * 18 : NULL;
    p->max_vel = b->max_vel; p->max_acc = b->max_acc;
    p->w_snap = b->w_snap; p->w_terminal = b->w_terminal; p->w_time = b->w_time;
    p->iter_max = b->iter_max; p->time_power = b->time_power;
    p->infeas = b->infeas ? b->infeas[i] : b->infeas_all;
    p->zero_init = b->zero_init; p->line_init = b->line_init; p->minvo = b->minvo;
}

static void fill_result(const oracle_out *o, int i, int N, ipddp_result *r) {
    memset(r, 0, sizeof *r);
    r->poly_coeff = o->poly_coeff ? o->poly_coeff + (size_t)i * N * 18 : NULL;
    r->bez_coeff = o->bez_coeff ? o->bez_coeff + (size_t)i * N * 18 : NULL;
    r->poly_time = o->poly_time ? o->poly_time + (size_t)i * N : NULL;
    r->jerk = o->jerk ? o->jerk + (size_t)i * N : NULL;
}

static void store_result(const oracle_out *o, int i, const ipddp_result *r) {
    if (o->rtn) o->rtn[i] = r->rtn;
    if (o->infeas_out) o->infeas_out[i] = r->infeas_out;
    if (o->line_failed_out) o->line_failed_out[i] = r->line_failed_out;
    if (o->iters) o->iters[i] = r->iters;
    if (o->cost) o->cost[i] = r->cost;
    if (o->x_final) memcpy(o->x_final + (size_t)i * 9, r->x_final, 72);
    if (o->stats) {
        long *S = o->stats + (size_t)i * 8;
        S[0] = r->n_bwd_sweeps; S[1] = r->n_bwd_knots; S[2] = r->n_fwd_trials; S[3] = r->n_fwd_knots;
        S[4] = S[5] = S[6] = S[7] = 0;
    }
}

/* One polyCurveGeneration call per trajectory. */
int ipddp_oracle_solve_batch(const oracle_batch *b, oracle_out *o, int nthreads) {
    int status = 0;
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
    for (int i = 0; i < b->B; i++) {
        ipddp_problem p; ipddp_result r;
        fill_problem(b, i, &p);
        fill_result(o, i, b->N, &r);
        int st = ipddp_oracle_solve(&p, &r);
        if (st) {
#pragma omp critical
            status = st;
        }
        store_result(o, i, &r);
    }
    return status;
}

/* The node's two-stage protocol (teach_repeat_planner.cpp:853-951) per trajectory. o0 may be NULL. */
int ipddp_oracle_two_stage_batch(const oracle_batch *b, const ipddp_two_stage_opts *opts, oracle_out *o0,
                                 oracle_out *o1, int nthreads) {
    int status = 0;
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
    for (int i = 0; i < b->B; i++) {
        ipddp_problem p; ipddp_result r0, r1;
        fill_problem(b, i, &p);
        if (o0) fill_result(o0, i, b->N, &r0); else memset(&r0, 0, sizeof r0);
        fill_result(o1, i, b->N, &r1);
        int st = ipddp_oracle_two_stage(&p, opts, &r0, &r1);
        if (st) {
#pragma omp critical
            status = st;
        }
        if (o0) store_result(o0, i, &r0);
        store_result(o1, i, &r1);
    }
    return status;
}

/* Single solve with a per-iteration trace, for debugging parity. */
int ipddp_oracle_solve_traced(const oracle_batch *b, int i, oracle_out *o, ipddp_iter_trace *trace, int cap,
                              int *len) {
    ipddp_problem p; ipddp_result r;
    fill_problem(b, i, &p);
    fill_result(o, 0, b->N, &r);
    r.trace = trace; r.trace_cap = cap;
    int st = ipddp_oracle_solve(&p, &r);
    store_result(o, 0, &r);
    *len = r.trace_len;
    return st;
}
