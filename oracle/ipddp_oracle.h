/*
 * ipddp_oracle.h -- CPU fp64 restatement of the reference's IPDDP trajectory optimiser.
 *
 * TEST INFRASTRUCTURE ONLY.  This is the parity checker for the CUDA path in
 * direct_b200/csrc.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may link or call it.  The product library (libdirect_ddp_b200.so)
 * never does and has no CPU fallback.
 *
 * What it restates (all paths relative to /root/reference/global_planner):
 *   src/ddp_optimizer.cpp:5-438     ddpTrajOptimizer::polyCurveGeneration  -> ipddp_oracle_solve
 *   src/ddp_optimizer.cpp:440-644   ddpTrajOptimizer::backwardpass
 *   src/ddp_optimizer.cpp:647-778   ddpTrajOptimizer::forwardpass
 *   src/ddp_optimizer.cpp:782-823   bez2polyFunc / poly2bezFunc / sysparam2polyFunc
 *   src/ddp_optimizer.cpp:836-1060  time2barFkbarGk / ...prime / time2barR / t2tau
 *   src/ddp_optimizer.cpp:1062,1132,1289,1294  computenextx / computecminvo / computep / computeq
 *   src/ddp_optimizer.cpp:1309-1368,1455-1604  computeall and friends
 *   src/ddp_optimizer.cpp:1608-1687 initialroll / finalroll / resetfilter / resetreg / initreg
 *   src/teach_repeat_planner.cpp:583-639   initTimeAllocation            -> ipddp_oracle_time_allocation
 *   src/teach_repeat_planner.cpp:853-951   two-stage protocol            -> ipddp_oracle_two_stage
 *
 * Third-party arithmetic: Eigen3 (system package, version unpinned by the reference,
 * CMakeLists.txt:16) -- LLT is restated as the textbook unblocked lower Cholesky with
 * Eigen's "pivot <= 0 -> NumericalIssue" test; MatrixXd::inverse() as LU with partial pivoting.
 *
 * Parity pin: the reference ships no golden vectors for this path.  The pin
 * is the reference's own translation unit compiled unmodified against oracle/shim (a minimal
 * stand-in for the absent Eigen/ROS/OOQP headers) into oracle/_ref, compared with this file
 * by tests/test_oracle_vs_ref.py and frozen as fixtures under tests/golden/.
 */
#ifndef IPDDP_ORACLE_H_
#define IPDDP_ORACLE_H_

#ifdef __cplusplus
extern "C" {
#endif

#define IPDDP_NX 9   /* dim*sys_order, ddp_optimizer.cpp:36-40 */
#define IPDDP_NU 10  /* 9 high-order monomial coefficients + segment time */
#define IPDDP_NCOEF 18

typedef struct {
    int N;                  /* number of polytopes = knots                                  */
    int P_max;              /* row stride of planes                                          */
    const double *planes;   /* [N][P_max][4]  (a,b,c,d): a x + b y + c z + d <= 0 inside     */
    const int *nplanes;     /* [N]                                                            */
    const double *durations;/* [N]                                                            */
    const double *seeds;    /* [N][3] polytope seed_coord; only read when line_init != 0      */
    double x0[9];           /* [pos, vel, acc] at start   (ddp_optimizer.cpp:115-121)        */
    double xd[9];           /* [pos, vel, acc] desired at end (ddp_optimizer.cpp:104-111)    */
    const double *init_bez; /* [N][18], each row [x*6, y*6, z*6]; may be NULL (zeros)         */
    double max_vel, max_acc;
    double w_snap, w_terminal, w_time;
    int iter_max;
    int time_power;         /* 1 or 2 */
    int infeas;             /* in: alg.infeas */
    int zero_init;
    int line_init;
    int minvo;
} ipddp_problem;

typedef struct {
    double cost, costq, logcost, err, mu, reg, stepsize, opterr;
    int step, fp_failed, n_bwd;
} ipddp_iter_trace;

typedef struct {
    int rtn;                /* 0, 1, 2, -3, -4  (ddp_optimizer.cpp:335-396)                  */
    int infeas_out;         /* value of the bool& infeas after the call                      */
    int line_failed_out;    /* value of the bool& line_failed after the call                 */
    int iters;              /* iter_used                                                     */
    double cost;            /* ddpobj                                                        */
    double costq;
    double x_final[9];      /* fp.x.back(), for getTerminalNorm                              */
    double *poly_coeff;     /* [N][18] caller-owned                                          */
    double *bez_coeff;      /* [N][18] caller-owned, [x*6,y*6,z*6]                           */
    double *poly_time;      /* [N]                                                           */
    double *jerk;           /* [N]                                                           */
    /* statistics for the roofline accounting */
    long n_bwd_sweeps;      /* backward passes started                                       */
    long n_bwd_knots;       /* knots processed by all backward passes                        */
    long n_fwd_trials;      /* line-search rollouts started                                  */
    long n_fwd_knots;       /* knots processed by all rollouts                               */
    double mu_final, opterr_final;
    /* optional per-iteration trace (caller-owned, trace_cap entries) */
    ipddp_iter_trace *trace;
    int trace_cap;
    int trace_len;
} ipddp_result;

/* One call of ddpTrajOptimizer::polyCurveGeneration.  Returns 0, or <0 on bad arguments. */
int ipddp_oracle_solve(const ipddp_problem *prob, ipddp_result *res);

/* teach_repeat_planner.cpp:583-639; points = [start, seed_1..seed_{N-1}, end].  */
void ipddp_oracle_time_allocation(int N, const double *start, const double *end,
                                  const double *seeds /*[N][3]*/, double max_vel,
                                  double max_acc, double *durations /*[N]*/);

typedef struct {
    double w_snap0, w_terminal0, w_time0; int iter_max0;  /* stage 0: 1,1,1,50 */
    double w_snap, w_terminal, w_time;    int iter_max;   /* stage 1: 1,100,20,100 */
    int time_power;
} ipddp_two_stage_opts;

/* teach_repeat_planner.cpp:853-951: stage 0 (zero init, infeasible IPDDP) then stage 1
 * warm-started from stage 0.  prob->durations/init_bez/weights/flags are overridden per stage.
 * res0 may be NULL. */
int ipddp_oracle_two_stage(const ipddp_problem *prob, const ipddp_two_stage_opts *opts,
                           ipddp_result *res0, ipddp_result *res1);

/* Batch helpers used by the CPU-baseline timing (OpenMP over trajectories when built with it). */
int ipddp_oracle_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
