// oracle/ref_driver.cpp -- C entry point around the REFERENCE's own ddpTrajOptimizer
// (global_planner/src/ddp_optimizer.cpp, compiled unmodified from /root/reference against
// oracle/shim).  TEST INFRASTRUCTURE ONLY: output goes to oracle/_ref/libddp_ref.so, which pins
// the C restatement (ipddp_oracle.c) and can serve as the CPU baseline of kind "reference".
// The call below is the one teach_repeat_planner.cpp:895-897 makes.
#include <global_planner/ddp_optimizer.h>

#include <cstring>

extern "C" {

struct oracle_batch {  // must match ipddp_batch.c
    int B, N, P_max;
    const double *planes;
    const int *nplanes;
    const double *durations;
    const double *seeds;
    const double *x0, *xd;
    const double *init_bez;
    double max_vel, max_acc, w_snap, w_terminal, w_time;
    int iter_max, time_power, zero_init, line_init, minvo;
    const int *infeas;
    int infeas_all;
};
struct oracle_out {
    int *rtn, *infeas_out, *line_failed_out, *iters;
    double *cost, *x_final;
    double *poly_coeff, *bez_coeff, *poly_time, *jerk;
    long *stats;
};

static void solve_one(const oracle_batch *b, const oracle_out *o, int i) {
    const int N = b->N;
    decomp_cvx_space::FlightCorridor corridor;
    for (int k = 0; k < N; k++) {
        decomp_cvx_space::Polytope p;
        const double *pl = b->planes + ((size_t)i * N + k) * b->P_max * 4;
        int np = b->nplanes[(size_t)i * N + k];
        for (int q = 0; q < np; q++) p.appendPlane(Eigen::Vector4d(pl[q * 4], pl[q * 4 + 1], pl[q * 4 + 2], pl[q * 4 + 3]));
        if (b->seeds) {
            const double *s = b->seeds + ((size_t)i * N + k) * 3;
            p.setSeed(Eigen::Vector3d(s[0], s[1], s[2]));
            p.setCenter(Eigen::Vector3d(s[0], s[1], s[2]));
        }
        corridor.appendPolytope(p);
        corridor.appendTime(b->durations[(size_t)i * N + k]);
    }
    Eigen::MatrixXd pos = Eigen::MatrixXd::Zero(2, 3), vel = pos, acc = pos, jer = pos;
    const double *x0 = b->x0 + (size_t)i * 9, *xd = b->xd + (size_t)i * 9;
    for (int a = 0; a < 3; a++) {
        pos(0, a) = x0[a]; vel(0, a) = x0[3 + a]; acc(0, a) = x0[6 + a];
        pos(1, a) = xd[a]; vel(1, a) = xd[3 + a]; acc(1, a) = xd[6 + a];
    }
    Eigen::MatrixXd initbez = Eigen::MatrixXd::Zero(N, 18);
    if (b->init_bez)
        for (int k = 0; k < N; k++)
            for (int c = 0; c < 18; c++) initbez(k, c) = b->init_bez[((size_t)i * N + k) * 18 + c];
    Eigen::MatrixXd Qo_u = Eigen::MatrixXd::Zero(1, 1), Qo_l = Qo_u;  // unused by the DDP (ddp_optimizer.cpp:7-8)
    bool infeas = (b->infeas ? b->infeas[i] : b->infeas_all) != 0;
    bool line_failed = true;
    ddpTrajOptimizer *opt = new ddpTrajOptimizer();
    int rtn = opt->polyCurveGeneration(corridor, Qo_u, Qo_l, pos, vel, acc, jer, 3.0, b->max_vel, b->max_acc, 10.0,
                                       initbez, b->w_snap, b->w_terminal, b->w_time, b->iter_max, infeas,
                                       b->zero_init != 0, b->line_init != 0, line_failed, b->time_power, b->minvo != 0);
    if (o->rtn) o->rtn[i] = rtn;
    if (rtn <= -100) { delete opt; return; }   // the B200 drop-in TU could not reach its device: nothing to read back
    if (o->infeas_out) o->infeas_out[i] = infeas;
    if (o->line_failed_out) o->line_failed_out[i] = line_failed;
    if (o->iters) o->iters[i] = opt->getIterUsed();
    if (o->cost) o->cost[i] = opt->getDDPObjective();
    Eigen::MatrixXd pc = opt->getPolyCoeff(), bz = opt->getBezCoeff();
    Eigen::VectorXd pt = opt->getPolyTime();
    for (int k = 0; k < N; k++) {
        for (int c = 0; c < 18; c++) {
            if (o->poly_coeff) o->poly_coeff[((size_t)i * N + k) * 18 + c] = pc(k, c);
            if (o->bez_coeff) o->bez_coeff[((size_t)i * N + k) * 18 + c] = bz(k, c);
        }
        if (o->poly_time) o->poly_time[(size_t)i * N + k] = pt(k);
    }
    // the reference exposes only the terminal norm and the summed jerk; x_final[0] carries the former,
    // jerk[0] the latter (the remaining entries are left untouched).
    if (o->x_final) o->x_final[(size_t)i * 9] = opt->getTerminalNorm();
    if (o->jerk) o->jerk[(size_t)i * N] = opt->getJerkCost();
    delete opt;
}

int ddp_ref_solve_batch(const oracle_batch *b, oracle_out *o, int nthreads) {
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
    for (int i = 0; i < b->B; i++) solve_one(b, o, i);
    return 0;
}
}
