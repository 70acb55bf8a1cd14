// tools/emulate.cpp -- DEBUG TOOL, not part of the product library and never loaded by it.
// Compiles direct_b200/csrc/ipddp_solver.h with -DDDP_EMULATE so that the warp-synchronous solver runs
// lane by lane on a CPU.  Used in the GPU-less dev container to debug the kernel logic against the
// oracle (tests/test_emulated_kernel.py); the shipped path is ipddp_kernels.cu on sm_100a.
#define DDP_EMULATE 1
#include "../direct_b200/csrc/ipddp_solver.h"

#include <cstdlib>
#include <cstring>
#include <vector>

extern "C" {
struct emu_batch {  // same layout as oracle/ipddp_batch.c::oracle_batch
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
struct emu_out {
    int *rtn, *infeas_out, *line_failed_out, *iters;
    double *cost, *x_final;
    double *poly_coeff, *bez_coeff, *poly_time, *jerk;
    long *stats;
};
struct emu_two_stage {
    double w_snap0, w_terminal0, w_time0; int iter_max0;
    double w_snap, w_terminal, w_time; int iter_max; int time_power;
};
}

template <class R>
static int run(const emu_batch *b, const emu_two_stage *ts, emu_out *o0, emu_out *o1, double *trace, int trace_cap,
               int *trace_len) {
    using namespace ddp;
    SolveArgs A;
    memset(&A, 0, sizeof A);
    A.B = b->B; A.N = b->N; A.PM = b->P_max;
    A.planes = b->planes; A.nplanes = b->nplanes; A.durations = b->durations; A.seeds = b->seeds;
    A.x0 = b->x0; A.xd = b->xd; A.init_bez = b->init_bez; A.infeas = b->infeas;
    A.max_vel = b->max_vel; A.max_acc = b->max_acc;
    std::vector<double> bez_tmp((size_t)b->B * b->N * 18), time_tmp((size_t)b->B * b->N);
    std::vector<int32_t> rtn0(b->B), inf0(b->B);
    auto fill = [&](OutPtrs &O, emu_out *o) {
        memset(&O, 0, sizeof O);
        if (!o) return;
        O.rtn = o->rtn; O.infeas_out = o->infeas_out; O.line_failed_out = o->line_failed_out; O.iters = o->iters;
        O.cost = o->cost; O.x_final = o->x_final; O.poly_coeff = o->poly_coeff; O.bez_coeff = o->bez_coeff;
        O.poly_time = o->poly_time; O.jerk = o->jerk; O.stats = (long long *)o->stats;
    };
    int max_iter;
    if (ts) {
        A.two_stage = 1;
        A.cfg[0] = StageCfg{ts->w_snap0, ts->w_terminal0, ts->w_time0, ts->iter_max0, ts->time_power, 1, 0, 0, 1};
        A.cfg[1] = StageCfg{ts->w_snap, ts->w_terminal, ts->w_time, ts->iter_max, ts->time_power, 0, 0, 0, 0};
        fill(A.out[0], o0); fill(A.out[1], o1);
        if (!A.out[0].rtn) A.out[0].rtn = rtn0.data();
        if (!A.out[0].infeas_out) A.out[0].infeas_out = inf0.data();
        A.bez_tmp = bez_tmp.data(); A.time_tmp = time_tmp.data();
        max_iter = ts->iter_max0 > ts->iter_max ? ts->iter_max0 : ts->iter_max;
    } else {
        A.two_stage = 0;
        A.cfg[0] = StageCfg{b->w_snap, b->w_terminal, b->w_time, b->iter_max, b->time_power, b->zero_init, b->line_init,
                            b->minvo, b->infeas_all};
        fill(A.out[1], o1);
        max_iter = b->iter_max;
    }
    A.fcap = max_iter + 2;
    A.trace = trace; A.trace_cap = trace_cap; A.trace_len = trace_len;
    const WsLay wl = ws_layout(A.N, A.PM, A.fcap);
    std::vector<R> ws((size_t)wl.total), sm((size_t)smem_elems_per_warp(A.PM)), tabs(360);
    const BasisTables &bt = basis_tables();
    for (int m = 0; m < 2; m++)
        for (int e = 0; e < 90; e++) { tabs[m * 180 + e] = (R)bt.val[m][e]; tabs[m * 180 + 90 + e] = (R)bt.dt[e]; }
    for (int i = 0; i < b->B; i++) {
        if (ts) solve_one<R>(A, 0, i, sm.data(), tabs.data(), ws.data(), 0);
        solve_one<R>(A, ts ? 1 : 0, i, sm.data(), tabs.data(), ws.data(), 0);
    }
    return 0;
}

extern "C" int emu_solve_batch(const emu_batch *b, emu_out *o, int fp32, double *trace, int trace_cap, int *trace_len) {
    return fp32 ? run<float>(b, nullptr, nullptr, o, trace, trace_cap, trace_len)
                : run<double>(b, nullptr, nullptr, o, trace, trace_cap, trace_len);
}
extern "C" int emu_two_stage_batch(const emu_batch *b, const emu_two_stage *ts, emu_out *o0, emu_out *o1, int fp32) {
    return fp32 ? run<float>(b, ts, o0, o1, nullptr, 0, nullptr) : run<double>(b, ts, o0, o1, nullptr, 0, nullptr);
}
