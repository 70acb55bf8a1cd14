// ddp_optimizer_b200.cpp -- drop-in replacement translation unit for the reference's
//   global_planner/src/ddp_optimizer.cpp
// The header global_planner/include/global_planner/ddp_optimizer.h stays byte-identical: this file
// defines the same out-of-line members of ddpTrajOptimizer (ddp_optimizer.h:267-297) and forwards the
// solve to libdirect_ddp_b200.so through the C-ABI of include/direct_ddp.h with a batch of one.
// All arithmetic happens on the GPU; this file only marshals (there is no CPU fallback: when the
// library cannot reach a B200 the call reports it and returns -100, a value the reference never uses).
//
// Build inside the reference's catkin package (INTEGRATION.md has the CMake diff):
//     replace  src/ddp_optimizer.cpp  by this file in add_executable(tr_node ...)
//     target_include_directories(tr_node PRIVATE <repo>/include)
//     target_link_libraries(tr_node <repo>/direct_b200/libdirect_ddp_b200.so)
//
// What the caller (teach_repeat_planner.cpp:853-951) reads afterwards and where it comes from:
//   return value, infeas, line_failed   <- direct_ddp_result::rtn / infeas_out / line_failed_out
//   getPolyCoeff / getBezCoeff          <- poly_coeff / bez_coeff   (N x 18, ddp_optimizer.cpp:427-436)
//   getPolyTime                         <- poly_time                 (N)
//   getDDPObjective / getIterUsed       <- cost / iters
//   getJerkCost (= jerkCost.sum())      <- jerk                      (N, finalroll ddp_optimizer.cpp:1624)
//   getTerminalNorm (fp.x.back(), fp.x_d, ddp_optimizer.h:327-330)   <- x_final and the goal state
//   getCompTime                         <- wall clock around the call, like ddp_optimizer.cpp:30,414-416
#include <global_planner/ddp_optimizer.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <cstring>
#include <vector>

#include "direct_ddp.h"

namespace {

// One solver handle per process (device ordinal and arithmetic from the environment, defaults: device 0,
// fp64 -- the reference's type).  polyCurveGeneration is called from the node's single spinner thread
// (teach_repeat_planner.cpp:1220-1226); the mutex only makes misuse from several threads safe.
struct SharedSolver {
    std::mutex mtx;
    direct_ddp_handle h = nullptr;
    int create_status = 0;
    SharedSolver() {
        direct_ddp_opts o;
        std::memset(&o, 0, sizeof o);   // fields added to the ABI later (ndevices, devices) default to 'single device'
        o.device = 0; o.precision = DIRECT_DDP_FP64; o.warps_per_block = 0; o.blocks_per_sm = 0; o.trace = 0;
        if (const char *e = std::getenv("DIRECT_DDP_DEVICE")) o.device = std::atoi(e);
        if (const char *e = std::getenv("DIRECT_DDP_PRECISION")) o.precision = (std::atoi(e) == 32) ? DIRECT_DDP_FP32 : DIRECT_DDP_FP64;
        create_status = direct_ddp_create(&o, &h);
    }
    ~SharedSolver() { direct_ddp_destroy(h); }
};
SharedSolver &shared_solver() {
    static SharedSolver s;
    return s;
}

}  // namespace

int ddpTrajOptimizer::polyCurveGeneration(
        const decomp_cvx_space::FlightCorridor &corridor,
        const Eigen::MatrixXd & /*MQM_u*/,   // unused by the reference as well (ddp_optimizer.cpp:7-8)
        const Eigen::MatrixXd & /*MQM_l*/,
        const Eigen::MatrixXd &pos,
        const Eigen::MatrixXd &vel,
        const Eigen::MatrixXd &acc,
        const Eigen::MatrixXd & /*jer*/,     // only read when sys_order == 4, unreachable (ddp_optimizer.cpp:37-39)
        const double /*minimize_order*/,
        const double maxVel,
        const double maxAcc,
        const double /*maxJer*/,
        Eigen::MatrixXd initbezCoeff,
        const double w_snap,
        const double w_terminal,
        const double w_time,
        const int iter_max,
        bool &infeas,
        bool zero_init_flag,
        bool line_init_flag,
        bool &line_failed,
        int time_power,
        bool minvo_flag)
{
    const auto t_begin = std::chrono::steady_clock::now();
    const int N = (int)corridor.polyhedrons.size();
    int P_max = 1;
    for (int i = 0; i < N; i++) P_max = std::max(P_max, (int)corridor.polyhedrons[i].planes.size());

    // ---- FlightCorridor (utils/data_type.h:190-245) -> flat arrays ---------------------------------------
    std::vector<double> planes((size_t)N * P_max * 4, 0.0), durations(N), seeds((size_t)N * 3, 0.0), bez((size_t)N * 18, 0.0);
    std::vector<int32_t> nplanes(N);
    for (int i = 0; i < N; i++) {
        const decomp_cvx_space::Polytope &pt = corridor.polyhedrons[i];
        nplanes[i] = (int32_t)pt.planes.size();
        for (int k = 0; k < nplanes[i]; k++)
            for (int c = 0; c < 4; c++) planes[((size_t)i * P_max + k) * 4 + c] = pt.planes[k](c);
        for (int k = nplanes[i]; k < P_max; k++) planes[((size_t)i * P_max + k) * 4 + 3] = -1.0;  // never-active padding
        durations[i] = corridor.durations[i];
        if (line_init_flag)
            for (int a = 0; a < 3; a++) seeds[(size_t)i * 3 + a] = pt.seed_coord(a);
    }
    const bool have_bez = (initbezCoeff.rows() == N && initbezCoeff.cols() == 18);
    if (have_bez)
        for (int i = 0; i < N; i++)
            for (int c = 0; c < 18; c++) bez[(size_t)i * 18 + c] = initbezCoeff(i, c);
    double x0[9], xd[9];
    for (int a = 0; a < 3; a++) {   // ddp_optimizer.cpp:104-121
        x0[a] = pos(0, a); x0[3 + a] = vel(0, a); x0[6 + a] = acc(0, a);
        xd[a] = pos(1, a); xd[3 + a] = vel(1, a); xd[6 + a] = acc(1, a);
    }

    direct_ddp_batch in;
    std::memset(&in, 0, sizeof in);   // fields added to the ABI later (nknots, ...) default to 'not used'
    in.B = 1; in.N = N; in.P_max = P_max;
    in.planes = planes.data(); in.nplanes = nplanes.data(); in.durations = durations.data();
    in.seeds = line_init_flag ? seeds.data() : nullptr;
    in.x0 = x0; in.xd = xd;
    in.init_bez = have_bez ? bez.data() : nullptr;
    in.infeas = nullptr; in.infeas_all = infeas ? 1 : 0;
    in.max_vel = maxVel; in.max_acc = maxAcc;
    in.w_snap = w_snap; in.w_terminal = w_terminal; in.w_time = w_time;
    in.iter_max = iter_max; in.time_power = time_power;
    in.zero_init = zero_init_flag ? 1 : 0; in.line_init = line_init_flag ? 1 : 0; in.minvo = minvo_flag ? 1 : 0;

    int32_t rtn = 0, infeas_out = 0, line_failed_out = 1, iters = 0;
    double cost = 0.0, x_final[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    std::vector<double> pc((size_t)N * 18), bz((size_t)N * 18), ptime(N), jerk(N);
    direct_ddp_result out;
    out.rtn = &rtn; out.infeas_out = &infeas_out; out.line_failed_out = &line_failed_out; out.iters = &iters;
    out.cost = &cost; out.x_final = x_final;
    out.poly_coeff = pc.data(); out.bez_coeff = bz.data(); out.poly_time = ptime.data(); out.jerk = jerk.data();
    out.stats = nullptr;

    SharedSolver &S = shared_solver();
    int status;
    {
        std::lock_guard<std::mutex> lock(S.mtx);
        status = (S.create_status == DIRECT_DDP_OK) ? direct_ddp_solve_batch(S.h, &in, &out) : S.create_status;
        if (status != DIRECT_DDP_OK) {
            ROS_ERROR("direct_ddp_b200: %s (status %d)", direct_ddp_last_error(S.h), status);
            std::fprintf(stderr, "direct_ddp_b200: %s (status %d)\n", direct_ddp_last_error(S.h), status);
        }
    }
    if (status != DIRECT_DDP_OK) {
        // The caller (teach_repeat_planner.cpp:899-944) does not know this value and goes on to call the getters: leave
        // every member they read in a defined state (zeros of the right size, the start state as "final" state).
        std::fill(pc.begin(), pc.end(), 0.0); std::fill(bz.begin(), bz.end(), 0.0);
        std::fill(ptime.begin(), ptime.end(), 0.0); std::fill(jerk.begin(), jerk.end(), 0.0);
        for (int q = 0; q < 9; q++) x_final[q] = x0[q];
        rtn = -100; cost = 0.0; iters = 0; infeas_out = infeas ? 1 : 0; line_failed_out = 1;
    }

    // ---- results -> the members the inline getters read (ddp_optimizer.h:299-340) ------------------------
    PolyCoeff = Eigen::MatrixXd::Zero(N, 18);
    BezCoeff = Eigen::MatrixXd::Zero(N, 18);
    PolyTime = Eigen::VectorXd::Zero(N);
    jerkCost = Eigen::VectorXd::Zero(N);
    for (int i = 0; i < N; i++) {
        for (int c = 0; c < 18; c++) { PolyCoeff(i, c) = pc[(size_t)i * 18 + c]; BezCoeff(i, c) = bz[(size_t)i * 18 + c]; }
        PolyTime(i) = ptime[i];
        jerkCost(i) = jerk[i];
    }
    ddpobj = cost;
    iter_used = iters;
    Eigen::VectorXd xf = Eigen::VectorXd::Zero(9), xdv = Eigen::VectorXd::Zero(9);
    for (int q = 0; q < 9; q++) { xf(q) = x_final[q]; xdv(q) = xd[q]; }
    fp.x.clear();
    fp.x.push_back(xf);   // getTerminalNorm reads fp.x.back() and fp.x_d
    fp.x_d = xdv;
    infeas = infeas_out != 0;
    line_failed = line_failed_out != 0;
    switch (rtn) {   // the reference's console messages (ddp_optimizer.cpp:336-406)
        case 2: ROS_WARN("found a feasible solution"); break;
        case -3: ROS_WARN("negative segment time"); break;
        case -4: ROS_WARN("backward pass stuck at the maximum regularisation"); break;
        default: break;
    }
    compTime = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_begin).count();
    ROS_WARN("Time consumation of %d ddp optimization is: %f", N, compTime);
    return rtn;
}

// The sweeps run inside the device kernel; the per-object entry points of the reference
// (ddp_optimizer.cpp:440, :647) have no host-side state to act on.  They exist so that every symbol the
// header declares resolves; the node never calls them (teach_repeat_planner.cpp:853-951).
void ddpTrajOptimizer::backwardpass(const decomp_cvx_space::FlightCorridor &) {
    ROS_WARN("ddpTrajOptimizer::backwardpass: the B200 build runs the sweep inside polyCurveGeneration");
}
void ddpTrajOptimizer::forwardpass(const decomp_cvx_space::FlightCorridor &) {
    ROS_WARN("ddpTrajOptimizer::forwardpass: the B200 build runs the line search inside polyCurveGeneration");
}
Eigen::MatrixXd ddpTrajOptimizer::poly2bezFunc() { return BezCoeff; }    // ddp_optimizer.cpp:799: BezCoeff of the final iterate
Eigen::MatrixXd ddpTrajOptimizer::bez2polyFunc() { return PolyCoeff; }   // ddp_optimizer.cpp:782
void ddpTrajOptimizer::sysparam2polyFunc() {}                            // ddp_optimizer.cpp:814: PolyCoeff/PolyTime are already filled
