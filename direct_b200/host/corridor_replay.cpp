// corridor_replay.cpp -- recorded-corridor loader and the authors' comparison loop on top of the C-ABI
// (SURVEY.md section 8(f) #2).  Host C++ only; every solve goes through direct_ddp_solve_two_stage.
//
// Reference behaviour restated here:
//   * msgs/msg/corridor.msg, polyhedron.msg, facet3.msg and readCorridorMsg (teach_repeat_planner.cpp:385-410):
//     a recorded corridor = path_id + N polyhedra, each {center, seed_coord, planes (a, b, c, d)}.
//   * corridorRecCallBack (teach_repeat_planner.cpp:309-350): for n = 2 .. 64 take the first n polyhedra, plan with
//     fastTrajPlanning (alg 0), print one 12-column row "%d %f x11" per n into .../alg{id}path{id}.
//   * fastTrajPlanning (teach_repeat_planner.cpp:792-951): start = center of polyhedron 0, end = center of polyhedron
//     n-1, zero boundary velocity / acceleration, initTimeAllocation (:583-639) over [start, seeds 1..n-1, end], the
//     two-stage solve, Results = {compTime, sum allocTime, sum initAllocTime, rtn0, iter_used0, jerkCost0, rtn,
//     iter_used, jerkCost, terminalNorm, 0 or -1 when a segment time is negative}.
// The reference replays the 63 prefixes one after the other; here they are ONE ragged batch (direct_ddp_batch::nknots).
// ROS bags cannot be parsed in this environment: the on-disk form is a plain-text dump of the message (format below).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/direct_ddp.h"

namespace {

// initTimeAllocation for one segment with v0 = 0 (teach_repeat_planner.cpp:596-637).
double trapezoid_time(const double *p0, const double *p1, double vel, double accl) {
    const double d0 = p1[0] - p0[0], d1 = p1[1] - p0[1], d2 = p1[2] - p0[2];
    const double D = std::sqrt(d0 * d0 + d1 * d1 + d2 * d2);
    const double V0 = 0.0 * (d0 / D) + 0.0 * (d1 / D) + 0.0 * (d2 / D);
    const double aV0 = std::fabs(V0);
    const double acct = (vel - V0) / accl * ((vel > V0) ? 1 : -1);
    const double accd = V0 * acct + (accl * acct * acct / 2) * ((vel > V0) ? 1 : -1);
    const double dcct = vel / accl, dccd = accl * dcct * dcct / 2;
    if (D < aV0 * aV0 / (2 * accl)) return ((V0 < 0) ? 2.0 * aV0 / accl : 0.0) + aV0 / accl;
    if (D < accd + dccd) {
        const double t1 = (V0 < 0) ? 2.0 * aV0 / accl : 0.0;
        const double t2 = (-aV0 + std::sqrt(aV0 * aV0 + accl * D - aV0 * aV0 / 2)) / accl;
        return t1 + t2 + (aV0 + accl * t2) / accl;
    }
    return acct + (D - accd - dccd) / vel + dcct;
}

}  // namespace

extern "C" {

void direct_ddp_corridor_free(direct_ddp_corridor *c) {
    if (!c) return;
    free(c->planes); free(c->nplanes); free(c->center); free(c->seed);
    free(c);
}

// Text dump of msgs/corridor:
//   corridor <path_id> <N>
//   polyhedron <cx> <cy> <cz> <sx> <sy> <sz> <P>      (N times, each followed by P lines "<a> <b> <c> <d>")
int direct_ddp_corridor_read(const char *path, direct_ddp_corridor **out) {
    if (!path || !out) return DIRECT_DDP_ERR_ARG;
    *out = nullptr;
    FILE *f = fopen(path, "r");
    if (!f) return DIRECT_DDP_ERR_ARG;
    int path_id = 0, N = 0;
    if (fscanf(f, " corridor %d %d", &path_id, &N) != 2 || N <= 0 || N > (1 << 20)) { fclose(f); return DIRECT_DDP_ERR_ARG; }
    std::vector<std::vector<double>> planes((size_t)N);
    std::vector<double> center((size_t)N * 3), seed((size_t)N * 3);
    int pmax = 0;
    for (int i = 0; i < N; i++) {
        int P = 0;
        if (fscanf(f, " polyhedron %lf %lf %lf %lf %lf %lf %d", &center[3 * i], &center[3 * i + 1], &center[3 * i + 2],
                   &seed[3 * i], &seed[3 * i + 1], &seed[3 * i + 2], &P) != 7 || P < 0 || P > DIRECT_DDP_MAX_PLANES) {
            fclose(f);
            return DIRECT_DDP_ERR_ARG;
        }
        planes[i].resize((size_t)P * 4);
        for (int k = 0; k < P; k++)
            if (fscanf(f, " %lf %lf %lf %lf", &planes[i][4 * k], &planes[i][4 * k + 1], &planes[i][4 * k + 2], &planes[i][4 * k + 3]) != 4) {
                fclose(f);
                return DIRECT_DDP_ERR_ARG;
            }
        if (P > pmax) pmax = P;
    }
    fclose(f);
    direct_ddp_corridor *c = (direct_ddp_corridor *)calloc(1, sizeof *c);
    if (!c) return DIRECT_DDP_ERR_NOMEM;
    c->path_id = path_id; c->N = N; c->P_max = pmax > 0 ? pmax : 1;
    c->planes = (double *)malloc((size_t)N * c->P_max * 4 * sizeof(double));
    c->nplanes = (int32_t *)malloc((size_t)N * sizeof(int32_t));
    c->center = (double *)malloc((size_t)N * 3 * sizeof(double));
    c->seed = (double *)malloc((size_t)N * 3 * sizeof(double));
    if (!c->planes || !c->nplanes || !c->center || !c->seed) { direct_ddp_corridor_free(c); return DIRECT_DDP_ERR_NOMEM; }
    for (int i = 0; i < N; i++) {
        const int P = (int)planes[i].size() / 4;
        c->nplanes[i] = P;
        for (int k = 0; k < c->P_max; k++) {
            double *dst = c->planes + ((size_t)i * c->P_max + k) * 4;
            if (k < P) memcpy(dst, &planes[i][4 * k], 4 * sizeof(double));
            else { dst[0] = dst[1] = dst[2] = 0.0; dst[3] = -1.0; }   // always-inactive padding (teach_repeat_planner.cpp:867-879)
        }
    }
    memcpy(c->center, center.data(), center.size() * sizeof(double));
    memcpy(c->seed, seed.data(), seed.size() * sizeof(double));
    *out = c;
    return DIRECT_DDP_OK;
}

int direct_ddp_corridor_write(const char *path, const direct_ddp_corridor *c) {
    if (!path || !c || !c->planes || !c->nplanes || !c->center || !c->seed) return DIRECT_DDP_ERR_ARG;
    FILE *f = fopen(path, "w");
    if (!f) return DIRECT_DDP_ERR_ARG;
    fprintf(f, "corridor %d %d\n", c->path_id, c->N);
    for (int i = 0; i < c->N; i++) {
        fprintf(f, "polyhedron %.17g %.17g %.17g %.17g %.17g %.17g %d\n", c->center[3 * i], c->center[3 * i + 1], c->center[3 * i + 2],
                c->seed[3 * i], c->seed[3 * i + 1], c->seed[3 * i + 2], (int)c->nplanes[i]);
        for (int k = 0; k < c->nplanes[i]; k++) {
            const double *p = c->planes + ((size_t)i * c->P_max + k) * 4;
            fprintf(f, "%.17g %.17g %.17g %.17g\n", p[0], p[1], p[2], p[3]);
        }
    }
    fclose(f);
    return DIRECT_DDP_OK;
}

int direct_ddp_replay(direct_ddp_handle h, const direct_ddp_corridor *c, int n_min, int n_max, const direct_ddp_two_stage *ts,
                      double max_vel, double max_acc, double *rows) {
    if (!h || !c || !ts || !rows || n_min < 1 || n_max < n_min || n_max > c->N) return DIRECT_DDP_ERR_ARG;
    const int B = n_max - n_min + 1, N = n_max, PM = c->P_max;
    std::vector<double> planes((size_t)B * N * PM * 4), dur((size_t)B * N, 1.0), x0((size_t)B * 9, 0.0), xd((size_t)B * 9, 0.0);
    std::vector<int32_t> npl((size_t)B * N, 0), nk((size_t)B);
    std::vector<double> init_sum((size_t)B, 0.0);
    for (int b = 0; b < B; b++) {
        const int n = n_min + b;
        nk[b] = n;
        // the first n polyhedra (teach_repeat_planner.cpp:805-810); knots past n are padded with an always-inactive cell
        for (int i = 0; i < N; i++) {
            double *dst = &planes[((size_t)b * N + i) * PM * 4];
            if (i < n) {
                memcpy(dst, c->planes + (size_t)i * PM * 4, (size_t)PM * 4 * sizeof(double));
                npl[(size_t)b * N + i] = c->nplanes[i];
            } else {
                for (int k = 0; k < PM; k++) { dst[4 * k] = dst[4 * k + 1] = dst[4 * k + 2] = 0.0; dst[4 * k + 3] = -1.0; }
            }
        }
        const double *start = c->center, *end = c->center + 3 * (size_t)(n - 1);
        for (int a = 0; a < 3; a++) { x0[(size_t)b * 9 + a] = start[a]; xd[(size_t)b * 9 + a] = end[a]; }
        // initTimeAllocation over [start, seeds 1..n-1, end] (teach_repeat_planner.cpp:583-639)
        for (int k = 0; k < n; k++) {
            const double *p0 = (k == 0) ? start : c->seed + 3 * (size_t)k;
            const double *p1 = (k == n - 1) ? end : c->seed + 3 * (size_t)(k + 1);
            const double t = trapezoid_time(p0, p1, max_vel, max_acc);
            dur[(size_t)b * N + k] = t;
            init_sum[b] += t;
        }
    }
    direct_ddp_batch in;
    memset(&in, 0, sizeof in);
    in.B = B; in.N = N; in.P_max = PM;
    in.planes = planes.data(); in.nplanes = npl.data(); in.durations = dur.data(); in.x0 = x0.data(); in.xd = xd.data();
    in.max_vel = max_vel; in.max_acc = max_acc; in.nknots = nk.data();
    std::vector<int32_t> rtn0(B), it0(B), rtn1(B), it1(B);
    std::vector<double> jerk0((size_t)B * N), jerk1((size_t)B * N), pt1((size_t)B * N), xf1((size_t)B * 9);
    std::vector<int64_t> st0((size_t)B * 8), st1((size_t)B * 8);
    direct_ddp_result o0, o1;
    memset(&o0, 0, sizeof o0); memset(&o1, 0, sizeof o1);
    o0.rtn = rtn0.data(); o0.iters = it0.data(); o0.jerk = jerk0.data(); o0.stats = st0.data();
    o1.rtn = rtn1.data(); o1.iters = it1.data(); o1.jerk = jerk1.data(); o1.poly_time = pt1.data(); o1.x_final = xf1.data();
    o1.stats = st1.data();
    const int st = direct_ddp_solve_two_stage(h, &in, ts, &o0, &o1);
    if (st) return st;
    double sm_hz = 0.0;
    direct_ddp_sm_clock_hz(h, &sm_hz);
    for (int b = 0; b < B; b++) {
        const int n = nk[b];
        double alloc = 0.0, j0 = 0.0, j1 = 0.0, tn = 0.0;
        bool positive = true;
        for (int i = 0; i < n; i++) {
            const double t = pt1[(size_t)b * N + i];
            alloc += t;
            if (t < 0.0) positive = false;
            j0 += jerk0[(size_t)b * N + i];
            j1 += jerk1[(size_t)b * N + i];
        }
        for (int a = 0; a < 9; a++) { const double d = xf1[(size_t)b * 9 + a] - xd[(size_t)b * 9 + a]; tn += d * d; }   // getTerminalNorm, ddp_optimizer.h:327-330
        double *r = rows + (size_t)b * 12;
        r[0] = n;
        // compTime: the reference adds the wall time of its two calls; here the time the owning warp spent on the two
        // stages of this trajectory (SM cycles / SM clock) - the batch as a whole takes direct_ddp_last_stats().kernel_ms
        r[1] = sm_hz > 0.0 ? (double)(st0[(size_t)b * 8 + 6] + st1[(size_t)b * 8 + 6]) / sm_hz : 0.0;
        r[2] = alloc; r[3] = init_sum[b]; r[4] = rtn0[b]; r[5] = it0[b]; r[6] = j0; r[7] = rtn1[b]; r[8] = it1[b]; r[9] = j1;
        r[10] = tn; r[11] = positive ? 0.0 : -1.0;
    }
    return DIRECT_DDP_OK;
}

// One line per prefix exactly like teach_repeat_planner.cpp:347.
int direct_ddp_replay_write(const char *path, const double *rows, int nrows) {
    if (!path || !rows || nrows < 0) return DIRECT_DDP_ERR_ARG;
    FILE *f = fopen(path, "w");
    if (!f) return DIRECT_DDP_ERR_ARG;
    for (int i = 0; i < nrows; i++) {
        const double *r = rows + (size_t)i * 12;
        fprintf(f, "%d %f %f %f %f %f %f %f %f %f %f %f\n", (int)r[0], r[1], r[2], r[3], r[4], r[5], r[6], r[7], r[8], r[9], r[10], r[11]);
    }
    fclose(f);
    return DIRECT_DDP_OK;
}

}  // extern "C"
