// ipddp_solver.h -- warp-per-trajectory interior-point DDP (IPDDP) solve for sm_100a.
//
// One warp owns one trajectory for its whole solve: setup, initial rollout, the outer barrier loop,
// backward Riccati/IP sweeps, filter line search and output conversion all run inside one kernel
// without host round trips.  This file is the device-side equivalent of the reference's
//   global_planner/src/ddp_optimizer.cpp   ("ddp.cpp" below)
//     polyCurveGeneration :5-438, backwardpass :440-644, forwardpass :647-778, bez2polyFunc :782,
//     poly2bezFunc :799, computenextx :1062, computecminvo :1132, computeq :1294,
//     computeall :1309-1368 + :1455-1604, initialroll :1608, finalroll :1624, resetfilter :1636
// re-designed rather than ported:
//   * no Jacobian is ever materialised: every constraint row is (basis row beta_g(T)) (x) (direction n)
//     plus a d/dT entry, so c, J*v, J^T*w and J^T D J are evaluated from 15 "row groups"
//     (6 position control points, 5 velocity, 4 acceleration) and the polytope planes;
//   * lane <-> plane for the 6P corridor rows, lane <-> (group, axis) for the 55 fixed rows;
//   * the 19x19 Hessian of Q in z = [u(10); x(9)] (plus the gradient as a 20th row/column) is held one
//     column per lane; ten right-looking Cholesky pivots over the u block leave V_xx, V_x in the
//     trailing block (a Schur complement) and the gains come from one back-substitution per lane;
//   * per-knot state (x,u,s,y, gains) streams through a per-warp workspace slot in global memory,
//     c and the slack gains ks,Ks,ky,Ky are recomputed on the fly instead of being stored.
// Arithmetic is re-associated with respect to the reference (documented in DESIGN.md), so results
// agree with the oracle to rounding, not bit for bit.
//
// The code is written against simt.h so that the very same source also runs lane-by-lane on a CPU
// for debugging (tools/emulate.cpp).
#ifndef DIRECT_B200_IPDDP_SOLVER_H_
#define DIRECT_B200_IPDDP_SOLVER_H_

#include "simt.h"

namespace ddp {

// ---------------------------------------------------------------------------------------------
// Basis tables (values of the reference's tables, ddp.cpp:62-96 and :1543-1560).  Row g: 0-5
// position control points, 6-10 velocity, 11-14 acceleration; column l = monomial order.
// DT is ALWAYS the MINVO derivative table, also when values use the Bezier tables: the reference
// does exactly that (hazard H1 in SURVEY.md) and parity needs it.
// ---------------------------------------------------------------------------------------------
struct BasisTables {
    double val[2][90];  // [minvo][g*6+l]
    double dt[90];
};
inline const BasisTables &basis_tables() {
    static const BasisTables t = {
        {{// Bezier, ddp.cpp:79-95
          1.0, 0, 0, 0, 0, 0, 1.0, 0.2, 0, 0, 0, 0, 1.0, 0.4, 0.1, 0, 0, 0, 1.0, 0.6, 0.3, 0.1, 0, 0, 1.0, 0.8, 0.6,
          0.4, 0.2, 0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0,
          0, 1.0, 0, 0, 0, 0, 0, 1.0, 0.5, 0, 0, 0, 0, 1.0, 1.0, 0.5, 0, 0, 0, 1.0, 1.5, 1.5, 1.0, 0, 0, 1.0, 2.0, 3.0,
          4.0, 5.0,
          0, 0, 2.0, 0, 0, 0, 0, 0, 2.0, 2.0, 0, 0, 0, 0, 2.0, 4.0, 4.0, 0, 0, 0, 2.0, 6.0, 12.0, 20.0},
         {// MINVO, ddp.cpp:63-77
          1.0, -0.06471861202, -0.03728008486, -0.02577637794, -0.02027573243, -0.01678273037,
          1.0, 0.03314986096, -0.06548114211, -0.05530463802, -0.04362718953, -0.03671639115,
          1.0, 0.3375528997, 0.05836232552, -0.02920033165, -0.04690387913, -0.04376447947,
          1.0, 0.6624471003, 0.3832565261, 0.1916286091, 0.06985980172, -0.002892843108,
          1.0, 0.966850139, 0.868219136, 0.7594116288, 0.6521050661, 0.5510660979,
          1.0, 1.064718612, 1.092157139, 1.108091959, 1.118023718, 1.123960059,
          0, 1.0, -0.1423379297, -0.1332742327, -0.1242105357, -0.126304257,
          0, 1.0, 0.1887439858, -0.1831318297, -0.2466606848, -0.2321393311,
          0, 1.0, 1.0, 0.5411016575, 0.08220331498, -0.2433474658,
          0, 1.0, 1.811256014, 2.250636213, 2.381669451, 2.282405938,
          0, 1.0, 2.14233793, 3.293739556, 4.445141183, 5.585385392,
          0, 0, 2.0, -0.4472869252, -0.6133793313, -0.6406553622,
          0, 0, 2.0, 1.223711659, -0.5552714346, -1.854618819,
          0, 0, 2.0, 4.776288341, 6.54988193, 6.841145057,
          0, 0, 2.0, 6.447286925, 13.17576837, 22.04662796}},
        {// d/dT of the MINVO tables, ddp.cpp:1544-1560
         0, -0.06471861202, -0.07456016972, -0.07732913382, -0.08110292972, -0.08391365186,
         0, 0.03314986096, -0.1309622842, -0.1659139141, -0.1745087581, -0.1835819558,
         0, 0.3375528997, 0.116724651, -0.08760099494, -0.1876155165, -0.2188223973,
         0, 0.6624471003, 0.7665130522, 0.5748858272, 0.2794392069, -0.01446421554,
         0, 0.966850139, 1.736438272, 2.278234886, 2.608420264, 2.755330489,
         0, 1.064718612, 2.184314278, 3.324275878, 4.472094873, 5.619800295,
         0, 0, -0.1423379297, -0.2665484655, -0.3726316072, -0.5052170278,
         0, 0, 0.1887439858, -0.3662636595, -0.7399820545, -0.9285573245,
         0, 0, 1.0, 1.082203315, 0.2466099449, -0.9733898632,
         0, 0, 1.811256014, 4.501272426, 7.145008354, 9.129623752,
         0, 0, 2.14233793, 6.587479113, 13.33542355, 22.34154157,
         0, 0, 0, -0.4472869252, -1.226758663, -1.921966087,
         0, 0, 0, 1.223711659, -1.110542869, -5.563856457,
         0, 0, 0, 4.776288341, 13.09976386, 20.52343517,
         0, 0, 0, 6.447286925, 26.35153674, 66.13988387}};
    return t;
}

// ---------------------------------------------------------------------------------------------
// Kernel arguments (device pointers).  I/O buffers are always double (the ABI's type).
// ---------------------------------------------------------------------------------------------
struct StageCfg {
    double w_snap, w_terminal, w_time;
    int iter_max, time_power, zero_init, line_init, minvo, infeas_all;
};
struct OutPtrs {
    int32_t *rtn, *infeas_out, *line_failed_out, *iters;
    double *cost, *x_final, *poly_coeff, *bez_coeff, *poly_time, *jerk;
    long long *stats;
};
struct SolveArgs {
    int B, N, PM;  // PM = P_max (row stride of planes)
    const double *planes;
    const int32_t *nplanes;
    const double *durations, *seeds, *x0, *xd, *init_bez;
    const int32_t *infeas;
    double max_vel, max_acc;
    StageCfg cfg[2];
    int two_stage;     // 0: cfg[0] only, outputs -> out[1]; 1: stage 0 (out[0] optional) then stage 1
    OutPtrs out[2];
    double *bez_tmp, *time_tmp;  // [B][N][18], [B][N] scratch carrying stage 0 -> stage 1
    void *ws;                    // workspace, ws_stride elements of Real per warp slot
    long long ws_stride;
    int fcap;                    // filter capacity per slot
    unsigned int *counter;       // work queue
    double *trace;               // optional [cap][12] trace of trajectory 0 (last stage), or NULL
    int trace_cap;
    int *trace_len;
};

// Shared-memory layout of one warp's scratch, in elements of Real.
struct Lay {
    int PM;
    DDP_DEVICE explicit Lay(int pm) : PM(pm) {}
    enum { BS = 0, BDT = 90, ZO = 180, ZN = 200, KC = 220, VXX = 320, VX = 404, XD = 416, PL = 428 };
    DDP_DEVICE int cend() const { return 428 + 4 * PM; }
    // backward-only region
    DDP_DEVICE int FG() const { return cend(); }
    DDP_DEVICE int FT() const { return cend() + 20; }
    DDP_DEVICE int RR() const { return cend() + 32; }
    DDP_DEVICE int TT9() const { return cend() + 52; }
    DDP_DEVICE int TRB() const { return cend() + 64; }
    DDP_DEVICE int G() const { return cend() + 154; }
    DDP_DEVICE int X() const { return cend() + 326; }
    DDP_DEVICE int XT() const { return cend() + 346; }
    DDP_DEVICE int XI() const { return cend() + 366; }
    DDP_DEVICE int LF() const { return cend() + 378; }
    DDP_DEVICE int W() const { return cend() + 578; }
    DDP_DEVICE int BV() const { return cend() + 578 + 30 * PM; }
    // forward-only region (aliases the backward one)
    DDP_DEVICE int BSN() const { return cend(); }
    DDP_DEVICE int TRF() const { return cend() + 90; }
    DDP_DEVICE int V1() const { return cend() + 318; }
    DDP_DEVICE int V2() const { return cend() + 338; }
    DDP_DEVICE int DX() const { return cend() + 358; }
    DDP_DEVICE int FGN() const { return cend() + 370; }
    DDP_DEVICE int total() const { return cend() + 578 + 40 * PM; }
};
DDP_HD int smem_elems_per_warp(int pm) { return 428 + 4 * pm + 578 + 40 * pm; }

// Workspace slot layout (elements of Real).
struct WsLay {
    long long xu, xun, s, sn, y, yn, K, filt, total;
    int MCS;
};
DDP_HD WsLay ws_layout(int N, int PM, int fcap) {
    WsLay w;
    w.MCS = (6 * PM + 55 + 3) & ~3;
    long long o = 0;
    w.xu = o; o += (long long)(N + 1) * 20;
    w.xun = o; o += (long long)(N + 1) * 20;
    w.s = o; o += (long long)N * w.MCS;
    w.sn = o; o += (long long)N * w.MCS;
    w.y = o; o += (long long)N * w.MCS;
    w.yn = o; o += (long long)N * w.MCS;
    w.K = o; o += (long long)N * 100;
    w.filt = o; o += 2LL * fcap;
    w.total = (o + 15) & ~15LL;
    return w;
}

DDP_DEVICE long long ddp_clock() {
#if DDP_GPU
    return clock64();
#else
    return 0;
#endif
}

// z-space index of monomial coefficient (l, axis): z = [u(0..8), T(9), x(10..18)].
DDP_DEVICE int zidx(int l, int a) { return l < 3 ? 10 + 3 * l + a : 3 * (l - 3) + a; }

template <class R> struct Traj {
    int N, PM, MCS, lane_;
    const double *planes;
    const int32_t *nplanes;
    R *sm;
    const R *tab;  // [0..89] value table (chosen basis), [90..179] dt table
    R *xu, *xun, *s, *sn, *y, *yn, *K, *filt;
    int fcap;
    R max_vel, max_acc, w_snap, w_terminal, w_time, margin;
    int time_power, infeas, zero_init, line_init;
    R mu, tol, reg, reg_base, opterr, cost, costq, logcost, err, stepsize;
    int step, failed, bfailed, nfilter;
    long long n_bwd_sweeps, n_bwd_knots, n_fwd_trials, n_fwd_knots;
    long long cyc_bwd, cyc_fwd, cyc_t0;  // clock64 accounting (0 in the emulation)
};

// Row slots of a lane: 0..5 corridor rows (control point j = slot, plane = lane), 6 = fixed "+" row,
// 7 = fixed "-" row.  Fixed lanes: 0..14 velocity (group 6+f/3, axis f%3), 15..26 acceleration,
// 27 the time row -T + 0.3 <= 0.  Reference row order (ddp.cpp:1181-1187, :1238, :1276, :1279).
DDP_DEVICE int row_index(int slot, int lane, int P) {
    if (slot < 6) return slot * P + lane;
    const int b = 6 * P;
    if (slot == 6) return lane < 15 ? b + lane : (lane < 27 ? b + 30 + (lane - 15) : b + 54);
    return lane < 15 ? b + 15 + lane : b + 42 + (lane - 15);
}
DDP_DEVICE bool row_valid(int slot, int lane, int P) {
    if (slot < 6) return lane < P;
    if (slot == 6) return lane < 28;
    return lane < 27;
}

template <class R> DDP_DEVICE void load_rows(const R *src, int P, Reg<R, 8> &dst, int lane_) {
    FOR_LANES(lane) {
        DDP_UNROLL
        for (int q = 0; q < 8; q++) dst(lane, q) = row_valid(q, lane, P) ? src[row_index(q, lane, P)] : R(1);
    }
}
template <class R> DDP_DEVICE void store_rows(R *dst, int P, const Reg<R, 8> &src, int lane_) {
    FOR_LANES(lane) {
        DDP_UNROLL
        for (int q = 0; q < 8; q++)
            if (row_valid(q, lane, P)) dst[row_index(q, lane, P)] = src(lane, q);
    }
}

// T-scaled basis rows: Bs[g*6+l] = tab[g][l] * T^(l-shift_g) * Ek_inv[l]  (ddp.cpp:1148-1160, :1203-1209,
// :1240-1247; Ek_inv = {1,1,1/2} folded in, ddp.cpp:1138-1143), Bdt likewise from the d/dT table
// (ddp.cpp:1543-1560).  Powers are formed by repeated multiplication like the reference's Tkv.
template <class R> DDP_DEVICE void scale_tables(const R *tab, R T, R *Bs, R *Bdt, int lane) {
    R tp1 = T, tp2 = tp1 * T, tp3 = tp2 * T, tp4 = tp3 * T, tp5 = tp4 * T;
    DDP_UNROLL
    for (int q = 0; q < 6; q++) {
        int e = lane + 32 * q;
        if (e >= 180) break;
        if (e >= 90 && Bdt == nullptr) break;
        int ee = e >= 90 ? e - 90 : e;
        int g = ee / 6, l = ee - 6 * g;
        int k = l - (g < 6 ? 0 : (g < 11 ? 1 : 2)) - (e >= 90 ? 1 : 0);
        R pw = k <= 0 ? R(1) : (k == 1 ? tp1 : (k == 2 ? tp2 : (k == 3 ? tp3 : (k == 4 ? tp4 : tp5))));
        R v = tab[e] * pw;
        if (l == 2) v = v * R(0.5);
        if (e >= 90) Bdt[ee] = v; else Bs[ee] = v;
    }
}

// dst[g*3+a] = sum_l B[g*6+l] * v[zidx(l,a)]  for the 45 (group, axis) pairs.
template <class R> DDP_DEVICE void transform45(const R *B, const R *v, R *dst, int lane) {
    DDP_UNROLL
    for (int q = 0; q < 2; q++) {
        int o = lane + 32 * q;
        if (o < 45) {
            int g = o / 3, a = o - 3 * g;
            R acc = R(0);
            DDP_UNROLL
            for (int l = 0; l < 6; l++) acc += B[g * 6 + l] * v[zidx(l, a)];
            dst[o] = acc;
        }
    }
}

// F, G of the segment dynamics x+ = (F (x) I3) x + (G (x) I3) u[0:9], ddp.cpp:862-871.  FG[o*6+l].
template <class R> DDP_DEVICE R fg_entry(int o, int l, R T) {
    R T2 = T * T, T3 = T2 * T, T4 = T3 * T, T5 = T4 * T;
    switch (o * 6 + l) {
        case 0: return R(1); case 1: return T; case 2: return T2 / R(2); case 3: return T3; case 4: return T4; case 5: return T5;
        case 6: return R(0); case 7: return R(1); case 8: return T; case 9: return R(3) * T2; case 10: return R(4) * T3; case 11: return R(5) * T4;
        case 12: return R(0); case 13: return R(0); case 14: return R(1); case 15: return R(6) * T; case 16: return R(12) * T2;
        default: return R(20) * T3;
    }
}
// d/dT of the above, ddp.cpp:930-935.
template <class R> DDP_DEVICE R fgp_entry(int o, int l, R T) {
    R T2 = T * T, T3 = T2 * T, T4 = T3 * T;
    switch (o * 6 + l) {
        case 1: return R(1); case 2: return T; case 3: return R(3) * T2; case 4: return R(4) * T3; case 5: return R(5) * T4;
        case 8: return R(1); case 9: return R(6) * T; case 10: return R(12) * T2; case 11: return R(20) * T3;
        case 15: return R(6); case 16: return R(24) * T; case 17: return R(60) * T2;
        default: return R(0);
    }
}
// Jerk-cost matrices R, R', R'' (3x3, index 3*i+j), ddp.cpp:991-999.
template <class R> DDP_DEVICE R rmat_entry(int which, int i, int j, R T) {
    R T2 = T * T, T3 = T2 * T, T4 = T3 * T, T5 = T4 * T;
    const int e = i * 3 + j;
    if (which == 0) {
        switch (e) { case 0: return R(36) * T; case 1: case 3: return R(72) * T2; case 2: case 6: return R(120) * T3;
                     case 4: return R(192) * T3; case 5: case 7: return R(360) * T4; default: return R(720) * T5; }
    } else if (which == 1) {
        switch (e) { case 0: return R(36); case 1: case 3: return R(144) * T; case 2: case 6: return R(360) * T2;
                     case 4: return R(576) * T2; case 5: case 7: return R(1440) * T3; default: return R(3600) * T4; }
    }
    switch (e) { case 0: return R(0); case 1: case 3: return R(144); case 2: case 6: return R(720) * T;
                 case 4: return R(1152) * T; case 5: case 7: return R(4320) * T2; default: return R(14400) * T3; }
}

// One lane's share of u^T (M (x) I3) u for lane c < 9 (c = 3*i + axis): u_c * sum_j M[i][j] u[3j+axis].
template <class R> DDP_DEVICE R quad_share(int which, const R *u, int c, R T) {
    int i = c / 3, a = c - 3 * i;
    R t = R(0);
    DDP_UNROLL
    for (int j = 0; j < 3; j++) t += rmat_entry<R>(which, i, j, T) * u[3 * j + a];
    return u[c] * t;
}

// Load the planes of knot i into shared memory (P x 4).
template <class R> DDP_DEVICE void load_planes(const Traj<R> &t, int knot, int P, int lane) {
    const double *pl = t.planes + (long long)knot * t.PM * 4;
    for (int e = lane; e < 4 * P; e += 32) t.sm[Lay::PL + e] = (R)pl[e];
}

// =============================================================================================
// Backward pass (ddp.cpp:440-644).  Returns with t.bfailed / t.opterr set; gains in t.K.
// =============================================================================================
template <class R> DDP_DEVICE_NOINLINE void backward_pass(Traj<R> &t) {
    const int lane_ = t.lane_;
    const Lay L(t.PM);
    R *sm = t.sm;
    const long long clk0 = ddp_clock();
    t.n_bwd_sweeps++;
    // regularisation schedule, ddp.cpp:452-474
    if (t.failed || t.bfailed) t.reg = t.reg + R(1);
    else if (t.step == 0) t.reg = t.reg - R(1);
    else if (t.step <= 3) t.reg = t.reg;
    else t.reg = t.reg + R(1);
    if (t.reg < R(0)) t.reg = R(0);
    else if (t.reg > R(24)) t.reg = R(24);
    const R regadd = rpow(t.reg_base, t.reg) - R(1);  // ddp.cpp:529
    const R sgn = t.infeas ? R(1) : R(-1);
    const R mu = t.mu;

    // terminal value function, ddp.cpp:1318-1323: Vx = P (x_N - x_d), Vxx = P = w_terminal I
    FOR_LANES(lane) {
        for (int e = lane; e < 81; e += 32) sm[Lay::VXX + e] = (e / 9 == e % 9) ? t.w_terminal : R(0);
        if (lane < 9) sm[Lay::VX + lane] = t.w_terminal * (t.xu[(long long)t.N * 20 + 10 + lane] - sm[Lay::XD + lane]);
    }
    WARP_SYNC();

    Reg<R, 3> errs;  // per-lane running maxima: |Qu|, |r|, |c+y|
    FOR_LANES(lane) { errs(lane, 0) = R(0); errs(lane, 1) = R(0); errs(lane, 2) = R(0); }

    for (int i = t.N - 1; i >= 0; i--) {
        t.n_bwd_knots++;
        const int P = t.nplanes[i];
        Reg<R, 8> s, y;
        load_rows(t.s + (long long)i * t.MCS, P, s, lane_);
        if (t.infeas) load_rows(t.y + (long long)i * t.MCS, P, y, lane_);
        Reg<R, 4> sums;  // per-lane partials: tt (H_TT from constraints), gt (grad_T), u'R'u, u'R''u
        // ---- phase A: knot state, scaled tables, dynamics/cost pieces ------------------------------
        FOR_LANES(lane) {
            if (lane < 20) sm[Lay::ZO + lane] = t.xu[(long long)i * 20 + lane];
            load_planes(t, i, P, lane);
            sums(lane, 0) = R(0); sums(lane, 1) = R(0); sums(lane, 2) = R(0); sums(lane, 3) = R(0);
        }
        WARP_SYNC();
        const R T = sm[Lay::ZO + 9];
        FOR_LANES(lane) {
            scale_tables(t.tab, T, sm + Lay::BS, sm + Lay::BDT, lane);
            if (lane < 18) sm[L.FG() + lane] = fg_entry<R>(lane / 6, lane % 6, T);
            if (lane < 18) sm[L.RR() + lane] = rmat_entry<R>(lane / 9, (lane % 9) / 3, lane % 3, T);
            if (lane < 9) {  // fT = d x+/dT = (F' (x) I) x + (G' (x) I) u, ddp.cpp:1332
                int o = lane / 3, a = lane - 3 * o;
                R s1 = R(0), s2 = R(0);
                DDP_UNROLL
                for (int b = 0; b < 3; b++) {
                    if (b > o) s1 += fgp_entry<R>(o, b, T) * sm[Lay::ZO + 10 + 3 * b + a];
                    s2 += fgp_entry<R>(o, 3 + b, T) * sm[Lay::ZO + 3 * b + a];
                }
                sm[L.FT() + lane] = s1 + s2;
                sums(lane, 2) = quad_share<R>(1, sm + Lay::ZO, lane, T);
                sums(lane, 3) = quad_share<R>(2, sm + Lay::ZO, lane, T);
            }
        }
        WARP_SYNC();
        // ---- phase B: control points and their d/dT ------------------------------------------------
        FOR_LANES(lane) {
            transform45(sm + Lay::BS, sm + Lay::ZO, sm + L.TRB(), lane);
            transform45(sm + Lay::BDT, sm + Lay::ZO, sm + L.TRB() + 45, lane);
            if (lane < 9) {  // tT = Vxx fT
                R acc = R(0);
                DDP_UNROLL
                for (int q = 0; q < 9; q++) acc += sm[Lay::VXX + lane * 9 + q] * sm[L.FT() + q];
                sm[L.TT9() + lane] = acc;
            }
        }
        WARP_SYNC();
        // ---- phase C: constraint rows -> weights --------------------------------------------------
        // per row: c, tcol = dc/dT, then (ddp.cpp:535-541 / :583-590)
        //   infeasible: D = s/y, r = s y - mu, tv2 = (s (c+y) - r)/y ;  feasible: D = s/c, r = s c + mu, tv2 = r/c
        //   H += sgn J^T D J,   grad += J^T (s + sgn tv2)
        FOR_LANES(lane) {
            R emu = errs(lane, 1), ecy = errs(lane, 2);
            if (lane < P) {
                const R n0 = sm[Lay::PL + 4 * lane], n1 = sm[Lay::PL + 4 * lane + 1], n2 = sm[Lay::PL + 4 * lane + 2],
                        n3 = sm[Lay::PL + 4 * lane + 3];
                R *bv = sm + L.BV() + 10 * lane;
                bv[0] = n0 * n0; bv[1] = n0 * n1; bv[2] = n0 * n2; bv[3] = n1 * n1; bv[4] = n1 * n2; bv[5] = n2 * n2;
                bv[6] = n0; bv[7] = n1; bv[8] = n2; bv[9] = R(1);
                DDP_UNROLL
                for (int j = 0; j < 6; j++) {
                    const R *cp = sm + L.TRB() + 3 * j, *cd = sm + L.TRB() + 45 + 3 * j;
                    R c = ((n0 * cp[0] + n1 * cp[1]) + n2 * cp[2]) + n3 - t.margin;
                    R tc = (n0 * cd[0] + n1 * cd[1]) + n2 * cd[2];
                    R sv = s(lane, j), D, r, tv2;
                    if (t.infeas) {
                        R yv = y(lane, j), yinv = R(1) / yv;
                        r = sv * yv - mu; D = sv * yinv; tv2 = yinv * (sv * (c + yv) - r);
                        ecy = amax(ecy, rabs(c + yv));
                    } else {
                        R cinv = R(1) / c;
                        r = sv * c + mu; D = sv * cinv; tv2 = cinv * r;
                    }
                    emu = amax(emu, rabs(r));
                    R Ds = sgn * D, gw = sv + sgn * tv2;
                    R *w = sm + L.W() + (j * t.PM + lane);
                    const int st = 6 * t.PM;
                    w[0] = Ds; w[st] = Ds * tc; w[2 * st] = (Ds * tc) * tc; w[3 * st] = gw; w[4 * st] = gw * tc;
                }
            }
            if (lane < 28) {
                R cplus, cminus = R(-1), tc;
                if (lane < 27) {
                    R val = sm[L.TRB() + 18 + lane], lim = lane < 15 ? t.max_vel : t.max_acc;
                    tc = sm[L.TRB() + 45 + 18 + lane];
                    cplus = val - lim - t.margin; cminus = -val - lim - t.margin;
                } else {
                    cplus = -T + R(0.3) - t.margin; tc = R(-1);
                }
                R Dsum = R(0), gdiff = R(0);
                DDP_UNROLL
                for (int pm = 0; pm < 2; pm++) {
                    if (pm == 1 && lane == 27) break;
                    R c = pm ? cminus : cplus, sv = s(lane, 6 + pm), D, r, tv2;
                    if (t.infeas) {
                        R yv = y(lane, 6 + pm), yinv = R(1) / yv;
                        r = sv * yv - mu; D = sv * yinv; tv2 = yinv * (sv * (c + yv) - r);
                        ecy = amax(ecy, rabs(c + yv));
                    } else {
                        R cinv = R(1) / c;
                        r = sv * c + mu; D = sv * cinv; tv2 = cinv * r;
                    }
                    emu = amax(emu, rabs(r));
                    Dsum += sgn * D;
                    gdiff += pm ? -(sv + sgn * tv2) : (sv + sgn * tv2);
                }
                // the "-" row has Jacobian -J+, so D adds and the gradient weight subtracts
                if (lane < 27) {
                    int g = lane < 15 ? lane / 3 : 5 + (lane - 15) / 3, a = lane < 15 ? lane % 3 : (lane - 15) % 3;
                    R *G = sm + L.G() + 90 + 9 * g;  // fixed groups 0..8 = row groups 6..14
                    G[a] = Dsum; G[3 + a] = Dsum * tc; G[6 + a] = gdiff;
                }
                sums(lane, 0) += (Dsum * tc) * tc;
                sums(lane, 1) += gdiff * tc;
            }
            errs(lane, 1) = emu; errs(lane, 2) = ecy;
        }
        WARP_SYNC();
        // ---- phase D: reduce the corridor rows over planes, per control point ----------------------
        // outputs per control point j: M (3x3 sym) , w (3), tt, gv (3), gt
        FOR_LANES(lane) {
            DDP_UNROLL
            for (int q = 0; q < 3; q++) {
                int o = lane + 32 * q;
                if (o < 84) {
                    int j = o / 14, e = o - 14 * j;
                    // weight selector / plane-vector selector
                    int ws = e < 6 ? 0 : (e < 9 ? 1 : (e == 9 ? 2 : (e < 13 ? 3 : 4)));
                    int bs = e < 6 ? e : (e < 9 ? e : (e == 9 ? 9 : (e < 13 ? e - 4 : 9)));
                    const R *w = sm + L.W() + ws * 6 * t.PM + j * t.PM;
                    const R *bv = sm + L.BV() + bs;
                    R acc = R(0);
                    for (int k = 0; k < P; k++) acc += w[k] * bv[10 * k];
                    R *G = sm + L.G() + 15 * j;
                    if (e < 6) {
                        const int a = e < 3 ? 0 : (e < 5 ? 1 : 2), b = e < 3 ? e : (e < 5 ? e - 2 : 2);
                        G[a * 3 + b] = acc; G[b * 3 + a] = acc;
                    } else if (e < 9) G[9 + (e - 6)] = acc;
                    else if (e == 9) sums(lane, 0) += acc;
                    else if (e < 13) G[12 + (e - 10)] = acc;
                    else sums(lane, 1) += acc;
                }
            }
        }
        WARP_SYNC();
        const R ttot = warp_sum(sums, 0, lane_), gtot = warp_sum(sums, 1, lane_);
        const R uRpu = warp_sum(sums, 2, lane_), uRppu = warp_sum(sums, 3, lane_);
        // ---- phase E: assemble column `lane` of the augmented Hessian ------------------------------
        // rows/cols 0..8 u-coefficients, 9 T, 10..18 x, 19 gradient.
        Reg<R, 20> col;
        FOR_LANES(lane) {
            DDP_UNROLL
            for (int r = 0; r < 20; r++) col(lane, r) = R(0);
            if (lane < 19 && lane != 9) {
                const int lp = lane < 9 ? 3 + lane / 3 : (lane - 10) / 3, ap = lane < 9 ? lane % 3 : (lane - 10) % 3;
                // constraints: sum over the 15 row groups of (beta beta^T) (x) M
                DDP_UNROLL
                for (int g = 0; g < 15; g++) {
                    const int lmin = g < 6 ? 0 : (g < 11 ? 1 : 2);
                    const R sc = sm[Lay::BS + g * 6 + lp];
                    R m0, m1, m2, wv, gv;
                    if (g < 6) {
                        const R *G = sm + L.G() + 15 * g;
                        m0 = sc * G[ap]; m1 = sc * G[3 + ap]; m2 = sc * G[6 + ap]; wv = G[9 + ap]; gv = G[12 + ap];
                    } else {
                        const R *G = sm + L.G() + 90 + 9 * (g - 6);
                        R d = sc * G[ap];
                        m0 = ap == 0 ? d : R(0); m1 = ap == 1 ? d : R(0); m2 = ap == 2 ? d : R(0);
                        wv = G[3 + ap]; gv = G[6 + ap];
                    }
                    DDP_UNROLL
                    for (int l = lmin; l < 6; l++) {
                        const R bl = sm[Lay::BS + g * 6 + l];
                        col(lane, zidx(l, 0)) += bl * m0;
                        col(lane, zidx(l, 1)) += bl * m1;
                        col(lane, zidx(l, 2)) += bl * m2;
                    }
                    col(lane, 9) += sc * wv;
                    col(lane, 19) += sc * gv;
                }
                // dynamics: A^T Vxx A and A^T Vx with A = [G (x) I | fT | F (x) I] (ddp.cpp:508-520)
                R fg0 = sm[L.FG() + lp], fg1 = sm[L.FG() + 6 + lp], fg2 = sm[L.FG() + 12 + lp];
                R tt[9];
                DDP_UNROLL
                for (int p = 0; p < 9; p++)
                    tt[p] = (sm[Lay::VXX + p * 9 + ap] * fg0 + sm[Lay::VXX + p * 9 + 3 + ap] * fg1) +
                            sm[Lay::VXX + p * 9 + 6 + ap] * fg2;
                DDP_UNROLL
                for (int l = 0; l < 6; l++) {
                    const R f0 = sm[L.FG() + l], f1 = sm[L.FG() + 6 + l], f2 = sm[L.FG() + 12 + l];
                    DDP_UNROLL
                    for (int a = 0; a < 3; a++) col(lane, zidx(l, a)) += (f0 * tt[a] + f1 * tt[3 + a]) + f2 * tt[6 + a];
                }
                R hT = R(0);
                DDP_UNROLL
                for (int p = 0; p < 9; p++) hT += sm[L.FT() + p] * tt[p];
                col(lane, 9) += hT;
                col(lane, 19) += (fg0 * sm[Lay::VX + ap] + fg1 * sm[Lay::VX + 3 + ap]) + fg2 * sm[Lay::VX + 6 + ap];
                // stage cost (ddp.cpp:1350-1355): quu = w [R (x) I, R'u; (R'u)^T, .], qu = w R u
                if (lane < 9) {
                    const int ip = lane / 3;
                    R Ru = R(0), Rpu = R(0);
                    DDP_UNROLL
                    for (int b = 0; b < 3; b++) {
                        const R rv = sm[L.RR() + ip * 3 + b];
                        Ru += rv * sm[Lay::ZO + 3 * b + ap];
                        Rpu += sm[L.RR() + 9 + ip * 3 + b] * sm[Lay::ZO + 3 * b + ap];
                        const R qv = t.w_snap * rv;
                        col(lane, 3 * b + 0) += ap == 0 ? qv : R(0);
                        col(lane, 3 * b + 1) += ap == 1 ? qv : R(0);
                        col(lane, 3 * b + 2) += ap == 2 ? qv : R(0);
                    }
                    col(lane, 9) += t.w_snap * Rpu;
                    col(lane, 19) += t.w_snap * Ru;
                }
                sm[L.XT() + lane] = col(lane, 9);
                sm[L.X() + lane] = col(lane, 19);
            }
            if (lane == 9) {
                R hTT = R(0), gT = R(0);
                DDP_UNROLL
                for (int p = 0; p < 9; p++) { hTT += sm[L.FT() + p] * sm[L.TT9() + p]; gT += sm[L.FT() + p] * sm[Lay::VX + p]; }
                R quT, quuTT;
                if (t.time_power == 2) { quT = t.w_time * T + R(0.5) * t.w_snap * uRpu; quuTT = t.w_time + R(0.5) * t.w_snap * uRppu; }
                else { quT = R(0.5) * t.w_time + R(0.5) * t.w_snap * uRpu; quuTT = R(0.5) * t.w_snap * uRppu; }
                col(lane, 9) = (quuTT + hTT) + ttot;
                col(lane, 19) = (quT + gT) + gtot;
                sm[L.X() + 9] = col(lane, 19);
            }
        }
        WARP_SYNC();
        FOR_LANES(lane) {
            if (lane == 9) {
                DDP_UNROLL
                for (int r = 0; r < 19; r++) if (r != 9) col(lane, r) = sm[L.XT() + r];
            }
            if (lane == 19) {
                DDP_UNROLL
                for (int r = 0; r < 19; r++) col(lane, r) = sm[L.X() + r];
            }
            if (lane < 10) errs(lane, 0) = amax(errs(lane, 0), rabs(col(lane, 19)));  // |Qu|, ddp.cpp:633
        }
        // ---- phase F: ten Cholesky pivots over the u block (Eigen::LLT, ddp.cpp:543/:592) -----------
        Reg<R, 10> mult;  // lane c keeps row p of L^-1 [H_u: | g_u] restricted to its column
        bool fail = false;
        DDP_UNROLL
        for (int p = 0; p < 10; p++) {
            const R d = warp_bcast(col, p, p, lane_) + regadd;
            if (d <= R(0)) { fail = true; break; }
            const R piv = rsqrt_(d), inv = R(1) / piv;
            FOR_LANES(lane) {
                R m = (lane == p) ? piv : col(lane, p) * inv;
                mult(lane, p) = m;
                if (lane < 20) sm[L.LF() + p * 20 + lane] = m;
            }
            FOR_LANES(lane) { if (lane == 0) sm[L.XI() + p] = inv; }
            WARP_SYNC();
            FOR_LANES(lane) {
                if (lane > p && lane < 20) {
                    const R m = mult(lane, p);
                    DDP_UNROLL
                    for (int r = p + 1; r < 20; r++) col(lane, r) -= sm[L.LF() + p * 20 + r] * m;
                }
            }
        }
        if (fail) {  // ddp.cpp:546-551 / :595-600
            t.bfailed = 1;
            t.opterr = R(INFINITY);
            t.cyc_bwd += ddp_clock() - clk0;
            return;
        }
        // ---- phase G: gains [ku | Ku] = -(L L^T)^-1 [Qu | Qux] (ddp.cpp:561-564 / :607-609) -----------
        Reg<R, 10> kx;
        FOR_LANES(lane) {
            if (lane >= 10 && lane < 20) {
                DDP_UNROLL
                for (int p = 9; p >= 0; p--) {
                    R v = mult(lane, p);
                    DDP_UNROLL
                    for (int q = p + 1; q < 10; q++) v -= sm[L.LF() + p * 20 + q] * (-kx(lane, q));
                    kx(lane, p) = -(v * sm[L.XI() + p]);
                }
                const int qc = lane == 19 ? 0 : lane - 9;
                DDP_UNROLL
                for (int p = 0; p < 10; p++) sm[Lay::KC + p * 10 + qc] = kx(lane, p);
            }
        }
        WARP_SYNC();
        // The reference backs up with the UNREGULARISED Quu (ddp.cpp:574/:615, :620-627).  With
        // K = -(Quu+rho I)^-1 Qux that equals the Schur complement above minus rho K^T K (and
        // minus rho K^T k for Vx).
        if (regadd != R(0)) {
            FOR_LANES(lane) {
                if (lane >= 10 && lane < 19) {
                    DDP_UNROLL
                    for (int r = 10; r < 19; r++) {
                        R acc = R(0);
                        DDP_UNROLL
                        for (int p = 0; p < 10; p++) acc += sm[Lay::KC + p * 10 + (r - 9)] * kx(lane, p);
                        col(lane, r) -= regadd * acc;
                    }
                    R acc = R(0);
                    DDP_UNROLL
                    for (int p = 0; p < 10; p++) acc += sm[Lay::KC + p * 10] * kx(lane, p);
                    col(lane, 19) -= regadd * acc;
                }
            }
        }
        FOR_LANES(lane) {
            if (lane >= 10 && lane < 19) {
                DDP_UNROLL
                for (int a = 0; a < 9; a++) sm[Lay::VXX + (lane - 10) * 9 + a] = col(lane, 10 + a);
                sm[Lay::VX + (lane - 10)] = col(lane, 19);
            }
            for (int e = lane; e < 100; e += 32) t.K[(long long)i * 100 + e] = sm[Lay::KC + e];
        }
        WARP_SYNC();
        Reg<R, 3> symv;  // Vxx = (Vxx + Vxx^T)/2, ddp.cpp:628
        FOR_LANES(lane) {
            DDP_UNROLL
            for (int q = 0; q < 3; q++) {
                int e = lane + 32 * q;
                if (e < 81) { int a = e / 9, b = e - 9 * a; symv(lane, q) = R(0.5) * (sm[Lay::VXX + e] + sm[Lay::VXX + b * 9 + a]); }
            }
        }
        WARP_SYNC();
        FOR_LANES(lane) {
            DDP_UNROLL
            for (int q = 0; q < 3; q++) { int e = lane + 32 * q; if (e < 81) sm[Lay::VXX + e] = symv(lane, q); }
        }
        WARP_SYNC();
    }
    t.bfailed = 0;
    const R e0 = warp_max(errs, 0, lane_), e1 = warp_max(errs, 1, lane_), e2 = warp_max(errs, 2, lane_);
    t.opterr = rmax(rmax(e0, t.infeas ? e2 : R(0)), e1);  // ddp.cpp:641
    t.cyc_bwd += ddp_clock() - clk0;
}

// =============================================================================================
// One rollout: initial roll (mode 0, ddp.cpp:1608-1620) or a line-search trial (mode 1,
// ddp.cpp:674-734) with step size alpha.  Writes the candidate into xun/sn/yn and returns false when
// the fraction-to-boundary test fails (ddp.cpp:683-687 / :699-703).
// =============================================================================================
template <class R> struct RollOut { R cost, costq, logcost, err; };

template <class R> DDP_DEVICE_NOINLINE bool rollout(Traj<R> &t, int mode, R alpha, R tau, RollOut<R> &out) {
    const int lane_ = t.lane_;
    const Lay L(t.PM);
    R *sm = t.sm;
    const R mu = t.mu;
    Reg<R, 3> acc;  // per-lane partials: stage cost, log barrier, |c+y|_1
    FOR_LANES(lane) {
        acc(lane, 0) = R(0); acc(lane, 1) = R(0); acc(lane, 2) = R(0);
        if (lane < 9) sm[Lay::ZN + 10 + lane] = t.xu[10 + lane];  // xnew[0] = xold[0]
    }
    WARP_SYNC();
    for (int i = 0; i < t.N; i++) {
        if (mode == 1) t.n_fwd_knots++;
        const int P = t.nplanes[i];
        Reg<R, 8> s, y;
        if (mode == 1) {
            load_rows(t.s + (long long)i * t.MCS, P, s, lane_);
            if (t.infeas) load_rows(t.y + (long long)i * t.MCS, P, y, lane_);
        }
        // ---- phase A ---------------------------------------------------------------------------------
        FOR_LANES(lane) {
            if (lane < 20) sm[Lay::ZO + lane] = t.xu[(long long)i * 20 + lane];
            load_planes(t, i, P, lane);
            if (mode == 1) for (int e = lane; e < 100; e += 32) sm[Lay::KC + e] = t.K[(long long)i * 100 + e];
        }
        WARP_SYNC();
        const R Told = sm[Lay::ZO + 9];
        FOR_LANES(lane) {
            if (mode == 1) {
                scale_tables(t.tab, Told, sm + Lay::BS, sm + Lay::BDT, lane);
                if (lane < 9) sm[L.DX() + lane] = sm[Lay::ZN + 10 + lane] - sm[Lay::ZO + 10 + lane];
            } else if (lane < 10) sm[Lay::ZN + lane] = sm[Lay::ZO + lane];  // initial roll keeps u
        }
        WARP_SYNC();
        // ---- phase B: unew = (uold + alpha ku) + Ku dx (ddp.cpp:689/:695); v1 = [ku;0], v2 = [Ku dx; dx] --
        if (mode == 1) {
            FOR_LANES(lane) {
                if (lane < 10) {
                    R kdx = R(0);
                    DDP_UNROLL
                    for (int b = 0; b < 9; b++) kdx += sm[Lay::KC + lane * 10 + 1 + b] * sm[L.DX() + b];
                    const R ku = sm[Lay::KC + lane * 10];
                    sm[Lay::ZN + lane] = (sm[Lay::ZO + lane] + alpha * ku) + kdx;
                    sm[L.V1() + lane] = ku; sm[L.V2() + lane] = kdx;
                } else if (lane < 19) {
                    sm[L.V1() + lane] = R(0); sm[L.V2() + lane] = sm[L.DX() + lane - 10];
                }
            }
            WARP_SYNC();
        }
        const R Tn = sm[Lay::ZN + 9];
        FOR_LANES(lane) {
            if (lane < 18) sm[L.FGN() + lane] = fg_entry<R>(lane / 6, lane % 6, Tn);
            if (mode == 1) {
                scale_tables<R>(t.tab, Tn, sm + L.BSN(), nullptr, lane);
                transform45(sm + Lay::BS, sm + Lay::ZO, sm + L.TRF(), lane);
                transform45(sm + Lay::BDT, sm + Lay::ZO, sm + L.TRF() + 45, lane);
                transform45(sm + Lay::BS, sm + L.V1(), sm + L.TRF() + 90, lane);
                transform45(sm + Lay::BS, sm + L.V2(), sm + L.TRF() + 135, lane);
            }
        }
        WARP_SYNC();
        if (mode == 1) {
            FOR_LANES(lane) { transform45(sm + L.BSN(), sm + Lay::ZN, sm + L.TRF() + 180, lane); }
            WARP_SYNC();
        }
        // ---- phase C: rows ---------------------------------------------------------------------------
        Reg<int, 1> bad;
        Reg<R, 8> sn, yn;
        const R v1T = mode == 1 ? sm[L.V1() + 9] : R(0), v2T = mode == 1 ? sm[L.V2() + 9] : R(0);
        if (mode == 1) FOR_LANES(lane) {
            int isbad = 0;
            R lp = R(1), e1 = R(0);
            DDP_UNROLL
            for (int q = 0; q < 8; q++) {
                if (!row_valid(q, lane, P)) { sn(lane, q) = R(1); yn(lane, q) = R(1); continue; }
                R cold = R(0), cnew, jv1 = R(0), jv2 = R(0);
                if (q < 6) {
                    const R n0 = sm[Lay::PL + 4 * lane], n1 = sm[Lay::PL + 4 * lane + 1], n2 = sm[Lay::PL + 4 * lane + 2],
                            n3 = sm[Lay::PL + 4 * lane + 3];
                    const R *f = sm + L.TRF() + 3 * q;
                    cnew = ((n0 * f[180] + n1 * f[181]) + n2 * f[182]) + n3 - t.margin;
                    if (mode == 1) {
                        cold = ((n0 * f[0] + n1 * f[1]) + n2 * f[2]) + n3 - t.margin;
                        const R tc = (n0 * f[45] + n1 * f[46]) + n2 * f[47];
                        jv1 = ((n0 * f[90] + n1 * f[91]) + n2 * f[92]) + tc * v1T;
                        jv2 = ((n0 * f[135] + n1 * f[136]) + n2 * f[137]) + tc * v2T;
                    }
                } else if (lane < 27) {
                    const R sg = q == 6 ? R(1) : R(-1), lim = lane < 15 ? t.max_vel : t.max_acc;
                    const R *f = sm + L.TRF() + 18 + lane;
                    cnew = sg * f[180] - lim - t.margin;
                    if (mode == 1) {
                        cold = sg * f[0] - lim - t.margin;
                        jv1 = sg * (f[90] + f[45] * v1T);
                        jv2 = sg * (f[135] + f[45] * v2T);
                    }
                } else {
                    cnew = -Tn + R(0.3) - t.margin;
                    if (mode == 1) { cold = -Told + R(0.3) - t.margin; jv1 = -v1T; jv2 = -v2T; }
                }
                if (mode == 1) {
                    const R sv = s(lane, q);
                    if (t.infeas) {  // ddp.cpp:535-536, :568-572, :680-684
                        const R yv = y(lane, q), yinv = R(1) / yv;
                        const R r = sv * yv - mu, rhat = sv * (cold + yv) - r, D = sv * yinv;
                        const R ks = yinv * (rhat + sv * jv1), ky = -(cold + yv) - jv1;
                        const R ynew = (yv + alpha * ky) + (-jv2), snew = (sv + alpha * ks) + D * jv2;
                        if (ynew < (R(1) - tau) * yv || snew < (R(1) - tau) * sv) isbad = 1;
                        sn(lane, q) = snew; yn(lane, q) = ynew;
                        lp *= ynew; e1 += rabs(cnew + ynew);
                    } else {  // ddp.cpp:583-586, :611-612, :694-700
                        const R cinv = R(1) / cold;
                        const R r = sv * cold + mu, D = sv * cinv;
                        const R ks = -(cinv * (r + sv * jv1));
                        const R snew = (sv + alpha * ks) + (-(D * jv2));
                        if (cnew > (R(1) - tau) * cold || snew < (R(1) - tau) * sv) isbad = 1;
                        sn(lane, q) = snew;
                        lp *= -cnew;
                    }
                }
            }
            bad(lane, 0) = isbad;
            if (mode == 1) { acc(lane, 1) += rlog(lp); acc(lane, 2) += e1; }
        }
        if (mode == 1) {
            if (warp_any(bad, 0, lane_)) return false;
            store_rows(t.sn + (long long)i * t.MCS, P, sn, lane_);
            if (t.infeas) store_rows(t.yn + (long long)i * t.MCS, P, yn, lane_);
        }
        // ---- phase D: stage cost (ddp.cpp:1294-1305), next state (ddp.cpp:1062-1067), store ---------------
        Reg<R, 1> xn;
        FOR_LANES(lane) {
            if (lane < 9) {
                acc(lane, 0) += R(0.5) * t.w_snap * quad_share<R>(0, sm + Lay::ZN, lane, Tn);
                const int o = lane / 3, a = lane - 3 * o;
                R s1 = R(0), s2 = R(0);
                DDP_UNROLL
                for (int b = 0; b < 3; b++) {
                    if (b >= o) s1 += sm[L.FGN() + o * 6 + b] * sm[Lay::ZN + 10 + 3 * b + a];
                    s2 += sm[L.FGN() + o * 6 + 3 + b] * sm[Lay::ZN + 3 * b + a];
                }
                xn(lane, 0) = s1 + s2;
            }
            if (lane == 9) acc(lane, 0) += t.time_power == 2 ? R(0.5) * Tn * t.w_time * Tn : R(0.5) * t.w_time * Tn;
            R *dst = mode == 1 ? t.xun : t.xu;
            if (lane < 19) dst[(long long)i * 20 + lane] = sm[Lay::ZN + lane];
        }
        WARP_SYNC();
        FOR_LANES(lane) { if (lane < 9) sm[Lay::ZN + 10 + lane] = xn(lane, 0); }
        WARP_SYNC();
    }
    // terminal cost (ddp.cpp:1289-1292) and totals
    Reg<R, 1> pt;
    FOR_LANES(lane) {
        pt(lane, 0) = R(0);
        if (lane < 9) {
            const R d = sm[Lay::ZN + 10 + lane] - sm[Lay::XD + lane];
            pt(lane, 0) = d * (t.w_terminal * d);
            R *dst = mode == 1 ? t.xun : t.xu;
            dst[(long long)t.N * 20 + 10 + lane] = sm[Lay::ZN + 10 + lane];
        }
    }
    WARP_SYNC();
    const R qs = warp_sum(acc, 0, lane_), p = R(0.5) * warp_sum(pt, 0, lane_);
    out.costq = qs;
    out.cost = qs + p;
    if (mode == 1) {
        const R ls = warp_sum(acc, 1, lane_);
        out.logcost = out.cost - mu * ls;
        out.err = t.infeas ? rmax(t.tol, warp_sum(acc, 2, lane_)) : R(0);
    }
    return true;
}

// Barrier cost and infeasibility at the current iterate + filter reset (ddp.cpp:1636-1662).
template <class R> DDP_DEVICE_NOINLINE void reset_filter(Traj<R> &t) {
    const int lane_ = t.lane_;
    const Lay L(t.PM);
    R *sm = t.sm;
    Reg<R, 2> acc;
    FOR_LANES(lane) { acc(lane, 0) = R(0); acc(lane, 1) = R(0); }
    for (int i = 0; i < t.N; i++) {
        const int P = t.nplanes[i];
        Reg<R, 8> y;
        if (t.infeas) load_rows(t.y + (long long)i * t.MCS, P, y, lane_);
        FOR_LANES(lane) {
            if (lane < 20) sm[Lay::ZO + lane] = t.xu[(long long)i * 20 + lane];
            load_planes(t, i, P, lane);
        }
        WARP_SYNC();
        const R T = sm[Lay::ZO + 9];
        FOR_LANES(lane) { scale_tables<R>(t.tab, T, sm + Lay::BS, nullptr, lane); }
        WARP_SYNC();
        FOR_LANES(lane) { transform45(sm + Lay::BS, sm + Lay::ZO, sm + L.TRF(), lane); }
        WARP_SYNC();
        FOR_LANES(lane) {
            R lp = R(1), e1 = R(0);
            DDP_UNROLL
            for (int q = 0; q < 8; q++) {
                if (!row_valid(q, lane, P)) continue;
                R c;
                if (q < 6) {
                    const R *f = sm + L.TRF() + 3 * q;
                    c = ((sm[Lay::PL + 4 * lane] * f[0] + sm[Lay::PL + 4 * lane + 1] * f[1]) + sm[Lay::PL + 4 * lane + 2] * f[2]) +
                        sm[Lay::PL + 4 * lane + 3] - t.margin;
                } else if (lane < 27) {
                    c = (q == 6 ? R(1) : R(-1)) * sm[L.TRF() + 18 + lane] - (lane < 15 ? t.max_vel : t.max_acc) - t.margin;
                } else c = -T + R(0.3) - t.margin;
                if (t.infeas) { lp *= y(lane, q); e1 += rabs(c + y(lane, q)); }
                else lp *= -c;
            }
            acc(lane, 0) += rlog(lp); acc(lane, 1) += e1;
        }
        WARP_SYNC();
    }
    t.logcost = t.cost - t.mu * warp_sum(acc, 0, lane_);
    t.err = R(0);
    if (t.infeas) { t.err = warp_sum(acc, 1, lane_); if (t.err < t.tol) t.err = R(0); }
    FOR_LANES(lane) { if (lane == 0) { t.filt[0] = t.logcost; t.filt[1] = t.err; } }
    WARP_SYNC();
    t.nfilter = 1;
    t.step = 0;
    t.failed = 0;
}

// Number of constraint rows with c >= 2e-4 at the current iterate (ddp.cpp:346-355).
template <class R> DDP_DEVICE_NOINLINE bool any_violation(Traj<R> &t, R thresh, bool strict) {
    const int lane_ = t.lane_;
    const Lay L(t.PM);
    R *sm = t.sm;
    Reg<int, 1> viol;
    FOR_LANES(lane) { viol(lane, 0) = 0; }
    for (int i = 0; i < t.N; i++) {
        const int P = t.nplanes[i];
        FOR_LANES(lane) {
            if (lane < 20) sm[Lay::ZO + lane] = t.xu[(long long)i * 20 + lane];
            load_planes(t, i, P, lane);
        }
        WARP_SYNC();
        const R T = sm[Lay::ZO + 9];
        FOR_LANES(lane) { scale_tables<R>(t.tab, T, sm + Lay::BS, nullptr, lane); }
        WARP_SYNC();
        FOR_LANES(lane) { transform45(sm + Lay::BS, sm + Lay::ZO, sm + L.TRF(), lane); }
        WARP_SYNC();
        FOR_LANES(lane) {
            DDP_UNROLL
            for (int q = 0; q < 8; q++) {
                if (!row_valid(q, lane, P)) continue;
                R c;
                if (q < 6) {
                    const R *f = sm + L.TRF() + 3 * q;
                    c = ((sm[Lay::PL + 4 * lane] * f[0] + sm[Lay::PL + 4 * lane + 1] * f[1]) + sm[Lay::PL + 4 * lane + 2] * f[2]) +
                        sm[Lay::PL + 4 * lane + 3] - t.margin;
                } else if (lane < 27) {
                    c = (q == 6 ? R(1) : R(-1)) * sm[L.TRF() + 18 + lane] - (lane < 15 ? t.max_vel : t.max_acc) - t.margin;
                } else c = -T + R(0.3) - t.margin;
                if (strict ? (c > thresh) : (c >= thresh)) viol(lane, 0) = 1;
            }
        }
        WARP_SYNC();
    }
    return warp_any(viol, 0, lane_);
}

// Line search with the filter (ddp.cpp:647-778).
template <class R> DDP_DEVICE_NOINLINE void forward_pass(Traj<R> &t) {
    const int lane_ = t.lane_;
    const long long clk0 = ddp_clock();
    const R tau = rmax(R(0.99), R(1) - t.mu);
    bool failed = true;
    RollOut<R> ro;
    R stepsize = R(0);
    int step;
    for (step = 0; step < 11; step++) {
        stepsize = R(1);
        for (int k = 0; k < step; k++) stepsize = stepsize * R(0.5);  // 2^-step exactly (ddp.cpp:670)
        t.n_fwd_trials++;
        if (!rollout(t, 1, stepsize, tau, ro)) continue;
        // filter acceptance, ddp.cpp:741-757: rejected if some entry is <= the candidate in both
        // coordinates; an accepted candidate evicts the entries it dominates.  Lane 0 owns the filter.
        FOR_LANES(lane) {
            if (lane == 0) {
                bool rej = false;
                for (int k = 0; k < t.nfilter; k++)
                    if (ro.logcost >= t.filt[2 * k] && ro.err >= t.filt[2 * k + 1]) { rej = true; break; }
                int nk = 0;
                if (!rej) {
                    for (int k = 0; k < t.nfilter; k++) {
                        const R f0 = t.filt[2 * k], f1 = t.filt[2 * k + 1];
                        if (ro.logcost > f0 || ro.err > f1) { t.filt[2 * nk] = f0; t.filt[2 * nk + 1] = f1; nk++; }
                    }
                    if (nk >= t.fcap) nk = t.fcap - 1;
                    t.filt[2 * nk] = ro.logcost; t.filt[2 * nk + 1] = ro.err;
                }
                t.sm[Lay::ZO + 19] = rej ? R(1) : R(0);
                t.sm[Lay::ZN + 19] = R(nk);
            }
        }
        WARP_SYNC();
        const bool rej = t.sm[Lay::ZO + 19] != R(0);
        const int nkeep = (int)t.sm[Lay::ZN + 19];
        WARP_SYNC();
        if (rej) continue;
        t.nfilter = nkeep + 1;
        failed = false;
        break;
    }
    if (failed) {
        t.failed = 1;
        t.stepsize = R(0);
    } else {
        t.cost = ro.cost; t.costq = ro.costq; t.logcost = ro.logcost; t.err = ro.err;
        R *tmp;
        tmp = t.xu; t.xu = t.xun; t.xun = tmp;
        tmp = t.s; t.s = t.sn; t.sn = tmp;
        if (t.infeas) { tmp = t.y; t.y = t.yn; t.yn = tmp; }
        t.stepsize = stepsize; t.step = step; t.failed = 0;
    }
    t.cyc_fwd += ddp_clock() - clk0;
}

// =============================================================================================
// One polyCurveGeneration (ddp.cpp:5-438) for trajectory `b`, stage `st` of the call.
// =============================================================================================
template <class R> DDP_DEVICE_NOINLINE void solve_one(const SolveArgs &A, int st, int b, R *sm, const R *tabs, R *ws,
                                                      int lane_) {
    const StageCfg &cfg = A.cfg[st];
    const int N = A.N;
    const WsLay wl = ws_layout(N, A.PM, A.fcap);
    Traj<R> t;
    t.N = N; t.PM = A.PM; t.MCS = wl.MCS; t.lane_ = lane_;
    t.planes = A.planes + (long long)b * N * A.PM * 4;
    t.nplanes = A.nplanes + (long long)b * N;
    t.sm = sm;
    t.tab = tabs + (cfg.minvo ? 180 : 0);
    t.xu = ws + wl.xu; t.xun = ws + wl.xun; t.s = ws + wl.s; t.sn = ws + wl.sn; t.y = ws + wl.y; t.yn = ws + wl.yn;
    t.K = ws + wl.K; t.filt = ws + wl.filt; t.fcap = A.fcap;
    t.max_vel = (R)A.max_vel; t.max_acc = (R)A.max_acc;
    t.w_snap = (R)cfg.w_snap; t.w_terminal = (R)cfg.w_terminal; t.w_time = (R)cfg.w_time;
    t.margin = cfg.minvo ? R(0) : R(2.0e-4);  // ddp.cpp:1281-1283
    t.time_power = cfg.time_power; t.zero_init = cfg.zero_init; t.line_init = cfg.line_init;
    t.tol = R(1.0e-7);                                 // ddp.cpp:43
    t.reg_base = cfg.zero_init ? R(1.6) : R(4.0);      // ddp.cpp:60-61
    t.n_bwd_sweeps = t.n_bwd_knots = t.n_fwd_trials = t.n_fwd_knots = 0;
    t.cyc_bwd = t.cyc_fwd = 0; t.cyc_t0 = ddp_clock();
    const bool from_stage0 = (A.two_stage && st == 1);
    int infeas_in;
    if (from_stage0) infeas_in = A.out[0].infeas_out ? A.out[0].infeas_out[b] : 1;
    else infeas_in = A.two_stage ? 1 : (A.infeas ? A.infeas[b] : cfg.infeas_all);
    t.infeas = infeas_in;
    const double *dur = A.durations + (long long)b * N;
    const double *ibez = A.init_bez ? A.init_bez + (long long)b * N * 18 : nullptr;
    if (from_stage0) {
        ibez = A.bez_tmp + (long long)b * N * 18;
        if (A.out[0].rtn[b] == 2) dur = A.time_tmp + (long long)b * N;  // UpdateTime, teach_repeat_planner.cpp:911-912
    }
    const Lay L(t.PM);

    // ---- setup (ddp.cpp:104-193) ---------------------------------------------------------------------
    FOR_LANES(lane) {
        if (lane < 9) {
            sm[Lay::XD + lane] = (R)A.xd[(long long)b * 9 + lane];
            t.xu[10 + lane] = (R)A.x0[(long long)b * 9 + lane];
        }
    }
    for (int i = 0; i < N; i++) {
        const int P = t.nplanes[i];
        const double T = dur[i];
        FOR_LANES(lane) {
            // warm start: Bezier [x*6,y*6,z*6] scaled by T -> monomial coefficients 3..5 (ddp.cpp:167-193,:782-796)
            if (lane < 9) {
                R uv = R(0);
                if (!cfg.zero_init && !cfg.line_init && ibez) {
                    const int cidx = 3 + lane / 3, a = lane % 3;
                    // (poly2bez * t2tauMat)(j, c) = BERN[c][j] / T^c
                    const double Tk = 1.0 / T;
                    double pw = 1.0;
                    for (int k = 0; k < cidx; k++) pw = (k == 0) ? Tk : pw * Tk;
                    const int bern[3][6] = {{-10, 30, -30, 10, 0, 0}, {5, -20, 30, -20, 5, 0}, {-1, 5, -10, 10, -5, 1}};
                    double accv = 0.0;
                    for (int j = 0; j < 6; j++) accv += (T * ibez[(long long)i * 18 + a * 6 + j]) * (bern[cidx - 3][j] * pw);
                    uv = (R)accv;
                }
                t.xu[(long long)i * 20 + lane] = uv;
            }
            if (lane == 9) t.xu[(long long)i * 20 + 9] = (R)T;
            const int mc = 6 * P + 55;
            for (int e = lane; e < mc; e += 32) {
                t.s[(long long)i * t.MCS + e] = R(0.1);   // ddp.cpp:150-151
                t.y[(long long)i * t.MCS + e] = R(0.01);
            }
        }
    }
    WARP_SYNC();
    RollOut<R> ro;
    rollout(t, 0, R(0), R(0), ro);  // initialroll, ddp.cpp:252
    t.cost = ro.cost; t.costq = ro.costq;
    t.mu = t.cost / R(N) / R(6 * t.nplanes[0] + 55);  // ddp.cpp:281 (hazard H3)
    reset_filter(t);
    t.reg = R(0); t.bfailed = 0;  // resetreg

    int rtn = 0, infeas_out = infeas_in, line_failed_out = 1;
    R cost_prev = t.cost;
    int iter = 0, bp_no_upd_count = 0;
    const int bp_no_upd_count_max = 20;
    int trace_n = 0;
    for (iter = 0; iter < cfg.iter_max; iter++) {
        int n_bwd = 0;
        while (true) {  // ddp.cpp:297-310
            backward_pass(t);
            n_bwd++;
            if (!t.bfailed) break;
            if (t.reg == R(24) && t.bfailed) bp_no_upd_count++;
            else bp_no_upd_count = 0;
            if (bp_no_upd_count > bp_no_upd_count_max) break;
        }
        forward_pass(t);
        if (A.trace && b == 0 && (st == 1 || !A.two_stage) && trace_n < A.trace_cap) {
            FOR_LANES(lane) {
                if (lane == 0) {
                    double *tr = A.trace + (long long)trace_n * 12;
                    tr[0] = t.cost; tr[1] = t.costq; tr[2] = t.logcost; tr[3] = t.err; tr[4] = t.mu; tr[5] = t.reg;
                    tr[6] = t.stepsize; tr[7] = t.opterr; tr[8] = t.step; tr[9] = t.failed; tr[10] = n_bwd; tr[11] = 0;
                }
            }
            trace_n++;
        }
        // negative segment time, ddp.cpp:317-326
        Reg<int, 1> neg;
        FOR_LANES(lane) {
            int f = 0;
            for (int i = lane; i < N; i += 32) if (t.xu[(long long)i * 20 + 9] < R(0)) f = 1;
            neg(lane, 0) = f;
        }
        if (warp_any(neg, 0, lane_)) { rtn = -3; break; }
        const R cost_m2 = cost_prev;
        cost_prev = t.cost;
        if (rmax(t.opterr, t.mu) <= t.tol) break;  // ddp.cpp:335-338
        if (t.opterr <= R(0.2) * t.mu) {           // ddp.cpp:340-344
            t.mu = rmax(t.tol / R(10), rmin(R(0.2) * t.mu, rpow(t.mu, R(1.2))));
            reset_filter(t);
            t.reg = R(0); t.bfailed = 0;
        }
        if (!any_violation(t, R(2.0e-4), false)) {  // ddp.cpp:346-390 (hazard H8)
            if (cfg.zero_init) { infeas_out = 0; rtn = 2; break; }
            if (!cfg.zero_init && !cfg.line_init) {
                const R dc = t.cost - cost_m2;
                if ((dc * dc < cost_m2 * R(1.0e-2)) && t.opterr < R(5.0e1)) { rtn = 1; break; }
            }
            if (cfg.line_init) {
                const R dc = t.cost - cost_m2;
                if (dc * dc < cost_m2 * R(0.01)) { line_failed_out = 0; break; }
            }
        }
        if (bp_no_upd_count > bp_no_upd_count_max) { rtn = -4; break; }  // ddp.cpp:392-396
    }
    if (A.trace && b == 0 && (st == 1 || !A.two_stage) && A.trace_len) {
        FOR_LANES(lane) { if (lane == 0) *A.trace_len = trace_n; }
    }

    // ---- outputs (ddp.cpp:418-437) -----------------------------------------------------------------------
    const OutPtrs &O = A.out[A.two_stage ? st : 1];
    const bool carry = (A.two_stage && st == 0);
    FOR_LANES(lane) {
        if (lane == 0) {
            if (O.rtn) O.rtn[b] = rtn;
            if (O.infeas_out) O.infeas_out[b] = infeas_out;
            if (O.line_failed_out) O.line_failed_out[b] = line_failed_out;
            if (O.iters) O.iters[b] = iter;
            if (O.cost) O.cost[b] = (double)t.cost;
            if (O.stats) {
                long long *S = O.stats + (long long)b * 8;
                S[0] = t.n_bwd_sweeps; S[1] = t.n_bwd_knots; S[2] = t.n_fwd_trials; S[3] = t.n_fwd_knots;
                S[4] = t.cyc_bwd; S[5] = t.cyc_fwd; S[6] = ddp_clock() - t.cyc_t0; S[7] = 0;
            }
        }
        if (lane < 9 && O.x_final) O.x_final[(long long)b * 9 + lane] = (double)t.xu[(long long)N * 20 + 10 + lane];
    }
    for (int i = 0; i < N; i++) {
        FOR_LANES(lane) { if (lane < 20) sm[Lay::ZO + lane] = t.xu[(long long)i * 20 + lane]; }
        WARP_SYNC();
        const R T = sm[Lay::ZO + 9];
        Reg<R, 1> jk;
        FOR_LANES(lane) {
            jk(lane, 0) = lane < 9 ? quad_share<R>(0, sm + Lay::ZO, lane, T) : R(0);  // finalroll, ddp.cpp:1624-1634
            if (lane < 18) {
                // PolyCoeff row = [Ek_inv * x, u[0:9]] (ddp.cpp:814-823); index l*3+a
                const int l = lane / 3, a = lane % 3;
                R pc = sm[Lay::ZO + zidx(l, a)];
                if (l == 2) pc = pc * R(0.5);
                if (O.poly_coeff) O.poly_coeff[((long long)b * N + i) * 18 + lane] = (double)pc;
                // BezCoeff = (1/T) * Bezier control points (poly2bezFunc, ddp.cpp:799-812: the inverse of
                // poly2bez*t2tauMat is the Bezier table with column k scaled by T^k), re-laid [x*6,y*6,z*6]
                const int j = lane / 3;
                R accv = R(0), pw = R(1);
                DDP_UNROLL
                for (int k = 0; k < 6; k++) {
                    R ck = sm[Lay::ZO + zidx(k, a)];
                    if (k == 2) ck = ck * R(0.5);
                    accv += (tabs[j * 6 + k] * pw) * ((R(1) / T) * ck);
                    pw = pw * T;
                }
                const double bzv = (double)accv;
                if (O.bez_coeff) O.bez_coeff[((long long)b * N + i) * 18 + a * 6 + j] = bzv;
                if (carry) A.bez_tmp[((long long)b * N + i) * 18 + a * 6 + j] = bzv;
            }
            if (lane == 9) {
                if (O.poly_time) O.poly_time[(long long)b * N + i] = (double)T;
                if (carry) A.time_tmp[(long long)b * N + i] = (double)T;
            }
        }
        const R jsum = warp_sum(jk, 0, lane_);
        FOR_LANES(lane) { if (lane == 0 && O.jerk) O.jerk[(long long)b * N + i] = (double)jsum; }
        WARP_SYNC();
    }
    WARP_SYNC();  // stage-0 outputs (bez_tmp/time_tmp/rtn) are read by other lanes of this warp in stage 1
}

}  // namespace ddp
#endif
