// ipddp_solver.h -- warp-per-trajectory interior-point DDP (IPDDP) solve for sm_100a.
//
// One warp owns one trajectory for its whole solve: setup, initial rollout, the outer barrier loop,
// backward Riccati/IP sweeps, filter line search and output conversion all run inside one kernel
// without host round trips.  This file is the device-side equivalent of the reference's
//   global_planner/src/ddp_optimizer.cpp   ("ddp.cpp" below)
//     polyCurveGeneration :5-438, backwardpass :440-644, forwardpass :647-778, bez2polyFunc :782,
//     poly2bezFunc :799, computenextx :1062, computecminvo :1132, computeq :1294,
//     computeall :1309-1368 + :1455-1604, initialroll :1608, finalroll :1624, resetfilter :1636
// re-designed rather than ported.  The batch is heavy-tailed (a few trajectories need 10-15x the mean
// work), so the kernel time is the critical path of the longest solves: everything is organised to keep
// the per-knot dependent chain of ONE warp short.
//   * Two kinds of phases.  KNOT-PARALLEL phases (lane <-> knot, no communication between lanes):
//     `linearize` (everything of the backward pass that does not depend on the Riccati recursion: constraint
//     values, interior-point weights, the constraint + stage-cost part of the Hessian/gradient of Q),
//     `evaluate_trial` (slack/dual updates, fraction-to-boundary tests, barrier cost of a line-search
//     trial), `scan_constraints` (filter reset, feasibility count).  SEQUENTIAL phases (lane <-> matrix
//     column / state element): `riccati` (dynamics term, ten pivots in five rounds, gains, value backup) and
//     the closed-loop state rollout inside `forward_trial`.
//   * No Jacobian is ever materialised: every constraint row is (basis row beta_g(T)) (x) (direction n)
//     plus a d/dT entry, so c, J*v, J^T*w and J^T D J come from 15 "row groups" (6 position control
//     points, 5 velocity, 4 acceleration) and the polytope planes.  J^T D J = sum_g (beta_g beta_g^T) (x) M_g
//     with 3x3 blocks M_g; only the 126 + 38 distinct entries are formed.
//   * The 19x19 Hessian of Q in z = [u(10); x(9)] plus the gradient as 20th row/column is held one column
//     per lane; the ten pivots of the u block are taken two per round as a unit-lower L D L^T (2 x 2 pivot block
//     broadcast by shuffle, both reciprocals side by side, multipliers through shared memory, one rank-2 update)
//     and leave V_xx, V_x in the trailing block (a Schur complement); the gains come from one back-substitution per lane.
//   * Per-knot state streams through a per-warp workspace slot in global memory (row arrays stored
//     [32-knot block][row slot][32] so lane <-> knot accesses coalesce and a knot's rows are 32 elements apart); c and the
//     slack gains ks,Ks,ky,Ky are recomputed on the fly instead of being stored.
//   * Data movement (round 2): whole tiles travel by bulk copies issued by ONE lane (cp.async.bulk + mbarrier: the
//     linearisation record of a knot into the Riccati recursion, gains + old point of a knot into the line search's state
//     recursion); the slack rows of the row loops travel by per-lane cp.async four rows ahead (RowStream).  Neither ties up
//     a scoreboard or a register; what each costs and what was measured slower is in profiles/r2a_kernel_ab_log.md.
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
    const int32_t *nknots;       // [B] knots of each trajectory (1 .. N), or null: all N
    double max_vel, max_acc;
    StageCfg cfg[2];
    int coop;          // 1: idle warps of a CTA help its remaining solves (tail balancing)
    int two_stage;     // 0: cfg[0] only, outputs -> out[1]; 1: stage 0 (out[0] optional) then stage 1
    OutPtrs out[2];
    double *bez_tmp, *time_tmp;  // [B][N][18], [B][N] scratch carrying stage 0 -> stage 1
    void *ws;                    // workspace, ws_stride elements of Real per warp slot
    long long ws_stride;
    int fcap;                    // filter capacity per slot
    unsigned int *counter;       // [0] work queue, [1] jobs posted, [2] units run by CTA helpers, [3] trajectories finished,
                                 // [4] warps of fully idle CTAs, [5] speculative line searches posted, [6] remote trials run,
                                 // [7] speculative backward sweeps claimed, [8] of which used
    int gspec;                   // 1: warps of idle CTAs run line-search trials of the remaining solves ("Speculative line search")
    int spec;                    // 1: ... and the backward sweep a failing line search would need ("Speculative backward sweep")
    void *gboards;               // GBoard<R>[GSPEC_BOARDS * slots]
    unsigned long long *gwords;  // their claim words, [GSPEC_BOARDS * slots], followed by a bitmap of the boards with unclaimed units
    double *trace;               // optional [cap][12] trace of trajectory 0 (last stage), or NULL
    int trace_cap;
    int *trace_len;
};

// ---------------------------------------------------------------------------------------------
// Per-warp shared-memory scratch (elements of Real).  Independent of the number of planes: the
// knot-parallel phases read planes straight from global memory (each lane its own knot).
// ---------------------------------------------------------------------------------------------
#ifndef DDP_CPR_DEPTH
#define DDP_CPR_DEPTH 4   // 4: 2 x 96.5 KB of shared memory per SM, still the 196 KB carve-out (8 would take the L1 from 60 to 28 KB)
#endif
struct Lay {
    enum {
        XD = 0,     // desired terminal state (9)
        S1 = 10,    // value-function Hessian as left by the Schur complement, S1[a*10+b] = S[a][b]
        S2 = 100,   // V = (S + S^T)/2 (ddp.cpp:628), V[b*10+a], formed by the writer of S1
        VX = 190,   // V_x (9)
        MB = 200,   // riccati: (x, y') of every lane per pivot round, MB[(kb*20+lane)*2 + {0,1}]  (5 x 20 x 2)
        XT = 400,   // row T of every column, for the T column (20)
        XH = 420,   // fT[p] * (V fT)[p]  (9)
        KC = 430,   // gains of the current knot, KC[p*10+qc], qc 0 = ku, 1..9 = Ku columns
        ZN = 530,   // rollout broadcast: new [u(10); x(9)] (20)
        DX = 550,   // rollout broadcast: xnew - xold (9)
        FL = 560,   // filter decision (2)
        FTN = 564,  // riccati: fT (9) and segment time (1) of the knot about to be processed, staged one knot ahead
        RI = 574,   // (free: the L D L^T pivots need no scale table)
        MSC = 584,  // per-lane 3x3 blocks M_g of the linearisation: MSC[e*32+lane], e < 63.  While no linearisation is in flight
                    // the same area holds: Riccati - multiplier table [0, 200) and the record ring [224, 896); line search - knot
                    // ring [0, 360), job partials / feed-forward staging [512, 1152).
        // Behind MSC (the linearisation overwrites all of MSC): the warp's three mbarriers, initialised once per kernel launch,
        // and the phase parity each of them completes next (one word, bit b = barrier b).
        RB = 584 + 63 * 32,         // three mbarriers, 8 bytes each
        RP = 584 + 63 * 32 + 6,     // phase word (unsigned), inside the 8 elements reserved here for float and double
        // Slack-row staging of the row loops (cp.async, one element per lane and row): CPR_DEPTH rows of s, then of y.  Outside
        // MSC because the linearisation fills all of MSC while its row loop runs.  2 x 88.3 KB -> 2 x 96.5 KB per SM at depth 4: the
        // same shared-memory carve-out (196 KB), so the L1 keeps its size.
        CR = 584 + 63 * 32 + 8,
        TOTAL = 584 + 63 * 32 + 8 + 2 * DDP_CPR_DEPTH * 32
    };
};
enum { CPR_DEPTH = DDP_CPR_DEPTH };   // rows in flight ahead of the row being processed
DDP_HD int smem_elems_per_warp(int /*pm*/) { return Lay::TOTAL; }

// Linearisation record of one knot, what the Riccati recursion reads: the symmetric 20 x 20 matrix Hc packed as its lower triangle,
// row-major (210), then fT (9) and the segment time (1), padded to a multiple of 16 bytes in float and double.  One record is ONE
// bulk copy (cp.async.bulk) into the recursion's shared-memory ring; the full square used to cost twice the HBM traffic and
// ten 128-bit loads per lane and knot.
enum { HREC = 224 };
DDP_DEVICE constexpr int hpk(int r, int c) { return r >= c ? r * (r + 1) / 2 + c : c * (c + 1) / 2 + r; }
// Workspace slot layout (elements of Real).  Row arrays (s, y and their trial copies) are stored
// [32-knot block][row slot][32] (row_ofs), so lane <-> knot accesses coalesce and a knot's rows are 32 elements apart.
// Row slots (row_slot below): corridor row (control point j, plane k) -> j*PM+k; then 6 rows per v / a group; time.
struct WsLay {
    long long xu, xun, K, K2, kdx, aux, H, s, sn, y, yn, filt, total;
    int MCS, NP;
};
DDP_HD WsLay ws_layout(int N, int PM, int fcap) {
    WsLay w;
    w.MCS = 6 * PM + 55;
    w.NP = (N + 31) & ~31;
    long long o = 0;
    w.xu = o; o += (long long)(N + 1) * 20;
    w.xun = o; o += (long long)(N + 1) * 20;
    w.K = o; o += (long long)N * 100;
    w.K2 = o; o += (long long)N * 100;   // target of a speculative backward sweep ("Speculative backward sweep"); swapped with K when it is used
    w.kdx = o; o += (long long)N * 10;
    w.aux = o; o += (long long)N * 12;
    o = (o + 3) & ~3LL;   // H columns are read with 128-bit loads, the row arrays by 16-byte-aligned bulk copies (float: 4 elements)
    w.H = o; o += (long long)N * HREC;
    w.s = o; o += (long long)w.MCS * w.NP;
    w.sn = o; o += (long long)w.MCS * w.NP;
    w.y = o; o += (long long)w.MCS * w.NP;
    w.yn = o; o += (long long)w.MCS * w.NP;
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
DDP_DEVICE constexpr int zidx(int l, int a) { return l < 3 ? 10 + 3 * l + a : 3 * (l - 3) + a; }
// index of the unordered pair (la <= lb) of monomial orders, 0..20
DDP_DEVICE constexpr int pidx(int la, int lb) { return la * 6 + lb - (la * (la + 1)) / 2; }

template <class R> struct Traj {
    int N, PM, NP, MCS, lane_;
    const double *planes;
    const int32_t *nplanes;
    R *sm;
    const R *tab;  // [0..89] value table (chosen basis), [90..179] d/dT table
    R *xu, *xun, *K, *K2, *kdx, *aux, *H, *s, *sn, *y, *yn, *filt;
    int fcap;
    R max_vel, max_acc, w_snap, w_terminal, w_time, margin;
    int time_power, infeas, zero_init, line_init;
    R mu, tol, reg, reg_base, opterr, cost, costq, logcost, err, stepsize;
    int step, failed, bfailed, nfilter;
    int cmax_valid;         // cmax is the largest constraint value at the current iterate (from the accepted trial)
    R cmax;
    int lin_valid;          // t.H / t.aux hold the linearisation of the current iterate at the current mu
    R lin_emu, lin_ecy;     // its max |r| and max |c + y| (ddp.cpp:636-637)
    long long n_bwd_sweeps, n_bwd_knots, n_fwd_trials, n_fwd_knots;
    long long cyc_bwd, cyc_fwd, cyc_t0;  // clock64 accounting (0 in the emulation)
    long long cyc_ric, cyc_seq;          // of which: Riccati recursion, sequential state rollout
    void *board, *ctl;                   // JobBoard<R> of this warp / BlockCtl of the CTA (null: no cooperation)
    int wpb;
    // speculative line search over the warps of idle CTAs (null / 0: off)
    void *gb;                            // this slot's two GBoard<R>
    unsigned long long *gw;              // and their claim words
    unsigned int *gbits;                 // bitmap of open boards (bit = board index)
    int gindex;                          // index of this slot's first board
    unsigned int *gctr;                  // SolveArgs::counter
    R *ws_all;                           // workspace of slot 0
    long long ws_stride, off_xun, off_sn, off_yn, off_kdx;
    int gflip, minvo;
    // speculative backward sweep (see backward_pass)
#ifdef DDP_TRACE_CYCLES
    long long cyc_seg[4];                // diagnostic build: cycles of the recursion per segment of a knot
#endif
    int spec_on;                         // enabled (SolveArgs::spec)
    int spec_posted;                     // a sweep is on the slot's third board and has not been closed yet
    R spec_regadd;                       // the regularisation it was posted with
    void *sweep_b;                       // that board, until the sweep has been taken or seen stopped
};

// ---------------------------------------------------------------------------------------------
// Model pieces shared by the phases (all per-lane, no communication).
// ---------------------------------------------------------------------------------------------
// Powers of the segment time by repeated multiplication like the reference's Tkv (ddp.cpp:1148-1160).
template <class R> DDP_DEVICE void time_powers(R T, R *tp) {
    tp[0] = R(1); tp[1] = T; tp[2] = tp[1] * T; tp[3] = tp[2] * T; tp[4] = tp[3] * T; tp[5] = tp[4] * T;
}
// T-scaled basis row of group g: b[l] = tab[g][l] * T^(l-SHIFT-DER) * Ek_inv[l] (ddp.cpp:1148-1160, :1203-1209,
// :1240-1247; Ek_inv = {1,1,1/2} folded in, ddp.cpp:1138-1143).  SHIFT = 0/1/2 for position / velocity /
// acceleration control points, DER = 1 for the d/dT tables (ddp.cpp:1543-1560).
template <class R, int SHIFT, int DER> DDP_DEVICE void basis_row(const R *tab, int g, const R *tp, R *b) {
    DDP_UNROLL
    for (int l = 0; l < 6; l++) {
        const int k = l - SHIFT - DER;
        R v = tab[DER * 90 + g * 6 + l];
        if (k > 0) v = v * tp[k];
        if (l == 2) v = v * R(0.5);
        b[l] = v;
    }
}
// (row b) . (coefficients of axis a of vector v in z layout); terms below order LMIN are structurally zero.
template <class R, int LMIN> DDP_DEVICE R dot_axis(const R *b, const R *v, int a) {
    R acc = R(0);
    DDP_UNROLL
    for (int l = LMIN; l < 6; l++) acc += b[l] * v[zidx(l, 0) + a];
    return acc;
}

// basis_row with the shift chosen at run time (sd = SHIFT + DER in 0..3), so that ONE copy of the row code serves the
// position, velocity and acceleration groups: multiplying by 1 where the template version does not multiply is exact.
template <class R> DDP_DEVICE void basis_row_rt(const R *tabrow, int sd, const R *tp, R *b) {
    DDP_UNROLL
    for (int l = 0; l < 6; l++) {
        R pw = l >= 1 ? tp[l] : R(1);
        if (sd == 1) pw = l >= 2 ? tp[l - 1] : R(1);
        else if (sd == 2) pw = l >= 3 ? tp[l - 2] : R(1);
        else if (sd == 3) pw = l >= 4 ? tp[l - 3] : R(1);
        R v = tabrow[l] * pw;
        if (l == 2) v = v * R(0.5);
        b[l] = v;
    }
}
// Row groups: g < 6 position control points (one row per plane), 6..10 velocity, 11..14 acceleration control points
// (six rows each, visited +x,-x,+y,-y,+z,-z like ddp.cpp:1228-1272).
DDP_DEVICE int group_shift(int g) { return g < 6 ? 0 : (g < 11 ? 1 : 2); }
// Row slot of row r of group g, in VISIT order (the slot order is internal to the workspace): corridor rows g*PM + k,
// then the six rows of every velocity / acceleration group, then the time row.  The row loops visit every slot in this
// order (a polytope with fewer than PM planes skips the body on its unused slots), so the row visited next always sits
// ROW_STRIDE elements further on (row_ofs), which is what RowStream relies on.
DDP_DEVICE int row_slot(int g, int r, int PM) { return g < 6 ? g * PM + r : 6 * PM + 6 * (g - 6) + r; }
// Unrolling of the row loops (rows of a knot are independent but for a running product and two maxima, and one row is one
// long dependent fp64 chain): SOLO = the loops of a warp working alone (the bulk of a batch, eight warps per SM share the
// instruction cache), COOP = the loops of the job units that only run when warps are idle (the tail of a batch).
#ifndef DDP_ROW_UNROLL_SOLO
#define DDP_ROW_UNROLL_SOLO 1
#endif
#ifndef DDP_ROW_UNROLL_COOP
#define DDP_ROW_UNROLL_COOP 1
#endif
#define DDP_PRAGMA_(x) _Pragma(#x)
#define DDP_PRAGMA(x) DDP_PRAGMA_(x)
#if DDP_GPU
#define DDP_ROWLOOP_SOLO DDP_PRAGMA(unroll DDP_ROW_UNROLL_SOLO)
#define DDP_ROWLOOP_COOP DDP_PRAGMA(unroll DDP_ROW_UNROLL_COOP)
#else
#define DDP_ROWLOOP_SOLO
#define DDP_ROWLOOP_COOP
#endif
// Direction and offset of a velocity / acceleration row: n = +-e_a, d = -limit (ddp.cpp:1238, :1276).
template <class R> DDP_DEVICE void fixed_row(int r, R lim, R *n) {
    const int a = r >> 1;
    const R sg = (r & 1) ? R(-1) : R(1);
    n[0] = a == 0 ? sg : R(0); n[1] = a == 1 ? sg : R(0); n[2] = a == 2 ? sg : R(0); n[3] = -lim;
}

// Plane k of knot `pl` (pointer to that knot's P_max x 4 block).
template <class R> DDP_DEVICE void load_plane(const double *pl, int k, R *n) {
#if DDP_GPU   // a plane is four doubles at a 32-byte-aligned address: two 128-bit loads instead of four 64-bit ones (-3 % kernel time)
    const double2 a = reinterpret_cast<const double2 *>(pl)[2 * k], b = reinterpret_cast<const double2 *>(pl)[2 * k + 1];
    n[0] = (R)a.x; n[1] = (R)a.y; n[2] = (R)b.x; n[3] = (R)b.y;
#else
    n[0] = (R)pl[4 * k]; n[1] = (R)pl[4 * k + 1]; n[2] = (R)pl[4 * k + 2]; n[3] = (R)pl[4 * k + 3];
#endif
}

// F, G of the segment dynamics x+ = (F (x) I3) x + (G (x) I3) u[0:9], ddp.cpp:862-871.  fg[o*6+l].
template <class R> DDP_DEVICE void fg_matrix(const R *tp, R *fg) {
    fg[0] = R(1); fg[1] = tp[1]; fg[2] = tp[2] / R(2); fg[3] = tp[3]; fg[4] = tp[4]; fg[5] = tp[5];
    fg[6] = R(0); fg[7] = R(1); fg[8] = tp[1]; fg[9] = R(3) * tp[2]; fg[10] = R(4) * tp[3]; fg[11] = R(5) * tp[4];
    fg[12] = R(0); fg[13] = R(0); fg[14] = R(1); fg[15] = R(6) * tp[1]; fg[16] = R(12) * tp[2]; fg[17] = R(20) * tp[3];
}
// Row o of [F | G] by selects: indexing fg[] with a lane-dependent o would put the array into local memory.
template <class R> DDP_DEVICE void fg_row(const R *fg, int o, R *c) {
    DDP_UNROLL
    for (int b = 0; b < 6; b++) c[b] = o == 0 ? fg[b] : (o == 1 ? fg[6 + b] : fg[12 + b]);
}
// fT = d x+/dT = (F' (x) I) x + (G' (x) I) u, ddp.cpp:930-935 and :1332.
template <class R> DDP_DEVICE void ft_vector(const R *tp, const R *z, R *fT) {
    const R fp[18] = {R(0), R(1), tp[1], R(3) * tp[2], R(4) * tp[3], R(5) * tp[4],
                      R(0), R(0), R(1), R(6) * tp[1], R(12) * tp[2], R(20) * tp[3],
                      R(0), R(0), R(0), R(6), R(24) * tp[1], R(60) * tp[2]};
    DDP_UNROLL
    for (int o = 0; o < 3; o++) {
        DDP_UNROLL
        for (int a = 0; a < 3; a++) {
            R s1 = R(0), s2 = R(0);
            DDP_UNROLL
            for (int b = 0; b < 3; b++) {
                if (b > o) s1 += fp[o * 6 + b] * z[10 + 3 * b + a];
                s2 += fp[o * 6 + 3 + b] * z[3 * b + a];
            }
            fT[3 * o + a] = s1 + s2;
        }
    }
}
// Jerk-cost matrices R, R', R'' (3x3, index 3*i+j), ddp.cpp:991-999.
template <class R> DDP_DEVICE void rmat(int which, const R *tp, R *m) {
    if (which == 0) {
        m[0] = R(36) * tp[1]; m[1] = m[3] = R(72) * tp[2]; m[2] = m[6] = R(120) * tp[3];
        m[4] = R(192) * tp[3]; m[5] = m[7] = R(360) * tp[4]; m[8] = R(720) * tp[5];
    } else if (which == 1) {
        m[0] = R(36); m[1] = m[3] = R(144) * tp[1]; m[2] = m[6] = R(360) * tp[2];
        m[4] = R(576) * tp[2]; m[5] = m[7] = R(1440) * tp[3]; m[8] = R(3600) * tp[4];
    } else {
        m[0] = R(0); m[1] = m[3] = R(144); m[2] = m[6] = R(720) * tp[1];
        m[4] = R(1152) * tp[1]; m[5] = m[7] = R(4320) * tp[2]; m[8] = R(14400) * tp[3];
    }
}
// (M (x) I3) u for the nine high-order coefficients u[3*i+a].
template <class R> DDP_DEVICE void rmat_times_u(const R *m, const R *u, R *out) {
    DDP_UNROLL
    for (int i = 0; i < 3; i++) {
        DDP_UNROLL
        for (int a = 0; a < 3; a++) {
            R acc = R(0);
            DDP_UNROLL
            for (int j = 0; j < 3; j++) acc += m[i * 3 + j] * u[3 * j + a];
            out[3 * i + a] = acc;
        }
    }
}
template <class R> DDP_DEVICE R dot9(const R *a, const R *b) {
    R acc = R(0);
    DDP_UNROLL
    for (int c = 0; c < 9; c++) acc += a[c] * b[c];
    return acc;
}
// Stage cost q(x,u), ddp.cpp:1294-1305.
template <class R> DDP_DEVICE R stage_cost(const Traj<R> &t, const R *tp, const R *u) {
    R m[9], mu9[9];
    rmat<R>(0, tp, m);
    rmat_times_u(m, u, mu9);
    const R T = tp[1];
    const R tterm = t.time_power == 2 ? R(0.5) * T * t.w_time * T : R(0.5) * t.w_time * T;
    return R(0.5) * t.w_snap * dot9(u, mu9) + tterm;
}

// Hot-loop view of the trajectory: plain by-value locals.  Traj<> is passed by reference to the out-of-line
// phases, so every t.field read inside a loop is a local-memory load that has to be repeated after each
// store (profiles/r1b); the phases copy what they need into a RowCtx once.
template <class R> struct RowCtx {
    const R *DDP_RESTRICT s;
    const R *DDP_RESTRICT y;
    R *DDP_RESTRICT sn;
    R *DDP_RESTRICT yn;
    const R *tab;
    const double *planes;
    const int32_t *nplanes;
    long long NP;
    int PM, MCS, infeas;
    R mu, margin, max_vel, max_acc;
};
// Offset of (row slot, knot i) in a row array.  Layout [block of 32 knots][row slot][32 knots]: lane <-> knot accesses coalesce,
// the row visited next sits 32 elements further on, and ALL rows of a 32-knot block are contiguous (MCS x 32 elements), so that
// the line search's bulk copies (RowRing) move any number of consecutive rows with one instruction.
DDP_DEVICE long long row_ofs(int MCS, int slot, int i) { return ((long long)(i >> 5) * MCS + slot) * 32 + (i & 31); }
enum { ROW_STRIDE = 32 };   // elements between consecutive row slots of a knot
template <class R> DDP_DEVICE RowCtx<R> row_ctx(const Traj<R> &t) {
    RowCtx<R> c;
    c.s = as_global(t.s); c.y = as_global(t.y); c.sn = as_global(t.sn); c.yn = as_global(t.yn); c.tab = t.tab;
    c.planes = as_global(t.planes); c.nplanes = as_global(t.nplanes);
    c.NP = t.NP; c.PM = t.PM; c.MCS = t.MCS; c.infeas = t.infeas; c.mu = t.mu; c.margin = t.margin; c.max_vel = t.max_vel; c.max_acc = t.max_acc;
    return c;
}

// A RowCtx that went through memory (job board) lost what the compiler knew about its pointers.
template <class R> DDP_DEVICE RowCtx<R> row_global(RowCtx<R> c) {
    c.s = as_global(c.s); c.y = as_global(c.y); c.sn = as_global(c.sn); c.yn = as_global(c.yn);
    c.planes = as_global(c.planes); c.nplanes = as_global(c.nplanes);
    return c;
}

// Slack rows of one knot through shared memory by cp.async (LDGSTS), one element per lane and row, CPR_DEPTH rows ahead of their
// use.  The row slots of a knot are visited in storage order (row_ofs: 32 elements apart), so the row wanted next is always
// one pointer step away.  Why not registers: every load of a warp that is in flight on the same scoreboard has to land before
// the oldest can be consumed, so a register pipeline k rows deep exposes the latency of the row issued LAST (measured: 181 ms
// against 164 ms, profiles/r2g); cp.async completes through its own group counter and costs no scoreboard.  Why not TMA here: a
// cp.async.bulk ring of 12 rows (one elected lane, an mbarrier wait and a warp synchronisation per batch of 4 rows) was built and
// measured SLOWER than plain loads (172.8 ms against 164.4 ms at B = 4096, 505 against 461 ms at 16 k; profiles/r2e): its extra
// code pushed the kernel over the instruction-cache cliff.  Each lane reads only what it copied itself, so cp.async.wait_group
// is all the synchronisation there is.
template <class R> struct RowStream {
    const R *ps, *py;   // next row to fetch (this lane's knot)
    R *rs;              // this lane's column of the staging area: rs[d * 32] = s, rs[(CPR_DEPTH + d) * 32] = y of ring slot d
    int slot, infeas;
};
template <class R> DDP_DEVICE void row_stream_fetch(RowStream<R> &q, int d) {
    cp_async<sizeof(R)>(q.rs + d * 32, q.ps);
    if (q.infeas) cp_async<sizeof(R)>(q.rs + (CPR_DEPTH + d) * 32, q.py);
    cp_commit();
    q.ps += ROW_STRIDE; q.py += ROW_STRIDE;
}
template <class R> DDP_DEVICE void row_stream_start(RowStream<R> &q, const RowCtx<R> &t, R *sm, int i, int lane, int slot0 = 0) {
    q.ps = t.s + row_ofs(t.MCS, slot0, i); q.py = t.y + row_ofs(t.MCS, slot0, i);
    q.rs = sm + Lay::CR + lane; q.slot = 0; q.infeas = t.infeas;
    DDP_UNROLL
    for (int d = 0; d < CPR_DEPTH; d++) row_stream_fetch(q, d);
}
// Slack (and dual slack) of the next row in visit order; refills the slot it frees with the row CPR_DEPTH further on.
template <class R> DDP_DEVICE void row_stream_next(RowStream<R> &q, R &sv, R &yv) {
    cp_wait<CPR_DEPTH - 1>();
    sv = q.rs[q.slot * 32];
    yv = q.infeas ? q.rs[(CPR_DEPTH + q.slot) * 32] : R(1);
    row_stream_fetch(q, q.slot);
    q.slot = q.slot + 1 == CPR_DEPTH ? 0 : q.slot + 1;
}

// Visit every constraint row of knot i at the point z (constraint VALUES only; computecminvo,
// ddp.cpp:1132-1285): f(row slot, c).
template <class R, class F> DDP_DEVICE void visit_rows(const RowCtx<R> &t, int i, const R *z, F &&f) {
    R tp[6];
    time_powers(z[9], tp);
    const int P = t.nplanes[i];
    const double *pl = t.planes + (long long)i * t.PM * 4;
    const R *tab = as_shared(t.tab);
    DDP_NOUNROLL
    for (int g = 0; g < 15; g++) {
        R b[6], cp[3];
        basis_row_rt(tab + g * 6, group_shift(g), tp, b);
        DDP_UNROLL
        for (int a = 0; a < 3; a++) cp[a] = dot_axis<R, 0>(b, z, a);
        const int nr = g < 6 ? P : 6;
        const R lim = g < 11 ? t.max_vel : t.max_acc;
        DDP_NOUNROLL
        for (int r = 0; r < nr; r++) {
            R n[4];
            if (g < 6) load_plane(pl, r, n);
            else fixed_row(r, lim, n);
            f(row_slot(g, r, t.PM), ((n[0] * cp[0] + n[1] * cp[1]) + n[2] * cp[2]) + n[3] - t.margin);
        }
    }
    f(6 * t.PM + 54, -z[9] + R(0.3) - t.margin);
}

// Interior-point weights of one row (ddp.cpp:535-541 infeasible / :583-590 feasible):
//   infeasible: D = s/y, r = s y - mu, tv2 = (s (c+y) - r)/y ;  feasible: D = s/c, r = s c + mu, tv2 = r/c
//   the row adds  sgn D a a^T  to the Hessian and  (s + sgn tv2) a  to the gradient (sgn = +1 / -1).
template <class R>
DDP_DEVICE void row_weights(int infeas, R mu, R sgn, R c, R sv, R yv, R &Ds, R &gw, R &emu, R &ecy) {
    R D, r, tv2;
    if (infeas) {
        const R yinv = rrcp(yv);
        r = sv * yv - mu; D = sv * yinv; tv2 = yinv * (sv * (c + yv) - r);
        ecy = amax(ecy, rabs(c + yv));
    } else {
        const R cinv = rrcp(c);
        r = sv * c + mu; D = sv * cinv; tv2 = cinv * r;
    }
    emu = amax(emu, rabs(r));
    Ds = sgn * D; gw = sv + sgn * tv2;
}

// Second pass of the linearisation, one group: h_a[(l,l')] += beta[l] beta[l'] M_g[a][a] for the three axes.
template <class R, int SHIFT>
DDP_DEVICE void lin_diag_group(const R *tab, int g, const R *tp, R m0, R m1, R m2, R *h0, R *h1, R *h2) {
    R b[6];
    basis_row<R, SHIFT, 0>(tab, g, tp, b);
    DDP_UNROLL
    for (int lb = SHIFT; lb < 6; lb++) {
        const R s0 = b[lb] * m0, s1 = b[lb] * m1, s2 = b[lb] * m2;
        DDP_UNROLL
        for (int la = SHIFT; la <= lb; la++) {
            h0[pidx(la, lb)] += b[la] * s0; h1[pidx(la, lb)] += b[la] * s1; h2[pidx(la, lb)] += b[la] * s2;
        }
    }
}

// =============================================================================================
// Intra-CTA cooperation.  A batch is heavy-tailed: a few solves run stage 1 to iter_max with ~5 full rollouts per
// iteration and take 10-15x the mean, so once the work queue is empty the kernel time is the latency of the last
// few solves, each on ONE warp (tools/cycle_report.py: 0.87 of the kernel at B = 4096).  A warp that finds the
// queue empty therefore stays as a HELPER of the warps of its CTA that still own a trajectory: the owner posts the
// knot-parallel phases as jobs of independent units on a board in shared memory (linearisation: one unit per 32
// knots; line-search rows of a 32-knot block: four units of row groups), owner and helpers claim units with a CAS
// on one packed word and the owner reduces the per-unit results in unit order, so the arithmetic - and the result,
// bit for bit - is the same with or without helpers.
// =============================================================================================
enum { JOB_LIN = 1, JOB_ROWS = 2, JOB_MAX_UNITS = 16 };
// Idle warps poll the boards with an exponentially growing nanosleep between polls (profiles/r1h: with a fixed 400-500 ns the
// polling loops executed as many warp instructions as the solves themselves and competed with the tail's owners for issue slots).
#ifndef DDP_IDLE_NS_MIN
#define DDP_IDLE_NS_MIN 250u
#endif
#ifndef DDP_IDLE_NS_MAX
#define DDP_IDLE_NS_MAX 4000u
#endif
template <class R> struct JobCtx {   // everything a unit needs; copied to registers by whoever runs the unit
    RowCtx<R> row;
    const R *xu, *xun, *K, *kdx;
    R *H, *aux;
    int N, base, time_power, type;
    R w_snap, w_time, alpha, tau;
};
template <class R> struct JobBoard {
    unsigned long long word;   // (job sequence number << 32) | (units << 16) | next unclaimed unit
    int done;                  // units completed
    int owner_seq;             // owner only: last sequence number used
    JobCtx<R> ctx;
    R res[JOB_MAX_UNITS][2];   // JOB_LIN: per-unit max |r|, max |c + y|
    R *part;                   // JOB_ROWS: [4][4][32] per-unit, per-lane {stage cost, sum log, |c + y|_1, max c} and
    int *badl;                 //           [4][32] per-lane first failing knot, both in the OWNER's MSC scratch
};
struct BlockCtl {
    int active_owners;         // warps of the CTA that still pull trajectories from the queue
    int unit_running;          // a warp of this (idle) CTA is running a remote line-search unit: the others only serve its rows
    unsigned int *jobs_ctr;    // global counter of posted jobs (statistics)
};
template <class R> DDP_HD int coop_smem_bytes(int warps_per_block) {
    return (int)(sizeof(JobBoard<R>) * warps_per_block + sizeof(BlockCtl));
}

// =============================================================================================
// Backward pass, knot-parallel part (lane <-> knot): everything of ddp.cpp:476-590 that does not depend
// on the value function.  Per knot it writes the constraint + stage-cost part of the augmented Hessian
//   Hc = [ H  g ; g^T . ]  (20 x 20, z order [u(9), T, x(9), gradient]), fT = dx+/dT and the segment time as one record (HREC) of t.H.
// Returns per-lane maxima of |r| and |c+y| in errs (ddp.cpp:636-637).
// =============================================================================================
template <class R> DDP_DEVICE_NOINLINE void lin_unit(const JobCtx<R> *cp_, int u, R *smw, int lane_, Reg<R, 2> &errs) {
    const JobCtx<R> c = *cp_;
    const RowCtx<R> t = row_global(c.row);
    smw = as_shared(smw);
    const R *tab = as_shared(t.tab);
    const int N = c.N, time_power = c.time_power;
    const R w_snap = c.w_snap, w_time = c.w_time;
    const R *DDP_RESTRICT xu = as_global(c.xu);
    R *DDP_RESTRICT Hout = as_global(c.H);
    (void)c.aux;
    const R sgn = t.infeas ? R(1) : R(-1);
    {
        const int base = 32 * (c.base + u);
        FOR_LANES(lane) {
            const int i = base + lane;
            if (i < N) {
                R emu = errs(lane, 0), ecy = errs(lane, 1);
                R z[19], tp[6];
                DDP_UNROLL
                for (int e = 0; e < 19; e++) z[e] = xu[(long long)i * 20 + e];
                time_powers(z[9], tp);
                const int P = t.nplanes[i];
                const double *pl = t.planes + (long long)i * t.PM * 4;
                R *msc = smw + Lay::MSC + lane;
                R accT[18], accG[18], tt = R(0), gt = R(0);
                DDP_UNROLL
                for (int e = 0; e < 18; e++) { accT[e] = R(0); accG[e] = R(0); }
                // ---- pass 1: rows -> weights -> per-group blocks.  One rolled loop over the 15 row groups and one over
                // the rows of a group (a single copy of the row code: the kernel is instruction-fetch bound,
                // profiles/r1c). ------------------------
                // Slack rows through the cp.async row stream (RowStream): every row slot of the knot is visited in storage order;
                // the slots a polytope with fewer than PM planes leaves unused are skipped by the body only.  The plane of the
                // next row is loaded one row ahead.
                RowStream<R> rows_in;
                row_stream_start(rows_in, t, smw, i, lane);
                R n_n[4] = {R(0), R(0), R(0), R(0)};
                if (P > 0) load_plane(pl, 0, n_n);
                DDP_NOUNROLL
                for (int g = 0; g < 15; g++) {
                    const int shift = group_shift(g), nr = g < 6 ? P : 6, nrw = g < 6 ? t.PM : 6;
                    const R lim = g < 11 ? t.max_vel : t.max_acc;
                    R b[6], bd[6], cp[3], cd[3];
                    basis_row_rt(tab + g * 6, shift, tp, b);
                    basis_row_rt(tab + 90 + g * 6, shift + 1, tp, bd);
                    DDP_UNROLL
                    for (int a = 0; a < 3; a++) { cp[a] = dot_axis<R, 0>(b, z, a); cd[a] = dot_axis<R, 0>(bd, z, a); }
                    R M[6] = {R(0), R(0), R(0), R(0), R(0), R(0)}, wv[3] = {R(0), R(0), R(0)}, gv[3] = {R(0), R(0), R(0)};
                    DDP_ROWLOOP_COOP
                    for (int r = 0; r < nrw; r++) {
                        R sv, yv;
                        row_stream_next(rows_in, sv, yv);
                        if (r < nr) {
                            R n[4];
                            if (g < 6) { n[0] = n_n[0]; n[1] = n_n[1]; n[2] = n_n[2]; n[3] = n_n[3]; }
                            else fixed_row(r, lim, n);
                            if (g < 6 && r + 1 < nr) load_plane(pl, r + 1, n_n);
                            const R c = ((n[0] * cp[0] + n[1] * cp[1]) + n[2] * cp[2]) + n[3] - t.margin;
                            const R tc = (n[0] * cd[0] + n[1] * cd[1]) + n[2] * cd[2];
                            R Ds, gw;
                            row_weights(t.infeas, t.mu, sgn, c, sv, yv, Ds, gw, emu, ecy);
                            M[0] += Ds * (n[0] * n[0]); M[1] += Ds * (n[0] * n[1]); M[2] += Ds * (n[0] * n[2]);
                            M[3] += Ds * (n[1] * n[1]); M[4] += Ds * (n[1] * n[2]); M[5] += Ds * (n[2] * n[2]);
                            const R dt = Ds * tc, gtc = gw * tc;
                            wv[0] += dt * n[0]; wv[1] += dt * n[1]; wv[2] += dt * n[2];
                            gv[0] += gw * n[0]; gv[1] += gw * n[1]; gv[2] += gw * n[2];
                            tt += dt * tc; gt += gtc;
                        }
                    }
                    if (g + 1 < 6 && P > 0) load_plane(pl, 0, n_n);   // first plane of the next position group
                    if (g < 6) {
                        DDP_UNROLL
                        for (int e = 0; e < 6; e++) msc[(g * 6 + e) * 32] = M[e];
                    } else {   // +-e_a rows: M is diagonal
                        msc[(36 + 3 * (g - 6)) * 32] = M[0]; msc[(37 + 3 * (g - 6)) * 32] = M[3]; msc[(38 + 3 * (g - 6)) * 32] = M[5];
                    }
                    DDP_UNROLL
                    for (int l = 0; l < 6; l++) {
                        DDP_UNROLL
                        for (int a = 0; a < 3; a++) { accT[l * 3 + a] += b[l] * wv[a]; accG[l * 3 + a] += b[l] * gv[a]; }
                    }
                }
                {   // the time row -T + 0.3 <= 0 (ddp.cpp:1279): Jacobian -1 in the T entry only; its slack is the stream's next row
                    R sv, yv, Ds, gw;
                    row_stream_next(rows_in, sv, yv);
                    row_weights(t.infeas, t.mu, sgn, -z[9] + R(0.3) - t.margin, sv, yv, Ds, gw, emu, ecy);
                    tt += Ds; gt -= gw;
                    cp_wait<0>();   // the stream's look-ahead copies (rows past the block) land before the staging area is reused
                }
                errs(lane, 0) = emu; errs(lane, 1) = ecy;
                // ---- stage cost (ddp.cpp:1338-1368): quu = w [R (x) I, R'u; (R'u)^T, .], qu = w [R u; .] --------
                R rm[9], Ru[9], Rpu[9], Rppu[9];
                rmat<R>(0, tp, rm);
                rmat_times_u(rm, z, Ru);
                {
                    R r1[9], r2[9];
                    rmat<R>(1, tp, r1); rmat_times_u(r1, z, Rpu);
                    rmat<R>(2, tp, r2); rmat_times_u(r2, z, Rppu);
                }
                const R uRpu = dot9(z, Rpu), uRppu = dot9(z, Rppu);
                R quT, quuTT;
                if (time_power == 2) { quT = w_time * z[9] + R(0.5) * w_snap * uRpu; quuTT = w_time + R(0.5) * w_snap * uRppu; }
                else { quT = R(0.5) * w_time + R(0.5) * w_snap * uRpu; quuTT = R(0.5) * w_snap * uRppu; }
                // ---- write row/column T and the gradient -------------------------------------------------------
                R *Hi = Hout + (long long)i * HREC;
                DDP_UNROLL
                for (int l = 0; l < 6; l++) {
                    DDP_UNROLL
                    for (int a = 0; a < 3; a++) {
                        const int r = zidx(l, a);
                        R hT = accT[l * 3 + a], gr = accG[l * 3 + a];
                        if (l >= 3) { hT += w_snap * Rpu[r]; gr += w_snap * Ru[r]; }
                        Hi[hpk(r, 9)] = hT;
                        Hi[hpk(r, 19)] = gr;
                    }
                }
                Hi[hpk(9, 9)] = quuTT + tt;
                Hi[hpk(9, 19)] = quT + gt;
                Hi[hpk(19, 19)] = R(0);
                R fT[9];
                ft_vector(tp, z, fT);
                DDP_UNROLL
                for (int q = 0; q < 9; q++) Hi[210 + q] = fT[q];
                Hi[219] = z[9];   // the segment time rides along: the recursion needs nothing else of the iterate
                // ---- pass 2: H[(l,a),(l',a')] = sum_g beta_g[l] beta_g[l'] M_g[a][a'] ------------------------------
                {   // same-axis blocks: all 15 groups, plus the stage cost w R (x) I on the u coefficients
                    R h0[21], h1[21], h2[21];
                    DDP_UNROLL
                    for (int e = 0; e < 21; e++) { h0[e] = R(0); h1[e] = R(0); h2[e] = R(0); }
                    DDP_NOUNROLL
                    for (int g = 0; g < 6; g++)
                        lin_diag_group<R, 0>(tab, g, tp, msc[(g * 6 + 0) * 32], msc[(g * 6 + 3) * 32], msc[(g * 6 + 5) * 32], h0, h1, h2);
                    DDP_NOUNROLL
                    for (int g = 6; g < 11; g++)
                        lin_diag_group<R, 1>(tab, g, tp, msc[(36 + 3 * (g - 6)) * 32], msc[(37 + 3 * (g - 6)) * 32],
                                             msc[(38 + 3 * (g - 6)) * 32], h0, h1, h2);
                    DDP_NOUNROLL
                    for (int g = 11; g < 15; g++)
                        lin_diag_group<R, 2>(tab, g, tp, msc[(36 + 3 * (g - 6)) * 32], msc[(37 + 3 * (g - 6)) * 32],
                                             msc[(38 + 3 * (g - 6)) * 32], h0, h1, h2);
                    DDP_UNROLL
                    for (int lb = 0; lb < 6; lb++) {
                        DDP_UNROLL
                        for (int la = 0; la <= lb; la++) {
                            const int e = pidx(la, lb);
                            R v0 = h0[e], v1 = h1[e], v2 = h2[e];
                            if (la >= 3) { const R q = w_snap * rm[(la - 3) * 3 + (lb - 3)]; v0 += q; v1 += q; v2 += q; }
                            const int r = zidx(la, 0), c = zidx(lb, 0);
                            Hi[hpk(r, c)] = v0;
                            Hi[hpk(r + 1, c + 1)] = v1;
                            Hi[hpk(r + 2, c + 2)] = v2;
                        }
                    }
                }
                {   // cross-axis blocks (0,1), (0,2), (1,2): only the position groups have off-diagonal M_g
                    R h0[21], h1[21], h2[21];
                    DDP_UNROLL
                    for (int e = 0; e < 21; e++) { h0[e] = R(0); h1[e] = R(0); h2[e] = R(0); }
                    DDP_NOUNROLL
                    for (int g = 0; g < 6; g++)
                        lin_diag_group<R, 0>(tab, g, tp, msc[(g * 6 + 1) * 32], msc[(g * 6 + 2) * 32], msc[(g * 6 + 4) * 32], h0, h1, h2);
                    DDP_UNROLL
                    for (int lb = 0; lb < 6; lb++) {
                        DDP_UNROLL
                        for (int la = 0; la <= lb; la++) {
                            const int e = pidx(la, lb);
                            const int ra = zidx(la, 0), rb = zidx(lb, 0);
                            // axes (0,1): entries ((la,0),(lb,1)) and ((lb,0),(la,1)) and their transposes; likewise (0,2), (1,2)
                            Hi[hpk(ra, rb + 1)] = h0[e];
                            Hi[hpk(rb, ra + 1)] = h0[e];
                            Hi[hpk(ra, rb + 2)] = h1[e];
                            Hi[hpk(rb, ra + 2)] = h1[e];
                            Hi[hpk(ra + 1, rb + 2)] = h2[e];
                            Hi[hpk(rb + 1, ra + 2)] = h2[e];
                        }
                    }
                }
            }
        }
    }
}

// ---- job board protocol (GPU only; the CPU emulation runs every unit in the owner) --------------------------------
// One unit of the line-search rows (defined below): per-lane partials of units u0 .. u1-1 into part / badk.
template <class R>
DDP_DEVICE_NOINLINE void rows_unit(const JobCtx<R> *cp_, int u0, int u1, int lane_, Reg<R, 16> &part, Reg<int, 4> &badk, R *smw);

#if DDP_GPU
// Claim the next unit of the job currently posted on `b` (any job when want_seq == 0): returns the unit or -1.
template <class R> DDP_DEVICE int job_claim(JobBoard<R> *b, unsigned want_seq) {
    int u = -1;
    if ((threadIdx.x & 31) == 0) {
        unsigned long long w = *(volatile unsigned long long *)&b->word;
        while (true) {
            const unsigned sq = (unsigned)(w >> 32);
            const int n = (int)((w >> 16) & 0xffff), nx = (int)(w & 0xffff);
            if (sq == 0 || (want_seq != 0 && sq != want_seq) || nx >= n) { u = -1; break; }
            const unsigned long long old = atomicCAS(&b->word, w, w + 1);
            if (old == w) { u = nx; break; }
            w = old;
        }
    }
    return __shfl_sync(0xffffffffu, u, 0);
}
// Run unit u of the job on `b` with this warp's scratch and leave its result on the board.
template <class R> DDP_DEVICE_NOINLINE void job_run_unit(JobBoard<R> *b, int u, R *smw, int lane_) {
    if (*(volatile int *)&b->ctx.type == JOB_LIN) {
        Reg<R, 2> errs;
        errs(lane_, 0) = R(0); errs(lane_, 1) = R(0);
        lin_unit(&b->ctx, u, smw, lane_, errs);
        const R e0 = warp_max(errs, 0, lane_), e1 = warp_max(errs, 1, lane_);
        if (lane_ == 0) { b->res[u][0] = e0; b->res[u][1] = e1; }
    } else {
        Reg<R, 16> part;
        Reg<int, 4> badk;
        rows_unit(&b->ctx, u, u + 1, lane_, part, badk, smw);
        DDP_UNROLL
        for (int e = 0; e < 4; e++) {
            if (e == u) {
                R *pp = b->part + e * 128 + lane_;
                pp[0] = part(lane_, 4 * e); pp[32] = part(lane_, 4 * e + 1); pp[64] = part(lane_, 4 * e + 2); pp[96] = part(lane_, 4 * e + 3);
                b->badl[e * 32 + lane_] = badk(lane_, e);
            }
        }
    }
    fence_async_all();       // the slack rows this unit wrote are read by the owner's bulk copies (RowRing) in the next trial
    __threadfence_block();   // the unit's global and shared writes are visible to the CTA before it counts as done
    __syncwarp();
    if (lane_ == 0) atomicAdd(&b->done, 1);
}
// Owner: post a job of n units, work on it together with the helpers, return when every unit is done.
template <class R> DDP_DEVICE_NOINLINE void job_run(Traj<R> &t, const JobCtx<R> &ctx, int n) {
    const int lane_ = t.lane_;
    JobBoard<R> *b = (JobBoard<R> *)t.board;
    __threadfence_block();   // this warp's global writes (candidate point, gains) before the job becomes visible
    __syncwarp();
    unsigned seq = 0;
    if (lane_ == 0) {
        atomicAdd(((BlockCtl *)t.ctl)->jobs_ctr, 1u);
        seq = (unsigned)(++b->owner_seq);
        b->ctx = ctx;
        b->part = t.sm + Lay::MSC + 512;                 // behind forward_trial's knot ring [0, 480)
        b->badl = (int *)(t.sm + Lay::MSC + 1024);
        b->done = 0;
        __threadfence_block();
        *(volatile unsigned long long *)&b->word = ((unsigned long long)seq << 32) | ((unsigned long long)n << 16);
    }
    seq = __shfl_sync(0xffffffffu, seq, 0);
    while (true) {
        const int u = job_claim(b, seq);
        if (u < 0) break;
        job_run_unit(b, u, t.sm, lane_);
    }
    if (lane_ == 0) {
        while (*(volatile int *)&b->done < n) __nanosleep(100);
    }
    __syncwarp();
    __threadfence_block();
}
// A warp without a trajectory serves the boards of its CTA until no warp of the CTA owns one any more.
template <class R>
DDP_DEVICE_NOINLINE void helper_loop(JobBoard<R> *boards, BlockCtl *ctl, int wpb, int me, R *sm, int lane_, unsigned int *units_ctr) {
    unsigned idle_ns = DDP_IDLE_NS_MIN;   // parked, not spinning: the sleep doubles while there is nothing to do
    while (*(volatile int *)&ctl->active_owners > 0) {
        bool found = false;
        for (int w = 0; w < wpb; w++) {
            if (w == me) continue;
            const int u = job_claim(boards + w, 0u);
            if (u < 0) continue;
            found = true;
            job_run_unit(boards + w, u, sm, lane_);
            if (lane_ == 0) atomicAdd(units_ctr, 1u);
        }
        if (found) idle_ns = DDP_IDLE_NS_MIN;
        else { __nanosleep(idle_ns); if (idle_ns < DDP_IDLE_NS_MAX) idle_ns *= 2; }
    }
}
#endif

template <class R> DDP_DEVICE bool coop_has_helpers(const Traj<R> &t) {
#if DDP_GPU
    return t.ctl != nullptr && *(volatile int *)&((BlockCtl *)t.ctl)->active_owners < t.wpb;
#else
    (void)t;
    return false;
#endif
}

// Line-search rows of one 32-knot block: per-lane partials of the four units.  Alone (one call, the knot's data
// loaded once) or, when a warp of the CTA is idle, as a job of four units; the partials are the same either way.
template <class R> DDP_DEVICE void run_rows(Traj<R> &t, const JobCtx<R> &ctx, Reg<R, 16> &part, Reg<int, 4> &badk) {
    const int lane_ = t.lane_;
#if DDP_GPU
    if (coop_has_helpers(t)) {
        job_run(t, ctx, 4);
        JobBoard<R> *b = (JobBoard<R> *)t.board;
        DDP_UNROLL
        for (int e = 0; e < 4; e++) {
            const R *pp = b->part + e * 128 + lane_;
            part(lane_, 4 * e) = pp[0]; part(lane_, 4 * e + 1) = pp[32]; part(lane_, 4 * e + 2) = pp[64]; part(lane_, 4 * e + 3) = pp[96];
            badk(lane_, e) = b->badl[e * 32 + lane_];
        }
        __syncwarp();
        return;
    }
#endif
    rows_unit(&ctx, 0, 4, lane_, part, badk, t.sm);
}

// Linearisation of the whole trajectory as a job of one unit per 32 knots (idle warps of the CTA available).
template <class R> DDP_DEVICE_NOINLINE void linearize_coop(Traj<R> &t, R &emu_max, R &ecy_max) {
    const int lane_ = t.lane_;
    JobCtx<R> c;
    c.row = row_ctx(t);
    c.xu = t.xu; c.xun = t.xun; c.K = t.K; c.kdx = t.kdx; c.H = t.H; c.aux = t.aux;
    c.N = t.N; c.base = 0; c.time_power = t.time_power; c.type = JOB_LIN;
    c.w_snap = t.w_snap; c.w_time = t.w_time; c.alpha = R(0); c.tau = R(0);
    const int n = (t.N + 31) / 32;
    emu_max = R(0); ecy_max = R(0);
#if DDP_GPU
    if (n > 1 && n <= JOB_MAX_UNITS && coop_has_helpers(t)) {
        job_run(t, c, n);
        JobBoard<R> *b = (JobBoard<R> *)t.board;
        for (int u = 0; u < n; u++) { emu_max = amax(emu_max, b->res[u][0]); ecy_max = amax(ecy_max, b->res[u][1]); }
        __syncwarp();
        return;
    }
#endif
    Reg<R, 2> errs;
    FOR_LANES(lane) { errs(lane, 0) = R(0); errs(lane, 1) = R(0); }
    for (int u = 0; u < n; u++) lin_unit(&c, u, t.sm, lane_, errs);
    emu_max = warp_max(errs, 0, lane_);
    ecy_max = warp_max(errs, 1, lane_);
}

// The same linearisation for a warp working alone (the common case): lin_unit's arithmetic over all 32-knot chunks in
// one function, so that the bulk of a batch runs the small code it ran before cooperation existed.
template <class R> DDP_DEVICE_NOINLINE void linearize_solo(Traj<R> &tt_, Reg<R, 2> &errs) {
    const int lane_ = tt_.lane_;
    const RowCtx<R> t = row_ctx(tt_);
    const int N = tt_.N, time_power = tt_.time_power;
    const R w_snap = tt_.w_snap, w_time = tt_.w_time;
    const R *DDP_RESTRICT xu = as_global(tt_.xu);
    R *DDP_RESTRICT Hout = as_global(tt_.H);

    R *smw = as_shared(tt_.sm);
    const R *tab = as_shared(t.tab);
    const R sgn = t.infeas ? R(1) : R(-1);
    FOR_LANES(lane) { errs(lane, 0) = R(0); errs(lane, 1) = R(0); }
    for (int base = 0; base < N; base += 32) {
        FOR_LANES(lane) {
            const int i = base + lane;
            if (i < N) {
                R emu = errs(lane, 0), ecy = errs(lane, 1);
                R z[19], tp[6];
                DDP_UNROLL
                for (int e = 0; e < 19; e++) z[e] = xu[(long long)i * 20 + e];
                time_powers(z[9], tp);
                const int P = t.nplanes[i];
                const double *pl = t.planes + (long long)i * t.PM * 4;
                R *msc = smw + Lay::MSC + lane;
                R accT[18], accG[18], tt = R(0), gt = R(0);
                DDP_UNROLL
                for (int e = 0; e < 18; e++) { accT[e] = R(0); accG[e] = R(0); }
                // ---- pass 1: rows -> weights -> per-group blocks.  One rolled loop over the 15 row groups and one over
                // the rows of a group (a single copy of the row code: the kernel is instruction-fetch bound,
                // profiles/r1c); the slack of row r+1 is loaded while row r is processed. ------------------------
                // Slack rows through the cp.async row stream (RowStream): every row slot of the knot is visited in storage order;
                // the slots a polytope with fewer than PM planes leaves unused are skipped by the body only.  The plane of the
                // next row is loaded one row ahead.
                RowStream<R> rows_in;
                row_stream_start(rows_in, t, smw, i, lane);
                R n_n[4] = {R(0), R(0), R(0), R(0)};
                if (P > 0) load_plane(pl, 0, n_n);
                DDP_NOUNROLL
                for (int g = 0; g < 15; g++) {
                    const int shift = group_shift(g), nr = g < 6 ? P : 6, nrw = g < 6 ? t.PM : 6;
                    const R lim = g < 11 ? t.max_vel : t.max_acc;
                    R b[6], bd[6], cp[3], cd[3];
                    basis_row_rt(tab + g * 6, shift, tp, b);
                    basis_row_rt(tab + 90 + g * 6, shift + 1, tp, bd);
                    DDP_UNROLL
                    for (int a = 0; a < 3; a++) { cp[a] = dot_axis<R, 0>(b, z, a); cd[a] = dot_axis<R, 0>(bd, z, a); }
                    R M[6] = {R(0), R(0), R(0), R(0), R(0), R(0)}, wv[3] = {R(0), R(0), R(0)}, gv[3] = {R(0), R(0), R(0)};
                    DDP_ROWLOOP_SOLO
                    for (int r = 0; r < nrw; r++) {
                        R sv, yv;
                        row_stream_next(rows_in, sv, yv);
                        if (r < nr) {
                            R n[4];
                            if (g < 6) { n[0] = n_n[0]; n[1] = n_n[1]; n[2] = n_n[2]; n[3] = n_n[3]; }
                            else fixed_row(r, lim, n);
                            if (g < 6 && r + 1 < nr) load_plane(pl, r + 1, n_n);
                            const R c = ((n[0] * cp[0] + n[1] * cp[1]) + n[2] * cp[2]) + n[3] - t.margin;
                            const R tc = (n[0] * cd[0] + n[1] * cd[1]) + n[2] * cd[2];
                            R Ds, gw;
                            row_weights(t.infeas, t.mu, sgn, c, sv, yv, Ds, gw, emu, ecy);
                            M[0] += Ds * (n[0] * n[0]); M[1] += Ds * (n[0] * n[1]); M[2] += Ds * (n[0] * n[2]);
                            M[3] += Ds * (n[1] * n[1]); M[4] += Ds * (n[1] * n[2]); M[5] += Ds * (n[2] * n[2]);
                            const R dt = Ds * tc, gtc = gw * tc;
                            wv[0] += dt * n[0]; wv[1] += dt * n[1]; wv[2] += dt * n[2];
                            gv[0] += gw * n[0]; gv[1] += gw * n[1]; gv[2] += gw * n[2];
                            tt += dt * tc; gt += gtc;
                        }
                    }
                    if (g + 1 < 6 && P > 0) load_plane(pl, 0, n_n);   // first plane of the next position group
                    if (g < 6) {
                        DDP_UNROLL
                        for (int e = 0; e < 6; e++) msc[(g * 6 + e) * 32] = M[e];
                    } else {   // +-e_a rows: M is diagonal
                        msc[(36 + 3 * (g - 6)) * 32] = M[0]; msc[(37 + 3 * (g - 6)) * 32] = M[3]; msc[(38 + 3 * (g - 6)) * 32] = M[5];
                    }
                    DDP_UNROLL
                    for (int l = 0; l < 6; l++) {
                        DDP_UNROLL
                        for (int a = 0; a < 3; a++) { accT[l * 3 + a] += b[l] * wv[a]; accG[l * 3 + a] += b[l] * gv[a]; }
                    }
                }
                {   // the time row -T + 0.3 <= 0 (ddp.cpp:1279): Jacobian -1 in the T entry only; its slack is the stream's next row
                    R sv, yv, Ds, gw;
                    row_stream_next(rows_in, sv, yv);
                    row_weights(t.infeas, t.mu, sgn, -z[9] + R(0.3) - t.margin, sv, yv, Ds, gw, emu, ecy);
                    tt += Ds; gt -= gw;
                    cp_wait<0>();   // the stream's look-ahead copies (rows past the block) land before the staging area is reused
                }
                errs(lane, 0) = emu; errs(lane, 1) = ecy;
                // ---- stage cost (ddp.cpp:1338-1368): quu = w [R (x) I, R'u; (R'u)^T, .], qu = w [R u; .] --------
                R rm[9], Ru[9], Rpu[9], Rppu[9];
                rmat<R>(0, tp, rm);
                rmat_times_u(rm, z, Ru);
                {
                    R r1[9], r2[9];
                    rmat<R>(1, tp, r1); rmat_times_u(r1, z, Rpu);
                    rmat<R>(2, tp, r2); rmat_times_u(r2, z, Rppu);
                }
                const R uRpu = dot9(z, Rpu), uRppu = dot9(z, Rppu);
                R quT, quuTT;
                if (time_power == 2) { quT = w_time * z[9] + R(0.5) * w_snap * uRpu; quuTT = w_time + R(0.5) * w_snap * uRppu; }
                else { quT = R(0.5) * w_time + R(0.5) * w_snap * uRpu; quuTT = R(0.5) * w_snap * uRppu; }
                // ---- write row/column T and the gradient -------------------------------------------------------
                R *Hi = Hout + (long long)i * HREC;
                DDP_UNROLL
                for (int l = 0; l < 6; l++) {
                    DDP_UNROLL
                    for (int a = 0; a < 3; a++) {
                        const int r = zidx(l, a);
                        R hT = accT[l * 3 + a], gr = accG[l * 3 + a];
                        if (l >= 3) { hT += w_snap * Rpu[r]; gr += w_snap * Ru[r]; }
                        Hi[hpk(r, 9)] = hT;
                        Hi[hpk(r, 19)] = gr;
                    }
                }
                Hi[hpk(9, 9)] = quuTT + tt;
                Hi[hpk(9, 19)] = quT + gt;
                Hi[hpk(19, 19)] = R(0);
                R fT[9];
                ft_vector(tp, z, fT);
                DDP_UNROLL
                for (int q = 0; q < 9; q++) Hi[210 + q] = fT[q];
                Hi[219] = z[9];   // the segment time rides along: the recursion needs nothing else of the iterate
                // ---- pass 2: H[(l,a),(l',a')] = sum_g beta_g[l] beta_g[l'] M_g[a][a'] ------------------------------
                {   // same-axis blocks: all 15 groups, plus the stage cost w R (x) I on the u coefficients
                    R h0[21], h1[21], h2[21];
                    DDP_UNROLL
                    for (int e = 0; e < 21; e++) { h0[e] = R(0); h1[e] = R(0); h2[e] = R(0); }
                    DDP_NOUNROLL
                    for (int g = 0; g < 6; g++)
                        lin_diag_group<R, 0>(tab, g, tp, msc[(g * 6 + 0) * 32], msc[(g * 6 + 3) * 32], msc[(g * 6 + 5) * 32], h0, h1, h2);
                    DDP_NOUNROLL
                    for (int g = 6; g < 11; g++)
                        lin_diag_group<R, 1>(tab, g, tp, msc[(36 + 3 * (g - 6)) * 32], msc[(37 + 3 * (g - 6)) * 32],
                                             msc[(38 + 3 * (g - 6)) * 32], h0, h1, h2);
                    DDP_NOUNROLL
                    for (int g = 11; g < 15; g++)
                        lin_diag_group<R, 2>(tab, g, tp, msc[(36 + 3 * (g - 6)) * 32], msc[(37 + 3 * (g - 6)) * 32],
                                             msc[(38 + 3 * (g - 6)) * 32], h0, h1, h2);
                    DDP_UNROLL
                    for (int lb = 0; lb < 6; lb++) {
                        DDP_UNROLL
                        for (int la = 0; la <= lb; la++) {
                            const int e = pidx(la, lb);
                            R v0 = h0[e], v1 = h1[e], v2 = h2[e];
                            if (la >= 3) { const R q = w_snap * rm[(la - 3) * 3 + (lb - 3)]; v0 += q; v1 += q; v2 += q; }
                            const int r = zidx(la, 0), c = zidx(lb, 0);
                            Hi[hpk(r, c)] = v0;
                            Hi[hpk(r + 1, c + 1)] = v1;
                            Hi[hpk(r + 2, c + 2)] = v2;
                        }
                    }
                }
                {   // cross-axis blocks (0,1), (0,2), (1,2): only the position groups have off-diagonal M_g
                    R h0[21], h1[21], h2[21];
                    DDP_UNROLL
                    for (int e = 0; e < 21; e++) { h0[e] = R(0); h1[e] = R(0); h2[e] = R(0); }
                    DDP_NOUNROLL
                    for (int g = 0; g < 6; g++)
                        lin_diag_group<R, 0>(tab, g, tp, msc[(g * 6 + 1) * 32], msc[(g * 6 + 2) * 32], msc[(g * 6 + 4) * 32], h0, h1, h2);
                    DDP_UNROLL
                    for (int lb = 0; lb < 6; lb++) {
                        DDP_UNROLL
                        for (int la = 0; la <= lb; la++) {
                            const int e = pidx(la, lb);
                            const int ra = zidx(la, 0), rb = zidx(lb, 0);
                            // axes (0,1): entries ((la,0),(lb,1)) and ((lb,0),(la,1)) and their transposes; likewise (0,2), (1,2)
                            Hi[hpk(ra, rb + 1)] = h0[e];
                            Hi[hpk(rb, ra + 1)] = h0[e];
                            Hi[hpk(ra, rb + 2)] = h1[e];
                            Hi[hpk(rb, ra + 2)] = h1[e];
                            Hi[hpk(ra + 1, rb + 2)] = h2[e];
                            Hi[hpk(rb + 1, ra + 2)] = h2[e];
                        }
                    }
                }
            }
        }
    }
}

// =============================================================================================
// Backward pass, sequential part (ddp.cpp:507-638): lane j < 20 owns column j of the augmented matrix
//   [ Quu Qux Qu ; Qxu Qxx Qx ; Qu^T Qx^T . ]   (z order [u(9), T, x(9), gradient]).
// Per knot: add the dynamics term A^T Vxx A, A^T Vx (A = [G (x) I | fT | F (x) I], ddp.cpp:508-520) to the
// linearised part, eliminate the ten u/T columns (Eigen::LLT, ddp.cpp:543/:592), back-substitute the gains
// [ku | Ku] (ddp.cpp:561-564 / :607-609) and keep the trailing block as Vxx, Vx (ddp.cpp:620-628).
// Returns false when a pivot is not positive (ddp.cpp:546-551 / :595-600).
// =============================================================================================
// `cancel` (speculative sweeps only, global memory): looked at once per knot; set -> the sweep stops like after a failed factorisation.
template <class R> DDP_DEVICE_NOINLINE bool riccati(Traj<R> &t, R regadd, Reg<R, 1> &errq, const volatile int *cancel = nullptr) {
    const int lane_ = t.lane_;
    R *sm = as_shared(t.sm);
    const int N = t.N;
    const R *DDP_RESTRICT Hin = as_global(t.H);
    const R *DDP_RESTRICT xu = as_global(t.xu);
    R *DDP_RESTRICT Kout = as_global(t.K);
    const R w_terminal = t.w_terminal;
    // terminal value function, ddp.cpp:1318-1323: Vx = P (x_N - x_d), Vxx = P = w_terminal I
    // Nothing inside the knot loop may touch local memory or wait on a global load: the linearisation records (HREC elements per
    // knot: packed Hessian, fT, segment time) arrive through a three-deep shared-memory ring filled by the copy engine, one
    // cp.async.bulk per knot issued by lane 0 three knots ahead, completion on the warp's mbarriers (phase word Lay::RP).
    // The ring sits in MSC behind the multiplier table [0, 200); MSC is idle while the recursion runs.
    R *tiles = sm + Lay::MSC + HREC;
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(sm + Lay::RB);
    unsigned rphase = *reinterpret_cast<volatile unsigned *>(sm + Lay::RP);
    const unsigned rec_bytes = (unsigned)(HREC * sizeof(R));
    int issued = 0;   // records handed to the copy engine: record k = knot N-1-k, ring slot k % 3
    Reg<R, 1> errl;
    FOR_LANES(lane) {
        errl(lane, 0) = R(0);
        for (int e = lane; e < 90; e += 32) {
            const R v = (e / 10 == e % 10 && e % 10 < 9) ? w_terminal : R(0);
            sm[Lay::S1 + e] = v; sm[Lay::S2 + e] = v;
        }
        if (lane < 9) sm[Lay::VX + lane] = w_terminal * (xu[(long long)N * 20 + 10 + lane] - sm[Lay::XD + lane]);
        fence_async_all();   // the records (st.global of this warp or of its helpers) and the last generic stores to MSC before the copies
    }
    WARP_SYNC();
    FOR_LANES(lane) {
        if (lane == 0) {
            for (int k = 0; k < 3 && k < N; k++) {
                mbar_expect_tx(bars + k, rec_bytes);
                bulk_g2s(tiles + k * HREC, Hin + (long long)(N - 1 - k) * HREC, rec_bytes, bars + k);
            }
        }
    }
    issued = N < 3 ? N : 3;
    Reg<R, 20> col;
    long long knots = 0;
    bool ok = true;
    for (int i = N - 1; i >= 0; i--) {
        if (cancel != nullptr && *cancel != 0) { ok = false; break; }
        const int slot = (int)(knots % 3);
        knots++;
#ifdef DDP_TRACE_CYCLES
        const long long pc0 = ddp_clock();
#endif
        mbar_wait(bars + slot, (rphase >> slot) & 1u);
        rphase ^= 1u << slot;
        const R *tile = tiles + slot * HREC;
        // uniform per-knot data: segment time, F/G, fT
        R tp[6], fg[18], fT[9];
        time_powers(tile[219], tp);
        fg_matrix(tp, fg);
        DDP_UNROLL
        for (int q = 0; q < 9; q++) fT[q] = tile[210 + q];
        FOR_LANES(lane) {
            if (lane < 20) {   // column `lane` of the symmetric matrix out of its packed lower triangle
                const int tri = lane * (lane + 1) / 2;
                DDP_UNROLL
                for (int r = 0; r < 20; r++) col(lane, r) = tile[r >= lane ? r * (r + 1) / 2 + lane : tri + r];
            } else {
                DDP_UNROLL
                for (int r = 0; r < 20; r++) col(lane, r) = R(0);
            }
            // Lanes 20 .. 28 ride along for the T-T entry: with tj = row (lane - 20) of V the dot product fT . tj below IS (V fT)[lane - 20],
            // in the instructions the other lanes need anyway (it used to be a divergent section of its own after this one).
            if (lane < 29 && lane != 9) {
                R tj[9];
                if (lane >= 19) {
                    const R *src = lane == 19 ? sm + Lay::VX : sm + Lay::S2 + (lane - 20) * 10;
                    DDP_UNROLL
                    for (int p = 0; p < 9; p++) tj[p] = src[p];
                } else {
                    const int lj = lane < 9 ? 3 + lane / 3 : (lane - 10) / 3, aj = lane < 9 ? lane % 3 : (lane - 10) % 3;
                    R f0 = fg[5], f1 = fg[11], f2 = fg[17];   // column lj of F|G by selects (a dynamic index would go through local memory)
                    DDP_UNROLL
                    for (int k = 4; k >= 0; k--) {
                        if (lj == k) { f0 = fg[k]; f1 = fg[6 + k]; f2 = fg[12 + k]; }
                    }
                    const R *V = sm + Lay::S2;   // V = (S + S^T)/2 (ddp.cpp:628), symmetrised by the writer below
                    DDP_UNROLL
                    for (int p = 0; p < 9; p++) {
                        const R v0 = V[aj * 10 + p], v1 = V[(3 + aj) * 10 + p], v2 = V[(6 + aj) * 10 + p];
                        tj[p] = (v0 * f0 + v1 * f1) + v2 * f2;
                    }
                    // gradient row of this column: (A e_j)^T Vx
                    col(lane, 19) += (f0 * sm[Lay::VX + aj] + f1 * sm[Lay::VX + 3 + aj]) + f2 * sm[Lay::VX + 6 + aj];
                }
                if (lane < 20) {
                    DDP_UNROLL
                    for (int l = 0; l < 6; l++) {
                        DDP_UNROLL
                        for (int a = 0; a < 3; a++)
                            col(lane, zidx(l, a)) += (fg[l] * tj[a] + fg[6 + l] * tj[3 + a]) + fg[12 + l] * tj[6 + a];
                    }
                }
                R hT = R(0);
                DDP_UNROLL
                for (int p = 0; p < 9; p++) hT += fT[p] * tj[p];
                if (lane < 20) {
                    col(lane, 9) += hT;
                    sm[Lay::XT + lane] = col(lane, 9);
                } else {
                    sm[Lay::XH + lane - 20] = tile[210 + lane - 20] * hT;   // fT[p] (V fT)[p]; fT[p] from shared memory: a dynamic register index would be local memory
                }
            }
        }
        WARP_SYNC();
        if (issued < N) {   // every lane has taken what it needs of this record: its ring slot gets the record three knots on
            FOR_LANES(lane) {
                if (lane == 0) {
                    mbar_expect_tx(bars + slot, rec_bytes);
                    bulk_g2s(tiles + slot * HREC, Hin + (long long)(N - 1 - issued) * HREC, rec_bytes, bars + slot);
                }
            }
            issued++;
        }
        FOR_LANES(lane) {
            // One straight-line block for every lane instead of a section for lane 9 and another for lane 19 (divergent sections run
            // one after the other and a lone warp pays the full latency of each): everybody forms the T-T sum and its column maxima,
            // lane 9 and lane 19 keep theirs.
            R hTT = R(0);
            DDP_UNROLL
            for (int p = 0; p < 9; p++) hTT += sm[Lay::XH + p];
            R m[5];   // |Qu|_inf, ddp.cpp:633: a tree instead of a chain of ten (the maximum does not depend on the order)
            DDP_UNROLL
            for (int r = 0; r < 5; r++) m[r] = amax(rabs(col(lane, 2 * r)), rabs(col(lane, 2 * r + 1)));
            const R e = amax(amax(errl(lane, 0), m[4]), amax(amax(m[0], m[1]), amax(m[2], m[3])));
            if (lane == 19) errl(lane, 0) = e;
            DDP_UNROLL
            for (int r = 0; r < 20; r++) {   // the T column is the T row of the others (the matrix is symmetric)
                if (r != 9) { const R v = sm[Lay::XT + r]; if (lane == 9) col(lane, r) = v; }
            }
            if (lane == 9) col(lane, 9) += hTT;
        }
#ifdef DDP_TRACE_CYCLES
        const long long pc1 = ddp_clock();
#endif
        // ---- the ten pivots of the u block, two per round ------------------------------------------------------
        // Eigen's LLT (ddp.cpp:543/:592) eliminates one column after the other; what the backup needs is the Schur complement
        // and K = -(Quu + rho I)^-1 [Qu | Qux], and the unit-lower L D L^T of the same ten pivots gives both with the same
        // stability and without square roots.  The recursion is latency-bound (one warp, profiles/r1h: 400 cycles per pivot in
        // broadcast -> reciprocal -> multiplier -> shared-memory round trip -> update), so two pivots share a round: the
        // 2 x 2 pivot block [a b; b c] is broadcast once, 1/a and 1/(a c - b^2) (pivot 2 = det / a) are formed side by side,
        // every lane takes both elimination steps on its own two entries (l1 = x / a, y' = y - b l1, l2 = y' / pivot2) and
        // the trailing matrix gets one rank-2 update.  Five rounds instead of ten, same pivots, same failure test
        // (a pivot <= 0, ddp.cpp:546-551 / :595-600).  After every round each lane shifts its column up by two rows, so
        // the pivot rows are always elements 0 and 1 and the rolled loop has static register indices.  Row p + r of the
        // matrix sits in element r; elements that would belong to rows >= 20 hold don't-care values that are never read.
        //   MB[(kb * 20 + lane) * 2 + {0, 1}] = (x, y') of the lane at round kb   (rows of the eliminated columns times D)
        //   MSC[(kb * 20 + lane) * 2 + {0, 1}] = (l1, l2)                         (unit-lower L: L[lane][2 kb], L[lane][2 kb + 1])
        // The pivot block of round kb + 1 is complete as soon as rows p + 2 and p + 3 have had round kb's update, so its broadcast
        // and the two reciprocals are issued there and run under the other sixteen row updates of round kb (same operations).
        R pa = warp_bcast(col, 0, 0, lane_) + regadd, pb = warp_bcast(col, 1, 0, lane_), pc = warp_bcast(col, 1, 1, lane_) + regadd;
        R pdet = pa * pc - pb * pb;   // pivot 2 = c - b^2 / a = det / a
        R pra = rrcp(pa), pr2 = pa * rrcp(pdet);
        DDP_NOUNROLL
        for (int kb = 0; kb < 5; kb++) {
            const int p = 2 * kb;
            const R a = pa, b = pb, det = pdet, ra = pra, r2 = pr2;
            if (a <= R(0)) { ok = false; break; }
            if (det <= R(0)) { ok = false; break; }
            Reg<R, 2> mreg;
            FOR_LANES(lane) {
                const R x = col(lane, 0);
                const R l1 = x * ra;
                const R yp = col(lane, 1) - b * l1;
                const R l2 = yp * r2;
                mreg(lane, 0) = l1; mreg(lane, 1) = l2;
                if (lane < 20) {
                    st2(sm + Lay::MB + (kb * 20 + lane) * 2, x, yp);
                    st2(sm + Lay::MSC + (kb * 20 + lane) * 2, l1, l2);
                }
            }
            WARP_SYNC();
            const R *Xp = sm + Lay::MB + (kb * 20 + p) * 2;   // (x, y') of row p + r at Xp[2 r], Xp[2 r + 1]
            FOR_LANES(lane) {
                const R l1 = mreg(lane, 0), l2 = mreg(lane, 1);
                DDP_UNROLL
                for (int r = 2; r < 4; r++) {
                    const Pair2<R> xy = ld2(Xp + 2 * r);   // one 128-bit shared load per row
                    col(lane, r - 2) = col(lane, r) - (xy.x * l1 + xy.y * l2);
                }
            }
            pa = warp_bcast(col, 0, p + 2, lane_) + regadd; pb = warp_bcast(col, 1, p + 2, lane_); pc = warp_bcast(col, 1, p + 3, lane_) + regadd;
            pdet = pa * pc - pb * pb;
            pra = rrcp(pa); pr2 = pa * rrcp(pdet);
            FOR_LANES(lane) {
                const R l1 = mreg(lane, 0), l2 = mreg(lane, 1);
                DDP_UNROLL
                for (int r = 4; r < 20; r++) {
                    const Pair2<R> xy = ld2(Xp + 2 * r);
                    col(lane, r - 2) = col(lane, r) - (xy.x * l1 + xy.y * l2);
                }
            }
        }
        if (!ok) break;
        // rows 10..19 of the matrix (V_xx block and gradient) now sit in elements 0..9 of lanes 10..19
#ifdef DDP_TRACE_CYCLES
        const long long pc2 = ddp_clock();
#endif
        // ---- gains [ku | Ku] = -(L D L^T)^-1 [Qu | Qux] --------------------------------------------------------
        Reg<R, 10> kx;
        FOR_LANES(lane) {
            if (lane >= 10 && lane < 20) {
                // this lane's multipliers are D^-1 L^-1 times its right-hand side; L^T k = -(that), column-oriented: once k_q is
                // known the partial sums above it are updated independently (dependent chain of 10 instead of 45 FMAs)
                R v[10];
                DDP_UNROLL
                for (int kb = 0; kb < 5; kb++) {
                    const Pair2<R> m = ld2(sm + Lay::MSC + (kb * 20 + lane) * 2);
                    v[2 * kb] = m.x; v[2 * kb + 1] = m.y;
                }
                DDP_UNROLL
                for (int q = 9; q >= 0; q--) {
                    kx(lane, q) = -v[q];
                    DDP_UNROLL
                    for (int kb = 0; 2 * kb < q; kb++) {   // L[q][2 kb], L[q][2 kb + 1]: one 128-bit load per round
                        const Pair2<R> m = ld2(sm + Lay::MSC + (kb * 20 + q) * 2);
                        v[2 * kb] += m.x * kx(lane, q);
                        if (2 * kb + 1 < q) v[2 * kb + 1] += m.y * kx(lane, q);
                    }
                }
                const int qc = lane == 19 ? 0 : lane - 9;
                DDP_UNROLL
                for (int q = 0; q < 10; q++) {
                    sm[Lay::KC + q * 10 + qc] = kx(lane, q);
                    Kout[(long long)i * 100 + q * 10 + qc] = kx(lane, q);
                }
            }
        }
#ifdef DDP_TRACE_CYCLES
        const long long pc3 = ddp_clock();
#endif
        // The reference backs up with the UNREGULARISED Quu (ddp.cpp:574/:615, :620-627).  With
        // K = -(Quu+rho I)^-1 Qux that equals the Schur complement above minus rho K^T K (and
        // minus rho K^T k for Vx).
        FOR_LANES(lane) {
            if (lane >= 10 && lane < 19) {
                const int b = lane - 10;
                DDP_UNROLL
                for (int a = 0; a < 9; a++) sm[Lay::S1 + a * 10 + b] = col(lane, a);   // S[a][b]
                sm[Lay::VX + b] = col(lane, 9);
            }
        }
        if (regadd != R(0)) {
            // rho K^T K and rho K^T k off the shared-memory copies, 30 lanes at once: lane < 27 takes rows r .. r+2 of column b of
            // K^T K (b = lane % 9, r = 3 (lane / 9)), lanes 27 .. 29 three entries each of K^T k; both are sum_p KC[p][c0 + i] KC[p][q],
            // ten terms in the order of the single-column loop this replaces (one lane per column, three row groups one after the other).
            WARP_SYNC();
            FOR_LANES(lane) {
                if (lane < 30) {
                    const int rg = lane / 9;
                    const int c0 = lane < 27 ? 3 * rg + 1 : 3 * (lane - 27) + 1, q = lane < 27 ? lane - 9 * rg + 1 : 0;
                    R a0 = R(0), a1 = R(0), a2 = R(0);
                    DDP_UNROLL
                    for (int p = 0; p < 10; p++) {
                        const R *kr = sm + Lay::KC + p * 10 + c0;
                        const R kq = sm[Lay::KC + p * 10 + q];
                        a0 += kr[0] * kq; a1 += kr[1] * kq; a2 += kr[2] * kq;
                    }
                    if (lane < 27) {
                        R *sr = sm + Lay::S1 + (c0 - 1) * 10 + (q - 1);
                        sr[0] -= regadd * a0; sr[10] -= regadd * a1; sr[20] -= regadd * a2;
                    } else {
                        R *vr = sm + Lay::VX + (c0 - 1);
                        vr[0] -= regadd * a0; vr[1] -= regadd * a1; vr[2] -= regadd * a2;
                    }
                }
            }
        }
        WARP_SYNC();
        FOR_LANES(lane) {   // V[b][a] = (S[a][b] + S[b][a]) / 2 (ddp.cpp:628), three entries per lane
            if (lane < 27) {
                const int b = lane / 3, a0 = 3 * (lane - 3 * b);
                DDP_UNROLL
                for (int a = a0; a < a0 + 3; a++) sm[Lay::S2 + b * 10 + a] = R(0.5) * (sm[Lay::S1 + b * 10 + a] + sm[Lay::S1 + a * 10 + b]);
            }
        }
        WARP_SYNC();
#ifdef DDP_TRACE_CYCLES   // where a knot of the recursion spends its time: assembly | pivot rounds | gains | backup
        { const long long pc4 = ddp_clock(); t.cyc_seg[0] += pc1 - pc0; t.cyc_seg[1] += pc2 - pc1; t.cyc_seg[2] += pc3 - pc2; t.cyc_seg[3] += pc4 - pc3; }
#endif
    }
    for (long long k = knots; k < issued; k++) {   // a failed factorisation leaves records in flight: they land before MSC is reused
        const int slot = (int)(k % 3);
        mbar_wait(bars + slot, (rphase >> slot) & 1u);
        rphase ^= 1u << slot;
    }
    FOR_LANES(lane) {
        errq(lane, 0) = errl(lane, 0);
        if (lane == 0) *reinterpret_cast<volatile unsigned *>(sm + Lay::RP) = rphase;
    }
    WARP_SYNC();
    t.n_bwd_knots += knots;
    return ok;
}

// ddp.cpp:440-644.
#if DDP_GPU   // "Speculative backward sweep" (GBoard, below)
template <class R> DDP_DEVICE_NOINLINE bool sweep_resolve(Traj<R> &t, bool want, bool &ok, R &e0);
template <class R> DDP_DEVICE_NOINLINE void sweep_post(Traj<R> &t, R regadd_next);
template <class R> DDP_DEVICE bool gspec_ready(const Traj<R> &t);
#endif
template <class R> DDP_DEVICE_NOINLINE void backward_pass(Traj<R> &t) {
    const int lane_ = t.lane_;
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
    Reg<R, 1> errq;
    bool spec_hit = false, ric_ok = false;
    R e0 = R(0);
#if DDP_GPU
    // A sweep posted by the previous backward pass is this pass iff the linearisation is the same and the regularisation is the
    // one it anticipated (compared bit for bit): a failed search or a failed factorisation in between.  Anything else withdraws it.
    if (t.spec_posted) spec_hit = sweep_resolve(t, t.lin_valid && regadd == t.spec_regadd, ric_ok, e0);
#endif
    // The linearisation depends on the iterate and on mu only: a retry after a failed factorisation or after a
    // failed line search (the reference does not relinearise either, ddp.cpp:476) reuses it.
    if (!t.lin_valid) {
        if (coop_has_helpers(t)) {
            linearize_coop(t, t.lin_emu, t.lin_ecy);
        } else {
            Reg<R, 2> errs;
            linearize_solo(t, errs);
            t.lin_emu = warp_max(errs, 0, lane_);
            t.lin_ecy = warp_max(errs, 1, lane_);
        }
        t.lin_valid = 1;
    }
    WARP_SYNC();
    const long long clk_r = ddp_clock();
#if DDP_GPU
    if (t.spec_on && gspec_ready(t)) {   // what the next backward pass needs if this one, or the search after it, fails
        R reg_next = t.reg + R(1);
        if (reg_next > R(24)) reg_next = R(24);
        sweep_post(t, rpow(t.reg_base, reg_next) - R(1));
    }
#endif
    if (!spec_hit) {
        ric_ok = riccati(t, regadd, errq);
        e0 = warp_max(errq, 0, lane_);
    }
    t.cyc_ric += ddp_clock() - clk_r;
    if (!ric_ok) {
        t.bfailed = 1;
        t.opterr = R(INFINITY);
    } else {
        t.bfailed = 0;
        t.opterr = rmax(rmax(e0, t.infeas ? t.lin_ecy : R(0)), t.lin_emu);  // ddp.cpp:641
    }
    t.cyc_bwd += ddp_clock() - clk0;
}

// =============================================================================================
// Forward pass.
// =============================================================================================
template <class R> struct RollOut { R cost, costq, logcost, err, cmax; };   // cmax: largest constraint value of the candidate

// Running product of barrier arguments with an occasional log, so that sum(log(.)) costs one log per ~dozens of rows.
template <class R> struct LogProd {
    R lp, lsum;
    DDP_DEVICE void mul(R v) {
        lp *= v;
        const R lo = sizeof(R) == 8 ? R(1e-200) : R(1e-25), hi = sizeof(R) == 8 ? R(1e200) : R(1e25);
        if (!(lp > lo && lp < hi)) { lsum += rlog(lp); lp = R(1); }
    }
    DDP_DEVICE R total() const { return lsum + rlog(lp); }
};

template <class R> struct TrialAcc {
    LogProd<R> lg;
    R e1, cmax;   // cmax is a running maximum over ALL rows of the knot (never reset between units: max is order-free)
    int bad;
};

// One constraint row of a line-search trial (ddp.cpp:680-703): slack/dual step from the gains that are
// recomputed here (ks, Ks dx, ky, Ky dx; ddp.cpp:568-572 / :611-612), fraction-to-boundary test, barrier terms.
// sv, yv are the row's slack / dual slack, loaded by the caller ahead of use.
template <class R>
DDP_DEVICE void trial_row(const RowCtx<R> &t, long long ro, R sv, R yv, R cold, R cnew, R jv1, R jv2, R alpha, R tau,
                          TrialAcc<R> &A) {
    if (t.infeas) {
        const R yinv = rrcp(yv);
        const R r = sv * yv - t.mu, rhat = sv * (cold + yv) - r, D = sv * yinv;
        const R ks = yinv * (rhat + sv * jv1), ky = -(cold + yv) - jv1;
        const R ynew = (yv + alpha * ky) + (-jv2), snew = (sv + alpha * ks) + D * jv2;
        if (ynew < (R(1) - tau) * yv || snew < (R(1) - tau) * sv) A.bad = 1;
        t.sn[ro] = snew; t.yn[ro] = ynew;
        A.lg.mul(ynew); A.e1 += rabs(cnew + ynew);
        A.cmax = rmax(A.cmax, cnew);
    } else {
        const R cinv = rrcp(cold);
        const R r = sv * cold + t.mu, D = sv * cinv;
        const R ks = -(cinv * (r + sv * jv1));
        const R snew = (sv + alpha * ks) + (-(D * jv2));
        if (cnew > (R(1) - tau) * cold || snew < (R(1) - tau) * sv) A.bad = 1;
        t.sn[ro] = snew;
        A.lg.mul(-cnew);
        A.cmax = rmax(A.cmax, cnew);
    }
}

// Units u0 .. u1-1 of the line-search rows of the 32-knot block c.base (lane <-> knot).  Unit u covers the row
// groups 4u .. 4u+3; the last one (groups 12-14) also takes the time row and the stage cost.  Per unit and lane:
// part(3u..3u+2) = {stage cost, sum log(barrier argument), |c + y|_1}, badk(u) = this lane's knot if it fails the
// fraction-to-boundary rule (ddp.cpp:683-687 / :699-703), else 0x7fffffff.
template <class R>
DDP_DEVICE_NOINLINE void rows_unit(const JobCtx<R> *cp_, int u0, int u1, int lane_, Reg<R, 16> &part, Reg<int, 4> &badk, R *smw) {
    (void)lane_;
    smw = as_shared(smw);   // staging area of the warp that RUNS the unit (owner or helper)
    const JobCtx<R> c = *cp_;
    const RowCtx<R> t = row_global(c.row);
    const int N = c.N, time_power = c.time_power;
    const R w_snap = c.w_snap, w_time = c.w_time, alpha = c.alpha, tau = c.tau;
    const R *DDP_RESTRICT xu = as_global(c.xu);
    const R *DDP_RESTRICT xun = as_global(c.xun);
    const R *DDP_RESTRICT Kin = as_global(c.K);
    const R *DDP_RESTRICT kdxo = as_global(c.kdx);
    const R *tab = as_shared(t.tab);
    // the knot's old/new point and gains are loaded once for all units of the call
    FOR_LANES(lane) {
        const int i = c.base + lane;
        DDP_UNROLL
        for (int e = 0; e < 4; e++) {
            part(lane, 4 * e) = R(0); part(lane, 4 * e + 1) = R(0); part(lane, 4 * e + 2) = R(0); part(lane, 4 * e + 3) = R(-INFINITY);
            badk(lane, e) = 0x7fffffff;
        }
        if (i < N) {
            R zo[19], zn[19], v1[10], v2[19], tpo[6], tpn[6];
            DDP_UNROLL
            for (int e = 0; e < 19; e++) { zo[e] = xu[(long long)i * 20 + e]; zn[e] = xun[(long long)i * 20 + e]; }
            DDP_UNROLL
            for (int e = 0; e < 10; e++) { v1[e] = Kin[(long long)i * 100 + e * 10]; v2[e] = kdxo[(long long)i * 10 + e]; }
            DDP_UNROLL
            for (int e = 10; e < 19; e++) v2[e] = zn[e] - zo[e];
            time_powers(zo[9], tpo);
            time_powers(zn[9], tpn);
            const int P = t.nplanes[i];
            const double *pl = t.planes + (long long)i * t.PM * 4;
            TrialAcc<R> A;
            A.lg.lp = R(1); A.lg.lsum = R(0); A.e1 = R(0); A.bad = 0; A.cmax = R(-INFINITY);
            R q_cost = R(0);
            // slack rows of the unit's groups through the cp.async row stream (RowStream): all row slots in storage order
            RowStream<R> rows_in;
            row_stream_start(rows_in, t, smw, i, lane, row_slot(4 * u0, 0, t.PM));
            R n_n[4] = {R(0), R(0), R(0), R(0)};   // plane of the next row, loaded one row ahead
            if (4 * u0 < 6 && P > 0) load_plane(pl, 0, n_n);
            DDP_NOUNROLL
            for (int g = 4 * u0; g < 16 && g < 4 * u1; g++) {   // g = 15: the time row and the stage cost
                if (g < 15) {   // one copy of the row code for all groups (see lin_unit)
                    const int shift = group_shift(g), nr = g < 6 ? P : 6, nrw = g < 6 ? t.PM : 6;
                    const R lim = g < 11 ? t.max_vel : t.max_acc;
                    R b[6], bd[6], bn[6], co[3], cd[3], cn[3], j1[3], j2[3];
                    basis_row_rt(tab + g * 6, shift, tpo, b);
                    basis_row_rt(tab + 90 + g * 6, shift + 1, tpo, bd);
                    basis_row_rt(tab + g * 6, shift, tpn, bn);
                    DDP_UNROLL
                    for (int a = 0; a < 3; a++) {
                        co[a] = dot_axis<R, 0>(b, zo, a); cd[a] = dot_axis<R, 0>(bd, zo, a); cn[a] = dot_axis<R, 0>(bn, zn, a);
                        j1[a] = dot_axis<R, 3>(b, v1, a); j2[a] = dot_axis<R, 0>(b, v2, a);
                    }
                    DDP_ROWLOOP_COOP
                    for (int r = 0; r < nrw; r++) {
                        R sv, yv;
                        row_stream_next(rows_in, sv, yv);
                        if (r < nr) {
                            const long long ro_cur = row_ofs(t.MCS, row_slot(g, r, t.PM), i);
                            R n[4];
                            if (g < 6) { n[0] = n_n[0]; n[1] = n_n[1]; n[2] = n_n[2]; n[3] = n_n[3]; }
                            else fixed_row(r, lim, n);
                            if (g < 6 && r + 1 < nr) load_plane(pl, r + 1, n_n);
                            const R cold = ((n[0] * co[0] + n[1] * co[1]) + n[2] * co[2]) + n[3] - t.margin;
                            const R cnew = ((n[0] * cn[0] + n[1] * cn[1]) + n[2] * cn[2]) + n[3] - t.margin;
                            const R tc = (n[0] * cd[0] + n[1] * cd[1]) + n[2] * cd[2];
                            const R jv1 = ((n[0] * j1[0] + n[1] * j1[1]) + n[2] * j1[2]) + tc * v1[9];
                            const R jv2 = ((n[0] * j2[0] + n[1] * j2[1]) + n[2] * j2[2]) + tc * v2[9];
                            trial_row(t, ro_cur, sv, yv, cold, cnew, jv1, jv2, alpha, tau, A);
                        }
                    }
                    if (g + 1 < 6 && P > 0) load_plane(pl, 0, n_n);   // first plane of the next position group
                } else {
                    const long long ro = row_ofs(t.MCS, 6 * t.PM + 54, i);
                    R sv, yv;
                    row_stream_next(rows_in, sv, yv);
                    trial_row(t, ro, sv, yv, -zo[9] + R(0.3) - t.margin, -zn[9] + R(0.3) - t.margin, -v1[9], -v2[9], alpha, tau, A);
                    // stage cost q(x,u), ddp.cpp:1294-1305
                    R m[9], mu9[9];
                    rmat<R>(0, tpn, m);
                    rmat_times_u(m, zn, mu9);
                    const R T = tpn[1];
                    const R tterm = time_power == 2 ? R(0.5) * T * w_time * T : R(0.5) * w_time * T;
                    q_cost = R(0.5) * w_snap * dot9(zn, mu9) + tterm;
                }
                if ((g & 3) == 3) {   // end of unit g / 4: bank its partials and start the next unit's accumulators
                    const int uu = g >> 2;
                    DDP_UNROLL
                    for (int e = 0; e < 4; e++) {
                        if (e == uu) {
                            part(lane, 4 * e) = q_cost; part(lane, 4 * e + 1) = A.lg.total(); part(lane, 4 * e + 2) = A.e1;
                            part(lane, 4 * e + 3) = A.cmax;
                            if (A.bad) badk(lane, e) = i;
                        }
                    }
                    A.lg.lp = R(1); A.lg.lsum = R(0); A.e1 = R(0); A.bad = 0; A.cmax = R(-INFINITY);
                }
            }
            cp_wait<0>();   // the stream's look-ahead copies land before anybody reuses the staging area
        }
    }
}

// Once per kernel launch and warp: the three mbarriers (one arrival each: the issuing lane's expect_tx) that the bulk copies of
// the warp complete on -- the record ring of the Riccati recursion and the knot ring of the line search, never both at once --
// and their phase word (bit b = parity barrier b completes next).
#if DDP_GPU
template <class R> DDP_DEVICE void ring_init(R *sm, int lane_) {
    if (lane_ < 3) mbar_init(reinterpret_cast<unsigned long long *>(sm + Lay::RB) + lane_, 1);
    if (lane_ == 0) *reinterpret_cast<unsigned *>(sm + Lay::RP) = 0u;
    fence_async_smem();
    __syncwarp();
}
#endif
// Knot ring of the line search: the gains [ku | Ku] (100) and the old point [u; x] (20) of knot k as two bulk copies into ring
// slot k % 3 = [K_k | xu_k], issued by lane 0 two knots ahead of the state recursion.  (Round 1 staged them with sixty 16-byte
// cp.async per knot spread over the lanes: 5 % of the kernel's issue slots, profiles/r2f.)
template <class R>
DDP_DEVICE void knot_issue(R *ring, unsigned long long *bars, const R *K, const R *xu, int k, int lane) {
    if (lane == 0) {
        R *slot = ring + (k % 3) * 120;
        mbar_expect_tx(bars + k % 3, (unsigned)(120 * sizeof(R)));
        bulk_g2s(slot, K + (long long)k * 100, (unsigned)(100 * sizeof(R)), bars + k % 3);
        bulk_g2s(slot + 100, xu + (long long)k * 20, (unsigned)(20 * sizeof(R)), bars + k % 3);
    }
}

// Closed-loop state / control recursion over knots base .. base+nk-1 of a line-search trial (ddp.cpp:689-697, :1062-1067):
// lanes 0-9 own u, lanes 10-18 own x.  Out of line on purpose: inside forward_trial the register allocator spilled this
// loop's state around the calls of the row phase.  xcur carries the state across blocks.
template <class R>
DDP_DEVICE_NOINLINE void rollout_block(R *sm, const R *DDP_RESTRICT xu, R *DDP_RESTRICT xun, const R *DDP_RESTRICT Kin,
                                       R *DDP_RESTRICT kdxo, int N, int base, int nk, R alpha, int lane_, Reg<R, 1> &xcur_io) {
    sm = as_shared(sm);
    xu = as_global(xu); xun = as_global(xun); Kin = as_global(Kin); kdxo = as_global(kdxo);
    R *ring = sm + Lay::MSC;
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(sm + Lay::RB);
    unsigned rphase = *reinterpret_cast<volatile unsigned *>(sm + Lay::RP);
    Reg<R, 1> xn, xcur;   // local copy: a by-reference Reg lives in local memory and would be re-read after every store
    FOR_LANES(lane) { xn(lane, 0) = R(0); xcur(lane, 0) = xcur_io(lane, 0); }
        for (int i = base; i < base + nk; i++) {
            const R *slot = ring + (i % 3) * 120;
            FOR_LANES(lane) {   // knot i + 2 into the slot knot i - 1 left (every lane is past the WARP_SYNC that ended it)
                if (i + 2 < N) knot_issue(ring, bars, Kin, xu, i + 2, lane);
            }
            mbar_wait(bars + i % 3, (rphase >> (i % 3)) & 1u);
            rphase ^= 1u << (i % 3);
            FOR_LANES(lane) {
                if (lane >= 10 && lane < 19) {
                    const R x = xcur(lane, 0);
                    sm[Lay::DX + lane - 10] = x - slot[100 + lane];
                    sm[Lay::ZN + lane] = x;
                    xun[(long long)i * 20 + lane] = x;
                }
            }
            WARP_SYNC();
            FOR_LANES(lane) {   // unew = (uold + alpha ku) + Ku dx (ddp.cpp:689/:695)
                if (lane < 10) {
                    const R *Kr = slot + lane * 10;
                    R kdx = R(0);
                    DDP_UNROLL
                    for (int b = 0; b < 9; b++) kdx += Kr[1 + b] * sm[Lay::DX + b];
                    const R un = step_u(slot[100 + lane], alpha, Kr[0], kdx);
                    sm[Lay::ZN + lane] = un;
                    xun[(long long)i * 20 + lane] = un;
                    kdxo[(long long)i * 10 + lane] = kdx;
                }
            }
            WARP_SYNC();
            R tp[6], fg[18];
            time_powers(sm[Lay::ZN + 9], tp);
            fg_matrix(tp, fg);
            FOR_LANES(lane) {   // x+ = (F (x) I) x + (G (x) I) u (ddp.cpp:1062-1067)
                if (lane >= 10 && lane < 19) {
                    const int o = (lane - 10) / 3, a = (lane - 10) % 3;
                    R s1 = R(0), s2 = R(0), fr[6];
                    fg_row(fg, o, fr);
                    DDP_UNROLL
                    for (int b = 0; b < 3; b++) {
                        if (b >= o) s1 += fr[b] * sm[Lay::ZN + 10 + 3 * b + a];
                        s2 += fr[3 + b] * sm[Lay::ZN + 3 * b + a];
                    }
                    xn(lane, 0) = s1 + s2;
                }
            }
            WARP_SYNC();
            FOR_LANES(lane) { xcur(lane, 0) = xn(lane, 0); }
        }
    FOR_LANES(lane) {
        xcur_io(lane, 0) = xcur(lane, 0);
        if (lane == 0) *reinterpret_cast<volatile unsigned *>(sm + Lay::RP) = rphase;
    }
    WARP_SYNC();
}
// Knots still in flight when a trial ends (two ahead of the last knot processed, `done` knots processed): they land before the
// ring area is reused.
template <class R> DDP_DEVICE void knot_ring_drain(R *sm, int N, int done, int lane_) {
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(sm + Lay::RB);
    unsigned rphase = *reinterpret_cast<volatile unsigned *>(sm + Lay::RP);
    const int issued = done + 2 < N ? done + 2 : N;
    for (int k = done; k < issued; k++) {
        mbar_wait(bars + k % 3, (rphase >> (k % 3)) & 1u);
        rphase ^= 1u << (k % 3);
    }
    FOR_LANES(lane) { if (lane == 0) *reinterpret_cast<volatile unsigned *>(sm + Lay::RP) = rphase; }
    WARP_SYNC();
}

// One line-search trial with step size alpha (ddp.cpp:674-734).  The closed-loop state/control recursion
// runs sequentially (lane <-> element) for 32 knots at a time; the constraint rows of those 32 knots are then
// evaluated lane <-> knot.  Writes the candidate into xun/sn/yn and returns false when the
// fraction-to-boundary test fails at some knot (ddp.cpp:683-687 / :699-703).
template <class R> DDP_DEVICE_NOINLINE bool forward_trial_coop(Traj<R> &tt_, R alpha, R tau, RollOut<R> &out) {
    const int lane_ = tt_.lane_;
    const RowCtx<R> t = row_ctx(tt_);
    R *sm = as_shared(tt_.sm);
    const R *tab = as_shared(t.tab);
    (void)tab;
    const int N = tt_.N;
    const R *DDP_RESTRICT xu = as_global(tt_.xu);
    R *DDP_RESTRICT xun = as_global(tt_.xun);
    const R *DDP_RESTRICT Kin = as_global(tt_.K);
    R *DDP_RESTRICT kdxo = as_global(tt_.kdx);
    const R w_terminal = tt_.w_terminal, w_snap = tt_.w_snap, w_time = tt_.w_time, tol = tt_.tol;
    const int time_power = tt_.time_power;
    long long fwd_knots = 0, cyc_seq = 0;
    Reg<R, 1> xcur;
    Reg<R, 4> acc;  // per-lane partials: stage cost, log barrier, |c+y|_1, max c
    // Old point and gains of the next knots stream into a three-slot shared-memory ring by bulk copies (knot_issue), two knots
    // ahead of the recursion (the MSC scratch of the linearisation is free during the line search): slot = [K_i (100) | xu_i (20)].
    R *ring = sm + Lay::MSC;
    FOR_LANES(lane) {
        acc(lane, 0) = R(0); acc(lane, 1) = R(0); acc(lane, 2) = R(0); acc(lane, 3) = R(-INFINITY);
        xcur(lane, 0) = (lane >= 10 && lane < 19) ? xu[lane] : R(0);  // xnew[0] = xold[0]
        fence_async_all();   // the gains (Riccati's st.global) and the last generic stores to the ring area before the copy engine
    }
    WARP_SYNC();
    FOR_LANES(lane) {
        for (int d = 0; d < 2 && d < N; d++) knot_issue(ring, reinterpret_cast<unsigned long long *>(sm + Lay::RB), Kin, xu, d, lane);
    }
    int knots_done = 0;
    bool ok = true;
    for (int base = 0; base < N && ok; base += 32) {
        const int nk = N - base < 32 ? N - base : 32;
        const long long clk_s = ddp_clock();
        rollout_block(sm, xu, xun, Kin, kdxo, N, base, nk, alpha, lane_, xcur);
        knots_done = base + nk;
        cyc_seq += ddp_clock() - clk_s;
        // ---- rows of knots base .. base+nk-1, lane <-> knot: four units of row groups (run_job) --------------
        {
            JobCtx<R> c;
            c.row = t;
            c.xu = xu; c.xun = xun; c.K = Kin; c.kdx = kdxo; c.H = nullptr; c.aux = nullptr;
            c.N = N; c.base = base; c.time_power = time_power; c.type = JOB_ROWS;
            c.w_snap = w_snap; c.w_time = w_time; c.alpha = alpha; c.tau = tau;
            Reg<R, 16> part;
            Reg<int, 4> badk;
            run_rows(tt_, c, part, badk);
            Reg<int, 1> bmin;
            FOR_LANES(lane) {   // units in fixed order per lane: the same sums whoever ran the units
                int m = 0x7fffffff;
                DDP_UNROLL
                for (int u = 0; u < 4; u++) {
                    acc(lane, 0) += part(lane, 4 * u); acc(lane, 1) += part(lane, 4 * u + 1); acc(lane, 2) += part(lane, 4 * u + 2);
                    acc(lane, 3) = rmax(acc(lane, 3), part(lane, 4 * u + 3));
                    if (badk(lane, u) < m) m = badk(lane, u);
                }
                bmin(lane, 0) = m;
            }
            const int first = warp_min_int(bmin, 0, lane_);
            if (first != 0x7fffffff) { fwd_knots += first - base + 1; ok = false; }
            else fwd_knots += nk;
        }
    }
    FOR_LANES(lane) { cp_wait<0>(); }   // nothing of the ring / staging may land after this trial
    WARP_SYNC();
    knot_ring_drain(sm, N, knots_done, lane_);
    tt_.n_fwd_knots += fwd_knots;
#ifndef DDP_SPEC_PROFILE
    tt_.cyc_seq += cyc_seq;
#endif
    if (!ok) return false;
    // terminal cost (ddp.cpp:1289-1292) and totals
    Reg<R, 1> pt;
    FOR_LANES(lane) {
        pt(lane, 0) = R(0);
        if (lane >= 10 && lane < 19) {
            const R d = xcur(lane, 0) - sm[Lay::XD + lane - 10];
            pt(lane, 0) = d * (w_terminal * d);
            xun[(long long)N * 20 + lane] = xcur(lane, 0);
        }
    }
    const R qs = warp_sum(acc, 0, lane_), p = R(0.5) * warp_sum(pt, 0, lane_);
    out.costq = qs;
    out.cost = qs + p;
    out.logcost = out.cost - t.mu * warp_sum(acc, 1, lane_);
    out.err = t.infeas ? rmax(tol, warp_sum(acc, 2, lane_)) : R(0);
    out.cmax = warp_max(acc, 3, lane_);
    WARP_SYNC();
    return true;
}

// The same trial for a warp working alone (the common case): recursion and rows of each 32-knot block in one function.
// The row partials are banked per unit (row groups 4u .. 4u+3) exactly like rows_unit does, so a trial gives the same
// bits whether it ran here or as jobs.
template <class R> DDP_DEVICE_NOINLINE bool forward_trial_solo(Traj<R> &tt_, R alpha, R tau, RollOut<R> &out) {
    const int lane_ = tt_.lane_;
    const RowCtx<R> t = row_ctx(tt_);
    R *sm = as_shared(tt_.sm);
    const R *tab = as_shared(t.tab);
    (void)tab;
    const int N = tt_.N;
    const R *DDP_RESTRICT xu = as_global(tt_.xu);
    R *DDP_RESTRICT xun = as_global(tt_.xun);
    const R *DDP_RESTRICT Kin = as_global(tt_.K);
    R *DDP_RESTRICT kdxo = as_global(tt_.kdx);
    const R w_terminal = tt_.w_terminal, w_snap = tt_.w_snap, w_time = tt_.w_time, tol = tt_.tol;
    const int time_power = tt_.time_power;
    long long fwd_knots = 0, cyc_seq = 0;
    Reg<R, 1> xcur, xn;
    Reg<R, 4> acc;  // per-lane partials: stage cost, log barrier, |c+y|_1, max c
    // Old point and gains of the next knots stream into a three-slot shared-memory ring by bulk copies (knot_issue), two knots
    // ahead of the recursion (the MSC scratch of the linearisation is free during the line search): slot = [K_i (100) | xu_i (20)].
    R *ring = sm + Lay::MSC;
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(sm + Lay::RB);
    unsigned rphase = *reinterpret_cast<volatile unsigned *>(sm + Lay::RP);
    FOR_LANES(lane) {
        acc(lane, 0) = R(0); acc(lane, 1) = R(0); acc(lane, 2) = R(0); acc(lane, 3) = R(-INFINITY);
        xcur(lane, 0) = (lane >= 10 && lane < 19) ? xu[lane] : R(0);  // xnew[0] = xold[0]
        xn(lane, 0) = R(0);
        fence_async_all();   // the gains (Riccati's st.global) and the last generic stores to the ring area before the copy engine
    }
    WARP_SYNC();
    FOR_LANES(lane) {
        for (int d = 0; d < 2 && d < N; d++) knot_issue(ring, bars, Kin, xu, d, lane);
    }
    int knots_done = 0;
    bool ok = true;
    for (int base = 0; base < N && ok; base += 32) {
        const int nk = N - base < 32 ? N - base : 32;
        const long long clk_s = ddp_clock();
        for (int i = base; i < base + nk; i++) {
            const R *slot = ring + (i % 3) * 120;
            FOR_LANES(lane) {   // knot i + 2 into the slot knot i - 1 left (every lane is past the WARP_SYNC that ended it)
                if (i + 2 < N) knot_issue(ring, bars, Kin, xu, i + 2, lane);
            }
            mbar_wait(bars + i % 3, (rphase >> (i % 3)) & 1u);
            rphase ^= 1u << (i % 3);
            FOR_LANES(lane) {
                if (lane >= 10 && lane < 19) {
                    const R x = xcur(lane, 0);
                    sm[Lay::DX + lane - 10] = x - slot[100 + lane];
                    sm[Lay::ZN + lane] = x;
                    xun[(long long)i * 20 + lane] = x;
                }
            }
            WARP_SYNC();
            FOR_LANES(lane) {   // unew = (uold + alpha ku) + Ku dx (ddp.cpp:689/:695)
                if (lane < 10) {
                    const R *Kr = slot + lane * 10;
                    R kdx = R(0);
                    DDP_UNROLL
                    for (int b = 0; b < 9; b++) kdx += Kr[1 + b] * sm[Lay::DX + b];
                    const R un = step_u(slot[100 + lane], alpha, Kr[0], kdx);
                    sm[Lay::ZN + lane] = un;
                    xun[(long long)i * 20 + lane] = un;
                    kdxo[(long long)i * 10 + lane] = kdx;
                }
            }
            WARP_SYNC();
            R tp[6], fg[18];
            time_powers(sm[Lay::ZN + 9], tp);
            fg_matrix(tp, fg);
            FOR_LANES(lane) {   // x+ = (F (x) I) x + (G (x) I) u (ddp.cpp:1062-1067)
                if (lane >= 10 && lane < 19) {
                    const int o = (lane - 10) / 3, a = (lane - 10) % 3;
                    R s1 = R(0), s2 = R(0), fr[6];
                    fg_row(fg, o, fr);
                    DDP_UNROLL
                    for (int b = 0; b < 3; b++) {
                        if (b >= o) s1 += fr[b] * sm[Lay::ZN + 10 + 3 * b + a];
                        s2 += fr[3 + b] * sm[Lay::ZN + 3 * b + a];
                    }
                    xn(lane, 0) = s1 + s2;
                }
            }
            WARP_SYNC();
            FOR_LANES(lane) { xcur(lane, 0) = xn(lane, 0); }
        }
        knots_done = base + nk;
        cyc_seq += ddp_clock() - clk_s;
        // ---- rows of knots base .. base+nk-1, lane <-> knot; slack rows from the TMA ring (RowRing) ---------------
        Reg<int, 1> badk;
        FOR_LANES(lane) {
            const int i = base + lane;
            const bool live = i < N;
            const int il = live ? i : N - 1;   // lanes past the end run along on the last knot and never store / accumulate
            badk(lane, 0) = 0x7fffffff;
            {
                // Live across the 15 groups: the old point zo and the new STATE zx in registers; the feed-forward v1 and the
                // closed-loop part v2 of the nine coefficient controls wait in shared memory (the job-partial area of MSC, idle
                // while a warp works alone), their T entries and the two segment times stay in registers.  The new controls and
                // the state step are re-formed per group with the recursion's own expression (step_u: same bits), the powers
                // of T per group as well: every value kept live here is a register the row loop below cannot use (the
                // 15 per-group constants co..j2 used to be spilled and re-read from local memory in every row).
                R zo[19], zx[9];
                DDP_UNROLL
                for (int e = 0; e < 19; e++) zo[e] = xu[(long long)il * 20 + e];
                DDP_UNROLL
                for (int e = 0; e < 9; e++) zx[e] = xun[(long long)il * 20 + 10 + e];
                R *vst = sm + Lay::MSC + 512 + lane;   // vst[e * 32]: v1[e], vst[(9 + e) * 32]: v2[e], e < 9
                DDP_UNROLL
                for (int e = 0; e < 9; e++) { vst[e * 32] = Kin[(long long)il * 100 + e * 10]; vst[(9 + e) * 32] = kdxo[(long long)il * 10 + e]; }
                const R v1T = Kin[(long long)il * 100 + 90], v2T = kdxo[(long long)il * 10 + 9];
                const R Tn = step_u(zo[9], alpha, v1T, v2T);   // = xun[9] (ddp.cpp:689/:695, rollout above)
                TrialAcc<R> A;
                A.lg.lp = R(1); A.lg.lsum = R(0); A.e1 = R(0); A.bad = 0; A.cmax = R(-INFINITY);
                const int P = live ? t.nplanes[il] : 0;
                const double *pl = t.planes + (long long)il * t.PM * 4;
                R n_n[4] = {R(0), R(0), R(0), R(0)};   // plane of the next row, loaded one row ahead
                if (P > 0) load_plane(pl, 0, n_n);
                RowStream<R> rows_in;
                row_stream_start(rows_in, t, sm, i, lane);
                DDP_NOUNROLL
                for (int g = 0; g < 16; g++) {   // one copy of the row code for all groups (see linearize); g = 15: the time row
                    const int shift = group_shift(g), nr = g < 6 ? P : (g < 15 ? (live ? 6 : 0) : (live ? 1 : 0));
                    const int nrw = g < 6 ? t.PM : (g < 15 ? 6 : 1);   // warp-uniform: every row slot is visited
                    const R lim = g < 11 ? t.max_vel : t.max_acc;
                    R co[3], cd[3], cn[3], j1[3], j2[3];
                    if (g < 15) {
                        R b[6], bd[6], bn[6], tpo[6], tpn[6];
                        time_powers(zo[9], tpo);
                        time_powers(Tn, tpn);
                        basis_row_rt(tab + g * 6, shift, tpo, b);
                        basis_row_rt(tab + 90 + g * 6, shift + 1, tpo, bd);
                        basis_row_rt(tab + g * 6, shift, tpn, bn);
                        DDP_UNROLL
                        for (int a = 0; a < 3; a++) {
                            co[a] = dot_axis<R, 0>(b, zo, a); cd[a] = dot_axis<R, 0>(bd, zo, a);
                            // new point [(uo + alpha ku) + Ku dx ; xn] and step [Ku dx ; xn - xo], orders l = 0..2 are the state
                            R accn = R(0), acc2 = R(0), acc1 = R(0);
                            DDP_UNROLL
                            for (int l = 0; l < 3; l++) {
                                accn += bn[l] * zx[3 * l + a];
                                acc2 += b[l] * (zx[3 * l + a] - zo[10 + 3 * l + a]);
                            }
                            DDP_UNROLL
                            for (int l = 3; l < 6; l++) {
                                const int e = 3 * (l - 3) + a;
                                const R k1 = vst[e * 32], k2 = vst[(9 + e) * 32];
                                accn += bn[l] * step_u(zo[e], alpha, k1, k2);
                                acc2 += b[l] * k2;
                                acc1 += b[l] * k1;
                            }
                            cn[a] = accn; j2[a] = acc2; j1[a] = acc1;
                        }
                    } else {   // the time row -T + 0.3 <= 0 (ddp.cpp:1279) as a row with n = (1, 0, 0 | 0.3): one copy of the row code
                        co[0] = -zo[9]; cn[0] = -Tn; cd[0] = R(0); j1[0] = -v1T; j2[0] = -v2T;
                        co[1] = co[2] = cn[1] = cn[2] = cd[1] = cd[2] = j1[1] = j1[2] = j2[1] = j2[2] = R(0);
                    }
                    DDP_ROWLOOP_SOLO
                    for (int r = 0; r < nrw; r++) {
                        const long long ro_cur = row_ofs(t.MCS, g < 15 ? row_slot(g, r, t.PM) : 6 * t.PM + 54, i);
                        R sv, yv;
                        row_stream_next(rows_in, sv, yv);
                        R n[4];
                        if (g < 6) { n[0] = n_n[0]; n[1] = n_n[1]; n[2] = n_n[2]; n[3] = n_n[3]; }
                        else if (g < 15) fixed_row(r, lim, n);
                        else { n[0] = R(1); n[1] = R(0); n[2] = R(0); n[3] = R(0.3); }
                        if (g < 6 && r + 1 < nr) load_plane(pl, r + 1, n_n);
                        if (r < nr) {
                            const R cold = ((n[0] * co[0] + n[1] * co[1]) + n[2] * co[2]) + n[3] - t.margin;
                            const R cnew = ((n[0] * cn[0] + n[1] * cn[1]) + n[2] * cn[2]) + n[3] - t.margin;
                            const R tc = (n[0] * cd[0] + n[1] * cd[1]) + n[2] * cd[2];
                            const R jv1 = ((n[0] * j1[0] + n[1] * j1[1]) + n[2] * j1[2]) + tc * v1T;
                            const R jv2 = ((n[0] * j2[0] + n[1] * j2[1]) + n[2] * j2[2]) + tc * v2T;
                            trial_row(t, ro_cur, sv, yv, cold, cnew, jv1, jv2, alpha, tau, A);
                        }
                    }
                    if (g + 1 < 6 && P > 0) load_plane(pl, 0, n_n);   // first plane of the next position group
                    if ((g & 3) == 3 && g < 15) {   // end of unit g / 4 (see rows_unit): bank its partials, restart the accumulators
                        if (live) { acc(lane, 1) += A.lg.total(); acc(lane, 2) += A.e1; }
                        if (A.bad) badk(lane, 0) = i;
                        A.lg.lp = R(1); A.lg.lsum = R(0); A.e1 = R(0); A.bad = 0;
                    }
                }
                if (A.bad) badk(lane, 0) = i;
                if (live) {   // stage cost q(x,u), ddp.cpp:1294-1305
                    R m[9], mu9[9], un[9], tpn[6];
                    DDP_UNROLL
                    for (int e = 0; e < 9; e++) un[e] = step_u(zo[e], alpha, vst[e * 32], vst[(9 + e) * 32]);
                    time_powers(Tn, tpn);
                    rmat<R>(0, tpn, m);
                    rmat_times_u(m, un, mu9);
                    const R T = tpn[1];
                    const R tterm = time_power == 2 ? R(0.5) * T * w_time * T : R(0.5) * w_time * T;
                    acc(lane, 0) += R(0.5) * w_snap * dot9(un, mu9) + tterm;
                    acc(lane, 1) += A.lg.total();
                    acc(lane, 2) += A.e1;
                    acc(lane, 3) = rmax(acc(lane, 3), A.cmax);
                }
            }
        }
        const int first = warp_min_int(badk, 0, lane_);
        if (first != 0x7fffffff) { fwd_knots += first - base + 1; ok = false; }
        else fwd_knots += nk;
    }
    FOR_LANES(lane) {   // nothing of the rings / staging may land after this trial
        cp_wait<0>();
        if (lane == 0) *reinterpret_cast<volatile unsigned *>(sm + Lay::RP) = rphase;
    }
    WARP_SYNC();
    knot_ring_drain(sm, N, knots_done, lane_);
    tt_.n_fwd_knots += fwd_knots;
#ifndef DDP_SPEC_PROFILE
    tt_.cyc_seq += cyc_seq;
#endif
    if (!ok) return false;
    // terminal cost (ddp.cpp:1289-1292) and totals
    Reg<R, 1> pt;
    FOR_LANES(lane) {
        pt(lane, 0) = R(0);
        if (lane >= 10 && lane < 19) {
            const R d = xcur(lane, 0) - sm[Lay::XD + lane - 10];
            pt(lane, 0) = d * (w_terminal * d);
            xun[(long long)N * 20 + lane] = xcur(lane, 0);
        }
    }
    const R qs = warp_sum(acc, 0, lane_), p = R(0.5) * warp_sum(pt, 0, lane_);
    out.costq = qs;
    out.cost = qs + p;
    out.logcost = out.cost - t.mu * warp_sum(acc, 1, lane_);
    out.err = t.infeas ? rmax(tol, warp_sum(acc, 2, lane_)) : R(0);
    out.cmax = warp_max(acc, 3, lane_);
    WARP_SYNC();
    return true;
}

// Open-loop rollout of the initial controls (initialroll, ddp.cpp:1608-1620) and its cost.
template <class R> DDP_DEVICE_NOINLINE void initial_roll(Traj<R> &t) {
    const int lane_ = t.lane_;
    R *sm = as_shared(t.sm);
    const int N = t.N;
    Reg<R, 1> xcur, xn;
    FOR_LANES(lane) { xcur(lane, 0) = (lane >= 10 && lane < 19) ? t.xu[lane] : R(0); xn(lane, 0) = R(0); }
    for (int i = 0; i < N; i++) {
        FOR_LANES(lane) {
            if (lane >= 10 && lane < 19) { sm[Lay::ZN + lane] = xcur(lane, 0); t.xu[(long long)i * 20 + lane] = xcur(lane, 0); }
            if (lane < 10) sm[Lay::ZN + lane] = t.xu[(long long)i * 20 + lane];
        }
        WARP_SYNC();
        R tp[6], fg[18];
        time_powers(sm[Lay::ZN + 9], tp);
        fg_matrix(tp, fg);
        FOR_LANES(lane) {
            if (lane >= 10 && lane < 19) {
                const int o = (lane - 10) / 3, a = (lane - 10) % 3;
                R s1 = R(0), s2 = R(0), fr[6];
                fg_row(fg, o, fr);
                DDP_UNROLL
                for (int b = 0; b < 3; b++) {
                    if (b >= o) s1 += fr[b] * sm[Lay::ZN + 10 + 3 * b + a];
                    s2 += fr[3 + b] * sm[Lay::ZN + 3 * b + a];
                }
                xn(lane, 0) = s1 + s2;
            }
        }
        WARP_SYNC();
        FOR_LANES(lane) { xcur(lane, 0) = xn(lane, 0); }
    }
    Reg<R, 2> acc;
    FOR_LANES(lane) {
        acc(lane, 0) = R(0); acc(lane, 1) = R(0);
        if (lane >= 10 && lane < 19) {
            t.xu[(long long)N * 20 + lane] = xcur(lane, 0);
            const R d = xcur(lane, 0) - sm[Lay::XD + lane - 10];
            acc(lane, 1) = d * (t.w_terminal * d);
        }
        for (int i = lane; i < N; i += 32) {
            R u[10], tp[6];
            DDP_UNROLL
            for (int e = 0; e < 10; e++) u[e] = t.xu[(long long)i * 20 + e];
            time_powers(u[9], tp);
            acc(lane, 0) += stage_cost(t, tp, u);
        }
    }
    t.costq = warp_sum(acc, 0, lane_);
    t.cost = t.costq + R(0.5) * warp_sum(acc, 1, lane_);
    WARP_SYNC();
}

// Constraint scan at the current iterate, lane <-> knot.
//   mode 0: barrier sum and infeasibility (resetfilter, ddp.cpp:1636-1662) -> lsum, e1; returns false
//   mode 1: any c >= thresh (ddp.cpp:346-355);  mode 2: any c > thresh (ddp.cpp:255-269)
template <class R> DDP_DEVICE_NOINLINE bool scan_constraints(Traj<R> &t, int mode, R thresh, R &lsum, R &e1sum) {
    const int lane_ = t.lane_;
    const RowCtx<R> c_ = row_ctx(t);
    const R *DDP_RESTRICT xu = t.xu;
    const int N = t.N;
    Reg<R, 2> acc;
    Reg<int, 1> viol;
    FOR_LANES(lane) {
        acc(lane, 0) = R(0); acc(lane, 1) = R(0); viol(lane, 0) = 0;
        for (int i = lane; i < N; i += 32) {
            R z[19];
            DDP_UNROLL
            for (int e = 0; e < 19; e++) z[e] = xu[(long long)i * 20 + e];
            LogProd<R> lg;
            lg.lp = R(1); lg.lsum = R(0);
            R e1 = R(0);
            int v = 0;
            visit_rows(c_, i, z, [&](int slot, R c) {
                if (mode == 0) {
                    if (c_.infeas) { const R yv = c_.y[row_ofs(c_.MCS, slot, i)]; lg.mul(yv); e1 += rabs(c + yv); }
                    else lg.mul(-c);
                } else if (mode == 1) { if (c >= thresh) v = 1; }
                else { if (c > thresh) v = 1; }
            });
            if (mode == 0) { acc(lane, 0) += lg.total(); acc(lane, 1) += e1; }
            if (v) viol(lane, 0) = 1;
        }
    }
    if (mode == 0) { lsum = warp_sum(acc, 0, lane_); e1sum = warp_sum(acc, 1, lane_); return false; }
    return warp_any(viol, 0, lane_);
}

// Barrier cost and infeasibility at the current iterate + filter reset (ddp.cpp:1636-1662).
template <class R> DDP_DEVICE void reset_filter(Traj<R> &t) {
    const int lane_ = t.lane_;
    R lsum = R(0), e1 = R(0);
    scan_constraints(t, 0, R(0), lsum, e1);
    t.logcost = t.cost - t.mu * lsum;
    t.err = R(0);
    if (t.infeas) { t.err = e1; if (t.err < t.tol) t.err = R(0); }
    FOR_LANES(lane) { if (lane == 0) { t.filt[0] = t.logcost; t.filt[1] = t.err; } }
    WARP_SYNC();
    t.nfilter = 1;
    t.step = 0;
    t.failed = 0;
}

// Filter acceptance of a trial that passed the fraction-to-boundary rule, ddp.cpp:741-757: rejected if some entry is <=
// the candidate in both coordinates; an accepted candidate evicts the entries it dominates.  Lane 0 owns the filter.
template <class R> DDP_DEVICE bool filter_try(Traj<R> &t, R logcost, R err) {
    const int lane_ = t.lane_;
    FOR_LANES(lane) {
        if (lane == 0) {
            bool rej = false;
            for (int k = 0; k < t.nfilter; k++)
                if (logcost >= t.filt[2 * k] && err >= t.filt[2 * k + 1]) { rej = true; break; }
            int nk = 0;
            if (!rej) {
                for (int k = 0; k < t.nfilter; k++) {
                    const R f0 = t.filt[2 * k], f1 = t.filt[2 * k + 1];
                    if (logcost > f0 || err > f1) { t.filt[2 * nk] = f0; t.filt[2 * nk + 1] = f1; nk++; }
                }
                if (nk >= t.fcap) nk = t.fcap - 1;
                t.filt[2 * nk] = logcost; t.filt[2 * nk + 1] = err;
            }
            t.sm[Lay::FL] = rej ? R(1) : R(0);
            t.sm[Lay::FL + 1] = R(nk);
        }
    }
    WARP_SYNC();
    const bool rej = t.sm[Lay::FL] != R(0);
    const int nkeep = (int)t.sm[Lay::FL + 1];
    WARP_SYNC();
    if (rej) return false;
    t.nfilter = nkeep + 1;
    return true;
}

// =============================================================================================
// Speculative line search.  The solves that dominate the tail of a batch run stage 1 to iter_max, and ~40 % of
// their line searches FAIL: eleven full-length trials one after the other, every one rejected (tools/tail_report.py).
// The trials of one line search are independent of each other (the filter only changes when a trial is accepted), so
// once whole CTAs are idle their warps run the trials 2^-1 .. 2^-10 of a remaining solve concurrently with the owner's
// own trial 2^0, each into the candidate buffers of its own (idle) workspace slot.  The owner then walks the results in
// step order - exactly the decisions of the sequential search - and copies the winning candidate, if any, into its own
// buffers.  Boards live in global memory (two per slot for the searches, used alternately so that a cancelled search never has to be
// waited for); claimed units always run to completion, and an owner that finds units unclaimed when its own trial is
// over closes the board and carries on sequentially, so nothing ever waits on a warp that is itself waiting.
// =============================================================================================
// From this many idle warps on, an idle CTA runs ONE remote unit at a time and keeps its other warps as row helpers of that unit
// (a trial with three row helpers takes 0.33 Mcycles, alone 0.73): latency of the owner's search instead of trial throughput.
#ifndef DDP_GSPEC_EXCL_IDLE
#define DDP_GSPEC_EXCL_IDLE 256
#endif
#ifndef DDP_GSPEC_MIN_IDLE
#define DDP_GSPEC_MIN_IDLE 12   // warps of fully idle CTAs from which on searches are posted (4 and 48 measured: within noise of 12)
#endif
enum { GSPEC_UNITS = 10, GSPEC_MIN_IDLE = DDP_GSPEC_MIN_IDLE, GSPEC_BOARDS = 3 };   // boards per slot: two for searches, one for sweeps
template <class R> struct GBoard {
    Traj<R> t;          // the owner's view of the trajectory when it posted the search
    R xd[9];            // desired terminal state (lives in the owner's shared memory)
    R tau;
    int seq_ctr;        // generations posted so far: searches count in the slot's first board, sweeps in its third; never reset in a launch
    int done;           // units finished
    int retired;        // last generation whose candidates the owner no longer needs
    int claimed_final;  // units that had been claimed when the last search on this board was closed
    struct Res { int ok, slot; long long knots; R cost, costq, logcost, err, cmax; } res[GSPEC_UNITS];
    // Speculative backward sweep (third board of a slot).  When a line search FAILS, or the factorisation does, the next backward
    // pass works on the same linearisation with the regularisation one step up (ddp.cpp:452-474, :476, :297-310): everything it
    // needs exists as soon as the linearisation does, and the Riccati recursion alone is more than half of an iteration of the
    // solves that are left at the end of a batch.  So the owner posts that sweep (one unit) before it starts its own recursion; a
    // warp of an idle CTA runs it into the owner's second gain buffer (K2) next to the owner's recursion and search, and a
    // backward pass that finds the linearisation unchanged and the regularisation it was posted for takes the result (gains by
    // pointer swap, |Qu|_inf, failure flag, knots visited) - the same function on the same inputs, hence the same bits.  Any other
    // backward pass withdraws it (sweep_cancel, looked at once per knot); K2 is not handed out again before sweep_done.
    int sweep;          // 1: this board carries one unit, the sweep (0: a line search, unit k = trial 2^-(k+1))
    int sweep_cancel, sweep_done, sweep_ok;
    long long sweep_knots;
    R sweep_regadd, sweep_errq;
};

#if DDP_GPU
template <class R> DDP_DEVICE bool gspec_ready(const Traj<R> &t) {
    return t.gb != nullptr && *(volatile unsigned int *)(t.gctr + 4) >= (unsigned)GSPEC_MIN_IDLE;
}
// Close a board: no further claims.  Returns how many units were claimed.
DDP_DEVICE int gspec_close(unsigned long long *w) {
    unsigned long long cur = *(volatile unsigned long long *)w;
    while (true) {
        const int n = (int)((cur >> 16) & 0xffff), nx = (int)(cur & 0xffff);
        const int claimed = nx < n ? nx : n;
        const unsigned long long closed = (cur & ~0xffffull) | (unsigned long long)n;
        const unsigned long long old = atomicCAS(w, cur, closed);
        if (old == cur) return claimed;
        cur = old;
    }
}
template <class R> DDP_DEVICE void warp_copy_global(R *DDP_RESTRICT dst, const R *DDP_RESTRICT src, long long n, int lane) {
    dst = as_global(dst); src = as_global(src);
    long long e = lane;
    for (; e + 96 < n; e += 128) {
        const R a = src[e], b = src[e + 32], c = src[e + 64], d = src[e + 96];
        dst[e] = a; dst[e + 32] = b; dst[e + 64] = c; dst[e + 96] = d;
    }
    for (; e < n; e += 32) dst[e] = src[e];
}
#endif

template <class R> DDP_DEVICE bool run_trial(Traj<R> &t, R alpha, R tau, RollOut<R> &ro) {
    return coop_has_helpers(t) ? forward_trial_coop(t, alpha, tau, ro) : forward_trial_solo(t, alpha, tau, ro);
}

// Line search with the filter (ddp.cpp:647-778).
template <class R> DDP_DEVICE_NOINLINE void forward_pass(Traj<R> &t) {
    const int lane_ = t.lane_;
    const long long clk0 = ddp_clock();
    const R tau = rmax(R(0.99), R(1) - t.mu);
    bool failed = true;
    RollOut<R> ro;
    R stepsize = R(0);
    int step = 0;
#if DDP_GPU
    if (gspec_ready(t)) {
        GBoard<R> *b = (GBoard<R> *)t.gb + t.gflip;
        unsigned long long *w = t.gw + t.gflip;
        const int bidx = t.gindex + t.gflip;
        t.gflip ^= 1;
        int seq = 0;
        __threadfence();   // every lane's part of the iterate and the gains before the board is posted
        __syncwarp();
        if (lane_ == 0) {
            // the previous search on this board may have been closed with units in flight: they are long done
            while (*(volatile int *)&b->done < *(volatile int *)&b->claimed_final) __nanosleep(200);
            __threadfence();
            GBoard<R> *b0 = (GBoard<R> *)t.gb;
            seq = b0->seq_ctr + 1;
            b0->seq_ctr = seq;
            b->t = t;
            for (int e = 0; e < 9; e++) b->xd[e] = t.sm[Lay::XD + e];
            b->tau = tau; b->done = 0; b->claimed_final = GSPEC_UNITS;
            b->sweep = 0;
            __threadfence();   // the board before the claim word
            *(volatile unsigned long long *)w = ((unsigned long long)(unsigned)seq << 32) | ((unsigned long long)GSPEC_UNITS << 16);
            __threadfence();
            atomicOr(t.gbits + (bidx >> 5), 1u << (bidx & 31));
            atomicAdd(t.gctr + 5, 1u);
        }
        seq = __shfl_sync(0xffffffffu, seq, 0);
        // the owner's own trial: step 2^0
        t.n_fwd_trials++;
        stepsize = R(1);
        bool acc0 = run_trial(t, stepsize, tau, ro);
        if (acc0) acc0 = filter_try(t, ro.logcost, ro.err);
        int claimed = 0;
#ifdef DDP_SPEC_PROFILE
        const long long clk_w = ddp_clock();   // profiling build: the "sequential rollout" counter shows wait + copy time
#endif
        if (lane_ == 0) {
            claimed = gspec_close(w);
            *(volatile int *)&b->claimed_final = claimed;
            atomicAnd(t.gbits + (bidx >> 5), ~(1u << (bidx & 31)));
        }
        claimed = __shfl_sync(0xffffffffu, claimed, 0);
        if (acc0) {
            if (lane_ == 0) { __threadfence(); *(volatile int *)&b->retired = seq; }   // nobody's candidate is needed
            failed = false;
            step = 0;
        } else {
            if (lane_ == 0) {
                while (*(volatile int *)&b->done < claimed) __nanosleep(200);
            }
            __syncwarp();
            __threadfence();   // acquire: results and candidate buffers of the finished units
            int win = -1;
            for (int u = 0; u < claimed; u++) {
                const typename GBoard<R>::Res *rs = &b->res[u];
                t.n_fwd_trials++;
                t.n_fwd_knots += *(volatile long long *)&rs->knots;
                if (*(volatile int *)&rs->ok) {
                    const R lc = *(volatile R *)&rs->logcost, er = *(volatile R *)&rs->err;
                    if (filter_try(t, lc, er)) { win = u; break; }
                }
            }
            if (win >= 0) {
                const typename GBoard<R>::Res *rs = &b->res[win];
                ro.cost = *(volatile R *)&rs->cost; ro.costq = *(volatile R *)&rs->costq;
                ro.logcost = *(volatile R *)&rs->logcost; ro.err = *(volatile R *)&rs->err; ro.cmax = *(volatile R *)&rs->cmax;
                const R *src = t.ws_all + (long long)(*(volatile int *)&rs->slot) * t.ws_stride;
                warp_copy_global(t.xun, src + t.off_xun, (long long)(t.N + 1) * 20, lane_);
                warp_copy_global(t.sn, src + t.off_sn, (long long)t.MCS * t.NP, lane_);
                if (t.infeas) warp_copy_global(t.yn, src + t.off_yn, (long long)t.MCS * t.NP, lane_);
                __syncwarp();
                failed = false;
                step = win + 1;
                stepsize = R(1);
                for (int k = 0; k < step; k++) stepsize = stepsize * R(0.5);
            } else {
                step = claimed + 1;   // carry on sequentially below
            }
            if (lane_ == 0) { __threadfence(); *(volatile int *)&b->retired = seq; }
        }
        __syncwarp();
#ifdef DDP_SPEC_PROFILE
        t.cyc_seq += ddp_clock() - clk_w;
#endif
    }
    if (failed)
#endif
    for (; step < 11; step++) {
        stepsize = R(1);
        for (int k = 0; k < step; k++) stepsize = stepsize * R(0.5);  // 2^-step exactly (ddp.cpp:670)
        t.n_fwd_trials++;
        if (!run_trial(t, stepsize, tau, ro)) continue;
        if (!filter_try(t, ro.logcost, ro.err)) continue;
        failed = false;
        break;
    }
    if (failed) {
        t.failed = 1;
        t.stepsize = R(0);
    } else {
        t.cost = ro.cost; t.costq = ro.costq; t.logcost = ro.logcost; t.err = ro.err;
        t.cmax = ro.cmax; t.cmax_valid = 1;
        R *tmp;
        tmp = t.xu; t.xu = t.xun; t.xun = tmp;
        tmp = t.s; t.s = t.sn; t.sn = tmp;
        if (t.infeas) { tmp = t.y; t.y = t.yn; t.yn = tmp; }
        t.stepsize = stepsize; t.step = step; t.failed = 0;
        t.lin_valid = 0;
    }
    t.cyc_fwd += ddp_clock() - clk0;
}

#if DDP_GPU
// Owner: post the sweep with regularisation regadd_next on the slot's third board (linearisation valid, K2 free or about to be).
template <class R> DDP_DEVICE_NOINLINE void sweep_post(Traj<R> &t, R regadd_next) {
    GBoard<R> *sb = (GBoard<R> *)t.gb + 2;
    unsigned long long *w = t.gw + 2;
    const int bidx = t.gindex + 2;
    __threadfence();   // the linearisation records (this warp's or its helpers') before the board
    __syncwarp();
    if (t.lane_ == 0) {
        if (t.sweep_b != nullptr)   // a withdrawn sweep may still be writing K2 (for one knot at most)
            while (*(volatile int *)&((GBoard<R> *)t.sweep_b)->sweep_done == 0) __nanosleep(100);
        const int seq = sb->seq_ctr + 1;
        sb->seq_ctr = seq;
        sb->t = t;
        for (int e = 0; e < 9; e++) sb->xd[e] = t.sm[Lay::XD + e];
        sb->sweep = 1; sb->sweep_cancel = 0; sb->sweep_done = 0; sb->sweep_regadd = regadd_next;
        __threadfence();   // the board before the claim word
        *(volatile unsigned long long *)w = ((unsigned long long)(unsigned)seq << 32) | (1ull << 16);
        __threadfence();
        atomicOr(t.gbits + (bidx >> 5), 1u << (bidx & 31));
    }
    t.sweep_b = sb; t.spec_posted = 1; t.spec_regadd = regadd_next;
    __syncwarp();
}
// Owner: close the posted sweep.  want: take its result if an idle warp claimed it (true: ok / e0 / knots / gains are those of the
// sweep).  Otherwise, or when nobody claimed it, the caller runs the recursion itself.
template <class R> DDP_DEVICE_NOINLINE bool sweep_resolve(Traj<R> &t, bool want, bool &ok, R &e0) {
    GBoard<R> *sb = (GBoard<R> *)t.sweep_b;
    const int bidx = t.gindex + 2;
    t.spec_posted = 0;
    int claimed = 0;
    if (t.lane_ == 0) {
        claimed = gspec_close(t.gw + 2);
        atomicAnd(t.gbits + (bidx >> 5), ~(1u << (bidx & 31)));
        if (claimed && !want) *(volatile int *)&sb->sweep_cancel = 1;
        if (claimed && want) {
            while (*(volatile int *)&sb->sweep_done == 0) __nanosleep(100);
            atomicAdd(t.gctr + 8, 1u);
        }
    }
    claimed = __shfl_sync(0xffffffffu, claimed, 0);
    if (!claimed) { t.sweep_b = nullptr; return false; }
    if (!want) return false;   // t.sweep_b stays set: K2 is not handed out again before that sweep has stopped
    __threadfence();   // acquire: the gains in K2 and the result
    ok = *(volatile int *)&sb->sweep_ok != 0;
    e0 = *(volatile R *)&sb->sweep_errq;
    t.n_bwd_knots += *(volatile long long *)&sb->sweep_knots;
    R *tmp = t.K; t.K = t.K2; t.K2 = tmp;
    t.sweep_b = nullptr;
    __syncwarp();
    return true;
}

// A warp of a fully idle CTA: run line-search trials posted by the remaining solves until every trajectory is finished.
template <class R>
DDP_DEVICE_NOINLINE void gspec_helper_loop(const SolveArgs &A, R *sm, const R *tabs, R *ws, int slot, int lane_,
                                           JobBoard<R> *cta_boards, BlockCtl *ctl, int wpb, int me) {
    GBoard<R> *boards = (GBoard<R> *)A.gboards;
    const int nboards = GSPEC_BOARDS * (int)(gridDim.x * (blockDim.x >> 5));
    const WsLay wl = ws_layout(A.N, A.PM, A.fcap);
    sm = as_shared(sm);
    if (lane_ == 0) atomicAdd(A.counter + 4, 1u);
    unsigned idle_ns = DDP_IDLE_NS_MIN;
    while (*(volatile unsigned int *)(A.counter + 3) < (unsigned)A.B) {
        // row units of a trial that another warp of this CTA is running come first: they are short and someone waits for them
        if (ctl != nullptr) {
            bool found = false;
            for (int w = 0; w < wpb; w++) {
                if (w == me) continue;
                const int u = job_claim(cta_boards + w, 0u);
                if (u < 0) continue;
                found = true;
                job_run_unit(cta_boards + w, u, sm, lane_);
                if (lane_ == 0) atomicAdd(A.counter + 2, 1u);
            }
            if (found) { idle_ns = DDP_IDLE_NS_MIN; continue; }
        }
        // find a board with unclaimed units in the bitmap (start position spread over the helpers), claim one unit
        int bi = -1, unit = -1;
        unsigned seq = 0;
        int excl = 0;   // 1: this warp holds the CTA's right to run a unit (many idle CTAs: one unit per CTA, the rest help with its rows)
        if (ctl != nullptr && *(volatile unsigned int *)(A.counter + 4) >= (unsigned)DDP_GSPEC_EXCL_IDLE) {
            int got = 0;
            if (lane_ == 0) got = atomicCAS(&ctl->unit_running, 0, 1) == 0;
            got = __shfl_sync(0xffffffffu, got, 0);
            if (!got) { __nanosleep(DDP_IDLE_NS_MIN); continue; }   // stays close: the running unit posts row jobs every few microseconds
            excl = 1;
        }
        {
            const unsigned int *bits = (const unsigned int *)(A.gwords + nboards);
            const int nwords = (nboards + 31) >> 5;
            const int rot = (slot * 7) % nwords;
            for (int base = 0; base < nwords && bi < 0; base += 32) {
                int wi = base + lane_;
                unsigned wv = 0;
                if (wi < nwords) { wi = (wi + rot) % nwords; wv = *(volatile unsigned int *)(bits + wi); }
                unsigned m = __ballot_sync(0xffffffffu, wv != 0);
                while (m && bi < 0) {
                    const int l = __ffs(m) - 1;
                    m &= m - 1;
                    unsigned word = __shfl_sync(0xffffffffu, wv, l);
                    const int widx = __shfl_sync(0xffffffffu, wi, l);
                    while (word && bi < 0) {
                        const int bit = __ffs(word) - 1;
                        word &= word - 1;
                        const int cand = (widx << 5) + bit;
                        if (cand >= nboards) continue;
                        int u = -1;
                        unsigned sq = 0;
                        if (lane_ == 0) {
                            unsigned long long cur = *(volatile unsigned long long *)(A.gwords + cand);
                            while (true) {
                                const int n = (int)((cur >> 16) & 0xffff), nx = (int)(cur & 0xffff);
                                if ((cur >> 32) == 0 || nx >= n) break;
                                const unsigned long long old = atomicCAS(A.gwords + cand, cur, cur + 1);
                                if (old == cur) { u = nx; sq = (unsigned)(cur >> 32); break; }
                                cur = old;
                            }
                        }
                        u = __shfl_sync(0xffffffffu, u, 0);
                        sq = __shfl_sync(0xffffffffu, sq, 0);
                        if (u >= 0) { bi = cand; unit = u; seq = sq; }
                    }
                }
            }
        }
        if (bi < 0) {
            if (excl && lane_ == 0) *(volatile int *)&ctl->unit_running = 0;
            __nanosleep(idle_ns); if (idle_ns < DDP_IDLE_NS_MAX) idle_ns *= 2;
            continue;
        }
        idle_ns = DDP_IDLE_NS_MIN;
        __threadfence();   // acquire: the board and the owner's arrays as of the posting
        GBoard<R> *b = boards + bi;
        Traj<R> t = b->t;
        t.lane_ = lane_; t.sm = sm; t.tab = tabs + (t.minvo ? 180 : 0);
        t.board = ctl ? (void *)(cta_boards + me) : nullptr; t.ctl = ctl; t.wpb = wpb; t.gb = nullptr;   // rows go to idle warps of THIS CTA
        t.xun = ws + wl.xun; t.sn = ws + wl.sn; t.yn = ws + wl.yn; t.kdx = ws + wl.kdx;
        t.n_fwd_knots = 0; t.cyc_seq = 0;
        if (lane_ < 9) sm[Lay::XD + lane_] = b->xd[lane_];
        __syncwarp();
        if (b->sweep) {
            {   // the backward sweep the owner needs if its iteration fails, into the owner's second gain buffer
                if (excl && lane_ == 0) *(volatile int *)&ctl->unit_running = 0;   // a sweep posts no row jobs: the CTA may run a trial next to it
                t.K = t.K2; t.n_bwd_knots = 0;
                Reg<R, 1> errq;
                const bool sok = riccati(t, b->sweep_regadd, errq, &b->sweep_cancel);
                const R e0 = warp_max(errq, 0, lane_);
                fence_async_all();   // the gains are read by bulk copies of the next trials
                __threadfence();
                __syncwarp();
                if (lane_ == 0) {
                    b->sweep_ok = sok ? 1 : 0; b->sweep_errq = e0; b->sweep_knots = t.n_bwd_knots;
                    __threadfence();
                    *(volatile int *)&b->sweep_done = 1;
                    atomicAdd(A.counter + 7, 1u);
                }
                __syncwarp();
                continue;
            }
        }
        const int step = unit + 1;
        R alpha = R(1);
        for (int k = 0; k < step; k++) alpha = alpha * R(0.5);
        RollOut<R> ro;
        ro.cost = ro.costq = ro.logcost = ro.err = ro.cmax = R(0);
        const bool ok = run_trial(t, alpha, b->tau, ro);
        if (lane_ == 0) {
            typename GBoard<R>::Res *rs = &b->res[unit];
            rs->ok = ok ? 1 : 0; rs->slot = slot; rs->knots = t.n_fwd_knots;
            rs->cost = ro.cost; rs->costq = ro.costq; rs->logcost = ro.logcost; rs->err = ro.err; rs->cmax = ro.cmax;
            atomicAdd(A.counter + 6, 1u);
        }
        __threadfence();   // the candidate buffers and the result before the unit counts as done
        __syncwarp();
        if (lane_ == 0) {
            atomicAdd(&b->done, 1);
            if (excl) *(volatile int *)&ctl->unit_running = 0;
            // the candidate stays untouched until the owner has taken it or discarded the search
            if (ok) while (*(volatile int *)&b->retired < (int)seq && *(volatile unsigned int *)(A.counter + 3) < (unsigned)A.B) __nanosleep(500);
        }
        __syncwarp();
    }
}
#endif

// =============================================================================================
// One polyCurveGeneration (ddp.cpp:5-438) for trajectory `b`, stage `st` of the call.
// =============================================================================================
template <class R> DDP_DEVICE_NOINLINE void solve_one(const SolveArgs &A, int st, int b, R *sm, const R *tabs, R *ws,
                                                      int lane_, void *board = nullptr, void *ctl = nullptr, int wpb = 1) {
    const StageCfg &cfg = A.cfg[st];
    sm = as_shared(sm);
    tabs = as_shared(tabs);
    // NS = knots per trajectory in the caller's arrays (their stride); N = knots of THIS trajectory (ragged batches, e.g. the
    // prefixes of a recorded corridor, teach_repeat_planner.cpp:316-350).  The workspace is laid out for NS.
    const int NS = A.N;
    const int N = A.nknots ? A.nknots[b] : A.N;
    const WsLay wl = ws_layout(NS, A.PM, A.fcap);
    Traj<R> t;
    t.N = N; t.PM = A.PM; t.NP = wl.NP; t.MCS = wl.MCS; t.lane_ = lane_;
    t.board = board; t.ctl = ctl; t.wpb = wpb;
    t.gb = nullptr; t.gw = nullptr; t.gctr = A.counter; t.gflip = 0; t.minvo = cfg.minvo;
    t.spec_on = 0; t.spec_posted = 0; t.spec_regadd = R(0); t.sweep_b = nullptr;
#if DDP_GPU
    if (A.gspec && A.gboards) {
        const long long slot = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
        t.gb = (GBoard<R> *)A.gboards + GSPEC_BOARDS * slot;
        t.gw = A.gwords + GSPEC_BOARDS * slot;
        t.gindex = (int)(GSPEC_BOARDS * slot);
        t.gbits = (unsigned int *)(A.gwords + GSPEC_BOARDS * (long long)gridDim.x * (blockDim.x >> 5));
        t.ws_all = (R *)A.ws; t.ws_stride = A.ws_stride;
        t.off_xun = wl.xun; t.off_sn = wl.sn; t.off_yn = wl.yn; t.off_kdx = wl.kdx;
        t.spec_on = A.spec ? 1 : 0;
    }
#endif
    t.planes = A.planes + (long long)b * NS * A.PM * 4;
    t.nplanes = A.nplanes + (long long)b * NS;
    t.sm = sm;
    t.tab = tabs + (cfg.minvo ? 180 : 0);
    t.xu = ws + wl.xu; t.xun = ws + wl.xun; t.K = ws + wl.K; t.K2 = ws + wl.K2; t.kdx = ws + wl.kdx; t.aux = ws + wl.aux; t.H = ws + wl.H;
    t.s = ws + wl.s; t.sn = ws + wl.sn; t.y = ws + wl.y; t.yn = ws + wl.yn;
    t.filt = ws + wl.filt; t.fcap = A.fcap;
    t.max_vel = (R)A.max_vel; t.max_acc = (R)A.max_acc;
    t.w_snap = (R)cfg.w_snap; t.w_terminal = (R)cfg.w_terminal; t.w_time = (R)cfg.w_time;
    t.margin = cfg.minvo ? R(0) : R(2.0e-4);  // ddp.cpp:1281-1283
    t.time_power = cfg.time_power; t.zero_init = cfg.zero_init; t.line_init = cfg.line_init;
    t.tol = R(1.0e-7);                                 // ddp.cpp:43
    t.reg_base = cfg.zero_init ? R(1.6) : R(4.0);      // ddp.cpp:60-61
    t.n_bwd_sweeps = t.n_bwd_knots = t.n_fwd_trials = t.n_fwd_knots = 0;
    t.cyc_bwd = t.cyc_fwd = 0; t.cyc_ric = t.cyc_seq = 0; t.cyc_t0 = ddp_clock();
#ifdef DDP_TRACE_CYCLES
    t.cyc_seg[0] = t.cyc_seg[1] = t.cyc_seg[2] = t.cyc_seg[3] = 0;
#endif
    t.mu = R(0); t.step = 0; t.failed = 0; t.bfailed = 0; t.lin_valid = 0; t.cmax_valid = 0; t.cmax = R(0);
    const bool from_stage0 = (A.two_stage && st == 1);
    int infeas_in;
    if (from_stage0) infeas_in = A.out[0].infeas_out ? A.out[0].infeas_out[b] : 1;
    else infeas_in = A.two_stage ? 1 : (A.infeas ? A.infeas[b] : cfg.infeas_all);
    t.infeas = infeas_in;
    const double *dur = A.durations + (long long)b * NS;
    const double *ibez = A.init_bez ? A.init_bez + (long long)b * NS * 18 : nullptr;
    if (from_stage0) {
        ibez = A.bez_tmp + (long long)b * NS * 18;
        if (A.out[0].rtn[b] == 2) dur = A.time_tmp + (long long)b * NS;  // UpdateTime, teach_repeat_planner.cpp:911-912
    }

    // ---- setup (ddp.cpp:104-250), lane <-> knot --------------------------------------------------------
    FOR_LANES(lane) {
        if (lane < 9) {
            sm[Lay::XD + lane] = (R)A.xd[(long long)b * 9 + lane];
            t.xu[10 + lane] = (R)A.x0[(long long)b * 9 + lane];
        }
        for (int i = lane; i < N; i += 32) {
            const double T = dur[i];
            R u[10];
            DDP_UNROLL
            for (int e = 0; e < 9; e++) u[e] = R(0);
            u[9] = (R)T;
            if (!cfg.zero_init && !cfg.line_init && ibez) {
                // warm start: Bezier [x*6,y*6,z*6] scaled by T -> monomial coefficients 3..5 (ddp.cpp:167-193, :782-796);
                // (poly2bez * t2tauMat)(j, c) = BERN[c][j] / T^c
                const double bern[3][6] = {{-10, 30, -30, 10, 0, 0}, {5, -20, 30, -20, 5, 0}, {-1, 5, -10, 10, -5, 1}};
                const double Tk = 1.0 / T;
                DDP_UNROLL
                for (int ci = 0; ci < 3; ci++) {
                    double pw = 1.0;
                    for (int k = 0; k < ci + 3; k++) pw = (k == 0) ? Tk : pw * Tk;
                    DDP_UNROLL
                    for (int a = 0; a < 3; a++) {
                        double accv = 0.0;
                        DDP_UNROLL
                        for (int j = 0; j < 6; j++) accv += (T * ibez[(long long)i * 18 + a * 6 + j]) * (bern[ci][j] * pw);
                        u[ci * 3 + a] = (R)accv;
                    }
                }
            }
            if (cfg.line_init) {
                // straight-line initialisation (ddp.cpp:195-247): rest-to-rest quintic between consecutive seeds,
                // segment time doubled (at most 5 times) until every constraint of the knot is negative
                R z[19];
                DDP_UNROLL
                for (int e = 0; e < 19; e++) z[e] = R(0);
                R p1[3];
                DDP_UNROLL
                for (int a = 0; a < 3; a++) {
                    z[10 + a] = (i == 0) ? (R)A.x0[(long long)b * 9 + a] : (R)A.seeds[((long long)b * NS + i) * 3 + a];
                    p1[a] = (i == N - 1) ? (R)A.xd[(long long)b * 9 + a] : (R)A.seeds[((long long)b * NS + i + 1) * 3 + a];
                }
                z[9] = u[9];
                int vio = 1, cnt = 0;
                while (vio && cnt <= 4) {
                    const R Tk = z[9], Tk2 = Tk * Tk, Tk3 = Tk2 * Tk, Tk4 = Tk3 * Tk, Tk5 = Tk4 * Tk;
                    const R Gi[3] = {R(10.0) / Tk3, R(-15.0) / Tk4, R(6.0) / Tk5};   // first column of G^-1
                    DDP_UNROLL
                    for (int a = 0; a < 3; a++) {
                        const R rhs = p1[a] - z[10 + a];   // only the position rows of xnext - F xcur are nonzero
                        DDP_UNROLL
                        for (int c = 0; c < 3; c++) z[c * 3 + a] = Gi[c] * rhs;
                    }
                    int all_neg = 1;
                    visit_rows(row_ctx(t), i, z, [&](int, R c) { if (!(c < R(0))) all_neg = 0; });
                    if (all_neg) vio = 0;
                    else { z[9] = R(2) * Tk; cnt++; }
                }
                DDP_UNROLL
                for (int e = 0; e < 10; e++) u[e] = z[e];
            }
            DDP_UNROLL
            for (int e = 0; e < 10; e++) t.xu[(long long)i * 20 + e] = u[e];
            for (int r = 0; r < t.MCS; r++) {
                t.s[row_ofs(t.MCS, r, i)] = R(0.1);   // ddp.cpp:150-151
                t.y[row_ofs(t.MCS, r, i)] = R(0.01);
            }
        }
    }
    WARP_SYNC();
    initial_roll(t);  // ddp.cpp:252
    R dummy0 = R(0), dummy1 = R(0);
    if (cfg.line_init) {  // ddp.cpp:255-269
        if (!scan_constraints(t, 2, R(0), dummy0, dummy1)) t.infeas = 0;
    }
    t.mu = t.cost / R(N) / R(6 * t.nplanes[0] + 55);  // ddp.cpp:281 (hazard H3)
    reset_filter(t);
    t.reg = R(0); t.bfailed = 0;  // resetreg
    if (cfg.line_init) t.reg = R(10);

    int rtn = 0, infeas_out = infeas_in, line_failed_out = 1;
    R cost_prev = t.cost;
    int iter = 0, bp_no_upd_count = 0, no_upd_count = 0;
    const int bp_no_upd_count_max = 20;
    int trace_n = 0;
#ifdef DDP_TRACE_CYCLES
    long long trc[4] = {0, 0, 0, 0};
    unsigned long long trace_t0 = 0;
#if DDP_GPU
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(trace_t0));
#endif
#endif
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
                    tr[6] = t.stepsize; tr[7] = t.opterr; tr[8] = t.step; tr[9] = t.failed; tr[10] = n_bwd;
                    tr[11] = (double)(ddp_clock() - t.cyc_t0);   // SM cycles since the solve began; the host converts with the device's clock
#ifdef DDP_TRACE_CYCLES   // diagnostic build (tools/timeline.py --cycles): kilo-cycles of this iteration per phase instead of four of the scalars
                    tr[1] = (double)((t.cyc_bwd - trc[0]) >> 10); tr[2] = (double)((t.cyc_fwd - trc[1]) >> 10);
                    tr[3] = (double)((t.cyc_ric - trc[2]) >> 10); tr[7] = (double)((t.cyc_seq - trc[3]) >> 10);
                    trc[0] = t.cyc_bwd; trc[1] = t.cyc_fwd; trc[2] = t.cyc_ric; trc[3] = t.cyc_seq;
                    tr[0] = (double)(t.cyc_seg[0] >> 10); tr[4] = (double)(t.cyc_seg[1] >> 10); tr[5] = (double)(t.cyc_seg[2] >> 10); tr[6] = (double)(t.cyc_seg[3] >> 10);
                    t.cyc_seg[0] = t.cyc_seg[1] = t.cyc_seg[2] = t.cyc_seg[3] = 0;
#endif
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
            t.lin_valid = 0;
        }
        // any c >= 2e-4 at the current iterate (ddp.cpp:346-355, hazard H8)?  The accepted trial evaluated exactly these
        // constraint values, so their maximum is at hand; a scan is only needed before the first accepted trial.
        const bool violated = t.cmax_valid ? (t.cmax >= R(2.0e-4)) : scan_constraints(t, 1, R(2.0e-4), dummy0, dummy1);
        if (!violated) {  // ddp.cpp:356-390
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
        if (cfg.line_init) {   // ddp.cpp:398-409
            if (t.stepsize < R(1.0e-6)) no_upd_count++;
            else no_upd_count = 0;
            if (no_upd_count > 100) break;
        }
    }
    if (A.trace && b == 0 && (st == 1 || !A.two_stage) && A.trace_len) {
        FOR_LANES(lane) { if (lane == 0) *A.trace_len = trace_n; }
    }
#if DDP_GPU
    {   // a sweep of the last backward pass: withdrawn, gone before the workspace is used again
        bool dummy_ok; R dummy_e;
        if (t.spec_posted) sweep_resolve(t, false, dummy_ok, dummy_e);
        if (t.sweep_b != nullptr) {
            if (lane_ == 0) while (*(volatile int *)&((GBoard<R> *)t.sweep_b)->sweep_done == 0) __nanosleep(100);
            __syncwarp();
        }
    }
#endif

    // ---- outputs (ddp.cpp:418-437), lane <-> knot ---------------------------------------------------------
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
                S[4] = t.cyc_bwd; S[5] = t.cyc_fwd; S[6] = ddp_clock() - t.cyc_t0;
#if defined(DDP_TRACE_CYCLES) && DDP_GPU   // diagnostic build (tools/tail_who.py): wall-clock start and end of the solve (ns) instead of the phase cycles
                { unsigned long long now; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now)); S[4] = (long long)trace_t0; S[5] = (long long)now; }
#endif
                S[7] = (t.cyc_ric >> 10) | ((t.cyc_seq >> 10) << 32);   // kilo-cycles: Riccati | sequential rollout
            }
        }
        if (lane < 9 && O.x_final) O.x_final[(long long)b * 9 + lane] = (double)t.xu[(long long)N * 20 + 10 + lane];
        for (int i = lane; i < N; i += 32) {
            R z[19], tp[6], m[9], mu9[9];
            DDP_UNROLL
            for (int e = 0; e < 19; e++) z[e] = t.xu[(long long)i * 20 + e];
            const R T = z[9];
            time_powers(T, tp);
            rmat<R>(0, tp, m);
            rmat_times_u(m, z, mu9);
            if (O.jerk) O.jerk[(long long)b * NS + i] = (double)dot9(z, mu9);  // finalroll, ddp.cpp:1624-1634
            if (O.poly_time) O.poly_time[(long long)b * NS + i] = (double)T;
            if (carry) A.time_tmp[(long long)b * NS + i] = (double)T;
            DDP_UNROLL
            for (int l = 0; l < 6; l++) {
                DDP_UNROLL
                for (int a = 0; a < 3; a++) {
                    // PolyCoeff row = [Ek_inv * x, u[0:9]] (ddp.cpp:814-823); index l*3+a
                    R pc = z[zidx(l, a)];
                    if (l == 2) pc = pc * R(0.5);
                    if (O.poly_coeff) O.poly_coeff[((long long)b * NS + i) * 18 + l * 3 + a] = (double)pc;
                }
            }
            // BezCoeff = (1/T) * Bezier control points (poly2bezFunc, ddp.cpp:799-812: the inverse of
            // poly2bez*t2tauMat is the Bezier table with column k scaled by T^k), re-laid [x*6,y*6,z*6]
            DDP_UNROLL
            for (int j = 0; j < 6; j++) {
                DDP_UNROLL
                for (int a = 0; a < 3; a++) {
                    R accv = R(0), pw = R(1);
                    DDP_UNROLL
                    for (int k = 0; k < 6; k++) {
                        R ck = z[zidx(k, a)];
                        if (k == 2) ck = ck * R(0.5);
                        accv += (tabs[j * 6 + k] * pw) * ((R(1) / T) * ck);
                        pw = pw * T;
                    }
                    const double bzv = (double)accv;
                    if (O.bez_coeff) O.bez_coeff[((long long)b * NS + i) * 18 + a * 6 + j] = bzv;
                    if (carry) A.bez_tmp[((long long)b * NS + i) * 18 + a * 6 + j] = bzv;
                }
            }
        }
    }
    WARP_SYNC();  // stage-0 outputs (bez_tmp/time_tmp/rtn) are read by other lanes of this warp in stage 1
}

}  // namespace ddp
#endif
