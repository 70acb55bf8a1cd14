/*
 * ipddp_oracle.c -- CPU fp64 restatement of ntu-caokun/DIRECT's IPDDP optimiser.
 * TEST INFRASTRUCTURE ONLY (see ipddp_oracle.h).  "ddp.cpp" below means
 * /root/reference/global_planner/src/ddp_optimizer.cpp, "trp.cpp" means
 * /root/reference/global_planner/src/teach_repeat_planner.cpp.
 *
 * The arithmetic is kept dense and in the reference's statement order on purpose (this is
 * the checker, not the product): Jacobians cx/cu are materialised, products are plain
 * triple loops.  Build with -O2 and WITHOUT -ffast-math: the filter and the
 * fraction-to-boundary tests rely on IEEE comparisons (including NaN behaviour).
 */
#include "ipddp_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define NX 9
#define NU 10
#define NZ 18
#define KC 6 /* num_ctrlP */

/* ---- basis tables, ddp.cpp:62-96 (values) and :1543-1560 (d/dT tables, always MINVO: hazard H1) ---- */
static const double MINVO_P[6][6] = {
    {1.0, -0.06471861202, -0.03728008486, -0.02577637794, -0.02027573243, -0.01678273037},
    {1.0, 0.03314986096, -0.06548114211, -0.05530463802, -0.04362718953, -0.03671639115},
    {1.0, 0.3375528997, 0.05836232552, -0.02920033165, -0.04690387913, -0.04376447947},
    {1.0, 0.6624471003, 0.3832565261, 0.1916286091, 0.06985980172, -0.002892843108},
    {1.0, 0.966850139, 0.868219136, 0.7594116288, 0.6521050661, 0.5510660979},
    {1.0, 1.064718612, 1.092157139, 1.108091959, 1.118023718, 1.123960059}};
static const double MINVO_V[5][6] = {
    {0, 1.0, -0.1423379297, -0.1332742327, -0.1242105357, -0.126304257},
    {0, 1.0, 0.1887439858, -0.1831318297, -0.2466606848, -0.2321393311},
    {0, 1.0, 1.0, 0.5411016575, 0.08220331498, -0.2433474658},
    {0, 1.0, 1.811256014, 2.250636213, 2.381669451, 2.282405938},
    {0, 1.0, 2.14233793, 3.293739556, 4.445141183, 5.585385392}};
static const double MINVO_A[4][6] = {
    {0, 0, 2.0, -0.4472869252, -0.6133793313, -0.6406553622},
    {0, 0, 2.0, 1.223711659, -0.5552714346, -1.854618819},
    {0, 0, 2.0, 4.776288341, 6.54988193, 6.841145057},
    {0, 0, 2.0, 6.447286925, 13.17576837, 22.04662796}};
static const double BEZ_P[6][6] = {
    {1.0, 0, 0, 0, 0, 0},       {1.0, 0.2, 0, 0, 0, 0},       {1.0, 0.4, 0.1, 0, 0, 0},
    {1.0, 0.6, 0.3, 0.1, 0, 0}, {1.0, 0.8, 0.6, 0.4, 0.2, 0}, {1.0, 1.0, 1.0, 1.0, 1.0, 1.0}};
static const double BEZ_V[5][6] = {{0, 1.0, 0, 0, 0, 0},
                                   {0, 1.0, 0.5, 0, 0, 0},
                                   {0, 1.0, 1.0, 0.5, 0, 0},
                                   {0, 1.0, 1.5, 1.5, 1.0, 0},
                                   {0, 1.0, 2.0, 3.0, 4.0, 5.0}};
static const double BEZ_A[4][6] = {
    {0, 0, 2.0, 0, 0, 0}, {0, 0, 2.0, 2.0, 0, 0}, {0, 0, 2.0, 4.0, 4.0, 0}, {0, 0, 2.0, 6.0, 12.0, 20.0}};
/* d/dT tables: entry = coefficient, multiplied by T^pw with pw given by the column. */
static const double DT_P[6][6] = {
    {0, -0.06471861202, -0.07456016972, -0.07732913382, -0.08110292972, -0.08391365186},
    {0, 0.03314986096, -0.1309622842, -0.1659139141, -0.1745087581, -0.1835819558},
    {0, 0.3375528997, 0.116724651, -0.08760099494, -0.1876155165, -0.2188223973},
    {0, 0.6624471003, 0.7665130522, 0.5748858272, 0.2794392069, -0.01446421554},
    {0, 0.966850139, 1.736438272, 2.278234886, 2.608420264, 2.755330489},
    {0, 1.064718612, 2.184314278, 3.324275878, 4.472094873, 5.619800295}};
static const double DT_V[5][6] = {{0, 0, -0.1423379297, -0.2665484655, -0.3726316072, -0.5052170278},
                                  {0, 0, 0.1887439858, -0.3662636595, -0.7399820545, -0.9285573245},
                                  {0, 0, 1.0, 1.082203315, 0.2466099449, -0.9733898632},
                                  {0, 0, 1.811256014, 4.501272426, 7.145008354, 9.129623752},
                                  {0, 0, 2.14233793, 6.587479113, 13.33542355, 22.34154157}};
static const double DT_A[4][6] = {{0, 0, 0, -0.4472869252, -1.226758663, -1.921966087},
                                  {0, 0, 0, 1.223711659, -1.110542869, -5.563856457},
                                  {0, 0, 0, 4.776288341, 13.09976386, 20.52343517},
                                  {0, 0, 0, 6.447286925, 26.35153674, 66.13988387}};
/* Bernstein <- monomial map of t2tau (ddp.cpp:1050-1055), tempm; poly2bez = tempm^T. */
static const double BERN[6][6] = {{1, 0, 0, 0, 0, 0},      {-5, 5, 0, 0, 0, 0},    {10, -20, 10, 0, 0, 0},
                                  {-10, 30, -30, 10, 0, 0}, {5, -20, 30, -20, 5, 0}, {-1, 5, -10, 10, -5, 1}};
static const double EK_INV[3] = {1.0, 1.0, 0.5}; /* ddp.cpp:101-103 */

typedef struct {
    const ipddp_problem *pb;
    int N, rows;          /* rows = sum of m_c */
    int *mc, *off;        /* per knot: constraint count and row offset */
    const double (*tp)[6];/* value tables chosen by minvo flag */
    const double (*tv)[6];
    const double (*ta)[6];
    /* fwdPass state */
    double *x, *u, *c, *s, *y, *q;
    double *fx, *fu, *qu, *quu, *cx, *cu;
    double px[NX];
    double cost, costq, logcost, err, stepsize;
    int step, failed;
    double *filter; int nfilter, capfilter;
    /* bwdPass state */
    double *ku, *Ku, *ks, *ky, *Ks, *Ky;
    double reg, opterr, dV[2];
    int bfailed;
    /* algParam */
    double mu, tol; int maxiter, infeas;
    double reg_base;
    /* scratch for forwardpass */
    double *xn, *un, *cn, *sn, *yn, *qn;
    /* hidden state left by computecminvo (hazard H2) */
    double Bp[6][6], Bv[5][6], Ba[4][6];
    /* stats */
    long n_bwd_sweeps, n_bwd_knots, n_fwd_trials, n_fwd_knots;
} ctx_t;

/* ---------------- model pieces ---------------- */

/* ddp.cpp:836-890 time2barFkbarGk (sys_order == 3). */
static void FG_of_T(double Tk, double F[3][3], double G[3][3]) {
    double Tk2 = Tk * Tk, Tk3 = Tk2 * Tk, Tk4 = Tk3 * Tk, Tk5 = Tk4 * Tk;
    F[0][0] = 1.0; F[0][1] = Tk;  F[0][2] = Tk2 / 2.0;
    F[1][0] = 0.0; F[1][1] = 1.0; F[1][2] = Tk;
    F[2][0] = 0.0; F[2][1] = 0.0; F[2][2] = 1.0;
    G[0][0] = Tk3;     G[0][1] = Tk4;      G[0][2] = Tk5;
    G[1][0] = 3 * Tk2; G[1][1] = 4 * Tk3;  G[1][2] = 5 * Tk4;
    G[2][0] = 6 * Tk;  G[2][1] = 12 * Tk2; G[2][2] = 20 * Tk3;
}
/* ddp.cpp:892-962 time2barFkprimebarGkprime. */
static void FGprime_of_T(double Tk, double Fp[3][3], double Gp[3][3]) {
    double Tk2 = Tk * Tk, Tk3 = Tk2 * Tk, Tk4 = Tk3 * Tk;
    memset(Fp, 0, 9 * sizeof(double));
    Fp[0][1] = 1; Fp[0][2] = Tk; Fp[1][2] = 1;
    Gp[0][0] = 3 * Tk2; Gp[0][1] = 4 * Tk3;  Gp[0][2] = 5 * Tk4;
    Gp[1][0] = 6 * Tk;  Gp[1][1] = 12 * Tk2; Gp[1][2] = 20 * Tk3;
    Gp[2][0] = 6;       Gp[2][1] = 24 * Tk;  Gp[2][2] = 60 * Tk2;
}
/* ddp.cpp:964-1015 time2barR. */
static void R_of_T(double Tk, double R[3][3], double Rp[3][3], double Rpp[3][3]) {
    double Tk2 = Tk * Tk, Tk3 = Tk2 * Tk, Tk4 = Tk3 * Tk, Tk5 = Tk4 * Tk;
    R[0][0] = 36 * Tk;   R[0][1] = 72 * Tk2;  R[0][2] = 120 * Tk3;
    R[1][0] = 72 * Tk2;  R[1][1] = 192 * Tk3; R[1][2] = 360 * Tk4;
    R[2][0] = 120 * Tk3; R[2][1] = 360 * Tk4; R[2][2] = 720 * Tk5;
    Rp[0][0] = 36;        Rp[0][1] = 144 * Tk;   Rp[0][2] = 360 * Tk2;
    Rp[1][0] = 144 * Tk;  Rp[1][1] = 576 * Tk2;  Rp[1][2] = 1440 * Tk3;
    Rp[2][0] = 360 * Tk2; Rp[2][1] = 1440 * Tk3; Rp[2][2] = 3600 * Tk4;
    Rpp[0][0] = 0;        Rpp[0][1] = 144;        Rpp[0][2] = 720 * Tk;
    Rpp[1][0] = 144;      Rpp[1][1] = 1152 * Tk;  Rpp[1][2] = 4320 * Tk2;
    Rpp[2][0] = 720 * Tk; Rpp[2][1] = 4320 * Tk2; Rpp[2][2] = 14400 * Tk3;
}

/* ddp.cpp:1062-1067 computenextx: x+ = (F (x) I3) x + (G (x) I3) u[0:9]. */
static void computenextx(const double *x, const double *u, double *xn) {
    double F[3][3], G[3][3];
    FG_of_T(u[9], F, G);
    for (int i = 0; i < 3; i++)
        for (int k = 0; k < 3; k++) {
            double a = 0.0, b = 0.0;
            for (int j = i; j < 3; j++) a += F[i][j] * x[j * 3 + k]; /* triangularView<Upper> */
            for (int j = 0; j < 3; j++) b += G[i][j] * u[j * 3 + k];
            xn[i * 3 + k] = a + b;
        }
}

/* u^T (M (x) I3) u for a 3x3 M. */
static double quad3(const double M[3][3], const double *u) {
    double acc = 0.0;
    for (int i = 0; i < 3; i++)
        for (int k = 0; k < 3; k++) {
            double t = 0.0;
            for (int j = 0; j < 3; j++) t += M[i][j] * u[j * 3 + k];
            acc += u[i * 3 + k] * t;
        }
    return acc;
}

/* ddp.cpp:1294-1305 computeq. */
static double computeq(const ctx_t *C, const double *u) {
    double R[3][3], Rp[3][3], Rpp[3][3];
    R_of_T(u[9], R, Rp, Rpp);
    double jerk = quad3(R, u);
    if (C->pb->time_power == 2) return 0.5 * C->pb->w_snap * jerk + 0.5 * u[9] * C->pb->w_time * u[9];
    return 0.5 * C->pb->w_snap * jerk + 0.5 * C->pb->w_time * u[9];
}

/* ddp.cpp:1289-1292 computep with Pmat = w_terminal * I (ddp.cpp:112). */
static double computep(const ctx_t *C, const double *x) {
    double acc = 0.0;
    for (int i = 0; i < NX; i++) {
        double d = x[i] - C->pb->xd[i];
        acc += d * (C->pb->w_terminal * d);
    }
    return 0.5 * acc;
}

/* ddp.cpp:1132-1285 computecminvo.  Leaves the T-scaled tables in C->Bp/Bv/Ba (hazard H2). */
static void computecminvo(ctx_t *C, const double *x, const double *u, int knot, double *c) {
    const ipddp_problem *pb = C->pb;
    const int P = pb->nplanes[knot];
    const double *pl = pb->planes + (size_t)knot * pb->P_max * 4;
    double coef[KC][3];
    for (int a = 0; a < 3; a++) {
        for (int k = 0; k < 3; k++) coef[k][a] = x[k * 3 + a] * EK_INV[k];
        for (int k = 3; k < KC; k++) coef[k][a] = u[(k - 3) * 3 + a];
    }
    double Tkv[7];
    Tkv[0] = u[9];
    for (int i = 1; i < 7; i++) Tkv[i] = Tkv[i - 1] * Tkv[0];
    for (int j = 0; j < KC; j++) {
        C->Bp[j][0] = 1.0;
        for (int i = 1; i < KC; i++) C->Bp[j][i] = C->tp[j][i] * Tkv[i - 1];
    }
    double pos[KC][3];
    for (int j = 0; j < KC; j++)
        for (int a = 0; a < 3; a++) {
            double acc = 0.0;
            for (int k = 0; k < KC; k++) acc += C->Bp[j][k] * coef[k][a];
            pos[j][a] = acc;
        }
    for (int j = 0; j < KC; j++)
        for (int k = 0; k < P; k++)
            c[j * P + k] = pl[k * 4 + 0] * pos[j][0] + pl[k * 4 + 1] * pos[j][1] + pl[k * 4 + 2] * pos[j][2] + pl[k * 4 + 3];

    double tempv[NZ];
    for (int k = 0; k < 3; k++)
        for (int a = 0; a < 3; a++) tempv[k * 3 + a] = EK_INV[k] * x[k * 3 + a];
    for (int i = 0; i < 9; i++) tempv[9 + i] = u[i];

    for (int j = 0; j < KC - 1; j++) {
        C->Bv[j][0] = 0.0;
        C->Bv[j][1] = C->tv[j][1];
        for (int i = 2; i < KC; i++) C->Bv[j][i] = C->tv[j][i] * Tkv[i - 2];
    }
    double *cv = c + KC * P;
    for (int j = 0; j < KC - 1; j++)
        for (int l = 0; l < 3; l++) {
            double acc = 0.0;
            for (int k = 1; k < KC; k++) acc += C->Bv[j][k] * tempv[k * 3 + l];
            cv[j * 3 + l] = acc - pb->max_vel;
            cv[15 + j * 3 + l] = -acc - pb->max_vel;
        }
    for (int j = 0; j < KC - 2; j++) {
        C->Ba[j][0] = 0.0;
        C->Ba[j][1] = 0.0;
        C->Ba[j][2] = C->ta[j][2];
        for (int i = 3; i < KC; i++) C->Ba[j][i] = C->ta[j][i] * Tkv[i - 3];
    }
    double *ca = cv + 30;
    for (int j = 0; j < KC - 2; j++)
        for (int l = 0; l < 3; l++) {
            double acc = 0.0;
            for (int k = 2; k < KC; k++) acc += C->Ba[j][k] * tempv[k * 3 + l];
            ca[j * 3 + l] = acc - pb->max_acc;
            ca[12 + j * 3 + l] = -acc - pb->max_acc;
        }
    ca[24] = -u[9] + 0.3;
    if (!pb->minvo) {
        int mc = KC * P + 55;
        for (int r = 0; r < mc; r++) c[r] = c[r] - 2.0e-4;
    }
}

/* ddp.cpp:1309-1316 computeall = computeprelated + computefrelated + computeqrelated +
 * computecrelatedminvo. */
static void computeall(ctx_t *C) {
    const ipddp_problem *pb = C->pb;
    const int N = C->N;
    /* computeprelated, ddp.cpp:1318-1323 */
    for (int i = 0; i < NX; i++) C->px[i] = pb->w_terminal * (C->x[(size_t)N * NX + i] - pb->xd[i]);

    for (int i = 0; i < N; i++) {
        const double *x = C->x + (size_t)i * NX;
        const double *u = C->u + (size_t)i * NU;
        const double T = u[9];
        /* computefrelated, ddp.cpp:1325-1336 */
        double F[3][3], G[3][3], Fp[3][3], Gp[3][3];
        FG_of_T(T, F, G);
        FGprime_of_T(T, Fp, Gp);
        double *fx = C->fx + (size_t)i * 81, *fu = C->fu + (size_t)i * 90;
        memset(fx, 0, 81 * sizeof(double));
        memset(fu, 0, 90 * sizeof(double));
        for (int a = 0; a < 3; a++)
            for (int b = 0; b < 3; b++)
                for (int k = 0; k < 3; k++) {
                    fx[(a * 3 + k) * 9 + (b * 3 + k)] = F[a][b];
                    fu[(a * 3 + k) * 10 + (b * 3 + k)] = G[a][b];
                }
        for (int a = 0; a < 3; a++)
            for (int k = 0; k < 3; k++) {
                double s1 = 0.0, s2 = 0.0;
                for (int b = a + 1; b < 3; b++) s1 += Fp[a][b] * x[b * 3 + k]; /* StrictlyUpper */
                for (int b = 0; b < 3; b++) s2 += Gp[a][b] * u[b * 3 + k];
                fu[(a * 3 + k) * 10 + 9] = s1 + s2;
            }
        /* computeqrelated, ddp.cpp:1338-1368 (qx = qxx = qxu = 0) */
        double R[3][3], Rp[3][3], Rpp[3][3];
        R_of_T(T, R, Rp, Rpp);
        double *qu = C->qu + (size_t)i * NU, *quu = C->quu + (size_t)i * 100;
        double Ru[9], Rpu[9];
        for (int a = 0; a < 3; a++)
            for (int k = 0; k < 3; k++) {
                double s1 = 0.0, s2 = 0.0;
                for (int b = 0; b < 3; b++) {
                    s1 += R[a][b] * u[b * 3 + k];
                    s2 += Rp[a][b] * u[b * 3 + k];
                }
                Ru[a * 3 + k] = s1;
                Rpu[a * 3 + k] = s2;
            }
        double uRpu = quad3(Rp, u), uRppu = quad3(Rpp, u);
        memset(quu, 0, 100 * sizeof(double));
        for (int r = 0; r < 9; r++) qu[r] = pb->w_snap * Ru[r];
        for (int a = 0; a < 3; a++)
            for (int b = 0; b < 3; b++)
                for (int k = 0; k < 3; k++) quu[(a * 3 + k) * 10 + (b * 3 + k)] = pb->w_snap * R[a][b];
        for (int r = 0; r < 9; r++) {
            quu[r * 10 + 9] = pb->w_snap * Rpu[r];
            quu[9 * 10 + r] = pb->w_snap * Rpu[r];
        }
        if (pb->time_power == 2) {
            qu[9] = pb->w_time * T + 0.5 * pb->w_snap * uRpu;
            quu[99] = pb->w_time + 0.5 * pb->w_snap * uRppu;
        } else {
            qu[9] = 0.5 * pb->w_time + 0.5 * pb->w_snap * uRpu;
            quu[99] = 0.5 * pb->w_snap * uRppu;
        }
        /* computecrelatedminvo, ddp.cpp:1455-1604 */
        const int P = pb->nplanes[i], mc = C->mc[i];
        const double *pl = pb->planes + (size_t)i * pb->P_max * 4;
        double *c = C->c + C->off[i];
        double *cx = C->cx + (size_t)C->off[i] * NX, *cu = C->cu + (size_t)C->off[i] * NU;
        computecminvo(C, x, u, i, c); /* refreshes C->Bp/Bv/Ba at (x_i,u_i) */
        double tempv[NZ];
        for (int k = 0; k < 3; k++)
            for (int a = 0; a < 3; a++) tempv[k * 3 + a] = EK_INV[k] * x[k * 3 + a];
        for (int r = 0; r < 9; r++) tempv[9 + r] = u[r];
        double T2 = T * T, T3 = T2 * T, T4 = T3 * T;
        double Tp[6] = {0, 1.0, T, T2, T3, T4};   /* DT_P column k carries T^(k-1) */
        double Tv[6] = {0, 0, 1.0, T, T2, T3};    /* DT_V column k carries T^(k-2) */
        double Ta[6] = {0, 0, 0, 1.0, T, T2};     /* DT_A column k carries T^(k-3) */
        memset(cx, 0, (size_t)mc * NX * sizeof(double));
        memset(cu, 0, (size_t)mc * NU * sizeof(double));
        for (int j = 0; j < KC; j++)
            for (int ld = 0; ld < P; ld++) {
                int r = j * P + ld;
                double dt = 0.0;
                for (int k = 0; k < KC; k++)
                    for (int a = 0; a < 3; a++) {
                        double h = C->Bp[j][k] * pl[ld * 4 + a];
                        if (k < 3) cx[r * NX + k * 3 + a] = h * EK_INV[k];
                        else cu[r * NU + (k - 3) * 3 + a] = h;
                        dt += (DT_P[j][k] * Tp[k]) * pl[ld * 4 + a] * tempv[k * 3 + a];
                    }
                cu[r * NU + 9] = dt;
            }
        int rv = KC * P;
        for (int j = 0; j < KC - 1; j++)
            for (int l = 0; l < 3; l++) {
                int r = rv + j * 3 + l;
                double dt = 0.0;
                for (int k = 1; k < KC; k++) {
                    double h = C->Bv[j][k];
                    if (k < 3) { cx[r * NX + k * 3 + l] = h * EK_INV[k]; cx[(r + 15) * NX + k * 3 + l] = -(h * EK_INV[k]); }
                    else { cu[r * NU + (k - 3) * 3 + l] = h; cu[(r + 15) * NU + (k - 3) * 3 + l] = -h; }
                    if (k >= 2) dt += (DT_V[j][k] * Tv[k]) * tempv[k * 3 + l];
                }
                cu[r * NU + 9] = dt;
                cu[(r + 15) * NU + 9] = -dt;
            }
        int ra = rv + 30;
        for (int j = 0; j < KC - 2; j++)
            for (int l = 0; l < 3; l++) {
                int r = ra + j * 3 + l;
                double dt = 0.0;
                for (int k = 2; k < KC; k++) {
                    double h = C->Ba[j][k];
                    if (k < 3) { cx[r * NX + k * 3 + l] = h * EK_INV[k]; cx[(r + 12) * NX + k * 3 + l] = -(h * EK_INV[k]); }
                    else { cu[r * NU + (k - 3) * 3 + l] = h; cu[(r + 12) * NU + (k - 3) * 3 + l] = -h; }
                    if (k >= 3) dt += (DT_A[j][k] * Ta[k]) * tempv[k * 3 + l];
                }
                cu[r * NU + 9] = dt;
                cu[(r + 12) * NU + 9] = -dt;
            }
        cu[(mc - 1) * NU + 9] = -1.0;
    }
}

/* ddp.cpp:1608-1620 initialroll. */
static void initialroll(ctx_t *C) {
    double qs = 0.0;
    for (int i = 0; i < C->N; i++) {
        const double *x = C->x + (size_t)i * NX, *u = C->u + (size_t)i * NU;
        computecminvo(C, x, u, i, C->c + C->off[i]);
        C->q[i] = computeq(C, u);
        qs += C->q[i];
        computenextx(x, u, C->x + (size_t)(i + 1) * NX);
    }
    C->cost = qs + computep(C, C->x + (size_t)C->N * NX);
    C->costq = qs;
}

/* ddp.cpp:1636-1662 resetfilter. */
static void resetfilter(ctx_t *C) {
    C->logcost = C->cost;
    C->err = 0.0;
    if (C->infeas) {
        for (int i = 0; i < C->N; i++) {
            double ls = 0.0, e1 = 0.0;
            for (int r = 0; r < C->mc[i]; r++) {
                ls += log(C->y[C->off[i] + r]);
                e1 += fabs(C->c[C->off[i] + r] + C->y[C->off[i] + r]);
            }
            C->logcost -= C->mu * ls;
            C->err += e1;
        }
        if (C->err < C->tol) C->err = 0.0;
    } else {
        for (int i = 0; i < C->N; i++) {
            double ls = 0.0;
            for (int r = 0; r < C->mc[i]; r++) ls += log(-C->c[C->off[i] + r]);
            C->logcost -= C->mu * ls;
        }
    }
    C->nfilter = 1;
    C->filter[0] = C->logcost;
    C->filter[1] = C->err;
    C->step = 0;
    C->failed = 0;
}

/* Eigen::LLT on an n x n matrix (lower part used), textbook unblocked algorithm (the one Eigen
 * runs below size 32); returns 0 on success, 1 on a non-positive pivot (NumericalIssue). */
static int llt10(double *A, int n) {
    for (int k = 0; k < n; k++) {
        double x = A[k * n + k];
        for (int j = 0; j < k; j++) x -= A[k * n + j] * A[k * n + j];
        if (x <= 0.0) return 1;
        x = sqrt(x);
        A[k * n + k] = x;
        for (int i = k + 1; i < n; i++) {
            double v = A[i * n + k];
            for (int j = 0; j < k; j++) v -= A[i * n + j] * A[k * n + j];
            A[i * n + k] = v / x;
        }
    }
    return 0;
}
static void llt_solve(const double *L, int n, double *b, int nrhs, int ldb) {
    for (int c = 0; c < nrhs; c++) {
        for (int i = 0; i < n; i++) {
            double v = b[i * ldb + c];
            for (int j = 0; j < i; j++) v -= L[i * n + j] * b[j * ldb + c];
            b[i * ldb + c] = v / L[i * n + i];
        }
        for (int i = n - 1; i >= 0; i--) {
            double v = b[i * ldb + c];
            for (int j = i + 1; j < n; j++) v -= L[j * n + i] * b[j * ldb + c];
            b[i * ldb + c] = v / L[i * n + i];
        }
    }
}

/* ddp.cpp:440-644 backwardpass. */
static void backwardpass(ctx_t *C) {
    const int N = C->N;
    double dV0 = 0.0, dV1 = 0.0, c_err = 0.0, mu_err = 0.0, Qu_err = 0.0;
    C->n_bwd_sweeps++;
    /* regularisation schedule, ddp.cpp:452-474 (hazard H4) */
    if (C->failed || C->bfailed) C->reg = C->reg + 1.0;
    else if (C->step == 0) C->reg = C->reg - 1.0;
    else if (C->step <= 3) C->reg = C->reg;
    else C->reg = C->reg + 1.0;
    if (C->reg < 0.0) C->reg = 0.0;
    else if (C->reg > 24.0) C->reg = 24.0;

    if (!C->failed) computeall(C); /* hazard H5 */

    double Vx[NX], Vxx[NX * NX];
    for (int i = 0; i < NX; i++) Vx[i] = C->px[i];
    memset(Vxx, 0, sizeof Vxx);
    for (int i = 0; i < NX; i++) Vxx[i * NX + i] = C->pb->w_terminal;
    const double regadd = pow(C->reg_base, C->reg) - 1;

    for (int i = N - 1; i >= 0; i--) {
        C->n_bwd_knots++;
        const int mc = C->mc[i], o = C->off[i];
        const double *fx = C->fx + (size_t)i * 81, *fu = C->fu + (size_t)i * 90;
        const double *qu = C->qu + (size_t)i * NU, *quu = C->quu + (size_t)i * 100;
        const double *cx = C->cx + (size_t)o * NX, *cu = C->cu + (size_t)o * NU;
        const double *c = C->c + o, *s = C->s + o, *y = C->y + o;
        double Qx[NX], Qu[NU], Qxx[NX * NX], Qxu[NX * NU], Quu[NU * NU], fxV[NX * NX], W[NX * NU];
        /* ddp.cpp:508-509 */
        for (int a = 0; a < NX; a++) {
            double t1 = 0.0, t2 = 0.0;
            for (int r = 0; r < mc; r++) t1 += cx[r * NX + a] * s[r];
            for (int b = 0; b < NX; b++) t2 += fx[b * 9 + a] * Vx[b];
            Qx[a] = (0.0 + t1) + t2;
        }
        for (int a = 0; a < NU; a++) {
            double t1 = 0.0, t2 = 0.0;
            for (int r = 0; r < mc; r++) t1 += cu[r * NU + a] * s[r];
            for (int b = 0; b < NX; b++) t2 += fu[b * 10 + a] * Vx[b];
            Qu[a] = (qu[a] + t1) + t2;
        }
        /* ddp.cpp:517-521 */
        for (int a = 0; a < NX; a++)
            for (int b = 0; b < NX; b++) {
                double t = 0.0;
                for (int k = 0; k < NX; k++) t += fx[k * 9 + a] * Vxx[k * NX + b];
                fxV[a * NX + b] = t;
            }
        for (int a = 0; a < NX; a++)
            for (int b = 0; b < NX; b++) {
                double t = 0.0;
                for (int k = 0; k < NX; k++) t += fxV[a * NX + k] * fx[k * 9 + b];
                Qxx[a * NX + b] = t;
            }
        for (int a = 0; a < NX; a++)
            for (int b = 0; b < NU; b++) {
                double t = 0.0;
                for (int k = 0; k < NX; k++) t += fxV[a * NX + k] * fu[k * 10 + b];
                Qxu[a * NU + b] = t;
            }
        for (int a = 0; a < NU; a++) /* W = fu^T Vxx  (10x9) */
            for (int b = 0; b < NX; b++) {
                double t = 0.0;
                for (int k = 0; k < NX; k++) t += fu[k * 10 + a] * Vxx[k * NX + b];
                W[a * NX + b] = t;
            }
        double Quu0[NU * NU];
        for (int a = 0; a < NU; a++)
            for (int b = 0; b < NU; b++) {
                double t = 0.0;
                for (int k = 0; k < NX; k++) t += W[a * NX + k] * fu[k * 10 + b];
                Quu0[a * NU + b] = quu[a * NU + b] + t;
            }
        for (int a = 0; a < NU; a++)
            for (int b = 0; b < NU; b++) Quu[a * NU + b] = 0.5 * (Quu0[a * NU + b] + Quu0[b * NU + a]);

        /* diagonal scalings: infeasible ddp.cpp:535-541, feasible :583-590 */
        double r[512], d[512], tv2[512];
        const double sgn = C->infeas ? 1.0 : -1.0;
        for (int k = 0; k < mc; k++) {
            if (C->infeas) {
                r[k] = s[k] * y[k] - C->mu;
                double rhat = s[k] * (c[k] + y[k]) - r[k];
                double yinv = 1.0 / y[k];
                d[k] = s[k] * yinv;
                tv2[k] = yinv * rhat;
            } else {
                r[k] = s[k] * c[k] + C->mu;
                double cinv = 1.0 / c[k];
                d[k] = s[k] * cinv;
                tv2[k] = cinv * r[k];
            }
        }
        double cDc[NU * NU], M[NU * NU];
        for (int a = 0; a < NU; a++)
            for (int b = 0; b < NU; b++) {
                double t = 0.0;
                for (int k = 0; k < mc; k++) t += cu[k * NU + a] * (d[k] * cu[k * NU + b]);
                cDc[a * NU + b] = t;
            }
        for (int a = 0; a < NU; a++)
            for (int b = 0; b < NU; b++)
                M[a * NU + b] = (Quu[a * NU + b] + (a == b ? regadd : 0.0)) + sgn * cDc[a * NU + b];
        if (llt10(M, NU)) { /* ddp.cpp:546-551 / :595-600 */
            C->bfailed = 1;
            C->opterr = INFINITY;
            return;
        }
        /* Qu, Qux (ddp.cpp:554-559 / :601-605) */
        for (int a = 0; a < NU; a++) {
            double t = 0.0;
            for (int k = 0; k < mc; k++) t += cu[k * NU + a] * tv2[k];
            Qu[a] += sgn * t;
        }
        double Qux[NU * NX]; /* tempQux */
        for (int a = 0; a < NU; a++)
            for (int b = 0; b < NX; b++) {
                double t = 0.0;
                for (int k = 0; k < mc; k++) t += cu[k * NU + a] * (d[k] * cx[k * NX + b]);
                Qux[a * NX + b] = Qxu[b * NU + a] + sgn * t;
            }
        double kK[NU * (NX + 1)];
        for (int a = 0; a < NU; a++) {
            kK[a * (NX + 1)] = Qu[a];
            for (int b = 0; b < NX; b++) kK[a * (NX + 1) + 1 + b] = Qux[a * NX + b];
        }
        llt_solve(M, NU, kK, NX + 1, NX + 1);
        double *ku = C->ku + (size_t)i * NU, *Ku = C->Ku + (size_t)i * 90;
        for (int a = 0; a < NU; a++) {
            ku[a] = -kK[a * (NX + 1)];
            for (int b = 0; b < NX; b++) Ku[a * NX + b] = -kK[a * (NX + 1) + 1 + b];
        }
        /* ks, Ks, ky, Ky (ddp.cpp:565-572 / :610-614) */
        double *ks = C->ks + o, *ky = C->ky + o, *Ks = C->Ks + (size_t)o * NX, *Ky = C->Ky + (size_t)o * NX;
        for (int k = 0; k < mc; k++) {
            double cuku = 0.0;
            for (int a = 0; a < NU; a++) cuku += cu[k * NU + a] * ku[a];
            for (int b = 0; b < NX; b++) {
                double t = 0.0;
                for (int a = 0; a < NU; a++) t += cu[k * NU + a] * Ku[a * NX + b];
                double cxk = cx[k * NX + b] + t;
                if (C->infeas) { Ks[k * NX + b] = d[k] * cxk; Ky[k * NX + b] = -cxk; }
                else { Ks[k * NX + b] = -(d[k] * cxk); Ky[k * NX + b] = 0.0; }
            }
            if (C->infeas) {
                double rhat = s[k] * (c[k] + y[k]) - r[k];
                ks[k] = (1.0 / y[k]) * (rhat + s[k] * cuku);
                ky[k] = -(c[k] + y[k]) - cuku;
            } else {
                ks[k] = -((1.0 / c[k]) * (r[k] + s[k] * cuku));
                ky[k] = 0.0;
            }
        }
        /* condensed (unregularised) blocks, ddp.cpp:574-578 / :615-618 (hazard H6) */
        for (int a = 0; a < NU * NU; a++) Quu[a] = Quu[a] + sgn * cDc[a];
        for (int a = 0; a < NX; a++)
            for (int b = 0; b < NU; b++) Qxu[a * NU + b] = Qux[b * NX + a];
        for (int a = 0; a < NX; a++)
            for (int b = 0; b < NX; b++) {
                double t = 0.0;
                for (int k = 0; k < mc; k++) t += cx[k * NX + a] * (d[k] * cx[k * NX + b]);
                Qxx[a * NX + b] += sgn * t;
            }
        for (int a = 0; a < NX; a++) {
            double t = 0.0;
            for (int k = 0; k < mc; k++) t += cx[k * NX + a] * tv2[k];
            Qx[a] += sgn * t;
        }
        /* value backup, ddp.cpp:620-628 */
        double kQu = 0.0;
        for (int a = 0; a < NU; a++) kQu += ku[a] * Qu[a];
        dV0 = dV0 + kQu;
        double QxuKu[NX * NX], KutQuu[NX * NU];
        for (int a = 0; a < NX; a++)
            for (int b = 0; b < NX; b++) {
                double t = 0.0;
                for (int k = 0; k < NU; k++) t += Qxu[a * NU + k] * Ku[k * NX + b];
                QxuKu[a * NX + b] = t;
            }
        for (int a = 0; a < NX; a++)
            for (int b = 0; b < NU; b++) {
                double t = 0.0;
                for (int k = 0; k < NU; k++) t += Ku[k * NX + a] * Quu[k * NU + b];
                KutQuu[a * NU + b] = t;
            }
        double kQk = 0.0;
        for (int a = 0; a < NU; a++) {
            double t = 0.0;
            for (int b = 0; b < NU; b++) t += Quu[a * NU + b] * ku[b];
            kQk += (0.5 * ku[a]) * t;
        }
        dV1 = dV1 + kQk;
        double Vxn[NX], Vxxn[NX * NX];
        for (int a = 0; a < NX; a++) {
            double t1 = 0.0, t2 = 0.0, t3 = 0.0;
            for (int k = 0; k < NU; k++) {
                t1 += Ku[k * NX + a] * Qu[k];
                t2 += KutQuu[a * NU + k] * ku[k];
                t3 += Qxu[a * NU + k] * ku[k];
            }
            Vxn[a] = ((Qx[a] + t1) + t2) + t3;
        }
        for (int a = 0; a < NX; a++)
            for (int b = 0; b < NX; b++) {
                double t = 0.0;
                for (int k = 0; k < NU; k++) t += KutQuu[a * NU + k] * Ku[k * NX + b];
                Vxxn[a * NX + b] = ((Qxx[a * NX + b] + QxuKu[b * NX + a]) + QxuKu[a * NX + b]) + t;
            }
        for (int a = 0; a < NX; a++) {
            Vx[a] = Vxn[a];
            for (int b = 0; b < NX; b++) Vxx[a * NX + b] = 0.5 * (Vxxn[a * NX + b] + Vxxn[b * NX + a]);
        }
        /* error norms, ddp.cpp:633-637 */
        for (int a = 0; a < NU; a++) Qu_err = fmax(Qu_err, fabs(Qu[a]));
        for (int k = 0; k < mc; k++) mu_err = fmax(mu_err, fabs(r[k]));
        if (C->infeas)
            for (int k = 0; k < mc; k++) c_err = fmax(c_err, fabs(c[k] + y[k]));
    }
    C->bfailed = 0;
    C->opterr = fmax(fmax(Qu_err, c_err), mu_err);
    C->dV[0] = dV0;
    C->dV[1] = dV1;
}

/* ddp.cpp:647-778 forwardpass. */
static void forwardpass(ctx_t *C) {
    const int N = C->N;
    memcpy(C->xn, C->x, (size_t)(N + 1) * NX * sizeof(double));
    memcpy(C->un, C->u, (size_t)N * NU * sizeof(double));
    memcpy(C->cn, C->c, (size_t)C->rows * sizeof(double));
    memcpy(C->yn, C->y, (size_t)C->rows * sizeof(double));
    memcpy(C->sn, C->s, (size_t)C->rows * sizeof(double));
    double cost = 0, costq = 0, logcost = 0, stepsize = 0, err = 0;
    const double tau = fmax(0.99, 1 - C->mu);
    int step, failed = 0;
    for (step = 0; step < 11; step++) {
        failed = 0;
        stepsize = pow(2.0, (double)(-step)); /* steplist = 2^{0,-1,...,-10}, ddp.cpp:670 */
        C->n_fwd_trials++;
        for (int a = 0; a < NX; a++) C->xn[a] = C->x[a];
        for (int i = 0; i < N; i++) {
            C->n_fwd_knots++;
            const int mc = C->mc[i], o = C->off[i];
            double dx[NX];
            for (int a = 0; a < NX; a++) dx[a] = C->xn[(size_t)i * NX + a] - C->x[(size_t)i * NX + a];
            const double *ku = C->ku + (size_t)i * NU, *Ku = C->Ku + (size_t)i * 90;
            double *un = C->un + (size_t)i * NU;
            if (C->infeas) {
                int bad = 0;
                for (int k = 0; k < mc; k++) {
                    double ty = 0.0, ts = 0.0;
                    for (int b = 0; b < NX; b++) {
                        ty += C->Ky[(size_t)(o + k) * NX + b] * dx[b];
                        ts += C->Ks[(size_t)(o + k) * NX + b] * dx[b];
                    }
                    C->yn[o + k] = (C->y[o + k] + stepsize * C->ky[o + k]) + ty;
                    C->sn[o + k] = (C->s[o + k] + stepsize * C->ks[o + k]) + ts;
                }
                for (int k = 0; k < mc; k++)
                    if (C->yn[o + k] < (1 - tau) * C->y[o + k] || C->sn[o + k] < (1 - tau) * C->s[o + k]) bad = 1;
                if (bad) { failed = 1; break; }
                for (int a = 0; a < NU; a++) {
                    double t = 0.0;
                    for (int b = 0; b < NX; b++) t += Ku[a * NX + b] * dx[b];
                    un[a] = (C->u[(size_t)i * NU + a] + stepsize * ku[a]) + t;
                }
                computenextx(C->xn + (size_t)i * NX, un, C->xn + (size_t)(i + 1) * NX);
            } else {
                int bad = 0;
                for (int k = 0; k < mc; k++) {
                    double ts = 0.0;
                    for (int b = 0; b < NX; b++) ts += C->Ks[(size_t)(o + k) * NX + b] * dx[b];
                    C->sn[o + k] = (C->s[o + k] + stepsize * C->ks[o + k]) + ts;
                }
                for (int a = 0; a < NU; a++) {
                    double t = 0.0;
                    for (int b = 0; b < NX; b++) t += Ku[a * NX + b] * dx[b];
                    un[a] = (C->u[(size_t)i * NU + a] + stepsize * ku[a]) + t;
                }
                computecminvo(C, C->xn + (size_t)i * NX, un, i, C->cn + o);
                for (int k = 0; k < mc; k++)
                    if (C->cn[o + k] > (1 - tau) * C->c[o + k] || C->sn[o + k] < (1 - tau) * C->s[o + k]) bad = 1;
                if (bad) { failed = 1; break; }
                computenextx(C->xn + (size_t)i * NX, un, C->xn + (size_t)(i + 1) * NX);
            }
        }
        if (failed) continue;
        double qs = 0.0;
        for (int i = 0; i < N; i++) {
            C->qn[i] = computeq(C, C->un + (size_t)i * NU);
            qs += C->qn[i];
        }
        cost = qs + computep(C, C->xn + (size_t)N * NX);
        costq = qs;
        logcost = cost;
        err = 0.0;
        if (C->infeas) {
            for (int i = 0; i < N; i++) {
                const int mc = C->mc[i], o = C->off[i];
                double ls = 0.0, e1 = 0.0;
                for (int k = 0; k < mc; k++) ls += log(C->yn[o + k]);
                logcost -= C->mu * ls;
                computecminvo(C, C->xn + (size_t)i * NX, C->un + (size_t)i * NU, i, C->cn + o);
                for (int k = 0; k < mc; k++) e1 += fabs(C->cn[o + k] + C->yn[o + k]);
                err += e1;
            }
            err = fmax(C->tol, err);
        } else {
            for (int i = 0; i < N; i++) {
                const int mc = C->mc[i], o = C->off[i];
                double ls = 0.0;
                computecminvo(C, C->xn + (size_t)i * NX, C->un + (size_t)i * NU, i, C->cn + o);
                for (int k = 0; k < mc; k++) ls += log(-C->cn[o + k]);
                logcost -= C->mu * ls;
            }
            err = 0.0;
        }
        /* filter test, ddp.cpp:737-757 (hazard H7) */
        int nkeep = 0;
        if (C->nfilter + 1 > C->capfilter) {
            C->capfilter = 2 * C->capfilter + 2;
            C->filter = (double *)realloc(C->filter, (size_t)C->capfilter * 2 * sizeof(double));
        }
        int *keep = (int *)malloc((size_t)(C->nfilter + 1) * sizeof(int));
        for (int i = 0; i < C->nfilter; i++) {
            double f0 = C->filter[2 * i], f1 = C->filter[2 * i + 1];
            if (logcost >= f0 && err >= f1) { failed = 1; break; }
            else if (logcost > f0 || err > f1) keep[nkeep++] = i;
        }
        if (failed) { free(keep); continue; }
        for (int i = 0; i < nkeep; i++) {
            C->filter[2 * i] = C->filter[2 * keep[i]];
            C->filter[2 * i + 1] = C->filter[2 * keep[i] + 1];
        }
        free(keep);
        C->filter[2 * nkeep] = logcost;
        C->filter[2 * nkeep + 1] = err;
        C->nfilter = nkeep + 1;
        break;
    }
    if (failed) {
        C->failed = 1;
        C->stepsize = 0.0;
    } else {
        C->cost = cost;
        C->costq = costq;
        C->logcost = logcost;
        memcpy(C->x, C->xn, (size_t)(N + 1) * NX * sizeof(double));
        memcpy(C->u, C->un, (size_t)N * NU * sizeof(double));
        memcpy(C->y, C->yn, (size_t)C->rows * sizeof(double));
        memcpy(C->s, C->sn, (size_t)C->rows * sizeof(double));
        memcpy(C->c, C->cn, (size_t)C->rows * sizeof(double));
        memcpy(C->q, C->qn, (size_t)N * sizeof(double));
        C->err = err;
        C->stepsize = stepsize;
        C->step = step;
        C->failed = 0;
    }
}

/* 6x6 inverse by LU with partial pivoting (Eigen's MatrixXd::inverse(), ddp.cpp:804). */
static void inv6(const double A[6][6], double Ai[6][6]) {
    double M[6][12];
    for (int i = 0; i < 6; i++)
        for (int j = 0; j < 6; j++) { M[i][j] = A[i][j]; M[i][6 + j] = (i == j); }
    for (int k = 0; k < 6; k++) {
        int p = k;
        for (int i = k + 1; i < 6; i++) if (fabs(M[i][k]) > fabs(M[p][k])) p = i;
        if (p != k) for (int j = 0; j < 12; j++) { double t = M[k][j]; M[k][j] = M[p][j]; M[p][j] = t; }
        for (int i = k + 1; i < 6; i++) {
            double f = M[i][k] / M[k][k];
            for (int j = k; j < 12; j++) M[i][j] -= f * M[k][j];
        }
    }
    for (int k = 5; k >= 0; k--) {
        for (int j = 6; j < 12; j++) {
            double v = M[k][j];
            for (int i = k + 1; i < 6; i++) v -= M[k][i] * M[i][j];
            M[k][j] = v / M[k][k];
        }
    }
    for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) Ai[i][j] = M[i][6 + j];
}

/* (poly2bez * t2tauMat)(r,c) = BERN[c][r] * (1/T)^c, ddp.cpp:1018-1060 + :788. */
static void bez2poly_matrix(double T, double B2P[6][6]) {
    double Tk = 1.0 / T, pw[6];
    pw[0] = 1.0; pw[1] = Tk;
    for (int i = 2; i < 6; i++) pw[i] = pw[i - 1] * Tk;
    for (int r = 0; r < 6; r++) for (int c = 0; c < 6; c++) B2P[r][c] = BERN[c][r] * pw[c];
}

int ipddp_oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

static void *xcalloc(size_t n, size_t sz) { return calloc(n ? n : 1, sz); }

int ipddp_oracle_solve(const ipddp_problem *pb, ipddp_result *res) {
    if (!pb || !res || pb->N <= 0) return -1;
    if (pb->time_power != 1 && pb->time_power != 2) return -2; /* hazard H11: UB in the reference */
    const int N = pb->N;
    ctx_t Cs; ctx_t *C = &Cs;
    memset(C, 0, sizeof *C);
    C->pb = pb; C->N = N;
    C->mc = (int *)xcalloc(N, sizeof(int));
    C->off = (int *)xcalloc(N + 1, sizeof(int));
    for (int i = 0; i < N; i++) {
        if (pb->nplanes[i] < 0 || pb->nplanes[i] > pb->P_max || KC * pb->nplanes[i] + 55 > 512) return -3;
        C->mc[i] = KC * pb->nplanes[i] + 55; /* ddp.cpp:145 */
        C->off[i + 1] = C->off[i] + C->mc[i];
    }
    C->rows = C->off[N];
    C->tp = pb->minvo ? MINVO_P : BEZ_P;
    C->tv = pb->minvo ? MINVO_V : BEZ_V;
    C->ta = pb->minvo ? MINVO_A : BEZ_A;
    C->maxiter = pb->iter_max; C->tol = 1.0e-7; C->infeas = pb->infeas; /* ddp.cpp:42-44 */
    C->reg_base = pb->zero_init ? 1.6 : 4.0;                            /* ddp.cpp:60-61 */
    const size_t R = (size_t)C->rows;
    C->x = (double *)xcalloc((size_t)(N + 1) * NX, 8); C->u = (double *)xcalloc((size_t)N * NU, 8);
    C->c = (double *)xcalloc(R, 8); C->s = (double *)xcalloc(R, 8); C->y = (double *)xcalloc(R, 8);
    C->q = (double *)xcalloc(N, 8);
    C->fx = (double *)xcalloc((size_t)N * 81, 8); C->fu = (double *)xcalloc((size_t)N * 90, 8);
    C->qu = (double *)xcalloc((size_t)N * NU, 8); C->quu = (double *)xcalloc((size_t)N * 100, 8);
    C->cx = (double *)xcalloc(R * NX, 8); C->cu = (double *)xcalloc(R * NU, 8);
    C->ku = (double *)xcalloc((size_t)N * NU, 8); C->Ku = (double *)xcalloc((size_t)N * 90, 8);
    C->ks = (double *)xcalloc(R, 8); C->ky = (double *)xcalloc(R, 8);
    C->Ks = (double *)xcalloc(R * NX, 8); C->Ky = (double *)xcalloc(R * NX, 8);
    C->xn = (double *)xcalloc((size_t)(N + 1) * NX, 8); C->un = (double *)xcalloc((size_t)N * NU, 8);
    C->cn = (double *)xcalloc(R, 8); C->sn = (double *)xcalloc(R, 8); C->yn = (double *)xcalloc(R, 8);
    C->qn = (double *)xcalloc(N, 8);
    C->capfilter = 16; C->filter = (double *)xcalloc((size_t)C->capfilter * 2, 8);

    int rtn = 0;
    int infeas_out = pb->infeas, line_failed_out = 1;
    /* setup, ddp.cpp:104-160 */
    for (int a = 0; a < NX; a++) C->x[a] = pb->x0[a];
    for (int i = 0; i < N; i++) {
        C->u[(size_t)i * NU + 9] = pb->durations[i];
        for (int r = 0; r < C->mc[i]; r++) { C->s[C->off[i] + r] = 1.0e-1; C->y[C->off[i] + r] = 0.01; }
    }
    /* warm start, ddp.cpp:163-193: Bezier [x*6,y*6,z*6] -> [xyz]*6 -> monomial in t */
    if (!pb->zero_init) {
        if (!pb->line_init) {
            for (int i = 0; i < N; i++) {
                double T = pb->durations[i], B2P[6][6], tb[NZ];
                bez2poly_matrix(T, B2P);
                for (int j = 0; j < 6; j++)
                    for (int a = 0; a < 3; a++) tb[j * 3 + a] = T * (pb->init_bez ? pb->init_bez[(size_t)i * NZ + a * 6 + j] : 0.0);
                for (int cidx = 3; cidx < 6; cidx++)
                    for (int a = 0; a < 3; a++) {
                        double acc = 0.0;
                        for (int j = 0; j < 6; j++) acc += tb[j * 3 + a] * B2P[j][cidx];
                        C->u[(size_t)i * NU + (cidx - 3) * 3 + a] = acc;
                    }
            }
        } else { /* straight-line initialisation, ddp.cpp:195-247 */
            for (int l = 0; l < N; l++) {
                double p0[3], p1[3];
                for (int a = 0; a < 3; a++) {
                    p0[a] = (l == 0) ? pb->x0[a] : pb->seeds[(size_t)l * 3 + a];
                    p1[a] = (l == N - 1) ? pb->xd[a] : pb->seeds[(size_t)(l + 1) * 3 + a];
                }
                int vio = 1, cnt = 0;
                double *u = C->u + (size_t)l * NU;
                double *ctmp = (double *)xcalloc(C->mc[l], 8);
                while (vio && cnt <= 4) {
                    double Tk = u[9], Tk2 = Tk * Tk, Tk3 = Tk2 * Tk, Tk4 = Tk3 * Tk, Tk5 = Tk4 * Tk;
                    double Gi[3][3] = {{10.0 / Tk3, -4.0 / Tk2, 0.5 / Tk},
                                       {-15.0 / Tk4, 7.0 / Tk3, -1.0 / Tk2},
                                       {6.0 / Tk5, -3.0 / Tk4, 0.5 / Tk3}};
                    double F[3][3] = {{1.0, Tk, Tk2 / 2.0}, {0.0, 1.0, Tk}, {0.0, 0.0, 1.0}};
                    double xcur[NX] = {0}, xnext[NX] = {0}, rhs[NX];
                    for (int a = 0; a < 3; a++) { xcur[a] = p0[a]; xnext[a] = p1[a]; }
                    for (int i = 0; i < 3; i++)
                        for (int k = 0; k < 3; k++) {
                            double t = 0.0;
                            for (int j = 0; j < 3; j++) t += F[i][j] * xcur[j * 3 + k];
                            rhs[i * 3 + k] = xnext[i * 3 + k] - t;
                        }
                    for (int i = 0; i < 3; i++)
                        for (int k = 0; k < 3; k++) {
                            double t = 0.0;
                            for (int j = 0; j < 3; j++) t += Gi[i][j] * rhs[j * 3 + k];
                            u[i * 3 + k] = t;
                        }
                    computecminvo(C, xcur, u, l, ctmp);
                    int all_neg = 1;
                    for (int r = 0; r < C->mc[l]; r++) if (!(ctmp[r] < 0)) all_neg = 0;
                    if (all_neg) vio = 0;
                    else { u[9] = 2 * Tk; cnt++; }
                }
                free(ctmp);
            }
        }
    }
    initialroll(C); /* ddp.cpp:252 */
    if (pb->line_init) { /* ddp.cpp:255-269 */
        int count = 0;
        for (size_t r = 0; r < R; r++) if (C->c[r] > 0) count++;
        if (count == 0) C->infeas = 0;
    }
    C->mu = C->cost / N / C->mc[0]; /* hazard H3, ddp.cpp:281 */
    resetfilter(C);
    C->reg = 0.0; C->bfailed = 0; /* resetreg */
    if (pb->line_init) C->reg = 10.0;

    double cost_prev = C->cost, costq_prev = C->costq; /* costTraj.end()[-2] after each push */
    int iter = 0, bp_no_upd_count = 0, no_upd_count = 0, opt_no_upd_count = 0;
    const int bp_no_upd_count_max = 20;
    res->trace_len = 0;
    for (iter = 0; iter < C->maxiter; iter++) {
        int n_bwd = 0;
        while (1) { /* ddp.cpp:297-310 */
            backwardpass(C);
            n_bwd++;
            if (!C->bfailed) break;
            if (C->reg == 24 && C->bfailed) bp_no_upd_count++;
            else bp_no_upd_count = 0;
            if (bp_no_upd_count > bp_no_upd_count_max) break;
        }
        forwardpass(C);
        if (res->trace && res->trace_len < res->trace_cap) {
            ipddp_iter_trace *t = &res->trace[res->trace_len++];
            t->cost = C->cost; t->costq = C->costq; t->logcost = C->logcost; t->err = C->err;
            t->mu = C->mu; t->reg = C->reg; t->stepsize = C->stepsize; t->opterr = C->opterr;
            t->step = C->step; t->fp_failed = C->failed; t->n_bwd = n_bwd;
        }
        int neg = 0; /* ddp.cpp:317-326 */
        for (int i = 0; i < N; i++) if (C->u[(size_t)i * NU + 9] < 0) neg = 1;
        if (neg) { rtn = -3; break; }
        double cost_m2 = cost_prev, costq_m2 = costq_prev;
        cost_prev = C->cost; costq_prev = C->costq;

        if (fmax(C->opterr, C->mu) <= C->tol) break; /* ddp.cpp:335-338 */
        if (C->opterr <= 0.2 * C->mu) {             /* ddp.cpp:340-344 */
            C->mu = fmax(C->tol / 10.0, fmin(0.2 * C->mu, pow(C->mu, 1.2)));
            resetfilter(C);
            C->reg = 0.0; C->bfailed = 0;
        }
        int count = 0; /* ddp.cpp:346-355 (hazard H8) */
        for (size_t r = 0; r < R; r++) if (C->c[r] >= 2.0e-4) count++;
        if (count == 0) {
            if (pb->zero_init) { infeas_out = 0; rtn = 2; break; }
            if (!pb->zero_init && !pb->line_init) {
                if (pow(C->costq - costq_m2, 2) < costq_m2 * 1.0e-2) opt_no_upd_count++;
                else opt_no_upd_count = 0;
                if ((pow(C->cost - cost_m2, 2) < cost_m2 * 1.0e-2) && C->opterr < 5.0e1) { rtn = 1; break; }
            }
            if (pb->line_init) {
                if (pow(C->cost - cost_m2, 2) < cost_m2 * 0.01) { line_failed_out = 0; break; }
            }
        }
        if (bp_no_upd_count > bp_no_upd_count_max) { rtn = -4; break; }
        if (pb->line_init) {
            if (C->stepsize < 1.0e-6) no_upd_count++;
            else no_upd_count = 0;
            if (no_upd_count > 100) break;
        }
    }
    (void)opt_no_upd_count;
    /* outputs, ddp.cpp:418-437 */
    res->rtn = rtn; res->infeas_out = infeas_out; res->line_failed_out = line_failed_out;
    res->iters = iter; res->cost = C->cost; res->costq = C->costq;
    res->mu_final = C->mu; res->opterr_final = C->opterr;
    for (int a = 0; a < NX; a++) res->x_final[a] = C->x[(size_t)N * NX + a];
    for (int i = 0; i < N; i++) {
        const double *x = C->x + (size_t)i * NX, *u = C->u + (size_t)i * NU;
        double R3[3][3], Rp[3][3], Rpp[3][3];
        R_of_T(u[9], R3, Rp, Rpp);
        if (res->jerk) res->jerk[i] = quad3(R3, u); /* finalroll, ddp.cpp:1624-1634 */
        double pc[NZ];
        for (int k = 0; k < 3; k++) for (int a = 0; a < 3; a++) pc[k * 3 + a] = EK_INV[k] * x[k * 3 + a];
        for (int r = 0; r < 9; r++) pc[9 + r] = u[r];
        if (res->poly_coeff) memcpy(res->poly_coeff + (size_t)i * NZ, pc, sizeof pc);
        if (res->poly_time) res->poly_time[i] = u[9];
        if (res->bez_coeff) { /* poly2bezFunc, ddp.cpp:799-812, then re-layout :430-436 */
            double B2P[6][6], P2B[6][6];
            bez2poly_matrix(u[9], B2P);
            inv6(B2P, P2B);
            double sc = 1.0 / u[9];
            for (int j = 0; j < 6; j++)
                for (int a = 0; a < 3; a++) {
                    double acc = 0.0;
                    for (int k = 0; k < 6; k++) acc += (sc * pc[k * 3 + a]) * P2B[k][j];
                    res->bez_coeff[(size_t)i * NZ + a * 6 + j] = acc;
                }
        }
    }
    res->n_bwd_sweeps = C->n_bwd_sweeps; res->n_bwd_knots = C->n_bwd_knots;
    res->n_fwd_trials = C->n_fwd_trials; res->n_fwd_knots = C->n_fwd_knots;

    free(C->mc); free(C->off); free(C->x); free(C->u); free(C->c); free(C->s); free(C->y); free(C->q);
    free(C->fx); free(C->fu); free(C->qu); free(C->quu); free(C->cx); free(C->cu);
    free(C->ku); free(C->Ku); free(C->ks); free(C->ky); free(C->Ks); free(C->Ky);
    free(C->xn); free(C->un); free(C->cn); free(C->sn); free(C->yn); free(C->qn); free(C->filter);
    return 0;
}

/* trp.cpp:583-639 initTimeAllocation (v0 = 0 so V0 = aV0 = 0). */
void ipddp_oracle_time_allocation(int N, const double *start, const double *end, const double *seeds,
                                  double max_vel, double max_acc, double *durations) {
    const double _Vel = max_vel, _Acc = max_acc;
    for (int k = 0; k < N; k++) {
        double p0[3], p1[3];
        for (int a = 0; a < 3; a++) {
            p0[a] = (k == 0) ? start[a] : seeds[(size_t)k * 3 + a];
            p1[a] = (k == N - 1) ? end[a] : seeds[(size_t)(k + 1) * 3 + a];
        }
        double d[3] = {p1[0] - p0[0], p1[1] - p0[1], p1[2] - p0[2]};
        double D = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
        double V0 = 0.0 * (d[0] / D) + 0.0 * (d[1] / D) + 0.0 * (d[2] / D);
        double aV0 = fabs(V0);
        double acct = (_Vel - V0) / _Acc * ((_Vel > V0) ? 1 : -1);
        double accd = V0 * acct + (_Acc * acct * acct / 2) * ((_Vel > V0) ? 1 : -1);
        double dcct = _Vel / _Acc;
        double dccd = _Acc * dcct * dcct / 2;
        double dtxyz;
        if (D < aV0 * aV0 / (2 * _Acc)) {
            double t1 = (V0 < 0) ? 2.0 * aV0 / _Acc : 0.0;
            double t2 = aV0 / _Acc;
            dtxyz = t1 + t2;
        } else if (D < accd + dccd) {
            double t1 = (V0 < 0) ? 2.0 * aV0 / _Acc : 0.0;
            double t2 = (-aV0 + sqrt(aV0 * aV0 + _Acc * D - aV0 * aV0 / 2)) / _Acc;
            double t3 = (aV0 + _Acc * t2) / _Acc;
            dtxyz = t1 + t2 + t3;
        } else {
            double t1 = acct;
            double t2 = (D - accd - dccd) / _Vel;
            double t3 = dcct;
            dtxyz = t1 + t2 + t3;
        }
        durations[k] = dtxyz;
    }
}

/* trp.cpp:853-951 (fastTrajPlanning): stage 0 zero-init infeasible solve, durations updated when
 * it returns 2, stage 1 warm-started from the stage-0 Bezier with the bool& infeas carried over. */
int ipddp_oracle_two_stage(const ipddp_problem *prob, const ipddp_two_stage_opts *o, ipddp_result *res0,
                           ipddp_result *res1) {
    const int N = prob->N;
    ipddp_problem p0 = *prob;
    ipddp_result r0local; memset(&r0local, 0, sizeof r0local);
    ipddp_result *r0 = res0 ? res0 : &r0local;
    double *bez0 = (double *)xcalloc((size_t)N * NZ, 8), *time0 = (double *)xcalloc(N, 8);
    double *save_bez = r0->bez_coeff, *save_time = r0->poly_time;
    r0->bez_coeff = bez0; r0->poly_time = time0;
    p0.init_bez = NULL;
    p0.w_snap = o->w_snap0; p0.w_terminal = o->w_terminal0; p0.w_time = o->w_time0; p0.iter_max = o->iter_max0;
    p0.infeas = 1; p0.zero_init = 1; p0.line_init = 0; p0.minvo = 0; p0.time_power = o->time_power;
    int st = ipddp_oracle_solve(&p0, r0);
    if (st) { free(bez0); free(time0); return st; }
    ipddp_problem p1 = *prob;
    if (r0->rtn == 2) p1.durations = time0; /* UpdateTime, trp.cpp:911-912 */
    p1.init_bez = bez0;
    p1.w_snap = o->w_snap; p1.w_terminal = o->w_terminal; p1.w_time = o->w_time; p1.iter_max = o->iter_max;
    p1.infeas = r0->infeas_out; p1.zero_init = 0; p1.line_init = 0; p1.minvo = 0; p1.time_power = o->time_power;
    st = ipddp_oracle_solve(&p1, res1);
    if (save_bez) memcpy(save_bez, bez0, (size_t)N * NZ * 8);
    if (save_time) memcpy(save_time, time0, (size_t)N * 8);
    r0->bez_coeff = save_bez; r0->poly_time = save_time;
    free(bez0); free(time0);
    return st;
}
