/* gddp_impl.h -- body of the generic unconstrained DDP oracle, included once per arithmetic type (REAL, SUF).
 * See gddp_oracle.c for what this is (and is not). */
#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SUF)

/* ---- models: continuous dynamics f(x,u) and its Jacobians F = df/dx (nx x nx), G = df/du (nx x nu) ---------------- */
static void FN(dint_f)(const REAL *x, const REAL *u, REAL *f, REAL *F, REAL *G) {
    /* 3D double integrator: x = [p, v], u = a */
    for (int a = 0; a < 3; a++) { f[a] = x[3 + a]; f[3 + a] = u[a]; }
    if (F) {
        memset(F, 0, 36 * sizeof(REAL)); memset(G, 0, 18 * sizeof(REAL));
        for (int a = 0; a < 3; a++) { F[a * 6 + 3 + a] = 1; G[(3 + a) * 3 + a] = 1; }
    }
}
static void FN(quad_f)(const REAL *x, const REAL *u, REAL *f, REAL *F, REAL *G) {
    /* rigid-body quadrotor, x = [p, v, (phi, theta, psi), (p, q, r)], u = [thrust, tau_x, tau_y, tau_z];
     * constants of simulation/so3_quadrotor_simulator/src/dynamics/Quadrotor.cpp:15-20 */
    const REAL m = (REAL)0.98, g = (REAL)9.81, Jx = (REAL)2.64e-3, Jy = (REAL)2.64e-3, Jz = (REAL)4.96e-3;
    const REAL ph = x[6], th = x[7], ps = x[8], p = x[9], q = x[10], r = x[11];
    const REAL sp = SIN(ph), cp = COS(ph), st = SIN(th), ct = COS(th), ss = SIN(ps), cs = COS(ps);
    const REAL tt = st / ct, ict = (REAL)1 / ct;
    const REAL b3[3] = {cp * st * cs + sp * ss, cp * st * ss - sp * cs, cp * ct};
    const REAL fm = u[0] / m;
    for (int a = 0; a < 3; a++) f[a] = x[3 + a];
    f[3] = fm * b3[0]; f[4] = fm * b3[1]; f[5] = fm * b3[2] - g;
    const REAL sqcr = sp * q + cp * r, cqsr = cp * q - sp * r;
    f[6] = p + tt * sqcr; f[7] = cqsr; f[8] = sqcr * ict;
    f[9] = (u[1] - (Jz - Jy) * q * r) / Jx; f[10] = (u[2] - (Jx - Jz) * p * r) / Jy; f[11] = (u[3] - (Jy - Jx) * p * q) / Jz;
    if (!F) return;
    memset(F, 0, 144 * sizeof(REAL)); memset(G, 0, 48 * sizeof(REAL));
    for (int a = 0; a < 3; a++) F[a * 12 + 3 + a] = 1;
    const REAL db_dph[3] = {-sp * st * cs + cp * ss, -sp * st * ss - cp * cs, -sp * ct};
    const REAL db_dth[3] = {cp * ct * cs, cp * ct * ss, -cp * st};
    const REAL db_dps[3] = {-cp * st * ss + sp * cs, cp * st * cs + sp * ss, (REAL)0};
    for (int a = 0; a < 3; a++) {
        F[(3 + a) * 12 + 6] = fm * db_dph[a]; F[(3 + a) * 12 + 7] = fm * db_dth[a]; F[(3 + a) * 12 + 8] = fm * db_dps[a];
        G[(3 + a) * 4 + 0] = b3[a] / m;
    }
    F[6 * 12 + 6] = tt * cqsr;        F[6 * 12 + 7] = sqcr * ict * ict;
    F[7 * 12 + 6] = -sqcr;
    F[8 * 12 + 6] = cqsr * ict;       F[8 * 12 + 7] = sqcr * st * ict * ict;
    F[6 * 12 + 9] = 1;  F[6 * 12 + 10] = sp * tt;  F[6 * 12 + 11] = cp * tt;
    F[7 * 12 + 10] = cp; F[7 * 12 + 11] = -sp;
    F[8 * 12 + 10] = sp * ict; F[8 * 12 + 11] = cp * ict;
    F[9 * 12 + 10] = -(Jz - Jy) * r / Jx;  F[9 * 12 + 11] = -(Jz - Jy) * q / Jx;
    F[10 * 12 + 9] = -(Jx - Jz) * r / Jy;  F[10 * 12 + 11] = -(Jx - Jz) * p / Jy;
    F[11 * 12 + 9] = -(Jy - Jx) * q / Jz;  F[11 * 12 + 10] = -(Jy - Jx) * p / Jz;
    G[9 * 4 + 1] = (REAL)1 / Jx; G[10 * 4 + 2] = (REAL)1 / Jy; G[11 * 4 + 3] = (REAL)1 / Jz;
}
static void FN(model_f)(int model, const REAL *x, const REAL *u, REAL *f, REAL *F, REAL *G) {
    if (model == GDDP_MODEL_QUAD12) FN(quad_f)(x, u, f, F, G);
    else FN(dint_f)(x, u, f, F, G);
}

/* rollout of (ub + alpha k + K (x - xb)) from x0; returns the cost.  With K == NULL: open loop. */
static REAL FN(rollout)(const gddp_problem *P, int b, const REAL *xb, const REAL *ub, const REAL *K, const REAL *kf, REAL alpha,
                        REAL *xn, REAL *un) {
    const int nx = P->nx, nu = P->nu, N = P->N;
    const REAL dt = (REAL)P->dt;
    REAL J = 0, f[GDDP_MAX_NX];
    for (int a = 0; a < nx; a++) xn[a] = (REAL)P->x0[(size_t)b * nx + a];
    for (int i = 0; i < N; i++) {
        const REAL *x = xn + (size_t)i * nx;
        REAL *u = un + (size_t)i * nu;
        for (int m = 0; m < nu; m++) {
            REAL v = ub[(size_t)i * nu + m];
            if (K) {
                v += alpha * kf[(size_t)i * nu + m];
                for (int a = 0; a < nx; a++) v += K[((size_t)i * nu + m) * nx + a] * (x[a] - xb[(size_t)i * nx + a]);
            }
            u[m] = v;
        }
        REAL c = 0;
        for (int a = 0; a < nx; a++) { const REAL d = x[a] - (REAL)P->xg[(size_t)b * nx + a]; c += (REAL)P->q[a] * d * d; }
        for (int m = 0; m < nu; m++) { const REAL d = u[m] - (REAL)P->uh[m]; c += (REAL)P->r[m] * d * d; }
        J += (REAL)0.5 * dt * c;
        FN(model_f)(P->model, x, u, f, NULL, NULL);
        for (int a = 0; a < nx; a++) xn[(size_t)(i + 1) * nx + a] = x[a] + dt * f[a];
    }
    REAL c = 0;
    for (int a = 0; a < nx; a++) { const REAL d = xn[(size_t)N * nx + a] - (REAL)P->xg[(size_t)b * nx + a]; c += (REAL)P->qf[a] * d * d; }
    return J + (REAL)0.5 * c;
}

/* backward sweep with the regularised Quu; 0 = a pivot was not positive */
static int FN(sweep)(const gddp_problem *P, int b, const REAL *xb, const REAL *ub, REAL rho, REAL *K, REAL *kf, REAL *dV1) {
    const int nx = P->nx, nu = P->nu, N = P->N;
    const REAL dt = (REAL)P->dt;
    REAL Vx[GDDP_MAX_NX], Vxx[GDDP_MAX_NX * GDDP_MAX_NX], F[GDDP_MAX_NX * GDDP_MAX_NX], G[GDDP_MAX_NX * GDDP_MAX_NU], f[GDDP_MAX_NX];
    REAL A[GDDP_MAX_NX * GDDP_MAX_NX], Bm[GDDP_MAX_NX * GDDP_MAX_NU], VA[GDDP_MAX_NX * GDDP_MAX_NX], VB[GDDP_MAX_NX * GDDP_MAX_NU];
    REAL Qx[GDDP_MAX_NX], Qu[GDDP_MAX_NU], Qxx[GDDP_MAX_NX * GDDP_MAX_NX], Qux[GDDP_MAX_NU * GDDP_MAX_NX], Quu[GDDP_MAX_NU * GDDP_MAX_NU];
    REAL L[GDDP_MAX_NU * GDDP_MAX_NU];
    memset(Vxx, 0, sizeof Vxx);
    REAL dv = 0;   /* expected first-order change of the cost, sum_i k_i' Qu_i = -sum_i |L_i^-1 Qu_i|^2 */
    for (int a = 0; a < nx; a++) {
        Vx[a] = (REAL)P->qf[a] * (xb[(size_t)N * nx + a] - (REAL)P->xg[(size_t)b * nx + a]);
        Vxx[a * nx + a] = (REAL)P->qf[a];
    }
    for (int i = N - 1; i >= 0; i--) {
        const REAL *x = xb + (size_t)i * nx, *u = ub + (size_t)i * nu;
        FN(model_f)(P->model, x, u, f, F, G);
        for (int r = 0; r < nx; r++) {
            for (int c = 0; c < nx; c++) A[r * nx + c] = (r == c ? (REAL)1 : (REAL)0) + dt * F[r * nx + c];
            for (int c = 0; c < nu; c++) Bm[r * nu + c] = dt * G[r * nu + c];
        }
        for (int r = 0; r < nx; r++) {
            for (int c = 0; c < nx; c++) { REAL s = 0; for (int k = 0; k < nx; k++) s += Vxx[r * nx + k] * A[k * nx + c]; VA[r * nx + c] = s; }
            for (int c = 0; c < nu; c++) { REAL s = 0; for (int k = 0; k < nx; k++) s += Vxx[r * nx + k] * Bm[k * nu + c]; VB[r * nu + c] = s; }
        }
        for (int c = 0; c < nx; c++) {
            REAL s = dt * (REAL)P->q[c] * (x[c] - (REAL)P->xg[(size_t)b * nx + c]);
            for (int k = 0; k < nx; k++) s += A[k * nx + c] * Vx[k];
            Qx[c] = s;
        }
        for (int c = 0; c < nu; c++) {
            REAL s = dt * (REAL)P->r[c] * (u[c] - (REAL)P->uh[c]);
            for (int k = 0; k < nx; k++) s += Bm[k * nu + c] * Vx[k];
            Qu[c] = s;
        }
        for (int r = 0; r < nx; r++)
            for (int c = 0; c < nx; c++) {
                REAL s = (r == c) ? dt * (REAL)P->q[c] : (REAL)0;
                for (int k = 0; k < nx; k++) s += A[k * nx + r] * VA[k * nx + c];
                Qxx[r * nx + c] = s;
            }
        for (int r = 0; r < nu; r++) {
            for (int c = 0; c < nx; c++) { REAL s = 0; for (int k = 0; k < nx; k++) s += Bm[k * nu + r] * VA[k * nx + c]; Qux[r * nx + c] = s; }
            for (int c = 0; c < nu; c++) {
                REAL s = (r == c) ? dt * (REAL)P->r[c] + rho : (REAL)0;
                for (int k = 0; k < nx; k++) s += Bm[k * nu + r] * VB[k * nu + c];
                Quu[r * nu + c] = s;
            }
        }
        /* Cholesky Quu = L L^T */
        memset(L, 0, sizeof L);
        for (int c = 0; c < nu; c++) {
            REAL d = Quu[c * nu + c];
            for (int k = 0; k < c; k++) d -= L[c * nu + k] * L[c * nu + k];
            if (!(d > (REAL)0)) return 0;
            const REAL ld = SQRT(d);
            L[c * nu + c] = ld;
            for (int r = c + 1; r < nu; r++) {
                REAL s = Quu[r * nu + c];
                for (int k = 0; k < c; k++) s -= L[r * nu + k] * L[c * nu + k];
                L[r * nu + c] = s / ld;
            }
        }
        /* [k | K] = -Quu^-1 [Qu | Qux] */
        for (int c = -1; c < nx; c++) {
            REAL y[GDDP_MAX_NU], z[GDDP_MAX_NU];
            for (int r = 0; r < nu; r++) {
                REAL s = (c < 0) ? Qu[r] : Qux[r * nx + c];
                for (int k = 0; k < r; k++) s -= L[r * nu + k] * y[k];
                y[r] = s / L[r * nu + r];
                if (c < 0) dv -= y[r] * y[r];
            }
            for (int r = nu - 1; r >= 0; r--) {
                REAL s = y[r];
                for (int k = r + 1; k < nu; k++) s -= L[k * nu + r] * z[k];
                z[r] = s / L[r * nu + r];
            }
            for (int r = 0; r < nu; r++) {
                if (c < 0) kf[(size_t)i * nu + r] = -z[r];
                else K[((size_t)i * nu + r) * nx + c] = -z[r];
            }
        }
        /* Vx = Qx + Qux^T k ; Vxx = sym(Qxx + Qux^T K) */
        REAL Vn[GDDP_MAX_NX * GDDP_MAX_NX];
        for (int r = 0; r < nx; r++) {
            REAL s = Qx[r];
            for (int m = 0; m < nu; m++) s += Qux[m * nx + r] * kf[(size_t)i * nu + m];
            Vx[r] = s;
            for (int c = 0; c < nx; c++) {
                REAL t = Qxx[r * nx + c];
                for (int m = 0; m < nu; m++) t += Qux[m * nx + r] * K[((size_t)i * nu + m) * nx + c];
                Vn[r * nx + c] = t;
            }
        }
        for (int r = 0; r < nx; r++)
            for (int c = 0; c < nx; c++) Vxx[r * nx + c] = (REAL)0.5 * (Vn[r * nx + c] + Vn[c * nx + r]);
    }
    *dV1 = dv;
    return 1;
}

static void FN(solve_one)(const gddp_problem *P, int b, gddp_result *O) {
    const int nx = P->nx, nu = P->nu, N = P->N;
    REAL *xb = malloc(sizeof(REAL) * (size_t)(N + 1) * nx), *xn = malloc(sizeof(REAL) * (size_t)(N + 1) * nx);
    REAL *ub = malloc(sizeof(REAL) * (size_t)N * nu), *un = malloc(sizeof(REAL) * (size_t)N * nu);
    REAL *K = malloc(sizeof(REAL) * (size_t)N * nu * nx), *kf = malloc(sizeof(REAL) * (size_t)N * nu);
    for (int i = 0; i < N; i++)
        for (int m = 0; m < nu; m++) un[(size_t)i * nu + m] = P->u_init ? (REAL)P->u_init[((size_t)b * N + i) * nu + m] : (REAL)P->uh[m];
    REAL J = FN(rollout)(P, b, NULL, un, NULL, NULL, 0, xb, ub);   /* open loop: ub <- un, xb <- states */
    REAL rho = 0;
    int rtn = 0, iter = 0;
    long long sweeps = 0, rollouts = 1;
    const REAL tol = (REAL)P->tol;
    for (iter = 0; iter < P->iter_max; iter++) {
        int ok = 0;
        REAL dV1 = 0;
        while (1) {
            sweeps++;
            ok = FN(sweep)(P, b, xb, ub, rho, K, kf, &dV1);
            if (ok) break;
            rho = rho * 4 > (REAL)1e-6 ? rho * 4 : (REAL)1e-6;
            if (rho > (REAL)1e10) break;
        }
        if (!ok) { rtn = -4; break; }
        if (-dV1 <= tol * ((REAL)1 + FABS(J))) { rtn = 1; break; }   /* stationary: nothing left to gain */
        int accepted = 0;
        REAL Jn = 0, alpha = 1;
        for (int s = 0; s < 11; s++, alpha *= (REAL)0.5) {
            rollouts++;
            Jn = FN(rollout)(P, b, xb, ub, K, kf, alpha, xn, un);
            if (Jn < J) { accepted = 1; break; }
        }
        if (!accepted) {
            rho = rho * 4 > (REAL)1e-6 ? rho * 4 : (REAL)1e-6;
            if (rho > (REAL)1e10) { rtn = -4; break; }
            continue;
        }
        const REAL dJ = J - Jn;
        J = Jn;
        REAL *t;
        t = xb; xb = xn; xn = t;
        t = ub; ub = un; un = t;
        rho = rho / 4;
        if (rho < (REAL)1e-9) rho = 0;
        if (dJ <= tol * ((REAL)1 + FABS(J))) { rtn = 1; iter++; break; }
    }
    O->rtn[b] = rtn; O->iters[b] = iter; O->cost[b] = (double)J;
    if (O->stats) { O->stats[(size_t)b * 2] = sweeps; O->stats[(size_t)b * 2 + 1] = rollouts; }
    for (size_t e = 0; e < (size_t)(N + 1) * nx; e++) O->x[(size_t)b * (N + 1) * nx + e] = (double)xb[e];
    for (size_t e = 0; e < (size_t)N * nu; e++) O->u[(size_t)b * N * nu + e] = (double)ub[e];
    free(xb); free(xn); free(ub); free(un); free(K); free(kf);
}
#undef FN
#undef CAT
#undef CAT_
