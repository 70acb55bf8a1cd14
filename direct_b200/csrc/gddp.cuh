// gddp.cuh -- generic unconstrained batched DDP for sm_100a (model (B) of SURVEY.md section 8(d); include/direct_gddp.h).
//
// One warp owns one trajectory for the whole solve (persistent kernel, atomic work queue).  The backward sweep keeps the
// augmented matrix  [Qxx Qxu Qx; Qux Quu Qu]  one COLUMN PER LANE in registers (z order [x(nx); u(nu)], lane nz holds the
// gradient column): per knot every lane builds its column a_j of [A | B] from the analytic Jacobian (selects, no
// indexing), forms w_j = Vxx a_j, and the matrix entries Q[r][j] = a_r . w_j from the columns published in shared memory;
// the nu control rows are then eliminated by right-looking Cholesky pivots (pivot broadcast by shuffle, multipliers through
// shared memory), the gains come from one nu x nu back-substitution per lane, and the trailing block IS the new Vxx, Vx
// (Schur complement with the regularised Quu).  Columns of [A | B] are never dense: beside the identity entry a column
// touches at most two groups of three rows (dp/dv; dv/d(rpy), d(rpy')/d(rpy); d(rpy')/d(w), dw'/dw; dv/d(thrust);
// dw'/d(tau)), so a column is six numbers and Vxx a, a_r . w cost 7 and <= 6 multiply-adds per entry instead of 12.  No Jacobian, no Q matrix and no value function ever touches HBM; per knot
// the sweep reads nx + nu numbers and writes nu (nx + 1) gains.  The rollout keeps the state in registers of every lane.
// Arithmetic type R = float (BASELINE.json's fp32 configuration) or double.
#ifndef DIRECT_B200_GDDP_CUH_
#define DIRECT_B200_GDDP_CUH_

namespace gddp {

template <class R> struct Args {
    int B, N, iter_max;
    R dt, tol;
    const double *x0, *xg, *u_init;
    R q[12], qf[12], r[4], uh[4];
    int32_t *rtn, *iters;
    double *cost, *x, *u;
    long long *stats;
    R *ws;
    long long ws_stride;
    unsigned int *counter;
};

template <int MODEL> struct Dim {
    static constexpr int NX = MODEL == 1 ? 12 : 6;
    static constexpr int NU = MODEL == 1 ? 4 : 3;
    static constexpr int NZ = NX + NU;
    static constexpr int NT = MODEL == 1 ? 6 : 0;   // transcendentals of the point cached next to it: sin/cos of roll, pitch, yaw
    static constexpr int XS = NX + NT;              // stride of a stored point [x; trig]
};

__device__ __forceinline__ float g_sin(float x) { return sinf(x); }
__device__ __forceinline__ double g_sin(double x) { return sin(x); }
__device__ __forceinline__ float g_cos(float x) { return cosf(x); }
__device__ __forceinline__ double g_cos(double x) { return cos(x); }
__device__ __forceinline__ float g_rsqrt(float x) { return rsqrtf(x); }
__device__ __forceinline__ void g_sincos(float x, float *s, float *c) { sincosf(x, s, c); }
__device__ __forceinline__ void g_sincos(double x, double *s, double *c) { sincos(x, s, c); }
__device__ __forceinline__ double g_rsqrt(double x) { return 1.0 / sqrt(x); }
__device__ __forceinline__ float g_abs(float x) { return fabsf(x); }
__device__ __forceinline__ double g_abs(double x) { return fabs(x); }
template <class R> __device__ __forceinline__ R sel3(R a, R b, R c, int k) { return k == 0 ? a : (k == 1 ? b : c); }

// ---- models: continuous dynamics f(x, u), evaluated identically by every lane (uniform registers) ---------------------
template <int MODEL, class R> struct Model;

template <class R> struct Model<0, R> {   // 3D double integrator
    static __device__ __forceinline__ void trig(const R *, int, R *) {}
    static __device__ __forceinline__ void f(const R *x, const R *u, const R *, R *fx) {
#pragma unroll
        for (int a = 0; a < 3; a++) { fx[a] = x[3 + a]; fx[3 + a] = u[a]; }
    }
    // Column j of [A | B] = [I + dt df/dx | dt df/du] = (identity entry if has_id) + aA on rows ga .. ga+2 + aB on rows gb .. gb+2.
    __host__ __device__ static constexpr int ga(int j) { return j < 6 ? 0 : 3; }
    __host__ __device__ static constexpr int gb(int) { return 0; }
    __host__ __device__ static constexpr bool has_a(int j) { return j >= 3; }
    __host__ __device__ static constexpr bool has_b(int) { return false; }
    __host__ __device__ static constexpr bool has_id(int j) { return j < 6; }
    static __device__ __forceinline__ void column(const R *, const R *, const R *, R dt, int j, R *aA, R *aB) {
        const int k = j < 6 ? j - 3 : j - 6;   // dp/dv (j = 3..5), dv/da (j = 6..8)
#pragma unroll
        for (int t = 0; t < 3; t++) { aA[t] = (j >= 3 && t == k) ? dt : R(0); aB[t] = R(0); }
    }
};

template <class R> struct Model<1, R> {   // rigid-body quadrotor, Euler angles (Quadrotor.cpp:15-20 constants)
    // tg = {sin, cos} of roll, pitch, yaw: one sincos per angle in lanes 0..2, broadcast (every lane needs all six)
    static __device__ __forceinline__ void trig(const R *x, int lane, R *tg) {
        const R ang = lane == 0 ? x[6] : (lane == 1 ? x[7] : x[8]);
        R s, c;
        g_sincos(ang, &s, &c);
#pragma unroll
        for (int t = 0; t < 3; t++) { tg[2 * t] = __shfl_sync(0xffffffffu, s, t); tg[2 * t + 1] = __shfl_sync(0xffffffffu, c, t); }
    }
    static __device__ __forceinline__ void f(const R *x, const R *u, const R *tg, R *fx) {
        // divisions by the constants are multiplications by their (compile-time) reciprocals; one division per point (1 / cos pitch)
        const R m = R(0.98), g = R(9.81), Jx = R(2.64e-3), Jy = R(2.64e-3), Jz = R(4.96e-3);
        const R im = R(1) / m, iJx = R(1) / Jx, iJy = R(1) / Jy, iJz = R(1) / Jz;
        const R p = x[9], q = x[10], r = x[11];
        const R sp = tg[0], cp = tg[1], st = tg[2], ct = tg[3], ss = tg[4], cs = tg[5];
        const R ict = R(1) / ct, tt = st * ict, fm = u[0] * im;
        fx[0] = x[3]; fx[1] = x[4]; fx[2] = x[5];
        fx[3] = fm * (cp * st * cs + sp * ss); fx[4] = fm * (cp * st * ss - sp * cs); fx[5] = fm * (cp * ct) - g;
        const R sqcr = sp * q + cp * r, cqsr = cp * q - sp * r;
        fx[6] = p + tt * sqcr; fx[7] = cqsr; fx[8] = sqcr * ict;
        fx[9] = (u[1] - (Jz - Jy) * q * r) * iJx; fx[10] = (u[2] - (Jx - Jz) * p * r) * iJy; fx[11] = (u[3] - (Jy - Jx) * p * q) * iJz;
    }
    __host__ __device__ static constexpr int ga(int j) { return j < 6 ? 0 : (j < 9 ? 3 : (j < 12 ? 6 : (j == 12 ? 3 : 9))); }
    __host__ __device__ static constexpr int gb(int j) { return j < 6 ? 0 : (j < 9 ? 6 : (j < 12 ? 9 : 0)); }
    __host__ __device__ static constexpr bool has_a(int j) { return j >= 3; }
    __host__ __device__ static constexpr bool has_b(int j) { return j >= 6 && j < 12; }
    __host__ __device__ static constexpr bool has_id(int j) { return j < 12; }
    static __device__ __forceinline__ void column(const R *x, const R *u, const R *tg, R dt, int j, R *aA, R *aB) {
        const R m = R(0.98), Jx = R(2.64e-3), Jy = R(2.64e-3), Jz = R(4.96e-3);
        const R im = R(1) / m, iJx = R(1) / Jx, iJy = R(1) / Jy, iJz = R(1) / Jz;
        const R p = x[9], q = x[10], r = x[11];
        const R sp = tg[0], cp = tg[1], st = tg[2], ct = tg[3], ss = tg[4], cs = tg[5];
        const R ict = R(1) / ct, tt = st * ict, fm = u[0] * im;
        const R sqcr = sp * q + cp * r, cqsr = cp * q - sp * r;
        const int grp = j / 3, k = j - 3 * grp;
        aA[0] = aA[1] = aA[2] = R(0); aB[0] = aB[1] = aB[2] = R(0);
        if (grp == 1) {          // velocity columns: dp/dv
#pragma unroll
            for (int t = 0; t < 3; t++) aA[t] = (t == k) ? dt : R(0);
        } else if (grp == 2) {   // attitude columns: dv/d(rpy) on the v rows, d(rpy rates)/d(rpy) on the rpy rows
            aA[0] = dt * (fm * sel3(-sp * st * cs + cp * ss, cp * ct * cs, -cp * st * ss + sp * cs, k));
            aA[1] = dt * (fm * sel3(-sp * st * ss - cp * cs, cp * ct * ss, cp * st * cs + sp * ss, k));
            aA[2] = dt * (fm * sel3(-sp * ct, -cp * st, R(0), k));
            aB[0] = dt * sel3(tt * cqsr, sqcr * ict * ict, R(0), k);
            aB[1] = dt * sel3(-sqcr, R(0), R(0), k);
            aB[2] = dt * sel3(cqsr * ict, sqcr * st * ict * ict, R(0), k);
        } else if (grp == 3) {   // body-rate columns: d(rpy rates)/d(omega) on the rpy rows, d(omega dot)/d(omega) on the omega rows
            aA[0] = dt * sel3(R(1), sp * tt, cp * tt, k);
            aA[1] = dt * sel3(R(0), cp, -sp, k);
            aA[2] = dt * sel3(R(0), sp * ict, cp * ict, k);
            aB[0] = dt * sel3(R(0), -(Jz - Jy) * r * iJx, -(Jz - Jy) * q * iJx, k);
            aB[1] = dt * sel3(-(Jx - Jz) * r * iJy, R(0), -(Jx - Jz) * p * iJy, k);
            aB[2] = dt * sel3(-(Jy - Jx) * q * iJz, -(Jy - Jx) * p * iJz, R(0), k);
        } else if (j == 12) {    // thrust column: dv/df
            aA[0] = dt * ((cp * st * cs + sp * ss) * im); aA[1] = dt * ((cp * st * ss - sp * cs) * im); aA[2] = dt * ((cp * ct) * im);
        } else if (j > 12) {     // torque columns: d(omega dot)/d(tau)
            aA[0] = j == 13 ? dt * iJx : R(0); aA[1] = j == 14 ? dt * iJy : R(0); aA[2] = j == 15 ? dt * iJz : R(0);
        }
    }
};

// Per-warp shared scratch (elements of R).
template <int MODEL> struct Smem {
    using D = Dim<MODEL>;
    enum {
        V0 = 0,                                  // value-function Hessian, two buffers used alternately
        V1 = D::NX * D::NX,
        S = 2 * D::NX * D::NX,                   // unsymmetrised Schur complement
        VX = 3 * D::NX * D::NX,                  // V_x
        AC = VX + D::NX,                         // published columns of [A | B]: AC[j * 6 + t] = {aA[3], aB[3]} of column j
        LM = AC + D::NZ * D::NX,                 // multipliers of the nu pivots: LM[p * (NZ + 1) + lane]
        XU = LM + D::NU * (D::NZ + 1),           // current knot's [x; trig; u] (and the rollout's staging)
        TOTAL = ((XU + D::XS + D::NU + 3) / 4) * 4
    };
};

template <int MODEL, class R> struct Ws {
    using D = Dim<MODEL>;
    long long xb, xn, ub, un, K, kf, total;
    __host__ __device__ explicit Ws(int N) {
        long long o = 0;
        xb = o; o += (long long)(N + 1) * D::XS;
        xn = o; o += (long long)(N + 1) * D::XS;
        ub = o; o += (long long)N * D::NU;
        un = o; o += (long long)N * D::NU;
        K = o; o += (long long)N * D::NU * D::NX;
        kf = o; o += (long long)N * D::NU;
        total = (o + 3) & ~3LL;
    }
};

// Closed-loop rollout from x0 with controls ub + alpha k + K (x - xb) (K == nullptr: open loop with the controls in `un`).
template <int MODEL, class R>
__device__ __forceinline__ R rollout(const Args<R> &A, int b, int lane, R *sm, const R *xb, const R *ub, const R *K, const R *kf, R alpha,
                                  R *xn, R *un) {
    using D = Dim<MODEL>;
    constexpr int NX = D::NX, NU = D::NU, NT = D::NT, XS = D::XS;
    R x[NX], xg[NX], tg[NT > 0 ? NT : 1];
#pragma unroll
    for (int a = 0; a < NX; a++) { x[a] = (R)A.x0[(long long)b * NX + a]; xg[a] = (R)A.xg[(long long)b * NX + a]; }
    if (lane < NX) xn[lane] = x[lane];
    R J = R(0);
    R *st = sm + Smem<MODEL>::XU;
    for (int i = 0; i < A.N; i++) {
        R u[NU];
        if (K) {
            // lanes 0..NU-1 each form one control; the old state comes through shared memory
            if (lane < NX) st[lane] = xb[(long long)i * XS + lane];
            __syncwarp();
            R v = R(0);
            if (lane < NU) {
                v = ub[(long long)i * NU + lane] + alpha * kf[(long long)i * NU + lane];
                const R *Kr = K + ((long long)i * NU + lane) * NX;
#pragma unroll
                for (int a = 0; a < NX; a++) v += Kr[a] * (x[a] - st[a]);
            }
#pragma unroll
            for (int m = 0; m < NU; m++) u[m] = __shfl_sync(0xffffffffu, v, m);
            __syncwarp();
        } else {
#pragma unroll
            for (int m = 0; m < NU; m++) u[m] = un[(long long)i * NU + m];
        }
        if (lane < NU) {
            R um = u[0];
#pragma unroll
            for (int m = 1; m < NU; m++) if (lane == m) um = u[m];
            un[(long long)i * NU + lane] = um;
        }
        R c = R(0);
#pragma unroll
        for (int a = 0; a < NX; a++) { const R d = x[a] - xg[a]; c += A.q[a] * d * d; }
#pragma unroll
        for (int m = 0; m < NU; m++) { const R d = u[m] - A.uh[m]; c += A.r[m] * d * d; }
        J += R(0.5) * A.dt * c;
        R fx[NX];
        Model<MODEL, R>::trig(x, lane, tg);
        if (NT > 0 && lane < NT) {   // the sweep linearises at this point: keep its sin/cos next to it
            R tv = tg[0];
#pragma unroll
            for (int a = 1; a < NT; a++) if (lane == a) tv = tg[a];
            xn[(long long)i * XS + NX + lane] = tv;
        }
        Model<MODEL, R>::f(x, u, tg, fx);
#pragma unroll
        for (int a = 0; a < NX; a++) x[a] = x[a] + A.dt * fx[a];
        if (lane < NX) {
            R xv = x[0];
#pragma unroll
            for (int a = 1; a < NX; a++) if (lane == a) xv = x[a];
            xn[(long long)(i + 1) * XS + lane] = xv;
        }
    }
    R c = R(0);
#pragma unroll
    for (int a = 0; a < NX; a++) { const R d = x[a] - xg[a]; c += A.qf[a] * d * d; }
    return J + R(0.5) * c;
}

// Backward sweep with the regularised Quu.  Returns false when a pivot is not positive; *dV1 = sum_i k_i' Qu_i.
template <int MODEL, class R>
__device__ __forceinline__ bool sweep(const Args<R> &A, int b, int lane, R *sm, const R *xb, const R *ub, R rho, R *K, R *kf, R *dV1,
                                   long long *knots) {
    using D = Dim<MODEL>;
    using SM = Smem<MODEL>;
    constexpr int NX = D::NX, NU = D::NU, NZ = D::NZ, NT = D::NT, XS = D::XS;
    const int N = A.N;
    R xg[NX];
#pragma unroll
    for (int a = 0; a < NX; a++) xg[a] = (R)A.xg[(long long)b * NX + a];
    // terminal value function
    for (int e = lane; e < NX * NX; e += 32) sm[SM::V0 + e] = (e / NX == e % NX) ? A.qf[e % NX] : R(0);
    if (lane < NX) {
        R qfl = A.qf[0], xgl = xg[0];
#pragma unroll
        for (int a = 1; a < NX; a++) if (lane == a) { qfl = A.qf[a]; xgl = xg[a]; }
        sm[SM::VX + lane] = qfl * (xb[(long long)N * XS + lane] - xgl);
    }
    // this lane's cost weights: Hessian diagonal entry of its own column
    R wdiag = R(0);
#pragma unroll
    for (int a = 0; a < NX; a++) if (lane == a) wdiag = A.dt * A.q[a];
#pragma unroll
    for (int m = 0; m < NU; m++) if (lane == NX + m) wdiag = A.dt * A.r[m] + rho;
    R dv = R(0);
    int vb = 0;   // which V buffer holds the current value function
    bool ok = true;
    R nxt = R(0);
    if (lane < XS) nxt = xb[(long long)(N - 1) * XS + lane];
    else if (lane < XS + NU) nxt = ub[(long long)(N - 1) * NU + lane - XS];
    __syncwarp();
    long long nk = 0;   // local counter: a by-pointer counter would be a local-memory read-modify-write per knot
    for (int i = N - 1; i >= 0; i--) {
        nk++;
        // ---- the knot's point, uniform in every lane; the next knot's is fetched meanwhile -------------------------
        if (lane < XS + NU) sm[SM::XU + lane] = nxt;
        if (i > 0) {
            if (lane < XS) nxt = xb[(long long)(i - 1) * XS + lane];
            else if (lane < XS + NU) nxt = ub[(long long)(i - 1) * NU + lane - XS];
        }
        __syncwarp();
        R x[NX], u[NU], tg[NT > 0 ? NT : 1];
#pragma unroll
        for (int a = 0; a < NX; a++) x[a] = sm[SM::XU + a];
#pragma unroll
        for (int a = 0; a < NT; a++) tg[a] = sm[SM::XU + NX + a];   // sin/cos cached by the rollout that produced the point
#pragma unroll
        for (int m = 0; m < NU; m++) u[m] = sm[SM::XU + XS + m];
        // ---- column of [A | B] (six numbers), w = Vxx a (gradient lane: w = Vx) ------------------------------------------
        R w[NX];
        const R *V = sm + (vb ? SM::V1 : SM::V0);
        if (lane < NZ) {
            R aA[3], aB[3];
            Model<MODEL, R>::column(x, u, tg, A.dt, lane, aA, aB);
            const int ga = Model<MODEL, R>::ga(lane), gb = Model<MODEL, R>::gb(lane);
            const R idw = Model<MODEL, R>::has_id(lane) ? R(1) : R(0);
            const R *v0 = V + (lane < NX ? lane : 0) * NX;   // V is symmetric: column k = row k
            const R *vA = V + ga * NX, *vB = V + gb * NX;
#pragma unroll
            for (int r = 0; r < NX; r++) {
                R acc = idw * v0[r];
#pragma unroll
                for (int t = 0; t < 3; t++) acc += vA[t * NX + r] * aA[t];
#pragma unroll
                for (int t = 0; t < 3; t++) acc += vB[t * NX + r] * aB[t];
                w[r] = acc;
            }
#pragma unroll
            for (int t = 0; t < 3; t++) { sm[SM::AC + lane * 6 + t] = aA[t]; sm[SM::AC + lane * 6 + 3 + t] = aB[t]; }
        } else {
#pragma unroll
            for (int r = 0; r < NX; r++) w[r] = sm[SM::VX + r];
        }
        __syncwarp();
        // ---- Q[r][lane] = a_r . w  (+ cost terms); the structure of column r is known at compile time ----------------------
        R qc[NZ];
#pragma unroll
        for (int r = 0; r < NZ; r++) {
            R acc = Model<MODEL, R>::has_id(r) ? w[r < NX ? r : 0] : R(0);
            if (Model<MODEL, R>::has_a(r)) {
#pragma unroll
                for (int t = 0; t < 3; t++) acc += sm[SM::AC + r * 6 + t] * w[Model<MODEL, R>::ga(r) + t];
            }
            if (Model<MODEL, R>::has_b(r)) {
#pragma unroll
                for (int t = 0; t < 3; t++) acc += sm[SM::AC + r * 6 + 3 + t] * w[Model<MODEL, R>::gb(r) + t];
            }
            qc[r] = acc;
        }
        if (lane < NZ) {
#pragma unroll
            for (int r = 0; r < NZ; r++) if (lane == r) qc[r] += wdiag;
        } else if (lane == NZ) {
#pragma unroll
            for (int r = 0; r < NX; r++) qc[r] += A.dt * A.q[r] * (x[r] - xg[r]);
#pragma unroll
            for (int m = 0; m < NU; m++) qc[NX + m] += A.dt * A.r[m] * (u[m] - A.uh[m]);
        }
        // ---- eliminate the nu control rows --------------------------------------------------------------------------------
        R mp[NU];
#pragma unroll
        for (int p = 0; p < NU; p++) {
            const R d = __shfl_sync(0xffffffffu, qc[NX + p], NX + p);
            if (!(d > R(0))) { ok = false; break; }
            const R ri = g_rsqrt(d);
            const R m = (lane == NX + p) ? d * ri : qc[NX + p] * ri;
            mp[p] = m;
            if (lane <= NZ) sm[SM::LM + p * (NZ + 1) + lane] = m;
            __syncwarp();
            const R *Lm = sm + SM::LM + p * (NZ + 1);
#pragma unroll
            for (int r = 0; r < NZ; r++) if (r < NX || r > NX + p) qc[r] -= Lm[r] * m;   // eliminated rows are never read again
            if (lane == NZ) dv -= m * m;
        }
        if (!ok) break;
        // ---- gains: L' z = y with y = this lane's multipliers; [k | K] = -z ------------------------------------------------
        if (lane < NX || lane == NZ) {
            R z[NU];
#pragma unroll
            for (int p = NU - 1; p >= 0; p--) {
                R s = mp[p];
#pragma unroll
                for (int t = p + 1; t < NU; t++) s -= sm[SM::LM + p * (NZ + 1) + NX + t] * z[t];
                z[p] = s / sm[SM::LM + p * (NZ + 1) + NX + p];
            }
            if (lane < NX) {
#pragma unroll
                for (int m = 0; m < NU; m++) K[((long long)i * NU + m) * NX + lane] = -z[m];
            } else {
#pragma unroll
                for (int m = 0; m < NU; m++) kf[(long long)i * NU + m] = -z[m];
            }
        }
        // ---- the trailing block is the new value function: Vx (gradient lane), Vxx = sym(.) --------------------------------
        if (lane < NX) {
#pragma unroll
            for (int r = 0; r < NX; r++) sm[SM::S + r * NX + lane] = qc[r];
        } else if (lane == NZ) {
#pragma unroll
            for (int r = 0; r < NX; r++) sm[SM::VX + r] = qc[r];
        }
        __syncwarp();
        if (lane < NX) {
            R *Vn = sm + (vb ? SM::V0 : SM::V1);
#pragma unroll
            for (int r = 0; r < NX; r++) Vn[lane * NX + r] = R(0.5) * (qc[r] + sm[SM::S + lane * NX + r]);
        }
        vb ^= 1;
        __syncwarp();
    }
    *dV1 = __shfl_sync(0xffffffffu, dv, NZ);
    *knots += nk;
    return ok;
}

// Resident CTAs per SM: the sweep is latency-bound, so occupancy pays until the register cap starts to spill.  Measured
// on B200 (fp32, B = 4096 / 65536, k solves/s): 2 CTAs (157 registers) 555 / 669, 4 CTAs (128) 803 / 1070, 5 CTAs (96) 758 / 1125.
#ifndef GDDP_MIN_BLOCKS
#define GDDP_MIN_BLOCKS (sizeof(R) == 4 ? 4 : 2)
#endif
template <int MODEL, class R> __global__ void __launch_bounds__(128, GDDP_MIN_BLOCKS) gddp_kernel(Args<R> A) {
    using D = Dim<MODEL>;
    constexpr int NX = D::NX, NU = D::NU;
    extern __shared__ __align__(16) unsigned char gsm_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    R *sm = reinterpret_cast<R *>(gsm_raw) + warp * Smem<MODEL>::TOTAL;
    const Ws<MODEL, R> wl(A.N);
    R *ws = A.ws + ((long long)blockIdx.x * wpb + warp) * A.ws_stride;
    const int N = A.N;
    while (true) {
        unsigned int b = 0;
        if (lane == 0) b = atomicAdd(A.counter, 1u);
        b = __shfl_sync(0xffffffffu, b, 0);
        if (b >= (unsigned int)A.B) break;
        const long long clk0 = clock64();
        R *xb = ws + wl.xb, *xn = ws + wl.xn, *ub = ws + wl.ub, *un = ws + wl.un, *K = ws + wl.K, *kf = ws + wl.kf;
        for (int e = lane; e < N * NU; e += 32) un[e] = A.u_init ? (R)A.u_init[(long long)b * N * NU + e] : A.uh[e % NU];
        __syncwarp();
        R J = rollout<MODEL, R>(A, (int)b, lane, sm, nullptr, nullptr, nullptr, nullptr, R(0), xb, un);
        { R *t = ub; ub = un; un = t; }
        __syncwarp();
        R rho = R(0);
        int rtn = 0, iter = 0;
        long long sweeps = 0, rollouts = 1, knots = 0;
        for (iter = 0; iter < A.iter_max; iter++) {
            bool ok = false;
            R dV1 = R(0);
            while (true) {
                sweeps++;
                ok = sweep<MODEL, R>(A, (int)b, lane, sm, xb, ub, rho, K, kf, &dV1, &knots);
                __syncwarp();
                if (ok) break;
                rho = rho * R(4) > R(1e-6) ? rho * R(4) : R(1e-6);
                if (rho > R(1e10)) break;
            }
            if (!ok) { rtn = -4; break; }
            if (-dV1 <= A.tol * (R(1) + g_abs(J))) { rtn = 1; break; }
            bool accepted = false;
            R Jn = R(0), alpha = R(1);
            for (int s = 0; s < 11; s++, alpha *= R(0.5)) {
                rollouts++;
                Jn = rollout<MODEL, R>(A, (int)b, lane, sm, xb, ub, K, kf, alpha, xn, un);
                __syncwarp();
                if (Jn < J) { accepted = true; break; }
            }
            if (!accepted) {
                rho = rho * R(4) > R(1e-6) ? rho * R(4) : R(1e-6);
                if (rho > R(1e10)) { rtn = -4; break; }
                continue;
            }
            const R dJ = J - Jn;
            J = Jn;
            { R *t = xb; xb = xn; xn = t; t = ub; ub = un; un = t; }
            rho = rho / R(4);
            if (rho < R(1e-9)) rho = R(0);
            if (dJ <= A.tol * (R(1) + g_abs(J))) { rtn = 1; iter++; break; }
        }
        if (lane == 0) {
            A.rtn[b] = rtn; A.iters[b] = iter; A.cost[b] = (double)J;
            if (A.stats) {
                long long *S = A.stats + (long long)b * 4;
                S[0] = sweeps; S[1] = rollouts; S[2] = knots; S[3] = clock64() - clk0;
            }
        }
        for (int e = lane; e < (N + 1) * NX; e += 32) A.x[(long long)b * (N + 1) * NX + e] = (double)xb[(long long)(e / NX) * D::XS + e % NX];
        for (int e = lane; e < N * NU; e += 32) A.u[(long long)b * N * NU + e] = (double)ub[e];
        __syncwarp();
    }
}

}  // namespace gddp
#endif
