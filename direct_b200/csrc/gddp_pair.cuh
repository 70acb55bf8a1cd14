// gddp_pair.cuh -- generic unconstrained batched DDP, TWO trajectories per warp (model (B) of SURVEY.md section 8(d)).
//
// gddp.cuh keeps the augmented matrix [Qxx Qxu Qx; Qux Quu Qu] one column per lane: nz = 16 columns plus the gradient column =
// 17 busy lanes of 32.  Here a trajectory lives in a HALF-warp: the 16 columns in 16 lanes, and the gradient column stored
// row-distributed - lane r keeps entry r of it as one extra scalar g.  Entry r is a_r . Vx + the cost gradient of row r, and
// lane r already holds column a_r; a pivot updates it by g_r -= m_r * (g_pivot / sqrt(d)), where m_r is the lane's own multiplier:
// one shuffle and one multiply-add instead of a 17th lane.  The two halves of a warp run two trajectories of a pair in lockstep
// (same code, width-16 shuffles, stores predicated per half); a half that has converged, or whose line search has accepted,
// idles until its partner catches up (measured on the bench problems: 92 % of the pair's steps are useful for both halves).
// Same arithmetic, in the same order, as gddp.cuh (one difference: the gains are multiplied by the pivot's rsqrt instead of divided by
// its square root); the checker is the same oracle (oracle/gddp_oracle.c, parity unpinned).
#ifndef DIRECT_B200_GDDP_PAIR_CUH_
#define DIRECT_B200_GDDP_PAIR_CUH_

#include "gddp.cuh"

namespace gddp {

template <class R> __device__ __forceinline__ R shfl16(R v, int src) { return __shfl_sync(0xffffffffu, v, src, 16); }

template <int MODEL, class R> __device__ __forceinline__ void trig16(const R *x, int hl, R *tg) {
    if constexpr (MODEL == 1) {   // one sincos per angle in lanes 0..2 of the half, broadcast inside the half
        const R ang = hl == 0 ? x[6] : (hl == 1 ? x[7] : x[8]);
        R s, c;
        g_sincos(ang, &s, &c);
#pragma unroll
        for (int t = 0; t < 3; t++) { tg[2 * t] = shfl16(s, t); tg[2 * t + 1] = shfl16(c, t); }
    }
}

// Per-half shared scratch.  Rows of V and of the unsymmetrised Schur complement are padded to NX + 1 elements: lane k reads / writes row k,
// and an odd row stride puts the 16 lanes of a half on 16 different banks (stride 12 gave two-way conflicts on every access).  TOTAL is
// 16 mod 32, so the two halves of a warp, which execute every access together, use complementary banks.
template <int MODEL> struct SmemP {
    using D = Dim<MODEL>;
    enum {
        LD = D::NX + 1,
        V0 = 0,
        V1 = D::NX * LD,
        S = 2 * D::NX * LD,
        VX = 3 * D::NX * LD,
        AC = VX + D::NX,
        LM = AC + D::NZ * 6,
        XU = LM + D::NU * (D::NZ + 1),
        END = XU + D::XS + D::NU,
        TOTAL = ((END + 15) / 32) * 32 + 16
    };
};

// Register array element selected by a lane index (a select chain: no local memory).
template <int n, class R> __device__ __forceinline__ R pick(const R *v, int k) {
    R r = v[0];
#pragma unroll
    for (int a = 1; a < n; a++) if (k == a) r = v[a];
    return r;
}

// Closed-loop rollout of the half's trajectory (closed == false: open loop with the controls in `un`).  `act` predicates the stores.
template <int MODEL, class R>
__device__ __forceinline__ R rollout2(const Args<R> &A, int b, int hl, bool act, R *sm, const R *xb, const R *ub, const R *K, const R *kf,
                                      bool closed, R alpha, R *xn, R *un) {
    using D = Dim<MODEL>;
    constexpr int NX = D::NX, NU = D::NU, NT = D::NT, XS = D::XS, NTT = NT > 0 ? NT : 1;
    R x[NX], xg[NX], tg[NTT];
#pragma unroll
    for (int a = 0; a < NX; a++) { x[a] = (R)A.x0[(long long)b * NX + a]; xg[a] = (R)A.xg[(long long)b * NX + a]; }
    if (act && hl < NX) xn[hl] = pick<NX>(x, hl);
    R J = R(0);
    R *st = sm + SmemP<MODEL>::XU;
    // the knot's old state, control, feedforward and gain row are fetched one knot ahead (the recursion would otherwise wait for
    // them at every knot: they were 14 % of the kernel's stall samples)
    R stn = R(0), ubn = R(0), kfn = R(0), Kn[NX];
#pragma unroll
    for (int a = 0; a < NX; a++) Kn[a] = R(0);
    if (closed) {
        if (hl < NX) stn = xb[hl];
        if (hl < NU) {
            ubn = ub[hl]; kfn = kf[hl];
#pragma unroll
            for (int a = 0; a < NX; a++) Kn[a] = K[(long long)hl * NX + a];
        }
    }
    for (int i = 0; i < A.N; i++) {
        R u[NU];
        if (closed) {
            if (hl < NX) st[hl] = stn;
            const R ubc = ubn, kfc = kfn;
            R Kc[NX];
#pragma unroll
            for (int a = 0; a < NX; a++) Kc[a] = Kn[a];
            if (i + 1 < A.N) {
                if (hl < NX) stn = xb[(long long)(i + 1) * XS + hl];
                if (hl < NU) {
                    ubn = ub[(long long)(i + 1) * NU + hl]; kfn = kf[(long long)(i + 1) * NU + hl];
                    const R *Kr = K + ((long long)(i + 1) * NU + hl) * NX;
#pragma unroll
                    for (int a = 0; a < NX; a++) Kn[a] = Kr[a];
                }
            }
            __syncwarp();
            R v = R(0);
            if (hl < NU) {
                v = ubc + alpha * kfc;
#pragma unroll
                for (int a = 0; a < NX; a++) v += Kc[a] * (x[a] - st[a]);
            }
#pragma unroll
            for (int m = 0; m < NU; m++) u[m] = shfl16(v, m);
            __syncwarp();
        } else {
#pragma unroll
            for (int m = 0; m < NU; m++) u[m] = un[(long long)i * NU + m];
        }
        if (act && hl < NU) un[(long long)i * NU + hl] = pick<NU>(u, hl);
        R c = R(0);
#pragma unroll
        for (int a = 0; a < NX; a++) { const R d = x[a] - xg[a]; c += A.q[a] * d * d; }
#pragma unroll
        for (int m = 0; m < NU; m++) { const R d = u[m] - A.uh[m]; c += A.r[m] * d * d; }
        J += R(0.5) * A.dt * c;
        R fx[NX];
        trig16<MODEL, R>(x, hl, tg);
        if (NT > 0 && act && hl < NT) xn[(long long)i * XS + NX + hl] = pick<NTT>(tg, hl);
        Model<MODEL, R>::f(x, u, tg, fx);
#pragma unroll
        for (int a = 0; a < NX; a++) x[a] = x[a] + A.dt * fx[a];
        if (act && hl < NX) xn[(long long)(i + 1) * XS + hl] = pick<NX>(x, hl);
    }
    R c = R(0);
#pragma unroll
    for (int a = 0; a < NX; a++) { const R d = x[a] - xg[a]; c += A.qf[a] * d * d; }
    return J + R(0.5) * c;
}

// Backward sweep of the half's trajectory.  Returns false when a pivot is not positive; *dV1 = sum_i k_i' Qu_i.  Both halves of the
// warp execute every knot until neither is (active and still ok); a half past its failure computes on but stores nothing.
template <int MODEL, class R>
__device__ __forceinline__ bool sweep2(const Args<R> &A, int b, int hl, bool act, R *sm, const R *xb, const R *ub, R rho, R *K, R *kf, R *dV1,
                                       long long *knots) {
    using D = Dim<MODEL>;
    using SM = SmemP<MODEL>;
    using M = Model<MODEL, R>;
    constexpr int NX = D::NX, NU = D::NU, NZ = D::NZ, NT = D::NT, XS = D::XS, NP = XS + NU, LD = SM::LD;   // NP entries of a staged point
    static_assert(NZ <= 16 && NP <= 32, "a trajectory must fit a half-warp");
    const int N = A.N;
    R xg[NX];
#pragma unroll
    for (int a = 0; a < NX; a++) xg[a] = (R)A.xg[(long long)b * NX + a];
    for (int e = hl; e < NX * NX; e += 16) sm[SM::V0 + (e / NX) * LD + e % NX] = (e / NX == e % NX) ? A.qf[e % NX] : R(0);
    if (hl < NX) sm[SM::VX + hl] = pick<NX>(A.qf, hl) * (xb[(long long)N * XS + hl] - pick<NX>(xg, hl));
    // this lane's row: Hessian diagonal entry, weight and target of the cost gradient, where its own z entry sits in the staged point
    R wdiag = R(0), gw = R(0), gt = R(0);
#pragma unroll
    for (int a = 0; a < NX; a++) if (hl == a) { wdiag = A.dt * A.q[a]; gw = A.dt * A.q[a]; gt = xg[a]; }
#pragma unroll
    for (int m = 0; m < NU; m++) if (hl == NX + m) { wdiag = A.dt * A.r[m] + rho; gw = A.dt * A.r[m]; gt = A.uh[m]; }
    const int gsrc = hl < NX ? hl : XS + (hl - NX);
    R dv = R(0);
    int vb = 0;
    bool ok = true;
    const int e0 = hl, e1 = hl + 16;
    R nxt0 = R(0), nxt1 = R(0);
    if (e0 < NP) nxt0 = e0 < XS ? xb[(long long)(N - 1) * XS + e0] : ub[(long long)(N - 1) * NU + e0 - XS];
    if (e1 < NP) nxt1 = e1 < XS ? xb[(long long)(N - 1) * XS + e1] : ub[(long long)(N - 1) * NU + e1 - XS];
    __syncwarp();
    long long nk = 0;
    for (int i = N - 1; i >= 0; i--) {
        if (ok && act) nk++;
        if (e0 < NP) sm[SM::XU + e0] = nxt0;
        if (e1 < NP) sm[SM::XU + e1] = nxt1;
        if (i > 0) {
            if (e0 < NP) nxt0 = e0 < XS ? xb[(long long)(i - 1) * XS + e0] : ub[(long long)(i - 1) * NU + e0 - XS];
            if (e1 < NP) nxt1 = e1 < XS ? xb[(long long)(i - 1) * XS + e1] : ub[(long long)(i - 1) * NU + e1 - XS];
        }
        __syncwarp();
        R x[NX], u[NU], tg[NT > 0 ? NT : 1];
#pragma unroll
        for (int a = 0; a < NX; a++) x[a] = sm[SM::XU + a];
#pragma unroll
        for (int a = 0; a < NT; a++) tg[a] = sm[SM::XU + NX + a];
#pragma unroll
        for (int m = 0; m < NU; m++) u[m] = sm[SM::XU + XS + m];
        // ---- column of [A | B], w = Vxx a, and this lane's entry of the gradient column g = a . Vx + cost gradient ---------------
        R w[NX], g = R(0);
        const R *V = sm + (vb ? SM::V1 : SM::V0);
        if (hl < NZ) {
            R aA[3], aB[3];
            M::column(x, u, tg, A.dt, hl, aA, aB);
            const int ga = M::ga(hl), gb = M::gb(hl);
            const bool id = M::has_id(hl);
            const R idw = id ? R(1) : R(0);
            const R *v0 = V + (hl < NX ? hl : 0) * LD;
            const R *vA = V + ga * LD, *vB = V + gb * LD;
#pragma unroll
            for (int r = 0; r < NX; r++) {
                R acc = idw * v0[r];
#pragma unroll
                for (int t = 0; t < 3; t++) acc += vA[t * LD + r] * aA[t];
#pragma unroll
                for (int t = 0; t < 3; t++) acc += vB[t * LD + r] * aB[t];
                w[r] = acc;
            }
#pragma unroll
            for (int t = 0; t < 3; t++) { sm[SM::AC + hl * 6 + t] = aA[t]; sm[SM::AC + hl * 6 + 3 + t] = aB[t]; }
            const R *vx = sm + SM::VX;
            R acc = id ? vx[hl < NX ? hl : 0] : R(0);
            if (M::has_a(hl)) {
#pragma unroll
                for (int t = 0; t < 3; t++) acc += aA[t] * vx[ga + t];
            }
            if (M::has_b(hl)) {
#pragma unroll
                for (int t = 0; t < 3; t++) acc += aB[t] * vx[gb + t];
            }
            g = acc + gw * (sm[SM::XU + gsrc] - gt);
        } else {
#pragma unroll
            for (int r = 0; r < NX; r++) w[r] = R(0);
        }
        __syncwarp();
        // ---- Q[r][lane] = a_r . w (+ the lane's diagonal cost term) ------------------------------------------------------------
        R qc[NZ];
#pragma unroll
        for (int r = 0; r < NZ; r++) {
            R acc = M::has_id(r) ? w[r < NX ? r : 0] : R(0);
            if (M::has_a(r)) {
#pragma unroll
                for (int t = 0; t < 3; t++) acc += sm[SM::AC + r * 6 + t] * w[M::ga(r) + t];
            }
            if (M::has_b(r)) {
#pragma unroll
                for (int t = 0; t < 3; t++) acc += sm[SM::AC + r * 6 + 3 + t] * w[M::gb(r) + t];
            }
            qc[r] = acc;
        }
#pragma unroll
        for (int r = 0; r < NZ; r++) if (hl == r) qc[r] += wdiag;
        // ---- eliminate the nu control rows -------------------------------------------------------------------------------------
        R mp[NU], mg[NU], rip[NU];
#pragma unroll
        for (int p = 0; p < NU; p++) {
            const R d = shfl16(qc[NX + p], NX + p);
            if (!(d > R(0))) ok = false;
            const R ri = g_rsqrt(d);
            const R m = (hl == NX + p) ? d * ri : qc[NX + p] * ri;
            const R gp = shfl16(g, NX + p) * ri;   // the gradient column's multiplier
            mp[p] = m; mg[p] = gp; rip[p] = ri;
            if (hl < NZ) sm[SM::LM + p * (NZ + 1) + hl] = m;
            __syncwarp();
            const R *Lm = sm + SM::LM + p * (NZ + 1);
#pragma unroll
            for (int r = 0; r < NZ; r++) if (r < NX || r > NX + p) qc[r] -= Lm[r] * m;
            g -= m * gp;
            dv -= gp * gp;
        }
        if (!__any_sync(0xffffffffu, ok && act)) break;
        const bool st = ok && act;
        // ---- gains: L' z = y; y = this lane's multipliers (feedback column) or the gradient multipliers (feedforward) ---------------
        {
            R z[NU], zg[NU];
#pragma unroll
            for (int p = NU - 1; p >= 0; p--) {
                R s = mp[p], sg = mg[p];
#pragma unroll
                for (int t = p + 1; t < NU; t++) { const R l = sm[SM::LM + p * (NZ + 1) + NX + t]; s -= l * z[t]; sg -= l * zg[t]; }
                z[p] = s * rip[p]; zg[p] = sg * rip[p];   // 1 / L[p][p] = 1 / sqrt(d) is the pivot's rsqrt
            }
            if (st && hl < NX) {
#pragma unroll
                for (int m = 0; m < NU; m++) K[((long long)i * NU + m) * NX + hl] = -z[m];
            }
            if (st && hl < NU) kf[(long long)i * NU + hl] = -pick<NU>(zg, hl);
        }
        // ---- the trailing block is the new value function ----------------------------------------------------------------------
        if (hl < NX) {
#pragma unroll
            for (int r = 0; r < NX; r++) sm[SM::S + r * LD + hl] = qc[r];
            sm[SM::VX + hl] = g;
        }
        __syncwarp();
        if (hl < NX) {
            R *Vn = sm + (vb ? SM::V0 : SM::V1);
#pragma unroll
            for (int r = 0; r < NX; r++) Vn[hl * LD + r] = R(0.5) * (qc[r] + sm[SM::S + hl * LD + r]);
        }
        vb ^= 1;
        __syncwarp();
    }
    *dV1 = dv;
    *knots += nk;
    return ok;
}

#ifndef GDDP_PAIR_MIN_BLOCKS
#define GDDP_PAIR_MIN_BLOCKS (sizeof(R) == 4 ? 4 : 2)
#endif
template <int MODEL, class R> __global__ void __launch_bounds__(128, GDDP_PAIR_MIN_BLOCKS) gddp_pair_kernel(Args<R> A) {
    using D = Dim<MODEL>;
    constexpr int NX = D::NX, NU = D::NU;
    extern __shared__ __align__(16) unsigned char gsm_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5, sub = lane >> 4, hl = lane & 15;
    R *sm = reinterpret_cast<R *>(gsm_raw) + (warp * 2 + sub) * SmemP<MODEL>::TOTAL;
    const Ws<MODEL, R> wl(A.N);
    R *ws = A.ws + (((long long)blockIdx.x * wpb + warp) * 2 + sub) * A.ws_stride;
    const int N = A.N;
    const unsigned FULL = 0xffffffffu;
    while (true) {
        unsigned int pr = 0;
        if (lane == 0) pr = atomicAdd(A.counter, 1u);
        pr = __shfl_sync(FULL, pr, 0);
        if (2ull * pr >= (unsigned long long)A.B) break;
        const bool valid = 2ll * pr + sub < A.B;
        const int b = valid ? (int)(2 * pr + sub) : (int)(2 * pr);   // an odd batch: the last half-warp shadows its partner, stores nothing
        const long long clk0 = clock64();
        R *xb = ws + wl.xb, *xn = ws + wl.xn, *ub = ws + wl.ub, *un = ws + wl.un, *K = ws + wl.K, *kf = ws + wl.kf;
        for (int e = hl; e < N * NU; e += 16) un[e] = A.u_init ? (R)A.u_init[(long long)b * N * NU + e] : A.uh[e % NU];
        __syncwarp();
        R J = rollout2<MODEL, R>(A, b, hl, true, sm, nullptr, nullptr, nullptr, nullptr, false, R(0), xb, un);
        { R *t = ub; ub = un; un = t; }
        __syncwarp();
        R rho = R(0);
        int rtn = 0, iters = A.iter_max;
        bool done = !valid;
        long long sweeps = 0, rollouts = 1, knots = 0;
        for (int it = 0; it < A.iter_max; it++) {
            if (__all_sync(FULL, done)) break;
            // ---- sweeps, until the half's own sweep succeeds ------------------------------------------------------------------------
            R dV1 = R(0);
            bool need = !done;
            while (__any_sync(FULL, need)) {
                if (need) sweeps++;
                R dv = R(0);
                const bool ok = sweep2<MODEL, R>(A, b, hl, need, sm, xb, ub, rho, K, kf, &dv, &knots);
                __syncwarp();
                if (need) {
                    if (ok) { need = false; dV1 = dv; }
                    else {
                        rho = rho * R(4) > R(1e-6) ? rho * R(4) : R(1e-6);
                        if (rho > R(1e10)) { need = false; done = true; rtn = -4; iters = it; }
                    }
                }
            }
            if (!done && -dV1 <= A.tol * (R(1) + g_abs(J))) { done = true; rtn = 1; iters = it; }
            // ---- line search: the halves still searching roll out, the others wait -------------------------------------------------
            bool ls = !done, accepted = false;
            R Jn = R(0), alpha = R(1);
            for (int s = 0; s < 11; s++, alpha *= R(0.5)) {
                if (!__any_sync(FULL, ls)) break;
                if (ls) rollouts++;
                const R Jt = rollout2<MODEL, R>(A, b, hl, ls, sm, xb, ub, K, kf, true, alpha, xn, un);
                __syncwarp();
                if (ls && Jt < J) { accepted = true; ls = false; Jn = Jt; }
            }
            if (!done) {
                if (!accepted) {
                    rho = rho * R(4) > R(1e-6) ? rho * R(4) : R(1e-6);
                    if (rho > R(1e10)) { done = true; rtn = -4; iters = it; }
                } else {
                    const R dJ = J - Jn;
                    J = Jn;
                    { R *t = xb; xb = xn; xn = t; t = ub; ub = un; un = t; }
                    rho = rho / R(4);
                    if (rho < R(1e-9)) rho = R(0);
                    if (dJ <= A.tol * (R(1) + g_abs(J))) { done = true; rtn = 1; iters = it + 1; }
                }
            }
        }
        if (valid) {
            if (hl == 0) {
                A.rtn[b] = rtn; A.iters[b] = iters; A.cost[b] = (double)J;
                if (A.stats) {
                    long long *S = A.stats + (long long)b * 4;
                    S[0] = sweeps; S[1] = rollouts; S[2] = knots; S[3] = clock64() - clk0;
                }
            }
            for (int e = hl; e < (N + 1) * NX; e += 16) A.x[(long long)b * (N + 1) * NX + e] = (double)xb[(long long)(e / NX) * D::XS + e % NX];
            for (int e = hl; e < N * NU; e += 16) A.u[(long long)b * N * NU + e] = (double)ub[e];
        }
        __syncwarp();
    }
}

}  // namespace gddp
#endif
