// se3mpc_soft.cuh -- "soft dynamics" pieces of the Riccati sweeps for the feasibility restoration phase (included by se3mpc_core.cuh
// ahead of the sweeps, namespace se3; see se3mpc_resto.cuh for the restoration phase itself).
//
// Eliminating the slacks n, p of  c(y) - p + n = 0  from the Newton system of the restoration problem leaves the regular KKT system
// with a negative diagonal block, [H J^T; J -D], D = n/z_n + p/z_p > 0.  In the Riccati recursion the value function (P', p') of stage
// k+1 then enters stage k through
//      P~ = P' - P' D^1/2 T^-1 D^1/2 P',   p~ = p' - P' D^1/2 T^-1 D^1/2 p',   T = I + D^1/2 P' D^1/2   (13 x 13, Cholesky per stage)
// and the forward sweep applies  dx+ = (I + D P')^-1 (A dx + B du + chat - D p') = w - D^1/2 T^-1 D^1/2 P' w,  w = A dx + B du + chat - D p'.
// T positive definite for every stage (together with every H_uu) <=> the KKT matrix has the inertia (n, m, 0).
#pragma once

// ---- dense 13 x 13 helpers on shared memory (row stride AS = 14), warp-cooperative --------------------------------------------
// Cholesky of the symmetric matrix T (lower triangle read): L overwrites the strict lower triangle, lr[j] = 1 / L[j][j].
// Returns 0 when a pivot is not safely positive (wrong inertia of the softened system).
SE3_HD int chol13_warp(double *T, double *lr)
{
    for (int j = 0; j < 13; j++) {
        double d = T[j * AS + j];
        for (int k = 0; k < j; k++) d -= T[j * AS + k] * T[j * AS + k];
        if (UNI(!(d > 4.0 * DBL_EPSILON * fabs(T[j * AS + j])) || !(d > 0))) return 0;
        const double r = rsqrt_d(d);
        WARP_FOR(lane) {
            if (lane == j) lr[j] = r;
            if (lane > j && lane < 13) {
                double v = T[lane * AS + j];
                for (int k = 0; k < j; k++) v -= T[lane * AS + k] * T[j * AS + k];
                T[lane * AS + j] = v * r;
            }
        }
        WARP_SYNC();
    }
    return 1;
}
// L y = b and L^T x = y on a register vector
SE3_HD void tri13_fwd(const double *L, const double *lr, double *v)
{
#pragma unroll
    for (int i = 0; i < 13; i++) {
        double a = v[i];
#pragma unroll
        for (int k = 0; k < i; k++) a -= L[i * AS + k] * v[k];
        v[i] = a * lr[i];
    }
}
SE3_HD void tri13_bwd(const double *L, const double *lr, double *v)
{
#pragma unroll
    for (int i = 12; i >= 0; i--) {
        double a = v[i];
#pragma unroll
        for (int k = i + 1; k < 13; k++) a -= L[k * AS + i] * v[k];
        v[i] = a * lr[i];
    }
}

// T = I + D^1/2 P D^1/2 (lower triangle) into Lm, then its Cholesky factor.  Pm: P (row stride AS), sq = D^1/2.
SE3_HD int soft_factor_T(const double *Pm, const double *sq, double *Lm, double *lr)
{
    WARP_FOR(lane) {
        if (lane < 13) {
            for (int j = 0; j <= lane; j++) Lm[lane * AS + j] = sq[lane] * Pm[lane * AS + j] * sq[j] + ((j == lane) ? 1.0 : 0.0);
        }
    }
    WARP_SYNC();
    return chol13_warp(Lm, lr);
}


/*
 * Factor sweep, top of stage k: (P', p') -> (P~, p~).  P' lives in the C fragments fr.pp (tiles (rt, ct)), p' in row 4 of FX^T.
 * Returns 0 when T is not positive definite.
 */
SE3_HD int soften_stage(const Ctx c, const int k, LANE_MREF(FacRegs, fr))
{
    WarpSmem &s = SE3_SMEM(c);
    double *Pm = s.u.soft.Pm, *Lm = s.u.soft.Lm, *Y = s.u.soft.Y, *sq = s.u.soft.sq, *lr = s.u.soft.lr;
    WARP_FOR(lane) {
        const FacRegs &r = LV(fr);
        const int g = lane >> 2, t = lane & 3;
#pragma unroll
        for (int tile = 0; tile < 4; tile++)
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const int i = 8 * (tile >> 1) + g, j = 8 * (tile & 1) + 2 * t + e;
                if (i < 13 && j < 13) Pm[i * AS + j] = r.pp[2 * tile + e];
            }
        if (lane < 13) {
            Pm[lane * AS + 13] = s.Qa[fxt(4, lane)];
            sq[lane] = dsqrt(row(c, R_RD + lane)[k]);
        }
    }
    WARP_SYNC();
    if (!uni(soft_factor_T(Pm, sq, Lm, lr))) return 0;
    // Y = L^-1 D^1/2 [P' | p']: lane j < 14 owns column j
    WARP_FOR(lane) {
        if (lane < 14) {
            double v[13];
#pragma unroll
            for (int i = 0; i < 13; i++) v[i] = sq[i] * Pm[i * AS + lane];
            tri13_fwd(Lm, lr, v);
#pragma unroll
            for (int i = 0; i < 13; i++) Y[i * AS + lane] = v[i];
        }
    }
    WARP_SYNC();
    // [P~ | p~] = [P' | p'] - Y^T Y: lane i < 13 owns row i
    WARP_FOR(lane) {
        if (lane < 13) {
#pragma unroll 1
            for (int j = 0; j < 14; j++) {
                double a = Pm[lane * AS + j];
#pragma unroll
                for (int l = 0; l < 13; l++) a -= Y[l * AS + lane] * Y[l * AS + j];
                Pm[lane * AS + j] = a;
            }
        }
    }
    WARP_SYNC();
    WARP_FOR(lane) {
        FacRegs &r = LV(fr);
        const int g = lane >> 2, t = lane & 3;
#pragma unroll
        for (int tile = 0; tile < 4; tile++)
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const int i = 8 * (tile >> 1) + g, j = 8 * (tile & 1) + 2 * t + e;
                r.pp[2 * tile + e] = (i < 13 && j < 13) ? Pm[i * AS + j] : 0.0;
            }
        if (lane < 13) s.Qa[fxt(4, lane)] = Pm[lane * AS + 13];
    }
    WARP_SYNC();
    return 1;
}

/*
 * Solve-only backward sweep, stage k after tv = P' c + p':  tv <- tv - P' D^1/2 T^-1 D^1/2 tv  (= P~ c + p~).
 * Pr: the stored P_{k+1} (row stride AS); tv: the 13-vector in shared memory.
 */
SE3_HD void soften_vector(const Ctx c, const int k, const double *Pr, double *tv)
{
    WarpSmem &s = SE3_SMEM(c);
    double *Lm = s.u.soft.Lm, *sq = s.u.soft.sq, *lr = s.u.soft.lr, *b = s.u.soft.b;
    WARP_FOR(lane) { if (lane < 13) sq[lane] = dsqrt(row(c, R_RD + lane)[k]); }
    WARP_SYNC();
    soft_factor_T(Pr, sq, Lm, lr);          // (positive definite: the factor sweep accepted this stage)
    WARP_FOR(lane) {
        if (lane < 13) {                    // every row lane solves the same 13-vector system (a serial chain either way)
            double v[13];
#pragma unroll
            for (int i = 0; i < 13; i++) v[i] = sq[i] * tv[i];
            tri13_fwd(Lm, lr, v);
            tri13_bwd(Lm, lr, v);
            double a = tv[lane];
#pragma unroll
            for (int l = 0; l < 13; l++) a -= Pr[lane * AS + l] * sq[l] * v[l];
            b[lane] = a;
        }
    }
    WARP_SYNC();
    WARP_FOR(lane) { if (lane < 13) tv[lane] = b[lane]; }
    WARP_SYNC();
}

/*
 * Forward sweep, stage k after v = A dx + B du + chat:  dx+ = (I + D P')^-1 (v - D p'), P' = P_{k+1}, p' = p_{k+1} from the stored factors.
 * xin/xout: v (13) in shared memory, overwritten by dx+.  Uses the (otherwise idle) Qa block as scratch; riccati_forward clears it.
 */
SE3_HD void soften_forward(const Ctx c, const int k, double *x)
{
    WarpSmem &s = SE3_SMEM(c);
    double *Pm = s.Qa, *Lm = s.Qa + 13 * AS, *sq = s.Qa + 26 * AS, *lr = sq + 16, *w = lr + 16, *b = w + 16;
    WARP_FOR(lane) {
        if (lane < 13) {
            const GP F = fac(c, k + 1);
#pragma unroll 1
            for (int j = 0; j < 13; j++) Pm[lane * AS + j] = F[FAC_P + fac_p(lane, j)];
            const double D = row(c, R_RD + lane)[k];
            sq[lane] = dsqrt(D);
            w[lane] = x[lane] - D * (double)F[FAC_PX + lane];
        }
    }
    WARP_SYNC();
    soft_factor_T(Pm, sq, Lm, lr);
    WARP_FOR(lane) {
        if (lane < 13) {
            double a = 0.0;
#pragma unroll
            for (int l = 0; l < 13; l++) a += Pm[lane * AS + l] * w[l];
            b[lane] = sq[lane] * a;
        }
    }
    WARP_SYNC();
    WARP_FOR(lane) {
        if (lane < 13) {
            double v[13];
#pragma unroll
            for (int i = 0; i < 13; i++) v[i] = b[i];
            tri13_fwd(Lm, lr, v);
            tri13_bwd(Lm, lr, v);
            x[lane] = w[lane] - sq[lane] * v[lane];
        }
    }
    WARP_SYNC();
}
