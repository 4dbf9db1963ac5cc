// se3mpc_resto.cuh -- feasibility restoration phase of the interior-point solver (included by se3mpc_core.cuh, namespace se3).
//
// IPOPT enters its restoration phase when the filter line search of the regular algorithm runs below alpha_min; the reference
// inherits that behaviour by calling nlpsol('ipopt') without options (quad_OC.py:170).  This is IPOPT's MinC_1NrmRestorationPhase
// (Waechter & Biegler 2006, section 3.3; IpRestoMinC_1Nrm.cpp, IpRestoIpoptNLP.cpp, IpRestoIterateInitializer.cpp,
// IpRestoConvCheck.cpp) for the stage-structured NLP of quad_OC.py:104-212.  With y = (x_1..x_N, u_0..u_{N-1}) and the reference
// point y_R = the iterate at entry it solves
//      min  rho sum(n + p) + eta(mu)/2 ||D_R (y - y_R)||^2     s.t.  c(y) - p + n = 0,  n, p >= 0,  y_L <= y <= y_U
// (rho = 1000, eta(mu) = sqrt(mu), D_R = diag(1 / max(1, |y_R|))) with the SAME interior-point algorithm -- monotone mu starting at
// max(mu, ||c||_inf), filter line search with second-order correction, inertia correction -- until the y part is acceptable to the
// ORIGINAL filter and has reduced the ORIGINAL constraint violation to kappa_resto = 0.9 of its value at entry.
//
// The slacks are eliminated from the Newton system analytically (IpAugRestoSystemSolver): what is left is the regular KKT system
// with a negative diagonal block, [H J^T; J -D], D = n/z_n + p/z_p: a Riccati recursion with "soft" dynamics.  (P', p') of stage
// k+1 enter stage k through  P~ = P' - P' D^1/2 T^-1 D^1/2 P',  p~ = p' - P' D^1/2 T^-1 D^1/2 p',  T = I + D^1/2 P' D^1/2  (a 13 x 13
// Cholesky per stage, soften_stage); the forward sweep applies dx+ = (I + D P')^-1 (A dx + B du + chat - D p').  Everything else --
// the DMMA products, the H_uu Cholesky, the stored factors, the stage-parallel record preparation (mode RESTO: H = eta D_R^2 +
// sum lam hess(c) + Sigma, no control-rate coupling) -- is the regular machinery of se3mpc_core.cuh.
//
// Problems that never fail their line search never execute any of this; the rows R_RN .. R_RHS of the slot are touched only here.
#pragma once

constexpr double RESTO_RHO = 1000.0;

// ---- stage-parallel phases of the restoration problem (lane <-> stages {lane, lane + 32}) --------------------------------------
// slack steps of one constraint component: dn = -(rn + dl) / sn, dp = -(rp - dl) / sp with dl = lam+ - lam (the eliminated rows of
// the Newton system), and the multiplier steps that belong to them
struct SlackStep { double dn, dp, dzn, dzp; };
SE3_HD SlackStep slack_step(double n, double p, double zn, double zp, double lam, double lamp, double mu)
{
    const double sn = ddiv(zn, n), sp = ddiv(zp, p);
    const double rn = RESTO_RHO + lam - ddiv(mu, n), rp = RESTO_RHO - lam - ddiv(mu, p);
    const double dl = lamp - lam;
    SlackStep r;
    r.dn = -ddiv(rn + dl, sn); r.dp = -ddiv(rp - dl, sp);
    r.dzn = ddiv(mu, n) - zn - sn * r.dn;
    r.dzp = ddiv(mu, p) - zp - sp * r.dp;
    return r;
}

// entry: reference point, multipliers at entry, initial slacks (IpRestoIterateInitializer), bound multipliers capped at rho
SE3_BIG void resto_begin(const Ctx c, double mu)
{
    const Prob &P = SE3_SMEM(c).P;
    const int N = uni(P.N);
    WARP_FOR(lane) {
#pragma unroll 1
        for (int s = lane; s <= N; s += 32) {
            for (int i = 0; i < 13; i++) row(c, R_RXR + i)[s] = row(c, R_X + i)[s];
            for (int b = 0; b < NB; b++) {
                const double zl = row(c, R_ZL + b)[s], zu = row(c, R_ZU + b)[s];
                row(c, R_RBZ + b)[s] = zl; row(c, R_RBZ + NB + b)[s] = zu;
                row(c, R_ZL + b)[s] = fmin(RESTO_RHO, zl); row(c, R_ZU + b)[s] = fmin(RESTO_RHO, zu);
            }
            if (s < N) {
                for (int j = 0; j < 4; j++) row(c, R_RUR + j)[s] = row(c, R_U + j)[s];
                for (int i = 0; i < 13; i++) {
                    row(c, R_RBLAM + i)[s] = row(c, R_LAM + i)[s];
                    const double cv = row(c, R_C + i)[s], a = ddiv(mu - RESTO_RHO * cv, 2.0 * RESTO_RHO);
                    const double n = a + dsqrt(a * a + ddiv(mu * cv, 2.0 * RESTO_RHO)), p = cv + n;
                    row(c, R_RN + i)[s] = n; row(c, R_RP + i)[s] = p;
                    row(c, R_RZN + i)[s] = ddiv(mu, n); row(c, R_RZP + i)[s] = ddiv(mu, p);
                    row(c, R_RLAM + i)[s] = 0.0;
                }
            }
        }
    }
    WARP_SYNC();
}

// D = n/z_n + p/z_p and the reduced constraint residual chat = h - rn/sn + rp/sp + D lam, h = c - p + n (rows R_RD, R_RCH)
SE3_BIG void resto_rows(const Ctx c, double mu)
{
    const Prob &P = SE3_SMEM(c).P;
    const int N = uni(P.N);
    WARP_FOR(lane) {
#pragma unroll 1
        for (int s = lane; s < N; s += 32) {
#pragma unroll 1
            for (int i = 0; i < 13; i++) {
                const double n = row(c, R_RN + i)[s], p = row(c, R_RP + i)[s], zn = row(c, R_RZN + i)[s], zp = row(c, R_RZP + i)[s];
                const double lam = row(c, R_RLAM + i)[s];
                const double sn = ddiv(zn, n), sp = ddiv(zp, p);
                const double rn = RESTO_RHO + lam - ddiv(mu, n), rp = RESTO_RHO - lam - ddiv(mu, p);
                const double D = ddiv(1.0, sn) + ddiv(1.0, sp), h = row(c, R_C + i)[s] - p + n;
                row(c, R_RD + i)[s] = D;
                row(c, R_RCH + i)[s] = h - ddiv(rn, sn) + ddiv(rp, sp) + D * lam;
            }
        }
    }
    WARP_SYNC();
}

// scaled optimality error of the restoration problem (same E_mu definition, all multipliers counted): returns E_0 and E_mu pieces
struct RKktOut { double dinf, cinf, cmp0, cmpmu, lam1, z1; };
SE3_HD RKktOut uni(RKktOut k) { k.dinf = uni(k.dinf); k.cinf = uni(k.cinf); k.cmp0 = uni(k.cmp0); k.cmpmu = uni(k.cmpmu); k.lam1 = uni(k.lam1); k.z1 = uni(k.z1); return k; }

SE3_BIG RKktOut resto_kkt(const Ctx c, double mu, double eta)
{
    const Prob &P = SE3_SMEM(c).P;
    const int N = uni(P.N);
    double out[6];
    warp_reduce<6, 0xFu>([&](int lane, double *acc) {
        for (int i = 0; i < 6; i++) acc[i] = 0;
#pragma unroll 1
        for (int s = lane; s <= N; s += 32) {
            double x[13], u[4], l[13], atl[13], btl[4];
            ld13(c, R_X, s, x);
            if (s < N) {
                ld4(c, R_U, s, u); ld13(c, R_RLAM, s, l);
                dyn_vjp(P, x, u, l, atl, btl);
                for (int j = 0; j < 4; j++) {
                    const double ur = row(c, R_RUR + j)[s], d = ddiv(1.0, fmax(1.0, fabs(ur)));
                    const double v = eta * d * d * (u[j] - ur) + btl[j] - row(c, R_ZLU + j)[s] + row(c, R_ZUU + j)[s];
                    acc[0] = dmax(acc[0], fabs(v));
                }
#pragma unroll 1
                for (int i = 0; i < 13; i++) {
                    const double n = row(c, R_RN + i)[s], p = row(c, R_RP + i)[s], zn = row(c, R_RZN + i)[s], zp = row(c, R_RZP + i)[s];
                    acc[1] = dmax(acc[1], fabs(row(c, R_C + i)[s] - p + n));
                    acc[4] += fabs(l[i]);
                    acc[0] = dmax(acc[0], dmax(fabs(RESTO_RHO + l[i] - zn), fabs(RESTO_RHO - l[i] - zp)));
                    acc[2] = dmax(acc[2], dmax(fabs(n * zn), fabs(p * zp)));
                    acc[3] = dmax(acc[3], dmax(fabs(n * zn - mu), fabs(p * zp - mu)));
                    acc[5] += fabs(zn) + fabs(zp);
                }
            } else
                for (int i = 0; i < 13; i++) atl[i] = 0.0;
    SE3_BL_KKT
            for (int b = 0; b < NB; b++) {
                const bool isw = b < 3;
                if (isw ? (s < 1) : (s >= N)) continue;
                const double v = row(c, R_BV + b)[s], zl = row(c, R_ZL + b)[s], zu = row(c, R_ZU + b)[s];
                const double sl = v - (isw ? P.wL : P.uL), su = (isw ? P.wU : P.uU) - v;
                acc[2] = dmax(acc[2], dmax(fabs(sl * zl), fabs(su * zu)));
                acc[3] = dmax(acc[3], dmax(fabs(sl * zl - mu), fabs(su * zu - mu)));
                acc[5] += fabs(zl) + fabs(zu);
            }
            if (s >= 1) {
                double lp[13];
                ld13(c, R_RLAM, s - 1, lp);
                for (int i = 0; i < 13; i++) {
                    const double xr = row(c, R_RXR + i)[s], d = ddiv(1.0, fmax(1.0, fabs(xr)));
                    double v = eta * d * d * (x[i] - xr) + atl[i] - lp[i];
                    if (i >= 10) v += -row(c, R_ZLW + i - 10)[s] + row(c, R_ZUW + i - 10)[s];
                    acc[0] = dmax(acc[0], fabs(v));
                }
            }
        }
    }, out);
    RKktOut k;
    k.dinf = out[0]; k.cinf = out[1]; k.cmp0 = out[2]; k.cmpmu = out[3]; k.lam1 = out[4]; k.z1 = out[5];
    return k;
}
SE3_HD double resto_kkt_scaled(const int N, const RKktOut &k, double cmp)
{
    const int m = 13 * N, nz = 2 * (4 * N + 3 * N) + 2 * 13 * N;
    const double sd = fmax(100.0, (k.lam1 + k.z1) / (double)(m + nz)) / 100.0;
    const double sc = fmax(100.0, k.z1 / (double)nz) / 100.0;
    return fmax(k.dinf / sd, fmax(k.cinf, cmp / sc));
}

// fraction to the boundary (primal incl. the slacks, dual incl. z_n, z_p), grad(phi_R)^T d, finiteness of the direction
SE3_BIG DirOut resto_direction(const Ctx c, double mu, double eta, double tau, int rdx, int rlamp)
{
    const Prob &P = SE3_SMEM(c).P;
    const int N = uni(P.N);
    double out[4];
    warp_reduce<4, 0xBu>([&](int lane, double *acc) {
        acc[0] = -1.0; acc[1] = -1.0; acc[2] = 0.0; acc[3] = 0.0;
#pragma unroll 1
        for (int s = lane; s <= N; s += 32) {
#pragma unroll 1
            for (int i = 0; i < 10; i++) {
                const double d = row(c, rdx + i)[s];
                if (!(fabs(d) < 1e300)) acc[3] = 1.0;
                if (s >= 1) {
                    const double xr = row(c, R_RXR + i)[s], dr = ddiv(1.0, fmax(1.0, fabs(xr)));
                    acc[2] += eta * dr * dr * (row(c, R_X + i)[s] - xr) * d;
                }
            }
    SE3_BL_DIR
            for (int b = 0; b < NB; b++) {
                const bool isw = b < 3;
                if (isw ? (s < 1) : (s >= N)) continue;
                const double v = row(c, R_BV + b)[s], d = row(c, rdx + 10 + b)[s];
                if (!(fabs(d) < 1e300)) acc[3] = 1.0;
                const double sl = v - (isw ? P.wL : P.uL), su = (isw ? P.wU : P.uU) - v;
                const double yr = row(c, R_RXR + 10 + b)[s], dr = ddiv(1.0, fmax(1.0, fabs(yr)));      // (R_RUR = R_RXR + 13: contiguous like R_BV)
                const double g = eta * dr * dr * (v - yr) - ddiv(mu, sl) + ddiv(mu, su);
                acc[2] += g * d;
                if (d < 0) acc[0] = dmax(acc[0], -ddiv(-tau * sl, d));
                else if (d > 0) acc[0] = dmax(acc[0], -ddiv(tau * su, d));
                const double zl = row(c, R_ZL + b)[s], zu = row(c, R_ZU + b)[s];
                const double dzl = ddiv(mu, sl) - zl - ddiv(zl, sl) * d, dzu = ddiv(mu, su) - zu + ddiv(zu, su) * d;
                if (dzl < 0) acc[1] = dmax(acc[1], -ddiv(-tau * zl, dzl));
                if (dzu < 0) acc[1] = dmax(acc[1], -ddiv(-tau * zu, dzu));
            }
            if (s < N) {
#pragma unroll 1
                for (int i = 0; i < 13; i++) {
                    const double n = row(c, R_RN + i)[s], p = row(c, R_RP + i)[s], zn = row(c, R_RZN + i)[s], zp = row(c, R_RZP + i)[s];
                    const SlackStep st = slack_step(n, p, zn, zp, row(c, R_RLAM + i)[s], row(c, rlamp + i)[s], mu);
                    if (!(fabs(st.dn) < 1e300) || !(fabs(st.dp) < 1e300)) acc[3] = 1.0;
                    acc[2] += (RESTO_RHO - ddiv(mu, n)) * st.dn + (RESTO_RHO - ddiv(mu, p)) * st.dp;
                    if (st.dn < 0) acc[0] = dmax(acc[0], -ddiv(-tau * n, st.dn));
                    if (st.dp < 0) acc[0] = dmax(acc[0], -ddiv(-tau * p, st.dp));
                    if (st.dzn < 0) acc[1] = dmax(acc[1], -ddiv(-tau * zn, st.dzn));
                    if (st.dzp < 0) acc[1] = dmax(acc[1], -ddiv(-tau * zp, st.dzp));
                }
            }
        }
    }, out);
    DirOut r;
    r.alpha_p = -out[0]; r.alpha_z = -out[1]; r.gphid = out[2]; r.finite = (out[3] == 0.0);
    return r;
}

// restoration objective, theta_R = ||c - p + n||_1 and the barrier sum at (y + alpha dy, n + alpha dn, p + alpha dp); optionally the
// second-order-correction residual h_soc (store_soc 1: alpha h + h(trial), 2: alpha h_soc + h(trial)) into rows R_RHS
SE3_BIG TrialOut resto_trial(const Ctx c, double alpha, double mu, double eta, int rdx, int rdu, int rlamp, int store_soc, double soc_alpha)
{
    const Prob &P = SE3_SMEM(c).P;
    const int N = uni(P.N);
    double out[4];
    warp_reduce<4, 8u>([&](int lane, double *acc) {
        acc[0] = acc[1] = acc[2] = acc[3] = 0;
#pragma unroll 1
        for (int s = lane; s <= N; s += 32) {
            double x[13], d[13];
            ld13(c, R_X, s, x); ld13(c, rdx, s, d);
#pragma unroll
            for (int i = 0; i < 13; i++) x[i] += alpha * d[i];
            if (s >= 1) {
                double q = 0.0;
                for (int i = 0; i < 13; i++) {
                    const double xr = row(c, R_RXR + i)[s], dr = ddiv(1.0, fmax(1.0, fabs(xr))), e = dr * (x[i] - xr);
                    q += e * e;
                }
                acc[0] += 0.5 * eta * q;
            }
    SE3_BL_TRIAL
            for (int b = 0; b < NB; b++) {
                const bool isw = b < 3;
                if (isw ? (s < 1) : (s >= N)) continue;
                const double v = row(c, R_BV + b)[s] + alpha * row(c, rdx + 10 + b)[s];
                if (!isw) {
                    const double ur = row(c, R_RXR + 10 + b)[s], dr = ddiv(1.0, fmax(1.0, fabs(ur))), e = dr * (v - ur);
                    acc[0] += 0.5 * eta * e * e;
                }
                const double sl = v - (isw ? P.wL : P.uL), su = (isw ? P.wU : P.uU) - v;
                if (!(sl > 0) || !(su > 0)) acc[3] = 1.0;
                else {
                    double ls = 0.0;
#pragma unroll 1
                    for (int q = 0; q < 2; q++) ls += log(q ? su : sl);
                    acc[2] += ls;
                }
            }
            if (s < N) {
                double u[4], xn[13], cc[13];
                ld4(c, R_U, s, u); ld4(c, rdu, s, cc);
#pragma unroll
                for (int j = 0; j < 4; j++) u[j] += alpha * cc[j];
                ld13(c, R_X, s + 1, xn); ld13(c, rdx, s + 1, d);
#pragma unroll
                for (int i = 0; i < 13; i++) xn[i] += alpha * d[i];
                dyn_residual(P, x, u, xn, cc);
#pragma unroll 1
                for (int i = 0; i < 13; i++) {
                    const double n = row(c, R_RN + i)[s], p = row(c, R_RP + i)[s];
                    const SlackStep st = slack_step(n, p, row(c, R_RZN + i)[s], row(c, R_RZP + i)[s], row(c, R_RLAM + i)[s], row(c, rlamp + i)[s], mu);
                    const double nt = n + alpha * st.dn, pt = p + alpha * st.dp;
                    acc[0] += RESTO_RHO * (nt + pt);
                    const double ht = cc[i] - pt + nt;
                    acc[1] += fabs(ht);
                    if (!(nt > 0) || !(pt > 0)) acc[3] = 1.0;
                    else {
                        double ls = 0.0;
#pragma unroll 1
                        for (int q = 0; q < 2; q++) ls += log(q ? pt : nt);
                        acc[2] += ls;
                    }
                    if (store_soc) {
                        const double hprev = (store_soc == 1) ? (row(c, R_C + i)[s] - p + n) : (double)row(c, R_RHS + i)[s];
                        row(c, R_RHS + i)[s] = soc_alpha * hprev + ht;
                    }
                }
            }
        }
    }, out);
    WARP_SYNC();
    TrialOut t;
    t.f = out[0]; t.theta = out[1]; t.logsum = out[2]; t.feasible = (out[3] == 0.0);
    return t;
}

// rows R_CS = chat + h_soc - h: the reduced residual of the corrected system (only the constraint residual changes)
SE3_BIG void resto_soc_rows(const Ctx c)
{
    const Prob &P = SE3_SMEM(c).P;
    const int N = uni(P.N);
    WARP_FOR(lane) {
#pragma unroll 1
        for (int s = lane; s < N; s += 32)
#pragma unroll 1
            for (int i = 0; i < 13; i++) {
                const double h = row(c, R_C + i)[s] - row(c, R_RP + i)[s] + row(c, R_RN + i)[s];
                row(c, R_CS + i)[s] = row(c, R_RCH + i)[s] + row(c, R_RHS + i)[s] - h;
            }
    }
    WARP_SYNC();
}

// accept the step: y, slacks, lam with alpha; all bound multipliers with alpha_z and the kappa_sigma safeguard
SE3_BIG void resto_apply(const Ctx c, double alpha, double alpha_z, double mu, int rdx, int rlamp)
{
    const Prob &P = SE3_SMEM(c).P;
    const int N = uni(P.N);
    const double ks = 1e10;
    WARP_FOR(lane) {
#pragma unroll 1
        for (int s = lane; s <= N; s += 32) {
    SE3_BL_APPLY
            for (int b = 0; b < NB; b++) {
                const bool isw = b < 3;
                if (isw ? (s < 1) : (s >= N)) continue;
                const double v = row(c, R_BV + b)[s], d = row(c, rdx + 10 + b)[s];
                const double lo = isw ? P.wL : P.uL, hi = isw ? P.wU : P.uU;
                const double sl = v - lo, su = hi - v;
                const double zl = row(c, R_ZL + b)[s], zu = row(c, R_ZU + b)[s];
                const double dzl = ddiv(mu, sl) - zl - ddiv(zl, sl) * d, dzu = ddiv(mu, su) - zu + ddiv(zu, su) * d;
                const double vn = v + alpha * d;
                const double sln = vn - lo, sun = hi - vn;
                if (!isw) row(c, R_BV + b)[s] = vn;
                row(c, R_ZL + b)[s] = fmax(fmin(zl + alpha_z * dzl, ddiv(ks * mu, sln)), ddiv(mu, ks * sln));
                row(c, R_ZU + b)[s] = fmax(fmin(zu + alpha_z * dzu, ddiv(ks * mu, sun)), ddiv(mu, ks * sun));
            }
            if (s >= 1) {
#pragma unroll
                for (int i = 0; i < 13; i++) row(c, R_X + i)[s] += alpha * row(c, rdx + i)[s];
            }
            if (s < N) {
#pragma unroll 1
                for (int i = 0; i < 13; i++) {
                    const double n = row(c, R_RN + i)[s], p = row(c, R_RP + i)[s], zn = row(c, R_RZN + i)[s], zp = row(c, R_RZP + i)[s];
                    const double lam = row(c, R_RLAM + i)[s], lamp = row(c, rlamp + i)[s];
                    const SlackStep st = slack_step(n, p, zn, zp, lam, lamp, mu);
                    const double nn = n + alpha * st.dn, pn = p + alpha * st.dp;
                    row(c, R_RN + i)[s] = nn; row(c, R_RP + i)[s] = pn;
                    row(c, R_RLAM + i)[s] = lam + alpha * (lamp - lam);
                    row(c, R_RZN + i)[s] = fmax(fmin(zn + alpha_z * st.dzn, ddiv(ks * mu, nn)), ddiv(mu, ks * nn));
                    row(c, R_RZP + i)[s] = fmax(fmin(zp + alpha_z * st.dzp, ddiv(ks * mu, pn)), ddiv(mu, ks * pn));
                }
            }
        }
    }
    WARP_SYNC();
}

// Back to the regular algorithm (MinC_1NrmRestorationPhase::PerformRestoration).  Success: equality multipliers are reset to zero
// (constr_mult_reset_threshold = 0); bound multipliers take the primal-dual step that belongs to the primal move from the entry point,
// dz = (mu - z (s_new - s_old)) / s_old - z, with the fraction-to-the-boundary rule, and are all reset to 1 when any of them exceeds
// bound_mult_reset_threshold = 1000.  Failure: the iterate of the regular algorithm is put back.
SE3_BIG void resto_finish(const Ctx c, int success_, double mu_orig, double tau_orig)
{
    const Prob &P = SE3_SMEM(c).P;
    const int N = uni(P.N);
    const int success = uni(success_);          // (arguments of a non-inlined function are not provably warp-uniform)
    if (!success) {
        WARP_FOR(lane) {
#pragma unroll 1
            for (int s = lane; s <= N; s += 32) {
                for (int i = 0; i < 13; i++) row(c, R_X + i)[s] = row(c, R_RXR + i)[s];
                for (int b = 0; b < NB; b++) { row(c, R_ZL + b)[s] = row(c, R_RBZ + b)[s]; row(c, R_ZU + b)[s] = row(c, R_RBZ + NB + b)[s]; }
                if (s < N) {
                    for (int j = 0; j < 4; j++) row(c, R_U + j)[s] = row(c, R_RUR + j)[s];
                    for (int i = 0; i < 13; i++) row(c, R_LAM + i)[s] = row(c, R_RBLAM + i)[s];
                }
            }
        }
        WARP_SYNC();
        return;
    }
    // pass 1: the largest step of the bound multipliers that keeps them positive (fraction to the boundary with tau of the regular problem)
    double o1[1];
    warp_reduce<1, 1u>([&](int lane, double *acc) {
        acc[0] = -1.0;
#pragma unroll 1
        for (int s = lane; s <= N; s += 32) {
#pragma unroll 1
            for (int b = 0; b < NB; b++) {
                const bool isw = b < 3;
                if (isw ? (s < 1) : (s >= N)) continue;
                const double lo = isw ? P.wL : P.uL, hi = isw ? P.wU : P.uU;
                const double vo = row(c, R_RXR + 10 + b)[s], vn = row(c, R_BV + b)[s];
                const double zl = row(c, R_RBZ + b)[s], zu = row(c, R_RBZ + NB + b)[s];
                const double dl = ddiv(mu_orig - zl * ((vn - lo) - (vo - lo)), vo - lo) - zl, dq = ddiv(mu_orig - zu * ((hi - vn) - (hi - vo)), hi - vo) - zu;
                if (dl < 0) acc[0] = dmax(acc[0], -ddiv(-tau_orig * zl, dl));
                if (dq < 0) acc[0] = dmax(acc[0], -ddiv(-tau_orig * zu, dq));
            }
        }
    }, o1);
    const double amax = -o1[0];
    double o2[1];
    warp_reduce<1, 1u>([&](int lane, double *acc) {
        acc[0] = 0.0;
#pragma unroll 1
        for (int s = lane; s <= N; s += 32) {
#pragma unroll 1
            for (int b = 0; b < NB; b++) {
                const bool isw = b < 3;
                if (isw ? (s < 1) : (s >= N)) continue;
                const double lo = isw ? P.wL : P.uL, hi = isw ? P.wU : P.uU;
                const double vo = row(c, R_RXR + 10 + b)[s], vn = row(c, R_BV + b)[s];
                const double zl = row(c, R_RBZ + b)[s], zu = row(c, R_RBZ + NB + b)[s];
                const double dl = ddiv(mu_orig - zl * ((vn - lo) - (vo - lo)), vo - lo) - zl, dq = ddiv(mu_orig - zu * ((hi - vn) - (hi - vo)), hi - vo) - zu;
                const double zln = zl + amax * dl, zun = zu + amax * dq;
                row(c, R_ZL + b)[s] = zln; row(c, R_ZU + b)[s] = zun;
                acc[0] = dmax(acc[0], dmax(zln, zun));
            }
            if (s < N)
                for (int i = 0; i < 13; i++) row(c, R_LAM + i)[s] = 0.0;
        }
    }, o2);
    WARP_SYNC();
    if (UNI(o2[0] > 1e3)) {
        WARP_FOR(lane) {
#pragma unroll 1
            for (int s = lane; s <= N; s += 32)
                for (int b = 0; b < NB; b++) {
                    const bool isw = b < 3;
                    if (isw ? (s < 1) : (s >= N)) continue;
                    row(c, R_ZL + b)[s] = 1.0; row(c, R_ZU + b)[s] = 1.0;
                }
        }
        WARP_SYNC();
    }
}

// Watchdog (ipm_solve): iterate (rows R_X .. R_ZU) and search direction (R_DX .. R_LAMP) of the point where the watchdog starts are kept
// in the restoration rows -- the two procedures never hold state at the same time (the watchdog is stopped before the restoration
// phase is entered).  restore = 0: save, 1: put back.
constexpr int WD_ROWS = R_SDX;          // rows [0, R_SDX): x, u, lam, zL, zU, dx, du, lam+
static_assert(R_ALL - R_RN >= WD_ROWS, "the watchdog copy lives in the restoration rows");
SE3_BIG void watchdog_rows(const Ctx c, int restore_)
{
    const int restore = uni(restore_);
    WARP_FOR(lane) {
#pragma unroll 1
        for (int r = 0; r < WD_ROWS; r++) {
            double *a = row(c, r), *b = row(c, R_RN + r);
            if (restore) { a[lane] = b[lane]; a[lane + 32] = b[lane + 32]; }
            else { b[lane] = a[lane]; b[lane + 32] = a[lane + 32]; }
        }
    }
    WARP_SYNC();
}

/*
 * The restoration phase.  The caller has added the current point to the original filter (IpFilterLSAcceptor::PrepareRestoPhaseStart).
 * ok = 1 when a point acceptable to the original filter was found (iterate and multipliers updated, rows R_C and theta, f, logsum of
 * the new point set), 0 when the restoration phase itself fails (locally infeasible / its own line search fails /
 * iteration limit): the iterate is then left where the regular algorithm stopped.
 */
struct RestoOut { int ok, iter, nls; double theta, f, logsum; };
SE3_HD RestoOut uni(RestoOut r) { r.ok = uni(r.ok); r.iter = uni(r.iter); r.nls = uni(r.nls); r.theta = uni(r.theta); r.f = uni(r.f); r.logsum = uni(r.logsum); return r; }

#ifndef SE3_RESTO_INLINE
#define SE3_RESTO_INLINE 0
#endif
#if SE3_RESTO_INLINE
SE3_HD
#else
SE3_BIG           // out of line: the interior-point loop of the kernel body stays compact (single-warp latency: profiles/README.md)
#endif
RestoOut restoration_phase(const Ctx c, double tol_, int max_iter_, int max_soc_, double mu_orig_, double tau_orig_, int nfilt_, double theta_max_orig_,
                           double theta_orig0_, double phi_orig0_, double cinf0_, int iter_in_, int nls_in_)
{
    // (arguments of a non-inlined function are not provably warp-uniform for ptxas: see uni())
    const double mu_orig = uni(mu_orig_), tau_orig = uni(tau_orig_), theta_max_orig = uni(theta_max_orig_), theta_orig0 = uni(theta_orig0_);
    const double phi_orig0 = uni(phi_orig0_), cinf0 = uni(cinf0_);
    const int nfilt = uni(nfilt_), iter_in = uni(iter_in_), nls_in = uni(nls_in_), max_iter = uni(max_iter_), max_soc = uni(max_soc_);
    const double tol = uni(tol_);
    int iter_v = iter_in, nls_v = nls_in;
    int *iter = &iter_v, *nls = &nls_v;
    double theta_o = 0.0, f_o = 0.0, logsum_o = 0.0;
    double *theta_out = &theta_o, *f_out = &f_o, *logsum_out = &logsum_o;
    WarpSmem &sm = SE3_SMEM(c);
    const int N = uni(sm.P.N);
    const double sf = uni(sm.P.sf);
    const double gamma_theta = 1e-5, gamma_phi = 1e-8, s_phi = 2.3, s_theta = 1.1, eta_phi = 1e-8, kappa_soc = 0.99, kappa_resto = 0.9;
    double mu = fmax(mu_orig, cinf0), tau = fmax(0.99, 1.0 - mu);
    resto_begin(c, mu);
    const double mu_min = tol / 11.0;
    int nrf = 0, success = 0, first = 1;
    double delta_last = 0.0;
    double rtheta_max = 0.0, rtheta_min = 0.0;

    for (;; (*iter)++) {
        // Loop-carried scalars are warp-uniform by construction, but ptxas cannot prove it across the back edge of THIS loop; one
        // possibly-divergent value here and every phase of the solver is compiled divergence-safe (se3mpc_core.cuh, uni()).  A dozen
        // REDUX per restoration iteration keep the throughput path clean (tests/test_codegen.py).
        mu = uni(mu); tau = uni(tau); delta_last = uni(delta_last); nrf = uni(nrf); rtheta_max = uni(rtheta_max); rtheta_min = uni(rtheta_min);
        first = uni(first); iter_v = uni(iter_v); nls_v = uni(nls_v);
        double eta = dsqrt(mu);
        // 1. progress for the ORIGINAL problem (IpRestoConvCheck): not in the first restoration iteration
        double th_o, ci_o;
        eval_constraints(c, &th_o, &ci_o);
        th_o = uni(th_o);
        if (UNI(!first)) {          // (a loop-carried flag is not provably warp-uniform for ptxas: see uni())
            const TrialOut t0 = uni(eval_trial(c, 0.0, R_DX, R_DU, 0, 0.0));
            const double ph_o = sf * t0.f - mu_orig * t0.logsum;
            int ok = t0.feasible && (fabs(th_o) <= 1e300) && (fabs(ph_o) <= 1e300) && th_o <= kappa_resto * theta_orig0 && th_o <= theta_max_orig;
            for (int e = 0; e < nfilt; e++)          // (no early exit: `ok` is not provably warp-uniform before the REDUX below)
                ok = ok && (compare_le(th_o, sm.filt_theta[e], sm.filt_theta[e]) || compare_le(ph_o, sm.filt_phi[e], sm.filt_phi[e]));
            ok = ok && (compare_le(th_o, (1 - gamma_theta) * theta_orig0, theta_orig0) || compare_le(ph_o - phi_orig0, -gamma_phi * theta_orig0, phi_orig0));
            if (uni(ok)) { success = 1; *theta_out = th_o; *f_out = t0.f; *logsum_out = t0.logsum; break; }
        }
        if (UNI(first)) {
            // theta of the restoration problem at its starting point (c - p + n = 0 by construction, up to rounding)
            const TrialOut tr0 = uni(resto_trial(c, 0.0, mu, eta, R_DX, R_DU, R_RLAM, 0, 0.0));
            rtheta_max = 1e4 * fmax(1.0, tr0.theta); rtheta_min = 1e-4 * fmax(1.0, tr0.theta);
        }
        first = 0;
        if (*iter >= max_iter) break;
        // 2. restoration problem solved to tolerance without becoming acceptable: local infeasibility
        RKktOut kk = uni(resto_kkt(c, mu, eta));
        const double E0 = resto_kkt_scaled(N, kk, kk.cmp0);
        if (UNI(!(E0 <= 1e300) || E0 <= tol)) break;
        // 3. monotone barrier update
        {
            double cmp = kk.cmpmu;
            for (;;) {
                const double Emu = resto_kkt_scaled(N, kk, cmp);
                if (UNI(!(Emu <= 10.0 * mu) || mu <= mu_min)) break;
                const double nm = fmax(mu_min, fmin(0.2 * mu, pow(mu, 1.5)));
                if (UNI(nm == mu)) break;
                mu = uni(nm); tau = fmax(0.99, 1.0 - mu); nrf = 0;
                kk = uni(resto_kkt(c, mu, dsqrt(mu)));           // (eta changes with mu: the dual infeasibility does too)
                cmp = kk.cmpmu;
            }
        }
        eta = dsqrt(mu);
        // 4. Newton system with inertia correction
        resto_rows(c, mu);
        double delta = 0.0;
        int ok = 0;
        DirOut dir;
        for (;;) {
            ok = uni(riccati(c, SE3_MODE_RESTO, 1, delta == 0.0 ? 1 : 0, delta, mu, R_RCH, R_DX, R_DU, R_LAMP));
            if (ok) {
                dir = uni(resto_direction(c, mu, eta, tau, R_DX, R_LAMP));
                if (!dir.finite) ok = 0;
            }
            if (ok) break;
            if (UNI(delta == 0.0)) delta = (delta_last == 0.0) ? 1e-4 : fmax(1e-20, delta_last / 3.0);
            else delta = (delta_last == 0.0 || 1e5 * delta_last < delta) ? 100.0 * delta : 8.0 * delta;
            if (UNI(delta > 1e20)) break;
        }
        if (!ok) break;
        if (UNI(delta > 0)) delta_last = delta;
        const double alpha_max = dir.alpha_p;
        double alpha_z = dir.alpha_z;
        const double gphid = dir.gphid;
        // filter line search on the restoration problem
        const TrialOut tk = uni(resto_trial(c, 0.0, mu, eta, R_DX, R_DU, R_LAMP, 0, 0.0));
        const double theta_k = tk.theta, phi_k = tk.f - mu * tk.logsum;
        double alpha_min;
        {
            double am = gamma_theta;
            if (UNI(gphid < 0)) {
                am = fmin(am, gamma_phi * theta_k / (-gphid));
                if (UNI(theta_k <= rtheta_min)) am = fmin(am, pow(theta_k, s_theta) / pow(-gphid, s_phi));
            }
            alpha_min = uni(0.05 * am);
        }
        double alpha = alpha_max, alpha_test = alpha_max;
        int accepted = 0, firsttrial = 1, use_soc = 0;
        double phi_acc = 0.0;
        auto acceptable = [&](const TrialOut &t) -> int {
            if (UNI(!t.feasible)) return 0;
            const double phi_t = t.f - mu * t.logsum;
            if (UNI(!(fabs(t.theta) <= 1e300) || !(fabs(phi_t) <= 1e300))) return 0;
            if (UNI(t.theta > rtheta_max)) return 0;
            const int ftype = UNI((gphid < 0) && (alpha_test * pow(-gphid, s_phi) > pow(theta_k, s_theta)));
            int a;
            if (UNI(ftype && theta_k <= rtheta_min)) a = compare_le(phi_t - phi_k, eta_phi * alpha_test * gphid, phi_k);
            else a = compare_le(t.theta, (1 - gamma_theta) * theta_k, theta_k) || compare_le(phi_t - phi_k, -gamma_phi * theta_k, phi_k);
            for (int e = 0; e < nrf; e++)
                if (!(compare_le(t.theta, sm.rfilt_theta[e], sm.rfilt_theta[e]) || compare_le(phi_t, sm.rfilt_phi[e], sm.rfilt_phi[e]))) a = 0;
            return uni(a);
        };
        while (!accepted) {
            (*nls)++;
            const int want_soc = firsttrial && max_soc > 0;
            TrialOut tr = uni(resto_trial(c, alpha, mu, eta, R_DX, R_DU, R_LAMP, want_soc ? 1 : 0, alpha));
            alpha_test = alpha;
            accepted = acceptable(tr);
            if (accepted) { phi_acc = tr.f - mu * tr.logsum; break; }
            if (UNI(want_soc && tr.theta >= theta_k)) {
                double theta_soc_old = theta_k, theta_tr = tr.theta;
                for (int it = 0; it < max_soc && !accepted; it++) {
                    if (UNI(it > 0 && !(theta_tr <= kappa_soc * theta_soc_old))) break;
                    theta_soc_old = theta_tr;
                    resto_soc_rows(c);
                    riccati(c, SE3_MODE_RESTO, 0, 0, delta, mu, R_CS, R_SDX, R_SDU, R_SLAMP);
                    const DirOut sd = uni(resto_direction(c, mu, eta, tau, R_SDX, R_SLAMP));
                    const double a_soc = sd.alpha_p;
                    const TrialOut t2 = uni(resto_trial(c, a_soc, mu, eta, R_SDX, R_SDU, R_SLAMP, 2, a_soc));
                    theta_tr = t2.theta;
                    if (acceptable(t2)) {
                        accepted = 1; use_soc = 1; phi_acc = t2.f - mu * t2.logsum;
                        alpha = a_soc; alpha_z = sd.alpha_z;
                    }
                }
                if (accepted) break;
            }
            firsttrial = 0;
            alpha *= 0.5;
            if (UNI(alpha < alpha_min)) break;
        }
        if (!accepted) break;                 // the restoration phase cannot restore itself: failure
        {   // filter augmentation
            const int ftype = (gphid < 0) && (alpha_test * pow(-gphid, s_phi) > pow(theta_k, s_theta));
            const int armijo = compare_le(phi_acc - phi_k, eta_phi * alpha_test * gphid, phi_k);
            if (UNI(!ftype || !armijo)) {
                const double et = (1 - gamma_theta) * theta_k, ep = phi_k - gamma_phi * theta_k;
                WARP_SYNC();
                WARP_FOR(lane) {
                    if (lane == 0) {
                        int w_ = 0;
                        for (int i = 0; i < nrf; i++)
                            if (!(sm.rfilt_theta[i] >= et && sm.rfilt_phi[i] >= ep)) { sm.rfilt_theta[w_] = sm.rfilt_theta[i]; sm.rfilt_phi[w_] = sm.rfilt_phi[i]; w_++; }
                        if (w_ < FILT_MAX) { sm.rfilt_theta[w_] = et; sm.rfilt_phi[w_] = ep; w_++; }
                        sm.nfilt = w_;          // (scratch: the caller keeps the original count in a register and writes it back)
                    }
                }
                WARP_SYNC();
                nrf = uni((int)sm.nfilt);
            }
        }
        if (use_soc) resto_apply(c, alpha, alpha_z, mu, R_SDX, R_SLAMP);
        else resto_apply(c, alpha, alpha_z, mu, R_DX, R_LAMP);
    }
    resto_finish(c, success, mu_orig, tau_orig);
    if (!success) {
        double th, ci;
        eval_constraints(c, &th, &ci);          // rows R_C of the restored iterate
    }
    RestoOut ro;
    ro.ok = success; ro.iter = iter_v; ro.nls = nls_v; ro.theta = theta_o; ro.f = f_o; ro.logsum = logsum_o;
    return ro;
}
