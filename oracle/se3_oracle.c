/*
 * se3_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).  See se3_oracle.h.
 *
 * fp64 restatement of the reference hot path.  Each block cites the reference file:line it follows.
 * The interior-point method restates IPOPT's published algorithm (Waechter & Biegler, Math. Prog. 2006;
 * IPOPT 3.12 option defaults) because the reference calls casadi.nlpsol('ipopt') (quad_OC.py:172), a
 * third-party wheel that is not vendored: PARITY UNPINNED for the solver iterations, pinned for the NLP
 * definition and the reward (tests/golden).
 */
#include "se3_oracle.h"

#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <stdatomic.h>
#include <unistd.h>

#define NX 13
#define NU 4
#define NMAX SE3O_NMAX

/* ------------------------------------------------------------------------------------------------ */
/* parameters: quad_policy.py:37-56                                                                  */
/* ------------------------------------------------------------------------------------------------ */
void se3o_default_params(se3o_params *p)
{
    memset(p, 0, sizeof(*p));
    p->Jx = 0.0023; p->Jy = 0.0023; p->Jz = 0.004; p->mass = 0.5; p->arm_l = 0.35; p->c_tau = 0.0245; /* quad_policy.py:37 */
    p->grav = 9.78;                                                                               /* quad_model.py:37 */
    p->wrt = 5; p->wqt = 80; p->wthrust = 0.1; p->wrf = 5; p->wvf = 5; p->wwf = 3;                /* quad_policy.py:38 */
    p->wdu = 1.0;                                                                                 /* quad_OC.py:150 */
    p->gamma_amp = 60.0; p->gamma_rate = 10.0;                                                    /* quad_OC.py:145 */
    p->dt = 0.1;                                                                                  /* quad_policy.py:43 */
    p->u_lb = 0.0; p->u_ub = 2.0 * 1.22;                                                          /* quad_policy.py:48-51 */
    p->w_bound = 3.14159265358979323846 / 2.0;                                                    /* quad_policy.py:47 */
    p->wing_len = 1.5;                                                                            /* quad_policy.py:19 */
    p->N = 50;                                                                                    /* quad_policy.py:17 */
    p->att_power = 2;                                                                             /* quad_model.py:210 */
    p->tol = 1e-8; p->acceptable_tol = 1e-6; p->mu_init = 0.1;
    p->max_iter = 3000; p->acceptable_iter = 15; p->max_soc = 4;
    p->flags = SE3O_FLAG_SCALING | SE3O_FLAG_LSINIT | SE3O_FLAG_RESTO | SE3O_FLAG_WATCHDOG | SE3O_FLAG_FILTER_RESET;
}

/* quad_policy.py:10-13 Rd2Rp, quad_model.py:818-825 toQuaternion */
void se3o_rd2quat(const double a[3], double q[4])
{
    double mag = sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
    double theta = 2.0 * atan(mag);
    double v[3] = {a[0] + 1e-8, a[1], a[2]};
    double vm = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    v[0] /= vm; v[1] /= vm; v[2] /= vm;                  /* solid_geometry.norm */
    double dn = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    v[0] /= dn; v[1] /= dn; v[2] /= dn;                  /* toQuaternion re-normalises */
    q[0] = cos(theta / 2.0);
    double s = sin(theta / 2.0);
    q[1] = s * v[0]; q[2] = s * v[1]; q[3] = s * v[2];
}

/* python round(t, 1) (quad_policy.py:70): nearest multiple of 0.1 to the exact binary value, ties to even digit */
double se3o_round1(double t)
{
    double y = t * 10.0;
    double k = floor(y);
    /* exact sign of t*10 - (k+0.5) via fma (the product error cannot hide the sign) */
    double e = fma(t, 10.0, -(k + 0.5));
    double r;
    if (e > 0) r = k + 1.0;
    else if (e < 0) r = k;
    else r = (fmod(k, 2.0) == 0.0) ? k : k + 1.0;
    /* floor(y) may be off by one when y rounded across an integer; re-check neighbours */
    if (fma(t, 10.0, -(r + 0.5)) > 0) r += 1.0;
    else if (fma(t, 10.0, -(r - 0.5)) < 0) r -= 1.0;
    return r / 10.0;
}

/* ------------------------------------------------------------------------------------------------ */
/* dynamics: quad_model.py:35-119, 637-660; Euler step quad_OC.py:47-53                              */
/* ------------------------------------------------------------------------------------------------ */
static void dyn_cont(const se3o_params *p, const double *x, const double *u, double *f)
{
    double q0 = x[6], q1 = x[7], q2 = x[8], q3 = x[9], wx = x[10], wy = x[11], wz = x[12];
    double T = u[0] + u[1] + u[2] + u[3];
    double tm = T / p->mass;
    f[0] = x[3]; f[1] = x[4]; f[2] = x[5];
    /* third row of C_B_I (= third column of C_I_B) times thrust, quad_model.py:103-108,637-643 */
    f[3] = tm * (2.0 * (q1 * q3 + q0 * q2));
    f[4] = tm * (2.0 * (q2 * q3 - q0 * q1));
    f[5] = tm * (1.0 - 2.0 * (q1 * q1 + q2 * q2)) - p->grav;
    /* 0.5*omega(w)*q, quad_model.py:110,653-660 */
    f[6] = 0.5 * (-wx * q1 - wy * q2 - wz * q3);
    f[7] = 0.5 * (wx * q0 + wz * q2 - wy * q3);
    f[8] = 0.5 * (wy * q0 - wz * q1 + wx * q3);
    f[9] = 0.5 * (wz * q0 + wy * q1 - wx * q2);
    /* inv(J)(M - w x Jw), quad_model.py:88-90,111 */
    double Mx = -u[1] * p->arm_l / 2 + u[3] * p->arm_l / 2;
    double My = -u[0] * p->arm_l / 2 + u[2] * p->arm_l / 2;
    double Mz = (u[0] - u[1] + u[2] - u[3]) * p->c_tau;
    f[10] = (Mx - (p->Jz - p->Jy) * wy * wz) / p->Jx;
    f[11] = (My - (p->Jx - p->Jz) * wx * wz) / p->Jy;
    f[12] = (Mz - (p->Jy - p->Jx) * wx * wy) / p->Jz;
}

void se3o_dyn(const se3o_params *p, const double *x, const double *u, double *xn)
{
    double f[NX];
    dyn_cont(p, x, u, f);
    for (int i = 0; i < NX; i++) xn[i] = x[i] + p->dt * f[i];
}

void se3o_dyn_jac(const se3o_params *p, const double *x, const double *u, double *A, double *B)
{
    double q0 = x[6], q1 = x[7], q2 = x[8], q3 = x[9], wx = x[10], wy = x[11], wz = x[12];
    double T = u[0] + u[1] + u[2] + u[3];
    double tm = T / p->mass, dt = p->dt;
    double F[NX][NX];
    memset(F, 0, sizeof(F));
    F[0][3] = 1; F[1][4] = 1; F[2][5] = 1;
    F[3][6] = tm * 2 * q2;  F[3][7] = tm * 2 * q3;  F[3][8] = tm * 2 * q0;  F[3][9] = tm * 2 * q1;
    F[4][6] = -tm * 2 * q1; F[4][7] = -tm * 2 * q0; F[4][8] = tm * 2 * q3;  F[4][9] = tm * 2 * q2;
    F[5][7] = -tm * 4 * q1; F[5][8] = -tm * 4 * q2;
    /* dq/dq = 0.5*Omega(w) */
    F[6][7] = -0.5 * wx; F[6][8] = -0.5 * wy; F[6][9] = -0.5 * wz;
    F[7][6] = 0.5 * wx;  F[7][8] = 0.5 * wz;  F[7][9] = -0.5 * wy;
    F[8][6] = 0.5 * wy;  F[8][7] = -0.5 * wz; F[8][9] = 0.5 * wx;
    F[9][6] = 0.5 * wz;  F[9][7] = 0.5 * wy;  F[9][8] = -0.5 * wx;
    /* dq/dw */
    F[6][10] = -0.5 * q1; F[6][11] = -0.5 * q2; F[6][12] = -0.5 * q3;
    F[7][10] = 0.5 * q0;  F[7][11] = -0.5 * q3; F[7][12] = 0.5 * q2;
    F[8][10] = 0.5 * q3;  F[8][11] = 0.5 * q0;  F[8][12] = -0.5 * q1;
    F[9][10] = -0.5 * q2; F[9][11] = 0.5 * q1;  F[9][12] = 0.5 * q0;
    /* dw/dw */
    double kx = (p->Jz - p->Jy) / p->Jx, ky = (p->Jx - p->Jz) / p->Jy, kz = (p->Jy - p->Jx) / p->Jz;
    F[10][11] = -kx * wz; F[10][12] = -kx * wy;
    F[11][10] = -ky * wz; F[11][12] = -ky * wx;
    F[12][10] = -kz * wy; F[12][11] = -kz * wx;
    for (int i = 0; i < NX; i++)
        for (int j = 0; j < NX; j++) A[i * NX + j] = (i == j ? 1.0 : 0.0) + dt * F[i][j];
    double px = 2.0 * (q1 * q3 + q0 * q2), py = 2.0 * (q2 * q3 - q0 * q1), pz = 1.0 - 2.0 * (q1 * q1 + q2 * q2);
    memset(B, 0, sizeof(double) * NX * NU);
    for (int j = 0; j < NU; j++) {
        B[3 * NU + j] = dt * px / p->mass;
        B[4 * NU + j] = dt * py / p->mass;
        B[5 * NU + j] = dt * pz / p->mass;
    }
    double ax = p->arm_l / 2 / p->Jx, ay = p->arm_l / 2 / p->Jy, az = p->c_tau / p->Jz;
    B[10 * NU + 1] = -dt * ax; B[10 * NU + 3] = dt * ax;
    B[11 * NU + 0] = -dt * ay; B[11 * NU + 2] = dt * ay;
    B[12 * NU + 0] = dt * az;  B[12 * NU + 1] = -dt * az; B[12 * NU + 2] = dt * az; B[12 * NU + 3] = -dt * az;
}

/* H (17x17, z=[x;u]) = dt * sum_i lam_i * hess(f_i)  (the identity part of x+dt*f has no curvature) */
void se3o_dyn_hess(const se3o_params *p, const double *x, const double *u, const double *lam, double *H)
{
    const int NZ = NX + NU;
    double q0 = x[6], q1 = x[7], q2 = x[8], q3 = x[9];
    double T = u[0] + u[1] + u[2] + u[3];
    double dt = p->dt, tm = T / p->mass;
    memset(H, 0, sizeof(double) * NZ * NZ);
#define HS(i, j, v) do { H[(i) * NZ + (j)] += (v); if ((i) != (j)) H[(j) * NZ + (i)] += (v); } while (0)
    double lx = lam[3], ly = lam[4], lz = lam[5];
    /* v-dot rows: (T/m) p(q) */
    HS(6, 8, dt * tm * 2 * lx);  HS(7, 9, dt * tm * 2 * lx);            /* p_x = 2(q1q3+q0q2) */
    HS(8, 9, dt * tm * 2 * ly);  HS(6, 7, -dt * tm * 2 * ly);           /* p_y = 2(q2q3-q0q1) */
    HS(7, 7, -dt * tm * 4 * lz); HS(8, 8, -dt * tm * 4 * lz);           /* p_z = 1-2(q1^2+q2^2) */
    /* q-u cross terms: d/dq [p(q)]^T lam_v / m */
    double gq[4];
    gq[0] = 2 * q2 * lx - 2 * q1 * ly;
    gq[1] = 2 * q3 * lx - 2 * q0 * ly - 4 * q1 * lz;
    gq[2] = 2 * q0 * lx + 2 * q3 * ly - 4 * q2 * lz;
    gq[3] = 2 * q1 * lx + 2 * q2 * ly;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < NU; j++) HS(6 + i, NX + j, dt * gq[i] / p->mass);
    /* q-dot rows: 0.5*Omega(w) q, bilinear in (q,w) */
    double l0 = lam[6], l1 = lam[7], l2 = lam[8], l3 = lam[9], h = 0.5 * dt;
    HS(6, 10, h * l1);  HS(6, 11, h * l2);  HS(6, 12, h * l3);
    HS(7, 10, -h * l0); HS(7, 11, h * l3);  HS(7, 12, -h * l2);
    HS(8, 10, -h * l3); HS(8, 11, -h * l0); HS(8, 12, h * l1);
    HS(9, 10, h * l2);  HS(9, 11, -h * l1); HS(9, 12, -h * l0);
    /* w-dot rows */
    double kx = (p->Jz - p->Jy) / p->Jx, ky = (p->Jx - p->Jz) / p->Jy, kz = (p->Jy - p->Jx) / p->Jz;
    HS(11, 12, -dt * kx * lam[10]);
    HS(10, 12, -dt * ky * lam[11]);
    HS(10, 11, -dt * kz * lam[12]);
#undef HS
}

/* ------------------------------------------------------------------------------------------------ */
/* costs: quad_model.py:121-213, quad_OC.py:145-151,167                                              */
/* ------------------------------------------------------------------------------------------------ */
static void dir_cosine(const double *q, double C[3][3])
{   /* quad_model.py:637-643 */
    C[0][0] = 1 - 2 * (q[2] * q[2] + q[3] * q[3]); C[0][1] = 2 * (q[1] * q[2] + q[0] * q[3]); C[0][2] = 2 * (q[1] * q[3] - q[0] * q[2]);
    C[1][0] = 2 * (q[1] * q[2] - q[0] * q[3]); C[1][1] = 1 - 2 * (q[1] * q[1] + q[3] * q[3]); C[1][2] = 2 * (q[2] * q[3] + q[0] * q[1]);
    C[2][0] = 2 * (q[1] * q[3] + q[0] * q[2]); C[2][1] = 2 * (q[2] * q[3] - q[0] * q[1]); C[2][2] = 1 - 2 * (q[1] * q[1] + q[2] * q[2]);
}

/* s(q) = trace(I - C(qt)^T C(q)) = s0 + q^T Nq q  (used for the gradient/Hessian; the value uses the trace) */
static void att_form(const double *qt, double Nq[4][4], double *s0)
{
    double Ct[3][3];
    dir_cosine(qt, Ct);
    *s0 = 3.0 - (Ct[0][0] + Ct[1][1] + Ct[2][2]);
    memset(Nq, 0, sizeof(double) * 16);
    Nq[1][1] = 2 * (Ct[1][1] + Ct[2][2]);
    Nq[2][2] = 2 * (Ct[0][0] + Ct[2][2]);
    Nq[3][3] = 2 * (Ct[0][0] + Ct[1][1]);
    Nq[1][2] = Nq[2][1] = -(Ct[0][1] + Ct[1][0]);
    Nq[1][3] = Nq[3][1] = -(Ct[0][2] + Ct[2][0]);
    Nq[2][3] = Nq[3][2] = -(Ct[1][2] + Ct[2][1]);
    Nq[0][3] = Nq[3][0] = -(Ct[0][1] - Ct[1][0]);
    Nq[0][2] = Nq[2][0] = (Ct[0][2] - Ct[2][0]);
    Nq[0][1] = Nq[1][0] = -(Ct[1][2] - Ct[2][1]);
}

static double att_trace(const double *qt, const double *q)
{   /* quad_model.py:207-210: trace(I - C(qt)^T C(q)) */
    double Ct[3][3], C[3][3], s = 3.0;
    dir_cosine(qt, Ct);
    dir_cosine(q, C);
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) s -= Ct[i][j] * C[i][j];
    return s;
}

static double gamma_k(const se3o_params *p, const se3o_problem *pb, int k)
{   /* quad_OC.py:145 */
    double d = p->dt * k - pb->t_tra;
    return p->gamma_amp * exp(-p->gamma_rate * d * d);
}

/* x-part of a running stage cost: gamma*l_tra(x) + l_path(x); Hxx optional */
static double cost_x(const se3o_params *p, const se3o_problem *pb, double gam, const double *x, double *gx, double *Hxx)
{
    double val = 0;
    if (gx) memset(gx, 0, sizeof(double) * NX);
    if (Hxx) memset(Hxx, 0, sizeof(double) * NX * NX);
    for (int i = 0; i < 3; i++) {
        double dt_ = x[i] - pb->tra_pos[i], dg = x[i] - pb->goal[i];
        val += gam * p->wrt * dt_ * dt_ + p->wrf * dg * dg;           /* quad_model.py:204,168,191 */
        val += p->wvf * x[3 + i] * x[3 + i];                          /* quad_model.py:173,192 (goal_velo=0) */
        val += p->wwf * x[10 + i] * x[10 + i];                        /* quad_model.py:183,193 */
        if (gx) {
            gx[i] = 2 * gam * p->wrt * dt_ + 2 * p->wrf * dg;
            gx[3 + i] = 2 * p->wvf * x[3 + i];
            gx[10 + i] = 2 * p->wwf * x[10 + i];
        }
        if (Hxx) {
            Hxx[i * NX + i] = 2 * gam * p->wrt + 2 * p->wrf;
            Hxx[(3 + i) * NX + 3 + i] = 2 * p->wvf;
            Hxx[(10 + i) * NX + 10 + i] = 2 * p->wwf;
        }
    }
    if (gam != 0.0) {
        double s = att_trace(pb->tra_q, x + 6);
        double Nq[4][4], s0, nq[4];
        att_form(pb->tra_q, Nq, &s0);
        for (int i = 0; i < 4; i++) {
            nq[i] = 0;
            for (int j = 0; j < 4; j++) nq[i] += Nq[i][j] * x[6 + j];
        }
        double w = gam * p->wqt;
        if (p->att_power == 2) {
            val += w * s * s;                                         /* quad_model.py:210 */
            if (gx) for (int i = 0; i < 4; i++) gx[6 + i] = w * 4 * s * nq[i];
            if (Hxx)
                for (int i = 0; i < 4; i++)
                    for (int j = 0; j < 4; j++) Hxx[(6 + i) * NX + 6 + j] = w * (8 * nq[i] * nq[j] + 4 * s * Nq[i][j]);
        } else {
            val += w * s;                                             /* gym_pybullet_drone/quad_model.py:200 */
            if (gx) for (int i = 0; i < 4; i++) gx[6 + i] = w * 2 * nq[i];
            if (Hxx)
                for (int i = 0; i < 4; i++)
                    for (int j = 0; j < 4; j++) Hxx[(6 + i) * NX + 6 + j] = w * 2 * Nq[i][j];
        }
    }
    return val;
}

/* C_k of quad_OC.py:149-150 */
double se3o_stage_cost(const se3o_params *p, const se3o_problem *pb, int k, const double *x, const double *u,
                       const double *uprev, double *gx, double *gu, double *guprev, double *Hxx)
{
    double val = cost_x(p, pb, gamma_k(p, pb, k), x, gx, Hxx);
    for (int j = 0; j < NU; j++) {
        double d = u[j] - uprev[j];
        val += p->wthrust * u[j] * u[j] + p->wdu * d * d;
        if (gu) gu[j] = 2 * p->wthrust * u[j] + 2 * p->wdu * d;
        if (guprev) guprev[j] = -2 * p->wdu * d;
    }
    return val;
}

/* final_cost, quad_model.py:195-198 (identical to goal_cost), quad_OC.py:167 */
double se3o_final_cost(const se3o_params *p, const se3o_problem *pb, const double *x, double *gx, double *Hxx)
{
    return cost_x(p, pb, 0.0, x, gx, Hxx);
}

/* ------------------------------------------------------------------------------------------------ */
/* dense NLP evaluation in the reference ordering (tests only): quad_OC.py:114-167                   */
/* ------------------------------------------------------------------------------------------------ */
void se3o_nlp_eval_dense(const se3o_params *p, const se3o_problem *pb, const double *w, const double *lam,
                         double obj_factor, double *f, double *grad, double *g, double *jac, double *hess)
{
    const int N = p->N, NZ = NX + NU, n = NZ * N + NX, m = NX * N;
    double fv = 0;
    if (grad) memset(grad, 0, sizeof(double) * n);
    if (jac) memset(jac, 0, sizeof(double) * (size_t)m * n);
    if (hess) memset(hess, 0, sizeof(double) * (size_t)n * n);
    for (int k = 0; k < N; k++) {
        const double *x = w + k * NZ, *u = x + NX, *xn = w + (k + 1) * NZ;
        const double *up = (k == 0) ? pb->ulast : (w + (k - 1) * NZ + NX);
        double gx[NX], gu[NU], gup[NU], Hxx[NX * NX];
        fv += se3o_stage_cost(p, pb, k, x, u, up, gx, gu, gup, Hxx);
        if (grad) {
            for (int i = 0; i < NX; i++) grad[k * NZ + i] += gx[i];
            for (int j = 0; j < NU; j++) {
                grad[k * NZ + NX + j] += gu[j];
                if (k > 0) grad[(k - 1) * NZ + NX + j] += gup[j];
            }
        }
        if (g) {
            double xe[NX];
            se3o_dyn(p, x, u, xe);
            for (int i = 0; i < NX; i++) g[k * NX + i] = xe[i] - xn[i];      /* quad_OC.py:162 */
        }
        if (jac) {
            double A[NX * NX], Bm[NX * NU];
            se3o_dyn_jac(p, x, u, A, Bm);
            for (int i = 0; i < NX; i++) {
                double *row = jac + (size_t)(k * NX + i) * n;
                for (int j = 0; j < NX; j++) row[k * NZ + j] = A[i * NX + j];
                for (int j = 0; j < NU; j++) row[k * NZ + NX + j] = Bm[i * NU + j];
                row[(k + 1) * NZ + i] = -1.0;
            }
        }
        if (hess) {
            for (int i = 0; i < NX; i++)
                for (int j = 0; j < NX; j++) hess[(size_t)(k * NZ + i) * n + k * NZ + j] += obj_factor * Hxx[i * NX + j];
            for (int j = 0; j < NU; j++) {
                int iu = k * NZ + NX + j;
                hess[(size_t)iu * n + iu] += obj_factor * (2 * p->wthrust + 2 * p->wdu);
                if (k > 0) {
                    int ip = (k - 1) * NZ + NX + j;
                    hess[(size_t)ip * n + ip] += obj_factor * 2 * p->wdu;
                    hess[(size_t)ip * n + iu] += -obj_factor * 2 * p->wdu;
                    hess[(size_t)iu * n + ip] += -obj_factor * 2 * p->wdu;
                }
            }
            if (lam) {
                double Hd[17 * 17];
                se3o_dyn_hess(p, x, u, lam + k * NX, Hd);
                for (int i = 0; i < NZ; i++)
                    for (int j = 0; j < NZ; j++) hess[(size_t)(k * NZ + i) * n + k * NZ + j] += Hd[i * NZ + j];
            }
        }
    }
    {
        double gx[NX], Hxx[NX * NX];
        fv += se3o_final_cost(p, pb, w + N * NZ, gx, Hxx);
        if (grad) for (int i = 0; i < NX; i++) grad[N * NZ + i] += gx[i];
        if (hess)
            for (int i = 0; i < NX; i++)
                for (int j = 0; j < NX; j++) hess[(size_t)(N * NZ + i) * n + N * NZ + j] += obj_factor * Hxx[i * NX + j];
    }
    if (f) *f = fv;
}

/* ------------------------------------------------------------------------------------------------ */
/* interior point method (IPOPT restatement)                                                         */
/* ------------------------------------------------------------------------------------------------ */
typedef struct {
    double gx[NMAX + 1][NX];   /* gradient wrt x_k (k>=1 used) of the (barrier) objective            */
    double gu[NMAX][NU];       /* stage-local gradient wrt u_k (without the -rho(u_{k+1}-u_k) part)   */
    double gub[NMAX][NU];      /* stage-local gradient wrt ubar_k = u_{k-1}                           */
    double c[NMAX][NX];        /* constraint residual x_k + dt f - x_{k+1}                            */
} rhs_t;

typedef struct {
    double A[NX][NX], B[NX][NU];
    double K[NU][NX], L[NU][NU], Hux[NU][NX], kff[NU];
    double Pxx[NX][NX], px[NX], pub[NU];   /* value function entering this stage (P_k, p_k) */
} stage_t;

typedef struct {
    const se3o_params *p;
    const se3o_problem *pb;
    int N;
    double sf;                         /* objective scaling */
    double uL, uU, wL, wU;             /* relaxed bounds */
    double gam[NMAX];
    double Nq[4][4], s0;
    /* iterate */
    double x[NMAX + 1][NX], u[NMAX][NU], lam[NMAX][NX];
    double zLu[NMAX][NU], zUu[NMAX][NU], zLw[NMAX + 1][3], zUw[NMAX + 1][3];
    /* direction */
    double dx[NMAX + 1][NX], du[NMAX][NU], lamp[NMAX][NX];
    double dzLu[NMAX][NU], dzUu[NMAX][NU], dzLw[NMAX + 1][3], dzUw[NMAX + 1][3];
    /* workspace */
    stage_t st[NMAX + 1];
    rhs_t rhs;
    double c[NMAX][NX];
    double ogx[NMAX + 1][NX], ogu[NMAX][NU];   /* unscaled objective gradient (flat NLP) */
    double xt[NMAX + 1][NX], ut[NMAX][NU];     /* trial point */
    int nfact;
} ipm_t;

static void nlp_constraints(const ipm_t *s, double (*x)[NX], double (*u)[NU], double (*c)[NX])
{
    for (int k = 0; k < s->N; k++) {
        double xe[NX];
        se3o_dyn(s->p, x[k], u[k], xe);
        for (int i = 0; i < NX; i++) c[k][i] = xe[i] - x[k + 1][i];
    }
}

static double nlp_objective(const ipm_t *s, double (*x)[NX], double (*u)[NU], double (*gx)[NX], double (*gu)[NU])
{
    double f = 0;
    const int N = s->N;
    if (gu) memset(gu, 0, sizeof(double) * NU * N);
    for (int k = 0; k < N; k++) {
        const double *up = (k == 0) ? s->pb->ulast : u[k - 1];
        double g_u[NU], g_up[NU];
        f += se3o_stage_cost(s->p, s->pb, k, x[k], u[k], up, gx ? gx[k] : NULL, g_u, g_up, NULL);
        if (gu)
            for (int j = 0; j < NU; j++) {
                gu[k][j] += g_u[j];
                if (k > 0) gu[k - 1][j] += g_up[j];
            }
    }
    f += se3o_final_cost(s->p, s->pb, x[N], gx ? gx[N] : NULL, NULL);
    return f;
}

static double theta_of(const ipm_t *s, double (*c)[NX])
{
    double t = 0;
    for (int k = 0; k < s->N; k++)
        for (int i = 0; i < NX; i++) t += fabs(c[k][i]);
    return t;
}

/* barrier objective phi_mu = sf*f - mu*sum ln(slacks); returns NaN/inf when a slack is non-positive */
static double barrier_obj(const ipm_t *s, double (*x)[NX], double (*u)[NU], double mu, double f)
{
    double b = 0;
    for (int k = 0; k < s->N; k++)
        for (int j = 0; j < NU; j++) {
            double sl = u[k][j] - s->uL, su = s->uU - u[k][j];
            if (!(sl > 0) || !(su > 0)) return INFINITY;
            b += log(sl) + log(su);
        }
    for (int k = 1; k <= s->N; k++)
        for (int j = 0; j < 3; j++) {
            double sl = x[k][10 + j] - s->wL, su = s->wU - x[k][10 + j];
            if (!(sl > 0) || !(su > 0)) return INFINITY;
            b += log(sl) + log(su);
        }
    return s->sf * f - mu * b;
}

static int chol4(double H[NU][NU], double L[NU][NU])
{
    memset(L, 0, sizeof(double) * NU * NU);
    for (int j = 0; j < NU; j++) {
        double d = H[j][j];
        for (int k = 0; k < j; k++) d -= L[j][k] * L[j][k];
        if (!(d > 4.0 * DBL_EPSILON * fabs(H[j][j])) || !(d > 0)) return 0;
        d = sqrt(d);
        L[j][j] = d;
        for (int i = j + 1; i < NU; i++) {
            double v = H[i][j];
            for (int k = 0; k < j; k++) v -= L[i][k] * L[j][k];
            L[i][j] = v / d;
        }
    }
    return 1;
}

static void chol4_solve(double L[NU][NU], const double *b, double *x)
{
    double y[NU];
    for (int i = 0; i < NU; i++) {
        double v = b[i];
        for (int k = 0; k < i; k++) v -= L[i][k] * y[k];
        y[i] = v / L[i][i];
    }
    for (int i = NU - 1; i >= 0; i--) {
        double v = y[i];
        for (int k = i + 1; k < NU; k++) v -= L[k][i] * x[k];
        x[i] = v / L[i][i];
    }
}

#define HESS_EXACT 0
#define HESS_IDENT 1

/*
 * Riccati sweep over the augmented state [x_k; ubar_k] (ubar_k = u_{k-1}); solves
 *     [H  J^T][d      ]   [-g]
 *     [J   0 ][lambda+] = [-c]
 * H = sf*hess(J) + sum lam*hess(c) + Sigma + delta_w*I (mode EXACT) or I (mode IDENT),
 * rho = 2*sf*wdu couples u_{k-1},u_k (EXACT only).
 * do_factor=1: (re)compute the matrix recursion and store it; 0: reuse (new right-hand side only).
 * Returns 1 iff every H_uu Cholesky succeeds <=> KKT inertia (n,m,0).
 */
static int riccati(ipm_t *s, int mode, int do_factor, double delta_w, const rhs_t *r,
                   double (*dx)[NX], double (*du)[NU], double (*lamp)[NX])
{
    const int N = s->N;
    const se3o_params *p = s->p;
    const double rho = (mode == HESS_EXACT) ? 2 * s->sf * p->wdu : 0.0;
    /* terminal value function */
    if (do_factor) {
        stage_t *T = &s->st[N];
        memset(T->Pxx, 0, sizeof(T->Pxx));
        if (mode == HESS_EXACT) {
            double Hxx[NX * NX];
            se3o_final_cost(p, s->pb, s->x[N], NULL, Hxx);
            for (int i = 0; i < NX; i++)
                for (int j = 0; j < NX; j++) T->Pxx[i][j] = s->sf * Hxx[i * NX + j];
            for (int j = 0; j < 3; j++) {
                double sl = s->x[N][10 + j] - s->wL, su = s->wU - s->x[N][10 + j];
                T->Pxx[10 + j][10 + j] += s->zLw[N][j] / sl + s->zUw[N][j] / su;
            }
            for (int i = 0; i < NX; i++) T->Pxx[i][i] += delta_w;
        } else {
            for (int i = 0; i < NX; i++) T->Pxx[i][i] = 1.0;
        }
        memset(T->K, 0, sizeof(T->K));
    }
    memcpy(s->st[N].px, r->gx[N], sizeof(double) * NX);
    memset(s->st[N].pub, 0, sizeof(double) * NU);

    for (int k = N - 1; k >= 0; k--) {
        stage_t *S = &s->st[k], *Sn = &s->st[k + 1];
        if (do_factor) {
            se3o_dyn_jac(p, s->x[k], s->u[k], &S->A[0][0], &S->B[0][0]);
            double Q[NX][NX], Sx[NX][NU], R[NU][NU];
            memset(Q, 0, sizeof(Q)); memset(Sx, 0, sizeof(Sx)); memset(R, 0, sizeof(R));
            if (mode == HESS_EXACT) {
                double Hd[17 * 17], Hxx[NX * NX], gxd[NX];
                se3o_dyn_hess(p, s->x[k], s->u[k], s->lam[k], Hd);
                cost_x(p, s->pb, s->gam[k], s->x[k], gxd, Hxx);
                for (int i = 0; i < NX; i++) {
                    for (int j = 0; j < NX; j++) Q[i][j] = s->sf * Hxx[i * NX + j] + Hd[i * 17 + j];
                    for (int j = 0; j < NU; j++) Sx[i][j] = Hd[i * 17 + NX + j];
                    Q[i][i] += delta_w;
                }
                if (k >= 1)
                    for (int j = 0; j < 3; j++) {
                        double sl = s->x[k][10 + j] - s->wL, su = s->wU - s->x[k][10 + j];
                        Q[10 + j][10 + j] += s->zLw[k][j] / sl + s->zUw[k][j] / su;
                    }
                for (int i = 0; i < NU; i++) {
                    for (int j = 0; j < NU; j++) R[i][j] = Hd[(NX + i) * 17 + NX + j];
                    double sl = s->u[k][i] - s->uL, su = s->uU - s->u[k][i];
                    R[i][i] += s->sf * 2 * p->wthrust + rho + s->zLu[k][i] / sl + s->zUu[k][i] / su + delta_w;
                }
            } else {
                for (int i = 0; i < NX; i++) Q[i][i] = 1.0;
                for (int i = 0; i < NU; i++) R[i][i] = 1.0;
            }
            /* W = P'xx A ; Y = P'xx B + P'xu,  P'xu = -rho K'^T (zero at the terminal stage) */
            double W[NX][NX], Y[NX][NU];
            for (int i = 0; i < NX; i++) {
                for (int j = 0; j < NX; j++) {
                    double v = 0;
                    for (int l = 0; l < NX; l++) v += Sn->Pxx[i][l] * S->A[l][j];
                    W[i][j] = v;
                }
                for (int j = 0; j < NU; j++) {
                    double v = 0;
                    for (int l = 0; l < NX; l++) v += Sn->Pxx[i][l] * S->B[l][j];
                    Y[i][j] = v + ((k + 1 < N) ? -rho * Sn->K[j][i] : 0.0);
                }
            }
            /* Hux = S^T + Y^T A ;  Huu = R + B^T Y + P'ux B + P'uu,  P'uu = rho I - rho^2 Huu'^-1 */
            double Huu[NU][NU];
            for (int a = 0; a < NU; a++) {
                for (int j = 0; j < NX; j++) {
                    double v = Sx[j][a];
                    for (int l = 0; l < NX; l++) v += Y[l][a] * S->A[l][j];
                    S->Hux[a][j] = v;
                }
                for (int b = 0; b < NU; b++) {
                    double v = R[a][b];
                    for (int l = 0; l < NX; l++) v += S->B[l][a] * Y[l][b];
                    if (k + 1 < N)
                        for (int l = 0; l < NX; l++) v += -rho * Sn->K[a][l] * S->B[l][b];
                    Huu[a][b] = v;
                }
            }
            if (k + 1 < N && rho != 0.0) {
                /* P'uu = rho I - rho^2 inv(Huu') */
                for (int b = 0; b < NU; b++) {
                    double e[NU] = {0, 0, 0, 0}, col[NU];
                    e[b] = 1.0;
                    chol4_solve(Sn->L, e, col);
                    for (int a = 0; a < NU; a++) Huu[a][b] += (a == b ? rho : 0.0) - rho * rho * col[a];
                }
            }
            /* symmetrise */
            for (int a = 0; a < NU; a++)
                for (int b = a + 1; b < NU; b++) Huu[a][b] = Huu[b][a] = 0.5 * (Huu[a][b] + Huu[b][a]);
            s->nfact += (k == N - 1);
            if (!chol4(Huu, S->L)) {
                if (getenv("SE3O_VERBOSE2"))
                    fprintf(stderr, "  chol fail stage %d delta %.2e diag %.3e %.3e %.3e %.3e\n", k, delta_w, Huu[0][0], Huu[1][1], Huu[2][2], Huu[3][3]);
                return 0;
            }
            /* K = -Huu^-1 Hux */
            for (int j = 0; j < NX; j++) {
                double col[NU], b4[NU];
                for (int a = 0; a < NU; a++) b4[a] = S->Hux[a][j];
                chol4_solve(S->L, b4, col);
                for (int a = 0; a < NU; a++) S->K[a][j] = -col[a];
            }
            /* Pxx = Q + A^T W + Hux^T K  (k>=1 only) */
            if (k >= 1) {
                for (int i = 0; i < NX; i++)
                    for (int j = 0; j < NX; j++) {
                        double v = Q[i][j];
                        for (int l = 0; l < NX; l++) v += S->A[l][i] * W[l][j];
                        for (int a = 0; a < NU; a++) v += S->Hux[a][i] * S->K[a][j];
                        S->Pxx[i][j] = v;
                    }
                for (int i = 0; i < NX; i++)
                    for (int j = i + 1; j < NX; j++) S->Pxx[i][j] = S->Pxx[j][i] = 0.5 * (S->Pxx[i][j] + S->Pxx[j][i]);
            }
        }
        /* vector recursion */
        double tv[NX], hx[NX], hu[NU];
        for (int i = 0; i < NX; i++) {
            double v = Sn->px[i];
            for (int l = 0; l < NX; l++) v += Sn->Pxx[i][l] * r->c[k][l];
            tv[i] = v;
        }
        for (int i = 0; i < NX; i++) {
            double v = r->gx[k][i];
            for (int l = 0; l < NX; l++) v += S->A[l][i] * tv[l];
            hx[i] = v;
        }
        for (int a = 0; a < NU; a++) {
            double v = r->gu[k][a] + Sn->pub[a];
            for (int l = 0; l < NX; l++) v += S->B[l][a] * tv[l];
            if (k + 1 < N)
                for (int l = 0; l < NX; l++) v += -rho * Sn->K[a][l] * r->c[k][l];   /* P'ux c */
            hu[a] = v;
        }
        double kk[NU];
        chol4_solve(S->L, hu, kk);
        for (int a = 0; a < NU; a++) S->kff[a] = -kk[a];
        for (int i = 0; i < NX; i++) {
            double v = hx[i];
            for (int a = 0; a < NU; a++) v += S->Hux[a][i] * S->kff[a];
            S->px[i] = v;
        }
        for (int a = 0; a < NU; a++) S->pub[a] = r->gub[k][a] - rho * S->kff[a];
    }
    /* forward sweep */
    double dub[NU] = {0, 0, 0, 0};
    memset(dx[0], 0, sizeof(double) * NX);
    for (int k = 0; k < N; k++) {
        stage_t *S = &s->st[k], *Sn = &s->st[k + 1];
        double e[NU], kub[NU];
        /* Kubar dubar = rho Huu^-1 dubar */
        for (int a = 0; a < NU; a++) e[a] = rho * dub[a];
        chol4_solve(S->L, e, kub);
        for (int a = 0; a < NU; a++) {
            double v = S->kff[a] + kub[a];
            for (int j = 0; j < NX; j++) v += S->K[a][j] * dx[k][j];
            du[k][a] = v;
        }
        for (int i = 0; i < NX; i++) {
            double v = r->c[k][i];
            for (int j = 0; j < NX; j++) v += S->A[i][j] * dx[k][j];
            for (int a = 0; a < NU; a++) v += S->B[i][a] * du[k][a];
            dx[k + 1][i] = v;
        }
        for (int a = 0; a < NU; a++) dub[a] = du[k][a];
        for (int i = 0; i < NX; i++) {
            double v = Sn->px[i];
            for (int j = 0; j < NX; j++) v += Sn->Pxx[i][j] * dx[k + 1][j];
            if (k + 1 < N)
                for (int a = 0; a < NU; a++) v += -rho * Sn->K[a][i] * dub[a];      /* P'xu dubar */
            lamp[k][i] = v;
        }
    }
    return 1;
}

/* right-hand side of the barrier problem at the current iterate: grad phi_mu (stage-local split) and c */
static void build_rhs(ipm_t *s, double mu, rhs_t *r)
{
    const int N = s->N;
    const se3o_params *p = s->p;
    for (int k = 0; k <= N; k++) {
        double gx[NX];
        cost_x(p, s->pb, k < N ? s->gam[k] : 0.0, s->x[k], gx, NULL);
        for (int i = 0; i < NX; i++) r->gx[k][i] = s->sf * gx[i];
        if (k >= 1)
            for (int j = 0; j < 3; j++) {
                double sl = s->x[k][10 + j] - s->wL, su = s->wU - s->x[k][10 + j];
                r->gx[k][10 + j] += -mu / sl + mu / su;
            }
    }
    for (int k = 0; k < N; k++) {
        const double *up = (k == 0) ? s->pb->ulast : s->u[k - 1];
        for (int j = 0; j < NU; j++) {
            double d = s->u[k][j] - up[j];
            double sl = s->u[k][j] - s->uL, su = s->uU - s->u[k][j];
            r->gu[k][j] = s->sf * (2 * p->wthrust * s->u[k][j] + 2 * p->wdu * d) - mu / sl + mu / su;
            r->gub[k][j] = -s->sf * 2 * p->wdu * d;
        }
        memcpy(r->c[k], s->c[k], sizeof(double) * NX);
    }
}

typedef struct { double theta, phi; } fentry;
#define FILT_MAX 32

static int compare_le(double lhs, double rhs, double base)
{   /* IPOPT Compare_le */
    return lhs - rhs <= 10.0 * DBL_EPSILON * fabs(base);
}

static double frac_to_bound_primal(const ipm_t *s, double (*dx)[NX], double (*du)[NU], double tau)
{
    double a = 1.0;
    for (int k = 0; k < s->N; k++)
        for (int j = 0; j < NU; j++) {
            double d = du[k][j];
            if (d < 0) { double v = -tau * (s->u[k][j] - s->uL) / d; if (v < a) a = v; }
            else if (d > 0) { double v = tau * (s->uU - s->u[k][j]) / d; if (v < a) a = v; }
        }
    for (int k = 1; k <= s->N; k++)
        for (int j = 0; j < 3; j++) {
            double d = dx[k][10 + j];
            if (d < 0) { double v = -tau * (s->x[k][10 + j] - s->wL) / d; if (v < a) a = v; }
            else if (d > 0) { double v = tau * (s->wU - s->x[k][10 + j]) / d; if (v < a) a = v; }
        }
    return a;
}

static void compute_dz(ipm_t *s, double mu, double (*dx)[NX], double (*du)[NU])
{
    for (int k = 0; k < s->N; k++)
        for (int j = 0; j < NU; j++) {
            double sl = s->u[k][j] - s->uL, su = s->uU - s->u[k][j];
            s->dzLu[k][j] = mu / sl - s->zLu[k][j] - s->zLu[k][j] / sl * du[k][j];
            s->dzUu[k][j] = mu / su - s->zUu[k][j] + s->zUu[k][j] / su * du[k][j];
        }
    for (int k = 1; k <= s->N; k++)
        for (int j = 0; j < 3; j++) {
            double sl = s->x[k][10 + j] - s->wL, su = s->wU - s->x[k][10 + j];
            s->dzLw[k][j] = mu / sl - s->zLw[k][j] - s->zLw[k][j] / sl * dx[k][10 + j];
            s->dzUw[k][j] = mu / su - s->zUw[k][j] + s->zUw[k][j] / su * dx[k][10 + j];
        }
}

static double frac_to_bound_dual(const ipm_t *s, double tau)
{
    double a = 1.0;
#define FTB(z, dz) do { if ((dz) < 0) { double v = -tau * (z) / (dz); if (v < a) a = v; } } while (0)
    for (int k = 0; k < s->N; k++)
        for (int j = 0; j < NU; j++) { FTB(s->zLu[k][j], s->dzLu[k][j]); FTB(s->zUu[k][j], s->dzUu[k][j]); }
    for (int k = 1; k <= s->N; k++)
        for (int j = 0; j < 3; j++) { FTB(s->zLw[k][j], s->dzLw[k][j]); FTB(s->zUw[k][j], s->dzUw[k][j]); }
#undef FTB
    return a;
}

/* scaled optimality error E_mu and its unscaled pieces */
static double kkt_error(ipm_t *s, double mu, double *dual_inf, double *cviol, double *compl_)
{
    const int N = s->N;
    double dinf = 0, cv = 0, cmp = 0, lam1 = 0, z1 = 0;
    for (int k = 0; k < N; k++) {
        for (int j = 0; j < NU; j++) {
            double v = s->sf * s->ogu[k][j] - s->zLu[k][j] + s->zUu[k][j];
            for (int l = 0; l < NX; l++) v += s->st[k].B[l][j] * s->lam[k][l];
            if (fabs(v) > dinf) dinf = fabs(v);
            double sl = s->u[k][j] - s->uL, su = s->uU - s->u[k][j];
            double c1 = fabs(sl * s->zLu[k][j] - mu), c2 = fabs(su * s->zUu[k][j] - mu);
            if (c1 > cmp) cmp = c1;
            if (c2 > cmp) cmp = c2;
            z1 += fabs(s->zLu[k][j]) + fabs(s->zUu[k][j]);
        }
        for (int i = 0; i < NX; i++) {
            if (fabs(s->c[k][i]) > cv) cv = fabs(s->c[k][i]);
            lam1 += fabs(s->lam[k][i]);
        }
    }
    for (int k = 1; k <= N; k++) {
        for (int i = 0; i < NX; i++) {
            double v = s->sf * s->ogx[k][i] - s->lam[k - 1][i];
            if (k < N)
                for (int l = 0; l < NX; l++) v += s->st[k].A[l][i] * s->lam[k][l];
            if (i >= 10) v += -s->zLw[k][i - 10] + s->zUw[k][i - 10];
            if (fabs(v) > dinf) dinf = fabs(v);
        }
        for (int j = 0; j < 3; j++) {
            double sl = s->x[k][10 + j] - s->wL, su = s->wU - s->x[k][10 + j];
            double c1 = fabs(sl * s->zLw[k][j] - mu), c2 = fabs(su * s->zUw[k][j] - mu);
            if (c1 > cmp) cmp = c1;
            if (c2 > cmp) cmp = c2;
            z1 += fabs(s->zLw[k][j]) + fabs(s->zUw[k][j]);
        }
    }
    int m = NX * N, nz = 2 * (NU * N + 3 * N);
    double sd = fmax(100.0, (lam1 + z1) / (m + nz)) / 100.0;
    double sc = fmax(100.0, z1 / nz) / 100.0;
    if (dual_inf) *dual_inf = dinf;
    if (cviol) *cviol = cv;
    if (compl_) *compl_ = cmp;
    return fmax(dinf / sd, fmax(cv, cmp / sc));
}

/* IPOPT's primal-dual system error (IpIpoptCalculatedQuantities::curr_primal_dual_system_error): averaged 1-norms of the dual
 * infeasibility, the constraint violation and the mu-complementarity at the current iterate (eval_point must be current). */
static double pd_system_error(ipm_t *s, double mu)
{
    const int N = s->N;
    double dual = 0, prim = 0, cmp = 0;
    for (int k = 0; k < N; k++) {
        for (int j = 0; j < NU; j++) {
            double v = s->sf * s->ogu[k][j] - s->zLu[k][j] + s->zUu[k][j];
            for (int l = 0; l < NX; l++) v += s->st[k].B[l][j] * s->lam[k][l];
            dual += fabs(v);
            double sl = s->u[k][j] - s->uL, su = s->uU - s->u[k][j];
            cmp += fabs(sl * s->zLu[k][j] - mu) + fabs(su * s->zUu[k][j] - mu);
        }
        for (int i = 0; i < NX; i++) prim += fabs(s->c[k][i]);
    }
    for (int k = 1; k <= N; k++) {
        for (int i = 0; i < NX; i++) {
            double v = s->sf * s->ogx[k][i] - s->lam[k - 1][i];
            if (k < N)
                for (int l = 0; l < NX; l++) v += s->st[k].A[l][i] * s->lam[k][l];
            if (i >= 10) v += -s->zLw[k][i - 10] + s->zUw[k][i - 10];
            dual += fabs(v);
        }
        for (int j = 0; j < 3; j++) {
            double sl = s->x[k][10 + j] - s->wL, su = s->wU - s->x[k][10 + j];
            cmp += fabs(sl * s->zLw[k][j] - mu) + fabs(su * s->zUw[k][j] - mu);
        }
    }
    return dual / (double)((NX + NU) * N) + prim / (double)(NX * N) + cmp / (double)(2 * (NU + 3) * N);
}

static void eval_point(ipm_t *s)
{   /* constraints, objective gradient and dynamics Jacobians at the current iterate */
    nlp_constraints(s, s->x, s->u, s->c);
    nlp_objective(s, s->x, s->u, s->ogx, s->ogu);
    for (int k = 0; k < s->N; k++) se3o_dyn_jac(s->p, s->x[k], s->u[k], &s->st[k].A[0][0], &s->st[k].B[0][0]);
}

static int all_finite(const double *v, int n)
{
    for (int i = 0; i < n; i++)
        if (!isfinite(v[i])) return 0;
    return 1;
}

/* =====================================================================================================================
 * Feasibility restoration phase -- IPOPT's MinC_1NrmRestorationPhase (Waechter & Biegler 2006, section 3.3; IPOPT 3.12
 * IpRestoMinC_1Nrm.cpp, IpRestoIpoptNLP.cpp, IpRestoIterateInitializer.cpp, IpRestoConvCheck.cpp), entered when the filter
 * line search of the regular algorithm runs below alpha_min.  With y = (x_1..x_N, u_0..u_{N-1}) and the reference point y_R =
 * the current iterate, it solves
 *      min  rho sum(n + p) + eta(mu)/2 ||D_R (y - y_R)||^2     s.t.  c(y) - p + n = 0,  n, p >= 0,  y_L <= y <= y_U
 * (rho = 1000, eta(mu) = sqrt(mu), D_R = diag(1 / max(1, |y_R|))) by the SAME interior-point algorithm (monotone mu starting at
 * max(mu, ||c||_inf), filter line search with second-order correction, inertia correction) until the y part is acceptable to the
 * ORIGINAL filter and has reduced the ORIGINAL constraint violation to kappa_resto = 0.9 of its value at entry.  The slack
 * variables are eliminated from the Newton system analytically (IpAugRestoSystemSolver): what is left is the regular KKT system
 * with a negative diagonal block,  [H  J^T; J  -D],  D = n/z_n + p/z_p, i.e. a Riccati recursion with "soft" dynamics:
 * (P', p') of a stage enter through  P~ = P' - P' D^1/2 T^-1 D^1/2 P',  T = I + D^1/2 P' D^1/2  (13 x 13 Cholesky per stage).
 * ===================================================================================================================== */
#define RESTO_RHO 1000.0
typedef struct {
    double n[NMAX][NX], p[NMAX][NX], zn[NMAX][NX], zp[NMAX][NX], lam[NMAX][NX];
    double dn[NMAX][NX], dp[NMAX][NX], dzn[NMAX][NX], dzp[NMAX][NX], lamp[NMAX][NX];
    double xR[NMAX + 1][NX], uR[NMAX][NU], dRx[NMAX + 1][NX], dRu[NMAX][NU];
    double D[NMAX][NX];                                   /* n/z_n + p/z_p */
    double LT[NMAX][NX][NX], Pn[NMAX + 1][NX][NX], pn[NMAX + 1][NX];        /* chol(T_k), P_{k+1}, p_{k+1} as seen by stage k */
    double gx[NMAX + 1][NX], gu[NMAX][NU], chat[NMAX][NX], h[NMAX][NX];
    double nt[NMAX][NX], pt[NMAX][NX];                    /* trial slacks */
} resto_t;

static int chol13(double T[NX][NX], double L[NX][NX])
{
    memset(L, 0, sizeof(double) * NX * NX);
    for (int j = 0; j < NX; j++) {
        double d = T[j][j];
        for (int k = 0; k < j; k++) d -= L[j][k] * L[j][k];
        if (!(d > 4.0 * DBL_EPSILON * fabs(T[j][j])) || !(d > 0)) return 0;
        L[j][j] = sqrt(d);
        for (int i = j + 1; i < NX; i++) {
            double v = T[i][j];
            for (int k = 0; k < j; k++) v -= L[i][k] * L[j][k];
            L[i][j] = v / L[j][j];
        }
    }
    return 1;
}
static void chol13_solve(double L[NX][NX], const double *b, double *x)
{
    double y[NX];
    for (int i = 0; i < NX; i++) { double v = b[i]; for (int k = 0; k < i; k++) v -= L[i][k] * y[k]; y[i] = v / L[i][i]; }
    for (int i = NX - 1; i >= 0; i--) { double v = y[i]; for (int k = i + 1; k < NX; k++) v -= L[k][i] * x[k]; x[i] = v / L[i][i]; }
}

/* Riccati sweep of the restoration system [H J^T; J -D][d; lam+] = -[g; chat].  H = eta D_R^2 + sum lam hess(c) + Sigma + delta_w I
 * (ident: H = I).  Returns 1 iff every T_k and H_uu,k is positive definite <=> inertia (n, m, 0). */
static int riccati_resto(ipm_t *s, resto_t *R, int ident, int do_factor, double delta_w, double eta,
                         double (*dx)[NX], double (*du)[NU], double (*lamp)[NX])
{
    const int N = s->N;
    if (do_factor) {
        memset(R->Pn[N], 0, sizeof(R->Pn[N]));
        for (int i = 0; i < NX; i++) {
            double q = ident ? 1.0 : eta * R->dRx[N][i] * R->dRx[N][i] + delta_w;
            if (!ident && i >= 10) {
                double sl = s->x[N][i] - s->wL, su = s->wU - s->x[N][i];
                q += s->zLw[N][i - 10] / sl + s->zUw[N][i - 10] / su;
            }
            R->Pn[N][i][i] = q;
        }
    }
    memcpy(R->pn[N], R->gx[N], sizeof(double) * NX);
    for (int k = N - 1; k >= 0; k--) {
        stage_t *S = &s->st[k];
        double (*Pp)[NX] = R->Pn[k + 1];
        const double *D = R->D[k];
        double sq[NX];
        for (int i = 0; i < NX; i++) sq[i] = sqrt(D[i]);
        if (do_factor) {
            se3o_dyn_jac(s->p, s->x[k], s->u[k], &S->A[0][0], &S->B[0][0]);
            double T[NX][NX];
            for (int i = 0; i < NX; i++)
                for (int j = 0; j < NX; j++) T[i][j] = sq[i] * Pp[i][j] * sq[j] + (i == j ? 1.0 : 0.0);
            if (!chol13(T, R->LT[k])) return 0;
            /* P~ = P' - Y^T Y',  column j of T^-1 D^1/2 P' by two triangular solves */
            double Pt[NX][NX];
            for (int j = 0; j < NX; j++) {
                double b[NX], z[NX];
                for (int i = 0; i < NX; i++) b[i] = sq[i] * Pp[i][j];
                chol13_solve(R->LT[k], b, z);
                for (int i = 0; i < NX; i++) {
                    double v = Pp[i][j];
                    for (int l = 0; l < NX; l++) v -= Pp[i][l] * sq[l] * z[l];
                    Pt[i][j] = v;
                }
            }
            for (int i = 0; i < NX; i++)
                for (int j = i + 1; j < NX; j++) Pt[i][j] = Pt[j][i] = 0.5 * (Pt[i][j] + Pt[j][i]);
            double Q[NX][NX], Sx[NX][NU], Rm[NU][NU];
            memset(Q, 0, sizeof(Q)); memset(Sx, 0, sizeof(Sx)); memset(Rm, 0, sizeof(Rm));
            if (!ident) {
                double Hd[17 * 17];
                se3o_dyn_hess(s->p, s->x[k], s->u[k], R->lam[k], Hd);
                for (int i = 0; i < NX; i++) {
                    for (int j = 0; j < NX; j++) Q[i][j] = Hd[i * 17 + j];
                    for (int j = 0; j < NU; j++) Sx[i][j] = Hd[i * 17 + NX + j];
                    Q[i][i] += eta * R->dRx[k][i] * R->dRx[k][i] + delta_w;
                }
                if (k >= 1)
                    for (int j = 0; j < 3; j++) {
                        double sl = s->x[k][10 + j] - s->wL, su = s->wU - s->x[k][10 + j];
                        Q[10 + j][10 + j] += s->zLw[k][j] / sl + s->zUw[k][j] / su;
                    }
                for (int i = 0; i < NU; i++) {
                    for (int j = 0; j < NU; j++) Rm[i][j] = Hd[(NX + i) * 17 + NX + j];
                    double sl = s->u[k][i] - s->uL, su = s->uU - s->u[k][i];
                    Rm[i][i] += eta * R->dRu[k][i] * R->dRu[k][i] + s->zLu[k][i] / sl + s->zUu[k][i] / su + delta_w;
                }
            } else {
                for (int i = 0; i < NX; i++) Q[i][i] = 1.0;
                for (int i = 0; i < NU; i++) Rm[i][i] = 1.0;
            }
            double W[NX][NX], Y[NX][NU], Huu[NU][NU];
            for (int i = 0; i < NX; i++) {
                for (int j = 0; j < NX; j++) { double v = 0; for (int l = 0; l < NX; l++) v += Pt[i][l] * S->A[l][j]; W[i][j] = v; }
                for (int j = 0; j < NU; j++) { double v = 0; for (int l = 0; l < NX; l++) v += Pt[i][l] * S->B[l][j]; Y[i][j] = v; }
            }
            for (int a = 0; a < NU; a++) {
                for (int j = 0; j < NX; j++) { double v = Sx[j][a]; for (int l = 0; l < NX; l++) v += Y[l][a] * S->A[l][j]; S->Hux[a][j] = v; }
                for (int b = 0; b < NU; b++) { double v = Rm[a][b]; for (int l = 0; l < NX; l++) v += S->B[l][a] * Y[l][b]; Huu[a][b] = v; }
            }
            for (int a = 0; a < NU; a++)
                for (int b = a + 1; b < NU; b++) Huu[a][b] = Huu[b][a] = 0.5 * (Huu[a][b] + Huu[b][a]);
            s->nfact += (k == N - 1);
            if (!chol4(Huu, S->L)) return 0;
            for (int j = 0; j < NX; j++) {
                double col[NU], b4[NU];
                for (int a = 0; a < NU; a++) b4[a] = S->Hux[a][j];
                chol4_solve(S->L, b4, col);
                for (int a = 0; a < NU; a++) S->K[a][j] = -col[a];
            }
            if (k >= 1) {
                for (int i = 0; i < NX; i++)
                    for (int j = 0; j < NX; j++) {
                        double v = Q[i][j];
                        for (int l = 0; l < NX; l++) v += S->A[l][i] * W[l][j];
                        for (int a = 0; a < NU; a++) v += S->Hux[a][i] * S->K[a][j];
                        R->Pn[k][i][j] = v;
                    }
                for (int i = 0; i < NX; i++)
                    for (int j = i + 1; j < NX; j++) R->Pn[k][i][j] = R->Pn[k][j][i] = 0.5 * (R->Pn[k][i][j] + R->Pn[k][j][i]);
            }
            memcpy(S->Pxx, Pt, sizeof(Pt));             /* P~ of this stage (vector recursion below) */
        }
        /* vector recursion with p~ = p' - P' D^1/2 T^-1 D^1/2 p' */
        double b[NX], z[NX], pt[NX], tv[NX], hx[NX], hu[NU], kk[NU];
        for (int i = 0; i < NX; i++) b[i] = sq[i] * R->pn[k + 1][i];
        chol13_solve(R->LT[k], b, z);
        for (int i = 0; i < NX; i++) { double v = R->pn[k + 1][i]; for (int l = 0; l < NX; l++) v -= Pp[i][l] * sq[l] * z[l]; pt[i] = v; }
        for (int i = 0; i < NX; i++) { double v = pt[i]; for (int l = 0; l < NX; l++) v += S->Pxx[i][l] * R->chat[k][l]; tv[i] = v; }
        for (int i = 0; i < NX; i++) { double v = R->gx[k][i]; for (int l = 0; l < NX; l++) v += S->A[l][i] * tv[l]; hx[i] = v; }
        for (int a = 0; a < NU; a++) { double v = R->gu[k][a]; for (int l = 0; l < NX; l++) v += S->B[l][a] * tv[l]; hu[a] = v; }
        chol4_solve(S->L, hu, kk);
        for (int a = 0; a < NU; a++) S->kff[a] = -kk[a];
        for (int i = 0; i < NX; i++) { double v = hx[i]; for (int a = 0; a < NU; a++) v += S->Hux[a][i] * S->kff[a]; R->pn[k][i] = v; }
    }
    memset(dx[0], 0, sizeof(double) * NX);
    for (int k = 0; k < N; k++) {
        stage_t *S = &s->st[k];
        double (*Pp)[NX] = R->Pn[k + 1];
        const double *D = R->D[k];
        for (int a = 0; a < NU; a++) { double v = S->kff[a]; for (int j = 0; j < NX; j++) v += S->K[a][j] * dx[k][j]; du[k][a] = v; }
        double v13[NX], b[NX], z[NX];
        for (int i = 0; i < NX; i++) {
            double v = R->chat[k][i] - D[i] * R->pn[k + 1][i];
            for (int j = 0; j < NX; j++) v += S->A[i][j] * dx[k][j];
            for (int a = 0; a < NU; a++) v += S->B[i][a] * du[k][a];
            v13[i] = v;
        }
        /* dx+ = (I + D P')^-1 v = v - D^1/2 T^-1 D^1/2 P' v */
        for (int i = 0; i < NX; i++) { double t = 0; for (int l = 0; l < NX; l++) t += Pp[i][l] * v13[l]; b[i] = sqrt(D[i]) * t; }
        chol13_solve(R->LT[k], b, z);
        for (int i = 0; i < NX; i++) dx[k + 1][i] = v13[i] - sqrt(D[i]) * z[i];
        for (int i = 0; i < NX; i++) { double v = R->pn[k + 1][i]; for (int j = 0; j < NX; j++) v += Pp[i][j] * dx[k + 1][j]; lamp[k][i] = v; }
    }
    return 1;
}

static double resto_objective(const ipm_t *s, const resto_t *R, double (*x)[NX], double (*u)[NU], double (*n)[NX], double (*p)[NX], double eta)
{
    double f = 0, q = 0;
    for (int k = 0; k < s->N; k++)
        for (int i = 0; i < NX; i++) f += n[k][i] + p[k][i];
    for (int k = 1; k <= s->N; k++)
        for (int i = 0; i < NX; i++) { double d = R->dRx[k][i] * (x[k][i] - R->xR[k][i]); q += d * d; }
    for (int k = 0; k < s->N; k++)
        for (int j = 0; j < NU; j++) { double d = R->dRu[k][j] * (u[k][j] - R->uR[k][j]); q += d * d; }
    return RESTO_RHO * f + 0.5 * eta * q;
}
static double resto_barrier(const ipm_t *s, double (*x)[NX], double (*u)[NU], double (*n)[NX], double (*p)[NX], double mu, double f)
{
    double b = 0;
    for (int k = 0; k < s->N; k++)
        for (int i = 0; i < NX; i++) {
            if (!(n[k][i] > 0) || !(p[k][i] > 0)) return INFINITY;
            b += log(n[k][i]) + log(p[k][i]);
        }
    ipm_t *ss = (ipm_t *)s;
    const double sf = ss->sf;
    ss->sf = 1.0;
    double v = barrier_obj(s, x, u, mu, f);          /* f - mu sum ln(slacks of y) */
    ss->sf = sf;
    return v - mu * b;
}
static double resto_theta(const ipm_t *s, double (*c)[NX], double (*n)[NX], double (*p)[NX])
{
    double t = 0;
    for (int k = 0; k < s->N; k++)
        for (int i = 0; i < NX; i++) t += fabs(c[k][i] - p[k][i] + n[k][i]);
    return t;
}

/* scaled optimality error of the restoration problem (same E_mu definition, all multipliers counted) */
static double resto_kkt_error(ipm_t *s, resto_t *R, double mu, double eta)
{
    const int N = s->N;
    double dinf = 0, cv = 0, cmp = 0, lam1 = 0, z1 = 0;
#define UPD(v) do { double a_ = fabs(v); if (a_ > dinf) dinf = a_; } while (0)
#define CMP(sl, z) do { double a_ = fabs((sl) * (z) - mu); if (a_ > cmp) cmp = a_; z1 += fabs(z); } while (0)
    for (int k = 0; k < N; k++) {
        for (int j = 0; j < NU; j++) {
            double v = eta * R->dRu[k][j] * R->dRu[k][j] * (s->u[k][j] - R->uR[k][j]) - s->zLu[k][j] + s->zUu[k][j];
            for (int l = 0; l < NX; l++) v += s->st[k].B[l][j] * R->lam[k][l];
            UPD(v);
            CMP(s->u[k][j] - s->uL, s->zLu[k][j]); CMP(s->uU - s->u[k][j], s->zUu[k][j]);
        }
        for (int i = 0; i < NX; i++) {
            double hv = s->c[k][i] - R->p[k][i] + R->n[k][i];
            if (fabs(hv) > cv) cv = fabs(hv);
            lam1 += fabs(R->lam[k][i]);
            UPD(RESTO_RHO + R->lam[k][i] - R->zn[k][i]);
            UPD(RESTO_RHO - R->lam[k][i] - R->zp[k][i]);
            CMP(R->n[k][i], R->zn[k][i]); CMP(R->p[k][i], R->zp[k][i]);
        }
    }
    for (int k = 1; k <= N; k++) {
        for (int i = 0; i < NX; i++) {
            double v = eta * R->dRx[k][i] * R->dRx[k][i] * (s->x[k][i] - R->xR[k][i]) - R->lam[k - 1][i];
            if (k < N)
                for (int l = 0; l < NX; l++) v += s->st[k].A[l][i] * R->lam[k][l];
            if (i >= 10) v += -s->zLw[k][i - 10] + s->zUw[k][i - 10];
            UPD(v);
        }
        for (int j = 0; j < 3; j++) { CMP(s->x[k][10 + j] - s->wL, s->zLw[k][j]); CMP(s->wU - s->x[k][10 + j], s->zUw[k][j]); }
    }
#undef UPD
#undef CMP
    const int m = NX * N, nz = 2 * (NU * N + 3 * N) + 2 * NX * N;
    double sd = fmax(100.0, (lam1 + z1) / (m + nz)) / 100.0;
    double sc = fmax(100.0, z1 / nz) / 100.0;
    return fmax(dinf / sd, fmax(cv, cmp / sc));
}

/* Returns 1 when a point acceptable to the original filter was found (s->x, s->u, bound multipliers and lam updated, *fcur_out set),
 * 0 when the restoration phase itself fails (locally infeasible / its own line search fails / iteration limit). */
static int restoration_phase(ipm_t *s, double mu_orig, double tau_orig, fentry *filt, int nfilt, double theta_max_orig, int *iter, int max_iter,
                             int *nls, double *fcur_out)
{
    const int N = s->N;
    const se3o_params *p = s->p;
    resto_t *R = (resto_t *)calloc(1, sizeof(resto_t));
    ipm_t *bk = (ipm_t *)malloc(sizeof(ipm_t));
    memcpy(bk, s, sizeof(ipm_t));                                    /* the original iterate (multipliers are needed at the end) */
    const double gamma_theta = 1e-5, gamma_phi = 1e-8, s_phi = 2.3, s_theta = 1.1, eta_phi = 1e-8, kappa_soc = 0.99, kappa_resto = 0.9;
    /* reference point, scaling D_R */
    memcpy(R->xR, s->x, sizeof(R->xR)); memcpy(R->uR, s->u, sizeof(R->uR));
    for (int k = 0; k <= N; k++)
        for (int i = 0; i < NX; i++) R->dRx[k][i] = 1.0 / fmax(1.0, fabs(s->x[k][i]));
    for (int k = 0; k < N; k++)
        for (int j = 0; j < NU; j++) R->dRu[k][j] = 1.0 / fmax(1.0, fabs(s->u[k][j]));
    const double theta_orig0 = theta_of(s, s->c);
    const double f_orig0 = nlp_objective(s, s->x, s->u, NULL, NULL);
    const double phi_orig0 = barrier_obj(s, s->x, s->u, mu_orig, f_orig0);
    /* initial restoration iterate (IpRestoIterateInitializer) */
    double cinf = 0;
    for (int k = 0; k < N; k++)
        for (int i = 0; i < NX; i++) cinf = fmax(cinf, fabs(s->c[k][i]));
    double mu = fmax(mu_orig, cinf), tau = fmax(0.99, 1.0 - mu);
    for (int k = 0; k < N; k++)
        for (int i = 0; i < NX; i++) {
            const double c = s->c[k][i], a = (mu - RESTO_RHO * c) / (2.0 * RESTO_RHO);
            R->n[k][i] = a + sqrt(a * a + mu * c / (2.0 * RESTO_RHO));
            R->p[k][i] = c + R->n[k][i];
            R->zn[k][i] = mu / R->n[k][i]; R->zp[k][i] = mu / R->p[k][i];
        }
    for (int k = 0; k < N; k++)
        for (int j = 0; j < NU; j++) { s->zLu[k][j] = fmin(RESTO_RHO, s->zLu[k][j]); s->zUu[k][j] = fmin(RESTO_RHO, s->zUu[k][j]); }
    for (int k = 1; k <= N; k++)
        for (int j = 0; j < 3; j++) { s->zLw[k][j] = fmin(RESTO_RHO, s->zLw[k][j]); s->zUw[k][j] = fmin(RESTO_RHO, s->zUw[k][j]); }
    /* equality multipliers of the restoration problem start at zero (IPOPT computes a least-squares estimate and keeps it when it is below
     * constr_mult_init_max; measured on the closed-loop failures the estimate changes the iteration count by ~1%, so it is omitted) */
    const double mu_min = p->tol / 11.0;
    fentry rf[FILT_MAX];
    int nrf = 0, success = 0, first = 1;
    double theta_r0 = resto_theta(s, s->c, R->n, R->p);
    const double rtheta_max = 1e4 * fmax(1.0, theta_r0), rtheta_min = 1e-4 * fmax(1.0, theta_r0);
    double delta_last = 0.0;
    double (*sdx)[NX] = (double (*)[NX])malloc(sizeof(double) * NX * (NMAX + 1));
    double (*sdu)[NU] = (double (*)[NU])malloc(sizeof(double) * NU * NMAX);
    double (*slam)[NX] = (double (*)[NX])malloc(sizeof(double) * NX * NMAX);

    for (;; (*iter)++) {
        double eta = sqrt(mu);
        /* 1. progress for the ORIGINAL problem (IpRestoConvCheck): not in the first restoration iteration */
        if (!first) {
            const double th_o = theta_of(s, s->c);
            const double f_o = nlp_objective(s, s->x, s->u, NULL, NULL);
            const double ph_o = barrier_obj(s, s->x, s->u, mu_orig, f_o);
            int ok = isfinite(th_o) && isfinite(ph_o) && th_o <= kappa_resto * theta_orig0 && th_o <= theta_max_orig;
            for (int e = 0; ok && e < nfilt; e++)
                if (!(compare_le(th_o, filt[e].theta, filt[e].theta) || compare_le(ph_o, filt[e].phi, filt[e].phi))) ok = 0;
            if (ok && !(compare_le(th_o, (1 - gamma_theta) * theta_orig0, theta_orig0) ||
                        compare_le(ph_o - phi_orig0, -gamma_phi * theta_orig0, phi_orig0))) ok = 0;
            if (ok) { success = 1; *fcur_out = f_o; break; }
        }
        first = 0;
        if (*iter >= max_iter) break;
        /* 2. restoration problem solved to tolerance without becoming acceptable: local infeasibility */
        double E0 = resto_kkt_error(s, R, 0.0, eta);
        if (!isfinite(E0) || E0 <= p->tol) break;
        if (getenv("SE3O_VERBOSE"))
            fprintf(stderr, "it %3dr fR=%.6e E0=%.3e mu=%.2e th_orig=%.3e th_R=%.3e\n", *iter, resto_objective(s, R, s->x, s->u, R->n, R->p, eta), E0, mu,
                    theta_of(s, s->c), resto_theta(s, s->c, R->n, R->p));
        /* 3. monotone barrier update */
        for (;;) {
            double Emu = resto_kkt_error(s, R, mu, sqrt(mu));
            if (!(Emu <= 10.0 * mu) || mu <= mu_min) break;
            double nm = fmax(mu_min, fmin(0.2 * mu, pow(mu, 1.5)));
            if (nm == mu) break;
            mu = nm; tau = fmax(0.99, 1.0 - mu); nrf = 0;
        }
        eta = sqrt(mu);
        /* 4. Newton system */
        for (int k = 0; k < N; k++)
            for (int i = 0; i < NX; i++) {
                const double sn = R->zn[k][i] / R->n[k][i], sp = R->zp[k][i] / R->p[k][i];
                const double rn = RESTO_RHO + R->lam[k][i] - mu / R->n[k][i], rp = RESTO_RHO - R->lam[k][i] - mu / R->p[k][i];
                R->D[k][i] = 1.0 / sn + 1.0 / sp;
                R->h[k][i] = s->c[k][i] - R->p[k][i] + R->n[k][i];
                R->chat[k][i] = R->h[k][i] - rn / sn + rp / sp + R->D[k][i] * R->lam[k][i];
            }
        for (int k = 0; k <= N; k++)
            for (int i = 0; i < NX; i++) {
                double g = (k >= 1) ? eta * R->dRx[k][i] * R->dRx[k][i] * (s->x[k][i] - R->xR[k][i]) : 0.0;
                if (k >= 1 && i >= 10) { double sl = s->x[k][i] - s->wL, su = s->wU - s->x[k][i]; g += -mu / sl + mu / su; }
                R->gx[k][i] = g;
            }
        for (int k = 0; k < N; k++)
            for (int j = 0; j < NU; j++) {
                double sl = s->u[k][j] - s->uL, su = s->uU - s->u[k][j];
                R->gu[k][j] = eta * R->dRu[k][j] * R->dRu[k][j] * (s->u[k][j] - R->uR[k][j]) - mu / sl + mu / su;
            }
        double delta = 0.0;
        int ok = 0;
        for (;;) {
            ok = riccati_resto(s, R, 0, 1, delta, eta, s->dx, s->du, R->lamp);
            if (ok && !all_finite(&s->dx[0][0], NX * (N + 1))) ok = 0;
            if (ok) break;
            if (delta == 0.0) delta = (delta_last == 0.0) ? 1e-4 : fmax(1e-20, delta_last / 3.0);
            else delta = (delta_last == 0.0 || 1e5 * delta_last < delta) ? 100.0 * delta : 8.0 * delta;
            if (delta > 1e20) break;
        }
        if (!ok) break;
        if (delta > 0) delta_last = delta;
        /* slack and multiplier steps */
#define SLACK_STEPS(LAMP) do {                                                                                             \
        for (int k = 0; k < N; k++)                                                                                        \
            for (int i = 0; i < NX; i++) {                                                                                 \
                const double sn = R->zn[k][i] / R->n[k][i], sp = R->zp[k][i] / R->p[k][i];                                 \
                const double rn = RESTO_RHO + R->lam[k][i] - mu / R->n[k][i], rp = RESTO_RHO - R->lam[k][i] - mu / R->p[k][i]; \
                const double dl = (LAMP)[k][i] - R->lam[k][i];                                                             \
                R->dn[k][i] = -(rn + dl) / sn; R->dp[k][i] = -(rp - dl) / sp;                                              \
                R->dzn[k][i] = mu / R->n[k][i] - R->zn[k][i] - sn * R->dn[k][i];                                           \
                R->dzp[k][i] = mu / R->p[k][i] - R->zp[k][i] - sp * R->dp[k][i];                                           \
            } } while (0)
#define FTB_SLACK(a_, DN, DP) do {                                                                                         \
        for (int k = 0; k < N; k++)                                                                                        \
            for (int i = 0; i < NX; i++) {                                                                                 \
                if ((DN)[k][i] < 0) { double v = -tau * R->n[k][i] / (DN)[k][i]; if (v < a_) a_ = v; }                     \
                if ((DP)[k][i] < 0) { double v = -tau * R->p[k][i] / (DP)[k][i]; if (v < a_) a_ = v; }                     \
            } } while (0)
        SLACK_STEPS(R->lamp);
        compute_dz(s, mu, s->dx, s->du);
        double alpha_max = frac_to_bound_primal(s, s->dx, s->du, tau);
        FTB_SLACK(alpha_max, R->dn, R->dp);
        double alpha_z = frac_to_bound_dual(s, tau);
        FTB_SLACK(alpha_z, R->dzn, R->dzp);          /* (same rule on z_n, z_p: step dz, value z) */
        {   /* FTB_SLACK tested n/dn; redo for the multipliers proper */
            alpha_z = frac_to_bound_dual(s, tau);
            for (int k = 0; k < N; k++)
                for (int i = 0; i < NX; i++) {
                    if (R->dzn[k][i] < 0) { double v = -tau * R->zn[k][i] / R->dzn[k][i]; if (v < alpha_z) alpha_z = v; }
                    if (R->dzp[k][i] < 0) { double v = -tau * R->zp[k][i] / R->dzp[k][i]; if (v < alpha_z) alpha_z = v; }
                }
        }
        /* filter line search on the restoration problem */
        const double theta_k = resto_theta(s, s->c, R->n, R->p);
        const double f_k = resto_objective(s, R, s->x, s->u, R->n, R->p, eta);
        const double phi_k = resto_barrier(s, s->x, s->u, R->n, R->p, mu, f_k);
        double gphid = 0;
        for (int k = 1; k <= N; k++)
            for (int i = 0; i < NX; i++) gphid += R->gx[k][i] * s->dx[k][i];
        for (int k = 0; k < N; k++) {
            for (int j = 0; j < NU; j++) gphid += R->gu[k][j] * s->du[k][j];
            for (int i = 0; i < NX; i++) gphid += (RESTO_RHO - mu / R->n[k][i]) * R->dn[k][i] + (RESTO_RHO - mu / R->p[k][i]) * R->dp[k][i];
        }
        double alpha_min;
        {
            double am = gamma_theta;
            if (gphid < 0) {
                am = fmin(am, gamma_phi * theta_k / (-gphid));
                if (theta_k <= rtheta_min) am = fmin(am, pow(theta_k, s_theta) / pow(-gphid, s_phi));
            }
            alpha_min = 0.05 * am;
        }
        double alpha = alpha_max, alpha_test = alpha;
        int accepted = 0, firsttrial = 1;
        double (*ddx)[NX] = s->dx, (*ddu)[NU] = s->du, (*ddn)[NX] = R->dn, (*ddp)[NX] = R->dp;
        double (*ct)[NX] = (double (*)[NX])malloc(sizeof(double) * NX * NMAX);
#define RTRIAL(a_, DX, DU, DN, DP, theta_out, phi_out) do {                                                                \
            for (int k = 0; k <= N; k++) for (int i = 0; i < NX; i++) s->xt[k][i] = s->x[k][i] + (a_) * (DX)[k][i];        \
            for (int k = 0; k < N; k++) {                                                                                  \
                for (int j = 0; j < NU; j++) s->ut[k][j] = s->u[k][j] + (a_) * (DU)[k][j];                                 \
                for (int i = 0; i < NX; i++) { R->nt[k][i] = R->n[k][i] + (a_) * (DN)[k][i]; R->pt[k][i] = R->p[k][i] + (a_) * (DP)[k][i]; } \
            }                                                                                                              \
            nlp_constraints(s, s->xt, s->ut, ct);                                                                          \
            theta_out = resto_theta(s, ct, R->nt, R->pt);                                                                  \
            phi_out = resto_barrier(s, s->xt, s->ut, R->nt, R->pt, mu, resto_objective(s, R, s->xt, s->ut, R->nt, R->pt, eta)); \
        } while (0)
#define RACCEPT(theta_t_, phi_t_, ok_) do {                                                                                \
            int a_ = 1;                                                                                                    \
            if (!isfinite(theta_t_) || !isfinite(phi_t_)) a_ = 0;                                                          \
            else if (theta_t_ > rtheta_max) a_ = 0;                                                                        \
            else {                                                                                                         \
                int ftype_ = (gphid < 0) && (alpha_test * pow(-gphid, s_phi) > pow(theta_k, s_theta));                     \
                if (ftype_ && theta_k <= rtheta_min) a_ = compare_le(phi_t_ - phi_k, eta_phi * alpha_test * gphid, phi_k); \
                else a_ = compare_le(theta_t_, (1 - gamma_theta) * theta_k, theta_k) || compare_le(phi_t_ - phi_k, -gamma_phi * theta_k, phi_k); \
                for (int e_ = 0; a_ && e_ < nrf; e_++)                                                                     \
                    if (!(compare_le(theta_t_, rf[e_].theta, rf[e_].theta) || compare_le(phi_t_, rf[e_].phi, rf[e_].phi))) a_ = 0; \
            }                                                                                                              \
            ok_ = a_;                                                                                                      \
        } while (0)
        double phi_acc = 0;
        while (!accepted) {
            (*nls)++;
            double theta_t, phi_t;
            RTRIAL(alpha, ddx, ddu, ddn, ddp, theta_t, phi_t);
            alpha_test = alpha;
            RACCEPT(theta_t, phi_t, accepted);
            if (accepted) { phi_acc = phi_t; break; }
            if (firsttrial && p->max_soc > 0 && theta_t >= theta_k) {
                /* second-order correction: only the constraint residual of the reduced system changes, chat += (alpha h + h(trial)) - h */
                double (*chat0)[NX] = (double (*)[NX])malloc(sizeof(R->chat));
                double (*hs)[NX] = (double (*)[NX])malloc(sizeof(R->h));
                memcpy(chat0, R->chat, sizeof(R->chat));
                for (int k = 0; k < N; k++)
                    for (int i = 0; i < NX; i++) hs[k][i] = alpha * R->h[k][i] + (ct[k][i] - R->pt[k][i] + R->nt[k][i]);
                double theta_soc_old = theta_k, theta_tr = theta_t;
                for (int it = 0; it < p->max_soc && !accepted; it++) {
                    if (it > 0 && !(theta_tr <= kappa_soc * theta_soc_old)) break;
                    theta_soc_old = theta_tr;
                    for (int k = 0; k < N; k++)
                        for (int i = 0; i < NX; i++) R->chat[k][i] = chat0[k][i] + hs[k][i] - R->h[k][i];
                    riccati_resto(s, R, 0, 0, delta, eta, sdx, sdu, slam);
                    /* slack steps of the corrected system: r_c is h_soc */
                    double (*sdn)[NX] = (double (*)[NX])malloc(sizeof(R->dn)), (*sdp)[NX] = (double (*)[NX])malloc(sizeof(R->dp));
                    for (int k = 0; k < N; k++)
                        for (int i = 0; i < NX; i++) {
                            const double sn = R->zn[k][i] / R->n[k][i], sp = R->zp[k][i] / R->p[k][i];
                            const double rn = RESTO_RHO + R->lam[k][i] - mu / R->n[k][i], rp = RESTO_RHO - R->lam[k][i] - mu / R->p[k][i];
                            const double dl = slam[k][i] - R->lam[k][i];
                            sdn[k][i] = -(rn + dl) / sn; sdp[k][i] = -(rp - dl) / sp;
                        }
                    double a_soc = frac_to_bound_primal(s, sdx, sdu, tau);
                    FTB_SLACK(a_soc, sdn, sdp);
                    double phi2;
                    RTRIAL(a_soc, sdx, sdu, sdn, sdp, theta_tr, phi2);
                    int okk;
                    RACCEPT(theta_tr, phi2, okk);
                    if (okk) {
                        accepted = 1; phi_acc = phi2; alpha = a_soc;
                        memcpy(s->dx, sdx, sizeof(double) * NX * (N + 1)); memcpy(s->du, sdu, sizeof(double) * NU * N);
                        memcpy(R->lamp, slam, sizeof(R->lamp));
                        SLACK_STEPS(R->lamp);
                        compute_dz(s, mu, s->dx, s->du);
                        alpha_z = frac_to_bound_dual(s, tau);
                        for (int k = 0; k < N; k++)
                            for (int i = 0; i < NX; i++) {
                                if (R->dzn[k][i] < 0) { double v = -tau * R->zn[k][i] / R->dzn[k][i]; if (v < alpha_z) alpha_z = v; }
                                if (R->dzp[k][i] < 0) { double v = -tau * R->zp[k][i] / R->dzp[k][i]; if (v < alpha_z) alpha_z = v; }
                            }
                    } else {
                        for (int k = 0; k < N; k++)
                            for (int i = 0; i < NX; i++) hs[k][i] = a_soc * hs[k][i] + (ct[k][i] - R->pt[k][i] + R->nt[k][i]);
                    }
                    free(sdn); free(sdp);
                }
                memcpy(R->chat, chat0, sizeof(R->chat));
                free(chat0); free(hs);
                if (accepted) break;
            }
            firsttrial = 0;
            alpha *= 0.5;
            if (alpha < alpha_min) break;
        }
        free(ct);
        if (!accepted) break;                 /* the restoration phase cannot restore itself: failure */
        {   /* filter augmentation */
            int ftype = (gphid < 0) && (alpha_test * pow(-gphid, s_phi) > pow(theta_k, s_theta));
            int armijo = compare_le(phi_acc - phi_k, eta_phi * alpha_test * gphid, phi_k);
            if (!ftype || !armijo) {
                fentry e = {(1 - gamma_theta) * theta_k, phi_k - gamma_phi * theta_k};
                int w_ = 0;
                for (int i = 0; i < nrf; i++)
                    if (!(rf[i].theta >= e.theta && rf[i].phi >= e.phi)) rf[w_++] = rf[i];
                nrf = w_;
                if (nrf < FILT_MAX) rf[nrf++] = e;
            }
        }
        /* accept the step */
        memcpy(s->x, s->xt, sizeof(double) * NX * (N + 1));
        memcpy(s->u, s->ut, sizeof(double) * NU * N);
        const double ks = 1e10;
#define ZUPD(z, dz, slack) do { double zn_ = (z) + alpha_z * (dz); double sl_ = (slack);                      \
            zn_ = fmax(fmin(zn_, ks * mu / sl_), mu / (ks * sl_)); (z) = zn_; } while (0)
        for (int k = 0; k < N; k++) {
            for (int i = 0; i < NX; i++) {
                R->n[k][i] = R->nt[k][i]; R->p[k][i] = R->pt[k][i];
                R->lam[k][i] += alpha * (R->lamp[k][i] - R->lam[k][i]);
                ZUPD(R->zn[k][i], R->dzn[k][i], R->n[k][i]); ZUPD(R->zp[k][i], R->dzp[k][i], R->p[k][i]);
            }
            for (int j = 0; j < NU; j++) {
                ZUPD(s->zLu[k][j], s->dzLu[k][j], s->u[k][j] - s->uL);
                ZUPD(s->zUu[k][j], s->dzUu[k][j], s->uU - s->u[k][j]);
            }
        }
        for (int k = 1; k <= N; k++)
            for (int j = 0; j < 3; j++) {
                ZUPD(s->zLw[k][j], s->dzLw[k][j], s->x[k][10 + j] - s->wL);
                ZUPD(s->zUw[k][j], s->dzUw[k][j], s->wU - s->x[k][10 + j]);
            }
#undef ZUPD
        eval_point(s);
    }
    free(sdx); free(sdu); free(slam);
    if (success) {
        /* back to the regular algorithm (MinC_1NrmRestorationPhase::PerformRestoration): equality multipliers are reset to zero
         * (constr_mult_reset_threshold = 0); bound multipliers take the primal-dual step that belongs to the primal move,
         * dz = (mu - z (s_new - s_old)) / s_old - z, with the fraction-to-the-boundary rule, and are all reset to 1 when any of
         * them exceeds bound_mult_reset_threshold = 1000 */
        memset(s->lam, 0, sizeof(s->lam));
        double amax = 1.0, zmax = 0.0;
#define DZ(zold, sold, snew) ((mu_orig - (zold) * ((snew) - (sold))) / (sold) - (zold))
#define FT(zold, dz) do { if ((dz) < 0) { double v = -tau_orig * (zold) / (dz); if (v < amax) amax = v; } } while (0)
        for (int pass = 0; pass < 2; pass++) {
            for (int k = 0; k < N; k++)
                for (int j = 0; j < NU; j++) {
                    const double zl = bk->zLu[k][j], zu = bk->zUu[k][j];
                    const double dl = DZ(zl, bk->u[k][j] - s->uL, s->u[k][j] - s->uL), dq = DZ(zu, s->uU - bk->u[k][j], s->uU - s->u[k][j]);
                    if (pass == 0) { FT(zl, dl); FT(zu, dq); }
                    else { s->zLu[k][j] = zl + amax * dl; s->zUu[k][j] = zu + amax * dq; zmax = fmax(zmax, fmax(s->zLu[k][j], s->zUu[k][j])); }
                }
            for (int k = 1; k <= N; k++)
                for (int j = 0; j < 3; j++) {
                    const double zl = bk->zLw[k][j], zu = bk->zUw[k][j];
                    const double dl = DZ(zl, bk->x[k][10 + j] - s->wL, s->x[k][10 + j] - s->wL), dq = DZ(zu, s->wU - bk->x[k][10 + j], s->wU - s->x[k][10 + j]);
                    if (pass == 0) { FT(zl, dl); FT(zu, dq); }
                    else { s->zLw[k][j] = zl + amax * dl; s->zUw[k][j] = zu + amax * dq; zmax = fmax(zmax, fmax(s->zLw[k][j], s->zUw[k][j])); }
                }
        }
#undef DZ
#undef FT
        if (zmax > 1e3) {
            for (int k = 0; k < N; k++)
                for (int j = 0; j < NU; j++) { s->zLu[k][j] = 1.0; s->zUu[k][j] = 1.0; }
            for (int k = 1; k <= N; k++)
                for (int j = 0; j < 3; j++) { s->zLw[k][j] = 1.0; s->zUw[k][j] = 1.0; }
        }
        eval_point(s);
    } else {
        const int nf = s->nfact;
        memcpy(s, bk, sizeof(ipm_t));            /* failure: leave the iterate where the regular algorithm stopped */
        s->nfact = nf;
    }
    free(bk); free(R);
    return success;
}

int se3o_solve(const se3o_params *p, const se3o_problem *pb, se3o_solution *sol)
{
    return se3o_solve_from(p, pb, NULL, NULL, sol);
}

/* Same algorithm started from a user-supplied primal point (IPOPT's treatment of any x0: pushed into the interior of the
 * bounds, z = 1, least-squares multipliers): the receding-horizon "warm start" of main.py-style closed loops. */
int se3o_solve_from(const se3o_params *p, const se3o_problem *pb, const double *Xinit, const double *Uinit, se3o_solution *sol)
{
    const int N = p->N;
    ipm_t *s = (ipm_t *)calloc(1, sizeof(ipm_t));
    s->p = p; s->pb = pb; s->N = N; s->sf = 1.0;
    memset(sol, 0, sizeof(*sol));
    att_form(pb->tra_q, s->Nq, &s->s0);
    for (int k = 0; k < N; k++) s->gam[k] = gamma_k(p, pb, k);
    /* bound_relax_factor = 1e-8 */
    s->uL = p->u_lb - 1e-8 * fmax(1.0, fabs(p->u_lb));
    s->uU = p->u_ub + 1e-8 * fmax(1.0, fabs(p->u_ub));
    s->wL = -p->w_bound - 1e-8 * fmax(1.0, fabs(p->w_bound));
    s->wU = p->w_bound + 1e-8 * fmax(1.0, fabs(p->w_bound));
    /* initial guess, quad_OC.py:128,142,158: x0 fixed, mid-point of bounds elsewhere (x_k = 0, u_k = (lb+ub)/2) */
    memcpy(s->x[0], pb->x0, sizeof(double) * NX);
    for (int k = 1; k <= N; k++) {
        if (Xinit) memcpy(s->x[k], Xinit + k * NX, sizeof(double) * NX);
        else memset(s->x[k], 0, sizeof(double) * NX);
    }
    for (int k = 0; k < N; k++)
        for (int j = 0; j < NU; j++) s->u[k][j] = Uinit ? Uinit[k * NU + j] : 0.5 * (p->u_lb + p->u_ub);
    /* push into the interior: bound_push = bound_frac = 1e-2 */
    {
        double pl = fmin(1e-2 * fmax(1.0, fabs(s->uL)), 1e-2 * (s->uU - s->uL));
        double pu = fmin(1e-2 * fmax(1.0, fabs(s->uU)), 1e-2 * (s->uU - s->uL));
        for (int k = 0; k < N; k++)
            for (int j = 0; j < NU; j++) s->u[k][j] = fmin(fmax(s->u[k][j], s->uL + pl), s->uU - pu);
        double ql = fmin(1e-2 * fmax(1.0, fabs(s->wL)), 1e-2 * (s->wU - s->wL));
        double qu = fmin(1e-2 * fmax(1.0, fabs(s->wU)), 1e-2 * (s->wU - s->wL));
        for (int k = 1; k <= N; k++)
            for (int j = 0; j < 3; j++) s->x[k][10 + j] = fmin(fmax(s->x[k][10 + j], s->wL + ql), s->wU - qu);
    }
    for (int k = 0; k <= N; k++)
        for (int j = 0; j < 3; j++) { s->zLw[k][j] = 1.0; s->zUw[k][j] = 1.0; }
    for (int k = 0; k < N; k++)
        for (int j = 0; j < NU; j++) { s->zLu[k][j] = 1.0; s->zUu[k][j] = 1.0; }

    eval_point(s);
    /* gradient-based scaling (nlp_scaling_max_gradient = 100) at the starting point */
    if (p->flags & SE3O_FLAG_SCALING) {
        double gmax = 0;
        for (int k = 1; k <= N; k++)
            for (int i = 0; i < NX; i++) gmax = fmax(gmax, fabs(s->ogx[k][i]));
        for (int k = 0; k < N; k++)
            for (int j = 0; j < NU; j++) gmax = fmax(gmax, fabs(s->ogu[k][j]));
        if (gmax > 100.0) s->sf = fmax(100.0 / gmax, 1e-8);
    }
    sol->obj_scale = s->sf;

    /* least-squares multiplier estimate; discarded when > constr_mult_init_max = 1e3 */
    if (p->flags & SE3O_FLAG_LSINIT) {
        rhs_t *r = &s->rhs;
        memset(r, 0, sizeof(*r));
        for (int k = 1; k <= N; k++)
            for (int i = 0; i < NX; i++) r->gx[k][i] = s->sf * s->ogx[k][i];   /* -zL+zU cancels (both 1) */
        for (int k = 0; k < N; k++)
            for (int j = 0; j < NU; j++) r->gu[k][j] = s->sf * s->ogu[k][j];
        int ok = riccati(s, HESS_IDENT, 1, 0.0, r, s->dx, s->du, s->lamp);
        double lmax = 0;
        for (int k = 0; k < N; k++)
            for (int i = 0; i < NX; i++) lmax = fmax(lmax, fabs(s->lamp[k][i]));
        if (ok && lmax <= 1e3 && isfinite(lmax)) memcpy(s->lam, s->lamp, sizeof(double) * NX * N);
        s->nfact = 0;
    }

    double mu = p->mu_init, tau = fmax(0.99, 1.0 - mu);
    const double mu_min = fmin(p->tol, s->sf * 1e-4) / 11.0;
    fentry filt[FILT_MAX];
    int nfilt = 0;
    double theta0 = theta_of(s, s->c);
    const double theta_max = 1e4 * fmax(1.0, theta0), theta_min = 1e-4 * fmax(1.0, theta0);
    const double gamma_theta = 1e-5, gamma_phi = 1e-8, s_phi = 2.3, s_theta = 1.1, eta_phi = 1e-8, kappa_soc = 0.99;
    double delta_last = 0.0;
    int acc_count = 0, status = 2, iter = 0;
    int nls = 0, nsoc = 0;
    int in_soft_resto = 0, soft_resto_counter = 0, nsoft = 0;
    int consec_short = 0, max_consec_short = 0, nresto = 0;       /* IPOPT's soft restoration phase (IpBacktrackingLineSearch.cpp) */
    /* IPOPT's watchdog (IpBacktrackingLineSearch.cpp StartWatchDog / StopWatchDog; FilterLSAcceptor reference values): after
     * watchdog_shortened_iter_trigger = 10 consecutive iterations whose accepted step needed backtracking, the full fraction-to-the-
     * boundary step is taken WITHOUT an acceptance test for up to watchdog_trial_iter_max = 3 iterations; every trial point is
     * tested against the point where the watchdog started (reference theta, phi, grad(phi)^T d).  If none is acceptable the
     * iterate and the search direction of the starting point are restored and the backtracking resumes at alpha_max / 2. */
    /* IPOPT's filter reset heuristic (filter_reset_trigger = 5, max_filter_resets = 5; IpFilterLSAcceptor.cpp): when in five
     * successive iterations the last rejected trial step was rejected because of the filter, the filter is cleared (at most five
     * times per solve).  Switch: SE3O_FLAG_FILTER_RESET. */
    int rej_filter = 0, n_rej_filter = 0, n_filter_resets = 0;
    int n_first_ok = 0;          /* iterations whose delta_w = 0 factorisation already had the right inertia (statistics) */
    int in_watchdog = 0, wd_short = 0, wd_trial = 0, nwatch = 0;
    double wd_theta = 0, wd_phi = 0, wd_gphid = 0, wd_fcur = 0, wd_delta = 0;
    ipm_t *wd_save = NULL;
    double fcur = nlp_objective(s, s->x, s->u, NULL, NULL);

    for (iter = 0;; iter++) {
        double dinf, cv, cmp;
        double E0 = kkt_error(s, 0.0, &dinf, &cv, &cmp);
        sol->kkt_error = E0; sol->dual_inf = dinf / s->sf; sol->constr_viol = cv; sol->compl_inf = cmp / s->sf;
        if (getenv("SE3O_VERBOSE"))
            fprintf(stderr, "it %3d f=%.10e E0=%.3e dinf=%.3e cv=%.3e cmp=%.3e mu=%.2e dl=%.1e th=%.3e\n", iter, fcur, E0, dinf, cv,
                    cmp, mu, delta_last, theta_of(s, s->c));
        if (!isfinite(E0)) { status = 4; break; }
        if (E0 <= p->tol && dinf / s->sf <= 1.0 && cv <= 1e-4 && cmp / s->sf <= 1e-4) { status = 0; break; }
        if (E0 <= p->acceptable_tol && dinf / s->sf <= 1e10 && cv <= 1e-2 && cmp / s->sf <= 1e-2) acc_count++;
        else acc_count = 0;
        if (acc_count >= p->acceptable_iter) { status = 1; break; }
        if (iter >= p->max_iter) { status = 2; break; }

        /* monotone barrier update (mu_allow_fast_monotone_decrease) */
        for (;;) {
            double Emu = kkt_error(s, mu, NULL, NULL, NULL);
            if (!(Emu <= 10.0 * mu) || mu <= mu_min) break;
            double nm = fmax(mu_min, fmin(0.2 * mu, pow(mu, 1.5)));
            if (nm == mu) break;
            mu = nm; tau = fmax(0.99, 1.0 - mu);
            nfilt = 0;
            in_watchdog = 0; wd_short = 0;                 /* BacktrackingLineSearch::Reset(): the watchdog is switched off, the iterate stays */
            rej_filter = 0; n_rej_filter = 0;               /* FilterLSAcceptor::Reset() */
        }

        /* search direction with inertia correction */
        build_rhs(s, mu, &s->rhs);
        double delta = 0.0;
        int ok = 0;
        for (;;) {
            ok = riccati(s, HESS_EXACT, 1, delta, &s->rhs, s->dx, s->du, s->lamp);
            if (ok && !all_finite(&s->dx[0][0], NX * (N + 1))) ok = 0;
            if (ok) break;
            if (delta == 0.0) delta = (delta_last == 0.0) ? 1e-4 : fmax(1e-20, delta_last / 3.0);
            else delta = (delta_last == 0.0 || 1e5 * delta_last < delta) ? 100.0 * delta : 8.0 * delta;
            if (delta > 1e20) break;
        }
        if (!ok) { status = 3; break; }
        if (delta > 0) delta_last = delta;
        else n_first_ok++;
        compute_dz(s, mu, s->dx, s->du);

        double alpha_max = frac_to_bound_primal(s, s->dx, s->du, tau);
        double alpha_z = frac_to_bound_dual(s, tau);

        /* filter line search */
        double theta_k = theta_of(s, s->c);
        double phi_k = barrier_obj(s, s->x, s->u, mu, fcur);
        double gphid = 0;
        for (int k = 1; k <= N; k++)
            for (int i = 0; i < NX; i++) gphid += s->rhs.gx[k][i] * s->dx[k][i];
        for (int k = 0; k < N; k++)
            for (int j = 0; j < NU; j++)
                gphid += (s->rhs.gu[k][j] + (k + 1 < N ? s->rhs.gub[k + 1][j] : 0.0)) * s->du[k][j];
        double alpha_min;
        {
            double am = gamma_theta;
            if (gphid < 0) {
                am = fmin(am, gamma_phi * theta_k / (-gphid));
                if (theta_k <= theta_min) am = fmin(am, pow(theta_k, s_theta) / pow(-gphid, s_phi));
            }
            alpha_min = 0.05 * am;
        }
        /* wake the watchdog up? */
        if ((p->flags & SE3O_FLAG_WATCHDOG) && !in_watchdog && !in_soft_resto && wd_short >= 10) {
            in_watchdog = 1; wd_trial = 0; nwatch++;
            if (!wd_save) wd_save = (ipm_t *)malloc(sizeof(ipm_t));
            memcpy(wd_save, s, sizeof(ipm_t));
            wd_theta = theta_k; wd_phi = phi_k; wd_gphid = gphid; wd_fcur = fcur; wd_delta = delta;
        }
        int forced = 0, skip_first = 0, n_steps = 0;
    line_search:
        if (in_watchdog) { theta_k = wd_theta; phi_k = wd_phi; gphid = wd_gphid; }      /* InitThisLineSearch(in_watchdog): reference values */
        double alpha = skip_first ? 0.5 * alpha_max : alpha_max;
        int accepted = 0, first = 1;
        double (*ddx)[NX] = s->dx, (*ddu)[NU] = s->du;
        double alpha_test = alpha;
        double ftrial = 0;
        n_steps = 0;
        while (!accepted && !in_soft_resto) {
            nls++;
            for (int k = 0; k <= N; k++)
                for (int i = 0; i < NX; i++) s->xt[k][i] = s->x[k][i] + alpha * s->dx[k][i];
            for (int k = 0; k < N; k++)
                for (int j = 0; j < NU; j++) s->ut[k][j] = s->u[k][j] + alpha * s->du[k][j];
            double ct[NMAX][NX];
            nlp_constraints(s, s->xt, s->ut, ct);
            double theta_t = theta_of(s, ct);
            ftrial = nlp_objective(s, s->xt, s->ut, NULL, NULL);
            double phi_t = barrier_obj(s, s->xt, s->ut, mu, ftrial);
            alpha_test = alpha;
#define ACCEPTABLE(theta_t_, phi_t_, ok_) do {                                                              \
                int a_ = 1;                                                                                  \
                if (!isfinite(theta_t_) || !isfinite(phi_t_)) a_ = 0;                                        \
                else if (theta_t_ > theta_max) a_ = 0;                                                       \
                else {                                                                                       \
                    int ftype_ = (gphid < 0) && (alpha_test * pow(-gphid, s_phi) > pow(theta_k, s_theta));   \
                    if (ftype_ && theta_k <= theta_min)                                                      \
                        a_ = compare_le(phi_t_ - phi_k, eta_phi * alpha_test * gphid, phi_k);                \
                    else                                                                                     \
                        a_ = compare_le(theta_t_, (1 - gamma_theta) * theta_k, theta_k) ||                   \
                             compare_le(phi_t_ - phi_k, -gamma_phi * theta_k, phi_k);                        \
                    if (!a_) rej_filter = 0;                   /* rejected by the test against the current iterate */    \
                    for (int e_ = 0; a_ && e_ < nfilt; e_++)                                                 \
                        if (!(compare_le(theta_t_, filt[e_].theta, filt[e_].theta) ||                        \
                              compare_le(phi_t_, filt[e_].phi, filt[e_].phi))) { a_ = 0; rej_filter = 1; }   \
                }                                                                                            \
                ok_ = a_;                                                                                    \
            } while (0)
            ACCEPTABLE(theta_t, phi_t, accepted);
            if (getenv("SE3O_VERBOSE4"))
                fprintf(stderr, "      trial alpha=%.3e theta_t=%.6e (theta_k %.6e) phi_t=%.8e (phi_k %.8e) gphid=%.3e alpha_min=%.3e acc=%d\n", alpha, theta_t,
                        theta_k, phi_t, phi_k, gphid, alpha_min, accepted);
            if (accepted) break;
            if (in_watchdog) break;                        /* only the full step is tried */
            /* second-order correction (first trial only, constraint violation not reduced) */
            if (first && !skip_first && p->max_soc > 0 && theta_t >= theta_k) {
                rhs_t *r2 = (rhs_t *)malloc(sizeof(rhs_t));
                memcpy(r2, &s->rhs, sizeof(rhs_t));
                double (*sdx)[NX] = (double (*)[NX])malloc(sizeof(double) * NX * (NMAX + 1));
                double (*sdu)[NU] = (double (*)[NU])malloc(sizeof(double) * NU * NMAX);
                double (*slam)[NX] = (double (*)[NX])malloc(sizeof(double) * NX * NMAX);
                double theta_soc_old = theta_k, theta_tr = theta_t;
                double alpha_soc = alpha;
                for (int k = 0; k < N; k++)
                    for (int i = 0; i < NX; i++) r2->c[k][i] = alpha * s->c[k][i] + ct[k][i];
                for (int it = 0; it < p->max_soc && !accepted; it++) {
                    if (it > 0 && !(theta_tr <= kappa_soc * theta_soc_old)) break;
                    theta_soc_old = theta_tr;
                    nsoc++;
                    riccati(s, HESS_EXACT, 0, delta, r2, sdx, sdu, slam);
                    alpha_soc = frac_to_bound_primal(s, sdx, sdu, tau);
                    for (int k = 0; k <= N; k++)
                        for (int i = 0; i < NX; i++) s->xt[k][i] = s->x[k][i] + alpha_soc * sdx[k][i];
                    for (int k = 0; k < N; k++)
                        for (int j = 0; j < NU; j++) s->ut[k][j] = s->u[k][j] + alpha_soc * sdu[k][j];
                    double ct2[NMAX][NX];
                    nlp_constraints(s, s->xt, s->ut, ct2);
                    theta_tr = theta_of(s, ct2);
                    double f2 = nlp_objective(s, s->xt, s->ut, NULL, NULL);
                    double phi2 = barrier_obj(s, s->xt, s->ut, mu, f2);
                    int okk;
                    ACCEPTABLE(theta_tr, phi2, okk);
                    if (okk) {
                        accepted = 1;
                        ftrial = f2;
                        memcpy(s->dx, sdx, sizeof(double) * NX * (N + 1));
                        memcpy(s->du, sdu, sizeof(double) * NU * N);
                        memcpy(s->lamp, slam, sizeof(double) * NX * N);
                        alpha = alpha_soc;
                        compute_dz(s, mu, s->dx, s->du);
                        alpha_z = frac_to_bound_dual(s, tau);
                    } else {
                        for (int k = 0; k < N; k++)
                            for (int i = 0; i < NX; i++) r2->c[k][i] = alpha_soc * r2->c[k][i] + ct2[k][i];
                    }
                }
                free(r2); free(sdx); free(sdu); free(slam);
                if (accepted) break;
            }
            first = 0;
            alpha *= 0.5;
            n_steps++;
            if (alpha < alpha_min) break;
        }
        (void)ddx; (void)ddu;
        if (in_watchdog) {
            if (accepted) in_watchdog = 0;                  /* "W": the watchdog procedure was successful */
            else if (++wd_trial > 3) {
                /* StopWatchDog: back to the stored iterate and direction, backtracking resumes below the full step */
                const int nfact_keep = s->nfact;
                memcpy(s, wd_save, sizeof(ipm_t));
                s->nfact = nfact_keep;
                in_watchdog = 0; wd_short = 0;
                fcur = wd_fcur; delta = wd_delta;
                theta_k = wd_theta; phi_k = wd_phi; gphid = wd_gphid;
                alpha_max = frac_to_bound_primal(s, s->dx, s->du, tau);
                alpha_z = frac_to_bound_dual(s, tau);
                {
                    double am = gamma_theta;
                    if (gphid < 0) {
                        am = fmin(am, gamma_phi * theta_k / (-gphid));
                        if (theta_k <= theta_min) am = fmin(am, pow(theta_k, s_theta) / pow(-gphid, s_phi));
                    }
                    alpha_min = 0.05 * am;
                }
                skip_first = 1;
                goto line_search;
            } else { forced = 1; accepted = 1; }             /* "w": take the step anyway */
        }
        if (!accepted && (p->flags & SE3O_FLAG_SOFT_RESTO)) {
            /* The backtracking line search ran below alpha_min (or the soft restoration phase is already active): IPOPT tries a
             * "soft restoration" step before it calls the restoration phase proper (TrySoftRestoStep): the ORIGINAL direction with
             * ONE step size for primal and dual variables, alpha = min(alpha_primal_max, alpha_dual_max) of the fraction-to-the-
             * boundary rule; accepted if the point is acceptable to the original filter criterion (then the phase ends) or if it
             * reduces the primal-dual system error by soft_resto_pderror_reduction_factor = 0.9999; at most max_soft_resto_iters =
             * 10 consecutive such steps.  The filter is not augmented by soft steps. */
            int try_it = 1;
            if (in_soft_resto) { soft_resto_counter++; if (soft_resto_counter > 10) try_it = 0; }
            if (try_it) {
                compute_dz(s, mu, s->dx, s->du);                     /* (a rejected SOC may have left dz of its own direction) */
                const double a_s = fmin(frac_to_bound_primal(s, s->dx, s->du, tau), frac_to_bound_dual(s, tau));
                const double pd_cur = pd_system_error(s, mu);
                ipm_t *bk = (ipm_t *)malloc(sizeof(ipm_t));
                memcpy(bk, s, sizeof(ipm_t));
                for (int k = 1; k <= N; k++)
                    for (int i = 0; i < NX; i++) s->x[k][i] += a_s * s->dx[k][i];
                for (int k = 0; k < N; k++) {
                    for (int j = 0; j < NU; j++) {
                        s->u[k][j] += a_s * s->du[k][j];
                        s->zLu[k][j] += a_s * s->dzLu[k][j]; s->zUu[k][j] += a_s * s->dzUu[k][j];
                    }
                    for (int i = 0; i < NX; i++) s->lam[k][i] += a_s * (s->lamp[k][i] - s->lam[k][i]);
                }
                for (int k = 1; k <= N; k++)
                    for (int j = 0; j < 3; j++) { s->zLw[k][j] += a_s * s->dzLw[k][j]; s->zUw[k][j] += a_s * s->dzUw[k][j]; }
                eval_point(s);
                nls++;
                const double theta_t = theta_of(s, s->c);
                const double f_t = nlp_objective(s, s->x, s->u, NULL, NULL);
                const double phi_t = barrier_obj(s, s->x, s->u, mu, f_t);
                int orig_ok = 0;
                alpha_test = 0.0;                                     /* CheckAcceptabilityOfTrialPoint(0.) */
                ACCEPTABLE(theta_t, phi_t, orig_ok);
                int soft_ok = orig_ok;
                if (!soft_ok && isfinite(theta_t) && isfinite(phi_t)) {
                    const double pd_t = pd_system_error(s, mu);
                    soft_ok = isfinite(pd_t) && pd_t <= 0.9999 * pd_cur;
                }
                if (soft_ok) {
                    /* kappa_sigma safeguard of the bound multipliers, as after every accepted step */
                    const double ks = 1e10;
#define ZFIX(z, slack) do { double sl_ = (slack); (z) = fmax(fmin((z), ks * mu / sl_), mu / (ks * sl_)); } while (0)
                    for (int k = 0; k < N; k++)
                        for (int j = 0; j < NU; j++) { ZFIX(s->zLu[k][j], s->u[k][j] - s->uL); ZFIX(s->zUu[k][j], s->uU - s->u[k][j]); }
                    for (int k = 1; k <= N; k++)
                        for (int j = 0; j < 3; j++) { ZFIX(s->zLw[k][j], s->x[k][10 + j] - s->wL); ZFIX(s->zUw[k][j], s->wU - s->x[k][10 + j]); }
#undef ZFIX
                    if (orig_ok) { in_soft_resto = 0; soft_resto_counter = 0; }
                    else if (!in_soft_resto) { in_soft_resto = 1; soft_resto_counter = 0; }
                    fcur = f_t;
                    nsoft++;
                    free(bk);
                    if (getenv("SE3O_VERBOSE3")) fprintf(stderr, "   soft restoration step alpha=%.3e orig_ok=%d\n", a_s, orig_ok);
                    continue;
                }
                const int nfact_keep = s->nfact;
                memcpy(s, bk, sizeof(ipm_t));
                s->nfact = nfact_keep;
                free(bk);
            }
        }
        if (!accepted && (p->flags & SE3O_FLAG_RESTO)) {
            /* feasibility restoration phase; the current point is added to the original filter first (IpFilterLSAcceptor::PrepareRestoPhaseStart) */
            {
                fentry e = {(1 - gamma_theta) * theta_k, phi_k - gamma_phi * theta_k};
                int w_ = 0;
                for (int i = 0; i < nfilt; i++)
                    if (!(filt[i].theta >= e.theta && filt[i].phi >= e.phi)) filt[w_++] = filt[i];
                nfilt = w_;
                if (nfilt < FILT_MAX) filt[nfilt++] = e;
            }
            in_soft_resto = 0; soft_resto_counter = 0;
            double f_new = fcur;
            iter++;
            if (restoration_phase(s, mu, tau, filt, nfilt, theta_max, &iter, p->max_iter, &nls, &f_new)) {
                fcur = f_new;
                nresto++;
                iter--;                                  /* (the for-loop increments) */
                continue;
            }
            iter--;
        }
        if (!accepted) { status = 3; break; }          /* restoration failed (or is switched off) */
        if (getenv("SE3O_VERBOSE3"))
            fprintf(stderr, "   ls: alpha_max=%.3e alpha=%.3e alpha_z=%.3e ntrial_total=%d gphid=%.3e theta_k=%.3e theta_t? nfilt=%d\n", alpha_max, alpha, alpha_z, nls, gphid, theta_k, nfilt);

        if (alpha < alpha_max) { consec_short++; if (consec_short > max_consec_short) max_consec_short = consec_short; } else consec_short = 0;
        if (n_steps == 0) wd_short = 0; else wd_short++;         /* watchdog_shortened_iter_ (a step accepted at the first trial, SOC included, resets it) */
        if ((p->flags & SE3O_FLAG_FILTER_RESET) && !forced) {
            if (rej_filter) {
                n_rej_filter++;
                if (n_filter_resets < 5 && n_rej_filter >= 5) { nfilt = 0; n_filter_resets++; n_rej_filter = 0; rej_filter = 0; }
            } else
                n_rej_filter = 0;
        }
        /* filter augmentation (uses the original alpha_test / direction); a forced watchdog step leaves the filter alone */
        if (!forced) {
            int ftype = (gphid < 0) && (alpha_test * pow(-gphid, s_phi) > pow(theta_k, s_theta));
            double phi_new = barrier_obj(s, s->xt, s->ut, mu, ftrial);
            int armijo = compare_le(phi_new - phi_k, eta_phi * alpha_test * gphid, phi_k);
            if (!ftype || !armijo) {
                fentry e = {(1 - gamma_theta) * theta_k, phi_k - gamma_phi * theta_k};
                int w_ = 0;
                for (int i = 0; i < nfilt; i++)   /* drop dominated entries */
                    if (!(filt[i].theta >= e.theta && filt[i].phi >= e.phi)) filt[w_++] = filt[i];
                nfilt = w_;
                if (nfilt < FILT_MAX) filt[nfilt++] = e;
            }
        }
        /* accept: primal with alpha, lambda with alpha (alpha_for_y=primal), z with alpha_z */
        memcpy(s->x, s->xt, sizeof(double) * NX * (N + 1));
        memcpy(s->u, s->ut, sizeof(double) * NU * N);
        for (int k = 0; k < N; k++)
            for (int i = 0; i < NX; i++) s->lam[k][i] += alpha * (s->lamp[k][i] - s->lam[k][i]);
        const double ks = 1e10;
#define ZUPD(z, dz, slack) do { double zn = (z) + alpha_z * (dz); double sl_ = (slack);                       \
            zn = fmax(fmin(zn, ks * mu / sl_), mu / (ks * sl_)); (z) = zn; } while (0)
        for (int k = 0; k < N; k++)
            for (int j = 0; j < NU; j++) {
                ZUPD(s->zLu[k][j], s->dzLu[k][j], s->u[k][j] - s->uL);
                ZUPD(s->zUu[k][j], s->dzUu[k][j], s->uU - s->u[k][j]);
            }
        for (int k = 1; k <= N; k++)
            for (int j = 0; j < 3; j++) {
                ZUPD(s->zLw[k][j], s->dzLw[k][j], s->x[k][10 + j] - s->wL);
                ZUPD(s->zUw[k][j], s->dzUw[k][j], s->wU - s->x[k][10 + j]);
            }
#undef ZUPD
        fcur = ftrial;
        eval_point(s);
    }

    /* outputs, quad_OC.py:175-210.  honor_original_bounds: project onto the un-relaxed box */
    sol->cost = fcur;
    for (int k = 0; k <= N; k++)
        for (int i = 0; i < NX; i++) {
            double v = s->x[k][i];
            if (k >= 1 && i >= 10) v = fmin(fmax(v, -p->w_bound), p->w_bound);
            sol->X[k * NX + i] = v;
            sol->XRAW[k * NX + i] = s->x[k][i];
        }
    for (int k = 0; k < N; k++) {
        for (int j = 0; j < NU; j++) sol->U[k * NU + j] = fmin(fmax(s->u[k][j], p->u_lb), p->u_ub);
        for (int j = 0; j < NU; j++) sol->URAW[k * NU + j] = s->u[k][j];
        for (int i = 0; i < NX; i++) sol->LAM[k * NX + i] = s->lam[k][i] / s->sf;   /* unscaled lam_g */
    }
    for (int k = 0; k <= N; k++)
        for (int j = 0; j < 3 && k >= 1; j++) { sol->ZL[k * 17 + 10 + j] = s->zLw[k][j]; sol->ZU[k * 17 + 10 + j] = s->zUw[k][j]; }
    for (int k = 0; k < N; k++)
        for (int j = 0; j < NU; j++) { sol->ZL[k * 17 + NX + j] = s->zLu[k][j]; sol->ZU[k * 17 + NX + j] = s->zUu[k][j]; }
    sol->status = status; sol->iters = iter; sol->nfact = s->nfact; sol->nls = nls; sol->nsoc = nsoc; sol->nresto = getenv("SE3O_STAT") ? max_consec_short : nresto + nsoft;
    sol->mu = mu;
    if (getenv("SE3O_STAT_WD")) sol->nresto = nwatch;
    if (getenv("SE3O_STAT_FIRST_OK")) sol->nresto = n_first_ok;
    free(wd_save);
    free(s);
    return status;
}

/* single Newton step for the dense cross-check (tests) */
int se3o_kkt_step(const se3o_params *p, const se3o_problem *pb, const double *w, const double *lam,
                  const double *zL, const double *zU, double mu, double delta_w, double obj_factor,
                  double *dw, double *lam_plus)
{
    const int N = p->N, NZ = NX + NU;
    ipm_t *s = (ipm_t *)calloc(1, sizeof(ipm_t));
    s->p = p; s->pb = pb; s->N = N; s->sf = obj_factor;
    att_form(pb->tra_q, s->Nq, &s->s0);
    for (int k = 0; k < N; k++) s->gam[k] = gamma_k(p, pb, k);
    s->uL = p->u_lb - 1e-8 * fmax(1.0, fabs(p->u_lb));
    s->uU = p->u_ub + 1e-8 * fmax(1.0, fabs(p->u_ub));
    s->wL = -p->w_bound - 1e-8 * fmax(1.0, fabs(p->w_bound));
    s->wU = p->w_bound + 1e-8 * fmax(1.0, fabs(p->w_bound));
    for (int k = 0; k <= N; k++) {
        memcpy(s->x[k], w + k * NZ, sizeof(double) * NX);
        for (int j = 0; j < 3; j++) { s->zLw[k][j] = zL[k * NZ + 10 + j]; s->zUw[k][j] = zU[k * NZ + 10 + j]; }
    }
    for (int k = 0; k < N; k++) {
        memcpy(s->u[k], w + k * NZ + NX, sizeof(double) * NU);
        memcpy(s->lam[k], lam + k * NX, sizeof(double) * NX);
        for (int j = 0; j < NU; j++) { s->zLu[k][j] = zL[k * NZ + NX + j]; s->zUu[k][j] = zU[k * NZ + NX + j]; }
    }
    eval_point(s);
    build_rhs(s, mu, &s->rhs);
    int ok = riccati(s, HESS_EXACT, 1, delta_w, &s->rhs, s->dx, s->du, s->lamp);
    for (int k = 0; k <= N; k++) memcpy(dw + k * NZ, s->dx[k], sizeof(double) * NX);
    for (int k = 0; k < N; k++) {
        memcpy(dw + k * NZ + NX, s->du[k], sizeof(double) * NU);
        memcpy(lam_plus + k * NX, s->lamp[k], sizeof(double) * NX);
    }
    free(s);
    return ok;
}

/* ------------------------------------------------------------------------------------------------ */
/* rotor-tip kinematics and reward: quad_model.py:239-276, solid_geometry.py, quad_policy.py:78-91   */
/* ------------------------------------------------------------------------------------------------ */
void se3o_rotor_positions(const se3o_params *p, const double *X, int nrow, double *pos)
{
    double a = p->wing_len * 0.5 / sqrt(2.0);
    const double rb[4][3] = {{a, a, 0}, {-a, a, 0}, {-a, -a, 0}, {a, -a, 0}};   /* quad_model.py:242-245 */
    for (int t = 0; t < nrow; t++) {
        const double *x = X + t * NX;
        double C[3][3];
        dir_cosine(x + 6, C);                          /* C_B_I; body->inertial is its transpose */
        for (int i = 0; i < 3; i++) pos[t * 15 + i] = x[i];
        for (int r = 0; r < 4; r++)
            for (int i = 0; i < 3; i++) {
                double v = 0;
                for (int j = 0; j < 3; j++) v += C[j][i] * rb[r][j];
                pos[t * 15 + 3 * (r + 1) + i] = x[i] + v;
            }
    }
}

static double dot3(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static void sub3(const double *a, const double *b, double *c) { c[0] = a[0] - b[0]; c[1] = a[1] - b[1]; c[2] = a[2] - b[2]; }
static void cross3(const double *a, const double *b, double *c)
{
    c[0] = a[1] * b[2] - a[2] * b[1]; c[1] = a[2] * b[0] - a[0] * b[2]; c[2] = a[0] * b[1] - a[1] * b[0];
}
static double magni3(const double *a) { return sqrt(dot3(a, a)); }          /* solid_geometry.py:7-8 */
static void norm3(const double *a, double *c) { double m = magni3(a); c[0] = a[0] / m; c[1] = a[1] / m; c[2] = a[2] / m; }

/* line.vertical / line.distance, solid_geometry.py:57-78 */
static double line_vertical(const double *pa, const double *pb, const double *X)
{
    double d[3], dir[3], v[3], n[3];
    sub3(pa, pb, d); norm3(d, dir);
    sub3(X, pa, v); cross3(v, dir, n);
    return magni3(n);
}
static double line_distance(const double *pa, const double *pb, const double *X)
{
    double a = line_vertical(pa, pb, X), v[3];
    sub3(X, pa, v); double b = magni3(v);
    sub3(X, pb, v); double c = magni3(v);
    sub3(pa, pb, v); double d = magni3(v);
    if (b > c) return ((b * b - d * d) > a * a) ? c : a;
    return ((c * c - d * d) > a * a) ? b : a;
}

double se3o_collis_det(const double gate[12], const double *V, int stride, int horizon, int *co_out)
{
    const double *P[4] = {gate, gate + 3, gate + 6, gate + 9};
    double cen[3];
    for (int i = 0; i < 3; i++) cen[i] = (P[0][i] + P[1][i] + P[2][i] + P[3][i]) / 4;   /* solid_geometry.py:89-90 */
    /* plane_i = plane(centroid, P_i, P_{i+1}): solid_geometry.py:16-40,93-96 */
    double nrm[4][3], n1[4][3], n2[4][3], n3[4][3];
    for (int i = 0; i < 4; i++) {
        const double *pa = P[i], *pb = P[(i + 1) % 4];
        double v1[3], v2[3], v3[3], t[3];
        sub3(pa, cen, v1); sub3(pb, cen, v2);
        cross3(v2, v1, t); norm3(t, nrm[i]);
        cross3(v1, nrm[i], t); norm3(t, n1[i]);
        cross3(nrm[i], v2, t); norm3(t, n2[i]);
        sub3(pb, pa, v3);
        cross3(nrm[i], v3, t); norm3(t, n3[i]);
    }
    double collision = 0, d_min = 0.2, rel[3];
    int co = 0;
    *co_out = 0;
    sub3(V, cen, rel);
    if (dot3(nrm[0], rel) < 0) return 0;                                            /* solid_geometry.py:110-111 */
    for (int t = 0; t < horizon; t++) {
        const double *vt = V + (size_t)t * stride;
        sub3(vt, cen, rel);
        if (dot3(nrm[0], rel) < 0) {
            const double *vp = V + (size_t)(t - 1) * stride;
            /* plane.interpoint, solid_geometry.py:43-47 */
            double d[3], dir[3], X[3];
            sub3(vt, vp, d); norm3(d, dir);
            double tt = 1.0 / dot3(dir, nrm[0]) * dot3(nrm[0], rel);
            for (int i = 0; i < 3; i++) X[i] = vt[i] - tt * dir[i];
            double xr[3];
            sub3(X, cen, xr);
            for (int i = 0; i < 4; i++) {                                           /* sequential, non-exclusive */
                if (dot3(n1[i], xr) > 0 && dot3(n2[i], xr) > 0) {
                    double pv[3];
                    sub3(P[i], X, pv);
                    if (dot3(pv, n3[i]) > 0) {
                        double m = line_vertical(P[0], P[1], X);
                        m = fmin(m, line_vertical(P[1], P[2], X));
                        m = fmin(m, line_vertical(P[2], P[3], X));
                        m = fmin(m, line_vertical(P[3], P[0], X));
                        double e = fmax(0.0, d_min - m);
                        collision = -(e * e);
                        co = 1;
                    } else {
                        int a = (i + 3) % 4, b = i, c = (i + 1) % 4;                /* lines i-1, i, i+1 */
                        double m = line_distance(P[a], P[(a + 1) % 4], X);
                        m = fmin(m, line_distance(P[b], P[(b + 1) % 4], X));
                        m = fmin(m, line_distance(P[c], P[(c + 1) % 4], X));
                        collision = -2 * d_min * m - d_min * d_min;
                    }
                }
            }
            break;
        }
    }
    *co_out = co;
    return collision;
}

double se3o_reward(const se3o_params *p, const double *X, const double gate[12], const double goal[3],
                   double *collision_out, double *path_out, int *co_out)
{
    const int N = p->N;
    double pos[(NMAX + 1) * 15];
    se3o_rotor_positions(p, X, N + 1, pos);
    double collision = 0, path = 0;
    int co = 0;
    for (int c = 0; c < 4; c++) {                                                   /* quad_policy.py:85-87 */
        int coc;
        collision += se3o_collis_det(gate, pos + 3 * (c + 1), 15, N, &coc);
        co += coc;
    }
    for (int q = 0; q < 4; q++) {                                                   /* quad_policy.py:88-89 */
        const int idx = (N - 1 - q >= 0) ? N - 1 - q : 2 * N - q;                    /* negative numpy index wraps (N < 4) */
        const double *r = pos + idx * 15;
        double d[3];
        sub3(r, goal, d);
        path += dot3(d, d);
    }
    if (collision_out) *collision_out = collision;
    if (path_out) *path_out = path;
    if (co_out) *co_out = co;
    return 1000 * collision - 0.5 * path + 100;                                     /* quad_policy.py:90 */
}

/* ------------------------------------------------------------------------------------------------ */
/* planner: quad_policy.py:67-112                                                                    */
/* ------------------------------------------------------------------------------------------------ */
double se3o_objective(const se3o_params *p, const double x0[13], const double goal[3], const double gate[12],
                      const double tra_pos[3], const double tra_ang[3], double t, const double *ulast,
                      se3o_solution *sol)
{
    se3o_problem pb;
    memcpy(pb.x0, x0, sizeof(pb.x0));
    memcpy(pb.goal, goal, sizeof(pb.goal));
    memcpy(pb.tra_pos, tra_pos, sizeof(pb.tra_pos));
    se3o_rd2quat(tra_ang, pb.tra_q);
    pb.t_tra = se3o_round1(t);                                                      /* quad_policy.py:70 */
    for (int j = 0; j < 4; j++) pb.ulast[j] = ulast ? ulast[j] : 0.0;               /* quad_OC.py:130-133 */
    se3o_solution local, *s = sol ? sol : &local;
    se3o_solve(p, &pb, s);
    return se3o_reward(p, s->X, gate, goal, NULL, NULL, NULL);
}

static double clip(double v, double lo, double hi) { return v < lo ? lo : (v > hi ? hi : v); }

void se3o_sol_gradient(const se3o_params *p, const double x0[13], const double goal[3], const double gate[12],
                       const double tra_pos[3], const double tra_ang[3], double t, const double *ulast,
                       double out[8], int status9[9])
{
    const double delta = 1e-3;
    se3o_solution sol;
    double j = se3o_objective(p, x0, goal, gate, tra_pos, tra_ang, t, NULL, &sol);  /* quad_policy.py:97: no Ulast */
    if (status9) status9[0] = sol.status;
    double dr[6];
    for (int i = 0; i < 6; i++) {
        double tp[3] = {tra_pos[0], tra_pos[1], tra_pos[2]}, ta[3] = {tra_ang[0], tra_ang[1], tra_ang[2]};
        if (i < 3) tp[i] += delta; else ta[i - 3] += delta;
        double v = se3o_objective(p, x0, goal, gate, tp, ta, t, ulast, &sol);
        if (status9) status9[1 + i] = sol.status;
        double scale = (i < 3) ? 0.1 : 1.0 / (500 * tra_ang[i - 3] * tra_ang[i - 3] + 5);
        dr[i] = clip(v - j, -0.5, 0.5) * scale;
    }
    double drdt = 0;
    double jm = se3o_objective(p, x0, goal, gate, tra_pos, tra_ang, t - 0.1, NULL, &sol);
    if (status9) status9[7] = sol.status;
    if (jm - j > 2) drdt = -0.05;
    double jp = se3o_objective(p, x0, goal, gate, tra_pos, tra_ang, t + 0.1, NULL, &sol);
    if (status9) status9[8] = sol.status;
    if (jp - j > 2) drdt = 0.05;
    for (int i = 0; i < 6; i++) out[i] = -dr[i];
    out[6] = -drdt;
    out[7] = j;
}

/* ------------------------------------------------------------------------------------------------ */
/* batch drivers: one problem per worker thread (the reference uses one OS process per sample,       */
/* deep_learning.py:66-72)                                                                           */
/* ------------------------------------------------------------------------------------------------ */
int se3o_num_threads(void)
{
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

typedef struct {
    void (*fn)(int b, void *ctx);
    void *ctx;
    int B;
    atomic_int next;
} pool_t;

static void *pool_worker(void *arg)
{
    pool_t *pl = (pool_t *)arg;
    for (;;) {
        int b = atomic_fetch_add(&pl->next, 1);
        if (b >= pl->B) break;
        pl->fn(b, pl->ctx);
    }
    return NULL;
}

static void parallel_for(int B, int nthreads, void (*fn)(int, void *), void *ctx)
{
    if (nthreads <= 0) nthreads = se3o_num_threads();
    if (nthreads > B) nthreads = B;
    pool_t pl;
    pl.fn = fn; pl.ctx = ctx; pl.B = B;
    atomic_init(&pl.next, 0);
    if (nthreads <= 1) { pool_worker(&pl); return; }
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * nthreads);
    int started = 0;
    for (int i = 0; i < nthreads - 1; i++)
        if (pthread_create(&th[started], NULL, pool_worker, &pl) == 0) started++;
    pool_worker(&pl);
    for (int i = 0; i < started; i++) pthread_join(th[i], NULL);
    free(th);
}

typedef struct {
    const se3o_params *p;
    const double *x0, *goal, *gate, *tra_pos, *tra_ang, *t, *ulast;
    int round_t;
    double *X, *U, *LAM, *cost, *reward;
    int *status, *iters, *nfact;
} batch_ctx;

static void solve_one(int b, void *vctx)
{
    batch_ctx *c = (batch_ctx *)vctx;
    const int N = c->p->N;
    se3o_problem pb;
    se3o_solution *sol = (se3o_solution *)malloc(sizeof(se3o_solution));
    memcpy(pb.x0, c->x0 + b * 13, sizeof(pb.x0));
    memcpy(pb.goal, c->goal + b * 3, sizeof(pb.goal));
    memcpy(pb.tra_pos, c->tra_pos + b * 3, sizeof(pb.tra_pos));
    se3o_rd2quat(c->tra_ang + b * 3, pb.tra_q);
    pb.t_tra = c->round_t ? se3o_round1(c->t[b]) : c->t[b];
    for (int j = 0; j < 4; j++) pb.ulast[j] = c->ulast ? c->ulast[b * 4 + j] : 0.0;
    se3o_solve(c->p, &pb, sol);
    if (c->X) memcpy(c->X + (size_t)b * (N + 1) * 13, sol->X, sizeof(double) * (N + 1) * 13);
    if (c->U) memcpy(c->U + (size_t)b * N * 4, sol->U, sizeof(double) * N * 4);
    if (c->LAM) memcpy(c->LAM + (size_t)b * N * 13, sol->LAM, sizeof(double) * N * 13);
    if (c->cost) c->cost[b] = sol->cost;
    if (c->status) c->status[b] = sol->status;
    if (c->iters) c->iters[b] = sol->iters;
    if (c->nfact) c->nfact[b] = sol->nfact;
    free(sol);
}

void se3o_solve_batch(const se3o_params *p, int B, const double *x0, const double *goal, const double *tra_pos,
                      const double *tra_ang, const double *t, const double *ulast, int round_t, int nthreads,
                      double *X, double *U, double *LAM, double *cost, int *status, int *iters, int *nfact)
{
    batch_ctx c;
    memset(&c, 0, sizeof(c));
    c.p = p; c.x0 = x0; c.goal = goal; c.tra_pos = tra_pos; c.tra_ang = tra_ang; c.t = t; c.ulast = ulast;
    c.round_t = round_t; c.X = X; c.U = U; c.LAM = LAM; c.cost = cost; c.status = status; c.iters = iters; c.nfact = nfact;
    parallel_for(B, nthreads, solve_one, &c);
}

static void objective_one(int b, void *vctx)
{
    batch_ctx *c = (batch_ctx *)vctx;
    se3o_solution *sol = (se3o_solution *)malloc(sizeof(se3o_solution));
    c->reward[b] = se3o_objective(c->p, c->x0 + b * 13, c->goal + b * 3, c->gate + b * 12, c->tra_pos + b * 3,
                                  c->tra_ang + b * 3, c->t[b], c->ulast ? c->ulast + b * 4 : NULL, sol);
    if (c->cost) c->cost[b] = sol->cost;
    if (c->status) c->status[b] = sol->status;
    if (c->iters) c->iters[b] = sol->iters;
    free(sol);
}

void se3o_objective_batch(const se3o_params *p, int B, const double *x0, const double *goal, const double *gate,
                          const double *tra_pos, const double *tra_ang, const double *t, const double *ulast,
                          int nthreads, double *reward, double *cost, int *status, int *iters)
{
    batch_ctx c;
    memset(&c, 0, sizeof(c));
    c.p = p; c.x0 = x0; c.goal = goal; c.gate = gate; c.tra_pos = tra_pos; c.tra_ang = tra_ang; c.t = t; c.ulast = ulast;
    c.reward = reward; c.cost = cost; c.status = status; c.iters = iters;
    parallel_for(B, nthreads, objective_one, &c);
}
