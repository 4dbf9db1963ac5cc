// Host warp-emulation harness (TESTS ONLY).  Compiles learningagileflight_se3_b200/csrc/se3mpc_core.cuh with
// -DSE3_EMULATE so the warp-cooperative solver can be exercised on a machine without a GPU (CPU test tier,
// debugging).  Never linked into libse3mpc.so; no product entry point reaches this code.
#define SE3_EMULATE 1
#include "../../learningagileflight_se3_b200/csrc/se3mpc_core.cuh"

#include <stdlib.h>
#include <string.h>

using namespace se3;

static LaneTab g_tab[32];
static void make_ctx(Ctx &c, const se3mpc_params *p)
{
    c.sm = (WarpSmem *)malloc(sizeof(WarpSmem));
    memset(c.sm, 0xFF, sizeof(WarpSmem));      // NaN pattern: shared memory is not zero on the GPU either
    c.ws = (double *)calloc(SLOT_DOUBLES, sizeof(double));
    for (int lane = 0; lane < 32; lane++) make_lane_tab(lane, g_tab[lane]);
    c.lt = g_tab;
}

static void make_instance(const double *x0, const double *goal, const double *tra_pos, const double *tra_ang, double t,
                          const double *ulast, int round_t, Instance &in)
{
    memcpy(in.x0, x0, sizeof(in.x0));
    memcpy(in.goal, goal, sizeof(in.goal));
    memcpy(in.tra_pos, tra_pos, sizeof(in.tra_pos));
    rd2quat(tra_ang, in.tra_q);
    in.t_tra = round_t ? round1(t) : t;
    for (int j = 0; j < 4; j++) in.ulast[j] = ulast ? ulast[j] : 0.0;
    in.xinit = nullptr; in.uinit = nullptr;
}

extern "C" int se3emu_solve(const se3mpc_params *p, const double *x0, const double *goal, const double *tra_pos,
                            const double *tra_ang, double t, const double *ulast, int round_t, double *X, double *U,
                            double *LAM, double *cost, int *stats4, double *kkt_error)
{
    Ctx c;
    make_ctx(c, p);
    Instance in;
    make_instance(x0, goal, tra_pos, tra_ang, t, ulast, round_t, in);
    SolveStats st = ipm_solve<true>(c, *p, in);
    write_outputs(c, X, U, LAM);
    if (cost) *cost = st.cost;
    if (stats4) { stats4[0] = st.status; stats4[1] = st.iters; stats4[2] = st.nfact; stats4[3] = st.nls; }
    if (kkt_error) *kkt_error = st.kkt_error;
    free(c.sm); free(c.ws);
    return st.status;
}

extern "C" double se3emu_reward(const se3mpc_params *p, const double *X, const double *gate, const double *goal,
                                double *collision, double *path, int *co)
{
    return reward_warp(*p, gate, goal, [&](int s, double *x) { for (int i = 0; i < 13; i++) x[i] = X[s * 13 + i]; },
                       collision, path, co);
}

extern "C" double se3emu_round1(double t) { return round1(t); }
extern "C" void se3emu_rd2quat(const double *a, double *q) { rd2quat(a, q); }
extern "C" size_t se3emu_smem_bytes(void) { return sizeof(WarpSmem); }
extern "C" size_t se3emu_slot_bytes(void) { return SLOT_DOUBLES * sizeof(double); }
// state order of the DMMA factor sweep and its list of structurally zero products (tests/test_state_order.py)
extern "C" int se3emu_perm(void) { return SE3_PERM; }
extern "C" int se3emu_xpos(int i) { return xpos<SE3_PERM != 0>(i); }
extern "C" int se3emu_xidx(int p) { return xidx<SE3_PERM != 0>(p); }
extern "C" int se3emu_ab_block(int tile, int ks4) { return ab_block(tile, ks4) ? 1 : 0; }

// one Newton step from a given iterate (reference ordering w=[x0,u0,...,xN]); mirrors se3o_kkt_step
extern "C" int se3emu_kkt_step(const se3mpc_params *p, const double *goal, const double *tra_pos, const double *tra_q,
                               double t_tra, const double *ulast, const double *w, const double *lam, const double *zL,
                               const double *zU, double mu, double delta_w, double obj_factor, double *dw,
                               double *lam_plus)
{
    Ctx c;
    make_ctx(c, p);
    Prob &P = c.sm->P;
    prob_from_params(*p, P);
    const int N = P.N;
    for (int i = 0; i < 3; i++) { P.goal[i] = goal[i]; P.ptra[i] = tra_pos[i]; }
    for (int j = 0; j < 4; j++) P.ulast[j] = ulast ? ulast[j] : 0.0;
    P.ttra = t_tra;
    att_form(tra_q, P.Nq, &P.s0);
    P.sf = obj_factor;
    c.sm->nfact = 0;
    for (int s = 0; s <= N; s++) {
        for (int i = 0; i < 13; i++) row(c, R_X + i)[s] = w[s * 17 + i];
        for (int j = 0; j < 3; j++) { row(c, R_ZLW + j)[s] = zL[s * 17 + 10 + j]; row(c, R_ZUW + j)[s] = zU[s * 17 + 10 + j]; }
        if (s < N) {
            for (int j = 0; j < 4; j++) {
                row(c, R_U + j)[s] = w[s * 17 + 13 + j];
                row(c, R_ZLU + j)[s] = zL[s * 17 + 13 + j];
                row(c, R_ZUU + j)[s] = zU[s * 17 + 13 + j];
            }
            for (int i = 0; i < 13; i++) row(c, R_LAM + i)[s] = lam[s * 13 + i];
            const double d = P.dt * s - P.ttra;
            row(c, R_GAM)[s] = p->gamma_amp * exp(-p->gamma_rate * d * d);
        }
    }
    init_AB(c);
    double th, ci;
    eval_constraints(c, &th, &ci);
    int ok = riccati(c, SE3_MODE_EXACT, 1, 1, delta_w, mu, R_C, R_DX, R_DU, R_LAMP);
    for (int s = 0; s <= N; s++) {
        for (int i = 0; i < 13; i++) dw[s * 17 + i] = row(c, R_DX + i)[s];
        if (s < N) {
            for (int j = 0; j < 4; j++) dw[s * 17 + 13 + j] = row(c, R_DU + j)[s];
            for (int i = 0; i < 13; i++) lam_plus[s * 13 + i] = row(c, R_LAMP + i)[s];
        }
    }
    free(c.sm); free(c.ws);
    return ok;
}
