/*
 * se3mpc.h -- C ABI of the B200-native batched SE(3) MPC engine (libse3mpc.so).
 *
 * Drop-in boundary for ONE hot path of yanrui89/LearningAgileFlight_SE3: thousands of independent
 * gate-traversal MPC solves + the gate-traversal reward.  The reference has no FFI of its own (pure
 * Python over casadi/IPOPT); the seam this library replaces is
 *     quad_OC.py:104-212      OCSys.ocSolver          -> se3mpc_solve
 *     quad_model.py:239-276   get_quadrotor_position  \
 *     solid_geometry.py:104   obstacle.collis_det      > se3mpc_reward
 *     quad_policy.py:81-91    reward assembly         /
 *     quad_policy.py:67-91    run_quad.objective      -> se3mpc_objective   (solve + reward fused)
 *     quad_policy.py:94-112   run_quad.sol_gradient   -> se3mpc_sol_gradient (9 solves / sample, one launch)
 *     quad_policy.py:202-211  run_quad.get_input      -> se3mpc_solve with round_t = 0, U[:,0,:]
 * The Python host side (learningagileflight_se3_b200/quad_OC.py, quad_policy.py) mirrors the
 * reference classes on top of these entry points; INTEGRATION.md shows the ctypes binding.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types cross the boundary.
 *   - `*_device` entry points: every data pointer is a DEVICE pointer owned by the caller; work is
 *     enqueued on `stream` (a cudaStream_t passed as void*) and the call returns without synchronising.
 *   - `*_host` entry points: every data pointer is a HOST pointer; the library stages through its own
 *     device buffers (H2D, kernel, D2H) and returns after the results are in the host buffers.
 *   - batch layouts are problem-major and contiguous: x0[B][13], X[B][N+1][13], U[B][N][4], LAM[B][N][13].
 *   - return value: 0 on success, negative se3mpc_error otherwise; nothing throws across the ABI.
 *   - there is NO CPU implementation behind this ABI: without a CUDA device every compute entry point
 *     returns SE3MPC_ERR_NO_DEVICE.
 */
#ifndef SE3MPC_H
#define SE3MPC_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SE3MPC_NX 13
#define SE3MPC_NU 4
#define SE3MPC_MAX_HORIZON 63

typedef enum {
    SE3MPC_OK = 0,
    SE3MPC_ERR_INVALID = -1,     /* bad argument (NULL pointer, horizon out of range, ...) */
    SE3MPC_ERR_NO_DEVICE = -2,   /* no CUDA device / driver */
    SE3MPC_ERR_CUDA = -3,        /* a CUDA runtime call failed; see se3mpc_last_error() */
    SE3MPC_ERR_WORKSPACE = -4    /* workspace too small */
} se3mpc_error;

/* per-problem solver status (the reference never reads IPOPT's status, quad_OC.py:174-175; we expose it) */
typedef enum {
    SE3MPC_STATUS_CONVERGED = 0,     /* scaled KKT error <= tol (1e-8) */
    SE3MPC_STATUS_ACCEPTABLE = 1,    /* <= acceptable_tol for acceptable_iter iterations */
    SE3MPC_STATUS_MAX_ITER = 2,
    SE3MPC_STATUS_FAILED = 3,        /* line search / inertia correction failed */
    SE3MPC_STATUS_NAN = 4
} se3mpc_status;

/* All literals of the reference's problem definition, one POD.  se3mpc_default_params() fills the
 * values of quad_policy.py:37-56 / quad_OC.py:145,150,170 / quad_model.py:37. */
typedef struct {
    double Jx, Jy, Jz, mass, arm_l, c_tau, grav;             /* quad_policy.py:37, quad_model.py:37 */
    double wrt, wqt, wthrust, wrf, wvf, wwf, wdu;             /* quad_policy.py:38, quad_OC.py:150 */
    double gamma_amp, gamma_rate;                             /* quad_OC.py:145 */
    double dt;                                                /* quad_policy.py:43 */
    double u_lb, u_ub;                                        /* quad_policy.py:50-51 */
    double w_bound;                                           /* quad_policy.py:47 */
    double wing_len;                                          /* quad_policy.py:19 */
    int32_t N;                                                /* horizon (<= SE3MPC_MAX_HORIZON), quad_policy.py:17 */
    int32_t att_power;                                        /* 2 root variant, 1 gym_pybullet_drone variant */
    double tol, acceptable_tol, mu_init;                      /* IPOPT defaults 1e-8, 1e-6, 0.1 */
    int32_t max_iter, acceptable_iter, max_soc, flags;        /* 3000, 15, 4, SCALING|LSINIT */
} se3mpc_params;

#define SE3MPC_FLAG_SCALING 1   /* gradient-based objective scaling (IPOPT nlp_scaling_method default) */
#define SE3MPC_FLAG_LSINIT  2   /* least-squares initial equality multipliers */
#define SE3MPC_FLAG_RESTO   8   /* IPOPT's feasibility restoration phase when the filter line search fails (on by default in IPOPT) */
#define SE3MPC_FLAG_FILTER_RESET 32 /* IPOPT's filter reset heuristic (filter_reset_trigger = 5, max_filter_resets = 5; IPOPT default) */
#define SE3MPC_FLAG_WATCHDOG 16 /* IPOPT's watchdog procedure (watchdog_shortened_iter_trigger = 10, watchdog_trial_iter_max = 3; IPOPT default) */

/* Optional per-problem solver statistics (device or host int32/double arrays; any may be NULL). */
typedef struct {
    int32_t *status;      /* [B] se3mpc_status */
    int32_t *iters;       /* [B] interior-point iterations */
    int32_t *nfact;       /* [B] Riccati (KKT) factorisations */
    int32_t *nls;         /* [B] line-search trial points */
    double  *kkt_error;   /* [B] final scaled optimality error E_0 */
} se3mpc_stats;

void se3mpc_default_params(se3mpc_params *p);
const char *se3mpc_last_error(void);
int se3mpc_device_count(void);
const char *se3mpc_version(void);

/* Bytes of device workspace se3mpc_*_device needs for a batch of `nproblems` on the current device. */
size_t se3mpc_workspace_bytes(const se3mpc_params *p, int nproblems);

/*
 * OCSys.ocSolver for B independent problems (quad_OC.py:104-212).
 *   x0[B][13] goal[B][3] tra_pos[B][3] t[B]; ulast[B][4] or NULL (= zeros, quad_OC.py:130-133)
 *   tra_att[B][att_dim]: att_dim == 3 -> the planner's Rodrigues-like vector tra_ang, converted on the device by
 *   Rd2Rp + toQuaternion (quad_policy.py:10-13, quad_model.py:818-825); att_dim == 4 -> the traversal quaternion
 *   itself (what Quadrotor.init_TraCost holds, quad_model.py:206).
 *   round_t != 0 rounds t to one decimal like run_quad.objective (quad_policy.py:70); get_input does not.
 *   outputs (any of X/U/LAM may be NULL): X[B][N+1][13] U[B][N][4] LAM[B][N][13] cost[B]
 */
int se3mpc_solve_device(const se3mpc_params *p, int B, const double *x0, const double *goal, const double *tra_pos,
                        const double *tra_att, int att_dim, const double *t, const double *ulast, int round_t,
                        double *X, double *U, double *LAM, double *cost, const se3mpc_stats *stats,
                        void *workspace, size_t workspace_bytes, void *stream);

/* get_quadrotor_position + 4 x collis_det + path + reward (quad_policy.py:78-91) for B trajectories.
 *   X[B][N+1][13] gate[B][12] goal[B][3] -> reward[B]; collision[B], path[B], co[B] optional. */
int se3mpc_reward_device(const se3mpc_params *p, int B, const double *X, const double *gate, const double *goal,
                         double *reward, double *collision, double *path, int32_t *co, void *stream);

/* obstacle.collis_det (solid_geometry.py:104-168) for B rotor-tip polylines V[B][horizon+1][3] against gate[B][12]:
 *   score[B] (0 when the plane is never crossed), co[B] optional. */
int se3mpc_collis_det_device(int B, int horizon, const double *V, const double *gate, double *score, int32_t *co,
                             void *stream);

/* run_quad.objective: solve (t rounded) + reward in one launch.  X/U/cost optional. */
int se3mpc_objective_device(const se3mpc_params *p, int B, const double *x0, const double *goal, const double *gate,
                            const double *tra_pos, const double *tra_ang, const double *t, const double *ulast,
                            double *reward, double *X, double *U, double *cost, const se3mpc_stats *stats,
                            void *workspace, size_t workspace_bytes, void *stream);

/* run_quad.sol_gradient for B samples: 9*B solves + rewards + the clip/scale epilogue, out[B][8].
 *   status9[B][9] optional.  workspace must be sized for 9*B problems. */
int se3mpc_sol_gradient_device(const se3mpc_params *p, int B, const double *x0, const double *goal, const double *gate,
                               const double *tra_pos, const double *tra_ang, const double *t, const double *ulast,
                               double *out, int32_t *status9, void *workspace, size_t workspace_bytes, void *stream);

/* se3mpc_solve_device started from a caller-supplied primal point (BASELINE.json configs[3]: receding-horizon re-solves warm
 * started from the previous tick's shifted solution).  X_init[B][N+1][13] (row 0 ignored: x_0 is fixed by quad_OC.py:127-129),
 * U_init[B][N][4]; either may be NULL (then that block keeps the reference's cold start of quad_OC.py:142,158).  The algorithm is
 * unchanged -- it is IPOPT's treatment of any user x0: the point is pushed into the interior of the bounds (bound_push 1e-2),
 * z = 1, least-squares multiplier estimate, mu_init from the parameters.  X_init / U_init must not alias X / U. */
int se3mpc_solve_warm_device(const se3mpc_params *p, int B, const double *x0, const double *goal, const double *tra_pos,
                             const double *tra_att, int att_dim, const double *t, const double *ulast, int round_t,
                             const double *X_init, const double *U_init, double *X, double *U, double *LAM, double *cost,
                             const se3mpc_stats *stats, void *workspace, size_t workspace_bytes, void *stream);

/* ---- moving-gate closed loop (main.py:65-116), one thread per episode ----
 * se3mpc_gate_frame: gate(gate[b]).translate(t[b] velo[b]); .rotate_y(t[b] w[b])  (quad_model.py:686-700, 720-729), then the
 *   DNN-2 input packing of main.py:86-94 (same as quad_moving.py:35-42):
 *   inputs18[b] = [gate.transform(state[b]) (13, quad_model.py:793-810: scipy Rotation round trip re-stated) |
 *                  gate.t_final(final_point[b]) (3) | width | pitch];  gate_pred[B][12] (optional) = the predicted corners.
 * se3mpc_plant_step: state[b] <- nsteps explicit Euler steps of quad_model.py:86-119 with step dt and the control u[b] held
 *   (main.py:108 with uav1.setDyn(0.01), main.py:35; 10 steps per MPC tick, main.py:76). state is updated in place. */
int se3mpc_gate_frame_device(int B, const double *gate, const double *velo, const double *w, const double *t,
                             const double *state, const double *final_point, double *inputs18, double *gate_pred, void *stream);
int se3mpc_plant_step_device(const se3mpc_params *p, int B, double dt, int nsteps, const double *u, double *state, void *stream);
/* se3mpc_episode_advance: everything between two re-solves of main.py's loop in ONE launch (main.py:108 and gate.move,
 *   quad_model.py:769-790), nsteps times: state[b] <- one Euler step of size dt with u[b] held; gate[b] (corners [12]) <- rotated about
 *   its centroid by w[b] dt and translated by (velo[b] + noise[j][b]) dt.  noise[nsteps][B][3] or NULL.  state and gate are updated
 *   in place; bit-identical to nsteps x (se3mpc_plant_step(1 step) ; se3mpc_gate_frame(t = dt) -> gate_pred). */
int se3mpc_episode_advance_device(const se3mpc_params *p, int B, double dt, int nsteps, const double *u, double *state, double *gate,
                                  const double *velo, const double *w, const double *noise, void *stream);

/* Quadrotor.get_quadrotor_position (quad_model.py:239-276): X[nstates][13] -> pos[nstates][15] = centre of mass + the four rotor
 * tips r + R(q)^T b_i, b_i = (+-a, +-a, 0), a = wing_len * 0.5 / sqrt(2), with the reference's un-normalised quaternion. */
int se3mpc_rotor_positions_device(int nstates, double wing_len, const double *X, double *pos, void *stream);

/* Host-buffer variants: same semantics, HOST pointers, synchronous (H2D + kernels + D2H inside). */
int se3mpc_solve_host(const se3mpc_params *p, int B, const double *x0, const double *goal, const double *tra_pos,
                      const double *tra_att, int att_dim, const double *t, const double *ulast, int round_t,
                      double *X, double *U, double *LAM, double *cost, const se3mpc_stats *stats);
int se3mpc_reward_host(const se3mpc_params *p, int B, const double *X, const double *gate, const double *goal,
                       double *reward, double *collision, double *path, int32_t *co);
int se3mpc_collis_det_host(int B, int horizon, const double *V, const double *gate, double *score, int32_t *co);
int se3mpc_gate_frame_host(int B, const double *gate, const double *velo, const double *w, const double *t, const double *state,
                           const double *final_point, double *inputs18, double *gate_pred);
int se3mpc_plant_step_host(const se3mpc_params *p, int B, double dt, int nsteps, const double *u, double *state);
int se3mpc_rotor_positions_host(int nstates, double wing_len, const double *X, double *pos);
int se3mpc_objective_host(const se3mpc_params *p, int B, const double *x0, const double *goal, const double *gate,
                          const double *tra_pos, const double *tra_ang, const double *t, const double *ulast,
                          double *reward, double *X, double *U, double *cost, const se3mpc_stats *stats);
int se3mpc_sol_gradient_host(const se3mpc_params *p, int B, const double *x0, const double *goal, const double *gate,
                             const double *tra_pos, const double *tra_ang, const double *t, const double *ulast,
                             double *out, int32_t *status9);
/* frees the staging buffers the *_host entry points cache */
void se3mpc_host_release(void);

/* ---- diagnostics (used by tests/: the DEVICE code of the derivative tables, the Riccati step and the libm-free helpers) ----
 * se3mpc_kkt_step_device: ONE Newton step of the interior-point method from a caller-supplied iterate, for B independent iterates of
 *   horizon N: w[B][17N+13] in the reference's ordering [x0,u0,x1,u1,...,xN] (quad_OC.py:125-158), lam[B][13N] (multipliers of
 *   g_k = x_k + dt f - x_{k+1}, quad_OC.py:162), zL / zU [B][17N+13] (bound multipliers; only the entries of u_k and of omega_k, k >= 1,
 *   are read).  tra_q[B][4] is the traversal quaternion, t[B] is used as given.  Runs the same device functions as the solver
 *   (prepare_stages, riccati_factor on DMMA tiles, riccati_forward) and returns dw[B][17N+13], lam_plus[B][13N], ok[B] (1 iff the KKT
 *   matrix has inertia (n, m, 0) for this delta_w).  stage_dump (optional) [B][16*28 + 24*24]: the dense stage matrices of stage
 *   `dump_stage` exactly as the factor sweep sees them: Rb = [A | 0 | B | c] (16 x 28, columns 0..12, 16..19, 20) followed by
 *   Qa (24 x 24: rows/cols 0..12 x, 16..19 u, 20 gradient row) -- Hessian of the Lagrangian incl. Sigma and delta_w.
 * se3mpc_device_math_device: out[n][8] = [ddiv(a,b), dsqrt(|a|), round1(a), rd2quat(a,b,c)(4), 0] with the device variants of the
 *   helpers (the CPU emulation tier runs their libm alternates), a/b/c = in[n][3]. */
int se3mpc_kkt_step_device(const se3mpc_params *p, int B, const double *goal, const double *tra_pos, const double *tra_q,
                           const double *t, const double *ulast, const double *w, const double *lam, const double *zL,
                           const double *zU, double mu, double delta_w, double obj_factor, double *dw, double *lam_plus,
                           int32_t *ok, int dump_stage, double *stage_dump, void *workspace, size_t workspace_bytes, void *stream);
int se3mpc_device_math_device(int n, const double *in3, double *out8, void *stream);

/* FP64 FMA-pipe peak probe (dependent-free DFMA chains on every SM): the denominator of the solver's roofline.
 * Enqueues one launch; *flops_per_launch = FLOPs it performs (2 per FMA).  Time it with CUDA events. */
int se3mpc_fp64_probe_device(int iters, double *sink, void *stream, double *flops_per_launch);

/* number of kernels this library has launched since load (bench.py reports it as gpu_launches) */
long long se3mpc_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif
