// se3mpc.cu -- kernels and C ABI of libse3mpc.so (sm_100a).  See include/se3mpc.h for the boundary and
// se3mpc_core.cuh for the warp-cooperative interior-point solver.
//
// Kernels
//   se3mpc_solve_kernel<CTAS>  persistent warps, one problem per warp at a time, problems handed out by an atomic ticket in
//                         longest-predicted-first order so that the very different iteration counts of the NLPs leave a short tail.
//                         Two instances: 8 warps/SM x 255 registers (small launches: latency) and 12 warps/SM x 168 registers.
//                         modes: solve (OCSys.ocSolver), objective (solve + reward fused in the epilogue),
//                         sol_gradient (9 perturbed objectives per sample, quad_policy.py:94-112).
//   se3mpc_reward_kernel  stand-alone rotor-tip roll-out + gate reward, one thread per trajectory, the states of 32 trajectories
//                         streamed through shared memory by per-lane TMA bulk copies (double-buffered): HBM-bound.
//   se3mpc_solgrad_epilogue  the clip/scale rule of sol_gradient over the 9 rewards of each sample.
//   se3mpc_gate_frame_kernel, se3mpc_plant_step_kernel  the per-tick pieces of the moving-gate closed loop (main.py:65-116).
#include <cuda_runtime.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "se3mpc.h"
#include "se3mpc_core.cuh"

namespace {

#ifndef SE3_WARPS_PER_CTA
#define SE3_WARPS_PER_CTA 4
#endif
#ifndef SE3_CTAS_PER_SM
#define SE3_CTAS_PER_SM 3   // measured on B200 (profiles/README.md): 12 warps/SM (168 registers); the GPC instruction cache, not occupancy, bounds the kernel
#endif
constexpr int WARPS_PER_CTA = SE3_WARPS_PER_CTA;
constexpr int CTAS_PER_SM = SE3_CTAS_PER_SM;          // resident warps (= problems) per SM = product; 13 KB shared memory per warp

enum : int { MODE_SOLVE = 0, MODE_OBJECTIVE = 1, MODE_SOLGRAD = 2 };

struct KArgs {
    se3mpc_params par;
    int nprob;                 // problems handed out (9*B in sol_gradient mode)
    int mode, round_t, att_dim;
    const double *x0, *goal, *gate, *tra_pos, *tra_ang, *t, *ulast;
    const double *Xinit, *Uinit;      // optional primal starting points [nprob][N+1][13] / [nprob][N][4]
    double *X, *U, *LAM, *cost, *reward;
    int32_t *status, *iters, *nfact, *nls;
    double *kkt_error;
    double *slots;
    unsigned int *ticket;      // [0] ticket of the main pass, [1] number of problems queued for the restoration pass, [2] its ticket
    const int32_t *order;      // longest-predicted-first processing order (or nullptr)
    int32_t *redo;             // problems whose line search failed in the main pass (the restoration pass re-solves them)
};

// Instantiated for 2 and for CTAS_PER_SM resident CTAs per SM: the launch bound sets the register budget (255 vs 168), and the
// host picks the instance by batch size (ctas_per_sm_for).
// FULL: IPOPT's restoration phase and watchdog are compiled in (se3mpc_resto.cuh).  The default build is single-pass: the throughput
// kernels are FULL.  -DSE3_TWO_PASS=1 builds them without (FULL = false) and adds a second-pass instance (REDO = true) that re-solves,
// from the start, the problems the first pass queued in a.redo (failed line search / watchdog wake-up); kept for A/B measurements
// of what the extra code costs the throughput path (profiles/README.md).
#ifndef SE3_TWO_PASS
#define SE3_TWO_PASS 0
#endif
constexpr bool TWO_PASS = SE3_TWO_PASS != 0;
template <int CTAS, bool FULL, bool REDO>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32, CTAS) se3mpc_solve_kernel(const __grid_constant__ KArgs a)
{
    constexpr bool RESTO = REDO;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // CTA-shared per-lane work table (identical for every warp), then one WarpSmem per warp
    if (warp == 0) se3::make_lane_tab(lane, reinterpret_cast<se3::LaneTab *>(se3_smem_raw)[lane]);
    __syncthreads();
    se3::Ctx c;
    c.sm = reinterpret_cast<se3::WarpSmem *>(se3_smem_raw + se3::LTAB_BYTES) + warp;
    c.lt = reinterpret_cast<const se3::LaneTab *>(se3_smem_raw);
    c.ws = a.slots + (size_t)(blockIdx.x * WARPS_PER_CTA + warp) * se3::SLOT_DOUBLES;
    const int N = a.par.N;

    const unsigned nwork = RESTO ? (unsigned)se3::uni((int)a.ticket[1]) : (unsigned)a.nprob;
    for (;;) {
        unsigned int pb = 0;
        if (lane == 0) pb = atomicAdd(a.ticket + (RESTO ? 2 : 0), 1u);
        pb = __shfl_sync(0xffffffffu, pb, 0);
        if (pb >= nwork) break;
        if (RESTO) pb = (unsigned)se3::uni((int)a.redo[pb]);
        else if (a.order) pb = (unsigned)se3::uni((int)a.order[pb]);      // (a loaded value is not provably warp-uniform: se3mpc_core.cuh, uni())

        // ---- instance (uniform registers) ----
        int src = (int)pb, variant = 0;
        if (a.mode == MODE_SOLGRAD) { src = (int)pb / 9; variant = (int)pb % 9; }
        se3::Instance in;
        // (loaded values are not provably warp-uniform for ptxas; the branches in round1 / rd2quat / the libm slow paths behind
        //  them would count as divergent and drag every phase into divergence-safe code: se3mpc_core.cuh, uni())
        for (int i = 0; i < 13; i++) in.x0[i] = se3::uni(a.x0[(size_t)src * 13 + i]);
        double ang[4];
        for (int i = 0; i < 3; i++) {
            in.goal[i] = se3::uni(a.goal[(size_t)src * 3 + i]);
            in.tra_pos[i] = se3::uni(a.tra_pos[(size_t)src * 3 + i]);
        }
        for (int i = 0; i < a.att_dim; i++) ang[i] = se3::uni(a.tra_ang[(size_t)src * a.att_dim + i]);
        double t = se3::uni(a.t[src]);
        bool with_ulast = (a.ulast != nullptr);
        int round_t = a.round_t;
        if (a.mode == MODE_SOLGRAD) {
            // quad_policy.py:97-110: nominal and t-/+0.1 without Ulast, the six one-sided perturbations with it
            const double delta = 1e-3;
            if (variant >= 1 && variant <= 3) in.tra_pos[variant - 1] += delta;
            else if (variant >= 4 && variant <= 6) ang[variant - 4] += delta;
            else if (variant == 7) t = t - 0.1;
            else if (variant == 8) t = t + 0.1;
            if (variant == 0 || variant >= 7) with_ulast = false;
            round_t = 1;
        }
        if (a.att_dim == 4) { for (int i = 0; i < 4; i++) in.tra_q[i] = ang[i]; }
        else se3::rd2quat(ang, in.tra_q);
        in.t_tra = round_t ? se3::round1(t) : t;
        for (int j = 0; j < 4; j++) in.ulast[j] = with_ulast ? se3::uni(a.ulast[(size_t)src * 4 + j]) : 0.0;
        in.xinit = a.Xinit ? a.Xinit + (size_t)pb * (N + 1) * 13 : nullptr;
        in.uinit = a.Uinit ? a.Uinit + (size_t)pb * N * 4 : nullptr;

#ifndef SE3_HOT_WATCHDOG
#define SE3_HOT_WATCHDOG 1
#endif
        const se3::SolveStats st = se3::ipm_solve<FULL, FULL || (SE3_HOT_WATCHDOG != 0)>(c, a.par, in);
        if (!FULL && st.want_resto && lane == 0) a.redo[atomicAdd(a.ticket + 1, 1u)] = (int32_t)pb;

        se3::write_outputs(c, a.X ? a.X + (size_t)pb * (N + 1) * 13 : nullptr, a.U ? a.U + (size_t)pb * N * 4 : nullptr,
                           a.LAM ? a.LAM + (size_t)pb * N * 13 : nullptr);
        double rew = 0.0;
        if (a.mode != MODE_SOLVE) {
            double gate[12], goal[3];
            for (int i = 0; i < 12; i++) gate[i] = a.gate[(size_t)src * 12 + i];
            for (int i = 0; i < 3; i++) goal[i] = in.goal[i];
            rew = se3::reward_warp(a.par, gate, goal,
                                   [&](int s, double *x) {
#pragma unroll
                                       for (int i = 0; i < 13; i++) x[i] = se3::row(c, se3::R_X + i)[s];
                                   },
                                   nullptr, nullptr, nullptr);
        }
        if (lane == 0) {
            if (a.cost) a.cost[pb] = st.cost;
            if (a.reward) a.reward[pb] = rew;
            if (a.status) a.status[pb] = st.status;
            if (a.iters) a.iters[pb] = st.iters;
            if (a.nfact) a.nfact[pb] = st.nfact;
            if (a.nls) a.nls[pb] = st.nls;
            if (a.kkt_error) a.kkt_error[pb] = st.kkt_error;
        }
        __syncwarp();
    }
}

// ---- longest-predicted-first ordering ------------------------------------------------------------------------------
// The NLPs need 35..130 interior-point iterations; with more problems than resident warps the makespan is set by the
// stragglers that start late.  A predictor of the factorisation count (lpt_bucket) buckets the problems; warps take tickets in
// descending-bucket order.  Results do not depend on the
// order (every problem is solved by one warp with a fixed arithmetic order).
constexpr int LPT_BUCKETS = 256;

// Quadratic predictor of the factorisation count (profiles/lpt/fit_predictor.py: ridge regression on the 210 monomials of 19 standardised
// instance features, 65 536 synthetic problems solved by the CPU oracle, held-out correlation 0.86; the linear four-feature predictor of
// round 1 had 0.70).  Single precision: only the ORDER of the problems depends on it, never a result.
#include "se3mpc_lpt_coef.h"
__device__ __forceinline__ int lpt_bucket(const KArgs &a, int pb)
{
    const int src = (a.mode == MODE_SOLGRAD) ? pb / 9 : pb;
    const double *x0 = a.x0 + (size_t)src * 13, *g = a.goal + (size_t)src * 3, *tp = a.tra_pos + (size_t)src * 3;
    const double *ta = a.tra_ang + (size_t)src * a.att_dim;
    float f[LPT_NF];
    for (int i = 0; i < 3; i++) { f[i] = (float)x0[i]; f[3 + i] = (float)g[i]; f[6 + i] = (float)tp[i]; }
    if (a.att_dim == 3) { for (int i = 0; i < 3; i++) f[9 + i] = (float)ta[i]; }
    else { const double q0 = (fabs(ta[0]) > 1e-3) ? ta[0] : 1e-3; for (int i = 0; i < 3; i++) f[9 + i] = (float)(ta[1 + i] / q0); }   // Rodrigues vector of the quaternion
    f[12] = (float)(a.round_t ? se3::round1(a.t[src]) : a.t[src]);
    f[13] = (float)x0[6]; f[14] = (float)x0[9];
    f[15] = sqrtf(f[0] * f[0] + f[1] * f[1] + f[2] * f[2]);
    f[16] = sqrtf((f[0] - f[3]) * (f[0] - f[3]) + (f[1] - f[4]) * (f[1] - f[4]) + (f[2] - f[5]) * (f[2] - f[5]));
    f[17] = sqrtf(f[6] * f[6] + f[7] * f[7] + f[8] * f[8]);
    f[18] = sqrtf(f[9] * f[9] + f[10] * f[10] + f[11] * f[11]);
#pragma unroll
    for (int i = 0; i < LPT_NF; i++) f[i] = (f[i] - LPT_MEAN[i]) * LPT_ISTD[i];
    float score = LPT_W[0];
    int m = 1;
#pragma unroll
    for (int i = 0; i < LPT_NF; i++) score += LPT_W[m++] * f[i];
#pragma unroll
    for (int i = 0; i < LPT_NF; i++)
#pragma unroll
        for (int j = i; j < LPT_NF; j++) score += LPT_W[m++] * f[i] * f[j];
    int b = (int)(score * 0.75f);           // 256 buckets over 0 .. 340 predicted factorisations
    b = b < 0 ? 0 : (b > LPT_BUCKETS - 1 ? LPT_BUCKETS - 1 : b);
    return LPT_BUCKETS - 1 - b;        // bucket 0 = predicted longest
}

__global__ void se3mpc_lpt_hist(const __grid_constant__ KArgs a, int32_t *hist, int32_t *key)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.nprob) return;
    const int b = lpt_bucket(a, i);
    key[i] = b;
    atomicAdd(&hist[b], 1);
}

__global__ void se3mpc_lpt_scan(int32_t *hist)
{
    __shared__ int32_t sh[LPT_BUCKETS];
    const int t = threadIdx.x;
    sh[t] = hist[t];
    __syncthreads();
    if (t == 0) {
        int32_t run = 0;
        for (int b = 0; b < LPT_BUCKETS; b++) { const int32_t c = sh[b]; sh[b] = run; run += c; }
    }
    __syncthreads();
    hist[t] = sh[t];
}

__global__ void se3mpc_lpt_scatter(int nprob, int32_t *offs, const int32_t *key, int32_t *order)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nprob) return;
    const int pos = atomicAdd(&offs[key[i]], 1);
    order[pos] = i;
}

// ---- moving-gate closed loop (main.py:65-116): one thread per episode -------------------------------------------------
__global__ void se3mpc_gate_frame_kernel(int B, const double *gate, const double *velo, const double *w, const double *t,
                                         const double *state, const double *fin, double *inputs18, double *gate_pred)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    double g[12], v[3], x[13], f[3], out[18], gp[12];
    for (int i = 0; i < 12; i++) g[i] = gate[(size_t)b * 12 + i];
    for (int i = 0; i < 3; i++) { v[i] = velo[(size_t)b * 3 + i]; f[i] = fin[(size_t)b * 3 + i]; }
    for (int i = 0; i < 13; i++) x[i] = state[(size_t)b * 13 + i];
    se3::gate_frame(g, v, w[b], t[b], x, f, out, gate_pred ? gp : nullptr);
    for (int i = 0; i < 18; i++) inputs18[(size_t)b * 18 + i] = out[i];
    if (gate_pred)
        for (int i = 0; i < 12; i++) gate_pred[(size_t)b * 12 + i] = gp[i];
}

__global__ void se3mpc_plant_step_kernel(const __grid_constant__ se3mpc_params par, int B, double dt, int nsteps, const double *u, double *state)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    se3::Prob P;
    se3::prob_from_params(par, P);
    double x[13], uu[4];
    for (int i = 0; i < 13; i++) x[i] = state[(size_t)b * 13 + i];
    for (int j = 0; j < 4; j++) uu[j] = u[(size_t)b * 4 + j];
    se3::plant_steps(P, dt, nsteps, uu, x);
    for (int i = 0; i < 13; i++) state[(size_t)b * 13 + i] = x[i];
}

// One MPC tick of the moving-gate closed loop between two re-solves (main.py:108 + gate.move, quad_model.py:769-790): nsteps times
// { plant Euler step with the control held; gate rotated about its centroid by w dt and translated by (velo + noise) dt }.
// One thread per episode, state and gate corners updated in place: replaces nsteps plant_step + nsteps gate_frame launches.
__global__ void se3mpc_episode_advance_kernel(const __grid_constant__ se3mpc_params par, int B, double dt, int nsteps, const double *u,
                                              double *state, double *gate, const double *velo, const double *w, const double *noise)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    se3::Prob P;
    se3::prob_from_params(par, P);
    double x[13], uu[4], g[12], v0[3];
    for (int i = 0; i < 13; i++) x[i] = state[(size_t)b * 13 + i];
    for (int j = 0; j < 4; j++) uu[j] = u[(size_t)b * 4 + j];
    for (int i = 0; i < 12; i++) g[i] = gate[(size_t)b * 12 + i];
    for (int d = 0; d < 3; d++) v0[d] = velo[(size_t)b * 3 + d];
    const double wb = w[b];
    for (int j = 0; j < nsteps; j++) {
        se3::plant_steps(P, dt, 1, uu, x);
        double v[3] = {v0[0], v0[1], v0[2]}, Pn[4][3], cen[3];
        if (noise)
            for (int d = 0; d < 3; d++) v[d] = v0[d] + noise[((size_t)j * B + b) * 3 + d];
        se3::gate_predict(g, v, wb, dt, Pn, cen);
        for (int i = 0; i < 4; i++)
            for (int d = 0; d < 3; d++) g[i * 3 + d] = Pn[i][d];
    }
    for (int i = 0; i < 13; i++) state[(size_t)b * 13 + i] = x[i];
    for (int i = 0; i < 12; i++) gate[(size_t)b * 12 + i] = g[i];
}

// quad_model.py:239-276 get_quadrotor_position: centre of mass + the four rotor tips of every state, one thread per state
__global__ void se3mpc_rotor_positions_kernel(int nstates, double arm, const double *X, double *pos)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nstates) return;
    double x[13];
    for (int k = 0; k < 13; k++) x[k] = X[(size_t)i * 13 + k];
    double *o = pos + (size_t)i * 15;
    o[0] = x[0]; o[1] = x[1]; o[2] = x[2];
    for (int r = 0; r < 4; r++) {
        double tip[3];
        se3::rotor_tip(x, r, arm, tip);
        o[3 + 3 * r] = tip[0]; o[4 + 3 * r] = tip[1]; o[5 + 3 * r] = tip[2];
    }
}

struct RArgs {
    se3mpc_params par;
    int B;
    const double *X, *gate, *goal;
    double *reward, *collision, *path;
    int32_t *co;
    int32_t *tcross;           // scratch [B]: the first plane crossing of the four rotor tips, 8 bits each (0xFF: none)
    double *pathsum;           // scratch [B]
};

// Stand-alone reward: one THREAD per trajectory; a warp streams its 32 trajectories through shared memory in chunks of RW_CH
// stages.  The chunk of a trajectory is RW_CH*13 contiguous doubles in HBM: each lane fetches its own chunk with ONE bulk
// asynchronous copy (cp.async.bulk, the TMA engine; completion on a per-warp mbarrier) into one of two buffers while the
// threads consume the other, so the copy costs one instruction per chunk and 6.5 KB per warp are always in flight.
// Trajectories are only 8-byte aligned (5304 B each), bulk copies need 16: a copy starts at the aligned-down address and
// the consumer skips the leading double.  Rows have an odd stride (conflict-free per-thread LDS.64).  HBM-bound: 5456 B per
// trajectory.
#ifndef SE3_RW_CH
#define SE3_RW_CH 4
#endif
#ifndef SE3_RW_CTAS
#define SE3_RW_CTAS 2
#endif
#ifndef SE3_RW_NBUF
#define SE3_RW_NBUF 2
#endif
constexpr int RW_CH = SE3_RW_CH, RW_ROW = RW_CH * 13, RW_STRIDE = (RW_ROW + 3) | 1, RW_WARPS = 4, RW_CTAS = SE3_RW_CTAS;
constexpr int RW_NBUF = SE3_RW_NBUF;                  // buffers (chunks in flight + 1) per warp
constexpr size_t RW_BUF = (size_t)RW_WARPS * RW_NBUF * 32 * RW_STRIDE * sizeof(double);
constexpr size_t RW_SMEM = RW_BUF + RW_WARPS * RW_NBUF * sizeof(unsigned long long);

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(RW_WARPS * 32, RW_CTAS) se3mpc_reward_kernel(const __grid_constant__ RArgs a)
{
    extern __shared__ __align__(16) double rw_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int N = a.par.N, nst = N + 1, nch = (nst + RW_CH - 1) / RW_CH;
    double *buf = rw_smem + (size_t)warp * RW_NBUF * 32 * RW_STRIDE;
    const unsigned bar0 = smem_u32(reinterpret_cast<unsigned char *>(rw_smem) + RW_BUF) + warp * RW_NBUF * 8;     // one mbarrier per buffer
    if (lane == 0) {
        for (int b = 0; b < RW_NBUF; b++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 32;" ::"r"(bar0 + b * 8));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    const size_t total = (size_t)a.B * nst * 13;          // doubles in X
    const int row0 = lane * RW_STRIDE + ((lane * RW_STRIDE) & 1);      // 16-byte aligned start of this lane's row
    unsigned gch = 0;                                     // chunks consumed so far: buffer = gch % NBUF, phase parity = (gch / NBUF) & 1
    const int ngroups = (a.B + 31) / 32;
    for (int grp = blockIdx.x * RW_WARPS + warp; grp < ngroups; grp += gridDim.x * RW_WARPS) {
        const int pb0 = grp * 32, pb = pb0 + lane;
        const bool live = pb < a.B;
        const int pbc = live ? pb : a.B - 1;
        const size_t base = (size_t)pbc * nst * 13;
        // chunk ch of this lane's trajectory -> its row of buffer g & 1; returns the offset of the first state in the row
        auto issue = [&](const int ch, const unsigned g) {
            const unsigned bar = bar0 + (g % RW_NBUF) * 8;
            if (live) {
                const int s0 = ch * RW_CH, cnt = min(RW_CH, nst - s0) * 13;
                const size_t off = base + (size_t)s0 * 13;
                const int shift = (int)((reinterpret_cast<size_t>(a.X + off) >> 3) & 1);
                int n = (cnt + shift + 1) & ~1;
                if (off - shift + n > total) n -= 2;                        // never read past the end of X (tail fixed up by the consumer)
                // ... nor before its start: an X that is only 8-byte aligned (a view such as X[1:]) has shift = 1 at off = 0; the copy
                // then starts one 16-byte unit later and the consumer fetches the first double by hand
                const int head = (off == 0 && shift == 1) ? 2 : 0;
                n -= head;
                const unsigned dst = smem_u32(buf + (g % RW_NBUF) * 32 * RW_STRIDE + row0 + head);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(n > 0 ? n * 8 : 0) : "memory");
                if (n > 0)
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"(dst), "l"(a.X + off + head - shift), "r"(n * 8), "r"(bar) : "memory");
            } else {
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
            }
        };
        __syncwarp();
        for (int d = 0; d < RW_NBUF - 1 && d < nch; d++) issue(d, gch + d);
        double gate[12], goal[3];
#pragma unroll
        for (int i = 0; i < 12; i++) gate[i] = __ldg(a.gate + (size_t)pbc * 12 + i);
#pragma unroll
        for (int i = 0; i < 3; i++) goal[i] = __ldg(a.goal + (size_t)pbc * 3 + i);
        se3::RewardAcc acc;
        acc.begin(a.par, gate);
        for (int ch = 0; ch < nch; ch++, gch++) {
            if (ch + RW_NBUF - 1 < nch) issue(ch + RW_NBUF - 1, gch + RW_NBUF - 1);
            {
                const unsigned bar = bar0 + (gch % RW_NBUF) * 8, parity = (gch / RW_NBUF) & 1;
                unsigned done = 0;
                while (!done)
                    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
            }
            const int s0 = ch * RW_CH, cnt = min(RW_CH, nst - s0) * 13;
            if (live) {
                const size_t off = base + (size_t)s0 * 13;
                const int shift = (int)((reinterpret_cast<size_t>(a.X + off) >> 3) & 1);
                double *cur = buf + (gch % RW_NBUF) * 32 * RW_STRIDE + row0 + shift;
                const int n = (cnt + shift + 1) & ~1;
                bool patched = false;
                if (off - shift + n > total) {                               // the copy stopped 16 bytes early: last double(s) by hand
                    for (int i = n - 2 - shift; i < cnt; i++) cur[i] = __ldg(a.X + off + i);
                    patched = true;
                }
                if (off == 0 && shift == 1) { cur[0] = __ldg(a.X); patched = true; }      // the copy started 16 bytes late (see issue())
                // generic-proxy stores into a buffer that a later bulk copy (async proxy) overwrites: order them before it
                if (patched) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#pragma unroll 1
                for (int q = 0; q * 13 < cnt; q++) {
                    double x[13];
#pragma unroll
                    for (int i = 0; i < 13; i++) x[i] = cur[q * 13 + i];
                    acc.consume(s0 + q, x, goal);
                }
            }
            __syncwarp();           // every lane is done with this buffer before the issue of the next iteration refills it
        }
        // the streaming pass ends here: the (rare, branchy, register-hungry) scoring of the four crossings is a second, small kernel
        if (live) {
            unsigned packed = 0;
#pragma unroll
            for (int r = 0; r < 4; r++) packed |= (unsigned)(acc.tcross[r] & 0xFF) << (8 * r);
            a.tcross[pb] = (int32_t)packed;
            a.pathsum[pb] = acc.path;
        }
    }
}

// Second pass of the stand-alone reward: one thread per trajectory builds the gate's wedge normals once and scores the recorded
// crossings of the four rotor tips from the two states around each (2 x 104 B gathered per rotor), then assembles
// reward = 1000 collision - 0.5 path + 100 (quad_policy.py:90) with the rotor scores summed in the reference's order.
__global__ void __launch_bounds__(128) se3mpc_reward_score_kernel(const __grid_constant__ RArgs a)
{
    const int pb = blockIdx.x * blockDim.x + threadIdx.x;
    if (pb >= a.B) return;
    const int nst = a.par.N + 1;
    double gate[12];
#pragma unroll
    for (int i = 0; i < 12; i++) gate[i] = __ldg(a.gate + (size_t)pb * 12 + i);
    se3::Gate G;
    se3::gate_setup(gate, G);
    const unsigned packed = (unsigned)a.tcross[pb];
    const double arm = a.par.wing_len * 0.5 / sqrt(2.0);
    const double *Xp = a.X + (size_t)pb * nst * 13;
    double col = 0.0;
    int cosum = 0;
#pragma unroll 1
    for (int r = 0; r < 4; r++) {
        const int tb = (int)((packed >> (8 * r)) & 0xFF), t = (tb == 0xFF) ? -1 : tb;
        int co = 0;
        col += se3::rotor_crossing_score(G, arm, r, t, [&](int tt, double *x) {
#pragma unroll
            for (int i = 0; i < 13; i++) x[i] = __ldg(Xp + tt * 13 + i);
        }, &co);
        cosum += co;
    }
    const double path = a.pathsum[pb];
    a.reward[pb] = 1000 * col - 0.5 * path + 100;          // quad_policy.py:90
    if (a.collision) a.collision[pb] = col;
    if (a.path) a.path[pb] = path;
    if (a.co) a.co[pb] = cosum;
}

__global__ void __launch_bounds__(256) se3mpc_collis_kernel(int B, int horizon, const double *V, const double *gate, double *score,
                                                            int32_t *co_out)
{
    const int warps_per_cta = blockDim.x >> 5;
    const int lane = threadIdx.x & 31;
    for (int pb = blockIdx.x * warps_per_cta + (threadIdx.x >> 5); pb < B; pb += gridDim.x * warps_per_cta) {
        const double *Vp = V + (size_t)pb * (horizon + 1) * 3;
        double g[12];
        for (int i = 0; i < 12; i++) g[i] = __ldg(gate + (size_t)pb * 12 + i);
        int co;
        const double sc = se3::collis_warp(g, horizon,
                                           [&](int s, double *v) { v[0] = __ldg(Vp + s * 3); v[1] = __ldg(Vp + s * 3 + 1); v[2] = __ldg(Vp + s * 3 + 2); },
                                           &co);
        if (lane == 0) {
            score[pb] = sc;
            if (co_out) co_out[pb] = co;
        }
    }
}

// FP64 peak probe: 8 independent DMMA.8x8x4 accumulator chains per warp (the instruction the factor sweep issues; on B200 the
// FP64 tensor rate is 37 TFLOP/s, the DFMA rate 34 TFLOP/s -- profiles/micro/dmma_probe.cu).  MEASURED_PEAKS.json has no fp64 entry.
__global__ void __launch_bounds__(256) se3mpc_fp64_probe_kernel(int iters, double seed, double *sink)
{
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; i++) { c[i][0] = seed + threadIdx.x + i; c[i][1] = seed - i; }
    const double a = 1.0 + 1e-9 * threadIdx.x, b = 1.0 - 1e-9 * threadIdx.x;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double r = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) r += c[i][0] + c[i][1];
    if (r == 123.456) sink[0] = r;
}

// quad_policy.py:98-112 over the 9 rewards of every sample
__global__ void se3mpc_solgrad_epilogue(int B, const double *rew9, const double *tra_ang, const int32_t *st_in, double *out,
                                        int32_t *status9)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const double *r = rew9 + (size_t)b * 9;
    const double j = r[0];
    double o[8];
    for (int i = 0; i < 3; i++) o[i] = -(fmin(fmax(r[1 + i] - j, -0.5), 0.5) * 0.1);
    for (int i = 0; i < 3; i++) {
        const double ai = tra_ang[(size_t)b * 3 + i];
        o[3 + i] = -(fmin(fmax(r[4 + i] - j, -0.5), 0.5) * (1.0 / (500.0 * ai * ai + 5.0)));
    }
    double drdt = 0.0;
    if (r[7] - j > 2.0) drdt = -0.05;
    if (r[8] - j > 2.0) drdt = 0.05;
    o[6] = -drdt;
    o[7] = j;
    for (int i = 0; i < 8; i++) out[(size_t)b * 8 + i] = o[i];
    if (status9)
        for (int i = 0; i < 9; i++) status9[(size_t)b * 9 + i] = st_in[(size_t)b * 9 + i];
}


// ---- diagnostics: one Newton step from a given iterate (tests; same contract as the host emulation's and the CPU checker's one-step entry points) ----
struct StepArgs {
    se3mpc_params par;
    int B, dump_stage;
    const double *goal, *tra_pos, *tra_q, *t, *ulast, *w, *lam, *zL, *zU;
    double mu, delta_w, obj_factor;
    double *dw, *lam_plus, *dump;
    int32_t *ok;
    double *slots;
};

__global__ void __launch_bounds__(WARPS_PER_CTA * 32, 2) se3mpc_kkt_step_kernel(const __grid_constant__ StepArgs a)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) se3::make_lane_tab(lane, reinterpret_cast<se3::LaneTab *>(se3_smem_raw)[lane]);
    __syncthreads();
    const int pb = blockIdx.x * WARPS_PER_CTA + warp;
    if (pb >= a.B) return;
    se3::Ctx c;
    c.sm = reinterpret_cast<se3::WarpSmem *>(se3_smem_raw + se3::LTAB_BYTES) + warp;
    c.lt = reinterpret_cast<const se3::LaneTab *>(se3_smem_raw);
    c.ws = a.slots + (size_t)pb * se3::SLOT_DOUBLES;
    se3::WarpSmem &sm = SE3_SMEM(c);
    se3::Prob &P = sm.P;
    const int N = a.par.N, nw = 17 * N + 13;
    se3::prob_from_params(a.par, P);
    for (int i = 0; i < 3; i++) { P.goal[i] = se3::uni(a.goal[(size_t)pb * 3 + i]); P.ptra[i] = se3::uni(a.tra_pos[(size_t)pb * 3 + i]); }
    for (int j = 0; j < 4; j++) P.ulast[j] = a.ulast ? se3::uni(a.ulast[(size_t)pb * 4 + j]) : 0.0;
    P.ttra = se3::uni(a.t[pb]);
    double tq[4];
    for (int i = 0; i < 4; i++) tq[i] = se3::uni(a.tra_q[(size_t)pb * 4 + i]);
    se3::att_form(tq, P.Nq, &P.s0);
    P.sf = a.obj_factor;
    if (lane == 0) { sm.nfact = 0; sm.nfilt = 0; }
    __syncwarp();
    const double *w = a.w + (size_t)pb * nw, *zL = a.zL + (size_t)pb * nw, *zU = a.zU + (size_t)pb * nw, *lam = a.lam + (size_t)pb * 13 * N;
    for (int e = lane; e < se3::R_END * se3::NS; e += 32) c.ws[e] = 0.0;
    __syncwarp();
    for (int s = lane; s <= N; s += 32) {
        for (int i = 0; i < 13; i++) se3::row(c, se3::R_X + i)[s] = w[s * 17 + i];
        for (int j = 0; j < 3; j++) { se3::row(c, se3::R_ZLW + j)[s] = zL[s * 17 + 10 + j]; se3::row(c, se3::R_ZUW + j)[s] = zU[s * 17 + 10 + j]; }
        if (s < N) {
            for (int j = 0; j < 4; j++) {
                se3::row(c, se3::R_U + j)[s] = w[s * 17 + 13 + j];
                se3::row(c, se3::R_ZLU + j)[s] = zL[s * 17 + 13 + j];
                se3::row(c, se3::R_ZUU + j)[s] = zU[s * 17 + 13 + j];
            }
            for (int i = 0; i < 13; i++) se3::row(c, se3::R_LAM + i)[s] = lam[s * 13 + i];
            const double d = P.dt * s - P.ttra;
            se3::row(c, se3::R_GAM)[s] = a.par.gamma_amp * exp(-a.par.gamma_rate * d * d);
        }
    }
    __syncwarp();
    se3::init_AB(c);
    double th, ci;
    se3::eval_constraints(c, &th, &ci);
    const int ok = se3::uni(se3::riccati(c, SE3_MODE_EXACT, 1, 1, a.delta_w, a.mu, se3::R_C, se3::R_DX, se3::R_DU, se3::R_LAMP));
    double *dw = a.dw + (size_t)pb * nw, *lp = a.lam_plus + (size_t)pb * 13 * N;
    for (int s = lane; s <= N; s += 32) {
        for (int i = 0; i < 13; i++) dw[s * 17 + i] = se3::row(c, se3::R_DX + i)[s];
        if (s < N) {
            for (int j = 0; j < 4; j++) dw[s * 17 + 13 + j] = se3::row(c, se3::R_DU + j)[s];
            for (int i = 0; i < 13; i++) lp[s * 13 + i] = se3::row(c, se3::R_LAMP + i)[s];
        }
    }
    if (lane == 0 && a.ok) a.ok[pb] = ok;
    if (a.dump && a.dump_stage >= 0 && a.dump_stage < N) {
        // the stage matrices as the factor sweep's S1 builds them: constant parts (init_AB), then the packed record through the lane table
        __syncwarp();
        // (in the state order of the factor sweep, xpos<SE3_PERM>; the dump is handed out in the natural order)
        constexpr bool PERM = SE3_PERM != 0;
        se3::init_AB(c, PERM);
        double *sbase = reinterpret_cast<double *>(&sm);
        const se3::LaneTab t = SE3_LTAB(c)[lane];
        const se3::GP R = se3::rec(c, a.dump_stage);
        const double r0 = R[lane], r1 = R[32 + lane], r2 = R[64 + lane], r3 = R[96 + lane];
        sbase[PERM ? t.pf0 : t.f0] = r0;
        const double w1 = r1 + (t.addd[0] ? a.delta_w : 0.0), w2 = r2 + (t.addd[1] ? a.delta_w : 0.0);
        for (int d = 0; d < 4; d++) {
            const int f1 = PERM ? t.pf1[d] : t.f1[d], f2 = PERM ? t.pf2[d] : t.f2[d];
            if (f1 >= 0) sbase[f1] = w1;
            if (f2 >= 0) sbase[f2] = w2;
        }
        if ((PERM ? t.pf3 : t.f3) >= 0) sbase[PERM ? t.pf3 : t.f3] = r3;
        __syncwarp();
        // slot of row / column e of the natural 16- resp. 24-wide images (13..15: the padding slots, in order)
        auto slot = [](int e) { return (e < 13) ? se3::xpos<PERM>(e) : (e < 16) ? (PERM ? (e == 13 ? 6 : e == 14 ? 7 : 15) : e) : e; };
        double *o = a.dump + (size_t)pb * (16 * se3::SR + 24 * se3::SQ);
        for (int e = lane; e < 16 * se3::SR; e += 32) {           // dense [16][SR] image of the transposed shared-memory operand
            const int i = e / se3::SR, j = e % se3::SR;
            o[e] = (j < 24) ? sm.Rb[se3::rbi(slot(i), slot(j))] : 0.0;
        }
        for (int e = lane; e < 24 * se3::SQ; e += 32) {
            const int i = e / se3::SQ, j = e % se3::SQ;
            o[16 * se3::SR + e] = (j < 16 || j >= 24) ? sm.Qa[slot(i) * se3::SQ + slot(j)] : sm.Qa[slot(i) * se3::SQ + j];
        }
    }
}

__global__ void se3mpc_device_math_kernel(int n, const double *in3, double *out8)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double a = in3[(size_t)i * 3], b = in3[(size_t)i * 3 + 1], cc = in3[(size_t)i * 3 + 2];
    double *o = out8 + (size_t)i * 8;
    o[0] = se3::ddiv(a, b);
    o[1] = se3::dsqrt(fabs(a));
    o[2] = se3::round1(a);
    const double v[3] = {a, b, cc};
    se3::rd2quat(v, o + 3);
    o[7] = 0.0;
}

// ------------------------------------------------------------------------------------------------------
thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};

int fail(int code, const char *what, cudaError_t e = cudaSuccess)
{
    if (e != cudaSuccess) snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
    else snprintf(g_err, sizeof(g_err), "%s", what);
    return code;
}

#define CK(call)                                                          \
    do {                                                                  \
        cudaError_t e_ = (call);                                          \
        if (e_ != cudaSuccess) return fail(SE3MPC_ERR_CUDA, #call, e_);   \
    } while (0)

int have_device()
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) { cudaGetLastError(); return 0; }
    return n;
}

int num_sms()
{
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms;
}

constexpr size_t HDR_BYTES = 256;
constexpr size_t SOLVER_SMEM = se3::LTAB_BYTES + sizeof(se3::WarpSmem) * WARPS_PER_CTA;
static_assert(CTAS_PER_SM * (SOLVER_SMEM + 1024) <= 228 * 1024, "CTAS_PER_SM resident CTAs (1 KB reserved each) must fit the 228 KB of an sm_100 SM");
constexpr int MAX_DEVICES = 64;

// Function attributes are per device (per primary context): the opt-in to more than 48 KB of dynamic shared memory has to be made on
// every device a launch goes to, not once per process.  Done on first use of each device, return codes checked.
int device_setup()
{
    static std::mutex mu;
    static bool done[MAX_DEVICES] = {};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return fail(SE3MPC_ERR_CUDA, "cudaGetDevice", e);
    if (dev < 0 || dev >= MAX_DEVICES) return fail(SE3MPC_ERR_INVALID, "device index out of range");
    std::lock_guard<std::mutex> lk(mu);
    if (done[dev]) return SE3MPC_OK;
    if ((e = cudaFuncSetAttribute(se3mpc_solve_kernel<2, !TWO_PASS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SOLVER_SMEM)) != cudaSuccess)
        return fail(SE3MPC_ERR_CUDA, "cudaFuncSetAttribute(se3mpc_solve_kernel<2>)", e);
    if ((e = cudaFuncSetAttribute(se3mpc_solve_kernel<CTAS_PER_SM, !TWO_PASS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SOLVER_SMEM)) != cudaSuccess)
        return fail(SE3MPC_ERR_CUDA, "cudaFuncSetAttribute(se3mpc_solve_kernel<CTAS>)", e);
    if (TWO_PASS && (e = cudaFuncSetAttribute(se3mpc_solve_kernel<2, true, TWO_PASS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SOLVER_SMEM)) != cudaSuccess)
        return fail(SE3MPC_ERR_CUDA, "cudaFuncSetAttribute(se3mpc_solve_kernel<2, second pass>)", e);
    if ((e = cudaFuncSetAttribute(se3mpc_kkt_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SOLVER_SMEM)) != cudaSuccess)
        return fail(SE3MPC_ERR_CUDA, "cudaFuncSetAttribute(se3mpc_kkt_step_kernel)", e);
    if ((e = cudaFuncSetAttribute(se3mpc_reward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RW_SMEM)) != cudaSuccess)
        return fail(SE3MPC_ERR_CUDA, "cudaFuncSetAttribute(se3mpc_reward_kernel)", e);
    done[dev] = true;
    return SE3MPC_OK;
}

// Resident CTAs per SM for a launch of `nproblems`; the solver kernel is instantiated for 2 (255 registers) and for CTAS_PER_SM
// (168 registers).  Fewer, fatter warps mean a lower per-problem latency (more registers, a larger share of the instruction
// cache).  Measured on B200 (profiles/README.md): up to ~2 waves of problems the makespan is set by single-problem latency and
// 2 CTAs/SM (8 warps) win by ~10%.  For large batches 3 CTAs/SM (12 warps) used to give ~3% more throughput; since the factor
// sweep lost its divergent branches (round 2) 8 warps win at every size (65 536 problems: 84.7 vs 83.3 k solves/s, 262 144: 85.1
// vs 83.8), so 2 is the policy and the 12-warp instance is kept for SE3MPC_CTAS_PER_SM=3 (tuning).
int ctas_per_sm_for(int nproblems)
{
    static const int forced = [] { const char *e = getenv("SE3MPC_CTAS_PER_SM"); return e ? atoi(e) : 0; }();
    if (forced >= 1 && forced <= CTAS_PER_SM) return forced;
    (void)nproblems;
    return 2;
}

int slots_for(int nproblems)
{
    const int cap = num_sms() * ctas_per_sm_for(nproblems) * WARPS_PER_CTA;
    int s = nproblems < cap ? nproblems : cap;
    s = (s + WARPS_PER_CTA - 1) / WARPS_PER_CTA * WARPS_PER_CTA;
    return s < WARPS_PER_CTA ? WARPS_PER_CTA : s;
}

bool params_ok(const se3mpc_params *p)
{
    return p && p->N >= 2 && p->N <= SE3MPC_MAX_HORIZON && (p->att_power == 1 || p->att_power == 2) && p->dt > 0 &&
           p->u_ub > p->u_lb && p->w_bound > 0 && p->mass > 0;
}

// workspace: [ticket | pad to 256 B][aux doubles: nprob][aux int32: nprob][slots]
struct WsLayout { unsigned int *ticket; double *aux_d; int32_t *aux_i, *lpt_hist, *lpt_key, *lpt_order, *redo; double *slots; int nslots; size_t bytes; };

WsLayout layout(void *ws, int nproblems)
{
    WsLayout L;
    unsigned char *b = static_cast<unsigned char *>(ws);
    size_t off = HDR_BYTES;
    L.ticket = reinterpret_cast<unsigned int *>(b);
    L.aux_d = reinterpret_cast<double *>(b + off);
    off += ((size_t)nproblems * sizeof(double) + 255) / 256 * 256;
    L.aux_i = reinterpret_cast<int32_t *>(b + off);
    off += ((size_t)nproblems * sizeof(int32_t) + 255) / 256 * 256;
    L.lpt_hist = reinterpret_cast<int32_t *>(b + off);
    off += (size_t)LPT_BUCKETS * sizeof(int32_t);
    L.lpt_key = reinterpret_cast<int32_t *>(b + off);
    off += ((size_t)nproblems * sizeof(int32_t) + 255) / 256 * 256;
    L.lpt_order = reinterpret_cast<int32_t *>(b + off);
    off += ((size_t)nproblems * sizeof(int32_t) + 255) / 256 * 256;
    L.redo = reinterpret_cast<int32_t *>(b + off);
    off += ((size_t)nproblems * sizeof(int32_t) + 255) / 256 * 256;
    L.slots = reinterpret_cast<double *>(b + off);
    L.nslots = slots_for(nproblems);
    off += (size_t)L.nslots * se3::SLOT_DOUBLES * sizeof(double);
    L.bytes = off;
    return L;
}

int launch_solver(const se3mpc_params *p, int mode, int nprob, int round_t, const double *x0, const double *goal,
                  const double *gate, const double *tra_pos, const double *tra_ang, int att_dim, const double *t, const double *ulast,
                  double *X, double *U, double *LAM, double *cost, double *reward, const se3mpc_stats *stats,
                  int32_t *status_override, void *workspace, size_t workspace_bytes, cudaStream_t stream,
                  const double *Xinit = nullptr, const double *Uinit = nullptr)
{
    if (nprob <= 0) return SE3MPC_OK;
    WsLayout L = layout(workspace, nprob);
    if (!workspace || workspace_bytes < L.bytes) return fail(SE3MPC_ERR_WORKSPACE, "workspace too small");
    KArgs a;
    memset(&a, 0, sizeof(a));
    a.par = *p; a.nprob = nprob; a.mode = mode; a.round_t = round_t; a.att_dim = att_dim;
    a.x0 = x0; a.goal = goal; a.gate = gate; a.tra_pos = tra_pos; a.tra_ang = tra_ang; a.t = t; a.ulast = ulast;
    a.X = X; a.U = U; a.LAM = LAM; a.cost = cost; a.reward = reward;
    a.Xinit = Xinit; a.Uinit = Uinit;
    if (stats) { a.status = stats->status; a.iters = stats->iters; a.nfact = stats->nfact; a.nls = stats->nls; a.kkt_error = stats->kkt_error; }
    if (status_override) a.status = status_override;
    a.slots = L.slots; a.ticket = L.ticket; a.redo = L.redo;
    CK(cudaMemsetAsync(L.ticket, 0, 4 * sizeof(unsigned int), stream));
    static const bool lpt_off = (getenv("SE3MPC_NO_LPT") != nullptr);      // A/B switch for profiling
    if (nprob > L.nslots && !lpt_off) {     // more problems than resident warps: longest-predicted-first order
        CK(cudaMemsetAsync(L.lpt_hist, 0, LPT_BUCKETS * sizeof(int32_t), stream));
        const int tb = 256, nb = (nprob + tb - 1) / tb;
        se3mpc_lpt_hist<<<nb, tb, 0, stream>>>(a, L.lpt_hist, L.lpt_key);
        se3mpc_lpt_scan<<<1, LPT_BUCKETS, 0, stream>>>(L.lpt_hist);
        se3mpc_lpt_scatter<<<nb, tb, 0, stream>>>(nprob, L.lpt_hist, L.lpt_key, L.lpt_order);
        g_launches += 3;
        CK(cudaGetLastError());
        a.order = L.lpt_order;
    }
    const size_t smem = SOLVER_SMEM;
    { const int rc_attr = device_setup(); if (rc_attr != SE3MPC_OK) return rc_attr; }
    const int grid = L.nslots / WARPS_PER_CTA;
    if (ctas_per_sm_for(nprob) <= 2) se3mpc_solve_kernel<2, !TWO_PASS, false><<<grid, WARPS_PER_CTA * 32, smem, stream>>>(a);
    else se3mpc_solve_kernel<CTAS_PER_SM, !TWO_PASS, false><<<grid, WARPS_PER_CTA * 32, smem, stream>>>(a);
    g_launches++;
    CK(cudaGetLastError());
    if (TWO_PASS && (p->flags & (SE3MPC_FLAG_RESTO | SE3MPC_FLAG_WATCHDOG))) {
        // restoration pass over the problems whose line search failed (usually none: every warp reads the count and leaves).  The
        // grid is at most two CTAs per SM; the slots of the main pass are reused.
        const int cap2 = num_sms() * 2;
        se3mpc_solve_kernel<2, true, TWO_PASS><<<grid < cap2 ? grid : cap2, WARPS_PER_CTA * 32, smem, stream>>>(a);
        g_launches++;
        CK(cudaGetLastError());
    }
    return SE3MPC_OK;
}

}  // namespace

extern "C" {

void se3mpc_default_params(se3mpc_params *p)
{
    memset(p, 0, sizeof(*p));
    p->Jx = 0.0023; p->Jy = 0.0023; p->Jz = 0.004; p->mass = 0.5; p->arm_l = 0.35; p->c_tau = 0.0245;   // quad_policy.py:37
    p->grav = 9.78;                                                                                   // quad_model.py:37
    p->wrt = 5; p->wqt = 80; p->wthrust = 0.1; p->wrf = 5; p->wvf = 5; p->wwf = 3;                    // quad_policy.py:38
    p->wdu = 1.0;                                                                                     // quad_OC.py:150
    p->gamma_amp = 60.0; p->gamma_rate = 10.0;                                                        // quad_OC.py:145
    p->dt = 0.1;                                                                                      // quad_policy.py:43
    p->u_lb = 0.0; p->u_ub = 2.0 * 1.22;                                                              // quad_policy.py:48-51
    p->w_bound = 3.14159265358979323846 / 2.0;                                                        // quad_policy.py:47
    p->wing_len = 1.5;                                                                                // quad_policy.py:19
    p->N = 50;                                                                                        // quad_policy.py:17
    p->att_power = 2;                                                                                 // quad_model.py:210
    p->tol = 1e-8; p->acceptable_tol = 1e-6; p->mu_init = 0.1;                                        // IPOPT defaults
    p->max_iter = 3000; p->acceptable_iter = 15; p->max_soc = 4;
    p->flags = SE3MPC_FLAG_SCALING | SE3MPC_FLAG_LSINIT | SE3MPC_FLAG_RESTO | SE3MPC_FLAG_WATCHDOG | SE3MPC_FLAG_FILTER_RESET;
}

const char *se3mpc_last_error(void) { return g_err; }
int se3mpc_device_count(void) { return have_device(); }
const char *se3mpc_version(void) { return "se3mpc 0.1 (sm_100a)"; }
long long se3mpc_launch_count(void) { return g_launches.load(); }

size_t se3mpc_workspace_bytes(const se3mpc_params *p, int nproblems)
{
    if (!params_ok(p) || nproblems <= 0 || !have_device()) return 0;
    return layout(nullptr, nproblems).bytes;
}

int se3mpc_solve_device(const se3mpc_params *p, int B, const double *x0, const double *goal, const double *tra_pos,
                        const double *tra_ang, int att_dim, const double *t, const double *ulast, int round_t, double *X, double *U,
                        double *LAM, double *cost, const se3mpc_stats *stats, void *workspace, size_t workspace_bytes,
                        void *stream)
{
    if (!params_ok(p) || B < 0 || !x0 || !goal || !tra_pos || !tra_ang || !t || (att_dim != 3 && att_dim != 4))
        return fail(SE3MPC_ERR_INVALID, "invalid argument");
    if (!have_device()) return fail(SE3MPC_ERR_NO_DEVICE, "no CUDA device");
    return launch_solver(p, MODE_SOLVE, B, round_t, x0, goal, nullptr, tra_pos, tra_ang, att_dim, t, ulast, X, U, LAM, cost, nullptr,
                         stats, nullptr, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

int se3mpc_solve_warm_device(const se3mpc_params *p, int B, const double *x0, const double *goal, const double *tra_pos,
                             const double *tra_ang, int att_dim, const double *t, const double *ulast, int round_t,
                             const double *X_init, const double *U_init, double *X, double *U, double *LAM, double *cost,
                             const se3mpc_stats *stats, void *workspace, size_t workspace_bytes, void *stream)
{
    if (!params_ok(p) || B < 0 || !x0 || !goal || !tra_pos || !tra_ang || !t || (att_dim != 3 && att_dim != 4) || (!X_init && !U_init))
        return fail(SE3MPC_ERR_INVALID, "invalid argument");
    if (!have_device()) return fail(SE3MPC_ERR_NO_DEVICE, "no CUDA device");
    return launch_solver(p, MODE_SOLVE, B, round_t, x0, goal, nullptr, tra_pos, tra_ang, att_dim, t, ulast, X, U, LAM, cost, nullptr,
                         stats, nullptr, workspace, workspace_bytes, static_cast<cudaStream_t>(stream), X_init, U_init);
}

int se3mpc_reward_device(const se3mpc_params *p, int B, const double *X, const double *gate, const double *goal,
                         double *reward, double *collision, double *path, int32_t *co, void *stream)
{
    if (!params_ok(p) || B < 0 || !X || !gate || !goal || !reward) return fail(SE3MPC_ERR_INVALID, "invalid argument");
    if (!have_device()) return fail(SE3MPC_ERR_NO_DEVICE, "no CUDA device");
    if (B == 0) return SE3MPC_OK;
    RArgs a;
    a.par = *p; a.B = B; a.X = X; a.gate = gate; a.goal = goal; a.reward = reward; a.collision = collision; a.path = path; a.co = co;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // 12 bytes of scratch per trajectory between the two passes, stream-ordered (freed when the second pass is done)
    void *scratch = nullptr;
    const size_t off_path = ((size_t)B * sizeof(int32_t) + 255) / 256 * 256;
    CK(cudaMallocAsync(&scratch, off_path + (size_t)B * sizeof(double), st));
    a.tcross = static_cast<int32_t *>(scratch);
    a.pathsum = reinterpret_cast<double *>(static_cast<unsigned char *>(scratch) + off_path);
    int grid = ((B + 31) / 32 + RW_WARPS - 1) / RW_WARPS;       // one thread per trajectory, RW_WARPS warps per CTA
    const int cap = num_sms() * RW_CTAS * 2;
    if (grid > cap) grid = cap;
    { const int rc_attr = device_setup(); if (rc_attr != SE3MPC_OK) { cudaFreeAsync(scratch, st); return rc_attr; } }
    se3mpc_reward_kernel<<<grid, RW_WARPS * 32, RW_SMEM, st>>>(a);
    se3mpc_reward_score_kernel<<<(B + 127) / 128, 128, 0, st>>>(a);
    g_launches += 2;
    const cudaError_t e_launch = cudaGetLastError();
    cudaFreeAsync(scratch, st);
    if (e_launch != cudaSuccess) return fail(SE3MPC_ERR_CUDA, "se3mpc_reward launch", e_launch);
    return SE3MPC_OK;
}

int se3mpc_fp64_probe_device(int iters, double *sink, void *stream, double *flops_per_launch)
{
    if (iters <= 0 || !sink) return fail(SE3MPC_ERR_INVALID, "invalid argument");
    if (!have_device()) return fail(SE3MPC_ERR_NO_DEVICE, "no CUDA device");
    const int grid = num_sms() * 8;
    se3mpc_fp64_probe_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(iters, 1.0, sink);
    g_launches++;
    CK(cudaGetLastError());
    if (flops_per_launch) *flops_per_launch = 512.0 * 8.0 * (double)iters * 8.0 * (double)grid;      // 512 flop per DMMA, 8 chains, 8 warps per CTA
    return SE3MPC_OK;
}

int se3mpc_kkt_step_device(const se3mpc_params *p, int B, const double *goal, const double *tra_pos, const double *tra_q,
                           const double *t, const double *ulast, const double *w, const double *lam, const double *zL,
                           const double *zU, double mu, double delta_w, double obj_factor, double *dw, double *lam_plus,
                           int32_t *ok, int dump_stage, double *stage_dump, void *workspace, size_t workspace_bytes, void *stream)
{
    if (!params_ok(p) || B < 0 || !goal || !tra_pos || !tra_q || !t || !w || !lam || !zL || !zU || !dw || !lam_plus || !(mu > 0))
        return fail(SE3MPC_ERR_INVALID, "invalid argument");
    if (!have_device()) return fail(SE3MPC_ERR_NO_DEVICE, "no CUDA device");
    if (B == 0) return SE3MPC_OK;
    const size_t need = (size_t)B * se3::SLOT_DOUBLES * sizeof(double);
    if (!workspace || workspace_bytes < need) return fail(SE3MPC_ERR_WORKSPACE, "workspace too small (B * slot bytes)");
    { const int rc_attr = device_setup(); if (rc_attr != SE3MPC_OK) return rc_attr; }
    StepArgs a;
    memset(&a, 0, sizeof(a));
    a.par = *p; a.B = B; a.dump_stage = dump_stage;
    a.goal = goal; a.tra_pos = tra_pos; a.tra_q = tra_q; a.t = t; a.ulast = ulast; a.w = w; a.lam = lam; a.zL = zL; a.zU = zU;
    a.mu = mu; a.delta_w = delta_w; a.obj_factor = obj_factor; a.dw = dw; a.lam_plus = lam_plus; a.dump = stage_dump; a.ok = ok;
    a.slots = static_cast<double *>(workspace);
    se3mpc_kkt_step_kernel<<<(B + WARPS_PER_CTA - 1) / WARPS_PER_CTA, WARPS_PER_CTA * 32, SOLVER_SMEM, static_cast<cudaStream_t>(stream)>>>(a);
    g_launches++;
    CK(cudaGetLastError());
    return SE3MPC_OK;
}

int se3mpc_device_math_device(int n, const double *in3, double *out8, void *stream)
{
    if (n < 0 || !in3 || !out8) return fail(SE3MPC_ERR_INVALID, "invalid argument");
    if (!have_device()) return fail(SE3MPC_ERR_NO_DEVICE, "no CUDA device");
    if (n == 0) return SE3MPC_OK;
    se3mpc_device_math_kernel<<<(n + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(n, in3, out8);
    g_launches++;
    CK(cudaGetLastError());
    return SE3MPC_OK;
}

int se3mpc_collis_det_device(int B, int horizon, const double *V, const double *gate, double *score, int32_t *co, void *stream)
{
    if (B < 0 || horizon < 1 || !V || !gate || !score) return fail(SE3MPC_ERR_INVALID, "invalid argument");
    if (!have_device()) return fail(SE3MPC_ERR_NO_DEVICE, "no CUDA device");
    if (B == 0) return SE3MPC_OK;
    const int warps_per_cta = 8;
    int grid = (B + warps_per_cta - 1) / warps_per_cta;
    const int cap = num_sms() * 8;
    if (grid > cap) grid = cap;
    se3mpc_collis_kernel<<<grid, warps_per_cta * 32, 0, static_cast<cudaStream_t>(stream)>>>(B, horizon, V, gate, score, co);
    g_launches++;
    CK(cudaGetLastError());
    return SE3MPC_OK;
}

int se3mpc_objective_device(const se3mpc_params *p, int B, const double *x0, const double *goal, const double *gate,
                            const double *tra_pos, const double *tra_ang, const double *t, const double *ulast,
                            double *reward, double *X, double *U, double *cost, const se3mpc_stats *stats, void *workspace,
                            size_t workspace_bytes, void *stream)
{
    if (!params_ok(p) || B < 0 || !x0 || !goal || !gate || !tra_pos || !tra_ang || !t || !reward)
        return fail(SE3MPC_ERR_INVALID, "invalid argument");
    if (!have_device()) return fail(SE3MPC_ERR_NO_DEVICE, "no CUDA device");
    return launch_solver(p, MODE_OBJECTIVE, B, 1, x0, goal, gate, tra_pos, tra_ang, 3, t, ulast, X, U, nullptr, cost, reward, stats,
                         nullptr, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

int se3mpc_sol_gradient_device(const se3mpc_params *p, int B, const double *x0, const double *goal, const double *gate,
                               const double *tra_pos, const double *tra_ang, const double *t, const double *ulast,
                               double *out, int32_t *status9, void *workspace, size_t workspace_bytes, void *stream)
{
    if (!params_ok(p) || B < 0 || !x0 || !goal || !gate || !tra_pos || !tra_ang || !t || !out)
        return fail(SE3MPC_ERR_INVALID, "invalid argument");
    if (!have_device()) return fail(SE3MPC_ERR_NO_DEVICE, "no CUDA device");
    if (B == 0) return SE3MPC_OK;
    const int nprob = 9 * B;
    WsLayout L = layout(workspace, nprob);
    if (!workspace || workspace_bytes < L.bytes) return fail(SE3MPC_ERR_WORKSPACE, "workspace too small");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    int rc = launch_solver(p, MODE_SOLGRAD, nprob, 1, x0, goal, gate, tra_pos, tra_ang, 3, t, ulast, nullptr, nullptr, nullptr, nullptr,
                           L.aux_d, nullptr, L.aux_i, workspace, workspace_bytes, s);
    if (rc != SE3MPC_OK) return rc;
    se3mpc_solgrad_epilogue<<<(B + 127) / 128, 128, 0, s>>>(B, L.aux_d, tra_ang, L.aux_i, out, status9);
    g_launches++;
    CK(cudaGetLastError());
    return SE3MPC_OK;
}

// ------------------------------------------------------------------------------------------------------
// host-buffer entry points: stage through cached device buffers
// ------------------------------------------------------------------------------------------------------
}  // extern "C"

namespace {
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    int ensure(size_t n)
    {
        if (n <= cap) return SE3MPC_OK;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMalloc(&p, n);
        if (e != cudaSuccess) return fail(SE3MPC_ERR_CUDA, "cudaMalloc", e);
        cap = n;
        return SE3MPC_OK;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};
// Staging buffers and the stream of the host-buffer entry points: HOST_LANES independent sets per device.  Per device, because a
// buffer or stream created while device A was current must not be used after the caller switched to device B (ops.py wraps calls in
// torch.cuda.device(dev)); several lanes, so that host threads calling concurrently (ctypes releases the GIL) each get their own
// stream + workspace and their launches overlap on the device -- the straggler tail of one batch is then filled by the next batch.
constexpr int HOST_LANES = 4;
struct HostCtx { std::mutex mu; DevBuf in, out, ws; cudaStream_t stream = nullptr; };
HostCtx g_host[MAX_DEVICES][HOST_LANES];
struct HostLease {
    HostCtx *h = nullptr;
    ~HostLease() { if (h) h->mu.unlock(); }
};
HostCtx *host_ctx(HostLease &lease)
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAX_DEVICES) return nullptr;
    for (int l = 0; l < HOST_LANES; l++)
        if (g_host[dev][l].mu.try_lock()) { lease.h = &g_host[dev][l]; return lease.h; }
    g_host[dev][0].mu.lock();           // every lane busy: queue behind lane 0
    lease.h = &g_host[dev][0];
    return lease.h;
}
#define HOST_CTX()                                                                         \
    HostLease lease_;                                                                      \
    HostCtx *hc_ = host_ctx(lease_);                                                       \
    if (!hc_) return fail(SE3MPC_ERR_CUDA, "cudaGetDevice failed or device index out of range"); \
    DevBuf &g_in = hc_->in, &g_out = hc_->out, &g_ws = hc_->ws;                            \
    cudaStream_t &g_stream = hc_->stream;                                                  \
    (void)g_in; (void)g_out; (void)g_ws;                                                   \
    if (!g_stream) CK(cudaStreamCreateWithFlags(&g_stream, cudaStreamNonBlocking))

struct Arena {
    unsigned char *base;
    size_t off = 0;
    template <class T> T *take(size_t n) { T *r = reinterpret_cast<T *>(base + off); off += (n * sizeof(T) + 255) / 256 * 256; return r; }
};
size_t pad(size_t n) { return (n + 255) / 256 * 256; }

int host_common(const se3mpc_params *p, int mode, int B, const double *x0, const double *goal, const double *gate,
                const double *tra_pos, const double *tra_ang, int att_dim, const double *t, const double *ulast, int round_t, double *X,
                double *U, double *LAM, double *cost, double *reward, double *out8, int32_t *status9, const se3mpc_stats *stats)
{
    if (!params_ok(p) || B < 0 || !x0 || !goal || !tra_pos || !tra_ang || !t || (att_dim != 3 && att_dim != 4))
        return fail(SE3MPC_ERR_INVALID, "invalid argument");
    if (mode != MODE_SOLVE && !gate) return fail(SE3MPC_ERR_INVALID, "invalid argument");
    if (!have_device()) return fail(SE3MPC_ERR_NO_DEVICE, "no CUDA device");
    if (B == 0) return SE3MPC_OK;
    HOST_CTX();
    const int N = p->N;
    const int nprob = (mode == MODE_SOLGRAD) ? 9 * B : B;
    const size_t in_bytes = pad((size_t)B * 13 * 8) + 2 * pad((size_t)B * 3 * 8) + pad((size_t)B * 4 * 8) + pad((size_t)B * 12 * 8) + pad((size_t)B * 8) + pad((size_t)B * 4 * 8);
    const size_t out_bytes = pad((size_t)nprob * (N + 1) * 13 * 8) + pad((size_t)nprob * N * 4 * 8) + pad((size_t)nprob * N * 13 * 8) +
                             3 * pad((size_t)nprob * 8) + 4 * pad((size_t)nprob * 4) + pad((size_t)B * 8 * 8) + pad((size_t)B * 9 * 4);
    int rc;
    if ((rc = g_in.ensure(in_bytes)) != SE3MPC_OK) return rc;
    if ((rc = g_out.ensure(out_bytes)) != SE3MPC_OK) return rc;
    const size_t ws_bytes = layout(nullptr, nprob).bytes;
    if ((rc = g_ws.ensure(ws_bytes)) != SE3MPC_OK) return rc;
    Arena ai{static_cast<unsigned char *>(g_in.p)}, ao{static_cast<unsigned char *>(g_out.p)};
    double *d_x0 = ai.take<double>((size_t)B * 13), *d_goal = ai.take<double>((size_t)B * 3), *d_tp = ai.take<double>((size_t)B * 3);
    double *d_ta = ai.take<double>((size_t)B * 4), *d_gate = ai.take<double>((size_t)B * 12), *d_t = ai.take<double>((size_t)B);
    double *d_ul = ai.take<double>((size_t)B * 4);
    double *d_X = ao.take<double>((size_t)nprob * (N + 1) * 13), *d_U = ao.take<double>((size_t)nprob * N * 4);
    double *d_LAM = ao.take<double>((size_t)nprob * N * 13), *d_cost = ao.take<double>((size_t)nprob), *d_rew = ao.take<double>((size_t)nprob);
    double *d_kkt = ao.take<double>((size_t)nprob);
    int32_t *d_st = ao.take<int32_t>((size_t)nprob), *d_it = ao.take<int32_t>((size_t)nprob), *d_nf = ao.take<int32_t>((size_t)nprob);
    int32_t *d_nl = ao.take<int32_t>((size_t)nprob);
    double *d_out8 = ao.take<double>((size_t)B * 8);
    int32_t *d_st9 = ao.take<int32_t>((size_t)B * 9);
    cudaStream_t s = g_stream;
    CK(cudaMemcpyAsync(d_x0, x0, (size_t)B * 13 * 8, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(d_goal, goal, (size_t)B * 3 * 8, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(d_tp, tra_pos, (size_t)B * 3 * 8, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(d_ta, tra_ang, (size_t)B * att_dim * 8, cudaMemcpyHostToDevice, s));
    if (gate) CK(cudaMemcpyAsync(d_gate, gate, (size_t)B * 12 * 8, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(d_t, t, (size_t)B * 8, cudaMemcpyHostToDevice, s));
    if (ulast) CK(cudaMemcpyAsync(d_ul, ulast, (size_t)B * 4 * 8, cudaMemcpyHostToDevice, s));
    se3mpc_stats dst = {d_st, d_it, d_nf, d_nl, d_kkt};
    if (mode == MODE_SOLVE)
        rc = se3mpc_solve_device(p, B, d_x0, d_goal, d_tp, d_ta, att_dim, d_t, ulast ? d_ul : nullptr, round_t, X ? d_X : nullptr,
                                 U ? d_U : nullptr, LAM ? d_LAM : nullptr, d_cost, &dst, g_ws.p, g_ws.cap, s);
    else if (mode == MODE_OBJECTIVE)
        rc = se3mpc_objective_device(p, B, d_x0, d_goal, d_gate, d_tp, d_ta, d_t, ulast ? d_ul : nullptr, d_rew, X ? d_X : nullptr,
                                     U ? d_U : nullptr, d_cost, &dst, g_ws.p, g_ws.cap, s);
    else
        rc = se3mpc_sol_gradient_device(p, B, d_x0, d_goal, d_gate, d_tp, d_ta, d_t, ulast ? d_ul : nullptr, d_out8, d_st9, g_ws.p,
                                        g_ws.cap, s);
    if (rc != SE3MPC_OK) return rc;
    if (mode == MODE_SOLGRAD) {
        CK(cudaMemcpyAsync(out8, d_out8, (size_t)B * 8 * 8, cudaMemcpyDeviceToHost, s));
        if (status9) CK(cudaMemcpyAsync(status9, d_st9, (size_t)B * 9 * 4, cudaMemcpyDeviceToHost, s));
    } else {
        if (X) CK(cudaMemcpyAsync(X, d_X, (size_t)B * (N + 1) * 13 * 8, cudaMemcpyDeviceToHost, s));
        if (U) CK(cudaMemcpyAsync(U, d_U, (size_t)B * N * 4 * 8, cudaMemcpyDeviceToHost, s));
        if (LAM) CK(cudaMemcpyAsync(LAM, d_LAM, (size_t)B * N * 13 * 8, cudaMemcpyDeviceToHost, s));
        if (cost) CK(cudaMemcpyAsync(cost, d_cost, (size_t)B * 8, cudaMemcpyDeviceToHost, s));
        if (reward) CK(cudaMemcpyAsync(reward, d_rew, (size_t)B * 8, cudaMemcpyDeviceToHost, s));
        if (stats) {
            if (stats->status) CK(cudaMemcpyAsync(stats->status, d_st, (size_t)B * 4, cudaMemcpyDeviceToHost, s));
            if (stats->iters) CK(cudaMemcpyAsync(stats->iters, d_it, (size_t)B * 4, cudaMemcpyDeviceToHost, s));
            if (stats->nfact) CK(cudaMemcpyAsync(stats->nfact, d_nf, (size_t)B * 4, cudaMemcpyDeviceToHost, s));
            if (stats->nls) CK(cudaMemcpyAsync(stats->nls, d_nl, (size_t)B * 4, cudaMemcpyDeviceToHost, s));
            if (stats->kkt_error) CK(cudaMemcpyAsync(stats->kkt_error, d_kkt, (size_t)B * 8, cudaMemcpyDeviceToHost, s));
        }
    }
    CK(cudaStreamSynchronize(s));
    return SE3MPC_OK;
}
}  // namespace

extern "C" {

int se3mpc_solve_host(const se3mpc_params *p, int B, const double *x0, const double *goal, const double *tra_pos,
                      const double *tra_ang, int att_dim, const double *t, const double *ulast, int round_t, double *X, double *U,
                      double *LAM, double *cost, const se3mpc_stats *stats)
{
    return host_common(p, MODE_SOLVE, B, x0, goal, nullptr, tra_pos, tra_ang, att_dim, t, ulast, round_t, X, U, LAM, cost, nullptr, nullptr,
                       nullptr, stats);
}

int se3mpc_objective_host(const se3mpc_params *p, int B, const double *x0, const double *goal, const double *gate,
                          const double *tra_pos, const double *tra_ang, const double *t, const double *ulast, double *reward,
                          double *X, double *U, double *cost, const se3mpc_stats *stats)
{
    if (!reward) return fail(SE3MPC_ERR_INVALID, "invalid argument");
    return host_common(p, MODE_OBJECTIVE, B, x0, goal, gate, tra_pos, tra_ang, 3, t, ulast, 1, X, U, nullptr, cost, reward, nullptr,
                       nullptr, stats);
}

int se3mpc_sol_gradient_host(const se3mpc_params *p, int B, const double *x0, const double *goal, const double *gate,
                             const double *tra_pos, const double *tra_ang, const double *t, const double *ulast, double *out,
                             int32_t *status9)
{
    if (!out) return fail(SE3MPC_ERR_INVALID, "invalid argument");
    return host_common(p, MODE_SOLGRAD, B, x0, goal, gate, tra_pos, tra_ang, 3, t, ulast, 1, nullptr, nullptr, nullptr, nullptr, nullptr,
                       out, status9, nullptr);
}

int se3mpc_reward_host(const se3mpc_params *p, int B, const double *X, const double *gate, const double *goal, double *reward,
                       double *collision, double *path, int32_t *co)
{
    if (!params_ok(p) || B < 0 || !X || !gate || !goal || !reward) return fail(SE3MPC_ERR_INVALID, "invalid argument");
    if (!have_device()) return fail(SE3MPC_ERR_NO_DEVICE, "no CUDA device");
    if (B == 0) return SE3MPC_OK;
    HOST_CTX();
    const int N = p->N;
    int rc;
    const size_t xb = (size_t)B * (N + 1) * 13 * 8;
    if ((rc = g_in.ensure(pad(xb) + pad((size_t)B * 12 * 8) + pad((size_t)B * 3 * 8))) != SE3MPC_OK) return rc;
    if ((rc = g_out.ensure(3 * pad((size_t)B * 8) + pad((size_t)B * 4))) != SE3MPC_OK) return rc;
    Arena ai{static_cast<unsigned char *>(g_in.p)}, ao{static_cast<unsigned char *>(g_out.p)};
    double *d_X = ai.take<double>((size_t)B * (N + 1) * 13), *d_gate = ai.take<double>((size_t)B * 12), *d_goal = ai.take<double>((size_t)B * 3);
    double *d_r = ao.take<double>(B), *d_c = ao.take<double>(B), *d_p = ao.take<double>(B);
    int32_t *d_co = ao.take<int32_t>(B);
    cudaStream_t s = g_stream;
    CK(cudaMemcpyAsync(d_X, X, xb, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(d_gate, gate, (size_t)B * 12 * 8, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(d_goal, goal, (size_t)B * 3 * 8, cudaMemcpyHostToDevice, s));
    rc = se3mpc_reward_device(p, B, d_X, d_gate, d_goal, d_r, d_c, d_p, d_co, s);
    if (rc != SE3MPC_OK) return rc;
    CK(cudaMemcpyAsync(reward, d_r, (size_t)B * 8, cudaMemcpyDeviceToHost, s));
    if (collision) CK(cudaMemcpyAsync(collision, d_c, (size_t)B * 8, cudaMemcpyDeviceToHost, s));
    if (path) CK(cudaMemcpyAsync(path, d_p, (size_t)B * 8, cudaMemcpyDeviceToHost, s));
    if (co) CK(cudaMemcpyAsync(co, d_co, (size_t)B * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return SE3MPC_OK;
}

int se3mpc_collis_det_host(int B, int horizon, const double *V, const double *gate, double *score, int32_t *co)
{
    if (B < 0 || horizon < 1 || !V || !gate || !score) return fail(SE3MPC_ERR_INVALID, "invalid argument");
    if (!have_device()) return fail(SE3MPC_ERR_NO_DEVICE, "no CUDA device");
    if (B == 0) return SE3MPC_OK;
    HOST_CTX();
    int rc;
    const size_t vb = (size_t)B * (horizon + 1) * 3 * 8;
    if ((rc = g_in.ensure(pad(vb) + pad((size_t)B * 12 * 8))) != SE3MPC_OK) return rc;
    if ((rc = g_out.ensure(pad((size_t)B * 8) + pad((size_t)B * 4))) != SE3MPC_OK) return rc;
    Arena ai{static_cast<unsigned char *>(g_in.p)}, ao{static_cast<unsigned char *>(g_out.p)};
    double *d_V = ai.take<double>((size_t)B * (horizon + 1) * 3), *d_gate = ai.take<double>((size_t)B * 12);
    double *d_s = ao.take<double>(B);
    int32_t *d_co = ao.take<int32_t>(B);
    cudaStream_t s = g_stream;
    CK(cudaMemcpyAsync(d_V, V, vb, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(d_gate, gate, (size_t)B * 12 * 8, cudaMemcpyHostToDevice, s));
    rc = se3mpc_collis_det_device(B, horizon, d_V, d_gate, d_s, d_co, s);
    if (rc != SE3MPC_OK) return rc;
    CK(cudaMemcpyAsync(score, d_s, (size_t)B * 8, cudaMemcpyDeviceToHost, s));
    if (co) CK(cudaMemcpyAsync(co, d_co, (size_t)B * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return SE3MPC_OK;
}

int se3mpc_gate_frame_device(int B, const double *gate, const double *velo, const double *w, const double *t, const double *state,
                             const double *final_point, double *inputs18, double *gate_pred, void *stream)
{
    if (B < 0 || !gate || !velo || !w || !t || !state || !final_point || !inputs18) return fail(SE3MPC_ERR_INVALID, "invalid argument");
    if (!have_device()) return fail(SE3MPC_ERR_NO_DEVICE, "no CUDA device");
    if (B == 0) return SE3MPC_OK;
    se3mpc_gate_frame_kernel<<<(B + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(B, gate, velo, w, t, state, final_point, inputs18, gate_pred);
    g_launches++;
    CK(cudaGetLastError());
    return SE3MPC_OK;
}

int se3mpc_plant_step_device(const se3mpc_params *p, int B, double dt, int nsteps, const double *u, double *state, void *stream)
{
    if (!params_ok(p) || B < 0 || !(dt > 0) || nsteps < 0 || !u || !state) return fail(SE3MPC_ERR_INVALID, "invalid argument");
    if (!have_device()) return fail(SE3MPC_ERR_NO_DEVICE, "no CUDA device");
    if (B == 0 || nsteps == 0) return SE3MPC_OK;
    se3mpc_plant_step_kernel<<<(B + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(*p, B, dt, nsteps, u, state);
    g_launches++;
    CK(cudaGetLastError());
    return SE3MPC_OK;
}

int se3mpc_episode_advance_device(const se3mpc_params *p, int B, double dt, int nsteps, const double *u, double *state, double *gate,
                                  const double *velo, const double *w, const double *noise, void *stream)
{
    if (!params_ok(p) || B < 0 || !(dt > 0) || nsteps < 0 || !u || !state || !gate || !velo || !w) return fail(SE3MPC_ERR_INVALID, "invalid argument");
    if (!have_device()) return fail(SE3MPC_ERR_NO_DEVICE, "no CUDA device");
    if (B == 0 || nsteps == 0) return SE3MPC_OK;
    se3mpc_episode_advance_kernel<<<(B + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(*p, B, dt, nsteps, u, state, gate, velo, w, noise);
    g_launches++;
    CK(cudaGetLastError());
    return SE3MPC_OK;
}

int se3mpc_rotor_positions_device(int nstates, double wing_len, const double *X, double *pos, void *stream)
{
    if (nstates < 0 || !X || !pos) return fail(SE3MPC_ERR_INVALID, "invalid argument");
    if (!have_device()) return fail(SE3MPC_ERR_NO_DEVICE, "no CUDA device");
    if (nstates == 0) return SE3MPC_OK;
    se3mpc_rotor_positions_kernel<<<(nstates + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(nstates, wing_len * 0.5 / sqrt(2.0), X, pos);
    g_launches++;
    CK(cudaGetLastError());
    return SE3MPC_OK;
}

int se3mpc_rotor_positions_host(int nstates, double wing_len, const double *X, double *pos)
{
    if (nstates < 0 || !X || !pos) return fail(SE3MPC_ERR_INVALID, "invalid argument");
    if (!have_device()) return fail(SE3MPC_ERR_NO_DEVICE, "no CUDA device");
    if (nstates == 0) return SE3MPC_OK;
    HOST_CTX();
    int rc;
    if ((rc = g_in.ensure(pad((size_t)nstates * 13 * 8))) != SE3MPC_OK) return rc;
    if ((rc = g_out.ensure(pad((size_t)nstates * 15 * 8))) != SE3MPC_OK) return rc;
    double *d_x = static_cast<double *>(g_in.p), *d_p = static_cast<double *>(g_out.p);
    cudaStream_t s = g_stream;
    CK(cudaMemcpyAsync(d_x, X, (size_t)nstates * 13 * 8, cudaMemcpyHostToDevice, s));
    rc = se3mpc_rotor_positions_device(nstates, wing_len, d_x, d_p, s);
    if (rc != SE3MPC_OK) return rc;
    CK(cudaMemcpyAsync(pos, d_p, (size_t)nstates * 15 * 8, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return SE3MPC_OK;
}

int se3mpc_gate_frame_host(int B, const double *gate, const double *velo, const double *w, const double *t, const double *state,
                           const double *final_point, double *inputs18, double *gate_pred)
{
    if (B < 0 || !gate || !velo || !w || !t || !state || !final_point || !inputs18) return fail(SE3MPC_ERR_INVALID, "invalid argument");
    if (!have_device()) return fail(SE3MPC_ERR_NO_DEVICE, "no CUDA device");
    if (B == 0) return SE3MPC_OK;
    HOST_CTX();
    int rc;
    if ((rc = g_in.ensure(pad((size_t)B * 12 * 8) + 2 * pad((size_t)B * 3 * 8) + 2 * pad((size_t)B * 8) + pad((size_t)B * 13 * 8))) != SE3MPC_OK) return rc;
    if ((rc = g_out.ensure(pad((size_t)B * 18 * 8) + pad((size_t)B * 12 * 8))) != SE3MPC_OK) return rc;
    Arena ai{static_cast<unsigned char *>(g_in.p)}, ao{static_cast<unsigned char *>(g_out.p)};
    double *d_g = ai.take<double>((size_t)B * 12), *d_v = ai.take<double>((size_t)B * 3), *d_f = ai.take<double>((size_t)B * 3);
    double *d_w = ai.take<double>(B), *d_t = ai.take<double>(B), *d_x = ai.take<double>((size_t)B * 13);
    double *d_o = ao.take<double>((size_t)B * 18), *d_gp = ao.take<double>((size_t)B * 12);
    cudaStream_t s = g_stream;
    CK(cudaMemcpyAsync(d_g, gate, (size_t)B * 12 * 8, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(d_v, velo, (size_t)B * 3 * 8, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(d_f, final_point, (size_t)B * 3 * 8, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(d_w, w, (size_t)B * 8, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(d_t, t, (size_t)B * 8, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(d_x, state, (size_t)B * 13 * 8, cudaMemcpyHostToDevice, s));
    rc = se3mpc_gate_frame_device(B, d_g, d_v, d_w, d_t, d_x, d_f, d_o, gate_pred ? d_gp : nullptr, s);
    if (rc != SE3MPC_OK) return rc;
    CK(cudaMemcpyAsync(inputs18, d_o, (size_t)B * 18 * 8, cudaMemcpyDeviceToHost, s));
    if (gate_pred) CK(cudaMemcpyAsync(gate_pred, d_gp, (size_t)B * 12 * 8, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return SE3MPC_OK;
}

int se3mpc_plant_step_host(const se3mpc_params *p, int B, double dt, int nsteps, const double *u, double *state)
{
    if (!params_ok(p) || B < 0 || !(dt > 0) || nsteps < 0 || !u || !state) return fail(SE3MPC_ERR_INVALID, "invalid argument");
    if (!have_device()) return fail(SE3MPC_ERR_NO_DEVICE, "no CUDA device");
    if (B == 0 || nsteps == 0) return SE3MPC_OK;
    HOST_CTX();
    int rc;
    if ((rc = g_in.ensure(pad((size_t)B * 4 * 8))) != SE3MPC_OK) return rc;
    if ((rc = g_out.ensure(pad((size_t)B * 13 * 8))) != SE3MPC_OK) return rc;
    double *d_u = static_cast<double *>(g_in.p), *d_x = static_cast<double *>(g_out.p);
    cudaStream_t s = g_stream;
    CK(cudaMemcpyAsync(d_u, u, (size_t)B * 4 * 8, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(d_x, state, (size_t)B * 13 * 8, cudaMemcpyHostToDevice, s));
    rc = se3mpc_plant_step_device(p, B, dt, nsteps, d_u, d_x, s);
    if (rc != SE3MPC_OK) return rc;
    CK(cudaMemcpyAsync(state, d_x, (size_t)B * 13 * 8, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return SE3MPC_OK;
}

void se3mpc_host_release(void)
{
    int cur = 0, n = 0;
    if (cudaGetDevice(&cur) != cudaSuccess || cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return; }
    for (int d = 0; d < n && d < MAX_DEVICES; d++)
        for (int l = 0; l < HOST_LANES; l++) {
            HostCtx &h = g_host[d][l];
            std::lock_guard<std::mutex> lk(h.mu);
            if (!h.in.p && !h.out.p && !h.ws.p && !h.stream) continue;
            cudaSetDevice(d);
            h.in.release(); h.out.release(); h.ws.release();
            if (h.stream) { cudaStreamDestroy(h.stream); h.stream = nullptr; }
        }
    cudaSetDevice(cur);
}

}  // extern "C"
