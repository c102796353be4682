// dps_backward.cu -- discrete adjoint of the stepping kernel (sm_100a) and its C-ABI entry points.
//
// What it replaces: the reference obtains gradients by recording every torch op of the time loop
// (simulator.py:523-543) and calling loss.sum().backward() (optimization_lib/ps_optimization.py:212) -- a graph of
// 1e5..1e6 nodes per run.  Here the adjoint of the *exact* discrete update is written by hand:
//
//   forward (dps_forward)  stores the states every `ckpt_stride` steps;
//   backward               walks the checkpoints from the last to the first; for each segment it re-integrates the
//                          segment forward (states of every step go to an HBM scratch, batch innermost), then sweeps
//                          the segment backwards.  One step of the sweep recomputes the stage inputs X_1..X_s and their
//                          network voltages, then applies the transposed stage Jacobians in reverse order
//                              dk_i = b_i dt lambda + c_{i+1} dt dX_{i+1},   dX_i = J(X_i)^T dk_i,   lambda <- lambda + sum_i dX_i
//                          (Euler b=[1]; Heun c=[0,1], b=[1/2,1/2]; RK4 c=[0,1/2,1/2,1], b=[1/6,1/3,1/3,1/6]).
//                          The network couples the generators of one trajectory in both directions: currents are
//                          exchanged through shared memory for V = Z I, voltage cotangents for dI = Z^T dV.
//
// Thread mapping is the forward kernel's: one thread per (trajectory, generator), a warp = 32 trajectories of one
// generator, named barrier per 32-trajectory group.  Gradients are produced for x0, for every constant slot the host
// marks as wanted and -- on request -- for the reduced impedance matrices (dL/dZsub per trajectory; the host chains it
// through the fault re-reduction and inv(Y_dyn) to x_d'', S_n, load parameters: SURVEY.md section 8f.1).
#include <cuda_runtime.h>
#include <string.h>
#include "dps_internal.h"
#include "dps_adjoint.cuh"
#include "dps_plan.h"

namespace dps {

struct BwdParams {
    long long P;
    int n_gen, solver, n_steps;
    double dt, fn;
    const double* upar;
    const double* vpar;
    const double* zsub;
    const double* ckpt;        // [n_ckpt][n_gen][13][P]
    const double* grad_rec;    // [n_rec][n_steps][P]
    const double* grad_xf;     // [n_gen][13][P] or null
    double* g_x0;              // [n_gen][13][P]
    double* gslot;             // [n_want][P]
    double* g_zsub;            // [n_cfg][z_rows][z_cols][2][P] accumulated, or null
    double* scratch;           // [ckpt_stride][n_gen][13][P]
    int n_obs, z_rows, z_cols, z_per_sim, ckpt_stride, n_ckpt, n_seg;
    int seg_step[MAX_SEG];
    int seg_cfg[MAX_SEG];
    int ctrl[MAX_G];
    unsigned rec_mask[MAX_G];
    short rec_chan[MAX_G][DPS_N_FIELDS];
    short obs_chan[MAX_EXTRA_OBS][2];
    short slot_vary[MAX_G * DPS_N_SLOTS];
    short want_row[MAX_G * DPS_N_SLOTS];
};

template <int G>
struct BwdCfg {
    static constexpr int GROUPS = (G >= 3) ? 1 : (G == 2 ? 2 : 4);
    static constexpr int THREADS = 32 * G * GROUPS;
};

__device__ __forceinline__ void bwd_barrier(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

template <int G, bool CTRL, bool ZGRAD>
__global__ void __launch_bounds__(BwdCfg<G>::THREADS, 1)
dps_adjoint_kernel(const BwdParams p) {
    constexpr int GROUPS = BwdCfg<G>::GROUPS;
    constexpr int NS = DPS_N_STATES;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* s_upar = reinterpret_cast<double*>(smem_raw);                        // [G][DPS_N_SLOTS]
    double* s_xch = s_upar + G * DPS_N_SLOTS;                                    // [2][GROUPS][G][2][32]
    short* s_vary = reinterpret_cast<short*>(s_xch + 2 * GROUPS * G * 64);       // [G][DPS_N_SLOTS]

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int gw = warp % G;
    const int grp = warp / G;
    const bool gen_live = gw < p.n_gen;
    const int g = gen_live ? gw : p.n_gen - 1;
    const long long P = p.P;
    const long long sim = ((long long)blockIdx.x * GROUPS + grp) * 32 + lane;
    const bool live = (sim < P) && gen_live;
    const long long s = (sim < P) ? sim : (P - 1);

    for (int i = threadIdx.x; i < G * DPS_N_SLOTS; i += blockDim.x) {
        const bool ok = (i / DPS_N_SLOTS) < p.n_gen;
        s_upar[i] = ok ? p.upar[i] : 0.0;
        s_vary[i] = ok ? p.slot_vary[i] : (short)-1;
    }
    __syncthreads();

    ParamReader<true> par;
    par.up = s_upar + g * DPS_N_SLOTS;
    par.sv = s_vary + g * DPS_N_SLOTS;
    par.vpar = p.vpar;
    par.P = P;
    par.s = s;
    const int ctrl = CTRL ? p.ctrl[g] : 0;
    const long long zb = p.z_per_sim ? P : 1;
    const long long zs = p.z_per_sim ? s : 0;
    const int bar_id = grp + 1;
    double* xch_grp = s_xch + (long long)grp * G * 64;
    int buf = 0;

    // exchange (a, b) with the other generators of this trajectory; afterwards xb[j*64 + lane] / [j*64+32+lane] hold
    // generator j's pair
    auto exchange = [&](double a, double b) -> const double* {
        double* xb = xch_grp + buf * (GROUPS * G * 64);
        xb[gw * 64 + lane] = gen_live ? a : 0.0;
        xb[gw * 64 + 32 + lane] = gen_live ? b : 0.0;
        bwd_barrier(bar_id, 32 * G);
        buf ^= 1;
        return xb;
    };
    auto z_at = [&](int cfg, int row, int col, double& zr, double& zi) {
        const double* e = p.zsub + ((((long long)cfg * p.z_rows + row) * p.z_cols + col) * 2) * zb + zs;
        zr = __ldg(e);
        zi = __ldg(e + zb);
    };
    // one stage evaluation at X: sin/cos, network voltage, derivatives
    auto eval_stage = [&](const double* X, int cfg, double& sd, double& cd, double& v_re, double& v_im, double* f,
                          Algebraics& alg) {
        dps_sincos(X[DPS_X_DELTA], &sd, &cd);
        double i_re, i_im;
        injection(X, par, sd, cd, i_re, i_im);
        const double* xb = exchange(i_re, i_im);
        v_re = 0.0;
        v_im = 0.0;
        for (int j = 0; j < p.n_gen; ++j) {
            double zr, zi;
            z_at(cfg, g, j, zr, zi);
            const double jr = xb[j * 64 + lane], ji = xb[j * 64 + 32 + lane];
            v_re += zr * jr - zi * ji;
            v_im += zr * ji + zi * jr;
        }
        stage_derivatives<CTRL>(X, par, ctrl, p.fn, sd, cd, v_re, v_im, f, alg);
    };
    auto cfg_of = [&](int step) {
        int seg = 0;
        while (seg + 1 < p.n_seg && p.seg_step[seg + 1] <= step) ++seg;
        return p.seg_cfg[seg];
    };

    const int n_stages = (p.solver == DPS_EULER) ? 1 : (p.solver == DPS_HEUN ? 2 : 4);
    const double dt = p.dt;
    double cst[4], bst[4];                     // c_i, b_i of the explicit scheme
    if (p.solver == DPS_EULER) { cst[0] = 0.0; bst[0] = 1.0; cst[1] = cst[2] = cst[3] = bst[1] = bst[2] = bst[3] = 0.0; }
    else if (p.solver == DPS_HEUN) { cst[0] = 0.0; cst[1] = 1.0; bst[0] = bst[1] = 0.5; cst[2] = cst[3] = bst[2] = bst[3] = 0.0; }
    else { cst[0] = 0.0; cst[1] = 0.5; cst[2] = 0.5; cst[3] = 1.0; bst[0] = 1.0 / 6.0; bst[1] = 1.0 / 3.0; bst[2] = 1.0 / 3.0; bst[3] = 1.0 / 6.0; }

    double lam[NS], gp[DPS_N_SLOTS];
#pragma unroll
    for (int k = 0; k < NS; ++k)
        lam[k] = (p.grad_xf != nullptr) ? p.grad_xf[((long long)g * NS + k) * P + s] : 0.0;
#pragma unroll
    for (int k = 0; k < DPS_N_SLOTS; ++k) gp[k] = 0.0;

    const unsigned rec_mask = p.rec_mask[g];
    const short* ch = p.rec_chan[g];
    const long long cs = (long long)p.n_steps * P;
    const int C = p.ckpt_stride;

    for (int ck = p.n_ckpt - 1; ck >= 0; --ck) {
        const int k0 = ck * C;
        const int k1 = (k0 + C < p.n_steps) ? (k0 + C) : p.n_steps;
        // ---- re-integrate the segment, keeping the state at the start of every step ----------------------------------
        {
            double x[NS], xs[NS], acc[NS], f[NS];
            const double* src = p.ckpt + (((long long)ck * p.n_gen + g) * NS) * P + s;
#pragma unroll
            for (int k = 0; k < NS; ++k) x[k] = src[(long long)k * P];
            for (int step = k0; step < k1; ++step) {
                if (live) {
                    double* dst = p.scratch + (((long long)(step - k0) * p.n_gen + g) * NS) * P + s;
#pragma unroll
                    for (int k = 0; k < NS; ++k) dst[(long long)k * P] = x[k];
                }
                if (step == k1 - 1) break;
                const int cfg = cfg_of(step);
#pragma unroll
                for (int k = 0; k < NS; ++k) xs[k] = x[k];
                for (int st = 0; st < n_stages; ++st) {
                    double sd, cd, v_re, v_im;
                    Algebraics alg;
                    eval_stage(x, cfg, sd, cd, v_re, v_im, f, alg);
                    if (p.solver == DPS_EULER) {
#pragma unroll
                        for (int k = 0; k < NS; ++k) x[k] = xs[k] + f[k] * dt;
                    } else if (p.solver == DPS_HEUN) {
                        if (st == 0) {
#pragma unroll
                            for (int k = 0; k < NS; ++k) { acc[k] = f[k]; x[k] = xs[k] + f[k] * dt; }
                        } else {
#pragma unroll
                            for (int k = 0; k < NS; ++k) x[k] = xs[k] + ((acc[k] + f[k]) * 0.5) * dt;
                        }
                    } else {
                        if (st == 0) {
#pragma unroll
                            for (int k = 0; k < NS; ++k) { acc[k] = f[k]; x[k] = xs[k] + (f[k] * dt) * 0.5; }
                        } else if (st == 1) {
#pragma unroll
                            for (int k = 0; k < NS; ++k) { acc[k] = acc[k] + 2.0 * f[k]; x[k] = xs[k] + (f[k] * dt) * 0.5; }
                        } else if (st == 2) {
#pragma unroll
                            for (int k = 0; k < NS; ++k) { acc[k] = acc[k] + 2.0 * f[k]; x[k] = xs[k] + f[k] * dt; }
                        } else {
#pragma unroll
                            for (int k = 0; k < NS; ++k) x[k] = xs[k] + dps_div6((acc[k] + f[k]) * dt);
                        }
                    }
                }
            }
        }
        // ---- backward sweep over the segment --------------------------------------------------------------------------
        for (int step = k1 - 1; step >= k0; --step) {
            const int cfg = cfg_of(step);
            double X[4][NS], sds[4], cds[4], vres[4], vims[4];
            {
                const double* src = p.scratch + (((long long)(step - k0) * p.n_gen + g) * NS) * P + s;
#pragma unroll
                for (int k = 0; k < NS; ++k) X[0][k] = src[(long long)k * P];
            }
            // recorded post-step states: rec[.., step] = x_{step+1}
            if (live && rec_mask) {
                const double* gr = p.grad_rec + (long long)step * P + sim;
#pragma unroll
                for (int k = 0; k < NS; ++k)
                    if (rec_mask & (1u << k)) lam[k] += gr[ch[k] * cs];
            }
            // stage inputs and their voltages
            for (int i = 0; i < n_stages; ++i) {
                double f[NS];
                Algebraics alg;
                eval_stage(X[i], cfg, sds[i], cds[i], vres[i], vims[i], f, alg);
                if (i + 1 < n_stages) {
                    const double ci = cst[i + 1];
#pragma unroll
                    for (int k = 0; k < NS; ++k) X[i + 1][k] = X[0][k] + ci * (f[k] * dt);
                }
            }
            // transposed stages, last to first
            double sum_dx[NS], dxn[NS];
#pragma unroll
            for (int k = 0; k < NS; ++k) { sum_dx[k] = 0.0; dxn[k] = 0.0; }
            for (int i = n_stages - 1; i >= 0; --i) {
                double dk[NS], dx[NS];
                const double bi = bst[i] * dt;
                const double cn = (i + 1 < n_stages) ? cst[i + 1] * dt : 0.0;
#pragma unroll
                for (int k = 0; k < NS; ++k) { dk[k] = bi * lam[k] + cn * dxn[k]; dx[k] = 0.0; }
                AlgCot ac = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
                if (i == n_stages - 1 && live && (rec_mask >> NS)) {
                    const double* gr = p.grad_rec + (long long)step * P + sim;
                    double* av[7] = {&ac.p_e, &ac.i_d, &ac.i_q, &ac.v_re, &ac.v_im, &ac.e_fd, &ac.p_m};
#pragma unroll
                    for (int k = 0; k < 7; ++k)
                        if (rec_mask & (1u << (NS + k))) *av[k] = gr[ch[NS + k] * cs];
                }
                double dsd, dcd, dv_re, dv_im;
                stage_vjp<CTRL>(X[i], par, ctrl, p.fn, sds[i], cds[i], vres[i], vims[i], dk, ac, dx, gp, dsd, dcd, dv_re, dv_im);
                // dI = Z^T dV over the generator rows ...
                const double* xb = exchange(dv_re, dv_im);
                double di_re = 0.0, di_im = 0.0;
                // ... and, when the network carries a gradient, dZ[r][g] += dV_r conj(I_g): this thread owns column g of
                // its trajectory's matrix, so the accumulation in HBM needs no atomics.  I_g is recomputed (6 flops).
                double own_re = 0.0, own_im = 0.0;
                double* gz = nullptr;
                if (ZGRAD) {
                    injection(X[i], par, sds[i], cds[i], own_re, own_im);
                    gz = p.g_zsub + ((((long long)cfg * p.z_rows) * p.z_cols + g) * 2) * P + s;
                }
                const long long gz_row = (long long)p.z_cols * 2 * P;
                for (int r = 0; r < p.n_gen; ++r) {
                    double zr, zi;
                    z_at(cfg, r, g, zr, zi);
                    const double ar = xb[r * 64 + lane], ai = xb[r * 64 + 32 + lane];
                    di_re += zr * ar + zi * ai;
                    di_im += -zi * ar + zr * ai;
                    if (ZGRAD && live) {
                        gz[r * gz_row] += ar * own_re + ai * own_im;
                        gz[r * gz_row + P] += -ar * own_im + ai * own_re;
                    }
                }
                // ... plus the recorded voltages of extra observed buses (last stage of the step)
                if (i == n_stages - 1 && p.n_obs > p.n_gen && sim < P) {
                    for (int r = p.n_gen; r < p.n_obs; ++r) {
                        const int e = r - p.n_gen;
                        const double ar = (p.obs_chan[e][0] >= 0) ? p.grad_rec[((long long)p.obs_chan[e][0] * p.n_steps + step) * P + sim] : 0.0;
                        const double ai = (p.obs_chan[e][1] >= 0) ? p.grad_rec[((long long)p.obs_chan[e][1] * p.n_steps + step) * P + sim] : 0.0;
                        double zr, zi;
                        z_at(cfg, r, g, zr, zi);
                        di_re += zr * ar + zi * ai;
                        di_im += -zi * ar + zr * ai;
                        if (ZGRAD && live) {
                            gz[r * gz_row] += ar * own_re + ai * own_im;
                            gz[r * gz_row + P] += -ar * own_im + ai * own_re;
                        }
                    }
                }
                injection_vjp(X[i], par, sds[i], cds[i], di_re, di_im, dx, gp, dsd, dcd);
                dx[DPS_X_DELTA] += dsd * cds[i] - dcd * sds[i];
#pragma unroll
                for (int k = 0; k < NS; ++k) { dxn[k] = dx[k]; sum_dx[k] += dx[k]; }
            }
#pragma unroll
            for (int k = 0; k < NS; ++k) lam[k] += sum_dx[k];
        }
    }

    if (live) {
        double* dst = p.g_x0 + (long long)g * NS * P + s;
#pragma unroll
        for (int k = 0; k < NS; ++k) dst[(long long)k * P] = lam[k];
        const short* wr = p.want_row + g * DPS_N_SLOTS;
#pragma unroll
        for (int k = 0; k < DPS_N_SLOTS; ++k)
            if (wr[k] >= 0) p.gslot[(long long)wr[k] * P + s] = gp[k];
    }
}

template <int G, bool CTRL, bool ZGRAD>
static cudaError_t launch_bwd_one(const BwdParams& p, cudaStream_t stream) {
    using C = BwdCfg<G>;
    const size_t smem = sizeof(double) * (G * DPS_N_SLOTS + 2 * C::GROUPS * G * 64) + sizeof(short) * G * DPS_N_SLOTS;
    const long long spb = 32LL * C::GROUPS;
    const long long blocks = (p.P + spb - 1) / spb;
    dps_adjoint_kernel<G, CTRL, ZGRAD><<<(unsigned)blocks, C::THREADS, smem, stream>>>(p);
    return cudaGetLastError();
}

template <int G>
static cudaError_t launch_bwd_g(const BwdParams& p, bool ctrl, cudaStream_t stream) {
    if (p.g_zsub != nullptr) return ctrl ? launch_bwd_one<G, true, true>(p, stream) : launch_bwd_one<G, false, true>(p, stream);
    return ctrl ? launch_bwd_one<G, true, false>(p, stream) : launch_bwd_one<G, false, false>(p, stream);
}

static cudaError_t launch_backward(const BwdParams& p, int g_template, bool ctrl, cudaStream_t stream) {
    switch (g_template) {
        case 1: return launch_bwd_g<1>(p, ctrl, stream);
        case 2: return launch_bwd_g<2>(p, ctrl, stream);
        case 3: return launch_bwd_g<3>(p, ctrl, stream);
        case 4: return launch_bwd_g<4>(p, ctrl, stream);
        case 5: return launch_bwd_g<5>(p, ctrl, stream);
        case 6: return launch_bwd_g<6>(p, ctrl, stream);
        case 8: return launch_bwd_g<8>(p, ctrl, stream);
        case 10: return launch_bwd_g<10>(p, ctrl, stream);
        case 16: return launch_bwd_g<16>(p, ctrl, stream);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace dps

using namespace dps;

extern "C" {

size_t dps_workspace_bytes(const dps_plan* plan, int64_t P) {
    if (!plan || plan->topo.ckpt_stride <= 0 || P < 1) return 0;
    return sizeof(double) * (size_t)plan->topo.ckpt_stride * plan->topo.n_gen * DPS_N_STATES * (size_t)P;
}

int dps_backward(const dps_plan* plan, int64_t P, const double* upar, const double* vpar, const double* zsub,
                 const double* ckpt, const double* grad_rec, const double* grad_x_final, const uint8_t* want,
                 double* g_x0, double* gslot, double* g_zsub, void* workspace, size_t workspace_bytes, void* stream) {
    if (!plan) { set_error("dps_backward: null plan"); return 1; }
    if (P < 1) { set_error("dps_backward: P must be >= 1"); return 1; }
    if (plan->topo.ckpt_stride <= 0) { set_error("dps_backward: the plan was created without checkpoints (ckpt_stride = 0)"); return 1; }
    if (!upar || !zsub || !ckpt || !want || !g_x0) { set_error("dps_backward: null input pointer"); return 1; }
    if (plan->topo.n_rec > 0 && !grad_rec) { set_error("dps_backward: grad_rec is null"); return 1; }
    if (plan->any_vary && !vpar) { set_error("dps_backward: per-sim constants requested but vpar is null"); return 1; }
    if ((int)plan->seg_step.size() > MAX_SEG) { set_error("dps_backward: more than %d configuration changes are not supported by the adjoint", MAX_SEG); return 1; }
    const size_t need = dps_workspace_bytes(plan, P);
    if (!workspace || workspace_bytes < need) { set_error("dps_backward: workspace too small (%zu < %zu bytes)", workspace_bytes, need); return 1; }

    BwdParams p;
    memset(&p, 0, sizeof(p));
    const FwdParams& b = plan->base;
    p.P = P; p.n_gen = b.n_gen; p.solver = b.solver; p.n_steps = b.n_steps; p.dt = b.dt; p.fn = b.fn;
    p.upar = upar; p.vpar = vpar; p.zsub = zsub; p.ckpt = ckpt; p.grad_rec = grad_rec; p.grad_xf = grad_x_final;
    p.g_x0 = g_x0; p.gslot = gslot; p.g_zsub = g_zsub; p.scratch = (double*)workspace;
    p.n_obs = b.n_obs; p.z_rows = b.z_rows; p.z_cols = b.z_cols; p.z_per_sim = b.z_per_sim;
    p.ckpt_stride = b.ckpt_stride; p.n_ckpt = (int)dps_n_checkpoints(plan);
    p.n_seg = (int)plan->seg_step.size();
    for (int i = 0; i < p.n_seg; ++i) { p.seg_step[i] = plan->seg_step[i]; p.seg_cfg[i] = plan->seg_cfg[i]; }
    memcpy(p.ctrl, b.ctrl, sizeof(p.ctrl));
    memcpy(p.rec_mask, b.rec_mask, sizeof(p.rec_mask));
    memcpy(p.rec_chan, b.rec_chan, sizeof(p.rec_chan));
    memcpy(p.obs_chan, b.obs_chan, sizeof(p.obs_chan));
    memcpy(p.slot_vary, b.slot_vary, sizeof(p.slot_vary));
    int n_want = 0;
    for (int i = 0; i < MAX_G * DPS_N_SLOTS; ++i) p.want_row[i] = -1;
    for (int i = 0; i < b.n_gen * DPS_N_SLOTS; ++i)
        if (want[i]) p.want_row[i] = (short)n_want++;
    if (n_want > 0 && !gslot) { set_error("dps_backward: slot gradients wanted but gslot is null"); return 1; }
    if (p.n_steps == 0) return 0;
    cudaError_t e = launch_backward(p, plan->g_template, plan->any_ctrl, (cudaStream_t)stream);
    if (e != cudaSuccess) { set_error("dps_backward: launch failed: %s", cudaGetErrorString(e)); return 3; }
    return 0;
}

}  // extern "C"
