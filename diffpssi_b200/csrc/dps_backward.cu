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
// Thread mapping is the forward kernel's (dps_forward.cuh): one thread per (trajectory, generator), the generators of a
// trajectory in adjacent lanes of one warp, exchanges through the warp's private shared-memory line.  Gradients are
// produced for x0, for every constant slot the host marks as wanted and -- on request -- for the reduced impedance
// matrices (dL/dZsub per trajectory; the host chains it through the fault re-reduction and inv(Y_dyn) to x_d'', S_n, load
// parameters: SURVEY.md section 8f.1).  Kernel: dps_backward.cuh (one translation unit per generator count).
#include <cuda_runtime.h>
#include <string.h>
#include "dps_internal.h"
#include "dps_plan.h"
#include "dps_backward_params.h"

namespace dps {

template <int G>
cudaError_t launch_bwd_g(const BwdParams& p, int cmask, int ngp, cudaStream_t stream);
#define DPS_EXTERN(G) extern template cudaError_t launch_bwd_g<G>(const BwdParams&, int, int, cudaStream_t);
DPS_EXTERN(1) DPS_EXTERN(2) DPS_EXTERN(3) DPS_EXTERN(4) DPS_EXTERN(5) DPS_EXTERN(6) DPS_EXTERN(8) DPS_EXTERN(10) DPS_EXTERN(16)
#undef DPS_EXTERN

cudaError_t launch_bwd_wide(const BwdWideParams& p, int cmask, int ngp, cudaStream_t stream);

static cudaError_t launch_backward(const BwdParams& p, int g_template, int ctrl, int ngp, cudaStream_t stream) {
#ifdef DPS_ONLY_G10      // kernel A/B experiment builds (tools/build_variant.sh)
    return g_template == 10 ? launch_bwd_g<10>(p, ctrl, ngp, stream) : cudaErrorInvalidValue;
#endif
    switch (g_template) {
        case 1: return launch_bwd_g<1>(p, ctrl, ngp, stream);
        case 2: return launch_bwd_g<2>(p, ctrl, ngp, stream);
        case 3: return launch_bwd_g<3>(p, ctrl, ngp, stream);
        case 4: return launch_bwd_g<4>(p, ctrl, ngp, stream);
        case 5: return launch_bwd_g<5>(p, ctrl, ngp, stream);
        case 6: return launch_bwd_g<6>(p, ctrl, ngp, stream);
        case 8: return launch_bwd_g<8>(p, ctrl, ngp, stream);
        case 10: return launch_bwd_g<10>(p, ctrl, ngp, stream);
        case 16: return launch_bwd_g<16>(p, ctrl, ngp, stream);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace dps

using namespace dps;

// Segment scratch of the adjoint: per step of a checkpoint segment and per stage of the scheme, the stage input (13 states,
// 6 for bare machines), sin/cos of the rotor angle and the bus voltage of every generator (dps_backward.cuh).
static size_t adjoint_scratch_bytes(const dps_plan* plan, int64_t P) {
    const int solver = plan->topo.solver;
    const size_t stages = (solver == DPS_EULER) ? 1 : (solver == DPS_HEUN ? 2 : 4);
    const size_t rows = stages * (((plan->any_ctrl ? DPS_N_STATES : 6) + 5) & ~(size_t)1);
    return sizeof(double) * (size_t)plan->topo.ckpt_stride * plan->topo.n_gen * rows * (size_t)P;
}

// 17..128 generators: same sweep, tables sized for MAX_GW, block-wide exchange (dps_backward_wide.cu)
static int backward_wide(const dps_plan* plan, int64_t P, const double* upar, const double* vpar, const double* zsub,
                         const double* ckpt, const double* grad_rec, const double* grad_x_final, const uint8_t* want,
                         double* g_x0, double* gslot, void* workspace, void* stream) {
    static thread_local BwdWideParams p;                 // ~28 KB: not on the stack of a Python thread
    memset(&p, 0, sizeof(p));
    const WideParams& w = plan->wbase;
    const int G = w.n_gen;
    p.P = P; p.n_gen = G; p.solver = w.solver; p.n_steps = w.n_steps; p.dt = w.dt; p.fn = w.fn;
    p.upar = upar; p.vpar = vpar; p.zsub = zsub; p.ckpt = ckpt; p.grad_rec = grad_rec; p.grad_xf = grad_x_final;
    p.g_x0 = g_x0; p.gslot = gslot; p.g_zsub = nullptr; p.scratch = (double*)workspace;
    p.q_counter = (int*)((char*)workspace + adjoint_scratch_bytes(plan, P));
    p.q_progress = p.q_counter + 1;
    p.q_capacity = (int)(P / 2 + 64);
    p.n_obs = w.n_obs; p.z_rows = w.z_rows; p.z_cols = w.z_cols; p.n_cfg = w.n_cfg; p.z_per_sim = 0;
    p.ckpt_stride = w.ckpt_stride; p.n_ckpt = (int)dps_n_checkpoints(plan);
    p.pmode = plan->pmode;
    p.n_seg = (int)plan->seg_step.size();
    for (int i = 0; i < p.n_seg; ++i) { p.seg_step[i] = plan->seg_step[i]; p.seg_cfg[i] = plan->seg_cfg[i]; }
    for (int g = 0; g < G; ++g) {
        p.ctrl[g] = w.ctrl[g];
        p.rec_mask[g] = w.rec_mask[g];
        memcpy(p.rec_chan[g], w.rec_chan[g], sizeof(p.rec_chan[g]));
        for (int k = 0; k < DPS_N_SLOTS; ++k) p.slot_vary[g * DPS_N_SLOTS + k] = w.slot_vary[g][k];
    }
    memcpy(p.obs_chan, w.obs_chan, sizeof(p.obs_chan));
    p.rec_uniform = 1;
    for (int g = 1; g < G; ++g) p.rec_uniform &= (p.rec_mask[g] == p.rec_mask[0]) ? 1 : 0;
    int n_want = 0;
    for (int i = 0; i < MAX_GW * DPS_N_SLOTS; ++i) p.want_row[i] = -1;
    for (int i = 0; i < G * DPS_N_SLOTS; ++i)
        if (want[i]) p.want_row[i] = (short)n_want++;
    if (n_want > 0 && !gslot) { set_error("dps_backward: slot gradients wanted but gslot is null"); return 1; }
    if (p.n_steps == 0) {
        cudaError_t e0 = grad_x_final ? cudaMemcpyAsync(g_x0, grad_x_final, sizeof(double) * (size_t)G * DPS_N_STATES * P,
                                                        cudaMemcpyDeviceToDevice, (cudaStream_t)stream)
                                      : cudaMemsetAsync(g_x0, 0, sizeof(double) * (size_t)G * DPS_N_STATES * P, (cudaStream_t)stream);
        if (e0 == cudaSuccess && n_want > 0) e0 = cudaMemsetAsync(gslot, 0, sizeof(double) * (size_t)n_want * P, (cudaStream_t)stream);
        if (e0 != cudaSuccess) { set_error("dps_backward: %s", cudaGetErrorString(e0)); return 3; }
        return 0;
    }
    int ngp = 0;
    for (int g = 0; g < G; ++g)
        for (int k = 0; k < DPS_N_SLOTS; ++k)
            if (want[g * DPS_N_SLOTS + k] && k + 1 > ngp) ngp = k + 1;
    cudaError_t e = launch_bwd_wide(p, plan->cmask, ngp, (cudaStream_t)stream);
    if (e != cudaSuccess) { set_error("dps_backward: launch failed: %s", cudaGetErrorString(e)); return 3; }
    return 0;
}

extern "C" {

size_t dps_workspace_bytes(const dps_plan* plan, int64_t P) {
    if (!plan || plan->topo.ckpt_stride <= 0 || P < 1) return 0;
    // per-step state scratch of one checkpoint segment + the work-queue words (one dequeue counter, one progress word per
    // block group; a group holds at least 2 trajectories)
    return adjoint_scratch_bytes(plan, P) + sizeof(int) * (size_t)(P / 2 + 66);
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

    if (plan->wide) {
        if (g_zsub) { set_error("dps_backward: gradients with respect to the network are not available for more than %d generators", MAX_G); return 2; }
        return backward_wide(plan, P, upar, vpar, zsub, ckpt, grad_rec, grad_x_final, want, g_x0, gslot, workspace, stream);
    }
    BwdParams p;
    memset(&p, 0, sizeof(p));
    const FwdParams& b = plan->base;
    p.P = P; p.n_gen = b.n_gen; p.solver = b.solver; p.n_steps = b.n_steps; p.dt = b.dt; p.fn = b.fn;
    p.upar = upar; p.vpar = vpar; p.zsub = zsub; p.ckpt = ckpt; p.grad_rec = grad_rec; p.grad_xf = grad_x_final;
    p.g_x0 = g_x0; p.gslot = gslot; p.g_zsub = g_zsub; p.scratch = (double*)workspace;
    p.q_counter = (int*)((char*)workspace + adjoint_scratch_bytes(plan, P));
    p.q_progress = p.q_counter + 1;
    p.q_capacity = (int)(P / 2 + 64);
    p.n_obs = b.n_obs; p.z_rows = b.z_rows; p.z_cols = b.z_cols; p.n_cfg = b.n_cfg; p.z_per_sim = b.z_per_sim;
    p.ckpt_stride = b.ckpt_stride; p.n_ckpt = (int)dps_n_checkpoints(plan);
    p.pmode = plan->pmode;
    p.n_seg = (int)plan->seg_step.size();
    for (int i = 0; i < p.n_seg; ++i) { p.seg_step[i] = plan->seg_step[i]; p.seg_cfg[i] = plan->seg_cfg[i]; }
    memcpy(p.ctrl, b.ctrl, sizeof(p.ctrl));
    memcpy(p.rec_mask, b.rec_mask, sizeof(p.rec_mask));
    memcpy(p.rec_chan, b.rec_chan, sizeof(p.rec_chan));
    memcpy(p.obs_chan, b.obs_chan, sizeof(p.obs_chan));
    p.rec_uniform = 1;
    for (int g = 1; g < b.n_gen; ++g) p.rec_uniform &= (p.rec_mask[g] == p.rec_mask[0]) ? 1 : 0;
    memcpy(p.slot_vary, b.slot_vary, sizeof(p.slot_vary));
    int n_want = 0;
    for (int i = 0; i < MAX_G * DPS_N_SLOTS; ++i) p.want_row[i] = -1;
    for (int i = 0; i < b.n_gen * DPS_N_SLOTS; ++i)
        if (want[i]) p.want_row[i] = (short)n_want++;
    if (n_want > 0 && !gslot) { set_error("dps_backward: slot gradients wanted but gslot is null"); return 1; }
    if (p.n_steps == 0) {
        // nothing to sweep: dL/dx0 = dL/dx_final, no slot gradient
        cudaError_t e0 = grad_x_final ? cudaMemcpyAsync(g_x0, grad_x_final, sizeof(double) * (size_t)b.n_gen * DPS_N_STATES * P,
                                                        cudaMemcpyDeviceToDevice, (cudaStream_t)stream)
                                      : cudaMemsetAsync(g_x0, 0, sizeof(double) * (size_t)b.n_gen * DPS_N_STATES * P, (cudaStream_t)stream);
        if (e0 == cudaSuccess && n_want > 0) e0 = cudaMemsetAsync(gslot, 0, sizeof(double) * (size_t)n_want * P, (cudaStream_t)stream);
        if (e0 != cudaSuccess) { set_error("dps_backward: %s", cudaGetErrorString(e0)); return 3; }
        return 0;
    }
    // how many leading slots need an accumulator: 0, the register slots, or all of them
    int ngp = 0;
    for (int g = 0; g < b.n_gen; ++g)
        for (int k = 0; k < DPS_N_SLOTS; ++k)
            if (want[g * DPS_N_SLOTS + k] && k + 1 > ngp) ngp = k + 1;
    cudaError_t e = launch_backward(p, plan->g_template, plan->cmask, ngp, (cudaStream_t)stream);
    if (e != cudaSuccess) { set_error("dps_backward: launch failed: %s", cudaGetErrorString(e)); return 3; }
    return 0;
}

}  // extern "C"
