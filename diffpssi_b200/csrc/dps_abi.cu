// dps_abi.cu -- extern "C" entry points declared in include/dps.h (host code; no torch types).
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <vector>
#include "dps_internal.h"
#include "dps_plan.h"

namespace dps {

static thread_local std::string g_error;

void set_error(const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_error = buf;
}

}  // namespace dps

using namespace dps;


extern "C" {

const char* dps_last_error(void) { return g_error.c_str(); }
int dps_abi_version(void) { return DPS_ABI_VERSION; }

int dps_plan_create(const dps_topology* t, dps_plan** out) {
    if (!t || !out) { set_error("dps_plan_create: null argument"); return 1; }
    if (t->abi_version != DPS_ABI_VERSION) { set_error("dps_plan_create: ABI version %d != %d", t->abi_version, DPS_ABI_VERSION); return 1; }
    if (t->n_gen < 1) { set_error("dps_plan_create: n_gen must be >= 1"); return 1; }
    if (t->n_gen > MAX_GW) {
        set_error("dps_plan_create: %d generators exceed the stepper limit (%d)", t->n_gen, MAX_GW);
        return 2;
    }
    const bool wide = t->n_gen > MAX_G;
    if (wide && t->z_per_sim) {
        set_error("dps_plan_create: systems with more than %d generators need a network shared by the batch", MAX_G);
        return 2;
    }
    if (t->solver < DPS_EULER || t->solver > DPS_RK4) { set_error("dps_plan_create: unknown solver %d", t->solver); return 1; }
    if (t->n_steps < 0 || !(t->dt > 0.0)) { set_error("dps_plan_create: bad n_steps/dt"); return 1; }
    if (t->n_obs < t->n_gen || t->n_obs - t->n_gen > MAX_EXTRA_OBS) { set_error("dps_plan_create: n_obs out of range"); return 1; }
    if (t->z_rows < t->n_obs || t->z_cols < t->n_gen) { set_error("dps_plan_create: z_rows/z_cols smaller than n_obs/n_gen"); return 1; }
    if (t->n_cfg < 1 || t->n_cfg > RED_MAX_CFG) { set_error("dps_plan_create: n_cfg %d out of range (1..%d)", t->n_cfg, RED_MAX_CFG); return 1; }
    if (!t->ctrl || !t->slot_vary || (t->n_steps > 0 && !t->step_cfg) || (t->n_rec > 0 && !t->rec)) {
        set_error("dps_plan_create: null ctrl / slot_vary / step_cfg / rec pointer"); return 1;
    }
    if (t->n_rec < 0 || t->n_rec > 32000) { set_error("dps_plan_create: n_rec out of range"); return 1; }

    dps_plan* pl = new dps_plan();
    pl->topo = *t;
    pl->ctrl.assign(t->ctrl, t->ctrl + t->n_gen);
    pl->step_cfg.assign(t->step_cfg, t->step_cfg + t->n_steps);
    pl->slot_vary.assign(t->slot_vary, t->slot_vary + (size_t)t->n_gen * DPS_N_SLOTS);
    if (t->n_rec) pl->rec.assign(t->rec, t->rec + t->n_rec);
    pl->topo.ctrl = pl->ctrl.data();
    pl->topo.step_cfg = pl->step_cfg.data();
    pl->topo.slot_vary = pl->slot_vary.data();
    pl->topo.rec = pl->rec.data();
    pl->topo.common_ctrl = nullptr;

    for (int k = 0; k < t->n_steps; ++k) {
        if (pl->step_cfg[k] >= t->n_cfg) { set_error("dps_plan_create: step_cfg[%d]=%d >= n_cfg", k, pl->step_cfg[k]); delete pl; return 1; }
        if (k == 0 || pl->step_cfg[k] != pl->step_cfg[k - 1]) { pl->seg_step.push_back(k); pl->seg_cfg.push_back(pl->step_cfg[k]); }
    }
    if (pl->seg_step.empty()) { pl->seg_step.push_back(0); pl->seg_cfg.push_back(0); }

    pl->wide = wide;
    pl->g_template = wide ? t->n_gen : fwd_template_g(t->n_gen);
    pl->any_ctrl = false;
    pl->any_vary = false;
    for (int g = 0; g < t->n_gen; ++g) pl->any_ctrl |= (pl->ctrl[g] != 0);
    for (int v : pl->slot_vary) {
        if (v >= t->n_vary || v < -1) { set_error("dps_plan_create: slot_vary entry %d outside -1..n_vary-1 (n_vary %d)", v, t->n_vary); delete pl; return 1; }
        pl->any_vary |= (v >= 0);
    }
    pl->cmask = pl->ctrl[0];
    for (int g = 1; g < t->n_gen; ++g)
        if (pl->ctrl[g] != pl->ctrl[0]) pl->cmask = -1;
    if (pl->cmask != 0 && pl->cmask != 7) pl->cmask = -1;
    // fast path: only the DPS_N_REG_SLOTS leading constants may differ between trajectories
    pl->pmode = 1;
    for (int g = 0; g < t->n_gen; ++g)
        for (int k = DPS_N_REG_SLOTS; k < DPS_N_SLOTS; ++k)
            if (pl->slot_vary[(size_t)g * DPS_N_SLOTS + k] >= 0) pl->pmode = 2;

    FwdParams& b = pl->base;
    WideParams& w = pl->wbase;
    memset(&b, 0, sizeof(b));
    memset(&w, 0, sizeof(w));
    if (t->common_ctrl != nullptr && pl->pmode == 1 && pl->cmask == 7) {
        pl->pmode = 3;
        for (int k = 0; k < DPS_N_SLOTS; ++k) b.cpar[k] = w.cpar[k] = t->common_ctrl[k];
    }
    b.n_gen = t->n_gen; b.solver = t->solver; b.n_steps = t->n_steps; b.dt = t->dt; b.fn = t->fn;
    b.n_obs = t->n_obs; b.z_rows = t->z_rows; b.z_cols = t->z_cols; b.n_cfg = t->n_cfg; b.z_per_sim = t->z_per_sim;
    b.ckpt_stride = t->ckpt_stride;
    w.n_gen = t->n_gen; w.solver = t->solver; w.n_steps = t->n_steps; w.dt = t->dt; w.fn = t->fn;
    w.n_obs = t->n_obs; w.z_rows = t->z_rows; w.z_cols = t->z_cols; w.n_cfg = t->n_cfg; w.z_per_sim = t->z_per_sim;
    w.ckpt_stride = t->ckpt_stride;
    for (int g = 0; g < t->n_gen; ++g) {
        if (wide) {
            w.ctrl[g] = (unsigned char)pl->ctrl[g];
            for (int k = 0; k < DPS_N_REG_SLOTS; ++k) w.reg_vary[g][k] = (short)pl->slot_vary[(size_t)g * DPS_N_SLOTS + k];
            for (int k = 0; k < DPS_N_SLOTS; ++k) w.slot_vary[g][k] = (short)pl->slot_vary[(size_t)g * DPS_N_SLOTS + k];
        } else {
            b.ctrl[g] = pl->ctrl[g];
            for (int k = 0; k < DPS_N_SLOTS; ++k) b.slot_vary[g * DPS_N_SLOTS + k] = (short)pl->slot_vary[(size_t)g * DPS_N_SLOTS + k];
        }
    }
    for (int e = 0; e < MAX_EXTRA_OBS; ++e) b.obs_chan[e][0] = b.obs_chan[e][1] = w.obs_chan[e][0] = w.obs_chan[e][1] = -1;
    for (int c = 0; c < t->n_rec; ++c) {
        const dps_channel& ch = pl->rec[c];
        if (ch.kind == DPS_CH_GEN) {
            if (ch.index < 0 || ch.index >= t->n_gen || ch.field < 0 || ch.field >= DPS_N_FIELDS) {
                set_error("dps_plan_create: recorder channel %d out of range", c); delete pl; return 1;
            }
            unsigned& mask = wide ? w.rec_mask[ch.index] : b.rec_mask[ch.index];
            if (mask & (1u << ch.field)) {
                set_error("dps_plan_create: recorder channel %d duplicates generator %d field %d", c, ch.index, ch.field);
                delete pl; return 1;
            }
            mask |= (1u << ch.field);
            (wide ? w.rec_chan[ch.index][ch.field] : b.rec_chan[ch.index][ch.field]) = (short)c;
        } else if (ch.kind == DPS_CH_BUS_VRE || ch.kind == DPS_CH_BUS_VIM) {
            const int e = ch.index - t->n_gen;
            if (e < 0 || ch.index >= t->n_obs) { set_error("dps_plan_create: bus channel %d must name an extra observed row", c); delete pl; return 1; }
            short& slot = wide ? w.obs_chan[e][ch.kind == DPS_CH_BUS_VIM] : b.obs_chan[e][ch.kind == DPS_CH_BUS_VIM];
            if (slot >= 0) { set_error("dps_plan_create: recorder channel %d duplicates an observed-bus channel", c); delete pl; return 1; }
            slot = (short)c;
        } else {
            set_error("dps_plan_create: recorder channel %d has unknown kind %d", c, ch.kind); delete pl; return 1;
        }
    }
    if (!wide) {
        // regular recorder layout?  (same field set for every generator, channel = base[field] + generator)
        bool regular = t->n_rec > 0;
        for (int g = 1; g < t->n_gen && regular; ++g) regular = (b.rec_mask[g] == b.rec_mask[0]);
        for (int k = 0; k < DPS_N_FIELDS && regular; ++k) {
            if (!(b.rec_mask[0] & (1u << k))) continue;
            for (int g = 0; g < t->n_gen && regular; ++g) regular = (b.rec_chan[g][k] == b.rec_chan[0][k] + g);
            b.rec_fbase[k] = b.rec_chan[0][k];
        }
        b.rec_regular = regular ? 1 : 0;
        b.rec_umask = regular ? b.rec_mask[0] : 0u;
    }
    *out = pl;
    return 0;
}

int dps_plan_destroy(dps_plan* plan) {
    delete plan;
    return 0;
}

int64_t dps_n_checkpoints(const dps_plan* plan) {
    if (!plan || plan->topo.ckpt_stride <= 0) return 0;
    return (plan->topo.n_steps + plan->topo.ckpt_stride - 1) / plan->topo.ckpt_stride;
}

int dps_plan_info(const dps_plan* plan, int32_t* block_threads, int32_t* sims_per_block, int32_t* regs_per_thread,
                  int32_t* g_template) {
    if (!plan) { set_error("dps_plan_info: null plan"); return 1; }
    int th = 0, spb = 0, regs = 0;
    if (plan->wide) {
        cudaError_t e = launch_forward_wide(plan->wbase, plan->cmask, plan->pmode, 0, &regs, &th, &spb);
        if (e != cudaSuccess) { set_error("dps_plan_info: %s", cudaGetErrorString(e)); return 3; }
    } else {
        fwd_geometry(plan->g_template, &th, &spb);
        FwdParams q = plan->base;
        q.P = plan->last_P;                    // the specialisation depends on the batch size (pair kernel)
        cudaError_t e = launch_forward(q, plan->g_template, plan->cmask, plan->pmode, 0, &regs, &spb);
        if (e != cudaSuccess) { set_error("dps_plan_info: %s", cudaGetErrorString(e)); return 3; }
    }
    if (block_threads) *block_threads = th;
    if (sims_per_block) *sims_per_block = spb;
    if (g_template) *g_template = plan->g_template;
    if (regs_per_thread) *regs_per_thread = regs;
    return 0;
}

// Steps [step_begin, step_end) from state x_in; one launch per MAX_SEG configuration change points (virtually always
// a single launch).
static int forward_range(const dps_plan* plan, int64_t P, int step_begin, int step_end, const double* x_in,
                         const double* upar, const double* vpar, const double* zsub, double* rec, double* x_out,
                         double* ckpt, void* workspace, size_t workspace_bytes, void* stream, const char* who) {
    if (!plan) { set_error("%s: null plan", who); return 1; }
    if (P < 1) { set_error("%s: P must be >= 1", who); return 1; }
    if (!x_in || !upar || !zsub) { set_error("%s: null input pointer", who); return 1; }
    if (plan->topo.n_rec > 0 && !rec) { set_error("%s: recorder channels requested but rec is null", who); return 1; }
    if (plan->any_vary && !vpar) { set_error("%s: per-sim constants requested but vpar is null", who); return 1; }
    if (plan->topo.ckpt_stride > 0 && !ckpt) { set_error("%s: checkpoints requested but ckpt is null", who); return 1; }
    if (step_begin < 0 || step_end > plan->topo.n_steps || step_begin > step_end) { set_error("%s: bad step range [%d, %d)", who, step_begin, step_end); return 1; }
    if (step_begin == step_end) return 0;
    plan->last_P = P;
    cudaStream_t st = (cudaStream_t)stream;
    const int n_seg = (int)plan->seg_step.size();
    // first segment that covers step_begin
    int seg0 = 0;
    while (seg0 + 1 < n_seg && plan->seg_step[seg0 + 1] <= step_begin) ++seg0;
    int begin = step_begin;
    while (begin < step_end) {
        int seg1 = seg0 + 1;
        while (seg1 < n_seg && seg1 - seg0 < MAX_SEG && plan->seg_step[seg1] < step_end) ++seg1;
        const int end = (seg1 < n_seg && plan->seg_step[seg1] < step_end) ? plan->seg_step[seg1] : step_end;
        if (end < step_end && !x_out) { set_error("%s: more than %d configuration changes need x_final as carry buffer", who, MAX_SEG); return 1; }
        cudaError_t e;
        if (plan->wide) {
            WideParams p = plan->wbase;
            p.P = P; p.x0 = x_in; p.upar = upar; p.vpar = vpar; p.zsub = zsub; p.rec = rec; p.ckpt = ckpt; p.x_final = x_out;
            p.step_begin = begin; p.step_end = end; p.n_seg = seg1 - seg0;
            for (int i = 0; i < p.n_seg; ++i) { p.seg_step[i] = plan->seg_step[seg0 + i]; p.seg_cfg[i] = plan->seg_cfg[seg0 + i]; }
            e = launch_forward_wide(p, plan->cmask, plan->pmode, st, nullptr, nullptr, nullptr);
        } else {
            FwdParams p = plan->base;
            p.P = P; p.x0 = x_in; p.upar = upar; p.vpar = vpar; p.zsub = zsub; p.rec = rec; p.ckpt = ckpt; p.x_final = x_out;
            p.step_begin = begin; p.step_end = end; p.n_seg = seg1 - seg0;
            for (int i = 0; i < p.n_seg; ++i) { p.seg_step[i] = plan->seg_step[seg0 + i]; p.seg_cfg[i] = plan->seg_cfg[seg0 + i]; }
            // optional work-queue scheduling (dps_internal.h): counter + one progress word per block group
            if (workspace != nullptr && workspace_bytes >= 2 * sizeof(int) && ((uintptr_t)workspace % sizeof(int)) == 0) {
                p.q_counter = (int*)workspace;
                p.q_progress = (int*)workspace + 1;
                const size_t cap = workspace_bytes / sizeof(int) - 1;
                p.q_capacity = cap > 0x3fffffff ? 0x3fffffff : (int)cap;
            }
            e = launch_forward(p, plan->g_template, plan->cmask, plan->pmode, st, nullptr);
        }
        if (e != cudaSuccess) { set_error("%s: launch failed: %s", who, cudaGetErrorString(e)); return 3; }
        x_in = x_out;
        begin = end;
        seg0 = seg1;
    }
    return 0;
}

int dps_forward(const dps_plan* plan, int64_t P, const double* x0, const double* upar, const double* vpar,
                const double* zsub, double* rec, double* x_final, double* ckpt, void* workspace, size_t workspace_bytes,
                void* stream) {
    if (!plan) { set_error("dps_forward: null plan"); return 1; }
    return forward_range(plan, P, 0, plan->topo.n_steps, x0, upar, vpar, zsub, rec, x_final, ckpt, workspace,
                         workspace_bytes, stream, "dps_forward");
}

int dps_forward_range(const dps_plan* plan, int64_t P, int32_t step_begin, int32_t step_end, const double* x_in,
                      const double* upar, const double* vpar, const double* zsub, double* rec, double* x_out,
                      double* ckpt, void* workspace, size_t workspace_bytes, void* stream) {
    return forward_range(plan, P, step_begin, step_end, x_in, upar, vpar, zsub, rec, x_out, ckpt, workspace,
                         workspace_bytes, stream, "dps_forward_range");
}

size_t dps_forward_workspace_bytes(const dps_plan* plan, int64_t P) {
    if (!plan || P < 1 || plan->wide) return 0;
    // one dequeue counter + one progress word per block group; a group holds at least 8 trajectories
    return sizeof(int) * (size_t)(P / 8 + 66);
}

int dps_network_reduce(int64_t Bz, int32_t n_bus, const double* z_re, const double* z_im, const double* yd_re,
                       const double* yd_im, int32_t n_obs, const int32_t* obs_bus, int32_t n_gen, const int32_t* gen_bus,
                       int32_t n_cfg, const int32_t* cfg_ptr, const int32_t* cfg_bus, double y_fault, double* zsub,
                       void* stream) {
    if (!z_re || !z_im || !yd_re || !yd_im || !obs_bus || !gen_bus || !cfg_ptr || !zsub) { set_error("dps_network_reduce: null argument"); return 1; }
    if (Bz < 1 || n_bus < 1) { set_error("dps_network_reduce: bad sizes"); return 1; }
    if (n_gen < 1 || n_gen > MAX_GW || n_obs < n_gen || n_obs > MAX_GW + MAX_EXTRA_OBS) { set_error("dps_network_reduce: n_gen/n_obs out of range"); return 1; }
    if (n_cfg < 1 || n_cfg > RED_MAX_CFG) { set_error("dps_network_reduce: n_cfg %d out of range (max %d)", n_cfg, RED_MAX_CFG); return 1; }
    ReduceParams p;
    memset(&p, 0, sizeof(p));
    p.Bz = Bz; p.n_bus = n_bus; p.n_obs = n_obs; p.n_gen = n_gen; p.n_cfg = n_cfg; p.y_fault = y_fault;
    p.z_re = z_re; p.z_im = z_im; p.yd_re = yd_re; p.yd_im = yd_im; p.zsub = zsub;
    for (int i = 0; i < n_obs; ++i) { if (obs_bus[i] < 0 || obs_bus[i] >= n_bus) { set_error("dps_network_reduce: obs_bus out of range"); return 1; } p.obs_bus[i] = obs_bus[i]; }
    for (int i = 0; i < n_gen; ++i) { if (gen_bus[i] < 0 || gen_bus[i] >= n_bus) { set_error("dps_network_reduce: gen_bus out of range"); return 1; } p.gen_bus[i] = gen_bus[i]; }
    int m_max = 0;
    for (int c = 0; c < n_cfg; ++c) {
        const int m = cfg_ptr[c + 1] - cfg_ptr[c];
        if (m < 0 || m > RED_MAX_FAULT) { set_error("dps_network_reduce: configuration %d has %d faulted buses (max %d)", c, m, RED_MAX_FAULT); return 1; }
        p.cfg_n[c] = m;
        for (int k = 0; k < m; ++k) {
            const int b = cfg_bus[cfg_ptr[c] + k];
            if (b < 0 || b >= n_bus) { set_error("dps_network_reduce: faulted bus out of range"); return 1; }
            p.cfg_bus[c][k] = b;
        }
        if (m > m_max) m_max = m;
    }
    p.m_max = m_max;
    cudaError_t e = launch_network_reduce_params(p, (cudaStream_t)stream);
    if (e != cudaSuccess) { set_error("dps_network_reduce: launch failed: %s", cudaGetErrorString(e)); return 3; }
    return 0;
}

int dps_recorder_to_host(double* host, const double* rec, int32_t n_rows, int64_t row_elems, int64_t first, int64_t count,
                         void* stream) {
    if (!host || !rec || n_rows < 1 || row_elems < 1 || first < 0 || count < 1 || first + count > row_elems) {
        set_error("dps_recorder_to_host: bad argument");
        return 1;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const size_t pitch = sizeof(double) * (size_t)row_elems;
    cudaError_t e = cudaSuccess;
    if (pitch <= ((size_t)1 << 31) - 1) {
        // one strided copy for the time slice of every channel (rows of rec[n_rows][T][P]) instead of one copy per channel
        e = cudaMemcpy2DAsync(host + first, pitch, rec + first, pitch, sizeof(double) * (size_t)count, (size_t)n_rows,
                              cudaMemcpyDeviceToHost, st);
    } else {
        for (int r = 0; r < n_rows && e == cudaSuccess; ++r)
            e = cudaMemcpyAsync(host + (size_t)r * row_elems + first, rec + (size_t)r * row_elems + first,
                                sizeof(double) * (size_t)count, cudaMemcpyDeviceToHost, st);
    }
    if (e != cudaSuccess) { set_error("dps_recorder_to_host: %s", cudaGetErrorString(e)); return 3; }
    return 0;
}

int dps_math_probe(int64_t n, const double* a, const double* b, double* out, void* stream) {
    if (n < 1 || !a || !b || !out) { set_error("dps_math_probe: bad argument"); return 1; }
    cudaError_t e = launch_math_probe(n, a, b, out, (cudaStream_t)stream);
    if (e != cudaSuccess) { set_error("dps_math_probe: %s", cudaGetErrorString(e)); return 3; }
    return 0;
}

int dps_fp64_peak_probe(int32_t iters, double* tflops, double* milliseconds) {
    if (!tflops || !milliseconds || iters < 1) { set_error("dps_fp64_peak_probe: bad argument"); return 1; }
    cudaError_t e = run_fp64_probe(iters, tflops, milliseconds);
    if (e != cudaSuccess) { set_error("dps_fp64_peak_probe: %s", cudaGetErrorString(e)); return 3; }
    return 0;
}

int dps_fp64_dmma_probe(int32_t iters, double* tflops, double* milliseconds) {
    if (!tflops || !milliseconds || iters < 1) { set_error("dps_fp64_dmma_probe: bad argument"); return 1; }
    cudaError_t e = run_fp64_dmma_probe(iters, tflops, milliseconds);
    if (e != cudaSuccess) { set_error("dps_fp64_dmma_probe: %s", cudaGetErrorString(e)); return 3; }
    return 0;
}

}  // extern "C"

// ---- adjoint entry points: implemented in dps_backward.cu ------------------------------------------------
