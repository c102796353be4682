// dps_forward_pair.cuh -- the stepping kernel with TWO trajectories per thread (sm_100a), a specialisation of
// dps_step_kernel (dps_forward.cuh) for its hot configuration:
//
//   all three controllers on every generator with batch-uniform constants (common to all generators: plan pmode 3,
//   constant-bank operands; or per generator: pmode 1, shared-memory table), network shared by the batch, Heun or Euler,
//   regular recorder over states (omega only: compile-time mask), no extra observed buses, no checkpoints, and
//   only the "identification" constants (1/(2H), D, the two X_d' differences, p_m0, e_fd0, v_set, bias, p_ref) differing
//   between trajectories.
//
// Why: the FP64 pipe of B200 only approaches its peak when a warp offers it independent instructions back to back
// (profiles/probe_dfma_latency_ilp.txt: one dependent chain per warp -> 66 % however many warps are resident, two -> 80 %),
// and every non-FP64 instruction costs issue bandwidth (profiles/probe_issue_mix.txt).  One stage evaluation is a long
// dependency chain; two trajectories in one thread give the scheduler a second, fully independent chain to interleave,
// and everything that does not depend on the trajectory -- the Z row (20 LDS per stage), the generator's batch-uniform
// machine constants, the controller and trigonometric constants -- is loaded once for both.
//
// Same equations (dps_model.cuh), same exchange scheme (warp-private shared-memory lines), same recorder layout and the
// same periodic block barrier as dps_step_kernel; results are bit-identical to it (tests/test_cuda_parity.py).  254
// registers, 8 warps per SM of 6 trajectories each: a "wave" is 7104 trajectories.  Measured (IEEE-39, Heun, T = 2000;
// "static" = one block per group for all steps, "queue" = (group, step chunk) work items, plan_work_queue below):
//   trajectories            8192    16384   32768   65536
//   pair kernel, static     5.49    8.97    17.61   33.21  ms     FP64 pipe 77 % busy, 218 instructions per trajectory-stage
//   pair kernel, queue      4.28    8.58    16.73   32.78  ms
//   dps_step_kernel, static 4.89    9.43    17.62   34.05  ms     FP64 pipe 73.5 %, 245 instructions
//   dps_step_kernel, queue  4.38    8.75    17.39   34.26  ms
// With the queue the pair kernel wins from the 8-GPU shard size (8192) on: that is the dispatcher's threshold.
#pragma once
#include <cuda_runtime.h>
#include <stdlib.h>
#include "dps_internal.h"
#include "dps_model.cuh"

namespace dps {

// per-trajectory register constants of the pair kernel (dps_lean_slot) vs batch-uniform shared-memory reads
__host__ __device__ constexpr bool dps_pair_reg_slot(int slot) {
    return slot == DPS_S_INV_2H || slot == DPS_S_D || slot == DPS_S_XD_XDT || slot == DPS_S_XDT_XDST ||
           slot == DPS_S_PM0 || slot == DPS_S_EFD0 || slot == DPS_S_A_VSET || slot == DPS_S_A_BIAS || slot == DPS_S_G_PREF;
}

struct ParamPair {
    const double* pm;        // registers, [DPS_N_REG_SLOTS]: only dps_pair_reg_slot entries are per trajectory
    const double* sh;        // registers, [DPS_N_REG_SLOTS]: the generator's batch-uniform machine constants (shared by both)
    const double* cp;        // kernel parameter block, [DPS_N_SLOTS]
    __device__ __forceinline__ double operator()(int slot) const {
        return (slot >= DPS_N_REG_SLOTS) ? cp[slot] : (dps_pair_reg_slot(slot) ? pm[slot] : sh[slot]);
    }
};

// Decide whether a launch of `groups` block groups runs as a work queue of (group, step chunk) items, one block each, (FwdParams::q_*, dps_internal.h) and set it up.  Worth it when the groups make only a few waves over the resident
// block slots: the last, partly filled wave of a static launch then runs latency-bound for every step of the run
// (8192 trajectories per GPU = 1.5 waves of dps_step_kernel: 4.89 ms against 4.15 ms for an eighth of 65 536).
// Needs the caller's forward workspace (q_counter / q_progress, zeroed here on the stream) and x_final (state carrier).
static cudaError_t plan_work_queue(FwdParams& q, const void* kern, int threads, size_t smem, long long groups, unsigned* grid,
                                   cudaStream_t stream) {
    q.q_items = 0;
    if (q.q_counter == nullptr || q.x_final == nullptr || groups > q.q_capacity) return cudaSuccess;
    const char* env = getenv("DPS_QUEUE_ROUNDS");              // experiments / tests: 0 = never queue
    const int rounds = env ? atoi(env) : 12;                   // measured at 8192 trajectories: 6..24 rounds within 1.5 %
    const int steps = q.step_end - q.step_begin;
    if (rounds == 0 || steps < 64 || groups > 0x3fffffff) return cudaSuccess;
    int dev = 0, sms = 0, per_sm = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem);
    if (e != cudaSuccess) return e;
    long long slots = (long long)sms * per_sm;
    if (const char* es = getenv("DPS_QUEUE_SLOTS")) slots = atoll(es);   // tests: exercise the queue at CI batch sizes
    if (slots < 1 || groups <= slots) return cudaSuccess;     // one wave: everything is resident at once anyway
    long long n_chunks = ((long long)rounds * slots + groups - 1) / groups;
    if (n_chunks < 2) return cudaSuccess;                      // many waves: the tail is already a small fraction
    int chunk_steps = (int)((steps + n_chunks - 1) / n_chunks);
    chunk_steps = (chunk_steps + 15) & ~15;                    // keeps the recorder's 16-step block barrier aligned
    if (chunk_steps < 32) chunk_steps = 32;
    n_chunks = (steps + chunk_steps - 1) / chunk_steps;
    if (n_chunks < 2 || groups * n_chunks > 0x3fffffff) return cudaSuccess;
    q.q_items = (int)(groups * n_chunks);
    q.q_groups = (int)groups;
    q.q_chunk_steps = chunk_steps;
    *grid = (unsigned)q.q_items;                              // one block per item; tickets give the order
    return cudaMemsetAsync(q.q_counter, 0, sizeof(int) * (size_t)(groups + 1), stream);   // counter + progress[groups]
}

template <int G>
struct PairCfg {
    static constexpr int SPW = 32 / G;             // trajectory PAIRS per warp
#ifndef DPS_PAIR_WARPS
#define DPS_PAIR_WARPS 4
#endif
#ifndef DPS_PAIR_MIN_BLOCKS
#define DPS_PAIR_MIN_BLOCKS 2
#endif
    static constexpr int WARPS = DPS_PAIR_WARPS;
    static constexpr int THREADS = 32 * WARPS;
    static constexpr int MIN_BLOCKS = DPS_PAIR_MIN_BLOCKS;
};

// PMODE 3: controller constants common to all generators, constant-bank operands (the headline configuration).
// PMODE 1: per-generator (batch-uniform) controller constants from a shared-memory table -- both trajectories of a thread
//          share every load, which is what keeps this "generic" path within reach of the specialised one: in
//          dps_step_kernel the 24 extra LDS per stage saturate the shared-memory pipe (47.6 ms against 34.1 ms).
// RMASK >= 0: compile-time recorder mask; -1: run-time mask over the 13 states of the regular recorder layout.
template <int G, int SOLVER, int RMASK, bool QUEUED = false, int PMODE = 3>
__global__ void __launch_bounds__(PairCfg<G>::THREADS, PairCfg<G>::MIN_BLOCKS)
dps_step_pair_kernel(const FwdParams p) {
    using C = PairCfg<G>;
    constexpr int SPW = C::SPW, NS = DPS_N_STATES, CM = 7;
    constexpr int UPS = DPS_N_SLOTS + 1;           // odd row stride: the generator rows of a warp fall into distinct banks
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* s_zr = reinterpret_cast<double*>(smem_raw);                               // [n_cfg][col j][row g]
    double* s_zi = s_zr + p.n_cfg * G * G;
    double2* s_cur = reinterpret_cast<double2*>(s_zi + p.n_cfg * G * G);              // [WARPS][2 buffers][2 trajectories][32]
    double* s_upar = reinterpret_cast<double*>(s_cur + C::WARPS * 128);               // PMODE 1: [G][UPS]

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int sl = lane / G;
    const int gl = lane - sl * G;
    const int base = ((sl < SPW) ? sl : SPW - 1) * G;
    const bool gen_live = (sl < SPW) && (gl < p.n_gen);
    const int g = (gl < p.n_gen) ? gl : p.n_gen - 1;
    const long long P = p.P;
    __shared__ int s_item;                         // work queue: the item this block is processing

    for (int i = threadIdx.x; i < C::WARPS * 128; i += blockDim.x) s_cur[i] = make_double2(0.0, 0.0);
    if (PMODE == 1) {
        for (int i = threadIdx.x; i < G * DPS_N_SLOTS; i += blockDim.x) {
            const int gi = i / DPS_N_SLOTS, k = i - gi * DPS_N_SLOTS;
            s_upar[gi * UPS + k] = (gi < p.n_gen) ? p.upar[i] : 0.0;
        }
    }
    for (int i = threadIdx.x; i < p.n_cfg * G * G; i += blockDim.x) {
        const int row = i % G, col = (i / G) % G, cfg = i / (G * G);
        const bool ok = row < p.n_gen && col < p.n_gen;
        const double* e = p.zsub + (((long long)cfg * p.z_rows + (ok ? row : 0)) * p.z_cols + (ok ? col : 0)) * 2;
        s_zr[i] = ok ? e[0] : 0.0;
        s_zi[i] = ok ? e[1] : 0.0;
    }
    __syncthreads();

    const double dt = p.dt, hdt = 0.5 * p.dt;
    constexpr bool queued = QUEUED;

    // one work item per block (see dps_step_kernel): block group `group` over steps [step_begin, step_end)
    long long group = blockIdx.x;
    int chunk = 0, n_chunks = 1;
    if (queued) {
        if (threadIdx.x == 0) s_item = atomicAdd(p.q_counter, 1);
        __syncthreads();
        const int item = __reduce_max_sync(0xffffffffu, s_item);     // REDUX: provably warp-uniform ticket
        chunk = item / p.q_groups;
        group = item - chunk * p.q_groups;
        n_chunks = p.q_items / p.q_groups;
        if (chunk > 0 && threadIdx.x == 0)
            while (dps_ld_acquire(p.q_progress + group) < chunk) __nanosleep(100);
        __syncthreads();
    }
    const int step_begin = queued ? p.step_begin + chunk * p.q_chunk_steps : p.step_begin;
    const int step_end = (queued && chunk + 1 < n_chunks) ? step_begin + p.q_chunk_steps : p.step_end;
    const long long sim0 = (((long long)(queued ? group : (long long)blockIdx.x) * C::WARPS + warp) * SPW + sl) * 2;   // this thread: sim0, sim0 + 1
    bool live[2];
    long long s[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        live[u] = gen_live && (sim0 + u < P);
        s[u] = (sim0 + u < P) ? sim0 + u : (P - 1);
    }

    // ---- constants ----------------------------------------------------------------------------------------------------
    double sh[DPS_N_REG_SLOTS], pm[2][DPS_N_REG_SLOTS];
#pragma unroll
    for (int k = 0; k < DPS_N_REG_SLOTS; ++k) {
        const int vi = p.slot_vary[g * DPS_N_SLOTS + k];
        const double uni = __ldg(p.upar + g * DPS_N_SLOTS + k);
        sh[k] = uni;                                                   // read only where !dps_pair_reg_slot(k)
#pragma unroll
        for (int u = 0; u < 2; ++u) pm[u][k] = (vi >= 0) ? __ldg(p.vpar + (long long)vi * P + s[u]) : uni;
    }
    const double* ctrl_consts = (PMODE == 1) ? (s_upar + g * UPS) : p.cpar;
    const ParamPair par[2] = {{pm[0], sh, ctrl_consts}, {pm[1], sh, ctrl_consts}};

    double x[2][NS];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const double* src = ((!queued || chunk == 0) ? p.x0 : p.x_final) + (long long)g * DPS_N_STATES * P + s[u];
#pragma unroll
        for (int k = 0; k < NS; ++k) x[u][k] = (!queued || chunk == 0) ? src[(long long)k * P] : __ldcg(src + (long long)k * P);
    }

    int seg = 0;
    while (seg + 1 < p.n_seg && p.seg_step[seg + 1] <= step_begin) ++seg;
    int cfg = p.seg_cfg[seg];
    int next_change = (seg + 1 < p.n_seg) ? p.seg_step[seg + 1] : 0x7fffffff;
    const double* zcol_r = s_zr + cfg * G * G + gl;
    const double* zcol_i = s_zi + cfg * G * G + gl;

    // recorder: one running pointer per recorded field; the thread's two trajectories are adjacent in every row
    constexpr int N_RUN = (RMASK >= 0) ? __builtin_popcount((unsigned)RMASK & 0x1FFFu) : 0;
    const long long rec_cs = (long long)p.n_steps * P;
    double* rec_run[N_RUN > 0 ? N_RUN : 1];
    double* rec_row = p.rec + (long long)g * rec_cs + (long long)step_begin * P + sim0;   // RMASK < 0: channel g, this step
    if (RMASK >= 0) {
        int q = 0;
#pragma unroll
        for (int k = 0; k < DPS_N_STATES; ++k)
            if ((RMASK >> k) & 1)
                rec_run[q++] = p.rec + ((long long)g + p.rec_fbase[k]) * rec_cs + (long long)step_begin * P + sim0;
    }
    int buf = 0;
    double2* cur_w = s_cur + warp * 128;

    // one stage evaluation of both trajectories
    auto eval = [&](double (*X)[NS], double (*f)[NS]) {
        double sd[2], cd[2];
        double2* cw = cur_w + buf * 64;
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            dps_sincos(X[u][DPS_X_DELTA], &sd[u], &cd[u]);
            double i_re, i_im;
            injection(X[u], par[u], sd[u], cd[u], i_re, i_im);
            if (gen_live) cw[u * 32 + lane] = make_double2(i_re, i_im);
        }
        __syncwarp();
        buf ^= 1;
        const double2* cg0 = cw + base;
        const double2* cg1 = cw + 32 + base;
        // V = Z I for both trajectories: every Z element is loaded once
        double vr[2][2] = {{0.0, 0.0}, {0.0, 0.0}}, vi[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
        for (int j = 0; j < G; ++j) {
            const double zre = zcol_r[j * G], zim = zcol_i[j * G];
            const double2 a = cg0[j], b = cg1[j];
            vr[0][0] = fma(zre, a.x, vr[0][0]);
            vr[0][1] = fma(-zim, a.y, vr[0][1]);
            vi[0][0] = fma(zre, a.y, vi[0][0]);
            vi[0][1] = fma(zim, a.x, vi[0][1]);
            vr[1][0] = fma(zre, b.x, vr[1][0]);
            vr[1][1] = fma(-zim, b.y, vr[1][1]);
            vi[1][0] = fma(zre, b.y, vi[1][0]);
            vi[1][1] = fma(zim, b.x, vi[1][1]);
        }
        // limiter vote over both trajectories of all lanes (see stage_derivatives in dps_model.cuh)
        const bool at = stage_at_limit(X[0], par[0], CM) || stage_at_limit(X[1], par[1], CM);
        Algebraics alg;
        if (__any_sync(0xffffffffu, at)) {
#pragma unroll
            for (int u = 0; u < 2; ++u)
                stage_body<true, true>(X[u], par[u], CM, p.fn, sd[u], cd[u], vr[u][0] + vr[u][1], vi[u][0] + vi[u][1], f[u], alg);
        } else {
#pragma unroll
            for (int u = 0; u < 2; ++u)
                stage_body<true, false>(X[u], par[u], CM, p.fn, sd[u], cd[u], vr[u][0] + vr[u][1], vi[u][0] + vi[u][1], f[u], alg);
        }
    };

    for (int step = step_begin; step < step_end; ++step) {
#if DPS_REC_SYNC_STEPS > 0
        if ((step & (DPS_REC_SYNC_STEPS - 1)) == 0) __syncthreads();      // see dps_step_kernel
#endif
        if (step == next_change) {
            ++seg;
            cfg = p.seg_cfg[seg];
            next_change = (seg + 1 < p.n_seg) ? p.seg_step[seg + 1] : 0x7fffffff;
            zcol_r = s_zr + cfg * G * G + gl;
            zcol_i = s_zi + cfg * G * G + gl;
        }
        double f[2][NS];
        if (SOLVER == DPS_EULER) {
            eval(x, f);
#pragma unroll
            for (int u = 0; u < 2; ++u)
#pragma unroll
                for (int k = 0; k < NS; ++k) x[u][k] = x[u][k] + f[u][k] * dt;
        } else {
            double xt[2][NS];
            eval(x, f);
#pragma unroll
            for (int u = 0; u < 2; ++u)
#pragma unroll
                for (int k = 0; k < NS; ++k) { xt[u][k] = fma(f[u][k], dt, x[u][k]); x[u][k] = fma(f[u][k], hdt, x[u][k]); }
            eval(xt, f);
#pragma unroll
            for (int u = 0; u < 2; ++u)
#pragma unroll
                for (int k = 0; k < NS; ++k) x[u][k] = fma(f[u][k], hdt, x[u][k]);
        }
        if (RMASK >= 0) {
            int q = 0;
#pragma unroll
            for (int k = 0; k < DPS_N_STATES; ++k)
                if ((RMASK >> k) & 1) {
                    if (live[0]) dps_rec_store(rec_run[q], x[0][k]);
                    if (live[1]) dps_rec_store(rec_run[q] + 1, x[1][k]);
                    rec_run[q] += P;
                    ++q;
                }
        } else {
            const unsigned um = p.rec_umask;           // launch-uniform: the tests below do not diverge
#pragma unroll
            for (int k = 0; k < DPS_N_STATES; ++k)
                if (um & (1u << k)) {
                    double* o = rec_row + (long long)p.rec_fbase[k] * rec_cs;
                    if (live[0]) dps_rec_store(o, x[0][k]);
                    if (live[1]) dps_rec_store(o + 1, x[1][k]);
                }
            rec_row += P;
        }
    }

    if (p.x_final != nullptr) {
#pragma unroll
        for (int u = 0; u < 2; ++u)
            if (live[u]) {
                double* dst = p.x_final + (long long)g * DPS_N_STATES * P + s[u];
#pragma unroll
                for (int k = 0; k < NS; ++k) dst[(long long)k * P] = x[u][k];
            }
    }
    if (queued) {
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) {
            // the ticket is re-read instead of kept: nothing of the queue stays live in registers across the step loop
            const int item = s_item, c = item / p.q_groups;
            dps_st_release(p.q_progress + (item - c * p.q_groups), c + 1);
        }
    }
}

// Can this launch use the pair kernel?  (see the header comment)
template <int G>
static bool pair_kernel_applies(const FwdParams& p, int cmask, int pmode, long long min_sims) {
    if (cmask != 7 || (pmode != 3 && pmode != 1) || p.z_per_sim || p.ckpt_stride > 0 || p.n_obs != p.n_gen) return false;
    if (p.solver != DPS_HEUN && p.solver != DPS_EULER) return false;
    // regular recorder over states only (algebraics are recorded by dps_step_kernel)
    if (!p.rec_regular || p.rec_umask == 0u || (p.rec_umask >> DPS_N_STATES) != 0u || p.P < min_sims) return false;
    for (int g = 0; g < p.n_gen; ++g)
        for (int k = 0; k < DPS_N_REG_SLOTS; ++k)
            if (!dps_pair_reg_slot(k) && p.slot_vary[g * DPS_N_SLOTS + k] >= 0) return false;
    return true;
}

template <int G, int SOLVER, int RMASK, int PMODE>
static cudaError_t launch_pair_solver(const FwdParams& p, cudaStream_t stream, int* regs) {
    using C = PairCfg<G>;
    auto kern = dps_step_pair_kernel<G, SOLVER, RMASK, false, PMODE>;
    auto kern_q = dps_step_pair_kernel<G, SOLVER, RMASK, true, PMODE>;
    if (regs) {
        cudaFuncAttributes a;
        cudaError_t e = cudaFuncGetAttributes(&a, kern);
        if (e != cudaSuccess) return e;
        *regs = a.numRegs;
        return cudaSuccess;
    }
    const size_t smem = sizeof(double) * p.n_cfg * G * G * 2 + sizeof(double2) * C::WARPS * 128 +
                        (PMODE == 1 ? sizeof(double) * G * (DPS_N_SLOTS + 1) : 0);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    const long long sims_per_block = (long long)C::WARPS * C::SPW * 2;
    const long long blocks = (p.P + sims_per_block - 1) / sims_per_block;
    FwdParams q = p;
    unsigned grid = (unsigned)blocks;
    cudaError_t e = plan_work_queue(q, (const void*)kern_q, C::THREADS, smem, blocks, &grid, stream);
    if (e != cudaSuccess) return e;
    if (q.q_items > 0) {
        if (smem > 48 * 1024) {
            e = cudaFuncSetAttribute(kern_q, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
        }
        kern_q<<<grid, C::THREADS, smem, stream>>>(q);
    } else {
        kern<<<(unsigned)blocks, C::THREADS, smem, stream>>>(p);
    }
    return cudaGetLastError();
}

// Hook called by launch_forward_g (dps_forward.cuh): true when the pair kernel took the launch.
template <int G>
bool try_pair_kernel(const FwdParams& p, int cmask, int pmode, cudaStream_t stream, int* regs, int* spb, cudaError_t* err) {
    const char* env = getenv("DPS_PAIR_MIN_SIMS");                  // experiments / tests: 0 = always, huge = never
    const long long min_sims = env ? atoll(env) : 8192;
    if (!pair_kernel_applies<G>(p, cmask, pmode, min_sims)) return false;
    if (spb) *spb = PairCfg<G>::WARPS * PairCfg<G>::SPW * 2;
    // four builds: controller constants common to all generators in the constant bank (pmode 3) or per generator from shared
    // memory (pmode 1)  x  the omega recorder folded in at compile time or the run-time recorder mask.  (Until round 2b only the
    // two corners existed and a grid that differed from the headline in ONE of the two paid for both.)
    const bool rec_w = (p.rec_umask == 0x1u);
    const bool heun = (p.solver == DPS_HEUN);
    if (pmode == 3 && rec_w)
        *err = heun ? launch_pair_solver<G, DPS_HEUN, 0x1, 3>(p, stream, regs) : launch_pair_solver<G, DPS_EULER, 0x1, 3>(p, stream, regs);
    else if (pmode == 3)
        *err = heun ? launch_pair_solver<G, DPS_HEUN, -1, 3>(p, stream, regs) : launch_pair_solver<G, DPS_EULER, -1, 3>(p, stream, regs);
    else if (rec_w)
        *err = heun ? launch_pair_solver<G, DPS_HEUN, 0x1, 1>(p, stream, regs) : launch_pair_solver<G, DPS_EULER, 0x1, 1>(p, stream, regs);
    else
        *err = heun ? launch_pair_solver<G, DPS_HEUN, -1, 1>(p, stream, regs) : launch_pair_solver<G, DPS_EULER, -1, 1>(p, stream, regs);
    return true;
}

}  // namespace dps
