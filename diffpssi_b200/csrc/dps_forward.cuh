// dps_forward.cuh -- the persistent fused stepping kernel (sm_100a), instantiated per generator count by
// dps_forward_inst.cu (one translation unit per G so the build parallelises).
//
// Replaces the whole ``for t in self.time`` loop of PowerSystemSimulation.run() (reference
// src/diffpssi/power_sim_lib/simulator.py:523-543) for all parallel_sims trajectories with ONE launch:
// fault switching, every solver stage (solvers.py:26-46 / 75-117 / 143-217), the network solve V = Z I
// (solvers.py:83,102), the model equations (dps_model.cuh) and the recorder (simulator.py:540-546).
//
// Mapping:
//   * one thread per (trajectory, generator).  The G generators of one trajectory sit in G adjacent lanes of ONE warp
//     (a warp holds SPW = 32/G trajectories), so the Norton currents the network solve needs are exchanged with warp
//     shuffles: no shared-memory round trip, no barrier, and warps never wait for each other -- the SM's schedulers
//     hide one warp's latency with the others (the earlier block-wide exchange ran 10 warps in lockstep and left the
//     FP64 pipe idle during every exchange; profiles/ncu_r1d_*);
//   * states x, the stage input and the stage accumulator stay in registers for all T steps; per step only the
//     recorder lines go to HBM.
// Specialisations:
//   PMODE 1 (fast): the DPS_N_REG_SLOTS leading constants (machine + init-derived set points) of the thread's
//           (trajectory, generator) are loaded ONCE into registers (per-sim or uniform alike); the remaining controller
//           constants must be batch-uniform and are read from a bank-conflict-free shared-memory table.
//   PMODE 3 (fastest): PMODE 1 plus: every generator has all three controllers with constants that are identical between
//           generators; they sit in the kernel parameter block and fold into the instructions as constant-bank operands.
//   PMODE 2 (generic): any constant may be per-sim; every read goes through the slot table (slow, kept for generality).
//   ZPS 0: the reduced impedance matrices of all fault configurations are shared by the batch and sit in shared
//          memory, transposed so the lanes of a warp read consecutive addresses;  ZPS 1: one matrix per trajectory,
//          row held in registers.
//   SOLVER: the stage sequence is unrolled per scheme, so no register copies are needed between steps.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "dps_internal.h"
#include "dps_model.cuh"

#ifndef DPS_Z_PAIRS
#define DPS_Z_PAIRS 0
#endif
#ifndef DPS_REC_SYNC_STEPS
#define DPS_REC_SYNC_STEPS 16
#endif
#include "dps_forward_pair.cuh"

namespace dps {

template <int G>
struct FwdCfg {
    static constexpr int SPW = 32 / G;             // trajectories per warp
    static constexpr int WARPS = DPS_FWD_WARPS;    // warps are independent; the block size only sets the launch geometry
    static constexpr int THREADS = 32 * WARPS;
#ifndef DPS_MIN_BLOCKS
#define DPS_MIN_BLOCKS (12 / DPS_FWD_WARPS)
#endif
    // 3 blocks -> 12 warps/SM at <= 168 registers: no spills.  (4 blocks / 128 registers with the Z row in registers was
    // 4 % faster but spilled ~200 B per thread, which doubled the kernel's DRAM traffic -- profiles/ncu_r1l_*.)
    static constexpr int MIN_BLOCKS = DPS_MIN_BLOCKS;
    // a batch-shared Z lives in shared memory (transposed); only a per-trajectory Z row is held in registers
    static constexpr bool Z_SMEM = true;
    static constexpr int UPS = DPS_N_SLOTS + 1;    // row stride of the constant table: odd -> conflict-free per-lane rows
};

// RMASK: compile-time copy of the regular recorder's field mask (dps_internal.h), or -1 for "decide at run time".
// The recorder of every benchmark configuration -- omega of each generator, mask 0x1 -- then is ONE predicated
// streaming store per step instead of ~45 instructions of mask tests and channel arithmetic.
// QUEUED: compile-time switch of the work-queue prologue / epilogue (dps_internal.h).  A template parameter rather than a
// run-time flag because the static launch then keeps its step bounds in the constant bank and compiles to exactly the
// pre-queue code (the run-time flag cost it 3 % through a different register allocation).
template <int G, int CM, int PMODE, int ZPS, int SOLVER, int RMASK = -1, bool QUEUED = false>
__global__ void __launch_bounds__(FwdCfg<G>::THREADS, FwdCfg<G>::MIN_BLOCKS)
dps_step_kernel(const FwdParams p) {
    using C = FwdCfg<G>;
    constexpr int SPW = C::SPW, UPS = C::UPS;
    constexpr bool CTRL = (CM != 0);                 // CM: -1 = per-generator controller flags, else the common flag set
    constexpr int NS = CTRL ? DPS_N_STATES : 6;      // live states per generator
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // ZPS==0: Z of every configuration in shared memory, real and imaginary parts in separate planes.
    //  default: transposed [n_cfg][col j][row g], one LDS.64 per element: the G lanes of a trajectory read G consecutive
    //    doubles and the trajectory groups of a warp hit the same addresses (2 conflict-free wavefronts per load).
    //  DPS_Z_PAIRS (experiment, off): row-major [n_cfg][row g][RS], RS = G rounded up to even; a lane reads TWO consecutive
    //    elements of its row per LDS.128.  10 loads of Z per stage instead of 20 and 3.6 % fewer instructions, but the
    //    quarter-warps that straddle two trajectory groups conflict and the loads arrive later: 37.0 ms vs 35.3 ms.
    //  (The interleaved (re, im) LDS.128 layout of r1m needed 7 wavefronts per load.)
    constexpr int RS = DPS_Z_PAIRS ? ((G + 1) & ~1) : G;
    double* s_zr = reinterpret_cast<double*>(smem_raw);
    double* s_zi = s_zr + (ZPS == 0 ? p.n_cfg * G * RS : 0);
    double2* s_cur = reinterpret_cast<double2*>(s_zi + (ZPS == 0 ? p.n_cfg * G * RS : 0));   // [WARPS][2][32] exchange lines
    double* s_upar = reinterpret_cast<double*>(s_cur + C::WARPS * 64);           // [G][UPS]
    short* s_vary = reinterpret_cast<short*>(s_upar + G * UPS);                  // [G][DPS_N_SLOTS]

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int sl = lane / G;                       // trajectory slot inside the warp
    const int gl = lane - sl * G;                  // generator slot
    const int base = ((sl < SPW) ? sl : SPW - 1) * G;   // first lane of this trajectory's generator group
    const bool gen_live = (sl < SPW) && (gl < p.n_gen);     // template G may exceed the real generator count
    const int g = (gl < p.n_gen) ? gl : p.n_gen - 1;
    const long long P = p.P;
    __shared__ int s_item;                         // work queue: the item this block is processing

    for (int i = threadIdx.x; i < G * DPS_N_SLOTS; i += blockDim.x) {
        const int gi = i / DPS_N_SLOTS, k = i - gi * DPS_N_SLOTS;
        const bool ok = gi < p.n_gen;
        s_upar[gi * UPS + k] = ok ? p.upar[i] : 0.0;
        if (PMODE == 2) s_vary[i] = ok ? p.slot_vary[i] : (short)-1;
    }
    // exchange lines start at zero and only live generator lanes ever store, so the slots of padding lanes
    // (template G > n_gen) contribute exactly 0 to V = Z I without a per-stage select
    for (int i = threadIdx.x; i < C::WARPS * 64; i += blockDim.x) s_cur[i] = make_double2(0.0, 0.0);
    if (ZPS == 0) {
        for (int i = threadIdx.x; i < p.n_cfg * G * RS; i += blockDim.x) {
            const int row = DPS_Z_PAIRS ? (i / RS) % G : i % G, col = DPS_Z_PAIRS ? i % RS : (i / G) % G, cfg = i / (G * RS);
            const bool ok = row < p.n_gen && col < p.n_gen;
            const double* e = p.zsub + (((long long)cfg * p.z_rows + (ok ? row : 0)) * p.z_cols + (ok ? col : 0)) * 2;
            s_zr[i] = ok ? e[0] : 0.0;
            s_zi[i] = ok ? e[1] : 0.0;
        }
    }
    __syncthreads();

    const double* up = s_upar + g * UPS;
    const int ctrl = (CM >= 0) ? CM : p.ctrl[g];
    const double dt = p.dt, hdt = 0.5 * p.dt;
    const int n_extra_iter = (p.n_obs > p.n_gen) ? (p.n_obs - p.n_gen + p.n_gen - 1) / p.n_gen : 0;
    constexpr bool queued = QUEUED;

    // A block integrates ONE work item: block group `group` (WARPS * SPW trajectories) over steps [step_begin, step_end).
    // Plain launch: group = blockIdx.x, all steps.  Queued launch (dps_internal.h): one block per (group, step chunk) item;
    // the block takes a ticket from an atomic counter (tickets, not blockIdx, define the order: whoever holds ticket i
    // runs, and every ticket below i was taken by a block that runs or has finished, so waiting for an earlier item cannot
    // deadlock whatever order the hardware dispatches blocks in) and the state of a group travels between its chunks
    // through x_final.
    long long group = blockIdx.x;
    int chunk = 0, n_chunks = 1;
    if (queued) {
        if (threadIdx.x == 0) s_item = atomicAdd(p.q_counter, 1);
        __syncthreads();
        // REDUX: tells the compiler that the ticket is warp-uniform, which keeps the step loop's bounds and the trigonometric
        // / controller constants on the uniform datapath (a plain shared-memory load made the whole loop "possibly
        // divergent": +10 % instructions, 20 % slower)
        const int item = __reduce_max_sync(0xffffffffu, s_item);
        chunk = item / p.q_groups;
        group = item - chunk * p.q_groups;
        n_chunks = p.q_items / p.q_groups;
        // chunk c of a group starts from the state chunk c-1 left behind; that item holds an earlier ticket
        if (chunk > 0 && threadIdx.x == 0)
            while (dps_ld_acquire(p.q_progress + group) < chunk) __nanosleep(100);
        __syncthreads();
    }
    const int step_begin = queued ? p.step_begin + chunk * p.q_chunk_steps : p.step_begin;
    const int step_end = (queued && chunk + 1 < n_chunks) ? step_begin + p.q_chunk_steps : p.step_end;
    const long long sim = ((long long)(queued ? group : (long long)blockIdx.x) * C::WARPS + warp) * SPW + sl;
    const bool live = gen_live && (sim < P);
    const long long s = (sim < P) ? sim : (P - 1);

    // ---- parameter access -------------------------------------------------------------------------------------------
    double pm[DPS_N_REG_SLOTS];
    if (PMODE == 1 || PMODE == 3) {
#pragma unroll
        for (int k = 0; k < DPS_N_REG_SLOTS; ++k) {
            const int vi = p.slot_vary[g * DPS_N_SLOTS + k];
            pm[k] = (vi >= 0) ? __ldg(p.vpar + (long long)vi * P + s) : up[k];
        }
    }
    ParamRegs par_fast{pm, up};
    ParamConst par_const{pm, p.cpar};
    ParamReader<true> par_gen;
    par_gen.up = up;
    par_gen.sv = s_vary + g * DPS_N_SLOTS;
    par_gen.vpar = p.vpar;
    par_gen.P = P;
    par_gen.s = s;

    double x[NS];
    {
        // state at step_begin: the caller's x0, or what the previous chunk of this group left in x_final (written by
        // another SM: read through L2, never from a stale L1 line)
        const double* src = ((!queued || chunk == 0) ? p.x0 : p.x_final) + (long long)g * DPS_N_STATES * P + s;
#pragma unroll
        for (int k = 0; k < NS; ++k) x[k] = (!queued || chunk == 0) ? src[(long long)k * P] : __ldcg(src + (long long)k * P);
    }

    // ---- network rows -------------------------------------------------------------------------------------------------
    double zr[ZPS ? G : 1], zi[ZPS ? G : 1];
    const long long zb = p.z_per_sim ? P : 1;
    const long long zs = p.z_per_sim ? s : 0;
    const double *zcol_r = nullptr, *zcol_i = nullptr;   // ZPS==0: &Zt[cfg][0][gl]; element j at zcol[j*G]
    auto load_z_row = [&](int cfg) {
        if (ZPS) {
            const double* zbase = p.zsub + (((long long)cfg * p.z_rows + g) * p.z_cols) * 2 * zb + zs;
#pragma unroll
            for (int j = 0; j < G; ++j) {
                const bool ok = j < p.n_gen;
                zr[ZPS ? j : 0] = ok ? __ldg(zbase + (long long)(2 * j) * zb) : 0.0;
                zi[ZPS ? j : 0] = ok ? __ldg(zbase + (long long)(2 * j + 1) * zb) : 0.0;
            }
        } else {
            zcol_r = s_zr + cfg * G * RS + (DPS_Z_PAIRS ? gl * RS : gl);
            zcol_i = s_zi + cfg * G * RS + (DPS_Z_PAIRS ? gl * RS : gl);
        }
    };

    // configuration schedule: (first step, cfg) change points, ascending
    int seg = 0;
    while (seg + 1 < p.n_seg && p.seg_step[seg + 1] <= step_begin) ++seg;
    int cfg = p.seg_cfg[seg];
    int next_change = (seg + 1 < p.n_seg) ? p.seg_step[seg + 1] : 0x7fffffff;
    load_z_row(cfg);

    const unsigned rec_mask = live ? p.rec_mask[g] : 0u;
    // regular recorder layout (dps_internal.h): launch-uniform field mask, channel = rec_fbase[field] + generator
    const bool rec_regular = (RMASK >= 0) ? true : (p.rec_regular != 0);
    const unsigned rec_umask = (RMASK >= 0) ? (unsigned)RMASK : p.rec_umask;
    const long long rec_cs = (long long)p.n_steps * P;                 // channel stride
    double* const rec_thread = p.rec + (long long)g * rec_cs + sim;    // channel g, step 0, this trajectory
    constexpr int N_RUN = (RMASK >= 0) ? __builtin_popcount((unsigned)RMASK & 0x1FFFu) : 0;
    double* rec_run[N_RUN > 0 ? N_RUN : 1];                            // RMASK: &rec[channel of field][step][sim]
    if (RMASK >= 0) {
        int q = 0;
#pragma unroll
        for (int k = 0; k < DPS_N_STATES; ++k)
            if ((RMASK >> k) & 1) rec_run[q++] = rec_thread + (long long)p.rec_fbase[k] * rec_cs + (long long)step_begin * P;
    }
    int step = step_begin;
    int buf = 0;
    // dummy lanes of a partially filled warp read the last real group
    double2* cur_w = s_cur + warp * 64;

    // One stage evaluation at state X (reference: one `voltages = matmul(...); for bus: update_voltages; differential()`
    // pass of solvers.py): Norton currents -> warp shuffles -> V = Z I -> derivatives f.
    auto eval = [&](const double* X, double* f, bool last_stage) {
        Algebraics alg;
        double sd, cd;
        dps_sincos(X[DPS_X_DELTA], &sd, &cd);
        double i_re, i_im;
        if (PMODE == 3) injection(X, par_const, sd, cd, i_re, i_im);
        else if (PMODE == 1) injection(X, par_fast, sd, cd, i_re, i_im);
        else injection(X, par_gen, sd, cd, i_re, i_im);
        // exchange inside the warp through its private, double-buffered shared-memory line: one STS.128, then every
        // lane reads the G currents of its own trajectory as broadcast LDS.128 (a quarter of the wavefronts the
        // equivalent 4*G 32-bit shuffles cost); __syncwarp orders the store against the loads.
        double2* cw = cur_w + buf * 32;
        if (gen_live) cw[lane] = make_double2(i_re, i_im);
        __syncwarp();
        buf ^= 1;
        const double2* cgrp = cw + base;
        // V_g = sum_j Z[g][j] I_j   (solvers.py:83; zero injections at non-generator buses are skipped).
        // Two partial sums per component keep four independent DFMA chains in flight.
        double vra = 0.0, vrb = 0.0, via = 0.0, vib = 0.0;
        if (ZPS == 0 && DPS_Z_PAIRS) {
#pragma unroll
            for (int jp = 0; jp < RS / 2; ++jp) {
                const double2 zr2 = reinterpret_cast<const double2*>(zcol_r)[jp];
                const double2 zi2 = reinterpret_cast<const double2*>(zcol_i)[jp];
                const double2 ca = cgrp[2 * jp];
                vra = fma(zr2.x, ca.x, vra);
                vrb = fma(-zi2.x, ca.y, vrb);
                via = fma(zr2.x, ca.y, via);
                vib = fma(zi2.x, ca.x, vib);
                if (2 * jp + 1 < G) {
                    const double2 cb = cgrp[2 * jp + 1];
                    vra = fma(zr2.y, cb.x, vra);
                    vrb = fma(-zi2.y, cb.y, vrb);
                    via = fma(zr2.y, cb.y, via);
                    vib = fma(zi2.y, cb.x, vib);
                }
            }
        } else {
#pragma unroll
            for (int j = 0; j < G; ++j) {
                const double2 cj = cgrp[j];
                const double jr = cj.x, ji = cj.y;
                double zre, zim;
                if (ZPS) { zre = zr[ZPS ? j : 0]; zim = zi[ZPS ? j : 0]; }
                else { zre = zcol_r[j * G]; zim = zcol_i[j * G]; }
                vra = fma(zre, jr, vra);
                vrb = fma(-zim, ji, vrb);
                via = fma(zre, ji, via);
                vib = fma(zim, jr, vib);
            }
        }
        const double v_re = vra + vrb, v_im = via + vib;
        if (PMODE == 3) stage_derivatives<CTRL, true>(X, par_const, ctrl, p.fn, sd, cd, v_re, v_im, f, alg);
        else if (PMODE == 1) stage_derivatives<CTRL, true>(X, par_fast, ctrl, p.fn, sd, cd, v_re, v_im, f, alg);
        else stage_derivatives<CTRL, true>(X, par_gen, ctrl, p.fn, sd, cd, v_re, v_im, f, alg);

        // recorded algebraics hold the values of the LAST stage evaluation of the step (simulator.py:540-543)
        if (last_stage && rec_regular) {
            if (rec_umask >> DPS_N_STATES) {       // uniform test; stores predicated on `live` only
                double* out = rec_thread + (long long)step * P;
                const double av[7] = {alg.p_e, alg.i_d, alg.i_q, alg.v_re, alg.v_im, alg.e_fd, alg.p_m};
#pragma unroll
                for (int k = 0; k < 7; ++k)
                    if (rec_umask & (1u << (DPS_N_STATES + k))) {
                        if (live) dps_rec_store(out + (long long)p.rec_fbase[DPS_N_STATES + k] * rec_cs, av[k]);
                    }
            }
        } else if (last_stage && (rec_mask >> DPS_N_STATES)) {
            const short* ch = p.rec_chan[g];
            double* out = p.rec + (long long)step * P + sim;
            const long long cs = (long long)p.n_steps * P;
            const double av[7] = {alg.p_e, alg.i_d, alg.i_q, alg.v_re, alg.v_im, alg.e_fd, alg.p_m};
#pragma unroll
            for (int k = 0; k < 7; ++k)
                if (rec_mask & (1u << (DPS_N_STATES + k))) dps_rec_store(out + ch[DPS_N_STATES + k] * cs, av[k]);
        }
        // extra observed buses (recorded voltages of non-generator buses): last stage only.  Every lane runs the same
        // number of iterations because the shuffles need the whole warp.
        if (last_stage && n_extra_iter > 0) {
            for (int it = 0; it < n_extra_iter; ++it) {
                const int r = p.n_gen + it * p.n_gen + g;
                const bool valid = live && (r < p.n_obs);
                const int rr = (r < p.n_obs) ? r : p.n_obs - 1;
                const double* zbase = p.zsub + (((long long)cfg * p.z_rows + rr) * p.z_cols) * 2 * zb + zs;
                double br = 0.0, bi = 0.0;
                for (int j = 0; j < p.n_gen; ++j) {
                    const double jr = cgrp[j].x, ji = cgrp[j].y;
                    const double zr_ = __ldg(zbase + (long long)(2 * j) * zb);
                    const double zi_ = __ldg(zbase + (long long)(2 * j + 1) * zb);
                    br += zr_ * jr - zi_ * ji;
                    bi += zr_ * ji + zi_ * jr;
                }
                if (valid) {
                    const int e = rr - p.n_gen;
                    if (p.obs_chan[e][0] >= 0) dps_rec_store(p.rec + ((long long)p.obs_chan[e][0] * p.n_steps + step) * P + sim, br);
                    if (p.obs_chan[e][1] >= 0) dps_rec_store(p.rec + ((long long)p.obs_chan[e][1] * p.n_steps + step) * P + sim, bi);
                }
            }
        }
    };

    for (; step < step_end; ++step) {
#if DPS_REC_SYNC_STEPS > 0
        // The warps of a block are independent; without a bound on their drift the 24-byte pieces they write into the
        // same 32-byte DRAM sectors of a recorder row arrive further apart than L2 keeps a dirty line, and every piece
        // costs a read-modify-write.  A block barrier every few steps keeps them within L2's reach.
        if ((step & (DPS_REC_SYNC_STEPS - 1)) == 0) __syncthreads();
#endif
        if (step == next_change) {
            ++seg;
            cfg = p.seg_cfg[seg];
            next_change = (seg + 1 < p.n_seg) ? p.seg_step[seg + 1] : 0x7fffffff;
            load_z_row(cfg);
        }
        if (p.ckpt_stride > 0 && (step % p.ckpt_stride) == 0 && live) {
            double* dst = p.ckpt + (((long long)(step / p.ckpt_stride) * p.n_gen + g) * DPS_N_STATES) * P + s;
#pragma unroll
            for (int k = 0; k < NS; ++k) dst[(long long)k * P] = x[k];
            if (NS < DPS_N_STATES)
                for (int k = NS; k < DPS_N_STATES; ++k) dst[(long long)k * P] = 0.0;
        }

        // explicit update.  x is the only loop-carried state vector; the stage input xt and the accumulator are
        // step-local, so no register copies are needed between steps.
        double f[DPS_N_STATES];
        if (SOLVER == DPS_EULER) {                    // solvers.py:34-46
            eval(x, f, true);
#pragma unroll
            for (int k = 0; k < NS; ++k) x[k] = x[k] + f[k] * dt;
        } else if (SOLVER == DPS_HEUN) {              // solvers.py:83-114: x0 + ((f0 + f1)/2)*dt
            // Evaluated as (x0 + f0*dt/2) + f1*dt/2: x0 and f0 die after the first stage, which frees 26 registers
            // for the price of one extra rounding per state and step (~1e-16 relative; parity budget 1e-9).
            double xt[NS];
            eval(x, f, false);
#pragma unroll
            for (int k = 0; k < NS; ++k) { xt[k] = fma(f[k], dt, x[k]); x[k] = fma(f[k], hdt, x[k]); }
            eval(xt, f, true);
#pragma unroll
            for (int k = 0; k < NS; ++k) x[k] = fma(f[k], hdt, x[k]);
        } else {                                      // solvers.py:150-214
            double xt[NS], acc[NS];
            eval(x, f, false);
#pragma unroll
            for (int k = 0; k < NS; ++k) { acc[k] = f[k]; xt[k] = x[k] + (f[k] * dt) * 0.5; }
            eval(xt, f, false);
#pragma unroll
            for (int k = 0; k < NS; ++k) { acc[k] = acc[k] + 2.0 * f[k]; xt[k] = x[k] + (f[k] * dt) * 0.5; }
            eval(xt, f, false);
#pragma unroll
            for (int k = 0; k < NS; ++k) { acc[k] = acc[k] + 2.0 * f[k]; xt[k] = x[k] + f[k] * dt; }
            eval(xt, f, true);
#pragma unroll
            for (int k = 0; k < NS; ++k) x[k] = x[k] + dps_div6((acc[k] + f[k]) * dt);
        }

        // recorder: post-step states (simulator.py:540-543); rec[chan][step][sim].  (Algebraics were stored by the last
        // stage evaluation.)  Compile-time mask: one running pointer per recorded field, advanced by P per step.
        if (RMASK >= 0) {
            int q = 0;
#pragma unroll
            for (int k = 0; k < DPS_N_STATES; ++k)
                if ((RMASK >> k) & 1) {
                    if (live) dps_rec_store(rec_run[q], (k < NS) ? x[k < NS ? k : 0] : 0.0);
                    rec_run[q] += P;
                    ++q;
                }
        } else if (rec_regular) {
            if (rec_umask & 0x1FFFu) {
                double* out = rec_thread + (long long)step * P;
#pragma unroll
                for (int grp4 = 0; grp4 < NS; grp4 += 4) {           // uniform tests in groups of four fields
                    if (rec_umask & (0xFu << grp4)) {
#pragma unroll
                        for (int k = grp4; k < grp4 + 4 && k < NS; ++k)
                            if (rec_umask & (1u << k)) {
                                if (live) dps_rec_store(out + (long long)p.rec_fbase[k] * rec_cs, x[k]);
                            }
                    }
                }
                if (NS < DPS_N_STATES) {
                    for (int k = NS; k < DPS_N_STATES; ++k)
                        if ((rec_umask & (1u << k)) && live) dps_rec_store(out + (long long)p.rec_fbase[k] * rec_cs, 0.0);
                }
            }
        } else if (rec_mask & 0x1FFFu) {
            const short* ch = p.rec_chan[g];
            double* out = p.rec + (long long)step * P + sim;
            const long long cs = (long long)p.n_steps * P;
#pragma unroll
            for (int grp4 = 0; grp4 < 16; grp4 += 4) {
                if (rec_mask & (0xFu << grp4) & 0x1FFFu) {
#pragma unroll
                    for (int k = grp4; k < grp4 + 4 && k < DPS_N_STATES; ++k)
                        if (rec_mask & (1u << k)) dps_rec_store(out + ch[k] * cs, (k < NS) ? x[k < NS ? k : 0] : 0.0);
                }
            }
        }
    }

    if (p.x_final != nullptr && live) {
        double* dst = p.x_final + (long long)g * DPS_N_STATES * P + s;
#pragma unroll
        for (int k = 0; k < NS; ++k) dst[(long long)k * P] = x[k];
        if (NS < DPS_N_STATES)
            for (int k = NS; k < DPS_N_STATES; ++k) dst[(long long)k * P] = 0.0;
    }
    if (queued) {
        // publish the chunk: every thread's state stores first (fence), then the group's progress counter
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) {
            // the ticket is re-read instead of kept: nothing of the queue stays live in registers across the step loop
            const int item = s_item, c = item / p.q_groups;
            dps_st_release(p.q_progress + (item - c * p.q_groups), c + 1);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// launch helpers (one instantiation set per G, see dps_forward_inst.cu)
// ---------------------------------------------------------------------------------------------------------
template <int G, int CM, int PMODE, int ZPS, int SOLVER, int RMASK = -1>
static cudaError_t launch_one(const FwdParams& p, cudaStream_t stream, int* regs) {
    using C = FwdCfg<G>;
    // shared Z + the usual recorder (omega of every generator, regular layout): compile-time recorder mask
    if constexpr (RMASK < 0 && ZPS == 0 && PMODE != 2 && CM >= 0) {
        if (p.rec_regular && p.rec_umask == 0x1u) return launch_one<G, CM, PMODE, ZPS, SOLVER, 0x1>(p, stream, regs);
    }
    auto kern = dps_step_kernel<G, CM, PMODE, ZPS, SOLVER, RMASK, false>;
    if (regs) {
        cudaFuncAttributes a;
        cudaError_t e = cudaFuncGetAttributes(&a, kern);
        if (e != cudaSuccess) return e;
        *regs = a.numRegs;
        return cudaSuccess;
    }
    constexpr int RS = DPS_Z_PAIRS ? ((G + 1) & ~1) : G;
    const size_t smem = sizeof(double) * (G * C::UPS + (ZPS == 0 ? p.n_cfg * G * RS * 2 : 0)) +
                        sizeof(short) * G * DPS_N_SLOTS + sizeof(double2) * C::WARPS * 64;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    const long long sims_per_block = (long long)C::WARPS * C::SPW;
    const long long blocks = (p.P + sims_per_block - 1) / sims_per_block;
    // the work-queue variant exists for the fast paths (batch-shared network, register / constant-bank constants)
    if constexpr (ZPS == 0 && PMODE != 2) {
        auto kern_q = dps_step_kernel<G, CM, PMODE, ZPS, SOLVER, RMASK, true>;
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
            return cudaGetLastError();
        }
    }
    kern<<<(unsigned)blocks, C::THREADS, smem, stream>>>(p);
    return cudaGetLastError();
}

template <int G>
cudaError_t launch_forward_g(const FwdParams& p, int cmask, int pmode, cudaStream_t stream, int* regs, int* spb) {
    {   // large batches of the hot configuration: two trajectories per thread (dps_forward_pair.cuh)
        cudaError_t e;
        if (try_pair_kernel<G>(p, cmask, pmode, stream, regs, spb, &e)) return e;
    }
    if (spb) *spb = FwdCfg<G>::WARPS * FwdCfg<G>::SPW;
    const int zps = (p.z_per_sim || !FwdCfg<G>::Z_SMEM) ? 1 : 0;
#define DPS_CASE(C, M, Z)                                                                          \
    if constexpr (Z == 1 || FwdCfg<G>::Z_SMEM) {                                                   \
        if (cmask == C && pmode == M && zps == Z) {                                                \
            if (p.solver == DPS_EULER) return launch_one<G, C, M, Z, DPS_EULER>(p, stream, regs);  \
            if (p.solver == DPS_HEUN) return launch_one<G, C, M, Z, DPS_HEUN>(p, stream, regs);    \
            return launch_one<G, C, M, Z, DPS_RK4>(p, stream, regs);                               \
        }                                                                                          \
    }
    DPS_CASE(0, 1, 0) DPS_CASE(0, 1, 1) DPS_CASE(0, 2, 0) DPS_CASE(0, 2, 1)
    DPS_CASE(7, 1, 0) DPS_CASE(7, 1, 1) DPS_CASE(7, 2, 0) DPS_CASE(7, 2, 1) DPS_CASE(7, 3, 0) DPS_CASE(7, 3, 1)
    DPS_CASE(-1, 1, 0) DPS_CASE(-1, 1, 1) DPS_CASE(-1, 2, 0) DPS_CASE(-1, 2, 1)
#undef DPS_CASE
    return cudaErrorInvalidValue;
}

}  // namespace dps
