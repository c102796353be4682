// dps_forward.cuh -- the persistent fused stepping kernel (sm_100a), instantiated per generator count by
// dps_forward_inst.cu (one translation unit per G so the build parallelises).
//
// Replaces the whole ``for t in self.time`` loop of PowerSystemSimulation.run() (reference
// src/diffpssi/power_sim_lib/simulator.py:523-543) for all parallel_sims trajectories with ONE launch:
// fault switching, every solver stage (solvers.py:26-46 / 75-117 / 143-217), the network solve V = Z I
// (solvers.py:83,102), the model equations (dps_model.cuh) and the recorder (simulator.py:540-546).
//
// Mapping:
//   * one thread per (trajectory, generator); a warp holds 32 consecutive trajectories of ONE generator, so every
//     global access is a coalesced 256-byte line and generator-dependent branches / constants are warp-uniform;
//   * the G warps of one 32-trajectory group exchange their Norton currents through double-buffered shared memory
//     with a single named barrier (bar.sync id, 32*G) per stage;
//   * states x, the step-start copy x0 and the stage accumulator stay in registers for all T steps; per step only the
//     recorder lines go to HBM.
// Specialisations:
//   PMODE 1 (fast): the DPS_N_REG_SLOTS leading constants (machine + init-derived set points) of the thread's
//           (trajectory, generator) are loaded ONCE into registers (per-sim or uniform alike); the remaining controller
//           constants must be batch-uniform and are warp-broadcast LDS.
//   PMODE 2 (generic): any constant may be per-sim; every read goes through the slot table (slow, kept for generality).
//   ZPS 0: the reduced impedance matrices of all fault configurations are shared by the batch and sit in shared
//          memory (row read as broadcast LDS.128);  ZPS 1: one matrix per trajectory, row held in registers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "dps_internal.h"
#include "dps_model.cuh"

namespace dps {

template <int G>
struct FwdCfg {
    // 32-trajectory groups per block: keep blocks at >= 96 threads
    static constexpr int GROUPS = (G >= 3) ? 1 : (G == 2 ? 2 : 4);
    static constexpr int THREADS = 32 * G * GROUPS;
    // resident blocks per SM the register allocator should aim for (register file: 64K x 32 bit per SM, warps are
    // allocated in groups of 4)
    static constexpr int MIN_BLOCKS = (THREADS >= 288) ? 1 : (THREADS >= 192 ? 2 : (THREADS >= 128 ? 3 : 4));
};

__device__ __forceinline__ void group_barrier(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

template <int G, bool CTRL, int PMODE, int ZPS, int SOLVER>
__global__ void __launch_bounds__(FwdCfg<G>::THREADS, FwdCfg<G>::MIN_BLOCKS)
dps_step_kernel(const FwdParams p) {
    constexpr int GROUPS = FwdCfg<G>::GROUPS;
    constexpr int NS = CTRL ? DPS_N_STATES : 6;                                  // live states per generator
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* s_upar = reinterpret_cast<double*>(smem_raw);                        // [G][DPS_N_SLOTS]
    double2* s_cur = reinterpret_cast<double2*>(s_upar + G * DPS_N_SLOTS);       // [2][GROUPS][G][32] (re, im)
    double* s_z = reinterpret_cast<double*>(s_cur + 2 * GROUPS * G * 32);        // ZPS==0: [n_cfg][G][G][2]
    short* s_vary = reinterpret_cast<short*>(s_z + (ZPS == 0 ? p.n_cfg * G * G * 2 : 0));   // [G][DPS_N_SLOTS]

    const int lane = threadIdx.x & 31;
    // broadcast from lane 0: tells the compiler the warp index -- and with it the generator index, the controller
    // flags, the recorder mask and every constant-table address -- is warp-uniform (uniform registers, plain branches)
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int gw = warp % G;                       // generator slot of this warp
    const int grp = warp / G;                      // trajectory group inside the block
    const bool gen_live = gw < p.n_gen;            // template G may exceed the real generator count
    const int g = gen_live ? gw : p.n_gen - 1;
    const long long P = p.P;
    const long long sim = ((long long)blockIdx.x * GROUPS + grp) * 32 + lane;
    const bool live = (sim < P) && gen_live;
    const long long s = (sim < P) ? sim : (P - 1);

    for (int i = threadIdx.x; i < G * DPS_N_SLOTS; i += blockDim.x) {
        const bool ok = (i / DPS_N_SLOTS) < p.n_gen;
        s_upar[i] = ok ? p.upar[i] : 0.0;
        if (PMODE == 2) s_vary[i] = ok ? p.slot_vary[i] : (short)-1;
    }
    if (ZPS == 0) {
        // all configurations of the shared reduced impedance matrix: [cfg][row < G][col < G](re, im)
        for (int i = threadIdx.x; i < p.n_cfg * G * G * 2; i += blockDim.x) {
            const int part = i & 1, col = (i >> 1) % G, row = (i >> 1) / G % G, cfg = (i >> 1) / (G * G);
            const bool ok = row < p.n_gen && col < p.n_gen;
            s_z[i] = ok ? p.zsub[(((long long)cfg * p.z_rows + row) * p.z_cols + col) * 2 + part] : 0.0;
        }
    }
    __syncthreads();

    const double* up = s_upar + g * DPS_N_SLOTS;
    const int ctrl = CTRL ? p.ctrl[g] : 0;

    // ---- parameter access -------------------------------------------------------------------------------------------
    double pm[DPS_N_REG_SLOTS];
    if (PMODE == 1) {
#pragma unroll
        for (int k = 0; k < DPS_N_REG_SLOTS; ++k) {
            const int vi = p.slot_vary[g * DPS_N_SLOTS + k];
            pm[k] = (vi >= 0) ? __ldg(p.vpar + (long long)vi * P + s) : up[k];
        }
    }
    ParamRegs par_fast{pm, up};
    ParamReader<true> par_gen;
    par_gen.up = up;
    par_gen.sv = s_vary + g * DPS_N_SLOTS;
    par_gen.vpar = p.vpar;
    par_gen.P = P;
    par_gen.s = s;

    double x[NS];
    {
        const double* src = p.x0 + (long long)g * DPS_N_STATES * P + s;   // state at step_begin
#pragma unroll
        for (int k = 0; k < NS; ++k) x[k] = src[(long long)k * P];
    }

    // ---- network rows -------------------------------------------------------------------------------------------------
    double zr[ZPS ? G : 1], zi[ZPS ? G : 1];
    const long long zb = p.z_per_sim ? P : 1;
    const long long zs = p.z_per_sim ? s : 0;
    const double2* zrow = nullptr;
    auto load_z_row = [&](int cfg) {
        if (ZPS) {
            const double* base = p.zsub + (((long long)cfg * p.z_rows + g) * p.z_cols) * 2 * zb + zs;
#pragma unroll
            for (int j = 0; j < G; ++j) {
                const bool ok = j < p.n_gen;
                zr[ZPS ? j : 0] = ok ? __ldg(base + (long long)(2 * j) * zb) : 0.0;
                zi[ZPS ? j : 0] = ok ? __ldg(base + (long long)(2 * j + 1) * zb) : 0.0;
            }
        } else {
            zrow = reinterpret_cast<const double2*>(s_z) + ((long long)cfg * G + gw) * G;
        }
    };

    // configuration schedule: (first step, cfg) change points, ascending
    int seg = 0;
    while (seg + 1 < p.n_seg && p.seg_step[seg + 1] <= p.step_begin) ++seg;
    int cfg = p.seg_cfg[seg];
    int next_change = (seg + 1 < p.n_seg) ? p.seg_step[seg + 1] : 0x7fffffff;
    load_z_row(cfg);

    const double dt = p.dt, hdt = 0.5 * p.dt;
    const int bar_id = grp + 1;
    double2* cur_grp = s_cur + grp * G * 32 + lane;            // + buf*GROUPS*G*32 + j*32
    const unsigned rec_mask = p.rec_mask[g];                   // warp-uniform
    int buf = 0;
    Algebraics alg;
    int step = p.step_begin;

    // One stage evaluation at state X (reference: the body of one `voltages = matmul(...); for bus: update_voltages;
    // differential()` pass of solvers.py): Norton currents -> shared memory -> V = Z I -> derivatives f.
    auto eval = [&](const double* X, double* f, bool last_stage) {
        double sd, cd;
        dps_sincos(X[DPS_X_DELTA], &sd, &cd);
        double i_re, i_im;
        if (PMODE == 1) injection(X, par_fast, sd, cd, i_re, i_im);
        else injection(X, par_gen, sd, cd, i_re, i_im);
        double2* cb = cur_grp + buf * (GROUPS * G * 32);
        cb[gw * 32] = gen_live ? make_double2(i_re, i_im) : make_double2(0.0, 0.0);
        group_barrier(bar_id, 32 * G);
        // V_g = sum_j Z[g][j] I_j   (solvers.py:83; zero injections at non-generator buses are skipped).
        // Two partial sums per component keep four independent DFMA chains in flight.
        double vra = 0.0, vrb = 0.0, via = 0.0, vib = 0.0;
#pragma unroll
        for (int j = 0; j < G; ++j) {
            const double2 cj = cb[j * 32];
            double zre, zim;
            if (ZPS) { zre = zr[ZPS ? j : 0]; zim = zi[ZPS ? j : 0]; }
            else { const double2 z = zrow[j]; zre = z.x; zim = z.y; }
            vra = fma(zre, cj.x, vra);
            vrb = fma(-zim, cj.y, vrb);
            via = fma(zre, cj.y, via);
            vib = fma(zim, cj.x, vib);
        }
        const double v_re = vra + vrb, v_im = via + vib;
        if (PMODE == 1) stage_derivatives<CTRL>(X, par_fast, ctrl, p.fn, sd, cd, v_re, v_im, f, alg);
        else stage_derivatives<CTRL>(X, par_gen, ctrl, p.fn, sd, cd, v_re, v_im, f, alg);

        // extra observed buses (recorded voltages of non-generator buses): last stage only
        if (last_stage && p.n_obs > p.n_gen && gen_live && sim < P) {
            for (int r = p.n_gen + g; r < p.n_obs; r += p.n_gen) {
                const double* base = p.zsub + (((long long)cfg * p.z_rows + r) * p.z_cols) * 2 * zb + zs;
                double br = 0.0, bi = 0.0;
                for (int j = 0; j < p.n_gen; ++j) {
                    const double zr_ = __ldg(base + (long long)(2 * j) * zb);
                    const double zi_ = __ldg(base + (long long)(2 * j + 1) * zb);
                    const double2 cj = cb[j * 32];
                    br += zr_ * cj.x - zi_ * cj.y;
                    bi += zr_ * cj.y + zi_ * cj.x;
                }
                const int e = r - p.n_gen;
                if (p.obs_chan[e][0] >= 0) p.rec[((long long)p.obs_chan[e][0] * p.n_steps + step) * P + sim] = br;
                if (p.obs_chan[e][1] >= 0) p.rec[((long long)p.obs_chan[e][1] * p.n_steps + step) * P + sim] = bi;
            }
        }
        buf ^= 1;
    };

    for (; step < p.step_end; ++step) {
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
            double xt[NS], acc[NS];
            eval(x, f, false);
#pragma unroll
            for (int k = 0; k < NS; ++k) { acc[k] = f[k]; xt[k] = x[k] + f[k] * dt; }
            eval(xt, f, true);
            // ((f0+f1)/2)*dt == (f0+f1)*(dt/2) bit for bit: halving is exact
#pragma unroll
            for (int k = 0; k < NS; ++k) x[k] = x[k] + (acc[k] + f[k]) * hdt;
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
            for (int k = 0; k < NS; ++k) x[k] = x[k] + ((acc[k] + f[k]) * dt) / 6.0;
        }

        // recorder: post-step states, last-stage algebraics (simulator.py:540-543); rec[chan][step][sim].
        // The mask is warp-uniform; fields are tested in groups of four so an all-clear group costs one branch.
        if (rec_mask && live) {
            const short* ch = p.rec_chan[g];
            double* out = p.rec + (long long)step * P + sim;
            const long long cs = (long long)p.n_steps * P;
            const double av[7] = {alg.p_e, alg.i_d, alg.i_q, alg.v_re, alg.v_im, alg.e_fd, alg.p_m};
#pragma unroll
            for (int grp4 = 0; grp4 < DPS_N_FIELDS; grp4 += 4) {
                if (rec_mask & (0xFu << grp4)) {
#pragma unroll
                    for (int k = grp4; k < grp4 + 4; ++k) {
                        if (rec_mask & (1u << k)) {
                            const double v = (k < NS) ? x[k < NS ? k : 0] : (k < DPS_N_STATES ? 0.0 : av[k >= DPS_N_STATES ? k - DPS_N_STATES : 0]);
                            out[ch[k] * cs] = v;
                        }
                    }
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
}

// ---------------------------------------------------------------------------------------------------------
// launch helpers (one instantiation set per G, see dps_forward_inst.cu)
// ---------------------------------------------------------------------------------------------------------
template <int G, bool CTRL, int PMODE, int ZPS, int SOLVER>
static cudaError_t launch_one(const FwdParams& p, cudaStream_t stream, int* regs) {
    using C = FwdCfg<G>;
    auto kern = dps_step_kernel<G, CTRL, PMODE, ZPS, SOLVER>;
    if (regs) {
        cudaFuncAttributes a;
        cudaError_t e = cudaFuncGetAttributes(&a, kern);
        if (e != cudaSuccess) return e;
        *regs = a.numRegs;
        return cudaSuccess;
    }
    const size_t smem = sizeof(double) * (G * DPS_N_SLOTS + 2 * C::GROUPS * G * 64 + (ZPS == 0 ? p.n_cfg * G * G * 2 : 0)) +
                        sizeof(short) * G * DPS_N_SLOTS;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    const long long sims_per_block = 32LL * C::GROUPS;
    const long long blocks = (p.P + sims_per_block - 1) / sims_per_block;
    kern<<<(unsigned)blocks, C::THREADS, smem, stream>>>(p);
    return cudaGetLastError();
}

template <int G>
cudaError_t launch_forward_g(const FwdParams& p, bool ctrl, int pmode, cudaStream_t stream, int* regs) {
    const int zps = p.z_per_sim ? 1 : 0;
#define DPS_CASE(C, M, Z)                                                                          \
    if (ctrl == C && pmode == M && zps == Z) {                                                     \
        if (p.solver == DPS_EULER) return launch_one<G, C, M, Z, DPS_EULER>(p, stream, regs);      \
        if (p.solver == DPS_HEUN) return launch_one<G, C, M, Z, DPS_HEUN>(p, stream, regs);        \
        return launch_one<G, C, M, Z, DPS_RK4>(p, stream, regs);                                   \
    }
    DPS_CASE(false, 1, 0) DPS_CASE(false, 1, 1) DPS_CASE(false, 2, 0) DPS_CASE(false, 2, 1)
    DPS_CASE(true, 1, 0) DPS_CASE(true, 1, 1) DPS_CASE(true, 2, 0) DPS_CASE(true, 2, 1)
#undef DPS_CASE
    return cudaErrorInvalidValue;
}

}  // namespace dps
