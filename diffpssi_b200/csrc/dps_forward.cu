// dps_forward.cu -- the persistent fused stepping kernel (sm_100a).
//
// Replaces the whole ``for t in self.time`` loop of PowerSystemSimulation.run() (reference
// src/diffpssi/power_sim_lib/simulator.py:523-543) for all parallel_sims trajectories with ONE launch:
// fault switching, every solver stage (solvers.py:26-46 / 75-117 / 143-217), the network solve V = Z I
// (solvers.py:83,102), the model equations (dps_model.cuh) and the recorder (simulator.py:540-546).
//
// Mapping:
//   * one thread per (trajectory, generator); a warp holds 32 consecutive trajectories of ONE generator, so
//     every global access is a coalesced 256-byte line and generator-dependent branches / constants are
//     warp-uniform;
//   * the G warps of one 32-trajectory group exchange their Norton currents through double-buffered shared
//     memory with a single named barrier (bar.sync id, 32*G) per stage;
//   * states x, the step-start copy x0, the stage accumulator and this generator's row of the reduced
//     impedance matrix stay in registers for all T steps; per step only the recorder lines go to HBM.
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
    // resident blocks per SM the register allocator should aim for (~12 warps/SM)
    static constexpr int MIN_BLOCKS = (THREADS >= 320) ? 1 : (THREADS >= 192 ? 2 : (THREADS >= 128 ? 3 : 4));
};

__device__ __forceinline__ void group_barrier(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

template <int G, bool CTRL, bool VARY>
__global__ void __launch_bounds__(FwdCfg<G>::THREADS, FwdCfg<G>::MIN_BLOCKS)
dps_step_kernel(const FwdParams p) {
    constexpr int GROUPS = FwdCfg<G>::GROUPS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* s_upar = reinterpret_cast<double*>(smem_raw);                        // [G][DPS_N_SLOTS]
    double* s_cur = s_upar + G * DPS_N_SLOTS;                                    // [2][GROUPS][G][2][32]
    short* s_vary = reinterpret_cast<short*>(s_cur + 2 * GROUPS * G * 2 * 32);   // [G][DPS_N_SLOTS]

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int gw = warp % G;                       // generator slot of this warp
    const int grp = warp / G;                      // trajectory group inside the block
    const bool gen_live = gw < p.n_gen;            // template G may exceed the real generator count
    const int g = gen_live ? gw : p.n_gen - 1;
    const long long P = p.P;
    const long long sim = ((long long)blockIdx.x * GROUPS + grp) * 32 + lane;
    const bool live = (sim < P) && gen_live;
    const long long s = (sim < P) ? sim : (P - 1);

    for (int i = threadIdx.x; i < G * DPS_N_SLOTS; i += blockDim.x) {
        const int gi = i / DPS_N_SLOTS;
        const bool ok = gi < p.n_gen;
        s_upar[i] = ok ? p.upar[i] : 0.0;
        s_vary[i] = ok ? p.slot_vary[i] : (short)-1;
    }
    __syncthreads();

    ParamReader<VARY> par;
    par.up = s_upar + g * DPS_N_SLOTS;
    par.sv = s_vary + g * DPS_N_SLOTS;
    par.vpar = p.vpar;
    par.P = P;
    par.s = s;
    const int ctrl = CTRL ? p.ctrl[g] : 0;

    double x[DPS_N_STATES], x0[DPS_N_STATES], acc[DPS_N_STATES];
    {
        const double* src = p.x0 + (long long)g * DPS_N_STATES * P + s;   // state at step_begin
#pragma unroll
        for (int k = 0; k < DPS_N_STATES; ++k) x[k] = src[(long long)k * P];
    }

    // this generator's row of the reduced impedance matrix, (re, im) per generator column
    double zr[G], zi[G];
    const long long zb = p.z_per_sim ? P : 1;
    const long long zs = p.z_per_sim ? s : 0;
    auto load_z_row = [&](int cfg) {
        const double* base = p.zsub + (((long long)cfg * p.z_rows + g) * p.z_cols) * 2 * zb + zs;
#pragma unroll
        for (int j = 0; j < G; ++j) {
            const bool ok = j < p.n_gen;
            zr[j] = ok ? __ldg(base + (long long)(2 * j) * zb) : 0.0;
            zi[j] = ok ? __ldg(base + (long long)(2 * j + 1) * zb) : 0.0;
        }
    };

    // configuration schedule: (first step, cfg) change points, ascending
    int seg = 0;
    while (seg + 1 < p.n_seg && p.seg_step[seg + 1] <= p.step_begin) ++seg;
    int cfg = p.seg_cfg[seg];
    int next_change = (seg + 1 < p.n_seg) ? p.seg_step[seg + 1] : 0x7fffffff;
    load_z_row(cfg);

    const int n_stages = (p.solver == DPS_EULER) ? 1 : (p.solver == DPS_HEUN ? 2 : 4);
    const double dt = p.dt;
    const int bar_id = grp + 1;
    double* cur_grp = s_cur + (long long)grp * G * 64;         // + buf*GROUPS*G*64 + j*64 + {0,32} + lane
    const unsigned rec_mask = live ? p.rec_mask[g] : 0u;
    int buf = 0;
    Algebraics alg;

    for (int step = p.step_begin; step < p.step_end; ++step) {
        if (step == next_change) {
            ++seg;
            cfg = p.seg_cfg[seg];
            next_change = (seg + 1 < p.n_seg) ? p.seg_step[seg + 1] : 0x7fffffff;
            load_z_row(cfg);
        }
        if (p.ckpt_stride > 0 && (step % p.ckpt_stride) == 0 && live) {
            double* dst = p.ckpt + (((long long)(step / p.ckpt_stride) * p.n_gen + g) * DPS_N_STATES) * P + s;
#pragma unroll
            for (int k = 0; k < DPS_N_STATES; ++k) dst[(long long)k * P] = x[k];
        }
#pragma unroll
        for (int k = 0; k < DPS_N_STATES; ++k) x0[k] = x[k];

        for (int st = 0; st < n_stages; ++st) {
            double sd, cd;
            sincos(x[DPS_X_DELTA], &sd, &cd);
            double i_re, i_im;
            injection(x, par, sd, cd, i_re, i_im);
            double* cb = cur_grp + buf * (GROUPS * G * 64);
            cb[gw * 64 + lane] = gen_live ? i_re : 0.0;
            cb[gw * 64 + 32 + lane] = gen_live ? i_im : 0.0;
            group_barrier(bar_id, 32 * G);
            // V_g = sum_j Z[g][j] I_j   (solvers.py:83; zero injections at non-generator buses are skipped)
            double v_re = 0.0, v_im = 0.0;
#pragma unroll
            for (int j = 0; j < G; ++j) {
                const double jr = cb[j * 64 + lane], ji = cb[j * 64 + 32 + lane];
                v_re += zr[j] * jr - zi[j] * ji;
                v_im += zr[j] * ji + zi[j] * jr;
            }
            double f[DPS_N_STATES];
            stage_derivatives<CTRL>(x, par, ctrl, p.fn, sd, cd, v_re, v_im, f, alg);

            // extra observed buses (recorded voltages of non-generator buses): last stage only
            if (p.n_obs > p.n_gen && st == n_stages - 1 && gen_live && sim < P) {
                for (int r = p.n_gen + g; r < p.n_obs; r += p.n_gen) {
                    const double* base = p.zsub + (((long long)cfg * p.z_rows + r) * p.z_cols) * 2 * zb + zs;
                    double br = 0.0, bi = 0.0;
                    for (int j = 0; j < p.n_gen; ++j) {
                        const double zr_ = __ldg(base + (long long)(2 * j) * zb);
                        const double zi_ = __ldg(base + (long long)(2 * j + 1) * zb);
                        const double jr = cb[j * 64 + lane], ji = cb[j * 64 + 32 + lane];
                        br += zr_ * jr - zi_ * ji;
                        bi += zr_ * ji + zi_ * jr;
                    }
                    const int e = r - p.n_gen;
                    if (p.obs_chan[e][0] >= 0) p.rec[((long long)p.obs_chan[e][0] * p.n_steps + step) * P + sim] = br;
                    if (p.obs_chan[e][1] >= 0) p.rec[((long long)p.obs_chan[e][1] * p.n_steps + step) * P + sim] = bi;
                }
            }
            buf ^= 1;

            // explicit update, solvers.py:43 (Euler), :94/:109-111 (Heun), :162/:180/:195/:211-212 (RK4)
            if (p.solver == DPS_EULER) {
#pragma unroll
                for (int k = 0; k < DPS_N_STATES; ++k) x[k] = x0[k] + f[k] * dt;
            } else if (p.solver == DPS_HEUN) {
                if (st == 0) {
#pragma unroll
                    for (int k = 0; k < DPS_N_STATES; ++k) { acc[k] = f[k]; x[k] = x0[k] + f[k] * dt; }
                } else {
#pragma unroll
                    for (int k = 0; k < DPS_N_STATES; ++k) x[k] = x0[k] + ((acc[k] + f[k]) / 2.0) * dt;
                }
            } else {
                if (st == 0) {
#pragma unroll
                    for (int k = 0; k < DPS_N_STATES; ++k) { acc[k] = f[k]; x[k] = x0[k] + (f[k] * dt) / 2.0; }
                } else if (st == 1) {
#pragma unroll
                    for (int k = 0; k < DPS_N_STATES; ++k) { acc[k] = acc[k] + 2.0 * f[k]; x[k] = x0[k] + (f[k] * dt) / 2.0; }
                } else if (st == 2) {
#pragma unroll
                    for (int k = 0; k < DPS_N_STATES; ++k) { acc[k] = acc[k] + 2.0 * f[k]; x[k] = x0[k] + f[k] * dt; }
                } else {
#pragma unroll
                    for (int k = 0; k < DPS_N_STATES; ++k) x[k] = x0[k] + ((acc[k] + f[k]) * dt) / 6.0;
                }
            }
        }

        // recorder: post-step states, last-stage algebraics (simulator.py:540-543); rec[chan][step][sim]
        if (rec_mask) {
            const short* ch = p.rec_chan[g];
            double* out = p.rec + (long long)step * P + sim;
            const long long cs = (long long)p.n_steps * P;
#pragma unroll
            for (int k = 0; k < DPS_N_STATES; ++k)
                if (rec_mask & (1u << k)) out[ch[k] * cs] = x[k];
            const double av[7] = {alg.p_e, alg.i_d, alg.i_q, alg.v_re, alg.v_im, alg.e_fd, alg.p_m};
#pragma unroll
            for (int k = 0; k < 7; ++k)
                if (rec_mask & (1u << (DPS_N_STATES + k))) out[ch[DPS_N_STATES + k] * cs] = av[k];
        }
    }

    if (p.x_final != nullptr && live) {
        double* dst = p.x_final + (long long)g * DPS_N_STATES * P + s;
#pragma unroll
        for (int k = 0; k < DPS_N_STATES; ++k) dst[(long long)k * P] = x[k];
    }
}

// ---------------------------------------------------------------------------------------------------------
// launch table
// ---------------------------------------------------------------------------------------------------------
template <int G, bool CTRL, bool VARY>
static cudaError_t launch_one(const FwdParams& p, cudaStream_t stream, int* regs) {
    using C = FwdCfg<G>;
    auto kern = dps_step_kernel<G, CTRL, VARY>;
    if (regs) {
        cudaFuncAttributes a;
        cudaError_t e = cudaFuncGetAttributes(&a, kern);
        if (e != cudaSuccess) return e;
        *regs = a.numRegs;
        return cudaSuccess;
    }
    const size_t smem = sizeof(double) * (G * DPS_N_SLOTS + 2 * C::GROUPS * G * 64) + sizeof(short) * G * DPS_N_SLOTS;
    const long long sims_per_block = 32LL * C::GROUPS;
    const long long blocks = (p.P + sims_per_block - 1) / sims_per_block;
    kern<<<(unsigned)blocks, C::THREADS, smem, stream>>>(p);
    return cudaGetLastError();
}

template <int G>
static cudaError_t launch_g(const FwdParams& p, bool ctrl, bool vary, cudaStream_t stream, int* regs) {
    if (ctrl) return vary ? launch_one<G, true, true>(p, stream, regs) : launch_one<G, true, false>(p, stream, regs);
    return vary ? launch_one<G, false, true>(p, stream, regs) : launch_one<G, false, false>(p, stream, regs);
}

int fwd_template_g(int n_gen) {
    static const int sizes[] = {1, 2, 3, 4, 5, 6, 8, 10, 12};
    for (int v : sizes)
        if (n_gen <= v) return v;
    return -1;
}

void fwd_geometry(int g_template, int* threads, int* sims_per_block) {
    int groups = (g_template >= 3) ? 1 : (g_template == 2 ? 2 : 4);
    *threads = 32 * g_template * groups;
    *sims_per_block = 32 * groups;
}

cudaError_t launch_forward(const FwdParams& p, int g_template, bool ctrl, bool vary, cudaStream_t stream, int* regs) {
    switch (g_template) {
        case 1: return launch_g<1>(p, ctrl, vary, stream, regs);
        case 2: return launch_g<2>(p, ctrl, vary, stream, regs);
        case 3: return launch_g<3>(p, ctrl, vary, stream, regs);
        case 4: return launch_g<4>(p, ctrl, vary, stream, regs);
        case 5: return launch_g<5>(p, ctrl, vary, stream, regs);
        case 6: return launch_g<6>(p, ctrl, vary, stream, regs);
        case 8: return launch_g<8>(p, ctrl, vary, stream, regs);
        case 10: return launch_g<10>(p, ctrl, vary, stream, regs);
        case 12: return launch_g<12>(p, ctrl, vary, stream, regs);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace dps
