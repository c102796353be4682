// dps_forward_wide.cu -- the stepping kernel for systems with 17..128 generators (e.g. the 500-bus / 100-generator
// configuration of BASELINE.json), where the generators of one trajectory no longer fit in one warp.
//
// Same per-thread work as dps_step_kernel (dps_forward.cuh): one thread per (trajectory, generator), states in
// registers for all steps, dps_model.cuh for the equations.  Differences:
//   * a trajectory occupies WG = ceil(G/32) warps; a block holds SPB trajectories; Norton currents are exchanged through
//     double-buffered shared memory with one __syncthreads per stage;
//   * the reduced impedance matrix of the ACTIVE fault configuration lives in shared memory, transposed
//     ([col j][row g], 16 G^2 bytes: 160 KB for G = 100), and is reloaded when the configuration changes;
//     each thread streams its row through LDS.128 (network solve: G complex MACs per thread and stage -- for G = 100
//     this is 84 % of the flops);
//   * the network must be shared by the batch (a per-trajectory 100x100 complex Z would be 160 KB per trajectory);
//   * constants: the register set (machine + set points) may differ per trajectory, controller constants are uniform
//     (PMODE 1) or common to all generators (PMODE 3).
// Bound: shared-memory bandwidth of the Z stream (one LDS.128 per complex MAC, no register blocking yet); see DESIGN.md.
#include <cuda_runtime.h>
#include <stdint.h>
#include "dps_internal.h"
#include "dps_model.cuh"

namespace dps {

constexpr int WIDE_THREADS = 512;

template <int CM, int PMODE, int SOLVER>
__global__ void __launch_bounds__(WIDE_THREADS, 1)
dps_step_wide_kernel(const WideParams p, int WG, int SPB) {
    constexpr bool CTRL = (CM != 0);
    constexpr int NS = CTRL ? DPS_N_STATES : 6;
    constexpr int UPS = DPS_N_SLOTS + 1;
    const int G = p.n_gen;
    const int GP = WG * 32;                                                       // padded generator count
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double2* s_z = reinterpret_cast<double2*>(smem_raw);                          // [G][G] active configuration, transposed
    double2* s_cur = s_z + G * G;                                                 // [2][SPB][GP]
    double* s_upar = reinterpret_cast<double*>(s_cur + 2 * SPB * GP);             // [G][UPS]

    const int t = threadIdx.x;
    const int sl = t / GP;                          // trajectory slot in the block
    const int gl = t - sl * GP;                     // generator slot
    const bool gen_live = (sl < SPB) && (gl < G);
    const int g = (gl < G) ? gl : G - 1;
    const long long P = p.P;
    const long long sim = (long long)blockIdx.x * SPB + sl;
    const bool live = gen_live && (sim < P);
    const long long s = (sim < P) ? sim : (P - 1);

    for (int i = t; i < G * DPS_N_SLOTS; i += blockDim.x) {
        const int gi = i / DPS_N_SLOTS, k = i - gi * DPS_N_SLOTS;
        s_upar[gi * UPS + k] = p.upar[i];
    }
    auto load_z = [&](int cfg) {
        // transposed copy of configuration cfg: s_z[j*G + g] = Z[g][j]
        for (int i = t; i < G * G; i += blockDim.x) {
            const int row = i % G, col = i / G;
            const double* e = p.zsub + (((long long)cfg * p.z_rows + row) * p.z_cols + col) * 2;
            s_z[i] = make_double2(e[0], e[1]);
        }
    };
    int seg = 0;
    while (seg + 1 < p.n_seg && p.seg_step[seg + 1] <= p.step_begin) ++seg;
    int cfg = p.seg_cfg[seg];
    int next_change = (seg + 1 < p.n_seg) ? p.seg_step[seg + 1] : 0x7fffffff;
    load_z(cfg);
    __syncthreads();

    const double* up = s_upar + g * UPS;
    const int ctrl = (CM >= 0) ? CM : p.ctrl[g];
    double pm[DPS_N_REG_SLOTS];
#pragma unroll
    for (int k = 0; k < DPS_N_REG_SLOTS; ++k) {
        const int vi = p.reg_vary[g][k];
        pm[k] = (vi >= 0) ? __ldg(p.vpar + (long long)vi * P + s) : up[k];
    }
    ParamRegs par_fast{pm, up};
    ParamConst par_const{pm, p.cpar};

    double x[NS];
    {
        const double* src = p.x0 + (long long)g * DPS_N_STATES * P + s;
#pragma unroll
        for (int k = 0; k < NS; ++k) x[k] = src[(long long)k * P];
    }

    const double dt = p.dt, hdt = 0.5 * p.dt;
    const unsigned rec_mask = live ? p.rec_mask[g] : 0u;
    const int n_extra = p.n_obs - G;
    int step = p.step_begin;
    int buf = 0;
    const double2* zcol = s_z + g;                                   // element j at zcol[j*G]

    auto eval = [&](const double* X, double* f, bool last_stage) {
        Algebraics alg;
        double sd, cd;
        dps_sincos(X[DPS_X_DELTA], &sd, &cd);
        double i_re, i_im;
        if (PMODE == 3) injection(X, par_const, sd, cd, i_re, i_im);
        else injection(X, par_fast, sd, cd, i_re, i_im);
        double2* cb = s_cur + (buf * SPB + ((sl < SPB) ? sl : SPB - 1)) * GP;
        if (sl < SPB) cb[gl] = gen_live ? make_double2(i_re, i_im) : make_double2(0.0, 0.0);
        __syncthreads();
        buf ^= 1;
        double vra = 0.0, vrb = 0.0, via = 0.0, vib = 0.0;
#pragma unroll 4
        for (int j = 0; j < G; ++j) {
            const double2 cj = cb[j];
            const double2 z = zcol[j * G];
            vra = fma(z.x, cj.x, vra);
            vrb = fma(-z.y, cj.y, vrb);
            via = fma(z.x, cj.y, via);
            vib = fma(z.y, cj.x, vib);
        }
        const double v_re = vra + vrb, v_im = via + vib;
        if (PMODE == 3) stage_derivatives<CTRL>(X, par_const, ctrl, p.fn, sd, cd, v_re, v_im, f, alg);
        else stage_derivatives<CTRL>(X, par_fast, ctrl, p.fn, sd, cd, v_re, v_im, f, alg);
        if (last_stage && (rec_mask >> DPS_N_STATES)) {
            const short* ch = p.rec_chan[g];
            double* out = p.rec + (long long)step * P + sim;
            const long long cs = (long long)p.n_steps * P;
            const double av[7] = {alg.p_e, alg.i_d, alg.i_q, alg.v_re, alg.v_im, alg.e_fd, alg.p_m};
#pragma unroll
            for (int k = 0; k < 7; ++k)
                if (rec_mask & (1u << (DPS_N_STATES + k))) __stcs(out + ch[DPS_N_STATES + k] * cs, av[k]);
        }
        // recorded voltages of non-generator buses: row r of the reduced matrix is read from global memory
        if (last_stage && n_extra > 0 && live) {
            for (int r = G + g; r < p.n_obs; r += G) {
                const double* zbase = p.zsub + (((long long)cfg * p.z_rows + r) * p.z_cols) * 2;
                double br = 0.0, bi = 0.0;
                for (int j = 0; j < G; ++j) {
                    const double zr_ = __ldg(zbase + 2 * j), zi_ = __ldg(zbase + 2 * j + 1);
                    const double2 cj = cb[j];
                    br += zr_ * cj.x - zi_ * cj.y;
                    bi += zr_ * cj.y + zi_ * cj.x;
                }
                const int e = r - G;
                if (p.obs_chan[e][0] >= 0) __stcs(p.rec + ((long long)p.obs_chan[e][0] * p.n_steps + step) * P + sim, br);
                if (p.obs_chan[e][1] >= 0) __stcs(p.rec + ((long long)p.obs_chan[e][1] * p.n_steps + step) * P + sim, bi);
            }
        }
    };

    for (; step < p.step_end; ++step) {
        if (step == next_change) {
            ++seg;
            cfg = p.seg_cfg[seg];
            next_change = (seg + 1 < p.n_seg) ? p.seg_step[seg + 1] : 0x7fffffff;
            __syncthreads();                          // everyone is done with the old matrix
            load_z(cfg);
            __syncthreads();
        }
        if (p.ckpt_stride > 0 && (step % p.ckpt_stride) == 0 && live) {
            double* dst = p.ckpt + (((long long)(step / p.ckpt_stride) * G + g) * DPS_N_STATES) * P + s;
#pragma unroll
            for (int k = 0; k < NS; ++k) dst[(long long)k * P] = x[k];
            if (NS < DPS_N_STATES)
                for (int k = NS; k < DPS_N_STATES; ++k) dst[(long long)k * P] = 0.0;
        }
        double f[DPS_N_STATES];
        if (SOLVER == DPS_EULER) {
            eval(x, f, true);
#pragma unroll
            for (int k = 0; k < NS; ++k) x[k] = x[k] + f[k] * dt;
        } else if (SOLVER == DPS_HEUN) {
            double xt[NS];
            eval(x, f, false);
#pragma unroll
            for (int k = 0; k < NS; ++k) { xt[k] = fma(f[k], dt, x[k]); x[k] = fma(f[k], hdt, x[k]); }
            eval(xt, f, true);
#pragma unroll
            for (int k = 0; k < NS; ++k) x[k] = fma(f[k], hdt, x[k]);
        } else {
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
        if (rec_mask & 0x1FFFu) {
            const short* ch = p.rec_chan[g];
            double* out = p.rec + (long long)step * P + sim;
            const long long cs = (long long)p.n_steps * P;
#pragma unroll
            for (int k = 0; k < DPS_N_STATES; ++k)
                if (rec_mask & (1u << k)) __stcs(out + ch[k] * cs, (k < NS) ? x[k < NS ? k : 0] : 0.0);
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

template <int CM, int PMODE, int SOLVER>
static cudaError_t launch_wide_one(const WideParams& p, cudaStream_t stream, int* regs, int* threads_out, int* spb_out) {
    auto kern = dps_step_wide_kernel<CM, PMODE, SOLVER>;
    const int G = p.n_gen;
    const int WG = (G + 31) / 32;
    const int SPB = WIDE_THREADS / (WG * 32);
    if (threads_out) *threads_out = WIDE_THREADS;
    if (spb_out) *spb_out = SPB;
    if (regs) {
        cudaFuncAttributes a;
        cudaError_t e = cudaFuncGetAttributes(&a, kern);
        if (e != cudaSuccess) return e;
        *regs = a.numRegs;
        return cudaSuccess;
    }
    const size_t smem = sizeof(double2) * ((size_t)G * G + 2 * SPB * WG * 32) + sizeof(double) * G * (DPS_N_SLOTS + 1);
    if (smem > 227 * 1024) return cudaErrorInvalidConfiguration;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const long long blocks = (p.P + SPB - 1) / SPB;
    kern<<<(unsigned)blocks, WIDE_THREADS, smem, stream>>>(p, WG, SPB);
    return cudaGetLastError();
}

cudaError_t launch_forward_wide(const WideParams& p, int cmask, int pmode, cudaStream_t stream, int* regs, int* threads,
                                int* sims_per_block) {
#define DPS_WCASE(C, M)                                                                                        \
    if (cmask == C && pmode == M) {                                                                            \
        if (p.solver == DPS_EULER) return launch_wide_one<C, M, DPS_EULER>(p, stream, regs, threads, sims_per_block); \
        if (p.solver == DPS_HEUN) return launch_wide_one<C, M, DPS_HEUN>(p, stream, regs, threads, sims_per_block);   \
        return launch_wide_one<C, M, DPS_RK4>(p, stream, regs, threads, sims_per_block);                       \
    }
    DPS_WCASE(0, 1) DPS_WCASE(7, 1) DPS_WCASE(7, 3) DPS_WCASE(-1, 1)
#undef DPS_WCASE
    return cudaErrorInvalidValue;
}

}  // namespace dps
