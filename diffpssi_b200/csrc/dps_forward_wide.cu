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
//
// MMA variant (MMA = true, used whenever its shared-memory footprint fits): the network product of a whole block is ONE
// real matrix product on the FP64 MMA instruction (mma.sync.m8n8k4.f64):
//     [V_re | V_im] (G x 2S)  =  Z_re (G x G) [I_re | I_im]  +  Z_im (G x G) [-I_im | I_re]         S = trajectories per block
// so the 8 columns of an MMA tile are the real and imaginary parts of FOUR trajectories and every tile is full.  Z_re and
// Z_im sit in shared memory pre-swizzled into A-fragment order (one conflict-free LDS.64 per lane and MMA), the currents
// are written by their owner threads as rows [column n][generator k] (B-fragment reads are conflict-free because the
// row stride is 4 mod 8 doubles), warps own (8-generator x 4-trajectory) output tiles and write V back through shared
// memory.  Per (thread, stage) this replaces 2 G LDS.128 + 4 G DFMA by G/2 MMAs + G LDS.64 per WARP.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include "dps_internal.h"
#include "dps_model.cuh"

namespace dps {

constexpr int WIDE_THREADS = 512;

__device__ __forceinline__ void dmma_m8n8k4(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// shared-memory geometry of the MMA variant
struct WideMmaGeom {
    int MT, KT, NT, KP, VS;          // 8-row tiles of Z, 4-column tiles of Z, 8-column tiles of I, padded row of s_B, row of s_V
    __host__ __device__ WideMmaGeom(int G, int SPB) {
        MT = (G + 7) / 8; KT = (G + 3) / 4; NT = (SPB + 3) / 4;
        KP = KT * 4 + (((KT * 4) % 8 == 0) ? 4 : 0);         // == 4 (mod 8): conflict-free B-fragment loads
        VS = NT * 8 + 1;
    }
    // split = true (PMODE 3): two voltage buffers (one per half of the k range), no constant table in shared memory
    __host__ __device__ size_t bytes(int G, bool split) const {
        return sizeof(double) * ((size_t)2 * MT * KT * 32 + (size_t)2 * NT * 8 * KP + (size_t)(split ? 2 : 1) * MT * 8 * VS +
                                 (split ? 0 : (size_t)G * (DPS_N_SLOTS + 1)));
    }
};

// PMODE 2 reader of the wide stepper: any controller constant may differ per trajectory.  The shared-memory budget is spent
// on Z, so there is no room for a second table saying which slots vary: a varying slot's entry of the uniform table holds
// a NaN whose payload is the row of vpar ("NaN boxing"; written once at block start from WideParams::slot_vary).
constexpr int WIDE_BOX_HI = 0x7FFDEAD0;
struct ParamWideGen {
    const double* pm;        // registers, [DPS_N_REG_SLOTS]
    const double* up;        // shared memory row of this generator: value, or boxed vpar row
    const double* vpar;
    long long P, s;
    __device__ __forceinline__ double operator()(int slot) const {
        if (slot < DPS_N_REG_SLOTS) return pm[slot];
        const double u = up[slot];
        if (__double2hiint(u) == WIDE_BOX_HI) return __ldg(vpar + (long long)__double2loint(u) * P + s);
        return u;
    }
};

template <int CM, int PMODE, int SOLVER, bool MMA>
__global__ void __launch_bounds__(WIDE_THREADS, 1)
dps_step_wide_kernel(const WideParams p, int WG, int SPB) {
    constexpr bool CTRL = (CM != 0);
    constexpr int NS = CTRL ? DPS_N_STATES : 6;
    constexpr int UPS = DPS_N_SLOTS + 1;
    const int G = p.n_gen;
    const int GP = WG * 32;                                                       // padded generator count
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // DFMA variant
    double2* s_z = reinterpret_cast<double2*>(smem_raw);                          // [G][G] active configuration, transposed
    double2* s_cur = s_z + G * G;                                                 // [2][SPB][GP]
    double* s_upar = reinterpret_cast<double*>(s_cur + 2 * SPB * GP);             // [G][UPS]
    // MMA variant
    const WideMmaGeom mg(G, SPB);
    double* s_ar = reinterpret_cast<double*>(smem_raw);                           // [MT][KT][32] A fragments of Z_re
    double* s_ai = s_ar + mg.MT * mg.KT * 32;                                     // ... of Z_im
    double* s_b = s_ai + mg.MT * mg.KT * 32;                                      // [2][NT*8][KP] currents, row n = column of B
    double* s_v = s_b + 2 * mg.NT * 8 * mg.KP;                                    // [MT*8][VS] voltages
    // PMODE 3 needs no constant table in shared memory (the controller constants are constant-bank operands, the machine
    // constants are read once into registers): the room goes to a second voltage buffer, so that an output tile can be
    // split over two warps (halves of the k range) without atomics
    // Two experiments of round 2, kept behind switches because they document what does NOT bound this kernel (config 5,
    // tools/gpu_wide_ab.sh, profiles/round2/wide_ab.txt): balancing the 13 output tiles over the four schedulers by splitting
    // every tile's k range over two warps (DPS_WIDE_SPLIT: 402.7 -> 429.5 ms: the second voltage buffer, its reads and the
    // shorter MMA chains cost more than the imbalance) and applying the sign of -I_im to the finished sums instead of to
    // every B fragment (DPS_WIDE_SIGN_ON_SUM: 75 integer instructions fewer per tile, 404.2 ms: no effect).
#ifndef DPS_WIDE_SPLIT
#define DPS_WIDE_SPLIT 0
#endif
#ifndef DPS_WIDE_SIGN_ON_SUM
#define DPS_WIDE_SIGN_ON_SUM 0
#endif
#ifndef DPS_WIDE_UNROLL
#define DPS_WIDE_UNROLL 5
#endif
    constexpr bool SPLIT = MMA && (PMODE == 3) && (DPS_WIDE_SPLIT != 0);
    constexpr int WIDE_UNROLL = DPS_WIDE_UNROLL;
    if (MMA) s_upar = s_v + (SPLIT ? 2 : 1) * mg.MT * 8 * mg.VS;

    const int t = threadIdx.x;
    const int sl = t / GP;                          // trajectory slot in the block
    const int gl = t - sl * GP;                     // generator slot
    const bool gen_live = (sl < SPB) && (gl < G);
    const int g = (gl < G) ? gl : G - 1;
    const long long P = p.P;
    const long long sim = (long long)blockIdx.x * SPB + sl;
    const bool live = gen_live && (sim < P);
    const long long s = (sim < P) ? sim : (P - 1);

    if (!SPLIT) {
        for (int i = t; i < G * DPS_N_SLOTS; i += blockDim.x) {
            const int gi = i / DPS_N_SLOTS, k = i - gi * DPS_N_SLOTS;
            const int vi = (PMODE == 2) ? p.slot_vary[gi][k] : -1;
            s_upar[gi * UPS + k] = (vi >= 0) ? __hiloint2double(WIDE_BOX_HI, vi) : p.upar[i];
        }
    }
    auto load_z = [&](int cfg) {
        if (MMA) {
            // A-fragment order: lane l of tile (mt, kt) holds Z[mt*8 + l/4][kt*4 + l%4]
            for (int i = t; i < mg.MT * mg.KT * 32; i += blockDim.x) {
                const int l = i & 31, tile = i >> 5;
                const int kt = tile % mg.KT, mt = tile / mg.KT;
                const int row = mt * 8 + (l >> 2), col = kt * 4 + (l & 3);
                const bool ok = row < G && col < G;
                const double* e = p.zsub + (((long long)cfg * p.z_rows + (ok ? row : 0)) * p.z_cols + (ok ? col : 0)) * 2;
                s_ar[i] = ok ? e[0] : 0.0;
                s_ai[i] = ok ? e[1] : 0.0;
            }
            // padding columns of the current rows stay zero for the whole run
            for (int i = t; i < 2 * mg.NT * 8 * mg.KP; i += blockDim.x) s_b[i] = 0.0;

        } else {
            // transposed copy of configuration cfg: s_z[j*G + g] = Z[g][j]
            for (int i = t; i < G * G; i += blockDim.x) {
                const int row = i % G, col = i / G;
                const double* e = p.zsub + (((long long)cfg * p.z_rows + row) * p.z_cols + col) * 2;
                s_z[i] = make_double2(e[0], e[1]);
            }
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
        pm[k] = (vi >= 0) ? __ldg(p.vpar + (long long)vi * P + s) : (SPLIT ? __ldg(p.upar + g * DPS_N_SLOTS + k) : up[k]);
    }
    ParamRegs par_fast{pm, up};
    ParamConst par_const{pm, p.cpar};
    ParamWideGen par_gen{pm, up, p.vpar, P, s};

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
        else if (PMODE == 2) injection(X, par_gen, sd, cd, i_re, i_im);
        else injection(X, par_fast, sd, cd, i_re, i_im);
        double2* cb = s_cur + (buf * SPB + ((sl < SPB) ? sl : SPB - 1)) * GP;
        double v_re, v_im;
        // MMA variant: rows of this trajectory in the current / voltage tiles (real part, imaginary part 4 rows below)
        const int slc = (sl < SPB) ? sl : SPB - 1;
        const int nrow = (slc >> 2) * 8 + (slc & 3);
        double* bcur = s_b + buf * (mg.NT * 8 * mg.KP);
        if (MMA) {
            if (gen_live) {
                bcur[nrow * mg.KP + gl] = i_re;
                bcur[(nrow + 4) * mg.KP + gl] = i_im;
            }
            __syncthreads();
            const int lane = t & 31, warp = t >> 5;
            const int fr = lane >> 2, fk = lane & 3;                     // fragment row (A) / column (B), and k index
            // SPLIT: work units = (output tile, half of the k range): 2 MT NT units over the 16 warps, so that the four
            // schedulers get (almost) equal shares -- 13 whole tiles gave scheduler 0 four tiles and the others three.
            // Each half writes its own voltage buffer; the reader adds the two.
            const int n_units = (SPLIT ? 2 : 1) * mg.MT * mg.NT;
            const int KH = SPLIT ? (mg.KT + 1) / 2 : mg.KT;
            for (int u = warp; u < n_units; u += WIDE_THREADS / 32) {
                const int half = SPLIT ? (u & 1) : 0, tile = SPLIT ? (u >> 1) : u;
                const int mt = tile % mg.MT, nt = tile / mg.MT;
                const int k0 = half ? KH : 0, k1 = half ? mg.KT : KH;
                const double* ar = s_ar + (mt * mg.KT) * 32 + lane;
                const double* ai = s_ai + (mt * mg.KT) * 32 + lane;
                const double* b1 = bcur + (nt * 8 + fr) * mg.KP + fk;              // [ I_re | I_im ]
                const double* b2 = bcur + (nt * 8 + ((fr + 4) & 7)) * mg.KP + fk;  // [ I_im | I_re ]: the sign of -I_im is applied
                double d0 = 0.0, d1 = 0.0, e0 = 0.0, e1 = 0.0;                     // to the finished sum (exact), not per element
#if !DPS_WIDE_SIGN_ON_SUM
                const int flip = (fr < 4) ? (int)0x80000000 : 0;
#endif
#pragma unroll WIDE_UNROLL
                for (int kt = k0; kt < k1; ++kt) {
                    const double a_r = ar[kt * 32], a_i = ai[kt * 32];
                    double x2 = b2[kt * 4];
#if !DPS_WIDE_SIGN_ON_SUM
                    x2 = __hiloint2double(__double2hiint(x2) ^ flip, __double2loint(x2));
#endif
                    dmma_m8n8k4(d0, d1, a_r, b1[kt * 4]);
                    dmma_m8n8k4(e0, e1, a_i, x2);
                }
                // columns 0..3 of the tile are V_re = Z_re I_re - Z_im I_im, columns 4..7 V_im = Z_re I_im + Z_im I_re
                double* vo = s_v + half * (mg.MT * 8 * mg.VS) + (mt * 8 + fr) * mg.VS + nt * 8 + 2 * fk;
#if DPS_WIDE_SIGN_ON_SUM
                vo[0] = (fk < 2) ? d0 - e0 : d0 + e0;
                vo[1] = (fk < 2) ? d1 - e1 : d1 + e1;
#else
                vo[0] = d0 + e0;
                vo[1] = d1 + e1;
#endif
            }
            __syncthreads();
            // (row g, not gl: the padding lanes of the last warp would read past the voltage buffers)
            v_re = s_v[g * mg.VS + nrow];
            v_im = s_v[g * mg.VS + nrow + 4];
            if (SPLIT) {
                v_re += s_v[mg.MT * 8 * mg.VS + g * mg.VS + nrow];
                v_im += s_v[mg.MT * 8 * mg.VS + g * mg.VS + nrow + 4];
            }
            buf ^= 1;
        } else {
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
        v_re = vra + vrb;
        v_im = via + vib;
        }
        if (PMODE == 3) stage_derivatives<CTRL>(X, par_const, ctrl, p.fn, sd, cd, v_re, v_im, f, alg);
        else if (PMODE == 2) stage_derivatives<CTRL>(X, par_gen, ctrl, p.fn, sd, cd, v_re, v_im, f, alg);
        else stage_derivatives<CTRL>(X, par_fast, ctrl, p.fn, sd, cd, v_re, v_im, f, alg);
        if (last_stage && (rec_mask >> DPS_N_STATES)) {
            const short* ch = p.rec_chan[g];
            double* out = p.rec + (long long)step * P + sim;
            const long long cs = (long long)p.n_steps * P;
            const double av[7] = {alg.p_e, alg.i_d, alg.i_q, alg.v_re, alg.v_im, alg.e_fd, alg.p_m};
#pragma unroll
            for (int k = 0; k < 7; ++k)
                if (rec_mask & (1u << (DPS_N_STATES + k))) dps_rec_store(out + ch[DPS_N_STATES + k] * cs, av[k]);
        }
        // recorded voltages of non-generator buses: row r of the reduced matrix is read from global memory
        if (last_stage && n_extra > 0 && live) {
            for (int r = G + g; r < p.n_obs; r += G) {
                const double* zbase = p.zsub + (((long long)cfg * p.z_rows + r) * p.z_cols) * 2;
                double br = 0.0, bi = 0.0;
                for (int j = 0; j < G; ++j) {
                    const double zr_ = __ldg(zbase + 2 * j), zi_ = __ldg(zbase + 2 * j + 1);
                    const double2 cj = MMA ? make_double2(bcur[nrow * mg.KP + j], bcur[(nrow + 4) * mg.KP + j]) : cb[j];
                    br += zr_ * cj.x - zi_ * cj.y;
                    bi += zr_ * cj.y + zi_ * cj.x;
                }
                const int e = r - G;
                if (p.obs_chan[e][0] >= 0) dps_rec_store(p.rec + ((long long)p.obs_chan[e][0] * p.n_steps + step) * P + sim, br);
                if (p.obs_chan[e][1] >= 0) dps_rec_store(p.rec + ((long long)p.obs_chan[e][1] * p.n_steps + step) * P + sim, bi);
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
            for (int k = 0; k < NS; ++k) x[k] = x[k] + dps_div6((acc[k] + f[k]) * dt);
        }
        if (rec_mask & 0x1FFFu) {
            const short* ch = p.rec_chan[g];
            double* out = p.rec + (long long)step * P + sim;
            const long long cs = (long long)p.n_steps * P;
#pragma unroll
            for (int k = 0; k < DPS_N_STATES; ++k)
                if (rec_mask & (1u << k)) dps_rec_store(out + ch[k] * cs, (k < NS) ? x[k < NS ? k : 0] : 0.0);
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

template <int CM, int PMODE, int SOLVER, bool MMA>
static cudaError_t launch_wide_kern(const WideParams& p, cudaStream_t stream, int* regs, int* threads_out, int* spb_out);

template <int CM, int PMODE, int SOLVER>
static cudaError_t launch_wide_one(const WideParams& p, cudaStream_t stream, int* regs, int* threads_out, int* spb_out) {
    // the MMA variant whenever its shared-memory footprint fits (G <= ~104 with all controller tables)
    const int G = p.n_gen;
    const int WG = (G + 31) / 32;
    const int SPB = WIDE_THREADS / (WG * 32);
    const bool mma = WideMmaGeom(G, SPB).bytes(G, PMODE == 3 && DPS_WIDE_SPLIT) <= 227 * 1024 && getenv("DPS_WIDE_NO_MMA") == nullptr;
    return mma ? launch_wide_kern<CM, PMODE, SOLVER, true>(p, stream, regs, threads_out, spb_out)
               : launch_wide_kern<CM, PMODE, SOLVER, false>(p, stream, regs, threads_out, spb_out);
}

template <int CM, int PMODE, int SOLVER, bool MMA>
static cudaError_t launch_wide_kern(const WideParams& p, cudaStream_t stream, int* regs, int* threads_out, int* spb_out) {
    auto kern = dps_step_wide_kernel<CM, PMODE, SOLVER, MMA>;
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
    const size_t smem = MMA ? WideMmaGeom(G, SPB).bytes(G, PMODE == 3 && DPS_WIDE_SPLIT)
                            : sizeof(double2) * ((size_t)G * G + 2 * SPB * WG * 32) + sizeof(double) * G * (DPS_N_SLOTS + 1);
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
    DPS_WCASE(0, 1) DPS_WCASE(7, 1) DPS_WCASE(7, 3) DPS_WCASE(-1, 1) DPS_WCASE(0, 2) DPS_WCASE(7, 2) DPS_WCASE(-1, 2)
#undef DPS_WCASE
    return cudaErrorInvalidValue;
}

}  // namespace dps
