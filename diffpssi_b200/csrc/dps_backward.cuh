// dps_backward.cuh -- the discrete-adjoint kernel (sm_100a), instantiated per generator count by dps_backward_inst.cu
// (one translation unit per G so the build parallelises); ABI entry points and dispatch: dps_backward.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdlib.h>
#include <string.h>
#include "dps_internal.h"
#include "dps_adjoint.cuh"
#include "dps_backward_params.h"

namespace dps {

// Thread mapping = the forward stepper's (dps_forward.cuh): one thread per (trajectory, generator), the G generators of a
// trajectory in G adjacent lanes of ONE warp (32/G trajectories per warp).  Currents (for V = Z I) and voltage
// cotangents (for dI = Z^T dV) are exchanged through the warp's private shared-memory line behind a __syncwarp: no block
// barrier, warps never wait for each other.  (Until r2p a warp held 32 trajectories of one generator and the G warps of a
// group met at a named barrier twice per stage: one block of 10 warps per SM in lockstep, 28 % FP64-pipe activity.)
// G = 0 selects the WIDE build (17..128 generators, run-time generator count): a trajectory occupies ceil(n_gen / 32) warps
// of a 256-thread block, currents and voltage cotangents are exchanged through block-wide shared-memory lines behind
// __syncthreads, and Z is read from global memory (L2-resident: 160 KB for 100 generators).  Same per-thread arithmetic
// (dps_adjoint.cuh), same checkpoint / scratch layout, so the forward wide stepper's checkpoints feed it directly.
#ifndef DPS_BWD_AOS
#define DPS_BWD_AOS 0              // segment scratch: 1 = one contiguous record per thread, 0 = row-major [row][trajectory][generator]
#endif
#ifndef DPS_BWD_L1PF
#define DPS_BWD_L1PF 1             // prefetch the next scratch block towards L1 while a stage is transposed
#endif
#ifndef DPS_BWD_ASYNC
#define DPS_BWD_ASYNC 1            // sweep: the next stage record is copied global -> shared asynchronously (cp.async) while a stage is transposed
#endif
#ifndef DPS_BWD_WARPS
#define DPS_BWD_WARPS 4            // warps per block of the (<= 16 generator) adjoint
#endif
#ifndef DPS_BWD_MIN_BLOCKS
#define DPS_BWD_MIN_BLOCKS 2       // resident 128-thread blocks per SM the register allocation aims at (A/B: tools/build_variant.sh)
#endif

__device__ __forceinline__ void dps_cp_async8(double* smem, const double* gmem) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void dps_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void dps_cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int G>
struct BwdCfg {
    static constexpr bool WIDE = (G == 0);
    static constexpr int SPW = WIDE ? 0 : 32 / (G ? G : 1);       // trajectories per warp
    static constexpr int WARPS = WIDE ? 8 : DPS_BWD_WARPS;
    static constexpr int THREADS = 32 * WARPS;
};


// Gradient accumulators of the constant slots.  N = how many leading slots are kept (0: no slot gradient is wanted, the
// whole parameter-gradient arithmetic of stage_vjp is dead code; DPS_N_REG_SLOTS: only machine constants / set points --
// the usual identification leaves H, D, X_d', ...; DPS_N_SLOTS: everything).  Indices are literals at every use, so
// writes beyond N fold into a dead dummy.
template <int N>
struct SlotGrads {
    double v[N > 0 ? N : 1];
    double dummy;
    __device__ __forceinline__ double& operator[](int k) { return (k < N) ? v[k < N ? k : 0] : dummy; }
};

// CM: controller set of the generators -- 0 none (bare machines), 7 all three on every generator (compile time: the
// has_avr / has_gov / has_pss tests of the stage code fold away and the transposed stage is one straight-line block),
// -1 mixed (read per generator).  PM: how constants are read -- 0 any slot may differ per trajectory (vary table), 1 the
// DPS_N_REG_SLOTS leading slots in registers + uniform controller constants from shared memory.
// INREG (ckpt_stride 1 only: the forward pass kept the state before EVERY step): nothing is re-integrated and there is no
// scratch; the sweep evaluates the stages of a step in registers from the checkpoint.  Wins for small systems (bare
// machines: a stage evaluation is cheaper than the round trip of its record through L2).
template <class PT, int G, int CM, bool ZGRAD, int SOLVER, int NGP, int PM, bool QUEUED = false, bool INREG = false>
#ifdef DPS_BWD_MAXNREG
__global__ void __maxnreg__(DPS_BWD_MAXNREG)
#else
__global__ void __launch_bounds__(BwdCfg<G>::THREADS, BwdCfg<G>::WIDE ? 1 : DPS_BWD_MIN_BLOCKS)
#endif
dps_adjoint_kernel(const PT p) {
    using Cfg = BwdCfg<G>;
    constexpr bool CTRL = (CM != 0);
    constexpr bool FASTPAR = (PM != 0);
    constexpr bool WIDE = Cfg::WIDE;
    constexpr int SPW = Cfg::SPW;
    const int GG = WIDE ? p.n_gen : G;                            // generators (table strides)
    const int GP = WIDE ? ((p.n_gen + 31) & ~31) : G;             // lanes a trajectory occupies
    const int SPB = WIDE ? Cfg::THREADS / GP : Cfg::WARPS * SPW;  // trajectories per block
    // live states per generator: a bare machine (no controllers anywhere) carries 6, its 7 controller rows are zero.  The
    // per-step scratch, the register arrays and the sweep only hold NS of them; NSG is the stride of the 13-row buffers.
    constexpr int NS = CTRL ? DPS_N_STATES : 6;
    constexpr int NSG = DPS_N_STATES;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double2* s_xch = reinterpret_cast<double2*>(smem_raw);                       // [WARPS][2][32] exchange lines (wide: [2][THREADS])
    double* s_upar = reinterpret_cast<double*>(s_xch + Cfg::WARPS * 64);         // [G][DPS_N_SLOTS]
    // network shared by the batch: Z of every configuration in shared memory, (re, im) planes, in both orientations --
    // s_zt[cfg][j][g] = Z[g][j] for V = Z I (lanes of a trajectory read consecutive doubles) and s_zn[cfg][r][g] = Z[r][g]
    // for dI = Z^T dV -- so the two products are LDS with compile-time offsets instead of pointer-chasing global loads
    const int n_cfg_s = (WIDE || p.z_per_sim) ? 0 : p.n_cfg;
    double* s_ztr = s_upar + GG * DPS_N_SLOTS;
    double* s_zti = s_ztr + n_cfg_s * G * G;
    double* s_znr = s_zti + n_cfg_s * G * G;
    double* s_zni = s_znr + n_cfg_s * G * G;
    short* s_vary = reinterpret_cast<short*>(s_zni + n_cfg_s * G * G);           // [G][DPS_N_SLOTS]
    // staging area of the sweep: [2][SST][THREADS] doubles, one stage record per buffer, a column per thread (see below)
    double* s_stage = reinterpret_cast<double*>(smem_raw + p.stage_offset) + threadIdx.x;
    // WIDE: the generator block of the ACTIVE fault configuration in shared memory when it fits (re / im planes, row
    // stride LDZ odd: a warp reads a row segment (V = Z I, lanes = rows) or a column segment (dI = Z^T dV, lanes = columns)
    // without bank conflicts from the one copy); reloaded when the configuration changes.  It takes the staging area's
    // place.  Until round 2b both products read Z from L2 (the 100-generator sweep was ~10x the forward pass per step).
    const bool wide_z = WIDE && p.wide_z_smem;
    const int LDZ = GG | 1;
    double* s_zw_re = reinterpret_cast<double*>(smem_raw + p.stage_offset);
    double* s_zw_im = s_zw_re + GG * LDZ;
    int z_loaded = -1;

    const int lane = WIDE ? (int)threadIdx.x : (int)(threadIdx.x & 31);          // index into the exchange line
    const int warp = threadIdx.x >> 5;
    const int sl = lane / GP;                      // trajectory slot inside the warp (wide: inside the block)
    const int gl = lane - sl * GP;                 // generator slot
    const int n_slots = WIDE ? SPB : SPW;
    const int base = ((sl < n_slots) ? sl : n_slots - 1) * GP;   // first lane of this trajectory's generator group
    const bool gen_live = (sl < n_slots) && (gl < p.n_gen);
    const int g = (gl < p.n_gen) ? gl : p.n_gen - 1;
    const long long P = p.P;
    // work item (see dps_step_kernel): block group + chunk of checkpoint segments, taken by ticket
    __shared__ int s_item;
    long long group = blockIdx.x;
    int chunk = 0;
    if (QUEUED) {
        if (threadIdx.x == 0) s_item = atomicAdd(p.q_counter, 1);
        __syncthreads();
        const int item = __reduce_max_sync(0xffffffffu, s_item);     // REDUX: provably warp-uniform
        chunk = item / p.q_groups;
        group = item - chunk * p.q_groups;
        if (chunk > 0 && threadIdx.x == 0)
            while (dps_ld_acquire(p.q_progress + group) < chunk) __nanosleep(100);
        __syncthreads();
    }
    const long long grp = QUEUED ? group : (long long)blockIdx.x;
    const long long sim = WIDE ? grp * SPB + sl : (grp * Cfg::WARPS + warp) * SPW + sl;
    const bool live = (sim < P) && gen_live;
    const long long s = (sim < P) ? sim : (P - 1);

    for (int i = threadIdx.x; i < GG * DPS_N_SLOTS; i += blockDim.x) {
        const bool ok = (i / DPS_N_SLOTS) < p.n_gen;
        s_upar[i] = ok ? p.upar[i] : 0.0;
        s_vary[i] = ok ? p.slot_vary[i] : (short)-1;
    }
    // exchange lines start at zero; only live generator lanes ever store (padding lanes contribute exact zeros)
    for (int i = threadIdx.x; i < Cfg::WARPS * 64; i += blockDim.x) s_xch[i] = make_double2(0.0, 0.0);
    constexpr int GD = WIDE ? 1 : G;               // (the wide build has no shared Z tables: n_cfg_s == 0)
    for (int i = threadIdx.x; i < n_cfg_s * G * G; i += blockDim.x) {
        const int a = i % GD, b = (i / GD) % GD, cfg = i / (GD * GD);           // s_zt[cfg][b][a] = Z[a][b];  s_zn[cfg][b][a] = Z[b][a]
        const bool ok = a < p.n_gen && b < p.n_gen;
        const double* et = p.zsub + (((long long)cfg * p.z_rows + (ok ? a : 0)) * p.z_cols + (ok ? b : 0)) * 2;
        const double* en = p.zsub + (((long long)cfg * p.z_rows + (ok ? b : 0)) * p.z_cols + (ok ? a : 0)) * 2;
        s_ztr[i] = ok ? et[0] : 0.0;
        s_zti[i] = ok ? et[1] : 0.0;
        s_znr[i] = ok ? en[0] : 0.0;
        s_zni[i] = ok ? en[1] : 0.0;
    }
    __syncthreads();

    // Constants.  FASTPAR (plan pmode 1/3: only the DPS_N_REG_SLOTS leading slots may differ per trajectory): those are
    // resolved once into registers, the controller constants are warp-uniform shared-memory reads.  Otherwise every read
    // goes through the per-slot vary table (two loads and a branch per use).
    ParamReader<true> par_gen;
    par_gen.up = s_upar + g * DPS_N_SLOTS;
    par_gen.sv = s_vary + g * DPS_N_SLOTS;
    par_gen.vpar = p.vpar;
    par_gen.P = P;
    par_gen.s = s;
    double pm[FASTPAR ? DPS_N_REG_SLOTS : 1];
    if (FASTPAR) {
#pragma unroll
        for (int k = 0; k < DPS_N_REG_SLOTS; ++k) pm[FASTPAR ? k : 0] = par_gen(k);
    }
    const ParamRegs par_fast{pm, s_upar + g * DPS_N_SLOTS};
    auto pick_reader = [&]() -> decltype(auto) {
        if constexpr (PM == 1) return (par_fast);
        else return (par_gen);
    };
    const auto& par = pick_reader();
    const int ctrl = (CM > 0) ? CM : (CM < 0 ? p.ctrl[g] : 0);
    const long long zb = p.z_per_sim ? P : 1;
    const long long zs = p.z_per_sim ? s : 0;
    double2* xch_w = WIDE ? s_xch : s_xch + warp * 64;
    int buf = 0;

    // exchange (a, b) with the other generators of this trajectory; afterwards xb[j] holds generator j's pair
    auto exchange = [&](double a, double b) -> const double2* {
        double2* cw = xch_w + buf * (WIDE ? Cfg::THREADS : 32);
        if (gen_live) cw[lane] = make_double2(a, b);
        if (WIDE) __syncthreads(); else __syncwarp();
        buf ^= 1;
        return cw + base;
    };
    // Z[cfg][row][col] = (zsub[..][0][b], zsub[..][1][b]).  Row g (V = Z I) and column g (dI = Z^T dV) of the active
    // configuration are walked with a running pointer: one 64-bit add per element instead of the full index product.
    const long long z_col_stride = 2 * zb;                              // (row, col) -> (row, col+1)
    const long long z_row_stride = (long long)p.z_cols * 2 * zb;        // (row, col) -> (row+1, col)
    auto z_cfg = [&](int cfg) { return p.zsub + ((long long)cfg * p.z_rows * p.z_cols * 2) * zb + zs; };
    auto ensure_z = [&](int cfg) {                    // (block-uniform: every thread of a block walks the same steps)
        if (WIDE && wide_z && cfg != z_loaded) {
            __syncthreads();
            const double* zc = p.zsub + (long long)cfg * p.z_rows * p.z_cols * 2;
            for (int i = threadIdx.x; i < GG * GG; i += blockDim.x) {
                const int r = i / GG, c = i - r * GG;
                s_zw_re[r * LDZ + c] = zc[((long long)r * p.z_cols + c) * 2];
                s_zw_im[r * LDZ + c] = zc[((long long)r * p.z_cols + c) * 2 + 1];
            }
            __syncthreads();
            z_loaded = cfg;
        }
    };
    // one stage evaluation at X: sin/cos, network voltage, derivatives
    auto eval_stage = [&](const double* X, int cfg, double& sd, double& cd, double& v_re, double& v_im, double* f,
                          Algebraics& alg) {
        dps_sincos(X[DPS_X_DELTA], &sd, &cd);
        double i_re, i_im;
        injection(X, par, sd, cd, i_re, i_im);
        const double2* xb = exchange(i_re, i_im);
        v_re = 0.0;
        v_im = 0.0;
        if (!WIDE && !p.z_per_sim) {
            const double* zr_ = s_ztr + cfg * G * G + gl;
            const double* zi_ = s_zti + cfg * G * G + gl;
#pragma unroll
            for (int j = 0; j < G; ++j) {
                const double zr = zr_[j * G], zi = zi_[j * G];
                const double jr = xb[j].x, ji = xb[j].y;
                v_re += zr * jr - zi * ji;
                v_im += zr * ji + zi * jr;
            }
        } else if (WIDE && wide_z) {
            const double* zr_ = s_zw_re + g * LDZ;
            const double* zi_ = s_zw_im + g * LDZ;
#pragma unroll 4
            for (int j = 0; j < p.n_gen; ++j) {
                const double zr = zr_[j], zi = zi_[j];
                const double jr = xb[j].x, ji = xb[j].y;
                v_re += zr * jr - zi * ji;
                v_im += zr * ji + zi * jr;
            }
        } else {
            const double* zp = z_cfg(cfg) + g * z_row_stride;
            for (int j = 0; j < p.n_gen; ++j, zp += z_col_stride) {
                const double zr = __ldg(zp), zi = __ldg(zp + zb);
                const double jr = xb[j].x, ji = xb[j].y;
                v_re += zr * jr - zi * ji;
                v_im += zr * ji + zi * jr;
            }
        }
        stage_derivatives<CTRL>(X, par, ctrl, p.fn, sd, cd, v_re, v_im, f, alg);
    };
    // index of the fault-configuration segment that contains `step` (searched once per checkpoint segment; the step loops
    // below move it incrementally)
    auto seg_of = [&](int step) {
        int seg = 0;
        while (seg + 1 < p.n_seg && p.seg_step[seg + 1] <= step) ++seg;
        return seg;
    };

    // the scheme is a template parameter: the stage loops unroll and X[i], sds[i], ... are registers, not a local array
    constexpr int n_stages = (SOLVER == DPS_EULER) ? 1 : (SOLVER == DPS_HEUN ? 2 : 4);
    // segment scratch: per (step, trajectory, generator) one record of n_stages blocks of SST doubles -- the stage input
    // (NS states), sin/cos of the rotor angle, the bus voltage, padded to an even count for 128-bit accesses
    constexpr int SST = (NS + 5) & ~1;
    constexpr int SROW = n_stages * SST;
    const double dt = p.dt;
    // c_i, b_i of the explicit scheme (Euler b=[1]; Heun c=[0,1], b=[1/2,1/2]; RK4 c=[0,1/2,1/2,1], b=[1/6,1/3,1/3,1/6])
    constexpr double cst[4] = {0.0, (SOLVER == DPS_HEUN) ? 1.0 : 0.5, 0.5, 1.0};
    constexpr double bst[4] = {(SOLVER == DPS_EULER) ? 1.0 : (SOLVER == DPS_HEUN ? 0.5 : 1.0 / 6.0),
                               (SOLVER == DPS_HEUN) ? 0.5 : 1.0 / 3.0, 1.0 / 3.0, 1.0 / 6.0};

    double lam[NS];
    SlotGrads<NGP> gp;
    const short* wr = p.want_row + g * DPS_N_SLOTS;
    if (!QUEUED || chunk == 0) {
#pragma unroll
        for (int k = 0; k < NS; ++k)
            lam[k] = (p.grad_xf != nullptr) ? p.grad_xf[((long long)g * NSG + k) * P + s] : 0.0;
#pragma unroll
        for (int k = 0; k < NGP; ++k) gp.v[k] = 0.0;
    } else {
        // continue where the previous chunk of this group stopped (written by another SM: read through L2)
#pragma unroll
        for (int k = 0; k < NS; ++k) lam[k] = __ldcg(p.g_x0 + ((long long)g * NSG + k) * P + s);
#pragma unroll
        for (int k = 0; k < NGP; ++k) gp.v[k] = (wr[k] >= 0) ? __ldcg(p.gslot + (long long)wr[k] * P + s) : 0.0;
    }
    gp.dummy = 0.0;

    const unsigned rec_mask = p.rec_mask[g];
    const bool rec_uniform = p.rec_uniform != 0;                 // every generator records the same set of states
    const unsigned rec_mask_u = p.rec_mask[0] & 0x1FFFu;
    const short* ch = p.rec_chan[g];
    const long long cs = (long long)p.n_steps * P;
    const int C = p.ckpt_stride;

    const int ck_hi = QUEUED ? p.n_ckpt - chunk * p.q_chunk_ckpts : p.n_ckpt;            // exclusive
    const int ck_lo = QUEUED ? ((ck_hi - p.q_chunk_ckpts > 0) ? ck_hi - p.q_chunk_ckpts : 0) : 0;
    for (int ck = ck_hi - 1; ck >= ck_lo; --ck) {
        const int k0 = ck * C;
        const int k1 = (k0 + C < p.n_steps) ? (k0 + C) : p.n_steps;
        // ---- re-integrate the segment ---------------------------------------------------------------------------------
        // Everything the transposed stages need -- the stage inputs X_i, sin/cos of the rotor angle and the bus voltage of
        // every stage -- goes to the scratch here, so that the sweep below evaluates no stage a second time (until round 2
        // it kept only the step's start state and re-evaluated every stage incl. the network product).  Scratch layout
        // [step][trajectory][generator][SROW]: a thread's record is contiguous (128-bit stores / loads at immediate
        // offsets, one running pointer), a warp's records are adjacent.
        const long long sstep = P * (long long)p.n_gen * SROW;                           // one step further
#if DPS_BWD_AOS
        double* const sbase = p.scratch + (s * p.n_gen + g) * SROW;
#else
        const long long srow_stride = P * (long long)p.n_gen;                            // one row further
        double* const sbase = p.scratch + (s * p.n_gen + g);
#endif
        int seg = seg_of(k0);
        int cfg = p.seg_cfg[seg];
        if constexpr (!INREG) {
            double xs[NS], xt[NS], acc[NS], f[NSG];
            const double* src = p.ckpt + (((long long)ck * p.n_gen + g) * NSG) * P + s;
#pragma unroll
            for (int k = 0; k < NS; ++k) xs[k] = src[(long long)k * P];
            int next_change = (seg + 1 < p.n_seg) ? p.seg_step[seg + 1] : 0x7fffffff;
            double* dst = sbase;
            for (int step = k0; step < k1; ++step, dst += sstep) {
                if (step == next_change) {
                    ++seg;
                    cfg = p.seg_cfg[seg];
                    next_change = (seg + 1 < p.n_seg) ? p.seg_step[seg + 1] : 0x7fffffff;
                }
                ensure_z(cfg);
#pragma unroll
                for (int st = 0; st < n_stages; ++st) {
                    const double* X = (st == 0) ? xs : xt;
                    double sd, cd, v_re, v_im;
                    Algebraics alg;
                    eval_stage(X, cfg, sd, cd, v_re, v_im, f, alg);
                    if (live) {
                        double rec[SST];
#pragma unroll
                        for (int k = 0; k < NS; ++k) rec[k] = X[k];
                        rec[NS] = sd; rec[NS + 1] = cd; rec[NS + 2] = v_re; rec[NS + 3] = v_im;
                        if (SST > NS + 4) rec[SST - 1] = 0.0;
#if DPS_BWD_AOS
                        double2* d2 = reinterpret_cast<double2*>(dst + st * SST);
#pragma unroll
                        for (int k = 0; k < SST / 2; ++k) d2[k] = make_double2(rec[2 * k], rec[2 * k + 1]);
#else
                        double* d = dst + (long long)(st * SST) * srow_stride;
#pragma unroll
                        for (int k = 0; k < NS + 4; ++k) d[k * srow_stride] = rec[k];
#endif
                    }
                    if (SOLVER == DPS_EULER) {
#pragma unroll
                        for (int k = 0; k < NS; ++k) xs[k] = xs[k] + f[k] * dt;
                    } else if (SOLVER == DPS_HEUN) {
                        if (st == 0) {
#pragma unroll
                            for (int k = 0; k < NS; ++k) { acc[k] = f[k]; xt[k] = xs[k] + f[k] * dt; }
                        } else {
#pragma unroll
                            for (int k = 0; k < NS; ++k) xs[k] = xs[k] + ((acc[k] + f[k]) * 0.5) * dt;
                        }
                    } else {
                        if (st == 0) {
#pragma unroll
                            for (int k = 0; k < NS; ++k) { acc[k] = f[k]; xt[k] = xs[k] + (f[k] * dt) * 0.5; }
                        } else if (st == 1) {
#pragma unroll
                            for (int k = 0; k < NS; ++k) { acc[k] = acc[k] + 2.0 * f[k]; xt[k] = xs[k] + (f[k] * dt) * 0.5; }
                        } else if (st == 2) {
#pragma unroll
                            for (int k = 0; k < NS; ++k) { acc[k] = acc[k] + 2.0 * f[k]; xt[k] = xs[k] + f[k] * dt; }
                        } else {
#pragma unroll
                            for (int k = 0; k < NS; ++k) xs[k] = xs[k] + dps_div6((acc[k] + f[k]) * dt);
                        }
                    }
                }
            }
        }
        // ---- backward sweep over the segment --------------------------------------------------------------------------
        // (seg / cfg are those of step k1 - 1, where the re-integration stopped)
        int seg_lo = p.seg_step[seg];
        const double* srec = sbase + (long long)(k1 - 1 - k0) * sstep;
#if DPS_BWD_ASYNC && !DPS_BWD_AOS
        // The stage records are staged through shared memory with cp.async, one record ahead: while a stage is transposed
        // the record the sweep needs next (the previous stage, or the last stage of the previous step) is already on its
        // way from L2, so its latency is off the dependency chain (it was 25 % of the kernel's stall samples).  Every
        // thread copies and reads only its own column: no barrier, cp.async.wait_group orders copy and read.
        constexpr bool STAGED = !INREG && !WIDE;
        int sbuf = 0;
        auto stage_fetch = [&](const double* rec_ptr, int b) {
            double* dst = s_stage + (long long)b * SST * Cfg::THREADS;
#pragma unroll
            for (int k = 0; k < NS + 4; ++k) dps_cp_async8(dst + k * Cfg::THREADS, rec_ptr + k * srow_stride);
            dps_cp_async_commit();
        };
        // ... and so are the cotangents of the recorded states, one step ahead (their load sat at the head of every step)
        double* const s_rec = s_stage + (long long)2 * SST * Cfg::THREADS;              // [2][NS][THREADS]
        int rbuf = 0;
        auto rec_fetch = [&](int st, int b) {
            const double* gr = p.grad_rec + (long long)st * P + s;
#pragma unroll
            for (int k = 0; k < NS; ++k)
                if (rec_mask_u & (1u << k)) dps_cp_async8(s_rec + (long long)(b * NS + k) * Cfg::THREADS, gr + ch[k] * cs);
        };
        if (STAGED) {
            if (rec_uniform) rec_fetch(k1 - 1, 0);
            stage_fetch(srec + (long long)((n_stages - 1) * SST) * srow_stride, 0);      // (commits both)
        }
#else
        constexpr bool STAGED = false;
#endif
        for (int step = k1 - 1; step >= k0; --step, srec -= sstep) {
            if (step < seg_lo) {
                --seg;
                cfg = p.seg_cfg[seg];
                seg_lo = p.seg_step[seg];
            }
            ensure_z(cfg);
            // recorded post-step states: rec[.., step] = x_{step+1}.  Every generator records the same fields in all the
            // usual cases: the field tests are then uniform branches (no divergence bookkeeping); dead lanes read a valid
            // (clamped) address and their co-state is never stored.
#if DPS_BWD_ASYNC && !DPS_BWD_AOS
            if (STAGED && rec_uniform) {
                dps_cp_async_wait<0>();                 // (the first stage of this step needs its record now anyway)
                const double* cur = s_rec + (long long)rbuf * NS * Cfg::THREADS;
#pragma unroll
                for (int k = 0; k < NS; ++k)
                    if (rec_mask_u & (1u << k)) lam[k] += cur[k * Cfg::THREADS];
                if (step > k0) rec_fetch(step - 1, rbuf ^ 1);     // committed with the next stage record
                rbuf ^= 1;
            } else
#endif
            if (rec_uniform) {
                const double* gr = p.grad_rec + (long long)step * P + s;
#pragma unroll
                for (int k = 0; k < NS; ++k)
                    if (rec_mask_u & (1u << k)) {
                        const double* e = gr + ch[k] * cs;
                        lam[k] += *e;
                        // the sweep walks the steps backwards: fetch the previous step's cotangent now, it is consumed
                        // a whole step later (this load sat on the critical path with full HBM latency)
                        if (step > 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(e - P));
                    }
            } else if (live && rec_mask) {
                const double* gr = p.grad_rec + (long long)step * P + sim;
#pragma unroll
                for (int k = 0; k < NS; ++k)
                    if (rec_mask & (1u << k)) {
                        const double* e = gr + ch[k] * cs;
                        lam[k] += *e;
                        if (step > 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(e - P));
                    }
            }
            // ckpt_stride 1: the stage inputs and their voltages are evaluated here, from the state the forward pass kept
            double XR[INREG ? 4 : 1][NS], sds[4], cds[4], vres[4], vims[4];
            if constexpr (INREG) {
                const double* src = p.ckpt + (((long long)step * p.n_gen + g) * NSG) * P + s;
#pragma unroll
                for (int k = 0; k < NS; ++k) XR[0][k] = __ldcs(src + (long long)k * P);      // read once: streaming
                // the state of the step the sweep visits next is fetched now and consumed a whole step later
                if (step > 0) {
                    const double* nxt = src - (long long)p.n_gen * NSG * P;
#pragma unroll
                    for (int k = 0; k < NS; ++k) asm volatile("prefetch.global.L2 [%0];" ::"l"(nxt + (long long)k * P));
                }
#pragma unroll
                for (int i = 0; i < n_stages; ++i) {
                    double f[NSG];
                    Algebraics alg;
                    eval_stage(XR[INREG ? i : 0], cfg, sds[i], cds[i], vres[i], vims[i], f, alg);
                    if (i + 1 < n_stages) {
                        const double ci = cst[i + 1];
#pragma unroll
                        for (int k = 0; k < NS; ++k) XR[INREG ? i + 1 : 0][k] = XR[0][k] + ci * (f[k] * dt);
                    }
                }
            }
            // transposed stages, last to first
            double sum_dx[NS], dxn[NS];
#pragma unroll
            for (int k = 0; k < NS; ++k) { sum_dx[k] = 0.0; dxn[k] = 0.0; }
#pragma unroll
            for (int i = n_stages - 1; i >= 0; --i) {
                double dk[NS], dx[NS];
                const double bi = bst[i] * dt;
                const double cn = (i + 1 < n_stages) ? cst[i + 1] * dt : 0.0;
#pragma unroll
                for (int k = 0; k < NS; ++k) { dk[k] = bi * lam[k] + cn * dxn[k]; dx[k] = 0.0; }
                AlgCot ac = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
                if (i == n_stages - 1 && live && (rec_mask >> NSG)) {
                    const double* gr = p.grad_rec + (long long)step * P + sim;
                    double* av[7] = {&ac.p_e, &ac.i_d, &ac.i_q, &ac.v_re, &ac.v_im, &ac.e_fd, &ac.p_m};
#pragma unroll
                    for (int k = 0; k < 7; ++k)
                        if (rec_mask & (1u << (NSG + k))) *av[k] = gr[ch[NSG + k] * cs];
                }
                // this stage's forward quantities, from the segment scratch; the block the sweep needs next (the previous
                // stage, or the last stage of the previous step) is pulled towards the SM meanwhile
                double Xi[NS], sdi, cdi, vri, vii;
                if constexpr (INREG) {
#pragma unroll
                    for (int k = 0; k < NS; ++k) Xi[k] = XR[INREG ? i : 0][k];
                    sdi = sds[i]; cdi = cds[i]; vri = vres[i]; vii = vims[i];
                } else {
                    double rec[SST];
#if DPS_BWD_ASYNC && !DPS_BWD_AOS
                    if (STAGED) {
                        if (i > 0 || step > k0) {
                            const double* nxt = (i > 0) ? srec + (long long)((i - 1) * SST) * srow_stride
                                                        : srec - sstep + (long long)((n_stages - 1) * SST) * srow_stride;
                            stage_fetch(nxt, sbuf ^ 1);
                            dps_cp_async_wait<1>();
                        } else {
                            dps_cp_async_wait<0>();
                        }
                        const double* cur = s_stage + (long long)sbuf * SST * Cfg::THREADS;
#pragma unroll
                        for (int k = 0; k < NS + 4; ++k) rec[k] = cur[k * Cfg::THREADS];
                        sbuf ^= 1;
                    } else
#endif
                    {
#if DPS_BWD_AOS
                        const double2* d2 = reinterpret_cast<const double2*>(srec + i * SST);
#pragma unroll
                        for (int k = 0; k < SST / 2; ++k) {
                            const double2 v = d2[k];
                            rec[2 * k] = v.x;
                            rec[2 * k + 1] = v.y;
                        }
#else
                        const double* d = srec + (long long)(i * SST) * srow_stride;
#pragma unroll
                        for (int k = 0; k < NS + 4; ++k) rec[k] = d[k * srow_stride];
#endif
                    }
#pragma unroll
                    for (int k = 0; k < NS; ++k) Xi[k] = rec[k];
                    sdi = rec[NS]; cdi = rec[NS + 1]; vri = rec[NS + 2]; vii = rec[NS + 3];
#if DPS_BWD_AOS && DPS_BWD_L1PF
                    const double* nxt = (i > 0) ? srec + (i - 1) * SST : srec - sstep + (n_stages - 1) * SST;
                    if (i > 0 || step > k0) {
                        asm volatile("prefetch.global.L1 [%0];" ::"l"(nxt));
                        asm volatile("prefetch.global.L1 [%0];" ::"l"(nxt + 16));
                    }
#endif
                }
                double dsd, dcd, dv_re, dv_im;
                stage_vjp<CTRL>(Xi, par, ctrl, p.fn, sdi, cdi, vri, vii, dk, ac, dx, gp, dsd, dcd, dv_re, dv_im);
                // dI = Z^T dV over the generator rows ...
                const double2* xb = exchange(dv_re, dv_im);
                double di_re = 0.0, di_im = 0.0;
                // ... and, when the network carries a gradient, dZ[r][g] += dV_r conj(I_g): this thread owns column g of
                // its trajectory's matrix, so the accumulation in HBM needs no atomics.  I_g is recomputed (6 flops).
                double own_re = 0.0, own_im = 0.0;
                double* gz = nullptr;
                if (ZGRAD) {
                    injection(Xi, par, sdi, cdi, own_re, own_im);
                    gz = p.g_zsub + ((((long long)cfg * p.z_rows) * p.z_cols + g) * 2) * P + s;
                }
                const long long gz_row = (long long)p.z_cols * 2 * P;
                if (!WIDE && !p.z_per_sim && !ZGRAD) {
                    const double* zr_ = s_znr + cfg * G * G + gl;
                    const double* zi_ = s_zni + cfg * G * G + gl;
#pragma unroll
                    for (int r = 0; r < G; ++r) {
                        const double zr = zr_[r * G], zi = zi_[r * G];
                        const double ar = xb[r].x, ai = xb[r].y;
                        di_re += zr * ar + zi * ai;
                        di_im += -zi * ar + zr * ai;
                    }
                } else if (WIDE && wide_z) {
                    const double* zr_ = s_zw_re + g;
                    const double* zi_ = s_zw_im + g;
#pragma unroll 4
                    for (int r = 0; r < p.n_gen; ++r) {
                        const double zr = zr_[r * LDZ], zi = zi_[r * LDZ];
                        const double ar = xb[r].x, ai = xb[r].y;
                        di_re += zr * ar + zi * ai;
                        di_im += -zi * ar + zr * ai;
                    }
                } else {
                    const double* zp = z_cfg(cfg) + g * z_col_stride;
                    for (int r = 0; r < p.n_gen; ++r, zp += z_row_stride) {
                        const double zr = __ldg(zp), zi = __ldg(zp + zb);
                        const double ar = xb[r].x, ai = xb[r].y;
                        di_re += zr * ar + zi * ai;
                        di_im += -zi * ar + zr * ai;
                        if (ZGRAD && live) {
                            gz[r * gz_row] += ar * own_re + ai * own_im;
                            gz[r * gz_row + P] += -ar * own_im + ai * own_re;
                        }
                    }
                }
                // ... plus the recorded voltages of extra observed buses (last stage of the step)
                if (i == n_stages - 1 && p.n_obs > p.n_gen && sim < P) {
                    for (int r = p.n_gen; r < p.n_obs; ++r) {
                        const int e = r - p.n_gen;
                        const double ar = (p.obs_chan[e][0] >= 0) ? p.grad_rec[((long long)p.obs_chan[e][0] * p.n_steps + step) * P + sim] : 0.0;
                        const double ai = (p.obs_chan[e][1] >= 0) ? p.grad_rec[((long long)p.obs_chan[e][1] * p.n_steps + step) * P + sim] : 0.0;
                        const double* ze = z_cfg(cfg) + r * z_row_stride + g * z_col_stride;
                        const double zr = __ldg(ze), zi = __ldg(ze + zb);
                        di_re += zr * ar + zi * ai;
                        di_im += -zi * ar + zr * ai;
                        if (ZGRAD && live) {
                            gz[r * gz_row] += ar * own_re + ai * own_im;
                            gz[r * gz_row + P] += -ar * own_im + ai * own_re;
                        }
                    }
                }
                injection_vjp(Xi, par, sdi, cdi, di_re, di_im, dx, gp, dsd, dcd);
                dx[DPS_X_DELTA] += dsd * cdi - dcd * sdi;
#pragma unroll
                for (int k = 0; k < NS; ++k) { dxn[k] = dx[k]; sum_dx[k] += dx[k]; }
            }
#pragma unroll
            for (int k = 0; k < NS; ++k) lam[k] += sum_dx[k];
        }
    }

    if (live) {
        double* dst = p.g_x0 + (long long)g * NSG * P + s;
#pragma unroll
        for (int k = 0; k < NS; ++k) dst[(long long)k * P] = lam[k];
        for (int k = NS; k < NSG; ++k) dst[(long long)k * P] = 0.0;
#pragma unroll
        for (int k = 0; k < NGP; ++k)
            if (wr[k] >= 0) p.gslot[(long long)wr[k] * P + s] = gp.v[k];
    }
    if (QUEUED) {
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) {
            const int item = s_item, c = item / p.q_groups;
            dps_st_release(p.q_progress + (item - c * p.q_groups), c + 1);
        }
    }
}

// Work-queue planning of the adjoint (see plan_work_queue in dps_forward_pair.cuh): chunks are whole checkpoint segments.
template <class PT>
static cudaError_t plan_adjoint_queue(PT& q, const void* kern, int threads, size_t smem, long long groups, unsigned* grid,
                                      cudaStream_t stream) {
    // default: one chunk per block group (a plain launch, blocks numbered by ticket)
    q.q_items = (int)groups;
    q.q_groups = (int)groups;
    q.q_chunk_ckpts = q.n_ckpt;
    *grid = (unsigned)groups;
    if (q.q_counter == nullptr || groups > q.q_capacity || groups > 0x3fffffff) return cudaErrorInvalidValue;
    const char* env = getenv("DPS_QUEUE_ROUNDS");
    const int rounds = env ? atoi(env) : 12;
    int dev = 0, sms = 0, per_sm = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem);
    if (e != cudaSuccess) return e;
    long long slots = (long long)sms * per_sm;
    if (const char* es = getenv("DPS_QUEUE_SLOTS")) slots = atoll(es);
    if (rounds > 0 && q.n_ckpt >= 4 && slots >= 1 && groups > slots) {
        long long n_chunks = ((long long)rounds * slots + groups - 1) / groups;
        int chunk_ckpts = (int)((q.n_ckpt + n_chunks - 1) / (n_chunks > 0 ? n_chunks : 1));
        if (chunk_ckpts < 1) chunk_ckpts = 1;
        n_chunks = (q.n_ckpt + chunk_ckpts - 1) / chunk_ckpts;
        if (n_chunks >= 2 && groups * n_chunks <= 0x3fffffff) {
            q.q_items = (int)(groups * n_chunks);
            q.q_chunk_ckpts = chunk_ckpts;
            *grid = (unsigned)q.q_items;
        }
    }
    return cudaMemsetAsync(q.q_counter, 0, sizeof(int) * (size_t)(groups + 1), stream);
}

template <class PT, int G, int CM, bool ZGRAD, int SOLVER, int NGP, int PM, bool INREG = false>
static cudaError_t launch_bwd_one(const PT& p, cudaStream_t stream) {
    using C = BwdCfg<G>;
    const int GG = C::WIDE ? p.n_gen : G;
    const size_t smem_tables = sizeof(double2) * C::WARPS * 64 + sizeof(double) * GG * DPS_N_SLOTS + sizeof(short) * GG * DPS_N_SLOTS +
                               ((C::WIDE || p.z_per_sim) ? 0 : sizeof(double) * 4 * (size_t)p.n_cfg * G * G);
    const long long spb = C::WIDE ? C::THREADS / ((p.n_gen + 31) & ~31) : (long long)C::WARPS * C::SPW;
    const long long blocks = (p.P + spb - 1) / spb;
    // staging area of the sweep (two stage records per thread, dps_adjoint_kernel), 16-byte aligned behind the tables
    constexpr int SST = ((CM != 0 ? DPS_N_STATES : 6) + 5) & ~1;
    const size_t stage_offset = (smem_tables + 15) & ~(size_t)15;
    size_t smem = stage_offset + ((DPS_BWD_ASYNC && !DPS_BWD_AOS && !INREG && !C::WIDE) ? sizeof(double) * 2 * (SST + SST - 4) * C::THREADS : 0);   // + [2][NS] cotangents
    int wide_z_smem = 0;
    if (C::WIDE && !p.z_per_sim && getenv("DPS_BWD_WIDE_Z_GLOBAL") == nullptr) {      // wide: Z of the active configuration instead
        const size_t zbytes = sizeof(double) * 2 * (size_t)GG * (GG | 1);
        if (stage_offset + zbytes <= 227 * 1024) { smem = stage_offset + zbytes; wide_z_smem = 1; }
    }
    PT pp = p;
    pp.stage_offset = (int)stage_offset;
    pp.wide_z_smem = wide_z_smem;
    // The fast-constant builds (no network gradient: those accumulate dL/dZ in HBM across chunks) only exist as work-queue
    // kernels; a launch that is not worth cutting into chunks runs as a queue of one chunk per block group.
    constexpr bool Q = !ZGRAD && (PM != 0);
    auto kern = dps_adjoint_kernel<PT, G, CM, ZGRAD, SOLVER, NGP, PM, Q, INREG>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    if constexpr (Q) {
        PT q = pp;
        unsigned grid = (unsigned)blocks;
        cudaError_t e = plan_adjoint_queue(q, (const void*)kern, C::THREADS, smem, blocks, &grid, stream);
        if (e != cudaSuccess) return e;
        kern<<<grid, C::THREADS, smem, stream>>>(q);
        return cudaGetLastError();
    } else {
        kern<<<(unsigned)blocks, C::THREADS, smem, stream>>>(pp);
        return cudaGetLastError();
    }
}

template <class PT, int G, int CM, bool ZGRAD, int SOLVER>
static cudaError_t launch_bwd_ngp(const PT& p, int ngp, cudaStream_t stream) {
    // Accumulators for the register slots only (machine constants / set points: the usual identification leaves) or for
    // every slot; fast constants whenever the plan allows it (pmode 1 / 3); the generic reader with all accumulators
    // covers everything else.  The network-gradient variant is only built generic (it is the rare one).  (Controller
    // constants as constant-bank operands, the forward stepper's PMODE 3, were measured here too: no gain -- the sweep is
    // bound by dependency latency, not by its shared-memory reads -- so the adjoint has no such build.)
    if constexpr (!ZGRAD) {
        if (ngp <= DPS_N_REG_SLOTS && p.pmode != 2) {
            if constexpr (G != 0) {
                if (p.ckpt_stride == 1) return launch_bwd_one<PT, G, CM, ZGRAD, SOLVER, DPS_N_REG_SLOTS, 1, true>(p, stream);
            }
            return launch_bwd_one<PT, G, CM, ZGRAD, SOLVER, DPS_N_REG_SLOTS, 1>(p, stream);
        }
        if (p.pmode != 2) return launch_bwd_one<PT, G, CM, ZGRAD, SOLVER, DPS_N_SLOTS, 1>(p, stream);
    }
    // generic: built for bare machines and for the run-time controller mask only
    if constexpr (CM == 7) return launch_bwd_one<PT, G, -1, ZGRAD, SOLVER, DPS_N_SLOTS, 0>(p, stream);
    else return launch_bwd_one<PT, G, CM, ZGRAD, SOLVER, DPS_N_SLOTS, 0>(p, stream);
}

template <class PT, int G, int CM, bool ZGRAD>
static cudaError_t launch_bwd_solver(const PT& p, int ngp, cudaStream_t stream) {
    if (p.solver == DPS_EULER) return launch_bwd_ngp<PT, G, CM, ZGRAD, DPS_EULER>(p, ngp, stream);
    if (p.solver == DPS_HEUN) return launch_bwd_ngp<PT, G, CM, ZGRAD, DPS_HEUN>(p, ngp, stream);
    return launch_bwd_ngp<PT, G, CM, ZGRAD, DPS_RK4>(p, ngp, stream);
}

// cmask: 0 = bare machines, 7 = every generator carries all three controllers, -1 = anything else
template <int G>
cudaError_t launch_bwd_g(const BwdParams& p, int cmask, int ngp, cudaStream_t stream) {
    if (p.g_zsub != nullptr)
        return cmask == 0 ? launch_bwd_solver<BwdParams, G, 0, true>(p, ngp, stream) : launch_bwd_solver<BwdParams, G, -1, true>(p, ngp, stream);
    if (cmask == 0) return launch_bwd_solver<BwdParams, G, 0, false>(p, ngp, stream);
    if (cmask == 7) return launch_bwd_solver<BwdParams, G, 7, false>(p, ngp, stream);
    return launch_bwd_solver<BwdParams, G, -1, false>(p, ngp, stream);
}

// 17..128 generators (G = 0: run-time generator count); only instantiated by dps_backward_wide.cu
#ifdef DPS_BWD_WIDE_TU
inline cudaError_t launch_bwd_wide_impl(const BwdWideParams& p, int cmask, int ngp, cudaStream_t stream) {
    if (p.g_zsub != nullptr || p.z_per_sim) return cudaErrorInvalidValue;
    if (cmask == 0) return launch_bwd_solver<BwdWideParams, 0, 0, false>(p, ngp, stream);
    if (cmask == 7) return launch_bwd_solver<BwdWideParams, 0, 7, false>(p, ngp, stream);
    return launch_bwd_solver<BwdWideParams, 0, -1, false>(p, ngp, stream);
}
#endif

}  // namespace dps
