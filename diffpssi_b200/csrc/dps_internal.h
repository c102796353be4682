// dps_internal.h -- structures shared between the ABI layer (dps_abi.cu) and the kernels.
#pragma once
#ifdef __CUDACC__
#include <cuda_runtime.h>
#endif
#include <stdint.h>
#include "../../include/dps.h"

namespace dps {

#ifndef DPS_FWD_WARPS
#define DPS_FWD_WARPS 4               // warps per block of the warp-exchange stepper (12 resident warps per SM)
#endif
constexpr int MAX_G = 16;            // generators handled by the warp-exchange stepper (dps_forward.cuh)
constexpr int MAX_GW = 128;          // generators handled by the wide stepper (dps_forward_wide.cuh)
constexpr int MAX_SEG = 48;          // network-configuration change points per launch
constexpr int MAX_EXTRA_OBS = 64;    // recorded non-generator buses (every bus of IEEE-39 can be recorded)

// Kernel parameter block (passed by value -> constant bank).
struct FwdParams {
    long long P;
    int n_gen, solver, n_steps, step_begin, step_end;
    double dt, fn;
    const double* x0;        // [n_gen][13][P] state at step_begin
    const double* upar;      // [n_gen][42]
    const double* vpar;      // [n_vary][P]
    const double* zsub;      // [n_cfg][z_rows][z_cols][2][Bz]
    double* rec;             // [n_rec][n_steps][P]
    double* x_final;         // [n_gen][13][P] or null
    double* ckpt;            // [n_ckpt][n_gen][13][P] or null
    int n_obs, z_rows, z_cols, n_cfg, z_per_sim, ckpt_stride, n_seg;
    int seg_step[MAX_SEG];
    int seg_cfg[MAX_SEG];
    int ctrl[MAX_G];
    unsigned rec_mask[MAX_G];               // bit k: field k of this generator is recorded
    short rec_chan[MAX_G][DPS_N_FIELDS];    // channel index of field k
    short obs_chan[MAX_EXTRA_OBS][2];       // channel of (re, im) of extra observed row e, or -1
    short slot_vary[MAX_G * DPS_N_SLOTS];   // row of vpar or -1
    double cpar[DPS_N_SLOTS];               // PMODE 3: controller constants common to all generators
    // "regular" recorder: every generator records the same fields (rec_umask) and field k of generator g is channel
    // rec_fbase[k] + g.  The stepper then needs no per-thread channel lookup and no divergent branch per store.
    int rec_regular;
    unsigned rec_umask;
    short rec_fbase[DPS_N_FIELDS];
    // Work queue (q_items > 0): the launch is a PERSISTENT grid whose blocks pull (block group, step chunk) items from an
    // atomic counter in chunk-major order instead of one block integrating one group for all steps.  With only a few
    // "waves" of blocks per SM a static launch leaves the SMs half empty for the whole last wave (8192 trajectories per GPU:
    // 1.5 waves); cut into chunks the same work spreads evenly.  The state travels between chunks through x_final
    // ([n_gen][13][P], the kernel's own state layout) and q_progress[group] counts the finished chunks of a group.
    int q_items, q_groups, q_chunk_steps;
    int q_capacity;          // block groups the caller's workspace has progress counters for
    int* q_counter;          // one int, zeroed before the launch
    int* q_progress;         // [q_groups], zeroed before the launch
};

// Acquire / release on the per-group progress counters of the work queue (gpu scope).
#ifdef __CUDACC__
__device__ __forceinline__ int dps_ld_acquire(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void dps_st_release(int* p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
#endif

// Parameter block of the wide stepper: same scalars, tables sized for MAX_GW generators (~11 KB, passed by value).
struct WideParams {
    long long P;
    int n_gen, solver, n_steps, step_begin, step_end;
    double dt, fn;
    const double* x0;
    const double* upar;
    const double* vpar;
    const double* zsub;
    double* rec;
    double* x_final;
    double* ckpt;
    int n_obs, z_rows, z_cols, n_cfg, z_per_sim, ckpt_stride, n_seg;
    int seg_step[MAX_SEG];
    int seg_cfg[MAX_SEG];
    unsigned char ctrl[MAX_GW];
    unsigned rec_mask[MAX_GW];
    short rec_chan[MAX_GW][DPS_N_FIELDS];
    short obs_chan[MAX_EXTRA_OBS][2];
    short reg_vary[MAX_GW][DPS_N_REG_SLOTS];   // row of vpar or -1, leading (register) slots only
    short slot_vary[MAX_GW][DPS_N_SLOTS];      // PMODE 2: row of vpar or -1 for every slot (read once, at block start)
    double cpar[DPS_N_SLOTS];
};

int fwd_template_g(int n_gen);
void fwd_geometry(int g_template, int* threads, int* sims_per_block);
#ifdef __CUDACC__
// regs_query / sims_per_block_query non-null: no launch, report the kernel that WOULD run for p (incl. p.P)
cudaError_t launch_forward(const FwdParams& p, int g_template, int cmask, int pmode, cudaStream_t stream,
                           int* regs_query, int* sims_per_block_query = nullptr);
cudaError_t launch_forward_wide(const WideParams& p, int cmask, int pmode, cudaStream_t stream, int* regs_query,
                                int* threads, int* sims_per_block);
#endif

constexpr int RED_MAX_CFG = 64;      // distinct faulted-bus sets per run
constexpr int RED_MAX_FAULT = 16;    // simultaneously faulted buses

struct ReduceParams {
    long long Bz;
    int n_bus, n_obs, n_gen, n_cfg, m_max;
    double y_fault;
    const double *z_re, *z_im, *yd_re, *yd_im;
    double* zsub;
    int obs_bus[MAX_GW + MAX_EXTRA_OBS];
    int gen_bus[MAX_GW];
    int cfg_n[RED_MAX_CFG];
    int cfg_bus[RED_MAX_CFG][RED_MAX_FAULT];
};
#ifdef __CUDACC__
cudaError_t launch_network_reduce_params(const ReduceParams& p, cudaStream_t stream);

cudaError_t run_fp64_probe(int iters, double* tflops, double* ms);
cudaError_t run_fp64_dmma_probe(int iters, double* tflops, double* ms);
cudaError_t launch_math_probe(long long n, const double* a, const double* b, double* out, cudaStream_t stream);
#endif

void set_error(const char* fmt, ...);

}  // namespace dps
