// dps_backward_params.h -- kernel parameter block of the adjoint (shared by dps_backward.cu and dps_backward.cuh).
#pragma once
#include "dps_internal.h"

namespace dps {

// MAXG = MAX_G for the warp-exchange adjoint (<= 16 generators), MAX_GW for the wide one (17..128 generators: block-wide
// exchange, tables of ~28 KB -- inside the 32 KB kernel-parameter limit).
template <int MAXG>
struct BwdParamsT {
    long long P;
    int n_gen, solver, n_steps;
    double dt, fn;
    const double* upar;
    const double* vpar;
    const double* zsub;
    const double* ckpt;        // [n_ckpt][n_gen][13][P]
    const double* grad_rec;    // [n_rec][n_steps][P]
    const double* grad_xf;     // [n_gen][13][P] or null
    double* g_x0;              // [n_gen][13][P]
    double* gslot;             // [n_want][P]
    double* g_zsub;            // [n_cfg][z_rows][z_cols][2][P] accumulated, or null
    double* scratch;           // [ckpt_stride][P][n_gen][stages * 18]  (dps_backward.cuh)
    int n_obs, z_rows, z_cols, n_cfg, z_per_sim, ckpt_stride, n_ckpt, n_seg;
    int pmode;                 // plan pmode: 1/3 = only the DPS_N_REG_SLOTS leading slots may differ per trajectory, 2 = any
    int seg_step[MAX_SEG];
    int seg_cfg[MAX_SEG];
    int ctrl[MAXG];
    unsigned rec_mask[MAXG];
    int wide_z_smem;           // wide build: the active configuration's generator block of Z is kept in shared memory
    int stage_offset;          // byte offset of the sweep's staging area in dynamic shared memory (set by the launcher)
    int rec_uniform;           // rec_mask identical for every generator (field tests become uniform branches)
    short rec_chan[MAXG][DPS_N_FIELDS];
    short obs_chan[MAX_EXTRA_OBS][2];
    short slot_vary[MAXG * DPS_N_SLOTS];
    short want_row[MAXG * DPS_N_SLOTS];
    // Work queue (q_items > 0; same scheme as the forward stepper, dps_internal.h): items = (block group, chunk of
    // q_chunk_ckpts checkpoint segments), chunk 0 = the LAST segments of the run (the sweep walks time backwards).  The
    // co-state and the slot-gradient accumulators of a group travel between its chunks through g_x0 / gslot.
    int q_items, q_groups, q_chunk_ckpts, q_capacity;
    int* q_counter;
    int* q_progress;
};
using BwdParams = BwdParamsT<MAX_G>;
using BwdWideParams = BwdParamsT<MAX_GW>;

}  // namespace dps
